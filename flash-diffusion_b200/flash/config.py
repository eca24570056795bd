"""Config base class (mirrors reference src/flash/config.py:13-141: same method names and semantics).

pydantic dataclasses silently drop unknown keyword arguments; the reference's examples and tests rely
on that (examples/train_flash_sdxl.py:163, tests/test_flash/test_flash_diffusion.py:78-84), so it is
part of the contract here too.
"""
import json
import os
import warnings
from dataclasses import asdict, field
from typing import Any, Dict, Union

import yaml
from pydantic.dataclasses import dataclass


@dataclass
class BaseConfig:
    name: str = field(init=False)

    def __post_init__(self):
        self.name = type(self).__name__

    # ---- construction
    @classmethod
    def from_dict(cls, config_dict: Dict[str, Any]) -> "BaseConfig":
        return cls(**config_dict)

    @classmethod
    def _dict_from_json(cls, json_path: Union[str, os.PathLike]) -> Dict[str, Any]:
        if not os.path.exists(json_path):
            raise FileNotFoundError(f"Config file not found. Please check path '{json_path}'")
        with open(json_path) as f:
            try:
                return json.load(f)
            except (TypeError, json.JSONDecodeError) as e:
                raise TypeError(f"File {json_path} not loadable. Maybe not json ? \n"
                                f"Catch Exception {type(e)} with message: {e}") from e

    @classmethod
    def _check_name(cls, d: Dict[str, Any]) -> Dict[str, Any]:
        """the serialised `name` is required and only WARNED about when it is another class's (reference :70-77, :131-138)"""
        name = d.pop("name")
        if name != cls.__name__:
            warnings.warn(f"You are trying to load a `{cls.__name__}` while a `{name}` is given.")
        return d

    @classmethod
    def from_json(cls, json_path: str) -> "BaseConfig":
        return cls.from_dict(cls._check_name(cls._dict_from_json(json_path)))

    @classmethod
    def from_yaml(cls, yaml_path: str) -> "BaseConfig":
        with open(yaml_path, "r") as f:
            try:
                d = yaml.safe_load(f)
            except yaml.YAMLError as e:
                raise yaml.YAMLError(f"File {yaml_path} not loadable. Maybe not yaml ? \n"
                                     f"Catch Exception {type(e)} with message: {e}") from e
        return cls.from_dict(cls._check_name(d))

    # ---- serialisation
    def to_dict(self) -> dict:
        return asdict(self)

    def to_json_string(self):
        return json.dumps(self.to_dict())

    def save_json(self, file_path: str):
        with open(file_path, "w", encoding="utf-8") as fp:
            fp.write(self.to_json_string())

    def save_yaml(self, file_path: str):
        with open(file_path, "w", encoding="utf-8") as fp:
            yaml.dump(self.to_dict(), fp)
