"""Config base class (mirrors reference src/flash/config.py:13-141: same method names and semantics).

pydantic dataclasses silently drop unknown keyword arguments; the reference's examples and tests rely
on that (examples/train_flash_sdxl.py:163, tests/test_flash/test_flash_diffusion.py:78-84), so it is
part of the contract here too.
"""
import json
import os
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
    def from_json(cls, json_path: str) -> "BaseConfig":
        d = cls._dict_from_json(json_path)
        name = d.pop("name", None)
        if name is not None and name != cls.__name__:
            raise ValueError(f"You are trying to load a `{cls.__name__}` while a `{name}` is given.")
        return cls.from_dict(d)

    @classmethod
    def from_yaml(cls, yaml_path: str) -> "BaseConfig":
        with open(yaml_path, "r") as f:
            d = yaml.safe_load(f)
        d = dict(d or {})
        name = d.pop("name", None)
        if name is not None and name != cls.__name__:
            raise ValueError(f"You are trying to load a `{cls.__name__}` while a `{name}` is given.")
        return cls.from_dict(d)

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
