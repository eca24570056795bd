"""reference: src/flash/data/mappers/{base,mappers,mappers_config,mappers_wrapper}.py (same names and semantics).
The Canny / MiDaS mappers need `controlnet_aux` (T2I-adapter recipe, out of the hot path): they import lazily."""
import json
from typing import Any, Callable, Dict, List, Literal, Optional, Union

from pydantic.dataclasses import dataclass

from ...config import BaseConfig


@dataclass
class BaseMapperConfig(BaseConfig):
    verbose: bool = False
    key: Optional[str] = None
    output_key: Optional[str] = None


@dataclass
class KeyRenameMapperConfig(BaseMapperConfig):
    key_map: Dict[str, str] = None
    condition_key: Optional[str] = None
    condition_fn: Optional[Callable[[Any], bool]] = None
    else_key_map: Optional[Dict[str, str]] = None


@dataclass
class TorchvisionMapperConfig(BaseMapperConfig):
    key: str = "image"
    transforms: List[str] = None
    transforms_kwargs: List[Dict[str, Any]] = None


@dataclass
class RescaleMapperConfig(BaseMapperConfig):
    key: str = "image"


@dataclass
class KeysFromJSONMapperConfig(BaseMapperConfig):
    key: str = "json"
    keys_to_extract: Union[str, List[str]] = None
    remove_original: bool = True
    strict: bool = True


@dataclass
class SelectKeysMapperConfig(BaseMapperConfig):
    keys: Union[str, List[str]] = None


@dataclass
class RemoveKeysMapperConfig(BaseMapperConfig):
    keys: Union[str, List[str]] = None


@dataclass
class SetValueConfig(BaseMapperConfig):
    value: Any = None


@dataclass
class CannyEdgeMapperConfig(BaseMapperConfig):
    key: str = "image"
    output_key: str = "edges"
    detect_resolution: int = 384
    image_resolution: int = 1024
    mode: Literal["L", "RGB"] = "RGB"


@dataclass
class MidasDepthMapperConfig(BaseMapperConfig):
    key: str = "image"
    output_key: str = "depth"
    detect_resolution: int = 512
    image_resolution: int = 1024
    mode: Literal["L", "RGB"] = "RGB"


def _as_list(x):
    return [x] if isinstance(x, str) else list(x or [])


class BaseMapper:
    def __init__(self, config: BaseMapperConfig):
        self.config = config
        self.key = config.key
        self.output_key = config.output_key if config.output_key is not None else config.key
        self.verbose = config.verbose

    def map(self):
        raise NotImplementedError("The __call__ method must be implemented")


class KeyRenameMapper(BaseMapper):
    """renames keys by `key_map` (by `else_key_map` when `condition_fn(batch[condition_key])` is false)"""

    def __init__(self, config: KeyRenameMapperConfig):
        super().__init__(config)
        self.key_map, self.condition_key = config.key_map, config.condition_key
        self.condition_fn, self.else_key_map = config.condition_fn, config.else_key_map

    def __call__(self, batch: dict):
        key_map = self.key_map
        if self.condition_key is not None and not self.condition_fn(batch[self.condition_key]):
            key_map = self.else_key_map or {}
        for old, new in key_map.items():
            if old in batch:
                batch[new] = batch.pop(old)
        return batch


class TorchvisionMapper(BaseMapper):
    """`torchvision.transforms.<name>(**kwargs)` composed in order on batch[key]"""

    def __init__(self, config: TorchvisionMapperConfig):
        super().__init__(config)
        from torchvision import transforms as T
        kwargs = config.transforms_kwargs or [{}] * len(config.transforms)
        self.transforms = T.Compose([getattr(T, n)(**kw) for n, kw in zip(config.transforms, kwargs)])

    def __call__(self, batch: Dict[str, Any]) -> Dict[str, Any]:
        if self.key in batch:
            batch[self.output_key] = self.transforms(batch[self.key])
        return batch


class RescaleMapper(BaseMapper):
    """[0, 1] -> [-1, 1]"""

    def __call__(self, batch: Dict[str, Any]) -> Dict[str, Any]:
        if isinstance(batch[self.key], list):
            batch[self.output_key] = [2 * v - 1.0 for v in batch[self.key]]
        else:
            batch[self.output_key] = 2 * batch[self.key] - 1.0
        return batch


class KeysFromJSONMapper(BaseMapper):
    def __init__(self, config: KeysFromJSONMapperConfig):
        super().__init__(config)
        self.keys_to_extract = _as_list(config.keys_to_extract)
        self.remove_original, self.strict = config.remove_original, config.strict

    def __call__(self, batch: Dict[str, Any]) -> Dict[str, Any]:
        obj = batch[self.key]
        if isinstance(obj, (bytes, str)):
            obj = json.loads(obj)
        for k in self.keys_to_extract:
            if k in obj:
                batch[k] = obj[k]
            elif self.strict:
                raise KeyError(f"{k} not in json at key {self.key}")
        if self.remove_original:
            del batch[self.key]
        return batch


class SelectKeysMapper(BaseMapper):
    def __init__(self, config: SelectKeysMapperConfig):
        super().__init__(config)
        self.keys = _as_list(config.keys)

    def __call__(self, batch: Dict[str, Any]) -> Dict[str, Any]:
        return {k: v for k, v in batch.items() if k in self.keys}


class RemoveKeysMapper(BaseMapper):
    def __init__(self, config: RemoveKeysMapperConfig):
        super().__init__(config)
        self.keys = _as_list(config.keys)

    def __call__(self, batch: Dict[str, Any]) -> Dict[str, Any]:
        for k in self.keys:
            batch.pop(k, None)
        return batch


class SetValueMapper(BaseMapper):
    def __init__(self, config: SetValueConfig):
        super().__init__(config)
        self.value = config.value

    def __call__(self, batch: dict):
        batch[self.output_key] = self.value
        return batch


class _DetectorMapper(BaseMapper):
    _detector = None

    def __init__(self, config):
        super().__init__(config)
        try:
            import controlnet_aux  # noqa: F401
        except ImportError as e:
            raise ImportError(f"{type(self).__name__} needs `controlnet_aux` (T2I-adapter recipe; not part of the "
                              "B200 hot path and not installable offline)") from e


class CannyEdgeMapper(_DetectorMapper):
    pass


class MidasDepthMapper(_DetectorMapper):
    pass


class MapperWrapper:
    """applies the mappers in order"""

    def __init__(self, mappers: Union[List[BaseMapper], None] = None):
        self.mappers = mappers or []

    def __call__(self, batch: Dict[str, Any]) -> Dict[str, Any]:
        for m in self.mappers:
            batch = m(batch)
        return batch


__all__ = ["BaseMapper", "BaseMapperConfig", "KeyRenameMapper", "KeyRenameMapperConfig", "KeysFromJSONMapper",
           "KeysFromJSONMapperConfig", "MapperWrapper", "RemoveKeysMapper", "RemoveKeysMapperConfig", "RescaleMapper",
           "RescaleMapperConfig", "SelectKeysMapper", "SelectKeysMapperConfig", "SetValueMapper", "SetValueConfig",
           "TorchvisionMapper", "TorchvisionMapperConfig", "CannyEdgeMapper", "CannyEdgeMapperConfig",
           "MidasDepthMapper", "MidasDepthMapperConfig"]
