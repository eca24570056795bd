"""reference: src/flash/data/filters/{base,filters,filters_config,filter_wrapper}.py (same names and semantics)."""
import logging
from typing import Any, Callable, Dict, List, Optional, Union

from pydantic.dataclasses import dataclass

from ...config import BaseConfig


@dataclass
class BaseFilterConfig(BaseConfig):
    verbose: bool = False


@dataclass
class KeyFilterConfig(BaseFilterConfig):
    keys: Union[str, List[str]] = "txt"


@dataclass
class FilterOnConditionConfig(BaseFilterConfig):
    condition_key: Optional[str] = None
    condition_fn: Optional[Callable[[Any], bool]] = None
    strict: bool = False


class BaseFilter:
    def __init__(self, config: BaseFilterConfig):
        self.verbose = config.verbose

    def __call__(self, sample: Dict[str, Any]) -> bool:
        raise NotImplementedError("The __call__ method must be implemented")


class KeyFilter(BaseFilter):
    """keeps a sample only if it holds every one of `keys`"""

    def __init__(self, config: KeyFilterConfig):
        super().__init__(config)
        self.keys = set([config.keys] if isinstance(config.keys, str) else config.keys)

    def __call__(self, batch: dict) -> bool:
        ok = self.keys.issubset(batch.keys())
        if not ok and self.verbose:
            logging.error(f"Missing keys: {self.keys - set(batch.keys())}")
        return ok


class FilterOnCondition(BaseFilter):
    """keeps a sample if `condition_fn(sample[condition_key])`; a missing key drops it only when `strict`"""

    def __init__(self, config: FilterOnConditionConfig):
        super().__init__(config)
        self.condition_key, self.condition_fn, self.strict = config.condition_key, config.condition_fn, config.strict

    def __call__(self, batch: dict) -> bool:
        if self.condition_key not in batch:
            return not self.strict
        return bool(self.condition_fn(batch[self.condition_key]))


class FilterWrapper:
    """all filters must accept"""

    def __init__(self, filters: Union[List[BaseFilter], None] = None):
        self.filters = filters or []

    def __call__(self, batch: Dict[str, Any]) -> bool:
        return all(f(batch) for f in self.filters)


__all__ = ["BaseFilter", "BaseFilterConfig", "KeyFilter", "KeyFilterConfig", "FilterOnCondition",
           "FilterOnConditionConfig", "FilterWrapper"]
