"""reference: src/flash/models/embedders/torch_nn/embedders.py:10-56, embedders_config.py"""
import importlib
from typing import Any, Dict, List

import torch.nn as nn
from pydantic.dataclasses import dataclass

from .base import BaseConditioner, BaseConditionerConfig


@dataclass
class TorchNNEmbedderConfig(BaseConditionerConfig):
    nn_modules: List[str] = None
    nn_modules_kwargs: List[Dict[str, Any]] = None
    flatten_output: bool = False
    input_key: str = "image"

    def __post_init__(self):
        super().__post_init__()
        self.nn_modules = self.nn_modules or []
        self.nn_modules_kwargs = self.nn_modules_kwargs or []
        assert len(self.nn_modules) == len(self.nn_modules_kwargs), "Number of modules and kwargs should be same"


class TorchNNEmbedder(BaseConditioner):
    """Chains torch.nn modules given by dotted name; the output rank selects the conditioning slot."""

    def __init__(self, config: TorchNNEmbedderConfig):
        super().__init__(config)
        self.flatten_output = config.flatten_output
        mods = []
        for path, kw in zip(config.nn_modules, config.nn_modules_kwargs):
            mod_name, cls_name = path.rsplit(".", 1)
            mods.append(getattr(importlib.import_module(mod_name), cls_name)(**kw))
        self.nn_modules = nn.Sequential(*mods)

    def forward(self, batch: Dict[str, Any], force_zero_embedding: bool = False, *args, **kwargs):
        x = self.nn_modules(batch[self.input_key])
        if force_zero_embedding:
            x = 0 * x
        if self.flatten_output:
            x = x.view(x.size(0), -1)
        return {self.dim2outputkey[x.dim()]: x}
