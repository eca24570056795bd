"""reference: src/flash/models/base/base_model.py:8-37 (same hooks)."""
from typing import Any, Dict

import torch.nn as nn

from .model_config import ModelConfig


class BaseModel(nn.Module):
    def __init__(self, config: ModelConfig):
        super().__init__()
        self.config = config
        self.input_key = config.input_key

    def forward(self, batch: Dict[str, Any], *args, **kwargs):
        raise NotImplementedError("forward method is not implemented")

    def freeze(self):
        self.eval()
        for p in self.parameters():
            p.requires_grad = False

    def compute_metrics(self, batch, *args, **kwargs):
        return {}

    def sample(self, batch, *args, **kwargs):
        return {}

    def log_samples(self, batch, *args, **kwargs):
        return None

    def on_train_batch_end(self, batch, *args, **kwargs):
        pass
