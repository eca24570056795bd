"""reference: src/flash/models/embedders/conditioners_wrapper.py:9-91"""
from typing import Any, Dict, List, Union

import torch
import torch.nn as nn

from .base import BaseConditioner

KEY2CATDIM = {"vector": 1, "crossattn": 2, "concat": 1}


class ConditionerWrapper(nn.Module):
    """Applies a list of conditioners and concatenates their outputs per conditioning slot.

    UCG semantics (conditioners_wrapper.py:62-71): a conditioner whose input_key is in `ucg_keys` is forced
    to a zero embedding; otherwise it is zeroed with probability `ucg_rate` unless `set_ucg_rate_zero`.
    """

    def __init__(self, conditioners: Union[List[BaseConditioner], None] = None):
        super().__init__()
        self.conditioners = nn.ModuleList(conditioners)

    def conditioner_sanity_check(self):
        """reference :32-37 — every key of `self.ucg_keys` (an attribute the caller sets; the reference never does)
        must be the input key of one of the conditioners."""
        keys = {c.input_key for c in self.conditioners}
        assert all(k in keys for k in self.ucg_keys)

    def forward(self, batch: Dict[str, Any], ucg_keys: List[str] = None, set_ucg_rate_zero=False,
                *args, **kwargs):
        ucg_keys = ucg_keys or []
        cond: Dict[str, torch.Tensor] = {}
        for conditioner in self.conditioners:
            if conditioner.input_key in ucg_keys:
                zero = True
            elif conditioner.ucg_rate > 0 and not set_ucg_rate_zero:
                zero = bool(torch.rand(1) < conditioner.ucg_rate)
            else:
                zero = False
            out = conditioner.forward(batch, force_zero_embedding=zero, *args, **kwargs)
            for key, value in out.items():
                cond[key] = value if key not in cond else torch.cat([cond[key], value], KEY2CATDIM[key])
        return {"cond": cond}

    def to(self, device):
        self.conditioners = self.conditioners.to(device)
        return self
