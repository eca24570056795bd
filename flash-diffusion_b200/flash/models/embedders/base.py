"""reference: src/flash/models/embedders/base/base_conditioner.py:6-58, base_conditioner_config.py"""
from typing import Any, Dict

from pydantic.dataclasses import dataclass

from ...config import BaseConfig
from ..base.base_model import BaseModel

# tensor rank of an embedder output -> conditioning slot (base_conditioner.py:6-10)
DIM2CONDITIONING = {2: "vector", 3: "crossattn", 4: "concat"}


@dataclass
class BaseConditionerConfig(BaseConfig):
    input_key: str = "text"
    unconditional_conditioning_rate: float = 0.0

    def __post_init__(self):
        super().__post_init__()
        assert 0.0 <= self.unconditional_conditioning_rate <= 1.0, \
            "Unconditional conditioning rate should be between 0 and 1"


class BaseConditioner(BaseModel):
    def __init__(self, config: BaseConditionerConfig):
        super().__init__(config)
        self.dim2outputkey = DIM2CONDITIONING
        self.ucg_rate = config.unconditional_conditioning_rate

    def forward(self, batch: Dict[str, Any], force_zero_embedding: bool = False, *args, **kwargs):
        raise NotImplementedError("Forward pass must be implemented in child class")
