"""reference: src/flash/models/embedders/clip/clip_embedder_config.py:8-58 (same fields, defaults and checks)."""
from typing import Literal, Optional

from pydantic.dataclasses import dataclass

from ..base import BaseConditionerConfig


@dataclass
class ClipEmbedderConfig(BaseConditionerConfig):
    version: str = "openai/clip-vit-large-patch14"
    text_embedder_subfolder: str = ""
    tokenizer_subfolder: str = ""
    text_embedder_revision: str = "main"
    tokenizer_revision: str = "main"
    layer: Literal["last", "pooled", "hidden"] = "last"
    layer_idx: int = None
    always_return_pooled: bool = False
    input_key: str = "text"
    pad_token: Optional[str] = None
    tokenizer_truncation: bool = True
    tokenizer_return_length: bool = True

    def __post_init__(self):
        super().__post_init__()
        if self.layer == "hidden":
            assert self.layer_idx is not None, "Layer index is required for hidden layer"
            assert 0 <= abs(self.layer_idx) <= 12, "Layer index should be between 0 and 12"
