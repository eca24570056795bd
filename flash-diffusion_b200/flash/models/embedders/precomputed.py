"""Pre-computed text embeddings + padding mask as a conditioner.

The reference's `T5TextEmbedder` (src/flash/models/embedders/t5/t5_embedder_model.py:48-104) returns
`{"crossattn": hidden_states, "attention_mask": mask}`; the T5-XXL encoder itself is out of scope (frozen, needs HF
weights), so this embedder feeds the same two conditioning slots from tensors already in the batch.
"""
from pydantic.dataclasses import dataclass

from .base import BaseConditioner, BaseConditionerConfig


@dataclass
class PrecomputedTextEmbedderConfig(BaseConditionerConfig):
    input_key: str = "text_emb"
    mask_key: str = "text_mask"


class PrecomputedTextEmbedder(BaseConditioner):
    def forward(self, batch, force_zero_embedding: bool = False, *args, **kwargs):
        x = batch[self.input_key]
        if force_zero_embedding:
            x = 0 * x
        out = {"crossattn": x}
        mask = batch.get(self.config.mask_key)
        if mask is not None:
            out["attention_mask"] = mask
        return out
