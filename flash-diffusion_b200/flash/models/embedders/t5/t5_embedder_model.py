"""T5 text conditioner (reference src/flash/models/embedders/t5/t5_embedder_model.py:11-104 and t5_embedder_config.py):
rank-3 "crossattn" output of the chosen encoder layer, optionally with the tokenizer's "attention_mask" (PixArt's
masked context, consumed by DiffusersTransformer2DWrapper as cond["attention_mask"]).  Encoder = `transformers`'
T5EncoderModel as in the reference; offline: architecture from the published config, random weights."""
from typing import Any, Dict, List, Literal, Optional

from pydantic.dataclasses import dataclass

from ..base import BaseConditioner, BaseConditionerConfig
from ..offline_text import load_text_model, load_tokenizer


@dataclass
class T5TextEmbedderConfig(BaseConditionerConfig):
    version: str = "google/flan-t5-xxl"
    text_embedder_subfolder: str = ""
    tokenizer_subfolder: str = ""
    text_embedder_revision: str = "main"
    tokenizer_revision: str = "main"
    layer: Literal["last", "hidden"] = "last"
    layer_idx: int = None
    input_key: str = "text"
    tokenizer_max_length: Optional[int] = None
    projection_nn_modules: Optional[List[str]] = None
    projection_nn_modules_kwargs: Optional[List[Dict[str, Any]]] = None
    returns_attention_mask: bool = False
    tokenizer_truncation: bool = True
    tokenizer_return_length: bool = True
    tokenizer_add_special_tokens: bool = True

    def __post_init__(self):
        """reference t5_embedder_config.py:49-66: a hidden layer needs an index within the 24 encoder blocks; the
        projection lists default to empty and must pair up (a pydantic ValidationError otherwise)."""
        super().__post_init__()
        if self.layer == "hidden":
            assert self.layer_idx is not None, "Layer index is required for hidden layer"
            assert 0 <= abs(self.layer_idx) <= 24, "Layer index should be between 0 and 24"
        self.projection_nn_modules = self.projection_nn_modules or []
        self.projection_nn_modules_kwargs = self.projection_nn_modules_kwargs or []
        assert len(self.projection_nn_modules) == len(self.projection_nn_modules_kwargs), \
            "Number of modules and kwargs should be same"


class T5TextEmbedder(BaseConditioner):
    def __init__(self, config: T5TextEmbedderConfig):
        BaseConditioner.__init__(self, config)
        from transformers import T5EncoderModel, T5Tokenizer
        self.tokenizer = load_tokenizer(T5Tokenizer, config.version, config.tokenizer_subfolder,
                                        config.tokenizer_revision, "t5")
        self.transformer = load_text_model(T5EncoderModel, config.version, config.text_embedder_subfolder,
                                           config.text_embedder_revision)
        self.max_length = config.tokenizer_max_length or self.tokenizer.model_max_length
        self.layer = config.layer
        self.layer_idx = config.layer_idx
        self.returns_attention_mask = config.returns_attention_mask
        self.tokenizer_truncation = config.tokenizer_truncation
        self.tokenizer_return_length = config.tokenizer_return_length
        self.tokenizer_add_special_tokens = config.tokenizer_add_special_tokens

    def freeze(self):
        super().freeze()
        self.transformer = self.transformer.eval()

    def forward(self, batch: Dict[str, Any], force_zero_embedding: bool = False, device: str = "cpu", *args,
                **kwargs) -> Dict[str, Any]:
        enc = self.tokenizer(batch[self.input_key], truncation=self.tokenizer_truncation, max_length=self.max_length,
                             return_length=self.tokenizer_return_length, return_overflowing_tokens=False,
                             padding="max_length", return_tensors="pt",
                             add_special_tokens=self.tokenizer_add_special_tokens)
        tokens, mask = enc["input_ids"].to(device), enc["attention_mask"].to(device)
        self.transformer = self.transformer.to(device)
        outputs = self.transformer(input_ids=tokens, attention_mask=mask, output_hidden_states=self.layer == "hidden")
        z = outputs.last_hidden_state if self.layer == "last" else outputs.hidden_states[self.layer_idx]
        if force_zero_embedding:
            z, mask = 0 * z, 0 * mask
        out = {self.dim2outputkey[z.dim()]: z}
        if self.returns_attention_mask:
            out["attention_mask"] = mask
        return out
