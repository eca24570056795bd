"""CLIP text conditioners (reference src/flash/models/embedders/clip/clip_embedder_model.py:10-201): the same classes,
constructor, `freeze` and `forward(batch, force_zero_embedding, device)` contract — a rank-3 "crossattn" output from
the chosen layer plus, with `always_return_pooled`, the rank-2 "vector" pooled output (`pooler_output` for
`ClipEmbedder`, `text_embeds` for `ClipEmbedderWithProjection`).  The encoder is `transformers`' own frozen CLIP text
model, as in the reference (the far side of the hot-path boundary); offline it is built from the
architecture's published config with random weights and a hashing tokenizer (offline_text.py)."""
from typing import Any, Dict

from ..base import BaseConditioner
from ..offline_text import load_text_model, load_tokenizer
from .clip_embedder_config import ClipEmbedderConfig


class _ClipBase(BaseConditioner):
    _with_projection = False

    def __init__(self, config: ClipEmbedderConfig):
        BaseConditioner.__init__(self, config)
        from transformers import CLIPTextModel, CLIPTextModelWithProjection, CLIPTokenizer
        kwargs, over = {}, {}
        if config.pad_token is not None:
            kwargs["pad_token"] = config.pad_token
        if self._with_projection and config.version in ("laion/CLIP-ViT-L-14-laion2B-s32B-b82K",
                                                        "laion/CLIP-ViT-L-14-DataComp.XL-s13B-b90K"):
            over["projection_dim"] = 768          # reference :121-126
        self.tokenizer = load_tokenizer(CLIPTokenizer, config.version, config.tokenizer_subfolder,
                                        config.tokenizer_revision, "clip", **kwargs)
        self.transformer = load_text_model(CLIPTextModelWithProjection if self._with_projection else CLIPTextModel,
                                           config.version, config.text_embedder_subfolder,
                                           config.text_embedder_revision, **over)
        self.max_length = self.tokenizer.model_max_length
        self.layer = config.layer
        self.layer_idx = config.layer_idx
        self.always_return_pooled = config.always_return_pooled
        self.tokenizer_truncation = config.tokenizer_truncation
        self.tokenizer_return_length = config.tokenizer_return_length

    def freeze(self):
        super().freeze()
        self.transformer = self.transformer.eval()

    def forward(self, batch: Dict[str, Any], force_zero_embedding: bool = False, device="cpu", *args, **kwargs):
        enc = self.tokenizer(batch[self.input_key], truncation=self.tokenizer_truncation, max_length=self.max_length,
                             return_length=self.tokenizer_return_length, return_overflowing_tokens=False,
                             padding="max_length", return_tensors="pt")
        tokens = enc["input_ids"].to(device)
        self.transformer = self.transformer.to(device)
        outputs = self.transformer(input_ids=tokens, output_hidden_states=self.layer == "hidden")
        pooled = outputs.text_embeds if self._with_projection else outputs.pooler_output
        if self.layer == "last":
            z = outputs.last_hidden_state
        elif self.layer == "pooled":
            z = pooled[:, None, :]
        else:
            z = outputs.hidden_states[self.layer_idx]
        if force_zero_embedding:
            z = 0 * z
        output = {self.dim2outputkey[z.dim()]: z}
        if self.always_return_pooled:
            output[self.dim2outputkey[2]] = 0 * pooled if force_zero_embedding else pooled
        return output


class ClipEmbedder(_ClipBase):
    _with_projection = False


class ClipEmbedderWithProjection(_ClipBase):
    _with_projection = True
