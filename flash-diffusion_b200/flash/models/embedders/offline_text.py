"""Offline stand-ins for `transformers` `from_pretrained` calls of the text conditioners.

The reference's CLIP / T5 embedders (src/flash/models/embedders/clip/clip_embedder_model.py:10-201,
t5/t5_embedder_model.py:11-104) download tokenizers and encoder weights from the HF hub.  There is no network here:
the ARCHITECTURE named by `config.version` is built from its published config with random weights (the encoders are
frozen library models on the far side of the hot-path boundary, SURVEY.md §8f-3 — they run on `transformers`' own
modules, exactly as in the reference), and the tokenizer falls back to a deterministic hashing tokenizer with the same
call contract (`input_ids`, `attention_mask`, padding to `model_max_length`).  If the files exist locally (HF cache or
a directory path) the real `from_pretrained` is used.
"""
import hashlib
import re

import torch

# text-encoder configs of the checkpoints the example scripts name (examples/train_flash_{sd,sdxl,pixart,sd3}.py)
_CLIP_L = dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12, projection_dim=768,
               vocab_size=49408, max_position_embeddings=77, hidden_act="quick_gelu")
_CLIP_BIGG = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=20,
                  projection_dim=1280, vocab_size=49408, max_position_embeddings=77, hidden_act="gelu")
_T5_XXL = dict(d_model=4096, d_kv=64, d_ff=10240, num_layers=24, num_heads=64, vocab_size=32128,
               feed_forward_proj="gated-gelu", model_max_length=512)
_T5_V11_BASE = dict(d_model=768, d_kv=64, d_ff=2048, num_layers=12, num_heads=12, vocab_size=32128,
                    feed_forward_proj="gated-gelu", model_max_length=512)     # the reference's tests/test_embedders/test_t5_embedder.py:22
OFFLINE_TEXT_CONFIGS = {
    ("google/t5-v1_1-base", ""): _T5_V11_BASE,
    ("openai/clip-vit-large-patch14", ""): _CLIP_L,
    ("runwayml/stable-diffusion-v1-5", "text_encoder"): _CLIP_L,
    ("stabilityai/stable-diffusion-xl-base-1.0", "text_encoder"): _CLIP_L,
    ("stabilityai/stable-diffusion-xl-base-1.0", "text_encoder_2"): _CLIP_BIGG,
    ("stabilityai/stable-diffusion-3-medium", "text_encoder"): _CLIP_L,
    ("stabilityai/stable-diffusion-3-medium", "text_encoder_2"): _CLIP_BIGG,
    ("google/flan-t5-xxl", ""): _T5_XXL,
    ("PixArt-alpha/PixArt-XL-2-1024-MS", "text_encoder"): _T5_XXL,
    ("stabilityai/stable-diffusion-3-medium", "text_encoder_3"): _T5_XXL,
}


class HashTokenizer:
    """Whitespace / punctuation split, each token hashed into the vocabulary; BOS / EOS / PAD as CLIP (49406 / 49407 /
    49407) or T5 (no BOS, EOS 1, PAD 0).  Same `__call__` contract as the HF tokenizers the embedders use."""

    def __init__(self, vocab_size, model_max_length, style="clip", pad_token=None):
        self.vocab_size, self.model_max_length, self.style = vocab_size, model_max_length, style
        if style == "clip":
            self.bos, self.eos, self.pad = vocab_size - 2, vocab_size - 1, vocab_size - 1
            if pad_token == "!":
                self.pad = 0
        else:
            self.bos, self.eos, self.pad = None, 1, 0

    def _ids(self, text):
        out = []
        for w in re.findall(r"[\w']+|[^\w\s]", text.lower()):
            h = int.from_bytes(hashlib.blake2s(w.encode(), digest_size=4).digest(), "little")
            out.append(2 + h % (self.vocab_size - 4))
        return out

    def __call__(self, text, truncation=True, max_length=None, padding="max_length", return_tensors="pt",
                 add_special_tokens=True, **unused):
        if isinstance(text, str):
            text = [text]
        L = max_length or self.model_max_length
        ids, mask = [], []
        for t in text:
            x = self._ids(t)
            if add_special_tokens:
                x = ([self.bos] if self.bos is not None else []) + x[: L - (2 if self.bos is not None else 1)] + [self.eos]
            x = x[:L]
            m = [1] * len(x) + [0] * (L - len(x))
            ids.append(x + [self.pad] * (L - len(x)))
            mask.append(m)
        return {"input_ids": torch.tensor(ids, dtype=torch.long), "attention_mask": torch.tensor(mask, dtype=torch.long),
                "length": torch.tensor([sum(m) for m in mask])}


def _lookup(version, subfolder):
    key = (version, subfolder or "")
    if key not in OFFLINE_TEXT_CONFIGS:
        raise ValueError(f"text encoder {key} is neither available locally nor a known architecture "
                         f"(offline build; known: {sorted(OFFLINE_TEXT_CONFIGS)})")
    return dict(OFFLINE_TEXT_CONFIGS[key])


def load_tokenizer(cls, version, subfolder, revision, style, **kw):
    try:
        tok = cls.from_pretrained(version, subfolder=subfolder, revision=revision, local_files_only=True, **kw)
        # transformers >= 5 hands back an EMPTY tokenizer instead of raising when no files are found
        if len(tok) < 100 or tok.model_max_length > 1_000_000:
            raise OSError("empty tokenizer")
        return tok
    except Exception:
        cfg = _lookup(version, subfolder.replace("tokenizer", "text_encoder") if subfolder else subfolder)
        max_len = cfg.get("max_position_embeddings", cfg.get("model_max_length", 512))
        return HashTokenizer(cfg["vocab_size"], max_len, style=style, pad_token=kw.get("pad_token"))


def load_text_model(cls, version, subfolder, revision, **overrides):
    """`cls.from_pretrained` when the files are local, else the architecture with random weights."""
    try:
        return cls.from_pretrained(version, subfolder=subfolder, revision=revision, local_files_only=True)
    except Exception:
        cfg = _lookup(version, subfolder)
        cfg.pop("model_max_length", None)
        cfg.update(overrides)
        from transformers import CLIPTextConfig, T5Config
        if cls.__name__.startswith("CLIP"):
            return cls(CLIPTextConfig(**cfg))
        return cls(T5Config(**cfg))
