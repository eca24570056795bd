"""Golden vectors from the REFERENCE's own text conditioners (src/flash/models/embedders/clip/clip_embedder_model.py:10-208
`ClipEmbedder`, `ClipEmbedderWithProjection`; t5/t5_embedder_model.py:11-100 `T5TextEmbedder`), imported unmodified from
/root/reference/src:   python tests/golden/make_reference_text_golden.py  ->  tests/golden/reference_text.pt
The encoders are transformers' REAL CLIPTextModel / CLIPTextModelWithProjection / T5EncoderModel classes (installed
here); only `from_pretrained` is redirected to tiny seeded configurations and a deterministic stand-in tokenizer, because
no checkpoint can be downloaded.  What the fixture pins is the reference's glue: which hidden state is returned (last /
hidden[layer_idx] / pooled), `always_return_pooled`, the projection variant's `text_embeds`, the T5 attention mask and
`force_zero_embedding`."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)

TEXTS = ["a photo of a cat", "", "an astronaut riding a horse on the moon, highly detailed, 8k"]


class TinyTokenizer:
    """deterministic stand-in for CLIPTokenizer / T5Tokenizer: word hashes, bos / eos, right padding"""
    model_max_length = 12

    def __init__(self, vocab=97, pad_id=0, bos=1, eos=2):
        self.vocab, self.pad_id, self.bos, self.eos = vocab, pad_id, bos, eos

    def __call__(self, text, truncation=True, max_length=None, return_length=False, return_overflowing_tokens=False,
                 padding="max_length", return_tensors="pt", add_special_tokens=True):
        L = max_length or self.model_max_length
        ids, mask = [], []
        for t in text:
            w = [3 + (sum(ord(c) * (i + 1) for i, c in enumerate(tok)) % (self.vocab - 3)) for tok in t.split()]
            seq = ([self.bos] if add_special_tokens else []) + w + ([self.eos] if add_special_tokens else [])
            seq = seq[:L] if truncation else seq
            m = [1] * len(seq) + [0] * (L - len(seq))
            ids.append(seq + [self.pad_id] * (L - len(seq)))
            mask.append(m)
        return {"input_ids": torch.tensor(ids), "attention_mask": torch.tensor(mask)}


def tiny_clip(projection, seed):
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection
    cfg = CLIPTextConfig(vocab_size=97, hidden_size=32, intermediate_size=64, num_hidden_layers=3, num_attention_heads=4,
                         max_position_embeddings=12, projection_dim=24, bos_token_id=1, eos_token_id=2, pad_token_id=0)
    torch.manual_seed(seed)
    return (CLIPTextModelWithProjection if projection else CLIPTextModel)(cfg).eval()


def tiny_t5(seed):
    from transformers import T5Config, T5EncoderModel
    cfg = T5Config(vocab_size=97, d_model=32, d_kv=8, d_ff=64, num_layers=2, num_heads=4, pad_token_id=0, eos_token_id=2)
    torch.manual_seed(seed)
    return T5EncoderModel(cfg).eval()


CLIP_CASES = [dict(layer="last"), dict(layer="hidden", layer_idx=-2), dict(layer="pooled"),
              dict(layer="hidden", layer_idx=1, always_return_pooled=True)]
T5_CASES = [dict(layer="last"), dict(layer="hidden", layer_idx=-2, returns_attention_mask=True),
            dict(layer="last", tokenizer_max_length=8, returns_attention_mask=True)]


def run(clip_cls, clip_proj_cls, t5_cls, clip_cfg_cls, t5_cfg_cls, build):
    """build(cls, cfg, transformer, tokenizer) -> embedder instance with the stand-ins in place"""
    out = {"clip": [], "clip_proj": [], "t5": []}
    batch = {"text": TEXTS}
    with torch.no_grad():
        for name, cls, proj in (("clip", clip_cls, False), ("clip_proj", clip_proj_cls, True)):
            for kw in CLIP_CASES:
                emb = build(cls, clip_cfg_cls(version="stand-in", input_key="text", **kw), tiny_clip(proj, 5), TinyTokenizer())
                for fz in (False, True):
                    o = emb(dict(batch), force_zero_embedding=fz)
                    out[name].append({k: v.clone() for k, v in o.items()})
        for kw in T5_CASES:
            emb = build(t5_cls, t5_cfg_cls(version="stand-in", input_key="text", **kw), tiny_t5(6), TinyTokenizer())
            for fz in (False, True):
                o = emb(dict(batch), force_zero_embedding=fz)
                out["t5"].append({k: v.clone() for k, v in o.items()})
    return out


def main():
    import transformers
    import make_reference_step_golden as G
    G.install_shims()
    sys.path.insert(0, G.REF_SRC)
    state = {}
    # the reference constructors call <Class>.from_pretrained(version, ...): hand them the stand-ins
    for cls_name in ("CLIPTokenizer", "T5Tokenizer", "CLIPTextModel", "CLIPTextModelWithProjection", "T5EncoderModel"):
        getattr(transformers, cls_name).from_pretrained = classmethod(
            lambda cls, *a, _n=cls_name, **k: state["tokenizer"] if "Tokenizer" in _n else state["transformer"])
    from flash.models.embedders import (ClipEmbedder, ClipEmbedderConfig, ClipEmbedderWithProjection, T5TextEmbedder,
                                        T5TextEmbedderConfig)
    import flash
    assert os.path.realpath(flash.__path__[0]).startswith(G.REF_SRC)

    def build(cls, cfg, transformer, tokenizer):
        state["transformer"], state["tokenizer"] = transformer, tokenizer
        return cls(cfg)
    out = run(ClipEmbedder, ClipEmbedderWithProjection, T5TextEmbedder, ClipEmbedderConfig, T5TextEmbedderConfig, build)
    out["generated_by"] = os.path.relpath(__file__, ROOT)
    path = os.path.join(HERE, "reference_text.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", {k: len(v) for k, v in out.items() if isinstance(v, list)})


if __name__ == "__main__":
    main()
