"""flash.models.embedders text conditioners against outputs of the REFERENCE's own classes (tests/golden/reference_text.pt,
written by tests/golden/make_reference_text_golden.py from the unmodified clip_embedder_model.py / t5_embedder_model.py):
the same tiny seeded transformers encoders and stand-in tokenizer are put behind the product's ClipEmbedder,
ClipEmbedderWithProjection and T5TextEmbedder."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = torch.load(os.path.join(HERE, "golden", "reference_text.pt"), weights_only=False)


def test_product_text_conditioners_match_reference_run(monkeypatch):
    import make_reference_text_golden as G
    from flash.models.embedders import (ClipEmbedder, ClipEmbedderConfig, ClipEmbedderWithProjection, T5TextEmbedder,
                                        T5TextEmbedderConfig)
    from flash.models.embedders.clip import clip_embedder_model as CM
    from flash.models.embedders.t5 import t5_embedder_model as TM
    state = {}
    for mod in (CM, TM):
        monkeypatch.setattr(mod, "load_tokenizer", lambda *a, **k: state["tokenizer"])
        monkeypatch.setattr(mod, "load_text_model", lambda *a, **k: state["transformer"])

    def build(cls, cfg, transformer, tokenizer):
        state["transformer"], state["tokenizer"] = transformer, tokenizer
        return cls(cfg)
    out = G.run(ClipEmbedder, ClipEmbedderWithProjection, T5TextEmbedder, ClipEmbedderConfig, T5TextEmbedderConfig, build)
    for fam in ("clip", "clip_proj", "t5"):
        assert len(out[fam]) == len(GOLD[fam])
        for i, (got, want) in enumerate(zip(out[fam], GOLD[fam])):
            assert list(got) == list(want), (fam, i, list(got), list(want))
            for k in want:
                assert got[k].shape == want[k].shape and got[k].dtype == want[k].dtype, (fam, i, k)
                assert torch.allclose(got[k].float(), want[k].float(), rtol=1e-5, atol=1e-6), (fam, i, k)
