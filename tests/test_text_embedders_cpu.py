"""Text conditioners (reference src/flash/models/embedders/{clip,t5}) and the data / compat surface
(Appendix A) on the CPU: contracts only — the encoders are `transformers` library modules with random weights."""
import io
import json
import os
import sys
import tarfile

import pytest
import torch

from flash.models.embedders import (ClipEmbedder, ClipEmbedderConfig, ClipEmbedderWithProjection, ConditionerWrapper,
                                    T5TextEmbedder, T5TextEmbedderConfig)
from flash.models.embedders import offline_text as OT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def tiny(monkeypatch):
    clip = dict(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4, projection_dim=24,
                vocab_size=49408, max_position_embeddings=77, hidden_act="quick_gelu")
    t5 = dict(d_model=32, d_kv=8, d_ff=64, num_layers=2, num_heads=4, vocab_size=32128, feed_forward_proj="gated-gelu",
              model_max_length=512)
    monkeypatch.setitem(OT.OFFLINE_TEXT_CONFIGS, ("openai/clip-vit-large-patch14", ""), clip)
    monkeypatch.setitem(OT.OFFLINE_TEXT_CONFIGS, ("google/flan-t5-xxl", ""), t5)


def test_clip_embedders_contract(tiny):
    e = ClipEmbedder(ClipEmbedderConfig(layer="last"))
    e.freeze()
    out = e({"text": ["a cat", "a raccoon reading a book in a lush forest"]})
    assert set(out) == {"crossattn"} and out["crossattn"].shape == (2, 77, 32)
    assert not any(p.requires_grad for p in e.parameters())
    p = ClipEmbedderWithProjection(ClipEmbedderConfig(layer="hidden", layer_idx=-2, always_return_pooled=True))
    o = p({"text": ["x", "y"]})
    assert o["crossattn"].shape == (2, 77, 32) and o["vector"].shape == (2, 24)
    z = p({"text": ["x"]}, force_zero_embedding=True)
    assert float(z["crossattn"].abs().max()) == 0 and float(z["vector"].abs().max()) == 0
    pooled = ClipEmbedder(ClipEmbedderConfig(layer="pooled"))({"text": ["x"]})
    assert pooled["crossattn"].shape == (1, 1, 32)
    # same text -> same tokens -> same embedding; different text differs
    a, b = e({"text": ["same words"]})["crossattn"], e({"text": ["same words"]})["crossattn"]
    assert torch.equal(a, b) and not torch.equal(a, e({"text": ["other words"]})["crossattn"])
    with pytest.raises(Exception, match="Layer index is required"):
        ClipEmbedderConfig(layer="hidden")


def test_t5_embedder_mask_and_wrapper(tiny):
    t = T5TextEmbedder(T5TextEmbedderConfig(tokenizer_max_length=120, returns_attention_mask=True))
    o = t({"text": ["a cat", "a much longer caption with many more words in it"]})
    assert o["crossattn"].shape == (2, 120, 32) and o["attention_mask"].shape == (2, 120)
    assert int(o["attention_mask"][0].sum()) < int(o["attention_mask"][1].sum())
    assert torch.equal(o["attention_mask"][0].cumsum(0).argmax(), o["attention_mask"][0].sum() - 1)   # ones then zeros
    cw = ConditionerWrapper([ClipEmbedder(ClipEmbedderConfig(layer="last", always_return_pooled=True))])
    c = cw({"text": ["a", "b"]}, ucg_keys=["text"])["cond"]
    assert float(c["crossattn"].abs().max()) == 0 and c["vector"].shape == (2, 32)


def test_unknown_text_encoder_raises():
    with pytest.raises(ValueError, match="neither available locally nor a known architecture"):
        ClipEmbedder(ClipEmbedderConfig(version="somebody/unknown-clip"))


def test_data_pipeline_filters_mappers(tmp_path):
    import numpy as np
    from PIL import Image
    from flash.data.datasets import DataModule, DataModuleConfig
    from flash.data.filters import FilterOnCondition, FilterOnConditionConfig, KeyFilter, KeyFilterConfig
    from flash.data.mappers import (KeyRenameMapper, KeyRenameMapperConfig, KeysFromJSONMapper, KeysFromJSONMapperConfig,
                                    MapperWrapper, RemoveKeysMapper, RemoveKeysMapperConfig, RescaleMapper,
                                    RescaleMapperConfig, SelectKeysMapper, SelectKeysMapperConfig, TorchvisionMapper,
                                    TorchvisionMapperConfig)
    shard = str(tmp_path / "000000.tar")
    rng = np.random.default_rng(0)
    with tarfile.open(shard, "w") as tf:
        for i in range(5):
            buf = io.BytesIO()
            Image.fromarray(rng.integers(0, 255, (80, 96, 3), dtype=np.uint8)).save(buf, format="JPEG")
            items = [(f"{i:04d}.jpg", buf.getvalue()),
                     (f"{i:04d}.json", json.dumps({"caption": f"cap {i}", "aesthetic_score": 7.0 if i % 2 == 0 else 2.0}).encode())]
            if i == 3:
                items = items[:1]                       # no json: dropped by the KeyFilter
            for name, data in items:
                info = tarfile.TarInfo(name)
                info.size = len(data)
                tf.addfile(info, io.BytesIO(data))
    chain = [KeyFilter(KeyFilterConfig(keys=["jpg", "json"])), SelectKeysMapper(SelectKeysMapperConfig(keys=["jpg", "json"])),
             MapperWrapper([KeysFromJSONMapper(KeysFromJSONMapperConfig(key="json", keys_to_extract=["caption", "aesthetic_score"],
                                                                         remove_original=False, strict=False)),
                            KeyRenameMapper(KeyRenameMapperConfig(key_map={"jpg": "image", "caption": "text"})),
                            TorchvisionMapper(TorchvisionMapperConfig(key="image", transforms=["CenterCrop", "ToTensor", "Resize"],
                                                                      transforms_kwargs=[{"size": (64, 64)}, {}, {"size": (32, 32)}])),
                            RemoveKeysMapper(RemoveKeysMapperConfig(keys=["json"])), RescaleMapper(RescaleMapperConfig(key="image"))]),
             FilterOnCondition(FilterOnConditionConfig(condition_key="aesthetic_score", condition_fn=lambda x: x >= 6.0))]
    dm = DataModule(train_config=DataModuleConfig(shards_path_or_urls=[f"pipe:cat {shard}"], decoder="pil",
                                                  per_worker_batch_size=3, num_workers=0,
                                                  shuffle_after_filter_mappers_buffer_size=2), train_filters_mappers=chain)
    dm.setup()
    batches = list(dm.train_dataloader())
    assert len(batches) == 1                                       # samples 0, 2, 4 survive -> one batch of 3
    b = batches[0]
    assert b["image"].shape == (3, 3, 32, 32) and float(b["image"].min()) >= -1 and float(b["image"].max()) <= 1
    assert sorted(b["text"]) == ["cap 0", "cap 2", "cap 4"] and "json" not in b and b["aesthetic_score"].tolist() == [7.0] * 3


def test_compat_shims_expose_the_names_the_examples_import(monkeypatch):
    compat = os.path.join(ROOT, "flash-diffusion_b200", "compat")
    monkeypatch.setattr(sys, "path", sys.path + [compat])
    for m in [k for k in sys.modules if k.split(".")[0] in ("diffusers", "peft", "pytorch_lightning", "braceexpand", "lpips")]:
        monkeypatch.delitem(sys.modules, m)
    import braceexpand
    from diffusers import (DiffusionPipeline, DPMSolverMultistepScheduler, EulerAncestralDiscreteScheduler,  # noqa: F401
                           EulerDiscreteScheduler, FlashFlowMatchEulerDiscreteScheduler, FlowMatchEulerDiscreteScheduler,
                           LCMScheduler, StableDiffusion3Pipeline, StableDiffusionXLPipeline)
    from peft import LoraConfig, get_peft_model  # noqa: F401
    from pytorch_lightning import Trainer, loggers  # noqa: F401
    from pytorch_lightning.callbacks import Callback, ModelCheckpoint  # noqa: F401
    from pytorch_lightning.utilities import rank_zero_only  # noqa: F401
    assert list(braceexpand.braceexpand("/d/{000008..000010}.tar")) == ["/d/000008.tar", "/d/000009.tar", "/d/000010.tar"]
    assert list(braceexpand.braceexpand("a{b,c}d")) == ["abd", "acd"]
    s = EulerDiscreteScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", subfolder="scheduler")
    s.set_timesteps(4)
    assert len(s.timesteps) == 4 and s.init_noise_sigma > 1
    with pytest.raises(OSError):
        DiffusionPipeline.from_pretrained("somebody/unknown")
    from flash.models.adapters import DiffusersT2IAdapterWrapper
    with pytest.raises(NotImplementedError):
        DiffusersT2IAdapterWrapper()
