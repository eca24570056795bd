"""`flash.models.utils` (Tiler, pad, update_ema, extract_into_tensor) against vectors the REFERENCE's own functions
produced (tests/golden/reference_tiler.pt, written by tests/golden/make_reference_tiler_golden.py from the unmodified
src/flash/models/utils.py:12-377): same tile grid, same geometry attributes, same merged images for the three merge
methods (average / gaussian / linear), incl. partial trailing tiles, an axis that is not tiled and zero overlap."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = torch.load(os.path.join(HERE, "golden", "reference_tiler.pt"), weights_only=False)


def test_tiler_pad_ema_match_reference_run():
    import make_reference_tiler_golden as G
    from flash.models.utils import Tiler, extract_into_tensor, pad, update_ema
    got = G.run(Tiler, pad, update_ema, extract_into_tensor)
    assert len(got["cases"]) == len(GOLD["cases"]) == 4
    for g, r in zip(got["cases"], GOLD["cases"]):
        assert g["tile_shapes"] == r["tile_shapes"], r["case"]
        assert g["geometry"] == r["geometry"], r["case"]
        for m in G.METHODS:
            a, b = g["merged"][m], r["merged"][m]
            assert a["shape"] == b["shape"], (r["case"], m)
            assert torch.allclose(a["blocks"], b["blocks"], rtol=1e-5, atol=1e-6), (r["case"], m)
            assert torch.allclose(a["proj"], b["proj"], rtol=1e-5, atol=1e-4), (r["case"], m, a["proj"], b["proj"])
    assert got["pad"] == GOLD["pad"]
    for a, b in zip(got["ema_out"], GOLD["ema_out"]):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)
    assert torch.equal(got["extract"], GOLD["extract"])


def test_tiler_rejects_unknown_method_and_oversized_overlap():
    from flash.models.utils import TILING_METHODS, Tiler
    t = Tiler()
    tiles = t.get_tiles(torch.zeros(1, 3, 8, 8), (4, 4), (1, 1))
    with pytest.raises(ValueError):
        t.merge_tiles(tiles, tiling_method="median")
    with pytest.raises(AssertionError):
        t.get_tiles(torch.zeros(1, 3, 8, 8), (4, 4), (5, 1))
    assert TILING_METHODS == ["average", "gaussian", "linear"]


def test_conditioner_sanity_check_and_trainer_hook():
    """`ConditionerWrapper.conditioner_sanity_check` (reference conditioners_wrapper.py:32-37) and
    `TrainingPipeline.on_train_batch_end` (trainer.py:62-74: forwards to the model's hook, logs the running average)."""
    import logging
    from flash.models.embedders import ConditionerWrapper, TimestepsEmbedder, TimestepsEmbedderConfig
    w = ConditionerWrapper([TimestepsEmbedder(TimestepsEmbedderConfig(input_key="a")),
                            TimestepsEmbedder(TimestepsEmbedderConfig(input_key="b"))])
    w.ucg_keys = ["a"]
    w.conditioner_sanity_check()
    w.ucg_keys = ["a", "text"]
    with pytest.raises(AssertionError):
        w.conditioner_sanity_check()

    from flash.trainer import TrainingConfig, TrainingPipeline

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(2, 2)
            self.seen = []

        def on_train_batch_end(self, batch, *a, **k):
            self.seen.append(batch)
    m = M()
    pipe = TrainingPipeline(model=m, pipeline_config=TrainingConfig(trainable_params=[["lin"]]))
    pipe.on_train_start()
    records = []
    h = logging.Handler()
    h.emit = lambda rec: records.append(rec.getMessage())
    logging.getLogger().addHandler(h)
    old = logging.getLogger().level
    logging.getLogger().setLevel(logging.INFO)
    try:
        pipe.on_train_batch_end({}, {"x": 1}, 0)
        pipe.on_train_batch_end({}, {"x": 2}, 3)
    finally:
        logging.getLogger().removeHandler(h)
        logging.getLogger().setLevel(old)
    assert m.seen == [{"x": 1}, {"x": 2}]
    assert sum("Average time per batch 0 took" in r for r in records) == 1 and not any("batch 3 " in r for r in records)
