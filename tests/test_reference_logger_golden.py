"""`flash.trainer.loggers.WandbSampleLogger` against a record of the REFERENCE's own callback
(tests/golden/reference_logger.pt, written by tests/golden/make_reference_logger_golden.py from the unmodified
src/flash/trainer/loggers.py:18-138): the same sequence of `experiment.log(payload, step)` calls — keys, steps, image
grids (make_grid nrow 4, [-1, 1] -> uint8 with clamping), text tables, tensor lists, metric dicts, scalars — the same
`log_batch_freq` gating and the same eval / train toggling around `log_samples`."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = torch.load(os.path.join(HERE, "golden", "reference_logger.pt"), weights_only=False)


def _same(a, b, path=""):
    assert type(a) is type(b) or (isinstance(a, (list, tuple)) and isinstance(b, (list, tuple))), (path, type(a), type(b))
    if isinstance(a, dict):
        assert list(a) == list(b), (path, list(a), list(b))
        for k in a:
            _same(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), (path, len(a), len(b))
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}[{i}]")
    elif isinstance(a, torch.Tensor):
        assert a.shape == b.shape and a.dtype == b.dtype, path
        assert torch.equal(a, b) if not a.is_floating_point() else torch.allclose(a, b, rtol=1e-6, atol=1e-7), path
    else:
        assert a == b, (path, a, b)


def test_sample_logger_matches_reference_run(monkeypatch):
    import make_reference_logger_golden as G
    monkeypatch.setitem(sys.modules, "wandb", G.wandb_stub())
    from flash.trainer.loggers import WandbSampleLogger
    got = G.run(WandbSampleLogger)
    assert got["toggles"] == GOLD["toggles"] == ["eval", "train"] and got["log_samples_calls"] == GOLD["log_samples_calls"]
    for phase in ("batch0", "batch1", "val3"):
        assert len(got[phase]) == len(GOLD[phase]), (phase, [list(c[0]) for c in got[phase]], [list(c[0]) for c in GOLD[phase]])
        for i, (g, r) in enumerate(zip(got[phase], GOLD[phase])):
            _same(g, r, f"{phase}[{i}]")
