"""Golden record of the REFERENCE's own `WandbSampleLogger` (src/flash/trainer/loggers.py:18-138), imported unmodified
from /root/reference/src:   python tests/golden/make_reference_logger_golden.py  ->  tests/golden/reference_logger.pt

A recording `trainer.logger.experiment` and a recording `wandb` stand-in (Image / Table keep what they were given) capture
every `experiment.log(payload, step=)` call of `on_train_batch_end` / `on_validation_batch_end` for a fixed pipeline whose
`log_samples` returns image batches, strings, tensor lists, metric dicts and scalars; also the train / eval toggling and the
`log_batch_freq` gating."""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)


class RecImage:
    def __init__(self, img):
        import numpy as np
        self.kind, self.array = "Image", torch.from_numpy(np.array(img).copy())


class RecTable:
    def __init__(self, data=None, columns=None):
        self.kind, self.data, self.columns = "Table", data, columns


def wandb_stub():
    m = types.ModuleType("wandb")
    m.Image, m.Table = RecImage, RecTable
    return m


class Experiment:
    def __init__(self):
        self.calls = []

    def log(self, payload, step=None):
        self.calls.append((summarise(payload), step))


def summarise(v):
    if isinstance(v, dict):
        return {k: summarise(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [summarise(x) for x in v]
    if getattr(v, "kind", None) == "Image":
        a = v.array
        return ("Image", tuple(a.shape), str(a.dtype), int(a.long().sum()), a[::7, ::5].clone())
    if getattr(v, "kind", None) == "Table":
        return ("Table", v.data, v.columns)
    if hasattr(v, "shape") and not isinstance(v, torch.Tensor):        # numpy
        return ("ndarray", tuple(v.shape), str(v.dtype), torch.as_tensor(v).double().reshape(-1)[:16].clone())
    if isinstance(v, torch.Tensor):
        return ("tensor", tuple(v.shape), v.double().reshape(-1)[:16].clone())
    return v


def make_logs():
    g = torch.Generator().manual_seed(8)
    return {"samples": torch.rand(5, 3, 16, 12, generator=g) * 2.4 - 1.2,          # out-of-range values get clamped
            "text": ["a red car", "a blue pig"],
            "latents": [torch.randn(2, 2, generator=g), torch.randn(3, generator=g)],
            "metrics": {"psnr": torch.tensor(21.5), "n": 3},
            "scalar": torch.tensor(0.25), "vector": torch.arange(4.0), "count": 7, "ratio": 0.5, "ignored": None}


class Pipeline:
    def __init__(self):
        self.training, self.toggles, self.calls = True, [], 0

    def eval(self):
        self.training = False
        self.toggles.append("eval")

    def train(self):
        self.training = True
        self.toggles.append("train")

    def log_samples(self, batch):
        self.calls += 1
        assert not self.training
        return make_logs()


def run(logger_cls):
    out = {}
    trainer = types.SimpleNamespace(logger=types.SimpleNamespace(experiment=Experiment()), global_step=11)
    pipe = Pipeline()
    cb = logger_cls(log_batch_freq=3)
    outputs = {"loss": torch.tensor(1.5), "loss_optimizer_0": torch.tensor(0.75), "student_output": torch.zeros(2, 3, 8, 8)}
    cb.on_train_batch_end(trainer, pipe, dict(outputs), {"image": None}, 0)
    out["batch0"] = list(trainer.logger.experiment.calls)
    trainer.logger.experiment.calls = []
    trainer.global_step = 12
    cb.on_train_batch_end(trainer, pipe, dict(outputs), {"image": None}, 1)           # not a logging batch: outputs only
    out["batch1"] = list(trainer.logger.experiment.calls)
    trainer.logger.experiment.calls = []
    pipe.training = False
    cb.on_validation_batch_end(trainer, pipe, {"loss": torch.tensor(2.0)}, {"image": None}, 3)
    out["val3"] = list(trainer.logger.experiment.calls)
    out["toggles"], out["log_samples_calls"] = list(pipe.toggles), pipe.calls
    return out


def main():
    import make_reference_step_golden as G
    G.install_shims()
    sys.modules["wandb"] = wandb_stub()
    pl = sys.modules["pytorch_lightning"]
    if not hasattr(pl, "Trainer"):
        pl.Trainer = object
    cbm = sys.modules.get("pytorch_lightning.callbacks") or types.ModuleType("pytorch_lightning.callbacks")
    if not hasattr(cbm, "Callback"):
        cbm.Callback = object
    sys.modules["pytorch_lightning.callbacks"] = cbm
    um = sys.modules.get("pytorch_lightning.utilities") or types.ModuleType("pytorch_lightning.utilities")
    if not hasattr(um, "rank_zero_only"):
        um.rank_zero_only = lambda fn: fn
    sys.modules["pytorch_lightning.utilities"] = um
    sys.path.insert(0, G.REF_SRC)
    from flash.trainer.loggers import WandbSampleLogger
    import flash
    assert os.path.realpath(flash.__path__[0]).startswith(G.REF_SRC)
    out = run(WandbSampleLogger)
    out["generated_by"] = os.path.relpath(__file__, ROOT)
    out["reference_files"] = ["src/flash/trainer/loggers.py:18-138"]
    path = os.path.join(HERE, "reference_logger.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", {k: len(v) for k, v in out.items() if isinstance(v, list)})
    for c in out["batch0"]:
        print("  ", list(c[0].keys()), c[1])


if __name__ == "__main__":
    main()
