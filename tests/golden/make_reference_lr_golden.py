"""Golden record of what the REFERENCE's own `TrainingPipeline.configure_optimizers` / `configure_lr_schedulers`
(src/flash/trainer/trainer.py:76-167) return for three optimizer / scheduler configurations, imported unmodified:
    python tests/golden/make_reference_lr_golden.py  ->  tests/golden/reference_lr.pt"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)

CASES = {
    "one_opt_steplr": dict(optimizers_name=["SGD"], learning_rates=[0.1], trainable_params=[["a"]],
                           lr_schedulers_name=["StepLR"], lr_schedulers_kwargs=[{"step_size": 1, "gamma": 0.5}],
                           lr_schedulers_interval=["step"], lr_schedulers_frequency=[2]),
    "two_opt_mixed": dict(optimizers_name=["AdamW", "SGD"], learning_rates=[1e-3, 1e-2], trainable_params=[["a"], ["b"]],
                          lr_schedulers_name=[None, "ExponentialLR"], lr_schedulers_kwargs=[{}, {"gamma": 0.9}],
                          lr_schedulers_interval=["step", "epoch"], lr_schedulers_frequency=[1, 3]),
    "two_opt_none": dict(optimizers_name=["AdamW", "AdamW"], learning_rates=[1e-3, 1e-3], trainable_params=[["a"], ["b"]]),
}


class _Cfg:
    def to_dict(self):
        return {"name": "Toy"}


class Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a, self.b, self.c = torch.nn.Linear(2, 2), torch.nn.Linear(2, 2), torch.nn.Linear(2, 2)
        self.config = _Cfg()          # the reference pipeline stores `model.config.to_dict()` as a hyper-parameter


def describe(ret):
    def sched(c):
        if c is None:
            return None
        return {k: (type(v).__name__ if k == "scheduler" else v) for k, v in c.items()}
    if isinstance(ret, tuple):
        opts, scheds = ret
        return dict(form="tuple", optimizers=[type(o).__name__ for o in opts], schedulers=[sched(c) for c in scheds])
    return dict(form="list", optimizers=[type(o).__name__ for o in ret], schedulers=None)


def run(pipeline_cls, config_cls):
    out = {}
    for name, kw in CASES.items():
        model = Toy()
        pipe = pipeline_cls(model=model, pipeline_config=config_cls(**kw))
        ret = pipe.configure_optimizers()
        out[name] = dict(describe(ret), automatic=bool(pipe.automatic_optimization),
                         frozen=sorted(n for n, p in model.named_parameters() if not p.requires_grad))
    return out


def main():
    import make_reference_step_golden as G
    G.install_shims()
    sys.path.insert(0, G.REF_SRC)
    from flash.trainer.trainer import TrainingPipeline
    from flash.trainer.training_config import TrainingConfig
    import flash
    assert os.path.realpath(flash.__path__[0]).startswith(G.REF_SRC)
    out = run(TrainingPipeline, TrainingConfig)
    out["generated_by"] = os.path.relpath(__file__, ROOT)
    torch.save(out, os.path.join(HERE, "reference_lr.pt"))
    for k, v in out.items():
        print(k, v)


if __name__ == "__main__":
    main()
