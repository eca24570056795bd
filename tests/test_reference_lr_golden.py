"""`TrainingPipeline.configure_optimizers` return forms and frozen sets against a record of the REFERENCE's own method
(tests/golden/reference_lr.pt from the unmodified src/flash/trainer/trainer.py:76-167), and the stepping of the learning-
rate schedulers by the Lightning-shaped `Trainer` (automatic optimisation: by `interval` / `frequency`; several
optimizers = the reference's manual optimisation, which never steps them)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = torch.load(os.path.join(HERE, "golden", "reference_lr.pt"), weights_only=False)


def test_configure_optimizers_matches_reference_run():
    import make_reference_lr_golden as G
    from flash.trainer import TrainingConfig, TrainingPipeline
    got = G.run(TrainingPipeline, TrainingConfig)
    for name in G.CASES:
        assert got[name] == GOLD[name], (name, got[name], GOLD[name])


class _Model(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a, self.b = torch.nn.Linear(2, 1), torch.nn.Linear(2, 1)

    def forward(self, batch, device=None, step=0, batch_idx=0, **kw):
        la, lb = self.a(batch["x"]).pow(2).mean(), self.b(batch["x"]).pow(2).mean()
        return {"loss": [la, lb]} if kw.get("two", self.two) else {"loss": la}

    def on_train_batch_end(self, batch, *a, **k):
        pass


def _fit(cfg_kw, two, steps):
    from flash.trainer import TrainingConfig, TrainingPipeline
    from flash.trainer.lightning import Trainer
    m = _Model()
    m.two = two
    pipe = TrainingPipeline(model=m, pipeline_config=TrainingConfig(**cfg_kw))
    loader = [{"x": torch.randn(4, 2)} for _ in range(3)]            # 3 batches per epoch
    Trainer(accelerator="cpu", max_epochs=10, max_steps=steps).fit(pipe, train_dataloaders=loader)
    return pipe


def test_trainer_steps_lr_schedulers_like_lightning():
    import make_reference_lr_golden as G
    pipe = _fit(G.CASES["one_opt_steplr"], two=False, steps=5)        # StepLR(step_size=1, gamma=.5) every 2nd step
    assert abs(pipe.optims[0].param_groups[0]["lr"] - 0.1 * 0.5 ** 2) < 1e-12
    kw = dict(G.CASES["one_opt_steplr"], lr_schedulers_interval=["epoch"], lr_schedulers_frequency=[1])
    pipe = _fit(kw, two=False, steps=7)                               # 7 steps = 2 full epochs + 1 step
    assert abs(pipe.optims[0].param_groups[0]["lr"] - 0.1 * 0.5 ** 2) < 1e-12
    pipe = _fit(G.CASES["two_opt_mixed"], two=True, steps=7)          # manual optimisation: schedulers untouched
    assert abs(pipe.optims[1].param_groups[0]["lr"] - 1e-2) < 1e-15
