"""Minimal Lightning-shaped runtime for the example scripts (`pytorch_lightning` is not installable offline).

The reference drives `TrainingPipeline` (a `pl.LightningModule`) with `pl.Trainer(accelerator="gpu", devices=...,
strategy="ddp_find_unused_parameters_true", precision="bf16-mixed", callbacks=[WandbSampleLogger, ModelCheckpoint],
logger=WandbLogger(...)).fit(pipeline, data_module)` (examples/train_flash_sdxl.py:414-447).  Here `Trainer.fit` is the
explicit loop: one process per GPU (torchrun env or a single process), NCCL process group when WORLD_SIZE > 1,
`pipeline.configure_optimizers()`, batches moved to the device, `pipeline.training_step`, callbacks.  Mixed precision /
DDP wrapping are not needed: the denoisers compute in bf16 on their own kernels and the gradient all-reduce is the
trainer's flat bucket.  `FLASH_MAX_STEPS` (env) bounds a run (the scripts pass only `max_epochs`)."""
import json
import logging
import os
import time
from typing import Any, List

import torch
import torch.distributed as dist


def rank_zero_only(fn):
    def wrapped(*a, **k):
        if int(os.environ.get("RANK", "0")) == 0:
            return fn(*a, **k)
        return None
    wrapped.__name__ = getattr(fn, "__name__", "wrapped")
    return wrapped


class Callback:
    def on_train_start(self, trainer, pl_module):
        pass

    def on_train_batch_end(self, trainer, pl_module, outputs, batch, batch_idx):
        pass

    def on_validation_batch_end(self, trainer, pl_module, outputs, batch, batch_idx):
        pass

    def on_train_end(self, trainer, pl_module):
        pass


class _Experiment:
    """what `trainer.logger.experiment.log(dict, step=)` writes to: a jsonl file under save_dir (wandb needs a network)"""

    def __init__(self, path):
        self.path, self.rows = path, 0

    def log(self, data, step=None):
        def enc(v):
            if isinstance(v, (int, float, str)) or v is None:
                return v
            if isinstance(v, torch.Tensor):
                return v.tolist() if v.numel() <= 16 else f"tensor{tuple(v.shape)}"
            if hasattr(v, "tolist") and getattr(v, "size", 17) <= 16:
                return v.tolist()
            if isinstance(v, dict):
                return {k: enc(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return [enc(x) for x in v[:8]]
            return type(v).__name__
        if self.path:
            with open(self.path, "a") as f:
                f.write(json.dumps({"step": step, **{k: enc(v) for k, v in data.items()}}) + "\n")
        self.rows += 1


class WandbLogger:
    def __init__(self, project=None, offline=True, save_dir=None, name=None, **unused):
        self.project, self.name, self.save_dir = project, name, save_dir
        path = None
        if save_dir and int(os.environ.get("RANK", "0")) == 0:
            os.makedirs(save_dir, exist_ok=True)
            path = os.path.join(save_dir, "metrics.jsonl")
        self.experiment = _Experiment(path)

    def log_metrics(self, metrics, step=None):
        self.experiment.log(metrics, step=step)


class Trainer:
    def __init__(self, accelerator="gpu", devices=1, num_nodes=1, strategy=None, default_root_dir=None, max_epochs=1,
                 max_steps=-1, logger=None, callbacks: List[Any] = None, num_sanity_val_steps=0, precision=None,
                 check_val_every_n_epoch=1, **unused):
        self.accelerator, self.devices, self.num_nodes = accelerator, devices, num_nodes
        self.default_root_dir, self.max_epochs = default_root_dir, max_epochs
        env = os.environ.get("FLASH_MAX_STEPS")
        self.max_steps = int(env) if env is not None else max_steps
        self.logger = logger if logger is not None else WandbLogger(save_dir=default_root_dir)
        self.callbacks = list(callbacks or [])
        self.global_step = 0
        self.current_epoch = 0
        self.global_rank = int(os.environ.get("RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))

    @staticmethod
    def _to_device(batch, device):
        if isinstance(batch, dict):
            return {k: Trainer._to_device(v, device) for k, v in batch.items()}
        if isinstance(batch, torch.Tensor):
            return batch.to(device, non_blocking=True)
        return batch

    def fit(self, model, datamodule=None, train_dataloaders=None):
        use_cuda = self.accelerator in ("gpu", "cuda", "auto") and torch.cuda.is_available()
        local = int(os.environ.get("LOCAL_RANK", "0"))
        device = torch.device("cuda", local) if use_cuda else torch.device("cpu")
        if use_cuda:
            torch.cuda.set_device(device)
        if self.world_size > 1 and not dist.is_initialized():
            dist.init_process_group("nccl" if use_cuda else "gloo", **({"device_id": device} if use_cuda else {}))
        plumbing_on_meta = self.max_steps == 0 and any(p.is_meta for p in model.parameters())
        if not plumbing_on_meta:        # FLASH_MAX_STEPS=0 under torch.device("meta"): shapes only, nothing to move
            model.to(device)
        model.trainer = self
        if hasattr(model, "configure_optimizers") and getattr(model, "optims", None) is None:
            model.configure_optimizers()
        if datamodule is not None:
            datamodule.setup("fit")
            loader = datamodule.train_dataloader()
        else:
            loader = train_dataloaders
        if hasattr(model, "on_train_start"):
            model.on_train_start()
        def fire(hook, *a):
            for cb in self.callbacks:
                fn = getattr(cb, hook, None)          # callbacks need not derive from Callback (ModelCheckpoint)
                if fn is not None:
                    fn(self, model, *a)

        fire("on_train_start")
        model.train()
        t0 = time.perf_counter()
        done = self.max_steps == 0
        summary = {"steps": 0, "device": str(device), "losses": [], "sanity_batch": None}
        if done and loader is not None:
            # plumbing-only run (FLASH_MAX_STEPS=0): pull ONE batch through the data pipeline, run no step
            with torch.device("cpu"):        # the data pipeline is host work whatever the ambient default device
                first = next(iter(loader))
            summary["sanity_batch"] = {k: (list(v.shape) if isinstance(v, torch.Tensor) else f"{type(v).__name__}[{len(v)}]")
                                       for k, v in first.items()}
        for epoch in range(self.max_epochs):
            self.current_epoch = epoch
            if done:
                break
            for batch_idx, batch in enumerate(loader):
                batch = self._to_device(batch, device)
                outputs = model.training_step(batch, batch_idx)
                self.global_step += 1
                if hasattr(model, "step_lr_schedulers"):
                    model.step_lr_schedulers("step", self.global_step)
                summary["losses"].append({k: float(v) for k, v in outputs.items()
                                          if k.startswith("loss") and (isinstance(v, (int, float)) or
                                                                       (isinstance(v, torch.Tensor) and v.numel() == 1))})
                if hasattr(model, "on_train_batch_end"):      # the LightningModule's own hook first, then the callbacks
                    model.on_train_batch_end(outputs, batch, batch_idx)
                fire("on_train_batch_end", outputs, batch, batch_idx)
                if self.max_steps > 0 and self.global_step >= self.max_steps:
                    done = True
                    break
            if hasattr(model, "step_lr_schedulers") and not done:       # only an epoch that ran all its batches counts
                model.step_lr_schedulers("epoch", epoch + 1)
        fire("on_train_end")
        summary["steps"] = self.global_step
        if self.default_root_dir and self.global_rank == 0:
            os.makedirs(self.default_root_dir, exist_ok=True)
            with open(os.path.join(self.default_root_dir, "fit_summary.json"), "w") as f:
                json.dump(summary, f)
        logging.info(f"Trainer.fit: {self.global_step} steps in {time.perf_counter() - t0:.1f} s")
        return self
