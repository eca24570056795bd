"""Step driver — mirrors reference `TrainingPipeline` (src/flash/trainer/trainer.py:16-251) without Lightning.

Semantics kept from the reference:
  * `configure_optimizers` (:76-139): optimizer i owns the parameters whose NAME regex-matches
    `trainable_params[i]` and that require grad; everything matched by no regex is frozen;
  * `training_step` (:169-218): with N>1 optimizers, for each optimizer i: a FULL `model(batch, step=i)`
    forward, `zero_grad`, backward of `loss[i]`, `opt.step()`.
What Lightning's DDP strategy did implicitly (`strategy="ddp_find_unused_parameters_true"`,
examples/train_flash_sdxl.py:427) is explicit here: the gradients of optimizer i's parameters live in ONE
flat fp32 bucket that is all-reduced (mean) across ranks with a single `torch.distributed.all_reduce`
(NCCL over NVLink on B200; gloo in CPU tests) between backward and `opt.step()`.
"""
import importlib
import logging
import re
import time
from typing import Any, Dict, List

import torch
import torch.distributed as dist

from .training_config import TrainingConfig


class _FlatGradBucket:
    """All gradients of one optimizer as views into one contiguous fp32 buffer.

    Data parallel: the buffer is all-reduced (mean) across ranks.  On CUDA the reduction is OVERLAPPED
    with the backward: the buffer is cut into `n_chunks` contiguous chunks; a post-accumulate hook on every parameter
    counts the chunk's outstanding gradients and, when a chunk is complete (backward fills the buffer roughly back to
    front), its NCCL all-reduce is enqueued on a side stream while the remaining dX / dA / dB kernels keep the compute
    stream busy.  `finish()` reduces whatever is left, makes the compute stream wait for the side stream and divides by
    the world size.  CUDA events bracket every chunk, so `stats()` can say how much of the all-reduce time was hidden
    behind the backward and how much was exposed before `opt.step()`.  CPU tensors (gloo tests): one blocking
    all-reduce in `finish()`."""

    def __init__(self, params: List[torch.nn.Parameter], n_chunks: int = 4):
        self.params = params
        self.flat = None
        self.n_chunks = n_chunks
        self.chunks = []            # (lo, hi) element ranges of the flat buffer
        self.chunk_of = {}          # id(param) -> chunk index
        self.pending = []
        self.launched = []
        self.hooks_installed = False
        self.side = None
        self.events = []            # per step: (e_backward_end, [(e_start, e_end) per chunk])
        self.overlap = True

    @staticmethod
    def _dp():
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _layout(self):
        n = sum(p.numel() for p in self.params)
        target = max(1, (n + self.n_chunks - 1) // self.n_chunks)
        self.chunks, self.chunk_of = [], {}
        lo = off = 0
        for p in self.params:
            self.chunk_of[id(p)] = len(self.chunks)
            off += p.numel()
            if off - lo >= target:
                self.chunks.append((lo, off))
                lo = off
        if off > lo:
            self.chunks.append((lo, off))

    def attach(self):
        if not self.params:
            return
        dev = self.params[0].device
        if self.flat is None or self.flat.device != dev:
            n = sum(p.numel() for p in self.params)
            self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
            self._layout()
        else:
            self.flat.zero_()
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        counts = [0] * len(self.chunks)
        for p in self.params:
            counts[self.chunk_of[id(p)]] += 1
        self.pending, self.launched = counts, [False] * len(self.chunks)
        self._cur = None
        if self._dp() and self.flat.is_cuda and self.overlap:
            if self.side is None:
                self.side = torch.cuda.Stream(device=dev)
            if not self.hooks_installed:
                for p in self.params:
                    p.register_post_accumulate_grad_hook(self._on_grad)
                self.hooks_installed = True
            self._cur = []

    def _on_grad(self, p):
        if self._cur is None or id(p) not in self.chunk_of:
            return
        c = self.chunk_of[id(p)]
        self.pending[c] -= 1
        if self.pending[c] == 0 and not self.launched[c]:
            self._launch(c)

    def _launch(self, c):
        lo, hi = self.chunks[c]
        self.launched[c] = True
        main = torch.cuda.current_stream(self.flat.device)
        self.side.wait_stream(main)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.side):
            e0.record()
            dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM)
            e1.record()
        self._cur.append((e0, e1))

    def finish(self):
        """all-reduce (mean) complete on the compute stream when this returns control to the optimizer step."""
        if self.flat is None or not self._dp():
            return
        world = dist.get_world_size()
        if self._cur is None:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(world)
            return
        main = torch.cuda.current_stream(self.flat.device)
        e_bwd = torch.cuda.Event(enable_timing=True)
        e_bwd.record(main)
        for c in range(len(self.chunks)):
            if not self.launched[c]:
                self._launch(c)
        main.wait_stream(self.side)
        self.flat.div_(world)
        self.events.append((e_bwd, self._cur))
        if len(self.events) > 64:
            self.events = self.events[-64:]
        self._cur = None

    all_reduce_mean = finish

    def stats(self):
        """{"steps", "ms_total", "ms_exposed"}: per-step means of the all-reduce time on the side stream and of the
        part of it that ran after the backward had finished.  Synchronises the device."""
        if not self.events:
            return None
        torch.cuda.synchronize()
        tot = exp = 0.0
        for e_bwd, pairs in self.events:
            tot += sum(a.elapsed_time(b) for a, b in pairs)
            exp += max(0.0, e_bwd.elapsed_time(pairs[-1][1])) if pairs else 0.0
        n = len(self.events)
        return {"steps": n, "ms_total": tot / n, "ms_exposed": exp / n, "chunks": len(self.chunks),
                "bytes": 4 * self.flat.numel()}


class TrainingPipeline(torch.nn.Module):
    def __init__(self, model, pipeline_config: TrainingConfig, verbose: bool = False, **kwargs):
        super().__init__()
        self.model = model
        self.pipeline_config = pipeline_config
        self.log_samples_model_kwargs = pipeline_config.log_samples_model_kwargs
        self.verbose = verbose
        log_keys = pipeline_config.log_keys
        self.log_keys = [log_keys] if isinstance(log_keys, str) else (log_keys or [])
        self.automatic_optimization = True
        self.optims = None
        self._buckets = None
        self.global_rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self.timer = None

    @property
    def device(self):
        return next(self.model.parameters()).device

    @property
    def allreduce_stats(self):
        """per optimizer: gradient all-reduce time per step, total vs exposed after the backward."""
        if not self._buckets:
            return None
        out = [b.stats() for b in self._buckets]
        return out if any(o is not None for o in out) else None

    # ---- reference :76-139
    def configure_optimizers(self):
        cfg = self.pipeline_config
        optimizers, buckets = [], []
        named = list(self.model.named_parameters())
        for i, opt_name in enumerate(cfg.optimizers_name):
            patterns = [re.compile(rx) for rx in cfg.trainable_params[i]]
            params = []
            for name, p in named:
                for pat in patterns:          # a parameter matched by two regexes is listed twice, as upstream
                    if re.match(pat, name) and p.requires_grad:
                        params.append(p)
            uniq, seen = [], set()
            for p in params:
                if id(p) not in seen:
                    seen.add(id(p))
                    uniq.append(p)
            logging.info(f"Number of trainable parameters for optimizer {i}: {sum(p.numel() for p in uniq)}")
            opt_cls = getattr(importlib.import_module("torch.optim"), opt_name)
            optimizers.append(opt_cls([{"params": uniq}], lr=cfg.learning_rates[i], **cfg.optimizers_kwargs[i]))
            buckets.append(_FlatGradBucket(uniq))
        if len(optimizers) > 1:
            self.automatic_optimization = False
        all_patterns = [re.compile(rx) for rxs in cfg.trainable_params for rx in rxs]
        for name, p in named:
            if not any(re.match(pat, name) for pat in all_patterns if p.requires_grad):
                p.requires_grad = False
        logging.info(f"Number of trainable parameters: {sum(p.numel() for p in self.model.parameters() if p.requires_grad)}")
        self.optims, self._buckets = optimizers, buckets
        self.lr_schedulers = self.configure_lr_schedulers()
        self.sync_replicas()
        if self.lr_schedulers is None:
            return optimizers
        return optimizers, list(self.lr_schedulers)          # Lightning's (optimizers, lr_scheduler configs) form

    def step_lr_schedulers(self, interval: str, count: int):
        """What Lightning does with the configs returned above under AUTOMATIC optimisation (one optimizer): a scheduler
        whose `interval` is "step" / "epoch" advances every `frequency` optimizer steps / epochs.  With several
        optimizers the reference runs manual optimisation and never calls `scheduler.step()` itself
        (trainer.py:169-218), so — like under Lightning — its schedulers stay where they started."""
        if not self.automatic_optimization or not self.lr_schedulers:
            return
        for cfg in self.lr_schedulers:
            if cfg is not None and cfg["interval"] == interval and count % max(int(cfg["frequency"]), 1) == 0:
                cfg["scheduler"].step()

    def sync_replicas(self):
        """What Lightning's DDP wrap does at construction: every rank starts from rank 0's parameters and buffers
        (LoRA A and the discriminator are randomly initialised per process, so without this the replicas would apply
        averaged gradients to different weights forever).  One flat broadcast per dtype/device group."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        groups = {}
        with torch.no_grad():
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                groups.setdefault((t.dtype, t.device), []).append(t)
            # frozen tensors are built from the same seed on every rank; broadcasting them too costs one pass over the
            # replica at start-up and removes the assumption.  Chunks of <= 256 Mi elements bound the staging buffer.
            for (dtype, dev), ts in groups.items():
                chunk, n = [], 0
                for t in ts + [None]:
                    if t is not None and (n == 0 or n + t.numel() <= (1 << 28)):
                        chunk.append(t)
                        n += t.numel()
                        continue
                    flat = torch.cat([c.detach().reshape(-1) for c in chunk])
                    dist.broadcast(flat, src=0)
                    off = 0
                    for c in chunk:
                        c.copy_(flat[off:off + c.numel()].view_as(c))
                        off += c.numel()
                    chunk, n = ([t], t.numel()) if t is not None else ([], 0)

    def configure_lr_schedulers(self):
        cfg = self.pipeline_config
        out = []
        for i, name in enumerate(cfg.lr_schedulers_name):
            if name is None:
                out.append(None)
                continue
            cls = getattr(importlib.import_module("torch.optim.lr_scheduler"), name)
            out.append({"scheduler": cls(self.optims[i], **cfg.lr_schedulers_kwargs[i]),
                        "interval": cfg.lr_schedulers_interval[i], "monitor": "val_loss",
                        "frequency": cfg.lr_schedulers_frequency[i]})
        return None if all(s is None for s in out) else out

    def optimizers(self):
        if self.optims is None:
            self.configure_optimizers()
        return self.optims

    # ---- reference :169-218
    def training_step(self, train_batch: Dict[str, Any], batch_idx: int = 0, draws=None) -> dict:
        optimizers = self.optimizers()
        outputs = {"batch_idx": batch_idx}
        if self.automatic_optimization:
            out = self.model(train_batch, device=self.device)
            loss = out["loss"]
            loss = loss[0] if isinstance(loss, (list, tuple)) else loss
            self._buckets[0].attach()
            loss.backward()
            self._buckets[0].all_reduce_mean()
            optimizers[0].step()
            return {"loss": loss.detach(), "batch_idx": batch_idx, "start_timestep": out.get("start_timestep")}
        for i, opt in enumerate(optimizers):
            kw = {} if draws is None else {"draws": draws[i] if isinstance(draws, (list, tuple)) else draws}
            model_output = self.model(train_batch, device=self.device, step=i, batch_idx=batch_idx, **kw)
            loss = model_output["loss"]
            if "start_timestep" in model_output:
                outputs["start_timestep"] = model_output["start_timestep"]
            li = loss[i]
            outputs[f"loss_optimizer_{i}"] = li.detach() if torch.is_tensor(li) else li
            self._buckets[i].attach()                     # == opt.zero_grad(), grads land in the flat bucket
            if torch.is_tensor(li) and li.requires_grad:
                # toggle_optimizer semantics: only optimizer i's parameters accumulate gradients
                others = [p for j, b in enumerate(self._buckets) if j != i for p in b.params]
                flags = [p.requires_grad for p in others]
                for p in others:
                    p.requires_grad_(False)
                try:
                    li.backward()
                finally:
                    for p, f in zip(others, flags):
                        p.requires_grad_(f)
            self._buckets[i].all_reduce_mean()
            opt.step()
        return outputs

    def on_train_start(self):
        if self.global_rank == 0:
            self.timer = time.perf_counter()

    def on_train_batch_end(self, outputs: Dict[str, Any], batch: Any, batch_idx: int) -> None:
        """reference trainer.py:62-74 — forward the hook to the model, and on rank 0 log the running average wall-clock
        seconds per batch every 10 batches (the reference's only step timer)."""
        self.model.on_train_batch_end(batch)
        if self.global_rank == 0 and batch_idx % 10 == 0 and self.timer is not None:
            delta = time.perf_counter() - self.timer
            logging.info(f"Average time per batch {batch_idx} took {delta / (batch_idx + 1)} seconds")

    def validation_step(self, val_batch, val_idx=0):
        loss = self.model(val_batch, device=self.device)["loss"]
        return {"loss": loss, "metrics": self.model.compute_metrics(val_batch)}

    def log_samples(self, batch):
        logs = self.model.log_samples(batch, device=self.device, **self.log_samples_model_kwargs)
        N = min(logs[k].shape[0] for k in logs) if logs is not None else 0
        for key in self.log_keys:
            if key in batch:
                logs = logs if logs is not None else {}
                logs[key] = batch[key][:N] if N > 0 else batch[key]
        return logs
