"""Sample logger callback (reference src/flash/trainer/loggers.py:18-138: `WandbSampleLogger(log_batch_freq)`): every
`log_batch_freq` batches the pipeline's `log_samples(batch)` output and every step's scalar outputs go to
`trainer.logger.experiment.log({f"{key}/{split}": value}, step=trainer.global_step)`; 4-D tensors become image grids
(`make_grid(nrow=4)`, [-1, 1] -> uint8), lists of strings a text table.  wandb images / tables are built when `wandb`
is importable, plain arrays / lists otherwise (the offline logger of flash.trainer.lightning writes jsonl)."""
import logging
from typing import Any, Dict

import torch

from .lightning import Callback, rank_zero_only


class WandbSampleLogger(Callback):
    def __init__(self, log_batch_freq: int = 100):
        super().__init__()
        self.log_batch_freq = log_batch_freq

    def on_train_batch_end(self, trainer, pl_module, outputs: Dict[str, Any], batch: Any, batch_idx: int) -> None:
        self.log_samples(trainer, pl_module, outputs, batch, batch_idx, split="train")
        self._process_logs(trainer, outputs, split="train")

    def on_validation_batch_end(self, trainer, pl_module, outputs: Dict[str, Any], batch: Any, batch_idx: int) -> None:
        self.log_samples(trainer, pl_module, outputs, batch, batch_idx, split="val")
        self._process_logs(trainer, outputs, split="val")

    @rank_zero_only
    @torch.no_grad()
    def log_samples(self, trainer, pl_module, outputs, batch, batch_idx, split="train") -> None:
        if not hasattr(pl_module, "log_samples"):
            logging.warning("log_img method not found in LightningModule. Skipping image logging.")
            return
        if batch_idx % self.log_batch_freq == 0:
            was_training = pl_module.training
            if was_training:
                pl_module.eval()
            logs = pl_module.log_samples(batch)
            self._process_logs(trainer, logs, split=split)
            if was_training:
                pl_module.train()

    @rank_zero_only
    def _process_logs(self, trainer, logs: Dict[str, Any], rescale=True, split="train") -> Dict[str, Any]:
        if not logs:
            return logs
        try:
            import wandb
            from PIL import Image
        except Exception:           # noqa: BLE001
            wandb = Image = None
        exp = trainer.logger.experiment
        for key, value in list(logs.items()):
            if isinstance(value, torch.Tensor):
                value = value.detach().cpu()
                if value.dim() == 4:
                    from torchvision.utils import make_grid
                    images = (value.float() + 1.0) / 2.0 if rescale else value.float()
                    grid = make_grid(images, nrow=4).permute(1, 2, 0).mul(255).clamp(0, 255).to(torch.uint8).numpy()
                    logs[key] = grid
                    payload = [wandb.Image(Image.fromarray(grid))] if wandb is not None else f"image{grid.shape}"
                    exp.log({f"{key}/{split}": payload}, step=trainer.global_step)
                elif value.dim() <= 1:
                    exp.log({f"{key}/{split}": value.float().numpy()}, step=trainer.global_step)
            elif isinstance(value, list) and value:
                if isinstance(value[0], str):
                    payload = (wandb.Table(data=[[c] for c in value], columns=["text"]) if wandb is not None else value)
                    exp.log({f"{key}/{split}": payload}, step=trainer.global_step)
                elif isinstance(value[0], torch.Tensor):
                    exp.log({f"{key}/{split}": [v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v
                                                 for v in value]}, step=trainer.global_step)
            elif isinstance(value, dict):
                exp.log({f"{key}/{split}": {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v)
                                             for k, v in value.items()}}, step=trainer.global_step)
            elif isinstance(value, (int, float)):
                exp.log({f"{key}/{split}": value}, step=trainer.global_step)
        return logs
