"""Shim of the `pytorch_lightning` names used by the example scripts and `flash.trainer` (see ../README.md)."""
import torch

from flash.trainer.lightning import Trainer  # noqa: F401

from . import callbacks, loggers, utilities  # noqa: F401


class LightningModule(torch.nn.Module):
    pass


class LightningDataModule:
    pass
