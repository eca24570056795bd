"""Checkpoint-surgery helpers of the reference's `flash.trainer.utils` (src/flash/trainer/utils.py: `setup_logging` :10-38,
`StateDictAdapter` :41-180, `StateDictRenamer` :183-222) — host-side utilities used when a teacher checkpoint is loaded
into a model whose conditioning widths differ (extra `concat` input channels, a wider `class_embedding`, a different
cross-attention width).  Same call signatures and results; tests/test_reference_utils_golden.py replays runs of the
reference's own classes."""
import logging
import os
import re
import time
from typing import Dict, List, Literal, Optional, Tuple

import torch

_FORMAT = "%(asctime)s [%(levelname)s]  %(message)s"


def setup_logging(output_dir: str, dir_logs: str, logger_id: str, logger_name: str = "logger",
                  level=logging.INFO) -> Tuple[logging.Logger, str]:
    """A logger writing to `<output_dir>/<dir_logs>/<logger_id>.log` (truncated) and to the console."""
    folder = os.path.join(output_dir, dir_logs)
    os.makedirs(folder, exist_ok=True)
    log_file = os.path.join(folder, logger_id + ".log")
    logger = logging.getLogger(logger_name)
    for handler in (logging.FileHandler(log_file, mode="w"), logging.StreamHandler()):
        handler.setFormatter(logging.Formatter(_FORMAT))
        logger.addHandler(handler)
    logger.setLevel(level)
    logger.info(f"Logging in {log_file}")
    return logger, log_file


class StateDictAdapter:
    """Makes the tensors of `checkpoint_state_dict` whose names match one of `regex_keys` fit the shapes the model
    expects: a rank mismatch between [a] and [a, b] is bridged by (un)squeezing the second axis, then every axis is
    either cut (`narrow`) or grown by a block of zeros / of normal noise with the tensor's own mean and std."""

    def _create_block(self, shape: List[int], strategy: Literal["zeros", "normal"], input: torch.Tensor = None):
        if strategy == "zeros":
            return torch.zeros(shape)
        if strategy == "normal":
            if input is None:
                return torch.randn(shape)
            return torch.randn(shape) * input.std().item() + input.mean().item()
        raise ValueError(f"Unknown strategy {strategy}")

    def _fit(self, key, src, want, strategy):
        if src.dim() != len(want):
            if src.dim() == 1:
                src = src.unsqueeze(1)
            elif len(want) == 1:
                src = src[:, 0]
            else:
                raise ValueError(f"Shapes of {key} are different: {tuple(want)} != {tuple(src.shape)}")
            assert src.dim() == len(want), f"Shapes of {key} are different: {tuple(want)} != {tuple(src.shape)}"
        if tuple(src.shape) == tuple(want):
            return src
        out = src.clone()
        for axis, (have_n, want_n) in enumerate(zip(src.shape, want)):
            if want_n > have_n:
                grow = list(out.shape)
                grow[axis] = want_n - have_n
                out = torch.cat((out, self._create_block(shape=grow, strategy=strategy, input=out)), dim=axis)
                logging.info(f"Adapting {key} with strategy:{strategy} from shape {tuple(src.shape)} to {tuple(want)}")
            elif want_n < have_n:
                out = out.narrow(axis, 0, want_n)
                logging.info(f"Adapting {key} by narrowing from shape {tuple(src.shape)} to {tuple(want)}")
        return out

    def __call__(self, model_state_dict: Dict[str, torch.Tensor], checkpoint_state_dict: Dict[str, torch.Tensor],
                 regex_keys: Optional[List[str]] = None, strategy: Literal["zeros", "normal"] = "normal"):
        t0 = time.perf_counter()
        patterns = list(model_state_dict.keys()) if regex_keys is None else regex_keys
        for key in list(checkpoint_state_dict.keys()):
            for pattern in patterns:            # a key matched by several patterns is visited once per pattern, as upstream
                if re.match(pattern, key):
                    checkpoint_state_dict[key] = self._fit(key, checkpoint_state_dict[key],
                                                           model_state_dict[key].shape, strategy)
        logging.info(f"StateDictAdapter took {time.perf_counter() - t0:.2f} seconds")
        return checkpoint_state_dict


class StateDictRenamer:
    """Moves entries of `checkpoint_state_dict` to new names (`rename_dict`: old -> new); a missing old name is only
    warned about, an already present new name is an error."""

    def __call__(self, checkpoint_state_dict: Dict[str, torch.Tensor], rename_dict: Dict[str, str]):
        for old, new in rename_dict.items():
            if old not in checkpoint_state_dict:
                logging.warning(f"Key {old} not found in checkpoint state dict")
                continue
            assert new not in checkpoint_state_dict, f"Key {new} already exists in checkpoint state dict"
            checkpoint_state_dict[new] = checkpoint_state_dict.pop(old)
            logging.info(f"Renaming {old} to {new}")
        return checkpoint_state_dict
