"""reference: src/flash/trainer/training_config.py:10-136 (same fields / defaults / length checks)."""
from dataclasses import field
from typing import List, Literal, Optional, Union

from pydantic.dataclasses import dataclass

from ..config import BaseConfig


@dataclass
class TrainingConfig(BaseConfig):
    experiment_id: Optional[str] = None
    optimizers_name: List[Literal["Adam", "AdamW", "Adadelta", "Adagrad", "RMSprop", "SGD"]] = field(
        default_factory=lambda: ["AdamW"])
    optimizers_kwargs: Optional[List[dict]] = field(default_factory=lambda: [{}])
    learning_rates: List[float] = field(default_factory=lambda: [1e-3])
    lr_schedulers_name: Optional[List[Literal["StepLR", "CosineAnnealingLR", "CosineAnnealingWarmRestarts",
                                              "ReduceLROnPlateau", "ExponentialLR", None]]] = field(
        default_factory=lambda: [None])
    lr_schedulers_kwargs: Optional[List[dict]] = field(default_factory=lambda: [{}])
    lr_schedulers_interval: Optional[List[Literal["step", "epoch", None]]] = field(default_factory=lambda: ["step"])
    lr_schedulers_frequency: Optional[List[Union[int, None]]] = field(default_factory=lambda: [1])
    metrics: Optional[List[str]] = None
    tracking_metrics: Optional[List[str]] = None
    backup_every: int = 50
    trainable_params: Optional[List[List[str]]] = field(default_factory=lambda: [["./*"]])
    log_keys: Optional[Union[str, List[str]]] = "txt"
    log_samples_model_kwargs: Optional[dict] = field(default_factory=lambda: {
        "max_samples": 8, "num_steps": 20, "input_shape": (4, 32, 32), "guidance_scale": 7.5})

    def __post_init__(self):
        n = len(self.optimizers_name)

        def _same_len(name, values):
            assert n == len(values), (f"The length of optimizers_name ({n}) must be equal to the length of "
                                      f"{name} ({len(values)})")

        if self.optimizers_kwargs != [{}]:
            _same_len("optimizers_kwargs", self.optimizers_kwargs)
        else:
            self.optimizers_kwargs = [{} for _ in range(n)]
        if self.trainable_params != [[".*"]]:
            _same_len("trainable_params", self.trainable_params)
        else:
            self.trainable_params = [[".*"] for _ in range(n)]
        m = len(self.lr_schedulers_name)
        if self.lr_schedulers_kwargs != [{}]:
            assert m == len(self.lr_schedulers_kwargs), "lr_schedulers_name / lr_schedulers_kwargs length mismatch"
            if self.lr_schedulers_frequency != [1]:
                assert m == len(self.lr_schedulers_frequency), "lr_schedulers_frequency length mismatch"
            else:
                self.lr_schedulers_frequency = [1] * m
            if self.lr_schedulers_interval != ["step"]:
                assert m == len(self.lr_schedulers_interval), "lr_schedulers_interval length mismatch"
            else:
                self.lr_schedulers_interval = ["step"] * m
        else:
            self.lr_schedulers_kwargs = [{} for _ in range(m)]
        _same_len("learning_rates", self.learning_rates)
        super().__post_init__()
