from .trainer import TrainingPipeline
from .training_config import TrainingConfig

__all__ = ["TrainingPipeline", "TrainingConfig"]
