from .export import ModelCheckpoint, load_lora, merge_lora_into_base, save_lora
from .trainer import TrainingPipeline
from .training_config import TrainingConfig

__all__ = ["TrainingPipeline", "TrainingConfig", "ModelCheckpoint", "save_lora", "load_lora", "merge_lora_into_base"]
