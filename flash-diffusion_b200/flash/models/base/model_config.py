from pydantic.dataclasses import dataclass

from ...config import BaseConfig


@dataclass
class ModelConfig(BaseConfig):
    """reference: src/flash/models/base/model_config.py:6-8"""
    input_key: str = "image"
