from flash.trainer.export import ModelCheckpoint  # noqa: F401
from flash.trainer.lightning import Callback  # noqa: F401
