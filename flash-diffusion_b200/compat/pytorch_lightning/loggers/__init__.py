from flash.trainer.lightning import WandbLogger  # noqa: F401
