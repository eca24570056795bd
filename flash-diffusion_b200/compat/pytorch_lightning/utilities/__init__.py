from flash.trainer.lightning import rank_zero_only  # noqa: F401
