"""reference: src/flash/models/adapters/{__init__,t2i_adapter}.py — `DiffusersT2IAdapterWrapper` subclasses diffusers'
`T2IAdapter` and is used by the canny-adapter recipe only (examples/train_flash_canny_adapter.py).  T2I adapters are out
of the B200 hot path (`down_intrablock_additional_residuals` is refused by the UNet engine): the
name exists so that `flash.models.adapters` imports, constructing it explains why it is not built."""


class DiffusersT2IAdapterWrapper:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("T2I adapters (canny-adapter recipe) are outside the B200 hot path: the UNet engine "
                                  "does not take down_intrablock_additional_residuals (SURVEY.md §2 row 7)")


__all__ = ["DiffusersT2IAdapterWrapper"]
