"""Shim of `peft.LoraConfig` / `peft.get_peft_model` (examples/train_flash_pixart.py:15,237-256) over flash.models.lora."""
from flash.models.lora import LoraConfig, inject_lora  # noqa: F401


def get_peft_model(model, peft_config):
    """peft wraps the model in a `PeftModel` that forwards keyword calls; the B200 wrappers carry the adapter
    themselves, so the adapted model is returned as is (with peft's `print_trainable_parameters`)."""
    if hasattr(model, "add_adapter"):
        model.add_adapter(peft_config)
    else:
        inject_lora(model, peft_config)

    def print_trainable_parameters():
        t = sum(p.numel() for p in model.parameters() if p.requires_grad)
        a = sum(p.numel() for p in model.parameters())
        print(f"trainable params: {t:,} || all params: {a:,} || trainable%: {100 * t / a:.4f}")

    model.print_trainable_parameters = print_trainable_parameters
    return model
