"""reference: src/flash/models/utils.py:316-361 (extract_into_tensor, append_dims)."""
import torch


def extract_into_tensor(a: torch.Tensor, t: torch.Tensor, x_shape):
    """Gather a[t] and reshape to broadcast against a tensor of shape x_shape."""
    b = t.shape[0]
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


def append_dims(x: torch.Tensor, target_dims: int) -> torch.Tensor:
    extra = target_dims - x.ndim
    if extra < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * extra]
