"""Host-side packing rules of flash.b200.ops.LinearPack that need no GPU: the zero padding of input widths that are not
multiples of 8 bf16 elements (TMA row pitch), of attention heads to the kernels' 16-channel granularity, and that the
padding is invisible to autograd."""
import torch
import torch.nn as nn


def test_k_padding_of_odd_input_widths():
    from flash.b200.ops import LinearPack
    lin_k, lin_v = nn.Linear(123, 32, bias=False), nn.Linear(123, 32, bias=False)
    pack = LinearPack([lin_k, lin_v])
    assert pack.K == 123 and pack.k_pad == 128
    w = pack._w2d(lin_k)
    assert w.shape == (32, 128) and torch.equal(w[:, :123], lin_k.weight.detach()) and float(w[:, 123:].abs().sum()) == 0
    x = torch.randn(5, 123, requires_grad=True)
    xp = pack.pad_k(x)
    assert xp.shape == (5, 128) and float(xp.detach()[:, 123:].abs().sum()) == 0 and pack.pad_k(xp) is xp
    # y = x W^T is unchanged by the padding, and the gradient of the padded input maps back onto the 123 real columns
    y = xp @ torch.cat([pack._w2d(lin_k), pack._w2d(lin_v)]).t()
    ref = torch.cat([lin_k(x), lin_v(x)], dim=1)
    assert torch.allclose(y, ref, atol=1e-6)
    y.sum().backward()
    assert x.grad.shape == (5, 123)
    # widths that already are multiples of 8 are left alone (no copy on the hot path)
    even = LinearPack(nn.Linear(2048, 64))
    z = torch.randn(3, 2048)
    assert even.k_pad == 0 and even.pad_k(z) is z and even._w2d(even.bases[0]).shape == (64, 2048)


def test_head_padding_rows_and_columns():
    from flash.b200.ops import LinearPack
    H, d, dp, C = 3, 8, 16, 24                      # UNet2DModel-style 8-channel heads -> 16
    q = nn.Linear(C, H * d)
    out = nn.Linear(H * d, C)
    pq = LinearPack(q, head_pad=(H, d, dp))
    po = LinearPack(out, head_pad=(H, d, dp), pad_cols=True)
    wq, bq, wo = pq._w2d(q), pq._bias(q), po._w2d(out)
    assert wq.shape == (H * dp, C) and bq.shape == (H * dp,) and wo.shape == (C, H * dp) and po.k_pad == 0
    x = torch.randn(4, C)
    full = (x @ wq.t() + bq).reshape(4, H, dp)
    assert torch.allclose(full[:, :, :d].reshape(4, H * d), q(x), atol=1e-6) and float(full[:, :, d:].abs().sum()) == 0
    o = torch.randn(4, H, dp)
    o[:, :, d:] = 123.0                              # whatever sits in the padded channels is multiplied by zero columns
    assert torch.allclose(o.reshape(4, H * dp) @ wo.t(), o[:, :, :d].reshape(4, H * d) @ out.weight.t(), atol=1e-5)
