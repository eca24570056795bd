"""GPU parity of the unconditional `DiffusersUNet2DWrapper` (UNet2DModel on the conv / GroupNorm / GEMM / attention
kernels) against the fp32 oracle oracle/unet2d.py.  Tolerance: bf16 kernels vs fp32 oracle rel-L2 <= 2e-2."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SMALL = dict(in_channels=4, out_channels=3, block_out_channels=(32, 64, 96), layers_per_block=1, norm_num_groups=8,
             down_block_types=("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D"),
             up_block_types=("AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"), attention_head_dim=8, num_class_embeds=10)


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def _pair(kw, seed=0):
    from flash.models.unets import DiffusersUNet2DWrapper
    from oracle.unet2d import UNet2DOracle
    torch.manual_seed(seed)
    ora = UNet2DOracle(**kw)
    with torch.no_grad():
        for n, p in ora.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    with torch.device("meta"):
        prod = DiffusersUNet2DWrapper(**kw)
    prod = prod.to_empty(device="cuda")
    ora = ora.cuda()
    prod.load_state_dict(ora.state_dict())
    prod.freeze(); ora.freeze()
    return prod, ora


@pytest.fixture(autouse=True)
def _fp32_reference():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


@pytest.mark.parametrize("kw,hw", [(SMALL, 16), (dict(SMALL, attention_head_dim=16, num_class_embeds=None), 32),
                                    (dict(SMALL, attention_head_dim=None, num_class_embeds=None), 16)])
def test_small_unconditional_unet_forward(kw, hw):
    prod, ora = _pair(kw)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(2, 4, hw, hw, device="cuda", generator=g)
    t = torch.randint(0, 1000, (2,), device="cuda", generator=g).float()
    cond = {"cond": {"vector": torch.tensor([3, 7], device="cuda")}} if kw.get("num_class_embeds") else None
    with torch.no_grad():
        ref = ora(x, t, cond)
        out = prod(x, t, cond)
        assert out.shape == ref.shape == (2, 3, hw, hw) and out.dtype == torch.float32
        assert _rel(out, ref) < 2e-2, _rel(out, ref)
        assert _rel(prod(x, 10.0, cond), ora(x, 10.0, cond)) < 2e-2       # scalar / int timestep forms of the reference test
        assert _rel(prod(x, 3, cond), ora(x, 3, cond)) < 2e-2


def test_default_unet2d_model_as_in_the_reference_test():
    """diffusers' default UNet2DModel (224-448-672-896 channels, 8-channel attention heads, 7 / 14 / 21 / 28 channels per
    GroupNorm group) with the inputs of tests/test_unet/test_unets_wrappers.py:29-42 (6 input channels incl. a
    concatenated map, 256 classes)."""
    prod, ora = _pair(dict(sample_size=(32, 32), in_channels=6, out_channels=3, num_class_embeds=256), seed=3)
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.rand(2, 4, 32, 32, device="cuda", generator=g)
    cond = {"cond": {"vector": torch.randint(0, 256, (2,), device="cuda", generator=g),
                     "concat": torch.randn(2, 2, 32, 32, device="cuda", generator=g)}}
    t = torch.randint(0, 1000, (2,), device="cuda", generator=g).float()
    with torch.no_grad():
        ref, out = ora(x, t, cond), prod(x, t, cond)
    assert out.shape == ref.shape == (2, 3, 32, 32)
    assert _rel(out, ref) < 2e-2, _rel(out, ref)
