"""Unconditional `DiffusersUNet2DWrapper` (reference src/flash/models/unets/unet.py:7-52 around diffusers' UNet2DModel):
module tree / state-dict keys / parameter count against the oracle restatement, and the no-CPU-fallback rule."""
import pytest
import torch

SMALL = dict(in_channels=4, out_channels=3, block_out_channels=(32, 64, 96), layers_per_block=1, norm_num_groups=8,
             down_block_types=("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D"),
             up_block_types=("AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"), attention_head_dim=8, num_class_embeds=10)


def test_keys_shapes_and_default_size_match_the_oracle():
    from flash.models.unets import DiffusersUNet2DWrapper
    from oracle.unet2d import UNet2DOracle
    for kw in (SMALL, dict(sample_size=(32, 32), in_channels=6, out_channels=3, num_class_embeds=256),
               dict(SMALL, attention_head_dim=None, add_attention=False, num_class_embeds=None)):
        with torch.device("meta"):
            prod, ora = DiffusersUNet2DWrapper(**kw), UNet2DOracle(**kw)
        ps, os_ = prod.state_dict(), ora.state_dict()
        assert set(ps) == set(os_)
        assert all(ps[k].shape == os_[k].shape for k in ps)
    # diffusers' default UNet2DModel (224-448-672-896 channels) with 6 input channels and 256 classes
    assert sum(v.numel() for v in ps.values()) > 0
    with torch.device("meta"):
        n = sum(p.numel() for p in DiffusersUNet2DWrapper(in_channels=3, out_channels=3).parameters())
    assert n == sum(p.numel() for p in UNet2DOracle(in_channels=3, out_channels=3).parameters())
    assert "down_blocks.1.attentions.0.group_norm.weight" in ps and "down_blocks.1.attentions.0.to_q.bias" in ps


def test_oracle_forward_contract():
    from oracle.unet2d import UNet2DOracle
    torch.manual_seed(0)
    ora = UNet2DOracle(**SMALL).eval()
    x = torch.rand(2, 4, 16, 16)
    cond = {"cond": {"vector": torch.tensor([3, 7])}}
    with torch.no_grad():
        y = ora(x, 10.0, cond)
        assert y.shape == (2, 3, 16, 16) and torch.isfinite(y).all()
        assert torch.allclose(y, ora(x, torch.tensor([10.0, 10.0]), cond), atol=1e-6)
        assert not torch.allclose(y, ora(x, 10.0, {"cond": {"vector": torch.tensor([4, 7])}}))   # class conditioning acts
        with pytest.raises(ValueError):
            ora(x, 3, None)


def test_product_refuses_cpu_and_unbuilt_variants():
    from flash.models.unets import DiffusersUNet2DWrapper
    net = DiffusersUNet2DWrapper(**dict(SMALL, num_class_embeds=None))
    with pytest.raises(RuntimeError):
        net(torch.rand(1, 4, 16, 16), 3)
    with pytest.raises(NotImplementedError):
        DiffusersUNet2DWrapper(time_embedding_type="fourier")
    with pytest.raises(NotImplementedError):
        DiffusersUNet2DWrapper(down_block_types=("SkipDownBlock2D",), up_block_types=("UpBlock2D",), block_out_channels=(32,))
    net.freeze()
    assert not any(p.requires_grad for p in net.parameters())
