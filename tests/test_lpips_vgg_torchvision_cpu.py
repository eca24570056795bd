"""The VGG16 feature stack of the LPIPS oracle (oracle/lpips.py) pinned to torchvision's own `vgg16` — the module the
lpips package wraps (`lpips/pretrained_networks.py`: `tv.vgg16(...).features` cut into the slices [0:4], [4:9], [9:16],
[16:23], [23:30] = relu1_2, relu2_2, relu3_3, relu4_3, relu5_3).  torchvision IS installed in this image, so this part of
the upstream arithmetic is checked against the real thing (random weights, loaded into both through the shared key
numbering); the lpips-specific parts (scaling layer, unit normalisation, lin layers) stay restated from the paper."""
import pytest
import torch


def test_oracle_vgg_taps_equal_torchvision_vgg16_features():
    tv = pytest.importorskip("torchvision")
    from oracle.lpips import CHNS, LPIPSOracle
    torch.manual_seed(0)
    feats = tv.models.vgg16(weights=None).features.eval()
    ora = LPIPSOracle()
    # the oracle keeps torchvision's layer numbering inside each slice: net.slice{k}.{N}.weight <-> features.{N}.weight
    sd = {}
    for k in range(5):
        for name, p in getattr(ora.net, f"slice{k + 1}").named_parameters():
            sd[name] = p.detach().clone()
    assert set(sd) == set(feats.state_dict())
    feats.load_state_dict(sd)
    x = torch.randn(2, 3, 64, 48)
    with torch.no_grad():
        taps = ora.net(x)
        h, ref, cuts = x, [], [4, 9, 16, 23, 30]
        for i, layer in enumerate(feats):
            if i >= cuts[-1]:
                break
            h = layer(h)
            if i + 1 in cuts:
                ref.append(h)
    assert [t.shape[1] for t in taps] == CHNS == [r.shape[1] for r in ref]
    for k, (a, b) in enumerate(zip(taps, ref)):
        assert a.shape == b.shape and torch.allclose(a, b, rtol=1e-5, atol=1e-6), k
