"""adaLN-single conditioning of the PixArt wrapper (reference src/flash/models/transformers/utils.py:8-102) against
vectors the REFERENCE's own class produced (tests/golden/reference_adaln.pt): the oracle restatement (oracle/dit.py) is
replayed on the CPU with the same seeded weights, and the product's parameter container carries the same keys.  (The
product's CUDA evaluation of it is checked against the oracle in tests/test_dit_gpu.py.)"""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = torch.load(os.path.join(HERE, "golden", "reference_adaln.pt"), weights_only=False)


@pytest.mark.parametrize("name", sorted(GOLD["cases"]))
def test_oracle_adaln_single_matches_reference_run(name):
    import make_reference_adaln_golden as G
    from make_golden import seeded_state_dict
    from flash.models.transformers.transformers import AdaLayerNormSingle as ProductAdaLN
    from oracle.dit import AdaLayerNormSingle
    c = next(x for x in G.CASES if x["name"] == name)
    rec = GOLD["cases"][name]
    net = AdaLayerNormSingle(**G.kwargs(c))
    assert sorted(net.state_dict()) == rec["keys"]
    assert sorted(ProductAdaLN(**G.kwargs(c)).state_dict()) == rec["keys"]
    net.load_state_dict(seeded_state_dict(net, rec["seed"]))
    t, v = G.inputs(c, 50 + [x["name"] for x in G.CASES].index(name))
    with torch.no_grad():
        t6, emb = net(t, v)
    assert torch.allclose(emb, rec["emb"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(t6, rec["t6"], rtol=1e-5, atol=1e-6)
