"""Golden vectors from the REFERENCE's own `AdaLayerNormSingle` (src/flash/models/transformers/utils.py:8-102: the adaLN-
single conditioning of the PixArt wrapper — timestep embedding, one or several vector embedders whose outputs are added
/ concatenated, SiLU, the 6x linear), imported unmodified from /root/reference/src:
    python tests/golden/make_reference_adaln_golden.py  ->  tests/golden/reference_adaln.pt

The two diffusers classes it builds on (`Timesteps`, `TimestepEmbedding`: a sinusoid and a two-layer MLP, UPSTREAM) are
served by the oracle's restatements; what the fixture pins is the reference's own glue around them."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)

CASES = [dict(name="concat3", time_embed_dim=96, timesteps_embedding_num_channels=32, projection_class_embeddings_input_dim=8,
              use_concat_conditioning=True, num_vector_conditionings=3, vec=24),
         dict(name="single", time_embed_dim=64, timesteps_embedding_num_channels=16, projection_class_embeddings_input_dim=12,
              use_concat_conditioning=False, num_vector_conditionings=None, vec=12),
         dict(name="none", time_embed_dim=32, timesteps_embedding_num_channels=8, projection_class_embeddings_input_dim=None,
              use_concat_conditioning=False, num_vector_conditionings=None, vec=0)]


def kwargs(c):
    return {k: c[k] for k in ("time_embed_dim", "timesteps_embedding_num_channels", "projection_class_embeddings_input_dim",
                              "use_concat_conditioning", "num_vector_conditionings")}


def inputs(c, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.randint(0, 1000, (3,), generator=g).float()
    v = torch.randn(3, c["vec"], generator=g) if c["vec"] else None
    return t, v


def main():
    import make_reference_step_golden as G
    from make_golden import seeded_state_dict
    G.install_shims()
    sys.path.insert(0, ROOT)
    from oracle.unet import TimestepEmbedding as OTE
    from oracle.unet import timestep_embedding

    class Timesteps(torch.nn.Module):
        def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift):
            super().__init__()
            self.n, self.flip, self.shift = num_channels, flip_sin_to_cos, downscale_freq_shift

        def forward(self, t):
            return timestep_embedding(t, self.n, self.flip, self.shift)

    class TimestepEmbedding(OTE):
        def __init__(self, in_channels, time_embed_dim):
            super().__init__(in_channels, time_embed_dim)
    emb = sys.modules["diffusers.models.embeddings"]
    emb.Timesteps, emb.TimestepEmbedding = Timesteps, TimestepEmbedding
    sys.path.insert(0, G.REF_SRC)
    from flash.models.transformers.utils import AdaLayerNormSingle
    import flash
    assert os.path.realpath(flash.__path__[0]).startswith(G.REF_SRC)
    out = {"cases": {}, "generated_by": os.path.relpath(__file__, ROOT),
           "reference_files": ["src/flash/models/transformers/utils.py:8-102"]}
    for ci, c in enumerate(CASES):
        net = AdaLayerNormSingle(**kwargs(c))
        net.load_state_dict(seeded_state_dict(net, 300 + ci))
        t, v = inputs(c, 50 + ci)
        with torch.no_grad():
            t6, e = net(t, {"vector_conditioning": v})
        out["cases"][c["name"]] = dict(keys=sorted(net.state_dict()), t6=t6.clone(), emb=e.clone(), seed=300 + ci)
        print(c["name"], tuple(t6.shape), tuple(e.shape), len(net.state_dict()))
    torch.save(out, os.path.join(HERE, "reference_adaln.pt"))


if __name__ == "__main__":
    main()
