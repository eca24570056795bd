"""Golden vectors from the REFERENCE's own `Tiler` (src/flash/models/utils.py:12-313: get_tiles + the three merge methods),
`pad` (:333-349), `update_ema` (:363-377) and `extract_into_tensor` (:316-330), imported unmodified from
/root/reference/src:     python tests/golden/make_reference_tiler_golden.py  ->  tests/golden/reference_tiler.pt

Every case cuts a seeded [B, C, H, W] tensor into tiles, applies a fixed non-pointwise per-tile map that also changes the
resolution by `scale` (so overlapping tiles genuinely disagree in the overlap), and merges.  Only summaries are stored
(tile shapes, a 4x4 block-mean image and two random projections of the merged output)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)

CASES = [dict(shape=(2, 4, 20, 28), tile=(8, 12), overlap=(2, 4), scale=2, out_channels=3),
         dict(shape=(1, 4, 16, 16), tile=(16, 16), overlap=(4, 4), scale=2, out_channels=3),      # fits: one tile
         dict(shape=(1, 3, 30, 10), tile=(12, 16), overlap=(3, 2), scale=1, out_channels=3),      # tiled along H only
         dict(shape=(2, 4, 24, 24), tile=(8, 8), overlap=(0, 0), scale=4, out_channels=2)]
METHODS = ["average", "gaussian", "linear"]


def process(tile, scale, out_channels, full):
    """stand-in for a decoder: channel mix + a 3x3 box blur (non-pointwise), nearest up-sampling by `scale`; partial
    trailing tiles are zero-padded to the tile size first (what the reference's VAE wrapper does, vae/autoencoderKL.py
    :97-104) decoded at the full tile size and cropped to the tile's own extent"""
    th, tw = full
    t = torch.zeros(tile.shape[0], tile.shape[1], th, tw, dtype=tile.dtype)
    t[:, :, :tile.shape[2], :tile.shape[3]] = tile
    mix = torch.linspace(-1, 1, out_channels * t.shape[1]).reshape(out_channels, t.shape[1])
    y = torch.einsum("oc,bchw->bohw", mix, t)
    y = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(y, (1, 1, 1, 1), mode="replicate"), 3, stride=1)
    y = torch.nn.functional.interpolate(y, scale_factor=scale, mode="nearest")
    return y[:, :, :tile.shape[2] * scale, :tile.shape[3] * scale].clone()       # ... and cropped back (:105-113)


def summarise(x):
    g = torch.Generator().manual_seed(99)
    p = torch.randn(2, x.numel(), generator=g, dtype=torch.float64)
    return dict(shape=tuple(x.shape), blocks=torch.nn.functional.adaptive_avg_pool2d(x.double(), 4).clone(),
                proj=(p @ x.double().reshape(-1)).clone())


def run(tiler_cls, pad_fn, ema_fn, extract_fn):
    out = {"cases": []}
    for ci, c in enumerate(CASES):
        g = torch.Generator().manual_seed(10 + ci)
        x = torch.randn(*c["shape"], generator=g)
        rec = dict(case=c, merged={})
        for m in METHODS:
            tiler = tiler_cls()
            tiles = tiler.get_tiles(x.clone(), c["tile"], c["overlap"], scale=c["scale"], out_channels=c["out_channels"])
            rec["tile_shapes"] = [[tuple(t.shape) for t in row] for row in tiles]
            rec["geometry"] = dict(output_shape=tuple(tiler.output_shape), output_tile_size=tuple(tiler.output_tile_size),
                                   output_overlap_size=tuple(tiler.output_overlap_size))
            done = [[process(t, c["scale"], c["out_channels"], c["tile"]) for t in row] for row in tiles]
            merged = tiler.merge_tiles(done, tiling_method=m)
            rec["merged"][m] = summarise(merged)
        out["cases"].append(rec)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 13, 22, generator=g)
    out["pad"] = [dict(base=b, shape=tuple(pad_fn(x, *b).shape), sum=float(pad_fn(x, *b).double().sum()),
                       corner=pad_fn(x, *b)[0, 0, -1, -1].item()) for b in ((8, 8), (13, 11), (1, 1), (16, 5))]
    tgt = [torch.randn(4, 3, generator=g), torch.randn(5, generator=g)]
    src = [torch.randn(4, 3, generator=g), torch.randn(5, generator=g)]
    out["ema_in"] = dict(target=[t.clone() for t in tgt], source=[s.clone() for s in src])
    ema_fn(tgt, src, rate=0.9)
    out["ema_out"] = [t.clone() for t in tgt]
    a = torch.linspace(0, 1, 50)
    out["extract"] = extract_fn(a, torch.tensor([3, 49, 0]), (3, 4, 8, 8)).clone()
    return out


def main():
    import make_reference_step_golden as G
    G.install_shims()
    sys.path.insert(0, G.REF_SRC)
    from flash.models.utils import Tiler, extract_into_tensor, pad, update_ema
    import flash
    assert os.path.realpath(flash.__path__[0]).startswith(G.REF_SRC)
    out = run(Tiler, pad, update_ema, extract_into_tensor)
    out["generated_by"] = os.path.relpath(__file__, ROOT)
    out["reference_files"] = ["src/flash/models/utils.py:12-377"]
    path = os.path.join(HERE, "reference_tiler.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")
    for rec in out["cases"]:
        print(rec["case"]["shape"], rec["geometry"], {m: v["shape"] for m, v in rec["merged"].items()})


if __name__ == "__main__":
    main()
