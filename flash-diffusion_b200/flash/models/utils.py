"""reference: src/flash/models/utils.py — `Tiler` (:12-313: cut a [B, C, H, W] tensor into overlapping tiles, merge the
processed tiles back by averaging, gaussian weighting or linear cross-fading), `extract_into_tensor` (:316-330), `pad`
(:333-349), `append_dims` (:352-360), `update_ema` (:363-377).  Pinned to runs of the reference's own functions:
tests/golden/reference_utils.pt (tests/test_reference_utils_golden.py)."""
import math
from typing import List

import torch
import torch.nn.functional as F

TILING_METHODS = ["average", "gaussian", "linear"]


class Tiler:
    """Same call contract as the reference's: `get_tiles` records the geometry (`output_shape`, `output_tile_size`,
    `output_overlap_size`, all scaled by `scale`) that `merge_tiles` then uses; tiles are rows of columns."""

    def get_tiles(self, input: torch.Tensor, tile_size: tuple, overlap_size: tuple, scale: int = 1,
                  out_channels: int = 3) -> List[List[torch.Tensor]]:
        for ax in (0, 1):
            assert overlap_size[ax] <= tile_size[ax], \
                f"Overlap size {overlap_size} must be smaller than tile size {tile_size}"
        B, C, H, W = input.shape
        th, tw = tile_size
        # an axis that fits into one tile is not tiled: no overlap along it
        oh = overlap_size[0] if H > th else 0
        ow = overlap_size[1] if W > tw else 0
        self.tile_size = tile_size
        self.output_overlap_size = (int(oh * scale), int(ow * scale))
        self.output_tile_size = (int(th * scale), int(tw * scale))
        self.output_shape = (B, out_channels, int(H * scale), int(W * scale))
        return [[input[:, :, i:i + th, j:j + tw].clone() for j in range(0, W, tw - ow)] for i in range(0, H, th - oh)]

    def merge_tiles(self, tiles: List[List[torch.Tensor]], tiling_method: str = "gaussian") -> torch.Tensor:
        if tiling_method == "average":
            return self._weighted_merge(tiles, lambda t: torch.ones((), dtype=t.dtype))
        if tiling_method == "gaussian":
            return self._weighted_merge(tiles, lambda t: self._gaussian_weights(t.shape[3], t.shape[2], 1, 1))
        if tiling_method == "linear":
            return self._linear_merge_tiles(tiles)
        raise ValueError(f"Unknown tiling method {tiling_method}. Available methods are {TILING_METHODS}")

    def _origins(self):
        _, _, H, W = self.output_shape
        sh = self.output_tile_size[0] - self.output_overlap_size[0]
        sw = self.output_tile_size[1] - self.output_overlap_size[1]
        return list(range(0, H, sh)), list(range(0, W, sw))

    def _weighted_merge(self, tiles, weight_of):
        """sum(tile * w) / sum(w) on the CPU in the default dtype, as the reference accumulates it"""
        out = torch.zeros(self.output_shape)
        norm = torch.zeros(self.output_shape)
        th, tw = self.output_tile_size
        rows, cols = self._origins()
        for a, i in enumerate(rows):
            for b, j in enumerate(cols):
                tile = tiles[a][b]
                w = weight_of(tile)
                out[:, :, i:i + th, j:j + tw] += tile.to(out.device) * w
                norm[:, :, i:i + th, j:j + tw] += w
        return out / norm

    def _gaussian_weights(self, tile_width: int, tile_height: int, nbatches: int, channels: int) -> torch.Tensor:
        """Outer product of two gaussians of relative variance 0.01, centred at (width - 1) / 2 and at height / 2 (the
        reference's asymmetry between the two axes is kept), float64, tiled to [nbatches, channels, h, w]."""
        var = 0.01
        norm = math.sqrt(2 * math.pi * var)

        def bell(n, mid):
            return [math.exp(-(k - mid) * (k - mid) / (n * n) / (2 * var)) / norm for k in range(n)]
        wx = torch.tensor(bell(tile_width, (tile_width - 1) / 2), dtype=torch.float64)
        wy = torch.tensor(bell(tile_height, tile_height / 2), dtype=torch.float64)
        return torch.outer(wy, wx).expand(nbatches, channels, tile_height, tile_width).clone()

    @staticmethod
    def _fade(prev_edge, cur_edge, dim):
        """linear cross-fade of the first n lines of `cur_edge` with the last n lines of the previous tile"""
        n = cur_edge.shape[dim]
        shape = [1, 1, 1, 1]
        shape[dim] = n
        t = (torch.arange(n, dtype=cur_edge.dtype, device=cur_edge.device) / n).view(shape)
        return prev_edge * (1 - t) + cur_edge * t

    def _blend_v(self, a: torch.Tensor, b: torch.Tensor, blend_extent: int) -> torch.Tensor:
        n = min(a.shape[2], b.shape[2], blend_extent)
        if n > 0:
            b[:, :, :n, :] = self._fade(a[:, :, a.shape[2] - n:, :], b[:, :, :n, :], 2)
        return b

    def _blend_h(self, a: torch.Tensor, b: torch.Tensor, blend_extent: int) -> torch.Tensor:
        n = min(a.shape[3], b.shape[3], blend_extent)
        if n > 0:
            b[:, :, :, :n] = self._fade(a[:, :, :, a.shape[3] - n:], b[:, :, :, :n], 3)
        return b

    def _linear_merge_tiles(self, tiles: List[List[torch.Tensor]]) -> torch.Tensor:
        """Each tile is cross-faded with its left neighbour (as given) and then with the already blended tile above it,
        and contributes its top-left (tile - overlap) block."""
        oh, ow = self.output_overlap_size
        keep_h, keep_w = self.output_tile_size[0] - oh, self.output_tile_size[1] - ow
        done = [[t.clone() for t in row] for row in tiles]
        rows = []
        for a, row in enumerate(done):
            strip = []
            for b in range(len(row)):
                cur = row[b]
                if b > 0:
                    cur = self._blend_h(row[b - 1], cur, ow)          # the neighbour as already blended
                if a > 0:
                    cur = self._blend_v(done[a - 1][b], cur, oh)
                row[b] = cur
                strip.append(cur[:, :, :keep_h, :keep_w])
            rows.append(torch.cat(strip, dim=3))
        return torch.cat(rows, dim=2)


def extract_into_tensor(a: torch.Tensor, t: torch.Tensor, x_shape):
    """Gather a[t] and reshape to broadcast against a tensor of shape x_shape."""
    b = t.shape[0]
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


def pad(x: torch.Tensor, base_h: int, base_w: int) -> torch.Tensor:
    """zero-pad the last two dims on the right / bottom up to the next multiples of base_h / base_w"""
    h, w = x.shape[-2:]
    dh, dw = -h % base_h, -w % base_w
    return F.pad(x, (0, dw, 0, dh)) if (dh or dw) else x


def append_dims(x: torch.Tensor, target_dims: int) -> torch.Tensor:
    extra = target_dims - x.ndim
    if extra < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * extra]


@torch.no_grad()
def update_ema(target_params: List[torch.Tensor], source_params: List[torch.Tensor], rate: float = 0.99):
    """target <- rate * target + (1 - rate) * source, in place"""
    for tgt, src in zip(target_params, source_params):
        tgt.detach().mul_(rate).add_(src, alpha=1 - rate)
