"""Golden record of the REFERENCE's own `custom_collation_fn` (src/flash/data/datasets/collation_fn.py:7-41), imported
unmodified:   python tests/golden/make_reference_collation_golden.py  ->  tests/golden/reference_collation.pt"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_FILE = "/root/reference/src/flash/data/datasets/collation_fn.py"


def samples():
    g = torch.Generator().manual_seed(2)
    out = []
    for i in range(3):
        s = {"image": torch.randn(3, 4, 5, generator=g), "score": 0.5 * i + 0.25, "idx": i, "flag": i % 2 == 0,
             "text": f"caption {i}", "arr": np.arange(4, dtype=np.float32) * (i + 1), "meta": {"k": i},
             "size": torch.tensor([512 + i, 640])}
        if i != 1:
            s["only_some"] = i                      # not common to all samples -> dropped
        out.append(s)
    return out


def plain(v):
    if isinstance(v, np.ndarray):
        return ("ndarray", str(v.dtype), tuple(v.shape), torch.from_numpy(np.ascontiguousarray(v)).clone())
    if isinstance(v, torch.Tensor):
        return ("tensor", str(v.dtype), tuple(v.shape), v.clone())
    return ("py", type(v).__name__, v)


def run(fn):
    out = {}
    for name, kw in (("default", {}), ("no_tensors", dict(combine_tensors=False)), ("no_scalars", dict(combine_scalars=False))):
        res = fn(samples(), **kw)
        out[name] = {k: plain(res[k]) for k in sorted(res)}
    return out


def main():
    spec = importlib.util.spec_from_file_location("ref_collation_fn", REF_FILE)     # the file has no package-level imports
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = run(mod.custom_collation_fn)
    out["generated_by"] = os.path.relpath(__file__, ROOT)
    path = os.path.join(HERE, "reference_collation.pt")
    torch.save(out, path)
    print("wrote", path, {k: list(v) for k, v in out.items() if isinstance(v, dict)})


if __name__ == "__main__":
    main()
