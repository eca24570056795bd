"""flash.data mappers / filters against outputs of the REFERENCE's own classes (tests/golden/reference_data.pt, written by
tests/golden/make_reference_data_golden.py from the unmodified src/flash/data/{mappers,filters}): the same samples and
configurations go through the product's classes (SURVEY.md §8 "next": the data formats on the input side of the step)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = torch.load(os.path.join(HERE, "golden", "reference_data.pt"), weights_only=False)


def _same(a, b, path=""):
    assert type(a) is type(b) or (isinstance(a, (int, float)) and isinstance(b, (int, float))), (path, type(a), type(b))
    if isinstance(a, dict):
        assert list(a) == list(b), (path, list(a), list(b))           # key ORDER too: collation depends on it
        for k in a:
            _same(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}[{i}]")
    elif torch.is_tensor(a):
        assert a.shape == b.shape and a.dtype == b.dtype and torch.equal(a, b), path
    else:
        assert a == b, (path, a, b)


def test_product_mappers_and_filters_match_reference_run():
    import make_reference_data_golden as G
    from flash.data import filters as PF
    from flash.data import mappers as PM
    out = G.run(PM, PF, PM.MapperWrapper, PF.FilterWrapper)
    for i, (got, want) in enumerate(zip(out["mappers"], GOLD["mappers"])):
        _same(got, want, f"mapper[{i}:{G.MAPPERS[i][0]}]")
    assert out["filters"] == GOLD["filters"]
    _same(out["wrapper"], GOLD["wrapper"], "wrapper")
    assert out["filter_wrapper"] == GOLD["filter_wrapper"]


def test_product_collation_matches_reference_run():
    """`custom_collation_fn` (reference data/datasets/collation_fn.py:7-41): common keys only; scalars (bools included)
    and numpy arrays become numpy arrays, tensors are stacked, everything else stays a list; a column whose combine flag
    is off is DROPPED.  Record of the reference's own function: tests/golden/reference_collation.pt."""
    import make_reference_collation_golden as GC
    from flash.data.datasets import custom_collation_fn
    gold = torch.load(os.path.join(HERE, "golden", "reference_collation.pt"), weights_only=False)
    got = GC.run(custom_collation_fn)
    for case in ("default", "no_tensors", "no_scalars"):
        assert list(got[case]) == list(gold[case]), (case, list(got[case]), list(gold[case]))
        for k, (a, b) in ((k, (got[case][k], gold[case][k])) for k in gold[case]):
            assert a[:-1] == b[:-1], (case, k, a[:-1], b[:-1])
            if isinstance(b[-1], torch.Tensor):
                assert torch.equal(a[-1], b[-1]), (case, k)
            else:
                assert a[-1] == b[-1], (case, k)
