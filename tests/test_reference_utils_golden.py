"""flash.trainer.utils (StateDictAdapter / StateDictRenamer) and flash.models.utils.append_dims against runs of the
REFERENCE's own helpers (tests/golden/reference_utils.pt, tests/golden/make_reference_utils_golden.py)."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = torch.load(os.path.join(HERE, "golden", "reference_utils.pt"), weights_only=False)


def test_state_dict_helpers_match_reference_run(tmp_path):
    import make_reference_utils_golden as G
    from flash.models.utils import append_dims
    from flash.trainer.utils import StateDictAdapter, StateDictRenamer, setup_logging
    out = G.run(StateDictAdapter, StateDictRenamer)
    for name, want in GOLD["adapter"].items():
        got = out["adapter"][name]
        assert list(got) == list(want), name
        for k in want:
            assert got[k].shape == want[k].shape and torch.equal(got[k], want[k]), (name, k)
    assert list(out["renamer"]) == list(GOLD["renamer"])
    for k, v in GOLD["renamer"].items():
        assert torch.equal(out["renamer"][k], v)
    assert [tuple(append_dims(torch.zeros(2, 3), n).shape) for n in (2, 3, 5)] == GOLD["append_dims"]
    with pytest.raises(ValueError):
        append_dims(torch.zeros(2, 3), 1)
    with pytest.raises(ValueError):
        StateDictAdapter()({"w": torch.zeros(2, 2, 2)}, {"w": torch.zeros(2, 2)}, strategy="zeros")
    logger, path = setup_logging(str(tmp_path), "logs", "run1", logger_name="flash-test-logger")
    logger.info("hello")
    assert path.endswith(os.path.join("logs", "run1.log")) and "hello" in open(path).read()
