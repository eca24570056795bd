"""`flash.config.BaseConfig` against a record of the REFERENCE's own class (tests/golden/reference_config.pt, written by
tests/golden/make_reference_config_golden.py from the unmodified src/flash/config.py:13-141): dict / json / yaml round
trips byte for byte, the WARNING (not an error) when another class's file is loaded, and the exception types for a
missing file, malformed json / yaml, a file without the `name` key and an invalid field."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = torch.load(os.path.join(HERE, "golden", "reference_config.pt"), weights_only=False)


def test_base_config_matches_reference_run():
    import make_reference_config_golden as G
    from flash.config import BaseConfig
    got = G.run(BaseConfig)
    for k in ("to_dict", "json", "json_file", "yaml_file", "from_json", "from_yaml", "mismatch_json", "mismatch_yaml",
              "errors", "from_dict"):
        assert got[k] == GOLD[k], (k, got[k], GOLD[k])
    assert got["errors"] == {"missing_file": "FileNotFoundError", "bad_json": "TypeError", "no_name_key": "KeyError",
                             "bad_field": "ValidationError", "bad_yaml": "YAMLError"}
    assert len(got["mismatch_json"]["warnings"]) == 1
