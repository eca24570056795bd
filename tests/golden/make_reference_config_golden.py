"""Golden record of the REFERENCE's own `BaseConfig` (src/flash/config.py:13-141), imported unmodified (the file depends
on pydantic / yaml only):   python tests/golden/make_reference_config_golden.py  ->  tests/golden/reference_config.pt"""
import importlib.util
import os
import sys
import tempfile
import warnings
from typing import List, Optional

import torch
from pydantic.dataclasses import dataclass

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_FILE = "/root/reference/src/flash/config.py"


def run(base_cls):
    @dataclass
    class DemoConfig(base_cls):
        K: List[int] = None
        rate: float = 0.5
        tag: Optional[str] = None

    @dataclass
    class OtherConfig(base_cls):
        K: List[int] = None
        rate: float = 0.25
        tag: Optional[str] = None

    out = {}
    c = DemoConfig(K=[4, 8], rate=0.75, tag="x")
    out["to_dict"] = c.to_dict()
    out["json"] = c.to_json_string()
    with tempfile.TemporaryDirectory() as d:
        jp, yp = os.path.join(d, "c.json"), os.path.join(d, "c.yaml")
        c.save_json(jp)
        c.save_yaml(yp)
        out["json_file"], out["yaml_file"] = open(jp).read(), open(yp).read()
        out["from_json"] = DemoConfig.from_json(jp).to_dict()
        out["from_yaml"] = DemoConfig.from_yaml(yp).to_dict()
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            other = OtherConfig.from_json(jp)
            out["mismatch_json"] = dict(result=other.to_dict(), warnings=[str(x.message) for x in w])
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            other = OtherConfig.from_yaml(yp)
            out["mismatch_yaml"] = dict(result=other.to_dict(), warnings=[str(x.message) for x in w])
        errs = {}
        for name, fn in (("missing_file", lambda: DemoConfig.from_json(os.path.join(d, "nope.json"))),
                         ("bad_json", lambda: (open(jp, "w").write("{not json"), DemoConfig.from_json(jp))),
                         ("no_name_key", lambda: (open(jp, "w").write('{"K": [1]}'), DemoConfig.from_json(jp))),
                         ("bad_field", lambda: DemoConfig.from_dict({"rate": "fast"})),
                         ("bad_yaml", lambda: (open(yp, "w").write("a: [1, 2"), DemoConfig.from_yaml(yp)))):
            try:
                fn()
                errs[name] = None
            except Exception as e:                                   # noqa: BLE001
                errs[name] = type(e).__name__
        out["errors"] = errs
    out["from_dict"] = DemoConfig.from_dict({"K": [1], "tag": None}).to_dict()
    return out


def main():
    spec = importlib.util.spec_from_file_location("ref_flash_config", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = run(mod.BaseConfig)
    out["generated_by"] = os.path.relpath(__file__, ROOT)
    torch.save(out, os.path.join(HERE, "reference_config.pt"))
    for k, v in out.items():
        print(k, "->", str(v)[:150])


if __name__ == "__main__":
    main()
