"""Golden vectors from the REFERENCE's own checkpoint-surgery helpers (src/flash/trainer/utils.py: StateDictAdapter
:41-180, StateDictRenamer :183-222) and tensor helpers (src/flash/models/utils.py: Tiler / pad are covered by
make_reference_vae_golden.py; append_dims :352-359), imported unmodified from /root/reference/src:
    python tests/golden/make_reference_utils_golden.py  ->  tests/golden/reference_utils.pt"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)


def dicts():
    g = torch.Generator().manual_seed(4)
    r = lambda *s: torch.randn(*s, generator=g)
    model = {"conv_in.weight": r(8, 9, 3, 3), "class_embedding.linear_1.weight": r(16, 40), "a.bias": r(16),
             "attn.to_k.weight": r(12, 20), "scale": r(6), "col": r(6, 1), "same": r(3, 3)}
    ckpt = {"conv_in.weight": r(8, 4, 3, 3), "class_embedding.linear_1.weight": r(16, 24), "a.bias": r(10),
            "attn.to_k.weight": r(12, 32), "scale": r(6, 1), "col": r(6), "same": r(3, 3), "untouched": r(2, 2)}
    return model, ckpt


CASES = [dict(name="zeros_all", regex_keys=None, strategy="zeros"),
         dict(name="normal_subset", regex_keys=[r"conv_in\.weight", r"class_embedding\..*", r"attn\.to_(k|v)\.weight"],
              strategy="normal"),
         dict(name="zeros_rank", regex_keys=[r"scale", r"col", r"a\.bias"], strategy="zeros")]
RENAME = {"add_embedding.linear_1.bias": "class_embedding.linear_1.bias", "missing.key": "whatever",
          "add_embedding.linear_1.weight": "class_embedding.linear_1.weight"}


def run(adapter_cls, renamer_cls):
    out = {"adapter": {}, "renamer": None}
    for case in CASES:
        model, ckpt = dicts()
        torch.manual_seed(31)                  # the "normal" strategy draws from the global generator
        res = adapter_cls()(model_state_dict=model, checkpoint_state_dict=ckpt, regex_keys=case["regex_keys"],
                            strategy=case["strategy"])
        out["adapter"][case["name"]] = {k: v.clone() for k, v in res.items()}
    sd = {"add_embedding.linear_1.bias": torch.arange(3.0), "add_embedding.linear_1.weight": torch.ones(2, 2), "x": torch.zeros(1)}
    res = renamer_cls()(checkpoint_state_dict=sd, rename_dict=RENAME)
    out["renamer"] = {k: v.clone() for k, v in res.items()}
    return out


def main():
    import make_reference_step_golden as G
    G.install_shims()
    sys.path.insert(0, G.REF_SRC)
    from flash.models.utils import append_dims
    from flash.trainer.utils import StateDictAdapter, StateDictRenamer
    import flash
    assert os.path.realpath(flash.__path__[0]).startswith(G.REF_SRC)
    out = run(StateDictAdapter, StateDictRenamer)
    out["append_dims"] = [tuple(append_dims(torch.zeros(2, 3), n).shape) for n in (2, 3, 5)]
    out["generated_by"] = os.path.relpath(__file__, ROOT)
    path = os.path.join(HERE, "reference_utils.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", {k: list(v) for k, v in out["adapter"].items()}.keys())


if __name__ == "__main__":
    main()
