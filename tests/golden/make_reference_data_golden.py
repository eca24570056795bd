"""Golden vectors from the REFERENCE's own sample mappers / filters (src/flash/data/mappers/mappers.py:24-252,
mappers_wrapper.py, src/flash/data/filters/filters.py:9-70, filter_wrapper.py), imported unmodified from
/root/reference/src in the build container:   python tests/golden/make_reference_data_golden.py
-> tests/golden/reference_data.pt.  tests/test_reference_data_golden.py replays the same samples through flash.data.

The reference modules import webdataset / controlnet_aux / pytorch_lightning at module scope; none is installed, so
they are replaced by empty stand-ins (nothing of them is executed by the classes exercised here)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)

COND = {"eq1": lambda x: x == 1, "gt5": lambda x: x > 5, "is_square": lambda x: x[0] == x[1]}


def image(seed, h=40, w=56):
    from PIL import Image
    return Image.fromarray(np.random.RandomState(seed).randint(0, 256, (h, w, 3), dtype=np.uint8))


def samples():
    import json
    meta = {"caption": "a photo", "aesthetic": 6.5, "size": [40, 56], "nested": {"k": 3}}
    return [
        {"jpg": image(0), "json": json.dumps(meta), "flag": 1, "score": 7},
        {"jpg": image(1, 64, 64), "json": {**meta, "aesthetic": 2.0, "size": [64, 64]}, "flag": 0,
         "score": 3},
        {"png": image(2), "flag": 1},
        {"t": torch.linspace(0, 1, 12).view(3, 2, 2)},
    ]


# (class name, config class name, config kwargs [condition functions by name], which samples)
MAPPERS = [
    ("KeyRenameMapper", "KeyRenameMapperConfig", dict(key_map={"jpg": "image"}), [0, 1, 2]),
    ("KeyRenameMapper", "KeyRenameMapperConfig",
     dict(key_map={"jpg": "image"}, condition_key="flag", condition_fn="eq1", else_key_map={"jpg": "other"}), [0, 1]),
    ("TorchvisionMapper", "TorchvisionMapperConfig",
     dict(key="jpg", output_key="image", transforms=["CenterCrop", "Resize", "ToTensor"],
          transforms_kwargs=[{"size": 32}, {"size": 16, "antialias": True}, {}]), [0, 1]),
    ("KeysFromJSONMapper", "KeysFromJSONMapperConfig", dict(key="json", keys_to_extract=["caption", "aesthetic"]), [0, 1]),
    ("KeysFromJSONMapper", "KeysFromJSONMapperConfig",
     dict(key="json", keys_to_extract=["caption", "missing"], remove_original=False, strict=False), [0, 1]),
    ("KeysFromJSONMapper", "KeysFromJSONMapperConfig", dict(key="json", keys_to_extract="size"), [1]),
    ("SelectKeysMapper", "SelectKeysMapperConfig", dict(keys=["flag", "score"]), [0, 1]),
    ("RemoveKeysMapper", "RemoveKeysMapperConfig", dict(keys=["json", "flag"]), [0, 1]),
    ("SetValueMapper", "SetValueConfig", dict(key="score", value=0), [0, 1]),
    ("SetValueMapper", "SetValueConfig", dict(key="text", value=""), [0]),
    ("RescaleMapper", "RescaleMapperConfig", dict(key="t", output_key="t2"), [3]),
]
FILTERS = [
    ("KeyFilter", "KeyFilterConfig", dict(keys=["jpg", "json"]), [0, 1, 2]),
    ("KeyFilter", "KeyFilterConfig", dict(keys="png"), [0, 2]),
    ("FilterOnCondition", "FilterOnConditionConfig", dict(condition_key="score", condition_fn="gt5"), [0, 1]),
    ("FilterOnCondition", "FilterOnConditionConfig", dict(condition_key="score", condition_fn="gt5", strict=True), [2]),
    ("FilterOnCondition", "FilterOnConditionConfig", dict(condition_key="score", condition_fn="gt5", strict=False), [2]),
]


def resolve(kw):
    return {k: (COND[v] if k == "condition_fn" else v) for k, v in kw.items()}


def plain(v):
    """PIL images -> uint8 arrays so that the fixture holds only tensors / python scalars"""
    if hasattr(v, "size") and hasattr(v, "mode"):
        return {"__pil__": torch.from_numpy(np.array(v))}
    return v


def run(mod_mappers, mod_filters, wrapper_cls=None, filter_wrapper_cls=None):
    out = {"mappers": [], "filters": []}
    for cls, cfg, kw, idx in MAPPERS:
        m = getattr(mod_mappers, cls)(getattr(mod_mappers, cfg)(**resolve(kw)))
        out["mappers"].append([{k: plain(v) for k, v in m(dict(samples()[i])).items()} for i in idx])
    for cls, cfg, kw, idx in FILTERS:
        f = getattr(mod_filters, cls)(getattr(mod_filters, cfg)(**resolve(kw)))
        out["filters"].append([bool(f(dict(samples()[i]))) for i in idx])
    if wrapper_cls is not None:
        chain = [getattr(mod_mappers, c)(getattr(mod_mappers, g)(**resolve(kw))) for c, g, kw, _ in
                 (MAPPERS[0], MAPPERS[3], MAPPERS[8])]
        w = wrapper_cls(chain)
        out["wrapper"] = [{k: plain(v) for k, v in w(dict(samples()[i])).items()} for i in (0, 1)]
        fw = filter_wrapper_cls([getattr(mod_filters, c)(getattr(mod_filters, g)(**resolve(kw))) for c, g, kw, _ in
                                 (FILTERS[0], FILTERS[2])])
        out["filter_wrapper"] = [bool(fw(dict(samples()[i]))) for i in (0, 1, 2)]
    return out


def main():
    import make_reference_step_golden as G
    G.install_shims()
    for name in ("webdataset", "controlnet_aux", "wandb"):
        m = types.ModuleType(name)
        m.DataPipeline = m.CannyDetector = m.MidasDetector = object
        m.warn_and_continue = m.reraise_exception = lambda *a, **k: True
        sys.modules[name] = m
    pl = sys.modules["pytorch_lightning"]
    pl.LightningDataModule = object
    sys.path.insert(0, G.REF_SRC)
    from flash.data import filters as RF
    from flash.data import mappers as RM
    import flash
    assert os.path.realpath(flash.__path__[0]).startswith(G.REF_SRC)
    out = run(RM, RF, RM.MapperWrapper, RF.FilterWrapper)
    out["generated_by"] = os.path.relpath(__file__, ROOT)
    path = os.path.join(HERE, "reference_data.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(out["mappers"]), "mapper cases,", len(out["filters"]),
          "filter cases")


if __name__ == "__main__":
    main()
