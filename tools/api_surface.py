"""Dump the public API surface of a `flash` package — every class / function exported by the modules the reference's
examples and tests import, with each public method's parameters (name, kind, default) and each config dataclass's fields
and defaults — as JSON on stdout.

    python tools/api_surface.py ref     # the reference (/root/reference/src, third-party imports stood in for)
    python tools/api_surface.py prod    # this repository's package

tests/test_reference_api_surface.py runs both and checks that nothing the reference offers is missing here."""
import dataclasses
import importlib
import inspect
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODULES = ["flash.models.embedders", "flash.models.flash", "flash.models.flash_sd3", "flash.models.vae", "flash.trainer",
           "flash.models.unets", "flash.models.transformers", "flash.trainer.loggers", "flash.trainer.utils",
           "flash.models.utils", "flash.models.base.base_model", "flash.config", "flash.data.datasets", "flash.data.filters",
           "flash.data.mappers", "flash.models.adapters"]


def show(v):
    return "<callable>" if callable(v) and not inspect.isclass(v) else repr(v)


def params(fn):
    return [[p.name, str(p.kind), show(p.default) if p.default is not p.empty else "-"]
            for p in inspect.signature(fn).parameters.values()]


def main(which):
    if which == "ref":
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        sys.path.insert(0, ROOT)
        import make_reference_step_golden as G
        G.install_shims()
        import types
        for name in ("webdataset", "controlnet_aux", "wandb"):      # imported at module scope by flash.data / loggers
            if name not in sys.modules:
                stub = types.ModuleType(name)
                stub.DataPipeline = stub.CannyDetector = stub.MidasDetector = stub.WebLoader = object
                stub.warn_and_continue = stub.reraise_exception = lambda *a, **k: True
                sys.modules[name] = stub
        if not hasattr(sys.modules["pytorch_lightning"], "LightningDataModule"):
            sys.modules["pytorch_lightning"].LightningDataModule = object
        sys.path.insert(0, G.REF_SRC)
    else:
        sys.path.insert(0, os.path.join(ROOT, "flash-diffusion_b200"))
        sys.path.append(os.path.join(ROOT, "flash-diffusion_b200", "compat"))
    out = {"methods": {}, "configs": {}, "import_errors": {}}
    for m in MODULES:
        try:
            mod = importlib.import_module(m)
        except Exception as e:                                  # noqa: BLE001
            out["import_errors"][m] = repr(e)[:300]
            continue
        for n in sorted(dir(mod)):
            c = getattr(mod, n)
            if n.startswith("_") or not (inspect.isclass(c) or inspect.isfunction(c)):
                continue
            if getattr(c, "__module__", "").split(".")[0] != "flash":
                continue
            if inspect.isfunction(c):
                try:
                    out["methods"][f"{m}.{n}"] = params(c)
                except (TypeError, ValueError):
                    pass
                continue
            if dataclasses.is_dataclass(c):
                out["configs"][n] = {f.name: (show(f.default) if f.default is not dataclasses.MISSING else
                                              ("factory" if f.default_factory is not dataclasses.MISSING else "REQUIRED"))
                                     for f in dataclasses.fields(c)}
            for mn, fn in inspect.getmembers(c, predicate=inspect.isfunction):
                if (mn.startswith("_") and mn not in ("__init__", "__call__")) or fn.__module__.split(".")[0] != "flash":
                    continue
                try:
                    out["methods"][f"{n}.{mn}"] = params(fn)
                except (TypeError, ValueError):
                    pass
    json.dump(out, sys.stdout, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1])
