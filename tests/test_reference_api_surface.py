"""Drop-in check of the API SURFACE: every public class, method and config field of the reference's `src/flash` modules
that its examples / tests import must exist in this repository's `flash` package with the same parameter names, order
and defaults (extra OPTIONAL parameters — `draws=`, `generator=`, `kv_cache=`, `noise=` — are B200-side extensions).
Both packages are introspected in fresh interpreters (tools/api_surface.py); the reference tree is only in the build
container, so the test is skipped elsewhere."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# documented deviations (DESIGN.md §9 / INTEGRATION.md)
ALLOWED_MISSING_METHODS = {
    # the T2I-adapter recipe (out of scope): the names exist and explain themselves when constructed
    "DiffusersT2IAdapterWrapper.forward", "DiffusersT2IAdapterWrapper.freeze",
    "CannyEdgeMapper.__call__", "MidasDepthMapper.__call__",                # controlnet_aux detectors of that recipe
}
ALLOWED_PARAM_DIFFS = {
    # rank_zero_only-wrapped callback hook: (*a, **k) forwarding wrapper
    "WandbSampleLogger.log_samples": {"self", "trainer", "pl_module", "outputs", "batch", "batch_idx", "split"},
}
# the reference's wrappers subclass diffusers models and take (*args, **kwargs); here the diffusers keyword arguments are
# spelled out
SPELLED_OUT_CTORS = {"DiffusersUNet2DCondWrapper.__init__", "DiffusersUNet2DWrapper.__init__", "DiffusersTransformer2DWrapper.__init__",
                     "DiffusersSD3Transformer2DWrapper.__init__"}


def _surface(which):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "api_surface.py"), which], capture_output=True,
                       text=True, timeout=900, cwd="/tmp")
    assert p.returncode == 0, p.stderr[-2000:]
    return json.loads(p.stdout)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/flash"), reason="the reference tree is not on this box")
def test_api_surface_covers_the_reference():
    ref, prod = _surface("ref"), _surface("prod")
    assert not ref["import_errors"] and not prod["import_errors"], (ref["import_errors"], prod["import_errors"])
    # config dataclasses: same fields, same defaults
    for name, fields in ref["configs"].items():
        assert name in prod["configs"], f"config class {name} is missing"
        assert prod["configs"][name] == fields, (name, {k: (fields.get(k), prod["configs"][name].get(k))
                                                        for k in set(fields) | set(prod["configs"][name])
                                                        if fields.get(k) != prod["configs"][name].get(k)})
    # methods: present, reference parameters kept (names, defaults, relative order)
    problems = []
    for key, rp in ref["methods"].items():
        if key not in prod["methods"]:
            if key not in ALLOWED_MISSING_METHODS:
                problems.append(f"missing: {key}")
            continue
        pp = prod["methods"][key]
        rnames, pnames = [x[0] for x in rp], [x[0] for x in pp]
        allowed = ALLOWED_PARAM_DIFFS.get(key, set())
        for name, kind, default in rp:
            if key in SPELLED_OUT_CTORS and name in ("args", "kwargs"):
                continue
            if name not in pnames:
                if name not in allowed:
                    problems.append(f"{key}: parameter `{name}` is missing")
                continue
            pd = pp[pnames.index(name)][2]
            if default != "-" and pd != default:
                problems.append(f"{key}: default of `{name}` is {pd}, reference {default}")
        common_r = [n for n in rnames if n in pnames]
        common_p = [n for n in pnames if n in rnames]
        if common_r != common_p:
            problems.append(f"{key}: parameter order differs: {common_r} vs {common_p}")
        for name, kind, default in pp:                       # extensions must be optional
            if name not in rnames and default == "-" and "VAR_" not in kind and key not in SPELLED_OUT_CTORS:
                problems.append(f"{key}: extra REQUIRED parameter `{name}`")
    assert not problems, "\n".join(problems)
    assert len(ref["methods"]) > 150 and len(ref["configs"]) >= 10
