"""Checker-side diagnostic (uses oracle/, like the tests; never imported by the product or bench.py): first op of the Transformer2D engine that yields a non-finite value (vector-conditioned variant)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "flash-diffusion_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from flash.b200 import ops, raw
import test_dit_gpu as T

def wrap(mod, name):
    fn = getattr(mod, name)
    def w(*a, **k):
        out = fn(*a, **k)
        outs = out if isinstance(out, (tuple, list)) else [out]
        for o in outs:
            if torch.is_tensor(o) and o.is_floating_point():
                bad = not bool(torch.isfinite(o.float()).all())
                ins = [float(x.float().abs().max()) for x in a if torch.is_tensor(x) and x.is_floating_point()]
                print(f"{mod.__name__.split('.')[-1]}.{name}: out shape {tuple(o.shape)} absmax {float(o.float().abs().nan_to_num(float('inf')).max()):.3e} finite={not bad} | in absmax {['%.2e' % v for v in ins]}", flush=True)
                if bad and not state["bad"]:
                    state["bad"] = True
                    print("   ^^^ FIRST NON-FINITE OUTPUT", flush=True)
        return out
    setattr(mod, name, w)
state = {"bad": False}
for n in ["linear", "modulate", "geglu", "gated_linear", "attention_self", "attention_cross", "unpatchify"]:
    wrap(ops, n)
for n in ["timestep_embedding", "silu_f32_to_bf16", "cast_scale"]:
    wrap(raw, n)

for scale_vec in (1.0, 1.0 / 64):
    state["bad"] = False
    print(f"===== vector scale {scale_vec}", flush=True)
    kw = dict(T.REF_TEST, in_channels=6, cross_attention_dim=None, projection_class_embeddings_input_dim=12, double_self_attention=True)
    prod, ora = T._pair(kw, seed=21)
    prod.freeze(); ora.freeze()
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.rand(2, 6, 32, 32, device="cuda", generator=g)
    t = torch.randint(0, 1000, (2,), device="cuda", generator=g).float()
    cond = {"vector": torch.randint(0, 256, (2, 12), device="cuda", generator=g).float() * scale_vec}
    with torch.no_grad():
        ref = ora(x, t, {"cond": cond})
        out = prod(x, t, {"cond": cond})
    print("ref absmax", float(ref.abs().max()), "out finite", bool(torch.isfinite(out).all()),
          "rel", float((out - ref).norm() / ref.norm()), flush=True)
