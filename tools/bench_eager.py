"""Same-box "kernel to beat" (BASELINE.md §3): what torch 2.11 eager dispatches on THIS B200 for the hot
path's ops — cuBLASLt GEMMs, cuDNN convs, SDPA (flash / cuDNN attention) under bf16 autocast — beside our kernels.

    python tools/bench_eager.py [--skip-unet]

Prints one JSON object: attention (ours vs F.scaled_dot_product_attention), GEMM shapes (fd_gemm vs torch.matmul bf16),
the SDXL oracle UNet under torch.autocast(bfloat16): one 2B teacher evaluation (B=8, no grad) and one LoRA student
forward+backward (B=4), vs the B200 engine on the same inputs."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-diffusion_b200"))
import torch
import torch.nn.functional as F

from flash.b200 import raw


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def attention_rows():
    rows = []
    for (B, H, Nq, Nkv) in [(8, 20, 1024, 1024), (8, 10, 4096, 4096), (8, 20, 1024, 77), (8, 10, 4096, 77)]:
        q = torch.randn(B, Nq, H * 64, device="cuda").bfloat16()
        k = torch.randn(B, Nkv, H * 64, device="cuda").bfloat16()
        v = torch.randn(B, Nkv, H * 64, device="cuda").bfloat16()
        fl = 4.0 * B * H * Nq * Nkv * 64
        ms = timeit(lambda: raw.attention_fwd(q, k, v, H), n=20)
        # the layout diffusers hands to SDPA: [B, H, N, d] views of the projection outputs
        q4, k4, v4 = (t.view(B, -1, H, 64).transpose(1, 2) for t in (q, k, v))
        ms_t = timeit(lambda: F.scaled_dot_product_attention(q4, k4, v4), n=20)
        o, lse = raw.attention_fwd(q, k, v, H, need_lse=True)
        do = torch.randn_like(o)
        ms_b = timeit(lambda: raw.attention_bwd(q, k, v, o, lse, do, H), n=5)
        qg, kg, vg = (t.detach().clone().requires_grad_(True) for t in (q4, k4, v4))
        og = F.scaled_dot_product_attention(qg, kg, vg)
        dog = torch.randn_like(og)
        ms_tb = timeit(lambda: torch.autograd.grad(og, (qg, kg, vg), dog, retain_graph=True), n=5)
        rows.append({"B": B, "H": H, "Nq": Nq, "Nkv": Nkv, "fd_fwd_tflops": fl / ms / 1e9, "sdpa_fwd_tflops": fl / ms_t / 1e9,
                     "fd_bwd_tflops": 2.5 * fl / ms_b / 1e9, "sdpa_bwd_tflops": 2.5 * fl / ms_tb / 1e9})
    return rows


def gemm_rows():
    rows = []
    for (M, N, K) in [(8192, 1280, 1280), (8192, 3840, 1280), (8192, 10240, 1280), (8192, 1280, 5120), (32768, 640, 640),
                      (32768, 1920, 640), (32768, 5120, 640), (32768, 640, 2560), (616, 2560, 2048)]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        b = torch.randn(N, K, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: raw.gemm(a, b, out=out), n=20)
        ms_t = timeit(lambda: torch.matmul(a, b.t(), out=out), n=20)
        fl = 2.0 * M * N * K
        rows.append({"M": M, "N": N, "K": K, "fd_tflops": fl / ms / 1e9, "cublas_tflops": fl / ms_t / 1e9})
    return rows


def unet_rows():
    import copy
    from flash.models.lora import LoraConfig
    from flash.models.unets import DiffusersUNet2DCondWrapper
    from oracle.unet import LoraConfig as OLoraConfig
    from oracle.unet import SDXL_KWARGS, UNet2DConditionOracle
    torch.manual_seed(0)
    lora = dict(r=64, lora_alpha=64, init_lora_weights="gaussian", target_modules=["to_k", "to_q", "to_v", "to_out.0"])
    with torch.device("cuda"):
        ora = UNet2DConditionOracle(**SDXL_KWARGS)
    for p in ora.parameters():
        p.requires_grad = False
    with torch.device("meta"):
        prod = DiffusersUNet2DCondWrapper(**SDXL_KWARGS)
    prod = prod.to_empty(device="cuda")
    prod.load_state_dict(ora.state_dict())
    prod.freeze()
    res = {}
    B = 8
    x = torch.randn(B, 4, 128, 128, device="cuda")
    t = torch.full((B,), 500.0, device="cuda")
    cond = {"cond": {"crossattn": torch.randn(B, 77, 2048, device="cuda"), "vector": torch.randn(B, 2816, device="cuda")}}

    def eager_fwd():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return ora(x, t, cond)

    from flash.b200.graphs import GraphedDenoiser
    g = GraphedDenoiser(prod)

    def ours_fwd():
        with torch.no_grad():
            return g(x, t, cond, clone=False)

    res["teacher_eval_2B_ms"] = {"torch_eager_autocast_bf16": timeit(eager_fwd, n=5, warm=2), "fd_b200": timeit(ours_fwd, n=5, warm=2),
                                 "batch": B, "flops": 6.76e12 * B}
    # LoRA student forward + backward at B = 4
    ora_s = copy.deepcopy(ora)
    ora_s.add_adapter(OLoraConfig(**lora))
    ora_s = ora_s.cuda()
    prod_s = copy.deepcopy(prod)
    prod_s.add_adapter(LoraConfig(**lora))
    prod_s.load_state_dict(ora_s.state_dict())
    prod_s.train()
    B = 4
    xs, ts = x[:B].clone(), t[:B].clone()
    cs = {"cond": {k: v[:B].clone() for k, v in cond["cond"].items()}}

    def eager_fb():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = ora_s(xs, ts, cs)
        y.float().square().mean().backward()

    def ours_fb():
        prod_s(xs, ts, cs).square().mean().backward()

    res["student_fwd_bwd_ms"] = {"torch_eager_autocast_bf16": timeit(eager_fb, n=3, warm=2), "fd_b200": timeit(ours_fb, n=3, warm=2),
                                 "batch": B}
    return res


if __name__ == "__main__":
    out = {"attention": attention_rows(), "gemm": gemm_rows()}
    if "--skip-unet" not in sys.argv:
        out["sdxl_unet"] = unet_rows()
    print(json.dumps(out))
