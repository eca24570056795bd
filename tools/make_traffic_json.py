"""profiles/r02_traffic.json from `ncu --set full` captures of the in-situ GEMMs (run here, on the CPU box):

  python tools/make_traffic_json.py <rep>:<case,case,...> [<rep>:<cases> ...]

Each report holds two consecutive launches per case, in the order given (tools/bench_gemm_insitu.py 1 <cases> under
`ncu --set full --clock-control none -k regex:gemm_pair -s <3 * ncases> -c <2 * ncases>`).  Algorithmic bytes = what the
GEMM must move once: A + W (+ residual) read, output written (bf16), per DESIGN.md §4."""
import csv
import json
import subprocess
import sys

SHAPES = {"ff1": (8192, 10240, 1280, "LayerNorm fold + bias + GEGLU", False, True),
          "ff1b": (32768, 5120, 640, "LayerNorm fold + bias + GEGLU", False, True),
          "o": (8192, 1280, 1280, "bias + residual + row statistics", True, False),
          "ff2": (8192, 1280, 5120, "bias + residual + row statistics", True, False),
          "qkv": (8192, 3840, 1280, "LayerNorm fold", False, False),
          "o640": (32768, 640, 640, "bias + residual + row statistics", True, False)}


def launches(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in raw.splitlines() if not l.startswith("==")]))
    hdr, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))

        def val(k):
            v = float(d[k].replace(",", ""))
            return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9}.get(u[k], 1.0)
        out.append({"kernel": d["Kernel Name"][:40], "read": val("dram__bytes_read.sum"), "write": val("dram__bytes_write.sum"),
                    "time_s": val("gpu__time_duration.sum"),
                    "tensor_pct": float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"])})
    return out


def main():
    shapes = {}
    for arg in sys.argv[1:]:
        rep, cases = arg.split(":")
        ls = launches(rep)
        for i, name in enumerate(cases.split(",")):
            M, N, K, epi, res, geglu = SHAPES[name]
            pair = ls[2 * i:2 * i + 2]
            n_out = N // 2 if geglu else N
            alg_read = 2 * (M * K + N * K) + (2 * M * n_out if res else 0)
            alg_write = 2 * M * n_out
            shapes[name] = {"M": M, "N": N, "K": K, "epilogue": epi, "kernel": pair[0]["kernel"],
                            "dram_bytes_read_per_launch": sum(p["read"] for p in pair) / len(pair),
                            "dram_bytes_write_per_launch": sum(p["write"] for p in pair) / len(pair),
                            "algorithmic_bytes_read": alg_read, "algorithmic_bytes_write": alg_write,
                            "ncu_time_us": 1e6 * sum(p["time_s"] for p in pair) / len(pair),
                            "tensor_pipe_pct": sum(p["tensor_pct"] for p in pair) / len(pair), "source": rep}
    dom = shapes.get("ff1") or next(iter(shapes.values()))
    out = {"dominant_shape": "ff1 8192x10240x1280 (LayerNorm fold + bias + GEGLU), 14 % of a step" if "ff1" in shapes else None,
           "dominant_dram_bytes_per_launch": dom["dram_bytes_read_per_launch"] + dom["dram_bytes_write_per_launch"],
           "dominant_algorithmic_bytes_per_launch": dom["algorithmic_bytes_read"] + dom["algorithmic_bytes_write"],
           "note": "dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full --clock-control none` capture, two "
                   "launches averaged; writes below the algorithmic output size = the tail of the output still sits in "
                   "the 126 MB L2 when the kernel ends",
           "shapes": shapes}
    json.dump(out, open("profiles/r02_traffic.json", "w"), indent=1)
    print(json.dumps(out, indent=1)[:1500])


if __name__ == "__main__":
    main()
