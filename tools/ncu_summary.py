"""Summarise ncu outputs into small text files for profiles/ (run here, on the CPU box).

  python tools/ncu_summary.py launches <launches.csv> <out.txt>      per-kernel share of device time
  python tools/ncu_summary.py report <file.ncu-rep> <out.txt>        key raw metrics of each captured launch
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic"]


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[name][0] += 1
        agg[name][1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({path})\n")
        f.write(f"# {n} launches captured, {tot / 1e6:.2f} ms total device time (cold-cache, serialised: compare SHARES)\n")
        f.write(f"{'share':>8} {'launches':>9} {'avg_us':>10}  kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{100 * v[1] / tot:7.2f}% {v[0]:9d} {v[1] / v[0] / 1e3:10.1f}  {k[:110]}\n")
    print(open(out).read()[:3000])


def report(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in raw.splitlines() if not l.startswith("==")]))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on  ({path})\n")
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            f.write(f"\n== {d.get('Kernel Name', '?')[:120]}  grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
            for i, h in enumerate(hdr):
                if h in KEYS:
                    f.write(f"   {h} [{units[i]}] = {row[i]}\n")
    print(open(out).read()[:4000])


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
