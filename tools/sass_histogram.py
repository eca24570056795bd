"""Per-kernel histogram of the Blackwell-specific SASS mnemonics in libflashb200.so (run on the CPU box):
    python tools/sass_histogram.py > profiles/r02_sass_mnemonics.txt
UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG = TMA tensor load / store,
UTMAPF = tensormap prefetch, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, UCGABAR = cluster barrier, UBLKCP = bulk copy,
REDG = global reductions (fp32 atomics), MUFU.EX2 = ex2.approx, FFMA2 / FADD2 = packed fp32 pairs."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "flash-diffusion_b200", "lib", "libflashb200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "UCGABAR", "UBLKCP",
        "REDG", "MUFU.EX2", "FFMA2", "FADD2", "HMMA", "LDGSTS"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k in KEYS:
        if re.search(r"\b" + re.escape(k) + r"(\b|_)", line):
            counts[cur][k] += 1
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a) — occurrences per kernel; only kernels using a listed mnemonic")
print("# " + sys.modules[__name__].__doc__.split("\n", 3)[3].replace("\n", "\n# "))
for name, c in counts.items():
    if sum(c.values()):
        print(f"{name}\n    " + "  ".join(f"{k}={c[k]}" for k in KEYS if c[k]))
