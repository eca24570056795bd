"""Prints the gpurun command that executes UNMODIFIED reference example scripts for 2 training steps on a B200 against
this repository (tests/test_examples_run_unchanged.py).  /root/reference does not exist on the GPU box and its sources
must not be copied into the repo, so the scripts and their yamls travel INSIDE the command (base64) and land in /tmp on
the box:
    gpurun --timeout 1500 -- "$(python tools/example_run_cmd.py sd)"            # examples/train_flash_sd.py
    gpurun --timeout 1500 -- "$(python tools/example_run_cmd.py sdxl sd3)"      # the full-size scripts
Each script's pytest tail is written to gpurun_out/r02_example_train_flash_<name>_gpu.txt."""
import base64
import gzip
import sys

b = lambda p: base64.b64encode(gzip.compress(open(p, "rb").read(), 9)).decode()
names = sys.argv[1:] or ["sd"]
cmd = ["mkdir -p /tmp/ref/configs gpurun_out"]
for n in names:
    cmd.append("echo %s | base64 -d | gunzip > /tmp/ref/train_flash_%s.py" % (b(f"/root/reference/examples/train_flash_{n}.py"), n))
    cmd.append("echo %s | base64 -d | gunzip > /tmp/ref/configs/flash_%s.yaml" % (b(f"/root/reference/examples/configs/flash_{n}.yaml"), n))
for n in names:
    if n == "sd":
        env, sel = "FLASH_REF_SCRIPT=/tmp/ref/train_flash_sd.py FLASH_REF_YAML=/tmp/ref/configs/flash_sd.yaml", "test_train_flash_sd_runs_unchanged"
    else:
        env, sel = "FLASH_REF_DIR=/tmp/ref", f"test_full_size_scripts_run_unchanged[{n}]"
    cmd.append(f"({env} timeout 600 python -m pytest 'tests/test_examples_run_unchanged.py::{sel}' -x -q -s 2>&1 | tail -40 "
               f"| tee gpurun_out/r02_example_train_flash_{n}_gpu.txt)")
print(" && ".join(cmd[: 1 + 2 * len(names)]) + " ; " + " ; ".join(cmd[1 + 2 * len(names):]))
