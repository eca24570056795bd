"""Prints the gpurun command that executes the UNMODIFIED reference script examples/train_flash_sd.py for 2 steps on a
B200 against this repository (tests/test_examples_run_unchanged.py).  /root/reference does not exist on the GPU box and
its sources must not be copied into the repo, so the script and its yaml travel INSIDE the command (base64) and land in
/tmp on the box:   gpurun --timeout 1500 -- "$(python tools/example_run_cmd.py)" """
import base64

b = lambda p: base64.b64encode(open(p, "rb").read()).decode()
print("mkdir -p /tmp/ref && echo %s | base64 -d > /tmp/ref/train_flash_sd.py && echo %s | base64 -d > /tmp/ref/flash_sd.yaml && "
      "FLASH_REF_SCRIPT=/tmp/ref/train_flash_sd.py FLASH_REF_YAML=/tmp/ref/flash_sd.yaml timeout 1200 python -m pytest "
      "tests/test_examples_run_unchanged.py -x -q -s 2>&1 | tail -40 | tee gpurun_out/r02_example_train_flash_sd_gpu.txt"
      % (b("/root/reference/examples/train_flash_sd.py"), b("/root/reference/examples/configs/flash_sd.yaml")))
