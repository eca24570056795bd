"""Prints the gpurun command that runs the REFERENCE'S OWN GPU-side test files, unmodified, on a B200 against this
repository (tests/test_reference_suite_unchanged.py).  /root/reference does not exist on the GPU box and its files must
not be copied into the repo, so the test directory travels inside the command (tar.gz, base64) and lands in /tmp:
    gpurun --timeout 1200 -- "$(python tools/reference_tests_cmd.py)"              # all GPU groups
    gpurun --timeout 1200 -- "$(python tools/reference_tests_cmd.py wrapper)"      # -k wrapper """
import base64
import io
import sys
import tarfile

buf = io.BytesIO()
with tarfile.open(fileobj=buf, mode="w:gz") as tf:
    tf.add("/root/reference/tests", arcname="tests")
print("mkdir -p /tmp/ref gpurun_out && echo %s | base64 -d | tar xzf - -C /tmp/ref && "
      "(FLASH_REF_TESTS=/tmp/ref/tests timeout 1000 python -m pytest tests/test_reference_suite_unchanged.py -m gpu -q -s %s "
      "2>&1 | grep -v Warning | tail -150 | tee gpurun_out/r02_reference_tests_gpu.txt)" % (base64.b64encode(buf.getvalue()).decode(), ("-k " + sys.argv[1]) if len(sys.argv) > 1 else ""))
