"""The REFERENCE'S OWN test files (gojasper/flash-diffusion `tests/`), unmodified, executed against this repository's
`flash` package (+ the compat shims for the packages that cannot be installed offline).  They are the reference's
statement of its wrapper / embedder / data API: shapes, conditioning keys, update invariants.

The files are read from /root/reference/tests (build container) or from $FLASH_REF_TESTS (a one-off GPU run ships them
inside the command, tools/reference_tests_cmd.py — the reference tree does not travel and must not be copied into the
repo).  Each group runs in a fresh interpreter whose sys.path ends with the compat directory.

CPU groups: data filters / mappers, timestep + torch.nn embedders (and, with FLASH_REF_TESTS_SLOW=1, the CLIP
conditioner tests and the T5 embedder's config + t5-v1_1-base cases, which build real-size text encoders: 42 + 3 tests,
15 minutes on the build container).  GPU groups (the denoisers / VAE are CUDA-only): denoiser
wrappers, VAE, FlashDiffusion forward.  Known, documented deviation: `test_flash_diffusion.py::test_optimizers*` train
ALL student parameters — the B200 path differentiates the LoRA adapters and the inputs only (BASELINE north_star: "the
student LoRA backward"), so those two are not collected."""
# (the T5 embedder file's T5-XXL cases need a 4.7 B-parameter encoder and are left to FLASH_REF_TESTS_SLOW runs)
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.environ.get("FLASH_REF_TESTS", "/root/reference/tests")
PKG = os.path.join(ROOT, "flash-diffusion_b200")

CPU_GROUPS = {        # one interpreter start (tens of seconds of imports on a slow box) for all four files
    "data_and_embedders": ["test_dataset/test_filters.py", "test_dataset/test_mappers.py",
                           "test_embedders/test_time_embedders.py", "test_embedders/test_torchnn_embedder.py"],
}
SLOW_CPU_GROUPS = {
    "clip_conditioners": ["test_embedders/test_conditioners_wrapper.py", "test_embedders/test_clip_embedders.py"],
    # config validation + the google/t5-v1_1-base cases (the other parameter set builds the 4.7 B-parameter T5-XXL)
    "t5_base": ["test_embedders/test_t5_embedder.py", "-k", "wrong_config or shape0"],
}
GPU_GROUPS = {
    "unet_wrapper": ["test_unet/test_unets_wrappers.py::TestDiffusersUNet2DCondWrapper"],
    "unet2d_wrapper": ["test_unet/test_unets_wrappers.py::TestDiffusersUNet2DWrapper"],
    "transformer_wrapper": ["test_transformers/test_transformers_wrappers.py"],
    "vae": ["test_vaes/test_autoencoderKL.py"],
    "flash_forward": ["test_flash/test_flash_diffusion.py::TestTurbo::test_model_forward"],
    "clip_conditioners": SLOW_CPU_GROUPS["clip_conditioners"],
}

RUNNER = """
import sys
sys.path.insert(0, {pkg!r})
sys.path.append({compat!r})          # AFTER site-packages: a real diffusers / peft / lightning install would win
import pytest
sys.exit(pytest.main({args!r}))
"""


def run_reference_tests(files, timeout=1500):
    args = [os.path.join(REF_TESTS, f) if f.startswith("test_") else f for f in files] + ["-q", "-p", "no:cacheprovider", "--rootdir", REF_TESTS,
                                                          "--tb=short"]
    code = RUNNER.format(pkg=PKG, compat=os.path.join(PKG, "compat"), args=args)
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    p = subprocess.run([sys.executable, "-c", code], cwd="/tmp", env=env, capture_output=True, text=True, timeout=timeout)
    tail = "\n".join((p.stdout + "\n" + p.stderr).strip().splitlines()[-80:])
    return p.returncode, tail


needs_ref = pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="the reference tests are not on this box")


@needs_ref
@pytest.mark.parametrize("group", sorted(CPU_GROUPS))
def test_reference_cpu_tests_pass_unchanged(group):
    rc, tail = run_reference_tests(CPU_GROUPS[group])
    assert rc == 0, tail


@needs_ref
@pytest.mark.skipif(os.environ.get("FLASH_REF_TESTS_SLOW") != "1", reason="builds real-size CLIP encoders on the CPU")
@pytest.mark.parametrize("group", sorted(SLOW_CPU_GROUPS))
def test_reference_slow_cpu_tests_pass_unchanged(group):
    rc, tail = run_reference_tests(SLOW_CPU_GROUPS[group], timeout=3000)
    assert rc == 0, tail


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("group", sorted(GPU_GROUPS))
def test_reference_gpu_tests_pass_unchanged(group):
    assert torch.cuda.is_available()
    rc, tail = run_reference_tests(GPU_GROUPS[group], timeout=900)
    print(tail)
    assert rc == 0, tail
