set -x
timeout 900 python -m pytest tests/test_vae_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_tests_vae.log
tail -12 gpurun_out/r02_tests_vae.log
for P in 4 3 2 0; do echo "POLY=$P"; FD_ATTN_POLY=$P timeout 300 python tools/bench_attn.py 2>&1 | grep "attn" | head -4; done | tee gpurun_out/r02_bench_attn_poly.txt
timeout 600 python tools/bench_gemm_insitu.py 30 | tee gpurun_out/r02_gemm_insitu.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 18 -c 6 -o gpurun_out/r02_prof_gemm_insitu -f python tools/bench_gemm_insitu.py 1 ff1,ff1b,o > gpurun_out/r02_ncu_gemm.log 2>&1
tail -3 gpurun_out/r02_ncu_gemm.log
cat > /tmp/attn1.py <<'PY'
import sys; sys.path.insert(0, "flash-diffusion_b200")
import torch
from flash.b200 import raw
q = torch.randn(8, 1024, 1280, device="cuda").bfloat16(); k = torch.randn_like(q); v = torch.randn_like(q)
for _ in range(4): raw.attention_fwd(q, k, v, 20)
q = torch.randn(8, 4096, 640, device="cuda").bfloat16(); k = torch.randn_like(q); v = torch.randn_like(q)
for _ in range(2): raw.attention_fwd(q, k, v, 10)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd1 -s 2 -c 3 -o gpurun_out/r02_prof_attn_fwd1 -f python /tmp/attn1.py > gpurun_out/r02_ncu_attn.log 2>&1
tail -3 gpurun_out/r02_ncu_attn.log
ls -la gpurun_out/*.ncu-rep | tail -4
