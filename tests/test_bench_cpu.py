"""bench.py pieces that need no GPU: the timed start-index order keeps the mean teacher-rollout length at the expected
5K/8 after every even number of steps, and the analytic step FLOPs follow SURVEY 8d."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def test_start_index_order_and_step_flops():
    import bench
    for name in ("sdxl", "sd15", "pixart", "sd3"):
        cfg = bench._cfg(name)
        d = bench.config_dict(name, cfg, 8)
        K, modes = cfg["K"], d["start_idx_schedule"]
        assert sorted(modes) == [0, K // 4, K // 2, 3 * K // 4] and d["expected_teacher_steps"] == 5 * K // 8
        for n in (2, 4, 6, 10, 20):
            assert sum(K - modes[i % 4] for i in range(n)) / n == 5 * K / 8, (name, n)
        assert d["global_batch"] == 8 * cfg["B"] and d["parallelism"] == "dp8"
    # FLOP/img = 2[(4 + 2n) F + 2 F_dm] + F + F_dm + 2 F_lora_dW   (SURVEY 8d), n = 20 -> ~617 TFLOP
    assert 600e12 < bench.flops_per_image(20) < 640e12
    assert bench.flops_per_image(32) > bench.flops_per_image(24) > bench.flops_per_image(8)
