"""Host logic of bench.py that needs no GPU: exactly K steps are timed for any K."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def test_launch_set_plan_times_exactly_k_steps():
    for n_ctx in (1, 2, 4, 8):
        for groups in (1, 4, 8, 16):
            for steps in list(range(1, 70)) + [100, 128, 255, 256, 257]:
                G, G_rem, full, assign = bench.plan_launch_sets(steps, groups, n_ctx)
                assert 1 <= G <= max(1, min(groups, steps)) and 0 <= G_rem < G
                assert full * G + G_rem == steps  # every batch is timed exactly once
                assert len(assign) == full + (1 if G_rem else 0) and all(0 <= a < n_ctx for a in assign)
                if G_rem:  # the remainder set is alone on the last context: its resident upload has G_rem groups
                    assert n_ctx > 1 and assign[-1] == n_ctx - 1 and n_ctx - 1 not in assign[:-1]
                if n_ctx == 1:
                    assert G_rem == 0


def test_algorithmic_work_model_counts():
    class BV:  # the vector_mul shape (12 points, 20 scalars, 8 squeezes, 6 shared bases, 2 multi-open points)
        n_points, n_scalars, n_challenges, n_shared, n_mo = 12, 20, 8, 6, 2

    geom = {"window_bits": 11 | (9 << 16), "windows": 24 | (29 << 16), "terms": 57350, "buckets": 32000}
    mm = bench.algorithmic_mm(BV, 4096, geom)
    assert mm["decompress"] == 4096 * 12 * 313 and set(mm) == {"decompress", "transcript", "scalar", "rlc_msm", "pairing"}
    assert 7500 < sum(mm.values()) / 4096 < 9500  # ~8.3 k Montgomery multiplications per proof
    assert bench.bucket_sum_mm(BV, 4096, geom) == (24 * (4096 * 12 + 6) + 29 * 4096) * 11
