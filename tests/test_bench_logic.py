"""Host logic of bench.py that needs no GPU: the launch-set plan and the algorithmic work model."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def test_launch_set_plan_is_balanced_and_rank_independent():
    for n_ctx in (1, 2, 4, 8):
        for steps in list(range(1, 70)) + [100, 128, 255, 256, 257]:
            assign = bench.plan_launch_sets(steps, n_ctx)
            assert len(assign) == steps and all(0 <= a < n_ctx for a in assign)  # every launch set is timed exactly once
            per = [assign.count(c) for c in range(n_ctx)]
            assert max(per) - min(per) <= 1
    assert bench.plan_launch_sets(20, 4) == [0, 1, 2, 3] * 5  # the driver's --steps 20 spreads evenly over the default 4 contexts


def test_algorithmic_work_model_counts():
    class BV:  # the vector_mul shape (12 points, 20 scalars, 8 squeezes, 6 shared bases, 2 multi-open points)
        n_points, n_scalars, n_challenges, n_shared, n_mo = 12, 20, 8, 6, 2
        multiopen = "shplonk"

        @staticmethod
        def work_model(rows):  # what h2v_ctx_work_model returns for this plan (checked on the GPU box in tests/test_exchange.py)
            return {"scalar": 700.0, "transcript": 60.0, "decompress_per_point": 313.0, "instance_scalars": float(rows)}

    geom = {"window_bits": 11 | (9 << 16), "windows": 24 | (29 << 16), "terms": 57350, "buckets": 32000}
    mm = bench.algorithmic_mm(BV, 4096, geom)
    assert mm["decompress"] == 4096 * 12 * 313 and set(mm) == {"decompress", "transcript", "scalar", "rlc_msm", "pairing"}
    assert 7500 < sum(mm.values()) / 4096 < 9500  # ~8.3 k Montgomery multiplications per proof
    assert bench.bucket_sum_mm(BV, 4096, geom) == (24 * (4096 * 12 + 6) + 29 * 4096) * 11
    BV.multiopen = "gwc"  # every multi-open point carries a left scalar
    assert bench.bucket_sum_mm(BV, 4096, geom) == (24 * (4096 * 12 + 6) + 29 * 4096 * 2) * 11


def test_timeline_report_splits_kernel_instances(tmp_path):
    """tools/timeline_report.py: block records {kernel, block, SM, tag, t0, t1} -> kernel instances (a gap of more than 30 us
    without a running block of the same kernel and context starts a new instance)."""
    import importlib.util
    import os

    import numpy as np

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("timeline_report", os.path.join(root, "tools", "timeline_report.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    def rec(kid, blk, tag, t0, t1):
        return [kid, blk, 0, tag, t0 & 0xFFFFFFFF, t0 >> 32, t1 & 0xFFFFFFFF, t1 >> 32]

    base = (7 << 32) + 1000  # timestamps beyond 32 bits: the two halves are recombined
    rows = [rec(1, b, 5, base + 1000 * b, base + 1000 * b + 50000) for b in range(10)]          # one instance of kernel 1
    rows += [rec(1, b, 5, base + 500000 + 1000 * b, base + 560000) for b in range(4)]           # a second one, 0.44 ms later
    rows += [rec(6, 0, 9, base + 20000, base + 90000), rec(0, 0, 0, 0, 0)]                      # another kernel; an empty slot
    path = tmp_path / "tl.npy"
    np.save(path, np.array(rows, dtype=np.uint32))
    kid, tag, t0, t1 = mod.load(str(path))
    assert len(kid) == 15 and int(t0.min()) == base
    inst = mod.instances(kid, tag, t0, t1)
    assert [(k, n) for _, _, k, _, n in inst] == [(1, 10), (6, 1), (1, 4)]
    assert inst[0][1] - inst[0][0] == 59000
