"""BASELINE.json configs[2] timing outside bench.py: 4096-proof GWC batch, clean and with 1 % corrupted proofs, through
h2v_verify_batch; with H2V_TRACE=1 the library prints host timestamps of the attribution phases (synchronising between
them, so the traced total is slower than the untraced one)."""
import os, sys, time, random, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import __graft_entry__ as g
import prover_sim as sim
from workloads import make_batch


def main():
    pkg = g.load_package()
    n, distinct, mo = 4096, 64, "gwc"
    params, vk, instances, proofs, rng = make_batch("vm", 10, distinct, mo, "blake2b", seed=31)
    proofs, insts = proofs * (n // distinct), [i[0] for i in instances] * (n // distinct)
    bad = list(proofs)
    idx = sorted(rng.sample(range(n), n // 100))
    kinds = list(sim.CORRUPTIONS)
    for t, i in enumerate(idx):
        bad[i], _ = sim.corrupt(proofs[i], vk, kinds[t % len(kinds)], rng, mo)
    bv = pkg.BatchVerifier(pkg.ParamsKZG.from_bytes(params.to_bytes()), pkg.VerifyingKey.from_bytes(vk.to_bytes(1)), mo, "blake2b", 0)
    reps = int(os.environ.get("REPS", "6"))

    def timed(pr, ins, groups):
        ts = []
        for _ in range(reps):
            t = time.perf_counter(); res = bv.verify_batch(pr, ins, fold_groups=groups); ts.append(time.perf_counter() - t)
        return res, statistics.median(ts[1:]) * 1e3

    res, ms = timed(proofs, insts, 1)
    assert res.verdict
    print("GWC, 4096 proofs, all valid:            %.2f ms per batch (incl. Python packing of the inputs)" % ms)
    res, ms_bad = timed(bad, insts, 1)
    assert not res.verdict and [i for i, s_ in enumerate(res.status) if s_] == idx
    print("GWC, 4096 proofs, 1 %% corrupted:        %.2f ms per batch (batch check rejects -> attribution), %d flagged == injected" % (ms_bad, len(idx)))
    if os.environ.get("H2V_TL"):  # per-kernel spans of ONE rejected batch incl. its attribution (block timeline, csrc/timeline.cuh)
        import ctypes
        import numpy as np
        lib, cap = pkg._lib, 1 << 20
        assert lib.h2v_debug_timeline_start(0, cap) == 0
        bv.verify_batch(bad, insts)
        buf, cnt = np.zeros(cap * 8, dtype=np.uint32), ctypes.c_uint32(0)
        assert lib.h2v_debug_timeline_stop(0, buf.ctypes.data, cap, ctypes.byref(cnt)) == 0
        rec = buf[: cnt.value * 8].reshape(-1, 8).astype(np.uint64)
        kid, t0, t1 = rec[:, 0], rec[:, 4] | (rec[:, 5] << np.uint64(32)), rec[:, 6] | (rec[:, 7] << np.uint64(32))
        names = {1: "decompress", 2: "transcript", 3: "scalar", 4: "digits", 5: "scatter", 6: "bucket_sum", 7: "chunk_reduce", 8: "window_reduce", 9: "lines",
                 10: "pairing", 11: "window_combine", 12: "pp_mul_list", 13: "pp_reduce_list", 14: "rlc_scan", 15: "shared_reduce", 16: "bucket_order"}
        base = int(t0.min())
        ev = []
        for k_ in np.unique(kid):
            m = kid == k_
            s0, s1 = np.sort(t0[m]), t1[m][np.argsort(t0[m])]
            start, mx = 0, int(s1[0])
            for i in range(1, len(s0)):
                if int(s0[i]) > mx + 20000:
                    ev.append((int(s0[start]) - base, int(s1[start:i].max()) - base, int(k_), i - start)); start = i
                mx = max(mx, int(s1[i]))
            ev.append((int(s0[start]) - base, int(s1[start:].max()) - base, int(k_), len(s0) - start))
        for a, b, k_, nblk in sorted(ev):
            print("  %8.3f -> %8.3f ms  (%6.3f)  %-14s blocks %d" % (a * 1e-6, b * 1e-6, (b - a) * 1e-6, names.get(k_, k_), nblk))
        return
    if os.environ.get("H2V_TRACE"):
        return
    G = 8
    res, ms_g = timed(proofs * (G - 1) + bad, insts * G, G)
    assert res.group_verdicts == [True] * (G - 1) + [False] and [i - n * (G - 1) for i, s_ in enumerate(res.status) if s_] == idx
    res2, ms_g0 = timed(proofs * G, insts * G, G)
    assert res2.verdict
    print("8 fold groups (32768 proofs), all valid: %.2f ms; last group 1 %% corrupted: %.2f ms (attribution only inside the rejected group)" % (ms_g0, ms_g))


if __name__ == "__main__":
    main()
