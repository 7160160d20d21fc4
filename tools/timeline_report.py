"""Text report of a per-block timeline recorded with H2V_BENCH_DIAG_TIMELINE=<file.npy> python bench.py ...
(csrc/timeline.cuh: one record per thread block {kernel id, block, SM, tag, start, end} on the global nanosecond timer).

Prints the kernel instances in start order (start, end, duration, kernel, blocks) and, per 0.25 ms slice, how many blocks of
the multiplier-bound kernels (decompress, bucket_sum) and of the other kernels were resident on average."""
import sys

import numpy as np

NAMES = {1: "decompress", 2: "transcript", 3: "scalar", 4: "digits", 5: "scatter", 6: "bucket_sum", 7: "chunk_reduce",
         8: "window_reduce", 9: "lines", 10: "pairing"}
HEAVY = (1, 6)


def load(path):
    a = np.load(path).astype(np.uint64)
    kid, tag = a[:, 0].astype(np.int64), a[:, 3].astype(np.int64)
    t0 = a[:, 4] | (a[:, 5] << np.uint64(32))
    t1 = a[:, 6] | (a[:, 7] << np.uint64(32))
    keep = kid > 0
    return kid[keep], tag[keep], t0[keep].astype(np.int64), t1[keep].astype(np.int64)


def instances(kid, tag, t0, t1):
    out = []
    for k in np.unique(kid):
        for tg in np.unique(tag[kid == k]):
            m = (kid == k) & (tag == tg)
            s0, s1 = t0[m], t1[m]
            o = np.argsort(s0)
            s0, s1 = s0[o], s1[o]
            start, mx = 0, s1[0]
            for i in range(1, len(s0)):
                if s0[i] > mx + 30000:  # 30 us without a running block of this (kernel, context): next instance
                    out.append((s0[start], s1[start:i].max(), int(k), int(tg), i - start))
                    start = i
                mx = max(mx, s1[i])
            out.append((s0[start], s1[start:].max(), int(k), int(tg), len(s0) - start))
    out.sort()
    return out


def main():
    kid, tag, t0, t1 = load(sys.argv[1])
    base = t0.min()
    inst = instances(kid, tag, t0, t1)
    print(f"records {len(kid)}, window {(t1.max() - base) * 1e-6:.2f} ms, kernel instances {len(inst)}")
    limit = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    for s, e, k, tg, nblk in inst[:limit]:
        print(f"  {(s - base) * 1e-6:8.3f} {(e - base) * 1e-6:8.3f}  {(e - s) * 1e-6:6.3f} ms  {NAMES.get(k, k):<13} blocks {nblk:5d}  tag {tg & 0xffff:04x}")
    # residency per slice
    dt = 250000
    nsl = int((t1.max() - base) // dt) + 1
    heavy, other = np.zeros(nsl), np.zeros(nsl)
    for hv, arr in ((True, heavy), (False, other)):
        m = np.isin(kid, HEAVY) == hv
        for a, b in zip(t0[m] - base, t1[m] - base):
            i0, i1 = a // dt, b // dt
            if i0 == i1:
                arr[i0] += (b - a) / dt
            else:
                arr[i0] += ((i0 + 1) * dt - a) / dt
                arr[i1] += (b - i1 * dt) / dt
                if i1 > i0 + 1:
                    arr[i0 + 1:i1] += 1
    print("slice(ms)  resident heavy blocks  resident other blocks")
    for i in range(nsl):
        print(f"  {i * dt * 1e-6:7.2f}  {heavy[i]:8.1f}  {other[i]:8.1f}")
    print(f"fraction of slices with < 148 heavy blocks resident: {(heavy < 148).mean():.3f}")


if __name__ == "__main__":
    main()
