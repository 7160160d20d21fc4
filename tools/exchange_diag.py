"""Diagnosis of the device-side exchange on ONE GPU: two contexts of one process play two ranks.  H2V_TRACE=1 prints host
timestamps of every step.  Variants: concurrent cold start; solo warm-ups first (short timeout), then concurrent."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import __graft_entry__ as g
import formats as F
from workloads import make_batch


def main():
    pkg = g.load_package()
    n, world = 64, 2
    per = n // world
    params, vk, instances, proofs, rng = make_batch("vm", 10, n, "shplonk", "blake2b", seed=61)
    insts = [i[0] for i in instances]
    mk = lambda: pkg.BatchVerifier(pkg.ParamsKZG.from_bytes(params.to_bytes()), pkg.VerifyingKey.from_bytes(vk.to_bytes(F.RAW_BYTES), F.RAW_BYTES), "shplonk", "blake2b", device=0)
    variant = sys.argv[1] if len(sys.argv) > 1 else "cold"
    bvs = [mk() for _ in range(world)]
    handles = [bv.comm_init(r, world, max_groups=3) for r, bv in enumerate(bvs)]
    [bv.comm_connect(handles) for bv in bvs]
    lib = bvs[0].lib
    if os.environ.get("DIAG_NOGRAPH"):
        [bv.set_graphs(False) for bv in bvs]

    def call(r, root, tag):
        t = time.time()
        try:
            v, st = bvs[r].verify_shard(proofs[r * per:(r + 1) * per], insts[r * per:(r + 1) * per], r * per, n, root, seed=5)
            print(f"[{tag}] rank {r}: verdicts {v} bad {sum(1 for x in st if x)} in {time.time() - t:.3f}s", flush=True)
        except Exception as e:
            print(f"[{tag}] rank {r}: ERROR {e} after {time.time() - t:.3f}s", flush=True)

    if variant == "warm":
        for r in range(world):  # solo calls: they time out (no peer), but allocate / load / capture everything
            lib.h2v_comm_set_timeout_ms(bvs[r]._ctx, 200)
            call(r, 0, "solo")
        # the channel's sequence numbers are now out of step by design (both ran one launch set alone): still equal on both ranks
        for r in range(world):
            lib.h2v_comm_set_timeout_ms(bvs[r]._ctx, 10000)
    for rep in range(3):
        ths = [threading.Thread(target=call, args=(r, rep % 2, f"{variant}{rep}")) for r in range(world)]
        [t.start() for t in ths]
        [t.join() for t in ths]
    [bv.close() for bv in bvs]


if __name__ == "__main__":
    main()
