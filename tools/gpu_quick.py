"""Ad-hoc GPU check used during development: field self test, smoke, timing of one batch."""
import os, sys, time, random, ctypes
from multiprocessing import Pool
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as g
import bn254 as bn, prover_sim as sim, verifier as orc

def gen(args):
    shape, k, s, mo, seed = args
    rng = random.Random(seed)
    params = sim.make_params(k, s); vk, dl = sim.make_vk(shape, k)
    inst = sim.random_instances(vk, rng, 10)
    return inst, sim.simulate_proof(params, vk, dl, s, inst, rng, mo)

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    pkg = g.load_package()
    lib = pkg.load_library()
    t = time.time(); bad = lib.h2v_selftest_field(0, 1 << 20, 12345); print("field selftest mismatches:", bad, "%.2fs" % (time.time() - t), flush=True)
    print("imad/s: %.3e" % lib.h2v_calibrate_imad(0), flush=True)
    g.smoke()
    shape, k, mo = "vm", 10, "shplonk"
    s = random.Random(2).randrange(bn.R)
    params = sim.make_params(k, s); vk, dl = sim.make_vk(shape, k)
    t = time.time()
    with Pool(min(32, os.cpu_count())) as p:
        items = p.map(gen, [(shape, k, s, mo, 1000 + i) for i in range(n)], chunksize=16)
    print("generated %d proofs in %.1fs on %d cpus" % (n, time.time() - t, os.cpu_count()), flush=True)
    instances = [it[0][0] for it in items]; proofs = [it[1] for it in items]
    bv = pkg.BatchVerifier(pkg.ParamsKZG.from_bytes(params.to_bytes()), pkg.VerifyingKey.from_bytes(vk.to_bytes(1)), mo, "blake2b", 0)
    for rep in range(4):
        t = time.time(); res = bv.verify_batch(proofs, instances, seed=7); dt = time.time() - t
        print("rep", rep, "verdict", res.verdict, "wall %.2f ms" % (dt * 1e3), {k_: round(v, 3) for k_, v in bv.timings().items()}, bv.msm_geometry(), flush=True)
    # spot-check a few proofs against the oracle
    res = bv.verify_batch(proofs[:16], instances[:16], seed=7, want_challenges=True, want_accum=True)
    for j in range(16):
        w = orc.verify_proof(params, vk, [instances[j]], proofs[j])
        C = bv.n_challenges
        got = [int.from_bytes(res.challenges[32*(j*C+c):32*(j*C+c+1)], "little") for c in range(C)]
        assert got == w.challenges and res.status[j] == w.status
        enc = lambda p: bytes(64) if p is None else p[0].to_bytes(32, "little") + p[1].to_bytes(32, "little")
        assert res.accum[128*j:128*j+128] == enc(w.L) + enc(w.R), j
    print("oracle spot-check ok (16 proofs: challenges, per-proof accumulators, statuses)")
    # corrupted batch -> attribution
    rng = random.Random(5); bad_idx = sorted(rng.sample(range(n), max(1, n // 100)))
    p2 = list(proofs)
    for i in bad_idx: p2[i], _ = sim.corrupt(p2[i], vk, "eval_flip", rng)
    t = time.time(); res = bv.verify_batch(p2, instances, seed=7); dt = time.time() - t
    print("corrupted batch: wall %.2f ms; flagged == injected:" % (dt * 1e3), [i for i, s_ in enumerate(res.status) if s_] == bad_idx, {k_: round(v, 3) for k_, v in bv.timings().items()})

if __name__ == "__main__":
    main()
