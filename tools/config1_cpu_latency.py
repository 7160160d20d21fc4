#!/usr/bin/env python
"""BASELINE.json configs[0] / BASELINE.md row 1: ONE `verify_proof` (KZG-SHPLONK, Blake2b) of the reference's small test
circuit (halo2_verifier/tests/vector_mul.rs: k = 8, 10 multiplications) with the reference's own params fixture
(halo2_verifier/params/kzg_bn254_8.srs -> g, g2, [s]g2), on ONE CPU core: p50 / p90 latency of 1,000 runs.

The implementation timed is the C restatement of the reference algorithm (oracle/c: SingleStrategy = the reference's serial
windowed MSM with c in {1,3,4} and one 2-pair pairing per proof; 4x64-bit Montgomery arithmetic) - the Rust binary cannot
be built in this image.  The proof is an HONEST proof of that circuit (oracle/honest_prover.py), the same bytes
tools/export_reference_fixtures.py writes for the reference-side kit.  CPU only.
    python tools/config1_cpu_latency.py [runs]
"""
import json
import os
import platform
import random
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import c_oracle  # noqa: E402
import formats as F  # noqa: E402
import honest_prover as hp  # noqa: E402
import prover_sim as sim  # noqa: E402


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    k, rows, s = 8, 10, sim.FIXTURE_SRS_SECRET
    rng = random.Random("config1")
    params, vk, pk = hp.keygen_vm(k, s, rows)
    lhs = [rng.randrange(1, 1 << 64) for _ in range(rows)]
    rhs = [rng.randrange(1, 1 << 64) for _ in range(rows)]
    proof, inst = hp.prove_vm(params, vk, pk, s, lhs, rhs, rng)
    co = c_oracle.COracle(params.to_bytes(F.RAW_BYTES), 1, vk.to_bytes(F.RAW_BYTES), 1)
    st, _, _ = co.verify(proof, inst[0])
    assert st == 0, "the honest proof must verify"
    bad = [list(inst[0][0])]
    bad[0][0] = (bad[0][0] + 1) % (1 << 200)
    assert co.verify(proof, bad)[0] == 4  # the reference's negative test: ConstraintSystemFailure
    for _ in range(20):
        co.verify(proof, inst[0])
    ts = []
    for _ in range(runs):
        a = time.perf_counter()
        co.verify(proof, inst[0])
        ts.append((time.perf_counter() - a) * 1e3)
    ts.sort()
    cpu = ""
    try:
        cpu = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        cpu = platform.processor()
    out = {"config": "BASELINE.json configs[0]: single verify_proof, vector_mul test circuit, k = 8, fixture params, SHPLONK / Blake2b, honest proof",
           "implementation": "oracle/c (C restatement of the reference algorithm, SingleStrategy), one thread, incl. the ctypes call (~2 us)",
           "runs": runs, "p50_ms": round(statistics.median(ts), 4), "p90_ms": round(ts[int(0.9 * runs)], 4), "min_ms": round(ts[0], 4),
           "proofs_per_s_one_core": round(1e3 / statistics.median(ts), 1), "proof_bytes": len(proof), "cpu": cpu}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
