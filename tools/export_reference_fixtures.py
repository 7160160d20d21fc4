#!/usr/bin/env python
"""Writes the honest mini-prover's output in the reference tooling's own on-disk layout, so that a maintainer
with cargo can feed the SAME bytes to the real `halo2_verifier::verify_proof` (closing "parity unpinned"):

    VALID_VK.bin      ParamsKZG::write (Processed form, 164 bytes: poly/kzg/commitment.rs:142-152,209-213)
                      immediately followed by VerifyingKey::write(SerdeFormat::RawBytes) (plonk/vk.rs:41-64)
                                                                        -- serialize/examples/vector_mul.rs:374-393
    VALID_PROOF.bin   the transcript bytes (Blake2b writer, SHPLONK)    -- serialize/examples/vector_mul.rs:361-365
    VALID_PUBS.bin    public inputs as consecutive 32-byte `to_bytes()` -- serialize/examples/vector_mul.rs:367-371
    EXPECTED.json     what this repository's oracle and CUDA path say about them: verdict, transcript challenges,
                      per-proof accumulators (L, R) - the values `integration/rust` asserts against the reference

The circuit is the reference's own test circuit (halo2_verifier/tests/vector_mul.rs: k = 8, 10 multiplications, the
fixture SRS `halo2_verifier/params/kzg_bn254_8.srs`), proved by oracle/honest_prover.py with a real witness.  The
reference then runs:  cargo run --example verify_bundle -- VALID_VK.bin VALID_PROOF.bin VALID_PUBS.bin  (or the three
lines of INTEGRATION.md).  Usage:  python tools/export_reference_fixtures.py [out_dir]   (CPU only, no GPU needed)
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import formats as F  # noqa: E402
import honest_prover as hp  # noqa: E402
import prover_sim as sim  # noqa: E402
import verifier as orc  # noqa: E402


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "integration", "fixtures")
    os.makedirs(out, exist_ok=True)
    k, rows, s = 8, 10, sim.FIXTURE_SRS_SECRET
    rng = random.Random("export-reference-fixtures")
    params, vk, pk = hp.keygen_vm(k, s, rows)
    lhs = [rng.randrange(1, 1 << 64) for _ in range(rows)]
    rhs = [rng.randrange(1, 1 << 64) for _ in range(rows)]
    proof, inst = hp.prove_vm(params, vk, pk, s, lhs, rhs, rng)
    res = orc.verify_proof(params, vk, inst, proof)
    assert res.status == orc.OK, "the honest proof must verify in the oracle"
    bundle = params.to_bytes(F.PROCESSED) + vk.to_bytes(F.RAW_BYTES)
    assert len(params.to_bytes(F.PROCESSED)) == 4 + 32 + 64 + 64
    pubs = b"".join(int(v).to_bytes(32, "little") for v in inst[0][0])
    open(os.path.join(out, "VALID_VK.bin"), "wb").write(bundle)
    open(os.path.join(out, "VALID_PROOF.bin"), "wb").write(proof)
    open(os.path.join(out, "VALID_PUBS.bin"), "wb").write(pubs)
    # a rejected companion, as the reference's own negative test makes it (tests/vector_mul.rs:327-330: first public input + 1)
    bad_pubs = bytearray(pubs)
    bad_pubs[:32] = ((int.from_bytes(pubs[:32], "little") + 1) % orc.bn.R).to_bytes(32, "little")
    open(os.path.join(out, "INVALID_PUBS.bin"), "wb").write(bytes(bad_pubs))
    bad_inst = [[list(inst[0][0])]]
    bad_inst[0][0][0] = (bad_inst[0][0][0] + 1) % orc.bn.R
    res_bad = orc.verify_proof(params, vk, bad_inst, proof)
    enc = lambda p: (bytes(64) if p is None else p[0].to_bytes(32, "little") + p[1].to_bytes(32, "little")).hex()
    json.dump({
        "circuit": "halo2_verifier/tests/vector_mul.rs (k = 8, 10 rows), fixture SRS, Blake2b transcript, SHPLONK",
        "proof_bytes": len(proof), "bundle_bytes": len(bundle),
        "valid": {"status": res.status, "challenges": [hex(c) for c in res.challenges], "L": enc(res.L), "R": enc(res.R)},
        "invalid_pubs": {"status": res_bad.status, "expected_error": "ConstraintSystemFailure"},
    }, open(os.path.join(out, "EXPECTED.json"), "w"), indent=1)
    print(f"wrote VALID_VK.bin ({len(bundle)} B), VALID_PROOF.bin ({len(proof)} B), VALID_PUBS.bin ({len(pubs)} B), INVALID_PUBS.bin, EXPECTED.json to {out}")
    print("oracle: valid ->", res.status, " first public input + 1 ->", res_bad.status)


if __name__ == "__main__":
    main()
