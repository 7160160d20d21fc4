"""Generates the committed golden vectors tests/golden/*.json from the Python oracle (oracle/).
Each file: params / vk bytes, a few proofs (valid and corrupted) with their public inputs, and the
oracle's outputs: status, transcript challenges, per-base MSM scalars, per-proof accumulators
(L_j, R_j), RLC scalars and the folded (L, R).  Run:  python tests/golden/make_golden.py"""
import json, os, random, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle")); sys.path.insert(0, os.path.join(HERE, ".."))
import bn254 as bn, formats as F, prover_sim as sim, verifier as orc
from workloads import setup, enc_point, oracle_scalars

CASES = [  # name, shape, k, multiopen, hash, vk format
    ("vm_k8_shplonk_blake2b", "vm", 8, "shplonk", "blake2b", F.RAW_BYTES),
    ("vm_k8_gwc_keccak", "vm", 8, "gwc", "keccak", F.PROCESSED),
    ("sh_k8_shplonk_keccak", "sh", 8, "shplonk", "keccak", F.RAW_BYTES),
    ("mix_k6_shplonk_blake2b", "mix", 6, "shplonk", "blake2b", F.PROCESSED),
    ("mix_k6_gwc_blake2b", "mix", 6, "gwc", "blake2b", F.RAW_BYTES),
]
for name, shape, k, mo, hk, vkfmt in CASES:
    params, vk, dl, s = setup(shape, k)
    rng = random.Random("golden-" + name)
    n = 4
    instances = [sim.random_instances(vk, rng, 5 + j) for j in range(n)]
    proofs = [sim.simulate_proof(params, vk, dl, s, inst, rng, mo, hk) for inst in instances]
    proofs[1], _ = sim.corrupt(proofs[1], vk, "eval_flip", rng, mo)
    proofs[3], _ = sim.corrupt(proofs[3], vk, "scalar_ge_r", rng, mo)
    rs = [rng.randrange(1, bn.R) for _ in range(n)]
    results = [orc.verify_proof(params, vk, inst, p, mo, hk) for inst, p in zip(instances, proofs)]
    items, first_mo = sim.proof_layout(vk, mo)
    n_points = items.count("P"); n_mo = len(items) - first_mo
    L, Rr, ok = orc.accumulate(params, results, rs)
    out = {
        "shape": shape, "k": k, "multiopen": mo, "hash": hk, "vk_format": vkfmt,
        "params": params.to_bytes().hex(), "vk": vk.to_bytes(vkfmt).hex(),
        "rlc_scalars": [hex(r) for r in rs], "folded": (enc_point(L) + enc_point(Rr)).hex(), "folded_ok": ok,
        "proofs": [],
    }
    for inst, p, res in zip(instances, proofs, results):
        e = {"proof": p.hex(), "instances": [[hex(v) for v in col] for col in inst[0]], "status": res.status,
             "challenges": [hex(c) for c in res.challenges]}
        if res.left is not None:
            e["accum"] = (enc_point(res.L) + enc_point(res.R)).hex()
            e["msm_scalars"] = [hex(v) for v in oracle_scalars(vk, res, n_points, n_mo)]
        out["proofs"].append(e)
    json.dump(out, open(os.path.join(HERE, name + ".json"), "w"), indent=0)
    print(name, [r.status for r in results], "folded_ok", ok)

# ---- proofs carrying m = 3 circuit instances in one transcript (h2v_ctx_create_multi; `instances.len() = 3` in the reference)
for name, shape, k, mo, hk, vkfmt, m in [("mix_k6_shplonk_blake2b_m3", "mix", 6, "shplonk", "blake2b", F.RAW_BYTES, 3)]:
    params, vk, dl, s = setup(shape, k)
    rng = random.Random("golden-" + name)
    n = 3
    instances = [[sim.random_instances(vk, rng, 4 + j)[0] for _ in range(m)] for j in range(n)]
    proofs = [sim.simulate_proof(params, vk, dl, s, inst, rng, mo, hk) for inst in instances]
    proofs[1], _ = sim.corrupt(proofs[1], vk, "eval_flip", rng, mo, m)
    rs = [rng.randrange(1, bn.R) for _ in range(n)]
    results = [orc.verify_proof(params, vk, inst, p, mo, hk) for inst, p in zip(instances, proofs)]
    items, first_mo = sim.proof_layout(vk, mo, m)
    n_points = items.count("P"); n_mo = len(items) - first_mo
    L, Rr, ok = orc.accumulate(params, results, rs)
    out = {"shape": shape, "k": k, "multiopen": mo, "hash": hk, "vk_format": vkfmt, "circuit_instances": m,
           "params": params.to_bytes().hex(), "vk": vk.to_bytes(vkfmt).hex(),
           "rlc_scalars": [hex(r) for r in rs], "folded": (enc_point(L) + enc_point(Rr)).hex(), "folded_ok": ok, "proofs": []}
    for inst, p, res in zip(instances, proofs, results):  # "instances": [circuit instance][column][row]
        out["proofs"].append({"proof": p.hex(), "instances": [[[hex(v) for v in col] for col in ci] for ci in inst], "status": res.status,
                              "challenges": [hex(c) for c in res.challenges], "accum": (enc_point(res.L) + enc_point(res.R)).hex(),
                              "msm_scalars": [hex(v) for v in oracle_scalars(vk, res, n_points, n_mo)]})
    json.dump(out, open(os.path.join(HERE, name + ".json"), "w"), indent=0)
    print(name, [r.status for r in results], "folded_ok", ok)

# ---- honest proofs of the vector_mul circuit (oracle/honest_prover.py: real witness, real polynomials, k = 8 fixture SRS secret)
import honest_prover as hp
rng = random.Random("golden-honest")
s = sim.FIXTURE_SRS_SECRET
params, vk, pk = hp.keygen_vm(8, s, 10)
proofs, instances = [], []
for j in range(3):
    lhs = [rng.randrange(bn.R) for _ in range(10)]
    rhs = [rng.randrange(bn.R) for _ in range(10)]
    p, inst = hp.prove_vm(params, vk, pk, s, lhs, rhs, rng, cheat_row=(4 if j == 1 else None))
    proofs.append(p); instances.append(inst)
inst_wrong = [[list(instances[2][0][0])]]
inst_wrong[0][0][3] = (inst_wrong[0][0][3] + 1) % bn.R  # the reference's own negative test: wrong public input (vector_mul.rs:327-330)
proofs.append(proofs[2]); instances.append(inst_wrong)
rs = [rng.randrange(1, bn.R) for _ in proofs]
results = [orc.verify_proof(params, vk, inst, p) for inst, p in zip(instances, proofs)]
assert [r.status for r in results] == [0, 4, 0, 4], [r.status for r in results]
L, Rr, ok = orc.accumulate(params, results, rs)
out = {"shape": "vm-honest", "k": 8, "multiopen": "shplonk", "hash": "blake2b", "vk_format": F.RAW_BYTES,
       "params": params.to_bytes().hex(), "vk": vk.to_bytes(F.RAW_BYTES).hex(),
       "rlc_scalars": [hex(r) for r in rs], "folded": (enc_point(L) + enc_point(Rr)).hex(), "folded_ok": ok, "proofs": []}
for inst, p, res in zip(instances, proofs, results):
    out["proofs"].append({"proof": p.hex(), "instances": [[hex(v) for v in col] for col in inst[0]], "status": res.status,
                          "challenges": [hex(c) for c in res.challenges], "accum": (enc_point(res.L) + enc_point(res.R)).hex(),
                          "msm_scalars": [hex(v) for v in oracle_scalars(vk, res, 12, 2)]})
json.dump(out, open(os.path.join(HERE, "vm_k8_honest_prover.json"), "w"), indent=0)
print("vm_k8_honest_prover", [r.status for r in results])

# ---- honest proofs of the lookup + shuffle + rotated-gate circuit (k = 6, seeded SRS secret)
rng = random.Random("golden-honest-lk")
s = rng.randrange(1, bn.R)
circ = hp.lookup_shuffle_circuit(6, 16)
params, vk, pk = hp.keygen(circ, s)
proofs, instances = [], []
for cheat in (None, "lookup", None, "shuffle"):
    adv, ins = hp.lookup_shuffle_assignment(circ, 16, rng, cheat)
    proofs.append(hp.prove(params, vk, pk, s, adv, ins, rng, "blake2b", expect_honest=cheat is None))
    instances.append([ins])
rs = [rng.randrange(1, bn.R) for _ in proofs]
results = [orc.verify_proof(params, vk, inst, p) for inst, p in zip(instances, proofs)]
assert [r.status for r in results] == [0, 4, 0, 4], [r.status for r in results]
items, first_mo = sim.proof_layout(vk, "shplonk")
n_points = items.count("P"); n_mo = len(items) - first_mo
L, Rr, ok = orc.accumulate(params, results, rs)
out = {"shape": "lookup-shuffle-honest", "k": 6, "multiopen": "shplonk", "hash": "blake2b", "vk_format": F.RAW_BYTES,
       "params": params.to_bytes().hex(), "vk": vk.to_bytes(F.RAW_BYTES).hex(),
       "rlc_scalars": [hex(r) for r in rs], "folded": (enc_point(L) + enc_point(Rr)).hex(), "folded_ok": ok, "proofs": []}
for inst, p, res in zip(instances, proofs, results):
    out["proofs"].append({"proof": p.hex(), "instances": [[hex(v) for v in col] for col in inst[0]], "status": res.status,
                          "challenges": [hex(c) for c in res.challenges], "accum": (enc_point(res.L) + enc_point(res.R)).hex(),
                          "msm_scalars": [hex(v) for v in oracle_scalars(vk, res, n_points, n_mo)]})
json.dump(out, open(os.path.join(HERE, "lk_k6_honest_prover.json"), "w"), indent=0)
print("lk_k6_honest_prover", [r.status for r in results])
