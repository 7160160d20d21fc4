"""Seeded synthetic workloads shared by the tests (oracle-side generation: TEST INFRASTRUCTURE)."""
import functools
import random

import bn254 as bn
import formats as F
import prover_sim as sim
import verifier as orc


@functools.lru_cache(maxsize=None)
def setup(shape, k, seed=0):
    rng = random.Random(("setup", shape, k, seed).__repr__())
    s = sim.FIXTURE_SRS_SECRET if k == 8 else rng.randrange(1, bn.R)
    params = sim.make_params(k, s)
    vk, dl = sim.make_vk(shape, k, seed)
    return params, vk, dl, s


def make_batch(shape, k, n, multiopen="shplonk", hash_kind="blake2b", seed=1, rows=10):
    params, vk, dl, s = setup(shape, k)
    rng = random.Random(("batch", shape, k, n, multiopen, hash_kind, seed).__repr__())
    instances = [sim.random_instances(vk, rng, rows) for _ in range(n)]
    proofs = [sim.simulate_proof(params, vk, dl, s, inst, rng, multiopen, hash_kind) for inst in instances]
    return params, vk, instances, proofs, rng


def enc_point(p):
    return bytes(64) if p is None else p[0].to_bytes(32, "little") + p[1].to_bytes(32, "little")


def split32(b, n):
    return [int.from_bytes(b[32 * i: 32 * i + 32], "little") for i in range(n)]


def oracle_scalars(vk, res, n_points, n_mo):
    """Per-base MSM scalars of an oracle Result in the C ABI's hook order:
    right[proof points] | shared[fixed | sigma | G] | left[multi-open points]."""
    rb, lb = res.right.by_base(), res.left.by_base()
    out = [rb.get(("proof", i), 0) for i in range(n_points)]
    out += [rb.get(("fixed", i), 0) for i in range(len(vk.fixed_commitments))]
    out += [rb.get(("sigma", i), 0) for i in range(len(vk.permutation_commitments))]
    out.append((rb.get(("g",), 0) - rb.get(("-g",), 0)) % bn.R)
    out += [lb.get(("proof", n_points - n_mo + i), 0) for i in range(n_mo)]
    return out
