"""The C oracle (oracle/c) against the Python oracle: two independent restatements of the reference
algorithm must agree on statuses, challenges and accumulators for every circuit shape, multi-open
scheme, transcript hash, VK format and corruption class."""
import random

import pytest

import bn254 as bn
import c_oracle
import formats as F
import prover_sim as sim
import verifier as orc
from workloads import enc_point, make_batch


@pytest.fixture(scope="module")
def clib(built):
    return c_oracle.load()


def test_selftest(clib):
    assert clib.h2vo_selftest() == 0


@pytest.mark.parametrize("shape,k", [("vm", 8), ("sh", 8), ("mix", 6)])
@pytest.mark.parametrize("mo", ["shplonk", "gwc"])
@pytest.mark.parametrize("hk", ["blake2b", "keccak"])
def test_c_oracle_matches_python_oracle(clib, shape, k, mo, hk):
    params, vk, instances, proofs, rng = make_batch(shape, k, 5, mo, hk, seed=21)
    vfmt = F.RAW_BYTES if mo == "shplonk" else F.PROCESSED
    pfmt = F.PROCESSED if hk == "blake2b" else F.RAW_BYTES
    co = c_oracle.COracle(params.to_bytes(pfmt), pfmt, vk.to_bytes(vfmt), vfmt)
    bad = list(proofs)
    kinds = list(sim.CORRUPTIONS)
    for i in range(1, len(bad)):
        bad[i], _ = sim.corrupt(proofs[i], vk, kinds[(i + len(shape)) % len(kinds)], rng, mo)
    lrs, include, want_all = b"", [], []
    for inst, p in list(zip(instances, proofs)) + list(zip(instances, bad)):
        want = orc.verify_proof(params, vk, inst, p, mo, hk)
        st, chal, lr = co.verify(p, inst[0], mo, hk)
        assert st == want.status
        assert chal == want.challenges
        if want.status in (orc.OK, orc.CONSTRAINT_SYSTEM_FAILURE):
            assert lr == enc_point(want.L) + enc_point(want.R)
        lrs += lr
        include.append(want.status in (orc.OK, orc.CONSTRAINT_SYSTEM_FAILURE))
        want_all.append(want)
    rs = [rng.randrange(1, bn.R) for _ in want_all]
    folded, ok = co.fold(lrs, rs, include)
    L, R_, ok_py = orc.accumulate(params, want_all, rs)
    assert folded == enc_point(L) + enc_point(R_) and ok == ok_py


def test_c_oracle_batch_threads_and_instances(clib):
    import numpy as np

    params, vk, instances, proofs, rng = make_batch("vm", 8, 12, "shplonk", "blake2b", seed=22)
    co = c_oracle.COracle(params.to_bytes(), 0, vk.to_bytes(1), 1)
    insts = [list(map(list, i[0])) for i in instances]
    insts[3][0][0] = (insts[3][0][0] + 1) % bn.R  # wrong public input: the reference's own negative test
    proofs[5] = proofs[5][:500]
    pb = b"".join(proofs)
    poff = np.cumsum([0] + [len(p) for p in proofs])
    ib = b"".join(int(v).to_bytes(32, "little") for inst in insts for col in inst for v in col)
    ioff = np.cumsum([0] + [sum(len(c) for c in inst) for inst in insts])
    st1, _, lr1, ch1 = co.verify_many(pb, poff, ib, ioff, 12, threads=1, want_lr=True, chal_cap=8)
    st4, _, lr4, ch4 = co.verify_many(pb, poff, ib, ioff, 12, threads=4, want_lr=True, chal_cap=8)
    assert list(st1) == list(st4) and lr1 == lr4 and ch1 == ch4
    want = [orc.verify_proof(params, vk, [inst], p) for inst, p in zip(insts, proofs)]
    assert list(st1) == [w.status for w in want]
    assert st1[3] == orc.CONSTRAINT_SYSTEM_FAILURE and st1[5] == orc.TRANSCRIPT
    # wrong number of columns -> InvalidInstances (lib.rs:51-55)
    assert co.verify(proofs[0], [], "shplonk", "blake2b")[0] == orc.INVALID_INSTANCES


def test_c_oracle_k18_lookup_heavy(clib):
    params, vk, instances, proofs, rng = make_batch("k18", 18, 1, "shplonk", "blake2b", seed=23)
    co = c_oracle.COracle(params.to_bytes(1), 1, vk.to_bytes(1), 1)
    want = orc.verify_proof(params, vk, instances[0], proofs[0])
    st, chal, lr = co.verify(proofs[0], instances[0][0])
    assert (st, chal, lr) == (want.status, want.challenges, enc_point(want.L) + enc_point(want.R)) and st == 0


@pytest.mark.parametrize("shape,k,m", [("vm", 8, 2), ("mix", 6, 3), ("sh", 8, 2)])
@pytest.mark.parametrize("mo", ["shplonk", "gwc"])
def test_c_oracle_multi_instance_proofs(clib, shape, k, m, mo):
    """Proofs that carry m circuit instances (`instances.len() = m`): the C restatement against the Python one on valid
    proofs, every corruption class, and a wrong public input in the last instance."""
    from workloads import setup

    params, vk, dl, s = setup(shape, k)
    rng = random.Random(f"c-multi-{shape}{k}{m}{mo}")
    co = c_oracle.COracle(params.to_bytes(), 0, vk.to_bytes(1), 1)
    insts = [sim.random_instances(vk, rng, 7)[0] for _ in range(m)]
    proof = sim.simulate_proof(params, vk, dl, s, insts, rng, mo, "blake2b")
    cases = [(proof, insts)] + [(sim.corrupt(proof, vk, kind, rng, mo, m)[0], insts) for kind in sim.CORRUPTIONS]
    if vk.cs.num_instance_columns:
        wrong = [[list(c) for c in inst] for inst in insts]
        wrong[-1][0][0] = (wrong[-1][0][0] + 1) % bn.R
        cases.append((proof, wrong))
        cases.append((proof, insts[:-1]))  # one instance short: InvalidInstances for this m
    for p, ins in cases:
        want = orc.verify_proof(params, vk, ins, p, mo, "blake2b")
        cols = [col for inst in ins for col in inst]
        if len(ins) != m:  # the C entry point is told m; a missing instance shows up as a wrong column count
            inst_b = b"".join(int(v).to_bytes(32, "little") for col in cols for v in col)
            import ctypes
            cl = (ctypes.c_uint32 * max(1, len(cols)))(*[len(c) for c in cols])
            st = co.lib.h2vo_verify_multi(co.h, p, len(p), inst_b, cl, len(cols), m, 0 if mo == "shplonk" else 1, 0, 1, None, None, None)
            assert st == orc.INVALID_INSTANCES
            continue
        st, chal, lr = co.verify_multi(p, ins, mo, "blake2b")
        assert st == want.status and chal == want.challenges
        if want.status in (orc.OK, orc.CONSTRAINT_SYSTEM_FAILURE):
            assert lr == enc_point(want.L) + enc_point(want.R)
    co.close()
