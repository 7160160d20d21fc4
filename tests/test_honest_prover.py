"""Honest proofs of the vector_mul circuit (oracle/honest_prover.py): a real witness, real polynomials and the
protocol's own definitions of the permutation / vanishing arguments.  Unlike the trapdoor simulator these only
verify if the restated verifier expressions vanish on the whole domain for an honest witness."""
import json
import os
import random

import pytest

import bn254 as bn
import c_oracle
import formats as F
import honest_prover as hp
import prover_sim as sim
import verifier as orc
from workloads import enc_point

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("k,rows,hk", [(8, 10, "blake2b"), (6, 3, "keccak"), (5, 1, "blake2b")])
def test_honest_proof_is_accepted_and_cheats_are_rejected(built, k, rows, hk):
    rng = random.Random(("honest", k, rows).__repr__())
    s = sim.FIXTURE_SRS_SECRET if k == 8 else rng.randrange(1, bn.R)
    params, vk, pk = hp.keygen_vm(k, s, rows)
    lhs = [rng.randrange(bn.R) for _ in range(rows)]
    rhs = [rng.randrange(bn.R) for _ in range(rows)]
    proof, inst = hp.prove_vm(params, vk, pk, s, lhs, rhs, rng, hk)
    assert inst[0][0] == [a * b % bn.R for a, b in zip(lhs, rhs)]
    res = orc.verify_proof(params, vk, inst, proof, "shplonk", hk)
    assert res.status == orc.OK
    co = c_oracle.COracle(params.to_bytes(1), 1, vk.to_bytes(1), 1)
    st, chal, lr = co.verify(proof, inst[0], "shplonk", hk)
    assert (st, chal, lr) == (0, res.challenges, enc_point(res.L) + enc_point(res.R))
    # wrong public input: the reference's own negative test (tests/vector_mul.rs:327-330)
    wrong = [[list(inst[0][0])]]
    wrong[0][0][0] = (wrong[0][0][0] + 1) % bn.R
    assert orc.verify_proof(params, vk, wrong, proof, "shplonk", hk).status == orc.CONSTRAINT_SYSTEM_FAILURE
    assert co.verify(proof, wrong[0], "shplonk", hk)[0] == orc.CONSTRAINT_SYSTEM_FAILURE
    # a witness that violates the gate (and a consistent public input): no polynomial quotient -> rejected
    bad, bad_inst = hp.prove_vm(params, vk, pk, s, lhs, rhs, rng, hk, cheat_row=rows - 1)
    assert orc.verify_proof(params, vk, bad_inst, bad, "shplonk", hk).status == orc.CONSTRAINT_SYSTEM_FAILURE
    assert co.verify(bad, bad_inst[0], "shplonk", hk)[0] == orc.CONSTRAINT_SYSTEM_FAILURE
    co.close()


@pytest.mark.parametrize("name", ["vm_k8_honest_prover", "lk_k6_honest_prover"])
def test_honest_golden_vector_against_both_oracles(built, name):
    g = json.load(open(os.path.join(HERE, "golden", name + ".json")))
    params = F.ParamsKZG.from_bytes(bytes.fromhex(g["params"]))
    vk = F.VerifyingKey.from_bytes(bytes.fromhex(g["vk"]), g["vk_format"])
    co = c_oracle.COracle(bytes.fromhex(g["params"]), 0, bytes.fromhex(g["vk"]), g["vk_format"])
    for e in g["proofs"]:
        inst = [[[int(v, 16) for v in col] for col in e["instances"]]]
        res = orc.verify_proof(params, vk, inst, bytes.fromhex(e["proof"]))
        assert res.status == e["status"] and [hex(c) for c in res.challenges] == e["challenges"]
        st, chal, lr = co.verify(bytes.fromhex(e["proof"]), inst[0])
        assert st == e["status"] and lr.hex() == e["accum"]
    co.close()


@pytest.mark.parametrize("hk", ["blake2b", "keccak"])
def test_honest_lookup_shuffle_circuit(built, hk):
    """lookup.rs:159-271, shuffle.rs:148-225, a rotated gate query and a permutation that includes a fixed column."""
    rng = random.Random(("honest-lk", hk).__repr__())
    s = rng.randrange(1, bn.R)
    circ = hp.lookup_shuffle_circuit(6, 16)
    params, vk, pk = hp.keygen(circ, s)
    co = c_oracle.COracle(params.to_bytes(1), 1, vk.to_bytes(1), 1)
    adv, ins = hp.lookup_shuffle_assignment(circ, 16, rng)
    proof = hp.prove(params, vk, pk, s, adv, ins, rng, hk)
    res = orc.verify_proof(params, vk, [ins], proof, "shplonk", hk)
    assert res.status == orc.OK
    st, chal, lr = co.verify(proof, ins, "shplonk", hk)
    assert (st, chal, lr) == (0, res.challenges, enc_point(res.L) + enc_point(res.R))
    for cheat in ("lookup", "shuffle", "gate", "copy"):
        adv, ins = hp.lookup_shuffle_assignment(circ, 16, rng, cheat)
        bad = hp.prove(params, vk, pk, s, adv, ins, rng, hk, expect_honest=False)
        assert orc.verify_proof(params, vk, [ins], bad, "shplonk", hk).status == orc.CONSTRAINT_SYSTEM_FAILURE, cheat
        assert co.verify(bad, ins, "shplonk", hk)[0] == orc.CONSTRAINT_SYSTEM_FAILURE, cheat
    co.close()


def _host_stage_status(params, vk, m, proof, instances, hk="blake2b"):
    """status of the host build of the CUDA stages (plan compiled for m circuit instances per proof)"""
    import ctypes

    lib = ctypes.CDLL(os.path.join(HERE, "hostlib", "libstage.so"))
    lib.s_err.restype = ctypes.c_char_p
    pb, vb = params.to_bytes(), vk.to_bytes(F.RAW_BYTES)
    assert lib.s_build_m(pb, len(pb), 0, vb, len(vb), F.RAW_BYTES, 0, 0 if hk == "blake2b" else 1, m) == 0, lib.s_err()
    info = (ctypes.c_uint32 * 8)()
    lib.s_info(info)
    _k, P, _S, C, _plen, _nic, nsh, nmo = list(info)
    cols = [col for inst in instances for col in inst]
    ib = b"".join(bn.fr_to_repr(v) for col in cols for v in col)
    cl = (ctypes.c_uint32 * max(1, len(cols)))(*[len(c) for c in cols])
    ch = (ctypes.c_uint8 * (32 * C))(); rt = (ctypes.c_uint8 * (32 * P))(); sh = (ctypes.c_uint8 * (32 * nsh))()
    lf = (ctypes.c_uint8 * (32 * nmo))(); LR = (ctypes.c_uint8 * 128)(); ok = ctypes.c_int(0)
    return lib.s_verify_one(proof, len(proof), ib, sum(len(c) for c in cols), cl, len(cols), ch, rt, sh, lf, LR, ctypes.byref(ok))


def test_honest_multi_instance_proofs(built):
    """ONE proof for several circuit instances (`instances.len() = m`, lib.rs:63,92,117,134) from real witnesses: accepted by
    the Python oracle, the C oracle and the host build of the CUDA stages; rejected when only ONE instance cheats, when a public
    input of one instance is wrong, and when the instances' public inputs are swapped (per-instance index plumbing)."""
    rng = random.Random("honest-multi")
    s = rng.randrange(1, bn.R)
    # vector_mul, 3 instances with different witnesses
    params, vk, pk = hp.keygen_vm(6, s, 4)
    co = c_oracle.COracle(params.to_bytes(1), 1, vk.to_bytes(1), 1)
    m = 3
    wit = [([rng.randrange(bn.R) for _ in range(4)], [rng.randrange(bn.R) for _ in range(4)]) for _ in range(m)]
    asg = [hp.vm_assignment(l, r) for l, r in wit]
    proof = hp.prove_multi(params, vk, pk, s, [a for a, _ in asg], [i for _, i in asg], rng)
    insts = [i for _, i in asg]
    res = orc.verify_proof(params, vk, insts, proof)
    assert res.status == orc.OK
    assert co.verify_multi(proof, insts) == (0, res.challenges, enc_point(res.L) + enc_point(res.R))
    assert _host_stage_status(params, vk, m, proof, insts) == 0
    swapped = [insts[1], insts[0], insts[2]]
    wrong = [[list(c) for c in i] for i in insts]
    wrong[2][0][1] = (wrong[2][0][1] + 1) % bn.R
    for bad_insts in (swapped, wrong):
        assert orc.verify_proof(params, vk, bad_insts, proof).status == orc.CONSTRAINT_SYSTEM_FAILURE
        assert co.verify_multi(proof, bad_insts)[0] == orc.CONSTRAINT_SYSTEM_FAILURE
        assert _host_stage_status(params, vk, m, proof, bad_insts) == orc.CONSTRAINT_SYSTEM_FAILURE
    asg_bad = list(asg)
    asg_bad[1] = hp.vm_assignment(*wit[1], cheat_row=2)  # only instance 1 violates the gate
    bad = hp.prove_multi(params, vk, pk, s, [a for a, _ in asg_bad], [i for _, i in asg_bad], rng, expect_honest=False)
    bad_insts = [i for _, i in asg_bad]
    assert orc.verify_proof(params, vk, bad_insts, bad).status == orc.CONSTRAINT_SYSTEM_FAILURE
    assert co.verify_multi(bad, bad_insts)[0] == orc.CONSTRAINT_SYSTEM_FAILURE
    assert _host_stage_status(params, vk, m, bad, bad_insts) == orc.CONSTRAINT_SYSTEM_FAILURE
    co.close()
    # lookup + shuffle + rotated gate + permutation with a fixed column, 2 instances
    circ = hp.lookup_shuffle_circuit(6, 16)
    params, vk, pk = hp.keygen(circ, s)
    co = c_oracle.COracle(params.to_bytes(1), 1, vk.to_bytes(1), 1)
    a0, i0 = hp.lookup_shuffle_assignment(circ, 16, rng)
    a1, i1 = hp.lookup_shuffle_assignment(circ, 16, rng)
    proof = hp.prove_multi(params, vk, pk, s, [a0, a1], [i0, i1], rng)
    res = orc.verify_proof(params, vk, [i0, i1], proof)
    assert res.status == orc.OK
    assert co.verify_multi(proof, [i0, i1]) == (0, res.challenges, enc_point(res.L) + enc_point(res.R))
    assert _host_stage_status(params, vk, 2, proof, [i0, i1]) == 0
    for cheat in ("lookup", "shuffle", "gate", "copy"):  # the SECOND instance cheats
        a1b, i1b = hp.lookup_shuffle_assignment(circ, 16, rng, cheat)
        bad = hp.prove_multi(params, vk, pk, s, [a0, a1b], [i0, i1b], rng, expect_honest=False)
        assert orc.verify_proof(params, vk, [i0, i1b], bad).status == orc.CONSTRAINT_SYSTEM_FAILURE, cheat
        assert co.verify_multi(bad, [i0, i1b])[0] == orc.CONSTRAINT_SYSTEM_FAILURE, cheat
        assert _host_stage_status(params, vk, 2, bad, [i0, i1b]) == orc.CONSTRAINT_SYSTEM_FAILURE, cheat
    co.close()


def test_honest_proofs_with_the_gwc_opening(built):
    """The honest prover with the GWC multi-open argument (gwc.rs:54-163): true evaluations, witnesses per distinct
    point.  Accepted by the Python oracle, the C oracle and the host build of the CUDA stages (GWC plan); cheats rejected."""
    import ctypes

    rng = random.Random("honest-gwc")
    s = rng.randrange(1, bn.R)
    lib = ctypes.CDLL(os.path.join(HERE, "hostlib", "libstage.so"))
    lib.s_err.restype = ctypes.c_char_p

    def host_status(params, vk, m, proof, insts):
        pb, vb = params.to_bytes(), vk.to_bytes(F.RAW_BYTES)
        assert lib.s_build_m(pb, len(pb), 0, vb, len(vb), F.RAW_BYTES, 1, 0, m) == 0, lib.s_err()
        info = (ctypes.c_uint32 * 8)()
        lib.s_info(info)
        _k, P, _S, C, plen, _nic, nsh, nmo = list(info)
        assert plen == len(proof)
        cols = [col for inst in insts for col in inst]
        ib = b"".join(bn.fr_to_repr(v) for col in cols for v in col)
        cl = (ctypes.c_uint32 * max(1, len(cols)))(*[len(c) for c in cols])
        bufs = [(ctypes.c_uint8 * (32 * k_))() for k_ in (C, P, nsh, nmo)]
        LR = (ctypes.c_uint8 * 128)(); ok = ctypes.c_int(0)
        return lib.s_verify_one(proof, len(proof), ib, sum(len(c) for c in cols), cl, len(cols), *bufs, LR, ctypes.byref(ok))

    # vector_mul: one instance, and two instances in one proof
    params, vk, pk = hp.keygen_vm(6, s, 4)
    co = c_oracle.COracle(params.to_bytes(1), 1, vk.to_bytes(1), 1)
    wit = [([rng.randrange(bn.R) for _ in range(4)], [rng.randrange(bn.R) for _ in range(4)]) for _ in range(2)]
    asg = [hp.vm_assignment(l, r) for l, r in wit]
    for m in (1, 2):
        proof = hp.prove_multi(params, vk, pk, s, [a for a, _ in asg[:m]], [i for _, i in asg[:m]], rng, multiopen="gwc")
        insts = [i for _, i in asg[:m]]
        res = orc.verify_proof(params, vk, insts, proof, "gwc")
        assert res.status == orc.OK
        assert co.verify_multi(proof, insts, "gwc") == (0, res.challenges, enc_point(res.L) + enc_point(res.R))
        assert host_status(params, vk, m, proof, insts) == 0
        wrong = [[list(c) for c in i] for i in insts]
        wrong[-1][0][0] = (wrong[-1][0][0] + 1) % bn.R
        assert orc.verify_proof(params, vk, wrong, proof, "gwc").status == orc.CONSTRAINT_SYSTEM_FAILURE
        assert co.verify_multi(proof, wrong, "gwc")[0] == orc.CONSTRAINT_SYSTEM_FAILURE
        assert host_status(params, vk, m, proof, wrong) == orc.CONSTRAINT_SYSTEM_FAILURE
    bad_a, bad_i = hp.vm_assignment(*wit[0], cheat_row=3)
    bad = hp.prove(params, vk, pk, s, bad_a, bad_i, rng, expect_honest=False, multiopen="gwc")
    assert orc.verify_proof(params, vk, [bad_i], bad, "gwc").status == orc.CONSTRAINT_SYSTEM_FAILURE
    assert co.verify_multi(bad, [bad_i], "gwc")[0] == orc.CONSTRAINT_SYSTEM_FAILURE
    co.close()
    # lookup + shuffle + rotated gate circuit
    circ = hp.lookup_shuffle_circuit(6, 16)
    params, vk, pk = hp.keygen(circ, s)
    co = c_oracle.COracle(params.to_bytes(1), 1, vk.to_bytes(1), 1)
    adv, ins = hp.lookup_shuffle_assignment(circ, 16, rng)
    proof = hp.prove(params, vk, pk, s, adv, ins, rng, multiopen="gwc")
    res = orc.verify_proof(params, vk, [ins], proof, "gwc")
    assert res.status == orc.OK and co.verify_multi(proof, [ins], "gwc")[0] == 0 and host_status(params, vk, 1, proof, [ins]) == 0
    for cheat in ("lookup", "shuffle", "gate", "copy"):
        adv, ins = hp.lookup_shuffle_assignment(circ, 16, rng, cheat)
        bad = hp.prove(params, vk, pk, s, adv, ins, rng, expect_honest=False, multiopen="gwc")
        assert orc.verify_proof(params, vk, [ins], bad, "gwc").status == orc.CONSTRAINT_SYSTEM_FAILURE, cheat
        assert host_status(params, vk, 1, bad, [ins]) == orc.CONSTRAINT_SYSTEM_FAILURE, cheat
    co.close()


@pytest.mark.parametrize("mo", ["shplonk", "gwc"])
def test_honest_two_phase_circuit_with_a_user_challenge(built, mo):
    """Advice columns in two phases and a user challenge squeezed between them (lib.rs:91-109; challenge variables in gate
    polynomials, vk.rs:490-500): the phase-1 column is assigned from the challenge.  Honest proofs (one and two circuit
    instances) are accepted by the Python oracle, the C oracle and the host build of the CUDA stages; a phase-1 cheat and a
    wrong public input are rejected."""
    import ctypes

    rng = random.Random("honest-phases-" + mo)
    s = rng.randrange(1, bn.R)
    circ = hp.two_phase_circuit(5, 6)
    params, vk, pk = hp.keygen(circ, s)
    co = c_oracle.COracle(params.to_bytes(1), 1, vk.to_bytes(1), 1)
    lib = ctypes.CDLL(os.path.join(HERE, "hostlib", "libstage.so"))
    lib.s_err.restype = ctypes.c_char_p

    def host_status(m, proof, insts):
        pb, vb = params.to_bytes(), vk.to_bytes(F.RAW_BYTES)
        assert lib.s_build_m(pb, len(pb), 0, vb, len(vb), F.RAW_BYTES, 0 if mo == "shplonk" else 1, 0, m) == 0, lib.s_err()
        info = (ctypes.c_uint32 * 8)()
        lib.s_info(info)
        _k, P, _S, C, plen, _nic, nsh, nmo = list(info)
        assert plen == len(proof)
        cols = [col for inst in insts for col in inst]
        ib = b"".join(bn.fr_to_repr(v) for col in cols for v in col)
        cl = (ctypes.c_uint32 * max(1, len(cols)))(*[len(c) for c in cols])
        bufs = [(ctypes.c_uint8 * (32 * k_))() for k_ in (C, P, nsh, nmo)]
        LR = (ctypes.c_uint8 * 128)(); ok = ctypes.c_int(0)
        return lib.s_verify_one(proof, len(proof), ib, sum(len(c) for c in cols), cl, len(cols), *bufs, LR, ctypes.byref(ok))

    asg = [hp.two_phase_assignment([rng.randrange(bn.R) for _ in range(6)]) for _ in range(2)]
    for m in (1, 2):
        advs, insts = [a for a, _ in asg[:m]], [i for _, i in asg[:m]]
        proof = hp.prove_multi(params, vk, pk, s, advs, insts, rng, multiopen=mo)
        res = orc.verify_proof(params, vk, insts, proof, mo)
        assert res.status == orc.OK and len(res.trace["user_challenges"]) == 1
        assert co.verify_multi(proof, insts, mo) == (0, res.challenges, enc_point(res.L) + enc_point(res.R))
        assert host_status(m, proof, insts) == 0
        wrong = [[list(c) for c in i] for i in insts]
        wrong[-1][0][2] = (wrong[-1][0][2] + 1) % bn.R
        assert orc.verify_proof(params, vk, wrong, proof, mo).status == orc.CONSTRAINT_SYSTEM_FAILURE
        assert co.verify_multi(proof, wrong, mo)[0] == orc.CONSTRAINT_SYSTEM_FAILURE
        assert host_status(m, proof, wrong) == orc.CONSTRAINT_SYSTEM_FAILURE
    cheat_adv, cheat_inst = hp.two_phase_assignment(asg[0][1][0], cheat=True)
    bad = hp.prove_multi(params, vk, pk, s, [cheat_adv], [cheat_inst], rng, expect_honest=False, multiopen=mo)
    assert orc.verify_proof(params, vk, [cheat_inst], bad, mo).status == orc.CONSTRAINT_SYSTEM_FAILURE
    assert co.verify_multi(bad, [cheat_inst], mo)[0] == orc.CONSTRAINT_SYSTEM_FAILURE
    assert host_status(1, bad, [cheat_inst]) == orc.CONSTRAINT_SYSTEM_FAILURE
    co.close()
