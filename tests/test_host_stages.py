"""CPU-side unit tests of the PRODUCT's device code: the HD headers (field / curve / tower / hash) and
the per-proof stages + plan compiler are compiled for the host (tests/hostlib) and compared with the
oracle.  This is the same source the CUDA kernels run; only the PTX multiplication path and the
kernel plumbing need a GPU (tests/test_gpu_parity.py)."""
import ctypes
import hashlib
import os
import random

import pytest

import bn254 as bn
import formats as F
import prover_sim as sim
import transcript as T
import verifier as orc
from workloads import oracle_scalars, setup

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def prim(built):
    return ctypes.CDLL(os.path.join(HERE, "hostlib", "libprim.so"))


@pytest.fixture(scope="module")
def stage(built):
    lib = ctypes.CDLL(os.path.join(HERE, "hostlib", "libstage.so"))
    lib.s_err.restype = ctypes.c_char_p
    return lib


def L(v):
    return (ctypes.c_uint32 * 8)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)])


def L16(x, y):
    return (ctypes.c_uint32 * 16)(*([(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)] + [(y >> (32 * i)) & 0xFFFFFFFF for i in range(8)]))


def I(a, off=0):
    return sum(int(a[off + i]) << (32 * i) for i in range(8))


M = 1 << 256


def test_montgomery_fields(prim):
    rng = random.Random(3)
    for field, mod in ((0, bn.P), (1, bn.R)):
        for t in range(200):
            x, y = rng.randrange(mod), rng.randrange(mod)
            if t < 4:
                x = [0, 1, mod - 1, mod - 2][t]
            out = (ctypes.c_uint32 * 8)()
            prim.t_field(field, 0, L(x), L(y), out); assert I(out) == x * y * pow(M, -1, mod) % mod
            prim.t_field(field, 1, L(x), L(y), out); assert I(out) == (x + y) % mod
            prim.t_field(field, 2, L(x), L(y), out); assert I(out) == (x - y) % mod
            prim.t_field(field, 5, L(x), L(y), out); assert I(out) == x * M % mod
        for _ in range(5):
            x = rng.randrange(1, mod)
            out = (ctypes.c_uint32 * 8)()
            prim.t_field(field, 3, L(x * M % mod), L(0), out); assert I(out) == pow(x, -1, mod) * M % mod
        # the binary-Euclid Montgomery inverse used by the scalar stage: the same element as Fermat's, on random and edge inputs
        for t in range(300):
            x = [1, 2, mod - 1, mod - 2, (mod + 1) // 2, 1 << 253, (1 << 253) - 1, 3][t] if t < 8 else rng.randrange(1, mod)
            out = (ctypes.c_uint32 * 8)()
            prim.t_field(field, 7, L(x * M % mod), L(0), out); assert I(out) == pow(x, -1, mod) * M % mod, hex(x)
        out = (ctypes.c_uint32 * 8)()
        prim.t_field(field, 7, L(0), L(0), out); assert I(out) == 0
    for _ in range(10):
        d = rng.randbytes(64)
        out = (ctypes.c_uint32 * 8)()
        prim.t_from_uniform(d, out)
        assert I(out) == int.from_bytes(d, "little") % bn.R


def test_decompression_and_group_law(prim):
    rng = random.Random(4)
    xy = (ctypes.c_uint32 * 16)()
    for _ in range(20):
        pt = bn.g1_mul_gen(rng.randrange(1, bn.R))
        enc = bn.g1_to_bytes(pt)
        assert prim.t_decompress(enc, xy) == 1 and (I(xy), I(xy, 8)) == pt
        bad = bytearray(enc); bad[31] |= 0x40
        assert prim.t_decompress(bytes(bad), xy) == 0
    for enc in (bytes(32), (bn.P + 1).to_bytes(32, "little"), (5).to_bytes(32, "little"), (bn.P - 1).to_bytes(32, "little")):
        ok, p = bn.g1_from_bytes(enc)
        assert (prim.t_decompress(enc, xy) == 1) == (ok and p is not None)
    out = (ctypes.c_uint32 * 16)()
    for _ in range(5):
        k, pt = rng.randrange(bn.R), bn.g1_mul_gen(rng.randrange(1, bn.R))
        prim.t_g1_mul(L16(*pt), L(k), out); assert (I(out), I(out, 8)) == bn.g1_mul(pt, k)
        q = bn.g1_mul(pt, 4)  # exceptional cases: 4p + 4p -> doubling path, 4p - 4p -> identity path
        prim.t_g1_add(L16(*pt), L16(*q), 0, out); assert (I(out), I(out, 8)) == bn.g1_mul(pt, 6)
        prim.t_g1_add(L16(*pt), L16(*q), 1, out); assert (I(out), I(out, 8)) == bn.g1_neg(bn.g1_mul(pt, 2))


def test_glv_decomposition_and_split_multiplication(prim):
    """glv.cuh: k = k1 + k2 * lambda (mod r) with both halves below 2^128, and [k]P rebuilt from the two half-length parts
    (the attribution kernels' scalar multiplication) equals the oracle's [k]P."""
    lam = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd
    beta = 0x59e26bcea0d48bacd4f263f1acdb5c4f5763473177fffffe
    assert pow(lam, 3, bn.R) == 1 and lam != 1 and pow(beta, 3, bn.P) == 1 and beta != 1
    g = bn.g1_mul_gen(1)
    assert bn.g1_mul(g, lam) == (beta * g[0] % bn.P, g[1])  # phi(P) = (beta x, y) = [lambda] P
    rng = random.Random(9)
    out12 = (ctypes.c_uint32 * 12)()
    ks = [0, 1, 2, bn.R - 1, bn.R - 2, lam, bn.R - lam, (bn.R - 1) // 2] + [rng.randrange(bn.R) for _ in range(3000)] + \
         [rng.randrange(1 << b) for b in range(1, 254)]
    for k in ks:
        prim.t_glv_decompose(L(k), out12)
        k1 = sum(out12[i] << (32 * i) for i in range(5)) * (-1 if out12[5] else 1)
        k2 = sum(out12[6 + i] << (32 * i) for i in range(5)) * (-1 if out12[11] else 1)
        assert (k1 + k2 * lam) % bn.R == k, hex(k)
        assert abs(k1) < 1 << 128 and abs(k2) < 1 << 128, hex(k)
    out = (ctypes.c_uint32 * 16)()
    for k in [1, 2, lam, bn.R - 1, (1 << 64) - 1, 1 << 64, (1 << 128) + 5] + [rng.randrange(bn.R) for _ in range(12)]:
        pt = bn.g1_mul_gen(rng.randrange(1, bn.R))
        prim.t_g1_mul_glv(L16(*pt), L(k), out)
        assert (I(out), I(out, 8)) == bn.g1_mul(pt, k), hex(k)


def test_pairing_value_equals_oracle(prim):
    S = sim.FIXTURE_SRS_SECRET
    sg2, ng2 = bn.g2_mul(bn.G2_GEN, S), bn.g2_neg(bn.G2_GEN)

    def G2L(q):
        (x0, x1), (y0, y1) = q
        return (ctypes.c_uint32 * 32)(*sum([[(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)] for v in (x0, x1, y0, y1)], []))

    a = 987654321
    Lp, Rp, Rbad = bn.g1_mul_gen(a), bn.g1_mul_gen(a * S % bn.R), bn.g1_mul_gen((a * S + 1) % bn.R)
    gt = (ctypes.c_uint32 * 96)()
    assert prim.t_pairing_check(L16(*Lp), L16(*Rp), 0, 0, G2L(sg2), G2L(ng2), gt) == 1
    assert prim.t_pairing_check(L16(*Lp), L16(*Rbad), 0, 0, G2L(sg2), G2L(ng2), gt) == 0
    ref = bn.final_exponentiation(bn.miller_loop([(Lp, sg2), (Rbad, ng2)]))
    assert tuple((I(gt, 16 * i), I(gt, 16 * i + 8)) for i in range(6)) == ref  # same GT element, not just the verdict


def test_hashes(prim):
    rng = random.Random(5)
    for n in (0, 1, 31, 32, 33, 119, 120, 121, 127, 128, 129, 135, 136, 137, 255, 256, 257, 1000):
        d = rng.randbytes(n)
        out = (ctypes.c_uint8 * 64)()
        prim.t_blake2b(d, n, out)
        assert bytes(out) == hashlib.blake2b(d, digest_size=64, person=b"Halo2-Transcript").digest(), n
        o2 = (ctypes.c_uint8 * 32)()
        prim.t_keccak(d, n, 10, o2)
        assert bytes(o2) == T.keccak256(b"Halo2-Transcript" + d + b"\x0a"), n


def _run_case(stage, shape, k, mo, hk, vkfmt, corruptions):
    params, vk, dl, s = setup(shape, k)
    rng = random.Random(f"host-{shape}{k}{mo}{hk}")
    pb, vb = params.to_bytes(), vk.to_bytes(vkfmt)
    rc = stage.s_build(pb, len(pb), 0, vb, len(vb), vkfmt, 0 if mo == "shplonk" else 1, 0 if hk == "blake2b" else 1)
    assert rc == 0, stage.s_err()
    info = (ctypes.c_uint32 * 8)()
    stage.s_info(info)
    _k, P, _S, C, plen, _nic, nsh, nmo = list(info)
    inst = sim.random_instances(vk, rng, 10)
    proof = sim.simulate_proof(params, vk, dl, s, inst, rng, mo, hk)
    assert plen == len(proof)

    def call(proof, inst, ncols=-1):
        ib = b"".join(bn.fr_to_repr(v) for col in inst[0] for v in col)
        tot = sum(len(c) for c in inst[0])
        ch = (ctypes.c_uint8 * (32 * C))(); rt = (ctypes.c_uint8 * (32 * P))(); sh = (ctypes.c_uint8 * (32 * nsh))()
        lf = (ctypes.c_uint8 * (32 * nmo))(); LR = (ctypes.c_uint8 * 128)(); ok = ctypes.c_int(0)
        st = stage.s_verify_one(proof, len(proof), ib, tot, None, ncols, ch, rt, sh, lf, LR, ctypes.byref(ok))
        g = lambda a, i: int.from_bytes(bytes(a[32 * i: 32 * i + 32]), "little")
        return st, [g(ch, i) for i in range(C)], [g(rt, i) for i in range(P)] + [g(sh, i) for i in range(nsh)] + [g(lf, i) for i in range(nmo)], [g(LR, i) for i in range(4)]

    st, ch, scalars, LR = call(proof, inst)
    res = orc.verify_proof(params, vk, inst, proof, mo, hk)
    assert st == res.status == 0 and ch == res.challenges
    assert scalars == oracle_scalars(vk, res, P, nmo)
    assert (LR[0], LR[1]) == res.L and (LR[2], LR[3]) == res.R
    if corruptions:
        for kind in sim.CORRUPTIONS:
            bad, exp = sim.corrupt(proof, vk, kind, rng, mo)
            assert call(bad, inst)[0] == exp, kind
        if vk.cs.num_instance_columns:
            inst2 = [[list(c) for c in inst[0]]]
            inst2[0][0][0] = (inst2[0][0][0] + 1) % bn.R
            assert call(proof, inst2)[0] == 4
            assert call(proof, [inst[0][:-1]], ncols=len(inst[0]) - 1)[0] == 1
        assert call(proof + b"\x01" * 40, inst)[0] == 0  # trailing bytes are never read by the reference


@pytest.mark.parametrize("shape,k", [("vm", 8), ("vm", 10), ("sh", 8), ("mix", 6)])
@pytest.mark.parametrize("mo", ["shplonk", "gwc"])
@pytest.mark.parametrize("hk", ["blake2b", "keccak"])
def test_stages_match_oracle(stage, shape, k, mo, hk):
    _run_case(stage, shape, k, mo, hk, F.RAW_BYTES if mo == "shplonk" else F.PROCESSED, corruptions=(hk == "blake2b" and shape in ("vm", "mix")))


@pytest.mark.parametrize("shape,k,m", [("vm", 8, 2), ("mix", 6, 3), ("sh", 8, 2)])
@pytest.mark.parametrize("mo", ["shplonk", "gwc"])
def test_stages_multi_instance_transcripts(stage, shape, k, m, mo):
    """Proofs that carry m circuit instances in one transcript (`instances.len() = m`, lib.rs:63,92,117,134): the plan
    compiler repeats the per-instance commitments, evaluations, expressions and queries in the reference's interleaving.
    Statuses, challenges, MSM scalars and accumulators against the Python oracle, plus every corruption class."""
    params, vk, dl, s = setup(shape, k)
    rng = random.Random(f"multi-{shape}{k}{m}{mo}")
    pb, vb = params.to_bytes(), vk.to_bytes(F.RAW_BYTES)
    assert stage.s_build_m(pb, len(pb), 0, vb, len(vb), F.RAW_BYTES, 0 if mo == "shplonk" else 1, 0, m) == 0, stage.s_err()
    info = (ctypes.c_uint32 * 8)()
    stage.s_info(info)
    _k, P, _S, C, plen, nic, nsh, nmo = list(info)
    assert nic == m * vk.cs.num_instance_columns
    insts = [sim.random_instances(vk, rng, 10)[0] for _ in range(m)]
    proof = sim.simulate_proof(params, vk, dl, s, insts, rng, mo, "blake2b")
    assert plen == len(proof) == 32 * len(sim.proof_layout(vk, mo, m)[0])

    def call(proof, insts):
        ib = b"".join(bn.fr_to_repr(v) for inst in insts for col in inst for v in col)  # instance-major, lib.rs:76-82
        tot = sum(len(c) for inst in insts for c in inst)
        ch = (ctypes.c_uint8 * (32 * C))(); rt = (ctypes.c_uint8 * (32 * P))(); sh = (ctypes.c_uint8 * (32 * nsh))()
        lf = (ctypes.c_uint8 * (32 * nmo))(); LR = (ctypes.c_uint8 * 128)(); ok = ctypes.c_int(0)
        st = stage.s_verify_one(proof, len(proof), ib, tot, None, -1, ch, rt, sh, lf, LR, ctypes.byref(ok))
        g = lambda a, i: int.from_bytes(bytes(a[32 * i: 32 * i + 32]), "little")
        return st, [g(ch, i) for i in range(C)], [g(rt, i) for i in range(P)] + [g(sh, i) for i in range(nsh)] + [g(lf, i) for i in range(nmo)], [g(LR, i) for i in range(4)]

    st, ch, scalars, LR = call(proof, insts)
    res = orc.verify_proof(params, vk, insts, proof, mo, "blake2b")
    assert st == res.status == 0 and ch == res.challenges
    assert scalars == oracle_scalars(vk, res, P, nmo)
    assert (LR[0], LR[1]) == res.L and (LR[2], LR[3]) == res.R
    for kind in sim.CORRUPTIONS:
        bad, exp = sim.corrupt(proof, vk, kind, rng, mo, m)
        assert orc.verify_proof(params, vk, insts, bad, mo, "blake2b").status == exp
        assert call(bad, insts)[0] == exp, kind
    if vk.cs.num_instance_columns:  # a wrong public input of the LAST instance
        bad_insts = [[list(c) for c in inst] for inst in insts]
        bad_insts[-1][0][0] = (bad_insts[-1][0][0] + 1) % bn.R
        assert call(proof, bad_insts)[0] == orc.verify_proof(params, vk, bad_insts, proof, mo, "blake2b").status == 4
    # the single-instance plan rejects / differs on the same bytes: the layouts are not interchangeable
    assert stage.s_build(pb, len(pb), 0, vb, len(vb), F.RAW_BYTES, 0 if mo == "shplonk" else 1, 0) == 0
    stage.s_info(info)
    assert list(info)[4] != plen


def test_stages_k18_lookup_heavy(stage):
    _run_case(stage, "k18", 18, "shplonk", "blake2b", F.RAW_BYTES, corruptions=False)


def test_plan_compiler_rejects_malformed_vk(stage):
    params, vk, _dl, _s = setup("vm", 8)
    pb, vb = params.to_bytes(), vk.to_bytes(F.RAW_BYTES)
    assert stage.s_build(pb, len(pb), 0, vb[:-5], len(vb) - 5, 1, 0, 0) != 0 and b"truncated" in stage.s_err()
    assert stage.s_build(pb[:100], 100, 0, vb, len(vb), 1, 0, 0) != 0
    import copy
    vk2 = copy.deepcopy(vk)
    vk2.cs.permutation_columns[1] = (7, 0)  # column without a query at Rotation::cur(): the reference panics at verify time
    vb2 = vk2.to_bytes(F.RAW_BYTES)
    assert stage.s_build(pb, len(pb), 0, vb2, len(vb2), 1, 0, 0) != 0 and b"permutation column" in stage.s_err()
    vk3 = copy.deepcopy(vk)
    vk3.cs.gates[0] = (5, [])
    vb3 = vk3.to_bytes(F.RAW_BYTES)
    assert stage.s_build(pb, len(pb), 0, vb3, len(vb3), 1, 0, 0) != 0 and b"empty polynomial" in stage.s_err()
    p10 = sim.make_params(10, 5).to_bytes()
    assert stage.s_build(p10, len(p10), 0, vb, len(vb), 1, 0, 0) != 0  # params.k != vk.k


def test_vk_lint_reports_the_reference_write_read_asymmetries(stage):
    """lookups with several expression pairs (write: inputs then tables; read: interleaved,
    plonk/lookup.rs:42-61) and query lists longer than the column counts (plonk/vk.rs:243-251 vs 310-322) are reported, a
    VK without them is clean."""
    stage.s_lint.restype = ctypes.c_char_p
    params, vk, _dl, _s = setup("vm", 8)
    pb, vb = params.to_bytes(), vk.to_bytes(F.RAW_BYTES)
    assert stage.s_build(pb, len(pb), 0, vb, len(vb), 1, 0, 0) == 0 and stage.s_lint() == b""
    vb2 = vb + b"\x00" * 12  # what a longer fixed-query list leaves behind after a misaligned but "successful" read
    assert stage.s_build(pb, len(pb), 0, vb2, len(vb2), 1, 0, 0) == 0
    assert b"12 bytes follow transcript_repr" in stage.s_lint() and b"vk.rs:243-251" in stage.s_lint()
    params, vk, _dl, _s = setup("k18", 18)  # 8 lookups with 2 expression pairs each
    pb, vb = params.to_bytes(), vk.to_bytes(F.RAW_BYTES)
    assert stage.s_build(pb, len(pb), 0, vb, len(vb), 1, 0, 0) == 0
    lint = stage.s_lint().decode().splitlines()
    assert len(lint) == 8 and all("expression pairs" in l and "lookup.rs:42-47 vs 58-61" in l for l in lint)
    assert stage.s_build(pb, len(pb), 0, vb[:-40], len(vb) - 40, 1, 0, 0) != 0 and b"VerifyingKey::write" in stage.s_err()


# ---- cooperative Fq12 engine (pairing_cta.cuh), lanes emulated on the host
def F12L(x):
    return (ctypes.c_uint32 * 96)(*sum([[(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)] for c in x for v in c], []))


def F12I(a):
    return tuple((I(a, 16 * i), I(a, 16 * i + 8)) for i in range(6))


def test_e12_engine_matches_oracle_tower(prim):
    rng = random.Random(77)
    rnd12 = lambda: tuple((rng.randrange(bn.P), rng.randrange(bn.P)) for _ in range(6))
    out = (ctypes.c_uint32 * 96)()
    edge = tuple((bn.P - 1, bn.P - 1) for _ in range(6))
    cases = [(rnd12(), rnd12()) for _ in range(12)] + [(edge, edge), (bn.F12_ONE, rnd12()), (rnd12(), tuple((0, 0) for _ in range(6)))]
    for x, y in cases:
        prim.t_e12_op(0, F12L(x), F12L(y), out); assert F12I(out) == bn.f12_mul(x, y)
        prim.t_e12_op(5, F12L(x), F12L(y), out); assert F12I(out) == bn.f12_mul(x, y)  # the 128-thread split of the map (k_pairing_check)
        prim.t_e12_op(1, F12L(x), F12L(y), out); assert F12I(out) == bn.f12_conj(x)
        prim.t_e12_op(2, F12L(x), F12L(y), out); assert F12I(out) == bn.f12_frob(x)
        prim.t_e12_op(3, F12L(x), F12L(y), out); assert F12I(out) == bn.f12_frob2(x)
    x = rnd12()
    prim.t_e12_op(4, F12L(x), F12L(x), out); assert F12I(out) == bn.f12_pow(x, bn.U)


def test_decomposed_pairing_check(prim):
    """prod_w e(S_w, [2^(c w)] Q) with projective S_w, no inversion: same verdict as the oracle's pairing_check."""
    rng = random.Random(78)
    S = sim.FIXTURE_SRS_SECRET
    sg2, ng2 = bn.g2_mul(bn.G2_GEN, S), bn.g2_neg(bn.G2_GEN)

    def jac(pt):  # random projective representative (X, Y, Z) of an affine point; None -> Z = 0
        if pt is None:
            return [1, 1, 0]
        z = rng.randrange(1, bn.P)
        return [pt[0] * z * z % bn.P, pt[1] * pow(z, 3, bn.P) % bn.P, z]

    def run(pairs):
        n = len(pairs)
        jw = (ctypes.c_uint32 * (24 * n))(*sum([[(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)] for p, _ in pairs for v in jac(p)], []))
        qw = (ctypes.c_uint32 * (32 * n))(*sum([[(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)] for _, q in pairs for c in q for v in c], []))
        return prim.t_pairing_windows(n, jw, qw)

    c, W = 7, 3
    # left = sum 2^(c w) A_w, right = s * left, split into windows on both sides
    a = [rng.randrange(bn.R) for _ in range(W)]
    A = [bn.g1_mul_gen(x) for x in a]
    B = [bn.g1_mul_gen(x * S % bn.R) for x in a]
    pairs = [(A[w], bn.g2_mul(sg2, 1 << (c * w))) for w in range(W)] + [(B[w], bn.g2_mul(ng2, 1 << (c * w))) for w in range(W)]
    assert bn.pairing_check(pairs)
    assert run(pairs) == 1
    bad = list(pairs)
    bad[4] = (bn.g1_mul_gen((a[1] * S + 1) % bn.R), bad[4][1])
    assert not bn.pairing_check(bad)
    assert run(bad) == 0
    # identity window sums are skipped; an all-identity product accepts
    pairs2 = [(A[0], sg2), (None, bn.g2_mul(sg2, 1 << c)), (B[0], ng2), (None, bn.g2_mul(ng2, 1 << c))]
    assert run(pairs2) == 1
    assert run([(None, sg2), (None, ng2)]) == 1
