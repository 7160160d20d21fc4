"""Honest mini keygen + prover for the vector_mul circuit (TEST INFRASTRUCTURE ONLY).

The trapdoor simulator (prover_sim.py) manufactures proofs that are self-consistent with the restated
verifier equations, so it cannot reveal a mis-transcribed gate / permutation / vanishing formula.  This
module closes that gap for the `tests/vector_mul.rs` circuit: it builds REAL polynomials from a REAL
witness following the definition of the Halo2 protocol (PLONKish arithmetisation, permutation argument
with delta-cosets and chained grand products, vanishing argument with quotient pieces, blinding rows),
derives every evaluation from those polynomials, and opens them with SHPLONK.  A proof produced here
only verifies if the verifier's expressions vanish on the whole domain for an honest witness, i.e. if
N(X) = sum_i y^i expr_i(X) is divisible by X^n - 1 (asserted below) -- which fails for a wrong delta
power, a wrong l_last / l_blind row, a wrong rotation of the chained product, or a wrong folding order.

Circuit (reference halo2_verifier/tests/vector_mul.rs:88-160): advice a0, a1, a2, instance i0, selector
s_mul (one fixed column), gate s_mul * (a0 * a1 - a2), equality enabled on i0, a0, a1, a2 (permutation
column order: instance first).  Witness: rows 0..m-1 hold lhs, rhs, out = lhs * rhs with s_mul = 1;
out of row j is copy-constrained to instance row j (expose_public).

Commitments use the known SRS secret as a shortcut: commit(p) = [p(s)] G, which is the same group
element as the MSM of p's coefficients with the SRS powers [s^i] G (the k = 8 fixture's secret is
public); for a seeded SRS the secret is ours anyway.
"""
import random

import bn254 as bn
from bn254 import R
from formats import COL_INSTANCE, ConstraintSystem, ParamsKZG, VerifyingKey
from transcript import TranscriptWrite
from verifier import Query, _shplonk_sets


# ---------------------------------------------------------------- domain helpers
def _root(k):
    return pow(bn.FR_ROOT_OF_UNITY, 1 << (bn.FR_S - k), R)


def _ntt(vals, omega):
    """In-order radix-2 NTT: out[i] = sum_j vals[j] * omega^(i j)."""
    n = len(vals)
    if n == 1:
        return list(vals)
    even = _ntt(vals[0::2], omega * omega % R)
    odd = _ntt(vals[1::2], omega * omega % R)
    out = [0] * n
    w = 1
    for i in range(n // 2):
        t = w * odd[i] % R
        out[i] = (even[i] + t) % R
        out[i + n // 2] = (even[i] - t) % R
        w = w * omega % R
    return out


def _intt(evals, omega):
    n = len(evals)
    ninv = bn.fr_inv(n)
    return [v * ninv % R for v in _ntt(evals, bn.fr_inv(omega))]


def _eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc


class Domain:
    def __init__(self, k):
        self.k, self.n = k, 1 << k
        self.omega = _root(k)
        self.ext_k = k + 2  # 4n points: enough for constraint degree 3
        self.ext_omega = _root(self.ext_k)
        self.shift = bn.FR_GENERATOR  # coset generator: keeps X^n - 1 away from zero

    def lagrange_to_coeff(self, vals):
        return _intt(vals, self.omega)

    def coeff_to_ext(self, coeffs):
        m = 1 << self.ext_k
        c = list(coeffs) + [0] * (m - len(coeffs))
        g = 1
        for i in range(m):
            c[i] = c[i] * g % R
            g = g * self.shift % R
        return _ntt(c, self.ext_omega)

    def ext_to_coeff(self, evals):
        c = _intt(evals, self.ext_omega)
        ginv = bn.fr_inv(self.shift)
        g = 1
        for i in range(len(c)):
            c[i] = c[i] * g % R
            g = g * ginv % R
        return c

    def rotate_ext(self, evals, rot):
        """evaluations of p(omega^rot X) on the extended coset from those of p(X)"""
        m = len(evals)
        step = (rot * (m // self.n)) % m
        return [evals[(i + step) % m] for i in range(m)]


# ---------------------------------------------------------------- keygen
def keygen_vm(k, s, rows, transcript_repr=0x1234567):
    """Returns (params, vk, pk) for the vector_mul circuit with `rows` multiplications."""
    dom = Domain(k)
    n = dom.n
    cs = ConstraintSystem()
    cs.num_fixed_columns, cs.num_advice_columns, cs.num_instance_columns = 1, 3, 1
    cs.num_selectors, cs.num_challenges = 1, 0
    cs.advice_column_phase = [0, 0, 0]
    cs.num_advice_queries = [1, 1, 1]
    cs.advice_queries = [(0, 0, 0), (1, 0, 0), (2, 0, 0)]
    cs.instance_queries = [(0, 0)]
    cs.fixed_queries = [(0, 0)]
    cs.permutation_columns = [(0, COL_INSTANCE), (0, 0), (1, 0), (2, 0)]
    cs.gates = [(5, [(0, [(0, 1), (1, 1), (3, 1)]), (1, [(2, 1), (3, 1)])])]  # s * a0 * a1 - s * a2
    cs.coeff_vals = [1, R - 1]
    cs_degree = 3
    bf = cs.blinding_factors()
    usable = n - (bf + 1)
    assert rows <= usable
    sel = [1 if j < rows else 0 for j in range(n)]
    # permutation: identity, except that (a2, row j) and (instance, row j) are swapped for j < rows
    # columns in permutation order: 0 = instance, 1 = a0, 2 = a1, 3 = a2
    mapping = {(c, j): (c, j) for c in range(4) for j in range(n)}
    for j in range(rows):
        mapping[(3, j)], mapping[(0, j)] = (0, j), (3, j)
    delta_pow = [pow(bn.FR_DELTA, c, R) for c in range(4)]
    omega_pow = [pow(dom.omega, j, R) for j in range(n)]
    sigma = [[delta_pow[mapping[(c, j)][0]] * omega_pow[mapping[(c, j)][1]] % R for j in range(n)] for c in range(4)]
    sel_c = dom.lagrange_to_coeff(sel)
    sigma_c = [dom.lagrange_to_coeff(col) for col in sigma]
    commit = lambda coeffs: bn.g1_mul_gen(_eval(coeffs, s))
    sel_bits = bytearray((n + 7) // 8)
    for j in range(n):
        if sel[j]:
            sel_bits[j // 8] |= 1 << (j % 8)
    vk = VerifyingKey(k=k, fixed_commitments=[commit(sel_c)], cs_degree=cs_degree, cs=cs,
                      permutation_commitments=[commit(c) for c in sigma_c], selectors=[bytes(sel_bits)],
                      transcript_repr=transcript_repr)
    params = ParamsKZG(k, bn.G1_GEN, bn.G2_GEN, bn.g2_mul(bn.G2_GEN, s % R))
    pk = {"dom": dom, "sel": sel, "sel_c": sel_c, "sigma": sigma, "sigma_c": sigma_c, "rows": rows, "bf": bf, "usable": usable}
    return params, vk, pk


# ---------------------------------------------------------------- prover
def prove_vm(params, vk, pk, s, lhs, rhs, rng, hash_kind="blake2b", cheat_row=None):
    """Returns (proof bytes, instances [[column]]) for out_j = lhs_j * rhs_j.  `cheat_row`: put a wrong
    product into that row of a2 AND of the public input (the copy constraint still holds, the gate does
    not): the quotient is then no polynomial, so the returned proof must be rejected."""
    dom, n, bf, usable, rows = pk["dom"], pk["dom"].n, pk["bf"], pk["usable"], pk["rows"]
    assert len(lhs) == len(rhs) == rows
    out = [a * b % R for a, b in zip(lhs, rhs)]
    if cheat_row is not None:
        out[cheat_row] = (out[cheat_row] + 1) % R
    blind = lambda col: col[:usable] + [rng.randrange(R) for _ in range(n - usable)]
    pad = lambda v: list(v) + [0] * (usable - len(v))
    a = [blind(pad(lhs)), blind(pad(rhs)), blind(pad(out))]
    inst = pad(out) + [0] * (n - usable)  # instance polynomial: public inputs then zeros (lib.rs:204-217 evaluates exactly this)
    cols = [inst] + a  # permutation order
    commit = lambda coeffs: bn.g1_mul_gen(_eval(coeffs, s))
    a_c = [dom.lagrange_to_coeff(col) for col in a]
    inst_c = dom.lagrange_to_coeff(inst)
    tr = TranscriptWrite(hash_kind)
    tr.common_scalar(vk.transcript_repr)
    for v in out:
        tr.common_scalar(v)
    for c in a_c:
        tr.write_point(commit(c))
    theta = tr.squeeze_challenge()  # noqa: F841  (no lookups in this circuit)
    beta = tr.squeeze_challenge()
    gamma = tr.squeeze_challenge()
    # permutation argument (plonk/permutation.rs, prover side of the upstream protocol): one column per set
    # (chunk = cs_degree - 2 = 1); z_i(omega^(j+1)) = z_i(omega^j) * (v + beta delta^i omega^j + gamma) / (v + beta sigma_i(omega^j) + gamma)
    # over the usable rows, z_0(1) = 1, z_i(1) = z_{i-1}(omega^usable), blinding rows random
    omega_pow = [pow(dom.omega, j, R) for j in range(n)]
    z, start = [], 1
    for i in range(4):
        zi = [0] * n
        zi[0] = start
        dp = pow(bn.FR_DELTA, i, R)
        dens = bn.batch_invert_skip_zero([(cols[i][j] + beta * pk["sigma"][i][j] + gamma) % R for j in range(usable)], R)
        for j in range(usable):
            num = (cols[i][j] + beta * dp % R * omega_pow[j] + gamma) % R
            zi[j + 1] = zi[j] * num % R * dens[j] % R
        start = zi[usable]
        for j in range(usable + 1, n):
            zi[j] = rng.randrange(R)
        z.append(zi)
    assert start == 1, "grand product of an honest permutation must close to 1"
    z_c = [dom.lagrange_to_coeff(zi) for zi in z]
    for c in z_c:
        tr.write_point(commit(c))
    rand_c = [rng.randrange(R) for _ in range(n)]  # vanishing.rs:49-57: random polynomial
    tr.write_point(commit(rand_c))
    y = tr.squeeze_challenge()

    # quotient: N(X) = fold_y(expressions)(X) on the extended coset, h = N / (X^n - 1)
    E = dom.coeff_to_ext
    a_e, inst_e, sel_e = [E(c) for c in a_c], E(inst_c), E(pk["sel_c"])
    sig_e, z_e = [E(c) for c in pk["sigma_c"]], [E(c) for c in z_c]
    col_e = [inst_e] + a_e
    lag = lambda row: E(dom.lagrange_to_coeff([1 if j == row else 0 for j in range(n)]))
    l0_e, llast_e = lag(0), lag(usable)
    lblind_e = E(dom.lagrange_to_coeff([1 if j > usable else 0 for j in range(n)]))
    m = len(sel_e)
    x_e = [dom.shift * pow(dom.ext_omega, i, R) % R for i in range(m)]
    z_next = [dom.rotate_ext(ze, 1) for ze in z_e]
    z_last = [dom.rotate_ext(ze, -(bf + 1)) for ze in z_e]
    exprs = [[sel_e[i] * ((a_e[0][i] * a_e[1][i] - a_e[2][i]) % R) % R for i in range(m)]]  # gate, vk.rs:478-512
    exprs.append([l0_e[i] * ((1 - z_e[0][i]) % R) % R for i in range(m)])  # permutation.rs:189-288
    exprs.append([llast_e[i] * ((z_e[3][i] * z_e[3][i] - z_e[3][i]) % R) % R for i in range(m)])
    for t in range(1, 4):
        exprs.append([l0_e[i] * ((z_e[t][i] - z_last[t - 1][i]) % R) % R for i in range(m)])
    for t in range(4):
        dp = pow(bn.FR_DELTA, t, R)
        exprs.append([(1 - llast_e[i] - lblind_e[i]) % R
                      * ((z_next[t][i] * ((col_e[t][i] + beta * sig_e[t][i] + gamma) % R)
                          - z_e[t][i] * ((col_e[t][i] + beta * dp % R * x_e[i] + gamma) % R)) % R) % R for i in range(m)])
    N = [0] * m
    for e in exprs:
        N = [(N[i] * y + e[i]) % R for i in range(m)]
    vinv = bn.batch_invert_skip_zero([(pow(xe, n, R) - 1) % R for xe in x_e], R)
    h_c = dom.ext_to_coeff([N[i] * vinv[i] % R for i in range(m)])
    honest = all(c == 0 for c in h_c[2 * n:])  # deg h <= 2n - 3 iff N vanishes on the domain (X^n - 1 | N)
    if cheat_row is None:
        assert honest, "constraint expressions do not vanish on the domain for an honest witness"
    else:
        assert not honest
    h_pieces = [h_c[0:n], h_c[n:2 * n]]
    for piece in h_pieces:
        tr.write_point(commit(piece))
    x = tr.squeeze_challenge()
    xn = pow(x, n, R)
    rot = lambda r: x * pow(dom.omega, r % n, R) % R
    ev = lambda coeffs, r=0: _eval(coeffs, rot(r))
    for c in a_c:
        tr.write_scalar(ev(c))
    tr.write_scalar(ev(pk["sel_c"]))
    tr.write_scalar(ev(rand_c))
    for c in pk["sigma_c"]:
        tr.write_scalar(ev(c))
    for t in range(4):
        tr.write_scalar(ev(z_c[t]))
        tr.write_scalar(ev(z_c[t], 1))
        if t != 3:
            tr.write_scalar(ev(z_c[t], -(bf + 1)))
    # ---- SHPLONK opening of every query (shplonk.rs): polynomials by commitment identity
    hmsm_c = [(h_pieces[0][i] + xn * h_pieces[1][i]) % R for i in range(n)]
    polys, queries = {}, []

    def add_q(ident, coeffs, r):
        polys[ident] = coeffs
        queries.append(Query(ident, rot(r), _eval(coeffs, rot(r)), None, r))

    for ci in range(3):
        add_q(("proof", ci), a_c[ci], 0)
    for t in range(4):
        add_q(("proof", 3 + t), z_c[t], 0)
        add_q(("proof", 3 + t), z_c[t], 1)
    for t in reversed(range(3)):
        add_q(("proof", 3 + t), z_c[t], -(bf + 1))
    add_q(("fixed", 0), pk["sel_c"], 0)
    for t in range(4):
        add_q(("sigma", t), pk["sigma_c"][t], 0)
    add_q(("hmsm",), hmsm_c, 0)
    add_q(("proof", 7), rand_c, 0)
    rotation_sets, super_points = _shplonk_sets(queries)
    yy = tr.squeeze_challenge()
    v = tr.squeeze_challenge()
    van = lambda roots, at: _prod((at - p) % R for p in roots)
    # h(X) = sum_i v^i (P_i(X) - R_i(X)) / Z_{T_i}(X) evaluated at s (commit = [h(s)] G)
    sets = []
    h_at_s, pow_v = 0, 1
    for points, commitments in rotation_sets:
        P_s, pow_y = 0, 1
        evals_comb = [0] * len(points)
        for ident, _c, evals in commitments:
            P_s = (P_s + pow_y * _eval(polys[ident], s)) % R
            evals_comb = [(e0 + pow_y * e1) % R for e0, e1 in zip(evals_comb, evals)]
            pow_y = pow_y * yy % R
        R_s = _interp_at(points, evals_comb, s)
        h_at_s = (h_at_s + pow_v * ((P_s - R_s) % R) % R * bn.fr_inv(van(points, s))) % R
        sets.append((points, P_s, evals_comb, pow_v))
        pow_v = pow_v * v % R
    tr.write_point(bn.g1_mul_gen(h_at_s))
    u = tr.squeeze_challenge()
    # L(X) = sum_i v^i Z_{T \\ T_i}(u) (P_i(X) - R_i(u)) - Z_T(u) h(X) vanishes at u; the verifier works with L / Z_{T \\ T_0}(u)
    zd0_inv = bn.fr_inv(van([p for p in super_points if p not in sets[0][0]], u))
    L_s = 0
    for points, P_s, evals_comb, pv in sets:
        zd = van([p for p in super_points if p not in points], u) * zd0_inv % R
        L_s = (L_s + pv * zd % R * ((P_s - _interp_at(points, evals_comb, u)) % R)) % R
    L_s = (L_s - van(sets[0][0], u) * h_at_s) % R
    tr.write_point(bn.g1_mul_gen(L_s * bn.fr_inv((s - u) % R) % R))
    return bytes(tr.out), [[out]]


def _prod(it):
    acc = 1
    for v in it:
        acc = acc * v % R
    return acc


def _interp_at(points, evals, at):
    """Lagrange interpolant through (points, evals) evaluated at `at`."""
    total = 0
    for j, (xj, ej) in enumerate(zip(points, evals)):
        num, den = 1, 1
        for k_, xk in enumerate(points):
            if k_ != j:
                num = num * ((at - xk) % R) % R
                den = den * ((xj - xk) % R) % R
        total = (total + ej * num % R * bn.fr_inv(den)) % R
    return total


def demo(seed=1, k=8, rows=10):
    from verifier import verify_proof

    rng = random.Random(seed)
    s = 0x1C59A59B6CFF4308740943526ADE1D8C09F71B337A67269CC89586BCDD6DFCBA if k == 8 else rng.randrange(1, R)
    params, vk, pk = keygen_vm(k, s, rows)
    lhs = [rng.randrange(R) for _ in range(rows)]
    rhs = [rng.randrange(R) for _ in range(rows)]
    proof, inst = prove_vm(params, vk, pk, s, lhs, rhs, rng)
    return params, vk, proof, inst, verify_proof(params, vk, inst, proof)


if __name__ == "__main__":
    res = demo()[-1]
    print("honest proof status:", res.status, res.error)
