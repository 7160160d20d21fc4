"""Honest mini keygen + prover for PLONKish circuits in the reference's VK format (TEST INFRASTRUCTURE
ONLY).

The trapdoor simulator (prover_sim.py) manufactures proofs that are self-consistent with the restated
verifier equations, so it cannot reveal a mis-transcribed gate / permutation / lookup / shuffle /
vanishing formula.  This module closes that gap: it builds REAL polynomials from a REAL assignment
following the DEFINITION of the Halo2 protocol -- custom gates over rotated queries, the permutation
argument with delta-cosets and chained grand products, the lookup argument (permuted input / table
columns, theta-compressed expressions, grand product), the shuffle argument, the vanishing argument
with quotient pieces and blinding rows -- derives every evaluation from those polynomials, and opens
them with SHPLONK.  A proof produced here only verifies if the verifier's expressions vanish on the
whole domain for an honest assignment, i.e. if N(X) = sum_i y^i expr_i(X) is divisible by X^n - 1
(asserted below): a wrong delta power, l_last / l_blind row, rotation of the chained product, folding
order or lookup identity breaks that.

Circuits: `vm_circuit` = reference halo2_verifier/tests/vector_mul.rs:88-160 (advice a0, a1, a2,
instance i0, selector s_mul, gate s_mul * (a0 * a1 - a2), out copy-constrained to the public input);
`lookup_shuffle_circuit` = a small circuit with a rotated gate, one lookup into a fixed table, one
shuffle and a permutation, exercising lookup.rs:159-271 and shuffle.rs:148-225.

Commitments use the known SRS secret as a shortcut: commit(p) = [p(s)] G, the same group element as the
MSM of p's coefficients with the SRS powers [s^i] G (the k = 8 fixture's secret is public, SURVEY.md
appendix A; for a seeded SRS the secret is ours anyway).  Single phase, one circuit instance.
"""
import random
from dataclasses import dataclass, field
from typing import List, Tuple

import bn254 as bn
from bn254 import R
from formats import COL_FIXED, COL_INSTANCE, ConstraintSystem, ParamsKZG, VerifyingKey
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
    ninv = bn.fr_inv(len(evals))
    return [v * ninv % R for v in _ntt(evals, bn.fr_inv(omega))]


def _eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc


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


class Domain:
    def __init__(self, k, degree):
        self.k, self.n = k, 1 << k
        self.omega = _root(k)
        self.ext_k = k + max(1, (degree - 1).bit_length())  # 2^ext_k >= degree * n: room for deg(N) < degree * n
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


# ---------------------------------------------------------------- circuits
@dataclass
class Circuit:
    k: int
    cs: ConstraintSystem
    cs_degree: int
    fixed: List[List[int]]  # [fixed column][row], zero-padded to n by keygen
    copies: List[Tuple[Tuple[int, int], Tuple[int, int]]] = field(default_factory=list)  # ((perm column index, row), (perm column index, row))


def vm_circuit(k, rows):
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
    # out (a2, row j) == public input (instance, row j)
    return Circuit(k, cs, 3, [[1] * rows], [((3, j), (0, j)) for j in range(rows)])


def vm_assignment(lhs, rhs, cheat_row=None):
    out = [a * b % R for a, b in zip(lhs, rhs)]
    if cheat_row is not None:
        out[cheat_row] = (out[cheat_row] + 1) % R
    return [list(lhs), list(rhs), list(out)], [list(out)]


def lookup_shuffle_circuit(k, rows):
    """advice a0, a1, a2; fixed q (selector), t, t2 (table and its squares); variables: a0@0 a0@1 a1@0 a2@0 | q t t2.
    gate      q * (a0(wX) - a0(X) - a1(X))               running sum, uses a rotated query
    lookup    theta-compressed (q * a1, a1 * a1)  in  (t, t2)    (constraint degree 1 + 1 + 2 + 1 = cs_degree = 5)
    shuffle   (a2)  is a shuffle of  (a1)
    permutation over a0, a2, t with copy constraints"""
    cs = ConstraintSystem()
    cs.num_fixed_columns, cs.num_advice_columns, cs.num_instance_columns = 3, 3, 0
    cs.num_selectors, cs.num_challenges = 0, 0
    cs.advice_column_phase = [0, 0, 0]
    cs.num_advice_queries = [2, 1, 1]
    cs.advice_queries = [(0, 0, 0), (0, 0, 1), (1, 0, 0), (2, 0, 0)]
    cs.fixed_queries = [(0, 0), (1, 0), (2, 0)]
    cs.permutation_columns = [(0, 0), (2, 0), (1, COL_FIXED)]
    nv = 7
    A0, A0N, A1, A2, Q, T, T2 = range(7)
    cs.coeff_vals = [1, R - 1]
    cs.gates = [(nv, [(0, [(A0N, 1), (Q, 1)]), (1, [(A0, 1), (Q, 1)]), (1, [(A1, 1), (Q, 1)])])]
    cs.lookups = [([(nv, [(0, [(A1, 1), (Q, 1)])]), (nv, [(0, [(A1, 2)])])], [(nv, [(0, [(T, 1)])]), (nv, [(0, [(T2, 1)])])])]
    cs.shuffles = [([(nv, [(0, [(A2, 1)])])], [(nv, [(0, [(A1, 1)])])])]
    table = [0] + list(range(3, 3 + 20))  # 0 first: rows with q = 0 look up (0, 0)
    fixed = [[1] * rows, table, [v * v % R for v in table]]
    # copy constraints: a2 row 0 must equal table row 5 (the assignment puts that value there), a0 row 1 == a0 row 1 (trivial cycle)
    return Circuit(k, cs, 5, fixed, [((1, 0), (2, 5))])


def lookup_shuffle_assignment(circ, rows, rng, cheat=None):
    n = 1 << circ.k
    usable = n - (circ.cs.blinding_factors() + 1)
    table = circ.fixed[1]

    def running(a1, first):
        a0 = [first]
        for j in range(rows):
            a0.append((a0[-1] + a1[j]) % R)
        return a0 + [0] * (usable - len(a0))

    a1 = [table[1 + rng.randrange(len(table) - 1)] for _ in range(rows)]
    a1[2] = table[5]  # the value the copy constraint pins into a2 row 0
    if cheat == "lookup":
        a1[0] = 2  # not in the table; gate and shuffle stay consistent
    a1 += [0] * (usable - rows)
    a0 = running(a1, rng.randrange(R))
    a2 = list(a1)
    i = a2.index(table[5])
    a2[0], a2[i] = a2[i], a2[0]  # a permutation of a1 with the pinned value in row 0
    tail = a2[1:]
    rng.shuffle(tail)
    a2 = [a2[0]] + tail
    if cheat == "shuffle":
        a2[3] = (a2[3] + 1) % R
    elif cheat == "gate":
        a0[2] = (a0[2] + 1) % R
    elif cheat == "copy":
        j = next(j for j in range(1, usable) if a2[j] != a2[0])
        a2[0], a2[j] = a2[j], a2[0]  # still a shuffle of a1, but row 0 no longer equals the table cell
    return [a0, a1, a2], []


def two_phase_circuit(k, rows):
    """advice a (phase 0), b (phase 1), one user challenge c squeezed after the phase-0 commitments (lib.rs:91-109), fixed
    selector q, instance column pub.  Variables: a b | q | pub | c  (challenges come last, vk.rs:490-500).
    gate   q * (b - c * a)                b = c * a can only be assigned once c is known
    gate   q * (b * b - c^2 * a^2)        a challenge variable with a power
    permutation over (a, phase 0) and the instance column: a row j == public input row j"""
    cs = ConstraintSystem()
    cs.num_fixed_columns, cs.num_advice_columns, cs.num_instance_columns = 1, 2, 1
    cs.num_selectors, cs.num_challenges = 0, 1
    cs.advice_column_phase = [0, 1]
    cs.challenge_phase = [0]
    cs.num_advice_queries = [1, 1]
    cs.advice_queries = [(0, 0, 0), (1, 1, 0)]
    cs.fixed_queries = [(0, 0)]
    cs.instance_queries = [(0, 0)]
    cs.permutation_columns = [(0, 0), (0, COL_INSTANCE)]
    A, B, Q, _PUB, CH = range(5)
    cs.coeff_vals = [1, R - 1]
    cs.gates = [(5, [(0, [(B, 1), (Q, 1)]), (1, [(A, 1), (CH, 1), (Q, 1)])]),
                (5, [(0, [(B, 2), (Q, 1)]), (1, [(A, 2), (CH, 2), (Q, 1)])])]
    return Circuit(k, cs, 4, [[1] * rows], [((0, j), (1, j)) for j in range(rows)])


def two_phase_assignment(a_vals, cheat=False):
    """Returns (advice as a function of the challenge list, instance): b = c * a is assigned in phase 1."""
    def advice(challenges):
        c = challenges[0]
        b = [c * a % R for a in a_vals]
        if cheat:
            b[0] = (b[0] + 1) % R
        return [list(a_vals), b]

    return advice, [list(a_vals)]


# ---------------------------------------------------------------- keygen
def keygen(circ: Circuit, s, transcript_repr=0x1234567):
    cs, k = circ.cs, circ.k
    dom = Domain(k, circ.cs_degree)
    n = dom.n
    bf = cs.blinding_factors()
    usable = n - (bf + 1)
    fixed = [list(col) + [0] * (n - len(col)) for col in circ.fixed]
    ncols = len(cs.permutation_columns)
    # permutation: cycles from the copy constraints (union by swapping images, like the upstream keygen)
    mapping = {(c, j): (c, j) for c in range(ncols) for j in range(n)}
    for a, b in circ.copies:
        mapping[a], mapping[b] = mapping[b], mapping[a]
    delta_pow = [pow(bn.FR_DELTA, c, R) for c in range(ncols)]
    omega_pow = [pow(dom.omega, j, R) for j in range(n)]
    sigma = [[delta_pow[mapping[(c, j)][0]] * omega_pow[mapping[(c, j)][1]] % R for j in range(n)] for c in range(ncols)]
    fixed_c = [dom.lagrange_to_coeff(col) for col in fixed]
    sigma_c = [dom.lagrange_to_coeff(col) for col in sigma]
    commit = lambda coeffs: bn.g1_mul_gen(_eval(coeffs, s))
    sel_bytes = (n + 7) // 8
    vk = VerifyingKey(k=k, fixed_commitments=[commit(c) for c in fixed_c], cs_degree=circ.cs_degree, cs=cs,
                      permutation_commitments=[commit(c) for c in sigma_c], selectors=[bytes(sel_bytes)] * cs.num_selectors,
                      transcript_repr=transcript_repr)
    params = ParamsKZG(k, bn.G1_GEN, bn.G2_GEN, bn.g2_mul(bn.G2_GEN, s % R))
    pk = {"dom": dom, "fixed": fixed, "fixed_c": fixed_c, "sigma": sigma, "sigma_c": sigma_c, "bf": bf, "usable": usable, "circ": circ}
    return params, vk, pk


def keygen_vm(k, s, rows, transcript_repr=0x1234567):
    params, vk, pk = keygen(vm_circuit(k, rows), s, transcript_repr)
    pk["rows"] = rows
    return params, vk, pk


# ---------------------------------------------------------------- prover
def prove(params, vk, pk, s, advice, instance, rng, hash_kind="blake2b", expect_honest=True, multiopen="shplonk"):
    """advice: [column][row] (at most `usable` rows, zero-padded), instance: [column][row].  Returns proof bytes.
    expect_honest = False: the assignment violates a constraint; the quotient is then no polynomial of the allowed
    degree (asserted) and the returned proof must be rejected.  multiopen: "shplonk" or "gwc" opening argument."""
    return prove_multi(params, vk, pk, s, [advice], [instance], rng, hash_kind, expect_honest, multiopen)


def prove_multi(params, vk, pk, s, advices, instances, rng, hash_kind="blake2b", expect_honest=True, multiopen="shplonk"):
    """One proof for m = len(advices) circuit instances of the same circuit (the `instances: &[&[&[Fr]]]` of the reference's
    verify_proof with instances.len() = m, lib.rs:63,92,117,134): advice / lookup / permutation / shuffle polynomials per
    instance, ONE random polynomial and ONE quotient over the y-fold of every instance's expressions, in the interleaving
    the verifier reads.  advices[pi]: [column][row], instances[pi]: [column][row]."""
    dom, bf, usable, circ = pk["dom"], pk["bf"], pk["usable"], pk["circ"]
    cs, n = circ.cs, dom.n
    M = len(advices)
    assert M == len(instances) and M >= 1
    chunk = circ.cs_degree - 2
    ncols = len(cs.permutation_columns)
    n_sets = -(-ncols // chunk) if ncols else 0
    commit = lambda coeffs: bn.g1_mul_gen(_eval(coeffs, s))
    blind = lambda col: list(col) + [0] * (usable - len(col)) + [rng.randrange(R) for _ in range(n - usable)]
    omega_pow = [pow(dom.omega, j, R) for j in range(n)]

    def rot_col(col, rot):
        return [col[(j + rot) % n] for j in range(n)]

    I = []  # per-instance prover state
    for instance in instances:
        st = {"adv": [None] * cs.num_advice_columns, "adv_c": [None] * cs.num_advice_columns}
        st["inst"] = [list(col) + [0] * (n - len(col)) for col in instance]  # lib.rs:204-217 evaluates exactly this polynomial
        st["inst_c"] = [dom.lagrange_to_coeff(col) for col in st["inst"]]
        I.append(st)

    def column(st, idx, typ):
        return pk["fixed"][idx] if typ == COL_FIXED else st["inst"][idx] if typ == COL_INSTANCE else st["adv"][idx]

    def poly_rows(st, poly):
        out = [0] * n
        for coeff, vars_ in poly[1]:
            for j in range(n):
                t = cs.coeff_vals[coeff]
                for v, pw in vars_:
                    t = t * pow(st["var_rows"][v][j], pw, R) % R
                out[j] = (out[j] + t) % R
        return out

    tr = TranscriptWrite(hash_kind)
    tr.common_scalar(vk.transcript_repr)
    for instance in instances:  # lib.rs:76-82
        for col in instance:
            for v in col:
                tr.common_scalar(v)
    # lib.rs:91-109: per phase, every instance's advice commitments of that phase, then the phase's challenges.  An
    # advice assignment that depends on earlier challenges is given as a function of the challenge list.
    challenges = [0] * cs.num_challenges
    for phase in sorted(set(cs.advice_column_phase)):
        for st, advice in zip(I, advices):
            cols = advice(list(challenges)) if callable(advice) else advice
            for c, ph in enumerate(cs.advice_column_phase):
                if ph == phase:
                    st["adv"][c] = blind(cols[c])
                    st["adv_c"][c] = dom.lagrange_to_coeff(st["adv"][c])
                    tr.write_point(commit(st["adv_c"][c]))
        for ci, ph in enumerate(cs.challenge_phase):
            if ph == phase:
                challenges[ci] = tr.squeeze_challenge()
    for st in I:  # values of every query variable on the base domain: advice | fixed | instance | challenges (vk.rs:490-500)
        st["var_rows"] = [rot_col(st["adv"][c], r) for c, _p, r in cs.advice_queries] + [rot_col(pk["fixed"][c], r) for c, r in cs.fixed_queries] \
            + [rot_col(st["inst"][c], r) for c, r in cs.instance_queries] + [[ch] * n for ch in challenges]
    theta = tr.squeeze_challenge()

    def compress_rows(st, polys):
        acc = [0] * n
        for p in polys:
            rows_ = poly_rows(st, p)
            acc = [(a * theta + b) % R for a, b in zip(acc, rows_)]
        return acc

    # ---- lookup argument, permuted columns (definition of the protocol; verifier side: lookup.rs:159-230)
    for st in I:
        st["lookups"] = []
        for inputs, tables in cs.lookups:
            A, S = compress_rows(st, inputs), compress_rows(st, tables)
            Ap = sorted(A[:usable])
            left = {}
            for v in S[:usable]:
                left[v] = left.get(v, 0) + 1
            Sp = [None] * usable
            for i in range(usable):
                if i == 0 or Ap[i] != Ap[i - 1]:
                    if left.get(Ap[i], 0) == 0:
                        assert not expect_honest, "lookup input not in the table"
                        left[Ap[i]] = left.get(Ap[i], 0) + 1  # dishonest prover: pretend
                    Sp[i] = Ap[i]
                    left[Ap[i]] -= 1
            rest = [v for v, c in left.items() for _ in range(max(c, 0))]
            for i in range(usable):
                if Sp[i] is None:
                    Sp[i] = rest.pop() if rest else 0
            Ap, Sp = blind(Ap), blind(Sp)
            Ap_c, Sp_c = dom.lagrange_to_coeff(Ap), dom.lagrange_to_coeff(Sp)
            tr.write_point(commit(Ap_c))
            tr.write_point(commit(Sp_c))
            st["lookups"].append({"A": A, "S": S, "Ap": Ap, "Sp": Sp, "Ap_c": Ap_c, "Sp_c": Sp_c})
    beta = tr.squeeze_challenge()
    gamma = tr.squeeze_challenge()
    # ---- permutation argument: chained grand products, `chunk` columns per set (verifier side: permutation.rs:189-288)
    for st in I:
        z, start = [], 1
        for t in range(n_sets):
            cols = list(range(t * chunk, min((t + 1) * chunk, ncols)))
            zi = [0] * n
            zi[0] = start
            dens = [1] * usable
            for c in cols:
                colv = column(st, *cs.permutation_columns[c])
                dens = [d * ((colv[j] + beta * pk["sigma"][c][j] + gamma) % R) % R for j, d in enumerate(dens)]
            dens = bn.batch_invert_skip_zero(dens, R)
            for j in range(usable):
                num = 1
                for c in cols:
                    colv = column(st, *cs.permutation_columns[c])
                    num = num * ((colv[j] + beta * pow(bn.FR_DELTA, c, R) % R * omega_pow[j] + gamma) % R) % R
                zi[j + 1] = zi[j] * num % R * dens[j] % R
            start = zi[usable]
            for j in range(usable + 1, n):
                zi[j] = rng.randrange(R)
            z.append(zi)
        if n_sets and expect_honest:
            assert start == 1, "grand product of an honest permutation must close to 1"
        st["z_c"] = [dom.lagrange_to_coeff(zi) for zi in z]
        for c in st["z_c"]:
            tr.write_point(commit(c))
    # ---- lookup / shuffle grand products
    for st in I:
        for L in st["lookups"]:
            zl = [0] * n
            zl[0] = 1
            dens = bn.batch_invert_skip_zero([((L["Ap"][j] + beta) % R) * ((L["Sp"][j] + gamma) % R) % R for j in range(usable)], R)
            for j in range(usable):
                zl[j + 1] = zl[j] * ((L["A"][j] + beta) % R) % R * ((L["S"][j] + gamma) % R) % R * dens[j] % R
            for j in range(usable + 1, n):
                zl[j] = rng.randrange(R)
            L["Z"], L["Z_c"] = zl, dom.lagrange_to_coeff(zl)
            tr.write_point(commit(L["Z_c"]))
    for st in I:
        st["shuffles"] = []
        for inputs, shufs in cs.shuffles:
            A, S = compress_rows(st, inputs), compress_rows(st, shufs)
            zs = [0] * n
            zs[0] = 1
            dens = bn.batch_invert_skip_zero([(S[j] + gamma) % R for j in range(usable)], R)
            for j in range(usable):
                zs[j + 1] = zs[j] * ((A[j] + gamma) % R) % R * dens[j] % R
            for j in range(usable + 1, n):
                zs[j] = rng.randrange(R)
            sh = {"A": A, "S": S, "Z": zs, "Z_c": dom.lagrange_to_coeff(zs)}
            tr.write_point(commit(sh["Z_c"]))
            st["shuffles"].append(sh)
    rand_c = [rng.randrange(R) for _ in range(n)]  # vanishing.rs:49-57: random polynomial
    tr.write_point(commit(rand_c))
    y = tr.squeeze_challenge()

    # ---- quotient: N(X) = fold_y(expressions of every instance)(X) on the extended coset, h = N / (X^n - 1)
    E = dom.coeff_to_ext
    fix_e = [E(c) for c in pk["fixed_c"]]
    m = 1 << dom.ext_k
    x_e = [dom.shift * pow(dom.ext_omega, i, R) % R for i in range(m)]
    lag = lambda rows_: E(dom.lagrange_to_coeff([1 if j in rows_ else 0 for j in range(n)]))
    l0_e, llast_e, lblind_e = lag({0}), lag({usable}), lag(set(range(usable + 1, n)))
    active_e = [(1 - llast_e[i] - lblind_e[i]) % R for i in range(m)]
    sig_e = [E(c) for c in pk["sigma_c"]] if n_sets else []
    exprs = []
    for st in I:  # lib.rs:273-344: every expression of instance pi, then the next instance
        adv_e, inst_e = [E(c) for c in st["adv_c"]], [E(c) for c in st["inst_c"]]
        var_e = [dom.rotate_ext(adv_e[c], r) for c, _p, r in cs.advice_queries] + [dom.rotate_ext(fix_e[c], r) for c, r in cs.fixed_queries] \
            + [dom.rotate_ext(inst_e[c], r) for c, r in cs.instance_queries] + [[ch] * m for ch in challenges]

        def poly_ext(poly):
            out = [0] * m
            for coeff, vars_ in poly[1]:
                for i in range(m):
                    t = cs.coeff_vals[coeff]
                    for v, pw in vars_:
                        t = t * pow(var_e[v][i], pw, R) % R
                    out[i] = (out[i] + t) % R
            return out

        def compress_ext(polys):
            acc = [0] * m
            for p in polys:
                pe = poly_ext(p)
                acc = [(a * theta + b) % R for a, b in zip(acc, pe)]
            return acc

        exprs += [poly_ext(g) for g in cs.gates]  # vk.rs:478-512
        if n_sets:  # permutation.rs:189-288
            z_e = [E(c) for c in st["z_c"]]

            def col_ext(idx, typ):
                return fix_e[idx] if typ == COL_FIXED else inst_e[idx] if typ == COL_INSTANCE else adv_e[idx]

            exprs.append([l0_e[i] * ((1 - z_e[0][i]) % R) % R for i in range(m)])
            exprs.append([llast_e[i] * ((z_e[-1][i] * z_e[-1][i] - z_e[-1][i]) % R) % R for i in range(m)])
            for t in range(1, n_sets):
                zl = dom.rotate_ext(z_e[t - 1], -(bf + 1))
                exprs.append([l0_e[i] * ((z_e[t][i] - zl[i]) % R) % R for i in range(m)])
            for t in range(n_sets):
                cols = list(range(t * chunk, min((t + 1) * chunk, ncols)))
                zn = dom.rotate_ext(z_e[t], 1)
                left, right = list(zn), list(z_e[t])
                for c in cols:
                    ce = col_ext(*cs.permutation_columns[c])
                    dp = pow(bn.FR_DELTA, c, R)
                    left = [left[i] * ((ce[i] + beta * sig_e[c][i] + gamma) % R) % R for i in range(m)]
                    right = [right[i] * ((ce[i] + beta * dp % R * x_e[i] + gamma) % R) % R for i in range(m)]
                exprs.append([active_e[i] * ((left[i] - right[i]) % R) % R for i in range(m)])
        for (inputs, tables), L in zip(cs.lookups, st["lookups"]):  # lookup.rs:159-230
            Z, Ap, Sp = E(L["Z_c"]), E(L["Ap_c"]), E(L["Sp_c"])
            Zn, Apm = dom.rotate_ext(Z, 1), dom.rotate_ext(Ap, -1)
            Ae, Se = compress_ext(inputs), compress_ext(tables)
            exprs.append([l0_e[i] * ((1 - Z[i]) % R) % R for i in range(m)])
            exprs.append([llast_e[i] * ((Z[i] * Z[i] - Z[i]) % R) % R for i in range(m)])
            exprs.append([active_e[i] * ((Zn[i] * ((Ap[i] + beta) % R) % R * ((Sp[i] + gamma) % R)
                                         - Z[i] * ((Ae[i] + beta) % R) % R * ((Se[i] + gamma) % R)) % R) % R for i in range(m)])
            exprs.append([l0_e[i] * ((Ap[i] - Sp[i]) % R) % R for i in range(m)])
            exprs.append([active_e[i] * ((Ap[i] - Sp[i]) % R) % R * ((Ap[i] - Apm[i]) % R) % R for i in range(m)])
        for (inputs, shufs), sh in zip(cs.shuffles, st["shuffles"]):  # shuffle.rs:148-203
            Z = E(sh["Z_c"])
            Zn = dom.rotate_ext(Z, 1)
            Ae, Se = compress_ext(inputs), compress_ext(shufs)
            exprs.append([l0_e[i] * ((1 - Z[i]) % R) % R for i in range(m)])
            exprs.append([llast_e[i] * ((Z[i] * Z[i] - Z[i]) % R) % R for i in range(m)])
            exprs.append([active_e[i] * ((Zn[i] * ((Se[i] + gamma) % R) - Z[i] * ((Ae[i] + gamma) % R)) % R) % R for i in range(m)])
    N = [0] * m
    for e in exprs:
        N = [(N[i] * y + e[i]) % R for i in range(m)]
    vinv = bn.batch_invert_skip_zero([(pow(xe, n, R) - 1) % R for xe in x_e], R)
    h_c = dom.ext_to_coeff([N[i] * vinv[i] % R for i in range(m)])
    n_h = circ.cs_degree - 1
    honest = all(c == 0 for c in h_c[n_h * n:])  # deg h < (cs_degree - 1) n  iff  N vanishes on the domain
    assert honest == expect_honest, "divisibility of the folded constraint polynomial by X^n - 1"
    h_pieces = [h_c[i * n:(i + 1) * n] for i in range(n_h)]
    for piece in h_pieces:
        tr.write_point(commit(piece))
    x = tr.squeeze_challenge()
    xn = pow(x, n, R)
    rot = lambda r: x * pow(dom.omega, r % n, R) % R
    ev = lambda coeffs, r=0: _eval(coeffs, rot(r))
    for st in I:  # lib.rs:219-253
        for c, _p, r in cs.advice_queries:
            tr.write_scalar(ev(st["adv_c"][c], r))
    for c, r in cs.fixed_queries:
        tr.write_scalar(ev(pk["fixed_c"][c], r))
    tr.write_scalar(ev(rand_c))
    for c in pk["sigma_c"]:
        tr.write_scalar(ev(c))
    for st in I:
        for t in range(n_sets):
            tr.write_scalar(ev(st["z_c"][t]))
            tr.write_scalar(ev(st["z_c"][t], 1))
            if t != n_sets - 1:
                tr.write_scalar(ev(st["z_c"][t], -(bf + 1)))
    for st in I:
        for L in st["lookups"]:  # product, product_next, permuted input, permuted input at omega^-1 x, permuted table
            for coeffs, r in ((L["Z_c"], 0), (L["Z_c"], 1), (L["Ap_c"], 0), (L["Ap_c"], -1), (L["Sp_c"], 0)):
                tr.write_scalar(ev(coeffs, r))
    for st in I:
        for sh in st["shuffles"]:
            tr.write_scalar(ev(sh["Z_c"]))
            tr.write_scalar(ev(sh["Z_c"], 1))
    # ---- SHPLONK opening of every query (shplonk.rs): polynomials by commitment identity, slots in transcript order
    n_adv, n_lk, n_sh = cs.num_advice_columns, len(cs.lookups), len(cs.shuffles)
    b_lk = M * n_adv
    b_perm = b_lk + 2 * M * n_lk
    b_lkz = b_perm + M * n_sets
    b_shz = b_lkz + M * n_lk
    slot_random = b_shz + M * n_sh
    polys, queries = {}, []

    def add_q(ident, coeffs, r):
        polys[ident] = coeffs
        queries.append(Query(ident, rot(r), _eval(coeffs, rot(r)), None, r))

    for pi, st in enumerate(I):  # lib.rs:349-414
        for c, _p, r in cs.advice_queries:
            add_q(("proof", pi * n_adv + c), st["adv_c"][c], r)
        for t in range(n_sets):
            add_q(("proof", b_perm + pi * n_sets + t), st["z_c"][t], 0)
            add_q(("proof", b_perm + pi * n_sets + t), st["z_c"][t], 1)
        for t in reversed(range(n_sets - 1)):
            add_q(("proof", b_perm + pi * n_sets + t), st["z_c"][t], -(bf + 1))
        for li, L in enumerate(st["lookups"]):
            s_in, s_tab, s_z = b_lk + 2 * (pi * n_lk + li), b_lk + 2 * (pi * n_lk + li) + 1, b_lkz + pi * n_lk + li
            add_q(("proof", s_z), L["Z_c"], 0)
            add_q(("proof", s_in), L["Ap_c"], 0)
            add_q(("proof", s_tab), L["Sp_c"], 0)
            add_q(("proof", s_in), L["Ap_c"], -1)
            add_q(("proof", s_z), L["Z_c"], 1)
        for si, sh in enumerate(st["shuffles"]):
            add_q(("proof", b_shz + pi * n_sh + si), sh["Z_c"], 0)
            add_q(("proof", b_shz + pi * n_sh + si), sh["Z_c"], 1)
    for c, r in cs.fixed_queries:
        add_q(("fixed", c), pk["fixed_c"][c], r)
    for t in range(ncols):
        add_q(("sigma", t), pk["sigma_c"][t], 0)
    hmsm_c = [0] * n
    for piece in reversed(h_pieces):
        hmsm_c = [(hmsm_c[i] * xn + piece[i]) % R for i in range(n)]
    add_q(("hmsm",), hmsm_c, 0)
    add_q(("proof", slot_random), rand_c, 0)
    if multiopen == "gwc":
        # GWC (gwc.rs:54-163): queries grouped by point in first-appearance order; per point z_i the witness
        # W_i = [ sum_j v^j (P_j(X) - e_j) / (X - z_i) ](s) G, the powers of v restarting in every group
        v = tr.squeeze_challenge()
        groups = []
        for q in queries:
            for pt, qs in groups:
                if pt == q.point:
                    qs.append(q)
                    break
            else:
                groups.append((q.point, [q]))
        for z, qs in groups:
            acc, pow_v = 0, 1
            for q in qs:
                acc = (acc + pow_v * ((_eval(polys[q.ident], s) - q.eval) % R)) % R
                pow_v = pow_v * v % R
            tr.write_point(bn.g1_mul_gen(acc * bn.fr_inv((s - z) % R) % R))
        return bytes(tr.out)
    assert multiopen == "shplonk"
    rotation_sets, super_points = _shplonk_sets(queries)
    yy = tr.squeeze_challenge()
    v = tr.squeeze_challenge()
    van = lambda roots, at: _prod((at - p) % R for p in roots)
    # h(X) = sum_i v^i (P_i(X) - R_i(X)) / Z_{T_i}(X), committed as [h(s)] G
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
    return bytes(tr.out)


def prove_vm(params, vk, pk, s, lhs, rhs, rng, hash_kind="blake2b", cheat_row=None):
    """Returns (proof bytes, instances [[column]]) for out_j = lhs_j * rhs_j.  `cheat_row`: put a wrong product into
    that row of a2 AND of the public input (the copy constraint still holds, the gate does not)."""
    advice, instance = vm_assignment(lhs, rhs, cheat_row)
    return prove(params, vk, pk, s, advice, instance, rng, hash_kind, expect_honest=cheat_row is None), [instance]


def demo(seed=1, k=8, rows=10):
    from verifier import verify_proof

    rng = random.Random(seed)
    s = 0x1C59A59B6CFF4308740943526ADE1D8C09F71B337A67269CC89586BCDD6DFCBA if k == 8 else rng.randrange(1, R)
    params, vk, pk = keygen_vm(k, s, rows)
    lhs = [rng.randrange(R) for _ in range(rows)]
    rhs = [rng.randrange(R) for _ in range(rows)]
    proof, inst = prove_vm(params, vk, pk, s, lhs, rhs, rng)
    r1 = verify_proof(params, vk, inst, proof)
    circ = lookup_shuffle_circuit(6, 16)
    params2, vk2, pk2 = keygen(circ, s)
    adv, ins = lookup_shuffle_assignment(circ, 16, rng)
    proof2 = prove(params2, vk2, pk2, s, adv, ins, rng)
    r2 = verify_proof(params2, vk2, [ins], proof2)
    return r1, r2


if __name__ == "__main__":
    a, b = demo()
    print("vector_mul honest proof status:", a.status, a.error)
    print("lookup/shuffle honest proof status:", b.status, b.error)
