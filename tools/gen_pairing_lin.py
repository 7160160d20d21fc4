#!/usr/bin/env python
"""Generates halo2-verifier_b200/csrc/pairing_lin.inc: the integer matrices of the CTA-cooperative
Fq12 engine (pairing_cta.cuh).

An Fq12 value x = x0 + x1 w (x0, x1 in Fq6 = Fq2[v]/(v^3 - xi), w^2 = v, xi = 9 + u) is kept in an
EXPANDED form of 54 Fq values: for each of the three Karatsuba operands g = x0, x1, x0 + x1 the six
Fq2 operands of a Karatsuba-3 Fq6 product (A0, A1, A2, A1+A2, A0+A1, A0+A2), each as (c0, c1, c0+c1).
A product is then 54 independent Fq multiplications (one per lane) followed by ONE linear map back to
the expanded form of the result.  FULL is that map (54 x 54, all interpolation steps and the
re-expansion composed); EXP expands 12 base coordinates.  Negative coefficients are expressed
against the negated copy of the source (index + N), so every coefficient is a small positive integer.
"""
import os
import numpy as np

NB, NE = 12, 54


def base(hh, j, part):
    return 2 * (3 * hh + j) + part


def eidx(g, i, q):
    return 3 * (6 * g + i) + q


def build():
    EXP = np.zeros((NE, NB), dtype=np.int64)
    pairs = {0: [0], 1: [1], 2: [2], 3: [1, 2], 4: [0, 1], 5: [0, 2]}
    for g in range(3):
        halves = [0] if g == 0 else [1] if g == 1 else [0, 1]
        for i in range(6):
            for hh in halves:
                for j in pairs[i]:
                    EXP[eidx(g, i, 0), base(hh, j, 0)] += 1
                    EXP[eidx(g, i, 1), base(hh, j, 1)] += 1
                    EXP[eidx(g, i, 2), base(hh, j, 0)] += 1
                    EXP[eidx(g, i, 2), base(hh, j, 1)] += 1

    def P(g, i):  # Fq2 product i of group g from its three Fq products (Karatsuba)
        re = np.zeros(NE, dtype=np.int64)
        im = np.zeros(NE, dtype=np.int64)
        re[eidx(g, i, 0)] += 1
        re[eidx(g, i, 1)] -= 1
        im[eidx(g, i, 2)] += 1
        im[eidx(g, i, 0)] -= 1
        im[eidx(g, i, 1)] -= 1
        return (re, im)

    add = lambda a, b: (a[0] + b[0], a[1] + b[1])
    sub = lambda a, b: (a[0] - b[0], a[1] - b[1])
    xi = lambda a: (9 * a[0] - a[1], 9 * a[1] + a[0])
    T = []
    for g in range(3):
        p = [P(g, i) for i in range(6)]
        T.append([add(p[0], xi(sub(sub(p[3], p[1]), p[2]))),
                  add(sub(sub(p[4], p[0]), p[1]), xi(p[2])),
                  add(sub(sub(p[5], p[0]), p[2]), p[1])])
    z0 = [add(T[0][0], xi(T[1][2])), add(T[0][1], T[1][0]), add(T[0][2], T[1][1])]
    z1 = [sub(sub(T[2][j], T[0][j]), T[1][j]) for j in range(3)]
    INT = np.zeros((NB, NE), dtype=np.int64)
    for hh, z in enumerate([z0, z1]):
        for j in range(3):
            INT[base(hh, j, 0)] = z[j][0]
            INT[base(hh, j, 1)] = z[j][1]
    return EXP, INT, EXP @ INT


def emit(name, M, out):
    n_src = M.shape[1]
    starts, terms = [0], []
    for row in M:
        pos = [(j, int(c)) for j, c in enumerate(row) if c > 0]
        neg = [(j + n_src, int(-c)) for j, c in enumerate(row) if c < 0]
        for j, c in pos + neg:
            assert 0 < c < 256 and j < 256
            terms.append(j | (c << 8))
        while len(terms) % 4:  # rows padded to a multiple of 4 terms (coefficient 0): lin_row is unrolled by 4
            terms.append(0)
        starts.append(len(terms))
    out.append(f"// {name}: {M.shape[0]} rows over {n_src} sources (+{n_src}: negated copy); "
               f"max terms/row {max(b - a for a, b in zip(starts, starts[1:]))}, max coefficient sum {int(np.abs(M).sum(axis=1).max())}")
    out.append(f"#define H2V_LIN_{name}_NTERMS {len(terms)}")
    out.append(f"#define H2V_LIN_{name}_START_INIT {{{', '.join(map(str, starts))}}}")
    out.append(f"#define H2V_LIN_{name}_TERMS_INIT {{{', '.join(map(str, terms))}}}")


def emit_fast(M, out, threads=128):
    """FULL for the latency-critical 128-thread group (k_pairing_check): every row is split over 2-3 CONSECUTIVE threads of
    one warp, each thread keeps its <= `cap` terms in registers for the whole kernel, partial column accumulators are
    combined with warp shuffles.  Per thread: cap packed terms (index | coefficient << 8), then a meta word
    row | lead << 8 | followers << 16 (lead = first thread of its row: adds the partials of its `followers` next lanes and
    reduces; row 255 = idle thread)."""
    n_src = M.shape[1]
    rows = []
    for row in M:
        pos = [(j, int(c)) for j, c in enumerate(row) if c > 0]
        neg = [(j + n_src, int(-c)) for j, c in enumerate(row) if c < 0]
        rows.append([j | (c << 8) for j, c in pos + neg])
    best = None
    for cap in range(8, 40):
        slots, ok = [], True
        for r, terms in enumerate(rows):
            k = -(-len(terms) // cap)
            if k > 4:
                ok = False
                break
            if (len(slots) % 32) + k > 32:  # the threads of a row share a warp
                slots += [None] * (32 - len(slots) % 32)
            per = -(-len(terms) // k)
            for i in range(k):
                slots.append((r, i == 0, k - 1 if i == 0 else 0, terms[i * per:(i + 1) * per]))
        if ok and len(slots) <= threads:
            best = (cap, slots)
            break
    cap, slots = best
    slots += [None] * (threads - len(slots))
    words = []
    for sl in slots:
        if sl is None:
            words += [0] * (cap // 2 + cap % 2) + [255]
            continue
        r, lead, fol, terms = sl
        terms = terms + [0] * (cap + cap % 2 - len(terms))
        words += [terms[i] | (terms[i + 1] << 16) for i in range(0, len(terms), 2)]
        words.append(r | (int(lead) << 8) | (fol << 16))
    out.append(f"// FAST: FULL split over {threads} threads, <= {cap} register-resident terms per thread ({sum(1 for x in slots if x)} busy threads)")
    out.append(f"#define H2V_LIN_FAST_CAP {cap + cap % 2}")
    out.append(f"#define H2V_LIN_FAST_THREADS {threads}")
    out.append(f"#define H2V_LIN_FAST_MAX_FOLLOWERS {max(sl[2] for sl in slots if sl)}")
    out.append(f"#define H2V_LIN_FAST_INIT {{{', '.join(map(str, words))}}}")


def main():
    EXP, INT, FULL = build()
    out = ["// generated by tools/gen_pairing_lin.py -- do not edit"]
    emit("EXP", EXP, out)
    emit("INT", INT, out)
    emit("FULL", FULL, out)
    emit_fast(FULL, out)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "halo2-verifier_b200", "csrc", "pairing_lin.inc")
    open(path, "w").write("\n".join(out) + "\n")
    print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    main()
