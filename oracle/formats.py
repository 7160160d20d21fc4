"""Params / VerifyingKey byte formats of the reference, restated (TEST INFRASTRUCTURE ONLY).

Follows:
  helpers.rs:7-19,40-98,120-164           SerdeFormat, curve/field (de)serialisation, BE integers
  poly/kzg/commitment.rs:142-207          ParamsKZG::{write_custom, read_custom}  (k is LE!)
  plonk/vk.rs:41-64,76-115                VerifyingKey::{write, read}
  plonk/vk.rs:214-365                     ConstraintSystem::{write, read}
  plonk/vk.rs:514-546                     IndexedExpressionPoly::{write, read}
  plonk/circuit.rs:36-65                  Column<Any> codec (255 fixed, 254 instance, 0..2 advice phase)
  plonk/permutation.rs:29-44,164-176      permutation Argument / VerifyingKey
  plonk/lookup.rs:51-68, shuffle.rs:85-102   READ side: (input, table) pairs interleaved
  poly/domain.rs:34-140                   EvaluationDomain::new (only omega, omega_inv, 1/n, quotient degree)

The writers emit what the reference's `read` expects (SURVEY.md section 4 caveats:
lookup/shuffle interleaving; exactly num_instance_columns / num_fixed_columns queries).
"""
import struct
from dataclasses import dataclass, field
from typing import List, Tuple

import bn254 as bn

PROCESSED, RAW_BYTES, RAW_BYTES_UNCHECKED = 0, 1, 2

COL_FIXED, COL_INSTANCE = 255, 254


class FormatError(Exception):
    pass


class Reader:
    def __init__(self, data: bytes):
        self.d = bytes(data)
        self.p = 0

    def take(self, n):
        if self.p + n > len(self.d):
            raise FormatError("failed to fill whole buffer")
        b = self.d[self.p : self.p + n]
        self.p += n
        return b

    def u8(self):
        return self.take(1)[0]

    def u16(self):
        return struct.unpack(">H", self.take(2))[0]

    def u32(self):
        return struct.unpack(">I", self.take(4))[0]

    def i32(self):
        return struct.unpack(">i", self.take(4))[0]

    def g1(self, fmt):
        if fmt == PROCESSED:
            ok, pt = bn.g1_from_bytes(self.take(32))
        else:
            ok, pt = bn.g1_read_raw(self.take(64))
        if not ok:
            raise FormatError("Invalid point encoding")
        return pt

    def g2(self, fmt):
        if fmt == PROCESSED:
            ok, pt = bn.g2_from_bytes(self.take(64))
        else:
            ok, pt = bn.g2_read_raw(self.take(128))
        if not ok:
            raise FormatError("Invalid point encoding")
        return pt

    def fr(self, fmt):
        if fmt == PROCESSED:
            v = bn.fr_from_repr(self.take(32))
        else:
            v = bn.fr_read_raw(self.take(32))
        if v is None:
            raise FormatError("Invalid prime field point encoding")
        return v


def _g1_bytes(pt, fmt):
    return bn.g1_to_bytes(pt) if fmt == PROCESSED else bn.g1_write_raw(pt)


def _g2_bytes(pt, fmt):
    return bn.g2_to_bytes(pt) if fmt == PROCESSED else bn.g2_write_raw(pt)


def _fr_bytes(v, fmt):
    return bn.fr_to_repr(v) if fmt == PROCESSED else bn.fr_write_raw(v)


# ---------------------------------------------------------------- ParamsKZG
@dataclass
class ParamsKZG:
    k: int
    g: tuple
    g2: tuple
    s_g2: tuple

    @property
    def n(self):
        return 1 << self.k

    def to_bytes(self, fmt=PROCESSED) -> bytes:
        return struct.pack("<I", self.k) + _g1_bytes(self.g, fmt) + _g2_bytes(self.g2, fmt) + _g2_bytes(self.s_g2, fmt)

    @staticmethod
    def from_bytes(data: bytes, fmt=PROCESSED) -> "ParamsKZG":
        r = Reader(data)
        k = struct.unpack("<I", r.take(4))[0]
        return ParamsKZG(k, r.g1(fmt), r.g2(fmt), r.g2(fmt))


def params_from_srs_fixture(raw: bytes) -> ParamsKZG:
    """serialize::convert_params (serialize/src/lib.rs:26-36) applied to the upstream PSE
    RawBytes SRS layout: u32le k | 2^k G1 g | 2^k G1 g_lagrange | G2 g2 | G2 s_g2."""
    k = struct.unpack("<I", raw[:4])[0]
    n = 1 << k
    ok, g = bn.g1_read_raw(raw[4:68])
    off = 4 + 2 * n * 64
    ok2, g2 = bn.g2_read_raw(raw[off : off + 128])
    ok3, s_g2 = bn.g2_read_raw(raw[off + 128 : off + 256])
    assert ok and ok2 and ok3 and len(raw) == off + 256
    return ParamsKZG(k, g, g2, s_g2)


# ---------------------------------------------------------------- VerifyingKey
Poly = Tuple[int, List[Tuple[int, List[Tuple[int, int]]]]]  # (num_vars, [(coeff_idx, [(var, pow)])])


@dataclass
class ConstraintSystem:
    num_fixed_columns: int = 0
    num_advice_columns: int = 0
    num_instance_columns: int = 0
    num_selectors: int = 0
    num_challenges: int = 0
    advice_column_phase: List[int] = field(default_factory=list)
    challenge_phase: List[int] = field(default_factory=list)
    num_advice_queries: List[int] = field(default_factory=list)
    advice_queries: List[Tuple[int, int, int]] = field(default_factory=list)  # (col, phase, rot)
    instance_queries: List[Tuple[int, int]] = field(default_factory=list)  # (col, rot)
    fixed_queries: List[Tuple[int, int]] = field(default_factory=list)
    permutation_columns: List[Tuple[int, int]] = field(default_factory=list)  # (index, type byte)
    gates: List[Poly] = field(default_factory=list)
    lookups: List[Tuple[List[Poly], List[Poly]]] = field(default_factory=list)
    shuffles: List[Tuple[List[Poly], List[Poly]]] = field(default_factory=list)
    coeff_vals: List[int] = field(default_factory=list)

    def blinding_factors(self):  # vk.rs:396-401
        factors = max(self.num_advice_queries) if self.num_advice_queries else 1
        return max(3, factors) + 2

    def phases(self):  # vk.rs:403-411
        return range(0, (max(self.advice_column_phase) if self.advice_column_phase else 0) + 1)

    def get_any_query_index(self, col, rot):  # vk.rs:413-455
        index, typ = col
        if typ == COL_FIXED:
            lst = [(c, r) for c, r in self.fixed_queries]
            key = (index, rot)
        elif typ == COL_INSTANCE:
            lst = [(c, r) for c, r in self.instance_queries]
            key = (index, rot)
        else:
            lst = list(self.advice_queries)
            key = (index, typ, rot)
        for i, q in enumerate(lst):
            if q == key:
                return i
        raise FormatError("get_query_index called for non-existent query")  # reference panics


def _write_poly(p: Poly) -> bytes:
    num_vars, terms = p
    out = struct.pack(">II", num_vars, len(terms))
    for coeff, vars_ in terms:
        out += struct.pack(">HI", coeff, len(vars_))
        for var, pw in vars_:
            out += struct.pack(">II", var, pw)
    return out


def _read_poly(r: Reader) -> Poly:
    num_vars = r.u32()
    num_terms = r.u32()
    terms = []
    for _ in range(num_terms):
        coeff = r.u16()
        n = r.u32()
        terms.append((coeff, [(r.u32(), r.u32()) for _ in range(n)]))
    return (num_vars, terms)


@dataclass
class VerifyingKey:
    k: int
    fixed_commitments: list
    cs_degree: int
    cs: ConstraintSystem
    permutation_commitments: list
    selectors: List[bytes]
    transcript_repr: int

    # --- EvaluationDomain::new(cs_degree, k), domain.rs:34-140 (verify path uses these only)
    @property
    def n(self):
        return 1 << self.k

    @property
    def omega(self):
        return pow(bn.FR_ROOT_OF_UNITY, 1 << (bn.FR_S - self.k), bn.R)

    @property
    def omega_inv(self):
        return bn.fr_inv(self.omega)

    @property
    def barycentric_weight(self):
        return bn.fr_inv(self.n % bn.R)

    @property
    def quotient_poly_degree(self):
        return self.cs_degree - 1

    def to_bytes(self, fmt=RAW_BYTES) -> bytes:
        cs = self.cs
        out = struct.pack(">II", self.k, len(self.fixed_commitments))
        for c in self.fixed_commitments:
            out += _g1_bytes(c, fmt)
        out += struct.pack(">I", self.cs_degree)
        out += struct.pack(
            ">9I",
            cs.num_fixed_columns,
            cs.num_advice_columns,
            cs.num_instance_columns,
            cs.num_selectors,
            cs.num_challenges,
            len(cs.gates),
            len(cs.lookups),
            len(cs.shuffles),
            len(cs.coeff_vals),
        )
        out += bytes(cs.advice_column_phase) + bytes(cs.challenge_phase)
        for n in cs.num_advice_queries:
            out += struct.pack(">I", n)
        for col, phase, rot in cs.advice_queries:
            out += struct.pack(">IBi", col, phase, rot)
        assert len(cs.instance_queries) == cs.num_instance_columns  # what `read` expects
        for col, rot in cs.instance_queries:
            out += struct.pack(">Ii", col, rot)
        assert len(cs.fixed_queries) == cs.num_fixed_columns
        for col, rot in cs.fixed_queries:
            out += struct.pack(">Ii", col, rot)
        out += struct.pack(">I", len(cs.permutation_columns))
        for idx, typ in cs.permutation_columns:
            out += struct.pack(">IB", idx, typ)
        for g in cs.gates:
            out += _write_poly(g)
        for inputs, tables in list(cs.lookups) + list(cs.shuffles):
            assert len(inputs) == len(tables)
            out += struct.pack(">I", len(inputs))
            for a, b in zip(inputs, tables):  # interleaved: what `read` expects
                out += _write_poly(a) + _write_poly(b)
        for v in cs.coeff_vals:
            out += _fr_bytes(v, fmt)
        assert len(self.permutation_commitments) == len(cs.permutation_columns)
        for c in self.permutation_commitments:
            out += _g1_bytes(c, fmt)
        assert len(self.selectors) == cs.num_selectors
        for s in self.selectors:
            assert len(s) == ((1 << self.k) + 7) // 8
            out += s
        out += _fr_bytes(self.transcript_repr, fmt)
        return out

    @staticmethod
    def from_bytes(data: bytes, fmt=RAW_BYTES) -> "VerifyingKey":
        r = Reader(data)
        k = r.u32()
        nfc = r.u32()
        fixed = [r.g1(fmt) for _ in range(nfc)]
        cs_degree = r.u32()
        cs = ConstraintSystem()
        (
            cs.num_fixed_columns,
            cs.num_advice_columns,
            cs.num_instance_columns,
            cs.num_selectors,
            cs.num_challenges,
        ) = (r.u32() for _ in range(5))
        num_gates, num_lookups, num_shuffles, num_coeff = (r.u32() for _ in range(4))
        cs.advice_column_phase = [r.u8() for _ in range(cs.num_advice_columns)]
        cs.challenge_phase = [r.u8() for _ in range(cs.num_challenges)]
        cs.num_advice_queries = [r.u32() for _ in range(cs.num_advice_columns)]
        cs.advice_queries = [(r.u32(), r.u8(), r.i32()) for _ in range(sum(cs.num_advice_queries))]
        cs.instance_queries = [(r.u32(), r.i32()) for _ in range(cs.num_instance_columns)]
        cs.fixed_queries = [(r.u32(), r.i32()) for _ in range(cs.num_fixed_columns)]
        nperm = r.u32()
        for _ in range(nperm):
            idx, typ = r.u32(), r.u8()
            if typ not in (COL_FIXED, COL_INSTANCE, 0, 1, 2):
                raise FormatError("Invalid phase for advice column")
            cs.permutation_columns.append((idx, typ))
        cs.gates = [_read_poly(r) for _ in range(num_gates)]
        for dst, cnt in ((cs.lookups, num_lookups), (cs.shuffles, num_shuffles)):
            for _ in range(cnt):
                m = r.u32()
                ins, tabs = [], []
                for _ in range(m):
                    ins.append(_read_poly(r))
                    tabs.append(_read_poly(r))
                dst.append((ins, tabs))
        cs.coeff_vals = [r.fr(fmt) for _ in range(num_coeff)]
        perm = [r.g1(fmt) for _ in range(nperm)]
        selectors = [r.take(((1 << k) + 7) // 8) for _ in range(cs.num_selectors)]
        repr_ = r.fr(fmt)
        return VerifyingKey(k, fixed, cs_degree, cs, perm, selectors, repr_)
