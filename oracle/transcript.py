"""Fiat-Shamir transcripts of the reference, restated (TEST INFRASTRUCTURE ONLY).

Follows halo2_verifier/src/transcript/mod.rs:
  prefixes                         :16-39
  Blake2bRead::init                :118-134  (Blake2b-512, personal "Halo2-Transcript")
  Keccak256Read::init              :136-151  (Keccak256 pre-loaded with "Halo2-Transcript")
  read_point / read_scalar         :153-203
  squeeze_challenge / common_*     :205-272
  Challenge255::new / get_scalar   :484-515
  Blake2bWrite / Keccak256Write    :274-438  (writer mirrors, used by the proof simulator)

Blake2b is Python's hashlib (same function as blake2b_simd 1.x with these
parameters).  Keccak-256 (sha3 0.9.1 `Keccak256`: ORIGINAL 0x01 padding, not
NIST SHA3) is implemented below because hashlib has no Keccak; KAT:
keccak256(b"") = c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470.
"""
import hashlib

from bn254 import (
    fq_to_repr,
    fr_from_repr,
    fr_from_uniform_bytes,
    fr_to_repr,
    g1_from_bytes,
    g1_to_bytes,
)

PREFIX_CHALLENGE = 0
PREFIX_POINT = 1
PREFIX_SCALAR = 2
KECCAK_PREFIX_LO = 10
KECCAK_PREFIX_HI = 11

# ---------------------------------------------------------------- Keccak-256
_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_ROT = [
    [0, 36, 3, 41, 18],
    [1, 44, 10, 45, 2],
    [62, 6, 43, 15, 61],
    [28, 55, 25, 21, 56],
    [27, 20, 39, 8, 14],
]
_M64 = (1 << 64) - 1


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & _M64 if n else v


def keccak_f1600(a):
    """a: 25 lanes, index x + 5*y."""
    for rnd in range(24):
        c = [a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [a[i] ^ d[i % 5] for i in range(25)]
        b = [0] * 25
        for x in range(5):
            for y in range(5):
                b[y + 5 * ((2 * x + 3 * y) % 5)] = _rol(a[x + 5 * y], _ROT[x][y])
        a = [b[i] ^ ((~b[(i % 5 + 1) % 5 + 5 * (i // 5)]) & b[(i % 5 + 2) % 5 + 5 * (i // 5)]) for i in range(25)]
        a[0] ^= _RC[rnd]
    return a


class Keccak256:
    RATE = 136

    def __init__(self):
        self.state = [0] * 25
        self.buf = b""

    def copy(self):
        k = Keccak256()
        k.state = list(self.state)
        k.buf = self.buf
        return k

    def _absorb_block(self, block):
        for i in range(self.RATE // 8):
            self.state[i] ^= int.from_bytes(block[8 * i : 8 * i + 8], "little")
        self.state = keccak_f1600(self.state)

    def update(self, data: bytes):
        self.buf += bytes(data)
        while len(self.buf) >= self.RATE:
            self._absorb_block(self.buf[: self.RATE])
            self.buf = self.buf[self.RATE :]

    def digest(self) -> bytes:
        k = self.copy()
        pad = bytearray(self.RATE - len(k.buf))
        pad[0] ^= 0x01
        pad[-1] ^= 0x80
        k._absorb_block(k.buf + bytes(pad))
        return b"".join(k.state[i].to_bytes(8, "little") for i in range(4))


def keccak256(data: bytes) -> bytes:
    k = Keccak256()
    k.update(data)
    return k.digest()


# ---------------------------------------------------------------- transcripts
class TranscriptError(Exception):
    """io::Error of the reference (a &'static str)."""


class _Common:
    def __init__(self, hash_kind):
        self.hash_kind = hash_kind
        if hash_kind == "blake2b":
            self.state = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
        elif hash_kind == "keccak":
            self.state = Keccak256()
            self.state.update(b"Halo2-Transcript")
        else:
            raise ValueError(hash_kind)
        self.squeezed = []  # every challenge in squeeze order (parity hook)

    def squeeze_challenge(self) -> int:
        self.state.update(bytes([PREFIX_CHALLENGE]))
        if self.hash_kind == "blake2b":
            result = self.state.copy().digest()
        else:
            lo = self.state.copy()
            hi = self.state.copy()
            lo.update(bytes([KECCAK_PREFIX_LO]))
            hi.update(bytes([KECCAK_PREFIX_HI]))
            result = lo.digest() + hi.digest()
        c = fr_from_uniform_bytes(result)
        self.squeezed.append(c)
        return c

    def common_point(self, pt):
        self.state.update(bytes([PREFIX_POINT]))
        if pt is None:
            raise TranscriptError("cannot write points at infinity to the transcript")
        self.state.update(fq_to_repr(pt[0]))
        self.state.update(fq_to_repr(pt[1]))

    def common_scalar(self, s):
        self.state.update(bytes([PREFIX_SCALAR]))
        self.state.update(fr_to_repr(s))


class TranscriptRead(_Common):
    def __init__(self, proof: bytes, hash_kind="blake2b"):
        super().__init__(hash_kind)
        self.proof = bytes(proof)
        self.pos = 0
        self.points = []  # every point read, in order (proof point slots)

    def _read_exact(self, n):
        if self.pos + n > len(self.proof):
            raise TranscriptError("failed to fill whole buffer")
        b = self.proof[self.pos : self.pos + n]
        self.pos += n
        return b

    def read_point(self):
        ok, pt = g1_from_bytes(self._read_exact(32))
        if not ok:
            raise TranscriptError("invalid point encoding in proof")
        self.common_point(pt)
        self.points.append(pt)
        return pt

    def read_scalar(self):
        s = fr_from_repr(self._read_exact(32))
        if s is None:
            raise TranscriptError("invalid field element encoding in proof")
        self.common_scalar(s)
        return s


class TranscriptWrite(_Common):
    def __init__(self, hash_kind="blake2b"):
        super().__init__(hash_kind)
        self.out = bytearray()

    def write_point(self, pt):
        self.common_point(pt)
        self.out += g1_to_bytes(pt)

    def write_scalar(self, s):
        self.common_scalar(s)
        self.out += fr_to_repr(s)
