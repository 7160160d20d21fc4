"""ctypes wrapper of the C oracle (oracle/c, built by oracle/Makefile).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and the CPU arm of bench.py; never by the product."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libh2v_oracle.so")
_lib = None


def build():
    subprocess.run(["make", "-s", "-C", HERE], check=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = ctypes.CDLL(LIB_PATH)
    vp = ctypes.c_void_p
    lib.h2vo_selftest.restype = ctypes.c_int
    lib.h2vo_load.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(vp)]
    lib.h2vo_free.argtypes = [vp]
    lib.h2vo_free.restype = None
    lib.h2vo_error.argtypes = [vp]
    lib.h2vo_error.restype = ctypes.c_char_p
    lib.h2vo_verify.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, vp, ctypes.c_uint32, ctypes.c_int, ctypes.c_int,
                                ctypes.c_int, vp, vp, vp]
    lib.h2vo_verify_multi.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, vp, vp, vp]
    lib.h2vo_verify_many.argtypes = [vp, ctypes.c_uint32, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp,
                                     ctypes.c_uint32, ctypes.POINTER(ctypes.c_double)]
    lib.h2vo_fold.argtypes = [vp, ctypes.c_uint32, vp, vp, vp, vp, ctypes.POINTER(ctypes.c_int)]
    _lib = lib
    return lib


_MO = {"shplonk": 0, "gwc": 1}
_HK = {"blake2b": 0, "keccak": 1, "keccak256": 1}


class COracle:
    """verify_proof of the reference restated in C, for one (params, vk)."""

    def __init__(self, params_bytes: bytes, params_fmt: int, vk_bytes: bytes, vk_fmt: int):
        self.lib = load()
        self.h = ctypes.c_void_p()
        rc = self.lib.h2vo_load(params_bytes, len(params_bytes), params_fmt, vk_bytes, len(vk_bytes), vk_fmt, ctypes.byref(self.h))
        if rc != 0:
            msg = self.lib.h2vo_error(self.h).decode()
            self.lib.h2vo_free(self.h)
            self.h = None
            raise ValueError(msg)

    def close(self):
        if getattr(self, "h", None):
            self.lib.h2vo_free(self.h)
            self.h = None

    __del__ = close

    def verify(self, proof: bytes, instance_columns, multiopen="shplonk", hash_kind="blake2b", check_pairing=True):
        """instance_columns: [column][row] ints.  Returns (status, challenges, L|R bytes)."""
        inst = b"".join(int(v).to_bytes(32, "little") for col in instance_columns for v in col)
        cl = (ctypes.c_uint32 * max(1, len(instance_columns)))(*[len(c) for c in instance_columns])
        ch = ctypes.create_string_buffer(64 * 32)
        nch = ctypes.c_uint32(0)
        lr = ctypes.create_string_buffer(128)
        st = self.lib.h2vo_verify(self.h, proof, len(proof), inst, cl, len(instance_columns), _MO[multiopen], _HK[hash_kind],
                                  1 if check_pairing else 0, ch, ctypes.byref(nch), lr)
        chal = [int.from_bytes(ch.raw[32 * i: 32 * i + 32], "little") for i in range(nch.value)]
        return st, chal, lr.raw

    def verify_multi(self, proof: bytes, instances, multiopen="shplonk", hash_kind="blake2b", check_pairing=True):
        """A proof that carries len(instances) circuit instances (`instances`: [instance][column][row] ints, the
        `instances` argument of the reference's verify_proof).  Returns (status, challenges, L|R bytes)."""
        cols = [col for inst in instances for col in inst]
        inst = b"".join(int(v).to_bytes(32, "little") for col in cols for v in col)
        cl = (ctypes.c_uint32 * max(1, len(cols)))(*[len(c) for c in cols])
        ch = ctypes.create_string_buffer(64 * 32)
        nch = ctypes.c_uint32(0)
        lr = ctypes.create_string_buffer(128)
        st = self.lib.h2vo_verify_multi(self.h, proof, len(proof), inst, cl, len(cols), len(instances), _MO[multiopen], _HK[hash_kind],
                                        1 if check_pairing else 0, ch, ctypes.byref(nch), lr)
        chal = [int.from_bytes(ch.raw[32 * i: 32 * i + 32], "little") for i in range(nch.value)]
        return st, chal, lr.raw

    def verify_many(self, proofs_buf, poff, inst_buf, ioff, n, multiopen="shplonk", hash_kind="blake2b", check_pairing=True, threads=1,
                    want_lr=False, chal_cap=0):
        """Packed batch (same layout as the C ABI of the product: numpy arrays / bytes).  Returns (statuses, seconds, LR, chal)."""
        import numpy as np

        pb = np.frombuffer(proofs_buf, dtype=np.uint8) if not isinstance(proofs_buf, np.ndarray) else proofs_buf
        ib = np.frombuffer(inst_buf, dtype=np.uint8) if not isinstance(inst_buf, np.ndarray) else inst_buf
        po = np.ascontiguousarray(poff, dtype=np.uint64)
        io = np.ascontiguousarray(ioff, dtype=np.uint64)
        status = np.zeros(n, dtype=np.uint8)
        lr = np.zeros(128 * n, dtype=np.uint8) if want_lr else None
        ch = np.zeros(32 * n * chal_cap, dtype=np.uint8) if chal_cap else None
        secs = ctypes.c_double(0)
        ptr = lambda a: a.ctypes.data if a is not None and a.size else None
        self.lib.h2vo_verify_many(self.h, n, ptr(pb), ptr(po), ptr(ib), ptr(io), _MO[multiopen], _HK[hash_kind], 1 if check_pairing else 0,
                                  int(threads), ptr(status), ptr(lr), ptr(ch), chal_cap, ctypes.byref(secs))
        return status, secs.value, (lr.tobytes() if lr is not None else None), (ch.tobytes() if ch is not None else None)

    def fold(self, lr_bytes: bytes, rs, include):
        n = len(rs)
        rb = b"".join(int(r).to_bytes(32, "little") for r in rs)
        inc = bytes(1 if x else 0 for x in include)
        out = ctypes.create_string_buffer(128)
        verdict = ctypes.c_int(0)
        rc = self.lib.h2vo_fold(self.h, n, lr_bytes, rb, inc, out, ctypes.byref(verdict))
        assert rc == 0
        return out.raw, bool(verdict.value)
