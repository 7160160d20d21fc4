"""The C-ABI library loads and exports every symbol include/h2v.h declares; without a CUDA device the
product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "h2v.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(h2v_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built, pkg):
    lib = ctypes.CDLL(pkg.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/h2v.h but not exported"
    assert sorted(pkg.EXPORTED_SYMBOLS) == syms


def test_malformed_inputs_are_rejected_before_touching_the_device(built, pkg):
    import formats as F
    from workloads import setup

    params, vk, _dl, _s = setup("vm", 8)
    with pytest.raises(pkg.BackendError, match="truncated"):
        pkg.BatchVerifier(pkg.ParamsKZG.from_bytes(params.to_bytes()), pkg.VerifyingKey.from_bytes(vk.to_bytes(F.RAW_BYTES)[:-3]))


def test_no_cpu_fallback(built, pkg):
    import torch
    import formats as F
    from workloads import setup

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    params, vk, _dl, _s = setup("vm", 8)
    with pytest.raises(pkg.BackendError, match="no usable CUDA device"):
        pkg.BatchVerifier(pkg.ParamsKZG.from_bytes(params.to_bytes()), pkg.VerifyingKey.from_bytes(vk.to_bytes(F.RAW_BYTES)))
    with pytest.raises(pkg.BackendError):
        pkg.verify_proof(pkg.ParamsKZG.from_bytes(params.to_bytes()), pkg.VerifyingKey.from_bytes(vk.to_bytes(F.RAW_BYTES)), bytes(1024), [[1] * 10])


def test_vk_bundle_helpers(pkg):
    import formats as F
    from workloads import setup

    params, vk, _dl, _s = setup("vm", 8)
    bundle = params.to_bytes(F.PROCESSED) + vk.to_bytes(F.RAW_BYTES)
    P, V = pkg.read_vk_bundle(bundle)
    assert P.data == params.to_bytes(F.PROCESSED) and P.format == pkg.SerdeFormat.Processed and P.k == 8
    assert V.data == vk.to_bytes(F.RAW_BYTES) and V.format == pkg.SerdeFormat.RawBytes
    assert pkg.instances_from_pubs((5).to_bytes(32, "little") + (7).to_bytes(32, "little")) == [[5, 7]]
    with pytest.raises(pkg.BackendError):
        pkg.read_vk_bundle(bundle[:100])
