"""Profiling driver (ncu): ONE 4096-proof batch alone (latency view) and launch sets of 16 fold groups (throughput view),
direct launches (no graph), nothing else in between.  Every batch kernel is launched inside the NVTX range
"h2v:launch_set", so `ncu --nvtx --nvtx-include "h2v:launch_set/"` sees exactly these kernels:
    launches   0..15   the workload generator's own pass (placeholder proofs)
    launches  16..63   three single batches (G = 1)
    launches  64..     two launch sets of 16 fold groups
Usage: python tools/prof_sets.py   (N = 1; never run the multi-GPU exchange under ncu: it serialises kernels)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
from importlib import import_module
import bench


def main():
    pkg = g.load_package()
    synth = import_module("halo2_verifier_b200.synth")
    k, n, G = 10, 4096, 16
    s = bench.srs_secret(k)
    vk_bytes, shared_dlogs = synth.make_vk_bytes("vm", k)
    params = pkg.ParamsKZG.from_bytes(synth.params_bytes_raw(k, s), pkg.SerdeFormat.RawBytes)
    vk = pkg.VerifyingKey.from_bytes(vk_bytes, pkg.SerdeFormat.RawBytes)
    bv = pkg.BatchVerifier(params, vk, "shplonk", "blake2b", device=0)
    bv.set_graphs(False)
    proofs, instances = synth.synthesize_shplonk_batch(bv, shared_dlogs, s, n, seed=("prof", 0))
    for _ in range(3):
        assert bv.verify_batch(proofs, instances, seed=7).verdict
    for _ in range(2):
        assert bv.verify_batch(proofs * G, instances * G, seed=7, fold_groups=G).verdict
    print("geometry", bv.msm_geometry(), "launches", bv.launch_count())
    bv.close()


if __name__ == "__main__":
    main()
