"""Host-side logic of the multi-GPU path on CPU: two gloo ranks shard one batch, each folds its shard
with GLOBALLY defined coefficients, partial accumulators are all-gathered and added (the device path
does the same with NCCL; the folding itself is done by the oracle here because there is no GPU)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rank_main(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import bn254 as bn
    import verifier as orc
    from workloads import enc_point, make_batch
    import __graft_entry__ as g

    pkg = g.load_package()
    sharding = __import__("importlib").import_module("halo2_verifier_b200.sharding")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 12
    params, vk, instances, proofs, rng = make_batch("vm", 8, n, "shplonk", "blake2b", seed=21)
    rs = [rng.randrange(1, bn.R) for _ in range(n)]
    lo, hi = sharding.shard_range(n, rank, world)
    results = [orc.verify_proof(params, vk, instances[j], proofs[j], check_pairing=False) for j in range(lo, hi)]
    cs = orc.rlc_coefficients(rs)  # GLOBAL coefficients
    L = R_ = None
    for w, c in zip(results, cs[lo:hi]):
        L, R_ = bn.g1_add(L, bn.g1_mul(w.L, c)), bn.g1_add(R_, bn.g1_mul(w.R, c))
    partial = torch.frombuffer(bytearray(enc_point(L) + enc_point(R_)), dtype=torch.uint8)
    parts = sharding.all_gather_partials(partial, world)
    assert len(parts) == world and all(len(p) == 128 for p in parts)
    if rank == 0:
        dec = lambda e: None if e == bytes(64) else (int.from_bytes(e[:32], "little"), int.from_bytes(e[32:], "little"))
        Lt = Rt = None
        for p in parts:
            Lt, Rt = bn.g1_add(Lt, dec(p[:64])), bn.g1_add(Rt, dec(p[64:]))
        all_res = [orc.verify_proof(params, vk, instances[j], proofs[j], check_pairing=False) for j in range(n)]
        Lw, Rw, ok = orc.accumulate(params, all_res, rs)
        q.put((Lt == Lw and Rt == Rw, ok))
    dist.barrier()
    dist.destroy_process_group()


def _rank_groups(rank, world, port, q):
    """fold groups across shards: G global batches, one [group]-blob per rank, [rank][group] after the all-gather"""
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import bn254 as bn
    import verifier as orc
    from workloads import enc_point, make_batch
    import __graft_entry__ as g

    g.load_package()
    sharding = __import__("importlib").import_module("halo2_verifier_b200.sharding")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G, n = 2, 6
    params, vk, instances, proofs, rng = make_batch("vm", 8, G * n, "shplonk", "blake2b", seed=22)
    rs = [rng.randrange(1, bn.R) for _ in range(G * n)]
    lo, hi = sharding.shard_range(n, rank, world)
    blob = b""
    for grp in range(G):  # this rank's shard of global batch grp, coefficients restart in every batch
        cs = orc.rlc_coefficients(rs[grp * n:(grp + 1) * n])
        L = R_ = None
        for j in range(lo, hi):
            w = orc.verify_proof(params, vk, instances[grp * n + j], proofs[grp * n + j], check_pairing=False)
            L, R_ = bn.g1_add(L, bn.g1_mul(w.L, cs[j])), bn.g1_add(R_, bn.g1_mul(w.R, cs[j]))
        blob += enc_point(L) + enc_point(R_)
    parts = sharding.all_gather_partials(torch.frombuffer(bytearray(blob), dtype=torch.uint8), world)
    if rank == 0:
        dec = lambda e: None if e == bytes(64) else (int.from_bytes(e[:32], "little"), int.from_bytes(e[32:], "little"))
        same = True
        for grp in range(G):
            Lt = Rt = None
            for p in parts:
                gp = sharding.split_group_partials(p, G)[grp]
                Lt, Rt = bn.g1_add(Lt, dec(gp[:64])), bn.g1_add(Rt, dec(gp[64:]))
            res = [orc.verify_proof(params, vk, instances[grp * n + j], proofs[grp * n + j], check_pairing=False) for j in range(n)]
            Lw, Rw, ok = orc.accumulate(params, res, rs[grp * n:(grp + 1) * n])
            same = same and ok and Lt == Lw and Rt == Rw
        q.put((same, True))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_fold_groups_layout():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_rank_groups, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    same, ok = q.get(timeout=300)
    [p.join(timeout=60) for p in procs]
    assert same and ok
    assert all(p.exitcode == 0 for p in procs)


def test_two_rank_sharded_fold_equals_whole_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    same, ok = q.get(timeout=300)
    [p.join(timeout=60) for p in procs]
    assert same and ok
    assert all(p.exitcode == 0 for p in procs)


def test_shard_range_partition():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    g.load_package()
    sharding = __import__("importlib").import_module("halo2_verifier_b200.sharding")
    for n in (1, 7, 4096, 65536):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
