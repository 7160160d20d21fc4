"""Device-side exchange of sharded batches (include/h2v.h `h2v_comm_*`, csrc/exchange.cuh) against the CPU oracle.

* one GPU (always runs under `-m gpu`): two contexts of ONE process play two ranks (each driven by its own host thread,
  as two processes would); the windows are mapped directly;
* two or more GPUs (self-skips below 2): one PROCESS per GPU, windows mapped through CUDA IPC, handles shipped with
  torch.distributed - the product path of sharding.py - including a corrupted proof on a rank that is NOT the root
  (reference contract: poly/strategy.rs:26-30) and the library-gather path (NCCL gather to the root).
"""
import os
import sys
import threading

import pytest

import bn254 as bn
import formats as F
import prover_sim as sim
import verifier as orc
from workloads import enc_point, make_batch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_bv(pkg, params, vk, mo="shplonk", hk="blake2b", device=0):
    return pkg.BatchVerifier(pkg.ParamsKZG.from_bytes(params.to_bytes()), pkg.VerifyingKey.from_bytes(vk.to_bytes(F.RAW_BYTES), F.RAW_BYTES),
                             mo, hk, device=device)


def run_ranks(fns):
    """one host thread per rank (ctypes releases the GIL): the ranks' launch sets overlap as those of two processes do"""
    out, errs = [None] * len(fns), []

    def wrap(i):
        try:
            out[i] = fns[i]()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ths = [threading.Thread(target=wrap, args=(i,)) for i in range(len(fns))]
    [t.start() for t in ths]
    [t.join() for t in ths]
    if errs:
        raise errs[0]
    return out


def test_exchange_two_contexts_one_gpu(pkg):
    """gather to the root + verdict broadcast inside the graphs; folded (L, R) == oracle; rejected batch attributed
    on the non-root rank; roots alternate; fold groups; graph replay (no recapture after the first of each kind)"""
    os.environ.setdefault("H2V_COMM_TIMEOUT_MS", "5000")
    n, world = 64, 2
    per = n // world
    params, vk, instances, proofs, rng = make_batch("vm", 10, n, "shplonk", "blake2b", seed=61)
    insts = [i[0] for i in instances]
    rs = [rng.randrange(1, bn.R) for _ in range(n)]
    want = [orc.verify_proof(params, vk, inst, p) for inst, p in zip(instances, proofs)]
    L, R_, ok = orc.accumulate(params, want, rs)
    assert ok
    bvs = [make_bv(pkg, params, vk) for _ in range(world)]
    try:
        handles = [bv.comm_init(r, world, max_groups=3) for r, bv in enumerate(bvs)]
        [bv.comm_connect(handles) for bv in bvs]
        for root in (0, 1, 1, 0):
            res = run_ranks([lambda r=r: bvs[r].verify_shard(proofs[r * per:(r + 1) * per], insts[r * per:(r + 1) * per], r * per, n, root, rlc_scalars=rs)
                             for r in range(world)])
            assert all(v == [True] and st == [0] * per for v, st in res)
            assert bvs[root].comm_last_batch_accum() == enc_point(L) + enc_point(R_), "folded accumulators differ from the oracle's"
        caps = [bv.cache_stats()["graph_captures"] for bv in bvs]
        assert caps == [2, 2], caps  # one graph as root, one as non-root; the repeats replayed
        # a corrupted proof on rank 1 while rank 0 is the root: every rank learns the rejection, rank 1 attributes it
        bad = list(proofs)
        bad[per + 5], _ = sim.corrupt(proofs[per + 5], vk, "eval_flip", rng)
        bad[3], _ = sim.corrupt(proofs[3], vk, "point_offcurve", rng)  # a transcript error on rank 0: excluded from the fold
        want_bad = [orc.verify_proof(params, vk, inst, p).status for inst, p in zip(instances, bad)]
        res = run_ranks([lambda r=r: bvs[r].verify_shard(bad[r * per:(r + 1) * per], insts[r * per:(r + 1) * per], r * per, n, 0, rlc_scalars=rs)
                         for r in range(world)])
        assert [v for v, _ in res] == [[False], [False]]
        assert res[0][1] + res[1][1] == want_bad and want_bad[per + 5] == orc.CONSTRAINT_SYSTEM_FAILURE
        # fold groups: 3 global batches of 16 per launch set, the middle one holds the bad proof (on rank 1)
        G, gn = 3, 16
        gp = gn // world
        pr = list(proofs[:3 * gn])
        pr[gn + gp + 2] = sim.corrupt(pr[gn + gp + 2], vk, "eval_flip", rng)[0]
        gi = insts[:3 * gn]
        rs3 = [rng.randrange(1, bn.R) for _ in range(G * gn)]
        sel = lambda r: [q * gn + r * gp + j for q in range(G) for j in range(gp)]
        res = run_ranks([lambda r=r: bvs[r].verify_shard([pr[i] for i in sel(r)], [gi[i] for i in sel(r)], r * gp, gn, 1, rlc_scalars=rs3, fold_groups=G)
                         for r in range(world)])
        assert [v for v, _ in res] == [[True, False, True]] * 2
        st = {i: s for r in range(world) for i, s in zip(sel(r), res[r][1])}
        assert [st[i] for i in range(G * gn)] == [4 if i == gn + gp + 2 else 0 for i in range(G * gn)]
        # fresh OS entropy per call: accepted, and two calls fold with different coefficients
        accs = []
        for _ in range(2):
            res = run_ranks([lambda r=r: bvs[r].verify_shard(proofs[r * per:(r + 1) * per], insts[r * per:(r + 1) * per], r * per, n, 0) for r in range(world)])
            assert all(v == [True] for v, _ in res) and bvs[0].rlc_source() == "os"
            accs.append(bvs[0].comm_last_batch_accum())
        assert accs[0] != accs[1]
    finally:
        [bv.close() for bv in bvs]


def test_exchange_timeout_is_an_error_not_a_hang(pkg):
    """a rank that never shows up: the root's wait is bounded and surfaces as a BackendError (never 'accepted')"""
    n, world = 8, 2
    params, vk, instances, proofs, rng = make_batch("vm", 8, n, "shplonk", "blake2b", seed=62)
    insts = [i[0] for i in instances]
    old = os.environ.get("H2V_COMM_TIMEOUT_MS")
    os.environ["H2V_COMM_TIMEOUT_MS"] = "300"
    bvs = [make_bv(pkg, params, vk) for _ in range(world)]
    try:
        handles = [bv.comm_init(r, world) for r, bv in enumerate(bvs)]
        [bv.comm_connect(handles) for bv in bvs]
        with pytest.raises(pkg.BackendError, match="timed out"):
            bvs[0].verify_shard(proofs[:4], insts[:4], 0, n, 0, seed=5)  # rank 1 never runs
    finally:
        if old is None:
            os.environ.pop("H2V_COMM_TIMEOUT_MS", None)
        else:
            os.environ["H2V_COMM_TIMEOUT_MS"] = old
        [bv.close() for bv in bvs]


def test_default_fold_randomness_is_os_entropy(pkg):
    """no public default coefficients - two default calls fold the same batch differently; seeds / scalars
    remain available as explicitly named parity hooks"""
    params, vk, instances, proofs, rng = make_batch("vm", 8, 6, "shplonk", "blake2b", seed=63)
    insts = [i[0] for i in instances]
    with make_bv(pkg, params, vk) as bv:
        a = bv.verify_batch(proofs, insts, want_batch_accum=True)
        assert bv.rlc_source() == "os"
        b = bv.verify_batch(proofs, insts, want_batch_accum=True)
        assert a.verdict and b.verdict and a.batch_accum != b.batch_accum
        c = bv.verify_batch(proofs, insts, seed=9, want_batch_accum=True)
        d = bv.verify_batch(proofs, insts, seed=9, want_batch_accum=True)
        assert bv.rlc_source() == "seed" and c.batch_accum == d.batch_accum
        key = bytes(range(32))
        e = bv.verify_batch(proofs, insts, key=key, want_batch_accum=True)
        f = bv.verify_batch(proofs, insts, key=key, want_batch_accum=True)
        assert bv.rlc_source() == "key" and e.batch_accum == f.batch_accum != c.batch_accum
        with pytest.raises(ValueError):
            bv.verify_batch(proofs, insts, seed=0)
        assert pkg.verify_proofs_batch(pkg.ParamsKZG.from_bytes(params.to_bytes()), pkg.VerifyingKey.from_bytes(vk.to_bytes(F.RAW_BYTES), F.RAW_BYTES),
                                       proofs, insts) == [None] * 6


def test_shape_keyed_caches(pkg):
    """alternating batch shapes on one context replays cached graphs and prepared window lines"""
    params, vk, instances, proofs, rng = make_batch("vm", 10, 32, "shplonk", "blake2b", seed=64)
    proofs, insts = proofs * 128, [i[0] for i in instances] * 128  # 4096
    with make_bv(pkg, params, vk) as bv:
        shapes = [(4096, 1), (1024, 1), (4096, 8)]
        for n, G in shapes * 2:  # second pass: a graph captured before a later shape grew a device buffer is captured once more
            assert bv.verify_batch(proofs[:n], insts[:n], fold_groups=G).verdict
        first = bv.cache_stats()
        for i in range(99):
            n, G = shapes[i % 3]
            assert bv.verify_batch(proofs[:n], insts[:n], fold_groups=G).verdict
        assert bv.cache_stats() == first, (first, bv.cache_stats())
        assert first["graph_captures"] <= 6 and first["lines_builds"] <= 3


# ---------------------------------------------------------------------------------------------------------------------
# one process per GPU
# ---------------------------------------------------------------------------------------------------------------------
def _rank_main(rank, world, port, q, use_exchange):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), H2V_COMM_TIMEOUT_MS="20000")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pkg = g.load_package()
    sharding = __import__("importlib").import_module("halo2_verifier_b200.sharding")
    try:
        n = 32 * world
        params, vk, instances, proofs, rng = make_batch("vm", 10, n, "shplonk", "blake2b", seed=71)
        insts = [i[0] for i in instances]
        rs = [rng.randrange(1, bn.R) for _ in range(n)]
        bv = make_bv(pkg, params, vk, device=rank)
        if use_exchange:
            assert sharding.connect_channel(bv, rank, world, max_groups=2), "CUDA IPC mapping of the peers' windows failed"
        lo, hi = sharding.shard_range(n, rank, world)
        # clean batch, every rank takes the root role once
        for root in range(world):
            ok, status = sharding.verify_batch_sharded(bv, proofs, insts, rank, world, rlc_scalars=rs, root=root)
            assert ok and status == [0] * (hi - lo)
            if use_exchange and rank == root:
                want = [orc.verify_proof(params, vk, inst, p) for inst, p in zip(instances, proofs)]
                L, R_, okk = orc.accumulate(params, want, rs)
                assert okk and bv.comm_last_batch_accum() == enc_point(L) + enc_point(R_)
        # default randomness: a fresh key from rank 0
        ok, status = sharding.verify_batch_sharded(bv, proofs, insts, rank, world)
        assert ok and bv.rlc_source() == "key"
        # one corrupted proof on the LAST rank, root = rank 0
        bad = list(proofs)
        j = n - 3
        bad[j], _ = sim.corrupt(proofs[j], vk, "eval_flip", rng)
        ok, status = sharding.verify_batch_sharded(bv, bad, insts, rank, world, rlc_scalars=rs, root=0)
        want_st = [orc.verify_proof(params, vk, inst, p).status for inst, p in zip(instances[lo:hi], bad[lo:hi])]
        assert not ok and status == want_st
        assert (rank == world - 1) == (orc.CONSTRAINT_SYSTEM_FAILURE in status)
        # two global batches per launch set (fold groups), the second one rejected, root = last rank
        half = n // 2
        batches = [(proofs[:half], insts[:half]), (bad[half:], insts[half:])]
        verdicts, sts = sharding.verify_batches_sharded(bv, batches, rank, world, seed=11, root=world - 1)
        assert verdicts == [True, False]
        l2, h2 = sharding.shard_range(half, rank, world)
        assert sts[0] == [0] * (h2 - l2)
        assert sts[1] == [orc.verify_proof(params, vk, instances[half + i], bad[half + i]).status for i in range(l2, h2)]
        bv.close()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback

        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("use_exchange", [True, False], ids=["device_exchange", "library_gather"])
def test_sharded_processes_multi_gpu(pkg, use_exchange):
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 200 + (7 if use_exchange else 0)
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, q, use_exchange)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=600) for _ in procs]
    [p.join(timeout=120) for p in procs]
    assert all(r[1] == "ok" for r in res), res
