"""Multi-GPU plumbing (one process per GPU): contiguous proof shards, globally defined RLC
coefficients, and an all-gather of the partial accumulators (NCCL on GPUs, gloo in the CPU tests).
A partial is the shard's per-window bucket sums (H2V_PARTIAL_BYTES, include/h2v.h).  NCCL cannot
reduce with an elliptic-curve addition, hence gather-then-add: rank 0 adds the partials window-wise
and runs the single pairing check (`h2v_finalize`).  SURVEY.md section 8(e)."""
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of global proof indices owned by `rank`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_partials(partial: torch.Tensor, world: int) -> List[bytes]:
    """partial: uint8[H2V_PARTIAL_BYTES] of this rank, on the device of the process group's backend.
    Returns every rank's partial as bytes, rank order."""
    assert partial.dtype == torch.uint8 and partial.dim() == 1
    if world == 1:
        return [bytes(partial.cpu().numpy().tobytes())]
    out = [torch.empty_like(partial) for _ in range(world)]
    dist.all_gather(out, partial)
    return [bytes(t.cpu().numpy().tobytes()) for t in out]


def verify_batch_sharded(bv, proofs, instances, rank: int, world: int, rlc_scalars=None, seed=0):
    """Every rank calls this with the WHOLE batch description (or at least its own shard's data at
    the right indices); returns (global verdict, statuses of this rank's shard)."""
    n = len(proofs)
    lo, hi = shard_range(n, rank, world)
    hint = shard_range(n, 0, world)[1]  # largest shard: common window geometry on every rank
    status, partial = bv.accumulate_shard(proofs[lo:hi], instances[lo:hi], lo, n, rlc_scalars=rlc_scalars, seed=seed, shard_hint=hint)
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    t = torch.frombuffer(bytearray(partial), dtype=torch.uint8).to(dev)
    parts = all_gather_partials(t, world)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    if rank == 0:
        ok, _ = bv.finalize(parts, want_batch_accum=False)
        flag[0] = 1 if ok else 0
    if world > 1:
        dist.broadcast(flag, src=0)
    ok = bool(flag.item())
    if not ok:  # rejected fold: every rank attributes inside its own shard, no further exchange
        status = bv.attribute_shard(status)
    return ok, status


def split_group_partials(rank_blob: bytes, groups: int) -> List[bytes]:
    """One rank's output of a launch set with `groups` fold groups = `groups` consecutive partials."""
    assert len(rank_blob) % groups == 0
    step = len(rank_blob) // groups
    return [rank_blob[q * step:(q + 1) * step] for q in range(groups)]


def verify_batches_sharded(bv, batches, rank: int, world: int, rlc_scalars=None, seed=0):
    """`batches`: G global batches [(proofs, instances), ...] of EQUAL size n, each sharded over the ranks; every rank
    processes its shard of all G batches in ONE set of kernel launches (fold groups, include/h2v.h), the G partials per
    rank are all-gathered as one blob ([rank][group]) and rank 0 runs the G pairing checks together.  `rlc_scalars`:
    G * n coefficients (batch-major) or None.  Returns (G verdicts, statuses of this rank's shard of every batch)."""
    G, n = len(batches), len(batches[0][0])
    assert all(len(p) == n and len(i) == n for p, i in batches) and n % world == 0, "equal batches, equal shards"
    lo, hi = shard_range(n, rank, world)
    proofs = [p for pr, _ in batches for p in pr[lo:hi]]
    insts = [i for _, ins in batches for i in ins[lo:hi]]
    status, blob = bv.accumulate_shard(proofs, insts, lo, n, rlc_scalars=rlc_scalars, seed=seed, fold_groups=G)
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    parts = all_gather_partials(torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev), world)
    flags = torch.zeros(G, dtype=torch.int32, device=dev)
    if rank == 0:
        flags[:] = torch.tensor([1 if v else 0 for v in bv.finalize_groups(parts, G)], dtype=torch.int32)
    if world > 1:
        dist.broadcast(flags, src=0)
    verdicts = [bool(v) for v in flags.tolist()]
    if not all(verdicts):  # attribution inside this rank's shards (proofs of accepted batches stay accepted)
        status = bv.attribute_shard(status)
    return verdicts, [status[q * (hi - lo):(q + 1) * (hi - lo)] for q in range(G)]
