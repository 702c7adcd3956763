"""Multi-GPU merge: one process per GPU, partitions sharded gpu = partition mod G (SURVEY.md §8 e),
no data-path collective during the scan, ONE NCCL all-reduce (SUM over u64) at the end for every
counter / histogram / extremum / HLL register, plus — only with -c — an all-gather of the compacted
alive-key stamps (NCCL has no OR / no 64-bit max over a 32 GiB table; last-writer-wins by global seq is
associative and commutative, so re-applying every rank's (hash, stamp) list on every rank is exact).

torch.distributed is plumbing here; the pack/unpack kernels are in csrc/kta_kernels.cuh."""
from __future__ import annotations

from .metrics import KtaEngine


def allreduce_merge(engine: KtaEngine, group=None, counters_only: bool = False) -> None:
    """After this, every rank's engine holds the merged state (call finalize() to read it).
    counters_only skips the exact alive-key exchange (used to time the all-reduce part alone)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return
    dev = torch.device("cuda", torch.cuda.current_device())
    words = engine.merge_words(world)
    buf = torch.empty(words, dtype=torch.int64, device=dev)   # u64 payload; SUM is bit-identical on i64
    # When the engine runs on torch's current stream (engine.set_stream) the three steps are ordered by that
    # stream alone — export kernel, NCCL all-reduce, import kernel — with no host synchronisation in between.
    engine.merge_export(rank, world, buf)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    if not engine.shares_caller_stream:
        torch.cuda.current_stream().synchronize()
    engine.merge_import(world, buf)
    if engine.count_alive_keys and not counters_only:
        n_local = engine.alive_export_count()
        counts = torch.zeros(world, dtype=torch.int64, device=dev)
        counts[rank] = n_local
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
        cap = int(counts.max().item())
        if cap == 0:
            return
        h_loc = torch.zeros(cap, dtype=torch.int32, device=dev)
        s_loc = torch.zeros(cap, dtype=torch.int64, device=dev)
        engine.alive_export(h_loc, s_loc, cap)
        h_all = torch.empty(world * cap, dtype=torch.int32, device=dev)
        s_all = torch.empty(world * cap, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(h_all, h_loc, group=group)
        dist.all_gather_into_tensor(s_all, s_loc, group=group)
        torch.cuda.current_stream().synchronize()
        cl = counts.tolist()
        for r in range(world):
            if r != rank and cl[r]:
                engine.alive_import(h_all[r * cap:], s_all[r * cap:], int(cl[r]))


# ---------------------------------------------------------------------------------------------------
# Host-side statement of the merge-buffer layout (what merge_export_kernel / merge_import_kernel do on the
# device, csrc/kta_kernels.cuh).  Used by the gloo CPU tests of the N>1 logic and usable for host-side merges.
#   [ sums (nsums u64) | world × 4 extrema slots | world × (nhll/8) words, eight one-byte registers per word ]
# Every rank fills only its own slots, so ONE SUM all-reduce hands every rank's values to every rank.
# ---------------------------------------------------------------------------------------------------
def merge_words(nsums: int, nhll: int, world: int) -> int:
    return nsums + world * 4 + world * (nhll // 8)


def pack_merge_buffer(sums, minmax, hll, rank: int, world: int):
    """sums u64[nsums]; minmax = (min_ts i64, max_ts i64, min_size u64, max_size u64); hll u32[nhll]."""
    import numpy as np
    nsums, nhll = len(sums), len(hll)
    buf = np.zeros(merge_words(nsums, nhll, world), dtype=np.uint64)
    buf[:nsums] = np.asarray(sums, dtype=np.uint64)
    mm = np.array([minmax[0], minmax[1]], dtype=np.int64).view(np.uint64)
    buf[nsums + 4 * rank: nsums + 4 * rank + 2] = mm
    buf[nsums + 4 * rank + 2] = np.uint64(minmax[2])
    buf[nsums + 4 * rank + 3] = np.uint64(minmax[3])
    if nhll:
        hw = nhll // 8
        regs = np.asarray(hll, dtype=np.uint8)
        o = nsums + 4 * world + rank * hw
        buf[o:o + hw] = regs.view(np.uint64) if regs.flags["C_CONTIGUOUS"] else np.ascontiguousarray(regs).view(np.uint64)
    return buf


def fold_merge_buffer(buf, nsums: int, nhll: int, world: int):
    """inverse of pack after the SUM all-reduce: returns (sums, (min_ts, max_ts, min_size, max_size), hll)."""
    import numpy as np
    sums = buf[:nsums].copy()
    mm = buf[nsums:nsums + 4 * world].reshape(world, 4)
    tmin = int(mm[:, 0].copy().view(np.int64).min())
    tmax = int(mm[:, 1].copy().view(np.int64).max())
    smin, smax = int(mm[:, 2].min()), int(mm[:, 3].max())
    hll = np.zeros(nhll, dtype=np.uint32)
    if nhll:
        hw = nhll // 8
        w = np.ascontiguousarray(buf[nsums + 4 * world:nsums + 4 * world + world * hw]).view(np.uint8).reshape(world, nhll)
        hll[:] = w.max(axis=0)
    return sums, (tmin, tmax, smin, smax), hll
