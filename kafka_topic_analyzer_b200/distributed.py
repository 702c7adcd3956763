"""Multi-GPU merge: one process per GPU, partitions sharded gpu = partition mod G (SURVEY.md §8 e),
no data-path collective during the scan, ONE NCCL all-reduce (SUM over u64) at the end for every
counter / histogram / extremum / HLL register, plus — only with -c — an all-gather of the compacted
alive-key stamps (NCCL has no OR / no 64-bit max over a 32 GiB table; last-writer-wins by global seq is
associative and commutative, so re-applying every rank's (hash, stamp) list on every rank is exact).

torch.distributed is plumbing here; the pack/unpack kernels are in csrc/kta_kernels.cuh."""
from __future__ import annotations

from .metrics import KtaEngine


def allreduce_merge(engine: KtaEngine, group=None) -> None:
    """After this, every rank's engine holds the merged state (call finalize() to read it)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return
    dev = torch.device("cuda", torch.cuda.current_device())
    words = engine.merge_words(world)
    buf = torch.empty(words, dtype=torch.int64, device=dev)   # u64 payload; SUM is bit-identical on i64
    engine.merge_export(rank, world, buf)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    torch.cuda.current_stream().synchronize()
    engine.merge_import(world, buf)
    if engine.count_alive_keys:
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
