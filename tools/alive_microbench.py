"""Microbenchmark of the alive-table stamping (atomicMax on the 32 GiB direct-mapped table) under different
hash orders, through kta_alive_import_device.  Answers: how fast are the atomics when the touched table
slice is L2 resident, and how much locality does the bucket pass need?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kafka_topic_analyzer_b200 as kta

N = 100_000_000
D = 10_000_000
dev = torch.device("cuda", 0)
e = kta.KtaEngine(4, count_alive_keys=True, device=0)
e.set_stream(torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(1)
distinct = torch.randint(0, 2**32, (D,), device=dev, dtype=torch.int64, generator=g)
idx = torch.randint(0, D, (N,), device=dev, dtype=torch.int64, generator=g)
h_rand = distinct[idx]
stamps = ((torch.arange(N, device=dev, dtype=torch.int64) + 1) << 1) | 1

def run(name, h64):
    h32 = h64.to(torch.int32) if h64.max() < 2**31 else (h64 - (h64 >= 2**31).to(torch.int64) * 2**32).to(torch.int32)
    torch.cuda.synchronize()
    best = 1e9
    for it in range(3):
        e.reset()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        e.alive_import(h32, stamps, N)
        t1.record(); torch.cuda.synchronize()
        best = min(best, t0.elapsed_time(t1))
    e.finalize()
    print("%-42s %8.3f ms  %6.1f G stamps/s  alive=%d" % (name, best, N / best / 1e6, e.alive_keys()), flush=True)

run("random order (10M distinct over 2^32)", h_rand)
run("bucketed by top 9 bits, random inside", h_rand[torch.argsort(h_rand >> 23, stable=True)])
run("bucketed by top 12 bits", h_rand[torch.argsort(h_rand >> 20, stable=True)])
run("bucketed by top 16 bits", h_rand[torch.argsort(h_rand >> 16, stable=True)])
run("fully sorted by hash", torch.sort(h_rand)[0])
small = torch.randint(0, 2**20, (N,), device=dev, dtype=torch.int64, generator=g)
run("2^20 distinct contiguous hashes (8 MiB)", small)
run("2^20 distinct hashes spread over 2^32", small << 12)
