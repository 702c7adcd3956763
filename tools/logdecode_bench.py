"""Throughput of the GPU RecordBatch v2 decoder on the synthetic topic stored broker-style.
Segments are encoded on the host (C++), staged to HBM once, then decoded + scanned from device memory."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import kafka_topic_analyzer_b200 as kta
from kafka_topic_analyzer_b200 import synth
from kafka_topic_analyzer_b200._native import lib, check

P, N, VM = 16, 8_000_000, int(sys.argv[1]) if len(sys.argv) > 1 else 256
BR = int(sys.argv[2]) if len(sys.argv) > 2 else 56   # ~16 KB batches (the producer default batch.size) at 256 B values
spec = synth.make_spec(N, P, value_mean=VM, distinct_keys=1_000_000)
segs, offs = [], []
t0 = time.time()
for p in range(P):
    s = synth.encode_segment(spec, p, batch_records=BR)
    # batch offsets by hopping headers on the host
    o, pos = [], 0
    while pos + 61 <= s.size:
        o.append(pos)
        pos += 12 + int.from_bytes(s[pos + 8:pos + 12].tobytes(), "big", signed=True)
    segs.append(torch.from_numpy(s).cuda())
    offs.append(torch.tensor(o, dtype=torch.int64).cuda())
raw = sum(int(s.numel()) for s in segs)
print("encoded %d records, %.2f GB raw log in %.1f s" % (N, raw / 1e9, time.time() - t0), flush=True)
for mode, kw in (("counters", {}), ("fused HLL", dict(hll_precision=14)), ("-c exact", dict(count_alive_keys=True))):
    e = kta.KtaEngine(P, **kw)
    best = 1e9
    for it in range(4):
        e.reset(); e.sync()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = C.c_int64(); tot = 0
        for p in range(P):
            check(lib().kta_scan_log_segment_device(e.handle, p, segs[p].data_ptr(), segs[p].numel(), offs[p].data_ptr(),
                                                    offs[p].numel(), C.byref(n)))
            tot += n.value   # (one call per segment: three small host round trips each)
        e.finalize()
        best = min(best, time.perf_counter() - t0)
    assert tot == N and e.message_metrics.overall_count() == N
    print("%-10s decode+scan %.3f ms  %.2e msg/s  %.0f GB/s of raw log (value mean %d B)" % (mode, best * 1e3, N / best, raw / best / 1e9, VM), flush=True)
    e.close()
