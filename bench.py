#!/usr/bin/env python3
"""bench.py — messages/s of the message-scan metric path on N B200s (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm   (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (oracle port) on host cores

A "step" is one pass of the hot path over one batch of the synthetic topic:
  reset state -> fused scan kernel over the rank's shard resident in HBM -> (N>1: one NCCL all-reduce
  merge) -> finalize (state to host).  Workload = BASELINE configs[1]: 64 partitions, 1e8 messages per GPU
  (weak scaling), 256 B mean value, 16-byte keys, counters + histograms + FNV32 per key + HLL sketch.
Inputs are 3.6 GB per GPU (>> 126 MB L2), so every step streams from HBM (no L2 flush needed).
`value` = all ranks' records / max-over-ranks device time (CUDA events on the scan stream).
`e2e`   = the same metric through the C-ABI host entry point (kta_push_batch_host) with the batch in pinned
          HOST memory: host->device copies and the state read-back are inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "messages/sec scanned (fused metric scan)"
UNIT = "msg/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="fused", choices=["fused", "counters", "alive"])
    ap.add_argument("--n", type=int, default=100_000_000, help="records per GPU")
    ap.add_argument("--partitions", type=int, default=64)
    ap.add_argument("--value-mean", type=int, default=256)
    ap.add_argument("--run-len", type=int, default=1)
    ap.add_argument("--distinct-keys", type=int, default=10_000_000)
    ap.add_argument("--hll", type=int, default=14)
    ap.add_argument("--key-mode", type=int, default=0, help="0 = 16-byte binary keys, 1 = ASCII key-<id>, 2 = variable 0..40 B")
    ap.add_argument("--zipf-keys", action="store_true", help="stress case: log-uniform (Zipf s = 1 staircase) key ids")
    ap.add_argument("--geometric-values", action="store_true", help="stress case: geometric-tailed value lengths")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="records in the CPU baseline sample (0 = auto)")
    return ap.parse_args()


def workload_name(a, world):
    keys = {0: "16 B", 1: "ASCII key-<id>", 2: "variable 0..40 B"}[a.key_mode & 0xFF] + " keys"
    if a.zipf_keys:
        keys += " (log-uniform ids)"
    if a.geometric_values:
        keys += ", geometric value tail"
    return ("C1 %d partitions, %.0e msgs/GPU x %d GPU, %d B mean value, %s, mode=%s "
            "(counters+histograms%s), run_len=%d; inputs %.1f GB/GPU > L2, no flush needed" %
            (a.partitions, a.n, world, a.value_mean, keys, a.mode,
             {"fused": "+FNV32+HLL p%d" % a.hll, "counters": "", "alive": "+FNV32+exact alive-key table"}[a.mode],
             a.run_len, a.n * 36 / 1e9))


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 9 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 9 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path (oracle port; the Rust original cannot be
# built here).  bench.py is one of the few places allowed to execute oracle/.
# --------------------------------------------------------------------------------------------------
def cpu_reference_rate(a, sample, threads, count_alive_keys, topic=None):
    """records/s of the C restatement of src/metric.rs:206-305 over `sample` records of the workload.
    threads == 1 is the reference as designed (one consumer thread, src/kafka.rs:92-135); threads > 1 runs
    one independent handler set per thread over an equal slice (what a partition-sharded rewrite could do)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from kafka_topic_analyzer_b200 import synth
    from oracle_lib import Oracle
    if topic is None:
        spec = synth.make_spec(a.partitions * a.run_len * max(1, sample // (a.partitions * a.run_len)), a.partitions,
                               run_len=a.run_len, distinct_keys=min(a.distinct_keys, max(a.partitions, sample // 10)),
                               value_mean=a.value_mean,
                               key_mode=a.key_mode, zipf_keys=a.zipf_keys, geometric_values=a.geometric_values)
        topic = synth.fill_host(spec)
    n = topic.n
    import numpy as np
    kl = np.maximum(topic.key_len.astype(np.int64), 0)
    koff = np.concatenate([[0], np.cumsum(kl)])
    bounds = [n * i // threads for i in range(threads + 1)]
    oracles = [Oracle(count_alive_keys=count_alive_keys, no_hist=True) for _ in range(threads)]

    def work(i):
        lo, hi = bounds[i], bounds[i + 1]
        oracles[i].handle_batch(topic.partition[lo:hi], topic.ts_ms[lo:hi], topic.key_len[lo:hi], topic.value_len[lo:hi],
                                topic.key_bytes[int(koff[lo]):int(koff[hi])])

    t0 = time.perf_counter()
    if threads == 1:
        work(0)
    else:
        ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    dt = time.perf_counter() - t0
    return n / dt, n, topic


def run_reference(a):
    """--impl reference: the reference's own CPU path.  It is single-threaded by construction (one consumer
    thread drives both handlers, src/kafka.rs:92-135; the alive-key BitSet is one global structure,
    src/metric.rs:262-264), so "all the host threads it can use" is 1; a partition-sharded multi-thread
    figure for the counter handler alone is added for context."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    alive = a.mode != "counters"
    probe_n = 1_000_000
    rate, _, _ = cpu_reference_rate(a, probe_n, 1, alive)
    # size the per-step sample so that warmup + steps take about two minutes in total
    sample = a.cpu_sample or int(min(20_000_000, max(1_000_000, rate * 120.0 / max(1, a.steps + a.warmup))))
    _, _, topic = cpu_reference_rate(a, sample, 1, alive)
    sample = topic.n
    for _ in range(max(0, a.warmup - 1)):
        cpu_reference_rate(a, sample, 1, alive, topic)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cpu_reference_rate(a, sample, 1, alive, topic)
    total = time.perf_counter() - t0
    value = sample * a.steps / total
    sharded = cpu_reference_rate(a, sample, cores, False, topic)[0]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * total / max(1, a.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(a, a.gpus)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": "%d synthetic records of the workload shape per step (fresh handlers each step); C "
                                   "restatement of src/metric.rs:206-305 + src/fnv32.rs (%s); 1 thread because the "
                                   "reference is single-threaded by construction; Rust original not buildable here"
                                   % (sample, "MessageMetrics + LogCompactionInMemoryMetrics (-c)" if alive else
                                      "MessageMetrics only"),
                         "counters_only_sharded_value": sharded, "counters_only_sharded_cores": cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    import kafka_topic_analyzer_b200 as kta
    from kafka_topic_analyzer_b200 import synth
    from kafka_topic_analyzer_b200.distributed import allreduce_merge

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE %d" % (a.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- the rank's shard of the topic, generated in HBM ----
    P = a.partitions
    n_total = a.n * world
    spec = synth.make_spec(n_total, P, run_len=a.run_len, distinct_keys=a.distinct_keys * world, value_mean=a.value_mean,
                           key_mode=a.key_mode, zipf_keys=a.zipf_keys, geometric_values=a.geometric_values)
    exact = a.mode == "alive"
    topic = synth.DeviceTopic(spec, rank=rank, world=world, device=local, with_seq=(exact and world > 1))
    n = topic.n
    alg_bytes = 20 * n + (topic.key_bytes_len if a.mode != "counters" else 0)

    eng = kta.KtaEngine(P, count_alive_keys=exact, hll_precision=a.hll if a.mode == "fused" else 0, device=local)
    stream = torch.cuda.current_stream(dev)
    eng.set_stream(stream.cuda_stream)

    def step():
        eng.reset()
        eng.scan_batch_device(topic.partition, topic.ts_ms, topic.key_len, topic.value_len,
                              key_bytes=topic.key_bytes if a.mode != "counters" else None,
                              key_bytes_len=topic.key_bytes_len if a.mode != "counters" else 0,
                              key_tile_base=topic.key_tile_base if a.mode != "counters" else None, seq=topic.seq)
        if world > 1:
            allreduce_merge(eng)
        eng.finalize()

    for _ in range(a.warmup):
        step()
    barrier()
    eng.set_timing(True)
    l0 = eng.stats()[0]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    barrier()
    e0.record(stream)
    for _ in range(a.steps):
        step()
        launches += eng.stats()[0]      # reset() zeroes the counter each step
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    kern_ms, kern_n = eng.scan_time_ms()
    eng.set_timing(False)
    t = torch.tensor([ms, kern_ms / max(1, kern_n)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, kern_avg_ms = t.tolist()
    value = n_total * a.steps / (ms / 1e3)

    merge_ms = None
    if world > 1:
        # the merge alone (export kernel + ONE NCCL all-reduce + import kernel), 20 back-to-back calls
        barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record(stream)
        for _ in range(20):
            allreduce_merge(eng)
        m1.record(stream)
        barrier()
        mt = torch.tensor([m0.elapsed_time(m1) / 20], dtype=torch.float64, device=dev)
        dist.all_reduce(mt, op=dist.ReduceOp.MAX)
        merge_ms = mt.item()
        step()   # leave the engine holding one clean, merged pass again

    # sanity: the result of the last step is the whole topic
    mm = eng.message_metrics
    assert mm.overall_count() == n_total, (mm.overall_count(), n_total)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (kern_avg_ms / 1e3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(a.mode)
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(a, world), "partitions": P, "records_per_gpu": n,
                   "mean_key_bytes": topic.key_bytes_len / n, "l2": "inputs larger than L2"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic,
                     "kernel": "kta::scan_kernel<MODE_%s>" % {"fused": "HLL", "counters": "COUNTERS", "alive": "EXACT"}[a.mode],
                     "kernel_ms": kern_avg_ms, "algorithmic_bytes_per_launch": alg_bytes,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650"},
        "logical_topic_gb_s": value * (topic.key_bytes_len / n + a.value_mean) / 1e9,
        "gpu_launches": launches, "clocks": clocks,
    }
    if merge_ms is not None:
        line["merge_ms"] = merge_ms

    if rank == 0 and world == 1 and not a.no_extra:
        line["extra_modes"] = extra_modes(a, topic, kta, torch, dev, peak)

    # ---- e2e: same metric through the host entry point, inputs in pinned host memory ----
    if not a.no_e2e:
        e2e_steps = max(3, min(a.steps, 8))
        cols = {}
        for name in ("partition", "ts_ms", "key_len", "value_len", "key_tile_base"):
            src = getattr(topic, name)
            cols[name] = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            cols[name].copy_(src)
        kb = torch.empty(topic.key_bytes_len, dtype=torch.uint8, pin_memory=True)
        kb.copy_(topic.key_bytes[: topic.key_bytes_len])
        use_keys = a.mode != "counters"
        h2d = sum(cols[c].numel() * cols[c].element_size() for c in ("partition", "ts_ms", "key_len", "value_len"))
        if use_keys:
            h2d += kb.numel() + cols["key_tile_base"].numel() * 8
        d2h = (P * 67 + 1) * 8 + 32 + ((1 << a.hll) * 4 if a.mode == "fused" else 0) + (8 if exact else 0)

        def e2e_step():
            eng.reset()
            eng.push_batch_host(cols["partition"], cols["ts_ms"], cols["key_len"], cols["value_len"],
                                kb if use_keys else None, cols["key_tile_base"] if use_keys else None)
            if world > 1:
                allreduce_merge(eng)
            eng.finalize()

        e2e_step()
        barrier()
        sampler2 = ClockSampler(local)
        if rank == 0:
            sampler2.start()
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        if rank == 0:
            line["clocks_e2e"] = sampler2.stop()
        t = torch.tensor([max(e0.elapsed_time(e1) / 1e3, wall)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert mm.overall_count() == n_total
        line["e2e"] = {"value": n_total * e2e_steps / t.item(), "unit": UNIT, "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                       "path": "kta_push_batch_host (pinned host SoA -> chunked cudaMemcpyAsync -> scan) + kta_finalize"}
    eng.close()

    # ---- cpu_baseline: the oracle port on the GPU box's host cores (rank 0, N=1 only) ----
    if rank == 0 and world == 1 and not a.no_cpu:
        alive = a.mode != "counters"
        sample = a.cpu_sample or 20_000_000
        m = min(n, sample)
        host = synth.HostTopic(topic.partition[:m].cpu().numpy(), None, topic.ts_ms[:m].cpu().numpy(),
                               topic.key_len[:m].cpu().numpy(), topic.value_len[:m].cpu().numpy(), None,
                               topic.key_bytes[: topic.key_bytes_len].cpu().numpy(), None)
        rate1, _, _ = cpu_reference_rate(a, m, 1, alive, host)
        cores = os.cpu_count() or 1
        rateN, _, _ = cpu_reference_rate(a, m, cores, False, host)   # counter handler only, equal slices
        line["cpu_baseline"] = {
            "value": rate1, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "first %d records of this workload, C restatement of src/metric.rs:206-305 + fnv32.rs (%s), "
                      "single thread as the reference is by construction (src/kafka.rs:92-135); Rust original not "
                      "buildable here (no toolchain)" % (m, "MessageMetrics + LogCompactionInMemoryMetrics" if alive
                                                         else "MessageMetrics"),
            "counters_only_sharded_value": rateN, "counters_only_sharded_cores": cores}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def extra_modes(a, topic, kta, torch, dev, peak):
    """Kernel-only numbers for the two reference-parity modes on the same topic (informational)."""
    out = {}
    for mode in ("counters", "alive"):
        if mode == a.mode:
            continue
        try:
            eng = kta.KtaEngine(a.partitions, count_alive_keys=(mode == "alive"), device=dev.index)
            eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
            keys = mode != "counters"

            def step():
                eng.reset()
                eng.scan_batch_device(topic.partition, topic.ts_ms, topic.key_len, topic.value_len,
                                      key_bytes=topic.key_bytes if keys else None,
                                      key_bytes_len=topic.key_bytes_len if keys else 0,
                                      key_tile_base=topic.key_tile_base if keys else None)
                eng.finalize()
            step()
            step()
            eng.set_timing(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                step()
            e1.record()
            torch.cuda.synchronize(dev)
            kms, kn = eng.scan_time_ms()
            alg = 20 * topic.n + (topic.key_bytes_len if keys else 0)
            ach = alg / (kms / kn / 1e3) / 1e9
            out[mode] = {"msg_per_s": topic.n * 5 / (e0.elapsed_time(e1) / 1e3), "kernel_ms": kms / kn,
                         "achieved_gb_s": ach, "frac": ach / peak,
                         "alive_keys": eng.alive_keys() if mode == "alive" else None}
            eng.close()
        except Exception as ex:  # informational only
            out[mode] = {"error": repr(ex)}
    return out


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
