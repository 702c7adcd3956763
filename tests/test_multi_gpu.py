"""N>1 path.  CPU (gloo, world_size 2 and 4): shard enumeration + the single-SUM-all-reduce merge layout, with
the per-shard states produced by the CPU oracle.  GPU (needs >= 2 devices, otherwise skipped): the real thing
under torchrun with NCCL, against the whole-topic oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
from kafka_topic_analyzer_b200 import synth
from kafka_topic_analyzer_b200.distributed import pack_merge_buffer, fold_merge_buffer, merge_words
from oracle_lib import Oracle, COUNTERS
import np_oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
P, HP = 16, 10
spec = synth.make_spec(P * 3 * 400, P, run_len=3, key_mode=2, distinct_keys=1600, tombstone_per_10k=2500,
                       null_key_per_10k=300, ts_missing_per_10k=200)
def state_of(t):
    """the device state layout (sums | minmax | hll), built from the oracle's view of a shard"""
    o = Oracle(track_stream=True)
    o.handle_batch(t.partition, t.ts_ms, t.key_len, t.value_len, t.key_bytes)
    sums = np.zeros(P * 67 + 1, dtype=np.uint64)
    for p in range(P):
        sums[p * 32:(p + 1) * 32] = o.hist(0, p)
        sums[(P + p) * 32:(P + p + 1) * 32] = o.hist(1, p)
        sums[P * 64 + p] = o.counter("key_size_sum", p)
        sums[P * 65 + p] = o.counter("value_size_sum", p)
        sums[P * 66 + p] = o.counter("key_null", p)
    raw = t.ts_ms
    valued = t.value_len >= 0
    size = np.where(t.key_len >= 0, t.key_len, 0).astype(np.int64) + t.value_len
    mm = (int(raw.min()), int(raw.max()), int(size[valued].min()) if valued.any() else 2**64 - 1,
          int(size[valued].max()) if valued.any() else 0)
    return sums, mm, o.hll_stream_regs(HP).astype(np.uint32)
shard = synth.fill_host(spec, rank=rank, world=world)
sums, mm, hll = state_of(shard)
buf = torch.from_numpy(pack_merge_buffer(sums, mm, hll, rank, world).view(np.int64))
assert buf.numel() == merge_words(len(sums), len(hll), world)
dist.all_reduce(buf, op=dist.ReduceOp.SUM)          # THE one collective of the merge
msums, mmm, mhll = fold_merge_buffer(buf.numpy().view(np.uint64), len(sums), len(hll), world)
wsums, wmm, whll = state_of(synth.fill_host(spec))  # whole topic, one process
assert np.array_equal(msums, wsums), "sums"
assert mmm == wmm, (mmm, wmm)
assert np.array_equal(mhll, whll), "hll"
# exact alive keys: all-gather (hash, stamp) of every shard, last writer wins by GLOBAL seq
h = np_oracle.fnv32_many(shard.key_len, shard.key_bytes)
keyed = shard.key_len >= 0
stamp = ((shard.seq.astype(np.int64) + 1) << 1) | (shard.value_len >= 0)
mine = torch.from_numpy(np.stack([h[keyed].astype(np.int64), stamp[keyed]], axis=1).copy())
sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
dist.all_gather(sizes, torch.tensor([mine.shape[0]]))
cap = max(int(s) for s in sizes)
pad = torch.zeros((cap, 2), dtype=torch.int64); pad[:mine.shape[0]] = mine
allv = [torch.zeros((cap, 2), dtype=torch.int64) for _ in range(world)]
dist.all_gather(allv, pad)
best = {}
for r in range(world):
    for hh, st in allv[r][: int(sizes[r])].tolist():
        if st > best.get(hh, 0): best[hh] = st
alive = sum(1 for st in best.values() if st & 1)
whole = synth.fill_host(spec)
ow = Oracle(count_alive_keys=True); ow.handle_batch(whole.partition, whole.ts_ms, whole.key_len, whole.value_len, whole.key_bytes)
assert alive == ow.scalar("sum_all_alive"), (alive, ow.scalar("sum_all_alive"))
dist.destroy_process_group()
sys.stdout.write(f"rank {rank} ok\n"); sys.stdout.flush()
'''

GPU_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
import kafka_topic_analyzer_b200 as kta
from kafka_topic_analyzer_b200 import synth
from kafka_topic_analyzer_b200.distributed import allreduce_merge
from parity import assert_parity, oracle_for
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
NOW = (4102444800, 5)
P = 16
spec = synth.make_spec(P * 5 * 2000, P, run_len=5, key_mode=2, distinct_keys=4000, tombstone_per_10k=2500, ts_missing_per_10k=20)
for exact in (False, True):
    topic = synth.DeviceTopic(spec, rank=rank, world=world, device=local, with_seq=True)
    e = kta.KtaEngine(P, count_alive_keys=exact, hll_precision=12, device=local, now=NOW)
    e.scan_batch_device(topic.partition, topic.ts_ms, topic.key_len, topic.value_len, key_bytes=topic.key_bytes,
                        key_bytes_len=topic.key_bytes_len, key_tile_base=topic.key_tile_base, seq=topic.seq)
    allreduce_merge(e)
    e.finalize()
    whole = synth.fill_host(spec)
    o = oracle_for(whole, count_alive_keys=exact, track_stream=not exact, now=NOW)
    assert_parity(e, o, P, check_alive=exact, hll_regs=o.hll_alive_regs(12) if exact else o.hll_stream_regs(12))
    e.close()
dist.destroy_process_group()
sys.stdout.write(f"rank {rank} ok\n"); sys.stdout.flush()
'''


def _run(world, script, extra_env=None, timeout=300):
    path = os.path.join("/tmp", "kta_worker_%d_%d.py" % (os.getpid(), world))
    open(path, "w").write(script)
    env = dict(os.environ, **(extra_env or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(29500 + (os.getpid() + world) % 400), path, ROOT]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)


@pytest.mark.parametrize("world", [2, 4])
def test_shard_and_merge_logic_gloo(world):
    r = _run(world, WORKER)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == world


@pytest.mark.gpu
def test_two_gpu_scan_and_nccl_merge():
    from kafka_topic_analyzer_b200 import lib
    if lib().kta_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = _run(2, GPU_WORKER)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2


def test_reference_arm_under_torchrun():
    """bench.py --impl reference launched like the driver does for N>1: rank 0 prints exactly one JSON line,
    the other ranks exit 0 without work."""
    import json
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29900 + os.getpid() % 90), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "1", "--cpu-sample", "200000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["unit"] == "msg/s"
