"""Shared comparison helpers for the parity tests (TEST INFRASTRUCTURE)."""
import numpy as np

from kafka_topic_analyzer_b200 import metrics as M
from oracle_lib import COUNTERS, Oracle


def oracle_for(topic, count_alive_keys=False, track_stream=False, now=(4102444800, 123456789), order=None):
    """Runs the CPU oracle over a HostTopic record by record, in seq order (src/kafka.rs:99)."""
    o = Oracle(count_alive_keys=count_alive_keys, track_stream=track_stream, now=now)
    o.handle_batch(topic.partition, topic.ts_ms, topic.key_len, topic.value_len, topic.key_bytes)
    return o


def assert_parity(engine, o, P, check_alive=False, hll_regs=None, extra_partitions=(-1,)):
    """Bit-exact comparison of everything the reference's report reads (src/main.rs:130-170)."""
    mm = engine.message_metrics
    for p in list(range(P)) + [P + 3] + list(extra_partitions):
        for i, name in enumerate(COUNTERS):
            assert engine.counter(i, p) == o.counter(name, p), (name, p)
        for name, which in (("key_size_avg", M.KEY_SIZE_AVG), ("value_size_avg", M.VALUE_SIZE_AVG),
                            ("message_size_avg", M.MESSAGE_SIZE_AVG)):
            try:
                want = o.avg(name, p)
            except ZeroDivisionError:
                want = "panic"
            try:
                got = engine.avg(which, p)
            except ZeroDivisionError:
                got = "panic"
            assert got == want, (name, p, got, want)
        assert mm.dirty_ratio(p) == o.dirty_ratio(p), ("dirty_ratio", p)   # f32, bit-exact
        if 0 <= p < P:
            assert engine.hist(0, p).tolist() == o.hist(0, p).tolist(), ("khist", p)
            assert engine.hist(1, p).tolist() == o.hist(1, p).tolist(), ("vhist", p)
    assert mm.smallest_message() == o.scalar("smallest_message")
    assert mm.largest_message() == o.scalar("largest_message")
    assert mm.overall_size() == o.scalar("overall_size")
    assert mm.overall_count() == o.scalar("overall_count")
    assert mm.earliest_message() == o.earliest()
    assert mm.latest_message() == o.latest()
    if check_alive:
        assert engine.alive_keys() == o.scalar("sum_all_alive")
    if hll_regs is not None:
        assert engine.hll_registers().tolist() == hll_regs.tolist()


def random_topic(rng, n, P, max_key=40, big=False):
    """Adversarial random SoA batch: nulls, empties, ragged key lengths, missing/negative timestamps."""
    from kafka_topic_analyzer_b200.synth import HostTopic, tile_base_from_key_len
    part = rng.integers(0, P, size=n).astype(np.int32)
    kl = rng.integers(-1, max_key + 1, size=n).astype(np.int32)
    vl = rng.choice(np.array([-1, -1, 0, 1, 2, 3, 127, 128, 255, 256, 1000, 65535, 65536, (1 << 31) - 1 if big else 99999],
                             dtype=np.int64), size=n).astype(np.int32)
    ts = (1_500_000_000_000 + rng.integers(-10**9, 10**9, size=n)).astype(np.int64)
    ts[rng.random(n) < 0.02] = -1
    ts[rng.random(n) < 0.01] = rng.integers(-5000, 5000)
    nkeys = max(4, n // 8)
    pool = [bytes(rng.integers(0, 256, size=int(l), dtype=np.uint8)) for l in rng.integers(0, max_key + 1, size=nkeys)]
    keys, blob = [], []
    for i in range(n):
        if kl[i] < 0:
            continue
        k = pool[int(rng.integers(0, nkeys))]
        kl[i] = len(k)
        blob.append(k)
    kb = np.frombuffer(b"".join(blob) or b"", dtype=np.uint8).copy()
    seq = np.arange(n, dtype=np.uint64)
    return HostTopic(part, np.zeros(n, dtype=np.int64), ts, kl, vl, seq, kb, tile_base_from_key_len(kl))
