"""The synthetic topic (the feeder standing in for consumer.poll, src/kafka.rs:93): determinism,
Kafka-like structure, and shard enumeration used by the multi-GPU path."""
import numpy as np
import pytest

from kafka_topic_analyzer_b200 import synth
import np_oracle


def test_deterministic_and_sliceable():
    s = synth.make_spec(64 * 500, 64, ts_missing_per_10k=50)
    a = synth.fill_host(s)
    b = synth.fill_host(s, start=1000, count=5000)
    assert np.array_equal(a.partition[1000:6000], b.partition)
    assert np.array_equal(a.ts_ms[1000:6000], b.ts_ms)
    assert np.array_equal(a.key_len[1000:6000], b.key_len)
    assert np.array_equal(a.value_len[1000:6000], b.value_len)
    k0 = int(np.maximum(a.key_len[:1000], 0).sum())
    assert np.array_equal(a.key_bytes[k0:k0 + b.key_bytes.size], b.key_bytes)
    assert np.array_equal(a.seq, np.arange(a.n, dtype=np.uint64))


@pytest.mark.parametrize("run_len", [1, 7, 500])
def test_offsets_are_per_partition_running_counts(run_len):
    P = 8
    s = synth.make_spec(P * run_len * 40, P, run_len=run_len)
    t = synth.fill_host(s)
    for p in range(P):
        o = t.offset[t.partition == p]
        assert np.array_equal(o, np.arange(o.size))          # exactly what a Kafka partition log looks like
        assert o.size == t.n // P


@pytest.mark.parametrize("key_mode", [0, 1, 2])
def test_same_key_same_partition_same_bytes(key_mode):
    P = 16
    s = synth.make_spec(P * 2000, P, key_mode=key_mode, distinct_keys=640, null_key_per_10k=0)
    t = synth.fill_host(s)
    h = np_oracle.fnv32_many(t.key_len, t.key_bytes)
    seen = {}
    for hh, p, kl in zip(h.tolist(), t.partition.tolist(), t.key_len.tolist()):
        if kl < 8:
            continue   # key_mode 2: very short random keys (and the EMPTY key) legitimately repeat across ids
        assert seen.setdefault(hh, p) == p
    assert len(seen) <= 640


def test_value_sizes_and_fractions():
    s = synth.make_spec(4 * 50_000, 4, value_mean=1024, tombstone_per_10k=500, null_key_per_10k=100)
    t = synth.fill_host(s)
    v = t.value_len[t.value_len >= 0]
    assert v.min() >= 512 and v.max() <= 1536 and abs(v.mean() - 1024) < 8
    assert abs((t.value_len < 0).mean() - 0.05) < 0.005
    assert abs((t.key_len < 0).mean() - 0.01) < 0.003


@pytest.mark.parametrize("world,run_len", [(2, 1), (4, 3), (8, 1)])
def test_shards_partition_the_topic(world, run_len):
    P = 16
    s = synth.make_spec(P * run_len * 100, P, run_len=run_len, key_mode=2)
    whole = synth.fill_host(s)
    seen = np.zeros(whole.n, dtype=bool)
    for r in range(world):
        sh = synth.fill_host(s, rank=r, world=world)
        assert sh.n == whole.n // world
        assert np.all(sh.partition % world == r)               # gpu = partition mod G (SURVEY.md §8 e)
        assert np.all(np.diff(sh.seq.astype(np.int64)) > 0)    # local order == global seq order
        idx = sh.seq.astype(np.int64)
        assert not seen[idx].any()
        seen[idx] = True
        assert np.array_equal(whole.partition[idx], sh.partition)
        assert np.array_equal(whole.value_len[idx], sh.value_len)
        assert np.array_equal(whole.key_len[idx], sh.key_len)
    assert seen.all()


def test_numpy_oracle_agrees_with_c_oracle():
    from parity import oracle_for
    from oracle_lib import COUNTERS
    s = synth.make_spec(12 * 3000, 12, key_mode=2, tombstone_per_10k=2000, null_key_per_10k=500,
                        ts_missing_per_10k=100, empty_value_per_10k=100, distinct_keys=1200)
    t = synth.fill_host(s)
    o = oracle_for(t, count_alive_keys=True)
    m = np_oracle.message_metrics(12, t.partition, t.ts_ms, t.key_len, t.value_len)
    for p in range(12):
        for name in COUNTERS:
            assert o.counter(name, p) == int(m[name][p]), (name, p)
        assert o.hist(0, p).tolist() == m["khist"][p].tolist()
        assert o.hist(1, p).tolist() == m["vhist"][p].tolist()
    assert o.scalar("largest_message") == m["largest"] and o.scalar("smallest_message") == m["smallest"]
    assert o.scalar("overall_size") == m["overall_size"] and o.scalar("overall_count") == m["overall_count"]
    assert o.earliest()[0] == m["min_ts_s"] and o.latest() == m["max_ts_s"]
    alive = np_oracle.alive_set(t.key_len, t.value_len, np_oracle.fnv32_many(t.key_len, t.key_bytes))
    assert o.scalar("sum_all_alive") == len(alive)


def test_log_uniform_keys_are_skewed_but_keep_the_topic_structure():
    """Stress case of SURVEY.md §8 d (Zipf s = 1 staircase): a few hot keys, a long tail; still one partition and one
    byte string per key, and the uniform default is untouched by the flag's code path."""
    P, K = 8, 4096                     # 4096 keys per partition
    base = synth.make_spec(P * 40_000, P, distinct_keys=P * K, null_key_per_10k=0)
    skew = synth.make_spec(P * 40_000, P, distinct_keys=P * K, null_key_per_10k=0, zipf_keys=True)
    assert skew.key_mode == base.key_mode | synth.KEYS_LOGUNIFORM
    tb, ts = synth.fill_host(base), synth.fill_host(skew)
    assert np.array_equal(tb.partition, ts.partition) and np.array_equal(tb.value_len, ts.value_len)   # only the ids change
    ids_b = tb.key_bytes.reshape(-1, 16)[:, :8].copy().view(np.uint64).ravel()
    ids_s = ts.key_bytes.reshape(-1, 16)[:, :8].copy().view(np.uint64).ravel()
    assert np.all(ids_s % P == ts.partition.astype(np.uint64))            # same key -> same partition
    within = (ids_s // P).astype(np.int64)
    assert within.min() == 0 and within.max() < K
    # the 13 bit lengths 0..12 are equally likely; indices of length b are [2^b - 1, 2^(b+1) - 2], and the top length
    # wraps around K = 2^12 (a uniform floor under the staircase, so every key stays reachable)
    for b in (1, 4, 8, 11):
        frac = float((within < (1 << b) - 1).mean())
        want = b / 13 + ((1 << b) - 1) / K / 13
        assert abs(frac - want) < 0.01, (b, frac, want)
    top = np.sort(np.bincount(within, minlength=K))[::-1]
    assert top[:41].sum() > 0.40 * ts.n                                   # 1 % of the keys carry > 40 % of the records
    uniform_top = np.sort(np.bincount((ids_b // P).astype(np.int64), minlength=K))[::-1]
    assert uniform_top[:41].sum() < 0.02 * tb.n


def test_geometric_value_tail():
    s = synth.make_spec(4 * 100_000, 4, value_mean=1024, tombstone_per_10k=500, geometric_values=True)
    u = synth.make_spec(4 * 100_000, 4, value_mean=1024, tombstone_per_10k=500)
    t, tu = synth.fill_host(s), synth.fill_host(u)
    assert np.array_equal(t.value_len < 0, tu.value_len < 0)              # same tombstones
    v, vu = t.value_len[t.value_len >= 0].astype(np.int64), tu.value_len[tu.value_len >= 0].astype(np.int64)
    g = v // vu                                                           # the power of two applied to each record
    assert np.array_equal(v, vu * g) and set(np.unique(g).tolist()) <= {1, 2, 4, 8, 16, 32, 64}
    for k in range(6):
        assert abs(float((g == 1 << k).mean()) - 2.0 ** -(k + 1)) < 0.004, k
    assert abs(float((g == 64).mean()) - 2.0 ** -6) < 0.002               # the cap collects the rest of the tail
    assert v.max() >= 32 * 1024 and abs(v.mean() / vu.mean() - 4.0) < 0.1   # E[2^g] = 6 * 1/2 + 64/64 = 4


def test_stress_distributions_through_both_oracles():
    from parity import oracle_for
    from oracle_lib import COUNTERS
    s = synth.make_spec(6 * 5000, 6, key_mode=1, distinct_keys=6 * 512, tombstone_per_10k=3000, zipf_keys=True,
                        geometric_values=True)
    t = synth.fill_host(s)
    o = oracle_for(t, count_alive_keys=True)
    m = np_oracle.message_metrics(6, t.partition, t.ts_ms, t.key_len, t.value_len)
    for p in range(6):
        for name in COUNTERS:
            assert o.counter(name, p) == int(m[name][p]), (name, p)
        assert o.hist(1, p).tolist() == m["vhist"][p].tolist()
    alive = np_oracle.alive_set(t.key_len, t.value_len, np_oracle.fnv32_many(t.key_len, t.key_bytes))
    assert o.scalar("sum_all_alive") == len(alive)
    assert int((m["vhist"][:, 12:]).sum()) > 0                            # values >= 2 KiB exist: the tail reaches the high buckets


def test_invalid_key_mode_flags_are_rejected():
    s = synth.make_spec(64, 4)
    s.key_mode = 3
    assert synth.lib().kta_synth_shard_records(s, 0, 1) == -1
    s.key_mode = 0x400
    assert synth.lib().kta_synth_shard_records(s, 0, 1) == -1
    s.key_mode = 2 | synth.KEYS_LOGUNIFORM | synth.VALUES_GEOMETRIC
    assert synth.lib().kta_synth_shard_records(s, 0, 1) == 64 and synth.lib().kta_synth_shard_records(s, 1, 4) == 16
