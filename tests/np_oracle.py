"""Second, independent restatement of the reference path in numpy (vectorised, order-free where the
reference is order-free) — TEST INFRASTRUCTURE.  Used to cross-check oracle/kta_oracle.c on seeded
inputs; follows /root/reference/src/metric.rs:206-253, :288-305 and src/fnv32.rs:92-101."""
import numpy as np


def fnv32_many(key_len, key_bytes):
    """reference hash of every packed key; null keys (len < 0) → 0"""
    kl = np.maximum(key_len.astype(np.int64), 0)
    off = np.concatenate([[0], np.cumsum(kl)[:-1]]) if kl.size else np.zeros(0, dtype=np.int64)
    h = np.full(kl.shape, 0x811C9DC5, dtype=np.uint64)
    maxlen = int(kl.max()) if kl.size else 0
    kb = np.concatenate([key_bytes.astype(np.uint64), np.zeros(1, dtype=np.uint64)])
    for j in range(maxlen):
        live = kl > j
        idx = np.where(live, off + j, len(kb) - 1)
        nh = ((h ^ kb[idx]) * np.uint64(0x811C9DC5)) & np.uint64(0xFFFFFFFF)  # fnv32.rs:96-97
        h = np.where(live, nh, h)
    h = h.astype(np.uint32)
    h[key_len < 0] = 0
    return h


def bucket(lens):
    lens = lens.astype(np.int64)
    out = np.zeros(lens.shape, dtype=np.int64)
    nz = lens > 0
    out[nz] = np.floor(np.log2(lens[nz])).astype(np.int64) + 1
    # guard against float rounding at exact powers of two
    fix = nz & ((np.int64(1) << np.clip(out - 1, 0, 62)) > lens)
    out[fix] -= 1
    fix = nz & ((np.int64(1) << np.clip(out, 0, 62)) <= lens)
    out[fix] += 1
    return out


def message_metrics(P, partition, ts_ms, key_len, value_len):
    """dict of per-partition u64 arrays + globals (metric.rs:206-253)."""
    p = partition.astype(np.int64)
    kl = key_len.astype(np.int64)
    vl = value_len.astype(np.int64)
    keyed, valued = kl >= 0, vl >= 0
    bc = lambda w=None, m=None: np.bincount(p if m is None else p[m], weights=None if w is None else (w if m is None else w[m]),
                                            minlength=P).astype(np.uint64)
    out = {
        "total": bc(),
        "tombstones": bc(m=~valued),
        "alive": bc(m=valued),
        "key_null": bc(m=~keyed),
        "key_non_null": bc(m=keyed),
    }
    # exact integer sums (bincount weights are float64: do it with add.at on uint64)
    ks = np.zeros(P, dtype=np.uint64)
    np.add.at(ks, p[keyed], kl[keyed].astype(np.uint64))
    vs = np.zeros(P, dtype=np.uint64)
    np.add.at(vs, p[valued], vl[valued].astype(np.uint64))
    out["key_size_sum"], out["value_size_sum"] = ks, vs
    ts = np.where(ts_ms == -1, 0, ts_ms)                      # :209
    ts_s = np.where(ts >= 0, ts // 1000, -((-ts) // 1000))    # :210 truncating division
    out["min_ts_s"] = int(ts_s.min()) if ts_s.size else None
    out["max_ts_s"] = int(ts_s.max()) if ts_s.size else None
    size = np.where(keyed, kl, 0) + vl
    out["largest"] = int(size[valued].max()) if valued.any() else 0       # :249-251
    out["smallest"] = int(size[valued].min()) if valued.any() else 0      # :177-183 (0 when unset)
    out["overall_size"] = int(ks.sum() + vs.sum())
    out["overall_count"] = int(p.shape[0])
    kh = np.zeros((P, 32), dtype=np.uint64)
    np.add.at(kh, (p[keyed], bucket(kl[keyed])), 1)
    vh = np.zeros((P, 32), dtype=np.uint64)
    np.add.at(vh, (p[valued], bucket(vl[valued])), 1)
    out["khist"], out["vhist"] = kh, vh
    return out


def alive_set(key_len, value_len, hashes, seq=None):
    """set of hash values whose LAST record (by seq) had a value (metric.rs:288-305 replayed)."""
    keyed = key_len >= 0
    h = hashes[keyed].astype(np.int64)
    alive = (value_len[keyed] >= 0)
    s = np.arange(key_len.shape[0], dtype=np.int64)[keyed] if seq is None else seq[keyed].astype(np.int64)
    order = np.lexsort((s, h))
    h, alive = h[order], alive[order]
    last = np.ones(h.shape, dtype=bool)
    last[:-1] = h[1:] != h[:-1]
    return set(h[last & alive].tolist())
