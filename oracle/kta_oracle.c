/*
 * kta_oracle.c — CPU ORACLE (test infrastructure, see kta_oracle.h).  Plain C11, single thread,
 * one record at a time, in sequence order — exactly the shape of the reference's loop
 * (src/kafka.rs:92-135).  Every function cites the reference lines it restates.
 * PARITY UNPINNED BY REFERENCE TESTS (the reference has none); pinned by tests/golden.
 */
#include "kta_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * std::collections::HashMap<i32, u64> as used by PartitionedCounterBucket (src/metric.rs:8-9).
 * Only entry(p).or_insert(0) += x (:75-99) and get(&p) (:198-203) are used, so iteration order
 * is irrelevant; any partition id (including negative) is a valid key.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t *keys;
    uint64_t *vals;
    uint8_t *used;
    size_t cap, len;
} pmap;

static size_t pmap_slot(const pmap *m, int32_t k) {
    uint32_t h = (uint32_t)k * 2654435761u;
    size_t i = h & (m->cap - 1);
    while (m->used[i] && m->keys[i] != k) i = (i + 1) & (m->cap - 1);
    return i;
}

static void pmap_grow(pmap *m) {
    pmap n;
    n.cap = m->cap ? m->cap * 2 : 16;
    n.len = 0;
    n.keys = (int32_t *)calloc(n.cap, sizeof(int32_t));
    n.vals = (uint64_t *)calloc(n.cap, sizeof(uint64_t));
    n.used = (uint8_t *)calloc(n.cap, 1);
    for (size_t i = 0; i < m->cap; i++)
        if (m->used[i]) {
            size_t s = pmap_slot(&n, m->keys[i]);
            n.used[s] = 1;
            n.keys[s] = m->keys[i];
            n.vals[s] = m->vals[i];
            n.len++;
        }
    free(m->keys);
    free(m->vals);
    free(m->used);
    *m = n;
}

/* *map.entry(p).or_insert(0u64) += amount   (src/metric.rs:75-99) */
static void pmap_add(pmap *m, int32_t p, uint64_t amount) {
    if ((m->len + 1) * 2 > m->cap) pmap_grow(m);
    size_t s = pmap_slot(m, p);
    if (!m->used[s]) {
        m->used[s] = 1;
        m->keys[s] = p;
        m->vals[s] = 0;
        m->len++;
    }
    m->vals[s] += amount;
}

/* fn metric(): Some(v) => *v, None => 0   (src/metric.rs:198-203) */
static uint64_t pmap_get(const pmap *m, int32_t p) {
    if (!m->cap) return 0;
    size_t s = pmap_slot(m, p);
    return m->used[s] ? m->vals[s] : 0;
}

static void pmap_free(pmap *m) {
    free(m->keys);
    free(m->vals);
    free(m->used);
}

/* ------------------------------------------------------------------------------------------
 * bit_set::BitSet<u32> (bit-set 0.5.2 over bit-vec 0.6.3): growable vector of u32 blocks.
 *   insert(v): if v >= len, grow to v+1 bits (zero filled); set bit; (returns whether new)
 *   remove(v): if !contains(v) return false; clear bit
 *   len():     sum of count_ones() over blocks
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    uint32_t *blocks;
    size_t nblocks; /* allocated + zeroed blocks */
    size_t nbits;   /* logical bit length of the BitVec */
} bitset;

static void bitset_grow(bitset *b, size_t nbits) {
    size_t need = (nbits + 31) / 32;
    if (need > b->nblocks) {
        size_t cap = b->nblocks ? b->nblocks : 64;
        while (cap < need) cap *= 2;
        if (cap > ((size_t)1 << 27)) cap = (size_t)1 << 27; /* 2^32 bits */
        b->blocks = (uint32_t *)realloc(b->blocks, cap * sizeof(uint32_t));
        memset(b->blocks + b->nblocks, 0, (cap - b->nblocks) * sizeof(uint32_t));
        b->nblocks = cap;
    }
    if (nbits > b->nbits) b->nbits = nbits;
}

static int bitset_contains(const bitset *b, size_t v) {
    return v < b->nbits && ((b->blocks[v / 32] >> (v % 32)) & 1u);
}

static int bitset_insert(bitset *b, size_t v) {
    if (v >= b->nbits) bitset_grow(b, v + 1);
    if (bitset_contains(b, v)) return 0;
    b->blocks[v / 32] |= 1u << (v % 32);
    return 1;
}

static int bitset_remove(bitset *b, size_t v) {
    if (!bitset_contains(b, v)) return 0;
    b->blocks[v / 32] &= ~(1u << (v % 32));
    return 1;
}

static size_t bitset_len(const bitset *b) {
    size_t n = 0;
    size_t nb = (b->nbits + 31) / 32;
    for (size_t i = 0; i < nb; i++) n += (size_t)__builtin_popcount(b->blocks[i]);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * src/fnv32.rs:74-101.  NOTE the multiplier is 0x811c9dc5 (the offset basis), NOT the FNV prime
 * 0x01000193 (fnv32.rs:97); init is 0x811c9dc5 (:80); xor then multiply (:96-97).
 * ------------------------------------------------------------------------------------------ */
uint32_t kto_fnv32(const uint8_t *bytes, size_t len) {
    uint32_t hash = 0x811c9dc5u; /* FnvHasher::default, fnv32.rs:79-81 */
    for (size_t i = 0; i < len; i++) {
        hash = hash ^ (uint32_t)bytes[i];  /* fnv32.rs:96 */
        hash = hash * 0x811c9dc5u;         /* fnv32.rs:97 wrapping_mul */
    }
    return hash; /* finish, fnv32.rs:87-89 */
}

/* ------------------------------------------------------------------------------------------ */
struct kto {
    /* struct MessageMetrics, src/metric.rs:11-26 */
    pmap total_messages, tombstones, alive, key_null, key_non_null, key_size_sum, value_size_sum;
    int64_t earliest_s;  /* DateTime<Utc>: seconds ... */
    int32_t earliest_ns; /* ... + nanoseconds (only Utc::now() has ns != 0) */
    int64_t latest_s;
    uint64_t smallest_message, largest_message, overall_size, overall_count;
    /* struct LogCompactionInMemoryMetrics, src/metric.rs:262-264 */
    int has_lc;
    bitset store;
    /* EXTENSIONS (not in the reference) */
    pmap khist[KTO_HIST_BUCKETS], vhist[KTO_HIST_BUCKETS];
    int track_stream;
    int no_hist; /* flags bit2: pure reference path, no extension work (used when timing) */
    bitset ever; /* every hash inserted with key && value (for the in-stream HLL) */
};

/* flags: bit0 = -c (count alive keys); bit1 = also track the in-stream insert set (HLL extension);
 * bit2 = skip the histogram extension (pure reference work, for the cpu_baseline timing) */
kto *kto_new(int flags, int64_t now_s, int32_t now_ns) {
    kto *o = (kto *)calloc(1, sizeof(kto));
    o->earliest_s = now_s;   /* earliest_message: Utc::now(), metric.rs:39 */
    o->earliest_ns = now_ns;
    o->latest_s = 0;         /* from_timestamp(0,0), metric.rs:40 */
    o->largest_message = 0;                 /* :41 */
    o->smallest_message = UINT64_MAX;       /* :42 */
    o->overall_size = 0;                    /* :43 */
    o->overall_count = 0;                   /* :44 */
    o->has_lc = flags & 1;                  /* main.rs:77-80 */
    o->track_stream = (flags >> 1) & 1;
    o->no_hist = (flags >> 2) & 1;
    return o;
}

void kto_free(kto *o) {
    if (!o) return;
    pmap_free(&o->total_messages);
    pmap_free(&o->tombstones);
    pmap_free(&o->alive);
    pmap_free(&o->key_null);
    pmap_free(&o->key_non_null);
    pmap_free(&o->key_size_sum);
    pmap_free(&o->value_size_sum);
    for (int b = 0; b < KTO_HIST_BUCKETS; b++) {
        pmap_free(&o->khist[b]);
        pmap_free(&o->vhist[b]);
    }
    free(o->store.blocks);
    free(o->ever.blocks);
    free(o);
}

static int hist_bucket(uint64_t len) { return len == 0 ? 0 : 64 - __builtin_clzll(len); }

/* cmp_and_set_message_size, metric.rs:56-63 */
static void cmp_and_set_message_size(kto *o, uint64_t size) {
    if (o->largest_message < size) o->largest_message = size;
    if (o->smallest_message > size) o->smallest_message = size;
}

/* cmp_and_set_message_timestamp, metric.rs:65-72; cmp always has 0 ns (from_timestamp(s, 0)) */
static void cmp_and_set_message_timestamp(kto *o, int64_t cmp_s) {
    /* earliest.gt(&cmp): (s, ns) > (cmp_s, 0) */
    if (o->earliest_s > cmp_s || (o->earliest_s == cmp_s && o->earliest_ns > 0)) {
        o->earliest_s = cmp_s;
        o->earliest_ns = 0;
    }
    if (o->latest_s < cmp_s) o->latest_s = cmp_s; /* latest.lt(&cmp) */
}

/* impl MetricHandler for MessageMetrics, metric.rs:206-253 */
static void message_metrics_handle(kto *o, int32_t partition, int64_t ts_ms, int32_t key_len,
                                   int32_t value_len) {
    /* :209 m.timestamp().to_millis().unwrap_or(0) */
    int64_t timestamp = (ts_ms == -1) ? 0 : ts_ms;
    /* :210 from_timestamp(timestamp / 1000, 0) — Rust i64 `/` truncates toward zero, as C does */
    int64_t ts_s = timestamp / 1000;
    uint64_t message_size = 0; /* :212 */
    int empty_value = 0;       /* :213 */

    o->overall_count += 1;                      /* :215 inc_overall_count */
    pmap_add(&o->total_messages, partition, 1); /* :216 inc_total */

    if (key_len >= 0) { /* :218 Some(k) */
        pmap_add(&o->key_non_null, partition, 1); /* :220 */
        uint64_t k_len = (uint64_t)key_len;       /* :221 */
        message_size += k_len;                    /* :222 */
        pmap_add(&o->key_size_sum, partition, k_len); /* :223 */
        o->overall_size += k_len;                 /* :224 */
        if (!o->no_hist) pmap_add(&o->khist[hist_bucket(k_len)], partition, 1); /* EXTENSION */
    } else {
        pmap_add(&o->key_null, partition, 1); /* :228 */
    }

    if (value_len >= 0) { /* :233 Some(v) */
        uint64_t v_len = (uint64_t)value_len;           /* :235 */
        message_size += v_len;                          /* :236 */
        pmap_add(&o->value_size_sum, partition, v_len); /* :237 */
        o->overall_size += v_len;                       /* :238 */
        pmap_add(&o->alive, partition, 1);              /* :239 */
        if (!o->no_hist) pmap_add(&o->vhist[hist_bucket(v_len)], partition, 1); /* EXTENSION */
    } else {
        empty_value = 1;                           /* :242 */
        pmap_add(&o->tombstones, partition, 1);    /* :243 */
    }

    cmp_and_set_message_timestamp(o, ts_s); /* :247 */

    if (!empty_value) cmp_and_set_message_size(o, message_size); /* :249-251 */
}

/* impl MetricHandler for LogCompactionInMemoryMetrics, metric.rs:288-305 */
static void log_compaction_handle(kto *o, const uint8_t *key, int32_t key_len, int32_t value_len) {
    if (key_len >= 0) { /* :291 Some(k) */
        size_t k = (size_t)kto_fnv32(key, (size_t)key_len); /* fnv1a, :256-260 */
        if (value_len >= 0) {
            if (o->has_lc) bitset_insert(&o->store, k); /* mark_key_alive :273-276 */
            if (o->track_stream) bitset_insert(&o->ever, k); /* EXTENSION */
        } else {
            if (o->has_lc) bitset_remove(&o->store, k); /* mark_key_dead :278-280 */
        }
    } /* :302 None => {} */
}

void kto_handle_message(kto *o, int32_t partition, int64_t ts_ms, const uint8_t *key,
                        int32_t key_len, int32_t value_len) {
    /* handlers run in registration order, src/kafka.rs:107-109; main.rs:108,112 */
    message_metrics_handle(o, partition, ts_ms, key_len, value_len);
    if (o->has_lc || o->track_stream) log_compaction_handle(o, key, key_len, value_len);
}

void kto_handle_batch(kto *o, int64_t n, const int32_t *partition, const int64_t *ts_ms,
                      const int32_t *key_len, const int32_t *value_len, const uint8_t *key_bytes) {
    size_t off = 0;
    for (int64_t i = 0; i < n; i++) {
        int32_t kl = key_len[i];
        kto_handle_message(o, partition[i], ts_ms[i], key_bytes ? key_bytes + off : NULL, kl,
                           value_len[i]);
        if (kl > 0) off += (size_t)kl;
    }
}

uint64_t kto_total(const kto *o, int32_t p) { return pmap_get(&o->total_messages, p); }
uint64_t kto_tombstones(const kto *o, int32_t p) { return pmap_get(&o->tombstones, p); }
uint64_t kto_alive(const kto *o, int32_t p) { return pmap_get(&o->alive, p); }
uint64_t kto_key_null(const kto *o, int32_t p) { return pmap_get(&o->key_null, p); }
uint64_t kto_key_non_null(const kto *o, int32_t p) { return pmap_get(&o->key_non_null, p); }
uint64_t kto_key_size_sum(const kto *o, int32_t p) { return pmap_get(&o->key_size_sum, p); }
uint64_t kto_value_size_sum(const kto *o, int32_t p) { return pmap_get(&o->value_size_sum, p); }

/* metric.rs:132-139 */
int kto_key_size_avg(const kto *o, int32_t p, uint64_t *out) {
    uint64_t s = kto_key_size_sum(o, p);
    if (s > 0) {
        uint64_t a = kto_alive(o, p);
        if (a == 0) return 1; /* Rust: attempt to divide by zero → panic */
        *out = s / a;
    } else {
        *out = 0;
    }
    return 0;
}

/* metric.rs:141-148 */
int kto_value_size_avg(const kto *o, int32_t p, uint64_t *out) {
    uint64_t s = kto_value_size_sum(o, p);
    if (s > 0) {
        uint64_t a = kto_alive(o, p);
        if (a == 0) return 1;
        *out = s / a;
    } else {
        *out = 0;
    }
    return 0;
}

/* metric.rs:150-157 */
int kto_message_size_avg(const kto *o, int32_t p, uint64_t *out) {
    uint64_t s = kto_key_size_sum(o, p) + kto_value_size_sum(o, p);
    if (s > 0) {
        uint64_t a = kto_alive(o, p);
        if (a == 0) return 1;
        *out = s / a;
    } else {
        *out = 0;
    }
    return 0;
}

/* metric.rs:159-167: tombstones as f32 / (total_messages as f32 / 100.0f32) */
float kto_dirty_ratio(const kto *o, int32_t p) {
    uint64_t total_messages = kto_total(o, p);
    uint64_t tombstones = kto_tombstones(o, p);
    if (total_messages > 0 && tombstones > 0) {
        volatile float t = (float)tombstones;
        volatile float d = (float)total_messages / 100.0f;
        return t / d;
    }
    return 0.0f;
}

void kto_earliest_message(const kto *o, int64_t *s, int32_t *ns) {
    *s = o->earliest_s;
    *ns = o->earliest_ns;
}
int64_t kto_latest_message_s(const kto *o) { return o->latest_s; }
/* metric.rs:177-183 */
uint64_t kto_smallest_message(const kto *o) {
    return o->smallest_message == UINT64_MAX ? 0 : o->smallest_message;
}
uint64_t kto_largest_message(const kto *o) { return o->largest_message; }
uint64_t kto_overall_count(const kto *o) { return o->overall_count; }
uint64_t kto_overall_size(const kto *o) { return o->overall_size; }
uint64_t kto_sum_all_alive(const kto *o) { return (uint64_t)bitset_len(&o->store); }
int kto_alive_contains(const kto *o, uint32_t hash) { return bitset_contains(&o->store, hash); }

/* test hook: overwrite one per-partition counter (index = order of the struct fields, metric.rs:13-19)
 * so the getter / derived arithmetic can be checked against demo_output.png without replaying 245 M
 * messages. */
void kto_test_set_counter(kto *o, int which, int32_t p, uint64_t v) {
    pmap *m[7] = {&o->total_messages, &o->tombstones,   &o->alive,         &o->key_null,
                  &o->key_non_null,   &o->key_size_sum, &o->value_size_sum};
    pmap_add(m[which], p, 0);
    m[which]->vals[pmap_slot(m[which], p)] = v;
}

/* ============================ EXTENSIONS — NOT IN THE REFERENCE ============================ */

void kto_hist(const kto *o, int which, int32_t p, uint64_t out[KTO_HIST_BUCKETS]) {
    for (int b = 0; b < KTO_HIST_BUCKETS; b++)
        out[b] = pmap_get(which ? &o->vhist[b] : &o->khist[b], p);
}

/* murmur3 fmix32 (Appleby): a bijection on 32 bits, so distinct reference hashes stay distinct */
uint32_t kto_hll_mix(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

/* Flajolet et al. 2007 over the 32-bit mixed hash: register index = top `precision` bits,
 * rho = 1 + leading zeros of the remaining 32 - precision bits (max 32 - precision + 1) */
void kto_hll_insert(uint8_t *regs, int precision, uint32_t hash) {
    uint32_t x = kto_hll_mix(hash);
    uint32_t idx = x >> (32 - precision);
    uint32_t rest = x << precision;
    int rho = rest ? __builtin_clz(rest) + 1 : (32 - precision + 1);
    if (regs[idx] < rho) regs[idx] = (uint8_t)rho;
}

static double hll_sigma(double x) {
    if (x == 1.0) return INFINITY;
    double y = 1.0, z = x, zo;
    do {
        x *= x;
        zo = z;
        z += x * y;
        y += y;
    } while (zo != z);
    return z;
}

static double hll_tau(double x) {
    if (x == 0.0 || x == 1.0) return 0.0;
    double y = 1.0, z = 1.0 - x, zo;
    do {
        x = sqrt(x);
        zo = z;
        y *= 0.5;
        z -= (1.0 - x) * (1.0 - x) * y;
    } while (zo != z);
    return z / 3.0;
}

/* Ertl 2017, "New cardinality estimation algorithms for HyperLogLog sketches", improved raw
 * estimator (no empirical bias tables, no small/large-range switch). */
double kto_hll_estimate(const uint8_t *regs, int precision) {
    int q = 32 - precision;
    size_t m = (size_t)1 << precision;
    double C[34];
    for (int k = 0; k <= q + 1; k++) C[k] = 0;
    for (size_t i = 0; i < m; i++) C[regs[i]] += 1.0;
    double z = (double)m * hll_tau(1.0 - C[q + 1] / (double)m);
    for (int k = q; k >= 1; k--) z = 0.5 * (z + C[k]);
    z += (double)m * hll_sigma(C[0] / (double)m);
    const double alpha_inf = 0.72134752044448170368; /* 1 / (2 ln 2) */
    return alpha_inf * (double)m * (double)m / z;
}

static void regs_from_bitset(const bitset *b, int precision, uint8_t *regs) {
    memset(regs, 0, (size_t)1 << precision);
    size_t nb = (b->nbits + 31) / 32;
    for (size_t i = 0; i < nb; i++) {
        uint32_t w = b->blocks[i];
        while (w) {
            int bit = __builtin_ctz(w);
            w &= w - 1;
            kto_hll_insert(regs, precision, (uint32_t)(i * 32 + (size_t)bit));
        }
    }
}

void kto_hll_stream_regs(const kto *o, int precision, uint8_t *regs_out) {
    regs_from_bitset(&o->ever, precision, regs_out);
}

void kto_hll_alive_regs(const kto *o, int precision, uint8_t *regs_out) {
    regs_from_bitset(&o->store, precision, regs_out);
}
