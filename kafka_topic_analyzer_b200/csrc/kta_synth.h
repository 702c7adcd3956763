// kta_synth.h — the synthetic in-memory Kafka topic (BASELINE.json configs[0..4]; SURVEY.md §8 d).
// Counter-based: record i of the topic is a pure function of (spec, i), evaluated by the SAME code
// on host and device, so multi-billion-record topics never need host storage and a CPU checker can
// regenerate any slice.  Stands in for consumer.poll() (src/kafka.rs:93): `seq` is the order in
// which the reference's single consumer thread would have seen the records (src/kafka.rs:99).
#pragma once
#include <stdint.h>

#include "../../include/kta.h"

#if defined(__CUDACC__)
#define KTA_HD __host__ __device__ __forceinline__
#else
#define KTA_HD static inline
#endif

#define KTA_SYNTH_MAX_KEY 40

KTA_HD uint64_t kta_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// independent stream `s` of the counter-based generator at counter `i`
KTA_HD uint64_t kta_synth_mix(uint64_t seed, uint64_t i, uint32_t s) {
    return kta_splitmix64(kta_splitmix64(seed ^ ((uint64_t)s * 0xD1B54A32D192ED03ull)) + i * 0x9E3779B97F4A7C15ull);
}

struct kta_synth_record {
    uint64_t seq;       // global index i
    int32_t partition;
    int64_t offset;     // per-partition running count (closed form)
    int64_t ts_ms;      // -1 = not available
    uint64_t key_id;    // valid iff key_len >= 0
    int32_t key_len;    // -1 = null
    int32_t value_len;  // -1 = tombstone
};

KTA_HD int32_t kta_synth_key_format(const kta_synth_spec &s) { return s.key_mode & 0xff; }

KTA_HD uint64_t kta_synth_keys_per_partition(const kta_synth_spec &s) {
    uint64_t d = s.distinct_keys / (uint64_t)s.num_partitions;
    return d ? d : 1;
}

// key bytes depend on key_id only (the same key always has the same bytes)
KTA_HD int32_t kta_synth_key_len(const kta_synth_spec &s, uint64_t key_id) {
    const int32_t fmt = kta_synth_key_format(s);
    if (fmt == 0) return 16;
    if (fmt == 1) {
        int32_t n = 5;
        for (uint64_t v = key_id; v >= 10; v /= 10) n++;
        return n;
    }
    return (int32_t)(kta_synth_mix(s.seed, key_id, 7) % (KTA_SYNTH_MAX_KEY + 1));
}

// writes kta_synth_key_len bytes to out (capacity KTA_SYNTH_MAX_KEY)
KTA_HD int32_t kta_synth_key_bytes(const kta_synth_spec &s, uint64_t key_id, uint8_t *out) {
    int32_t len = kta_synth_key_len(s, key_id);
    const int32_t fmt = kta_synth_key_format(s);
    if (fmt == 0) {
        uint64_t a = key_id, b = key_id * 0x9E3779B97F4A7C15ull;
        for (int j = 0; j < 8; j++) out[j] = (uint8_t)(a >> (8 * j));
        for (int j = 0; j < 8; j++) out[8 + j] = (uint8_t)(b >> (8 * j));
    } else if (fmt == 1) {
        out[0] = 'k'; out[1] = 'e'; out[2] = 'y'; out[3] = '-';
        uint64_t v = key_id;
        for (int j = len - 1; j >= 4; j--) { out[j] = (uint8_t)('0' + v % 10); v /= 10; }
    } else {
        for (int j = 0; j < len; j++)
            out[j] = (uint8_t)(kta_synth_mix(s.seed, key_id, 8 + (uint32_t)(j >> 3)) >> (8 * (j & 7)));
    }
    return len;
}

// global index i -> record
KTA_HD void kta_synth_record_at(const kta_synth_spec &s, uint64_t i, kta_synth_record &r) {
    const uint64_t P = (uint64_t)s.num_partitions, R = (uint64_t)s.run_len;
    const uint64_t run = i / R, within = i % R;
    const uint64_t cycle = run / P, slot = run % P;
    const uint64_t shift = kta_synth_mix(s.seed, cycle, 0) % P;
    const uint64_t p = (slot + shift) % P;
    r.seq = i;
    r.partition = (int32_t)p;
    r.offset = (int64_t)(cycle * R + within);
    const bool null_key = (kta_synth_mix(s.seed, i, 1) % 10000u) < (uint64_t)s.null_key_per_10k;
    const uint64_t K = kta_synth_keys_per_partition(s);
    uint64_t kidx = kta_synth_mix(s.seed, i, 2);
    if (s.key_mode & KTA_SYNTH_KEYS_LOGUNIFORM) {
        // pick a bit length b uniformly, then an index uniformly among those of that length: [2^b - 1, 2^(b+1) - 2]
        int nbits = 0;
        while (nbits < 63 && (2ull << nbits) <= K) nbits++;          // floor(log2 K)
        const uint32_t b = (uint32_t)((kidx >> 40) % (uint64_t)(nbits + 1));
        kidx = ((1ull << b) - 1) + (kidx & ((1ull << b) - 1));
    }
    r.key_id = (kidx % K) * P + p;
    r.key_len = null_key ? -1 : kta_synth_key_len(s, r.key_id);
    const uint64_t rv = kta_synth_mix(s.seed, i, 3);
    if ((rv % 10000u) < (uint64_t)s.tombstone_per_10k) {
        r.value_len = -1;
    } else if (((rv >> 20) % 10000u) < (uint64_t)s.empty_value_per_10k) {
        r.value_len = 0;
    } else {
        const uint64_t m = (uint64_t)s.value_mean;
        uint64_t v = m / 2 + kta_synth_mix(s.seed, i, 4) % (m + 1);
        if (s.key_mode & KTA_SYNTH_VALUES_GEOMETRIC) {
            const uint64_t g = kta_synth_mix(s.seed, i, 6);
            int k = 0;
            while (k < 6 && ((g >> k) & 1) == 0) k++;                // P(k) = 2^-(k+1), capped at 6
            v <<= k;
            if (v > 0x7fffffffull) v = 0x7fffffffull;
        }
        r.value_len = (int32_t)v;
    }
    const uint64_t rt = kta_synth_mix(s.seed, i, 5);
    r.ts_ms = ((rt % 10000u) < (uint64_t)s.ts_missing_per_10k)
                  ? -1
                  : (int64_t)(1500000000000ull + i * 7 + (rt >> 32) % 1000u);
}

// local index j of the shard (partitions p % world == rank), in global seq order -> global index.
// Requires num_partitions % world == 0.
KTA_HD uint64_t kta_synth_local_to_global(const kta_synth_spec &s, int32_t rank, int32_t world, uint64_t j) {
    if (world <= 1) return j;
    const uint64_t P = (uint64_t)s.num_partitions, R = (uint64_t)s.run_len, G = (uint64_t)world;
    const uint64_t per_cycle = P / G;  // runs of this shard per cycle
    const uint64_t lrun = j / R, within = j % R;
    const uint64_t cycle = lrun / per_cycle, m = lrun % per_cycle;
    const uint64_t shift = kta_synth_mix(s.seed, cycle, 0) % P;
    // owned slots satisfy (slot + shift) % G == rank  (G divides P)
    const uint64_t s0 = ((uint64_t)rank + G - shift % G) % G;
    const uint64_t slot = s0 + m * G;
    return (cycle * P + slot) * R + within;
}
