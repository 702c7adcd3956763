/* minimal.c — the drop-in boundary from plain C: what a host that already has messages in hand does.
 * Mirrors the reference's use of its two handlers (src/main.rs:77-82 construct, src/kafka.rs:107-109 one
 * handle_message per polled record, src/main.rs:130-170 read the getters).
 *
 *   gcc -std=c11 -I../../include minimal.c -L../../kafka_topic_analyzer_b200 -lkta_gpu \
 *       -Wl,-rpath,'$ORIGIN/../../kafka_topic_analyzer_b200' -o minimal && ./minimal
 *
 * There is no CPU fallback: without a usable CUDA device kta_create fails and this program says why and exits 2. */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "kta.h"

static int die(const char *what) {
    fprintf(stderr, "%s: %s\n", what, kta_last_error());
    return 2;
}

int main(void) {
    kta_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = (int32_t)sizeof cfg;
    cfg.device = -1;               /* current device */
    cfg.num_partitions = 2;        /* from the topic metadata (src/kafka.rs:60-72) */
    cfg.count_alive_keys = 1;      /* -c given once (src/main.rs:77-80) */
    cfg.now_s = INT64_MIN;         /* earliest_message starts at Utc::now() (src/metric.rs:39) */
    kta_handle *h = NULL;
    if (kta_create(&cfg, &h) != KTA_OK) return die("kta_create");

    /* partition, offset, timestamp (ms, -1 = not available), key (NULL = null key), key_len, value_len (-1 = tombstone) */
    static const struct { int32_t p; int64_t off, ts; const char *key; int32_t vlen; } msgs[] = {
        {0, 0, 1500000000000, "user-1", 120}, {1, 0, 1500000001000, "user-2", 80}, {0, 1, 1500000002000, "user-1", -1},
        {1, 1, -1, NULL, 10},                 {0, 2, 1500000003000, "user-3", 0},
    };
    for (size_t i = 0; i < sizeof msgs / sizeof msgs[0]; i++) {
        const int32_t kl = msgs[i].key ? (int32_t)strlen(msgs[i].key) : -1;
        if (kta_push(h, msgs[i].p, msgs[i].off, msgs[i].ts, (const uint8_t *)msgs[i].key, kl, msgs[i].vlen) != KTA_OK)
            return die("kta_push");
    }
    if (kta_finalize(h) != KTA_OK) return die("kta_finalize");

    for (int32_t p = 0; p < cfg.num_partitions; p++) {
        uint64_t total = 0, tomb = 0, kbytes = 0, vbytes = 0;
        float dirty = 0;
        kta_counter(h, KTA_TOTAL, p, &total);
        kta_counter(h, KTA_TOMBSTONES, p, &tomb);
        kta_counter(h, KTA_KEY_SIZE_SUM, p, &kbytes);
        kta_counter(h, KTA_VALUE_SIZE_SUM, p, &vbytes);
        kta_dirty_ratio(h, p, &dirty);
        printf("partition %d: total %llu tombstones %llu key bytes %llu value bytes %llu dirty ratio %.4f\n", p,
               (unsigned long long)total, (unsigned long long)tomb, (unsigned long long)kbytes, (unsigned long long)vbytes, dirty);
    }
    uint64_t alive = 0;
    if (kta_alive_keys(h, &alive) != KTA_OK) return die("kta_alive_keys");
    printf("alive keys: %llu\n", (unsigned long long)alive);   /* user-2 and user-3: user-1's last record is a tombstone */
    kta_destroy(h);
    return 0;
}
