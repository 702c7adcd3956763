// kta_logdecode.cuh — Kafka RecordBatch v2 (magic 2) → SoA columns, on the GPU.
//
// This is the step BEFORE the metric path: in the reference it happens inside librdkafka's fetch parser, which
// hands one BorrowedMessage per record to the handlers (src/kafka.rs:93,107-109).  Here a whole log segment
// (the concatenated record batches of one partition, exactly what a broker stores in <topic>-<p>/*.log and
// sends in a fetch response) is decoded where it lies in HBM:
//   batch header (61 bytes, big-endian): baseOffset i64 | batchLength i32 | partitionLeaderEpoch i32 | magic i8 |
//     crc u32 | attributes i16 | lastOffsetDelta i32 | baseTimestamp i64 | maxTimestamp i64 | producerId i64 |
//     producerEpoch i16 | baseSequence i32 | recordsCount i32
//   record: length varint | attributes i8 | timestampDelta varlong | offsetDelta varint | keyLength varint | key |
//     valueLength varint | value | headersCount varint | headers…          (varints are zig-zag, LSB group first)
// Semantics kept from the consumer: control batches (attributes bit 5) are not delivered to the application;
// LogAppendTime batches (attributes bit 3) stamp every record with maxTimestamp; a record's timestamp is baseTimestamp +
// timestampDelta as the consumer computes it, and only a RESULT of -1 means "not available"; key/value length -1 means
// null.  CRCs are not verified (librdkafka's default check.crcs=false).
// Not handled: records of aborted transactions are delivered (a read_committed consumer would filter them through the
// .txnindex / abort markers), legacy magic 0/1 message sets are flagged as malformed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kta {

constexpr int LOG_HEADER_BYTES = 61;
enum LogBatchFlags { LOGB_OK = 0, LOGB_SKIP_CONTROL = 1, LOGB_BAD = 2, LOGB_COMPRESSED = 4 };

__device__ __forceinline__ uint64_t be_u64(const uint8_t *p) {
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) v = (v << 8) | __ldg(p + i);
    return v;
}
__device__ __forceinline__ uint32_t be_u32(const uint8_t *p) {
    return ((uint32_t)__ldg(p) << 24) | ((uint32_t)__ldg(p + 1) << 16) | ((uint32_t)__ldg(p + 2) << 8) | __ldg(p + 3);
}
__device__ __forceinline__ uint32_t be_u16(const uint8_t *p) { return ((uint32_t)__ldg(p) << 8) | __ldg(p + 1); }

__device__ __forceinline__ int64_t unzigzag(uint64_t u) { return (int64_t)(u >> 1) ^ -(int64_t)(u & 1); }

struct LogBatchInfo {      // one per record batch, filled by log_header_kernel
    uint64_t off;          // byte offset of the batch in the segment buffer
    uint32_t len;          // 12 + batchLength
    uint32_t flags;        // LogBatchFlags
    int32_t partition;
    int32_t records;       // records delivered to the handlers (0 for skipped batches)
    int64_t base_offset, base_ts, max_ts;
    uint32_t log_append_time;
    uint32_t pad;
};

// thread per batch: validate + read the header
__global__ void log_header_kernel(const uint8_t *bytes, int64_t nbytes, const uint64_t *batch_off, int64_t nbatches,
                                  int32_t partition, const int32_t *batch_partition /* per batch, or NULL = `partition` */,
                                  LogBatchInfo *info, uint64_t *rec_count /*[nbatches+1], [b+1]*/, uint32_t *error_flags) {
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nbatches; b += (int64_t)gridDim.x * blockDim.x) {
        LogBatchInfo bi{};
        bi.off = batch_off[b];
        bi.partition = batch_partition ? batch_partition[b] : partition;
        bi.flags = LOGB_BAD;
        if (bi.off + LOG_HEADER_BYTES <= (uint64_t)nbytes) {
            const uint8_t *p = bytes + bi.off;
            const int32_t batch_len = (int32_t)be_u32(p + 8);
            const int magic = (int8_t)__ldg(p + 16);
            const uint32_t attrs = be_u16(p + 21);
            const int32_t count = (int32_t)be_u32(p + 57);
            // recordsCount sizes the output columns, so it must be plausible before anything is allocated for it: the
            // smallest record is 7 bytes (length, attributes, two deltas, key length, value length, header count)
            if (magic == 2 && batch_len >= LOG_HEADER_BYTES - 12 && bi.off + 12 + (uint64_t)batch_len <= (uint64_t)nbytes && count >= 0 &&
                (uint64_t)count * 7u + (uint64_t)(LOG_HEADER_BYTES - 12) <= (uint64_t)batch_len) {
                bi.len = 12u + (uint32_t)batch_len;
                bi.base_offset = (int64_t)be_u64(p);
                bi.base_ts = (int64_t)be_u64(p + 27);
                bi.max_ts = (int64_t)be_u64(p + 35);
                bi.log_append_time = (attrs >> 3) & 1u;
                if (attrs & 0x7u) bi.flags = LOGB_COMPRESSED;
                else if (attrs & 0x20u) bi.flags = LOGB_SKIP_CONTROL;
                else {
                    bi.flags = LOGB_OK;
                    bi.records = count;
                }
            }
        }
        if (bi.flags & (LOGB_BAD | LOGB_COMPRESSED)) atomicOr(error_flags, bi.flags);
        else if (bi.flags == LOGB_OK) atomicMax(error_flags + 1, bi.len);   // [1]: the longest batch (sizes the decode stage)
        info[b] = bi;
        rec_count[b + 1] = (uint64_t)bi.records;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) rec_count[0] = 0;
}

// unsigned LEB128 at p (bounded by end); returns bytes consumed, 0 on malformed input.  Generic byte loads: the batch may
// sit in shared memory or in global memory.
__device__ __forceinline__ int uvarint_g(const uint8_t *p, const uint8_t *end, uint64_t &out) {
    uint64_t v = 0;
    int shift = 0, n = 0;
    while (p + n < end && n < 10) {
        const uint8_t b = p[n];
        n++;
        v |= (uint64_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) {
            out = v;
            return n;
        }
        shift += 7;
    }
    return 0;
}

// One WARP per batch.  Records are length-prefixed, so finding where record i starts is a serial chain: lane 0 hops through
// 32 record-length varints at a time and publishes the 32 start positions; then the 32 lanes parse their records in
// parallel and write the columns coalesced.
// STAGED: the whole batch (what a producer's batch.size bounds: 16 KiB by default) is first brought global→shared by ONE
// bulk async copy per batch per warp (cp.async.bulk → UBLKCP, mbarrier completion), so every hop of the chain is a
// 29-cycle shared-memory read instead of a dependent global access to a new line (measured: the chain of global hops was
// the decoder's bottleneck).  Batches that do not fit the stage, or whose 16-byte-aligned copy would run past the readable
// bytes, are read in place.
// Output: the header columns and, per record, the position of its key bytes in the segment buffer (key_src, only when the
// keys will be hashed) — the keys themselves are packed afterwards by log_gather_keys_kernel, without a second walk.
constexpr int LOG_DECODE_THREADS = 128;
constexpr int LOG_WARP_HEADER = 192;   // per warp: mbarrier (8 B) + 33 record starts (132 B), padded

template <bool STAGED>
__global__ void __launch_bounds__(LOG_DECODE_THREADS) log_decode_kernel(
    const uint8_t *bytes, uint64_t readable /* bytes that may be read from `bytes` */, const LogBatchInfo *info, int64_t nbatches,
    const uint64_t *rec_base, int32_t *partition, int64_t *offset, int64_t *ts_ms, int32_t *key_len, int32_t *value_len,
    uint64_t *key_src, uint32_t stage_bytes /* per warp, multiple of 16 */, uint32_t *error_flags) {
    extern __shared__ __align__(128) unsigned char log_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu;
    unsigned char *wsm = log_smem + (size_t)wib * (LOG_WARP_HEADER + (STAGED ? stage_bytes : 0u));
    uint32_t *s_start = reinterpret_cast<uint32_t *>(wsm + 16);
    unsigned char *stage = wsm + LOG_WARP_HEADER;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(wsm);
    uint32_t phase = 0;
    if (STAGED && lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, gs = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = gw; b < nbatches; b += gs) {
        const LogBatchInfo bi = info[b];
        if (bi.flags == LOGB_OK && bi.records > 0) {
            const uint8_t *base = bytes + bi.off;
            if (STAGED) {
                const uint32_t lead = (uint32_t)(bi.off & 15u), span = (lead + bi.len + 15u) & ~15u;
                if (span <= stage_bytes && (bi.off - lead) + span <= readable) {
                    if (lane == 0) {
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(span) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"((uint32_t)__cvta_generic_to_shared(stage)), "l"(bytes + (bi.off - lead)), "r"(span), "r"(bar)
                                     : "memory");
                    }
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "LOGW_%=:\n\t"
                        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                        "@p bra LOGD_%=;\n\t"
                        "bra LOGW_%=;\n\t"
                        "LOGD_%=:\n\t}" ::"r"(bar), "r"(phase) : "memory");
                    phase ^= 1u;
                    base = stage + lead;
                }
            }
            const uint8_t *end = base + bi.len;
            uint32_t pos = LOG_HEADER_BYTES;          // offset of the next record inside the batch
            const uint64_t r0 = rec_base[b];
            bool ok = true;
            for (int32_t i0 = 0; i0 < bi.records && ok; i0 += 32) {
                const int cnt = min(32, bi.records - i0);
                if (lane == 0) {
                    for (int j = 0; j < cnt; j++) {
                        s_start[j] = pos;
                        uint64_t u;
                        const int n = uvarint_g(base + pos, end, u);
                        const int64_t rec_len = unzigzag(u);
                        if (n <= 0 || rec_len < 0 || (uint64_t)pos + n + rec_len > bi.len) { ok = false; break; }
                        pos += (uint32_t)n + (uint32_t)rec_len;
                    }
                    s_start[32] = ok ? pos : 0xffffffffu;
                }
                __syncwarp();
                pos = s_start[32];
                ok = pos != 0xffffffffu;
                if (!ok) break;
                int64_t klen = -1, vlen = -1, ts_delta = 0, off_delta = 0;
                uint32_t key_at = 0;
                bool lane_ok = true;
                if (lane < cnt) {
                    const uint8_t *q = base + s_start[lane];
                    const uint8_t *rec_end = lane + 1 < cnt ? base + s_start[lane + 1] : base + pos;
                    uint64_t u;
                    int n = uvarint_g(q, rec_end, u); q += n;            // record length (validated by lane 0)
                    q += 1;                                               // record attributes (unused)
                    n = uvarint_g(q, rec_end, u); lane_ok = lane_ok && n > 0; q += n;
                    ts_delta = unzigzag(u);
                    n = uvarint_g(q, rec_end, u); lane_ok = lane_ok && n > 0; q += n;
                    off_delta = unzigzag(u);
                    n = uvarint_g(q, rec_end, u); lane_ok = lane_ok && n > 0; q += n;
                    klen = unzigzag(u);
                    lane_ok = lane_ok && klen >= -1 && klen <= 0x7fffffff && (klen <= 0 || q + klen <= rec_end);
                    key_at = (uint32_t)(q - base);
                    if (lane_ok && klen > 0) q += klen;
                    n = lane_ok ? uvarint_g(q, rec_end, u) : 0; lane_ok = lane_ok && n > 0; q += n;
                    vlen = unzigzag(u);
                    lane_ok = lane_ok && vlen >= -1 && vlen <= 0x7fffffff && (vlen <= 0 || q + vlen <= rec_end);
                }
                __syncwarp();   // every lane has read its start before lane 0 overwrites them
                ok = __all_sync(full, lane_ok);
                if (!ok) break;
                if (lane < cnt) {
                    const uint64_t r = r0 + (uint64_t)i0 + lane;
                    partition[r] = bi.partition;
                    if (offset) offset[r] = bi.base_offset + off_delta;
                    ts_ms[r] = bi.log_append_time ? bi.max_ts : bi.base_ts + ts_delta;
                    key_len[r] = (int32_t)klen;
                    value_len[r] = (int32_t)vlen;
                    if (key_src) key_src[r] = bi.off + key_at;
                }
            }
            if (!ok && lane == 0) atomicOr(error_flags, (uint32_t)LOGB_BAD);
            __syncwarp();   // the stage is free for the next batch's copy
        }
    }
}

// Packs the key bytes in record order (what the scan kernel hashes): one warp per 128-record tile, a lane owns four
// consecutive records; byte offsets from the tile base (key_tile_base, derived from key_len beforehand) plus an in-tile scan.
__global__ void __launch_bounds__(256) log_gather_keys_kernel(const uint8_t *bytes, const uint64_t *key_src, const int32_t *key_len,
                                                              int64_t n, const uint64_t *tile_base, uint8_t *key_out) {
    const int lane = threadIdx.x & 31;
    const int64_t ntiles = (n + 127) / 128;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, gs = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t tile = gw; tile < ntiles; tile += gs) {
        int32_t len[4];
        uint64_t src[4];
        uint32_t mine = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int64_t r = tile * 128 + (int64_t)lane * 4 + k;
            len[k] = r < n ? key_len[r] : -1;
            src[k] = len[k] > 0 ? key_src[r] : 0;
            mine += len[k] > 0 ? (uint32_t)len[k] : 0u;
        }
        uint32_t inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        uint8_t *o = key_out + tile_base[tile] + (inc - mine);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (len[k] > 0) {
                const uint8_t *sp = bytes + src[k];
                for (int j = 0; j < len[k]; j++) o[j] = __ldg(sp + j);
                o += len[k];
            }
        }
    }
}

}  // namespace kta
