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
// LogAppendTime batches (attributes bit 3) stamp every record with maxTimestamp; a timestamp of -1 means
// "not available"; key/value length -1 means null.  CRCs are not verified (librdkafka's default check.crcs=false).
// Compressed batches (attributes bits 0-2) are rejected: no decompressor here.
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

// unsigned LEB128 at p (bounded by end); returns bytes consumed, 0 on malformed input
__device__ __forceinline__ int uvarint(const uint8_t *p, const uint8_t *end, uint64_t &out) {
    uint64_t v = 0;
    int shift = 0, n = 0;
    while (p + n < end && n < 10) {
        const uint8_t b = __ldg(p + n);
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
                                  int32_t partition, LogBatchInfo *info, uint64_t *rec_count /*[nbatches+1], [b+1]*/,
                                  uint32_t *error_flags) {
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nbatches; b += (int64_t)gridDim.x * blockDim.x) {
        LogBatchInfo bi{};
        bi.off = batch_off[b];
        bi.partition = partition;
        bi.flags = LOGB_BAD;
        if (bi.off + LOG_HEADER_BYTES <= (uint64_t)nbytes) {
            const uint8_t *p = bytes + bi.off;
            const int32_t batch_len = (int32_t)be_u32(p + 8);
            const int magic = (int8_t)__ldg(p + 16);
            const uint32_t attrs = be_u16(p + 21);
            const int32_t count = (int32_t)be_u32(p + 57);
            if (magic == 2 && batch_len >= LOG_HEADER_BYTES - 12 && bi.off + 12 + (uint64_t)batch_len <= (uint64_t)nbytes && count >= 0) {
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
        info[b] = bi;
        rec_count[b + 1] = (uint64_t)bi.records;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) rec_count[0] = 0;
}

// thread per batch walks its records.  PASS 0: header columns + key-byte total of the batch; PASS 1: copy the keys.
template <int PASS>
__global__ void log_decode_kernel(const uint8_t *bytes, const LogBatchInfo *info, int64_t nbatches, const uint64_t *rec_base,
                                  int32_t *partition, int64_t *offset, int64_t *ts_ms, int32_t *key_len, int32_t *value_len,
                                  uint64_t *key_total /*[nbatches+1], PASS 0 out, PASS 1 in as exclusive bases*/,
                                  uint8_t *key_out, uint32_t *error_flags) {
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nbatches; b += (int64_t)gridDim.x * blockDim.x) {
        const LogBatchInfo bi = info[b];
        uint64_t kbytes = 0;
        if (bi.flags == LOGB_OK && bi.records > 0) {
            const uint8_t *p = bytes + bi.off + LOG_HEADER_BYTES;
            const uint8_t *end = bytes + bi.off + bi.len;
            uint64_t r = rec_base[b];
            uint64_t kdst = PASS == 1 ? key_total[b] : 0;
            bool ok = true;
            for (int32_t i = 0; i < bi.records && ok; i++, r++) {
                uint64_t u;
                int n = uvarint(p, end, u);
                const int64_t rec_len = unzigzag(u);
                ok = n > 0 && rec_len >= 0 && p + n + rec_len <= end;
                if (!ok) break;
                const uint8_t *q = p + n, *rec_end = q + rec_len;
                p = rec_end;
                q += 1;  // record attributes (unused)
                n = uvarint(q, rec_end, u); ok = ok && n > 0; q += n;
                const int64_t ts_delta = unzigzag(u);
                n = uvarint(q, rec_end, u); ok = ok && n > 0; q += n;
                const int64_t off_delta = unzigzag(u);
                n = uvarint(q, rec_end, u); ok = ok && n > 0; q += n;
                const int64_t klen = unzigzag(u);
                ok = ok && klen >= -1 && klen <= 0x7fffffff && (klen <= 0 || q + klen <= rec_end);
                if (!ok) break;
                const uint8_t *key = q;
                if (klen > 0) q += klen;
                n = uvarint(q, rec_end, u); ok = ok && n > 0; q += n;
                const int64_t vlen = unzigzag(u);
                ok = ok && vlen >= -1 && vlen <= 0x7fffffff && (vlen <= 0 || q + vlen <= rec_end);
                if (!ok) break;
                if (PASS == 0) {
                    const int64_t t = bi.log_append_time ? bi.max_ts : (bi.base_ts == -1 ? -1 : bi.base_ts + ts_delta);
                    partition[r] = bi.partition;
                    if (offset) offset[r] = bi.base_offset + off_delta;
                    ts_ms[r] = t;
                    key_len[r] = (int32_t)klen;
                    value_len[r] = (int32_t)vlen;
                    if (klen > 0) kbytes += (uint64_t)klen;
                } else if (klen > 0) {
                    for (int64_t j = 0; j < klen; j++) key_out[kdst + j] = __ldg(key + j);
                    kdst += (uint64_t)klen;
                }
            }
            if (!ok) atomicOr(error_flags, (uint32_t)LOGB_BAD);
        }
        if (PASS == 0) key_total[b + 1] = kbytes;
    }
    if (PASS == 0 && blockIdx.x == 0 && threadIdx.x == 0) key_total[0] = 0;
}

}  // namespace kta
