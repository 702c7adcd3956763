// TEST INFRASTRUCTURE: the LZ4-frame / Snappy / gzip walks of the RecordBatch decoder (csrc/kta_logdecode.cuh: the
// __host__ __device__ statements log_unc_size_kernel and log_decompress_kernel run on the GPU) on the host, one "lane".
// stdin: cases of u8 codec (1 gzip, 2 snappy, 3 lz4) + u32 length + bytes; stdout per case: u8 ok, u32 size-pass length,
// u32 length, bytes.  The output buffer is allocated at exactly the size pass's length, so that an overrun of the copy
// pass is a heap overflow an address-sanitizer build reports.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../kafka_topic_analyzer_b200/csrc/kta_logdecode.cuh"

struct HostInflateOut {
    uint8_t *out;
    uint64_t op, cap;
    bool lit(uint8_t b) {
        if (op >= cap) return false;
        out[op++] = b;
        return true;
    }
    bool match(uint32_t dist, uint32_t len) {
        if (dist > op || op + len > cap) return false;
        kta::lz_emit_match<true>(out, op, dist, len, 0);
        op += len;
        return true;
    }
    bool stored(const uint8_t *src, uint32_t len) {
        if (op + len > cap) return false;
        kta::lz_emit_literals<true>(out, op, src, len, 0);
        op += len;
        return true;
    }
};

int main() {
    uint8_t codec;
    uint32_t n;
    while (fread(&codec, 1, 1, stdin) == 1 && fread(&n, 4, 1, stdin) == 1) {
        // exact-size heap copies: reads past the input are heap overflows too
        uint8_t *in = (uint8_t *)malloc(n ? n : 1);
        if (n && fread(in, 1, n, stdin) != n) return 2;
        kta::LzWalk size{0, false}, copy{0, false};
        if (codec == 1) {
            if (kta::gzip_header_len(in, n) && (uint64_t)kta::gzip_isize(in, n) <= (uint64_t)n * 1032u + 64u) size = kta::LzWalk{kta::gzip_isize(in, n), true};
        } else if (codec == 2) size = kta::snappy_walk<false>(in, n, nullptr, 0, 0);
        else size = kta::lz4_frame_walk<false>(in, n, nullptr, 0, 0);
        uint8_t *out = (uint8_t *)malloc(size.out_len ? size.out_len : 1);
        if (size.ok && size.out_len <= (64u << 20)) {
            if (codec == 1) {
                const uint32_t hl = kta::gzip_header_len(in, n);
                kta::InfBits s{in + hl, n - hl - 8u, 0u, 0ull, 0, false};
                HostInflateOut o{out, 0, size.out_len};
                kta::InfWork w;
                const bool ok = kta::inf_stream(s, o, w, 0);
                copy = kta::LzWalk{o.op, ok && o.op == size.out_len};
            } else if (codec == 2) copy = kta::snappy_walk<true>(in, n, out, size.out_len, 0);
            else copy = kta::lz4_frame_walk<true>(in, n, out, size.out_len, 0);
        }
        const uint8_t okb = size.ok && copy.ok && copy.out_len == size.out_len ? 1 : 0;
        const uint32_t sl = (uint32_t)size.out_len, len = okb ? (uint32_t)copy.out_len : 0;
        fwrite(&okb, 1, 1, stdout);
        fwrite(&sl, 4, 1, stdout);
        fwrite(&len, 4, 1, stdout);
        if (len) fwrite(out, 1, len, stdout);
        free(out);
        free(in);
    }
    return 0;
}
