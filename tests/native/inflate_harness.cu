// TEST INFRASTRUCTURE: runs the product's inflate statements (kafka_topic_analyzer_b200/csrc/kta_inflate.cuh, the
// __host__ __device__ code log_decompress_kernel calls on the GPU) on the host, so that tests/test_inflate_host.py can
// compare them with zlib without a GPU.  stdin: cases of u32 length + gzip member; stdout: per case u8 ok, u32 length, bytes.
#include <cstdio>
#include <cstdint>
#include <vector>

#include "../../kafka_topic_analyzer_b200/csrc/kta_inflate.cuh"

struct HostOut {
    std::vector<uint8_t> out;
    uint64_t cap;
    bool lit(uint8_t b) {
        if (out.size() >= cap) return false;
        out.push_back(b);
        return true;
    }
    bool match(uint32_t dist, uint32_t len) {
        if (dist > out.size() || out.size() + len > cap) return false;
        for (uint32_t i = 0; i < len; i++) out.push_back(out[out.size() - dist]);
        return true;
    }
    bool stored(const uint8_t *src, uint32_t len) {
        if (out.size() + len > cap) return false;
        out.insert(out.end(), src, src + len);
        return true;
    }
};

int main() {
    uint32_t n;
    while (fread(&n, 4, 1, stdin) == 1) {
        std::vector<uint8_t> in(n);
        if (n && fread(in.data(), 1, n, stdin) != n) return 2;
        bool ok = false;
        HostOut o{{}, 0};
        const uint32_t hl = kta::gzip_header_len(in.data(), n);
        if (hl) {
            o.cap = kta::gzip_isize(in.data(), n);
            kta::InfBits s{in.data() + hl, n - hl - 8u, 0u, 0ull, 0, false};
            kta::InfWork w;
            ok = kta::inf_stream(s, o, w, 0) && o.out.size() == o.cap;
        }
        const uint8_t okb = ok ? 1 : 0;
        const uint32_t len = (uint32_t)o.out.size();
        fwrite(&okb, 1, 1, stdout);
        fwrite(&len, 4, 1, stdout);
        if (len) fwrite(o.out.data(), 1, len, stdout);
    }
    return 0;
}
