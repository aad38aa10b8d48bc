// Seed-addressed synthetic data: every value is a pure function of (seed, stream, index)
// through splitmix64, so the GPU run, the CPU baseline and the tests (tests/helpers.py mirrors these
// formulas in numpy) see byte-identical genomes and reads.  Host code, multi-threaded.
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include <algorithm>
#include "../../include/fermi_b200.h"

namespace {

inline uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

inline uint64_t stream_base(uint64_t seed, uint64_t stream) { return splitmix64(seed ^ (stream * 0xD1B54A32D192ED03ull)); }

template <class F> void parallel_for(int64_t n, F f) {
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
    if (n < 1 << 16) nt = 1;
    std::vector<std::thread> th;
    const int64_t step = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t) {
        const int64_t b = t * step, e = std::min<int64_t>(n, b + step);
        if (b >= e) break;
        th.emplace_back([=]() { f(b, e); });
    }
    for (auto &x : th) x.join();
}

inline uint8_t comp(uint8_t c) { return c >= 1 && c <= 4 ? 5 - c : c; }

} // namespace

extern "C" {

void fmg_synth_genome(uint64_t seed, int64_t n, uint8_t *nt6) {
    const uint64_t base = stream_base(seed, 0);
    parallel_for((n + 31) / 32, [=](int64_t b, int64_t e) {
        for (int64_t w = b; w < e; ++w) {
            uint64_t h = splitmix64(base + (uint64_t)w);
            const int64_t lim = std::min<int64_t>(32, n - w * 32);
            for (int64_t k = 0; k < lim; ++k, h >>= 2) nt6[w * 32 + k] = (uint8_t)((h & 3) + 1);
        }
    });
}

void fmg_synth_reads(uint64_t seed, int64_t genome_len, const uint8_t *genome, int64_t n_reads, int len, double err, uint8_t *reads) {
    const uint64_t b1 = stream_base(seed, 1), b2 = stream_base(seed, 2), b3 = stream_base(seed, 3);
    const uint64_t thr = (uint64_t)(err * 4294967296.0);
    const uint64_t span = (uint64_t)(genome_len - len + 1);
    parallel_for(n_reads, [=](int64_t b, int64_t e) {
        for (int64_t r = b; r < e; ++r) {
            const uint64_t start = splitmix64(b1 + (uint64_t)r) % span;
            const bool rev = (splitmix64(b2 + (uint64_t)r) >> 17) & 1;
            uint8_t *out = reads + r * len;
            for (int j = 0; j < len; ++j) {
                uint8_t c = genome[start + j];
                const uint64_t h = splitmix64(b3 + (uint64_t)r * (uint64_t)len + (uint64_t)j);
                if ((h & 0xffffffffull) < thr) c = (uint8_t)(((c - 1 + (h >> 32) % 3 + 1) & 3) + 1);
                out[j] = c;
            }
            if (rev) {
                for (int j = 0; j < len / 2; ++j) {
                    const uint8_t t = comp(out[len - 1 - j]);
                    out[len - 1 - j] = comp(out[j]); out[j] = t;
                }
                if (len & 1) out[len / 2] = comp(out[len / 2]);
            }
        }
    });
}

// r0 $ rc(r0) $ r1 $ rc(r1) $ ... (cmd.c:457-469); an even-length sequence equal to its own reverse
// complement loses its last base first (cmd.c:458-463)
int64_t fmg_fmd_text(int64_t n_seq, int len, const uint8_t *seqs, uint8_t *text) {
    int64_t total = 0;
    std::vector<int64_t> off;
    if (text) off.resize(n_seq);
    std::vector<uint8_t> eff(text ? n_seq : 0);
    for (int64_t i = 0; i < n_seq; ++i) {
        const uint8_t *s = seqs + i * len;
        int l = len;
        if ((l & 1) == 0 && l > 0) {
            int k = 0;
            for (; k < l / 2; ++k) if (s[k] + s[l - 1 - k] != 5) break;
            if (k == l / 2) --l;
        }
        if (text) off[i] = total, eff[i] = (uint8_t)(len - l);
        total += 2 * (int64_t)(l + 1);
    }
    if (!text) return total;
    parallel_for(n_seq, [&](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; ++i) {
            const uint8_t *s = seqs + i * len;
            const int l = len - eff[i];
            uint8_t *o = text + off[i];
            std::memcpy(o, s, l); o[l] = 0;
            for (int j = 0; j < l; ++j) o[l + 1 + j] = comp(s[l - 1 - j]);
            o[2 * l + 1] = 0;
        }
    });
    return total;
}

} // extern "C"
