// Host-side .fmd container of libfermi_b200 (product code; the mirror of rld_t, rld.h:20-39).
#pragma once
#include <cstdint>
#include <cstdio>
#include <vector>
#include <string>

namespace fmg {

constexpr int kAlphabet = 6;          // $ A C G T N (seq.c:12-21)
constexpr int kBlockWords = 8;        // sbits = 3  => 64-byte blocks (cmd.c:380, rld.c:69)
constexpr int kHeaderWords16 = 2;     // 7 x u16 header (rld.c:76)
constexpr int kHeaderWords32 = 4;     // 7 x u32 header (rld.c:77)
constexpr uint64_t kChunkWords = 1ull << 23;   // rld.h:9-10

struct Run { uint64_t len; int sym; };

// The RLD bit stream + sparse rank directory exactly as stored in an "RLD\2" file.
struct FmdImage {
    std::vector<uint64_t> words;      // n_bytes/8 stream words (+2 zero pad words kept past the end)
    std::vector<uint64_t> frame;      // n_frames x 7 (rld.c:186-224)
    uint64_t n_bytes = 0, n_frames = 0;
    uint64_t mcnt[8] = {0};           // [0] total, [c+1] count of symbol c
    uint64_t cnt[8] = {0};            // cnt[c] = #symbols < c
    int ibits = 0;

    uint64_t n_symbols() const { return mcnt[0]; }
    uint64_t n_stream_words() const { return n_bytes >> 3; }
    uint64_t n_blocks() const { return (n_bytes >> 3) / kBlockWords; }   // payload-carrying blocks (last is header-only)

    void finish_counts();             // cnt[] and mcnt[0] from mcnt[1..6]
    void build_frames();              // rld_rank_index, rld.c:186-224
    bool write(const char *fn) const; // rld_dump, rld.c:242-263
    // iterate all runs in BWT order; F(uint64_t len, int sym)
    template <class F> void for_each_run(F &&f) const;
};

// Streaming Elias-delta encoder producing the reference's exact bit layout (rld.c:111-184,226-236).
class FmdEncoder {
public:
    FmdEncoder();
    void put(uint64_t len, int sym);  // rld_enc: adjacent runs of one symbol are merged
    FmdImage finish();                // rld_enc_finish
private:
    void emit(uint64_t len, int sym);
    void open_block();
    void reserve(uint64_t need);
    std::vector<uint64_t> w_;
    uint64_t p_ = kHeaderWords16, head_ = 0, tail_ = 0;
    int room_ = 64;
    int pend_sym_ = -1;
    uint64_t pend_len_ = 0;
    uint64_t tot_[7] = {0}, mark_[7] = {0};
};

FmdImage *load_fmd(const char *fn);   // rld_restore, rld.c:288-325

// ---- bit helpers shared with the decoder ------------------------------------------------------
inline bool header_is32(uint64_t head) { return (uint32_t)head >> 31; }   // rld.h:68
inline uint64_t tail_word(uint64_t blk) {                                   // rld.h:66
    uint64_t end = blk + kBlockWords;
    return end - (((end & (kChunkWords - 1)) == 0) ? 2 : 1);
}

template <class F> void FmdImage::for_each_run(F &&f) const {
    const uint64_t last = n_blocks() * kBlockWords;
    const uint64_t *w = words.data();
    for (uint64_t blk = 0; blk < last; blk += kBlockWords) {
        uint64_t bit = (blk + (header_is32(w[blk]) ? kHeaderWords32 : kHeaderWords16)) * 64;
        const uint64_t end = (tail_word(blk) + 1) * 64;
        while (bit < end) {
            const uint64_t i = bit >> 6; const int s = bit & 63;
            uint64_t x = s ? (w[i] << s) | (w[i + 1] >> (64 - s)) : w[i];
            if (end - bit < 64) x &= ~0ull << (64 - (end - bit));     // nothing is read past the tail word (rld.h:82)
            if (x == 0) break;
            uint64_t len; int sym, used;
            if (x >> 63) { len = 1; sym = (x >> 60) & 7; used = 4; }
            else {
                const int z = __builtin_clzll(x);
                if (z > 5) break;                                       // rld.h:84
                const int g = 2 * z + 1, y = (int)(x >> (64 - g)) - 1;
                len = ((x << g) >> (64 - y)) | (1ull << y);
                sym = (int)((x << (g + y)) >> 61);
                used = g + y + 3;
            }
            if (sym > kAlphabet) break;                                 // rld.h:100
            f(len, sym);
            bit += used;
        }
    }
}

} // namespace fmg
