// Host-side transcoder: RLD runs (.fmd) -> occ blocks.  See fmd_device.cuh for the layout.
#include "occ_layout.hpp"
#include <cstring>

namespace fmg {

// OR the bit range [pos, pos+len) into plane `pl` of the block array
static inline void fill_plane(uint32_t *blocks, int pl, uint64_t pos, uint64_t len) {
    while (len) {
        const uint64_t blk = pos >> 7, within = pos & 127;
        uint32_t *w = blocks + blk * 16 + 4 + pl * 4 + (within >> 5);
        const unsigned bit = within & 31;
        const uint64_t take = (32 - bit) < len ? (32 - bit) : len;
        const uint32_t mask = (take == 32 ? 0xffffffffu : ((1u << take) - 1u)) << bit;
        *w |= mask;
        pos += take; len -= take;
    }
}

OccHost build_occ_host(const FmdImage &img) {
    OccHost o;
    o.n_sym = img.n_symbols();
    o.n_blocks = occ_n_blocks(o.n_sym);
    o.n_super = occ_n_super(o.n_sym);
    o.blocks.assign(o.n_blocks * 16, 0);
    o.cs.assign(o.n_super * 8, 0);
    uint32_t *B = o.blocks.data();
    uint64_t pos = 0;
    img.for_each_run([&](uint64_t len, int sym) {
        for (int pl = 0; pl < 3; ++pl)
            if (sym >> pl & 1) fill_plane(B, pl, pos, len);
        pos += len;
    });
    // positions past the end hold symbol 7, which no rank counts
    for (int pl = 0; pl < 3; ++pl) fill_plane(B, pl, o.n_sym, o.n_blocks * 128 - o.n_sym);

    uint64_t run[6] = {0, 0, 0, 0, 0, 0}, base[6] = {0, 0, 0, 0, 0, 0};
    for (uint64_t b = 0; b < o.n_blocks; ++b) {
        uint32_t *w = B + b * 16;
        if ((b & ((1ull << 17) - 1)) == 0) {                 // 2^24 symbols = 2^17 blocks per superblock
            for (int c = 0; c < 6; ++c) base[c] = run[c], o.cs[(b >> 17) * 8 + c] = img.cnt[c] + run[c];
        }
        uint64_t half[2][6];
        for (int h = 0; h < 2; ++h)
            for (int c = 0; c < 6; ++c) {
                uint64_t n = 0;
                for (int k = 0; k < 2; ++k) {
                    const uint32_t p0 = w[4 + 2 * h + k], p1 = w[8 + 2 * h + k], p2 = w[12 + 2 * h + k];
                    const uint32_t m = ((c & 1) ? p0 : ~p0) & ((c & 2) ? p1 : ~p1) & ((c & 4) ? p2 : ~p2);
                    n += __builtin_popcount(m);
                }
                half[h][c] = n;
            }
        uint8_t *cb = reinterpret_cast<uint8_t *>(w);
        for (int c = 0; c < 5; ++c) {
            const uint32_t v = (uint32_t)(run[c] + half[0][c] - base[c]);     // < 2^24
            cb[3 * c] = v & 0xff; cb[3 * c + 1] = (v >> 8) & 0xff; cb[3 * c + 2] = (v >> 16) & 0xff;
        }
        cb[15] = 0;
        for (int c = 0; c < 6; ++c) run[c] += half[0][c] + half[1][c];
    }
    return o;
}

} // namespace fmg
