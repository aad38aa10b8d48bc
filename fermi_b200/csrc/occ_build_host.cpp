// Host-side transcoder: RLD runs (.fmd) -> occ lines.  See fmd_device.cuh for the layout.
#include "occ_layout.hpp"
#include <cstring>

namespace fmg {

// OR the bit range [pos, pos+len) into plane `pl` of the line array
static inline void fill_plane(uint64_t *lines, int pl, uint64_t pos, uint64_t len) {
    while (len) {
        const uint64_t line = pos >> 8, within = pos & 255;
        uint64_t *w = lines + line * 16 + 4 + pl * 4 + (within >> 6);
        const unsigned bit = within & 63;
        const uint64_t take = (64 - bit) < len ? (64 - bit) : len;
        const uint64_t mask = (take == 64 ? ~0ull : ((1ull << take) - 1)) << bit;
        *w |= mask;
        pos += take; len -= take;
    }
}

OccHost build_occ_host(const FmdImage &img) {
    OccHost o;
    o.n_sym = img.n_symbols();
    o.n_lines = occ_n_lines(o.n_sym);
    o.lines.assign(o.n_lines * 16, 0);
    uint64_t *L = o.lines.data();
    uint64_t pos = 0;
    img.for_each_run([&](uint64_t len, int sym) {
        for (int pl = 0; pl < 3; ++pl)
            if (sym >> pl & 1) fill_plane(L, pl, pos, len);
        pos += len;
    });
    // positions past the end hold symbol 7, which no rank counts
    for (int pl = 0; pl < 3; ++pl) fill_plane(L, pl, o.n_sym, o.n_lines * 256 - o.n_sym);

    const bool big = occ_needs_super(o.n_sym);
    if (big) o.super.assign(((o.n_sym >> 31) + 1) * 8, 0);
    uint64_t run[6] = {0, 0, 0, 0, 0, 0}, base[6] = {0, 0, 0, 0, 0, 0};
    for (uint64_t ln = 0; ln < o.n_lines; ++ln) {
        uint64_t *w = L + ln * 16;
        if (big && (ln & ((1ull << 23) - 1)) == 0) {
            for (int c = 0; c < 6; ++c) base[c] = run[c], o.super[(ln >> 23) * 8 + c] = run[c];
        }
        uint64_t half[2][6];
        for (int h = 0; h < 2; ++h)
            for (int c = 0; c < 6; ++c) {
                uint64_t n = 0;
                for (int k = 0; k < 2; ++k) {
                    const uint64_t p0 = w[4 + 2 * h + k], p1 = w[8 + 2 * h + k], p2 = w[12 + 2 * h + k];
                    const uint64_t m = ((c & 1) ? p0 : ~p0) & ((c & 2) ? p1 : ~p1) & ((c & 4) ? p2 : ~p2);
                    n += __builtin_popcountll(m);
                }
                half[h][c] = n;
            }
        uint32_t *cw = reinterpret_cast<uint32_t *>(w);
        for (int c = 0; c < 6; ++c) {
            cw[c] = (uint32_t)(run[c] + half[0][c] - base[c]);
            run[c] += half[0][c] + half[1][c];
        }
        cw[6] = cw[7] = 0;
    }
    return o;
}

} // namespace fmg
