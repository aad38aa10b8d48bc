// Device-side core of libfermi_b200: the "occ line" rank, fm6_extend and the per-lane SMEM
// state machine.  Everything here is a __host__ __device__ inline so that tests/emu can compile
// the SAME source for the host and check it against the oracle where no GPU exists; the product
// only ever instantiates these functions inside __global__ kernels (kernels.cu).
//
// Query layout in HBM ("occ lines", built by occ_build.cu from the .fmd image):
//   one line = 128 B = 8 x uint4 covering 256 consecutive BWT symbols
//     uint4 #0,#1 : u32 cnt[6] (+2 pad) = number of $,A,C,G,T,N in BWT[superblock_start, line*256+128)
//                   i.e. cumulative counts at the MIDDLE of the line
//     uint4 #2,#3 : bit plane 0 of symbols   0..127 | 128..255   (bit i of a plane = bit of symbol i)
//     uint4 #4,#5 : bit plane 1
//     uint4 #6,#7 : bit plane 2
//   counts are relative to a 2^31-symbol superblock; `super` holds 8 x u64 absolute counts per
//   superblock (NULL when the whole BWT has < 2^32-256 symbols, then the u32 counts are absolute).
//   A rank touches exactly one line and popcounts at most 128 symbols on one side of the middle:
//   rank(p) = mid -/+ popcount(symbols between p and the middle).  Replaces rld_locate_blk +
//   rld_dec0 run decoding (rld.c:352-446): one 128-byte HBM line per rank instead of a frame row,
//   ~7 block headers and a serial Elias-delta decode.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FMG_HD __host__ __device__ __forceinline__
#else
#define FMG_HD inline
#endif

#if !defined(__CUDACC__)
struct uint4 { uint32_t x, y, z, w; };
#endif

namespace fmg {

constexpr int kLineShift = 8;                 // 256 symbols per line
constexpr int kSuperShift = 31;               // symbols per superblock
constexpr uint64_t kLineU4 = 8;               // uint4 per line

struct OccView {
    const uint4 *lines;
    const uint64_t *super;    // [n_super][8], or nullptr
    uint64_t n_sym;           // mcnt[0]
    uint64_t n_seq;           // mcnt[1]: number of sentinels = sequences (both strands)
    uint64_t C[8];            // C[c] = #symbols < c (rld cnt[], rld.c:233)
};

struct Intv { uint64_t x0, x1, x2, info; };   // fmintv_t (fermi.h:13-16)

FMG_HD uint32_t popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return (uint32_t)__builtin_popcount(v);
#endif
}

FMG_HD uint4 ld_line(const uint4 *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

FMG_HD uint8_t ld_u8(const uint8_t *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// word w (0..3) of the 128-bit mask with the low `oo` bits set
FMG_HD uint32_t low_mask_word(uint32_t oo, int w) {
    const uint32_t lo = 32u * w;
    return oo >= lo + 32u ? 0xffffffffu : (oo <= lo ? 0u : ((1u << (oo - lo)) - 1u));
}

// number of positions where the 3-bit symbol (p2 p1 p0) equals SYM, among the bits of m
template <int SYM>
FMG_HD uint32_t match32(uint32_t p0, uint32_t p1, uint32_t p2m, uint32_t np2m) {
    const uint32_t a = (SYM & 1) ? p0 : ~p0;
    const uint32_t b = (SYM & 2) ? p1 : ~p1;
    return a & b & ((SYM & 4) ? p2m : np2m);
}

struct LineRegs { uint4 c0, c1, a, b, d; };

FMG_HD LineRegs load_line(const OccView &ix, uint64_t p) {
    const uint4 *L = ix.lines + (p >> kLineShift) * kLineU4;
    const uint32_t half = (uint32_t)(p >> 7) & 1u;
    LineRegs r;
    r.c0 = ld_line(L);
    r.c1 = ld_line(L + 1);
    r.a = ld_line(L + 2 + half);
    r.b = ld_line(L + 4 + half);
    r.d = ld_line(L + 6 + half);
    return r;
}

// cnt[c] = number of symbols c in BWT[0, p), for p in [0, n_sym]
FMG_HD void rank_from_line(const OccView &ix, const LineRegs &r, uint64_t p, uint64_t cnt[6]) {
    const uint32_t o = (uint32_t)p & 255u, half = o >> 7, oo = o & 127u;
    const uint32_t flip = half ? 0u : 0xffffffffu;     // below the middle: count [oo,128) and subtract
    const uint32_t m0 = low_mask_word(oo, 0) ^ flip, m1 = low_mask_word(oo, 1) ^ flip;
    const uint32_t m2 = low_mask_word(oo, 2) ^ flip, m3 = low_mask_word(oo, 3) ^ flip;
    const uint32_t q0 = r.d.x & m0, q1 = r.d.y & m1, q2 = r.d.z & m2, q3 = r.d.w & m3;
    const uint32_t n0 = ~r.d.x & m0, n1 = ~r.d.y & m1, n2 = ~r.d.z & m2, n3 = ~r.d.w & m3;
    uint32_t pc[5];
#define FMG_PC(S) pc[S] = popc32(match32<S>(r.a.x, r.b.x, q0, n0)) + popc32(match32<S>(r.a.y, r.b.y, q1, n1)) + \
                          popc32(match32<S>(r.a.z, r.b.z, q2, n2)) + popc32(match32<S>(r.a.w, r.b.w, q3, n3));
    FMG_PC(0) FMG_PC(1) FMG_PC(2) FMG_PC(3) FMG_PC(4)
#undef FMG_PC
    const uint32_t base[5] = { r.c0.x, r.c0.y, r.c0.z, r.c0.w, r.c1.x };
    uint64_t sum = 0;
    const uint64_t *sb = ix.super ? ix.super + ((p >> kSuperShift) << 3) : nullptr;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const uint32_t rel = half ? base[c] + pc[c] : base[c] - pc[c];
        cnt[c] = (sb ? sb[c] : 0ull) + rel;
        sum += cnt[c];
    }
    cnt[5] = p - sum;      // N: the six counts add up to p
}

// All six extensions of one bi-interval: fm6_extend, exact.c:72-88 (A.5).
// far side = x[!is_back]; size[c] = ok[c].x[2]; far[c] = ok[c].x[!is_back]; near[c] = ok[c].x[is_back]
struct Ext6 { uint64_t size[6], far[6], near[6]; };

FMG_HD void extend6(const OccView &ix, uint64_t x_near, uint64_t x_far, uint64_t size, Ext6 &e) {
    // rld_rank2a(x_far-1, x_far-1+size): counts in BWT[0,x_far) and BWT[0,x_far+size)  (k=-1 <=> p=0)
    const uint64_t pk = x_far, pl = x_far + size;
    const LineRegs rk = load_line(ix, pk);
    const LineRegs rl = load_line(ix, pl);
    uint64_t tk[6], tl[6];
    rank_from_line(ix, rk, pk, tk);
    rank_from_line(ix, rl, pl, tl);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        e.size[c] = tl[c] - tk[c];
        e.far[c] = ix.C[c] + tk[c];
    }
    e.near[0] = x_near;                       // cumulative in the order $,T,G,C,A,N (exact.c:81-86)
    e.near[4] = e.near[0] + e.size[0];
    e.near[3] = e.near[4] + e.size[4];
    e.near[2] = e.near[3] + e.size[3];
    e.near[1] = e.near[2] + e.size[2];
    e.near[5] = e.near[1] + e.size[1];
}

FMG_HD uint64_t pick6(const uint64_t v[6], int c) {
    uint64_t r = v[0];
    r = c == 1 ? v[1] : r; r = c == 2 ? v[2] : r; r = c == 3 ? v[3] : r;
    r = c == 4 ? v[4] : r; r = c == 5 ? v[5] : r;
    return r;
}

FMG_HD int comp6(int c) { return (c >= 1 && c <= 4) ? 5 - c : c; }   // fm6_comp, fermi.h:52

// fm6_set_intv, fermi.h:53
FMG_HD Intv base_intv(const OccView &ix, int c) {
    Intv k;
    k.x0 = ix.C[c]; k.x2 = ix.C[c + 1] - ix.C[c]; k.x1 = ix.C[comp6(c)]; k.info = 0;
    return k;
}

FMG_HD Intv ld_intv(const uint4 *p) {
    const uint4 a = p[0], b = p[1];
    Intv k;
    k.x0 = (uint64_t)a.y << 32 | a.x; k.x1 = (uint64_t)a.w << 32 | a.z;
    k.x2 = (uint64_t)b.y << 32 | b.x; k.info = (uint64_t)b.w << 32 | b.z;
    return k;
}

FMG_HD void st_intv(uint4 *p, const Intv &k) {
    uint4 a, b;
    a.x = (uint32_t)k.x0; a.y = (uint32_t)(k.x0 >> 32); a.z = (uint32_t)k.x1; a.w = (uint32_t)(k.x1 >> 32);
    b.x = (uint32_t)k.x2; b.y = (uint32_t)(k.x2 >> 32); b.z = (uint32_t)k.info; b.w = (uint32_t)(k.info >> 32);
    p[0] = a; p[1] = b;
}

// ------------------------------------------------------------------------------------------------
// SMEM: fm6_smem (smem.c:397-410) over fm6_smem1_core (smem.c:13-80) as a per-lane state machine
// that performs exactly ONE fm6_extend per loop iteration, so the 32 lanes of a warp (each on its
// own read, in its own phase) stay converged on the expensive part (two line loads + popcounts).
//
// Per-lane scratch (thread-contiguous, 32 B entries):
//   F[cap]  intervals pushed by the forward sweep, in push order (the reference's `curr` before the
//           reversal at smem.c:45; it is read back to front instead of being reversed)
//   W[cap]  the candidate list of the backward sweep, updated IN PLACE (a pass writes entry n<=j
//           only after it has read entry j; the reference swaps two vectors, smem.c:73)
//   cap >= 2*max_len+2 (at most two pushes per forward step, smem.c:25-30,35-44)
// Records go to a fixed slot of out_cap entries per read (compacted afterwards); rec_cnt[r] is the
// true number of records, so rec_cnt[r] > out_cap flags an overflow that the host re-runs.
struct SmemArgs {
    OccView ix;
    const uint8_t *seq;
    const uint64_t *off;
    int64_t n_reads;
    int self_match;
    uint4 *F, *W;
    int cap;
    uint4 *out;
    int out_cap;
    uint32_t *rec_cnt;
    unsigned long long *next_read;
};

enum { PH_FETCH = 0, PH_BEGIN, PH_START_BWD, PH_FWD, PH_FWD_TAIL, PH_BWD, PH_DONE };

FMG_HD bool warp_any(bool v) {
#if defined(__CUDA_ARCH__)
    return __any_sync(0xffffffffu, v);      // also the reconvergence point of the 32 lanes
#else
    return v;
#endif
}

// Loop shape (per trip):  [divergent, short]  advance the lane's state machine to its next extension request
//                         [warp vote]         leave when no lane has a request; reconverges the warp
//                         [converged, long]   extend6: two line loads + popcounts for all requesting lanes
//                         [divergent, short]  consume the result according to the lane's phase
template <class FetchFn>
FMG_HD void smem_lane(const SmemArgs &A, int64_t lane_slot, FetchFn fetch) {
    uint4 *F = A.F + (size_t)lane_slot * A.cap * 2;
    uint4 *W = A.W + (size_t)lane_slot * A.cap * 2;
    const int sm = A.self_match;
    int ph = PH_FETCH;
    int64_t r = 0;
    const uint8_t *q = nullptr;
    uint4 *out = nullptr;
    int len = 0, x = 0, i = 0, j = 0, nF = 0, nprev = 0, ncurr = 0, first_pass = 0, ret = 0;
    int call_base = 0, nmem = 0, last_start = 0;
    uint64_t last_x2 = 0;
    Intv ik = {0, 0, 0, 0};     // FWD: the interval being extended; BWD: the candidate p being extended

    for (;;) {
        // ---- advance to the next extension request (no index access in here)
        while (ph < PH_FWD) {
            if (ph == PH_FETCH) {
                r = fetch();
                if (r >= A.n_reads) { ph = PH_DONE; break; }
                const uint64_t o = A.off[r];
                len = (int)(A.off[r + 1] - o);
                if (len <= 0) { A.rec_cnt[r] = 0; continue; }
                q = A.seq + o;
                out = A.out + (size_t)r * A.out_cap * 2;
                nmem = 0; x = 0;
                ph = PH_BEGIN;
            } else if (ph == PH_BEGIN) {                // begin fm6_smem1_core at x (smem.c:19-21)
                ik = base_intv(A.ix, ld_u8(q + x));
                ik.info = (uint64_t)(x + 1);
                i = x + 1; nF = 0;
                ph = PH_FWD;
                if (i == len) {                         // smem.c:35-36
                    st_intv(F + 2 * nF++, ik);
                    ph = sm ? PH_START_BWD : PH_FWD_TAIL;
                }
            } else {                                    // PH_START_BWD: forward sweep finished (smem.c:45-50)
                if (nF == 0) {                          // undefined in the reference: no SMEM here
                    x = i < len ? i : len;
                    if (x >= len) { A.rec_cnt[r] = (uint32_t)nmem; ph = PH_FETCH; } else ph = PH_BEGIN;
                    continue;
                }
                ik = ld_intv(F + 2 * (nF - 1));         // the longest match is the last push = first candidate
                ret = (int)ik.info;
                nprev = nF; first_pass = 1; i = x - 1; j = 0; ncurr = 0; call_base = nmem;
                ph = PH_BWD;
            }
        }
        const bool active = ph != PH_DONE;
        if (!warp_any(active)) return;
        if (!active) continue;

        // ---- the one extension of this trip (converged across the warp)
        const int back = (ph == PH_BWD);
        Ext6 e;
        extend6(A.ix, back ? ik.x1 : ik.x0, back ? ik.x0 : ik.x1, ik.x2, e);

        // ---- consume it
        // x[0]/x[1] of ok[c]: the far side is x[1] for a forward, x[0] for a backward extension
#define FMG_OK(c, dst) do { const uint64_t nr_ = pick6(e.near, c), fr_ = pick6(e.far, c); \
                            (dst).x0 = back ? fr_ : nr_; (dst).x1 = back ? nr_ : fr_; (dst).x2 = pick6(e.size, c); } while (0)
        if (ph == PH_FWD) {                             // smem.c:22-34
            const int c = comp6(ld_u8(q + i));
            const uint64_t sc = pick6(e.size, c);
            if (sc != ik.x2) {
                if (ik.x2 != e.size[0]) st_intv(F + 2 * nF++, ik);
                if (!sm && e.size[0]) {
                    Intv s0; s0.x0 = e.near[0]; s0.x1 = e.far[0]; s0.x2 = e.size[0]; s0.info = (uint64_t)i;
                    st_intv(F + 2 * nF++, s0);
                }
            }
            const bool stop = sm ? sc < 2 : sc == 0;
            if (stop) ph = PH_START_BWD;
            else {
                FMG_OK(c, ik); ik.info = (uint64_t)(i + 1);
                if (++i == len) {                       // reached the end of the read (smem.c:35-36)
                    st_intv(F + 2 * nF++, ik);
                    ph = sm ? PH_START_BWD : PH_FWD_TAIL;
                }
            }
        } else if (ph == PH_FWD_TAIL) {                 // smem.c:37-43
            if (e.size[0]) {
                Intv s0; s0.x0 = e.near[0]; s0.x1 = e.far[0]; s0.x2 = e.size[0]; s0.info = (uint64_t)len;
                st_intv(F + 2 * nF++, s0);
            }
            ph = PH_START_BWD;
        } else {                                        // backward sweep, smem.c:51-75; ik is the candidate p
            const int c = i < 0 ? 0 : (int)ld_u8(q + i);
            const uint64_t sc = pick6(e.size, c);
            const bool fl = e.size[0] != 0 && ik.x1 < A.ix.n_seq;
            const bool cont = sm ? sc > 1 : sc != 0;
            if ((!cont || fl || i == -1) && (ncurr == 0 || fl) &&
                (fl || nmem == call_base || i + 1 < last_start)) {
                Intv m = ik;
                m.info |= (uint64_t)(e.size[0] != 0) << 63 | (uint64_t)(i + 1) << 32;
                if (nmem < A.out_cap) st_intv(out + 2 * nmem, m);
                ++nmem; last_start = i + 1;
            }
            if (cont && (ik.x1 < A.ix.n_seq || ncurr == 0 || sc != last_x2)) {
                Intv n; FMG_OK(c, n); n.info = ik.info;
                st_intv(W + 2 * ncurr++, n);
                last_x2 = sc;
            }
            if (++j == nprev) {
                if (ncurr != 0 && i != -1) {            // next backward position over the survivors
                    nprev = ncurr; ncurr = 0; j = 0; --i; first_pass = 0;
                } else {
                    // end of this fm6_smem1_core call: records were pushed by decreasing start (smem.c:76)
                    int lo = call_base, hi = (nmem < A.out_cap ? nmem : A.out_cap) - 1;
                    for (; lo < hi; ++lo, --hi) {
                        const Intv a = ld_intv(out + 2 * lo), b = ld_intv(out + 2 * hi);
                        st_intv(out + 2 * lo, b); st_intv(out + 2 * hi, a);
                    }
                    x = ret;
                    if (x >= len) { A.rec_cnt[r] = (uint32_t)nmem; ph = PH_FETCH; }
                    else ph = PH_BEGIN;
                }
            }
            if (ph == PH_BWD) ik = ld_intv(first_pass ? F + 2 * (nF - 1 - j) : W + 2 * j);   // next candidate
        }
#undef FMG_OK
    }
}

} // namespace fmg
