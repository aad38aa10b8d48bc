// Device-side core of libfermi_b200: the "occ block" rank, fm6_extend and the per-lane SMEM
// state machine.  Everything here is a __host__ __device__ inline so that tests/emu can compile
// the SAME source for the host and check it against the oracle where no GPU exists; the product
// only ever instantiates these functions inside __global__ kernels (fmg_cuda.cu).
//
// Query layout in HBM ("occ blocks", built from the .fmd image at upload):
//   one block = 64 B = 16 x u32 covering 128 consecutive BWT symbols, read with two 256-bit loads
//     w[0..3]   five 24-bit counts, byte-packed ($ A C G T; byte 15 spare): number of each symbol in
//               BWT[superblock_start, 128*b + 64), i.e. cumulative counts at the MIDDLE of the block,
//               relative to the 2^24-symbol superblock the block lies in
//     w[4..7]   bit plane 0 of the 128 symbols (bit i of a plane = bit of symbol i)
//     w[8..11]  bit plane 1
//     w[12..15] bit plane 2          (positions past the end of the BWT hold symbol 7)
//   cs[sb][c] (u64, 8 per superblock) = C[c] + number of c before the superblock, so that the SA
//   coordinate C[c] + rank(c, p) is cs[sb][c] + rel[c].
//   A rank touches exactly one 64-byte block and popcounts at most 64 symbols on one side of the
//   middle: rel(p) = mid -/+ popcount(symbols between p and the middle); N is derived from the total.
//   Replaces rld_locate_blk + rld_dec0 run decoding (rld.c:352-446): one 64-byte HBM access per rank
//   instead of a frame row, ~7 block headers in distinct lines and a serial Elias-delta decode.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define FMG_HD __host__ __device__ __forceinline__
#else
#define FMG_HD inline
#endif

#if !defined(__CUDACC__)
struct uint4 { uint32_t x, y, z, w; };
#endif

namespace fmg {

constexpr int kBlkShift = 7;                  // 128 symbols per occ block
constexpr int kSuperShift = 24;               // symbols per superblock
constexpr uint64_t kBlkWords = 16;            // u32 per block

struct OccView {
    const uint32_t *blocks;   // n_blocks x 16 u32, 64-byte aligned
    const uint64_t *cs;       // [n_super][8]
    uint64_t n_sym;           // mcnt[0]
    uint64_t n_seq;           // mcnt[1]: number of sentinels = sequences (both strands)
    uint64_t C[8];            // C[c] = #symbols < c (rld cnt[], rld.c:233)
};

struct Intv { uint64_t x0, x1, x2, info; };   // fmintv_t (fermi.h:13-16)
struct alignas(32) Vec8 { uint32_t v[8]; };

FMG_HD uint32_t popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return (uint32_t)__builtin_popcount(v);
#endif
}

// 256-bit read-only load (LDG.E.256 on sm_100a)
FMG_HD Vec8 ld256_nc(const void *p) {
    Vec8 r;
#if defined(__CUDA_ARCH__)
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
        : "l"(p));
#else
    memcpy(&r, p, 32);
#endif
    return r;
}

// 256-bit load / store of read-write scratch
FMG_HD Vec8 ld256(const void *p) {
    Vec8 r;
#if defined(__CUDA_ARCH__)
    asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
                 : "l"(p) : "memory");
#else
    memcpy(&r, p, 32);
#endif
    return r;
}

FMG_HD void st256(void *p, const Vec8 &r) {
#if defined(__CUDA_ARCH__)
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]),
                 "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]) : "memory");
#else
    memcpy(p, &r, 32);
#endif
}

FMG_HD uint8_t ld_u8(const uint8_t *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

FMG_HD uint64_t ld_u64(const uint64_t *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

struct Blk { Vec8 lo, hi; };       // lo = counts + plane 0, hi = planes 1 and 2

FMG_HD Blk load_blk(const OccView &ix, uint64_t p) {
    const uint32_t *b = ix.blocks + (p >> kBlkShift) * kBlkWords;
    Blk r;
    r.lo = ld256_nc(b);
    r.hi = ld256_nc(b + 8);
    return r;
}

// 64-byte block gather for a CONVERGED warp (all 32 lanes call it; `need` = this lane wants the block at position p).
// A warp-level load instruction costs one L1-miss request per distinct 128-byte line it touches, and the rate of those requests
// -- not bytes -- is what bounds dependent random gathers from an index that does not fit L2 (tools/micro/gather_bench2.cu on a
// B200, 1 GB array: two ld.v8 per thread 27 G blocks/s; both 32-byte halves of a block requested by two adjacent lanes in ONE
// instruction 53 G blocks/s).  So lanes work in pairs: instruction 1 fetches the blocks of the even lanes (even lane: lower
// half, odd lane: upper half), instruction 2 those of the odd lanes, and eight shuffles hand each lane the half it lacks.
// The gather is split in two so that several of them can be in flight: pair_issue puts the two loads of a block on the wire,
// pair_finish (which waits for the data) shuffles the halves into place.
struct PairReq { Vec8 r0, r1; };
FMG_HD PairReq pair_issue(const OccView &ix, uint64_t p, bool need) {
    PairReq q;
#if defined(__CUDA_ARCH__)
    constexpr uint32_t kNoBlk = 0xffffffffu;                       // block numbers fit 32 bits (checked at upload)
    const uint32_t b = need ? (uint32_t)(p >> kBlkShift) : kNoBlk;
    const uint32_t pb = __shfl_xor_sync(0xffffffffu, b, 1);
    const uint32_t odd = threadIdx.x & 1u;
    const uint32_t be = odd ? pb : b, bo = odd ? b : pb;
    const Vec8 z = {{0, 0, 0, 0, 0, 0, 0, 0}};
    q.r0 = z; q.r1 = z;
    if (be != kNoBlk) q.r0 = ld256_nc(ix.blocks + (uint64_t)be * kBlkWords + odd * 8u);
    if (bo != kNoBlk) q.r1 = ld256_nc(ix.blocks + (uint64_t)bo * kBlkWords + odd * 8u);
#else
    if (need) { const Blk B = load_blk(ix, p); q.r0 = B.lo; q.r1 = B.hi; }
    else { memset(&q, 0, sizeof q); }
#endif
    return q;
}
FMG_HD Blk pair_finish(const PairReq &q) {
    Blk r;
#if defined(__CUDA_ARCH__)
    const uint32_t odd = threadIdx.x & 1u;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t got = __shfl_xor_sync(0xffffffffu, odd ? q.r0.v[i] : q.r1.v[i], 1);
        r.lo.v[i] = odd ? got : q.r0.v[i];
        r.hi.v[i] = odd ? q.r1.v[i] : got;
    }
#else
    r.lo = q.r0; r.hi = q.r1;
#endif
    return r;
}
FMG_HD Blk load_blk_pair(const OccView &ix, uint64_t p, bool need) { return pair_finish(pair_issue(ix, p, need)); }

// positions where the 3-bit symbol (p2 p1 p0) equals SYM; p2m / np2m are p2 / ~p2 already ANDed with the mask
template <int SYM>
FMG_HD uint32_t match32(uint32_t p0, uint32_t p1, uint32_t p2m, uint32_t np2m) {
    const uint32_t a = (SYM & 1) ? p0 : ~p0;
    const uint32_t b = (SYM & 2) ? p1 : ~p1;
    return a & b & ((SYM & 4) ? p2m : np2m);
}

// rel[c] = number of symbols c in BWT[superblock_start(p), p), c = 0..5, for p in [0, n_sym]
FMG_HD void rank_rel(const Blk &B, uint64_t p, uint32_t rel[6]) {
    const uint32_t o = (uint32_t)p & 127u, half = o >> 6, oo = o & 63u;
    // the 64-symbol half that holds p: words 2*half, 2*half+1 of each plane
    const uint32_t a0 = half ? B.lo.v[6] : B.lo.v[4], a1 = half ? B.lo.v[7] : B.lo.v[5];
    const uint32_t b0 = half ? B.hi.v[2] : B.hi.v[0], b1 = half ? B.hi.v[3] : B.hi.v[1];
    const uint32_t d0 = half ? B.hi.v[6] : B.hi.v[4], d1 = half ? B.hi.v[7] : B.hi.v[5];
    // upper half: count [64, 64+oo) and add; lower half: count [oo, 64) and subtract
    const uint32_t lo0 = oo >= 32u ? 0xffffffffu : ((1u << oo) - 1u);
    const uint32_t lo1 = oo <= 32u ? 0u : ((1u << (oo - 32u)) - 1u);
    const uint32_t flip = half ? 0u : 0xffffffffu;
    const uint32_t m0 = lo0 ^ flip, m1 = lo1 ^ flip;
    const uint32_t q0 = d0 & m0, q1 = d1 & m1, n0 = ~d0 & m0, n1 = ~d1 & m1;
    uint32_t pc[5];
#define FMG_PC(S) pc[S] = popc32(match32<S>(a0, b0, q0, n0)) + popc32(match32<S>(a1, b1, q1, n1));
    FMG_PC(0) FMG_PC(1) FMG_PC(2) FMG_PC(3) FMG_PC(4)
#undef FMG_PC
    // five byte-packed 24-bit counts in w[0..3]
    const uint32_t w0 = B.lo.v[0], w1 = B.lo.v[1], w2 = B.lo.v[2], w3 = B.lo.v[3];
    uint32_t base[5];
    base[0] = w0 & 0xffffffu;
    base[1] = (w0 >> 24) | ((w1 & 0xffffu) << 8);
    base[2] = (w1 >> 16) | ((w2 & 0xffu) << 16);
    base[3] = w2 >> 8;
    base[4] = w3 & 0xffffffu;
    uint32_t sum = 0;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        rel[c] = half ? base[c] + pc[c] : base[c] - pc[c];
        sum += rel[c];
    }
    rel[5] = ((uint32_t)p & ((1u << kSuperShift) - 1u)) - sum;     // N: the six counts add up to p - superblock_start
}

// All six extensions of one bi-interval: fm6_extend, exact.c:72-88 (A.5).
//   size[c] = ok[c].x[2];  near[c] = ok[c].x[is_back];  ok[c].x[!is_back] = cs[sbk][c] + relk[c] (far_of)
// U is the coordinate type: uint32_t when the whole BWT has < 2^32 symbols (half the registers, ALU
// work and scratch bytes), uint64_t otherwise.
template <typename U> struct IntvT { U x0, x1, x2, info; };
template <typename U> struct Ext6T { U size[6], near[6]; uint32_t relk[6]; uint64_t sbk; };
typedef Ext6T<uint64_t> Ext6;

// the arithmetic of extend6 on blocks that are already in registers: bk holds position pk = x_far, bl holds pl = x_far + size
template <typename U>
FMG_HD void extend6_with(const OccView &ix, U x_near, uint64_t pk, uint64_t pl, const Blk &bk, const Blk &bl, Ext6T<U> &e) {
    uint32_t rl[6];
    rank_rel(bk, pk, e.relk);
    rank_rel(bl, pl, rl);
    // a position that is a multiple of 2^24 belongs to the superblock it starts (all its counts are 0 there)
    const uint64_t sbk = pk >> kSuperShift, sbl = pl >> kSuperShift;
    e.sbk = sbk;
#pragma unroll
    for (int c = 0; c < 6; ++c) e.size[c] = (U)rl[c] - (U)e.relk[c];
    if (sbk != sbl) {                               // rare: the interval straddles a superblock boundary
        const uint64_t *ck = ix.cs + sbk * 8, *cl = ix.cs + sbl * 8;
#pragma unroll
        for (int c = 0; c < 6; ++c) e.size[c] += (U)(ld_u64(cl + c) - ld_u64(ck + c));
    }
    e.near[0] = x_near;                       // cumulative in the order $,T,G,C,A,N (exact.c:81-86)
    e.near[4] = e.near[0] + e.size[0];
    e.near[3] = e.near[4] + e.size[4];
    e.near[2] = e.near[3] + e.size[3];
    e.near[1] = e.near[2] + e.size[2];
    e.near[5] = e.near[1] + e.size[1];
}

// the same from the two rank results (rel counts at pk = x_far and pl = x_far + size)
template <typename U>
FMG_HD void extend6_rel(const OccView &ix, U x_near, uint64_t pk, uint64_t pl, const uint32_t rk[6], const uint32_t rl[6], Ext6T<U> &e) {
    const uint64_t sbk = pk >> kSuperShift, sbl = pl >> kSuperShift;
    e.sbk = sbk;
#pragma unroll
    for (int c = 0; c < 6; ++c) { e.relk[c] = rk[c]; e.size[c] = (U)rl[c] - (U)rk[c]; }
    if (sbk != sbl) {
        const uint64_t *ck = ix.cs + sbk * 8, *cl = ix.cs + sbl * 8;
#pragma unroll
        for (int c = 0; c < 6; ++c) e.size[c] += (U)(ld_u64(cl + c) - ld_u64(ck + c));
    }
    e.near[0] = x_near;
    e.near[4] = e.near[0] + e.size[0];
    e.near[3] = e.near[4] + e.size[4];
    e.near[2] = e.near[3] + e.size[3];
    e.near[1] = e.near[2] + e.size[2];
    e.near[5] = e.near[1] + e.size[1];
}

FMG_HD bool warp_any(bool v) {
#if defined(__CUDA_ARCH__)
    return __any_sync(0xffffffffu, v);      // also the reconvergence point of the 32 lanes
#else
    return v;
#endif
}

// Re-form the full warp.  Lanes that took different branches run as separate groups until a convergence barrier, and a barrier
// the compiler places ends with the function it is in: after a call, or around a state-machine loop, the groups would execute the
// same code one after the other.
FMG_HD void warp_rejoin() {
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
}

// fm6_extend for a CONVERGED warp: every lane calls, `active` lanes get their extension; the blocks come through
// load_blk_pair (half the L1-miss requests of extend6)
template <typename U>
FMG_HD void extend6_conv(const OccView &ix, bool active, U x_near, U x_far, U size, Ext6T<U> &e) {
    const uint64_t pk = x_far, pl = (uint64_t)x_far + size;
    const bool two = active && (pk >> kBlkShift) != (pl >> kBlkShift);
    // one lane in a dozen needs a second block, so most warps do: both gathers are issued before either is waited for
    const bool any_two = warp_any(two);
    const PairReq qk = pair_issue(ix, pk, active);
    PairReq ql;
    if (any_two) ql = pair_issue(ix, pl, two);
    uint32_t rk[6], rl[6];
    {
        const Blk bk = pair_finish(qk);
        if (active) { rank_rel(bk, pk, rk); if (!two) rank_rel(bk, pl, rl); }
    }
    if (any_two) {
        const Blk bl = pair_finish(ql);
        if (two) rank_rel(bl, pl, rl);
    }
    if (active) extend6_rel<U>(ix, x_near, pk, pl, rk, rl, e);
}

template <typename U>
FMG_HD void extend6(const OccView &ix, U x_near, U x_far, U size, Ext6T<U> &e) {
    // rld_rank2a(x_far-1, x_far-1+size): counts in BWT[0,x_far) and BWT[0,x_far+size)  (k=-1 <=> p=0)
    const uint64_t pk = x_far, pl = (uint64_t)x_far + size;
    const bool same = (pk >> kBlkShift) == (pl >> kBlkShift);
    const Blk bk = load_blk(ix, pk);
    Blk bl = bk;
    if (!same) bl = load_blk(ix, pl);              // small intervals: both ranks read the same 64 bytes
    extend6_with<U>(ix, x_near, pk, pl, bk, bl, e);
}

// BWT[k] from the block that holds position k
FMG_HD int blk_symbol(const Blk &B, uint64_t k) {
    const uint32_t w = ((uint32_t)k >> 5) & 3u, bit = (uint32_t)k & 31u;
    const uint32_t p0 = w == 0 ? B.lo.v[4] : w == 1 ? B.lo.v[5] : w == 2 ? B.lo.v[6] : B.lo.v[7];
    const uint32_t p1 = w == 0 ? B.hi.v[0] : w == 1 ? B.hi.v[1] : w == 2 ? B.hi.v[2] : B.hi.v[3];
    const uint32_t p2 = w == 0 ? B.hi.v[4] : w == 1 ? B.hi.v[5] : w == 2 ? B.hi.v[6] : B.hi.v[7];
    return (int)((p0 >> bit & 1u) | (p1 >> bit & 1u) << 1 | (p2 >> bit & 1u) << 2);
}

template <typename T>
FMG_HD T pick6(const T v[6], int c) {
    T r = v[0];
    r = c == 1 ? v[1] : r; r = c == 2 ? v[2] : r; r = c == 3 ? v[3] : r;
    r = c == 4 ? v[4] : r; r = c == 5 ? v[5] : r;
    return r;
}

// ok[c].x[!is_back] = C[c] + rank(c, x_far)
template <typename U>
FMG_HD U far_of(const OccView &ix, const Ext6T<U> &e, int c) { return (U)(ld_u64(ix.cs + e.sbk * 8 + c) + pick6(e.relk, c)); }

FMG_HD int comp6(int c) { return (c >= 1 && c <= 4) ? 5 - c : c; }   // fm6_comp, fermi.h:52

// fm6_set_intv, fermi.h:53
template <typename U>
FMG_HD IntvT<U> base_intv(const OccView &ix, int c) {
    IntvT<U> k;
    k.x0 = (U)ix.C[c]; k.x2 = (U)(ix.C[c + 1] - ix.C[c]); k.x1 = (U)ix.C[comp6(c)]; k.info = 0;
    return k;
}

// 32-byte records (fmintv_t) in global memory
FMG_HD Intv ld_intv(const uint4 *p) {
    const Vec8 a = ld256(p);
    Intv k;
    k.x0 = (uint64_t)a.v[1] << 32 | a.v[0]; k.x1 = (uint64_t)a.v[3] << 32 | a.v[2];
    k.x2 = (uint64_t)a.v[5] << 32 | a.v[4]; k.info = (uint64_t)a.v[7] << 32 | a.v[6];
    return k;
}

FMG_HD void st_intv(uint4 *p, const Intv &k) {
    Vec8 a;
    a.v[0] = (uint32_t)k.x0; a.v[1] = (uint32_t)(k.x0 >> 32); a.v[2] = (uint32_t)k.x1; a.v[3] = (uint32_t)(k.x1 >> 32);
    a.v[4] = (uint32_t)k.x2; a.v[5] = (uint32_t)(k.x2 >> 32); a.v[6] = (uint32_t)k.info; a.v[7] = (uint32_t)(k.info >> 32);
    st256(p, a);
}

// candidate entries of the per-lane scratch lists: 4 x U (32 bytes for u64, 16 bytes for u32 coordinates)
FMG_HD IntvT<uint64_t> ld_cand(const IntvT<uint64_t> *p) {
    const Intv k = ld_intv(reinterpret_cast<const uint4 *>(p));
    IntvT<uint64_t> r; r.x0 = k.x0; r.x1 = k.x1; r.x2 = k.x2; r.info = k.info;
    return r;
}
FMG_HD void st_cand(IntvT<uint64_t> *p, const IntvT<uint64_t> &k) {
    Intv r; r.x0 = k.x0; r.x1 = k.x1; r.x2 = k.x2; r.info = k.info;
    st_intv(reinterpret_cast<uint4 *>(p), r);
}
FMG_HD IntvT<uint32_t> ld_cand(const IntvT<uint32_t> *p) {
    const uint4 a = *reinterpret_cast<const uint4 *>(p);
    IntvT<uint32_t> r; r.x0 = a.x; r.x1 = a.y; r.x2 = a.z; r.info = a.w;
    return r;
}
FMG_HD void st_cand(IntvT<uint32_t> *p, const IntvT<uint32_t> &k) {
    uint4 a; a.x = k.x0; a.y = k.x1; a.z = k.x2; a.w = k.info;
    *reinterpret_cast<uint4 *>(p) = a;
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
    void *F, *W;              // per-lane candidate lists: cap entries of 4 x U each
    int cap;
    uint4 *out;
    int out_cap;
    uint32_t *rec_cnt;
    unsigned long long *next_read;
    int max_len;                      // reads longer than this do not fit F / W: they are skipped and counted in *too_long
    unsigned long long *too_long;
};

enum { PH_FETCH = 0, PH_BEGIN, PH_START_BWD, PH_FWD, PH_FWD_TAIL, PH_BWD, PH_DONE };

// Loop shape (per trip):  [divergent, short]  advance the lane's state machine to its next extension request
//                         [warp vote]         leave when no lane has a request; reconverges the warp
//                         [converged, long]   extend6: two line loads + popcounts for all requesting lanes
//                         [divergent, short]  consume the result according to the lane's phase
// PAIR: the blocks come through the paired gather (extend6_conv) -- for indexes that do not fit L2, where the rate of L1-miss
// requests bounds the kernel; an L2-resident index is served as fast by plain loads with fewer instructions.
template <typename U, bool PAIR = false, class FetchFn>
FMG_HD void smem_lane(const SmemArgs &A, int64_t lane_slot, FetchFn fetch) {
    typedef IntvT<U> Cand;
    Cand *F = static_cast<Cand *>(A.F) + (size_t)lane_slot * A.cap;
    Cand *W = static_cast<Cand *>(A.W) + (size_t)lane_slot * A.cap;
    const int sm = A.self_match;
    int ph = PH_FETCH;
    int64_t r = 0;
    const uint8_t *q = nullptr;
    uint4 *out = nullptr;
    int len = 0, x = 0, i = 0, j = 0, nF = 0, nprev = 0, ncurr = 0, first_pass = 0, ret = 0;
    int call_base = 0, nmem = 0, last_start = 0;
    U last_x2 = 0;
    Cand ik = {0, 0, 0, 0};     // FWD: the interval being extended; BWD: the candidate p being extended

    for (;;) {
        // ---- advance to the next extension request (no index access in here)
        while (ph < PH_FWD) {
            if (ph == PH_FETCH) {
                r = fetch();
                if (r >= A.n_reads) { ph = PH_DONE; break; }
                const uint64_t o = A.off[r];
                len = (int)(A.off[r + 1] - o);
                if (len <= 0) { A.rec_cnt[r] = 0; continue; }
                if (len > A.max_len) {                  // would overrun the lane's lists: report it, never corrupt memory
                    A.rec_cnt[r] = 0;
#if defined(__CUDA_ARCH__)
                    atomicAdd(A.too_long, 1ull);
#else
                    ++*A.too_long;
#endif
                    continue;
                }
                q = A.seq + o;
                out = A.out + (size_t)r * A.out_cap * 2;
                nmem = 0; x = 0;
                ph = PH_BEGIN;
            } else if (ph == PH_BEGIN) {                // begin fm6_smem1_core at x (smem.c:19-21)
                ik = base_intv<U>(A.ix, ld_u8(q + x));
                ik.info = (U)(x + 1);
                i = x + 1; nF = 0;
                ph = PH_FWD;
                if (i == len) {                         // smem.c:35-36
                    st_cand(F + nF++, ik);
                    ph = sm ? PH_START_BWD : PH_FWD_TAIL;
                }
            } else {                                    // PH_START_BWD: forward sweep finished (smem.c:45-50)
                if (nF == 0) {                          // undefined in the reference: no SMEM here
                    x = i < len ? i : len;
                    if (x >= len) { A.rec_cnt[r] = (uint32_t)nmem; ph = PH_FETCH; } else ph = PH_BEGIN;
                    continue;
                }
                ik = ld_cand(F + (nF - 1));         // the longest match is the last push = first candidate
                ret = (int)ik.info;
                nprev = nF; first_pass = 1; i = x - 1; j = 0; ncurr = 0; call_base = nmem;
                ph = PH_BWD;
            }
        }
        const bool active = ph != PH_DONE;
        if (!warp_any(active)) return;
        if (!PAIR && !active) continue;

        // ---- the one extension of this trip (converged across the warp).  The query base it will be
        // consumed with and the next candidate of the backward pass are requested first, so that their
        // latency overlaps the two block loads instead of following them.
        const int back = (ph == PH_BWD);
        const int qc = (!active || ph == PH_FWD_TAIL || i < 0) ? 0 : (int)ld_u8(q + i);
        const bool have_nxt = active && back && j + 1 < nprev;       // entry j+1 of the list being read is final (writes go to <= j)
        Cand nxt = ik;
        if (have_nxt) nxt = ld_cand(first_pass ? F + (nF - 2 - j) : W + (j + 1));
        Ext6T<U> e;
        if (PAIR) {
            extend6_conv<U>(A.ix, active, back ? ik.x1 : ik.x0, back ? ik.x0 : ik.x1, ik.x2, e);
            if (!active) continue;
        } else extend6(A.ix, back ? ik.x1 : ik.x0, back ? ik.x0 : ik.x1, ik.x2, e);

        // ---- consume it
        // x[0]/x[1] of ok[c]: the far side is x[1] for a forward, x[0] for a backward extension
#define FMG_OK(c, dst) do { const U nr_ = pick6(e.near, c), fr_ = far_of(A.ix, e, c); \
                            (dst).x0 = back ? fr_ : nr_; (dst).x1 = back ? nr_ : fr_; (dst).x2 = pick6(e.size, c); } while (0)
        if (ph == PH_FWD) {                             // smem.c:22-34
            const int c = comp6(qc);
            const U sc = pick6(e.size, c);
            if (sc != ik.x2) {
                if (ik.x2 != e.size[0]) st_cand(F + nF++, ik);
                if (!sm && e.size[0]) {
                    Cand s0; s0.x0 = e.near[0]; s0.x1 = far_of(A.ix, e, 0); s0.x2 = e.size[0]; s0.info = (U)i;
                    st_cand(F + nF++, s0);
                }
            }
            const bool stop = sm ? sc < 2 : sc == 0;
            if (stop) ph = PH_START_BWD;
            else {
                FMG_OK(c, ik); ik.info = (U)(i + 1);
                if (++i == len) {                       // reached the end of the read (smem.c:35-36)
                    st_cand(F + nF++, ik);
                    ph = sm ? PH_START_BWD : PH_FWD_TAIL;
                }
            }
        } else if (ph == PH_FWD_TAIL) {                 // smem.c:37-43
            if (e.size[0]) {
                Cand s0; s0.x0 = e.near[0]; s0.x1 = far_of(A.ix, e, 0); s0.x2 = e.size[0]; s0.info = (U)len;
                st_cand(F + nF++, s0);
            }
            ph = PH_START_BWD;
        } else {                                        // backward sweep, smem.c:51-75; ik is the candidate p
            const int c = qc;
            const U sc = pick6(e.size, c);
            const bool fl = e.size[0] != 0 && ik.x1 < A.ix.n_seq;
            const bool cont = sm ? sc > 1 : sc != 0;
            if ((!cont || fl || i == -1) && (ncurr == 0 || fl) &&
                (fl || nmem == call_base || i + 1 < last_start)) {
                Intv m; m.x0 = ik.x0; m.x1 = ik.x1; m.x2 = ik.x2;
                m.info = (uint64_t)ik.info | (uint64_t)(e.size[0] != 0) << 63 | (uint64_t)(i + 1) << 32;
                if (nmem < A.out_cap) st_intv(out + 2 * nmem, m);
                ++nmem; last_start = i + 1;
            }
            if (cont && (ik.x1 < A.ix.n_seq || ncurr == 0 || sc != last_x2)) {
                Cand n; FMG_OK(c, n); n.info = ik.info;
                st_cand(W + ncurr++, n);
                last_x2 = sc;
            }
            if (++j < nprev) ik = nxt;
            else {
                if (ncurr != 0 && i != -1) {            // next backward position over the survivors
                    nprev = ncurr; ncurr = 0; j = 0; --i; first_pass = 0;
                    ik = ld_cand(W);
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
        }
#undef FMG_OK
    }
}

} // namespace fmg
