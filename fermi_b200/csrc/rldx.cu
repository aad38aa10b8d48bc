// Rank on the .fmd stream ITSELF (north_star's design): the run-length / Elias-delta blocks of rld.c stay as they are in HBM
// (0.8-2.2 bits per symbol for read sets instead of the 4 bits of the occ-block layout), a warp serves a query:
//   locate   the block holding position p: a coarse table (one entry per 256 symbols) + 32 lanes comparing 32 consecutive block
//            starts, one ballot -- instead of rld_locate_blk's frame row + linear scan of ~7 block headers (rld.c:352-392);
//   stage    the 64-byte block and its 64-byte directory line (start coordinate + the six cumulative counts: the "128 bytes per
//            lookup" of SURVEY.md 8d) travel to shared memory as bulk asynchronous copies (TMA unit) completing on an mbarrier;
//   decode   every lane takes 12 of the up-to-384 payload bit offsets and computes the length a code would have if it started
//            there (rld_dec0, rld.h:77-94); the offsets that ARE code starts are the chain start -> start + length -> ...,
//            marked by pointer doubling in 7 rounds (a block holds at most 96 codes); the marked codes, in offset order, are
//            the runs: a warp prefix sum of their lengths places every run, and six warp reductions add up, per symbol, the
//            part of each run that lies before the target position.
// Memory: the stream as it is on disk plus one 64-byte directory line per 64-byte block, i.e. about twice the .fmd file (1.6-4.4
// bits per symbol for read sets, against 4 for the occ blocks; random text, 4.8 bits per symbol on disk, is larger this way).
// The occ-block layout (fmd_device.cuh) remains the fast path: a rank there is ~100 instructions of ONE thread, here ~700 of a
// whole warp.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <vector>
#include <algorithm>
#include "fmd_host.hpp"
#include "dev_pool.hpp"
#include "fmg_internal.hpp"
#include "../../include/fermi_b200.h"

using namespace fmg;
extern std::atomic<uint64_t> g_launches;

#define RX_TRY(call, fail)                                                                            \
    do {                                                                                              \
        cudaError_t err__ = (call);                                                                   \
        if (err__ != cudaSuccess) {                                                                   \
            if (fmg_verbose >= 1)                                                                     \
                std::fprintf(stderr, "[E::%s] %s failed: %s\n", __func__, #call, cudaGetErrorString(err__)); \
            fail;                                                                                     \
        }                                                                                             \
    } while (0)

struct RldView {
    const uint64_t *words;        // the bit stream: n_blk blocks of 8 words (+ the closing header-only block)
    const uint64_t *dir;          // n_blk + 1 lines of 8 words: [0] first BWT coordinate of the block, [1..6] symbols $..N before it
    const uint32_t *coarse;       // coarse[q] = block holding position q << kCoarseShift
    uint64_t n_blk, n_sym;
    uint64_t C[8];
};
constexpr int kCoarseShift = 8;
constexpr int kRxWarps = 4;       // warps per thread block
constexpr uint32_t kEnd = 511;    // chain terminator (bit offsets of a block are < 512)

struct fmg_rldx_s {
    int device = 0, n_sm = 0;
    uint64_t *d_words = nullptr, *d_dir = nullptr;
    uint32_t *d_coarse = nullptr;
    uint64_t bytes = 0;
    RldView view;
};

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct WarpSmem {
    alignas(128) uint64_t blk[2][8];     // staged payload blocks (k side, l side)
    alignas(64) uint64_t dir[2][8];      // their directory lines
    uint16_t jmp[2][512];                // pointer-doubling tables (ping-pong)
    uint32_t mark[16];                   // code starts, one bit per bit offset
    alignas(8) uint64_t bar;             // mbarrier of the bulk copies
};

// 64 bits of the block starting at bit offset `bit` (MSB first), nothing read past bit `end` (rld.h:82)
__device__ __forceinline__ uint64_t window(const uint64_t *w, uint32_t bit, uint32_t end) {
    const uint32_t i = bit >> 6, s = bit & 63;
    uint64_t x = w[i];
    if (s) x = (x << s) | (i + 1 < 8 ? w[i + 1] >> (64 - s) : 0ull);
    if (end - bit < 64) x &= ~0ull << (64 - (end - bit));
    return x;
}
// one code (rld_dec0, rld.h:77-94): bits used (0 = no code here: end of block), run length, symbol
__device__ __forceinline__ uint32_t decode_at(uint64_t x, uint64_t *len, int *sym) {
    if (x == 0) return 0;
    if (x >> 63) { *len = 1; *sym = (int)((x >> 60) & 7); return *sym > 6 ? 0u : 4u; }
    const int z = __clzll((long long)x);
    if (z > 5) return 0;
    const int g = 2 * z + 1, y = (int)(x >> (64 - g)) - 1;
    *len = ((x << g) >> (64 - y)) | (1ull << y);
    *sym = (int)((x << (g + y)) >> 61);
    return *sym > 6 ? 0u : (uint32_t)(g + y + 3);
}

// Counts of the six symbols among the first t0 (and t1) symbols of the staged block w; every lane returns the warp totals.
__device__ void warp_block_counts(WarpSmem &S, const uint64_t *w, bool last_of_chunk, uint64_t t0, uint64_t t1, uint32_t c0[6], uint32_t c1[6]) {
    const int lane = threadIdx.x & 31;
    const uint32_t hb = (((uint32_t)w[0] >> 31) ? kHeaderWords32 : kHeaderWords16) * 64, tb = (last_of_chunk ? 7u : 8u) * 64;
    // 1. the length a code would have at each of my 12 offsets
    const uint32_t o0 = hb + 12u * lane;
    if (lane < 16) S.mark[lane] = 0;
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        const uint32_t o = o0 + j;
        if (o >= 512) continue;
        uint32_t nx = kEnd;
        if (o < tb) {
            uint64_t len; int sym;
            const uint32_t used = decode_at(window(w, o, tb), &len, &sym);
            if (used && o + used <= tb) nx = o + used < tb ? o + used : kEnd;
            if (!used) nx = kEnd;
        }
        S.jmp[0][o] = (uint16_t)nx;
    }
    if (lane == 0) { S.jmp[0][kEnd] = (uint16_t)kEnd; S.jmp[1][kEnd] = (uint16_t)kEnd; }
    __syncwarp();
    if (lane == 0) S.mark[hb >> 5] = 1u << (hb & 31);
    __syncwarp();
    // 2. the offsets reachable from the first one: pointer doubling, the marked prefix of the chain doubles every round
    int cur = 0;
    for (int r = 0; r < 7; ++r) {
        uint16_t nj[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            const uint32_t o = o0 + j;
            nj[j] = (uint16_t)kEnd;
            if (o >= 511) continue;
            const uint32_t t = S.jmp[cur][o];
            if ((S.mark[o >> 5] >> (o & 31) & 1u) && t != kEnd) atomicOr(&S.mark[t >> 5], 1u << (t & 31));
            nj[j] = S.jmp[cur][t];
        }
#pragma unroll
        for (int j = 0; j < 12; ++j) { const uint32_t o = o0 + j; if (o < 511) S.jmp[cur ^ 1][o] = nj[j]; }
        cur ^= 1;
        __syncwarp();
    }
    // 3. my runs (codes that start at a marked offset of mine), in offset order
    uint64_t len[3] = {0, 0, 0};
    int sym[3] = {0, 0, 0}, nr = 0;
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        const uint32_t o = o0 + j;
        if (o >= tb || !(S.mark[o >> 5] >> (o & 31) & 1u)) continue;
        uint64_t l; int s;
        if (decode_at(window(w, o, tb), &l, &s) && nr < 3) { len[nr] = l; sym[nr] = s; ++nr; }
    }
    const uint64_t mine = len[0] + len[1] + len[2];
    uint64_t incl = mine;
    for (int d = 1; d < 32; d <<= 1) { const uint64_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    uint64_t cum = incl - mine;                       // symbols of the block before my first run
    uint32_t p0[6] = {0, 0, 0, 0, 0, 0}, p1[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        if (q < nr) {
            const uint64_t a0 = cum < t0 ? (t0 - cum < len[q] ? t0 - cum : len[q]) : 0, a1 = cum < t1 ? (t1 - cum < len[q] ? t1 - cum : len[q]) : 0;
#pragma unroll
            for (int c = 0; c < 6; ++c) if (sym[q] == c) { p0[c] += (uint32_t)a0; p1[c] += (uint32_t)a1; }
            cum += len[q];
        }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) { c0[c] = __reduce_add_sync(0xffffffffu, p0[c]); c1[c] = __reduce_add_sync(0xffffffffu, p1[c]); }
    __syncwarp();
}

// block holding position p (p <= n_sym; p == n_sym belongs to the last block)
__device__ __forceinline__ uint64_t warp_locate(const RldView &ix, uint64_t p) {
    const int lane = threadIdx.x & 31;
    const uint64_t b0 = ix.coarse[p >> kCoarseShift], b = b0 + lane;
    const bool le = b < ix.n_blk && ix.dir[b * 8] <= p;
    const unsigned m = __ballot_sync(0xffffffffu, le);
    return b0 + (uint64_t)__popc(m) - 1;
}

// one warp per query: counts of the six symbols in BWT[0, pk) and BWT[0, pl)
__global__ void __launch_bounds__(32 * kRxWarps) k_rldx_rank2(RldView ix, int64_t n, const uint64_t *__restrict__ pk_, const uint64_t *__restrict__ pl_,
                                                            uint64_t *__restrict__ ok, uint64_t *__restrict__ ol) {
    __shared__ WarpSmem smem[kRxWarps];
    WarpSmem &S = smem[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = (int64_t)gridDim.x * kRxWarps;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0;
    for (int64_t q = (int64_t)blockIdx.x * kRxWarps + (threadIdx.x >> 5); q < n; q += n_warps) {
        const uint64_t pk = pk_[q], pl = pl_[q];
        const uint64_t bk = warp_locate(ix, pk), bl = warp_locate(ix, pl);
        const bool two = bl != bk;
        // stage: payload block(s) + directory line(s), 64 bytes each, as bulk copies on one mbarrier
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const uint32_t bytes = two ? 256u : 128u, bar = smem_u32(&S.bar);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];" ::"r"(smem_u32(S.blk[0])), "l"(ix.words + bk * 8), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];" ::"r"(smem_u32(S.dir[0])), "l"(ix.dir + bk * 8), "r"(bar) : "memory");
            if (two) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];" ::"r"(smem_u32(S.blk[1])), "l"(ix.words + bl * 8), "r"(bar) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];" ::"r"(smem_u32(S.dir[1])), "l"(ix.dir + bl * 8), "r"(bar) : "memory");
            }
        }
        {
            const uint32_t bar = smem_u32(&S.bar);
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "W_%=:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra D_%=;\n\t"
                "bra W_%=;\n\t"
                "D_%=:\n\t}" ::"r"(bar), "r"(phase) : "memory");
            phase ^= 1;
        }
        uint32_t ck[6], cl[6], dummy[6];
        const bool lastk = ((bk + 1) * 8 & (kChunkWords - 1)) == 0, lastl = ((bl + 1) * 8 & (kChunkWords - 1)) == 0;
        const uint64_t sk = S.dir[0][0];
        if (!two) warp_block_counts(S, S.blk[0], lastk, pk - sk, pl - sk, ck, cl);
        else {
            warp_block_counts(S, S.blk[0], lastk, pk - sk, 0, ck, dummy);
            warp_block_counts(S, S.blk[1], lastl, pl - S.dir[1][0], 0, cl, dummy);
        }
        if (lane < 6) {
            ok[6 * q + lane] = S.dir[0][1 + lane] + ck[lane];
            ol[6 * q + lane] = S.dir[two ? 1 : 0][1 + lane] + cl[lane];
        }
        __syncwarp();
    }
}

}  // namespace

extern "C" {

fmg_rldx_t *fmg_rldx_upload(const fmg_fmd_t *e, int device) {
    if (!e) return nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] no CUDA device available; libfermi_b200 has no CPU path\n", __func__);
        return nullptr;
    }
    RX_TRY(cudaSetDevice(device), return nullptr);
    const FmdImage &img = e->img;
    const uint64_t n_blk = img.n_blocks(), n_sym = img.mcnt[0];
    // dense directory from the block headers: header b holds the symbol counts of block b-1 (rld.c:111-134), so their running
    // sums are the counts before every block and the total of a header's counts the length of the block before it
    std::vector<uint64_t> dir((n_blk + 1) * 8, 0);
    const uint64_t *w = img.words.data();
    uint64_t cum[7] = {0, 0, 0, 0, 0, 0, 0};
    for (uint64_t b = 0; b <= n_blk; ++b) {
        if (b > 0) {                                               // header of block b = counts of block b-1: [0] total, [1..6] symbols
            const uint64_t *h = w + b * kBlockWords;
            if (header_is32(h[0])) {
                const uint32_t *u = reinterpret_cast<const uint32_t *>(h);
                cum[0] += u[0] & 0x7fffffffu;
                for (int c = 1; c < 7; ++c) cum[c] += u[c];
            } else {
                const uint16_t *u = reinterpret_cast<const uint16_t *>(h);
                for (int c = 0; c < 7; ++c) cum[c] += u[c];
            }
        }
        dir[b * 8] = cum[0];
        for (int c = 0; c < 6; ++c) dir[b * 8 + 1 + c] = cum[1 + c];
    }
    if (dir[n_blk * 8] != n_sym) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] block headers add up to %llu symbols, the index holds %llu\n", __func__,
                                           (unsigned long long)dir[n_blk * 8], (unsigned long long)n_sym);
        return nullptr;
    }
    const uint64_t n_coarse = (n_sym >> kCoarseShift) + 2;
    std::vector<uint32_t> coarse(n_coarse);
    {
        uint64_t b = 0;
        for (uint64_t q = 0; q < n_coarse; ++q) {
            const uint64_t p = q << kCoarseShift;
            while (b + 1 < n_blk && dir[(b + 1) * 8] <= p) ++b;
            coarse[q] = (uint32_t)b;
        }
    }
    fmg_rldx_s *x = new fmg_rldx_s;
    x->device = device;
    cudaDeviceGetAttribute(&x->n_sm, cudaDevAttrMultiProcessorCount, device);
    const uint64_t n_words = (n_blk + 1) * kBlockWords;
    bool ok = false;
    do {
        RX_TRY(cudaMalloc(&x->d_words, (n_words + 8) * 8), break);
        RX_TRY(cudaMemset(x->d_words, 0, (n_words + 8) * 8), break);
        RX_TRY(cudaMemcpy(x->d_words, w, std::min<uint64_t>(n_words, img.words.size()) * 8, cudaMemcpyHostToDevice), break);
        RX_TRY(cudaMalloc(&x->d_dir, dir.size() * 8), break);
        RX_TRY(cudaMemcpy(x->d_dir, dir.data(), dir.size() * 8, cudaMemcpyHostToDevice), break);
        RX_TRY(cudaMalloc(&x->d_coarse, n_coarse * 4), break);
        RX_TRY(cudaMemcpy(x->d_coarse, coarse.data(), n_coarse * 4, cudaMemcpyHostToDevice), break);
        ok = true;
    } while (0);
    if (!ok) { cudaFree(x->d_words); cudaFree(x->d_dir); cudaFree(x->d_coarse); delete x; return nullptr; }
    x->bytes = (n_words + 8) * 8 + dir.size() * 8 + n_coarse * 4;
    x->view.words = x->d_words; x->view.dir = x->d_dir; x->view.coarse = x->d_coarse; x->view.n_blk = n_blk; x->view.n_sym = n_sym;
    for (int c = 0; c < 8; ++c) x->view.C[c] = img.cnt[c];
    if (fmg_verbose >= 3)
        std::fprintf(stderr, "[M::%s] %llu symbols in %llu RLD blocks: %.1f MB on device %d (stream %.1f MB + directory %.1f MB)\n", __func__, (unsigned long long)n_sym,
                     (unsigned long long)n_blk, x->bytes / 1e6, device, n_words * 8 / 1e6, (dir.size() * 8 + n_coarse * 4) / 1e6);
    return x;
}

void fmg_rldx_free(fmg_rldx_t *x) {
    if (!x) return;
    cudaSetDevice(x->device);
    cudaFree(x->d_words); cudaFree(x->d_dir); cudaFree(x->d_coarse);
    delete x;
}

uint64_t fmg_rldx_bytes(const fmg_rldx_t *x) { return x ? x->bytes : 0; }

int fmg_rldx_rank2a_batch(const fmg_rldx_t *x, int64_t n, const uint64_t *k, const uint64_t *l, uint64_t *ok, uint64_t *ol) {
    if (!x || !k || !l || !ok || !ol) return -1;
    RX_TRY(cudaSetDevice(x->device), return -1);
    if (n <= 0) return 0;
    std::vector<uint64_t> pk(n), pl(n);
    for (int64_t i = 0; i < n; ++i) {                              // rld_rank2a(k, l) counts BWT[0..k] and BWT[0..l]; k == -1: nothing
        pk[i] = k[i] + 1; pl[i] = l[i] + 1;
        if (pk[i] > x->view.n_sym || pl[i] > x->view.n_sym) {
            if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] position beyond the end of the BWT\n", __func__);
            return -1;
        }
    }
    uint64_t *dk = nullptr, *dl = nullptr, *dok = nullptr, *dol = nullptr;
    int rc = -1;
    do {
        RX_TRY(cudaMalloc(&dk, n * 8), break); RX_TRY(cudaMalloc(&dl, n * 8), break);
        RX_TRY(cudaMalloc(&dok, n * 48), break); RX_TRY(cudaMalloc(&dol, n * 48), break);
        RX_TRY(cudaMemcpy(dk, pk.data(), n * 8, cudaMemcpyHostToDevice), break);
        RX_TRY(cudaMemcpy(dl, pl.data(), n * 8, cudaMemcpyHostToDevice), break);
        const int grid = (int)std::min<int64_t>((n + kRxWarps - 1) / kRxWarps, (int64_t)x->n_sm * 8);
        k_rldx_rank2<<<grid, 32 * kRxWarps>>>(x->view, n, dk, dl, dok, dol);
        ++g_launches;
        RX_TRY(cudaGetLastError(), break);
        RX_TRY(cudaMemcpy(ok, dok, n * 48, cudaMemcpyDeviceToHost), break);
        RX_TRY(cudaMemcpy(ol, dol, n * 48, cudaMemcpyDeviceToHost), break);
        rc = 0;
    } while (0);
    cudaFree(dk); cudaFree(dl); cudaFree(dok); cudaFree(dol);
    return rc;
}

// fm6_extend (exact.c:72-88) from the two ranks
int fmg_rldx_extend_batch(const fmg_rldx_t *x, int64_t n, const fmg_intv_t *ik, const uint8_t *is_back, fmg_intv_t *ok6) {
    if (!x || !ik || !is_back || !ok6) return -1;
    if (n <= 0) return 0;
    std::vector<uint64_t> k(n), l(n), tk(6 * n), tl(6 * n);
    for (int64_t i = 0; i < n; ++i) { const int b = is_back[i] != 0; k[i] = ik[i].x[!b] - 1; l[i] = ik[i].x[!b] - 1 + ik[i].x[2]; }
    if (fmg_rldx_rank2a_batch(x, n, k.data(), l.data(), tk.data(), tl.data()) != 0) return -1;
    for (int64_t i = 0; i < n; ++i) {
        const int b = is_back[i] != 0;
        fmg_intv_t *o = ok6 + 6 * i;
        uint64_t sz[6];
        for (int c = 0; c < 6; ++c) { o[c].x[!b] = x->view.C[c] + tk[6 * i + c]; o[c].x[2] = sz[c] = tl[6 * i + c] - tk[6 * i + c]; o[c].info = 0; }
        o[0].x[b] = ik[i].x[b];
        o[4].x[b] = o[0].x[b] + sz[0]; o[3].x[b] = o[4].x[b] + sz[4]; o[2].x[b] = o[3].x[b] + sz[3]; o[1].x[b] = o[2].x[b] + sz[2]; o[5].x[b] = o[1].x[b] + sz[1];
    }
    return 0;
}

} // extern "C"
