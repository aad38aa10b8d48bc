// BCR (Bauer-Cox-Rosone) BWT construction on the GPU: the builder behind `fermi ropebwt -a bcr`
// (bcr.c:358-521, ropebwt.c:47-158), for collections of any total size that fits HBM (no 2^32 limit).
//
// Same cycle structure as the reference -- cycle `pos` inserts, for every sequence still active, the base
// at distance `pos` from its end as the BWT symbol of the suffix that was completed in cycle pos-1, and the
// sentinel when the sequence is exhausted (bcr.c:417-449) -- but laid out for a B200:
//   * the partial BWT is ONE plain byte array in HBM (15 GB for 50 M x 150 bp; two buffers), not six
//     run-length-encoded ropes: a cycle is a streaming merge of the old array with the sorted inserts;
//   * the insert position of a suffix in the next cycle is the LF mapping C[a] + rank_a(B, F): ranks come
//     from a per-tile symbol histogram (k_bcr_merge), an exclusive scan over the tiles and the in-tile
//     prefix, all computed inside the merge pass -- no separate rank structure is maintained;
//   * LF is monotone, so the order of the sequences in the next cycle is the STABLE partition of the current
//     order by the inserted symbol: one 3-bit radix pass replaces the reference's per-bucket radix sort of
//     64-bit positions (rs_sort, bcr.c:212-249,426).
// The result is the BWT of  r0 $ r1 $ ...  with sentinels ordered by sequence number: byte-identical (after
// RLD encoding) to what `fermi build` and `fermi ropebwt` + `recode` produce.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <vector>
#include <memory>
#include <chrono>
#include <thread>
#include "fmd_host.hpp"
#include "dev_pool.hpp"
#include "bcr_tile.cuh"
#include "../../include/fermi_b200.h"

int fmg_rld_encode_device(const uint8_t *d_bwt, uint64_t n, fmg::FmdImage *out);     // rld_enc.cu
struct fmg_fmd_s { fmg::FmdImage img; };

extern std::atomic<uint64_t> g_launches;

namespace {

#define BCR_TRY(call)                                                                                   \
    do {                                                                                                \
        cudaError_t err__ = (call);                                                                     \
        if (err__ != cudaSuccess) {                                                                     \
            if (fmg_verbose >= 1)                                                                       \
                std::fprintf(stderr, "[E::fmg_bcr_build] %s failed: %s\n", #call, cudaGetErrorString(err__)); \
            return -1;                                                                                  \
        }                                                                                               \
    } while (0)

// the merge pass: kPer output symbols per thread (16 or 32: one or two 16-byte vectors), a tile = kPer x block size symbols

struct Vec4 { uint64_t v[4]; };      // occurrences of A,C,G,T
struct Vec4Add { __host__ __device__ Vec4 operator()(const Vec4 &a, const Vec4 &b) const { Vec4 r; for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] + b.v[i]; return r; } };
struct Item { uint64_t f; uint32_t id; uint32_t pad; };      // (position of the suffix in the current BWT, sequence)

// Device buffers come from the pool the overlap pass uses (dev_pool.hpp): a build allocates and frees tens of gigabytes, and
// cudaMalloc / cudaFree of those cost as much as several cycles.  fmg_release_cache hands the pool back to the driver.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int dev = 0;
    ~DevBuf() { release(); }
    void release() { if (p) fmg::g_pool.put(p, cap, dev); p = nullptr; cap = 0; }
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        release();
        cudaGetDevice(&dev);
        return fmg::g_pool.get(bytes ? bytes : 1, dev, &p, &cap);
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

// Pageable host memory -> device through pinned staging buffers filled by several threads: a plain cudaMemcpy of pageable memory is
// one thread copying into the driver's staging buffer (~10 GB/s); the reads of a build are gigabytes.
struct Stager {
    static constexpr int kThreads = 4, kBufs = 2;
    static constexpr size_t kChunk = 32u << 20;
    uint8_t *pin = nullptr;                  // kept until the process exits
    std::mutex lock;
    cudaError_t upload(void *dst, const void *src, size_t bytes) {
        if (bytes < 4 * kChunk) return cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
        std::lock_guard<std::mutex> guard(lock);
        if (!pin) { const cudaError_t e = cudaHostAlloc((void **)&pin, kThreads * kBufs * kChunk, cudaHostAllocDefault); if (e != cudaSuccess) { pin = nullptr; cudaGetLastError(); return cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice); } }
        int dev = 0;
        cudaGetDevice(&dev);
        const size_t n_chunks = (bytes + kChunk - 1) / kChunk;
        std::atomic<int> fail{0};
        auto work = [&](int t) {
            cudaSetDevice(dev);
            cudaStream_t st; cudaEvent_t ev[kBufs];
            if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { fail = 1; return; }
            for (auto &e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            int k = 0;
            for (size_t c = (size_t)t; c < n_chunks && !fail; c += kThreads, ++k) {
                uint8_t *buf = pin + ((size_t)t * kBufs + (k % kBufs)) * kChunk;
                const size_t o = c * kChunk, len = std::min(kChunk, bytes - o);
                if (k >= kBufs) cudaEventSynchronize(ev[k % kBufs]);                  // the copy that last used this buffer
                std::memcpy(buf, static_cast<const uint8_t *>(src) + o, len);
                if (cudaMemcpyAsync(static_cast<uint8_t *>(dst) + o, buf, len, cudaMemcpyHostToDevice, st) != cudaSuccess) fail = 1;
                cudaEventRecord(ev[k % kBufs], st);
            }
            if (cudaStreamSynchronize(st) != cudaSuccess) fail = 1;
            for (auto &e : ev) cudaEventDestroy(e);
            cudaStreamDestroy(st);
        };
        std::vector<std::thread> th;
        for (int t = 0; t < kThreads; ++t) th.emplace_back(work, t);
        for (auto &x : th) x.join();
        return fail ? cudaErrorUnknown : cudaSuccess;
    }
};
Stager g_stager;

// symbol each active sequence inserts in cycle `pos`: its base at distance pos from the end, 0 when exhausted (bcr.c:430)
__global__ void k_bcr_symbols(const Item *__restrict__ item, uint64_t n_act, const uint8_t *__restrict__ seq, const uint64_t *__restrict__ off,
                              int pos, uint8_t *__restrict__ sym) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_act) return;
    const uint32_t id = item[k].id;
    const uint64_t b = off[id], e = off[id + 1];
    sym[k] = (uint64_t)pos < e - b ? seq[e - 1 - pos] : 0;
}

// The same from the sequences transposed by cycle (what bcr_append keeps, bcr.c:358-376): col[pos][id] = the base of sequence id
// at distance pos from its end, 0 past its start.  A cycle then gathers single bytes from ONE row of n_seq bytes (L2-resident up to
// ~10^8 sequences) instead of an offset pair and a byte from the whole read set.
__global__ void k_bcr_transpose(const uint8_t *__restrict__ seq, const uint64_t *__restrict__ off, uint64_t n_seq, int max_len, uint8_t *__restrict__ col) {
    const uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_seq) return;
    const uint64_t b = off[id], e = off[id + 1];
    for (int pos = 0; pos < max_len; ++pos) col[(uint64_t)pos * n_seq + id] = (uint64_t)pos < e - b ? seq[e - 1 - pos] : 0;
}
__global__ void k_bcr_symbols_col(const Item *__restrict__ item, uint64_t n_act, const uint8_t *__restrict__ row, uint8_t *__restrict__ sym) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_act) sym[k] = row[item[k].id];
}

__global__ void k_bcr_iota(Item *item, uint64_t n) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { item[k].f = k; item[k].id = (uint32_t)k; item[k].pad = 0; }
}

// tile_lo[t] = first insert whose position falls into output tile t or later (the inserts are sorted by position): every
// insert fills the entries between its predecessor's tile and its own -- one coalesced pass instead of a binary search
// over the whole insert list at the start of every tile
__global__ void k_bcr_bounds(const Item *__restrict__ item, uint64_t n_act, uint64_t n_tiles, uint64_t *__restrict__ tile_lo, uint32_t kTile) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_act) return;
    const uint64_t t1 = item[k].f / kTile;
    const uint64_t t0 = k ? item[k - 1].f / kTile + 1 : 0;
    for (uint64_t t = t0; t <= t1; ++t) tile_lo[t] = k;
    if (k == n_act - 1) for (uint64_t t = t1 + 1; t <= n_tiles; ++t) tile_lo[t] = n_act;
}

// One tile of the new BWT: old symbols and inserts merged, the tile's A/C/G/T histogram, and for every insert
// the number of equal symbols before it inside the tile.  16 output symbols per thread: the old symbols a tile keeps are
// one contiguous piece of the old array, staged in shared memory with aligned 16-byte loads; a thread whose 16 positions
// hold no insert (most of them: one insert per pos symbols in cycle pos) copies 16 staged bytes with word operations, and
// every thread writes one 16-byte vector.
// ---- bulk asynchronous copies (TMA unit, 1-D) and their mbarriers: the old symbols of the next tiles travel to shared memory while
// the block merges the current one
constexpr int kStages = 4;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar)) : "memory");
}

// PERSISTENT: blocks walk over the tiles (blockIdx.x, + gridDim.x, ...); thread 0 keeps kStages bulk copies of old symbols in
// flight per block, each completing on its own mbarrier.
template <int kThreads, int kMinBlocks, int kPer>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_bcr_merge(const uint8_t *__restrict__ old_bwt, uint64_t m_new, const Item *__restrict__ item,
                                                        const uint8_t *__restrict__ sym, const uint64_t *__restrict__ tile_lo, uint8_t *__restrict__ new_bwt,
                                                        Vec4 *__restrict__ tile_hist, uint32_t *__restrict__ rank_in_tile, uint64_t n_tiles) {
    constexpr int kTile = kThreads * kPer, kWords = kPer / 4, kVecs = kPer / 16;
    static_assert(kPer == 16 || kPer == 32, "the bit planes below hold one bit per word in every byte");
    __shared__ __align__(16) uint8_t s_sym[kTile];
    __shared__ __align__(128) uint8_t s_srcs[kStages][kTile + 64];
    __shared__ __align__(16) uint8_t s_flag[kTile];
    __shared__ uint32_t s_warp_ins[kThreads / 32];
    __shared__ uint64_t s_warp_cnt[kThreads / 32];
    __shared__ __align__(8) uint64_t s_bar[kStages];
    __shared__ uint32_t s_spread[16];              // byte-permute selector that spreads consecutive bytes over the non-insert positions of a word
    const int tid = threadIdx.x;
    // source range of a tile: old symbols [src_lo, src_lo + n_src), fetched from the 16-byte boundary below src_lo
    auto issue = [&](uint64_t tile, int stage) {
        const uint64_t tq0 = tile * kTile, lo = tile_lo[tile], hi = tile_lo[tile + 1];
        const uint64_t len = tq0 + kTile <= m_new ? (uint64_t)kTile : m_new - tq0;
        const uint64_t ns = len - (hi - lo), sl = tq0 - lo, al = sl & ~15ull;
        const uint32_t bytes = ns ? (uint32_t)((ns + (sl - al) + 15) & ~15ull) : 0u;
        if (bytes) { mbar_expect_tx(&s_bar[stage], bytes); bulk_g2s(s_srcs[stage], old_bwt + al, bytes, &s_bar[stage]); }
        else mbar_arrive(&s_bar[stage]);
    };
    if (tid < 16) s_spread[tid] = fmg::bcr_spread_selector((uint32_t)tid);
    if (tid == 0) {
        for (int q = 0; q < kStages; ++q) mbar_init(&s_bar[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int q = 0; q < kStages; ++q) { const uint64_t t = (uint64_t)blockIdx.x + (uint64_t)q * gridDim.x; if (t < n_tiles) issue(t, q); }
    // Software pipeline of the metadata: the insert range of a tile is requested two tiles ahead and the thread's first insert of
    // it (position, symbol) one tile ahead, so that no tile waits for a chain of dependent global loads
    const uint64_t stride = gridDim.x;
    auto range_of = [&](uint64_t t, uint64_t &lo, uint64_t &hi) { if (t < n_tiles) { lo = tile_lo[t]; hi = tile_lo[t + 1]; } else lo = hi = 0; };
    uint64_t k_lo, k_hi, n_lo, n_hi, nn_lo, nn_hi;
    range_of(blockIdx.x, k_lo, k_hi);
    range_of(blockIdx.x + stride, n_lo, n_hi);
    uint64_t my_f = 0, nx_f = 0;
    uint8_t my_s = 0, nx_s = 0;
    if (k_lo + tid < k_hi) { my_f = item[k_lo + tid].f; my_s = sym[k_lo + tid]; }
    for (uint64_t it = 0;; ++it) {
    const uint64_t tile = (uint64_t)blockIdx.x + it * stride;
    if (tile >= n_tiles) break;
    const int stage = (int)(it % kStages);
    const uint8_t *s_src = s_srcs[stage];
    const uint64_t q0 = tile * kTile;
    range_of(tile + 2 * stride, nn_lo, nn_hi);                          // requested now, used two tiles on
    if (n_lo + tid < n_hi) { nx_f = item[n_lo + tid].f; nx_s = sym[n_lo + tid]; }      // the next tile's insert of this thread
#pragma unroll
    for (int v = 0; v < kVecs; ++v) reinterpret_cast<uint4 *>(s_flag)[tid * kVecs + v] = make_uint4(0, 0, 0, 0);
    // the old symbols this tile keeps
    const uint64_t src_lo = q0 - k_lo;            // old symbols before this tile
    const uint64_t a0 = src_lo & ~15ull;
    const int shift = (int)(src_lo - a0);
    __syncthreads();
    if (k_lo + tid < k_hi) { const int j = (int)(my_f - q0); s_flag[j] = 1; s_sym[j] = my_s; }
    for (uint64_t k = k_lo + tid + kThreads; k < k_hi; k += kThreads) {  // more than 256 inserts in the tile: the first cycles only
        const int j = (int)(item[k].f - q0);
        s_flag[j] = 1; s_sym[j] = sym[k];
    }
    mbar_wait(&s_bar[stage], (uint32_t)((it / kStages) & 1));        // the bulk copy of this tile's old symbols has landed
    __syncthreads();
    // inserts before each thread's kPer positions
    const int j0 = tid * kPer;
    uint32_t flw[kWords], syw[kWords];
#pragma unroll
    for (int v = 0; v < kVecs; ++v) {
        const uint4 fl = reinterpret_cast<const uint4 *>(s_flag)[tid * kVecs + v], sy = reinterpret_cast<const uint4 *>(s_sym)[tid * kVecs + v];
        flw[4 * v] = fl.x; flw[4 * v + 1] = fl.y; flw[4 * v + 2] = fl.z; flw[4 * v + 3] = fl.w;
        syw[4 * v] = sy.x; syw[4 * v + 1] = sy.y; syw[4 * v + 2] = sy.z; syw[4 * v + 3] = sy.w;
    }
    uint32_t fl_sum = 0;
#pragma unroll
    for (int w = 0; w < kWords; ++w) fl_sum += flw[w];
    const uint32_t my_ins = fmg::bcr_flag_count(fl_sum);                 // flag bytes are 0/1: at most kWords per byte lane
    uint32_t incl = my_ins;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += v; }
    if ((tid & 31) == 31) s_warp_ins[tid >> 5] = incl;
    __syncthreads();
    uint32_t ins_before = incl - my_ins;
    for (int w = 0; w < (tid >> 5); ++w) ins_before += s_warp_ins[w];
    // The thread's 16 output symbols in four words, a word at a time and without a branch per symbol (every warp holds a few
    // inserts, so a per-symbol path taken by the threads that hold one is a path the whole warp waits for): the old symbols of a
    // word are consecutive staged bytes; a byte permute spreads them over the positions that are not inserts (selector from a
    // 16-entry table indexed by the word's four flags) and the inserted symbols are blended in under the flag mask.
    uint32_t o[kWords];
    const bool full = q0 + j0 + kPer <= m_new;
    {
        uint32_t ib = ins_before;
#pragma unroll
        for (int w = 0; w < kWords; ++w) {
            // staged byte of the word's first old symbol: the thread's first position less the inserts before the word
            o[w] = fmg::bcr_merge_word(reinterpret_cast<const uint32_t *>(s_src), shift + j0 + 4 * w - (int)ib, flw[w], syw[w], s_spread);
            ib += fmg::bcr_flag_count(flw[w]);
        }
        if (!full) {                                                 // the last tile: positions past the end count as no symbol
#pragma unroll
            for (int t = 0; t < kPer; ++t) if (q0 + j0 + t >= m_new) o[t >> 2] |= 7u << (8 * (t & 3));
        }
    }
    // Bit planes of the thread's symbols (symbol 4w+b at bit 8b+w of each plane) and from them one mask per base: packed A/C/G/T
    // counts of the thread (4 x 16 bit) by population count, then their block-wide exclusive prefix
    uint32_t eq[4];                                                      // A = 1, C = 2, G = 3, T = 4
    fmg::bcr_base_masks<kWords>(o, eq);
    const uint64_t my_cnt = fmg::bcr_pack_counts(eq);
    uint64_t cincl = my_cnt;
    for (int q = 1; q < 32; q <<= 1) { const uint64_t v = __shfl_up_sync(0xffffffffu, cincl, q); if ((tid & 31) >= q) cincl += v; }
    if ((tid & 31) == 31) s_warp_cnt[tid >> 5] = cincl;
    __syncthreads();
    uint64_t cnt_before = cincl - my_cnt;
    for (int w = 0; w < (tid >> 5); ++w) cnt_before += s_warp_cnt[w];
    if (tid == kThreads - 1) {
        const uint64_t tot = cnt_before + my_cnt;
        Vec4 h;
        for (int c = 0; c < 4; ++c) h.v[c] = (tot >> (16 * c)) & 0xffff;
        tile_hist[tile] = h;
    }
    // rank of every insert among equal symbols inside the tile: the prefix of the thread plus the equal symbols at earlier
    // positions of its 16 (a population count under the mask of those positions)
    if (my_ins) {
        uint32_t ib = ins_before;
#pragma unroll
        for (int w = 0; w < kWords; ++w) {
            uint32_t f = flw[w];
            while (f) {
                const int sh = (__ffs(f) - 1) & ~7;                  // 8 x byte of the insert in the word
                f &= f - 1;
                rank_in_tile[k_lo + ib] = fmg::bcr_insert_rank(o[w], w, sh, eq, cnt_before);
                ++ib;
            }
        }
    }
    // write the tile (16 bytes per thread)
    if (full) {
#pragma unroll
        for (int v = 0; v < kVecs; ++v) *reinterpret_cast<uint4 *>(new_bwt + q0 + j0 + 16 * v) = make_uint4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
    }
    else for (int t = 0; t < kPer; ++t) if (q0 + j0 + t < m_new) new_bwt[q0 + j0 + t] = (uint8_t)(o[t >> 2] >> (8 * (t & 3)));
    __syncthreads();                              // every thread is done with this stage's bytes, the flags and the warp sums
    if (tid == 0) {
        const uint64_t nxt = tile + (uint64_t)kStages * gridDim.x;
        if (nxt < n_tiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy reads of the buffer precede the async-proxy write
            issue(nxt, stage);
        }
    }
    k_lo = n_lo; k_hi = n_hi; n_lo = nn_lo; n_hi = nn_hi; my_f = nx_f; my_s = nx_s;
    }
}

// LF mapping of every insert: position of the extended suffix in the next cycle's BWT (bcr.c:442 + set_bwt bookkeeping)
__global__ void k_bcr_lf(Item *__restrict__ item, const uint8_t *__restrict__ sym, const uint32_t *__restrict__ rank_in_tile,
                         const Vec4 *__restrict__ tile_pref, const Vec4 *__restrict__ total, uint64_t n_act, uint64_t n_seq, uint32_t kTile) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_act) return;
    const uint32_t a = sym[k];
    if (a < 1 || a > 4) return;                 // sequence finished: dropped by the partition
    const uint64_t f = item[k].f;
    uint64_t c = n_seq;                         // the sentinel bucket always holds one suffix per sequence
    for (uint32_t b = 1; b < a; ++b) c += total->v[b - 1];
    item[k].f = c + tile_pref[f / kTile].v[a - 1] + rank_in_tile[k];
}

__global__ void k_bcr_total(const Vec4 *tile_hist, const Vec4 *tile_pref, uint64_t n_tiles, Vec4 *total) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        Vec4 t;
        for (int c = 0; c < 4; ++c) t.v[c] = tile_pref[n_tiles - 1].v[c] + tile_hist[n_tiles - 1].v[c];
        *total = t;
    }
}

inline unsigned blocks_for(uint64_t n, int t) { return (unsigned)((n + t - 1) / t); }

} // namespace

struct fmg_bcr_s {
    int device = 0;
    std::vector<uint8_t> seq;        // appended sequences, nt6 1..4
    std::vector<uint64_t> off{0};
    int max_len = 0;
    DevBuf d_bwt;                    // result: one nt6 byte per symbol, kept in HBM; copied out / RLD-encoded on request
    bool built = false;
    uint64_t n_sym = 0;
};

static int bcr_build_device(fmg_bcr_s *b) {
    const uint64_t n_seq = b->off.size() - 1;
    const uint64_t total = b->seq.size() + n_seq;
    if (n_seq == 0) { b->built = true; return 0; }
    if (n_seq >= 0xffffffffull) { if (fmg_verbose >= 1) std::fprintf(stderr, "[E::fmg_bcr_build] more than 2^32-1 sequences\n"); return -1; }
    const auto t_start = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(); };
    DevBuf d_seq, d_off, d_bwt[2], d_item[2], d_sym[2], d_rank, d_hist, d_pref, d_total, d_tmp, d_lo;
    BCR_TRY(d_seq.reserve(b->seq.size())); BCR_TRY(d_off.reserve((n_seq + 1) * 8));
    BCR_TRY(g_stager.upload(d_seq.p, b->seq.data(), b->seq.size()));
    BCR_TRY(g_stager.upload(d_off.p, b->off.data(), (n_seq + 1) * 8));
    const double t_copy = since();
    // reads of (nearly) one length: transposed by cycle; a ragged set keeps the gather through the offsets
    DevBuf d_col;
    const bool by_cycle = (uint64_t)b->max_len * n_seq <= b->seq.size() + b->seq.size() / 4 + (1u << 20) && !std::getenv("FMG_BCR_NO_TRANSPOSE");
    if (by_cycle) {
        BCR_TRY(d_col.reserve((uint64_t)b->max_len * n_seq + 1));
        k_bcr_transpose<<<blocks_for(n_seq, 128), 128>>>(d_seq.as<uint8_t>(), d_off.as<uint64_t>(), n_seq, b->max_len, d_col.as<uint8_t>()); ++g_launches;
        BCR_TRY(cudaGetLastError());
        BCR_TRY(cudaDeviceSynchronize());
        d_seq.release(); d_off.release();               // every cycle reads the columns; the last one (all exhausted) needs neither
    }
    BCR_TRY(d_bwt[0].reserve(total + 64)); BCR_TRY(d_bwt[1].reserve(total + 64));         // the merge stages whole 16-byte words
    BCR_TRY(d_item[0].reserve(n_seq * sizeof(Item))); BCR_TRY(d_item[1].reserve(n_seq * sizeof(Item)));
    BCR_TRY(d_sym[0].reserve(n_seq)); BCR_TRY(d_sym[1].reserve(n_seq)); BCR_TRY(d_rank.reserve(n_seq * 4));
    // block size of the merge pass: small blocks put more tiles in flight per SM (a tile is a chain of barriers and dependent
    // loads: its latency, not its work, bounds the pass)
    // measured (10 M x 150 bp, best of three builds on a shared box): 64 x 32 -> 0.77 s, 64 x 16 -> 0.82 s, 128 x 16 -> 1.01 s, 128 x 32 -> 1.04 s
    int merge_threads = 64, merge_per = 32;
    if (const char *e = std::getenv("FMG_BCR_THREADS")) merge_threads = std::atoi(e) >= 256 ? 256 : std::atoi(e) >= 128 ? 128 : 64;
    if (const char *e = std::getenv("FMG_BCR_PER")) merge_per = std::atoi(e) >= 32 ? 32 : 16;
    const void *merge_kernel = merge_per == 16
        ? (merge_threads == 256 ? (const void *)k_bcr_merge<256, 4, 16> : merge_threads == 128 ? (const void *)k_bcr_merge<128, 8, 16> : (const void *)k_bcr_merge<64, 16, 16>)
        : (merge_threads >= 128 ? (const void *)k_bcr_merge<128, 8, 32> : (const void *)k_bcr_merge<64, 16, 32>);       // static shared memory: no 8192-symbol tile
    // (64 x 32 at 16 blocks per SM: 62 registers without spills, half the warp slots; at 12 blocks and 80 registers the cycles took 6 % longer)
    if (merge_per == 32 && merge_threads > 128) merge_threads = 128;
    const uint32_t kTile = (uint32_t)(merge_threads * merge_per);
    const uint64_t max_tiles = (total + kTile - 1) / kTile;
    BCR_TRY(d_lo.reserve((max_tiles + 2) * 8));
    BCR_TRY(d_hist.reserve(max_tiles * sizeof(Vec4))); BCR_TRY(d_pref.reserve(max_tiles * sizeof(Vec4))); BCR_TRY(d_total.reserve(sizeof(Vec4)));
    size_t tmp_scan = 0, tmp_sort = 0;
    {
        Vec4 zero{};
        BCR_TRY(cub::DeviceScan::ExclusiveScan(nullptr, tmp_scan, d_hist.as<Vec4>(), d_pref.as<Vec4>(), Vec4Add(), zero, (int64_t)max_tiles));
        cub::DoubleBuffer<uint8_t> dk(d_sym[0].as<uint8_t>(), d_sym[1].as<uint8_t>());
        cub::DoubleBuffer<Item> dv(d_item[0].as<Item>(), d_item[1].as<Item>());
        BCR_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, dk, dv, (int64_t)n_seq, 0, 3));
    }
    BCR_TRY(d_tmp.reserve(tmp_scan > tmp_sort ? tmp_scan : tmp_sort));

    uint64_t merge_grid = 148 * 6;
    {
        int n_sm = 0, per_sm = 0;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, b->device);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, merge_kernel, merge_threads, 0);
        if (n_sm > 0 && per_sm > 0) merge_grid = (uint64_t)n_sm * per_sm;
        if (const char *e = std::getenv("FMG_BCR_BLOCKS")) merge_grid = (uint64_t)n_sm * std::max(1, std::atoi(e));
    }
    const double t_in = since();
    cub::DoubleBuffer<uint8_t> syms(d_sym[0].as<uint8_t>(), d_sym[1].as<uint8_t>());
    cub::DoubleBuffer<Item> items(d_item[0].as<Item>(), d_item[1].as<Item>());
    k_bcr_iota<<<blocks_for(n_seq, 256), 256>>>(items.Current(), n_seq); ++g_launches;
    uint64_t n_act = n_seq, m = 0;
    int cur = 0;
    Vec4 h_total{}, h_prev{};
    for (int pos = 0; n_act > 0; ++pos) {
        const uint64_t m_new = m + n_act, n_tiles = (m_new + kTile - 1) / kTile;
        if (by_cycle && pos < b->max_len) k_bcr_symbols_col<<<blocks_for(n_act, 256), 256>>>(items.Current(), n_act, d_col.as<uint8_t>() + (uint64_t)pos * n_seq, syms.Current());
        else if (by_cycle) BCR_TRY(cudaMemsetAsync(syms.Current(), 0, n_act));       // cycle max_len: every sequence left inserts its sentinel
        else k_bcr_symbols<<<blocks_for(n_act, 256), 256>>>(items.Current(), n_act, d_seq.as<uint8_t>(), d_off.as<uint64_t>(), pos, syms.Current());
        ++g_launches;
        k_bcr_bounds<<<blocks_for(n_act, 256), 256>>>(items.Current(), n_act, n_tiles, d_lo.as<uint64_t>(), kTile); ++g_launches;
        {
            const uint8_t *a_old = d_bwt[cur].as<uint8_t>(); uint64_t a_m = m_new; const Item *a_item = items.Current(); const uint8_t *a_sym = syms.Current();
            const uint64_t *a_lo = d_lo.as<uint64_t>(); uint8_t *a_new = d_bwt[cur ^ 1].as<uint8_t>(); Vec4 *a_hist = d_hist.as<Vec4>(); uint32_t *a_rank = d_rank.as<uint32_t>();
            uint64_t a_tiles = n_tiles;
            void *kargs[] = {&a_old, &a_m, &a_item, &a_sym, &a_lo, &a_new, &a_hist, &a_rank, &a_tiles};
            BCR_TRY(cudaLaunchKernel(merge_kernel, dim3((unsigned)std::min<uint64_t>(n_tiles, merge_grid)), dim3(merge_threads), kargs, 0, nullptr));
            ++g_launches;
        }
        size_t need = d_tmp.cap;
        Vec4 zero{};
        BCR_TRY(cub::DeviceScan::ExclusiveScan(d_tmp.p, need, d_hist.as<Vec4>(), d_pref.as<Vec4>(), Vec4Add(), zero, (int64_t)n_tiles));
        k_bcr_total<<<1, 32>>>(d_hist.as<Vec4>(), d_pref.as<Vec4>(), n_tiles, d_total.as<Vec4>()); ++g_launches;
        k_bcr_lf<<<blocks_for(n_act, 256), 256>>>(items.Current(), syms.Current(), d_rank.as<uint32_t>(), d_pref.as<Vec4>(), d_total.as<Vec4>(), n_act, n_seq, kTile); ++g_launches;
        // next order = stable partition by the inserted symbol; finished sequences (symbol 0) sort first and are dropped
        need = d_tmp.cap;
        BCR_TRY(cub::DeviceRadixSort::SortPairs(d_tmp.p, need, syms, items, (int64_t)n_act, 0, 3));
        BCR_TRY(cudaMemcpy(&h_total, d_total.p, sizeof(Vec4), cudaMemcpyDeviceToHost));
        // sequences that inserted their sentinel in this cycle: all inserts that were not A/C/G/T
        uint64_t bases_now = 0, bases_prev = 0;
        for (int c = 0; c < 4; ++c) bases_now += h_total.v[c], bases_prev += h_prev.v[c];
        const uint64_t finished = n_act - (bases_now - bases_prev);
        h_prev = h_total;
        // drop them: the sorted arrays start with the `finished` zero-symbol entries
        if (finished) {
            BCR_TRY(cudaMemcpyAsync(items.Alternate(), items.Current() + finished, (n_act - finished) * sizeof(Item), cudaMemcpyDeviceToDevice));
            items.selector ^= 1;
        }
        n_act -= finished;
        m = m_new;
        cur ^= 1;
        if (fmg_verbose >= 4) std::fprintf(stderr, "[M::fmg_bcr_build] cycle %d: %llu symbols, %llu sequences still active\n", pos, (unsigned long long)m, (unsigned long long)n_act);
    }
    BCR_TRY(cudaGetLastError());
    BCR_TRY(cudaDeviceSynchronize());
    const double t_cycles = since();
    b->n_sym = m;
    std::swap(b->d_bwt.p, d_bwt[cur].p); std::swap(b->d_bwt.cap, d_bwt[cur].cap); std::swap(b->d_bwt.dev, d_bwt[cur].dev);     // the BWT stays in HBM with the handle
    b->built = true;
    if (fmg_verbose >= 3)
        std::fprintf(stderr, "[M::fmg_bcr_build] %llu sequences, %llu symbols: allocation + copy in %.3f s, %d cycles %.3f s (copy in alone %.3f s)\n", (unsigned long long)n_seq,
                     (unsigned long long)m, t_in, b->max_len + 1, t_cycles - t_in, t_copy);
    return 0;
}

extern "C" {

fmg_bcr_t *fmg_bcr_init(int device) {
    fmg_bcr_s *b = new fmg_bcr_s;
    b->device = device;
    return b;
}

void fmg_bcr_destroy(fmg_bcr_t *b) { delete b; }

int fmg_bcr_append(fmg_bcr_t *b, int len, const uint8_t *seq) {
    if (!b || len < 1) return -1;
    for (int i = 0; i < len; ++i)
        if (seq[i] < 1 || seq[i] > 4) {
            if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] only A/C/G/T (nt6 1..4) are supported, like bcr_append (ropebwt.c:98,118)\n", __func__);
            return -1;
        }
    b->seq.insert(b->seq.end(), seq, seq + len);
    b->off.push_back(b->seq.size());
    if (len > b->max_len) b->max_len = len;
    return 0;
}

int fmg_bcr_append_batch(fmg_bcr_t *b, int64_t n, int len, const uint8_t *seqs) {
    if (!b || len < 1 || n < 0) return -1;
    const size_t base = b->seq.size(), add = (size_t)n * len;
    unsigned bad = 0;                                     // branch-free so that the compiler vectorises the scan of gigabytes
    for (size_t i = 0; i < add; ++i) bad |= (unsigned)(uint8_t)(seqs[i] - 1) > 3u;
    if (bad) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] only A/C/G/T (nt6 1..4) are supported\n", __func__);
        return -1;
    }
    b->seq.insert(b->seq.end(), seqs, seqs + add);
    b->off.reserve(b->off.size() + (size_t)n);
    for (int64_t i = 1; i <= n; ++i) b->off.push_back(base + (size_t)i * len);
    if (len > b->max_len) b->max_len = len;
    return 0;
}

int fmg_bcr_build(fmg_bcr_t *b) {
    if (!b) return -1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] no CUDA device available; libfermi_b200 has no CPU path\n", __func__);
        return -1;
    }
    if (cudaSetDevice(b->device) != cudaSuccess) return -1;
    return bcr_build_device(b);
}

int64_t fmg_bcr_size(const fmg_bcr_t *b) { return b && b->built ? (int64_t)b->n_sym : -1; }

// `fermi ropebwt | fermi recode` in one step: the BWT in HBM through the RLD encoder on the device; the caller owns the image
fmg_fmd_t *fmg_bcr_fmd(fmg_bcr_t *b) {
    if (!b || !b->built || !b->n_sym || cudaSetDevice(b->device) != cudaSuccess) return nullptr;
    fmg_fmd_t *e = new fmg_fmd_s;
    if (fmg_rld_encode_device(b->d_bwt.as<uint8_t>(), b->n_sym, &e->img) != 0) { delete e; return nullptr; }
    return e;
}

int fmg_bcr_bwt(const fmg_bcr_t *b, uint8_t *bwt) {
    if (!b || !b->built || cudaSetDevice(b->device) != cudaSuccess) return -1;
    if (b->n_sym && cudaMemcpy(bwt, b->d_bwt.p, b->n_sym, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return 0;
}

// the byte run-length stream of bcr_itr_next / `fermi ropebwt -b` (ropebwt.c:127-144): bytes len<<3|sym, len <= 31
int fmg_bcr_rle(const fmg_bcr_t *b, uint8_t **rle, int64_t *n) {
    if (!b || !b->built) return -1;
    std::unique_ptr<uint8_t[]> w(new uint8_t[b->n_sym ? b->n_sym : 1]);
    if (fmg_bcr_bwt(b, w.get()) != 0) return -1;
    std::vector<uint8_t> out;
    out.reserve(b->n_sym / 2 + 16);
    for (size_t i = 0; i < b->n_sym;) {
        size_t j = i;
        while (j < b->n_sym && w[j] == w[i] && j - i < 31) ++j;
        out.push_back((uint8_t)((j - i) << 3 | w[i]));
        i = j;
    }
    *rle = (uint8_t *)std::malloc(out.size() ? out.size() : 1);
    std::memcpy(*rle, out.data(), out.size());
    *n = (int64_t)out.size();
    return 0;
}

} // extern "C"
