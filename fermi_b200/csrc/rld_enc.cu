// RLD encoder on the GPU: BWT symbols in HBM -> the reference's exact "RLD\2" bit stream (rld_enc / rld_enc1 /
// enc_next_block / rld_enc_finish, rld.c:111-236; layout summary in SURVEY.md 8a R2-R3).  Replaces the serial host
// encoder (FmdEncoder, fmd_host.cpp) wherever the BWT is produced on the device (fmg_build_fmd, fmg_bcr_*).
//
// The stream is a sequence of 64-byte blocks, each filled greedily with Elias-delta run codes: a code goes into the
// current block iff the payload bits used so far plus its width stay BELOW the payload size (rld.c:164), the payload
// size depends on the header width (7 x u16, or 7 x u32 when the previous block held >= 0x8000 symbols, rld.c:119-124)
// and on whether the block is the last of a 2^23-word chunk (one word less, rld.h:66).  So where block b+1 starts is a
// function of where block b starts -- a serial chain over ~n/100 blocks.  It is cut into segments of kSeg runs that
// are chased speculatively in parallel: every segment assumes an entry state (first run of a block, header width, block
// number), chases its blocks, and hands its exit state to the next segment; rounds repeat until no segment's entry
// changes.  Greedy packings started a few runs apart fall into step after a handful of blocks, so the fixpoint -- which is
// exactly the serial result, by induction from segment 0 -- arrives after a few rounds.  Then one thread per block packs
// its codes and header into registers and writes the 64 bytes.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <atomic>
#include <chrono>
#include <vector>
#include <algorithm>
#include "fmd_host.hpp"
#include "dev_pool.hpp"
#include "../../include/fermi_b200.h"

using namespace fmg;

extern std::atomic<uint64_t> g_launches;

#define RE_TRY(call)                                                                                  \
    do {                                                                                              \
        cudaError_t err__ = (call);                                                                   \
        if (err__ != cudaSuccess) {                                                                   \
            if (fmg_verbose >= 1)                                                                     \
                std::fprintf(stderr, "[E::%s] %s failed: %s\n", __func__, #call, cudaGetErrorString(err__)); \
            return -1;                                                                                \
        }                                                                                             \
    } while (0)

namespace {

constexpr int kTile = 256 * 16;              // symbols per thread block in the run detection
constexpr int kSeg = 4096;                   // runs per speculative segment
constexpr uint64_t kChunkBlocks = 1ull << 20; // blocks per 2^23-word chunk (rld.h:9-10)

__device__ __forceinline__ int ilog2_u32(uint32_t v) { return 31 - __clz(v); }

// width and bits of the code of one run (rld_delta_enc1, rld.c:47-53, + 3-bit symbol, rld.c:159-173)
__device__ __forceinline__ int code_width(uint64_t len) {
    const int y = ilog2_u32((uint32_t)len), z = ilog2_u32((uint32_t)(y + 1));
    return 2 * z + 1 + y + 3;
}
__device__ __forceinline__ uint64_t code_bits(uint64_t len, int sym) {
    const int y = ilog2_u32((uint32_t)len);
    return ((((len ^ (1ull << y)) | (uint64_t)(y + 1) << y)) << 3) | (uint64_t)sym;
}

// ---- runs: a run starts at i when bwt[i] != bwt[i-1]
__global__ void __launch_bounds__(256) k_run_count(const uint8_t *__restrict__ bwt, uint64_t n, uint64_t *tile_cnt, unsigned long long *hist) {
    __shared__ unsigned int sh[8];
    if (threadIdx.x < 8) sh[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kTile;
    unsigned c = 0;
    for (int k = 0; k < 16; ++k) {
        const uint64_t i = base + (uint64_t)k * 256 + threadIdx.x;
        if (i < n) {
            const uint8_t s = bwt[i];
            c += (i == 0 || bwt[i - 1] != s);
            atomicAdd(&sh[s & 7], 1u);
        }
    }
    typedef cub::BlockReduce<unsigned, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const unsigned tot = BR(tmp).Sum(c);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = tot;
    __syncthreads();
    if (threadIdx.x < 8 && sh[threadIdx.x]) atomicAdd(hist + threadIdx.x, (unsigned long long)sh[threadIdx.x]);
}

__global__ void __launch_bounds__(256) k_run_scatter(const uint8_t *__restrict__ bwt, uint64_t n, const uint64_t *__restrict__ tile_off,
                                                    uint64_t *__restrict__ run_pos, uint8_t *__restrict__ run_sym) {
    typedef cub::BlockScan<unsigned, 256> BS;
    __shared__ typename BS::TempStorage tmp;
    const uint64_t base = (uint64_t)blockIdx.x * kTile + (uint64_t)threadIdx.x * 16;      // 16 consecutive symbols per thread
    unsigned f = 0, c = 0;
    uint8_t s[16];
    uint8_t prev = base > 0 && base <= n ? bwt[base - 1] : 0xff;
    for (int k = 0; k < 16; ++k) {
        const uint64_t i = base + k;
        s[k] = i < n ? bwt[i] : 0xff;
        if (i < n && (i == 0 || s[k] != prev)) { f |= 1u << k; ++c; }
        prev = s[k];
    }
    unsigned ex;
    BS(tmp).ExclusiveSum(c, ex);
    uint64_t o = tile_off[blockIdx.x] + ex;
    for (int k = 0; k < 16; ++k)
        if (f >> k & 1) { run_pos[o] = base + k; run_sym[o] = s[k]; ++o; }
}

__global__ void __launch_bounds__(256) k_widths(const uint64_t *__restrict__ pos, uint64_t n_runs, uint64_t *__restrict__ W) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n_runs) W[i] = i < n_runs ? (uint64_t)code_width(pos[i + 1] - pos[i]) : 0;
}

// ---- greedy block chain
struct Chain {
    const uint64_t *W;        // n_runs + 1: code bits before run i
    const uint64_t *pos;      // n_runs + 1: symbols before run i (pos[n_runs] = n)
    uint64_t n_runs;
};

// the block starting at run r with header width h (0: 7 x u16, 1: 7 x u32) as block number b: first run of the next block
__device__ __forceinline__ uint64_t next_block(const Chain &c, uint64_t r, int h, uint64_t b) {
    const int words = ((b + 1) % kChunkBlocks == 0 ? 7 : 8) - (h ? kHeaderWords32 : kHeaderWords16);
    const uint64_t limit = c.W[r] + (uint64_t)words * 64 - 1;         // runs r..q fit iff W[q+1] <= limit
    uint64_t lo = r + 1, hi = min(c.n_runs, r + 100);                 // a block holds at most 96 codes (4 bits each at least)
    // largest t in [lo, hi] with W[t] <= limit (W[r+1] - W[r] <= 48 always fits)
    while (lo < hi) {
        const uint64_t mid = (lo + hi + 1) >> 1;
        if (c.W[mid] <= limit) lo = mid; else hi = mid - 1;
    }
    return lo;
}

struct SegState { uint64_t run; uint64_t blk; int h; };

// chase the blocks of one segment from its entry state; returns the exit state = entry of the next segment
__device__ __forceinline__ SegState chase(const Chain &c, SegState s, uint64_t seg_end, uint64_t *bstart) {
    while (s.run < seg_end) {
        if (bstart) bstart[s.blk] = s.run;
        const uint64_t nx = next_block(c, s.run, s.h, s.blk);
        s.h = c.pos[nx] - c.pos[s.run] >= 0x8000;
        s.run = nx;
        ++s.blk;
    }
    return s;
}

// One round: segment s chases its blocks from its entry state (ent_run[s], ent_h[s]) numbered from blk0[s], records how many
// it holds and hands its exit to segment s+1.  Block numbers only matter for the shortened last block of a chunk; they come
// from a scan of n_blk between rounds.  A round that changes no entry and no count has reached the serial answer.
__global__ void __launch_bounds__(128) k_seg_chase(Chain c, uint64_t n_seg, uint64_t *ent_run, uint8_t *ent_h, const uint64_t *__restrict__ blk0,
                                                  uint64_t *n_blk, unsigned long long *changed, uint64_t *bstart) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    SegState e;
    e.run = ent_run[s]; e.blk = blk0[s]; e.h = ent_h[s];
    const uint64_t seg_end = min(c.n_runs, (s + 1) * (uint64_t)kSeg);
    const SegState x = chase(c, e, seg_end, bstart);
    if (bstart) return;
    const uint64_t nb = x.blk - e.blk;
    bool chg = false;
    if (n_blk[s] != nb) { n_blk[s] = nb; chg = true; }
    // slot n_seg holds the end of the chain
    if (ent_run[s + 1] != x.run || ent_h[s + 1] != (uint8_t)x.h) { ent_run[s + 1] = x.run; ent_h[s + 1] = (uint8_t)x.h; chg = true; }
    if (chg) atomicAdd(changed, 1ull);
}

// ---- one thread per block: header (counts of the previous block, rld.c:111-134) + codes, 64 bytes
__global__ void __launch_bounds__(128) k_pack(Chain c, const uint8_t *__restrict__ run_sym, const uint64_t *__restrict__ bstart, uint64_t n_blocks,
                                             uint64_t *__restrict__ words) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b > n_blocks) return;                                   // block n_blocks is the trailing header-only block (rld.c:230)
    uint64_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int hw = kHeaderWords16;
    if (b > 0) {
        uint64_t d[7] = {0, 0, 0, 0, 0, 0, 0};
        const uint64_t p0 = bstart[b - 1], p1 = b < n_blocks ? bstart[b] : c.n_runs;
        for (uint64_t r = p0; r < p1; ++r) {
            const uint64_t len = c.pos[r + 1] - c.pos[r];
            const int s = run_sym[r];
            d[0] += len;
#pragma unroll
            for (int q = 0; q < 6; ++q) d[q + 1] += s == q ? len : 0;
        }
        if (d[0] >= 0x8000) {
            hw = kHeaderWords32;
            w[0] = (d[0] | 1ull << 31) | d[1] << 32; w[1] = d[2] | d[3] << 32; w[2] = d[4] | d[5] << 32; w[3] = d[6];
        } else {
            w[0] = d[0] | d[1] << 16 | d[2] << 32 | d[3] << 48; w[1] = d[4] | d[5] << 16 | d[6] << 32;
        }
    }
    uint64_t *dst = words + b * kBlockWords;
    if (b == n_blocks) {
        for (int q = 0; q < hw; ++q) dst[q] = w[q];
        return;
    }
    const uint64_t r0 = bstart[b], r1 = b + 1 < n_blocks ? bstart[b + 1] : c.n_runs;
    int p = hw, room = 64;
    for (uint64_t r = r0; r < r1; ++r) {
        const uint64_t len = c.pos[r + 1] - c.pos[r];
        const uint64_t code = code_bits(len, run_sym[r]);
        int width = code_width(len);
        if (width > room) {                                     // straddles two words (rld.c:166-170)
            width -= room;
            w[p & 7] |= code >> width;
            ++p;
            room = 64 - width;
            w[p & 7] = code << room;
        } else {
            room -= width;
            w[p & 7] |= code << room;
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) dst[q] = w[q];
}

inline unsigned grid_for(uint64_t n, int per) { return (unsigned)((n + per - 1) / per); }

}  // namespace

// d_bwt: n nt6 symbols (0..5) in the HBM of the current device -> the .fmd image (stream words, counts, rank directory)
int fmg_rld_encode_device(const uint8_t *d_bwt, uint64_t n, FmdImage *out) {
    if (!out || n == 0) return -1;
    const auto t0 = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); };
    Dev d_tile, d_hist, d_pos, d_sym, d_W, d_tmp, d_run[2], d_blk[2], d_h[2], d_chg, d_bstart, d_words;
    const uint64_t n_tiles = (n + kTile - 1) / kTile;
    RE_TRY(d_tile.alloc((n_tiles + 1) * 8)); RE_TRY(d_hist.alloc(64));
    RE_TRY(cudaMemset(d_hist.p, 0, 64));
    RE_TRY(cudaMemset(d_tile.as<uint64_t>() + n_tiles, 0, 8));
    k_run_count<<<(unsigned)n_tiles, 256>>>(d_bwt, n, d_tile.as<uint64_t>(), d_hist.as<unsigned long long>());
    ++g_launches;
    size_t need = 0;
    RE_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, d_tile.as<uint64_t>(), d_tile.as<uint64_t>(), n_tiles + 1));
    RE_TRY(d_tmp.alloc(need + 256));
    RE_TRY(cub::DeviceScan::ExclusiveSum(d_tmp.p, need, d_tile.as<uint64_t>(), d_tile.as<uint64_t>(), n_tiles + 1));
    uint64_t n_runs = 0, hist[8];
    RE_TRY(cudaMemcpy(&n_runs, d_tile.as<uint64_t>() + n_tiles, 8, cudaMemcpyDeviceToHost));
    RE_TRY(cudaMemcpy(hist, d_hist.p, 64, cudaMemcpyDeviceToHost));
    if (hist[6] || hist[7]) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] the BWT holds symbols outside 0..5\n", __func__);
        return -1;
    }
    RE_TRY(d_pos.alloc((n_runs + 1) * 8)); RE_TRY(d_sym.alloc(n_runs + 1)); RE_TRY(d_W.alloc((n_runs + 2) * 8));
    k_run_scatter<<<(unsigned)n_tiles, 256>>>(d_bwt, n, d_tile.as<uint64_t>(), d_pos.as<uint64_t>(), d_sym.as<uint8_t>());
    ++g_launches;
    RE_TRY(cudaMemcpy(d_pos.as<uint64_t>() + n_runs, &n, 8, cudaMemcpyHostToDevice));
    // W[i] = code bits before run i (n_runs + 1 entries; the last one is the length of the whole code stream)
    {
        k_widths<<<(unsigned)((n_runs + 256) / 256), 256>>>(d_pos.as<uint64_t>(), n_runs, d_W.as<uint64_t>());
        size_t need2 = 0;
        RE_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need2, d_W.as<uint64_t>(), d_W.as<uint64_t>(), n_runs + 1));
        if (need2 > need) { RE_TRY(d_tmp.alloc(need2 + 256)); need = need2; }
        RE_TRY(cub::DeviceScan::ExclusiveSum(d_tmp.p, need2, d_W.as<uint64_t>(), d_W.as<uint64_t>(), n_runs + 1));
        g_launches += 2;
    }
    Chain c;
    c.W = d_W.as<uint64_t>(); c.pos = d_pos.as<uint64_t>(); c.n_runs = n_runs;
    const double t_runs = since(t0);

    // ---- speculative segments until the entry states and block counts stop changing
    const uint64_t n_seg = (n_runs + kSeg - 1) / kSeg;
    Dev &d_ent = d_run[0], &d_blk0 = d_blk[0], &d_nblk = d_blk[1], &d_eh = d_h[0];
    RE_TRY(d_ent.alloc((n_seg + 1) * 8)); RE_TRY(d_blk0.alloc((n_seg + 1) * 8)); RE_TRY(d_nblk.alloc((n_seg + 1) * 8));
    RE_TRY(d_eh.alloc(n_seg + 1)); RE_TRY(d_chg.alloc(8));
    {   // initial guess: every segment starts a block at its first run with a 16-bit header
        std::vector<uint64_t> r0(n_seg + 1);
        for (uint64_t s = 0; s <= n_seg; ++s) r0[s] = std::min<uint64_t>(s * kSeg, n_runs);
        RE_TRY(cudaMemcpy(d_ent.p, r0.data(), (n_seg + 1) * 8, cudaMemcpyHostToDevice));
        RE_TRY(cudaMemset(d_blk0.p, 0, (n_seg + 1) * 8));
        RE_TRY(cudaMemset(d_nblk.p, 0, (n_seg + 1) * 8));
        RE_TRY(cudaMemset(d_eh.p, 0, n_seg + 1));
    }
    size_t need3 = 0;
    RE_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need3, d_nblk.as<uint64_t>(), d_blk0.as<uint64_t>(), n_seg + 1));
    if (need3 > need) { RE_TRY(d_tmp.alloc(need3 + 256)); need = need3; }
    int rounds = 0;
    for (;; ++rounds) {
        if (rounds > 4096) {
            if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] the block chain did not settle after %d rounds\n", __func__, rounds);
            return -2;
        }
        RE_TRY(cudaMemset(d_chg.p, 0, 8));
        // entries are updated in place: a slot rewritten during this round is read by its owner now or in the next round, and the
        // loop only ends after a round in which nothing was rewritten (then every read was of settled values)
        k_seg_chase<<<grid_for(n_seg, 128), 128>>>(c, n_seg, d_ent.as<uint64_t>(), d_eh.as<uint8_t>(), d_blk0.as<uint64_t>(), d_nblk.as<uint64_t>(),
                                                   d_chg.as<unsigned long long>(), nullptr);
        ++g_launches;
        unsigned long long chg = 0;
        RE_TRY(cudaMemcpy(&chg, d_chg.p, 8, cudaMemcpyDeviceToHost));
        if (chg == 0) break;
        RE_TRY(cub::DeviceScan::ExclusiveSum(d_tmp.p, need3, d_nblk.as<uint64_t>(), d_blk0.as<uint64_t>(), n_seg + 1));
        ++g_launches;
    }
    uint64_t n_blocks = 0;
    RE_TRY(cudaMemcpy(&n_blocks, d_blk0.as<uint64_t>() + n_seg, 8, cudaMemcpyDeviceToHost));
    uint8_t h_last = 0;
    RE_TRY(cudaMemcpy(&h_last, d_eh.as<uint8_t>() + n_seg, 1, cudaMemcpyDeviceToHost));
    RE_TRY(d_bstart.alloc((n_blocks + 1) * 8));
    k_seg_chase<<<grid_for(n_seg, 128), 128>>>(c, n_seg, d_ent.as<uint64_t>(), d_eh.as<uint8_t>(), d_blk0.as<uint64_t>(), nullptr, nullptr, d_bstart.as<uint64_t>());
    ++g_launches;
    const double t_chain = since(t0);

    // ---- pack
    const uint64_t n_words = n_blocks * kBlockWords + (h_last ? kHeaderWords32 : kHeaderWords16);
    RE_TRY(d_words.alloc((n_blocks + 1) * kBlockWords * 8));
    k_pack<<<grid_for(n_blocks + 1, 128), 128>>>(c, d_sym.as<uint8_t>(), d_bstart.as<uint64_t>(), n_blocks, d_words.as<uint64_t>());
    ++g_launches;
    RE_TRY(cudaGetLastError());
    out->words.assign(n_words + 2, 0);
    RE_TRY(cudaMemcpy(out->words.data(), d_words.p, n_words * 8, cudaMemcpyDeviceToHost));
    out->n_bytes = n_words * 8;
    for (int q = 0; q < 8; ++q) out->mcnt[q] = 0;
    for (int q = 0; q < 6; ++q) out->mcnt[q + 1] = hist[q];
    out->finish_counts();
    out->build_frames();                      // rld_rank_index (rld.c:186-224): one pass over the block headers, host
    if (fmg_verbose >= 4)
        std::fprintf(stderr, "[M::%s] %llu symbols, %llu runs, %llu blocks: runs %.3f s, block chain %.3f s (%d rounds over %llu segments), pack + copy + directory %.3f s\n",
                     __func__, (unsigned long long)n, (unsigned long long)n_runs, (unsigned long long)n_blocks, t_runs, t_chain - t_runs, rounds + 1,
                     (unsigned long long)n_seg, since(t0) - t_chain);
    return 0;
}
