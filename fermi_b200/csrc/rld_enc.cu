// RLD encoder on the GPU: BWT symbols in HBM -> the reference's exact "RLD\2" bit stream (rld_enc / rld_enc1 /
// enc_next_block / rld_enc_finish, rld.c:111-236; layout summary in SURVEY.md 8a R2-R3).  Replaces the serial host
// encoder (FmdEncoder, fmd_host.cpp) wherever the BWT is produced on the device (fmg_build_fmd, fmg_bcr_*).
//
// The stream is a sequence of 64-byte blocks, each filled greedily with Elias-delta run codes: a code goes into the
// current block iff the payload bits used so far plus its width stay BELOW the payload size (rld.c:164), the payload
// size depends on the header width (7 x u16, or 7 x u32 when the previous block held >= 0x8000 symbols, rld.c:119-124)
// and on whether the block is the last of a 2^23-word chunk (one word less, rld.h:66).  So where block b+1 starts is a
// function of where block b starts -- a serial chain over ~n/100 blocks.  It is resolved exactly by pointer doubling over
// the runs (see k_jump0 .. k_fill_blocks below); then one thread per block packs its codes and header into registers and
// writes the 64 bytes.
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

// ---- the chain by pointer doubling.  Almost every block has a 7 x u16 header and the full payload ("regular"): for those,
// where the next block starts is a function of where this one starts alone, jump[r] = next_block(r) - r.  K rounds of
// doubling give every run the start of the block 2^K blocks further on (as a distance, with a flag that is set when one of
// those blocks holds >= 0x8000 symbols, i.e. changes the header width of its successor).  One warp then walks the chain in
// groups of 2^K blocks -- one table look-up per regular group, block by block through a group with a flagged block or with
// the shortened last block of a 2^23-word chunk (which is always the last block of its group) -- and finally one thread per
// group fills in the block starts between two anchors with the exact rule.
constexpr int kJumpLog = 10;
constexpr uint64_t kGroup = 1ull << kJumpLog;     // divides kChunkBlocks
constexpr uint32_t kBig = 1u << 31;

__global__ void __launch_bounds__(256) k_jump0(Chain c, uint32_t *__restrict__ jump) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > c.n_runs) return;
    if (r == c.n_runs) { jump[r] = 0; return; }                  // the end of the chain is absorbing
    const uint64_t nx = next_block(c, r, 0, 0);
    jump[r] = (uint32_t)(nx - r) | (c.pos[nx] - c.pos[r] >= 0x8000 ? kBig : 0u);
}

__global__ void __launch_bounds__(256) k_jump_double(uint64_t n_runs, const uint32_t *__restrict__ in, uint32_t *__restrict__ out) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_runs) return;
    const uint32_t a = in[r], b = in[r + (a & ~kBig)];
    out[r] = ((a & ~kBig) + (b & ~kBig)) | ((a | b) & kBig);
}

// next_block with the 32 lanes of a warp probing the <= 100 candidate ends at once (one round trip instead of seven)
__device__ __forceinline__ uint64_t next_block_warp(const Chain &c, uint64_t r, int h, uint64_t b, int lane) {
    const int words = ((b + 1) % kChunkBlocks == 0 ? 7 : 8) - (h ? kHeaderWords32 : kHeaderWords16);
    const uint64_t limit = c.W[r] + (uint64_t)words * 64 - 1;
    uint64_t best = r + 1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint64_t t = r + 1 + (uint64_t)(q * 32 + lane);
        const bool ok = t <= c.n_runs && t <= r + 100 && c.W[t] <= limit;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (m) best = r + 1 + (uint64_t)(q * 32 + (31 - __clz(m)));      // W is increasing: the fitting ends are a prefix
    }
    return best;
}

// anchors[g] = first run of block g * 2^K (bit 63: that block has a 7 x u32 header); *n_groups = number of anchors
__global__ void __launch_bounds__(32) k_anchor_chase(Chain c, const uint32_t *__restrict__ jumpK, uint64_t *anchors, uint64_t *n_groups) {
    const int lane = threadIdx.x;
    uint64_t r = 0, g = 0;
    int h = 0;
    while (r < c.n_runs) {
        if (lane == 0) anchors[g] = r | (uint64_t)h << 63;
        const uint32_t j = jumpK[r];
        if (h == 0 && !(j & kBig) && (g + 1) % (kChunkBlocks / kGroup) != 0) r += j;
        else
            for (uint64_t i = 0; i < kGroup && r < c.n_runs; ++i) {
                const uint64_t nx = next_block_warp(c, r, h, g * kGroup + i, lane);
                h = c.pos[nx] - c.pos[r] >= 0x8000;
                r = nx;
            }
        ++g;
    }
    if (lane == 0) *n_groups = g;
}

// block starts of one group from its anchor with the exact rule; the last group reports how many blocks it holds
__global__ void __launch_bounds__(128) k_fill_blocks(Chain c, const uint64_t *__restrict__ anchors, uint64_t n_groups, uint64_t *__restrict__ bstart,
                                                    uint64_t *last_count, uint8_t *last_h) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    uint64_t r = anchors[g] & ~(1ull << 63), i = 0;
    int h = (int)(anchors[g] >> 63);
    for (; i < kGroup && r < c.n_runs; ++i) {
        const uint64_t b = g * kGroup + i;
        bstart[b] = r;
        const uint64_t nx = next_block(c, r, h, b);
        h = c.pos[nx] - c.pos[r] >= 0x8000;
        r = nx;
    }
    if (g == n_groups - 1) { *last_count = i; *last_h = (uint8_t)h; }
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

    // ---- the block chain: jump distances, K doublings, anchors every 2^K blocks, block starts
    const uint64_t max_groups = n_runs / kGroup + 2;                     // every block holds at least one run
    Dev &d_j0 = d_run[0], &d_j1 = d_run[1], &d_anchor = d_blk[0], &d_cnt = d_blk[1];
    RE_TRY(d_j0.alloc((n_runs + 2) * 4)); RE_TRY(d_j1.alloc((n_runs + 2) * 4)); RE_TRY(d_anchor.alloc(max_groups * 8)); RE_TRY(d_cnt.alloc(64));
    k_jump0<<<grid_for(n_runs + 1, 256), 256>>>(c, d_j0.as<uint32_t>());
    uint32_t *ja = d_j0.as<uint32_t>(), *jb = d_j1.as<uint32_t>();
    for (int k = 0; k < kJumpLog; ++k) {
        k_jump_double<<<grid_for(n_runs + 1, 256), 256>>>(n_runs, ja, jb);
        std::swap(ja, jb);
    }
    k_anchor_chase<<<1, 32>>>(c, ja, d_anchor.as<uint64_t>(), d_cnt.as<uint64_t>());
    g_launches += kJumpLog + 2;
    uint64_t n_groups = 0;
    RE_TRY(cudaMemcpy(&n_groups, d_cnt.p, 8, cudaMemcpyDeviceToHost));
    if (n_groups == 0 || n_groups > max_groups) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] the block chain has an impossible length\n", __func__);
        return -2;
    }
    RE_TRY(d_bstart.alloc((n_groups * kGroup + 1) * 8));
    k_fill_blocks<<<grid_for(n_groups, 128), 128>>>(c, d_anchor.as<uint64_t>(), n_groups, d_bstart.as<uint64_t>(), d_cnt.as<uint64_t>() + 1,
                                                    reinterpret_cast<uint8_t *>(d_cnt.as<uint64_t>() + 2));
    ++g_launches;
    uint64_t tail[2] = {0, 0};
    RE_TRY(cudaMemcpy(tail, d_cnt.as<uint64_t>() + 1, 16, cudaMemcpyDeviceToHost));
    const uint64_t n_blocks = (n_groups - 1) * kGroup + tail[0];
    const uint8_t h_last = (uint8_t)(tail[1] & 0xff);
    const int rounds = kJumpLog - 1;
    const uint64_t n_seg = n_groups;
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
        std::fprintf(stderr, "[M::%s] %llu symbols, %llu runs, %llu blocks: runs %.3f s, block chain %.3f s (%d doublings, %llu groups), pack + copy + directory %.3f s\n",
                     __func__, (unsigned long long)n, (unsigned long long)n_runs, (unsigned long long)n_blocks, t_runs, t_chain - t_runs, rounds + 1,
                     (unsigned long long)n_seg, since(t0) - t_chain);
    return 0;
}
