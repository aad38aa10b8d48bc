// Index upload: .fmd image -> occ blocks in HBM (layout in fmd_device.cuh), transcoded ON the GPU.
//
//   k_rld_block_sizes   symbols per 64-byte RLD block, read from the header of the next block (rld.c:111-134)
//   (exclusive scan)    -> BWT coordinate of the first symbol of every RLD block
//   k_rld_decode        one thread per RLD block: Elias-delta run decode (rld.h:77-94) and bit-plane fill
//   k_occ_pad           symbol 7 past the end of the BWT
//   k_occ_block_counts  per occ block: symbol totals           (exclusive scan over blocks)
//   k_occ_finalize      mid-block counts relative to the superblock, byte-packed; cs[] table
// The host-side builder (occ_build_host.cpp) produces the same bytes and is kept for tests (FMG_HOST_OCC_BUILD=1).
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include "fmg_internal.hpp"
#include "occ_layout.hpp"
#include "../../include/fermi_b200.h"

using namespace fmg;
extern std::atomic<uint64_t> g_launches;

namespace {

#define OB_TRY(call)                                                                                  \
    do {                                                                                              \
        cudaError_t err__ = (call);                                                                   \
        if (err__ != cudaSuccess) {                                                                   \
            if (fmg_verbose >= 1)                                                                     \
                std::fprintf(stderr, "[E::fmg_index_upload] %s failed: %s\n", #call, cudaGetErrorString(err__)); \
            return -1;                                                                                \
        }                                                                                             \
    } while (0)

struct Vec6 { uint64_t v[6]; };
struct Vec6Add { __host__ __device__ Vec6 operator()(const Vec6 &a, const Vec6 &b) const { Vec6 r; for (int i = 0; i < 6; ++i) r.v[i] = a.v[i] + b.v[i]; return r; } };

struct Dev {
    void *p = nullptr;
    ~Dev() { cudaFree(p); }
    cudaError_t alloc(size_t b) { cudaFree(p); p = nullptr; return cudaMalloc(&p, b ? b : 1); }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

__global__ void k_rld_block_sizes(const uint64_t *__restrict__ words, uint64_t n_blk, uint64_t *__restrict__ size) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blk) return;
    const uint64_t h = words[(b + 1) * 8];                      // the header of block b+1 counts the symbols of block b
    size[b] = ((uint32_t)h >> 31) ? ((uint32_t)h & 0x7fffffffu) : (h & 0xffffu);
}

// OR the bit range [pos, pos+len) into plane `pl`
__device__ __forceinline__ void fill_plane(uint32_t *blocks, int pl, uint64_t pos, uint64_t len) {
    while (len) {
        const uint64_t blk = pos >> 7, within = pos & 127;
        uint32_t *w = blocks + blk * 16 + 4 + pl * 4 + (within >> 5);
        const unsigned bit = within & 31;
        const uint64_t take = (32 - bit) < len ? (32 - bit) : len;
        if (take == 32) *w = 0xffffffffu;                       // a word wholly inside one run has a single writer
        else atomicOr(w, ((1u << take) - 1u) << bit);
        pos += take; len -= take;
    }
}

__global__ void __launch_bounds__(128) k_rld_decode(const uint64_t *__restrict__ words, uint64_t n_blk, const uint64_t *__restrict__ start,
                                                    const uint64_t *__restrict__ size, uint32_t *__restrict__ blocks) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blk) return;
    const uint64_t *w = words + b * 8;
    uint64_t bit = ((uint32_t)w[0] >> 31) ? 256 : 128;          // payload after a 4- or 2-word header (rld.c:76-77)
    uint64_t pos = start[b];
    const uint64_t end = pos + size[b];
    while (pos < end) {
        const uint64_t i = bit >> 6; const int s = bit & 63;
        const uint64_t x = s ? (w[i] << s) | (w[i + 1] >> (64 - s)) : w[i];     // a code never leaves its block, w[i+1] is at most the next header
        uint64_t len; int sym, used;
        if (x >> 63) { len = 1; sym = (x >> 60) & 7; used = 4; }
        else {
            const int z = __clzll(x);
            const int g = 2 * z + 1, y = (int)(x >> (64 - g)) - 1;
            len = ((x << g) >> (64 - y)) | (1ull << y);
            sym = (int)((x << (g + y)) >> 61);
            used = g + y + 3;
        }
        if (sym & 1) fill_plane(blocks, 0, pos, len);
        if (sym & 2) fill_plane(blocks, 1, pos, len);
        if (sym & 4) fill_plane(blocks, 2, pos, len);
        pos += len; bit += used;
    }
}

__global__ void k_occ_pad(uint32_t *blocks, uint64_t n_sym, uint64_t n_occ) {
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int pl = 0; pl < 3; ++pl) fill_plane(blocks, pl, n_sym, n_occ * 128 - n_sym);
}

__device__ __forceinline__ void half_counts(const uint32_t *w, int h, uint32_t cnt[6]) {
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        uint32_t n = 0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t p0 = w[4 + 2 * h + k], p1 = w[8 + 2 * h + k], p2 = w[12 + 2 * h + k];
            n += __popc(((c & 1) ? p0 : ~p0) & ((c & 2) ? p1 : ~p1) & ((c & 4) ? p2 : ~p2));
        }
        cnt[c] = n;
    }
}

__global__ void __launch_bounds__(256) k_occ_block_counts(const uint32_t *__restrict__ blocks, uint64_t n_occ, Vec6 *__restrict__ total) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_occ) return;
    uint32_t a[6], c[6];
    half_counts(blocks + b * 16, 0, a);
    half_counts(blocks + b * 16, 1, c);
    Vec6 t;
#pragma unroll
    for (int k = 0; k < 6; ++k) t.v[k] = a[k] + c[k];
    total[b] = t;
}

__global__ void __launch_bounds__(256) k_occ_finalize(uint32_t *__restrict__ blocks, uint64_t n_occ, const Vec6 *__restrict__ prefix,
                                                      uint64_t *__restrict__ cs, OccView ix) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_occ) return;
    uint32_t *w = blocks + b * 16;
    uint32_t a[6];
    half_counts(w, 0, a);
    const Vec6 here = prefix[b], base = prefix[b & ~((1ull << 17) - 1)];     // 2^17 occ blocks per superblock
    uint8_t *cb = reinterpret_cast<uint8_t *>(w);
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const uint32_t v = (uint32_t)(here.v[c] - base.v[c]) + a[c];
        cb[3 * c] = v & 0xff; cb[3 * c + 1] = (v >> 8) & 0xff; cb[3 * c + 2] = (v >> 16) & 0xff;
    }
    cb[15] = 0;
    if ((b & ((1ull << 17) - 1)) == 0) {
        uint64_t *row = cs + (b >> 17) * 8;
#pragma unroll
        for (int c = 0; c < 6; ++c) row[c] = ix.C[c] + here.v[c];
        row[6] = row[7] = 0;
    }
}

int build_on_host(const FmdImage &img, fmg_index_s *idx) {
    OccHost occ = build_occ_host(img);
    OB_TRY(cudaMemcpy(idx->d_blocks, occ.blocks.data(), occ.blocks.size() * 4, cudaMemcpyHostToDevice));
    OB_TRY(cudaMemcpy(idx->d_cs, occ.cs.data(), occ.cs.size() * 8, cudaMemcpyHostToDevice));
    return 0;
}

int build_on_device(const FmdImage &img, fmg_index_s *idx) {
    const uint64_t n_words = img.n_stream_words(), n_blk = img.n_blocks(), n_occ = idx->n_blocks;
    Dev d_words, d_size, d_start, d_total, d_prefix, d_tmp;
    OB_TRY(d_words.alloc((n_words + 2) * 8));
    OB_TRY(cudaMemcpy(d_words.p, img.words.data(), (n_words + 2) * 8, cudaMemcpyHostToDevice));
    OB_TRY(cudaMemset(idx->d_blocks, 0, n_occ * 64));
    if (n_blk) {
        OB_TRY(d_size.alloc(n_blk * 8)); OB_TRY(d_start.alloc(n_blk * 8));
        k_rld_block_sizes<<<(unsigned)((n_blk + 255) / 256), 256>>>(d_words.as<uint64_t>(), n_blk, d_size.as<uint64_t>()); ++g_launches;
        size_t need = 0;
        OB_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, d_size.as<uint64_t>(), d_start.as<uint64_t>(), (int64_t)n_blk));
        OB_TRY(d_tmp.alloc(need));
        OB_TRY(cub::DeviceScan::ExclusiveSum(d_tmp.p, need, d_size.as<uint64_t>(), d_start.as<uint64_t>(), (int64_t)n_blk));
        k_rld_decode<<<(unsigned)((n_blk + 127) / 128), 128>>>(d_words.as<uint64_t>(), n_blk, d_start.as<uint64_t>(), d_size.as<uint64_t>(), idx->d_blocks); ++g_launches;
    }
    k_occ_pad<<<1, 32>>>(idx->d_blocks, img.n_symbols(), n_occ); ++g_launches;
    OB_TRY(d_total.alloc(n_occ * sizeof(Vec6))); OB_TRY(d_prefix.alloc(n_occ * sizeof(Vec6)));
    k_occ_block_counts<<<(unsigned)((n_occ + 255) / 256), 256>>>(idx->d_blocks, n_occ, d_total.as<Vec6>()); ++g_launches;
    size_t need = 0;
    Vec6 zero{};
    OB_TRY(cub::DeviceScan::ExclusiveScan(nullptr, need, d_total.as<Vec6>(), d_prefix.as<Vec6>(), Vec6Add(), zero, (int64_t)n_occ));
    OB_TRY(d_tmp.alloc(need));
    OB_TRY(cub::DeviceScan::ExclusiveScan(d_tmp.p, need, d_total.as<Vec6>(), d_prefix.as<Vec6>(), Vec6Add(), zero, (int64_t)n_occ));
    OccView v;
    for (int c = 0; c < 8; ++c) v.C[c] = img.cnt[c];
    k_occ_finalize<<<(unsigned)((n_occ + 255) / 256), 256>>>(idx->d_blocks, n_occ, d_prefix.as<Vec6>(), idx->d_cs, v); ++g_launches;
    OB_TRY(cudaGetLastError());
    OB_TRY(cudaDeviceSynchronize());
    return 0;
}

} // namespace

int occ_build_device(const FmdImage &img, fmg_index_s *idx) {
    idx->n_blocks = occ_n_blocks(img.n_symbols());
    const uint64_t n_super = occ_n_super(img.n_symbols());
    idx->bytes = idx->n_blocks * 64 + n_super * 64;
    cudaError_t err = cudaMalloc(&idx->d_blocks, idx->n_blocks * 64);
    if (err == cudaSuccess) err = cudaMalloc(&idx->d_cs, n_super * 64);
    int rc = -1;
    if (err == cudaSuccess) rc = std::getenv("FMG_HOST_OCC_BUILD") ? build_on_host(img, idx) : build_on_device(img, idx);
    else if (fmg_verbose >= 1) std::fprintf(stderr, "[E::fmg_index_upload] %s\n", cudaGetErrorString(err));
    if (rc != 0) {
        cudaFree(idx->d_blocks); cudaFree(idx->d_cs);
        idx->d_blocks = nullptr; idx->d_cs = nullptr;
    }
    return rc;
}

// copy the query layout back to the host (tests, serialisation): blocks = n_blocks*16 u32, cs = n_super*8 u64
extern "C" int fmg_index_export(const fmg_index_t *idx, uint32_t *blocks, uint64_t *cs, uint64_t *n_blocks, uint64_t *n_super) {
    if (!idx) return -1;
    const uint64_t ns = occ_n_super(idx->mcnt[0]);
    if (n_blocks) *n_blocks = idx->n_blocks;
    if (n_super) *n_super = ns;
    if (cudaSetDevice(idx->device) != cudaSuccess) return -1;
    if (blocks && cudaMemcpy(blocks, idx->d_blocks, idx->n_blocks * 64, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (cs && cudaMemcpy(cs, idx->d_cs, ns * 64, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return 0;
}
