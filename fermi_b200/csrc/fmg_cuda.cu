// libfermi_b200: sm_100a kernels and the C-ABI around them (see include/fermi_b200.h).
//
// Kernels (all integer / bit work, HBM-latency and -bandwidth bound; no tensor cores):
//   k_rank2a            batched rld_rank2a           (rld.c:457-492)
//   k_extend            batched fm6_extend           (exact.c:72-88)
//   k_backward_search   batched fm_backward_search   (exact.c:7-23)
//   k_smem              fm6_smem over fm6_smem1_core (smem.c:13-80,397-410), persistent lanes,
//                       one read per lane, one extension per loop trip (fmd_device.cuh)
//   k_compact_*         per-read record slots -> dense record array + offsets
// There is no host execution path: every entry point needs a CUDA device and fails loudly otherwise.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <vector>
#include <algorithm>
#include <mutex>
#include "fmd_device.cuh"
#include "occ_layout.hpp"
#include "fmg_internal.hpp"
#include "../../include/fermi_b200.h"

using namespace fmg;

std::atomic<uint64_t> g_launches{0};

#define CUDA_TRY(call, fail)                                                                         \
    do {                                                                                             \
        cudaError_t err__ = (call);                                                                  \
        if (err__ != cudaSuccess) {                                                                  \
            if (fmg_verbose >= 1)                                                                    \
                std::fprintf(stderr, "[E::%s] %s failed: %s\n", __func__, #call, cudaGetErrorString(err__)); \
            fail;                                                                                    \
        }                                                                                            \
    } while (0)

#define LAUNCH_CHECK(fail)                                                                           \
    do {                                                                                             \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                          \
        cudaError_t err__ = cudaGetLastError();                                                      \
        if (err__ != cudaSuccess) {                                                                  \
            if (fmg_verbose >= 1)                                                                    \
                std::fprintf(stderr, "[E::%s] kernel launch failed: %s\n", __func__, cudaGetErrorString(err__)); \
            fail;                                                                                    \
        }                                                                                            \
    } while (0)

// ------------------------------------------------------------------------------------ kernels

__global__ void __launch_bounds__(256) k_rank2a(OccView ix, int64_t n, const uint64_t *__restrict__ k,
                                                const uint64_t *__restrict__ l, uint64_t *__restrict__ ok,
                                                uint64_t *__restrict__ ol) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t pk = k[i] + 1, pl = l[i] + 1;          // k == (uint64_t)-1 -> counts of the empty prefix
    const Blk bk = load_blk(ix, pk), bl = load_blk(ix, pl);
    uint32_t a[6], b[6];
    rank_rel(bk, pk, a);
    rank_rel(bl, pl, b);
    const uint64_t *ck = ix.cs + (pk >> kSuperShift) * 8, *cl = ix.cs + (pl >> kSuperShift) * 8;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        ok[6 * i + c] = ck[c] - ix.C[c] + a[c];
        ol[6 * i + c] = cl[c] - ix.C[c] + b[c];
    }
}

// rld_rank1a (rld.c:424-446): ok[c] = #c in BWT[0..k] and the symbol BWT[k]; k == -1 -> zeros and -1.  ks == nullptr: k = first + i
__global__ void __launch_bounds__(256) k_rank1a(OccView ix, int64_t n, const uint64_t *__restrict__ ks, uint64_t first, uint64_t *__restrict__ ok,
                                                int32_t *__restrict__ sym) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = ks ? ks[i] : first + (uint64_t)i;
    const uint64_t p = k + 1;                              // counts of BWT[0, k]
    const Blk b = load_blk(ix, p);
    uint32_t a[6];
    rank_rel(b, p, a);
    const uint64_t *ck = ix.cs + (p >> kSuperShift) * 8;
#pragma unroll
    for (int c = 0; c < 6; ++c) ok[6 * i + c] = ck[c] - ix.C[c] + a[c];
    if (sym) sym[i] = k == ~0ull ? -1 : blk_symbol((k >> kBlkShift) == (p >> kBlkShift) ? b : load_blk(ix, k), k);
}

// the check of `fermi chkbwt -r` (cmd.c:90-105) for every position at once: rank1a(k)[c] must equal the number of c among
// BWT[0..k] counted directly.  The direct count is a scan of the symbols; here it is checked incrementally, which is equivalent:
// rank(k) - rank(k-1) must be exactly the one-hot vector of BWT[k] (rank(-1) = 0), and rank(n-1) must equal the marginal counts.
__global__ void __launch_bounds__(256) k_chk_rank(OccView ix, uint64_t n, unsigned long long *bad, unsigned long long *first_bad) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t a[6], b[6];
    const uint64_t p1 = k + 1;
    const Blk B1 = load_blk(ix, p1);
    rank_rel(B1, p1, a);
    const Blk B0 = (k >> kBlkShift) == (p1 >> kBlkShift) ? B1 : load_blk(ix, k);
    rank_rel(B0, k, b);
    const int c0 = blk_symbol(B0, k);
    const uint64_t *c1 = ix.cs + (p1 >> kSuperShift) * 8, *cz = ix.cs + (k >> kSuperShift) * 8;
    bool ok = c0 >= 0 && c0 < 6;
#pragma unroll
    for (int c = 0; c < 6; ++c) ok = ok && (c1[c] + a[c]) - (cz[c] + b[c]) == (uint64_t)(c == c0);
    if (!ok) { atomicAdd(bad, 1ull); atomicMin(first_bad, (unsigned long long)k); }
}

__global__ void __launch_bounds__(256) k_extend(OccView ix, int64_t n, const uint4 *__restrict__ ik,
                                                const uint8_t *__restrict__ is_back, uint4 *__restrict__ ok6) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Intv p = ld_intv(ik + 2 * i);
    const int b = is_back[i] != 0;
    Ext6 e;
    extend6<uint64_t>(ix, b ? p.x1 : p.x0, b ? p.x0 : p.x1, p.x2, e);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        Intv o;
        const uint64_t fr = far_of(ix, e, c);
        o.x0 = b ? fr : e.near[c]; o.x1 = b ? e.near[c] : fr; o.x2 = e.size[c]; o.info = 0;
        st_intv(ok6 + 2 * (6 * i + c), o);
    }
}

__global__ void __launch_bounds__(256) k_backward_search(OccView ix, int64_t n, const uint8_t *__restrict__ seq,
                                                         const uint64_t *__restrict__ off, uint64_t *__restrict__ sa_beg,
                                                         uint64_t *__restrict__ sa_end, uint64_t *__restrict__ size) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint8_t *q = seq + off[r];
    const int len = (int)(off[r + 1] - off[r]);
    uint64_t beg = 0, end = 0, sz = 0;
    if (len > 0) {
        int c = q[len - 1];
        uint64_t k = ix.C[c], l = ix.C[c + 1] - 1;        // closed interval [k,l]
        int i;
        for (i = len - 2; i >= 0; --i) {
            c = q[i];
            const uint64_t pk = k, pl = l + 1;             // rank11(k-1), rank11(l)
            const Blk bk = load_blk(ix, pk), bl = load_blk(ix, pl);
            uint32_t a[6], b[6];
            rank_rel(bk, pk, a);
            rank_rel(bl, pl, b);
            k = ix.cs[(pk >> kSuperShift) * 8 + c] + pick6(a, c);
            l = ix.cs[(pl >> kSuperShift) * 8 + c] + pick6(b, c) - 1;
            if (k > l) break;
        }
        if (!(k > l)) beg = k, end = l, sz = l - k + 1;
    }
    sa_beg[r] = beg; sa_end[r] = end; size[r] = sz;
}

// U = coordinate type: k_smem<uint32_t> for indexes of < 2^32 symbols, k_smem<uint64_t> beyond
// PAIR = paired block gathers (fmd_device.cuh: load_blk_pair), chosen for indexes that do not fit L2
template <typename U, bool PAIR>
__global__ void __launch_bounds__(SMEM_BLOCK, sizeof(U) == 4 ? SMEM_MIN_BLOCKS_U32 : SMEM_MIN_BLOCKS_U64) k_smem(SmemArgs A) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    smem_lane<U, PAIR>(A, slot, [&]() -> int64_t { return (int64_t)atomicAdd(A.next_read, 1ull); });
}
static const void *smem_kernel(bool wide, bool pair) {
    if (wide) return pair ? (const void *)k_smem<uint64_t, true> : (const void *)k_smem<uint64_t, false>;
    return pair ? (const void *)k_smem<uint32_t, true> : (const void *)k_smem<uint32_t, false>;
}

// ---- compaction of the per-read record slots -------------------------------------------------
constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;                 // reads per thread
constexpr int kScanTile = kScanBlock * kScanItems;

__device__ __forceinline__ uint32_t clamp_cnt(uint32_t c, int cap) { return c > (uint32_t)cap ? (uint32_t)cap : c; }

__global__ void __launch_bounds__(kScanBlock) k_compact_tile_sums(const uint32_t *__restrict__ cnt, int64_t n, int cap,
                                                                  uint64_t *__restrict__ tile_sum, unsigned long long *overflow) {
    __shared__ uint32_t wsum[kScanBlock / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    uint32_t s = 0, ov = 0;
    for (int t = 0; t < kScanItems; ++t) {
        const int64_t i = base + t * kScanBlock + threadIdx.x;
        if (i < n) { const uint32_t c = cnt[i]; s += clamp_cnt(c, cap); ov += c > (uint32_t)cap; }
    }
    if (ov) atomicAdd(overflow, (unsigned long long)ov);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int w = 0; w < kScanBlock / 32; ++w) t += wsum[w];
        tile_sum[blockIdx.x] = t;
    }
}

// one block: exclusive scan of the tile sums (n_tiles is small: n_reads / 2048)
__global__ void __launch_bounds__(1024) k_compact_scan_tiles(uint64_t *tile_sum, int64_t n_tiles, uint64_t *total) {
    __shared__ uint64_t sh[1024];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_tiles; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const uint64_t v = i < n_tiles ? tile_sum[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const uint64_t a = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += a;
            __syncthreads();
        }
        if (i < n_tiles) tile_sum[i] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// exclusive offsets per read + gather of the records
__global__ void __launch_bounds__(kScanBlock) k_compact_gather(const uint32_t *__restrict__ cnt, int64_t n, int cap,
                                                               const uint64_t *__restrict__ tile_sum,
                                                               const uint4 *__restrict__ slots, uint4 *__restrict__ mem,
                                                               uint64_t *__restrict__ mem_off) {
    __shared__ uint32_t wsum[kScanBlock / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t c[kScanItems], s = 0;
#pragma unroll
    for (int t = 0; t < kScanItems; ++t) {
        c[t] = base + t < n ? clamp_cnt(cnt[base + t], cap) : 0;
        s += c[t];
    }
    uint32_t incl = s;                                     // warp inclusive scan of the per-thread sums
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += v;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += wsum[w];
    uint64_t o = tile_sum[blockIdx.x] + wbase + incl - s;
#pragma unroll
    for (int t = 0; t < kScanItems; ++t) {
        const int64_t r = base + t;
        if (r < n) {
            mem_off[r] = o;
            const uint4 *src = slots + (size_t)r * cap * 2;
            uint4 *dst = mem + o * 2;
            for (uint32_t k = 0; k < 2 * c[t]; ++k) dst[k] = src[k];
            o += c[t];
        }
    }
    if (base <= n && n < base + kScanItems) mem_off[n] = o;     // the thread owning index n writes the end offset
}

int64_t fmg_compact_tiles(int64_t n) { return (n + 1 + kScanTile - 1) / kScanTile; }

// slots[n][cap] (32-byte records, cnt[i] used) -> dense mem + mem_off[n+1]; ctrl[1] += overflowing rows, ctrl[2] = total
int fmg_compact_slots(const uint32_t *cnt, int64_t n, int cap, const uint4 *slots, uint4 *mem, uint64_t *mem_off, uint64_t *tile_sum,
                      unsigned long long *ctrl, cudaStream_t st) {
    const int64_t n_tiles = fmg_compact_tiles(n);
    k_compact_tile_sums<<<(unsigned)n_tiles, kScanBlock, 0, st>>>(cnt, n, cap, tile_sum, ctrl + 1);
    LAUNCH_CHECK(return -1);
    k_compact_scan_tiles<<<1, 1024, 0, st>>>(tile_sum, n_tiles, (uint64_t *)(ctrl + 2));
    LAUNCH_CHECK(return -1);
    k_compact_gather<<<(unsigned)n_tiles, kScanBlock, 0, st>>>(cnt, n, cap, tile_sum, slots, mem, mem_off);
    LAUNCH_CHECK(return -1);
    return 0;
}

// ------------------------------------------------------------------------------------ index

static int use_device(int device, const char *who) {
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess || n == 0) {
        if (fmg_verbose >= 1)
            std::fprintf(stderr, "[E::%s] no CUDA device available (%s); libfermi_b200 has no CPU path\n", who,
                         err == cudaSuccess ? "device count is 0" : cudaGetErrorString(err));
        return -1;
    }
    if (device < 0 || device >= n) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] device %d out of range (have %d)\n", who, device, n);
        return -1;
    }
    err = cudaSetDevice(device);
    if (err != cudaSuccess) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] cudaSetDevice(%d): %s\n", who, device, cudaGetErrorString(err));
        return -1;
    }
    return 0;
}

extern "C" {

fmg_index_t *fmg_index_upload(const fmg_fmd_t *e, int device) {
    if (!e || use_device(device, __func__)) return nullptr;
    fmg_index_t *idx = new fmg_index_s;
    idx->device = device;
    const FmdImage &img = e->img;
    std::memcpy(idx->mcnt, img.mcnt, sizeof(idx->mcnt));
    std::memcpy(idx->cnt, img.cnt, sizeof(idx->cnt));
    if (occ_build_device(img, idx) != 0) { delete idx; return nullptr; }
    OccView &v = idx->view;
    v.blocks = idx->d_blocks; v.cs = idx->d_cs;
    v.n_sym = img.mcnt[0]; v.n_seq = img.mcnt[1];
    for (int c = 0; c < 8; ++c) v.C[c] = img.cnt[c];
    cudaDeviceGetAttribute(&idx->n_sm, cudaDevAttrMultiProcessorCount, device);
    if (fmg_verbose >= 3)
        std::fprintf(stderr, "[M::%s] %llu symbols, %llu sequences -> %.1f MB of occ blocks on device %d (%d SMs)\n", __func__,
                     (unsigned long long)v.n_sym, (unsigned long long)v.n_seq, idx->bytes / 1e6, device, idx->n_sm);
    return idx;
}

void fmg_index_free(fmg_index_t *idx) {
    if (!idx) return;
    cudaSetDevice(idx->device);
    fmg_pipe_destroy(idx->pipe);
    fmg_ovcache_destroy(idx->ovc);
    cudaFree(idx->d_blocks);
    cudaFree(idx->d_cs);
    delete idx;
}

uint64_t fmg_index_bytes(const fmg_index_t *idx) { return idx->bytes; }
int fmg_index_device(const fmg_index_t *idx) { return idx->device; }
uint64_t fmg_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------ simple batches

int fmg_rank2a_batch(const fmg_index_t *idx, int64_t n, const uint64_t *k, const uint64_t *l, uint64_t *ok, uint64_t *ol) {
    if (!idx || use_device(idx->device, __func__)) return -1;
    if (n <= 0) return 0;
    uint64_t *dk = nullptr, *dl = nullptr, *dok = nullptr, *dol = nullptr;
    int rc = -1;
    do {
        CUDA_TRY(cudaMalloc(&dk, n * 8), break);
        CUDA_TRY(cudaMalloc(&dl, n * 8), break);
        CUDA_TRY(cudaMalloc(&dok, n * 48), break);
        CUDA_TRY(cudaMalloc(&dol, n * 48), break);
        CUDA_TRY(cudaMemcpy(dk, k, n * 8, cudaMemcpyHostToDevice), break);
        CUDA_TRY(cudaMemcpy(dl, l, n * 8, cudaMemcpyHostToDevice), break);
        k_rank2a<<<(unsigned)((n + 255) / 256), 256>>>(idx->view, n, dk, dl, dok, dol);
        LAUNCH_CHECK(break);
        CUDA_TRY(cudaMemcpy(ok, dok, n * 48, cudaMemcpyDeviceToHost), break);
        CUDA_TRY(cudaMemcpy(ol, dol, n * 48, cudaMemcpyDeviceToHost), break);
        rc = 0;
    } while (0);
    cudaFree(dk); cudaFree(dl); cudaFree(dok); cudaFree(dol);
    return rc;
}

int fmg_rank1a_batch(const fmg_index_t *idx, int64_t n, const uint64_t *k, uint64_t *ok, int32_t *sym) {
    if (!idx || !k || !ok || use_device(idx->device, __func__)) return -1;
    if (n <= 0) return 0;
    uint64_t *dk = nullptr, *dok = nullptr;
    int32_t *ds = nullptr;
    int rc = -1;
    do {
        CUDA_TRY(cudaMalloc(&dk, n * 8), break);
        CUDA_TRY(cudaMalloc(&dok, n * 48), break);
        CUDA_TRY(cudaMalloc(&ds, n * 4), break);
        CUDA_TRY(cudaMemcpy(dk, k, n * 8, cudaMemcpyHostToDevice), break);
        k_rank1a<<<(unsigned)((n + 255) / 256), 256>>>(idx->view, n, dk, 0, dok, ds);
        LAUNCH_CHECK(break);
        CUDA_TRY(cudaMemcpy(ok, dok, n * 48, cudaMemcpyDeviceToHost), break);
        if (sym) CUDA_TRY(cudaMemcpy(sym, ds, n * 4, cudaMemcpyDeviceToHost), break);
        rc = 0;
    } while (0);
    cudaFree(dk); cudaFree(dok); cudaFree(ds);
    return rc;
}

int fmg_check_rank(const fmg_index_t *idx, uint64_t *n_bad, uint64_t *first_bad) {
    if (!idx || use_device(idx->device, __func__)) return -1;
    const uint64_t n = idx->view.n_sym;
    unsigned long long *d = nullptr, h[2] = {0, ~0ull};
    int rc = -1;
    do {
        CUDA_TRY(cudaMalloc(&d, 16), break);
        CUDA_TRY(cudaMemcpy(d, h, 16, cudaMemcpyHostToDevice), break);
        if (n) {
            k_chk_rank<<<(unsigned)((n + 255) / 256), 256>>>(idx->view, n, d, d + 1);
            LAUNCH_CHECK(break);
        }
        CUDA_TRY(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost), break);
        // the last position must reproduce the marginal counts (cmd.c:108-116)
        uint64_t last = n - 1, ok[6];
        if (n && fmg_rank1a_batch(idx, 1, &last, ok, nullptr) == 0)
            for (int c = 0; c < 6; ++c) if (ok[c] != idx->mcnt[c + 1]) { ++h[0]; if (h[1] == ~0ull) h[1] = last; }
        rc = 0;
    } while (0);
    cudaFree(d);
    if (n_bad) *n_bad = h[0];
    if (first_bad) *first_bad = h[1];
    return rc;
}

int fmg_extend_batch(const fmg_index_t *idx, int64_t n, const fmg_intv_t *ik, const uint8_t *is_back, fmg_intv_t *ok6) {
    if (!idx || use_device(idx->device, __func__)) return -1;
    if (n <= 0) return 0;
    uint4 *din = nullptr, *dout = nullptr;
    uint8_t *db = nullptr;
    int rc = -1;
    do {
        CUDA_TRY(cudaMalloc(&din, n * 32), break);
        CUDA_TRY(cudaMalloc(&dout, n * 192), break);
        CUDA_TRY(cudaMalloc(&db, n), break);
        CUDA_TRY(cudaMemcpy(din, ik, n * 32, cudaMemcpyHostToDevice), break);
        CUDA_TRY(cudaMemcpy(db, is_back, n, cudaMemcpyHostToDevice), break);
        k_extend<<<(unsigned)((n + 255) / 256), 256>>>(idx->view, n, din, db, dout);
        LAUNCH_CHECK(break);
        CUDA_TRY(cudaMemcpy(ok6, dout, n * 192, cudaMemcpyDeviceToHost), break);
        rc = 0;
    } while (0);
    cudaFree(din); cudaFree(dout); cudaFree(db);
    return rc;
}

int fmg_backward_search_batch(const fmg_index_t *idx, int64_t n, const uint8_t *seq, const uint64_t *off,
                              uint64_t *sa_beg, uint64_t *sa_end, uint64_t *size) {
    if (!idx || use_device(idx->device, __func__)) return -1;
    if (n <= 0) return 0;
    uint8_t *dseq = nullptr;
    uint64_t *doff = nullptr, *dres = nullptr;
    const uint64_t nb = off[n];
    int rc = -1;
    do {
        CUDA_TRY(cudaMalloc(&dseq, nb ? nb : 1), break);
        CUDA_TRY(cudaMalloc(&doff, (n + 1) * 8), break);
        CUDA_TRY(cudaMalloc(&dres, n * 24), break);
        CUDA_TRY(cudaMemcpy(dseq, seq, nb, cudaMemcpyHostToDevice), break);
        CUDA_TRY(cudaMemcpy(doff, off, (n + 1) * 8, cudaMemcpyHostToDevice), break);
        k_backward_search<<<(unsigned)((n + 255) / 256), 256>>>(idx->view, n, dseq, doff, dres, dres + n, dres + 2 * n);
        LAUNCH_CHECK(break);
        CUDA_TRY(cudaMemcpy(sa_beg, dres, n * 8, cudaMemcpyDeviceToHost), break);
        CUDA_TRY(cudaMemcpy(sa_end, dres + n, n * 8, cudaMemcpyDeviceToHost), break);
        CUDA_TRY(cudaMemcpy(size, dres + 2 * n, n * 8, cudaMemcpyDeviceToHost), break);
        rc = 0;
    } while (0);
    cudaFree(dseq); cudaFree(doff); cudaFree(dres);
    return rc;
}

// ------------------------------------------------------------------------------------ SMEM session

struct fmg_smem_session_s {
    const fmg_index_s *idx = nullptr;
    int64_t max_reads = 0;
    int max_len = 0, cap = 0, out_cap = 0;
    int grid = 0, n_lanes = 0;
    void *F = nullptr, *W = nullptr;
    uint4 *slots = nullptr, *mem = nullptr;
    bool wide = true;                          // 64-bit coordinates (index of >= 2^32 symbols)
    bool pair = false;                         // paired block gathers (index larger than L2)
    uint32_t *rec_cnt = nullptr;
    uint64_t *mem_off = nullptr, *tile_sum = nullptr;
    unsigned long long *ctrl = nullptr;        // [0] next read, [1] overflow count, [2] total records
    unsigned long long *h_ctrl = nullptr;      // pinned mirror of ctrl
    int64_t last_n = 0;
    // arguments of the last run (kept to re-run after a slot overflow)
    const uint8_t *last_seq = nullptr;
    const uint64_t *last_off = nullptr;
    int last_self = 0;
    cudaStream_t last_stream = nullptr;
    // optional CUDA-event timing of k_smem alone (bench.py's roofline figure)
    bool timing = false;
    std::vector<cudaEvent_t> ev;               // pairs (start, stop), one per k_smem launch since the last query
    std::vector<cudaEvent_t> ev_free;
};

static void session_free_slots(fmg_smem_session_t *s) {
    cudaFree(s->slots); cudaFree(s->mem);
    s->slots = s->mem = nullptr;
}

static int session_alloc_slots(fmg_smem_session_t *s, int out_cap) {
    session_free_slots(s);
    s->out_cap = out_cap;
    const size_t bytes = (size_t)s->max_reads * out_cap * 32;
    CUDA_TRY(cudaMalloc(&s->slots, bytes), return -1);
    CUDA_TRY(cudaMalloc(&s->mem, bytes), return -1);
    return 0;
}

fmg_smem_session_t *fmg_smem_session_create(const fmg_index_t *idx, int64_t max_reads, int max_len) {
    if (!idx || max_reads <= 0 || max_len <= 0 || use_device(idx->device, __func__)) return nullptr;
    fmg_smem_session_t *s = new fmg_smem_session_s;
    s->idx = idx; s->max_reads = max_reads; s->max_len = max_len;
    s->cap = 2 * max_len + 2;
    s->wide = idx->view.n_sym + 256 >= (1ull << 32) || std::getenv("FMG_FORCE_WIDE") != nullptr;
    // paired gathers when the index does not fit L2 (126 MB on B200): there the L1-miss request rate bounds the kernel
    s->pair = idx->bytes > ((size_t)112 << 20);
    if (const char *e = std::getenv("FMG_SMEM_PAIR")) s->pair = std::atoi(e) != 0;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smem_kernel(s->wide, s->pair), SMEM_BLOCK, 0);
    if (per_sm < 1) per_sm = 1;
    s->grid = idx->n_sm * per_sm;
    // long queries (contigs, unitigs: fm6_remap, `exact` on assemblies): every lane needs 2 x (2 len + 2) candidate slots, so
    // the lane count shrinks until the two lists fit a fixed share of HBM; the kernel hands out reads to however many lanes run
    {
        const size_t per_lane = (size_t)2 * s->cap * (s->wide ? 32 : 16), budget = (size_t)24 << 30;
        const int64_t max_blocks = (int64_t)(budget / (per_lane * SMEM_BLOCK));
        if (max_blocks < 1) {
            if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] queries of %d bases need more candidate scratch than fits\n", __func__, max_len);
            delete s;
            return nullptr;
        }
        if (s->grid > max_blocks) s->grid = (int)max_blocks;
    }
    s->n_lanes = s->grid * SMEM_BLOCK;
    const int64_t n_tiles = (max_reads + 1 + kScanTile - 1) / kScanTile;
    bool ok = false;
    do {
        CUDA_TRY(cudaMalloc(&s->F, (size_t)s->n_lanes * s->cap * (s->wide ? 32 : 16)), break);
        CUDA_TRY(cudaMalloc(&s->W, (size_t)s->n_lanes * s->cap * (s->wide ? 32 : 16)), break);
        CUDA_TRY(cudaMalloc(&s->rec_cnt, (size_t)(max_reads + 1) * 4), break);
        CUDA_TRY(cudaMalloc(&s->mem_off, (size_t)(max_reads + 1) * 8), break);
        CUDA_TRY(cudaMalloc(&s->tile_sum, (size_t)n_tiles * 8), break);
        CUDA_TRY(cudaMalloc(&s->ctrl, 4 * sizeof(unsigned long long)), break);
        CUDA_TRY(cudaMallocHost(&s->h_ctrl, 4 * sizeof(unsigned long long)), break);
        if (session_alloc_slots(s, SMEM_DEFAULT_OUT_CAP)) break;
        ok = true;
    } while (0);
    if (!ok) { fmg_smem_session_destroy(s); return nullptr; }
    if (fmg_verbose >= 4)
        std::fprintf(stderr, "[M::%s] %d blocks x %d lanes (%d blocks/SM), %d-bit coordinates, %d candidate slots/lane, %d record slots/read\n",
                     __func__, s->grid, SMEM_BLOCK, per_sm, s->wide ? 64 : 32, s->cap, s->out_cap);
    return s;
}

void fmg_smem_session_destroy(fmg_smem_session_t *s) {
    if (!s) return;
    cudaSetDevice(s->idx->device);
    session_free_slots(s);
    cudaFree(s->F); cudaFree(s->W); cudaFree(s->rec_cnt); cudaFree(s->mem_off); cudaFree(s->tile_sum); cudaFree(s->ctrl);
    cudaFreeHost(s->h_ctrl);
    for (cudaEvent_t e : s->ev) cudaEventDestroy(e);
    for (cudaEvent_t e : s->ev_free) cudaEventDestroy(e);
    delete s;
}

static int session_enqueue(fmg_smem_session_t *s, int64_t n, const uint8_t *d_seq, const uint64_t *d_off, int self_match,
                           cudaStream_t st) {
    CUDA_TRY(cudaMemsetAsync(s->ctrl, 0, 4 * sizeof(unsigned long long), st), return -1);
    SmemArgs A;
    A.ix = s->idx->view; A.seq = d_seq; A.off = d_off; A.n_reads = n; A.self_match = self_match;
    A.F = s->F; A.W = s->W; A.cap = s->cap; A.out = s->slots; A.out_cap = s->out_cap;
    A.rec_cnt = s->rec_cnt; A.next_read = s->ctrl; A.max_len = s->max_len; A.too_long = s->ctrl + 3;
    const int64_t need_blocks = (n + SMEM_BLOCK - 1) / SMEM_BLOCK;
    const int grid = (int)std::min<int64_t>(s->grid, need_blocks);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (s->timing) {
        for (cudaEvent_t *e : {&e0, &e1}) {
            if (!s->ev_free.empty()) { *e = s->ev_free.back(); s->ev_free.pop_back(); }
            else CUDA_TRY(cudaEventCreate(e), return -1);
        }
        CUDA_TRY(cudaEventRecord(e0, st), return -1);
    }
    {
        void *kargs[] = {(void *)&A};
        CUDA_TRY(cudaLaunchKernel(smem_kernel(s->wide, s->pair), dim3((unsigned)grid), dim3(SMEM_BLOCK), kargs, 0, st), return -1);
        ++g_launches;
    }
    if (s->timing) {
        CUDA_TRY(cudaEventRecord(e1, st), return -1);
        s->ev.push_back(e0); s->ev.push_back(e1);
    }
    if (fmg_compact_slots(s->rec_cnt, n, s->out_cap, s->slots, s->mem, s->mem_off, s->tile_sum, s->ctrl, st)) return -1;
    CUDA_TRY(cudaMemcpyAsync(s->h_ctrl, s->ctrl, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st), return -1);
    return 0;
}

int fmg_smem_session_run(fmg_smem_session_t *s, int64_t n, const uint8_t *d_seq, const uint64_t *d_off, int self_match,
                         void *stream) {
    if (!s || n < 0 || n > s->max_reads) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] bad session or batch size\n", __func__);
        return -1;
    }
    if (use_device(s->idx->device, __func__)) return -1;
    s->last_n = n; s->last_seq = d_seq; s->last_off = d_off; s->last_self = self_match; s->last_stream = (cudaStream_t)stream;
    if (n == 0) { s->h_ctrl[1] = s->h_ctrl[2] = 0; return 0; }
    return session_enqueue(s, n, d_seq, d_off, self_match, (cudaStream_t)stream);
}

int fmg_smem_session_result(fmg_smem_session_t *s, uint64_t *n_records, const fmg_intv_t **d_mem, const uint64_t **d_mem_off) {
    if (!s) return -1;
    if (use_device(s->idx->device, __func__)) return -1;
    CUDA_TRY(cudaStreamSynchronize(s->last_stream), return -1);
    if (s->last_n > 0 && s->h_ctrl[3] != 0) {
        if (fmg_verbose >= 1)
            std::fprintf(stderr, "[E::%s] %llu reads are longer than the %d bases this session was created for; they were skipped\n", __func__, s->h_ctrl[3], s->max_len);
        return -2;
    }
    while (s->last_n > 0 && s->h_ctrl[1] != 0) {
        // some read produced more records than its slot holds: grow the slots and run the batch again
        const int bigger = s->out_cap * 4;
        if (fmg_verbose >= 3)
            std::fprintf(stderr, "[M::%s] %llu reads overflowed %d record slots; re-running with %d\n", __func__,
                         s->h_ctrl[1], s->out_cap, bigger);
        if (session_alloc_slots(s, bigger)) return -1;
        if (session_enqueue(s, s->last_n, s->last_seq, s->last_off, s->last_self, s->last_stream)) return -1;
        CUDA_TRY(cudaStreamSynchronize(s->last_stream), return -1);
    }
    if (n_records) *n_records = s->last_n > 0 ? s->h_ctrl[2] : 0;
    if (d_mem) *d_mem = reinterpret_cast<const fmg_intv_t *>(s->mem);
    if (d_mem_off) *d_mem_off = s->mem_off;
    return 0;
}

void fmg_smem_session_set_timing(fmg_smem_session_t *s, int on) { if (s) s->timing = on != 0; }

// sum of the k_smem durations (CUDA events on the launching stream) since the last call; *n_launches = how many
double fmg_smem_session_kernel_ms(fmg_smem_session_t *s, int *n_launches) {
    double total = 0;
    if (n_launches) *n_launches = 0;
    if (!s) return 0;
    for (size_t i = 0; i + 1 < s->ev.size(); i += 2) {
        float ms = 0;
        if (cudaEventSynchronize(s->ev[i + 1]) == cudaSuccess && cudaEventElapsedTime(&ms, s->ev[i], s->ev[i + 1]) == cudaSuccess) {
            total += ms;
            if (n_launches) ++*n_launches;
        }
        s->ev_free.push_back(s->ev[i]); s->ev_free.push_back(s->ev[i + 1]);
    }
    s->ev.clear();
    return total;
}

// ------------------------------------------------------------------------------------ SMEM, host buffers
// Reads are cut into batches; three streams (H2D, compute, D2H) and two sets of device buffers overlap
// the copies of batch b+1 / b-1 with the kernels of batch b.

int fmg_smem_batch_into(const fmg_index_t *idx, int64_t n, const uint8_t *seq, const uint64_t *off, int self_match,
                        fmg_intv_t *mem, uint64_t mem_cap, uint64_t *mem_off, uint64_t *n_records, int64_t batch_reads);

int fmg_smem_batch(const fmg_index_t *idx, int64_t n, const uint8_t *seq, const uint64_t *off, int self_match,
                   fmg_intv_t **mem, uint64_t *mem_off) {
    if (!idx || !mem || !mem_off) return -1;
    // two passes are avoided by sizing from an estimate and growing on demand
    uint64_t cap = (uint64_t)std::max<int64_t>(n, 1) * 16, got = 0;
    for (;;) {
        fmg_intv_t *buf = (fmg_intv_t *)std::malloc(cap * sizeof(fmg_intv_t));
        if (!buf) { if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] out of host memory\n", __func__); return -1; }
        const int rc = fmg_smem_batch_into(idx, n, seq, off, self_match, buf, cap, mem_off, &got, 0);
        if (rc == 0) { *mem = buf; return 0; }
        std::free(buf);
        if (rc != 1) return rc;                 // 1 = capacity too small, `got` holds the need
        cap = got;
    }
}

struct BatchBuf {
    uint8_t *d_seq = nullptr;
    uint64_t *d_off = nullptr;
    fmg_smem_session_t *sess = nullptr;
    cudaEvent_t h2d_done = nullptr, run_done = nullptr, d2h_done = nullptr;
    int64_t first = 0, count = 0;
    uint64_t rec_base = 0, n_rec = 0;
};

} // extern "C"

// Streams, events, device input buffers and the two sessions of the host-buffer pipeline.  They are kept with the
// index between calls (allocating ~10 GB of slots per call costs more than a batch of kernels).
struct fmg_pipe_s {
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    BatchBuf buf[2];
    int64_t batch_reads = 0;
    int max_len = 0;
    uint64_t max_bytes = 0;
};

void fmg_pipe_destroy(fmg_pipe_s *p) {
    if (!p) return;
    for (int k = 0; k < 2; ++k) {
        if (p->buf[k].sess) fmg_smem_session_destroy(p->buf[k].sess);
        cudaFree(p->buf[k].d_seq); cudaFree(p->buf[k].d_off);
        if (p->buf[k].h2d_done) cudaEventDestroy(p->buf[k].h2d_done);
        if (p->buf[k].run_done) cudaEventDestroy(p->buf[k].run_done);
        if (p->buf[k].d2h_done) cudaEventDestroy(p->buf[k].d2h_done);
    }
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_run) cudaStreamDestroy(p->s_run);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    delete p;
}

static fmg_pipe_s *pipe_create(const fmg_index_t *idx, int64_t batch_reads, int max_len, uint64_t max_bytes) {
    fmg_pipe_s *p = new fmg_pipe_s;
    p->batch_reads = batch_reads; p->max_len = max_len; p->max_bytes = max_bytes;
    bool ok = false;
    do {
        CUDA_TRY(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking), break);
        CUDA_TRY(cudaStreamCreateWithFlags(&p->s_run, cudaStreamNonBlocking), break);
        CUDA_TRY(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking), break);
        int k = 0;
        for (; k < 2; ++k) {
            CUDA_TRY(cudaMalloc(&p->buf[k].d_seq, max_bytes), break);
            CUDA_TRY(cudaMalloc(&p->buf[k].d_off, (size_t)(batch_reads + 1) * 8), break);
            CUDA_TRY(cudaEventCreateWithFlags(&p->buf[k].h2d_done, cudaEventDisableTiming), break);
            CUDA_TRY(cudaEventCreateWithFlags(&p->buf[k].run_done, cudaEventDisableTiming), break);
            CUDA_TRY(cudaEventCreateWithFlags(&p->buf[k].d2h_done, cudaEventDisableTiming), break);
            p->buf[k].sess = fmg_smem_session_create(idx, batch_reads, max_len);
            if (!p->buf[k].sess) break;
        }
        ok = k == 2;
    } while (0);
    if (!ok) { fmg_pipe_destroy(p); return nullptr; }
    return p;
}

extern "C" {

__global__ void k_rebase_offsets(uint64_t *off, int64_t n, uint64_t sub) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) off[i] -= sub;
}

// fmintv_t records (32 bytes) -> fmg_intv16_t (16 bytes): x[0..2] as 32-bit values, info = end | start << 16 | left_closed << 31
__global__ void __launch_bounds__(256) k_pack16(uint64_t n, const uint4 *__restrict__ in, uint4 *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 a = in[2 * i], b = in[2 * i + 1];              // a = x0 lo, x0 hi, x1 lo, x1 hi;  b = x2 lo, x2 hi, info lo (end), info hi (start | flag << 31)
    uint4 o;
    o.x = a.x; o.y = a.z; o.z = b.x;
    o.w = (b.z & 0xffffu) | (b.w & 0x7fffu) << 16 | (b.w & 0x80000000u);
    out[i] = o;
}

static int smem_batch_into_impl(const fmg_index_t *idx, int64_t n, const uint8_t *seq, const uint64_t *off, int self_match,
                                void *mem_v, int rec_bytes, uint64_t mem_cap, uint64_t *mem_off, uint64_t *n_records, int64_t batch_reads);

int fmg_smem_batch_into(const fmg_index_t *idx, int64_t n, const uint8_t *seq, const uint64_t *off, int self_match,
                        fmg_intv_t *mem, uint64_t mem_cap, uint64_t *mem_off, uint64_t *n_records, int64_t batch_reads) {
    return smem_batch_into_impl(idx, n, seq, off, self_match, mem, 32, mem_cap, mem_off, n_records, batch_reads);
}

int fmg_smem_batch_into16(const fmg_index_t *idx, int64_t n, const uint8_t *seq, const uint64_t *off, int self_match,
                          fmg_intv16_t *mem, uint64_t mem_cap, uint64_t *mem_off, uint64_t *n_records, int64_t batch_reads) {
    if (idx && idx->view.n_sym >= (1ull << 32)) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] packed records need an index of < 2^32 symbols\n", __func__);
        return -3;
    }
    return smem_batch_into_impl(idx, n, seq, off, self_match, mem, 16, mem_cap, mem_off, n_records, batch_reads);
}

void fmg_intv16_expand(uint64_t n, const fmg_intv16_t *in, fmg_intv_t *out) {
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t w = in[i].info;
        out[i].x[0] = in[i].x[0]; out[i].x[1] = in[i].x[1]; out[i].x[2] = in[i].x[2];
        out[i].info = (uint64_t)(w & 0xffffu) | (uint64_t)(w >> 16 & 0x7fffu) << 32 | (uint64_t)(w >> 31) << 63;
    }
}

static int smem_batch_into_impl(const fmg_index_t *idx, int64_t n, const uint8_t *seq, const uint64_t *off, int self_match,
                                void *mem_v, int rec_bytes, uint64_t mem_cap, uint64_t *mem_off, uint64_t *n_records, int64_t batch_reads) {
    const char *const __func__name = rec_bytes == 16 ? "fmg_smem_batch_into16" : "fmg_smem_batch_into";
    (void)__func__name;
    uint8_t *const mem = static_cast<uint8_t *>(mem_v);
    if (!idx || use_device(idx->device, __func__)) return -1;
    if (n_records) *n_records = 0;
    mem_off[0] = 0;
    if (n <= 0) return 0;
    if (batch_reads <= 0) batch_reads = 1 << 20;
    batch_reads = std::min<int64_t>(batch_reads, n);
    // Batch boundaries.  The first copy in and the last copy out cannot overlap any kernel, so a call of several full batches
    // starts and ends with short ones (1/4, 1/2, 1, 1, ..., 1/2, 1/4 of batch_reads): the exposed copies shrink fourfold.
    std::vector<int64_t> cut{0};
    {
        const int64_t q = batch_reads / 4, h = batch_reads / 2;
        if (n >= 3 * batch_reads && q >= 1024) {
            cut.push_back(q); cut.push_back(q + h);
            const int64_t tail = h + q, body_end = n - tail;
            for (int64_t p0 = q + h; p0 < body_end;) { p0 = std::min(body_end, p0 + batch_reads); cut.push_back(p0); }
            cut.push_back(n - q); cut.push_back(n);
        } else
            for (int64_t p0 = 0; p0 < n;) { p0 = std::min(n, p0 + batch_reads); cut.push_back(p0); }
    }
    const int64_t n_batches = (int64_t)cut.size() - 1;
    uint64_t max_bytes = 1;
    for (int64_t b = 0; b < n_batches; ++b) max_bytes = std::max<uint64_t>(max_bytes, off[cut[b + 1]] - off[cut[b]]);
    // The longest read sizes the per-lane candidate lists.  Scanning all n lengths up front would sit in front of the first
    // kernel (10 ms for 10 M reads), so only the first batch is scanned here; every later batch is scanned just before it is
    // issued, while the GPU works on its predecessors, and a longer read than the pipeline was built for rebuilds it.
    auto batch_max_len = [&](int64_t b) {
        int m = 1;
        for (int64_t i = cut[b]; i < cut[b + 1]; ++i) m = std::max<int>(m, (int)(off[i + 1] - off[i]));
        return m;
    };
    int max_len = batch_max_len(0);
    auto fits16 = [&](int ml) {                       // packed records hold the start in 15 bits and the end in 16
        if (rec_bytes == 32 || ml < 32768) return true;
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] packed records need reads shorter than 32768 bases (one has %d)\n", __func__name, ml);
        return false;
    };
    if (!fits16(max_len)) return -3;

    // one pipeline per index, rebuilt only when a call needs larger batches or longer reads
    std::lock_guard<std::mutex> guard(idx->pipe_lock);
    auto ensure_pipe = [&](int need_len) -> bool {
        if (idx->pipe && (idx->pipe->batch_reads < batch_reads || idx->pipe->max_len < need_len || idx->pipe->max_bytes < max_bytes)) {
            fmg_pipe_destroy(idx->pipe);
            idx->pipe = nullptr;
        }
        if (!idx->pipe) idx->pipe = pipe_create(idx, batch_reads, need_len, max_bytes);
        return idx->pipe != nullptr;
    };
    if (!ensure_pipe(max_len)) return -1;
    cudaStream_t s_in = idx->pipe->s_in, s_run = idx->pipe->s_run, s_out = idx->pipe->s_out;
    BatchBuf *buf = idx->pipe->buf;
    int rc = -1;
    uint64_t total = 0;
    bool short_cap = false;
    do {
        auto issue = [&](int64_t b) -> int {          // H2D + kernels of batch b
            BatchBuf &B = buf[b & 1];
            B.first = cut[b]; B.count = cut[b + 1] - cut[b];
            const uint64_t o0 = off[B.first], nb = off[B.first + B.count] - o0;
            // the device buffers of this slot are free once the D2H of batch b-2 has finished
            CUDA_TRY(cudaStreamWaitEvent(s_in, B.d2h_done, 0), return -1);
            CUDA_TRY(cudaMemcpyAsync(B.d_seq, seq + o0, nb, cudaMemcpyHostToDevice, s_in), return -1);
            CUDA_TRY(cudaMemcpyAsync(B.d_off, off + B.first, (size_t)(B.count + 1) * 8, cudaMemcpyHostToDevice, s_in), return -1);
            CUDA_TRY(cudaEventRecord(B.h2d_done, s_in), return -1);
            CUDA_TRY(cudaStreamWaitEvent(s_run, B.h2d_done, 0), return -1);
            if (o0) {
                k_rebase_offsets<<<(unsigned)((B.count + 1 + 255) / 256), 256, 0, s_run>>>(B.d_off, B.count + 1, o0);
                LAUNCH_CHECK(return -1);
            }
            if (fmg_smem_session_run(B.sess, B.count, B.d_seq, B.d_off, self_match, s_run)) return -1;
            CUDA_TRY(cudaEventRecord(B.run_done, s_run), return -1);
            return 0;
        };
        auto drain = [&](int64_t b) -> int {          // D2H of batch b
            BatchBuf &B = buf[b & 1];
            fmg_smem_session_t *S = B.sess;
            CUDA_TRY(cudaEventSynchronize(B.run_done), return -1);       // only this batch, not the one queued behind it
            if (S->h_ctrl[1] != 0 && fmg_smem_session_result(S, nullptr, nullptr, nullptr)) return -1;  // slot overflow: slow path
            B.n_rec = S->h_ctrl[2];
            B.rec_base = total;
            total += B.n_rec;
            if (total > mem_cap) short_cap = true;
            if (B.rec_base) {
                k_rebase_offsets<<<(unsigned)((B.count + 1 + 255) / 256), 256, 0, s_out>>>(S->mem_off, B.count + 1, 0 - B.rec_base);
                LAUNCH_CHECK(return -1);
            }
            if (!short_cap && B.n_rec) {
                const void *src = S->mem;
                if (rec_bytes == 16) {                                   // pack into the slot array, which the compaction has finished with
                    k_pack16<<<(unsigned)((B.n_rec + 255) / 256), 256, 0, s_out>>>(B.n_rec, S->mem, S->slots);
                    LAUNCH_CHECK(return -1);
                    src = S->slots;
                }
                CUDA_TRY(cudaMemcpyAsync(mem + B.rec_base * (uint64_t)rec_bytes, src, B.n_rec * (uint64_t)rec_bytes, cudaMemcpyDeviceToHost, s_out), return -1);
            }
            CUDA_TRY(cudaMemcpyAsync(mem_off + B.first, S->mem_off, (size_t)(B.count + 1) * 8, cudaMemcpyDeviceToHost, s_out), return -1);
            CUDA_TRY(cudaEventRecord(B.d2h_done, s_out), return -1);
            return 0;
        };
        bool fail = false;
        for (int64_t b = 0; b <= n_batches && !fail; ++b) {
            bool drained = false;
            if (b >= 1 && b < n_batches) {
                const int ml = batch_max_len(b);
                if (!fits16(ml)) { fail = true; break; }
                if (ml > idx->pipe->max_len) {
                    // a longer read than any before: finish what is in flight, then rebuild the sessions for it
                    if (drain(b - 1)) { fail = true; break; }
                    drained = true;
                    CUDA_TRY(cudaStreamSynchronize(s_out), fail = true; break);
                    if (!ensure_pipe(ml)) { fail = true; break; }
                    s_in = idx->pipe->s_in; s_run = idx->pipe->s_run; s_out = idx->pipe->s_out; buf = idx->pipe->buf;
                }
            }
            if (b < n_batches && issue(b)) { fail = true; break; }
            if (b >= 1 && !drained && drain(b - 1)) { fail = true; break; }
        }
        if (fail) break;
        CUDA_TRY(cudaStreamSynchronize(s_out), break);
        mem_off[n] = total;
        if (n_records) *n_records = total;
        rc = short_cap ? 1 : 0;
    } while (0);
    return rc;
}

} // extern "C"
