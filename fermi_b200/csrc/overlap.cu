// Overlap path of `fermi unitig` on the GPU: kernels + C-ABI (see fmd_overlap.cuh for the algorithm).
//   k_retrieve     fm_retrieve (exact.c:59-70), one sequence per thread
//   k_overlap<U>   fm6_is_contained + fm6_get_nei + check_left_simple (unitig.c:77-204), persistent lanes
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <vector>
#include <algorithm>
#include <chrono>
#include <mutex>
#include "fmd_overlap.cuh"
#include "fmg_internal.hpp"
#include "../../include/fermi_b200.h"

using namespace fmg;

extern std::atomic<uint64_t> g_launches;

#define OV_TRY(call)                                                                                  \
    do {                                                                                              \
        cudaError_t err__ = (call);                                                                   \
        if (err__ != cudaSuccess) {                                                                   \
            if (fmg_verbose >= 1)                                                                     \
                std::fprintf(stderr, "[E::%s] %s failed: %s\n", __func__, #call, cudaGetErrorString(err__)); \
            return -1;                                                                                \
        }                                                                                             \
    } while (0)

__global__ void __launch_bounds__(256) k_retrieve(RetrieveArgs A) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < A.n) retrieve_one(A, t);
}

template <typename U>
__global__ void __launch_bounds__(OVLP_BLOCK, OVLP_MIN_BLOCKS) k_overlap(OverlapArgs A) {
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    overlap_lane<U>(A, lane, [&]() -> int64_t { return (int64_t)atomicAdd(A.next, 1ull); });
}

// the record compaction kernels of the SMEM path (fmg_cuda.cu) are reused for the neighbour slots
int fmg_compact_slots(const uint32_t *cnt, int64_t n, int cap, const uint4 *slots, uint4 *mem, uint64_t *mem_off, uint64_t *tile_sum,
                      unsigned long long *ctrl, cudaStream_t st);
int64_t fmg_compact_tiles(int64_t n);

namespace {
// Device scratch is recycled between calls: a unitig run issues one call per 2 M sequences and each needs ~3 GB of
// lists and slots; cudaMalloc/cudaFree of those costs more than the kernels.  Blocks return to a small pool and are
// handed out again when they fit (released by fmg_release_cache or at process exit).
struct Pool {
    struct Blk { void *p; size_t cap; int dev; };
    std::vector<Blk> free_list;
    std::mutex lock;
    cudaError_t get(size_t bytes, int dev, void **out, size_t *cap) {
        {
            std::lock_guard<std::mutex> g(lock);
            for (size_t i = 0; i < free_list.size(); ++i)
                if (free_list[i].dev == dev && free_list[i].cap >= bytes && free_list[i].cap <= 2 * bytes + (1 << 20)) {
                    *out = free_list[i].p; *cap = free_list[i].cap;
                    free_list.erase(free_list.begin() + i);
                    return cudaSuccess;
                }
        }
        *cap = bytes;
        cudaError_t e = cudaMalloc(out, bytes);
        if (e != cudaSuccess) { release(); cudaGetLastError(); e = cudaMalloc(out, bytes); }
        return e;
    }
    void put(void *p, size_t cap, int dev) { std::lock_guard<std::mutex> g(lock); free_list.push_back(Blk{p, cap, dev}); }
    void release() { std::lock_guard<std::mutex> g(lock); for (auto &b : free_list) cudaFree(b.p); free_list.clear(); }
} g_pool;

struct Dev {
    void *p = nullptr;
    size_t cap = 0;
    int dev = 0;
    ~Dev() { if (p) g_pool.put(p, cap, dev); }
    cudaError_t alloc(size_t b) {
        if (p) { g_pool.put(p, cap, dev); p = nullptr; }
        cudaGetDevice(&dev);
        return g_pool.get(b ? b : 1, dev, &p, &cap);
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
};
}

extern "C" void fmg_release_cache(void) { g_pool.release(); }

extern "C" {

int fmg_overlap_batch(const fmg_index_t *idx, int min_match, int64_t n, const uint64_t *ids, uint64_t first, uint64_t step,
                      int max_len, int64_t *rec, fmg_intv_t **nei, uint64_t *nei_off, uint8_t *seq, int32_t *len, uint8_t *ext) {
    if (!idx || n < 0 || max_len <= 0 || !rec || !nei || !nei_off) return -1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] no CUDA device available; libfermi_b200 has no CPU path\n", __func__);
        return -1;
    }
    OV_TRY(cudaSetDevice(idx->device));
    nei_off[0] = 0;
    *nei = nullptr;
    if (n == 0) { *nei = (fmg_intv_t *)std::malloc(32); return 0; }
    const bool wide = idx->view.n_sym + 256 >= (1ull << 32) || std::getenv("FMG_FORCE_WIDE") != nullptr;

    Dev d_ids, d_seq, d_len, d_ret, d_rec, d_ext, d_cnt, d_slots, d_mem, d_off, d_tiles, d_ctrl, d_sbuf, d_A, d_B, d_cat;
    OV_TRY(d_seq.alloc((size_t)n * max_len)); OV_TRY(d_len.alloc((size_t)n * 4)); OV_TRY(d_ret.alloc((size_t)n * 8));
    OV_TRY(d_rec.alloc((size_t)n * OV_NREC * 8)); OV_TRY(d_ext.alloc((size_t)n * max_len)); OV_TRY(d_cnt.alloc((size_t)(n + 1) * 4));
    OV_TRY(d_off.alloc((size_t)(n + 1) * 8)); OV_TRY(d_tiles.alloc((size_t)fmg_compact_tiles(n) * 8)); OV_TRY(d_ctrl.alloc(64));
    if (ids) { OV_TRY(d_ids.alloc((size_t)n * 8)); OV_TRY(cudaMemcpy(d_ids.p, ids, (size_t)n * 8, cudaMemcpyHostToDevice)); }

    const auto t0 = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); };
    // ---- sequences
    RetrieveArgs R;
    R.ix = idx->view; R.n = n; R.ids = ids ? d_ids.as<uint64_t>() : nullptr; R.first = first; R.step = step;
    R.seq = d_seq.as<uint8_t>(); R.max_len = max_len; R.len = d_len.as<int32_t>(); R.ret = d_ret.as<int64_t>();
    k_retrieve<<<(unsigned)((n + 255) / 256), 256>>>(R);
    ++g_launches;
    OV_TRY(cudaGetLastError());
    std::vector<int32_t> h_len(n);
    OV_TRY(cudaMemcpy(h_len.data(), d_len.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    const double t_retrieve = since(t0);
    const auto t1 = std::chrono::steady_clock::now();
    for (int64_t i = 0; i < n; ++i)
        if (h_len[i] < 0) {
            if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] a sequence of %d bases exceeds max_len=%d\n", __func__, -h_len[i], max_len);
            return 2;
        }

    // ---- overlap records; scratch capacities grow until nothing overflows
    int per_sm = 0;
    if (wide) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_overlap<uint64_t>, OVLP_BLOCK, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_overlap<uint32_t>, OVLP_BLOCK, 0);
    if (per_sm < 1) per_sm = 1;
    const int grid = (int)std::min<int64_t>((int64_t)idx->n_sm * per_sm, (n + OVLP_BLOCK - 1) / OVLP_BLOCK);
    const int64_t n_lanes = (int64_t)grid * OVLP_BLOCK;
    int cap = 4 * max_len, nei_cap = 8;
    const int s_cap = 2 * max_len + 8;
    int64_t *h_rec = rec;
    std::vector<uint32_t> h_cnt(n);
    for (int attempt = 0;; ++attempt) {
        const size_t esz = wide ? 32 : 16;
        OV_TRY(d_sbuf.alloc((size_t)n_lanes * s_cap)); OV_TRY(d_A.alloc((size_t)n_lanes * cap * esz)); OV_TRY(d_B.alloc((size_t)n_lanes * cap * esz));
        OV_TRY(d_cat.alloc((size_t)n_lanes * cap * 4)); OV_TRY(d_slots.alloc((size_t)n * nei_cap * 32)); OV_TRY(d_mem.alloc((size_t)n * nei_cap * 32));
        OV_TRY(cudaMemset(d_ctrl.p, 0, 64));
        OverlapArgs O;
        O.ix = idx->view; O.min_match = min_match; O.n = n; O.seq = d_seq.as<uint8_t>(); O.len = d_len.as<int32_t>(); O.max_len = max_len;
        O.sbuf = d_sbuf.as<uint8_t>(); O.s_cap = s_cap; O.A = d_A.p; O.B = d_B.p; O.cap = cap; O.cat = d_cat.as<int32_t>();
        O.rec = d_rec.as<int64_t>(); O.nei = d_slots.as<uint4>(); O.nei_cap = nei_cap; O.nei_cnt = d_cnt.as<uint32_t>();
        O.ext = d_ext.as<uint8_t>(); O.next = d_ctrl.as<unsigned long long>();
        if (wide) k_overlap<uint64_t><<<grid, OVLP_BLOCK>>>(O); else k_overlap<uint32_t><<<grid, OVLP_BLOCK>>>(O);
        ++g_launches;
        OV_TRY(cudaGetLastError());
        OV_TRY(cudaDeviceSynchronize());
        if (fmg_verbose >= 4) std::fprintf(stderr, "[M::%s] k_retrieve %.3f s, k_overlap (attempt %d) %.3f s for %lld sequences\n", __func__, t_retrieve, attempt, since(t1), (long long)n);
        OV_TRY(cudaMemcpy(h_rec, d_rec.p, (size_t)n * OV_NREC * 8, cudaMemcpyDeviceToHost));
        OV_TRY(cudaMemcpy(h_cnt.data(), d_cnt.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
        bool list_ovf = false, nei_ovf = false;
        for (int64_t i = 0; i < n; ++i) {
            if (h_rec[i * OV_NREC + OV_CONTAINED] == -100) list_ovf = true;
            if (h_cnt[i] > (uint32_t)nei_cap) nei_ovf = true;
        }
        if (!list_ovf && !nei_ovf) break;
        if (attempt == 6) {
            if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] scratch overflow persists (cap=%d, nei_cap=%d)\n", __func__, cap, nei_cap);
            return -1;
        }
        if (list_ovf) cap *= 4;
        if (nei_ovf || list_ovf) nei_cap *= 4;
        if (fmg_verbose >= 3) std::fprintf(stderr, "[M::%s] scratch overflow; re-running with %d candidate / %d neighbour slots\n", __func__, cap, nei_cap);
    }
    // the value fm_retrieve returned goes to rec[OV_K]
    std::vector<int64_t> h_ret(n);
    OV_TRY(cudaMemcpy(h_ret.data(), d_ret.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; ++i) h_rec[i * OV_NREC + OV_K] = h_ret[i];

    // ---- neighbour slots -> dense array + offsets
    if (fmg_compact_slots(d_cnt.as<uint32_t>(), n, nei_cap, d_slots.as<uint4>(), d_mem.as<uint4>(), d_off.as<uint64_t>(),
                          d_tiles.as<uint64_t>(), d_ctrl.as<unsigned long long>(), nullptr)) return -1;
    OV_TRY(cudaMemcpy(nei_off, d_off.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost));
    const uint64_t tot = nei_off[n];
    *nei = (fmg_intv_t *)std::malloc((tot ? tot : 1) * 32);
    if (tot) OV_TRY(cudaMemcpy(*nei, d_mem.p, tot * 32, cudaMemcpyDeviceToHost));
    if (seq) OV_TRY(cudaMemcpy(seq, d_seq.p, (size_t)n * max_len, cudaMemcpyDeviceToHost));
    if (len) std::memcpy(len, h_len.data(), (size_t)n * 4);
    if (ext) OV_TRY(cudaMemcpy(ext, d_ext.p, (size_t)n * max_len, cudaMemcpyDeviceToHost));
    if (fmg_verbose >= 4) std::fprintf(stderr, "[M::%s] batch total %.3f s\n", __func__, since(t0));
    return 0;
}

} // extern "C"
