// Device scratch pool and the pinned host cache shared by the overlap pass (overlap.cu) and the device unitig
// assembly (unitig_gpu.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <mutex>
#include <vector>

namespace fmg {
// Device scratch is recycled between calls: a unitig run needs gigabytes of lists, slots and record arrays, and
// cudaMalloc/cudaFree of those costs more than the kernels.  Blocks return to a small pool and are handed out again
// when they fit (released by fmg_release_cache or at process exit).
struct Pool {
    struct Blk { void *p; size_t cap; int dev; };
    std::vector<Blk> free_list;
    std::mutex lock;
    cudaError_t get(size_t bytes, int dev, void **out, size_t *cap) {
        {
            std::lock_guard<std::mutex> g(lock);
            // best fit: a repeated run finds, for every request, exactly the block its previous run returned
            size_t best = free_list.size();
            for (size_t i = 0; i < free_list.size(); ++i)
                if (free_list[i].dev == dev && free_list[i].cap >= bytes && free_list[i].cap <= 2 * bytes + (1 << 20) &&
                    (best == free_list.size() || free_list[i].cap < free_list[best].cap))
                    best = i;
            if (best != free_list.size()) {
                *out = free_list[best].p; *cap = free_list[best].cap;
                free_list.erase(free_list.begin() + best);
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
};
extern Pool g_pool;               // overlap.cu

struct Dev {
    void *p = nullptr;
    size_t cap = 0;
    int dev = 0;
    Dev() = default;
    Dev(const Dev &) = delete;
    Dev &operator=(const Dev &) = delete;
    ~Dev() { if (p) g_pool.put(p, cap, dev); }
    cudaError_t alloc(size_t b) {
        if (p) { g_pool.put(p, cap, dev); p = nullptr; }
        cudaGetDevice(&dev);
        return g_pool.get(b ? b : 1, dev, &p, &cap);
    }
    void swap(Dev &o) { std::swap(p, o.p); std::swap(cap, o.cap); std::swap(dev, o.dev); }
    template <class T> T *as() const { return static_cast<T *>(p); }
};
// device-resident result of the whole-index overlap pass (overlap.cu: fmg_overlap_pass)
struct OvDevice {
    Dev pack, rank, ext, spill;     // OvPack[n_seq] by rank; rank of each BWT row; appended bases; neighbour lists of forks
    uint64_t n_seq = 0, ext_total = 0, spill_total = 0;
    int max_len = 0;
};

// a shard of the pass written into caller-owned device buffers (multi-GPU: fmg_overlap_shard)
struct OvShard {
    uint64_t row_lo, row_hi;        // BWT rows [row_lo, row_hi)
    void *pack;                     // OvPack[n_seq], zeroed by the caller; only the ranks of the shard's rows are written
    int64_t *rank;                  // row_hi - row_lo
    uint8_t *ext; uint64_t ext_cap;
    void *spill; uint64_t spill_cap;        // entries of 32 bytes
    uint64_t ext_total, spill_total;        // out (the need, when a capacity was too small)
    int max_len;                            // out
};
// raw view of device-resident records for the assembly
struct OvDevView {
    const void *pack; const int64_t *rank; const uint8_t *ext; const void *spill;
    uint64_t n_seq, ext_total, spill_total;
};
}  // namespace fmg

struct fmg_index_s;
struct OvHost;
// Overlap records of every sequence of the index: into `dev` (kept in HBM for the device unitig assembly) and / or
// into `host` (pinned arrays for the host walk); either may be nullptr.  Returns 0, or -1 on a CUDA error.
int fmg_overlap_pass(const fmg_index_s *idx, int min_match, int max_len, fmg::OvDevice *dev, OvHost *host, fmg::OvShard *shard = nullptr);
// unitig_gpu.cu: unitigs from the device-resident records; 0 = written, 1 = the link graph is irregular (cycles or
// one-sided links: the caller falls back to the host walk, which reproduces the reference's seed order), -1 = error
// part / n_parts: only the chains whose head rank % n_parts == part are emitted (several GPUs share the emission); sink != nullptr
// keeps the MAG text in memory instead of writing it
struct fmg_magpart_s;
int fmg_unitig_device(const fmg_index_s *idx, const fmg::OvDevView &D, int min_match, const char *out_path, uint64_t *n_unitigs,
                      uint32_t part, uint32_t n_parts, fmg_magpart_s *sink);

// Pinned host arrays of the whole-index pass and of the device unitig assembly.  Page-locking gigabytes costs more
// than the pass itself, so the arrays stay with the index handle and are reused (grown on demand) by the next call.
struct fmg_ovcache_s {
    struct Pin {
        void *p = nullptr;
        size_t cap = 0;
        cudaError_t need(size_t bytes) {
            if (bytes <= cap) return cudaSuccess;
            if (p) cudaFreeHost(p);
            p = nullptr; cap = 0;
            const cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
            if (e == cudaSuccess) cap = bytes;
            return e;
        }
        ~Pin() { if (p) cudaFreeHost(p); }
        template <class T> T *as() const { return static_cast<T *>(p); }
    } pack, rank, seq, ext, spill, ctrl, text;
};
