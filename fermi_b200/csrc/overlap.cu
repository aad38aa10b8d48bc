// Overlap path of `fermi unitig` on the GPU: kernels + C-ABI (see fmd_overlap.cuh for the algorithm).
//   k_ov_chain<U,1|3>  fm_retrieve (exact.c:59-70) fused with fm6_is_contained / the overlap_intv of check_left_simple (unitig.c:77-91,186-190), one sequence per thread
//   k_ov_lists<U,2|4>  fm6_get_nei / the candidate loop of check_left_simple (unitig.c:93-179,191-203), persistent lanes
//   k_ov_pack          batch records -> rank-indexed 64-byte records + compact ext / spill arrays (whole-index pass)
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <vector>
#include <algorithm>
#include <chrono>
#include <mutex>
#include "fmd_overlap.cuh"
#include "ov_records.hpp"
#include "dev_pool.hpp"
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

// phase 1 is run by whole warps (paired block loads, fmd_device.cuh: load_blk_pair): threads past the end of the batch take part
// without a sequence.  Phase 3: one sequence per thread, a straight chain of extensions.
template <typename U, int PHASE>
__global__ void __launch_bounds__(OVCH_BLOCK) k_ov_chain(const __grid_constant__ OverlapArgs A) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (PHASE == 1) overlap_chain<U, 1>(A, t, t < A.n);
    else if (t < A.n) overlap_chain<U, PHASE>(A, t);
}

// phases 2 and 4: persistent lanes, sequences handed out by atomicAdd, every extension behind a warp vote; the level lists of
// phase 2 start in (dynamic) shared memory
template <typename U, int PHASE>
__global__ void __launch_bounds__(OVLP_BLOCK, OVLP_MIN_BLOCKS) k_ov_lists(const __grid_constant__ OverlapArgs A) {     // &A.ix is taken: no stack copy
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    overlap_lane_sync<U, PHASE>(A, lane, [&]() -> int64_t { return (int64_t)atomicAdd(A.next, 1ull); });
}
// phase 2 (fm6_get_nei): persistent lanes of the flat state machine nei_lane, one converged gather per trip; the first entries of
// the level lists in (dynamic) shared memory
// MINB = resident blocks per SM the register allocation aims at (4: 128 registers, 5: 96, 6: 80 with some spilling); the best one is
// a measured choice (FMG_NEI_BLOCKS overrides it for experiments)
template <typename U, int MINB>
__global__ void __launch_bounds__(OVLP_BLOCK, MINB) k_ov_nei(const __grid_constant__ OverlapArgs A) {
    extern __shared__ uint4 fmg_ov_shared[];
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    nei_lane<U, MINB>(A, lane, [&]() -> int64_t {
        const int64_t j = (int64_t)atomicAdd(A.next, 1ull);
        return A.order && j < A.n ? (int64_t)A.order[j] : j;
    }, fmg_ov_shared, (int)blockDim.x, (int)threadIdx.x);
}

// The lanes of a warp stay in step only while their sequences take the same trips: a sequence costs (levels) trips of (candidates)
// rank triples each, levels = length - longest overlap.  The batch is therefore handed out sorted by that pair (heaviest first), so
// that the 32 sequences a warp works on at any time are alike: key = levels << 8 | candidates, 0 for sequences without work.
template <typename U>
__global__ void __launch_bounds__(256) k_nei_key(OverlapArgs A, uint32_t *__restrict__ key, uint32_t *__restrict__ idx) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= A.n) return;
    const int np = A.np0[t];
    uint32_t k = 0;
    if (np > 0) {
        const int npc = np < A.pcap ? np : A.pcap;
        const IntvT<U> first = ld_cand(static_cast<const IntvT<U> *>(A.P0) + (size_t)t * A.pcap + (A.pcap - npc));     // the longest overlap
        const int levels = A.len[t] - (int)first.info;
        k = (uint32_t)(levels < 1 ? 1 : levels > 4095 ? 4095 : levels) << 8 | (uint32_t)(np > 255 ? 255 : np);
    }
    key[t] = k; idx[t] = (uint32_t)t;
}
static int nei_minb() {
    static const int v = [] { const char *e = std::getenv("FMG_NEI_BLOCKS"); const int x = e ? std::atoi(e) : 4; return x <= 3 ? 3 : x >= 5 ? 5 : 4; }();
    return v;
}
template <typename U> static const void *nei_kernel() {
    const int m = nei_minb();
    return m == 3 ? (const void *)k_ov_nei<U, 3> : m == 5 ? (const void *)k_ov_nei<U, 5> : (const void *)k_ov_nei<U, 4>;
}
template <typename U> static constexpr size_t lists_shared_bytes() { return NeiLists<U>::shared_bytes(OVLP_BLOCK); }

// ---- whole-index pass (fmg_overlap_all): per-batch records -> rank-indexed packed records + compact ext / spill arrays
enum { OVC_NEXT = 0, OVC_NEXT2, OVC_EXT, OVC_SPILL, OVC_FLAGS, OVC_MAXLEN, OVC_NLEFT, OVC_N };
enum { OVF_LIST = 1, OVF_NEI = 2, OVF_EXT = 4, OVF_SPILL = 8, OVF_FIELD = 16, OVF_RANK = 32 };

struct PackArgs {
    int64_t n;                          // sequences (BWT rows) of this batch
    const int64_t *rec;                 // n x OV_NREC
    const int64_t *ret;                 // n: fm_retrieve's return value = rank of the sequence
    const int32_t *len;                 // n: < 0 flags a sequence longer than max_len
    const uint32_t *nei_cnt;
    const uint4 *nei_slots; int nei_cap;
    const uint8_t *ext; int max_len;
    OvPack *pack; uint64_t n_seq;       // by_row == 0: n_seq records indexed by rank; by_row == 1: the records of this batch in row order
    int by_row;
    uint8_t *ext_out; uint64_t ext_cap;
    uint4 *spill_out; uint64_t spill_cap;
    unsigned long long *ctrl;
};

// Space in the ext / spill arrays is handed out by one atomicAdd per warp (warp-inclusive scan of the needs); the
// order of the entries therefore varies from run to run, the records address them through ext_first / nx0.
__global__ void __launch_bounds__(256) k_ov_pack(PackArgs A) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < A.n;
    const int64_t *rec = A.rec + (live ? t : 0) * OV_NREC;
    uint32_t flags = 0, el = 0, sp = 0, cnt = 0;
    if (live) {
        cnt = A.nei_cnt[t];
        if (rec[OV_CONTAINED] == -100) flags |= OVF_LIST;
        if (cnt > (uint32_t)A.nei_cap) flags |= OVF_NEI;
        if (A.len[t] < 0) atomicMax(A.ctrl + OVC_MAXLEN, (unsigned long long)(-(int64_t)A.len[t]));
        if (!flags) { el = ov_ext_len(rec); sp = cnt > 1 ? cnt : 0; }
    }
    uint32_t ie = el, is = sp;
    const int lane = threadIdx.x & 31;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, ie, o), b = __shfl_up_sync(0xffffffffu, is, o);
        if (lane >= o) ie += a, is += b;
    }
    unsigned long long be = 0, bs = 0;
    if (lane == 31) {
        if (ie) be = atomicAdd(A.ctrl + OVC_EXT, (unsigned long long)ie);
        if (is) bs = atomicAdd(A.ctrl + OVC_SPILL, (unsigned long long)is);
    }
    be = __shfl_sync(0xffffffffu, be, 31) + ie - el;
    bs = __shfl_sync(0xffffffffu, bs, 31) + is - sp;
    if (live && !flags) {
        const uint64_t k = (uint64_t)A.ret[t];
        if (k >= A.n_seq) flags |= OVF_RANK;
        if (be + el > A.ext_cap) flags |= OVF_EXT;
        if (bs + sp > A.spill_cap) flags |= OVF_SPILL;
        if (!flags) {
            const uint8_t *e = A.ext + (size_t)t * A.max_len;
            for (uint32_t i = 0; i < el; ++i) A.ext_out[be + i] = e[i];
            const uint4 *nb = A.nei_slots + (size_t)t * A.nei_cap * 2;
            uint64_t nx0 = 0, nx1 = 0, nx2 = 0;
            if (cnt == 1) {
                const Intv v = ld_intv(nb);
                nx0 = v.x0; nx1 = v.x1; nx2 = v.x2;
            } else if (cnt > 1) {
                nx0 = bs;
                for (uint32_t i = 0; i < 2 * cnt; ++i) A.spill_out[2 * bs + i] = nb[i];
            }
            OvPack o;
            if (!ov_pack(rec, nx0, nx1, nx2, be, &o)) flags |= OVF_FIELD;
            uint4 *dst = reinterpret_cast<uint4 *>(A.pack + (A.by_row ? (uint64_t)t : k));
            const uint4 *src = reinterpret_cast<const uint4 *>(&o);
            dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
        }
    }
    if (flags) atomicOr(A.ctrl + OVC_FLAGS, (unsigned long long)flags);
}

// ---- the deferred left check.  check_left (unitig.c:206-225) only decides something when check_left_simple (unitig.c:186-204)
// fails AND the reverse complement of the neighbour has more than one right neighbour; whether a record links to its unique
// neighbour therefore depends on OV_LEFT only if pack[nx1].nnei > 1 (unitig_gpu.cu: links / stop_nei, unitig_host.cpp).  The pass
// computes the records without the left check (phases 1 and 2) and evaluates it afterwards for exactly those rows -- none to
// a few per mille on error-free reads, where phases 3 and 4 used to cost a third of the kernel time for every sequence.
__global__ void __launch_bounds__(256) k_left_select(const OvPack *__restrict__ pack, const int64_t *__restrict__ rank_of_row, uint64_t row_lo, uint64_t row_hi,
                                                    uint64_t *ids, unsigned long long *count) {
    const uint64_t row = row_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool want = false;
    if (row < row_hi) {
        const OvPack p = pack[rank_of_row[row]];
        want = p.nnei == 1 && p.contained == 0 && p.rbeg >= 0 && p.left == 1 && pack[p.nx1].nnei > 1;
    }
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (want) ids[base + __popc(m & ((1u << lane) - 1u))] = row;
}

__global__ void __launch_bounds__(256) k_left_patch(int64_t n, const int64_t *__restrict__ rec, const int64_t *__restrict__ ret, const uint32_t *__restrict__ nei_cnt, int nei_cap,
                                                   OvPack *pack, unsigned long long *ctrl) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int64_t *r = rec + t * OV_NREC;
    if (r[OV_CONTAINED] == -100 || nei_cnt[t] > (uint32_t)nei_cap) { atomicOr(ctrl + OVC_FLAGS, (unsigned long long)OVF_LIST); return; }
    pack[ret[t]].left = (int8_t)r[OV_LEFT];
}

// the sequences of the odd rows (the seeds of unitig_core, unitig.c:333-334) of a batch whose first row is even, left-aligned
// for the host walk (the rows of the batch are right-aligned, fmd_overlap.cuh: OverlapArgs::seq_of)
__global__ void __launch_bounds__(256) k_seq_odd(const uint8_t *__restrict__ seq, const int32_t *__restrict__ len, int max_len, int64_t n_odd, uint8_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_odd * max_len) return;
    const int64_t row = i / max_len, col = i - row * max_len;
    const int l = len[2 * row + 1], lc = l < 0 ? max_len : l;
    out[i] = col < lc ? seq[(2 * row + 2) * max_len - lc + col] : 0;
}

// fm6_seqsort (seqsort.c:12-35): from fm6_retrieve's k / k2 / containment of every even row i, the rank table
//   sorted[k] = i << 2 | flag,  sorted[rank of the reverse complement] = (i | 1) << 2 | flag
__global__ void __launch_bounds__(256) k_seqsort_scatter(int64_t n, uint64_t row0, const int64_t *__restrict__ rec, const int64_t *__restrict__ ret,
                                                        uint64_t *__restrict__ sorted) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int64_t *r = rec + t * OV_NREC;
    const uint64_t i = row0 + 2 * (uint64_t)t, k = (uint64_t)ret[t];
    const uint64_t x0 = (uint64_t)r[OV_X0], x1 = (uint64_t)r[OV_X1], x2 = (uint64_t)r[OV_X2];
    const uint64_t flag = (uint64_t)(r[OV_CONTAINED] != 0) << 1 | (uint64_t)(x2 > 1 && k != x0);
    sorted[k] = i << 2 | flag;
    if (x0 != x1) sorted[x1 + (k - x0)] = (i | 1) << 2 | flag;        // seq and reverse complement are different
    else sorted[k + 1] = (i | 1) << 2 | flag;
}

// the record compaction kernels of the SMEM path (fmg_cuda.cu) are reused for the neighbour slots
int fmg_compact_slots(const uint32_t *cnt, int64_t n, int cap, const uint4 *slots, uint4 *mem, uint64_t *mem_off, uint64_t *tile_sum,
                      unsigned long long *ctrl, cudaStream_t st);
int64_t fmg_compact_tiles(int64_t n);

// phases 1 + 2 (the record without the left check) and phases 3 + 4 (check_left_simple) over one batch, back to back on `st`;
// ctrl2 = two work counters (zeroed by launch_records)
template <typename U>
static cudaError_t launch_records(OverlapArgs O, int grid, unsigned long long *ctrl2, cudaStream_t st, cudaEvent_t *ev = nullptr,
                                  uint32_t *ord = nullptr, void *ord_tmp = nullptr, size_t ord_tmp_bytes = 0) {
    cudaError_t e = cudaMemsetAsync(ctrl2, 0, 16, st);
    if (e != cudaSuccess) return e;
    const unsigned gch = (unsigned)((O.n + OVCH_BLOCK - 1) / OVCH_BLOCK);
    int g = (int)std::min<int64_t>(grid, (O.n + OVLP_BLOCK - 1) / OVLP_BLOCK);
    {
        // persistent lanes: no more blocks than are resident at once (the grid is sized for the roomier of the two list kernels)
        static int n_sm = 0;
        if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nei_kernel<U>(), OVLP_BLOCK, lists_shared_bytes<U>());
        if (per_sm > 0) g = std::min(g, n_sm * per_sm);
    }
    if (ev) cudaEventRecord(ev[0], st);
    k_ov_chain<U, 1><<<gch, OVCH_BLOCK, 0, st>>>(O);
    if (ev) cudaEventRecord(ev[1], st);
    O.next = ctrl2;
    O.order = nullptr;
    if (ord && O.n > 4096) {                     // hand the sequences out sorted by their cost (k_nei_key)
        uint32_t *k0 = ord, *k1 = ord + O.n, *i0 = ord + 2 * O.n, *i1 = ord + 3 * O.n;
        k_nei_key<U><<<(unsigned)((O.n + 255) / 256), 256, 0, st>>>(O, k0, i0);
        size_t need = ord_tmp_bytes;
        e = cub::DeviceRadixSort::SortPairsDescending(ord_tmp, need, k0, k1, i0, i1, (int64_t)O.n, 0, 20, st);
        if (e != cudaSuccess) return e;
        O.order = i1;
        g_launches += 2;
    }
    {
        void *kargs[] = {(void *)&O};
        e = cudaLaunchKernel(nei_kernel<U>(), dim3((unsigned)g), dim3(OVLP_BLOCK), kargs, lists_shared_bytes<U>(), st);
        if (e != cudaSuccess) return e;
    }
    if (ev) cudaEventRecord(ev[2], st);
    g_launches += 2;
    return cudaGetLastError();
}
template <typename U>
static cudaError_t launch_left(OverlapArgs O, int grid, unsigned long long *ctrl2, cudaStream_t st) {
    const unsigned gch = (unsigned)((O.n + OVCH_BLOCK - 1) / OVCH_BLOCK);
    const int g = (int)std::min<int64_t>(grid, (O.n + OVLP_BLOCK - 1) / OVLP_BLOCK);
    k_ov_chain<U, 3><<<gch, OVCH_BLOCK, 0, st>>>(O);
    O.next = ctrl2 + 1;
    k_ov_lists<U, 4><<<g, OVLP_BLOCK, 0, st>>>(O);
    g_launches += 2;
    return cudaGetLastError();
}

static int lists_blocks_per_sm(bool wide) {
    int a = 0, b = 0;
    if (wide) {
        cudaFuncSetAttribute(nei_kernel<uint64_t>(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lists_shared_bytes<uint64_t>());
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, nei_kernel<uint64_t>(), OVLP_BLOCK, lists_shared_bytes<uint64_t>());
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_ov_lists<uint64_t, 4>, OVLP_BLOCK, 0);
    } else {
        cudaFuncSetAttribute(nei_kernel<uint32_t>(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lists_shared_bytes<uint32_t>());
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, nei_kernel<uint32_t>(), OVLP_BLOCK, lists_shared_bytes<uint32_t>());
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_ov_lists<uint32_t, 4>, OVLP_BLOCK, 0);
    }
    return std::max(1, std::max(a, b));       // lane scratch is sized for the larger grid
}

// narrow (32-bit) candidate entries pack the read position in 16 bits (fmd_overlap.cuh: OvBits): sequences of 64 kb or more, or
// an index of 2^32 symbols, need the wide kernels
static bool ov_wide(const fmg_index_s *idx, int max_len) {
    return idx->view.n_sym + 256 >= (1ull << 32) || max_len >= 65536 || std::getenv("FMG_FORCE_WIDE") != nullptr;
}
static inline int round8(int v) { return (v + 7) & ~7; }

namespace fmg { Pool g_pool; }

// device scratch of the phase kernels for batches of up to `nb` rows
struct OvScratch {
    Dev seq, len, rec, ext, cnt, slots, P0, S0, np0, A, B, cat, S, ord, ord_tmp;
    size_t ord_tmp_bytes = 0;
    int max_len = 0, pcap = 0, cap = 0, nei_cap = 0, grid = 0;
    bool wide = false;
    cudaError_t alloc(int64_t nb, int max_len_, int pcap_, int cap_, int nei_cap_, bool wide_, int grid_) {
        max_len = max_len_; pcap = pcap_; cap = cap_; nei_cap = nei_cap_; wide = wide_; grid = grid_;
        const size_t esz = wide ? 32 : 16;
        const int64_t n_lanes = (int64_t)grid * OVLP_BLOCK;
        cudaError_t e;
#define OVS_A(call) if ((e = (call)) != cudaSuccess) return e
        OVS_A(seq.alloc((size_t)nb * max_len)); OVS_A(len.alloc((size_t)nb * 4)); OVS_A(rec.alloc((size_t)nb * OV_NREC * 8));
        OVS_A(ext.alloc((size_t)nb * max_len)); OVS_A(cnt.alloc((size_t)(nb + 1) * 4)); OVS_A(slots.alloc((size_t)nb * nei_cap * 32));
        OVS_A(P0.alloc((size_t)nb * pcap * esz)); OVS_A(S0.alloc((size_t)nb * pcap * (esz / 4))); OVS_A(np0.alloc((size_t)nb * 4));
        OVS_A(A.alloc((size_t)n_lanes * cap * esz)); OVS_A(B.alloc((size_t)n_lanes * cap * esz)); OVS_A(cat.alloc((size_t)n_lanes * cap * 8));
        OVS_A(S.alloc((size_t)n_lanes * cap * 2 * (esz / 4)));
        OVS_A(ord.alloc((size_t)nb * 16));                           // keys and sequence numbers, two buffers each (k_nei_key + radix sort)
        {
            size_t need = 0;
            OVS_A(cub::DeviceRadixSort::SortPairsDescending(nullptr, need, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr, nb, 0, 20));
            ord_tmp_bytes = need + 256;
            OVS_A(ord_tmp.alloc(ord_tmp_bytes));
        }
#undef OVS_A
        return cudaSuccess;
    }
    OverlapArgs args(const fmg_index_s *idx, int min_match, int64_t m) const {
        OverlapArgs O;
        O.ix = idx->view; O.min_match = min_match; O.mode = 0; O.n = m; O.seq = seq.as<uint8_t>(); O.len = len.as<int32_t>(); O.max_len = max_len;
        O.ids = nullptr; O.first = 0; O.step = 1; O.ret = nullptr;
        O.P0 = P0.p; O.S0 = S0.p; O.pcap = pcap; O.np0 = np0.as<int32_t>(); O.A = A.p; O.B = B.p; O.cap = cap; O.cat = cat.as<int32_t>(); O.S = S.p;
        O.rec = rec.as<int64_t>(); O.nei = slots.as<uint4>(); O.nei_cap = nei_cap; O.nei_cnt = cnt.as<uint32_t>();
        O.ext = ext.as<uint8_t>(); O.next = nullptr; O.order = nullptr;
        return O;
    }
};

// CUDA-event durations of the kernels of the last whole-index pass (bench.py's roofline figure for the unitig path)
static std::mutex g_stats_lock;
static double g_pass_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
extern "C" void fmg_overlap_stats(double ms[8]) {
    std::lock_guard<std::mutex> g(g_stats_lock);
    for (int k = 0; k < 8; ++k) ms[k] = g_pass_ms[k];
}

extern "C" void fmg_release_cache(void) { g_pool.release(); }

void fmg_ovcache_destroy(fmg_ovcache_s *p) { delete p; }

// The deferred left check over rows [row_lo, row_hi) of a COMPLETE rank-indexed record array (see k_left_select): selected rows
// are recomputed through all four phases in batches and their OV_LEFT is patched into `pack`.  Returns 0, 1 when a scratch
// capacity was exceeded (the caller re-runs with larger ones), -1 on a CUDA error; *n_left = rows evaluated, *ms = device time.
static int left_fix(const fmg_index_s *idx, int min_match, const OvScratch &S, int64_t nb_max, OvPack *pack, const int64_t *rank_of_row, uint64_t row_lo, uint64_t row_hi,
                    unsigned long long *d_ctrl, unsigned long long *h_ctrl, cudaStream_t st, uint64_t *n_left, double *ms) {
    *n_left = 0; *ms = 0;
    if (row_hi <= row_lo) return 0;
    Dev d_ids, d_ret;
    OV_TRY(d_ids.alloc((row_hi - row_lo) * 8));
    OV_TRY(d_ret.alloc((size_t)nb_max * 8));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    OV_TRY(cudaEventCreate(&e0)); OV_TRY(cudaEventCreate(&e1));
    struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evg{e0, e1};
    OV_TRY(cudaEventRecord(e0, st));
    OV_TRY(cudaMemsetAsync(d_ctrl + OVC_NLEFT, 0, 8, st));
    k_left_select<<<(unsigned)((row_hi - row_lo + 255) / 256), 256, 0, st>>>(pack, rank_of_row, row_lo, row_hi, d_ids.as<uint64_t>(), d_ctrl + OVC_NLEFT);
    ++g_launches;
    OV_TRY(cudaGetLastError());
    OV_TRY(cudaMemcpyAsync(h_ctrl + OVC_NLEFT, d_ctrl + OVC_NLEFT, 8, cudaMemcpyDeviceToHost, st));
    OV_TRY(cudaStreamSynchronize(st));
    const uint64_t n = h_ctrl[OVC_NLEFT];
    *n_left = n;
    for (uint64_t o = 0; o < n; o += (uint64_t)nb_max) {
        const int64_t m = (int64_t)std::min<uint64_t>((uint64_t)nb_max, n - o);
        OverlapArgs O = S.args(idx, min_match, m);
        O.ids = d_ids.as<uint64_t>() + o; O.ret = d_ret.as<int64_t>();
        unsigned long long *c2 = d_ctrl + OVC_NEXT;
        OV_TRY(S.wide ? launch_records<uint64_t>(O, S.grid, c2, st, nullptr, S.ord.as<uint32_t>(), S.ord_tmp.p, S.ord_tmp_bytes)
                      : launch_records<uint32_t>(O, S.grid, c2, st, nullptr, S.ord.as<uint32_t>(), S.ord_tmp.p, S.ord_tmp_bytes));
        OV_TRY(S.wide ? launch_left<uint64_t>(O, S.grid, c2, st) : launch_left<uint32_t>(O, S.grid, c2, st));
        k_left_patch<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(m, S.rec.as<int64_t>(), d_ret.as<int64_t>(), S.cnt.as<uint32_t>(), S.nei_cap, pack, d_ctrl);
        ++g_launches;
        OV_TRY(cudaGetLastError());
    }
    OV_TRY(cudaEventRecord(e1, st));
    OV_TRY(cudaMemcpyAsync(h_ctrl + OVC_FLAGS, d_ctrl + OVC_FLAGS, 8, cudaMemcpyDeviceToHost, st));
    OV_TRY(cudaStreamSynchronize(st));
    float f = 0;
    cudaEventElapsedTime(&f, e0, e1);
    *ms = f;
    return h_ctrl[OVC_FLAGS] ? 1 : 0;
}

// Overlap records of EVERY sequence of the index (fm_retrieve + fm6_is_contained + fm6_get_nei + check_left_simple per
// BWT row, unitig.c:77-204) for the unitig walk.  Rows are processed in batches queued back to back on one stream with
// no host synchronisation in between: phases 1 + 2 -> k_ov_pack scatter the batch into device-resident, rank-indexed
// 64-byte records plus compact ext / spill arrays; the seed sequences of batch b travel to pinned host memory on a second
// stream while batch b+1 computes; the left check follows for the rows where it matters (left_fix).  Overflow of any scratch
// or output capacity is flagged on the device and answered by ONE re-run of the whole pass with larger capacities.
int fmg_overlap_all(const fmg_index_s *idx, int min_match, int max_len, OvHost *out) {
    return out ? fmg_overlap_pass(idx, min_match, max_len, nullptr, out) : -1;
}

static inline uint64_t row_lo_arg(const OvShard *shard) { return shard ? shard->row_lo : 0; }
static inline uint64_t row_hi_arg(const OvShard *shard, uint64_t n_seq) { return shard ? shard->row_hi : n_seq; }

int fmg_overlap_pass(const fmg_index_s *idx, int min_match, int max_len, OvDevice *dev_out, OvHost *out, OvShard *shard) {
    if (!idx || (!out && !dev_out && !shard) || (shard && (out || dev_out))) return -1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] no CUDA device available; libfermi_b200 has no CPU path\n", __func__);
        return -1;
    }
    OV_TRY(cudaSetDevice(idx->device));
    std::lock_guard<std::mutex> ov_guard(idx->ov_lock);
    const uint64_t n_seq = idx->mcnt[1];
    if (max_len <= 0) max_len = (int)((idx->mcnt[0] - n_seq + n_seq - 1) / (n_seq ? n_seq : 1)) + 8;
    max_len = round8(max_len);
    if (!idx->ovc) idx->ovc = new fmg_ovcache_s;
    fmg_ovcache_s &H = *idx->ovc;
    // Rows per batch (even, so that the odd rows of a batch are its local odd rows).  The persistent lanes of the neighbour phase
    // finish a batch at different times -- the cost of a sequence is heavy-tailed, and a block retires only when its slowest lane
    // does -- so a batch should hold many sequences per lane: 2 M rows (21 per lane) left the lanes idle half of the time.  Scratch
    // is ~1.9 KB per row; a batch takes up to a quarter of the free HBM, at most 32 M rows.
    int64_t batch = 1 << 25;
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            const int64_t fit = (int64_t)(free_b / 4 / (size_t)(18 * max_len + 256 + 20 * std::max(8, max_len - min_match + 8)));
            batch = std::max<int64_t>(1 << 20, std::min<int64_t>(batch, fit & ~(int64_t)1));
        }
        if (const char *e = std::getenv("FMG_OV_BATCH")) batch = std::max<int64_t>(2, std::atoll(e) & ~1ll);
    }
    const int64_t nb_max = (int64_t)std::min<uint64_t>(batch, std::max<uint64_t>(row_hi_arg(shard, n_seq) - row_lo_arg(shard), 2));
    const auto t0 = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); };

    cudaStream_t s_run = nullptr, s_copy = nullptr;
    cudaEvent_t run_done[2] = {nullptr, nullptr}, copy_done[2] = {nullptr, nullptr};
    struct Guard {
        cudaStream_t &a, &b; cudaEvent_t *e, *f;
        ~Guard() { for (int k = 0; k < 2; ++k) { if (e[k]) cudaEventDestroy(e[k]); if (f[k]) cudaEventDestroy(f[k]); }
                   if (a) cudaStreamDestroy(a); if (b) cudaStreamDestroy(b); }
    } guard{s_run, s_copy, run_done, copy_done};
    OV_TRY(cudaStreamCreateWithFlags(&s_run, cudaStreamNonBlocking));
    OV_TRY(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
        OV_TRY(cudaEventCreateWithFlags(&run_done[k], cudaEventDisableTiming));
        OV_TRY(cudaEventCreateWithFlags(&copy_done[k], cudaEventDisableTiming));
    }

    int cap = 4 * max_len, nei_cap = 8, pcap_mul = 1;
    // 6 events per batch = [start | retrieve + contained | neighbours | pack | seed rows | end]
    constexpr int kEv = 6;
    std::vector<cudaEvent_t> phase_ev;
    uint64_t ext_cap = std::max<uint64_t>(n_seq * 24, 1 << 20), spill_cap = std::max<uint64_t>(n_seq, 1 << 16);
    const uint64_t row_lo = shard ? shard->row_lo : 0, row_hi = shard ? shard->row_hi : n_seq;
    if (shard) { ext_cap = shard->ext_cap; spill_cap = shard->spill_cap; }
    if (row_lo > row_hi || row_hi > n_seq || (row_lo & 1)) return -1;          // shards start at a read (even row)
    OV_TRY(H.ctrl.need(OVC_N * 8));
    unsigned long long *h_ctrl = static_cast<unsigned long long *>(H.ctrl.p);
    Dev d_pack, d_ret, d_extout, d_spill, d_ctrl, d_odd[2];
    OvScratch S;
    uint64_t n_left = 0;
    double left_ms = 0;
    for (int attempt = 0;; ++attempt) {
        const bool wide = ov_wide(idx, max_len);
        const int per_sm = lists_blocks_per_sm(wide);
        const int grid = (int)std::min<int64_t>((int64_t)idx->n_sm * per_sm, (nb_max + OVLP_BLOCK - 1) / OVLP_BLOCK);
        const int pcap = std::max(8, max_len - min_match + 8) * pcap_mul;
        const uint64_t n_odd_all = n_seq / 2;
        if (!shard) {
            OV_TRY(d_pack.alloc(n_seq * sizeof(OvPack))); OV_TRY(d_ret.alloc(n_seq * 8));
            OV_TRY(d_extout.alloc(ext_cap)); OV_TRY(d_spill.alloc(spill_cap * 32));
        }
        OV_TRY(d_ctrl.alloc(OVC_N * 8));
        // the four output arrays: the pass's own (records by rank), or the caller's shard buffers (records in row order; rank[] is addressed by absolute row)
        OvPack *o_pack = shard ? static_cast<OvPack *>(shard->pack) : d_pack.as<OvPack>();
        int64_t *o_ret = shard ? shard->rank - row_lo : d_ret.as<int64_t>();
        uint8_t *o_ext = shard ? shard->ext : d_extout.as<uint8_t>();
        uint4 *o_spill = shard ? static_cast<uint4 *>(shard->spill) : d_spill.as<uint4>();
        OV_TRY(S.alloc(nb_max, max_len, pcap, cap, nei_cap, wide, grid));
        for (int k = 0; k < 2; ++k) OV_TRY(d_odd[k].alloc((size_t)(nb_max / 2 + 1) * max_len));
        if (out) OV_TRY(H.seq.need(std::max<uint64_t>(n_odd_all, 1) * (uint64_t)max_len));
        OV_TRY(cudaMemsetAsync(d_ctrl.p, 0, OVC_N * 8, s_run));
        // records of sequences that overflow are not written: keep the array defined
        if (!shard) OV_TRY(cudaMemsetAsync(d_pack.p, 0, n_seq * sizeof(OvPack), s_run));
        else if (row_hi > row_lo) OV_TRY(cudaMemsetAsync(o_pack, 0, (row_hi - row_lo) * sizeof(OvPack), s_run));
        int64_t b = 0;
        for (uint64_t row0 = row_lo; row0 < row_hi; row0 += batch, ++b) {
            const int64_t m = (int64_t)std::min<uint64_t>(batch, row_hi - row0);
            const size_t e0 = phase_ev.size();
            phase_ev.resize(e0 + kEv, nullptr);
            for (int k = 0; k < kEv; ++k) OV_TRY(cudaEventCreate(&phase_ev[e0 + k]));
            cudaEvent_t *ev = &phase_ev[e0];
            OverlapArgs O = S.args(idx, min_match, m);
            O.first = row0; O.ret = o_ret + row0;
            unsigned long long *c2 = d_ctrl.as<unsigned long long>() + OVC_NEXT;       // OVC_NEXT, OVC_NEXT2: the work counters
            OV_TRY(wide ? launch_records<uint64_t>(O, grid, c2, s_run, ev, S.ord.as<uint32_t>(), S.ord_tmp.p, S.ord_tmp_bytes)
                        : launch_records<uint32_t>(O, grid, c2, s_run, ev, S.ord.as<uint32_t>(), S.ord_tmp.p, S.ord_tmp_bytes));
            PackArgs P;
            P.n = m; P.rec = S.rec.as<int64_t>(); P.ret = o_ret + row0; P.len = S.len.as<int32_t>(); P.nei_cnt = S.cnt.as<uint32_t>();
            P.nei_slots = S.slots.as<uint4>(); P.nei_cap = nei_cap; P.ext = S.ext.as<uint8_t>(); P.max_len = max_len;
            P.pack = shard ? o_pack + (row0 - row_lo) : o_pack; P.n_seq = n_seq; P.by_row = shard ? 1 : 0; P.ext_out = o_ext; P.ext_cap = ext_cap;
            P.spill_out = o_spill; P.spill_cap = spill_cap; P.ctrl = d_ctrl.as<unsigned long long>();
            k_ov_pack<<<(unsigned)((m + 255) / 256), 256, 0, s_run>>>(P);
            ++g_launches;
            OV_TRY(cudaGetLastError());
            OV_TRY(cudaEventRecord(ev[3], s_run));
            const int64_t n_odd = out ? m / 2 : 0;             // the seed sequences only travel for the host walk
            if (n_odd) {
                OV_TRY(cudaStreamWaitEvent(s_run, copy_done[b & 1], 0));     // the staging buffer of batch b-2 has left the device
                k_seq_odd<<<(unsigned)((n_odd * max_len + 255) / 256), 256, 0, s_run>>>(S.seq.as<uint8_t>(), S.len.as<int32_t>(), max_len, n_odd, d_odd[b & 1].as<uint8_t>());
                ++g_launches;
                OV_TRY(cudaGetLastError());
                OV_TRY(cudaEventRecord(run_done[b & 1], s_run));
                OV_TRY(cudaStreamWaitEvent(s_copy, run_done[b & 1], 0));
                OV_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(H.seq.p) + (row0 / 2) * (uint64_t)max_len, d_odd[b & 1].p, (size_t)n_odd * max_len,
                                       cudaMemcpyDeviceToHost, s_copy));
                OV_TRY(cudaEventRecord(copy_done[b & 1], s_copy));
            }
            OV_TRY(cudaEventRecord(ev[4], s_run));
        }
        OV_TRY(cudaMemcpyAsync(h_ctrl, d_ctrl.p, OVC_N * 8, cudaMemcpyDeviceToHost, s_run));
        OV_TRY(cudaStreamSynchronize(s_run));
        unsigned long long flags = h_ctrl[OVC_FLAGS];
        const unsigned long long too_long = h_ctrl[OVC_MAXLEN];
        // the left check where it decides a link; a shard cannot know (the records of other rows are elsewhere): its caller runs fmg_overlap_left_fix on the merged array
        if (!flags && !too_long && !shard) {
            const int rc = left_fix(idx, min_match, S, nb_max, o_pack, o_ret, row_lo, row_hi, d_ctrl.as<unsigned long long>(), h_ctrl, s_run, &n_left, &left_ms);
            if (rc < 0) return -1;
            if (rc == 1) flags |= OVF_LIST;
        }
        if (!phase_ev.empty()) {
            double tot[kEv] = {0, 0, 0, 0, 0, 0};
            for (size_t e0 = 0; e0 + kEv <= phase_ev.size(); e0 += kEv)
                for (int k = 0; k < 4; ++k) {
                    float ms = 0;
                    cudaEventElapsedTime(&ms, phase_ev[e0 + k], phase_ev[e0 + k + 1]);
                    tot[k] += ms;
                }
            if (fmg_verbose >= 4)
                std::fprintf(stderr, "[M::%s] kernels over %zu batches (ms): retrieve + contained %.1f, neighbours %.1f, pack %.1f, seed rows (+ wait for the copy engine) %.1f; left check of %llu rows %.1f\n",
                             __func__, phase_ev.size() / kEv, tot[0], tot[1], tot[2], tot[3], (unsigned long long)n_left, left_ms);
            {
                std::lock_guard<std::mutex> g(g_stats_lock);
                g_pass_ms[0] = 0; g_pass_ms[1] = tot[0]; g_pass_ms[2] = tot[1]; g_pass_ms[3] = left_ms; g_pass_ms[4] = (double)n_left;
                g_pass_ms[5] = tot[2]; g_pass_ms[6] = tot[3]; g_pass_ms[7] = (double)(phase_ev.size() / kEv);
            }
            for (cudaEvent_t e : phase_ev) cudaEventDestroy(e);
            phase_ev.clear();
        }
        if (fmg_verbose >= 4)
            std::fprintf(stderr, "[M::%s] attempt %d: %lld batches, %.3f s; ext %llu B, spill %llu, flags %llx, longest clipped %llu\n", __func__, attempt,
                         (long long)b, since(t0), h_ctrl[OVC_EXT], h_ctrl[OVC_SPILL], flags, too_long);
        if (!flags && !too_long) break;
        OV_TRY(cudaStreamSynchronize(s_copy));
        if (attempt == 6 || (flags & (OVF_FIELD | OVF_RANK))) {
            if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] cannot represent the overlap records (flags %llx, cap=%d, nei_cap=%d)\n", __func__, flags, cap, nei_cap);
            return -1;
        }
        if (shard && (flags & (OVF_EXT | OVF_SPILL))) {        // the caller owns these buffers: report the need
            shard->ext_total = h_ctrl[OVC_EXT] + (h_ctrl[OVC_EXT] >> 3) + 4096; shard->spill_total = h_ctrl[OVC_SPILL] + (h_ctrl[OVC_SPILL] >> 3) + 256;
            return 1;
        }
        if (too_long) { max_len = round8((int)too_long + 8); cap = std::max(cap, 4 * max_len); }
        if (flags & OVF_LIST) cap *= 4, pcap_mul *= 4;
        if (flags & (OVF_LIST | OVF_NEI)) nei_cap *= 4;
        // the totals keep counting past the capacity, so they are the true need unless other sequences were skipped
        if (flags & OVF_EXT) ext_cap = std::max<uint64_t>(2 * ext_cap, h_ctrl[OVC_EXT] + (h_ctrl[OVC_EXT] >> 3));
        if (flags & OVF_SPILL) spill_cap = std::max<uint64_t>(2 * spill_cap, h_ctrl[OVC_SPILL] + (h_ctrl[OVC_SPILL] >> 3));
        if (fmg_verbose >= 3)
            std::fprintf(stderr, "[M::%s] capacity exceeded (flags %llx); re-running with max_len=%d, %d candidate / %d neighbour slots, %llu ext bytes, %llu spill entries\n",
                         __func__, flags, max_len, cap, nei_cap, (unsigned long long)ext_cap, (unsigned long long)spill_cap);
    }
    const uint64_t ext_total = h_ctrl[OVC_EXT], spill_total = h_ctrl[OVC_SPILL];
    if (out) {
        OV_TRY(H.pack.need(std::max<uint64_t>(n_seq, 1) * sizeof(OvPack))); OV_TRY(H.rank.need(std::max<uint64_t>(n_seq, 1) * 8));
        OV_TRY(H.ext.need(std::max<uint64_t>(ext_total, 1))); OV_TRY(H.spill.need(std::max<uint64_t>(spill_total, 1) * 32));
        if (n_seq) {
            OV_TRY(cudaMemcpyAsync(H.pack.p, d_pack.p, n_seq * sizeof(OvPack), cudaMemcpyDeviceToHost, s_run));
            OV_TRY(cudaMemcpyAsync(H.rank.p, d_ret.p, n_seq * 8, cudaMemcpyDeviceToHost, s_run));
        }
        if (ext_total) OV_TRY(cudaMemcpyAsync(H.ext.p, d_extout.p, ext_total, cudaMemcpyDeviceToHost, s_run));
        if (spill_total) OV_TRY(cudaMemcpyAsync(H.spill.p, d_spill.p, spill_total * 32, cudaMemcpyDeviceToHost, s_run));
        OV_TRY(cudaStreamSynchronize(s_run));
        OV_TRY(cudaStreamSynchronize(s_copy));
        out->n_seq = n_seq; out->max_len = max_len;
        out->pack = static_cast<const OvPack *>(H.pack.p); out->rank_of_row = static_cast<const uint64_t *>(H.rank.p);
        out->seq = static_cast<const uint8_t *>(H.seq.p); out->seq_stride = (uint64_t)max_len; out->seq_odd_only = 1;
        out->ext = static_cast<const uint8_t *>(H.ext.p); out->spill = static_cast<const fmg_intv_t *>(H.spill.p);
        out->ext_total = ext_total; out->spill_total = spill_total;
    }
    if (shard) { shard->ext_total = ext_total; shard->spill_total = spill_total; shard->max_len = max_len; }
    if (dev_out) {
        dev_out->pack.swap(d_pack); dev_out->rank.swap(d_ret); dev_out->ext.swap(d_extout); dev_out->spill.swap(d_spill);
        dev_out->n_seq = n_seq; dev_out->ext_total = ext_total; dev_out->spill_total = spill_total; dev_out->max_len = max_len;
    }
    if (fmg_verbose >= 4) std::fprintf(stderr, "[M::%s] %llu sequences, records %s after %.3f s\n", __func__, (unsigned long long)n_seq, out ? "on the host" : "in HBM", since(t0));
    return 0;
}

// the left check over ALL rows of a merged, rank-indexed record array in caller-owned device memory (multi-GPU path: every rank
// runs it on its copy after the exchange; fmg_overlap_pass does the same internally for a single GPU)
// OV_LEFT of rows [row_lo, row_lo + n) read out of (apply = 0) or written into (apply = 1) the rank-indexed records: the exchange of the
// sharded left fix
__global__ void __launch_bounds__(256) k_left_flags(OvPack *pack, const int64_t *__restrict__ rank_of_row, uint64_t row_lo, uint64_t n, int8_t *flags, int apply) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    OvPack &p = pack[rank_of_row[row_lo + i]];
    if (apply) p.left = flags[i];
    else flags[i] = p.left;
}
int fmg_overlap_left_flags_dev(const fmg_index_s *idx, void *d_pack, const int64_t *d_rank_of_row, uint64_t row_lo, uint64_t row_hi, int8_t *d_flags, int apply) {
    if (!idx || !d_pack || !d_rank_of_row || !d_flags || row_hi < row_lo || row_hi > idx->mcnt[1]) return -1;
    OV_TRY(cudaSetDevice(idx->device));
    const uint64_t n = row_hi - row_lo;
    if (n == 0) return 0;
    k_left_flags<<<(unsigned)((n + 255) / 256), 256>>>(static_cast<OvPack *>(d_pack), d_rank_of_row, row_lo, n, d_flags, apply);
    ++g_launches;
    OV_TRY(cudaGetLastError());
    return 0;
}

int fmg_overlap_left_fix_dev(const fmg_index_s *idx, int min_match, int max_len, void *d_pack, const int64_t *d_rank_of_row, uint64_t row_lo, uint64_t row_hi, uint64_t *n_left_out) {
    if (!idx || !d_pack || !d_rank_of_row || row_hi < row_lo || row_hi > idx->mcnt[1]) return -1;
    OV_TRY(cudaSetDevice(idx->device));
    std::lock_guard<std::mutex> ov_guard(idx->ov_lock);
    const uint64_t n_seq = idx->mcnt[1];
    if (max_len <= 0) max_len = (int)((idx->mcnt[0] - n_seq + n_seq - 1) / (n_seq ? n_seq : 1)) + 8;
    max_len = round8(max_len);
    if (!idx->ovc) idx->ovc = new fmg_ovcache_s;
    fmg_ovcache_s &H = *idx->ovc;
    OV_TRY(H.ctrl.need(OVC_N * 8));
    unsigned long long *h_ctrl = static_cast<unsigned long long *>(H.ctrl.p);
    const int64_t nb_max = (int64_t)std::min<uint64_t>(1 << 21, n_seq ? n_seq : 1);
    Dev d_ctrl;
    OV_TRY(d_ctrl.alloc(OVC_N * 8));
    int cap = 4 * max_len, nei_cap = 8, pcap_mul = 1;
    uint64_t n_left = 0;
    for (int attempt = 0;; ++attempt) {
        const bool wide = ov_wide(idx, max_len);
        const int per_sm = lists_blocks_per_sm(wide);
        const int grid = (int)std::min<int64_t>((int64_t)idx->n_sm * per_sm, (nb_max + OVLP_BLOCK - 1) / OVLP_BLOCK);
        OvScratch S;
        OV_TRY(S.alloc(nb_max, max_len, std::max(8, max_len - min_match + 8) * pcap_mul, cap, nei_cap, wide, grid));
        OV_TRY(cudaMemset(d_ctrl.p, 0, OVC_N * 8));
        double ms = 0;
        const int rc = left_fix(idx, min_match, S, nb_max, static_cast<OvPack *>(d_pack), d_rank_of_row, row_lo, row_hi, d_ctrl.as<unsigned long long>(), h_ctrl, nullptr, &n_left, &ms);
        if (rc < 0) return -1;
        {
            std::lock_guard<std::mutex> g(g_stats_lock);
            g_pass_ms[3] = ms; g_pass_ms[4] = (double)n_left;
        }
        if (rc == 0) break;
        if (attempt == 6) return -1;
        cap *= 4; pcap_mul *= 4; nei_cap *= 4;
    }
    if (n_left_out) *n_left_out = n_left;
    return 0;
}

extern "C" {

// fm6_seqsort (seqsort.c:37-70) / `fermi seqrank`: sorted[mcnt[1]] as the reference fills it; stats = #zeros, #contained, #duplicates
int fmg_seqsort(const fmg_index_t *idx, uint64_t *sorted, int64_t stats[3]) {
    if (!idx || !sorted) return -1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] no CUDA device available; libfermi_b200 has no CPU path\n", __func__);
        return -1;
    }
    OV_TRY(cudaSetDevice(idx->device));
    const uint64_t n_seq = idx->mcnt[1];
    const bool wide = ov_wide(idx, 0);
    Dev d_sorted, d_rec, d_ret, d_cnt, d_np0, d_len;
    OV_TRY(d_sorted.alloc(std::max<uint64_t>(n_seq, 1) * 8));
    OV_TRY(cudaMemset(d_sorted.p, 0, std::max<uint64_t>(n_seq, 1) * 8));         // calloc in the reference (seqsort.c:49)
    const int64_t batch = 1 << 22;
    const int64_t n_even = (int64_t)((n_seq + 1) / 2), nb_max = std::min<int64_t>(batch, std::max<int64_t>(n_even, 1));
    OV_TRY(d_rec.alloc((size_t)nb_max * OV_NREC * 8)); OV_TRY(d_ret.alloc((size_t)nb_max * 8)); OV_TRY(d_cnt.alloc((size_t)nb_max * 4));
    OV_TRY(d_np0.alloc((size_t)nb_max * 4)); OV_TRY(d_len.alloc((size_t)nb_max * 4));
    for (int64_t t0 = 0; t0 < n_even; t0 += batch) {
        const int64_t m = std::min(batch, n_even - t0);
        OverlapArgs O;
        std::memset(&O, 0, sizeof O);
        O.ix = idx->view; O.min_match = 0; O.mode = 1; O.n = m; O.max_len = 0;
        O.ids = nullptr; O.first = 2 * (uint64_t)t0; O.step = 2; O.ret = d_ret.as<int64_t>(); O.len = d_len.as<int32_t>();
        O.np0 = d_np0.as<int32_t>(); O.rec = d_rec.as<int64_t>(); O.nei_cnt = d_cnt.as<uint32_t>();
        const unsigned gch = (unsigned)((m + OVCH_BLOCK - 1) / OVCH_BLOCK);
        if (wide) k_ov_chain<uint64_t, 1><<<gch, OVCH_BLOCK>>>(O); else k_ov_chain<uint32_t, 1><<<gch, OVCH_BLOCK>>>(O);
        k_seqsort_scatter<<<(unsigned)((m + 255) / 256), 256>>>(m, 2 * (uint64_t)t0, d_rec.as<int64_t>(), d_ret.as<int64_t>(), d_sorted.as<uint64_t>());
        g_launches += 2;
        OV_TRY(cudaGetLastError());
    }
    OV_TRY(cudaMemcpy(sorted, d_sorted.p, n_seq * 8, cudaMemcpyDeviceToHost));
    if (stats) {                                                                    // the tally fm6_seqsort prints (seqsort.c:61-68)
        stats[0] = stats[1] = stats[2] = 0;
        for (uint64_t i = 0; i < n_seq; ++i)
            if (sorted[i] == 0) ++stats[0];
            else if (sorted[i] & 2) ++stats[1];
            else if (sorted[i] & 1) ++stats[2];
        if (fmg_verbose >= 3)
            std::fprintf(stderr, "[M::%s] #zeros=%ld, #contained=%ld, #duplicates=%ld\n", __func__, (long)stats[0], (long)stats[1], (long)stats[2]);
    }
    return 0;
}

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
    const int ml = round8(max_len);                              // row stride on the device
    const bool wide = ov_wide(idx, ml);

    Dev d_ids, d_ret, d_mem, d_off, d_tiles, d_ctrl;
    OV_TRY(d_ret.alloc((size_t)n * 8));
    OV_TRY(d_off.alloc((size_t)(n + 1) * 8)); OV_TRY(d_tiles.alloc((size_t)fmg_compact_tiles(n) * 8)); OV_TRY(d_ctrl.alloc(64));
    if (ids) { OV_TRY(d_ids.alloc((size_t)n * 8)); OV_TRY(cudaMemcpy(d_ids.p, ids, (size_t)n * 8, cudaMemcpyHostToDevice)); }

    const auto t0 = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); };
    // ---- overlap records (all four phases for every sequence); scratch capacities grow until nothing overflows
    const int per_sm = lists_blocks_per_sm(wide);
    const int grid = (int)std::min<int64_t>((int64_t)idx->n_sm * per_sm, (n + OVLP_BLOCK - 1) / OVLP_BLOCK);
    int cap = 4 * ml, nei_cap = 8, pcap = std::max(8, ml - min_match + 8);
    int64_t *h_rec = rec;
    std::vector<uint32_t> h_cnt(n);
    std::vector<int32_t> h_len(n);
    OvScratch S;
    for (int attempt = 0;; ++attempt) {
        OV_TRY(S.alloc(n, ml, pcap, cap, nei_cap, wide, grid));
        OV_TRY(d_mem.alloc((size_t)n * nei_cap * 32));
        OverlapArgs O = S.args(idx, min_match, n);
        O.ids = ids ? d_ids.as<uint64_t>() : nullptr; O.first = first; O.step = step; O.ret = d_ret.as<int64_t>();
        OV_TRY(wide ? launch_records<uint64_t>(O, grid, d_ctrl.as<unsigned long long>(), nullptr) : launch_records<uint32_t>(O, grid, d_ctrl.as<unsigned long long>(), nullptr));
        OV_TRY(wide ? launch_left<uint64_t>(O, grid, d_ctrl.as<unsigned long long>(), nullptr) : launch_left<uint32_t>(O, grid, d_ctrl.as<unsigned long long>(), nullptr));
        OV_TRY(cudaDeviceSynchronize());
        if (fmg_verbose >= 4) std::fprintf(stderr, "[M::%s] overlap phases (attempt %d) %.3f s for %lld sequences\n", __func__, attempt, since(t0), (long long)n);
        OV_TRY(cudaMemcpy(h_len.data(), S.len.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < n; ++i)
            if (h_len[i] < 0 || h_len[i] > max_len) {
                if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] a sequence of %d bases exceeds max_len=%d\n", __func__, h_len[i] < 0 ? -h_len[i] : h_len[i], max_len);
                return 2;
            }
        OV_TRY(cudaMemcpy(h_rec, S.rec.p, (size_t)n * OV_NREC * 8, cudaMemcpyDeviceToHost));
        OV_TRY(cudaMemcpy(h_cnt.data(), S.cnt.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
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
        if (list_ovf) cap *= 4, pcap *= 4;
        if (nei_ovf || list_ovf) nei_cap *= 4;
        if (fmg_verbose >= 3) std::fprintf(stderr, "[M::%s] scratch overflow; re-running with %d candidate / %d neighbour slots\n", __func__, cap, nei_cap);
    }
    // the value fm_retrieve returned goes to rec[OV_K]
    std::vector<int64_t> h_ret(n);
    OV_TRY(cudaMemcpy(h_ret.data(), d_ret.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; ++i) h_rec[i * OV_NREC + OV_K] = h_ret[i];

    // ---- neighbour slots -> dense array + offsets
    if (fmg_compact_slots(S.cnt.as<uint32_t>(), n, nei_cap, S.slots.as<uint4>(), d_mem.as<uint4>(), d_off.as<uint64_t>(),
                          d_tiles.as<uint64_t>(), d_ctrl.as<unsigned long long>(), nullptr)) return -1;
    OV_TRY(cudaMemcpy(nei_off, d_off.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost));
    const uint64_t tot = nei_off[n];
    *nei = (fmg_intv_t *)std::malloc((tot ? tot : 1) * 32);
    if (tot) OV_TRY(cudaMemcpy(*nei, d_mem.p, tot * 32, cudaMemcpyDeviceToHost));
    if (seq || ext) {                                            // device rows: stride ml, sequences right-aligned -> caller's n x max_len, left-aligned
        std::vector<uint8_t> h((size_t)n * ml);
        if (seq) {
            OV_TRY(cudaMemcpy(h.data(), S.seq.p, (size_t)n * ml, cudaMemcpyDeviceToHost));
            std::memset(seq, 0, (size_t)n * max_len);
            for (int64_t i = 0; i < n; ++i) std::memcpy(seq + (size_t)i * max_len, h.data() + (size_t)(i + 1) * ml - h_len[i], (size_t)h_len[i]);
        }
        if (ext) {
            OV_TRY(cudaMemcpy(h.data(), S.ext.p, (size_t)n * ml, cudaMemcpyDeviceToHost));
            for (int64_t i = 0; i < n; ++i) std::memcpy(ext + (size_t)i * max_len, h.data() + (size_t)i * ml, (size_t)max_len);
        }
    }
    if (len) std::memcpy(len, h_len.data(), (size_t)n * 4);
    if (fmg_verbose >= 4) std::fprintf(stderr, "[M::%s] batch total %.3f s\n", __func__, since(t0));
    return 0;
}

} // extern "C"
