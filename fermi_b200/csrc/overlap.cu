// Overlap path of `fermi unitig` on the GPU: kernels + C-ABI (see fmd_overlap.cuh for the algorithm).
//   k_ov_chain<U,1|3>  fm_retrieve (exact.c:59-70) fused with fm6_is_contained / the overlap_intv of check_left_simple (unitig.c:77-91,186-190), one sequence per thread
//   k_ov_lists<U,2|4>  fm6_get_nei / the candidate loop of check_left_simple (unitig.c:93-179,191-203), persistent lanes
//   k_ov_pack          batch records -> rank-indexed 64-byte records + compact ext / spill arrays (whole-index pass)
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

// phases 1 and 3: one sequence per thread, a straight chain of extensions (converged for equal-length reads)
template <typename U, int PHASE>
__global__ void __launch_bounds__(OVCH_BLOCK) k_ov_chain(OverlapArgs A) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < A.n) overlap_chain<U, PHASE>(A, t);
}

// phases 2 and 4: persistent lanes, sequences handed out by atomicAdd, every extension behind a warp vote
template <typename U, int PHASE>
__global__ void __launch_bounds__(OVLP_BLOCK, OVLP_MIN_BLOCKS) k_ov_lists(const __grid_constant__ OverlapArgs A) {     // &A.ix is taken: no stack copy
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    overlap_lane_sync<U, PHASE>(A, lane, [&]() -> int64_t { return (int64_t)atomicAdd(A.next, 1ull); });
}

// ---- whole-index pass (fmg_overlap_all): per-batch records -> rank-indexed packed records + compact ext / spill arrays
enum { OVC_NEXT = 0, OVC_NEXT2, OVC_EXT, OVC_SPILL, OVC_FLAGS, OVC_MAXLEN, OVC_N };
enum { OVF_LIST = 1, OVF_NEI = 2, OVF_EXT = 4, OVF_SPILL = 8, OVF_FIELD = 16, OVF_RANK = 32 };

struct PackArgs {
    int64_t n;                          // sequences (BWT rows) of this batch
    const int64_t *rec;                 // n x OV_NREC
    const int64_t *ret;                 // n: fm_retrieve's return value = rank of the sequence
    const int32_t *len;                 // n: < 0 flags a sequence longer than max_len
    const uint32_t *nei_cnt;
    const uint4 *nei_slots; int nei_cap;
    const uint8_t *ext; int max_len;
    OvPack *pack; uint64_t n_seq;
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
            uint4 *dst = reinterpret_cast<uint4 *>(A.pack + k);
            const uint4 *src = reinterpret_cast<const uint4 *>(&o);
            dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
        }
    }
    if (flags) atomicOr(A.ctrl + OVC_FLAGS, (unsigned long long)flags);
}

// the sequences of the odd rows (the seeds of unitig_core, unitig.c:333-334) of a batch whose first row is even
__global__ void __launch_bounds__(256) k_seq_odd(const uint8_t *__restrict__ seq, int max_len, int64_t n_odd, uint8_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_odd * max_len) return;
    const int64_t row = i / max_len, col = i - row * max_len;
    out[i] = seq[(2 * row + 1) * max_len + col];
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

// the four phases of the overlap record over one batch, back to back on `st`; ctrl2 = two work counters (zeroed here)
template <typename U>
static cudaError_t launch_phases(OverlapArgs O, int grid, unsigned long long *ctrl2, cudaStream_t st, cudaEvent_t *ev = nullptr) {
    cudaError_t e = cudaMemsetAsync(ctrl2, 0, 16, st);
    if (e != cudaSuccess) return e;
    const unsigned gch = (unsigned)((O.n + OVCH_BLOCK - 1) / OVCH_BLOCK);
    const int g = (int)std::min<int64_t>(grid, (O.n + OVLP_BLOCK - 1) / OVLP_BLOCK);
    if (ev) cudaEventRecord(ev[0], st);
    k_ov_chain<U, 1><<<gch, OVCH_BLOCK, 0, st>>>(O);
    if (ev) cudaEventRecord(ev[1], st);
    O.next = ctrl2;
    k_ov_lists<U, 2><<<g, OVLP_BLOCK, 0, st>>>(O);
    if (ev) cudaEventRecord(ev[2], st);
    k_ov_chain<U, 3><<<gch, OVCH_BLOCK, 0, st>>>(O);
    if (ev) cudaEventRecord(ev[3], st);
    O.next = ctrl2 + 1;
    k_ov_lists<U, 4><<<g, OVLP_BLOCK, 0, st>>>(O);
    if (ev) cudaEventRecord(ev[4], st);
    g_launches += 4;
    return cudaGetLastError();
}

static int lists_blocks_per_sm(bool wide) {
    int a = 0, b = 0;
    if (wide) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_ov_lists<uint64_t, 2>, OVLP_BLOCK, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_ov_lists<uint64_t, 4>, OVLP_BLOCK, 0);
    } else {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_ov_lists<uint32_t, 2>, OVLP_BLOCK, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_ov_lists<uint32_t, 4>, OVLP_BLOCK, 0);
    }
    return std::max(1, std::max(a, b));       // lane scratch is sized for the larger grid
}

namespace fmg { Pool g_pool; }

// CUDA-event durations of the kernels of the last whole-index pass (bench.py's roofline figure for the unitig path)
static std::mutex g_stats_lock;
static double g_pass_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
extern "C" void fmg_overlap_stats(double ms[8]) {
    std::lock_guard<std::mutex> g(g_stats_lock);
    for (int k = 0; k < 8; ++k) ms[k] = g_pass_ms[k];
}

extern "C" void fmg_release_cache(void) { g_pool.release(); }

void fmg_ovcache_destroy(fmg_ovcache_s *p) { delete p; }

// Overlap records of EVERY sequence of the index (fm_retrieve + fm6_is_contained + fm6_get_nei + check_left_simple per
// BWT row, unitig.c:77-204) for the unitig walk.  Rows are processed in batches queued back to back on one stream with
// no host synchronisation in between: the four overlap phases -> k_ov_pack scatter the batch into device-resident,
// rank-indexed 64-byte records plus compact ext / spill arrays; the seed sequences of batch b travel to pinned host
// memory on a second stream while batch b+1 computes.  Overflow of any scratch or output capacity is flagged on the
// device and answered by ONE re-run of the whole pass with larger capacities.
int fmg_overlap_all(const fmg_index_s *idx, int min_match, int max_len, OvHost *out) {
    return out ? fmg_overlap_pass(idx, min_match, max_len, nullptr, out) : -1;
}

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
    if (!idx->ovc) idx->ovc = new fmg_ovcache_s;
    fmg_ovcache_s &H = *idx->ovc;
    const bool wide = idx->view.n_sym + 256 >= (1ull << 32) || std::getenv("FMG_FORCE_WIDE") != nullptr;
    const int per_sm = lists_blocks_per_sm(wide);
    const int64_t batch = 1 << 21;                           // even, so that the odd rows of a batch are its local odd rows
    const int64_t nb_max = (int64_t)std::min<uint64_t>(batch, n_seq ? n_seq : 1);
    const int grid = (int)std::min<int64_t>((int64_t)idx->n_sm * per_sm, (nb_max + OVLP_BLOCK - 1) / OVLP_BLOCK);
    const int64_t n_lanes = (int64_t)grid * OVLP_BLOCK;
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
    // fmg_verbose >= 4: 8 events per batch = [start | (unused) | retrieve + contained | neighbours | left chain | left lists | pack | seed rows]
    std::vector<cudaEvent_t> phase_ev;
    uint64_t ext_cap = std::max<uint64_t>(n_seq * 24, 1 << 20), spill_cap = std::max<uint64_t>(n_seq, 1 << 16);
    const uint64_t row_lo = shard ? shard->row_lo : 0, row_hi = shard ? shard->row_hi : n_seq;
    if (shard) { ext_cap = shard->ext_cap; spill_cap = shard->spill_cap; }
    if (row_lo > row_hi || row_hi > n_seq || (row_lo & 1)) return -1;          // shards start at a read (even row)
    OV_TRY(H.ctrl.need(OVC_N * 8));
    unsigned long long *h_ctrl = static_cast<unsigned long long *>(H.ctrl.p);
    Dev d_pack, d_ret, d_extout, d_spill, d_ctrl, d_seq, d_len, d_rec, d_ext, d_cnt, d_slots, d_P0, d_np0, d_A, d_B, d_cat, d_odd[2];
    for (int attempt = 0;; ++attempt) {
        const int pcap = std::max(8, max_len - min_match + 8) * pcap_mul;
        const size_t esz = wide ? 32 : 16;
        const uint64_t n_odd_all = n_seq / 2;
        if (!shard) {
            OV_TRY(d_pack.alloc(n_seq * sizeof(OvPack))); OV_TRY(d_ret.alloc(n_seq * 8));
            OV_TRY(d_extout.alloc(ext_cap)); OV_TRY(d_spill.alloc(spill_cap * 32));
        }
        OV_TRY(d_ctrl.alloc(OVC_N * 8));
        // the four output arrays: the pass's own, or the caller's shard buffers (rank[] is addressed by absolute row)
        OvPack *o_pack = shard ? static_cast<OvPack *>(shard->pack) : d_pack.as<OvPack>();
        int64_t *o_ret = shard ? shard->rank - row_lo : d_ret.as<int64_t>();
        uint8_t *o_ext = shard ? shard->ext : d_extout.as<uint8_t>();
        uint4 *o_spill = shard ? static_cast<uint4 *>(shard->spill) : d_spill.as<uint4>();
        OV_TRY(d_seq.alloc((size_t)nb_max * max_len)); OV_TRY(d_len.alloc((size_t)nb_max * 4)); OV_TRY(d_rec.alloc((size_t)nb_max * OV_NREC * 8));
        OV_TRY(d_ext.alloc((size_t)nb_max * max_len)); OV_TRY(d_cnt.alloc((size_t)(nb_max + 1) * 4)); OV_TRY(d_slots.alloc((size_t)nb_max * nei_cap * 32));
        OV_TRY(d_P0.alloc((size_t)nb_max * pcap * esz)); OV_TRY(d_np0.alloc((size_t)nb_max * 4));
        OV_TRY(d_A.alloc((size_t)n_lanes * cap * esz)); OV_TRY(d_B.alloc((size_t)n_lanes * cap * esz));
        OV_TRY(d_cat.alloc((size_t)n_lanes * cap * 8));
        for (int k = 0; k < 2; ++k) OV_TRY(d_odd[k].alloc((size_t)(nb_max / 2 + 1) * max_len));
        if (out) OV_TRY(H.seq.need(std::max<uint64_t>(n_odd_all, 1) * (uint64_t)max_len));
        OV_TRY(cudaMemsetAsync(d_ctrl.p, 0, OVC_N * 8, s_run));
        // records of sequences that overflow are not written: keep the array defined
        if (!shard) OV_TRY(cudaMemsetAsync(d_pack.p, 0, n_seq * sizeof(OvPack), s_run));
        int64_t b = 0;
        for (uint64_t row0 = row_lo; row0 < row_hi; row0 += batch, ++b) {
            const int64_t m = (int64_t)std::min<uint64_t>(batch, row_hi - row0);
            cudaEvent_t *ev = nullptr;
            {
                const size_t e0 = phase_ev.size();
                phase_ev.resize(e0 + 8, nullptr);
                for (int k = 0; k < 8; ++k) OV_TRY(cudaEventCreate(&phase_ev[e0 + k]));
                ev = &phase_ev[e0];
                OV_TRY(cudaEventRecord(ev[0], s_run));
            }
            OverlapArgs O;
            O.ix = idx->view; O.min_match = min_match; O.mode = 0; O.n = m; O.seq = d_seq.as<uint8_t>(); O.len = d_len.as<int32_t>(); O.max_len = max_len;
            O.ids = nullptr; O.first = row0; O.step = 1; O.ret = o_ret + row0;
            O.P0 = d_P0.p; O.pcap = pcap; O.np0 = d_np0.as<int32_t>(); O.A = d_A.p; O.B = d_B.p; O.cap = cap; O.cat = d_cat.as<int32_t>();
            O.rec = d_rec.as<int64_t>(); O.nei = d_slots.as<uint4>(); O.nei_cap = nei_cap; O.nei_cnt = d_cnt.as<uint32_t>();
            O.ext = d_ext.as<uint8_t>(); O.next = nullptr;
            unsigned long long *c2 = d_ctrl.as<unsigned long long>() + OVC_NEXT;       // OVC_NEXT, OVC_NEXT2: the work counters
            OV_TRY(wide ? launch_phases<uint64_t>(O, grid, c2, s_run, ev ? ev + 1 : nullptr) : launch_phases<uint32_t>(O, grid, c2, s_run, ev ? ev + 1 : nullptr));
            PackArgs P;
            P.n = m; P.rec = d_rec.as<int64_t>(); P.ret = o_ret + row0; P.len = d_len.as<int32_t>(); P.nei_cnt = d_cnt.as<uint32_t>();
            P.nei_slots = d_slots.as<uint4>(); P.nei_cap = nei_cap; P.ext = d_ext.as<uint8_t>(); P.max_len = max_len;
            P.pack = o_pack; P.n_seq = n_seq; P.ext_out = o_ext; P.ext_cap = ext_cap;
            P.spill_out = o_spill; P.spill_cap = spill_cap; P.ctrl = d_ctrl.as<unsigned long long>();
            k_ov_pack<<<(unsigned)((m + 255) / 256), 256, 0, s_run>>>(P);
            ++g_launches;
            OV_TRY(cudaGetLastError());
            if (ev) OV_TRY(cudaEventRecord(ev[6], s_run));
            const int64_t n_odd = out ? m / 2 : 0;             // the seed sequences only travel for the host walk
            if (n_odd) {
                OV_TRY(cudaStreamWaitEvent(s_run, copy_done[b & 1], 0));     // the staging buffer of batch b-2 has left the device
                k_seq_odd<<<(unsigned)((n_odd * max_len + 255) / 256), 256, 0, s_run>>>(d_seq.as<uint8_t>(), max_len, n_odd, d_odd[b & 1].as<uint8_t>());
                ++g_launches;
                OV_TRY(cudaGetLastError());
                OV_TRY(cudaEventRecord(run_done[b & 1], s_run));
                OV_TRY(cudaStreamWaitEvent(s_copy, run_done[b & 1], 0));
                OV_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(H.seq.p) + (row0 / 2) * (uint64_t)max_len, d_odd[b & 1].p, (size_t)n_odd * max_len,
                                       cudaMemcpyDeviceToHost, s_copy));
                OV_TRY(cudaEventRecord(copy_done[b & 1], s_copy));
            }
            if (ev) OV_TRY(cudaEventRecord(ev[7], s_run));
        }
        OV_TRY(cudaMemcpyAsync(h_ctrl, d_ctrl.p, OVC_N * 8, cudaMemcpyDeviceToHost, s_run));
        OV_TRY(cudaStreamSynchronize(s_run));
        const unsigned long long flags = h_ctrl[OVC_FLAGS], too_long = h_ctrl[OVC_MAXLEN];
        if (!phase_ev.empty()) {
            double tot[7] = {0, 0, 0, 0, 0, 0, 0};
            for (size_t e0 = 0; e0 + 8 <= phase_ev.size(); e0 += 8)
                for (int k = 0; k < 7; ++k) {
                    float ms = 0;
                    cudaEventElapsedTime(&ms, phase_ev[e0 + k], phase_ev[e0 + k + 1]);
                    tot[k] += ms;
                }
            if (fmg_verbose >= 4)
                std::fprintf(stderr, "[M::%s] kernels over %zu batches (ms): memset %.1f, retrieve + contained %.1f, neighbours %.1f, left chain %.1f, left lists %.1f, pack %.1f, seed rows (+ wait for the copy engine) %.1f\n",
                             __func__, phase_ev.size() / 8, tot[0], tot[1], tot[2], tot[3], tot[4], tot[5], tot[6]);
            {
                std::lock_guard<std::mutex> g(g_stats_lock);
                for (int k = 0; k < 7; ++k) g_pass_ms[k] = tot[k];
                g_pass_ms[7] = (double)(phase_ev.size() / 8);
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
        if (too_long) { max_len = (int)too_long + 8; cap = std::max(cap, 4 * max_len); }
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
    const bool wide = idx->view.n_sym + 256 >= (1ull << 32) || std::getenv("FMG_FORCE_WIDE") != nullptr;
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
    const bool wide = idx->view.n_sym + 256 >= (1ull << 32) || std::getenv("FMG_FORCE_WIDE") != nullptr;

    Dev d_ids, d_seq, d_len, d_ret, d_rec, d_ext, d_cnt, d_slots, d_mem, d_off, d_tiles, d_ctrl, d_A, d_B, d_cat;
    OV_TRY(d_seq.alloc((size_t)n * max_len)); OV_TRY(d_len.alloc((size_t)n * 4)); OV_TRY(d_ret.alloc((size_t)n * 8));
    OV_TRY(d_rec.alloc((size_t)n * OV_NREC * 8)); OV_TRY(d_ext.alloc((size_t)n * max_len)); OV_TRY(d_cnt.alloc((size_t)(n + 1) * 4));
    OV_TRY(d_off.alloc((size_t)(n + 1) * 8)); OV_TRY(d_tiles.alloc((size_t)fmg_compact_tiles(n) * 8)); OV_TRY(d_ctrl.alloc(64));
    if (ids) { OV_TRY(d_ids.alloc((size_t)n * 8)); OV_TRY(cudaMemcpy(d_ids.p, ids, (size_t)n * 8, cudaMemcpyHostToDevice)); }

    const auto t0 = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); };
    const auto t1 = std::chrono::steady_clock::now();
    // ---- overlap records; scratch capacities grow until nothing overflows
    const int per_sm = lists_blocks_per_sm(wide);
    const int grid = (int)std::min<int64_t>((int64_t)idx->n_sm * per_sm, (n + OVLP_BLOCK - 1) / OVLP_BLOCK);
    const int64_t n_lanes = (int64_t)grid * OVLP_BLOCK;
    int cap = 4 * max_len, nei_cap = 8, pcap = std::max(8, max_len - min_match + 8);
    int64_t *h_rec = rec;
    std::vector<uint32_t> h_cnt(n);
    Dev d_P0, d_np0;
    std::vector<int32_t> h_len(n);
    OV_TRY(d_np0.alloc((size_t)n * 4));
    for (int attempt = 0;; ++attempt) {
        const size_t esz = wide ? 32 : 16;
        OV_TRY(d_P0.alloc((size_t)n * pcap * esz)); OV_TRY(d_A.alloc((size_t)n_lanes * cap * esz)); OV_TRY(d_B.alloc((size_t)n_lanes * cap * esz));
        OV_TRY(d_cat.alloc((size_t)n_lanes * cap * 8)); OV_TRY(d_slots.alloc((size_t)n * nei_cap * 32)); OV_TRY(d_mem.alloc((size_t)n * nei_cap * 32));
        OverlapArgs O;
        O.ix = idx->view; O.min_match = min_match; O.mode = 0; O.n = n; O.seq = d_seq.as<uint8_t>(); O.len = d_len.as<int32_t>(); O.max_len = max_len;
        O.ids = ids ? d_ids.as<uint64_t>() : nullptr; O.first = first; O.step = step; O.ret = d_ret.as<int64_t>();
        O.P0 = d_P0.p; O.pcap = pcap; O.np0 = d_np0.as<int32_t>(); O.A = d_A.p; O.B = d_B.p; O.cap = cap; O.cat = d_cat.as<int32_t>();
        O.rec = d_rec.as<int64_t>(); O.nei = d_slots.as<uint4>(); O.nei_cap = nei_cap; O.nei_cnt = d_cnt.as<uint32_t>();
        O.ext = d_ext.as<uint8_t>(); O.next = nullptr;
        OV_TRY(wide ? launch_phases<uint64_t>(O, grid, d_ctrl.as<unsigned long long>(), nullptr) : launch_phases<uint32_t>(O, grid, d_ctrl.as<unsigned long long>(), nullptr));
        OV_TRY(cudaDeviceSynchronize());
        if (fmg_verbose >= 4) std::fprintf(stderr, "[M::%s] overlap phases (attempt %d) %.3f s for %lld sequences\n", __func__, attempt, since(t1), (long long)n);
        OV_TRY(cudaMemcpy(h_len.data(), d_len.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < n; ++i)
            if (h_len[i] < 0) {
                if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] a sequence of %d bases exceeds max_len=%d\n", __func__, -h_len[i], max_len);
                return 2;
            }
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
        if (list_ovf) cap *= 4, pcap *= 4;
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
