// BWT of the multi-sentinel FMD text on the GPU by prefix doubling.
// Replaces fm_build / ksa_bwt (build.c:33-50, ksa.c:131-242) for texts of < 2^32 symbols that fit one
// GPU.  Sentinels compare by their position in the text (ksa.c:54), so the output equals the SA-IS
// one symbol for symbol: BWT[j] = T[SA[j]-1], 0 when SA[j]==0 (ksa.c:237-238).
//
// Round 0 sorts every suffix by its first K symbols with ONE 64-bit radix sort; a window that runs into
// a sentinel is cut there and carries the sentinel's ordinal in the low bits, which makes it unique.
// Every later round doubles the compared length for the suffixes that are still tied, sorting only
// those by (rank[s], rank[s+h]).  Radix sort / scan / select are CUB device primitives; the key
// construction, rank update and BWT gather kernels are ours.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdint>
#include <atomic>
#include "fmd_host.hpp"
#include "../../include/fermi_b200.h"

int fmg_rld_encode_device(const uint8_t *d_bwt, uint64_t n, fmg::FmdImage *out);     // rld_enc.cu
struct fmg_fmd_s { fmg::FmdImage img; };

extern std::atomic<uint64_t> g_launches;

namespace {

#define BW_TRY(call)                                                                                        \
    do {                                                                                                    \
        cudaError_t err__ = (call);                                                                         \
        if (err__ != cudaSuccess) {                                                                         \
            if (fmg_verbose >= 1)                                                                           \
                std::fprintf(stderr, "[E::fmg_build_bwt] %s failed: %s\n", #call, cudaGetErrorString(err__)); \
            return -1;                                                                                      \
        }                                                                                                   \
    } while (0)

constexpr int kThreads = 256;
inline unsigned blocks_for(uint64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

__global__ void k_is_sentinel(const uint8_t *__restrict__ T, uint32_t n, uint32_t *__restrict__ flag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = T[i] == 0;
}

// key of suffix i = its first K symbols (3 bits each); cut at the first sentinel, whose ordinal goes to the low B bits
__global__ void k_init_keys(const uint8_t *__restrict__ T, const uint32_t *__restrict__ sent_ord, uint32_t n, int K, int B,
                            uint64_t *__restrict__ key, uint32_t *__restrict__ val) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = 0, ord = 0;
    int t = 0;
    for (; t < K; ++t) {
        const uint32_t p = i + t;
        const uint32_t c = p < n ? T[p] : 0;
        if (c == 0) { ord = p < n ? sent_ord[p] : 0; break; }
        k = k << 3 | c;
    }
    k <<= 3 * (K - t);
    key[i] = k << B | ord;
    val[i] = i;
}

// head[j] = j if element j starts a new group of equal keys, else 0
__global__ void k_group_heads(const uint64_t *__restrict__ key, uint32_t m, const uint32_t *__restrict__ pos,
                              uint32_t *__restrict__ head) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const bool is_head = j == 0 || key[j] != key[j - 1];
    head[j] = is_head ? (pos ? pos[j] : j) : 0;
}

// after the max-scan head[j] is the group rank of element j: publish it and flag elements of groups larger than one
__global__ void k_set_rank(const uint64_t *__restrict__ key, const uint32_t *__restrict__ val, const uint32_t *__restrict__ head,
                           uint32_t m, uint32_t *__restrict__ rank, uint8_t *__restrict__ tied) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    rank[val[j]] = head[j];
    const bool first = j == 0 || key[j] != key[j - 1];
    const bool last = j == m - 1 || key[j] != key[j + 1];
    tied[j] = !(first && last);
}

__global__ void k_iota(uint32_t *a, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}

__global__ void k_double_keys(const uint32_t *__restrict__ val, uint32_t m, const uint32_t *__restrict__ rank, uint32_t n,
                              uint32_t h, uint64_t *__restrict__ key) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const uint32_t s = val[t];
    const uint64_t nxt = (uint64_t)s + h < n ? (uint64_t)rank[s + h] + 1 : 0;
    key[t] = (uint64_t)rank[s] << 32 | nxt;
}

__global__ void k_scatter_sa(const uint32_t *__restrict__ pos, const uint32_t *__restrict__ val, uint32_t m, uint32_t *__restrict__ sa) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m) sa[pos[t]] = val[t];
}

__global__ void k_gather_bwt(const uint8_t *__restrict__ T, const uint32_t *__restrict__ sa, uint32_t n, uint8_t *__restrict__ bwt) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) { const uint32_t s = sa[j]; bwt[j] = s ? T[s - 1] : 0; }
}

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <class T> T *as() { return static_cast<T *>(p); }
};

int build_bwt_device(uint32_t n, const uint8_t *h_text, uint8_t *h_bwt, fmg::FmdImage *img) {
    DevBuf dT, dSA, dRank, dKeyA, dKeyB, dValA, dValB, dPosA, dPosB, dHead, dTied, dTmp, dCount;
    BW_TRY(dT.alloc(n)); BW_TRY(dSA.alloc((size_t)n * 4)); BW_TRY(dRank.alloc((size_t)n * 4));
    BW_TRY(dKeyA.alloc((size_t)n * 8)); BW_TRY(dKeyB.alloc((size_t)n * 8));
    BW_TRY(dValA.alloc((size_t)n * 4)); BW_TRY(dValB.alloc((size_t)n * 4));
    BW_TRY(dPosA.alloc((size_t)n * 4)); BW_TRY(dPosB.alloc((size_t)n * 4));
    BW_TRY(dHead.alloc((size_t)n * 4)); BW_TRY(dTied.alloc(n)); BW_TRY(dCount.alloc(8));
    BW_TRY(cudaMemcpy(dT.p, h_text, n, cudaMemcpyHostToDevice));

    uint8_t *T = dT.as<uint8_t>();
    uint32_t *sa = dSA.as<uint32_t>(), *rank = dRank.as<uint32_t>(), *head = dHead.as<uint32_t>();
    uint8_t *tied = dTied.as<uint8_t>();
    uint32_t *count = dCount.as<uint32_t>();

    // CUB scratch: size it for the largest call (the round-0 sort over n pairs)
    size_t tmp_bytes = 0, need = 0;
    {
        cub::DoubleBuffer<uint64_t> dk(dKeyA.as<uint64_t>(), dKeyB.as<uint64_t>());
        cub::DoubleBuffer<uint32_t> dv(dValA.as<uint32_t>(), dValB.as<uint32_t>());
        BW_TRY(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, (int64_t)n, 0, 64));
        tmp_bytes = need;
        BW_TRY(cub::DeviceScan::InclusiveScan(nullptr, need, head, head, cub::Max(), (int64_t)n));
        tmp_bytes = need > tmp_bytes ? need : tmp_bytes;
        BW_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, head, head, (int64_t)n));
        tmp_bytes = need > tmp_bytes ? need : tmp_bytes;
        BW_TRY(cub::DeviceSelect::Flagged(nullptr, need, sa, tied, sa, count, (int64_t)n));
        tmp_bytes = need > tmp_bytes ? need : tmp_bytes;
    }
    BW_TRY(dTmp.alloc(tmp_bytes));

    // sentinel ordinals (exclusive prefix count of zeros); head[] is reused as the ordinal array
    k_is_sentinel<<<blocks_for(n), kThreads>>>(T, n, head); ++g_launches;
    need = tmp_bytes;
    BW_TRY(cub::DeviceScan::ExclusiveSum(dTmp.p, need, head, head, (int64_t)n));
    uint32_t last_ord = 0;
    BW_TRY(cudaMemcpy(&last_ord, head + (n - 1), 4, cudaMemcpyDeviceToHost));
    const uint32_t n_sent = last_ord + 1;       // T[n-1] is a sentinel
    int B = 1;
    while ((1ull << B) < n_sent) ++B;
    int K = (64 - B) / 3;
    if (K > 20) K = 20;

    cub::DoubleBuffer<uint64_t> keys(dKeyA.as<uint64_t>(), dKeyB.as<uint64_t>());
    cub::DoubleBuffer<uint32_t> vals(dValA.as<uint32_t>(), dValB.as<uint32_t>());
    k_init_keys<<<blocks_for(n), kThreads>>>(T, head, n, K, B, keys.Current(), vals.Current()); ++g_launches;
    need = tmp_bytes;
    BW_TRY(cub::DeviceRadixSort::SortPairs(dTmp.p, need, keys, vals, (int64_t)n, 0, 3 * K + B));
    BW_TRY(cudaMemcpyAsync(sa, vals.Current(), (size_t)n * 4, cudaMemcpyDeviceToDevice));
    k_group_heads<<<blocks_for(n), kThreads>>>(keys.Current(), n, nullptr, head); ++g_launches;
    need = tmp_bytes;
    BW_TRY(cub::DeviceScan::InclusiveScan(dTmp.p, need, head, head, cub::Max(), (int64_t)n));
    k_set_rank<<<blocks_for(n), kThreads>>>(keys.Current(), vals.Current(), head, n, rank, tied); ++g_launches;

    // active list = (position in SA, suffix) of every suffix still tied with a neighbour
    cub::DoubleBuffer<uint32_t> pos(dPosA.as<uint32_t>(), dPosB.as<uint32_t>());
    k_iota<<<blocks_for(n), kThreads>>>(pos.Current(), n); ++g_launches;
    uint32_t m = n, h = (uint32_t)K;
    int round = 0;
    for (;;) {
        // compact (pos, val) by tied[]
        uint32_t m2 = 0;
        need = tmp_bytes;
        BW_TRY(cub::DeviceSelect::Flagged(dTmp.p, need, pos.Current(), tied, pos.Alternate(), count, (int64_t)m));
        need = tmp_bytes;
        BW_TRY(cub::DeviceSelect::Flagged(dTmp.p, need, vals.Current(), tied, vals.Alternate(), count, (int64_t)m));
        BW_TRY(cudaMemcpy(&m2, count, 4, cudaMemcpyDeviceToHost));
        pos.selector ^= 1; vals.selector ^= 1;
        if (fmg_verbose >= 4)
            std::fprintf(stderr, "[M::fmg_build_bwt] round %d: h=%u, %u of %u suffixes still tied\n", round, h, m2, n);
        m = m2;
        if (m == 0) break;
        if ((uint64_t)h >= n) {
            if (fmg_verbose >= 1) std::fprintf(stderr, "[E::fmg_build_bwt] suffixes still tied at h >= n (text without final sentinel?)\n");
            return -1;
        }
        k_double_keys<<<blocks_for(m), kThreads>>>(vals.Current(), m, rank, n, h, keys.Current()); ++g_launches;
        need = tmp_bytes;
        BW_TRY(cub::DeviceRadixSort::SortPairs(dTmp.p, need, keys, vals, (int64_t)m, 0, 64));
        k_scatter_sa<<<blocks_for(m), kThreads>>>(pos.Current(), vals.Current(), m, sa); ++g_launches;
        k_group_heads<<<blocks_for(m), kThreads>>>(keys.Current(), m, pos.Current(), head); ++g_launches;
        need = tmp_bytes;
        BW_TRY(cub::DeviceScan::InclusiveScan(dTmp.p, need, head, head, cub::Max(), (int64_t)m));
        k_set_rank<<<blocks_for(m), kThreads>>>(keys.Current(), vals.Current(), head, m, rank, tied); ++g_launches;
        h = h > (1u << 30) ? 0xffffffffu : h * 2;
        ++round;
    }
    // BWT (reuse tied[] as the output buffer)
    k_gather_bwt<<<blocks_for(n), kThreads>>>(T, sa, n, tied); ++g_launches;
    BW_TRY(cudaGetLastError());
    if (h_bwt) BW_TRY(cudaMemcpy(h_bwt, tied, n, cudaMemcpyDeviceToHost));
    if (img && fmg_rld_encode_device(tied, n, img) != 0) return -1;      // the BWT never leaves the device
    return 0;
}

} // namespace

extern "C" int fmg_build_bwt(int device, int64_t n, const uint8_t *text, uint8_t *bwt) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] no CUDA device available; libfermi_b200 has no CPU path\n", __func__);
        return -1;
    }
    if (n <= 0 || n >= 0xffffffffll || text[n - 1] != 0) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] text must have 1..2^32-2 symbols and end with a sentinel\n", __func__);
        return -1;
    }
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    return build_bwt_device((uint32_t)n, text, bwt, nullptr);
}

// fm_build (build.c:33-50) entirely on the device: suffix sort, BWT and RLD encoding; only the .fmd image comes back
extern "C" fmg_fmd_t *fmg_build_fmd(int device, int64_t n, const uint8_t *text) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] no CUDA device available; libfermi_b200 has no CPU path\n", __func__);
        return nullptr;
    }
    if (n <= 0 || n >= 0xffffffffll || text[n - 1] != 0) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] text must have 1..2^32-2 symbols and end with a sentinel\n", __func__);
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    fmg_fmd_t *e = new fmg_fmd_s;
    if (build_bwt_device((uint32_t)n, text, nullptr, &e->img) != 0) { delete e; return nullptr; }
    return e;
}

// fm_bwtenc (build.c:11-31) on the device for a BWT that lives on the host: copy in, encode, image out
extern "C" fmg_fmd_t *fmg_fmd_from_bwt_device(int device, int64_t n, const uint8_t *bwt) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || n <= 0 || !bwt) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] no CUDA device available (or empty input); libfermi_b200 has no CPU path\n", __func__);
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    uint8_t *d = nullptr;
    if (cudaMalloc(&d, (size_t)n) != cudaSuccess) return nullptr;
    fmg_fmd_t *e = nullptr;
    if (cudaMemcpy(d, bwt, (size_t)n, cudaMemcpyHostToDevice) == cudaSuccess) {
        e = new fmg_fmd_s;
        if (fmg_rld_encode_device(d, (uint64_t)n, &e->img) != 0) { delete e; e = nullptr; }
    }
    cudaFree(d);
    return e;
}
