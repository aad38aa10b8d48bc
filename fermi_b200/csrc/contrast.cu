// `fermi contrast` on the GPU: fm6_contrast (cmp.c:45-126), the lock-step walk of the backward-extension tries of TWO FMD-indexes.
// The reference walks depth first, one 4^SUF_LEN-th of the trie per thread, and ORs sequence ranks into two shared bitmaps
// (cmp.c:33-37); the bitmaps are a set, so the order of the walk does not matter.  Here both tries are expanded breadth first:
//   k_pair_expand   one node = the pair of bi-intervals of one string in the two indexes; a node whose string is absent from one
//                   index hands the other index's interval to that index's tip list (collect_tips, cmp.c:22-43); otherwise,
//                   below depth k, both are extended and the children with >= min_occ occurrences in either index survive
//                   (cmp.c:61-71)
//   k_tips_expand   collect_tips: every sequence that starts with the string (ok[0]) is marked, all non-empty children followed
// Frontiers live in HBM and are appended to with atomics; coordinates are 64-bit (the command is not on the benchmark path).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <algorithm>
#include "fmd_device.cuh"
#include "dev_pool.hpp"
#include "fmg_internal.hpp"
#include "../../include/fermi_b200.h"

using namespace fmg;
extern std::atomic<uint64_t> g_launches;

#define CT_TRY(call)                                                                                  \
    do {                                                                                              \
        cudaError_t err__ = (call);                                                                   \
        if (err__ != cudaSuccess) {                                                                   \
            if (fmg_verbose >= 1)                                                                     \
                std::fprintf(stderr, "[E::fmg_contrast] %s failed: %s\n", #call, cudaGetErrorString(err__)); \
            return -1;                                                                                \
        }                                                                                             \
    } while (0)

namespace {

constexpr int kSufLen = 4;                       // SUF_LEN, cmp.c:8
struct Iv { uint64_t x0, x1, x2; };              // a bi-interval without its info
struct Pair { Iv a, b; };                        // the same string in index 0 and index 1

// all six backward extensions of iv (fm6_extend(e, &ik, ok, 1), exact.c:72-88); an empty interval extends to empty ones
__device__ __forceinline__ void extend_back(const OccView &ix, const Iv &iv, Iv ok[6]) {
    Ext6T<uint64_t> e;
    extend6<uint64_t>(ix, iv.x1, iv.x0, iv.x2, e);
#pragma unroll
    for (int c = 0; c < 6; ++c) { ok[c].x0 = far_of(ix, e, c); ok[c].x1 = e.near[c]; ok[c].x2 = e.size[c]; }
}

// descend (cmp.c:10-20) one level: all four children, empty or not
__global__ void __launch_bounds__(256) k_pair_descend(OccView e0, OccView e1, const Pair *__restrict__ in, uint64_t n_in, Pair *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    Iv a[6], b[6];
    extend_back(e0, in[i].a, a);
    extend_back(e1, in[i].b, b);
#pragma unroll
    for (int c = 1; c <= 4; ++c) { Pair p; p.a = a[c]; p.b = b[c]; out[4 * i + (c - 1)] = p; }
}

// one level of contrast_core (cmp.c:57-72); ctr: [0] next pairs, [1] tips of index 0, [2] tips of index 1
__global__ void __launch_bounds__(256) k_pair_expand(OccView e0, OccView e1, const Pair *__restrict__ in, uint64_t n_in, int depth, int kmer, uint64_t min_occ,
                                                     Pair *__restrict__ out, uint64_t cap_out, Iv *tips0, Iv *tips1, uint64_t cap_tips, unsigned long long *ctr) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    const Pair p = in[i];
    if (p.a.x2 == 0) {                                             // absent from index 0: what index 1 has below is unique to it
        if (p.b.x2 == 0) return;
        const unsigned long long s = atomicAdd(ctr + 2, 1ull);
        if (s < cap_tips) tips1[s] = p.b;
        return;
    }
    if (p.b.x2 == 0) {
        const unsigned long long s = atomicAdd(ctr + 1, 1ull);
        if (s < cap_tips) tips0[s] = p.a;
        return;
    }
    if (depth >= kmer) return;
    Iv a[6], b[6];
    extend_back(e0, p.a, a);
    extend_back(e1, p.b, b);
#pragma unroll
    for (int c = 1; c <= 4; ++c) {
        if (a[c].x2 < min_occ && b[c].x2 < min_occ) continue;
        const unsigned long long s = atomicAdd(ctr, 1ull);
        if (s < cap_out) { Pair q; q.a = a[c]; q.b = b[c]; out[s] = q; }
    }
}

// one level of collect_tips (cmp.c:22-43)
__global__ void __launch_bounds__(256) k_tips_expand(OccView ix, const Iv *__restrict__ in, uint64_t n_in, Iv *__restrict__ out, uint64_t cap_out,
                                                     unsigned long long *n_out, unsigned long long *sub) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    Iv ok[6];
    extend_back(ix, in[i], ok);
    for (uint64_t k = 0; k < ok[0].x2; ++k) {
        const uint64_t x = ok[0].x0 + k;
        atomicOr(sub + (x >> 6), 1ull << (x & 63));
    }
#pragma unroll
    for (int c = 1; c <= 4; ++c)
        if (ok[c].x2) {
            const unsigned long long s = atomicAdd(n_out, 1ull);
            if (s < cap_out) out[s] = ok[c];
        }
}

inline unsigned nb(uint64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

extern "C" int fmg_contrast(const fmg_index_t *idx0, const fmg_index_t *idx1, int k, int min_occ, uint64_t *sub0, uint64_t *sub1) {
    if (!idx0 || !idx1 || !sub0 || !sub1 || k <= kSufLen || idx0->device != idx1->device) return -1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] no CUDA device available; libfermi_b200 has no CPU path\n", __func__);
        return -1;
    }
    CT_TRY(cudaSetDevice(idx0->device));
    const OccView &e0 = idx0->view, &e1 = idx1->view;
    const uint64_t w0 = (e0.n_seq + 63) / 64, w1 = (e1.n_seq + 63) / 64;
    Dev d_sub0, d_sub1, d_ctr, d_pair[2], d_tips[2], d_tw[2];
    CT_TRY(d_sub0.alloc(std::max<uint64_t>(w0, 1) * 8)); CT_TRY(d_sub1.alloc(std::max<uint64_t>(w1, 1) * 8)); CT_TRY(d_ctr.alloc(64));
    CT_TRY(cudaMemset(d_sub0.p, 0, std::max<uint64_t>(w0, 1) * 8)); CT_TRY(cudaMemset(d_sub1.p, 0, std::max<uint64_t>(w1, 1) * 8));
    unsigned long long *ctr = d_ctr.as<unsigned long long>(), h[8];
    uint64_t cap = 1 << 20, cap_tips = 1 << 20;
    for (int attempt = 0;; ++attempt) {
        CT_TRY(d_pair[0].alloc(cap * sizeof(Pair))); CT_TRY(d_pair[1].alloc(cap * sizeof(Pair)));
        CT_TRY(d_tips[0].alloc(cap_tips * sizeof(Iv))); CT_TRY(d_tips[1].alloc(cap_tips * sizeof(Iv)));
        CT_TRY(cudaMemset(ctr, 0, 64));
        // depth 1: the four single-base intervals of both indexes (fm6_set_intv, cmp.c:14), then descend to depth SUF_LEN
        Pair root[4];
        for (int c = 1; c <= 4; ++c) {
            root[c - 1].a = Iv{e0.C[c], e0.C[5 - c], e0.C[c + 1] - e0.C[c]};
            root[c - 1].b = Iv{e1.C[c], e1.C[5 - c], e1.C[c + 1] - e1.C[c]};
        }
        CT_TRY(cudaMemcpy(d_pair[0].p, root, sizeof root, cudaMemcpyHostToDevice));
        uint64_t n = 4;
        int cur = 0;
        for (int d = 1; d < kSufLen; ++d) {
            k_pair_descend<<<nb(n), 256>>>(e0, e1, d_pair[cur].as<Pair>(), n, d_pair[cur ^ 1].as<Pair>());
            ++g_launches;
            n *= 4; cur ^= 1;
        }
        bool overflow = false;
        for (int d = kSufLen; n > 0; ++d) {
            CT_TRY(cudaMemset(ctr, 0, 8));
            k_pair_expand<<<nb(n), 256>>>(e0, e1, d_pair[cur].as<Pair>(), n, d, k, (uint64_t)(int64_t)min_occ, d_pair[cur ^ 1].as<Pair>(), cap,
                                          d_tips[0].as<Iv>(), d_tips[1].as<Iv>(), cap_tips, ctr);
            ++g_launches;
            CT_TRY(cudaMemcpy(h, ctr, 24, cudaMemcpyDeviceToHost));
            if (h[0] > cap || h[1] > cap_tips || h[2] > cap_tips) { overflow = true; break; }
            n = h[0]; cur ^= 1;
        }
        if (overflow) {                                           // a frontier did not fit: grow and start over (the bitmaps are untouched so far)
            if (attempt == 8) return -1;
            cap = std::max<uint64_t>(cap * 4, h[0] + (h[0] >> 2));
            cap_tips = std::max<uint64_t>(cap_tips * 4, std::max(h[1], h[2]) * 2);
            continue;
        }
        // collect_tips from the accumulated roots, one index at a time; a level has at most four children per node
        bool tips_overflow = false;
        for (int which = 0; which < 2 && !tips_overflow; ++which) {
            const OccView &ix = which ? e1 : e0;
            unsigned long long *sub = which ? d_sub1.as<unsigned long long>() : d_sub0.as<unsigned long long>();
            const Iv *in = d_tips[which].as<Iv>();
            uint64_t nt = h[1 + which], capw[2] = {0, 0};
            int tcur = 0;
            while (nt > 0) {
                if (capw[tcur] < 4 * nt) { capw[tcur] = 4 * nt + 1024; CT_TRY(d_tw[tcur].alloc(capw[tcur] * sizeof(Iv))); }
                CT_TRY(cudaMemset(ctr + 3, 0, 8));
                k_tips_expand<<<nb(nt), 256>>>(ix, in, nt, d_tw[tcur].as<Iv>(), capw[tcur], ctr + 3, sub);
                ++g_launches;
                CT_TRY(cudaMemcpy(h + 3, ctr + 3, 8, cudaMemcpyDeviceToHost));
                if (h[3] > capw[tcur]) { tips_overflow = true; break; }
                nt = h[3];
                in = d_tw[tcur].as<Iv>();
                tcur ^= 1;
            }
        }
        if (tips_overflow) return -1;
        break;
    }
    CT_TRY(cudaGetLastError());
    CT_TRY(cudaMemcpy(sub0, d_sub0.p, w0 * 8, cudaMemcpyDeviceToHost));
    CT_TRY(cudaMemcpy(sub1, d_sub1.p, w1 * 8, cudaMemcpyDeviceToHost));
    return 0;
}
