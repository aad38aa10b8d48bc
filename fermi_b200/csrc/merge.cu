// `fermi merge` on the GPU: fm_compute_gap_bits (merge.c:31-94) + fm_merge (merge.c:100-137).
//   k_gap_bits     one thread per sequence of the second index e1: the LF walk of the sequence through e1 (as fm_retrieve,
//                  exact.c:59-70) in lock-step with the position i the same suffix would take in e0 (one rld_rank1a in each
//                  index per base, merge.c:44-56); bit k + i + 1 of the gap vector marks a symbol that comes from e1.  The
//                  reference stripes the sequences over threads and ORs into the shared vector atomically (merge.c:21-29);
//                  here every sequence is a thread and the OR is an atomicOr in HBM.
//   k_merge_pick   merged BWT[j] = bits[j] ? bwt1[ones before j] : bwt0[zeros before j]   (popcount prefix over the words)
// The merged BWT goes through the device RLD encoder (rld_enc.cu): the image is byte-identical to what fm_merge builds with
// rld_dec_enc run by run, because both encode the maximal runs of the same symbol string.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <vector>
#include "fmd_device.cuh"
#include "dev_pool.hpp"
#include "fmg_internal.hpp"
#include "../../include/fermi_b200.h"

using namespace fmg;

extern std::atomic<uint64_t> g_launches;
int fmg_rld_encode_device(const uint8_t *d_bwt, uint64_t n, fmg::FmdImage *out);     // rld_enc.cu

#define MG_TRY(call)                                                                                  \
    do {                                                                                              \
        cudaError_t err__ = (call);                                                                   \
        if (err__ != cudaSuccess) {                                                                   \
            if (fmg_verbose >= 1)                                                                     \
                std::fprintf(stderr, "[E::%s] %s failed: %s\n", __func__, #call, cudaGetErrorString(err__)); \
            return -1;                                                                                \
        }                                                                                             \
    } while (0)

namespace {

// rld_rank1a on the occ blocks: counts of BWT[0..k] for symbol BWT[k], and that symbol
__device__ __forceinline__ int rank1_at(const OccView &ix, uint64_t k, uint64_t *cnt_c) {
    const uint64_t p = k + 1;
    const Blk B = load_blk(ix, p);
    const int c = blk_symbol((k >> kBlkShift) == (p >> kBlkShift) ? B : load_blk(ix, k), k);
    uint32_t rel[6];
    rank_rel(B, p, rel);
    *cnt_c = ld_u64(ix.cs + (p >> kSuperShift) * 8 + c) + pick6(rel, c);          // C[c] + #c in BWT[0..k]
    return c;
}
// C[c] + #c in BWT[0..i] for a given symbol
__device__ __forceinline__ uint64_t rank_of(const OccView &ix, uint64_t i, int c) {
    const uint64_t p = i + 1;
    uint32_t rel[6];
    rank_rel(load_blk(ix, p), p, rel);
    return ld_u64(ix.cs + (p >> kSuperShift) * 8 + c) + pick6(rel, c);
}

__device__ __forceinline__ void set_bit(unsigned long long *bits, uint64_t q) { atomicOr(bits + (q >> 6), 1ull << (q & 63)); }

__global__ void __launch_bounds__(256) k_gap_bits(OccView e0, OccView e1, unsigned long long *bits) {
    const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= e1.n_seq) return;
    uint64_t k = x, i = e0.n_seq - 1;                  // merge.c:41-42
    set_bit(bits, i + k + 1);
    for (;;) {
        uint64_t lf;
        const int c = rank1_at(e1, k, &lf);            // c = rld_rank1a(e1, k, ok)
        if (c == 0 || c > 5) break;                    // the sequence is spelled out (merge.c:46-50: the next one is another thread)
        k = lf - 1;                                    // k = cnt[c] + ok[c] - 1
        i = rank_of(e0, i, c) - 1;                     // rld_rank1a(e0, i, ok); i = cnt[c] + ok[c] - 1
        set_bit(bits, k + i + 1);
    }
}

__global__ void __launch_bounds__(256) k_word_pop(const unsigned long long *__restrict__ bits, uint64_t n_words, uint64_t *__restrict__ pop) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < n_words) pop[w] = (uint64_t)__popcll(bits[w]);
}

__global__ void __launch_bounds__(256) k_merge_pick(uint64_t n, const unsigned long long *__restrict__ bits, const uint64_t *__restrict__ ones_before_word,
                                                   const uint8_t *__restrict__ bwt0, const uint8_t *__restrict__ bwt1, uint8_t *__restrict__ out) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned long long w = bits[j >> 6];
    const uint64_t ones = ones_before_word[j >> 6] + (uint64_t)__popcll(w & ((1ull << (j & 63)) - 1));
    out[j] = (w >> (j & 63) & 1) ? bwt1[ones] : bwt0[j - ones];
}

}  // namespace

extern "C" {

int fmg_gap_bits(const fmg_index_t *idx0, const fmg_index_t *idx1, uint64_t *bits) {
    if (!idx0 || !idx1 || !bits || idx0->device != idx1->device) return -1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] no CUDA device available; libfermi_b200 has no CPU path\n", __func__);
        return -1;
    }
    MG_TRY(cudaSetDevice(idx0->device));
    const uint64_t n = idx0->view.n_sym + idx1->view.n_sym, n_words = (n + 63) / 64;
    Dev d_bits;
    MG_TRY(d_bits.alloc(n_words * 8));
    MG_TRY(cudaMemset(d_bits.p, 0, n_words * 8));
    if (idx1->view.n_seq) {
        k_gap_bits<<<(unsigned)((idx1->view.n_seq + 255) / 256), 256>>>(idx0->view, idx1->view, d_bits.as<unsigned long long>());
        ++g_launches;
        MG_TRY(cudaGetLastError());
    }
    MG_TRY(cudaMemcpy(bits, d_bits.p, n_words * 8, cudaMemcpyDeviceToHost));
    return 0;
}

fmg_fmd_t *fmg_merge(const fmg_fmd_t *e0, const fmg_fmd_t *e1, int device) {
    if (!e0 || !e1) return nullptr;
    fmg_index_t *i0 = fmg_index_upload(e0, device), *i1 = i0 ? fmg_index_upload(e1, device) : nullptr;
    fmg_fmd_t *res = nullptr;
    do {
        if (!i0 || !i1) break;
        const uint64_t n0 = e0->img.mcnt[0], n1 = e1->img.mcnt[0], n = n0 + n1, n_words = (n + 63) / 64;
        Dev d_bits, d_pop, d_tmp, d_b0, d_b1, d_out;
        auto ok = [&](cudaError_t e, const char *what) {
            if (e == cudaSuccess) return true;
            if (fmg_verbose >= 1) std::fprintf(stderr, "[E::fmg_merge] %s failed: %s\n", what, cudaGetErrorString(e));
            return false;
        };
        if (!ok(d_bits.alloc(n_words * 8), "alloc") || !ok(d_pop.alloc((n_words + 1) * 8), "alloc") || !ok(d_b0.alloc(n0 + 1), "alloc") ||
            !ok(d_b1.alloc(n1 + 1), "alloc") || !ok(d_out.alloc(n + 1), "alloc")) break;
        if (!ok(cudaMemset(d_bits.p, 0, n_words * 8), "memset")) break;
        k_gap_bits<<<(unsigned)((i1->view.n_seq + 255) / 256), 256>>>(i0->view, i1->view, d_bits.as<unsigned long long>());
        k_word_pop<<<(unsigned)((n_words + 255) / 256), 256>>>(d_bits.as<unsigned long long>(), n_words, d_pop.as<uint64_t>());
        g_launches += 2;
        size_t need = 0;
        if (!ok(cub::DeviceScan::ExclusiveSum(nullptr, need, d_pop.as<uint64_t>(), d_pop.as<uint64_t>(), (int64_t)n_words), "scan")) break;
        if (!ok(d_tmp.alloc(need + 256), "alloc")) break;
        if (!ok(cub::DeviceScan::ExclusiveSum(d_tmp.p, need, d_pop.as<uint64_t>(), d_pop.as<uint64_t>(), (int64_t)n_words), "scan")) break;
        // the two symbol strings (decoded on the host from the .fmd images)
        {
            std::vector<uint8_t> h(std::max(n0, n1) + 1);
            fmg_fmd_decode_bwt(e0, h.data());
            if (!ok(cudaMemcpy(d_b0.p, h.data(), n0, cudaMemcpyHostToDevice), "copy")) break;
            fmg_fmd_decode_bwt(e1, h.data());
            if (!ok(cudaMemcpy(d_b1.p, h.data(), n1, cudaMemcpyHostToDevice), "copy")) break;
        }
        k_merge_pick<<<(unsigned)((n + 255) / 256), 256>>>(n, d_bits.as<unsigned long long>(), d_pop.as<uint64_t>(), d_b0.as<uint8_t>(), d_b1.as<uint8_t>(), d_out.as<uint8_t>());
        ++g_launches;
        if (!ok(cudaGetLastError(), "k_merge_pick")) break;
        fmg_fmd_t *e = new fmg_fmd_s;
        if (fmg_rld_encode_device(d_out.as<uint8_t>(), n, &e->img) != 0) { delete e; break; }
        res = e;
    } while (0);
    if (i1) fmg_index_free(i1);
    if (i0) fmg_index_free(i0);
    return res;
}

} // extern "C"
