// k-mer collection of `fermi correct` on the GPU: fm6_traverse (exact.c:141-171) + ec_collect (correct.c:35-87).
//
// The reference walks the backward-extension trie depth first, one 4^SUF_LEN-th of it per thread, with a
// stack per thread.  Here the trie is expanded breadth first, one level per launch: every node of the
// frontier does its one fm6_extend in lock-step (fully converged warps, coalesced node reads) and appends its
// surviving children to the next frontier through warp-aggregated atomics.  The frontier (<= one entry per
// distinct k-mer with >= min_occ occurrences, 16-32 bytes each) lives in HBM.  At depth w every node emits
// the packed (suffix, key, val) triple that ec_collect puts into solid[suffix] (correct.c:56-75); the host
// fills the hash tables from them (the correction itself, ec_fix*, is host code and out of scope).
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <cstdio>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <algorithm>
#include <cmath>
#include "fmd_device.cuh"
#include "fmg_internal.hpp"
#include "dev_pool.hpp"
#include "../../include/fermi_b200.h"

using namespace fmg;
extern std::atomic<uint64_t> g_launches;

#define EC_TRY(call)                                                                                   \
    do {                                                                                               \
        cudaError_t err__ = (call);                                                                    \
        if (err__ != cudaSuccess) {                                                                    \
            if (fmg_verbose >= 1)                                                                      \
                std::fprintf(stderr, "[E::fmg_ec_collect] %s failed: %s\n", #call, cudaGetErrorString(err__)); \
            return -1;                                                                                 \
        }                                                                                              \
    } while (0)

namespace {

template <typename U> struct Frontier { U *x0, *x1, *x2; uint64_t *path; };

// one level of the trie: node -> children with at least `thr` occurrences (thr = 1 above depth SUF_LEN: fm6_traverse
// keeps every non-empty child, exact.c:158-164; below it ec_collect prunes with min_occ, correct.c:77-82)
template <typename U>
__global__ void __launch_bounds__(256) k_trie_expand(OccView ix, Frontier<U> in, uint64_t n_in, int depth, uint64_t thr,
                                                     Frontier<U> out, unsigned long long *n_out, uint64_t cap_out, uint32_t part, uint32_t n_parts) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    const U x0 = in.x0[i], x1 = in.x1[i], x2 = in.x2[i];
    const uint64_t path = in.path[i];
    Ext6T<U> e;
    extend6<U>(ix, x1, x0, x2, e);                       // backward: far = x[0], near = x[1]
    // n_parts > 1 (set on the level that completes the suffix): keep only the subtrees of suffixes s with s % n_parts == part --
    // the unit the reference hands to its threads (correct.c:346-350) and this library to its GPUs
    int n_child = 0;
#pragma unroll
    for (int c = 1; c <= 4; ++c) {
        if (n_parts > 1 && (uint32_t)((path | (uint64_t)(c - 1) << (2 * depth)) % n_parts) != part) e.size[c] = 0;
        n_child += e.size[c] >= thr && e.size[c] != 0;
    }
    if (n_child == 0) return;
    const unsigned long long base = atomicAdd(n_out, (unsigned long long)n_child);      // (ptxas aggregates per warp)
    if (base + n_child > cap_out) return;               // overflow: detected by the host from *n_out
    int k = 0;
#pragma unroll
    for (int c = 1; c <= 4; ++c)
        if (e.size[c] >= thr && e.size[c] != 0) {
            out.x0[base + k] = far_of(ix, e, c); out.x1[base + k] = e.near[c]; out.x2[base + k] = e.size[c];
            out.path[base + k] = path | (uint64_t)(c - 1) << (2 * depth);
            ++k;
        }
}

// depth w: ec_collect's "keep the k-mer" branch, correct.c:56-75
template <typename U>
__global__ void __launch_bounds__(256) k_ec_emit(OccView ix, Frontier<U> in, uint64_t n_in, int suf_len, uint64_t min_occ,
                                                 uint64_t *triples, unsigned long long *ctr /* [0] n_out, [1] cnt0, [2] cnt1 */) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    const U x0 = in.x0[i], x1 = in.x1[i], x2 = in.x2[i];
    const uint64_t path = in.path[i];
    Ext6T<U> e;
    extend6<U>(ix, x1, x0, x2, e);
    uint64_t mx = 0; int max_c = 6;
#pragma unroll
    for (int c = 1; c <= 4; ++c) if ((uint64_t)e.size[c] > mx) mx = e.size[c], max_c = c;
    if (mx < min_occ) return;
    const uint64_t rest = (uint64_t)x2 - mx - e.size[0] - e.size[5];
    double r = rest == 0 ? (double)mx : (double)mx / (double)rest;
    if (r > 31.) r = 31.;
    const uint64_t suffix = path & ((1ull << (2 * suf_len)) - 1);
    const uint64_t key = (path >> (2 * suf_len)) << 2 | (uint64_t)(max_c - 1);
    const uint64_t val = (uint64_t)(int)(r + .499) << 3 | (rest < 7 ? rest : 7);
    const unsigned long long slot = atomicAdd(ctr, 1ull);
    triples[slot] = suffix << 40 | (key & 0xffffffffull) << 8 | val;
    atomicAdd(ctr + 1, 1ull);
    if (rest <= 7 && r >= (double)min_occ) atomicAdd(ctr + 2, 1ull);
}

// frontier storage comes from the library's device pool (dev_pool.hpp) and grows geometrically: a level that does not fit is
// run again, and cudaMalloc / cudaFree of gigabyte buffers on every level cost more than the expansion itself
template <typename U> struct FrontierBuf {
    fmg::Dev xs, ps;
    U *x = nullptr; uint64_t *path = nullptr; uint64_t cap = 0;
    cudaError_t reserve(uint64_t n) {
        if (n <= cap) return cudaSuccess;
        if (n < 2 * cap) n = 2 * cap;
        cap = 0; x = nullptr; path = nullptr;
        cudaError_t err = xs.alloc(n * 3 * sizeof(U));
        if (err == cudaSuccess) err = ps.alloc(n * 8);
        if (err == cudaSuccess) { cap = n; x = xs.template as<U>(); path = ps.template as<uint64_t>(); }
        return err;
    }
    Frontier<U> view() const { return Frontier<U>{x, x + cap, x + 2 * cap, path}; }
};

template <typename U>
int ec_collect_impl(const fmg_index_s *idx, int w, int suf_len, uint64_t min_occ, uint32_t part, uint32_t n_parts, uint64_t **triples, uint64_t *n_triples,
                    int64_t cnt[2]) {
    const OccView &ix = idx->view;
    const auto t_start = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(); };
    FrontierBuf<U> fb[2];
    unsigned long long *d_ctr = nullptr, h_ctr[4];
    uint64_t *d_tri = nullptr;
    EC_TRY(cudaMalloc(&d_ctr, 4 * sizeof(unsigned long long)));
    uint64_t cap = 1 << 20;
    EC_TRY(fb[0].reserve(cap)); EC_TRY(fb[1].reserve(cap));
    // depth 1: the four single-base intervals (fm6_set_intv, exact.c:153-156)
    {
        U h[12]; uint64_t hp[4]; int n = 0;
        for (int c = 1; c <= 4; ++c) {
            const uint64_t sz = ix.C[c + 1] - ix.C[c];
            if (sz == 0) continue;
            h[n] = (U)ix.C[c]; h[4 + n] = (U)ix.C[5 - c]; h[8 + n] = (U)sz; hp[n] = (uint64_t)(c - 1);
            ++n;
        }
        Frontier<U> v = fb[0].view();
        EC_TRY(cudaMemcpy(v.x0, h, n * sizeof(U), cudaMemcpyHostToDevice));
        EC_TRY(cudaMemcpy(v.x1, h + 4, n * sizeof(U), cudaMemcpyHostToDevice));
        EC_TRY(cudaMemcpy(v.x2, h + 8, n * sizeof(U), cudaMemcpyHostToDevice));
        EC_TRY(cudaMemcpy(v.path, hp, n * 8, cudaMemcpyHostToDevice));
        h_ctr[0] = n;
    }
    uint64_t n_cur = h_ctr[0];
    int cur = 0;
    for (int depth = 1; depth < w && n_cur > 0; ++depth) {          // nodes at `depth` -> nodes at depth+1
        const uint64_t thr = depth < suf_len ? 1 : min_occ;
        for (;;) {
            if (fb[cur ^ 1].cap < n_cur + (n_cur >> 1)) EC_TRY(fb[cur ^ 1].reserve(n_cur + (n_cur >> 1)));   // grown further on demand below
            EC_TRY(cudaMemset(d_ctr, 0, 4 * sizeof(unsigned long long)));
            // children of this level sit at depth + 1: the suffix (the first suf_len bases) is complete when depth + 1 == suf_len
            k_trie_expand<U><<<(unsigned)((n_cur + 255) / 256), 256>>>(ix, fb[cur].view(), n_cur, depth, thr, fb[cur ^ 1].view(), d_ctr, fb[cur ^ 1].cap,
                                                                       part, depth + 1 == suf_len ? n_parts : 1u);
            ++g_launches;
            EC_TRY(cudaMemcpy(h_ctr, d_ctr, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            if (h_ctr[0] <= fb[cur ^ 1].cap) break;
            EC_TRY(fb[cur ^ 1].reserve(h_ctr[0]));                    // the level needs more room: run it again
        }
        n_cur = h_ctr[0];
        cur ^= 1;
        if (fmg_verbose >= 4) std::fprintf(stderr, "[M::fmg_ec_collect] depth %d: %llu nodes\n", depth + 1, (unsigned long long)n_cur);
    }
    cnt[0] = cnt[1] = 0;
    *n_triples = 0;
    *triples = nullptr;
    const double t_expand = since();
    if (n_cur > 0) {
        fmg::Dev b_tri, b_sorted, b_tmp;
        EC_TRY(b_tri.alloc(n_cur * 8));
        d_tri = b_tri.as<uint64_t>();
        EC_TRY(cudaMemset(d_ctr, 0, 4 * sizeof(unsigned long long)));
        k_ec_emit<U><<<(unsigned)((n_cur + 255) / 256), 256>>>(ix, fb[cur].view(), n_cur, suf_len, min_occ, d_tri, d_ctr);
        ++g_launches;
        EC_TRY(cudaMemcpy(h_ctr, d_ctr, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        *n_triples = h_ctr[0];
        cnt[0] = (int64_t)h_ctr[1]; cnt[1] = (int64_t)h_ctr[2];
        *triples = (uint64_t *)std::malloc((h_ctr[0] ? h_ctr[0] : 1) * 8);
        if (h_ctr[0]) {                                               // canonical order: by suffix, then key (radix sort on the device)
            size_t need = 0;
            EC_TRY(b_sorted.alloc(h_ctr[0] * 8));
            uint64_t *d_sorted = b_sorted.as<uint64_t>();
            EC_TRY(cub::DeviceRadixSort::SortKeys(nullptr, need, d_tri, d_sorted, (int64_t)h_ctr[0]));
            EC_TRY(b_tmp.alloc(need));
            EC_TRY(cub::DeviceRadixSort::SortKeys(b_tmp.p, need, d_tri, d_sorted, (int64_t)h_ctr[0]));
            EC_TRY(cudaMemcpy(*triples, d_sorted, h_ctr[0] * 8, cudaMemcpyDeviceToHost));
        }
    } else *triples = (uint64_t *)std::malloc(8);
    cudaFree(d_ctr);
    if (fmg_verbose >= 3)
        std::fprintf(stderr, "[M::fmg_ec_collect] k=%d: %llu k-mers; trie expansion %.3f s, emit + sort + copy out %.3f s\n", w, (unsigned long long)*n_triples,
                     t_expand, since() - t_expand);
    return 0;
}

} // namespace

extern "C" {

int fmg_ec_kmer_length(uint64_t n_symbols) {      // fm6_ec_correct, correct.c:313-318
    int w = (int)(log((double)n_symbols) / log(4.0) + 8.499);
    return w >= 27 ? 27 : w;
}

int fmg_ec_collect(const fmg_index_t *idx, int w, int min_occ, uint64_t **triples, uint64_t *n_triples, int64_t cnt[2]) {
    return fmg_ec_collect_part(idx, w, min_occ, 0, 1, triples, n_triples, cnt);
}

int fmg_ec_collect_part(const fmg_index_t *idx, int w, int min_occ, int part, int n_parts, uint64_t **triples, uint64_t *n_triples, int64_t cnt[2]) {
    if (!idx || !triples || !n_triples || !cnt || n_parts < 1 || part < 0 || part >= n_parts) return -1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] no CUDA device available; libfermi_b200 has no CPU path\n", __func__);
        return -1;
    }
    if (cudaSetDevice(idx->device) != cudaSuccess) return -1;
    if (w < 0) w = fmg_ec_kmer_length(idx->mcnt[0]);
    if (w < 2 || w > 27) { if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] k-mer length %d out of range [2,27]\n", __func__, w); return -1; }
    const int suf_len = w > 15 ? w - 15 : 1;        // compute_SUF, correct.c:319
    const bool wide = idx->view.n_sym + 256 >= (1ull << 32) || std::getenv("FMG_FORCE_WIDE") != nullptr;
    if (suf_len == 1 && n_parts > 1) {                  // the depth-1 nodes are set up on the host: a suffix of one base cannot be filtered on a level
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] k-mer length %d has one-base suffixes: nothing to shard\n", __func__, w);
        return -1;
    }
    return wide ? ec_collect_impl<uint64_t>(idx, w, suf_len, (uint64_t)min_occ, (uint32_t)part, (uint32_t)n_parts, triples, n_triples, cnt)
                : ec_collect_impl<uint32_t>(idx, w, suf_len, (uint64_t)min_occ, (uint32_t)part, (uint32_t)n_parts, triples, n_triples, cnt);
}

} // extern "C"
