// Unitig construction on the GPU from the device-resident overlap records (replaces the seed-by-seed walk of
// unitig_core / unitig1 / unitig_unidir, unitig.c:227-362, for regular link graphs).
//
// The walk of the reference advances from a read to its neighbour when (unitig.c:236-251): the read has exactly one
// right neighbour k, k is not the read's own reverse complement, k has no other left neighbour (check_left), and k is
// not the read itself.  Every one of these tests is a function of the packed records (OvPack), so the links
// succ(v) = k form a static graph over the sequence ranks.  It contains every unitig twice -- v1 > v2 > ... > vm and
// rc(vm) > ... > rc(v1) -- and the reference emits whichever orientation its first unused seed happens to walk.
// Here:   k_links        succ(v) for every primary, non-contained sequence
//         k_pred         pred = inverse of succ (must be injective) and the check that rc(k) links back to rc(v)
//         k_rank_*, k_jump  head, number of reads and base offset of every read in its chain: list ranking over splitters
//                        (walks between splitters, pointer jumping over the splitters only)
//         k_tails/k_select  one orientation per chain (head <= rc(tail)), sizes for the output scans
//         k_emit_*       consensus = head sequence (LF walk, exact.c:59-70) + appended bases of every link (unitig.c:141),
//                        coverage = 33 + min(93, reads covering the base) via a difference array + scan (unitig.c:253-257),
//                        end ids and end neighbour lists exactly as unitig1 leaves them (unitig.c:300-316)
// A link graph with a cycle, a one-sided link or a shared successor has no order-free answer (the reference resolves it
// by seed order): the function then reports 1 and fmg_unitig runs the host walk, which reproduces that order.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <atomic>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <algorithm>
#include "fmd_overlap.cuh"
#include "ov_records.hpp"
#include "dev_pool.hpp"
#include "fmg_internal.hpp"
#include "../../include/fermi_b200.h"

using namespace fmg;

extern std::atomic<uint64_t> g_launches;

#define UG_TRY(call)                                                                                  \
    do {                                                                                              \
        cudaError_t err__ = (call);                                                                   \
        if (err__ != cudaSuccess) {                                                                   \
            if (fmg_verbose >= 1)                                                                     \
                std::fprintf(stderr, "[E::%s] %s failed: %s\n", __func__, #call, cudaGetErrorString(err__)); \
            return -1;                                                                                \
        }                                                                                             \
    } while (0)

namespace {

constexpr uint32_t kNone = 0xffffffffu;
enum { UGF_NONINJ = 1, UGF_ASYM = 2, UGF_NOTNODE = 4, UGF_CHANGED = 8 };

struct UMeta {                  // one unitig, as mag_v_write prints it (mag.c:149-174)
    uint64_t k0, k1;            // end ids
    uint64_t seq_off;           // first base in useq / ucov
    uint64_t nei_off;           // first neighbour entry: n0 entries of end 0, then n1 of end 1
    uint32_t len, nsr, n0, n1;
};
struct UNei { uint64_t x; int64_t ovlp; };

struct G {
    const OvPack *pack;
    const uint4 *spill;         // fmg_intv_t entries
    const uint8_t *ext;
    const int64_t *rank_of_row;
    uint64_t n;
    int min_match;
    uint32_t *succ, *pred, *row_of_rank, *tail_of;
    uint2 *step;                // (succ, rbeg) of every row in one 8-byte record: what a walk along a chain reads per read
    uint32_t *flags;
    unsigned long long *counts;     // [0] nodes, [1] nodes reachable from a head (they differ when the graph has a cycle)
};

__device__ __forceinline__ bool is_node(const OvPack &p, uint64_t v, int min_match) {
    return p.x0 == v && p.contained == 0 && (int)p.len > min_match;
}

// does the walk step from the read with record p to its unique neighbour?  (unitig.c:236-251 without the `bend` cache)
__device__ __forceinline__ bool links(const G &g, const OvPack &p) {
    if (p.rbeg < 0 || p.nnei != 1) return false;
    if (p.nx0 == p.x1) return false;                                   // the neighbour is the read's own reverse complement
    if (p.left != 0 && g.pack[p.nx1].nnei > 1) return false;           // check_left: backward bifurcation
    if (p.nx1 == p.x1) return false;                                   // the neighbour is the read itself
    return true;
}

__global__ void __launch_bounds__(256) k_links(G g) {
    const uint64_t v0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = v0 < g.n;
    const uint64_t v = in ? v0 : 0;
    if (in) g.row_of_rank[(uint64_t)g.rank_of_row[v]] = (uint32_t)v;   // v is a BWT row here; the array maps rank -> row
    const OvPack p = g.pack[v];
    uint32_t s = kNone;
    const bool node = in && is_node(p, v, g.min_match);
    if (node && links(g, p)) s = (uint32_t)p.nx0;
    if (in) { g.succ[v] = s; g.step[v] = make_uint2(s, (uint32_t)p.rbeg); }
    const int c = __syncthreads_count(node);
    if (threadIdx.x == 0 && c) atomicAdd(g.counts, (unsigned long long)c);
}

__global__ void __launch_bounds__(256) k_pred(G g) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= g.n) return;
    const uint32_t k = g.succ[v];
    if (k == kNone) return;
    uint32_t f = 0;
    if (atomicCAS(&g.pred[k], kNone, (uint32_t)v) != kNone) f |= UGF_NONINJ;
    const OvPack pk = g.pack[k];
    if (!is_node(pk, k, g.min_match)) f |= UGF_NOTNODE;
    else if (g.succ[pk.x1] != (uint32_t)g.pack[v].x1) f |= UGF_ASYM;  // rc(k) must link back to rc(v)
    if (f) atomicOr(g.flags, f);
}

// Rank of every read in its chain: head, dn = reads before this one, db = start of the read in the consensus.
//
// Pointer jumping over all reads costs a random gather per read and round (log2 of the longest chain: 14 rounds on 10x reads);
// here it runs on SPLITTERS only, list ranking with linear work:
//   k_rank_init    splitters = the heads and one read in 16 by a hash of its row; each gets an ordinal
//   k_rank_walk    a thread per splitter follows succ up to the next splitter (16 reads on average, one 8-byte gather each): the
//                  reads passed get (ordinal, dn, db) relative to the splitter, the splitter reached gets its jump record
//   k_jump         pointer jumping over the splitters (a 16th of the reads, L2-resident records)
//   k_rank_final   every read adds the totals of its splitter
// Reads on a cycle without a head are either never reached by a walk or keep the jumps from converging; both end in the host walk
// (the count check after k_select, the round limit).
struct alignas(16) Jmp { uint32_t ptr, dn; uint64_t db; };
__device__ __forceinline__ Jmp ld_jmp(const Jmp *p) { const uint4 a = *reinterpret_cast<const uint4 *>(p); Jmp j; j.ptr = a.x; j.dn = a.y; j.db = (uint64_t)a.w << 32 | a.z; return j; }
__device__ __forceinline__ void st_jmp(Jmp *p, const Jmp &j) { *reinterpret_cast<uint4 *>(p) = make_uint4(j.ptr, j.dn, (uint32_t)j.db, (uint32_t)(j.db >> 32)); }
__device__ __forceinline__ bool hashed_splitter(uint32_t v) { return (v * 0x9E3779B1u) >> 28 == 0; }

// rec[v]: (kNone, 0, 0) for a read that waits for a walk, (ordinal, 0, 0) for a splitter; reads without links are their own chain
__global__ void __launch_bounds__(256) k_rank_init(G g, Jmp *rec, uint32_t *list, Jmp *sj, unsigned long long *n_split) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = v < g.n;
    bool linked = false, split = false;
    if (in) {
        const uint32_t p = g.pred[v];
        linked = p != kNone || g.succ[v] != kNone;
        split = linked && (p == kNone || hashed_splitter((uint32_t)v));
    }
    const unsigned m = __ballot_sync(0xffffffffu, split);
    uint32_t base = 0;
    if ((threadIdx.x & 31) == 0 && m) base = (uint32_t)atomicAdd(n_split, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (!in) return;
    Jmp r; r.ptr = kNone; r.dn = 0; r.db = 0;
    if (split) {
        const uint32_t i = base + __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
        list[i] = (uint32_t)v;
        Jmp j; j.ptr = i; j.dn = 0; j.db = 0;       // a head; a hashed splitter is overwritten by the walk that reaches it
        st_jmp(sj + i, j);
        r.ptr = i;
    }
    if (linked) st_jmp(rec + v, r);
}

__global__ void __launch_bounds__(256) k_rank_walk(G g, uint64_t n_split, const uint32_t *__restrict__ list, Jmp *rec, Jmp *sj) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_split) return;
    uint32_t cur = list[i];
    Jmp r; r.ptr = (uint32_t)i; r.dn = 0; r.db = 0;
    for (;;) {
        const uint2 s = g.step[cur];
        if (s.x == kNone) break;                    // the tail of the chain
        ++r.dn; r.db += s.y;
        cur = s.x;
        if (hashed_splitter(cur)) { st_jmp(sj + rec[cur].ptr, r); break; }       // its ordinal was stored by k_rank_init
        st_jmp(rec + cur, r);
    }
}

__global__ void __launch_bounds__(256) k_jump(uint64_t n, const Jmp *__restrict__ J, Jmp *__restrict__ J2, uint32_t *flags) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const Jmp a = ld_jmp(J + v), b = ld_jmp(J + a.ptr);
    Jmp o;
    o.ptr = b.ptr;
    o.dn = a.dn + (a.ptr != v ? b.dn : 0u);
    o.db = a.db + (a.ptr != v ? b.db : 0ull);
    st_jmp(J2 + v, o);
    if (b.ptr != a.ptr) atomicOr(flags, (uint32_t)UGF_CHANGED);
}

__global__ void __launch_bounds__(256) k_rank_final(G g, const Jmp *__restrict__ rec, const Jmp *__restrict__ sj, const uint32_t *__restrict__ list,
                                                   uint32_t *__restrict__ head, uint32_t *__restrict__ dn, uint64_t *__restrict__ db) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= g.n) return;
    uint32_t h = (uint32_t)v, n = 0;
    uint64_t b = 0;
    if (g.pred[v] != kNone || g.succ[v] != kNone) {
        const Jmp r = ld_jmp(rec + v);
        if (r.ptr != kNone) {                       // kNone: no walk came by (a cycle without a splitter); the read stays its own head
            const Jmp j = ld_jmp(sj + r.ptr);
            h = list[j.ptr]; n = r.dn + j.dn; b = r.db + j.db;
        }
    }
    head[v] = h; dn[v] = n; db[v] = b;
}

__global__ void __launch_bounds__(256) k_tails(G g, const uint32_t *__restrict__ head) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= g.n) return;
    if (g.succ[v] == kNone && is_node(g.pack[v], v, g.min_match)) g.tail_of[head[v]] = (uint32_t)v;
}

// neighbour list the walk holds when it stops at u (a->nei after unitig_unidir, unitig.c:233-250); out may be nullptr
__device__ __forceinline__ uint32_t stop_nei(const G &g, uint64_t u, UNei *out) {
    const OvPack p = g.pack[u];
    if (p.rbeg < 0 || p.nnei == 0) return 0;
    if (p.nnei > 1) {
        if (out)
            for (uint32_t i = 0; i < p.nnei; ++i) {
                const Intv v = ld_intv(g.spill + 2 * (p.nx0 + i));
                out[i].x = v.x0; out[i].ovlp = (int64_t)(int32_t)v.info;
            }
        return p.nnei;
    }
    // one neighbour the walk did not step to; the link "b>>c>>a>>a" is cut, the others are reported (unitig.c:246)
    const bool self_rc = p.nx0 == p.x1, back_fork = p.left != 0 && g.pack[p.nx1].nnei > 1;
    if (!self_rc && !back_fork && p.nx1 == p.x1) return 0;
    if (out) { out[0].x = p.nx0; out[0].ovlp = (int64_t)p.len - p.rbeg; }
    return 1;
}

// one orientation per chain: the head h emits when h <= rc(tail)
__global__ void __launch_bounds__(256) k_select(G g, const uint32_t *__restrict__ dn, const uint64_t *__restrict__ db,
                                               uint64_t *e_cnt, uint64_t *e_len, uint64_t *e_nei, uint32_t part, uint32_t n_parts) {
    const uint64_t h0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = h0 < g.n;
    const uint64_t h = in ? h0 : 0;
    uint64_t c = 0, l = 0, m = 0;
    uint32_t reads = 0;                              // reads of the chain this thread is the head of
    const OvPack ph = g.pack[h];
    if (in && g.pred[h] == kNone && is_node(ph, h, g.min_match)) {
        const uint32_t t = g.tail_of[h];
        const OvPack pt = g.pack[t];
        reads = dn[t] + 1;
        if (h <= pt.x1 && (n_parts <= 1 || (uint32_t)(h % n_parts) == part)) {      // several GPUs: the chains are dealt out by head rank
            c = 1;
            l = db[t] + pt.len;
            m = stop_nei(g, ph.x1, nullptr) + stop_nei(g, t, nullptr);
        }
    }
    // one atomic per warp: with noisy reads most sequences are heads, and 10^7 atomics on one address serialise
    const uint32_t warp_reads = __reduce_add_sync(0xffffffffu, reads);
    if ((threadIdx.x & 31) == 0 && warp_reads) atomicAdd(g.counts + 1, (unsigned long long)warp_reads);
    if (in) { e_cnt[h] = c; e_len[h] = l; e_nei[h] = m; }
}

__global__ void __launch_bounds__(256) k_emit_meta(G g, const uint32_t *__restrict__ dn, const uint64_t *__restrict__ db, const uint64_t *__restrict__ e_cnt,
                                                  const uint64_t *__restrict__ u_idx, const uint64_t *__restrict__ u_off, const uint64_t *__restrict__ n_off,
                                                  UMeta *meta, UNei *nei) {
    const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= g.n || e_cnt[h] == 0) return;
    const OvPack ph = g.pack[h];
    const uint32_t t = g.tail_of[h];
    const OvPack pt = g.pack[t];
    UMeta m;
    // the consensus reads v1 -> vm: the record is "@v1:rc(vm)" with the neighbours of rc(v1), then those of vm
    m.k0 = h; m.k1 = pt.x1;
    m.seq_off = u_off[h]; m.nei_off = n_off[h];
    m.len = (uint32_t)(db[t] + pt.len); m.nsr = dn[t] + 1;
    m.n0 = stop_nei(g, ph.x1, nei + m.nei_off);
    m.n1 = stop_nei(g, t, nei + m.nei_off + m.n0);
    meta[u_idx[h]] = m;
}

// every read of an emitted chain: +1 / -1 of the coverage difference array, the bases it appends, and for the head its own
// sequence spelled by LF steps from its BWT row (fm_retrieve, exact.c:59-70: last base first)
__global__ void __launch_bounds__(256) k_emit_nodes(G g, OccView ix, const uint32_t *__restrict__ head, const uint64_t *__restrict__ db, const uint64_t *__restrict__ e_cnt,
                                                   const uint64_t *__restrict__ u_off, uint8_t *useq, int32_t *diff) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= g.n) return;
    const OvPack p = g.pack[v];
    if (!is_node(p, v, g.min_match)) return;
    const uint32_t h = head[v];
    if (e_cnt[h] == 0) return;
    const uint64_t base = u_off[h] + db[v];
    atomicAdd(diff + base, 1);
    atomicAdd(diff + base + p.len, -1);
    if (g.succ[v] != kNone) {
        const uint8_t *e = g.ext + p.ext_first;
        for (uint32_t i = p.len; i < p.slen; ++i) useq[base + i] = e[i - p.len];
    }
    if (h == v) {
        uint64_t k = g.row_of_rank[v];
        for (uint32_t n = 0; n < p.len; ++n) {
            const Blk B = load_blk(ix, k);
            uint32_t rel[6];
            rank_rel(B, k, rel);
            const int c = blk_symbol(B, k);
            k = ld_u64(ix.cs + (k >> kSuperShift) * 8 + c) + pick6(rel, c);
            useq[base + p.len - 1 - n] = (uint8_t)c;
        }
    }
}

// ---- MAG text on the device (mag_v_write, mag.c:149-174):  @k0:k1 \t nsr \t nei0 \t nei1 \n SEQ \n + \n COV \n   with nei = "x,ovlp;"...
// or "." -- sizes per unitig, one scan, then one thread per unitig for the header and one per eight bases for the two strings.
__device__ __forceinline__ uint32_t dec_digits(uint64_t v) {
    uint32_t n = 1;
    while (v >= 10) { v /= 10; ++n; }
    return n;
}
__device__ __forceinline__ uint32_t dec_len_i64(int64_t v) { return v < 0 ? 1 + dec_digits(0 - (uint64_t)v) : dec_digits((uint64_t)v); }
__device__ __forceinline__ char *put_dec(char *o, int64_t v) {
    uint64_t u = v < 0 ? 0 - (uint64_t)v : (uint64_t)v;
    if (v < 0) *o++ = '-';
    const uint32_t n = dec_digits(u);
    for (uint32_t i = n; i-- > 0;) { o[i] = (char)('0' + u % 10); u /= 10; }
    return o + n;
}

__global__ void __launch_bounds__(256) k_mag_len(uint64_t n_u, const UMeta *__restrict__ meta, const UNei *__restrict__ nei, uint64_t total,
                                                uint64_t *rec_len, uint32_t *hdr_len, uint64_t *soff) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u == n_u) { soff[u] = total; rec_len[u] = 0; }
    if (u >= n_u) return;
    const UMeta m = meta[u];
    uint32_t h = 1 + dec_digits(m.k0) + 1 + dec_digits(m.k1) + 1 + dec_digits(m.nsr);
    const UNei *e = nei + m.nei_off;
    for (int j = 0; j < 2; ++j) {
        const uint32_t c = j ? m.n1 : m.n0;
        h += 1;
        for (uint32_t i = 0; i < c; ++i) h += dec_len_i64((int64_t)e[i].x) + 1 + dec_len_i64((int64_t)(int32_t)e[i].ovlp) + 1;
        if (c == 0) h += 1;
        e += c;
    }
    h += 1;
    hdr_len[u] = h;
    soff[u] = m.seq_off;
    rec_len[u] = (uint64_t)h + 2ull * m.len + 4;                 // header, SEQ, "\n+\n", COV, "\n"
}

__global__ void __launch_bounds__(256) k_mag_hdr(uint64_t n_u, const UMeta *__restrict__ meta, const UNei *__restrict__ nei, const uint64_t *__restrict__ text_off,
                                                const uint32_t *__restrict__ hdr_len, char *text) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_u) return;
    const UMeta m = meta[u];
    char *o = text + text_off[u];
    *o++ = '@'; o = put_dec(o, (int64_t)m.k0); *o++ = ':'; o = put_dec(o, (int64_t)m.k1); *o++ = '\t'; o = put_dec(o, (int64_t)m.nsr);
    const UNei *e = nei + m.nei_off;
    for (int j = 0; j < 2; ++j) {
        const uint32_t c = j ? m.n1 : m.n0;
        *o++ = '\t';
        for (uint32_t i = 0; i < c; ++i) { o = put_dec(o, (int64_t)e[i].x); *o++ = ','; o = put_dec(o, (int64_t)(int32_t)e[i].ovlp); *o++ = ';'; }
        if (c == 0) *o++ = '.';
        e += c;
    }
    *o++ = '\n';
    char *q = text + text_off[u] + hdr_len[u] + m.len;
    q[0] = '\n'; q[1] = '+'; q[2] = '\n';
    q[3 + m.len] = '\n';
}

// bases -> letters, depth -> coverage character (unitig.c:253-257: '"' for the first read, +1 per further read, saturating at '~'),
// both written to their place in the text; a thread takes eight consecutive bases of the flat consensus array
__global__ void __launch_bounds__(256) k_mag_body(uint64_t total, uint64_t n_u, const int32_t *__restrict__ depth, const uint8_t *__restrict__ useq,
                                                 const uint64_t *__restrict__ soff, const uint64_t *__restrict__ text_off, const uint32_t *__restrict__ hdr_len,
                                                 char *text) {
    const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i0 >= total) return;
    uint64_t lo = 0, hi = n_u;                                    // the unitig of base i0: the last u with soff[u] <= i0
    while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (soff[mid] <= i0) lo = mid; else hi = mid;
    }
    uint64_t u = lo, end = soff[u + 1], ulen = end - soff[u];
    char *ps = text + text_off[u] + hdr_len[u] - soff[u];         // + i = place of base i in SEQ; + ulen + 3 in COV
    const uint64_t i1 = i0 + 8 < total ? i0 + 8 : total;
    for (uint64_t i = i0; i < i1; ++i) {
        while (i >= end) { ++u; end = soff[u + 1]; ulen = end - soff[u]; ps = text + text_off[u] + hdr_len[u] - soff[u]; }
        const uint8_t c = useq[i];
        const int d = depth[i];
        ps[i] = c == 1 ? 'A' : c == 2 ? 'C' : c == 3 ? 'G' : c == 4 ? 'T' : 'N';
        ps[i + ulen + 3] = (char)(33 + (d < 93 ? d : 93));
    }
}

inline unsigned nblk(uint64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

// MAG text of one part of the unitigs, held by the caller between the two steps of a multi-GPU run (all ranks learn the sizes
// of all parts before they write them side by side into one file)
// The text lives in the pinned cache of the index handle (fmg_ovcache_s::text): valid until the next unitig call on that handle.
// The MAG text of one part (multi-GPU: the chains a rank owns).  The text leaves the device in slices on a stream of its own while
// the caller exchanges the part sizes; fmg_magpart_write waits for a slice, copies it into the file, waits for the next.
struct fmg_magpart_s {
    const char *text = nullptr;      // pinned host buffer of the index handle
    uint64_t bytes = 0;
    unsigned threads = 1;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<cudaEvent_t> landed; // slice k = [bytes * k / n, bytes * (k + 1) / n) is on the host once landed[k] completed
    fmg::Dev d_text;                 // the device copy lives until the slices have landed
    ~fmg_magpart_s() {
        if (stream) { cudaSetDevice(device); cudaStreamSynchronize(stream); }
        for (auto e : landed) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }
};

// `bytes` of text to `out_path` at `offset` by `nt` threads.  The file must already have its final size (ensure_size): the
// threads copy into a shared mapping of their own slices, which -- unlike write() on one inode -- neither serialises the threads
// of a process nor the processes of a multi-GPU run that fill one file side by side.  pwrite is the fallback where a mapping is
// not possible (pipes, devices).
static int ensure_size(const char *out_path, uint64_t total, const char *who) {
    // no O_TRUNC: dropping the pages of a previous output of the same size costs more than overwriting them
    int fd = ::open(out_path, O_RDWR | O_CREAT, 0644);
    if (fd < 0 || ::ftruncate(fd, (off_t)total) != 0) {
        if (fd >= 0) ::close(fd);
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] cannot write '%s'\n", who, out_path);
        return -1;
    }
    ::close(fd);
    return 0;
}

static int write_text(const char *text, uint64_t bytes, unsigned nt, const char *out_path, uint64_t offset, const char *who) {
    int fd = ::open(out_path, O_RDWR | O_CREAT, 0644);
    if (fd < 0) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] cannot write '%s'\n", who, out_path);
        return -1;
    }
    if (bytes < (1u << 22)) nt = 1;
    std::atomic<int> io_fail{0};
    const uint64_t page = (uint64_t)::sysconf(_SC_PAGESIZE);
    static const bool use_pwrite = [] { const char *e = std::getenv("FMG_MAG_WRITE"); return e && std::strcmp(e, "pwrite") == 0; }();
    auto put = [&](unsigned t) {
        const uint64_t a = bytes * t / nt, b = bytes * (t + 1) / nt;
        if (b <= a) return;
        const uint64_t f0 = (offset + a) & ~(page - 1), span = offset + b - f0;
        void *m = use_pwrite ? MAP_FAILED : ::mmap(nullptr, span, PROT_READ | PROT_WRITE, MAP_SHARED, fd, (off_t)f0);
        if (m != MAP_FAILED) {
            std::memcpy(static_cast<char *>(m) + (offset + a - f0), text + a, b - a);
            ::munmap(m, span);
            return;
        }
        for (uint64_t done = a; done < b;) {
            const ssize_t w = ::pwrite(fd, text + done, (size_t)std::min<uint64_t>(b - done, 1u << 26), (off_t)(offset + done));
            if (w <= 0) { io_fail = 1; return; }
            done += (uint64_t)w;
        }
    };
    if (nt <= 1) put(0);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(put, t);
        for (auto &x : th) x.join();
    }
    ::close(fd);
    if (io_fail) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] short write to '%s'\n", who, out_path);
        return -1;
    }
    return 0;
}

int fmg_unitig_device(const fmg_index_s *idx, const OvDevView &D, int min_match, const char *out_path, uint64_t *n_unitigs,
                      uint32_t part, uint32_t n_parts, fmg_magpart_s *sink) {
    const uint64_t n = D.n_seq;
    if (n_unitigs) *n_unitigs = 0;
    if (n == 0 || n >= kNone) return 1;
    UG_TRY(cudaSetDevice(idx->device));
    std::lock_guard<std::mutex> ov_guard(idx->ov_lock);
    const auto t0 = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); };
    if (!idx->ovc) idx->ovc = new fmg_ovcache_s;
    fmg_ovcache_s &H = *idx->ovc;
    cudaStream_t st = nullptr;                       // legacy default stream: ordered after the pass (which synchronised its streams)
    Dev d_succ, d_pred, d_row, d_tail, d_step, d_flags, d_jmp[2], d_rec, d_list, d_ptr, d_dn, d_db, d_cnt, d_len, d_nei, d_tmp, d_meta, d_unei, d_useq, d_diff;
    UG_TRY(d_succ.alloc(n * 4)); UG_TRY(d_pred.alloc(n * 4)); UG_TRY(d_row.alloc(n * 4)); UG_TRY(d_tail.alloc(n * 4)); UG_TRY(d_flags.alloc(64));
    // jump records of the splitters (a 16th of the linked reads plus the heads: n / 2 bounds them, a chain has two reads or more and
    // the hash takes a 16th) -- sized for the worst case all the same
    for (int k = 0; k < 2; ++k) UG_TRY(d_jmp[k].alloc(n * sizeof(Jmp)));
    UG_TRY(d_rec.alloc(n * sizeof(Jmp))); UG_TRY(d_list.alloc(n * 4)); UG_TRY(d_step.alloc(n * 8));
    UG_TRY(d_ptr.alloc(n * 4)); UG_TRY(d_dn.alloc(n * 4)); UG_TRY(d_db.alloc(n * 8));
    UG_TRY(d_cnt.alloc((n + 1) * 8)); UG_TRY(d_len.alloc((n + 1) * 8)); UG_TRY(d_nei.alloc((n + 1) * 8));
    UG_TRY(H.ctrl.need(64));
    uint32_t *h_flags = H.ctrl.as<uint32_t>();
    G g;
    g.pack = static_cast<const OvPack *>(D.pack); g.spill = static_cast<const uint4 *>(D.spill); g.ext = D.ext; g.rank_of_row = D.rank;
    g.n = n; g.min_match = min_match;
    g.succ = d_succ.as<uint32_t>(); g.pred = d_pred.as<uint32_t>(); g.row_of_rank = d_row.as<uint32_t>(); g.tail_of = d_tail.as<uint32_t>();
    g.step = d_step.as<uint2>();
    g.flags = d_flags.as<uint32_t>(); g.counts = d_flags.as<unsigned long long>() + 1;
    UG_TRY(cudaMemsetAsync(d_flags.p, 0, 64, st));
    UG_TRY(cudaMemsetAsync(d_pred.p, 0xff, n * 4, st));
    UG_TRY(cudaMemsetAsync(d_tail.p, 0xff, n * 4, st));
    k_links<<<nblk(n), 256, 0, st>>>(g);
    k_pred<<<nblk(n), 256, 0, st>>>(g);
    k_rank_init<<<nblk(n), 256, 0, st>>>(g, d_rec.as<Jmp>(), d_list.as<uint32_t>(), d_jmp[0].as<Jmp>(), g.counts + 2);
    g_launches += 3;
    UG_TRY(cudaGetLastError());
    int cur = 0, rounds = 0;
    uint64_t n_split = 0;
    for (;; ++rounds) {
        UG_TRY(cudaMemcpyAsync(h_flags, d_flags.p, 32, cudaMemcpyDeviceToHost, st));
        UG_TRY(cudaStreamSynchronize(st));
        const uint32_t f = *h_flags;
        if (f & (UGF_NONINJ | UGF_ASYM | UGF_NOTNODE)) {
            if (fmg_verbose >= 3) std::fprintf(stderr, "[M::%s] irregular link graph (flags %x): falling back to the host walk\n", __func__, f);
            return 1;
        }
        if (rounds == 0) {
            n_split = reinterpret_cast<const uint64_t *>(h_flags)[3];      // counts[2]
            if (n_split == 0) break;                                       // no links at all: every read is its own chain
            k_rank_walk<<<nblk(n_split), 256, 0, st>>>(g, n_split, d_list.as<uint32_t>(), d_rec.as<Jmp>(), d_jmp[0].as<Jmp>());
            ++g_launches;
        }
        if (rounds > 0 && !(f & UGF_CHANGED)) break;
        if (rounds >= 40) {
            if (fmg_verbose >= 3) std::fprintf(stderr, "[M::%s] the link graph has a cycle: falling back to the host walk\n", __func__);
            return 1;
        }
        UG_TRY(cudaMemsetAsync(d_flags.p, 0, 4, st));
        // two jumps per flag read-back (a converged jump is the identity)
        for (int r = 0; r < 2; ++r) {
            k_jump<<<nblk(n_split), 256, 0, st>>>(n_split, d_jmp[cur].as<Jmp>(), d_jmp[cur ^ 1].as<Jmp>(), g.flags);
            ++g_launches;
            cur ^= 1;
        }
        UG_TRY(cudaGetLastError());
    }
    k_rank_final<<<nblk(n), 256, 0, st>>>(g, d_rec.as<Jmp>(), d_jmp[cur].as<Jmp>(), d_list.as<uint32_t>(), d_ptr.as<uint32_t>(), d_dn.as<uint32_t>(), d_db.as<uint64_t>());
    ++g_launches;
    const uint32_t *head = d_ptr.as<uint32_t>(), *dn = d_dn.as<uint32_t>();
    const uint64_t *db = d_db.as<uint64_t>();
    const double t_rank = since(t0);
    k_tails<<<nblk(n), 256, 0, st>>>(g, head);
    uint64_t *e_cnt = d_cnt.as<uint64_t>(), *e_len = d_len.as<uint64_t>(), *e_nei = d_nei.as<uint64_t>();
    k_select<<<nblk(n), 256, 0, st>>>(g, dn, db, e_cnt, e_len, e_nei, part, n_parts);
    g_launches += 2;
    UG_TRY(cudaGetLastError());
    // exclusive scans (n + 1 items: the last one is the total); e_cnt keeps the flags, the offsets go to the other buffers
    Dev d_uidx, d_uoff, d_noff;
    UG_TRY(d_uidx.alloc((n + 1) * 8)); UG_TRY(d_uoff.alloc((n + 1) * 8)); UG_TRY(d_noff.alloc((n + 1) * 8));
    UG_TRY(cudaMemsetAsync(e_cnt + n, 0, 8, st)); UG_TRY(cudaMemsetAsync(e_len + n, 0, 8, st)); UG_TRY(cudaMemsetAsync(e_nei + n, 0, 8, st));
    size_t need = 0;
    UG_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, e_cnt, d_uidx.as<uint64_t>(), n + 1, st));
    UG_TRY(d_tmp.alloc(need + 256));
    UG_TRY(cub::DeviceScan::ExclusiveSum(d_tmp.p, need, e_cnt, d_uidx.as<uint64_t>(), n + 1, st));
    UG_TRY(cub::DeviceScan::ExclusiveSum(d_tmp.p, need, e_len, d_uoff.as<uint64_t>(), n + 1, st));
    UG_TRY(cub::DeviceScan::ExclusiveSum(d_tmp.p, need, e_nei, d_noff.as<uint64_t>(), n + 1, st));
    g_launches += 3;
    uint64_t *h_tot = H.ctrl.as<uint64_t>() + 1;
    UG_TRY(cudaMemcpyAsync(h_tot + 3, d_flags.as<uint64_t>() + 1, 16, cudaMemcpyDeviceToHost, st));
    UG_TRY(cudaMemcpyAsync(h_tot + 0, d_uidx.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
    UG_TRY(cudaMemcpyAsync(h_tot + 1, d_uoff.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
    UG_TRY(cudaMemcpyAsync(h_tot + 2, d_noff.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
    UG_TRY(cudaStreamSynchronize(st));
    const uint64_t n_u = h_tot[0], total = h_tot[1], n_nei = h_tot[2];
    if (h_tot[3] != h_tot[4]) {
        if (fmg_verbose >= 3)
            std::fprintf(stderr, "[M::%s] %llu of %llu reads lie on cycles of the link graph: falling back to the host walk\n", __func__,
                         (unsigned long long)(h_tot[3] - h_tot[4]), (unsigned long long)h_tot[3]);
        return 1;
    }
    UG_TRY(d_meta.alloc(std::max<uint64_t>(n_u, 1) * sizeof(UMeta))); UG_TRY(d_unei.alloc(std::max<uint64_t>(n_nei, 1) * sizeof(UNei)));
    UG_TRY(d_useq.alloc(total + 1)); UG_TRY(d_diff.alloc((total + 1) * 4));
    UG_TRY(cudaMemsetAsync(d_diff.p, 0, (total + 1) * 4, st));
    k_emit_meta<<<nblk(n), 256, 0, st>>>(g, dn, db, e_cnt, d_uidx.as<uint64_t>(), d_uoff.as<uint64_t>(), d_noff.as<uint64_t>(), d_meta.as<UMeta>(), d_unei.as<UNei>());
    k_emit_nodes<<<nblk(n), 256, 0, st>>>(g, idx->view, head, db, e_cnt, d_uoff.as<uint64_t>(), d_useq.as<uint8_t>(), d_diff.as<int32_t>());
    g_launches += 2;
    UG_TRY(cudaGetLastError());
    if (total) {
        size_t need2 = 0;
        UG_TRY(cub::DeviceScan::InclusiveSum(nullptr, need2, d_diff.as<int32_t>(), d_diff.as<int32_t>(), total, st));
        if (need2 > need) UG_TRY(d_tmp.alloc(need2 + 256));
        UG_TRY(cub::DeviceScan::InclusiveSum(d_tmp.p, need2, d_diff.as<int32_t>(), d_diff.as<int32_t>(), total, st));
        ++g_launches;
    }
    // ---- MAG text (mag_v_write, mag.c:149-174) formatted on the device: record sizes, one scan, headers, the two strings
    Dev d_rlen, d_toff, d_hlen, d_soff, d_text;
    UG_TRY(d_rlen.alloc((n_u + 1) * 8)); UG_TRY(d_toff.alloc((n_u + 1) * 8)); UG_TRY(d_hlen.alloc((n_u + 1) * 4)); UG_TRY(d_soff.alloc((n_u + 1) * 8));
    k_mag_len<<<nblk(n_u + 1), 256, 0, st>>>(n_u, d_meta.as<UMeta>(), d_unei.as<UNei>(), total, d_rlen.as<uint64_t>(), d_hlen.as<uint32_t>(), d_soff.as<uint64_t>());
    {
        size_t need3 = 0;
        UG_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need3, d_rlen.as<uint64_t>(), d_toff.as<uint64_t>(), n_u + 1, st));
        UG_TRY(d_tmp.alloc(need3 + 256));
        UG_TRY(cub::DeviceScan::ExclusiveSum(d_tmp.p, need3, d_rlen.as<uint64_t>(), d_toff.as<uint64_t>(), n_u + 1, st));
    }
    g_launches += 2;
    UG_TRY(cudaMemcpyAsync(h_tot + 5, d_toff.as<uint64_t>() + n_u, 8, cudaMemcpyDeviceToHost, st));
    UG_TRY(cudaStreamSynchronize(st));
    const uint64_t text_bytes = h_tot[5];
    UG_TRY(d_text.alloc(text_bytes + 1));
    if (n_u) {
        k_mag_hdr<<<nblk(n_u), 256, 0, st>>>(n_u, d_meta.as<UMeta>(), d_unei.as<UNei>(), d_toff.as<uint64_t>(), d_hlen.as<uint32_t>(), d_text.as<char>());
        if (total)
            k_mag_body<<<nblk((total + 7) / 8), 256, 0, st>>>(total, n_u, d_diff.as<int32_t>(), d_useq.as<uint8_t>(), d_soff.as<uint64_t>(), d_toff.as<uint64_t>(),
                                                          d_hlen.as<uint32_t>(), d_text.as<char>());
        g_launches += 2;
        UG_TRY(cudaGetLastError());
    }
    UG_TRY(H.text.need(text_bytes + 1));
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency() / std::max(1u, n_parts), 32u));
    if (const char *e = std::getenv("FMG_THREADS")) nt = (unsigned)std::max(1, std::atoi(e));
    const bool to_stdout = !sink && std::strcmp(out_path, "-") == 0;
    double t_dev = 0;
    if (sink) {
        // asynchronous hand-over: slices on the part's own stream, ordered after the formatting kernels
        sink->text = H.text.as<char>(); sink->bytes = text_bytes; sink->threads = nt; sink->device = idx->device;
        UG_TRY(cudaStreamCreateWithFlags(&sink->stream, cudaStreamNonBlocking));
        cudaEvent_t formatted;
        UG_TRY(cudaEventCreateWithFlags(&formatted, cudaEventDisableTiming));
        UG_TRY(cudaEventRecord(formatted, st));
        UG_TRY(cudaStreamWaitEvent(sink->stream, formatted, 0));
        cudaEventDestroy(formatted);
        const unsigned n_slice = text_bytes < (1u << 24) ? 1u : std::max(2u, std::min(nt, 8u));
        sink->landed.resize(n_slice);
        for (auto &e : sink->landed) UG_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (unsigned k = 0; k < n_slice; ++k) {
            const uint64_t a = text_bytes * k / n_slice, b = text_bytes * (k + 1) / n_slice;
            if (b > a) UG_TRY(cudaMemcpyAsync(H.text.as<char>() + a, d_text.as<char>() + a, b - a, cudaMemcpyDeviceToHost, sink->stream));
            UG_TRY(cudaEventRecord(sink->landed[k], sink->stream));
        }
        sink->d_text.swap(d_text);
        UG_TRY(cudaStreamSynchronize(st));          // the scratch of this call returns to the pool when it ends: its kernels must be done
        t_dev = since(t0);
    } else if (to_stdout || text_bytes < (1u << 24)) {
        if (text_bytes) UG_TRY(cudaMemcpyAsync(H.text.p, d_text.p, text_bytes, cudaMemcpyDeviceToHost, st));
        UG_TRY(cudaStreamSynchronize(st));
        t_dev = since(t0);
        if (to_stdout) { std::fwrite(H.text.p, 1, text_bytes, stdout); std::fflush(stdout); }
        else if (ensure_size(out_path, text_bytes, __func__) != 0 || write_text(H.text.as<char>(), text_bytes, nt, out_path, 0, __func__) != 0) return -1;
    } else {
        // a file: the text leaves the device in slices, and each slice is written while the next one is copied
        if (ensure_size(out_path, text_bytes, __func__) != 0) return -1;
        const unsigned n_slice = std::max(2u, std::min(nt, 16u));
        std::vector<cudaEvent_t> ev(n_slice);
        for (auto &e : ev) UG_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (unsigned k = 0; k < n_slice; ++k) {
            const uint64_t a = text_bytes * k / n_slice, b = text_bytes * (k + 1) / n_slice;
            UG_TRY(cudaMemcpyAsync(H.text.as<char>() + a, d_text.as<char>() + a, b - a, cudaMemcpyDeviceToHost, st));
            UG_TRY(cudaEventRecord(ev[k], st));
        }
        std::atomic<int> fail{0};
        std::vector<std::thread> th;
        for (unsigned k = 0; k < n_slice; ++k)
            th.emplace_back([&, k]() {
                const uint64_t a = text_bytes * k / n_slice, b = text_bytes * (k + 1) / n_slice;
                if (cudaEventSynchronize(ev[k]) != cudaSuccess) { fail = 1; return; }
                if (write_text(H.text.as<char>() + a, b - a, 1, out_path, a, "fmg_unitig_device") != 0) fail = 1;
            });
        for (auto &x : th) x.join();
        for (auto &e : ev) cudaEventDestroy(e);
        t_dev = since(t0);
        if (fail) return -1;
    }
    if (n_unitigs) *n_unitigs = n_u;
    if (fmg_verbose >= 4)
        std::fprintf(stderr, "[M::%s] %llu unitigs, %llu bases from %llu sequences: chains %.3f s (%d jump rounds), assembly + copies %.3f s, text %.3f s\n", __func__,
                     (unsigned long long)n_u, (unsigned long long)total, (unsigned long long)n, t_rank, 2 * rounds, t_dev - t_rank, since(t0) - t_dev);
    return 0;
}

extern "C" {

// fm6_unitig (unitig.c:378-407) + main_unitig (cmd.c:184-216): overlap records of every sequence on the GPU, then the
// walk; MAG records go to `out_path` ("-" = stdout).  max_len = upper bound of the sequence length in the index
// (0: estimate from the symbol counts, grown on demand).
int fmg_unitig(const fmg_index_t *idx, int min_match, int max_len, const char *out_path, uint64_t *n_unitigs) {
    if (!idx) return -1;
    if (n_unitigs) *n_unitigs = 0;
    const auto t0 = std::chrono::steady_clock::now();
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    // 1. records stay in HBM and the unitigs are assembled there (unitig_gpu.cu)
    const char *force_host = std::getenv("FMG_UNITIG_HOST");
    if (!force_host || !*force_host || *force_host == '0') {
        int rc;
        {
            fmg::OvDevice D;
            rc = fmg_overlap_pass(idx, min_match, max_len, &D, nullptr);
            if (rc != 0) return rc;
            const auto t1 = std::chrono::steady_clock::now();
            const fmg::OvDevView V = {D.pack.p, D.rank.as<int64_t>(), D.ext.as<uint8_t>(), D.spill.p, D.n_seq, D.ext_total, D.spill_total};
            rc = fmg_unitig_device(idx, V, min_match, out_path, n_unitigs, 0, 1, nullptr);
            if (rc == 0 && fmg_verbose >= 3)
                std::fprintf(stderr, "[M::%s] %llu sequences: overlap records %.3f s, unitig assembly + output %.3f s (GPU)\n", __func__,
                             (unsigned long long)D.n_seq, secs(t0, t1), secs(t1, std::chrono::steady_clock::now()));
            max_len = D.max_len;
        }
        if (rc <= 0) return rc;
    }
    // 2. an irregular link graph (or FMG_UNITIG_HOST=1): records to the host, walk in the reference's seed order
    const auto t1 = std::chrono::steady_clock::now();
    OvHost R;
    const int rc0 = fmg_overlap_all(idx, min_match, max_len, &R);
    if (rc0 != 0) return rc0;
    const auto t2 = std::chrono::steady_clock::now();
    const int rc = fmg_unitig_walk(R, min_match, out_path, n_unitigs);
    if (fmg_verbose >= 3)
        std::fprintf(stderr, "[M::%s] %llu sequences: overlap records %.3f s (GPU, incl. copies), unitig walk + output %.3f s (host)\n", __func__,
                     (unsigned long long)R.n_seq, secs(t1, t2), secs(t2, std::chrono::steady_clock::now()));
    return rc;
}

} // extern "C"

// ---- multi-GPU `fermi unitig` (fermi_b200/parallel.py drives it with NCCL through torch.distributed):
//   fmg_overlap_shard      every GPU: records of its rows, in row order, into caller-owned device buffers
//   -- one all-gather of the record / rank / ext / spill shards, each padded to the largest shard --
//   fmg_overlap_merge      every GPU: gathered shards -> the rank-indexed record array the assembly chases
//   fmg_overlap_left_fix_rows / _flags  every GPU: the deferred left check (overlap.cu: left_fix) for its own rows of the merged
//                          array, then one all-gather of a byte per row (fmg_overlap_left_fix does all rows on one GPU)
//   fmg_unitig_part        every GPU: link graph + list ranking over all records, then emission and MAG text of the chains
//                          it owns (head rank % n_parts == part)
//   -- all-gather of the text sizes --
//   fmg_magpart_write      every GPU: its text at its offset of the one output file
struct MergeArgs {
    int n_shards;
    uint64_t rows[64];              // rows of each shard (shards are consecutive row ranges, in order)
    uint64_t row_pad, ext_pad, spill_pad;   // slot sizes of the gathered arrays (rows, bytes, entries per shard)
};
__global__ void __launch_bounds__(256) k_ov_merge(MergeArgs M, const OvPack *__restrict__ rec_all, const int64_t *__restrict__ rank_all, uint64_t n_seq,
                                                 OvPack *pack, int64_t *rank_of_row, uint32_t *flags) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;       // index into the padded gathered arrays
    const uint64_t s = i / M.row_pad, t = i - s * M.row_pad;
    if (s >= (uint64_t)M.n_shards || t >= M.rows[s]) return;
    uint64_t row = t;
    for (uint64_t q = 0; q < s; ++q) row += M.rows[q];
    const int64_t r = rank_all[i];
    rank_of_row[row] = r;
    OvPack p = rec_all[i];
    if (p.nnei == 1) p.ext_first += s * M.ext_pad;
    else if (p.nnei > 1) p.nx0 += s * M.spill_pad;
    if ((uint64_t)r >= n_seq) { atomicOr(flags, 1u); return; }
    uint4 *dst = reinterpret_cast<uint4 *>(pack + r);
    const uint4 *src = reinterpret_cast<const uint4 *>(&p);
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
}

int fmg_overlap_left_fix_dev(const fmg_index_s *idx, int min_match, int max_len, void *d_pack, const int64_t *d_rank_of_row, uint64_t row_lo, uint64_t row_hi, uint64_t *n_left_out);
int fmg_overlap_left_flags_dev(const fmg_index_s *idx, void *d_pack, const int64_t *d_rank_of_row, uint64_t row_lo, uint64_t row_hi, int8_t *d_flags, int apply);

extern "C" {

int fmg_overlap_shard(const fmg_index_t *idx, int min_match, int max_len, uint64_t row_lo, uint64_t row_hi, void *d_rec, int64_t *d_rank,
                      uint8_t *d_ext, uint64_t ext_cap, void *d_spill, uint64_t spill_cap, uint64_t totals[2]) {
    if (!idx || !d_rec || !d_rank || !d_ext || !d_spill || !totals) return -1;
    fmg::OvShard S;
    S.row_lo = row_lo; S.row_hi = row_hi; S.pack = d_rec; S.rank = d_rank; S.ext = d_ext; S.ext_cap = ext_cap; S.spill = d_spill; S.spill_cap = spill_cap;
    S.ext_total = S.spill_total = 0; S.max_len = 0;
    const int rc = fmg_overlap_pass(idx, min_match, max_len, nullptr, nullptr, &S);
    totals[0] = S.ext_total; totals[1] = S.spill_total;
    return rc;
}

int fmg_overlap_merge(const fmg_index_t *idx, int n_shards, const uint64_t *rows, uint64_t row_pad, uint64_t ext_pad, uint64_t spill_pad,
                      const void *d_rec_all, const int64_t *d_rank_all, void *d_pack, int64_t *d_rank_of_row) {
    if (!idx || n_shards < 1 || n_shards > 64 || !rows || !d_rec_all || !d_rank_all || !d_pack || !d_rank_of_row) return -1;
    UG_TRY(cudaSetDevice(idx->device));
    const uint64_t n_seq = idx->mcnt[1];
    MergeArgs M;
    M.n_shards = n_shards; M.row_pad = row_pad ? row_pad : 1; M.ext_pad = ext_pad; M.spill_pad = spill_pad;
    uint64_t tot = 0;
    for (int q = 0; q < n_shards; ++q) { M.rows[q] = rows[q]; tot += rows[q]; if (rows[q] > M.row_pad) return -1; }
    if (tot != n_seq) return -1;
    Dev d_flag;
    UG_TRY(d_flag.alloc(4));
    UG_TRY(cudaMemset(d_flag.p, 0, 4));
    UG_TRY(cudaMemset(d_pack, 0, n_seq * sizeof(OvPack)));
    k_ov_merge<<<nblk(M.row_pad * (uint64_t)n_shards), 256>>>(M, static_cast<const OvPack *>(d_rec_all), d_rank_all, n_seq, static_cast<OvPack *>(d_pack), d_rank_of_row,
                                                           d_flag.as<uint32_t>());
    ++g_launches;
    UG_TRY(cudaGetLastError());
    uint32_t f = 0;
    UG_TRY(cudaMemcpy(&f, d_flag.p, 4, cudaMemcpyDeviceToHost));
    return f ? -1 : 0;
}

int fmg_overlap_left_fix(const fmg_index_t *idx, int min_match, int max_len, void *d_pack, const int64_t *d_rank_of_row, uint64_t *n_left) {
    return idx ? fmg_overlap_left_fix_dev(idx, min_match, max_len, d_pack, d_rank_of_row, 0, idx->mcnt[1], n_left) : -1;
}
int fmg_overlap_left_fix_rows(const fmg_index_t *idx, int min_match, int max_len, void *d_pack, const int64_t *d_rank_of_row, uint64_t row_lo, uint64_t row_hi,
                              uint64_t *n_left) {
    return fmg_overlap_left_fix_dev(idx, min_match, max_len, d_pack, d_rank_of_row, row_lo, row_hi, n_left);
}
int fmg_overlap_left_flags(const fmg_index_t *idx, void *d_pack, const int64_t *d_rank_of_row, uint64_t row_lo, uint64_t row_hi, int8_t *d_flags, int apply) {
    return fmg_overlap_left_flags_dev(idx, d_pack, d_rank_of_row, row_lo, row_hi, d_flags, apply);
}

int fmg_unitig_from_device(const fmg_index_t *idx, int min_match, const void *d_pack, const int64_t *d_rank, const uint8_t *d_ext, uint64_t ext_total,
                           const void *d_spill, uint64_t spill_total, const char *out_path, uint64_t *n_unitigs) {
    if (!idx || !d_pack || !d_rank || !out_path) return -1;
    UG_TRY(cudaSetDevice(idx->device));
    UG_TRY(cudaDeviceSynchronize());
    const fmg::OvDevView V = {d_pack, d_rank, d_ext, d_spill, idx->mcnt[1], ext_total, spill_total};
    return fmg_unitig_device(idx, V, min_match, out_path, n_unitigs, 0, 1, nullptr);
}

int fmg_unitig_part(const fmg_index_t *idx, int min_match, const void *d_pack, const int64_t *d_rank_of_row, const uint8_t *d_ext, const void *d_spill,
                    int part, int n_parts, fmg_magpart_t **out, uint64_t *n_unitigs, uint64_t *n_bytes) {
    if (!idx || !d_pack || !d_rank_of_row || !out || part < 0 || n_parts < 1 || part >= n_parts) return -1;
    *out = nullptr;
    UG_TRY(cudaSetDevice(idx->device));
    UG_TRY(cudaDeviceSynchronize());
    fmg_magpart_s *sink = new fmg_magpart_s;
    const fmg::OvDevView V = {d_pack, d_rank_of_row, d_ext, d_spill, idx->mcnt[1], 0, 0};
    const int rc = fmg_unitig_device(idx, V, min_match, "", n_unitigs, (uint32_t)part, (uint32_t)n_parts, sink);
    if (rc != 0) { delete sink; return rc; }
    if (n_bytes) *n_bytes = sink->bytes;
    *out = sink;
    return 0;
}

int fmg_magpart_write(const fmg_magpart_t *p, const char *path, uint64_t offset, uint64_t total_bytes) {
    if (!p || !path) return -1;
    if (total_bytes && ensure_size(path, total_bytes, __func__) != 0) return -1;
    if (cudaSetDevice(p->device) != cudaSuccess) return -1;
    const unsigned n_slice = (unsigned)p->landed.size();
    if (n_slice <= 1) {
        if (n_slice == 1 && cudaEventSynchronize(p->landed[0]) != cudaSuccess) return -1;
        return write_text(p->text, p->bytes, p->threads, path, offset, __func__);
    }
    // a thread per slice: each waits for its slice to land and copies it into the file while the later ones are still in flight
    std::atomic<int> fail{0};
    std::vector<std::thread> th;
    for (unsigned k = 0; k < n_slice; ++k)
        th.emplace_back([&, k]() {
            const uint64_t a = p->bytes * k / n_slice, b = p->bytes * (k + 1) / n_slice;
            if (cudaSetDevice(p->device) != cudaSuccess || cudaEventSynchronize(p->landed[k]) != cudaSuccess) { fail = 1; return; }
            if (b > a && write_text(p->text + a, b - a, 1, path, offset + a, "fmg_magpart_write") != 0) fail = 1;
        });
    for (auto &x : th) x.join();
    return fail ? -1 : 0;
}

void fmg_magpart_free(fmg_magpart_t *p) { delete p; }

} // extern "C"
