// `fermi unitig` on top of the GPU overlap records (host side of the path).
//
// The reference walks unitigs seed by seed and asks the FMD-index at every step (unitig.c:227-362).
// Every such question is a pure function of one read -- its right neighbours, the consensus extension
// and the simple left check (fmd_overlap.cuh) -- so here the index work is done once per sequence on
// the GPU (fmg_overlap_batch) and the walk below only chases records.  It reproduces the single-thread
// seed order of unitig_core (unitig.c:319-362) including the `used` / `bend` / `visited` bitmaps, so the
// MAG records equal those of `fermi unitig -t1` (as a canonicalised set: SURVEY.md A.8).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <chrono>
#include <memory>
#include <thread>
#include <mutex>
#include "fmd_overlap.cuh"
#include "ov_records.hpp"
#include "fmg_internal.hpp"
#include "../../include/fermi_b200.h"

using namespace fmg;

namespace {

struct Bits {                                   // shared by the walker threads like the reference's bitmaps (unitig.c:15-20)
    std::vector<uint64_t> w;
    explicit Bits(uint64_t n) : w((n + 63) / 64, 0) {}
    bool get(uint64_t i) const { return __atomic_load_n(&w[i >> 6], __ATOMIC_RELAXED) >> (i & 63) & 1; }
    bool test_and_set(uint64_t i) { const uint64_t m = 1ull << (i & 63); return __atomic_fetch_or(&w[i >> 6], m, __ATOMIC_RELAXED) & m; }
    void set(uint64_t i) {
        const uint64_t m = 1ull << (i & 63);
        if (!(__atomic_load_n(&w[i >> 6], __ATOMIC_RELAXED) & m)) __atomic_fetch_or(&w[i >> 6], m, __ATOMIC_RELAXED);
    }
    void set_intv(uint64_t x0, uint64_t x1, uint64_t x2) {           // set_bits, unitig.c:22-36
        for (uint64_t k = 0; k < x2; ++k) set(x0 + k), set(x1 + k);
    }
};

struct Nei { uint64_t x; uint64_t y; };

// Records are indexed by the rank of the sequence among all sequences (LF of a sentinel counts the '$' above it):
// the ids the walk sees -- fm_retrieve's return value, intv0.x[], neighbour x[] -- address OvHost::pack directly.
// Only the seeds are BWT rows; they go through rank_of_row once.
struct Walker {
    const OvHost &R;
    int min_match;
    Bits &used, &bend, &visited;
    std::string s, cov;
    std::vector<Nei> last_nei;           // a->nei after unitig_unidir

    Walker(const OvHost &r, int mm, Bits &u, Bits &b, Bits &v) : R(r), min_match(mm), used(u), bend(b), visited(v) {}

    // unitig_unidir, unitig.c:227-262.  `cur` = rank of the last read of s, which starts at s[beg].
    int unidir(uint64_t cur, int beg, uint64_t k0, uint64_t *end, int *is_loop) {
        int ori_l = (int)s.size(), n_reads = 0;
        *is_loop = 0;
        last_nei.clear();
        for (;;) {
            const OvPack &p = R.pack[cur];
            last_nei.clear();
            if (p.rbeg < 0 || p.nnei == 0) break;                          // try_right() < 0
            if (p.nnei > 1) {                                              // forward bifurcation
                const fmg_intv_t *nb = R.spill + p.nx0;
                for (int k = 0; k < (int)p.nnei; ++k) last_nei.push_back(Nei{nb[k].x[0], nb[k].info});
                bend.set(*end);
                break;
            }
            const uint64_t k = p.nx0;
            last_nei.push_back(Nei{k, (uint64_t)((int64_t)p.len - p.rbeg)});
            const int rbeg = beg + p.rbeg;
            if (k == *end) break;                                          // a loop like b>>c>>a><a
            bool back_fork = bend.get(k);
            if (!back_fork && p.left != 0)
                // check_left (unitig.c:206-225): the simple test failed; confirm with the right neighbours of
                // the reverse complement of the neighbour
                back_fork = R.pack[p.nx1].nnei > 1;
            if (back_fork) { bend.set(k); break; }                         // backward bifurcation
            if (k == k0) { *is_loop = 1; break; }                          // a loop like a>>b>>c>>a
            if (p.nx1 == *end) { last_nei.clear(); break; }                // a loop like b>>c>>a>>a; cut the last link
            *end = p.nx1;
            __builtin_prefetch(&R.pack[k]);
            used.set_intv(p.nx0, p.nx1, p.nx2);
            ++n_reads;
            // the consensus grows by the extension recorded for `cur` (unitig.c:141,253-257)
            const int new_l = beg + (int)p.slen;
            const uint8_t *ext = R.ext + p.ext_first;
            s.resize(new_l); cov.resize(new_l);
            for (int i = ori_l; i < new_l; ++i) s[i] = (char)ext[i - ori_l], cov[i] = '"';
            for (int i = rbeg; i < ori_l; ++i) cov[i] += cov[i] != '~';
            beg = rbeg; ori_l = new_l; cur = k;
        }
        s.resize(ori_l); cov.resize(ori_l);
        return n_reads;
    }

    // unitig1, unitig.c:274-317
    int unitig1(uint64_t seed, uint64_t end[2], std::vector<Nei> nei[2], int *n_reads) {
        const uint64_t k = R.rank_of_row[seed];
        *n_reads = 0; nei[0].clear(); nei[1].clear();
        if (used.get(k)) {                                                 // -2, unless the read is too short (-1); both skip the seed
            return -2;
        }
        const OvPack &p = R.pack[k];
        const int seed_len = (int)p.len;
        if (seed_len <= min_match) return -1;                              // too short
        used.set_intv(p.x0, p.x1, p.x2);
        if (p.contained < 0) return -3;
        *n_reads = 1;
        const uint8_t *sq = R.seq + (R.seq_odd_only ? seed >> 1 : seed) * R.seq_stride;
        s.assign((const char *)sq, seed_len);
        cov.assign(seed_len, '"');
        end[0] = p.x1; end[1] = p.x0;
        int is_loop = 0;
        // (the reference skips this call when the read has no overlap candidate at all; the call is then a no-op)
        *n_reads += unidir(k, 0, p.x0, &end[0], &is_loop);
        nei[0] = last_nei;
        if (is_loop) {
            nei[1].push_back(Nei{end[0], last_nei[0].y});
            return 0;
        }
        // the other direction: reverse complement the consensus, reverse the coverage
        std::reverse(s.begin(), s.end());
        for (auto &c : s) c = (c >= 1 && c <= 4) ? 5 - c : c;
        std::reverse(cov.begin(), cov.end());
        *n_reads += unidir(p.x1, (int)s.size() - seed_len, p.x1, &end[1], &is_loop);
        nei[1] = last_nei;
        return 0;
    }
};


void append_i64(std::string &o, int64_t v) {
    char b[24];
    int n = 24;
    uint64_t u = v < 0 ? 0 - (uint64_t)v : (uint64_t)v;
    do { b[--n] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) b[--n] = '-';
    o.append(b + n, 24 - n);
}

// mag_v_write, mag.c:149-174
void write_mag(std::string &o, const uint64_t k[2], int nsr, const std::vector<Nei> nei[2], const std::string &seq, const std::string &cov) {
    o += '@'; append_i64(o, (int64_t)k[0]); o += ':'; append_i64(o, (int64_t)k[1]); o += '\t'; append_i64(o, nsr);
    for (int j = 0; j < 2; ++j) {
        o += '\t';
        for (const Nei &z : nei[j]) { append_i64(o, (int64_t)z.x); o += ','; append_i64(o, (int32_t)z.y); o += ';'; }
        if (nei[j].empty()) o += '.';
    }
    o += '\n';
    for (char c : seq) o += "ACGT"[(int)c - 1];
    o += "\n+\n";
    o += cov;
    o += '\n';
}

} // namespace

// The walk alone (host code): MAG records from the packed overlap records of ALL sequences of an index.
// Strided worker threads over the seeds, sharing the three bitmaps through atomics: the scheme of fm6_unitig
// (unitig.c:378-407).  One thread reproduces `fermi unitig -t1` record for record; with more threads the record
// order and orientation vary but the canonicalised set does not.
int fmg_unitig_walk(const OvHost &R, int min_match, const char *out_path, uint64_t *n_unitigs) {
    const uint64_t n_seq = R.n_seq;
    FILE *fp = std::strcmp(out_path, "-") ? std::fopen(out_path, "wb") : stdout;
    if (!fp) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] cannot write '%s'\n", __func__, out_path);
        return -1;
    }
    const auto tw0 = std::chrono::steady_clock::now();
    int n_threads = 1;
    // one thread by default, like `fermi unitig` (-t 1, cmd.c:186): on an irregular link graph the result depends on the seed
    // order, and only the single-thread walk reproduces the reference record for record; FMG_THREADS / -t opts in to more
    if (const char *e = std::getenv("FMG_THREADS")) n_threads = std::atoi(e);
    if (n_threads < 1 || n_seq < 64) n_threads = 1;
    Bits used(n_seq), bend(n_seq), visited(n_seq);
    std::vector<uint64_t> counts(n_threads, 0);
    std::mutex out_lock;
    auto work = [&](int tid) {
        Walker W(R, min_match, used, bend, visited);
        std::string out;
        out.reserve(1 << 20);
        uint64_t end[2];
        std::vector<Nei> nb[2];
        int n_reads;
        // seeds = odd sentinel ranks in the order of unitig_core (unitig.c:333-334), start = tid, step = n_threads
        for (uint64_t j = tid; j <= n_seq >> 2; j += n_threads)
            for (uint64_t i = j << 2 | 1; i < (j << 2) + 4 && i < n_seq; i += 2) {
                if (W.unitig1(i, end, nb, &n_reads) < 0) continue;
                // unitig.c:337-339: the first walk that claims both end ids emits the unitig.  One thread claims in the
                // reference's order (k[0] then k[1], short-circuit).  With several threads the reference's order lets two
                // walks of the same unitig in opposite orientations knock each other out (each claims its own k[0] first
                // and then finds the other's); claiming the smaller id first makes exactly one of them win.
                const int f = (n_threads > 1 && end[1] < end[0]) ? 1 : 0;
                if (visited.test_and_set(end[f]) || visited.test_and_set(end[f ^ 1])) continue;
                write_mag(out, end, n_reads, nb, W.s, W.cov);
                ++counts[tid];
                if (out.size() > (1 << 20)) {
                    std::lock_guard<std::mutex> g(out_lock);
                    std::fwrite(out.data(), 1, out.size(), fp);
                    out.clear();
                }
            }
        std::lock_guard<std::mutex> g(out_lock);
        std::fwrite(out.data(), 1, out.size(), fp);
    };
    if (n_threads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
        for (auto &t : th) t.join();
    }
    uint64_t count = 0;
    for (uint64_t c : counts) count += c;
    if (fp != stdout) std::fclose(fp); else std::fflush(fp);
    if (n_unitigs) *n_unitigs = count;
    if (fmg_verbose >= 4)
        std::fprintf(stderr, "[M::%s] %llu unitigs from %llu sequences in %.3f s (%d threads)\n", __func__, (unsigned long long)count, (unsigned long long)n_seq,
                     std::chrono::duration<double>(std::chrono::steady_clock::now() - tw0).count(), n_threads);
    return 0;
}

extern "C" {

// The walk over records that arrive as separate arrays (rec: n_seq x 10, nei/nei_off, seq/ext: n_seq x max_len; the layout
// fmg_overlap_batch produces, e.g. after the all-gather of per-GPU shards): packs them by rank, then walks.  Host code.
int fmg_unitig_assemble(uint64_t n_seq, int max_len, int min_match, const int64_t *rec, const fmg_intv_t *nei,
                        const uint64_t *nei_off, const uint8_t *seq, const uint8_t *ext, const char *out_path, uint64_t *n_unitigs) {
    std::unique_ptr<OvPack[]> pack(new OvPack[n_seq ? n_seq : 1]);
    std::vector<uint64_t> rank_of_row(n_seq);
    std::vector<uint8_t> xt;
    std::vector<fmg_intv_t> spill;
    for (uint64_t t = 0; t < n_seq; ++t) {
        const int64_t *r = rec + t * OV_NREC;
        const uint64_t k = (uint64_t)r[OV_K];
        if (k >= n_seq) {
            if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] record %llu has rank %llu >= %llu sequences\n", __func__, (unsigned long long)t, (unsigned long long)k, (unsigned long long)n_seq);
            return -1;
        }
        rank_of_row[t] = k;
        const fmg_intv_t *nb = nei + nei_off[t];
        uint64_t nx0 = 0, nx1 = 0, nx2 = 0;
        if (r[OV_NNEI] == 1) nx0 = nb[0].x[0], nx1 = nb[0].x[1], nx2 = nb[0].x[2];
        else if (r[OV_NNEI] > 1) { nx0 = spill.size(); spill.insert(spill.end(), nb, nb + r[OV_NNEI]); }
        const uint32_t el = ov_ext_len(r);
        if (!ov_pack(r, nx0, nx1, nx2, xt.size(), &pack[k])) {
            if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] record %llu does not fit the packed layout\n", __func__, (unsigned long long)t);
            return -1;
        }
        xt.insert(xt.end(), ext + t * (uint64_t)max_len, ext + t * (uint64_t)max_len + el);
    }
    OvHost R;
    R.n_seq = n_seq; R.max_len = max_len; R.pack = pack.get(); R.rank_of_row = rank_of_row.data();
    R.seq = seq; R.seq_stride = (uint64_t)max_len; R.seq_odd_only = 0;
    R.ext = xt.data(); R.spill = spill.data(); R.ext_total = xt.size(); R.spill_total = spill.size();
    return fmg_unitig_walk(R, min_match, out_path, n_unitigs);
}

} // extern "C"
