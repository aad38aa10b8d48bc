// Device-side core of the overlap path of `fermi unitig` (unitig.c:38-204):
//   OvLane          per sequence: fm_retrieve (exact.c:59-70) fused with fm6_is_contained (unitig.c:77-91) -> fm6_get_nei
//                   (unitig.c:93-179) -> check_left_simple (unitig.c:186-204), in four phases (see OvLane)
// Everything the unitig walk (unitig_unidir / unitig1, unitig.c:227-317) asks the index is a pure function
// of ONE read: its right neighbours, the consensus extension towards them and the simple left check of a
// unique neighbour.  The GPU computes that record for every sequence of the index in parallel; the walk
// itself then only chases these records (unitig_host.cpp).
//
// Control flow is written as the natural nested loops of the reference.  In the list-chasing phases the
// lanes of a warp are on different sequences and in different loops, so every extension goes through
// ext_sync(): a NOINLINE function, i.e. one copy of the code that all call sites jump to, which starts with
// a full-warp vote.  The vote lines the 32 lanes up in time at the same program counter, so the expensive
// part (block loads + popcounts) executes converged, whatever loop each lane came from.
//
// Same source compiles for the host (tests/emu) -- checker only, never linked by the product.
#pragma once
#include "fmd_device.cuh"

#if defined(__CUDACC__)
#define FMG_NOINLINE __device__ __noinline__
#else
#define FMG_NOINLINE
#endif

namespace fmg {

struct OverlapArgs {
    OccView ix;
    int min_match;
    int mode;                   // 0: unitig records; 1: fm6_retrieve for seqsort (seqsort.c:12-35): phase 1 only, no candidate lists, no length cut
    int64_t n;                  // sequences in this batch
    // phase 1 spells the sequences itself (fm_retrieve, exact.c:59-70, fused with the backward search of fm6_is_contained)
    const uint64_t *ids;        // BWT rows (sentinel ranks) of the batch, or nullptr for row t = first + t * step
    uint64_t first, step;
    uint8_t *seq;               // n x max_len nt6 bytes, reading order
    int32_t *len;               // n; a sequence longer than max_len is flagged by len = -(true length)
    int64_t *ret;               // n: the value fm_retrieve returns = rank of the sequence among all sequences
    int max_len;
    // per-sequence scratch handed from phase to phase
    void *P0; int pcap;         // candidate list of overlap_intv (unitig.c:38-64): n x pcap entries of 4 x U
    int32_t *np0;               // n: entries in P0 for the next phase; -1 = the next phase has nothing to do
    // per-lane scratch of the two list-chasing phases
    void *A, *B; int cap;       // candidate interval lists (4 x U per entry)
    int32_t *cat;               // category per candidate (unitig.c:105-151): 2 x cap per lane (previous / current level)
    // per-sequence output
    int64_t *rec;               // 10 per sequence, see OV_* below
    uint4 *nei; int nei_cap;    // neighbour records (fmintv_t: x = interval of the neighbour, info = overlap length)
    uint32_t *nei_cnt;
    uint8_t *ext;               // n x max_len: bases fm6_get_nei appends to the read (s[len .. s_len), unitig.c:141)
    unsigned long long *next;
};

// rec[] layout (the first nine match oracle/ref_harness.c:refh_overlap_batch)
enum { OV_K = 0, OV_LEN, OV_CONTAINED, OV_X0, OV_X1, OV_X2, OV_RBEG, OV_NNEI, OV_SLEN, OV_LEFT, OV_NREC };
// OV_CONTAINED: 0, -1 contained (unitig.c:86,89), -9 not longer than min_match (unitig.c:288), -100 scratch overflow
// OV_LEFT: check_left_simple of the unique neighbour: 0 ok, -1 backward bifurcation, 1 not evaluated

// ---------------------------------------------------------------------------------------------
// Packed record the unitig walk chases: one 64-byte line per sequence, indexed by the RANK of the sequence
// (the id the walk sees: fm_retrieve's return value, intv0.x[], neighbour x[]), so that one step of
// unitig_unidir (unitig.c:227-262) costs one cache miss for the record and one for the appended bases.
// The overlap length of a unique neighbour is len - rbeg (unitig.c:118,155), so it is not stored.
struct alignas(64) OvPack {
    uint64_t x0, x1;          // intv0.x[0], intv0.x[1]                              (OV_X0, OV_X1)
    uint64_t nx0, nx1;        // nnei == 1: x[0], x[1] of the neighbour;  nnei > 1: nx0 = first entry in the spill array
    uint64_t ext_first;       // nnei == 1: offset of the slen - len appended bases in the ext array
    uint32_t x2, nx2;         // intv0.x[2]; x[2] of the unique neighbour
    uint32_t len, slen;       // OV_LEN, OV_SLEN
    int32_t  rbeg;            // OV_RBEG (-1: no neighbour)
    uint16_t nnei;            // OV_NNEI
    int8_t   contained, left; // OV_CONTAINED (0, -1, -9), OV_LEFT (0, -1, 1)
};
static_assert(sizeof(OvPack) == 64, "OvPack must be one 64-byte line");

// number of appended bases the walk will read for this record (only a unique neighbour extends the consensus)
FMG_HD uint32_t ov_ext_len(const int64_t *rec) {
    return (rec[OV_NNEI] == 1 && rec[OV_SLEN] > rec[OV_LEN]) ? (uint32_t)(rec[OV_SLEN] - rec[OV_LEN]) : 0u;
}

// rec[] (OV_*) + first neighbour / spill index + ext offset -> OvPack; returns false when a field does not fit
FMG_HD bool ov_pack(const int64_t *rec, uint64_t nx0, uint64_t nx1, uint64_t nx2, uint64_t ext_first, OvPack *o) {
    o->x0 = (uint64_t)rec[OV_X0]; o->x1 = (uint64_t)rec[OV_X1];
    o->nx0 = nx0; o->nx1 = nx1; o->ext_first = ext_first;
    o->x2 = (uint32_t)rec[OV_X2]; o->nx2 = (uint32_t)nx2;
    o->len = (uint32_t)rec[OV_LEN]; o->slen = (uint32_t)rec[OV_SLEN];
    o->rbeg = (int32_t)rec[OV_RBEG];
    o->nnei = (uint16_t)rec[OV_NNEI];
    o->contained = (int8_t)rec[OV_CONTAINED]; o->left = (int8_t)rec[OV_LEFT];
    return (uint64_t)rec[OV_X2] < (1ull << 32) && nx2 < (1ull << 32) && rec[OV_NNEI] < 65536 && rec[OV_LEN] < (1ll << 31) && rec[OV_SLEN] < (1ll << 31);
}

// Result intervals of fm6_extend (exact.c:72-88) indexed by symbol, info = 0: storage for the ok[1..4] a call site walks over.
template <typename U> struct Ok6 { IntvT<U> v[6]; };

// What most call sites need of an extension: the size of ok[0] and the interval of ONE selected symbol.  These come back in
// registers together with a mask of the non-empty base extensions; the intervals of those (ok[1..4], usually one) are only
// written -- through `kids`, to the caller's stack -- where a call site walks over them.  Everything is computed inside
// ext_sync, i.e. while the warp is converged, so that the (divergent) callers only pick values.
template <typename U> struct ExtSel { U s0, ssel, x0, x1; int flags; };       // flags: bit 0 = some lane of the warp was active, bits 1..4 = ok[c] is not empty

// TAG gives every kernel its own copy of the function (ptxas 12.9 crashes on a noinline function shared by two entries)

template <typename U, int TAG>
FMG_NOINLINE ExtSel<U> ext_sync(const OccView &ix, bool active, U x0, U x1, U x2, int back, int csel, Ok6<U> *kids) {
    ExtSel<U> R;
    R.s0 = R.ssel = R.x0 = R.x1 = 0;
#if defined(__CUDA_ARCH__)
    R.flags = __any_sync(0xffffffffu, active) ? 1 : 0;
#else
    R.flags = active ? 1 : 0;
#endif
    if (active) {
        Ext6T<U> e;
        extend6<U>(ix, back ? x1 : x0, back ? x0 : x1, x2, e);
        const uint64_t *row = ix.cs + e.sbk * 8;
        const U nr = pick6(e.near, csel), fr = (U)(ld_u64(row + csel) + pick6(e.relk, csel));
        R.s0 = e.size[0]; R.ssel = pick6(e.size, csel);
        R.x0 = back ? fr : nr; R.x1 = back ? nr : fr;
#pragma unroll
        for (int c = 1; c < 5; ++c) {
            if (e.size[c] == 0) continue;
            R.flags |= 1 << c;
            if (kids) {
                const U f6 = (U)(ld_u64(row + c) + e.relk[c]);
                kids->v[c].x0 = back ? f6 : e.near[c];
                kids->v[c].x1 = back ? e.near[c] : f6;
                kids->v[c].x2 = e.size[c];
                kids->v[c].info = 0;
            }
        }
    }
    return R;
}

template <typename U> struct OvBits {     // packing of the candidate `info` word (unitig.c:132,150)
    static constexpr int pos_bits = sizeof(U) == 8 ? 32 : 16;
    static constexpr U pos_mask = (U)(((uint64_t)1 << pos_bits) - 1);
    static constexpr U base_mask = (U)((uint64_t)0xf << pos_bits);
    static constexpr int cat_shift = pos_bits + 4;
    static constexpr int max_cat = sizeof(U) == 8 ? (1 << 27) : ((1 << 12) - 1);
};

// the consensus string of fm6_get_nei: the read itself followed by the appended bases
struct SeqView {
    const uint8_t *a; int la; const uint8_t *b;
    FMG_HD int at(int i) const { return i < la ? a[i] : b[i - la]; }
};

// The per-sequence record is computed in FOUR phases, each its own kernel over the batch, handing the
// candidate list of overlap_intv from one to the next through P0 / np0:
//   1 phase_contained  fm6_is_contained (unitig.c:77-91): a chain of len+1 extensions, no lists read    [SYNC = false]
//   2 phase_nei        fm6_get_nei (unitig.c:93-179): breadth-first over the candidate list             [SYNC = true]
//   3 phase_left1      overlap_intv of check_left_simple (unitig.c:186-190): a chain of extensions      [SYNC = false]
//   4 phase_left2      the candidate loop of check_left_simple (unitig.c:191-203)                       [SYNC = true]
// Phases 1 and 3 are the same straight loop for every lane of a warp (equal-length reads: exactly the same trip
// count), so they run converged without help and with few registers.  Phases 2 and 4 have data-dependent nested
// loops; there every extension goes through ext_sync (warp vote + one shared copy of the code).  Splitting them
// keeps the lanes of a warp in the same phase: one fused kernel had lanes in all four at once, and every distinct
// path between two votes costs the warp its own serialized round trips to the scratch lists.
template <typename U, bool SYNC, int TAG>
struct OvLane {
    typedef IntvT<U> Cand;
    typedef OvBits<U> BT;
    const OverlapArgs &A;
    Cand *P, *Q;      // lane lists (phases 2 and 4)
    int32_t *cat;
    bool ovf;
    Ok6<U> em;        // SYNC: the non-empty ok[1..4] of the last extension that asked for them (phase 2)
    ExtSel<U> rs;     // SYNC: size of ok[0], the selected interval and the mask of non-empty ok[1..4] of the last extension (registers)
    Ext6T<U> e;       // !SYNC: the same, before the far coordinates are formed
    int eback;

    FMG_HD OvLane(const OverlapArgs &a, int64_t lane)
        : A(a), P(static_cast<Cand *>(a.A) + (size_t)lane * a.cap), Q(static_cast<Cand *>(a.B) + (size_t)lane * a.cap),
          cat(a.cat + (size_t)lane * a.cap * 2), ovf(false), eback(0) {}

    // SYNC phases: rs for the selected symbol, with (ext_kids) or without (extend_sel) the non-empty ok[1..4] in `em`
    FMG_HD void ext_kids(const Cand &k, int back, int csel) { rs = ext_sync<U, TAG>(A.ix, true, k.x0, k.x1, k.x2, back, csel, &em); }
    FMG_HD void extend_sel(const Cand &k, int back, int csel) { rs = ext_sync<U, TAG>(A.ix, true, k.x0, k.x1, k.x2, back, csel, nullptr); }
    FMG_HD Cand sel() const { Cand o; o.x0 = rs.x0; o.x1 = rs.x1; o.x2 = rs.ssel; o.info = 0; return o; }
    // chain phases (!SYNC): the plain inlined extension
    FMG_HD void extend(const Cand &k, int back) { extend6<U>(A.ix, back ? k.x1 : k.x0, back ? k.x0 : k.x1, k.x2, e); eback = back; }
    FMG_HD U size(int c) const { return pick6(e.size, c); }
    FMG_HD Cand ok(int c) const {
        Cand o;
        const U nr = pick6(e.near, c), fr = far_of(A.ix, e, c);
        o.x0 = eback ? fr : nr; o.x1 = eback ? nr : fr; o.x2 = pick6(e.size, c); o.info = 0;
        return o;
    }
    FMG_HD void push(Cand *list, int lcap, int &n, const Cand &k) {
        if (n < lcap) st_cand(list + n, k); else ovf = true;
        ++n;
    }
    static FMG_HD void reverse(Cand *list, int n, int lcap) {
        if (n > lcap) n = lcap;
        for (int a = 0, b = n - 1; a < b; ++a, --b) {
            const Cand x = ld_cand(list + a), y = ld_cand(list + b);
            st_cand(list + a, y); st_cand(list + b, x);
        }
    }
    FMG_HD Cand *list0(int64_t t) const { return static_cast<Cand *>(A.P0) + (size_t)t * A.pcap; }

    // overlap_intv, unitig.c:38-64.  Returns the final interval; list = candidates, smallest interval first.
    FMG_HD Cand overlap_intv(int len, const SeqView &sv, int min, int j, int at5, Cand *list, int lcap, int &n, int inc_sentinel) {
        const int dir = at5 ? 1 : -1, end = at5 ? len : -1;
        Cand ik = base_intv<U>(A.ix, sv.at(j));
        n = 0;
        int depth = 1;
        for (j += dir; j != end; j += dir, ++depth) {
            const int b = sv.at(j);
            const int c = at5 ? comp6(b) : b;
            extend(ik, !at5);
            if (size(c) == 0) break;
            if (depth >= min && size(0) != 0) {
                if (inc_sentinel) { Cand t = ok(0); t.info = (U)(j - dir); push(list, lcap, n, t); }
                else { ik.info = (U)(j - dir); push(list, lcap, n, ik); }
            }
            ik = ok(c);
        }
        reverse(list, n, lcap);
        return ik;
    }

    // ---- phase 1: fm_retrieve (exact.c:59-70) + fm6_is_contained (unitig.c:77-91) in ONE chain.
    // fm_retrieve spells the sequence from its last base to its first by LF steps from BWT row k; fm6_is_contained
    // extends the bi-interval of the growing suffix backward by exactly those bases (overlap_intv, unitig.c:38-64,
    // at5 = 0).  Row k always lies inside that interval [x0, x0+x2), so for the small intervals of most steps the block
    // that yields BWT[k] and LF(k) is the block the extension reads anyway: one 64-byte access per base for both.
    // The read length is only known at the end, so candidates are pushed with the suffix length in `info` and get the
    // start position (unitig.c:55) when the list is reversed.
    FMG_HD void phase_contained(int64_t t) {
        int64_t *rec = A.rec + t * OV_NREC;
        const int min_match = A.min_match;
        uint8_t *out = A.mode == 0 ? A.seq + (size_t)t * A.max_len : nullptr;
        Cand *list = A.mode == 0 ? list0(t) : nullptr;
        ovf = false;
        uint64_t k = A.ids ? A.ids[t] : A.first + (uint64_t)t * A.step;
        int n = 0, np = 0, ret = 0;
        Cand ik = {0, 0, 0, 0}, intv0 = {0, 0, 0, 0};
        for (;;) {
            // blocks: the two ends of the interval (from the second base on) and the one holding row k
            const uint64_t pk = ik.x0, pl = (uint64_t)ik.x0 + ik.x2;
            Blk bk, bl, bq;
            if (n > 0) {
                bk = load_blk(A.ix, pk);
                bl = bk;
                if ((pk >> kBlkShift) != (pl >> kBlkShift)) bl = load_blk(A.ix, pl);
                const uint64_t qb = k >> kBlkShift;
                if (qb == (pk >> kBlkShift)) bq = bk;
                else if (qb == (pl >> kBlkShift)) bq = bl;
                else bq = load_blk(A.ix, k);
            } else bq = load_blk(A.ix, k);
            uint32_t rel[6];
            rank_rel(bq, k, rel);                             // counts in [superblock_start, k)
            const int c = blk_symbol(bq, k);                  // BWT[k]
            // rank of c in BWT[0..k] is rel+1, so LF(k) = C[c] + rel (exact.c:66)
            k = ld_u64(A.ix.cs + (k >> kSuperShift) * 8 + c) + pick6(rel, c);
            const bool last = c == 0 || c > 5;
            if (n > 0) {
                extend6_with<U>(A.ix, ik.x1, pk, pl, bk, bl, e);      // backward extension of the suffix read so far
                eback = 1;
                if (last) {                                   // fm6_is_contained after overlap_intv: extend(ik, 1)
                    if (ik.x2 != size(0)) ret = -1;           // left contained
                    intv0 = ok(0);
                    break;
                }
                if (A.mode == 0 && n >= min_match && size(0) != 0) { ik.info = (U)n; push(list, A.pcap, np, ik); }
                ik = ok(c);
            } else {
                if (last) break;                              // an empty sequence
                ik = base_intv<U>(A.ix, c);
            }
            if (A.mode == 0 && n < A.max_len) out[n] = (uint8_t)c;
            ++n;
        }
        A.ret[t] = (int64_t)k;
        for (int q = 0; q < OV_NREC; ++q) rec[q] = 0;
        if (A.mode != 0) {                                            // fm6_retrieve (exact.c:100-127): k, k2 and the containment flags
            extend(intv0, 0);
            if (intv0.x2 != size(0)) ret = -1;
            intv0 = ok(0);
            rec[OV_LEN] = n; rec[OV_CONTAINED] = ret; rec[OV_X0] = (int64_t)intv0.x0; rec[OV_X1] = (int64_t)intv0.x1; rec[OV_X2] = (int64_t)intv0.x2;
            return;
        }
        const int L = n <= A.max_len ? n : -n;
        A.len[t] = L;
        for (int a = 0, b = (n < A.max_len ? n : A.max_len) - 1; a < b; ++a, --b) { const uint8_t x = out[a]; out[a] = out[b]; out[b] = x; }   // seq_reverse (unitig.c:285)
        rec[OV_LEN] = L; rec[OV_RBEG] = -1; rec[OV_LEFT] = 1;
        A.nei_cnt[t] = 0;
        A.np0[t] = -1;
        if (L <= min_match) { rec[OV_CONTAINED] = -9; return; }       // unitig.c:288 (also: clipped sequences, re-run by the host)
        // candidates: smallest interval first, info = start of the suffix in the read
        {
            const int m = np < A.pcap ? np : A.pcap;
            for (int a = 0, b = m - 1; a <= b; ++a, --b) {
                Cand x = ld_cand(list + a), y = ld_cand(list + b);
                x.info = (U)(L - (int)x.info); y.info = (U)(L - (int)y.info);
                st_cand(list + a, y);
                if (a != b) st_cand(list + b, x);
            }
        }
        extend(intv0, 0);
        if (intv0.x2 != size(0)) ret = -1;              // right contained
        intv0 = ok(0);
        rec[OV_CONTAINED] = ret; rec[OV_X0] = (int64_t)intv0.x0; rec[OV_X1] = (int64_t)intv0.x1; rec[OV_X2] = (int64_t)intv0.x2;
        if (ovf) { rec[OV_CONTAINED] = -100; return; }
        if (ret < 0 || np == 0) return;
        A.np0[t] = np;
    }

    // ---- phase 2: fm6_get_nei, unitig.c:93-179 (beg = 0; the first `prev` list was filled by phase 1)
    // List entries keep the RAW sort key of unitig.c:132 in `info` (category of the parent | base | position); the category
    // an entry gets at the end of its level (unitig.c:142-151) lives in a parallel array and is merged in when the entry
    // is read.  Children arrive almost always in key order already (one category, one base), so the category is computed
    // while pushing and the sort + renumbering pass of the reference only runs for a level whose pushes were out of order.
    FMG_HD void phase_nei(int64_t t) {
        int np = A.np0[t];
        if (np <= 0) return;
        A.np0[t] = -1;
        int64_t *rec = A.rec + t * OV_NREC;
        const int ori_l = A.len[t];
        const uint8_t *sq = A.seq + (size_t)t * A.max_len;
        uint8_t *xt = A.ext + (size_t)t * A.max_len;
        int sl = ori_l, nq = 0;
        ovf = false;
        int nnei = 0, is_forked = 0;
        uint4 *nei = A.nei + (size_t)t * A.nei_cap * 2;
        Cand nei0 = {0, 0, 0, 0};
        Cand *prev = list0(t), *curr = P;
        int32_t *pcat = cat, *ccat = cat + A.cap;                        // categories of `prev` / `curr`
        int pcap = A.pcap;                                               // capacity of `prev`; `curr` always holds A.cap
        for (int j = 0; j < np && j < A.cap; ++j) pcat[j] = 0;
        while (np) {
            nq = 0;
            int first_base = 0, cat_run = 0;
            bool sorted = true;
            U last_key = 0, last_hi = 0;
            const int npc = np < pcap ? np : pcap;
            Cand p = ld_cand(prev);
            int cj = pcat[0];
            for (int j = 0; j < npc; ++j) {
                // the next candidate is requested before this one is extended: its latency hides behind the extension
                const bool more = j + 1 < npc;
                Cand pn = p;
                int cn = -1;
                if (more) { pn = ld_cand(prev + j + 1); cn = pcat[j + 1]; }
                do {
                    if (cj < 0) break;
                    p.info = (U)((p.info & BT::pos_mask) | ((U)cj << BT::cat_shift));
                    ext_kids(p, 0, 0);                                   // forward extension; the non-empty ok[1..4] stay in `em` across the probes
                    const U s0 = rs.s0;
                    const int kids = rs.flags;
                    if (s0 != 0 && ori_l != sl) {                        // some (partial) reads end here
                        extend_sel(sel(), 1, 0);                         // fm6_extend0(ok[0], back)
                        if (rs.s0 != 0) {                                // bounded by sentinels on both sides: a full read
                            if (s0 == p.x2 && p.x2 == rs.s0) {           // not contained in a longer read
                                Cand nb = sel();                         // fm6_extend0's ok0 (x[0] = tk[0]; cnt[0] is 0)
                                nb.info = (U)(ori_l - (int)(p.info & BT::pos_mask));
                                for (int i = j; i < npc && pcat[i] == cj; ++i) pcat[i] = -1;
                                if (more) cn = pcat[j + 1];              // the request above may predate the masking
                                if (nnei < A.nei_cap) {
                                    Intv o; o.x0 = nb.x0; o.x1 = nb.x1; o.x2 = nb.x2; o.info = nb.info;
                                    st_intv(nei + 2 * nnei, o);
                                } else ovf = true;
                                if (nnei == 0) nei0 = nb;
                                ++nnei;
                                break;
                            }   // else: a read contained in another one (the reference only marks it `used`)
                        }
                    }
                    for (int c = 1; c < 5; ++c) {                        // collect extensible intervals
                        if (!(kids >> c & 1)) continue;
                        Cand kc = em.v[c];
                        extend_sel(kc, 1, 0);                            // fm6_extend0(ok[c], back)
                        if (rs.s0 != 0) {                                // left end still bounded by a sentinel
                            kc.info = (U)((p.info & ~BT::base_mask) | ((U)c << BT::pos_bits));
                            const U hi = (U)(kc.info >> BT::pos_bits);
                            if (nq == 0) first_base = c;
                            else if (kc.info < last_key) sorted = false;
                            if (nq == 0 || hi != last_hi) { last_hi = hi; cat_run = nq; }
                            last_key = kc.info;
                            if (nq < A.cap) ccat[nq] = cat_run;
                            push(curr, A.cap, nq, kc);
                        }
                    }
                } while (0);
                p = pn; cj = cn;
            }
            if (nq) {                                                    // update categories, unitig.c:137-153
                const int nqc = nq < A.cap ? nq : A.cap;
                if (sl - ori_l < A.max_len) { xt[sl - ori_l] = (uint8_t)comp6(first_base); ++sl; } else ovf = true;
                if (!sorted) {
                    for (int a = 1; a < nqc; ++a) {                      // insertion sort by info (keys are unique)
                        const Cand x = ld_cand(curr + a);
                        int b = a - 1;
                        while (b >= 0) {
                            const Cand y = ld_cand(curr + b);
                            if (y.info <= x.info) break;
                            st_cand(curr + b + 1, y); --b;
                        }
                        st_cand(curr + b + 1, x);
                    }
                    U last = 0;
                    cat_run = 0;
                    for (int j = 0; j < nqc; ++j) {
                        const U hi = (U)(ld_cand(curr + j).info >> BT::pos_bits);
                        if (j == 0 || hi != last) { last = hi; cat_run = j; }
                        ccat[j] = cat_run;
                    }
                }
                if (cat_run != 0) is_forked = 1;
                if (cat_run > BT::max_cat) ovf = true;
            }
            prev = curr; curr = curr == P ? Q : P;
            int32_t *tc = pcat; pcat = ccat; ccat = tc;
            pcap = A.cap;
            np = nq;
        }
        A.nei_cnt[t] = (uint32_t)nnei;
        rec[OV_NNEI] = nnei;
        if (nnei == 0) { rec[OV_SLEN] = sl; if (ovf) rec[OV_CONTAINED] = -100; return; }     // unitig.c:154 (returns -1, s keeps its growth)
        const int rbeg = ori_l - (int)nei0.info;
        if (nnei == 1 && is_forked) {             // contained reads forked the path: rebuild it along the one neighbour
            Cand k0 = base_intv<U>(A.ix, 0);
            for (int i = rbeg; i < ori_l; ++i) { extend_sel(k0, 0, comp6(sq[i])); k0 = sel(); }
            int i = ori_l;
            for (; i < sl; ++i) {
                int c0 = -1, hits = 0;
                ext_kids(k0, 0, 0);
                for (int c = 1; c < 5; ++c) {
                    if (!(rs.flags >> c & 1)) continue;
                    const Cand kc = em.v[c];
                    if (kc.x0 <= nei0.x0 && kc.x0 + kc.x2 >= nei0.x0 + nei0.x2) ++hits, c0 = c;
                }
                if (hits == 0 && rs.s0 != 0) break;
                if (hits != 1) { ovf = true; break; }                    // the reference asserts hits == 1 (unitig.c:171)
                xt[i - ori_l] = (uint8_t)comp6(c0);
                k0 = em.v[c0];
            }
            sl = i;
        }
        if (nnei > 1) sl = ori_l;
        rec[OV_RBEG] = rbeg; rec[OV_SLEN] = sl;
        if (ovf) rec[OV_CONTAINED] = -100;
    }

    // ---- phase 3: the overlap_intv call of check_left_simple, unitig.c:186-190, for a unique neighbour (beg = 0)
    FMG_HD void phase_left1(int64_t t) {
        int64_t *rec = A.rec + t * OV_NREC;
        if (rec[OV_NNEI] != 1 || rec[OV_CONTAINED] != 0) return;
        const int L = A.len[t], sl = (int)rec[OV_SLEN], rbeg = (int)rec[OV_RBEG];
        const SeqView sv = {A.seq + (size_t)t * A.max_len, L, A.ext + (size_t)t * A.max_len};
        ovf = false;
        int np = 0;
        overlap_intv(sl, sv, A.min_match, rbeg, 1, list0(t), A.pcap, np, 1);
        if (ovf) { rec[OV_CONTAINED] = -100; return; }
        A.np0[t] = np;
    }

    // ---- phase 4: the candidate loop of check_left_simple, unitig.c:191-203
    FMG_HD void phase_left2(int64_t t) {
        int np = A.np0[t];
        if (np < 0) return;
        int64_t *rec = A.rec + t * OV_NREC;
        const uint8_t *sq = A.seq + (size_t)t * A.max_len;
        const int rbeg = (int)rec[OV_RBEG];
        ovf = false;
        int left = 0, pcap = A.pcap;
        Cand *prev = list0(t), *curr = P;
        for (int i = rbeg - 1; i >= 0 && left == 0; --i) {
            int nq = 0;
            const int npc = np < pcap ? np : pcap, c = sq[i];
            for (int j = 0; j < npc; ++j) {
                const Cand p = ld_cand(prev + j);
                extend_sel(p, 1, c);
                if ((U)(rs.s0 + rs.ssel) != p.x2) { left = -1; break; }         // potential backward bifurcation
                push(curr, A.cap, nq, sel());
            }
            prev = curr; curr = curr == P ? Q : P;
            pcap = A.cap;
            np = nq;
        }
        rec[OV_LEFT] = left;
        if (ovf) rec[OV_CONTAINED] = -100;
    }
};

// list-chasing phases: persistent lanes; a lane that runs out of work keeps answering the warp votes
template <typename U, int PHASE, class FetchFn>
FMG_HD void overlap_lane_sync(const OverlapArgs &A, int64_t lane, FetchFn fetch) {
    OvLane<U, true, PHASE> ln(A, lane);
    for (;;) {
        const int64_t t = fetch();
        if (t >= A.n) break;
        if (PHASE == 2) ln.phase_nei(t); else ln.phase_left2(t);
    }
    while (ext_sync<U, PHASE>(A.ix, false, 0, 0, 0, 0, 0, nullptr).flags & 1) {}
}

// chain phases: one sequence per thread
template <typename U, int PHASE>
FMG_HD void overlap_chain(const OverlapArgs &A, int64_t t) {
    OvLane<U, false, PHASE> ln(A, 0);
    if (PHASE == 1) ln.phase_contained(t); else ln.phase_left1(t);
}

} // namespace fmg
