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
    uint8_t *seq;               // n rows of max_len nt6 bytes; a sequence is RIGHT-aligned in its row (phase 1 spells it last base
                                // first and fills the row from its end, eight bases per store): see seq_of()
    int32_t *len;               // n; a sequence longer than max_len is flagged by len = -(true length)
    int64_t *ret;               // n: the value fm_retrieve returns = rank of the sequence among all sequences
    int max_len;                // row stride of seq / ext: a multiple of 8
    // per-sequence scratch handed from phase to phase
    void *P0; int pcap;         // candidate list of overlap_intv (unitig.c:38-64): n x pcap entries of 4 x U.  Phase 1 fills a slot from
                                // its END (the list wanted is the push order reversed) with info = suffix length; phase 3 from its start
    void *S0;                   // phase 1 -> 2: n x pcap values of U, |$P| of every candidate (see nei_lane)
    int32_t *np0;               // n: entries in P0 for the next phase; -1 = the next phase has nothing to do
    // per-lane scratch of the two list-chasing phases
    void *A, *B; int cap;       // candidate interval lists (4 x U per entry)
    int32_t *cat;               // category per candidate (unitig.c:105-151): 2 x cap per lane (previous / current level)
    void *S;                    // |$P| per candidate of the two level lists of phase 2: 2 x cap values of U per lane
    // per-sequence output
    int64_t *rec;               // 10 per sequence, see OV_* below
    uint4 *nei; int nei_cap;    // neighbour records (fmintv_t: x = interval of the neighbour, info = overlap length)
    uint32_t *nei_cnt;
    uint8_t *ext;               // n x max_len: bases fm6_get_nei appends to the read (s[len .. s_len), unitig.c:141)
    unsigned long long *next;
    const uint32_t *order;      // phase 2: the sequences of the batch in the order the lanes take them (see k_nei_key), or nullptr
    // row t of seq, given the (clipped) length of the sequence
    FMG_HD const uint8_t *seq_of(int64_t t, int len) const { return seq + (size_t)(t + 1) * max_len - (len < max_len ? len : max_len); }
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

// What the call sites of the list phase need of an extension (fm6_extend, exact.c:72-88): the size of ok[0] and the interval of ONE
// selected symbol.  These come back in registers together with a mask of the non-empty base extensions.  Everything is computed
// inside ext_sync, i.e. while the warp is converged, so that the (divergent) callers only pick values.
template <typename U> struct ExtSel { U s0, ssel, x0, x1; int flags; };       // flags: bit 0 = some lane of the warp was active, bits 1..4 = ok[c] is not empty

// TAG gives every kernel its own copy of the function (ptxas 12.9 crashes on a noinline function shared by two entries)

template <typename U, int TAG>
FMG_NOINLINE ExtSel<U> ext_sync(const OccView &ix, bool active, U x0, U x1, U x2, int back, int csel) {
    ExtSel<U> R;
    R.s0 = R.ssel = R.x0 = R.x1 = 0;
#if defined(__CUDA_ARCH__)
    R.flags = __any_sync(0xffffffffu, active) ? 1 : 0;
#else
    R.flags = active ? 1 : 0;
#endif
    if (!R.flags) return R;
    Ext6T<U> e;
    extend6_conv<U>(ix, active, back ? x1 : x0, back ? x0 : x1, x2, e);       // all 32 lanes are here: paired block loads
    if (active) {
        const uint64_t *row = ix.cs + e.sbk * 8;
        const U nr = pick6(e.near, csel), fr = (U)(ld_u64(row + csel) + pick6(e.relk, csel));
        R.s0 = e.size[0]; R.ssel = pick6(e.size, csel);
        R.x0 = back ? fr : nr; R.x1 = back ? nr : fr;
#pragma unroll
        for (int c = 1; c < 5; ++c) if (e.size[c] != 0) R.flags |= 1 << c;
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
//   2 nei_lane         fm6_get_nei (unitig.c:93-179): breadth-first over the candidate list, one gather per level (below)
//   3 phase_left1      overlap_intv of check_left_simple (unitig.c:186-190): a chain of extensions      [SYNC = false]
//   4 phase_left2      the candidate loop of check_left_simple (unitig.c:191-203)                       [SYNC = true]
// Phases 1 and 3 are the same straight loop for every lane of a warp (equal-length reads: exactly the same trip
// count), so they run converged without help and with few registers.  Phases 2 and 4 have data-dependent nested
// loops: phase 2 is a flat state machine of its own (nei_lane), in phase 4 every extension goes through ext_sync (warp vote +
// one shared copy of the code).  Splitting them
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
    ExtSel<U> rs;     // SYNC: size of ok[0], the selected interval and the mask of non-empty ok[1..4] of the last extension (registers)
    Ext6T<U> e;       // !SYNC: the same, before the far coordinates are formed
    int eback;

    FMG_HD OvLane(const OverlapArgs &a, int64_t lane)
        : A(a), P(static_cast<Cand *>(a.A) + (size_t)lane * a.cap), Q(static_cast<Cand *>(a.B) + (size_t)lane * a.cap),
          cat(a.cat + (size_t)lane * a.cap * 2), ovf(false), eback(0) {}

    // SYNC phases: rs for the selected symbol
    FMG_HD void extend_sel(const Cand &k, int back, int csel) { rs = ext_sync<U, TAG>(A.ix, true, k.x0, k.x1, k.x2, back, csel); }
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

    // ---- phase 1: fm_retrieve (exact.c:59-70) + fm6_is_contained (unitig.c:77-91) in ONE chain, run by a CONVERGED warp
    // (every lane its own sequence; `live` = the lane has one; finished lanes keep answering the votes and the paired loads).
    // fm_retrieve spells the sequence from its last base to its first by LF steps from BWT row k; fm6_is_contained
    // extends the bi-interval of the growing suffix backward by exactly those bases (overlap_intv, unitig.c:38-64,
    // at5 = 0).  Row k always lies inside that interval [x0, x0+x2), so for the small intervals of most steps the block
    // that yields BWT[k] and LF(k) is the block the extension reads anyway: one 64-byte gather per base for both, fetched
    // through load_blk_pair (one L1-miss request per block instead of two).  Stores are kept off the request path too: the
    // bases go out eight per store into a row filled from its END (no seq_reverse pass, unitig.c:285), and so do the
    // candidates (the list wanted is the push order reversed, unitig.c:62), with the suffix length in `info` because the
    // read length -- hence the start position of unitig.c:55 -- is only known at the end.
    FMG_HD void phase_contained(int64_t t, bool live) {
        const int min_match = A.min_match, ml = A.max_len;
        uint8_t *row_end = A.mode == 0 && live ? A.seq + (size_t)(t + 1) * ml : nullptr;
        Cand *list = A.mode == 0 && live ? list0(t) : nullptr;
        ovf = false;
        uint64_t k = live ? (A.ids ? A.ids[t] : A.first + (uint64_t)t * A.step) : 0;
        int n = 0, np = 0, ret = 0;
        int stage = live ? 0 : 2;                         // 0: the chain; 1: the forward extension of intv0 (unitig.c:88-90); 2: done
        bool empty = false;
        uint64_t acc = 0;                                 // the last bases spelled, newest in the low byte
        Cand ik = {0, 0, 0, 0}, intv0 = {0, 0, 0, 0};
        while (warp_any(stage < 2)) {
            const bool run = stage < 2, chain = stage == 0;
            const bool have = run && (n > 0 || !chain);   // there is an interval to extend
            // backward extension of ik reads the blocks of x0 and x0 + x2; the forward extension of intv0 those of x1 and x1 + x2
            const uint64_t pk = chain ? ik.x0 : intv0.x1, pl = pk + (chain ? ik.x2 : intv0.x2);
            const uint64_t bpk = pk >> kBlkShift, bpl = pl >> kBlkShift, bq = k >> kBlkShift;
            uint32_t rk[6], rl[6], rq[6];
            int c = 0;
            // Up to three blocks per step (both ends of the interval and row k): rare for a lane after the first ~15 bases, yet most
            // steps for SOME lane of the warp.  The gathers are nevertheless taken one after the other: the phase is bound by the
            // rate of L1-miss requests, not by latency (enough warps are resident), and one block in flight keeps the registers
            // at 72 (28 warps per SM) where three in flight need 113 (16 warps).
            bool need_l = false, need_q = false;
            {
                const Blk B = load_blk_pair(A.ix, have ? pk : k, run);
                if (have) {
                    rank_rel(B, pk, rk);
                    need_l = bpl != bpk;
                    if (!need_l) rank_rel(B, pl, rl);
                    need_q = chain && bq != bpk;
                    if (chain && !need_q) { rank_rel(B, k, rq); c = blk_symbol(B, k); }
                } else if (run) { rank_rel(B, k, rq); c = blk_symbol(B, k); }
            }
            if (warp_any(need_l)) {                       // large intervals (the first ~15 bases): the other end is in another block
                const Blk B = load_blk_pair(A.ix, pl, need_l);
                if (need_l) {
                    rank_rel(B, pl, rl);
                    if (need_q && bq == bpl) { rank_rel(B, k, rq); c = blk_symbol(B, k); need_q = false; }
                }
            }
            if (warp_any(need_q)) {
                const Blk B = load_blk_pair(A.ix, k, need_q);
                if (need_q) { rank_rel(B, k, rq); c = blk_symbol(B, k); }
            }
            if (!run) continue;
            if (!chain) {                                 // fm6_is_contained, second extension: right containment
                extend6_rel<U>(A.ix, intv0.x0, pk, pl, rk, rl, e);
                eback = 0;
                if (intv0.x2 != size(0)) ret = -1;
                intv0 = ok(0);
                stage = 2;
                continue;
            }
            // rank of c in BWT[0..k] is rel+1, so LF(k) = C[c] + rel (exact.c:66)
            k = ld_u64(A.ix.cs + (k >> kSuperShift) * 8 + c) + pick6(rq, c);
            const bool last = c == 0 || c > 5;
            if (n > 0) {
                extend6_rel<U>(A.ix, ik.x1, pk, pl, rk, rl, e);       // backward extension of the suffix read so far
                eback = 1;
                if (last) {                                   // fm6_is_contained after overlap_intv: extend(ik, 1)
                    if (ik.x2 != size(0)) ret = -1;           // left contained
                    intv0 = ok(0);
                    stage = 1;
                    continue;
                }
                if (A.mode == 0 && n >= min_match && size(0) != 0) {      // some sequences start with this suffix P: a candidate overlap
                    // phase 2 wants the suffix P as (x0 of $P, x1, x2, suffix length) and |$P| (see nei_lane)
                    Cand cd; cd.x0 = far_of(A.ix, e, 0); cd.x1 = ik.x1; cd.x2 = ik.x2; cd.info = (U)n;
                    if (np < A.pcap) {
                        st_cand(list + (A.pcap - 1 - np), cd);
                        static_cast<U *>(A.S0)[(size_t)t * A.pcap + (A.pcap - 1 - np)] = size(0);
                    } else ovf = true;
                    ++np;
                }
                ik = ok(c);
            } else {
                if (last) { empty = true; stage = 2; continue; }      // an empty sequence
                ik = base_intv<U>(A.ix, c);
            }
            acc = acc << 8 | (uint64_t)c;
            ++n;
            if (A.mode == 0 && (n & 7) == 0 && n <= ml) *reinterpret_cast<uint64_t *>(row_end - n) = acc;
        }
        if (!live) return;
        int64_t *rec = A.rec + t * OV_NREC;
        A.ret[t] = (int64_t)k;
        for (int q = 0; q < OV_NREC; ++q) rec[q] = 0;
        if (A.mode != 0) {                                            // fm6_retrieve (exact.c:100-127): k, k2 and the containment flags
            rec[OV_LEN] = n; rec[OV_CONTAINED] = ret; rec[OV_X0] = (int64_t)intv0.x0; rec[OV_X1] = (int64_t)intv0.x1; rec[OV_X2] = (int64_t)intv0.x2;
            return;
        }
        for (int j = 0, r = n & 7; j < r; ++j)                        // the bases not yet stored (the first ones of the sequence)
            if (n - j <= ml) row_end[-(n - j)] = (uint8_t)(acc >> (8 * j));
        const int L = n <= ml ? n : -n;
        A.len[t] = L;
        rec[OV_LEN] = L; rec[OV_RBEG] = -1; rec[OV_LEFT] = 1;
        A.nei_cnt[t] = 0;
        A.np0[t] = -1;
        if (L <= min_match || empty) { rec[OV_CONTAINED] = -9; return; }       // unitig.c:288 (also: clipped sequences, re-run by the host)
        rec[OV_CONTAINED] = ret; rec[OV_X0] = (int64_t)intv0.x0; rec[OV_X1] = (int64_t)intv0.x1; rec[OV_X2] = (int64_t)intv0.x2;
        if (ovf) { rec[OV_CONTAINED] = -100; return; }
        if (ret < 0 || np == 0) return;
        A.np0[t] = np;
    }

    // ---- phase 3: the overlap_intv call of check_left_simple, unitig.c:186-190, for a unique neighbour (beg = 0)
    FMG_HD void phase_left1(int64_t t) {
        int64_t *rec = A.rec + t * OV_NREC;
        if (rec[OV_NNEI] != 1 || rec[OV_CONTAINED] != 0) return;
        const int L = A.len[t], sl = (int)rec[OV_SLEN], rbeg = (int)rec[OV_RBEG];
        const SeqView sv = {A.seq_of(t, L), L, A.ext + (size_t)t * A.max_len};
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
        const uint8_t *sq = A.seq_of(t, A.len[t]);
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

// ---------------------------------------------------------------------------------------------
// Phase 2: fm6_get_nei (unitig.c:93-179, beg = 0, the first `prev` list filled by phase 1) as a flat per-lane state machine with
// ONE index gather per LEVEL (nearly always), executed converged by the warp (persistent lanes, like smem_lane).
//
// The reference spends, per candidate P and level: fm6_extend(P, forward) for the children P.c and |P$|; fm6_extend0(ok[c], back)
// per non-empty child to learn whether some sequence still STARTS with P.c ($P.c not empty); fm6_extend0(ok[0], back) for $P$ when a
// sequence ends here -- two to three dependent rank2a calls at unrelated places of the BWT.  All of these are sizes of intervals of
// $P-strings, and the bi-interval of $P shares its x[1] with that of P (rc($P) = rc(P)$ sorts first among the suffixes starting with
// rc(P)).  So a candidate is kept as  (z0 = x0($P), x1, x2 = |P|, s = |$P|)  and ONE forward extension step ranks three positions
// of the same neighbourhood -- x1, x1 + s, x1 + x2, nearly always one 64-byte block:
//     |P.c|  = rank_c(x1 + x2) - rank_c(x1)      |$P.c| = rank_c(x1 + s) - rank_c(x1)      x1(P.c) = x1($P.c) = C[c] + rank_c(x1)
//     x0($P.c) = z0 + |$P.c'| summed over the c' before c in the order $,T,G,C,A (exact.c:81-86)
//     |P$| = |P.0|,  $P$ = ($P).0:  x0 = z0, x1 = C[0] + rank_0(x1), x2 = |$P.0|      -- the neighbour record of unitig.c:118-123
// which is everything fm6_get_nei reads off its three calls (x0 of P itself is never used).  The values are those of the
// reference because a bi-interval is a function of the string alone.  The rebuild of unitig.c:156-177 is a plain chain of forward
// extensions and runs through the same gather (s = 0).
//
// List entries keep the RAW sort key of unitig.c:132 in `info` (category of the parent | base | position); the category an entry
// gets at the end of its level (unitig.c:142-151) lives in a parallel array and is merged in when the entry is read.  Children
// arrive almost always in key order already (one category, one base), so the category is computed while pushing and the sort +
// renumbering pass of the reference only runs for a level whose pushes were out of order.  The first kShared entries of the two
// level lists (one entry per overlapping read: rarely more) live in shared memory, one column per lane; the rest in global lists.
template <typename U>
struct NeiLists {
    typedef IntvT<U> Cand;
    static constexpr int kShared = 6;
    const OverlapArgs &A;
    Cand *P, *Q; U *S; int32_t *cat;        // global overflow lists of the lane
    Cand *sh; U *shs; int32_t *shc; int stride;
    FMG_HD NeiLists(const OverlapArgs &a, int64_t lane, void *shared, int n_threads, int tid)
        : A(a), P(static_cast<Cand *>(a.A) + (size_t)lane * a.cap), Q(static_cast<Cand *>(a.B) + (size_t)lane * a.cap),
          S(static_cast<U *>(a.S) + (size_t)lane * a.cap * 2), cat(a.cat + (size_t)lane * a.cap * 2) {
        sh = static_cast<Cand *>(shared) + tid;
        shs = reinterpret_cast<U *>(static_cast<Cand *>(shared) + (size_t)n_threads * 2 * kShared) + tid;
        shc = reinterpret_cast<int32_t *>(reinterpret_cast<U *>(static_cast<Cand *>(shared) + (size_t)n_threads * 2 * kShared) + (size_t)n_threads * 2 * kShared) + tid;
        stride = n_threads;
    }
    static constexpr size_t shared_bytes(int n_threads) { return (size_t)n_threads * 2 * kShared * (sizeof(Cand) + sizeof(U) + sizeof(int32_t)); }
    FMG_HD Cand lget(int w, int i) const { return i < kShared ? sh[(w * kShared + i) * stride] : ld_cand((w ? Q : P) + i); }
    FMG_HD void lput(int w, int i, const Cand &k) { if (i < kShared) sh[(w * kShared + i) * stride] = k; else st_cand((w ? Q : P) + i, k); }
    FMG_HD U sget(int w, int i) const { return i < kShared ? shs[(w * kShared + i) * stride] : S[w * A.cap + i]; }
    FMG_HD void sput(int w, int i, U v) { if (i < kShared) shs[(w * kShared + i) * stride] = v; else S[w * A.cap + i] = v; }
    FMG_HD int cget(int w, int i) const { return i < kShared ? shc[(w * kShared + i) * stride] : cat[w * A.cap + i]; }
    FMG_HD void cput(int w, int i, int v) { if (i < kShared) shc[(w * kShared + i) * stride] = v; else cat[w * A.cap + i] = v; }
};

enum { NP_FETCH = 0, NP_NEXT, NP_END_LEVEL, NP_FINISH, NP_FINISH2, NP_CAND, NP_FIX1, NP_FIX2, NP_DONE };     // >= NP_CAND: wants a gather (or is done)

template <typename U, int TAG, class FetchFn>
FMG_HD void nei_lane(const OverlapArgs &A, int64_t lane, FetchFn fetch, void *shared, int n_threads, int tid) {
    typedef IntvT<U> Cand;
    typedef OvBits<U> BT;
    NeiLists<U> L(A, lane, shared, n_threads, tid);
    int ph = NP_FETCH;
    int64_t t = 0, t_next = fetch();                                     // the next sequence is claimed one ahead: the atomic's latency hides behind the current one
    int ori_l = 0, sl = 0, np = 0, npc = 0, np_first = 0, j = 0, nq = 0, pv = 0, cv = 1;
    int first_base = 0, cat_run = 0, nnei = 0, is_forked = 0, cj = 0, fi = 0, rbeg = 0;
    bool sorted = true, ovf = false;
    U last_key = 0, last_hi = 0, ps = 0;
    Cand p = {0, 0, 0, 0}, nei0 = {0, 0, 0, 0}, k0 = {0, 0, 0, 0};      // p: the candidate (z0, x1, x2, key); k0: the interval of the rebuild chain
    const uint8_t *sq = nullptr;
    uint8_t *xt = nullptr;
    int64_t *rec = nullptr;
    uint4 *nei = nullptr;

    for (;;) {
        // ---- advance to the next gather request (no index access in here)
        warp_rejoin();
        while (ph < NP_CAND) {
            if (ph == NP_FETCH) {
                t = t_next;
                if (t >= A.n) { ph = NP_DONE; break; }
                t_next = fetch();
                np_first = A.np0[t];
                if (np_first <= 0) continue;
                A.np0[t] = -1;
                rec = A.rec + t * OV_NREC;
                ori_l = A.len[t]; sl = ori_l;
                sq = A.seq_of(t, ori_l);
                xt = A.ext + (size_t)t * A.max_len;
                nei = A.nei + (size_t)t * A.nei_cap * 2;
                ovf = false; nnei = 0; is_forked = 0;
                np = np_first; npc = np < A.pcap ? np : A.pcap;
                {
                    // the list of phase 1 (its slot was filled from the end: the smallest interval, pushed last, comes first) becomes the
                    // first `prev` list, all in category 0 (unitig.c:105-106); four entries are requested at a time
                    const size_t o = (size_t)t * A.pcap + (A.pcap - np_first);
                    const Cand *src = static_cast<const Cand *>(A.P0) + o;
                    const U *ssrc = static_cast<const U *>(A.S0) + o;
                    for (int i = 0; i < npc; i += 4) {
                        Cand c4[4]; U s4[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) if (i + q < npc) { c4[q] = ld_cand(src + i + q); s4[q] = ssrc[i + q]; }
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (i + q < npc) {
                                c4[q].info = (U)(ori_l - (int)c4[q].info);       // suffix length -> start of the suffix in the read (unitig.c:55)
                                L.lput(0, i + q, c4[q]); L.sput(0, i + q, s4[q]); L.cput(0, i + q, 0);
                            }
                    }
                }
                pv = 0; cv = 1;
                j = 0; nq = 0; first_base = 0; cat_run = 0; sorted = true; last_key = 0; last_hi = 0;
                ph = NP_NEXT;
            } else if (ph == NP_NEXT) {                                  // the next candidate of `prev` that is still alive (unitig.c:109)
                if (j >= npc) { ph = NP_END_LEVEL; continue; }
                cj = L.cget(pv, j);
                if (cj < 0) { ++j; continue; }
                p = L.lget(pv, j); ps = L.sget(pv, j);
                p.info = (U)((p.info & BT::pos_mask) | ((U)cj << BT::cat_shift));
                ph = NP_CAND;
            } else if (ph == NP_END_LEVEL) {                             // update categories, unitig.c:137-153
                if (nq) {
                    const int nqc = nq < A.cap ? nq : A.cap;
                    if (sl - ori_l < A.max_len) { xt[sl - ori_l] = (uint8_t)comp6(first_base); ++sl; } else ovf = true;
                    if (!sorted) {
                        for (int a = 1; a < nqc; ++a) {                  // insertion sort by info (keys are unique)
                            const Cand x = L.lget(cv, a);
                            const U xs = L.sget(cv, a);
                            int b = a - 1;
                            while (b >= 0) {
                                const Cand y = L.lget(cv, b);
                                if (y.info <= x.info) break;
                                L.lput(cv, b + 1, y); L.sput(cv, b + 1, L.sget(cv, b)); --b;
                            }
                            L.lput(cv, b + 1, x); L.sput(cv, b + 1, xs);
                        }
                        U last = 0;
                        cat_run = 0;
                        for (int a = 0; a < nqc; ++a) {
                            const U hi = (U)(L.lget(cv, a).info >> BT::pos_bits);
                            if (a == 0 || hi != last) { last = hi; cat_run = a; }
                            L.cput(cv, a, cat_run);
                        }
                    }
                    if (cat_run != 0) is_forked = 1;
                    if (cat_run > BT::max_cat) ovf = true;
                }
                pv = cv; cv ^= 1;
                np = nq;
                if (np) {
                    npc = np < A.cap ? np : A.cap;
                    j = 0; nq = 0; first_base = 0; cat_run = 0; sorted = true; last_key = 0; last_hi = 0;
                    ph = NP_NEXT;
                } else ph = NP_FINISH;
            } else if (ph == NP_FINISH) {
                A.nei_cnt[t] = (uint32_t)nnei;
                rec[OV_NNEI] = nnei;
                if (nnei == 0) {                                         // unitig.c:154 (returns -1, s keeps its growth)
                    rec[OV_SLEN] = sl;
                    if (ovf) rec[OV_CONTAINED] = -100;
                    ph = NP_FETCH;
                    continue;
                }
                rbeg = ori_l - (int)nei0.info;
                ph = NP_FINISH2;
                if (nnei == 1 && is_forked) {                            // contained reads forked the path: rebuild it along the one neighbour
                    k0 = base_intv<U>(A.ix, 0);
                    fi = rbeg;
                    if (fi < ori_l) ph = NP_FIX1;
                    else if (fi < sl) ph = NP_FIX2;
                }
            } else {                                                     // NP_FINISH2
                if (nnei > 1) sl = ori_l;
                rec[OV_RBEG] = rbeg; rec[OV_SLEN] = sl;
                if (ovf) rec[OV_CONTAINED] = -100;
                ph = NP_FETCH;
            }
        }
        // ---- the one gather of this trip (converged): the block(s) around x1 of the lane's candidate -- the candidates of a level are
        // suffixes of one read, i.e. rc(P) are prefixes of one string, so their x1 lie within a few BWT rows of each other and the
        // gather nearly always serves the whole level
        const bool active = ph != NP_DONE;
        if (!warp_any(active)) return;
        const bool cand = ph == NP_CAND;
        const uint64_t pa = cand ? (uint64_t)p.x1 : (uint64_t)k0.x1, pb = pa + (cand ? (uint64_t)p.x2 : (uint64_t)k0.x2);
        const uint64_t b1 = pa >> kBlkShift, b2 = pb >> kBlkShift;
        const bool two = active && b2 != b1;
        const bool any_two = warp_any(two);          // one lane in a dozen needs a second block, so most warps do: both gathers are issued before either is waited for
        const PairReq qa = pair_issue(A.ix, pa, active);
        PairReq qb;
        if (any_two) qb = pair_issue(A.ix, pb, two);
        const uint64_t sba = pa >> kSuperShift;      // C[c] + occurrences before the superblock of x1, requested with the blocks (the row is in L1 / L2)
        const uint64_t *rowa = A.ix.cs + sba * 8;
        U csa[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) csa[c] = active ? (U)ld_u64(rowa + c) : (U)0;
        const Blk B1 = pair_finish(qa);
        Blk B2 = B1;
        if (any_two) { const Blk Bx = pair_finish(qb); if (two) B2 = Bx; }
        warp_rejoin();
        if (!active) continue;
        // rank at position q of one of the two blocks, relative to the superblock of x1 (q's superblock differs: rare)
        auto rank_at = [&](uint64_t q, U r[6]) {
            uint32_t r32[6];
            rank_rel((q >> kBlkShift) == b1 ? B1 : B2, q, r32);
#pragma unroll
            for (int c = 0; c < 6; ++c) r[c] = r32[c];
            if ((q >> kSuperShift) != sba) {
                const uint64_t *r2 = A.ix.cs + (q >> kSuperShift) * 8;
#pragma unroll
                for (int c = 0; c < 6; ++c) r[c] += (U)(ld_u64(r2 + c) - ld_u64(rowa + c));
            }
        };
        auto have_blk = [&](uint64_t q) { const uint64_t bq = q >> kBlkShift; return bq == b1 || (two && bq == b2); };

        // ---- consume
        if (cand) {
            // candidate j, then the following ones of the level while the fetched blocks hold their three positions
            for (;;) {
                const uint64_t qa_ = (uint64_t)p.x1, qm_ = qa_ + ps, qb_ = qa_ + p.x2;
                U ra[6], rm[6], rb[6];
                rank_at(qa_, ra);
                rank_at(qb_, rb);
                if (have_blk(qm_)) rank_at(qm_, rm);
                else {                                                   // an interval wider than a block with $P ending in between: rare
                    uint32_t r32[6];
                    rank_rel(load_blk(A.ix, qm_), qm_, r32);
#pragma unroll
                    for (int c = 0; c < 6; ++c) rm[c] = r32[c];
                    if ((qm_ >> kSuperShift) != sba) {
                        const uint64_t *r2 = A.ix.cs + (qm_ >> kSuperShift) * 8;
#pragma unroll
                        for (int c = 0; c < 6; ++c) rm[c] += (U)(ld_u64(r2 + c) - ld_u64(rowa + c));
                    }
                }
                const U s0 = (U)(rb[0] - ra[0]);                         // |P$|: sequences ending with P
                const U t0 = (U)(rm[0] - ra[0]);                         // |$P$|: sequences equal to P
                bool found = false;
                if (s0 != 0 && ori_l != sl && t0 != 0 && s0 == p.x2 && p.x2 == t0) {     // a full read, not contained in a longer one (unitig.c:112-125)
                    Cand nb;                                             // fm6_extend0's ok0 (x[0] = tk[0]; cnt[0] is 0)
                    nb.x0 = p.x0; nb.x1 = (U)(csa[0] + ra[0]); nb.x2 = t0;
                    nb.info = (U)(ori_l - (int)(p.info & BT::pos_mask));
                    for (int i = j; i < npc && L.cget(pv, i) == cj; ++i) L.cput(pv, i, -1);      // mask out the other intervals of the category
                    if (nnei < A.nei_cap) {
                        Intv o; o.x0 = nb.x0; o.x1 = nb.x1; o.x2 = nb.x2; o.info = nb.info;
                        st_intv(nei + 2 * nnei, o);
                    } else ovf = true;
                    if (nnei == 0) nei0 = nb;
                    ++nnei;
                    found = true;
                }   // a read contained in another one is only marked `used` by the reference
                if (!found) {                                            // collect the extensible intervals (unitig.c:127-135)
                    U zc[5];                                             // x0 of $P.c, cumulative in the order $,T,G,C,A (exact.c:81-86)
                    zc[4] = (U)(p.x0 + t0); zc[3] = (U)(zc[4] + (U)(rm[4] - ra[4])); zc[2] = (U)(zc[3] + (U)(rm[3] - ra[3])); zc[1] = (U)(zc[2] + (U)(rm[2] - ra[2]));
#pragma unroll
                    for (int c = 1; c < 5; ++c) {
                        const U sc = (U)(rm[c] - ra[c]);                 // |$P.c|: left end still bounded by a sentinel
                        if (sc == 0) continue;
                        Cand kc;
                        kc.x0 = zc[c]; kc.x1 = (U)(csa[c] + ra[c]); kc.x2 = (U)(rb[c] - ra[c]);
                        kc.info = (U)((p.info & ~BT::base_mask) | ((U)c << BT::pos_bits));
                        const U hi = (U)(kc.info >> BT::pos_bits);
                        if (nq == 0) first_base = c;
                        else if (kc.info < last_key) sorted = false;
                        if (nq == 0 || hi != last_hi) { last_hi = hi; cat_run = nq; }
                        last_key = kc.info;
                        if (nq < A.cap) { L.cput(cv, nq, cat_run); L.lput(cv, nq, kc); L.sput(cv, nq, sc); } else ovf = true;
                        ++nq;
                    }
                }
                // the next live candidate of the level, if the blocks in registers serve it too
                ++j;
                while (j < npc && (cj = L.cget(pv, j)) < 0) ++j;
                if (j >= npc) break;
                const Cand pn = L.lget(pv, j);
                const U sn = L.sget(pv, j);
                const uint64_t na = (uint64_t)pn.x1;
                if (!have_blk(na) || !have_blk(na + sn) || !have_blk(na + pn.x2) || (na >> kSuperShift) != sba) break;
                p = pn; ps = sn;
                p.info = (U)((p.info & BT::pos_mask) | ((U)cj << BT::cat_shift));
            }
            ph = NP_NEXT;
        } else {                                                         // the rebuild chain, unitig.c:158-176: forward extension of k0
            U ra[6], rb[6], sz[6], nr[6];
            rank_at(pa, ra);
            rank_at(pb, rb);
#pragma unroll
            for (int c = 0; c < 6; ++c) sz[c] = (U)(rb[c] - ra[c]);
            nr[0] = k0.x0; nr[4] = (U)(nr[0] + sz[0]); nr[3] = (U)(nr[4] + sz[4]); nr[2] = (U)(nr[3] + sz[3]); nr[1] = (U)(nr[2] + sz[2]); nr[5] = (U)(nr[1] + sz[1]);
            int csel = -1;
            if (ph == NP_FIX1) csel = comp6(sq[fi]);
            else {
                int hits = 0;
#pragma unroll
                for (int c = 1; c < 5; ++c)
                    if (sz[c] != 0 && nr[c] <= nei0.x0 && (uint64_t)nr[c] + sz[c] >= (uint64_t)nei0.x0 + nei0.x2) ++hits, csel = c;
                if (hits == 0 && sz[0] != 0) csel = -1;
                else if (hits != 1) { ovf = true; csel = -1; }           // the reference asserts hits == 1 (unitig.c:171)
                if (csel < 0) { sl = fi; ph = NP_FINISH2; continue; }
                xt[fi - ori_l] = (uint8_t)comp6(csel);
            }
            k0.x0 = pick6(nr, csel); k0.x1 = (U)(pick6(csa, csel) + pick6(ra, csel)); k0.x2 = pick6(sz, csel);
            ++fi;
            if (ph == NP_FIX1) { if (fi >= ori_l) ph = fi < sl ? NP_FIX2 : NP_FINISH2; }
            else if (fi >= sl) { sl = fi; ph = NP_FINISH2; }
        }
    }
}

// the candidate loop of check_left_simple (phase 4): persistent lanes; a lane that runs out of work keeps answering the warp votes
template <typename U, int PHASE, class FetchFn>
FMG_HD void overlap_lane_sync(const OverlapArgs &A, int64_t lane, FetchFn fetch) {
    OvLane<U, true, PHASE> ln(A, lane);
    for (;;) {
        const int64_t t = fetch();
        if (t >= A.n) break;
        ln.phase_left2(t);
    }
    while (ext_sync<U, PHASE>(A.ix, false, 0, 0, 0, 0, 0).flags & 1) {}
}

// chain phases: one sequence per thread; phase 1 is run by whole warps (`live` = this lane has a sequence)
template <typename U, int PHASE>
FMG_HD void overlap_chain(const OverlapArgs &A, int64_t t, bool live = true) {
    OvLane<U, false, PHASE> ln(A, 0);
    if (PHASE == 1) ln.phase_contained(t, live); else if (live) ln.phase_left1(t);
}

} // namespace fmg
