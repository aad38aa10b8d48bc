// TEST INFRASTRUCTURE ONLY.
// Compiles the product's device-side core (fermi_b200/csrc/fmd_device.cuh) for the HOST so that the
// CPU-only test suite can check the exact code the kernels run (occ-line rank, fm6_extend, the SMEM
// lane state machine) against the oracle without a GPU.  The product never links this file and has
// no CPU execution path; libfermi_b200.so fails loudly without a CUDA device.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../fermi_b200/csrc/fmd_device.cuh"
#include "../../fermi_b200/csrc/fmd_overlap.cuh"
#include "../../fermi_b200/csrc/occ_layout.hpp"
#include "../../fermi_b200/csrc/bcr_tile.cuh"
#include "../../include/fermi_b200.h"

using namespace fmg;

struct fmg_fmd_s { FmdImage img; };

struct EmuIndex {
    OccHost occ;
    OccView view;
};

extern "C" {

void *emu_index_build(const fmg_fmd_t *e) {
    EmuIndex *x = new EmuIndex;
    x->occ = build_occ_host(e->img);
    x->view.blocks = x->occ.blocks.data();
    x->view.cs = x->occ.cs.data();
    x->view.n_sym = e->img.mcnt[0];
    x->view.n_seq = e->img.mcnt[1];
    for (int c = 0; c < 8; ++c) x->view.C[c] = e->img.cnt[c];
    return x;
}

void emu_index_free(void *x) { delete static_cast<EmuIndex *>(x); }

void emu_rank2a(const void *_x, int64_t n, const uint64_t *k, const uint64_t *l, uint64_t *ok, uint64_t *ol) {
    const OccView &ix = static_cast<const EmuIndex *>(_x)->view;
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t pk = k[i] + 1, pl = l[i] + 1;      // k == -1 -> p = 0
        uint32_t rk[6], rl[6];
        rank_rel(load_blk(ix, pk), pk, rk);
        rank_rel(load_blk(ix, pl), pl, rl);
        for (int c = 0; c < 6; ++c) {
            ok[6 * i + c] = ix.cs[(pk >> kSuperShift) * 8 + c] - ix.C[c] + rk[c];
            ol[6 * i + c] = ix.cs[(pl >> kSuperShift) * 8 + c] - ix.C[c] + rl[c];
        }
    }
}

void emu_extend(const void *_x, int64_t n, const fmg_intv_t *ik, const uint8_t *is_back, fmg_intv_t *ok6) {
    const OccView &ix = static_cast<const EmuIndex *>(_x)->view;
    for (int64_t i = 0; i < n; ++i) {
        const int b = is_back[i] != 0;
        Ext6 e;
        extend6<uint64_t>(ix, ik[i].x[b], ik[i].x[!b], ik[i].x[2], e);
        for (int c = 0; c < 6; ++c) {
            fmg_intv_t &o = ok6[6 * i + c];
            o.x[!b] = far_of(ix, e, c); o.x[b] = e.near[c]; o.x[2] = e.size[c]; o.info = 0;
        }
    }
}

// runs smem_lane with `n_lanes` emulated lanes (executed one after the other; lanes are independent)
int emu_smem(const void *_x, int64_t n, const uint8_t *seq, const uint64_t *off, int self_match, int n_lanes,
             int out_cap, int wide, fmg_intv_t **mem, uint64_t *mem_off) {
    const EmuIndex *x = static_cast<const EmuIndex *>(_x);
    int max_len = 1;
    for (int64_t i = 0; i < n; ++i) if ((int)(off[i + 1] - off[i]) > max_len) max_len = (int)(off[i + 1] - off[i]);
    const int cap = 2 * max_len + 2;
    std::vector<uint4> F((size_t)n_lanes * cap * 2), W((size_t)n_lanes * cap * 2), out((size_t)n * out_cap * 2);
    std::vector<uint32_t> cnt(n, 0);
    unsigned long long next = 0;
    SmemArgs A;
    A.ix = x->view; A.seq = seq; A.off = off; A.n_reads = n; A.self_match = self_match;
    A.F = F.data(); A.W = W.data(); A.cap = cap; A.out = out.data(); A.out_cap = out_cap;
    unsigned long long too_long = 0;
    A.rec_cnt = cnt.data(); A.next_read = &next; A.max_len = max_len; A.too_long = &too_long;
    // interleave the lanes' reads like concurrent lanes would: lane t takes reads t, t+n_lanes, ...
    for (int t = 0; t < n_lanes; ++t) {
        int64_t cur = t;
        auto fetch = [&]() { int64_t r = cur; cur += n_lanes; return r; };
        if (wide) smem_lane<uint64_t>(A, t, fetch);      // 64-bit coordinates (any index size)
        else smem_lane<uint32_t>(A, t, fetch);           // 32-bit coordinates (BWT < 2^32 symbols)
    }
    int overflow = 0;
    mem_off[0] = 0;
    for (int64_t i = 0; i < n; ++i) {
        if ((int)cnt[i] > out_cap) overflow = 1;
        mem_off[i + 1] = mem_off[i] + (cnt[i] > (uint32_t)out_cap ? (uint32_t)out_cap : cnt[i]);
    }
    *mem = (fmg_intv_t *)std::malloc((mem_off[n] ? mem_off[n] : 1) * sizeof(fmg_intv_t));
    for (int64_t i = 0; i < n; ++i)
        std::memcpy(*mem + mem_off[i], out.data() + (size_t)i * out_cap * 2, (mem_off[i + 1] - mem_off[i]) * 32);
    return overflow;
}

// retrieve + overlap record of the sequences ids[0..n) (both through the product's lane code)
int emu_overlap(const void *_x, int min_match, int64_t n, const uint64_t *ids, int max_len, int cap, int nei_cap, int wide,
                int64_t *rec, fmg_intv_t *nei, uint32_t *nei_cnt, uint8_t *seq, int32_t *len, uint8_t *ext) {
    const EmuIndex *x = static_cast<const EmuIndex *>(_x);
    std::vector<int64_t> ret(n);
    // the four phases in the order the product's fmg_overlap_batch launches them, over the whole batch.  Device rows have a
    // stride that is a multiple of 8 and hold the sequence right-aligned (fmd_overlap.cuh: OverlapArgs::seq_of).
    const int ml = (max_len + 7) & ~7;
    const int n_lanes = 3, pcap = (ml - min_match + 8 > 8 ? ml - min_match + 8 : 8);
    std::vector<uint64_t> P0((size_t)n * pcap * 4), A((size_t)n_lanes * cap * 4), B((size_t)n_lanes * cap * 4);
    std::vector<int32_t> cat((size_t)n_lanes * cap * 2), np0(n);
    std::vector<uint64_t> S0((size_t)n * pcap), S((size_t)n_lanes * cap * 2);
    std::vector<uint8_t> dseq((size_t)n * ml + 8), dext((size_t)n * ml + 8);
    // stand-in for the shared-memory part of the level lists of phase 2 (one lane at a time)
    std::vector<uint64_t> shared(2 * 8 * 6 + 8);
    OverlapArgs O;
    O.ix = x->view; O.min_match = min_match; O.mode = 0; O.n = n; O.seq = dseq.data(); O.len = len; O.max_len = ml;
    O.ids = ids; O.first = 0; O.step = 1; O.ret = ret.data();
    O.P0 = P0.data(); O.S0 = S0.data(); O.S = S.data(); O.pcap = pcap; O.np0 = np0.data(); O.A = A.data(); O.B = B.data(); O.cap = cap; O.cat = cat.data();
    O.rec = rec; O.nei = reinterpret_cast<uint4 *>(nei); O.nei_cap = nei_cap; O.nei_cnt = nei_cnt; O.ext = dext.data(); O.next = nullptr; O.order = nullptr;
    auto lists = [&](int phase) {
        for (int t = 0; t < n_lanes; ++t) {
            int64_t cur = t;
            auto fetch = [&]() { int64_t r = cur; cur += n_lanes; return r; };
            if (wide) { if (phase == 2) nei_lane<uint64_t, 0>(O, t, fetch, shared.data(), 1, 0); else overlap_lane_sync<uint64_t, 4>(O, t, fetch); }
            else { if (phase == 2) nei_lane<uint32_t, 0>(O, t, fetch, shared.data(), 1, 0); else overlap_lane_sync<uint32_t, 4>(O, t, fetch); }
        }
    };
    for (int64_t t = 0; t < n; ++t) { if (wide) overlap_chain<uint64_t, 1>(O, t); else overlap_chain<uint32_t, 1>(O, t); }
    lists(2);
    for (int64_t t = 0; t < n; ++t) { if (wide) overlap_chain<uint64_t, 3>(O, t); else overlap_chain<uint32_t, 3>(O, t); }
    lists(4);
    for (int64_t t = 0; t < n; ++t) {
        const int l = len[t] < 0 ? ml : len[t];
        std::memset(seq + (size_t)t * max_len, 0, max_len);
        std::memcpy(seq + (size_t)t * max_len, dseq.data() + (size_t)(t + 1) * ml - l, l < max_len ? l : max_len);
        std::memcpy(ext + (size_t)t * max_len, dext.data() + (size_t)t * ml, max_len);
    }
    for (int64_t t = 0; t < n; ++t) rec[t * OV_NREC + OV_K] = ret[t];
    return 0;
}

// One tile of the BCR merge pass (bcr.cu: k_bcr_merge) with the kernel's per-thread arithmetic (bcr_tile.cuh) and its block scans
// replaced by a serial loop over the threads.  tile_len <= n_threads * k_per output positions; flags / syms per position (syms valid
// where flagged); staged = the old symbols of the tile starting at byte `shift` of a word-aligned buffer (8 bytes of slack after
// the last one).  Out: the merged symbols, the A,C,G,T histogram and the in-tile rank of every insert in position order.
int emu_bcr_tile(int k_per, int n_threads, int tile_len, const uint8_t *flags_in, const uint8_t *syms_in, const uint32_t *staged, int shift,
                 uint8_t *out, uint64_t hist[4], uint32_t *ranks) {
    if ((k_per != 16 && k_per != 32) || tile_len > n_threads * k_per) return -1;
    const int words = k_per / 4, tile = n_threads * k_per;
    std::vector<uint8_t> flags(tile, 0), syms(tile, 0xee);           // garbage where no insert: the kernel's s_sym is stale there too
    std::memcpy(flags.data(), flags_in, tile_len);
    for (int i = 0; i < tile_len; ++i) if (flags[i]) syms[i] = syms_in[i];
    uint32_t spread[16];
    for (uint32_t f = 0; f < 16; ++f) spread[f] = bcr_spread_selector(f);
    uint32_t ins_before = 0;
    uint64_t cnt_before = 0;
    for (int tid = 0; tid < n_threads; ++tid) {
        const int j0 = tid * k_per;
        uint32_t flw[8], syw[8], o[8], eq[4];
        std::memcpy(flw, flags.data() + j0, k_per); std::memcpy(syw, syms.data() + j0, k_per);
        uint32_t ib = ins_before;
        for (int w = 0; w < words; ++w) {
            o[w] = bcr_merge_word(staged, shift + j0 + 4 * w - (int)ib, flw[w], syw[w], spread);
            ib += bcr_flag_count(flw[w]);
        }
        for (int t = 0; t < k_per; ++t) if (j0 + t >= tile_len) o[t >> 2] |= 7u << (8 * (t & 3));
        if (words == 4) bcr_base_masks<4>(o, eq); else bcr_base_masks<8>(o, eq);
        const uint64_t my_cnt = bcr_pack_counts(eq);
        uint32_t k = ins_before;
        for (int w = 0; w < words; ++w)
            for (int b = 0; b < 4; ++b)
                if ((flw[w] >> (8 * b)) & 1u) ranks[k++] = bcr_insert_rank(o[w], w, 8 * b, eq, cnt_before);
        for (int t = 0; t < k_per; ++t) if (j0 + t < tile_len) out[j0 + t] = (uint8_t)(o[t >> 2] >> (8 * (t & 3)));
        ins_before = ib; cnt_before += my_cnt;
    }
    for (int c = 0; c < 4; ++c) hist[c] = (cnt_before >> (16 * c)) & 0xffff;
    return (int)ins_before;
}

} // extern "C"
