// Host-side view of the overlap records of ALL sequences of an index, as the unitig walk consumes them
// (unitig_host.cpp).  Filled by fmg_overlap_all (overlap.cu: GPU pass into pinned memory) or, for records that
// arrive through the C-ABI as separate arrays (multi-GPU gather, tests), by fmg_unitig_assemble.
#pragma once
#include <cstdint>
#include "fmd_overlap.cuh"
#include "../../include/fermi_b200.h"

struct OvHost {
    uint64_t n_seq = 0;
    int max_len = 0;
    const fmg::OvPack *pack = nullptr;        // [n_seq], indexed by the rank of the sequence
    const uint64_t *rank_of_row = nullptr;    // [n_seq]: value fm_retrieve returns for BWT row t (exact.c:59-70)
    const uint8_t *seq = nullptr;             // sequences of the seed rows; row t at seq[(seq_odd_only ? t >> 1 : t) * seq_stride]
    uint64_t seq_stride = 0;
    int seq_odd_only = 0;                     // only the odd rows (the seeds of unitig_core, unitig.c:333-334) are present
    const uint8_t *ext = nullptr;             // appended bases, addressed by OvPack::ext_first
    const fmg_intv_t *spill = nullptr;        // neighbour lists of the records with more than one neighbour (OvPack::nx0)
    uint64_t ext_total = 0, spill_total = 0;
};

struct fmg_index_s;
// overlap.cu: records of every sequence of the index; the arrays live in a pinned host cache owned by the library
// (valid until the next call or fmg_release_cache).  Returns 0, or -1 on a CUDA error.
int fmg_overlap_all(const fmg_index_s *idx, int min_match, int max_len, OvHost *out);
// unitig_host.cpp: the walk (unitig.c:227-362) + MAG text
int fmg_unitig_walk(const OvHost &R, int min_match, const char *out_path, uint64_t *n_unitigs);
