// Internal definitions shared by the translation units of libfermi_b200.
#pragma once
#include <cstdint>
#include <mutex>
#include "fmd_host.hpp"
#include "fmd_device.cuh"

// k_smem launch shape: 128-thread blocks, as many as the register budget allows per SM
#define SMEM_BLOCK 128
#define SMEM_MIN_BLOCKS_U32 6
#define SMEM_MIN_BLOCKS_U64 5
#define SMEM_DEFAULT_OUT_CAP 64
// overlap kernels: list-chasing phases (persistent lanes) and chain phases (one sequence per thread)
#define OVLP_BLOCK 128
#ifndef OVLP_MIN_BLOCKS
#define OVLP_MIN_BLOCKS 4
#endif
#define OVCH_BLOCK 128

struct fmg_fmd_s { fmg::FmdImage img; };

struct fmg_pipe_s;
void fmg_pipe_destroy(fmg_pipe_s *p);
struct fmg_ovcache_s;
void fmg_ovcache_destroy(fmg_ovcache_s *p);

struct fmg_index_s {
    int device = 0, n_sm = 0;
    uint32_t *d_blocks = nullptr;
    uint64_t *d_cs = nullptr;
    uint64_t n_blocks = 0, bytes = 0;
    uint64_t mcnt[8] = {0}, cnt[8] = {0};
    fmg::OccView view;
    // lazily created host-buffer SMEM pipeline (fmg_cuda.cu), reused between calls
    mutable fmg_pipe_s *pipe = nullptr;
    mutable std::mutex pipe_lock;
    // pinned host arrays of the last whole-index overlap pass (overlap.cu: fmg_overlap_all), reused between calls
    mutable fmg_ovcache_s *ovc = nullptr;
    mutable std::mutex ov_lock;          // one overlap pass / unitig assembly at a time per index handle (they share ovc)
};

// occ_build.cu: build the occ blocks of `img` in the HBM of the current device; fills d_blocks/d_cs/n_blocks/bytes
int occ_build_device(const fmg::FmdImage &img, fmg_index_s *idx);
