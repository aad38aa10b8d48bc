// Host description of the "occ block" query layout (see fmd_device.cuh) and its host-side builder.
#pragma once
#include <cstdint>
#include <vector>
#include "fmd_host.hpp"

namespace fmg {

struct OccHost {
    std::vector<uint32_t> blocks;    // n_blocks x 16 u32 (64 B per block of 128 symbols)
    std::vector<uint64_t> cs;        // n_super x 8: C[c] + count of c before the superblock
    uint64_t n_blocks = 0, n_super = 0, n_sym = 0;
};

// rank positions run over [0, n], so position n must be addressable
inline uint64_t occ_n_blocks(uint64_t n_sym) { return (n_sym >> 7) + 1; }
inline uint64_t occ_n_super(uint64_t n_sym) { return (n_sym >> 24) + 1; }

OccHost build_occ_host(const FmdImage &img);

} // namespace fmg
