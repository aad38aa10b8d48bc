// Host description of the "occ line" query layout (see fmd_device.cuh) and its host-side builder
// used for small indexes and by the host emulation tests; large images are transcoded on the GPU
// (occ_build.cu).
#pragma once
#include <cstdint>
#include <vector>
#include "fmd_host.hpp"

namespace fmg {

struct OccHost {
    std::vector<uint64_t> lines;     // n_lines x 16 u64 (128 B per line)
    std::vector<uint64_t> super;     // n_super x 8 (empty => u32 counts are absolute)
    uint64_t n_lines = 0, n_sym = 0;
};

// number of lines needed for n symbols: rank positions run over [0, n], so position n must be addressable
inline uint64_t occ_n_lines(uint64_t n_sym) { return (n_sym >> 8) + 1; }
inline bool occ_needs_super(uint64_t n_sym) { return n_sym + 256 >= (1ull << 32); }

OccHost build_occ_host(const FmdImage &img);

} // namespace fmg
