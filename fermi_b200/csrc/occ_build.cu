// Index upload: .fmd image -> occ blocks in HBM (layout in fmd_device.cuh).
#include <cuda_runtime.h>
#include <cstdio>
#include "fmg_internal.hpp"
#include "occ_layout.hpp"
#include "../../include/fermi_b200.h"

using namespace fmg;

int occ_build_device(const FmdImage &img, fmg_index_s *idx) {
    OccHost occ = build_occ_host(img);
    idx->n_blocks = occ.n_blocks;
    idx->bytes = occ.blocks.size() * 4 + occ.cs.size() * 8;
    cudaError_t err = cudaMalloc(&idx->d_blocks, occ.blocks.size() * 4);
    if (err == cudaSuccess) err = cudaMemcpy(idx->d_blocks, occ.blocks.data(), occ.blocks.size() * 4, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMalloc(&idx->d_cs, occ.cs.size() * 8);
    if (err == cudaSuccess) err = cudaMemcpy(idx->d_cs, occ.cs.data(), occ.cs.size() * 8, cudaMemcpyHostToDevice);
    if (err != cudaSuccess) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::fmg_index_upload] %s\n", cudaGetErrorString(err));
        cudaFree(idx->d_blocks); cudaFree(idx->d_cs);
        idx->d_blocks = nullptr; idx->d_cs = nullptr;
        return -1;
    }
    return 0;
}
