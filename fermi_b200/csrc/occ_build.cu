// Index upload: .fmd image -> occ lines in HBM (layout in fmd_device.cuh).
#include <cuda_runtime.h>
#include <cstdio>
#include "fmg_internal.hpp"
#include "occ_layout.hpp"
#include "../../include/fermi_b200.h"

using namespace fmg;

int occ_build_device(const FmdImage &img, fmg_index_s *idx) {
    OccHost occ = build_occ_host(img);
    idx->n_lines = occ.n_lines;
    idx->bytes = occ.lines.size() * 8 + occ.super.size() * 8;
    cudaError_t err = cudaMalloc(&idx->d_lines, occ.lines.size() * 8);
    if (err == cudaSuccess) err = cudaMemcpy(idx->d_lines, occ.lines.data(), occ.lines.size() * 8, cudaMemcpyHostToDevice);
    if (err == cudaSuccess && !occ.super.empty()) {
        err = cudaMalloc(&idx->d_super, occ.super.size() * 8);
        if (err == cudaSuccess) err = cudaMemcpy(idx->d_super, occ.super.data(), occ.super.size() * 8, cudaMemcpyHostToDevice);
    }
    if (err != cudaSuccess) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::fmg_index_upload] %s\n", cudaGetErrorString(err));
        cudaFree(idx->d_lines); cudaFree(idx->d_super);
        idx->d_lines = nullptr; idx->d_super = nullptr;
        return -1;
    }
    return 0;
}
