// Micro-benchmark: dependent random gathers of 64-byte (or 128-byte) blocks from an array >> L2, the access pattern of the
// FMD-index extension chains (one block per step, the next address depends on the loaded data).  Sweeps the L2 fetch
// granularity limit, the resident warps per SM and the number of independent chains per thread, so that the ceiling of the
// pattern on this GPU is a measured number:   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bench gather_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

struct alignas(32) V8 { uint32_t v[8]; };
__device__ __forceinline__ V8 ld256(const void *p) {
    V8 r;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 29; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 32; return x; }

template <int K, int BYTES>
__global__ void k_gather(const uint32_t *__restrict__ a, uint64_t n_blk, int steps, uint64_t *out) {
    extern __shared__ uint8_t pad[];
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t idx[K], acc = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) idx[k] = mix(t * K + k + 1) % n_blk;
    for (int s = 0; s < steps; ++s) {
        V8 lo[K], hi[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t *p = a + idx[k] * (BYTES / 4);
            lo[k] = ld256(p); hi[k] = ld256(p + 8);
            if (BYTES == 128) { V8 x = ld256(p + 16), y = ld256(p + 24); lo[k].v[1] ^= x.v[3]; hi[k].v[2] ^= y.v[5]; }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            uint64_t h = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) h += __popc(lo[k].v[j] & hi[k].v[j]) + lo[k].v[j];
            acc += h;
            idx[k] = mix(idx[k] + h + s) % n_blk;
        }
    }
    if (acc == 0x1234567) out[t] = acc + pad[0];
}

template <int K, int BYTES>
static void run(const uint32_t *a, uint64_t n_blk, int steps, int blocks_per_sm, int n_sm, uint64_t *out, const char *tag) {
    // occupancy control through dynamic shared memory: blocks of 128 threads
    int smem = blocks_per_sm >= 16 ? 0 : (227 * 1024 / blocks_per_sm - 1024) & ~1023;
    cudaFuncSetAttribute(k_gather<K, BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gather<K, BYTES>, 128, smem);
    const int waves = 4;
    const int grid = n_sm * occ * waves;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_gather<K, BYTES><<<grid, 128, smem>>>(a, n_blk, 10, out);
    cudaEventRecord(e0);
    k_gather<K, BYTES><<<grid, 128, smem>>>(a, n_blk, steps, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double g = (double)grid * 128 * K * steps;
    printf("%-10s B=%3d K=%d warps/SM=%2d grid=%6d  %8.2f ms  %7.2f Ggather/s  %7.1f GB/s useful\n", tag, BYTES, K, occ * 4, grid, ms, g / ms / 1e6, g * BYTES / ms / 1e6);
    fflush(stdout);
}

int main(int argc, char **argv) {
    const uint64_t bytes = (argc > 1 ? atoll(argv[1]) : 1024ull) << 20;
    int n_sm = 0;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *a; uint64_t *out;
    cudaMalloc(&a, bytes); cudaMalloc(&out, 1 << 28);
    cudaMemset(a, 0x5a, bytes);
    size_t lim = 0;
    cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
    printf("array %llu MB, %d SMs, default cudaLimitMaxL2FetchGranularity = %zu\n", (unsigned long long)(bytes >> 20), n_sm, lim);
    const int steps = 200;
    for (int gran : {0, 32, 64, 128}) {
        char tag[32];
        if (gran) {
            cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
            cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
            snprintf(tag, sizeof tag, "gran=%zu%s", lim, e == cudaSuccess ? "" : "!");
        } else snprintf(tag, sizeof tag, "default");
        for (int bps : {4, 8, 12, 16}) {
            run<1, 64>(a, bytes / 64, steps, bps, n_sm, out, tag);
            run<2, 64>(a, bytes / 64, steps, bps, n_sm, out, tag);
            run<4, 64>(a, bytes / 64, steps, bps, n_sm, out, tag);
        }
        run<1, 128>(a, bytes / 128, steps, 16, n_sm, out, tag);
        run<2, 128>(a, bytes / 128, steps, 16, n_sm, out, tag);
        run<2, 128>(a, bytes / 128, steps, 8, n_sm, out, tag);
    }
    return 0;
}
