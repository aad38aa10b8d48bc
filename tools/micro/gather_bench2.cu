// Micro-benchmark 2: what limits dependent random block gathers on B200 -- sectors, requests, DRAM lines, TLB reach or
// outstanding misses per SM?  Variants of one gather of a BYTES-byte block:
//   V256   BYTES/32 x ld.global.nc.v8.u32 per thread          V128   BYTES/16 x ld.global.nc.v4.u32 per thread
//   COOP2  two adjacent lanes share a chain and load one 32-byte half each (one warp instruction = 16 lines x 2 sectors)
//   COOP4  four lanes, 16 bytes each (one warp instruction = 8 lines x 2 sectors)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bench2 gather_bench2.cu
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
__device__ __forceinline__ uint4 ld128(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 29; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 32; return x; }

enum { V256 = 0, V128, COOP2, COOP4 };

template <int MODE, int BYTES>
__global__ void k_gather(const uint32_t *__restrict__ a, uint64_t n_blk, int steps, uint64_t *out) {
    extern __shared__ uint8_t pad[];
    const uint64_t t0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int share = MODE == COOP2 ? 2 : MODE == COOP4 ? 4 : 1;
    const uint64_t t = t0 / share;
    const int sub = (int)(t0 % share);
    uint64_t idx = mix(t + 1) % n_blk, acc = 0;
    for (int s = 0; s < steps; ++s) {
        const uint32_t *p = a + idx * (BYTES / 4);
        uint64_t h = 0;
        if (MODE == V256) {
#pragma unroll
            for (int j = 0; j < BYTES / 32; ++j) { const V8 x = ld256(p + 8 * j); h += __popc(x.v[0] & x.v[5]) + x.v[3] + x.v[7]; }
        } else if (MODE == V128) {
#pragma unroll
            for (int j = 0; j < BYTES / 16; ++j) { const uint4 x = ld128(p + 4 * j); h += __popc(x.x & x.y) + x.w; }
        } else if (MODE == COOP2) {
            const V8 x = ld256(p + 8 * sub);
            h = __popc(x.v[0] & x.v[5]) + x.v[3] + x.v[7];
            h += __shfl_xor_sync(0xffffffffu, h, 1);
        } else {
            const uint4 x = ld128(p + 4 * sub);
            h = __popc(x.x & x.y) + x.w;
            h += __shfl_xor_sync(0xffffffffu, h, 1);
            h += __shfl_xor_sync(0xffffffffu, h, 2);
        }
        acc += h;
        idx = mix(idx + h + s) % n_blk;
    }
    if (acc == 0x1234567) out[t0] = acc + pad[0];
}

template <int MODE, int BYTES>
static void run(const uint32_t *a, uint64_t n_blk, int steps, int warps_per_sm, int n_sm, uint64_t *out, const char *tag) {
    const int bt = warps_per_sm < 4 ? 32 * warps_per_sm : 128, bps = warps_per_sm < 4 ? 1 : warps_per_sm / 4;
    int smem = bps >= 16 ? 0 : (227 * 1024 / bps - 1024) & ~1023;
    cudaFuncSetAttribute(k_gather<MODE, BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gather<MODE, BYTES>, bt, smem);
    const int grid = n_sm * occ * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_gather<MODE, BYTES><<<grid, bt, smem>>>(a, n_blk, 10, out);
    cudaEventRecord(e0);
    k_gather<MODE, BYTES><<<grid, bt, smem>>>(a, n_blk, steps, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const int share = MODE == COOP2 ? 2 : MODE == COOP4 ? 4 : 1;
    const double g = (double)grid * bt / share * steps;
    static const char *names[] = {"v256", "v128", "coop2", "coop4"};
    printf("%-8s %-5s B=%3d warps/SM=%2d  %8.2f ms  %7.2f Ggather/s  %7.1f GB/s useful  %6.1f Gsector/s\n", tag, names[MODE], BYTES, occ * bt / 32, ms, g / ms / 1e6,
           g * BYTES / ms / 1e6, g * BYTES / 32 / ms / 1e6);
    fflush(stdout);
}

int main(int argc, char **argv) {
    int n_sm = 0;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *a; uint64_t *out;
    const uint64_t max_bytes = 8192ull << 20;
    cudaMalloc(&a, max_bytes); cudaMalloc(&out, 1 << 28);
    cudaMemset(a, 0x5a, max_bytes);
    const int steps = 200;
    for (uint64_t mb : {64ull, 256ull, 1024ull, 8192ull}) {
        char tag[32];
        snprintf(tag, sizeof tag, "%lluMB", (unsigned long long)mb);
        const uint64_t bytes = mb << 20;
        for (int w : {1, 2, 4, 8, 16, 32, 64}) run<V256, 64>(a, bytes / 64, steps, w, n_sm, out, tag);
        for (int w : {16, 64}) {
            run<V256, 32>(a, bytes / 32, steps, w, n_sm, out, tag);
            run<V128, 64>(a, bytes / 64, steps, w, n_sm, out, tag);
            run<COOP2, 64>(a, bytes / 64, steps, w, n_sm, out, tag);
            run<COOP4, 64>(a, bytes / 64, steps, w, n_sm, out, tag);
            run<V256, 128>(a, bytes / 128, steps, w, n_sm, out, tag);
        }
    }
    return 0;
}
