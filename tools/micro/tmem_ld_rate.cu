// Micro-benchmark: tcgen05.ld throughput (32x32b.x32 and .x16) with 4 / 8 / 16 warps of one CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mvlpt_b200/csrc -o tools/micro/tmem_ld_rate.bin tools/micro/tmem_ld_rate.cu
#include <cstdio>
#include "ptx_sm100.cuh"
using namespace mvlpt;

__global__ void __launch_bounds__(1024, 1) k(int nwarps, int iters, int x16, long long* out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot + (uint32_t((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    if (warp < nwarps) {
        for (int i = 0; i < iters; ++i) {
            if (x16) {
                uint32_t a[16], b[16];
                tmem_ld_32x32b_x16(tb + ((i * 32) & 255), a);
                tmem_ld_32x32b_x16(tb + ((i * 32 + 16) & 255), b);
                tmem_ld_wait();
                acc += a[0] + b[15];
            } else {
                uint32_t a[32];
                tmem_ld_32x32(tb + ((i * 32) & 255), a);
                tmem_ld_wait();
                acc += a[0] + a[31];
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
    long long* d; uint32_t* s;
    cudaMalloc(&d, 16); cudaMalloc(&s, 4096);
    for (int x16 = 0; x16 < 2; ++x16)
        for (int nw : {1, 4, 8, 16, 32}) {
            const int iters = 4096;
            for (int rep = 0; rep < 2; ++rep) k<<<1, 1024>>>(nw, iters, x16, d, s);
            long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)nw * iters * 32 * 32 * 4;
            printf("%s warps=%2d: %7.1f cyc per 32-col load per warp, %6.1f B/cyc/SM  (%s)\n", x16 ? "2 x .x16" : "  1 x .x32", nw,
                   (double)h / iters, bytes / h, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
