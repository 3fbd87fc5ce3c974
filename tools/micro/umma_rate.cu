// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16) for several N and operand forms, one CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mvlpt_b200/csrc -o gpurun_out/umma_rate tools/micro/umma_rate.cu
#include <cstdio>
#include "ptx_sm100.cuh"
using namespace mvlpt;

// mode 0: SS, A K-major, B K-major; 1: SS, A K-major, B MN-major; 2: SS, A MN-major, B MN-major; 3: TS, B MN-major;
// 4: TS, B K-major
__global__ void __launch_bounds__(128, 1) k(int mode, int N, int iters, int distinct, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 1) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 65536);
        const uint32_t idesc = umma_idesc_f16(128, N, mode == 2, mode == 1 || mode == 2 || mode == 3);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int j = i % distinct;
            const uint32_t acc = tb + (j & 1) * 256;
            if (mode == 0) umma_f16_ss(acc, umma_desc_k_sw128(a0) + 2 * (j & 3), umma_desc_k_sw128(b0) + 2 * (j & 3), idesc, 1);
            else if (mode == 1) umma_f16_ss(acc, umma_desc_k_sw128(a0) + 2 * (j & 3), umma_desc_mn_sw128(b0 + (j & 7) * 2048, 16384), idesc, 1);
            else if (mode == 2) umma_f16_ss(acc, umma_desc_mn_sw128(a0 + (j & 7) * 2048, 16384), umma_desc_mn_sw128(b0 + (j & 7) * 2048, 16384), idesc, 1);
            else if (mode == 3) umma_f16_ts(acc, tb + 128 + (j & 7) * 8, umma_desc_mn_sw128(b0 + (j & 7) * 2048, 16384), idesc, 1);
            else umma_f16_ts(acc, tb + 128 + (j & 7) * 8, umma_desc_k_sw128(b0) + 2 * (j & 3), idesc, 1);
        }
        long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char* names[] = {"SS A:K B:K", "SS A:K B:MN", "SS A:MN B:MN", "TS B:MN", "TS B:K"};
    for (int mode = 0; mode < 5; ++mode)
        for (int N : {16, 32, 64, 128, 256}) {
            if (mode >= 3 && N > 128) continue;  // accumulator at column 0/256, A at 128..
            for (int rep = 0; rep < 2; ++rep) {
                const int iters = 2048;
                k<<<1, 128, 200 * 1024>>>(mode, N, iters, 8, d);
                long long h[2];
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                cudaError_t e = cudaGetLastError();
                if (rep) printf("%-13s N=%3d: issue %6.1f cyc/mma, complete %6.1f cyc/mma  (%s)\n", names[mode], N,
                                (double)h[0] / iters, (double)h[1] / iters, cudaGetErrorString(e));
            }
        }
    return 0;
}
