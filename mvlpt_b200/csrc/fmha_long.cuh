// Attention for sequences LONGER than the single-pass tcgen05 kernels hold (L > 272): ViT-L/14@336px has 577 image
// tokens + prompts (configs/trainers/MVLPT/vit_l14_336.yaml; clip/model.py:181-183 for the arithmetic).  Streaming
// (online-softmax) forward and a two-kernel backward (dQ per query tile, dK/dV per key tile; no atomics, deterministic),
// fp32 arithmetic on the CUDA cores with fp16 operands staged in shared memory.  This is the functional path for the one
// shipped config outside the BASELINE shapes, not a tuned one: every BASELINE shape (L <= 265) takes the tcgen05 kernels.
//
// Thread layout of all three kernels: 256 threads = 32 owned rows x 8 lanes; lane g of a row handles the streamed
// rows j = 8*jj + g of each 64-row tile for the dot products and the output columns [8g, 8g+8) for the accumulation.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace mvlpt {

constexpr int kLongRows = 32;  // rows owned by a block (queries in fwd / dQ, keys in dK/dV)
constexpr int kLongTile = 64;  // rows of a streamed tile
constexpr int kLongLd = 66;    // shared-memory row stride in halfs (33 words: rows land on different banks)

// rows [r0, r0+64) x 64 halfs of one head slice -> smem (zero-fill rows >= L)
__device__ __forceinline__ void long_load_tile(__half* dst, const __half* base, size_t row_stride, int r0, int L, int tid) {
    for (int idx = tid; idx < kLongTile * 8; idx += 256) {
        const int r = idx >> 3, c = idx & 7;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r0 + r < L) v = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(r0 + r) * row_stride + c * 8));
        uint32_t* d = reinterpret_cast<uint32_t*>(dst + r * kLongLd + c * 8);
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
}
// one 64-wide row of a head slice -> 64 floats (zeros when !live), times `mul`
__device__ __forceinline__ void long_load_row(float (&x)[64], const __half* row, bool live, float mul) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (live) v = __ldg(reinterpret_cast<const uint4*>(row) + c);
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(h[k]);
            x[c * 8 + 2 * k] = f.x * mul;
            x[c * 8 + 2 * k + 1] = f.y * mul;
        }
    }
}
__device__ __forceinline__ float long_dot(const float (&x)[64], const __half* row) {
    const __half2* h = reinterpret_cast<const __half2*>(row);
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        const float2 f = __half22float2(h[c]);
        a = fmaf(x[2 * c], f.x, a);
        b = fmaf(x[2 * c + 1], f.y, b);
    }
    return a + b;
}
__device__ __forceinline__ float group8_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
}
__device__ __forceinline__ float group8_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v + __shfl_xor_sync(0xffffffffu, v, 4);
}
// acc[0..8) += w * row[8g .. 8g+8)
__device__ __forceinline__ void long_axpy8(float (&acc)[8], float w, const __half* row8) {
    const __half2* h = reinterpret_cast<const __half2*>(row8);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(h[k]);
        acc[2 * k] = fmaf(w, f.x, acc[2 * k]);
        acc[2 * k + 1] = fmaf(w, f.y, acc[2 * k + 1]);
    }
}
__device__ __forceinline__ void long_store8(__half* dst, const float (&acc)[8], float mul) {
    uint4 v;
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = __floats2half2_rn(acc[2 * k] * mul, acc[2 * k + 1] * mul);
    *reinterpret_cast<uint4*>(dst) = v;
}

// ------------------------------------------------------------------------------------------------ forward
// grid (ceil(L/32), heads, N).  out[q] = softmax_j(scale q.k_j [causal]) v_j ; lse[q] = log sum_j exp(scale q.k_j)
__global__ void __launch_bounds__(256)
fmha_long_fwd_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, float* __restrict__ lse, int L, int d,
                     int heads, int causal, float scale) {
    __shared__ __align__(16) __half sK[kLongTile * kLongLd];
    __shared__ __align__(16) __half sV[kLongTile * kLongLd];
    __shared__ float sP[kLongRows][kLongTile + 1];
    const int tid = threadIdx.x, r = tid >> 3, g = tid & 7;
    const int q0 = blockIdx.x * kLongRows, h = blockIdx.y, n = blockIdx.z;
    const size_t rs = (size_t)3 * d;
    const __half* base = qkv + (size_t)n * L * rs + h * 64;
    const int q = q0 + r;
    const bool live = q < L;
    float qv[64];
    long_load_row(qv, base + (size_t)q * rs, live, scale);  // the reference scales q before QK^T
    float m = -INFINITY, l = 0.f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int kend = causal ? min(L, q0 + kLongRows) : L;
    for (int k0 = 0; k0 < kend; k0 += kLongTile) {
        __syncthreads();
        long_load_tile(sK, base + d, rs, k0, L, tid);
        long_load_tile(sV, base + 2 * d, rs, k0, L, tid);
        __syncthreads();
        float s[8], mx = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = jj * 8 + g, kj = k0 + j;
            const float v = long_dot(qv, sK + j * kLongLd);
            const bool ok = live && kj < L && (!causal || kj <= q);
            s[jj] = ok ? v : -INFINITY;
            mx = fmaxf(mx, s[jj]);
        }
        mx = group8_max(mx);
        const float m_new = fmaxf(m, mx);
        const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
        const float corr = (m == -INFINITY) ? 0.f : __expf(m - m_safe);
        float ps = 0.f;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const float p = (s[jj] == -INFINITY) ? 0.f : __expf(s[jj] - m_safe);
            sP[r][jj * 8 + g] = p;
            ps += p;
        }
        ps = group8_sum(ps);
        l = l * corr + ps;
        m = m_new;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] *= corr;
        __syncwarp();  // row r of sP is written and read by the 8 lanes of one warp only
        for (int j = 0; j < kLongTile; ++j) long_axpy8(acc, sP[r][j], sV + j * kLongLd + g * 8);
    }
    if (live) {
        long_store8(out + ((size_t)n * L + q) * d + h * 64 + g * 8, acc, 1.f / l);
        if (g == 0) lse[((size_t)n * heads + h) * L + q] = m + __logf(l);
    }
}

// ------------------------------------------------------------------------------------------------ backward, dQ
// grid (ceil(L/32), heads, N).  P = exp(scale q.k - lse), dP = dO.v, dS = P (dP - D), D = dO.O ; dQ = scale * dS K
__global__ void __launch_bounds__(256)
fmha_long_dq_kernel(const __half* __restrict__ qkv, const __half* __restrict__ o, const __half* __restrict__ d_o,
                    const float* __restrict__ lse, __half* __restrict__ dqkv, int L, int d, int heads, int causal,
                    float scale) {
    __shared__ __align__(16) __half sK[kLongTile * kLongLd];
    __shared__ __align__(16) __half sV[kLongTile * kLongLd];
    __shared__ float sS[kLongRows][kLongTile + 1];
    const int tid = threadIdx.x, r = tid >> 3, g = tid & 7;
    const int q0 = blockIdx.x * kLongRows, h = blockIdx.y, n = blockIdx.z;
    const size_t rs = (size_t)3 * d;
    const __half* base = qkv + (size_t)n * L * rs + h * 64;
    const int q = q0 + r;
    const bool live = q < L;
    float qv[64], dov[64];
    long_load_row(qv, base + (size_t)q * rs, live, scale);
    long_load_row(dov, d_o + ((size_t)n * L + q) * d + h * 64, live, 1.f);
    float D = 0.f;
    if (live) D = long_dot(dov, o + ((size_t)n * L + q) * d + h * 64);
    const float lq = live ? lse[((size_t)n * heads + h) * L + q] : 0.f;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int kend = causal ? min(L, q0 + kLongRows) : L;
    for (int k0 = 0; k0 < kend; k0 += kLongTile) {
        __syncthreads();
        long_load_tile(sK, base + d, rs, k0, L, tid);
        long_load_tile(sV, base + 2 * d, rs, k0, L, tid);
        __syncthreads();
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = jj * 8 + g, kj = k0 + j;
            const bool ok = live && kj < L && (!causal || kj <= q);
            float ds = 0.f;
            if (ok) {
                const float p = __expf(long_dot(qv, sK + j * kLongLd) - lq);
                ds = p * (long_dot(dov, sV + j * kLongLd) - D);
            }
            sS[r][j] = ds;
        }
        __syncwarp();
        for (int j = 0; j < kLongTile; ++j) long_axpy8(acc, sS[r][j], sK + j * kLongLd + g * 8);
    }
    if (live) long_store8(dqkv + ((size_t)n * L + q) * rs + h * 64 + g * 8, acc, scale);
}

// ------------------------------------------------------------------------------------------------ backward, dK and dV
// grid (ceil(L/32), heads, N): the block owns 32 KEY rows and streams the query tiles.
//   dV[k] = sum_q P[q,k] dO[q] ;  dK[k] = scale * sum_q dS[q,k] q
__global__ void __launch_bounds__(256)
fmha_long_dkv_kernel(const __half* __restrict__ qkv, const __half* __restrict__ o, const __half* __restrict__ d_o,
                     const float* __restrict__ lse, __half* __restrict__ dqkv, int L, int d, int heads, int causal,
                     float scale) {
    __shared__ __align__(16) __half sQ[kLongTile * kLongLd];
    __shared__ __align__(16) __half sdO[kLongTile * kLongLd];
    __shared__ float sP[kLongRows][kLongTile + 1], sS[kLongRows][kLongTile + 1];
    __shared__ float sL[kLongTile], sD[kLongTile];
    const int tid = threadIdx.x, r = tid >> 3, g = tid & 7;
    const int k0 = blockIdx.x * kLongRows, h = blockIdx.y, n = blockIdx.z;
    const size_t rs = (size_t)3 * d;
    const __half* base = qkv + (size_t)n * L * rs + h * 64;
    const int kr = k0 + r;
    const bool live = kr < L;
    float kv[64], vv[64];
    long_load_row(kv, base + d + (size_t)kr * rs, live, scale);  // scale folded into k: s = q.(scale k)
    long_load_row(vv, base + 2 * d + (size_t)kr * rs, live, 1.f);
    float accK[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, accV[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int qbeg = causal ? (k0 / kLongTile) * kLongTile : 0;  // queries before the first owned key see none of them
    for (int q0 = qbeg; q0 < L; q0 += kLongTile) {
        __syncthreads();
        long_load_tile(sQ, base, rs, q0, L, tid);
        long_load_tile(sdO, d_o + (size_t)n * L * d + h * 64, (size_t)d, q0, L, tid);
        {
            // D and lse of the 64 streamed queries: 4 lanes per query, 16 columns each
            const int qi = tid >> 2, part = tid & 3, qq = q0 + qi;
            float dsum = 0.f;
            if (qq < L) {
                const size_t off = ((size_t)n * L + qq) * d + h * 64 + part * 16;
                const uint4* a = reinterpret_cast<const uint4*>(d_o + off);
                const uint4* b = reinterpret_cast<const uint4*>(o + off);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint4 x = __ldg(a + c), y = __ldg(b + c);
                    const __half2* hx = reinterpret_cast<const __half2*>(&x);
                    const __half2* hy = reinterpret_cast<const __half2*>(&y);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 fx = __half22float2(hx[k]), fy = __half22float2(hy[k]);
                        dsum = fmaf(fx.x, fy.x, fmaf(fx.y, fy.y, dsum));
                    }
                }
            }
            dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
            dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
            if (part == 0) {
                sD[qi] = dsum;
                sL[qi] = qq < L ? lse[((size_t)n * heads + h) * L + qq] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = jj * 8 + g, qq = q0 + j;
            const bool ok = live && qq < L && (!causal || kr <= qq);
            float p = 0.f, ds = 0.f;
            if (ok) {
                p = __expf(long_dot(kv, sQ + j * kLongLd) - sL[j]);
                ds = p * (long_dot(vv, sdO + j * kLongLd) - sD[j]);
            }
            sP[r][j] = p;
            sS[r][j] = ds;
        }
        __syncwarp();
        for (int j = 0; j < kLongTile; ++j) {
            long_axpy8(accV, sP[r][j], sdO + j * kLongLd + g * 8);
            long_axpy8(accK, sS[r][j], sQ + j * kLongLd + g * 8);
        }
    }
    if (live) {
        long_store8(dqkv + ((size_t)n * L + kr) * rs + d + h * 64 + g * 8, accK, scale);
        long_store8(dqkv + ((size_t)n * L + kr) * rs + 2 * d + h * 64 + g * 8, accV, 1.f);
    }
}

inline int fmha_long_fwd(const void* qkv, void* out, void* lse, int N, int L, int d, int heads, int causal, cudaStream_t s) {
    dim3 grid((L + kLongRows - 1) / kLongRows, heads, N);
    fmha_long_fwd_kernel<<<grid, 256, 0, s>>>(static_cast<const __half*>(qkv), static_cast<__half*>(out),
                                              static_cast<float*>(lse), L, d, heads, causal, 0.125f);
    return launched("fmha_long_fwd");
}

inline int fmha_long_bwd(const void* qkv, const void* o, const void* d_o, const void* lse, void* dqkv, int N, int L, int d,
                         int heads, int causal, cudaStream_t s) {
    dim3 grid((L + kLongRows - 1) / kLongRows, heads, N);
    fmha_long_dq_kernel<<<grid, 256, 0, s>>>(static_cast<const __half*>(qkv), static_cast<const __half*>(o),
                                             static_cast<const __half*>(d_o), static_cast<const float*>(lse),
                                             static_cast<__half*>(dqkv), L, d, heads, causal, 0.125f);
    int rc = launched("fmha_long_dq");
    if (rc) return rc;
    fmha_long_dkv_kernel<<<grid, 256, 0, s>>>(static_cast<const __half*>(qkv), static_cast<const __half*>(o),
                                              static_cast<const __half*>(d_o), static_cast<const float*>(lse),
                                              static_cast<__half*>(dqkv), L, d, heads, causal, 0.125f);
    return launched("fmha_long_dkv");
}

}  // namespace mvlpt
