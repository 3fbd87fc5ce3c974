// CoCoOp branch of the MVLPT prompt learner (trainers/mvlpt.py:260-290 meta_net, :348-374 forward_cocoop, :556-571
// instance-conditioned logits) — the small kernels around the text tower, which itself runs the usual kernels on
// B*C sequences.  Everything here is tiny next to that tower; plain CUDA-core kernels, fp32 arithmetic.
#include "common.cuh"
#include <cuda_fp16.h>

using namespace mvlpt;

namespace {

__device__ __forceinline__ float ldp(const void* p, size_t i, int f16) {
    return f16 ? __half2float(static_cast<const __half*>(p)[i]) : static_cast<const float*>(p)[i];
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// h1[b] = relu(W1 . imf[b] + b1) ; bias[b] = W2 . h1[b] + b2          one block per image, one warp per output
__global__ void metanet_fwd_kernel(const float* __restrict__ imf, const void* W1, const void* b1, const void* W2,
                                   const void* b2, int pf16, float* __restrict__ h1, float* __restrict__ bias, int e, int H,
                                   int dt) {
    extern __shared__ float sh[];  // [e] image feature, then [H] hidden
    float* s_in = sh;
    float* s_h = sh + e;
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < e; i += blockDim.x) s_in[i] = imf[(size_t)b * e + i];
    __syncthreads();
    for (int j = warp; j < H; j += nw) {
        float a = 0.f;
        for (int i = lane; i < e; i += 32) a += ldp(W1, (size_t)j * e + i, pf16) * s_in[i];
        a = warp_sum(a) + ldp(b1, j, pf16);
        a = a > 0.f ? a : 0.f;
        if (lane == 0) {
            s_h[j] = a;
            h1[(size_t)b * H + j] = a;
        }
    }
    __syncthreads();
    for (int k = warp; k < dt; k += nw) {
        float a = 0.f;
        for (int j = lane; j < H; j += 32) a += ldp(W2, (size_t)k * H + j, pf16) * s_h[j];
        a = warp_sum(a);
        if (lane == 0) bias[(size_t)b * dt + k] = a + ldp(b2, k, pf16);
    }
}

// d_h1[b,j] = (h1[b,j] > 0) * sum_k d_bias[b,k] W2[k,j]               one block per image
__global__ void metanet_dh_kernel(const float* __restrict__ d_bias, const float* __restrict__ h1, const void* W2, int pf16,
                                  float* __restrict__ d_h1, int H, int dt) {
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int j = warp; j < H; j += nw) {
        float a = 0.f;
        for (int k = lane; k < dt; k += 32) a += d_bias[(size_t)b * dt + k] * ldp(W2, (size_t)k * H + j, pf16);
        a = warp_sum(a);
        if (lane == 0) d_h1[(size_t)b * H + j] = h1[(size_t)b * H + j] > 0.f ? a : 0.f;
    }
}

// dW[r, c] = sum_b dY[b, r] * X[b, c] ; db[r] = sum_b dY[b, r]         one block per output row r
__global__ void outer_sum_kernel(const float* __restrict__ dY, const float* __restrict__ X, float* __restrict__ dW,
                                 float* __restrict__ db, int B, int R, int Cc) {
    const int r = blockIdx.x;
    for (int c = threadIdx.x; c < Cc; c += blockDim.x) {
        float a = 0.f;
        for (int b = 0; b < B; ++b) a += dY[(size_t)b * R + r] * X[(size_t)b * Cc + c];
        dW[(size_t)r * Cc + c] = a;
    }
    if (threadIdx.x == 0) {
        float a = 0.f;
        for (int b = 0; b < B; ++b) a += dY[(size_t)b * R + r];
        db[r] = a;
    }
}

// d_imf[b, i] (+)= scale * sum_j d_h1[b, j] W1[j, i]                    one block per image
__global__ void metanet_din_kernel(const float* __restrict__ d_h1, const void* W1, int pf16, float* __restrict__ d_imf, int e,
                                   int H, float scale) {
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < e; i += blockDim.x) {
        float a = 0.f;
        for (int j = 0; j < H; ++j) a += d_h1[(size_t)b * H + j] * ldp(W1, (size_t)j * e + i, pf16);
        d_imf[(size_t)b * e + i] += scale * a;
    }
}

// x0[(b,c), t] = (slot[c,t] >= 0 ? ctx[slot] + bias[b] : emb[c,t]) + pos[t]       one warp per row
__global__ void assemble_kernel(const float* __restrict__ emb, const void* ctx, int ctx_f16, const float* __restrict__ bias,
                                const int* __restrict__ slot, const float* __restrict__ pos, float* __restrict__ x0, int B,
                                int C, int Lk, int d) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= (long long)B * C * Lk) return;
    const int t = (int)(row % Lk);
    const int c = (int)((row / Lk) % C);
    const int b = (int)(row / ((long long)Lk * C));
    const int s = slot[c * Lk + t];
    float* dst = x0 + row * d;
    for (int i = lane; i < d; i += 32) {
        const float v = s >= 0 ? ldp(ctx, (size_t)s * d + i, ctx_f16) + bias[(size_t)b * d + i]
                               : emb[((size_t)c * Lk + t) * d + i];
        dst[i] = v + pos[(size_t)t * d + i];
    }
}

// logits[b, c] = s * <img[b], txt[b*C + c]>                             one warp per (b, c)
__global__ void pair_logits_fwd_kernel(const float* __restrict__ img, const float* __restrict__ txt, float s,
                                       float* __restrict__ logits, int ldc, int B, int C, int e) {
    const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= (long long)B * C) return;
    const int b = (int)(p / C), c = (int)(p % C);
    float a = 0.f;
    for (int i = lane; i < e; i += 32) a += img[(size_t)b * e + i] * txt[(size_t)p * e + i];
    a = warp_sum(a);
    if (lane == 0) logits[(size_t)b * ldc + c] = s * a;
}

// d_txt[(b,c), :] = s * dz[b,c] * img[b, :]                              one warp per (b, c)
__global__ void pair_logits_dtxt_kernel(const __half* __restrict__ dz, int ldc, const float* __restrict__ img, float s,
                                        float* __restrict__ d_txt, int B, int C, int e) {
    const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= (long long)B * C) return;
    const int b = (int)(p / C), c = (int)(p % C);
    const float g = s * __half2float(dz[(size_t)b * ldc + c]);
    for (int i = lane; i < e; i += 32) d_txt[(size_t)p * e + i] = g * img[(size_t)b * e + i];
}
// d_img[b, :] = s * sum_c dz[b,c] * txt[(b,c), :]                        one block per image
__global__ void pair_logits_dimg_kernel(const __half* __restrict__ dz, int ldc, const float* __restrict__ txt, float s,
                                        float* __restrict__ d_img, int C, int e) {
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < e; i += blockDim.x) {
        float a = 0.f;
        for (int c = 0; c < C; ++c) a += __half2float(dz[(size_t)b * ldc + c]) * txt[((size_t)b * C + c) * e + i];
        d_img[(size_t)b * e + i] = s * a;
    }
}

// part[b, j, :] = sum_c dx[((b*C + c)*Lk + pos[c, j]), :]                one block per (j, b)
__global__ void ctx_partial_kernel(const __half* __restrict__ dx, const int* __restrict__ ctx_pos, float* __restrict__ part,
                                   int C, int Lk, int n, int d) {
    const int j = blockIdx.x, b = blockIdx.y;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        float a = 0.f;
        for (int c = 0; c < C; ++c)
            a += __half2float(dx[(((size_t)b * C + c) * Lk + ctx_pos[c * n + j]) * d + i]);
        part[((size_t)b * n + j) * d + i] = a;
    }
}
// d_bias[b, :] = inv * sum_j part[b, j, :] ; d_ctx[j, :] = inv * sum_b part[b, j, :]
__global__ void ctx_finish_kernel(const float* __restrict__ part, float* __restrict__ d_ctx, float* __restrict__ d_bias, int B,
                                  int n, int d, float inv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    const int which = blockIdx.y;  // [0, B): d_bias row ; [B, B+n): d_ctx row
    float a = 0.f;
    if (which < B) {
        for (int j = 0; j < n; ++j) a += part[((size_t)which * n + j) * d + i];
        d_bias[(size_t)which * d + i] = inv * a;
    } else {
        const int j = which - B;
        for (int b = 0; b < B; ++b) a += part[((size_t)b * n + j) * d + i];
        d_ctx[(size_t)j * d + i] = inv * a;
    }
}

}  // namespace

extern "C" {

int mvlpt_metanet_fwd(const void* imf, const void* W1, const void* b1, const void* W2, const void* b2, int param_f16,
                      void* h1, void* bias, int B, int e, int H, int dt, mvlpt_stream_t stream) {
    if (!imf || !W1 || !b1 || !W2 || !b2 || !h1 || !bias) return fail(MVLPT_EINVAL, "mvlpt_metanet_fwd: null argument");
    if (B <= 0 || e <= 0 || H <= 0 || dt <= 0) return fail(MVLPT_EINVAL, "mvlpt_metanet_fwd: bad sizes");
    int rc = require_sm100();
    if (rc) return rc;
    metanet_fwd_kernel<<<B, 256, (size_t)(e + H) * 4, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float*>(imf), W1, b1, W2, b2, param_f16, static_cast<float*>(h1), static_cast<float*>(bias), e, H,
        dt);
    return launched("metanet_fwd");
}

int mvlpt_metanet_bwd(const void* d_bias, const void* imf, const void* h1, const void* W1, const void* W2, int param_f16,
                      void* d_h1_ws, void* dW1, void* db1, void* dW2, void* db2, void* d_imf, float d_imf_scale, int B, int e,
                      int H, int dt, mvlpt_stream_t stream) {
    if (!d_bias || !imf || !h1 || !W1 || !W2 || !d_h1_ws || !dW1 || !db1 || !dW2 || !db2 || !d_imf)
        return fail(MVLPT_EINVAL, "mvlpt_metanet_bwd: null argument");
    if (B <= 0 || e <= 0 || H <= 0 || dt <= 0) return fail(MVLPT_EINVAL, "mvlpt_metanet_bwd: bad sizes");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const float* db = static_cast<const float*>(d_bias);
    float* dh = static_cast<float*>(d_h1_ws);
    metanet_dh_kernel<<<B, 256, 0, s>>>(db, static_cast<const float*>(h1), W2, param_f16, dh, H, dt);
    if ((rc = launched("metanet_dh"))) return rc;
    outer_sum_kernel<<<dt, 64, 0, s>>>(db, static_cast<const float*>(h1), static_cast<float*>(dW2), static_cast<float*>(db2),
                                       B, dt, H);
    if ((rc = launched("metanet_dW2"))) return rc;
    outer_sum_kernel<<<H, 256, 0, s>>>(dh, static_cast<const float*>(imf), static_cast<float*>(dW1), static_cast<float*>(db1),
                                       B, H, e);
    if ((rc = launched("metanet_dW1"))) return rc;
    metanet_din_kernel<<<B, 256, 0, s>>>(dh, W1, param_f16, static_cast<float*>(d_imf), e, H, d_imf_scale);
    return launched("metanet_din");
}

int mvlpt_cocoop_assemble(const void* emb, const void* ctx, int ctx_f16, const void* bias, const void* slot, const void* pos,
                          void* x0, int B, int C, int Lk, int d, mvlpt_stream_t stream) {
    if (!emb || !ctx || !bias || !slot || !pos || !x0) return fail(MVLPT_EINVAL, "mvlpt_cocoop_assemble: null argument");
    if (B <= 0 || C <= 0 || Lk <= 0 || d <= 0) return fail(MVLPT_EINVAL, "mvlpt_cocoop_assemble: bad sizes");
    int rc = require_sm100();
    if (rc) return rc;
    const long long rows = (long long)B * C * Lk;
    assemble_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float*>(emb), ctx, ctx_f16, static_cast<const float*>(bias), static_cast<const int*>(slot),
        static_cast<const float*>(pos), static_cast<float*>(x0), B, C, Lk, d);
    return launched("cocoop_assemble");
}

int mvlpt_pair_logits_fwd(const void* img, const void* txt, float scale, void* logits, int ldc, int B, int C, int e,
                          mvlpt_stream_t stream) {
    if (!img || !txt || !logits) return fail(MVLPT_EINVAL, "mvlpt_pair_logits_fwd: null argument");
    if (B <= 0 || C <= 0 || e <= 0 || ldc < C) return fail(MVLPT_EINVAL, "mvlpt_pair_logits_fwd: bad sizes");
    int rc = require_sm100();
    if (rc) return rc;
    const long long pairs = (long long)B * C;
    pair_logits_fwd_kernel<<<(unsigned)((pairs + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float*>(img), static_cast<const float*>(txt), scale, static_cast<float*>(logits), ldc, B, C, e);
    return launched("pair_logits_fwd");
}

int mvlpt_pair_logits_bwd(const void* dz16, int ldc, const void* img, const void* txt, float scale, void* d_txt, void* d_img,
                          int B, int C, int e, mvlpt_stream_t stream) {
    if (!dz16 || !img || !txt || !d_txt || !d_img) return fail(MVLPT_EINVAL, "mvlpt_pair_logits_bwd: null argument");
    if (B <= 0 || C <= 0 || e <= 0 || ldc < C) return fail(MVLPT_EINVAL, "mvlpt_pair_logits_bwd: bad sizes");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long pairs = (long long)B * C;
    pair_logits_dtxt_kernel<<<(unsigned)((pairs + 7) / 8), 256, 0, s>>>(static_cast<const __half*>(dz16), ldc,
                                                                        static_cast<const float*>(img), scale,
                                                                        static_cast<float*>(d_txt), B, C, e);
    if ((rc = launched("pair_logits_dtxt"))) return rc;
    pair_logits_dimg_kernel<<<B, 256, 0, s>>>(static_cast<const __half*>(dz16), ldc, static_cast<const float*>(txt), scale,
                                              static_cast<float*>(d_img), C, e);
    return launched("pair_logits_dimg");
}

int mvlpt_cocoop_ctx_grad(const void* dx16, const void* ctx_pos, void* part_ws, void* d_ctx, void* d_bias, int B, int C, int Lk,
                          int n_ctx, int d, float inv_scale, mvlpt_stream_t stream) {
    if (!dx16 || !ctx_pos || !part_ws || !d_ctx || !d_bias) return fail(MVLPT_EINVAL, "mvlpt_cocoop_ctx_grad: null argument");
    if (B <= 0 || C <= 0 || Lk <= 0 || n_ctx <= 0 || d <= 0) return fail(MVLPT_EINVAL, "mvlpt_cocoop_ctx_grad: bad sizes");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ctx_partial_kernel<<<dim3(n_ctx, B), 256, 0, s>>>(static_cast<const __half*>(dx16), static_cast<const int*>(ctx_pos),
                                                      static_cast<float*>(part_ws), C, Lk, n_ctx, d);
    if ((rc = launched("cocoop_ctx_partial"))) return rc;
    ctx_finish_kernel<<<dim3(cdiv(d, 128), B + n_ctx), 128, 0, s>>>(static_cast<const float*>(part_ws), static_cast<float*>(d_ctx),
                                                                   static_cast<float*>(d_bias), B, n_ctx, d, inv_scale);
    return launched("cocoop_ctx_finish");
}

}  // extern "C"
