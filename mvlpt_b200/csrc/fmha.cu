// Multi-head attention core, forward and backward, head width 64, whole key range resident in shared memory
// (L <= 288: the reference's sequences are 50..257 image tokens + prompts, <= 77 text tokens).
//
// Replaces the inside of nn.MultiheadAttention as called at clip/model.py:181-183
// (F.multi_head_attention_forward: scale q by hd^-1/2, QK^T, additive causal mask for the text tower
// clip/model.py:324-330, softmax over keys, PV) and its autograd backward.
// qkv is the packed in_proj output [N*L, 3d] (Q | K | V, head j = columns [64j, 64j+64) of each part).
//
// This revision runs the four contractions on the legacy tensor path (nvcuda::wmma -> HMMA); they are 4.3 % of
// the block FLOPs.  csrc/fmha_sm100.cuh holds the tcgen05 forward used for the image tower.
#include <mma.h>

#include "common.cuh"
#include "fmha_sm100.cuh"
#include "fmha_bwd_sm100.cuh"
#include "fmha_long.cuh"

using namespace nvcuda;
using namespace mvlpt;

namespace {

constexpr int HD = 64;
constexpr int LDH = 72;  // padded smem row stride (halfs) for 64-wide tiles

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// copy rows [r0, r0+rows) x 64 halfs of one head slice into smem (row stride LDH), zero-filling rows >= L
__device__ __forceinline__ void load_head_tile(__half* dst, const __half* src_base, size_t row_stride, int r0, int rows,
                                               int L, int tid, int nthreads) {
    for (int idx = tid; idx < rows * 8; idx += nthreads) {
        const int r = idx >> 3, c = idx & 7;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r0 + r < L) v = *reinterpret_cast<const uint4*>(src_base + (size_t)(r0 + r) * row_stride + c * 8);
        *reinterpret_cast<uint4*>(dst + r * LDH + c * 8) = v;
    }
}

// ------------------------------------------------------------------------------------------ forward
// grid (q_tiles, heads, N); block 128 threads = 4 warps x 16 query rows
__global__ void __launch_bounds__(128) fmha_fwd_wmma_kernel(const __half* __restrict__ qkv, __half* __restrict__ out,
                                                            float* __restrict__ lse, int L, int Lp, int d, int heads,
                                                            int causal, float scale) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int ldS = (Lp + 8 > 72) ? Lp + 8 : 72;
    __half* Qs = reinterpret_cast<__half*>(smem);
    __half* Ks = Qs + 64 * LDH;
    __half* Vs = Ks + Lp * LDH;
    float* S = reinterpret_cast<float*>(Vs + Lp * LDH);

    const int q0 = blockIdx.x * 64, h = blockIdx.y, n = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t rs = (size_t)3 * d;
    const __half* base = qkv + (size_t)n * L * rs + h * HD;
    load_head_tile(Qs, base, rs, q0, 64, L, tid, 128);
    load_head_tile(Ks, base + d, rs, 0, Lp, L, tid, 128);
    load_head_tile(Vs, base + 2 * d, rs, 0, Lp, L, tid, 128);
    __syncthreads();

    const int nct = Lp / 16;
    float* Sw = S + warp * 16 * ldS;
    // S = Q K^T for this warp's 16 rows
    {
        wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::row_major> a[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) wmma::load_matrix_sync(a[kk], Qs + warp * 16 * LDH + kk * 16, LDH);
        const int q_hi = q0 + warp * 16 + 15;
        for (int j = 0; j < nct; ++j) {
            if (causal && j * 16 > q_hi) break;
            wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc;
            wmma::fill_fragment(acc, 0.f);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::col_major> b;
                wmma::load_matrix_sync(b, Ks + j * 16 * LDH + kk * 16, LDH);
                wmma::mma_sync(acc, a[kk], b, acc);
            }
            wmma::store_matrix_sync(Sw + j * 16, acc, ldS, wmma::mem_row_major);
        }
    }
    __syncwarp();
    // softmax, one row at a time; P (fp16, normalised) overwrites the front of the fp32 row
    constexpr int NI = 9;  // 9 * 32 = 288 >= Lp
    for (int rr = 0; rr < 16; ++rr) {
        const int qi = q0 + warp * 16 + rr;
        float* row = Sw + rr * ldS;
        float v[NI];
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int c = lane + 32 * i;
            const bool ok = (c < L) && (!causal || c <= qi);
            v[i] = ok ? row[c] * scale : -INFINITY;
            m = fmaxf(m, v[i]);
        }
        m = warp_max(m);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            v[i] = (v[i] == -INFINITY) ? 0.f : __expf(v[i] - m);
            sum += v[i];
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        __syncwarp();
        __half* prow = reinterpret_cast<__half*>(row);
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int c = lane + 32 * i;
            if (c < Lp) prow[c] = __float2half_rn(v[i] * inv);
        }
        if (lane == 0 && qi < L) lse[((size_t)n * heads + h) * L + qi] = m + __logf(sum);
    }
    __syncwarp();
    // O = P V
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> oacc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) wmma::fill_fragment(oacc[j], 0.f);
    const __half* Pw = reinterpret_cast<const __half*>(Sw);
    int kmax = nct;
    if (causal) {
        const int lim = (q0 + warp * 16 + 15) / 16 + 1;
        kmax = lim < nct ? lim : nct;
    }
    for (int kk = 0; kk < kmax; ++kk) {
        wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::row_major> a;
        wmma::load_matrix_sync(a, Pw + kk * 16, 2 * ldS);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::row_major> b;
            wmma::load_matrix_sync(b, Vs + kk * 16 * LDH + j * 16, LDH);
            wmma::mma_sync(oacc[j], a, b, oacc[j]);
        }
    }
    __syncwarp();
    float* stage = Sw;  // 16 x 64 fp32 (ldS >= 72 keeps this inside the warp's own rows)
#pragma unroll
    for (int j = 0; j < 4; ++j) wmma::store_matrix_sync(stage + j * 16, oacc[j], 64, wmma::mem_row_major);
    __syncwarp();
    {
        const int rr = lane >> 1, c0 = (lane & 1) * 32;
        const int qi = q0 + warp * 16 + rr;
        if (qi < L) {
            __half* op = out + ((size_t)n * L + qi) * d + h * HD + c0;
            const float* sp = stage + rr * 64 + c0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 u;
                __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) h2[j] = __floats2half2_rn(sp[q * 8 + 2 * j], sp[q * 8 + 2 * j + 1]);
                reinterpret_cast<uint4*>(op)[q] = u;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ backward
// grid (heads, N); block 256 threads; loops over 32-row query tiles, dK/dV accumulate in registers.
template <int MAXT>
__global__ void __launch_bounds__(256, 1)
fmha_bwd_wmma_kernel(const __half* __restrict__ qkv, const __half* __restrict__ o, const __half* __restrict__ d_o,
                     const float* __restrict__ lse, __half* __restrict__ dqkv, int L, int Lp, int d, int heads,
                     int causal, float scale) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int ldS = Lp + 8, ldP = Lp + 8;
    __half* Ks = reinterpret_cast<__half*>(smem);
    __half* Vs = Ks + Lp * LDH;
    __half* Qs = Vs + Lp * LDH;
    __half* dOs = Qs + 32 * LDH;
    float* S = reinterpret_cast<float*>(dOs + 32 * LDH);
    float* dP = S + 32 * ldS;
    __half* P16 = reinterpret_cast<__half*>(dP + 32 * ldS);
    __half* dS16 = P16 + 32 * ldP;
    float* Dv = reinterpret_cast<float*>(dS16 + 32 * ldP);
    float* lse_s = Dv + 32;

    const int h = blockIdx.x, n = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t rs = (size_t)3 * d;
    const __half* base = qkv + (size_t)n * L * rs + h * HD;
    const __half* obase = o + (size_t)n * L * d + h * HD;
    const __half* dobase = d_o + (size_t)n * L * d + h * HD;
    __half* dbase = dqkv + (size_t)n * L * rs + h * HD;
    load_head_tile(Ks, base + d, rs, 0, Lp, L, tid, 256);
    load_head_tile(Vs, base + 2 * d, rs, 0, Lp, L, tid, 256);

    const int nct = Lp / 16;
    const int ntile = nct * 4;  // dK / dV output tiles (16x16)
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> dv[MAXT], dk[MAXT];
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
        wmma::fill_fragment(dv[t], 0.f);
        wmma::fill_fragment(dk[t], 0.f);
    }

    for (int q0 = 0; q0 < L; q0 += 32) {
        __syncthreads();  // previous iteration finished with Qs/dOs/S/dP/P16/dS16
        load_head_tile(Qs, base, rs, q0, 32, L, tid, 256);
        load_head_tile(dOs, dobase, (size_t)d, q0, 32, L, tid, 256);
        // D[r] = sum_c dO[r,c] * O[r,c]; 4 rows per warp
        for (int rr = 0; rr < 4; ++rr) {
            const int r = warp * 4 + rr, qi = q0 + r;
            float acc = 0.f;
            if (qi < L) {
                const __half2 a = *reinterpret_cast<const __half2*>(obase + (size_t)qi * d + 2 * lane);
                const __half2 b = *reinterpret_cast<const __half2*>(dobase + (size_t)qi * d + 2 * lane);
                const float2 fa = __half22float2(a), fb = __half22float2(b);
                acc = fa.x * fb.x + fa.y * fb.y;
            }
            acc = warp_sum(acc);
            if (lane == 0) {
                Dv[r] = acc;
                lse_s[r] = (qi < L) ? lse[((size_t)n * heads + h) * L + qi] : 0.f;
            }
        }
        __syncthreads();
        // S = Q K^T, dP = dO V^T   (2 row tiles x nct col tiles each)
        for (int t = warp; t < 2 * nct; t += 8) {
            const int ri = t / nct, cj = t % nct;
            if (causal && cj * 16 > q0 + ri * 16 + 15) continue;  // fully masked tile: never read below
            wmma::fragment<wmma::accumulator, 16, 16, 16, float> s_acc, p_acc;
            wmma::fill_fragment(s_acc, 0.f);
            wmma::fill_fragment(p_acc, 0.f);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::row_major> a;
                wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::col_major> b;
                wmma::load_matrix_sync(a, Qs + ri * 16 * LDH + kk * 16, LDH);
                wmma::load_matrix_sync(b, Ks + cj * 16 * LDH + kk * 16, LDH);
                wmma::mma_sync(s_acc, a, b, s_acc);
                wmma::load_matrix_sync(a, dOs + ri * 16 * LDH + kk * 16, LDH);
                wmma::load_matrix_sync(b, Vs + cj * 16 * LDH + kk * 16, LDH);
                wmma::mma_sync(p_acc, a, b, p_acc);
            }
            wmma::store_matrix_sync(S + ri * 16 * ldS + cj * 16, s_acc, ldS, wmma::mem_row_major);
            wmma::store_matrix_sync(dP + ri * 16 * ldS + cj * 16, p_acc, ldS, wmma::mem_row_major);
        }
        __syncthreads();
        // P = exp(scale*S - lse), dS = scale * P * (dP - D)
        for (int idx = tid; idx < 32 * Lp; idx += 256) {
            const int r = idx / Lp, c = idx - r * Lp;
            const int qi = q0 + r;
            const bool ok = (qi < L) && (c < L) && (!causal || c <= qi);
            float p = 0.f, ds = 0.f;
            if (ok) {
                p = __expf(S[r * ldS + c] * scale - lse_s[r]);
                ds = p * (dP[r * ldS + c] - Dv[r]) * scale;
            }
            P16[r * ldP + c] = __float2half_rn(p);
            dS16[r * ldP + c] = __float2half_rn(ds);
        }
        __syncthreads();
        // dV += P^T dO ; dK += dS^T Q      (tiles owned round-robin by warps)
#pragma unroll
        for (int t = 0; t < MAXT; ++t) {
            const int id = warp + 8 * t;
            if (id < ntile) {
                const int i = id >> 2, j = id & 3;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::col_major> a;
                    wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::row_major> b;
                    wmma::load_matrix_sync(a, P16 + kk * 16 * ldP + i * 16, ldP);
                    wmma::load_matrix_sync(b, dOs + kk * 16 * LDH + j * 16, LDH);
                    wmma::mma_sync(dv[t], a, b, dv[t]);
                    wmma::load_matrix_sync(a, dS16 + kk * 16 * ldP + i * 16, ldP);
                    wmma::load_matrix_sync(b, Qs + kk * 16 * LDH + j * 16, LDH);
                    wmma::mma_sync(dk[t], a, b, dk[t]);
                }
            }
        }
        // dQ = dS K  : 2 x 4 tiles, one per warp
        {
            const int ri = warp >> 2, j = warp & 3;
            wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc;
            wmma::fill_fragment(acc, 0.f);
            int kmax = nct;
            if (causal) {
                const int lim = (q0 + ri * 16 + 15) / 16 + 1;
                kmax = lim < nct ? lim : nct;
            }
            for (int kk = 0; kk < kmax; ++kk) {
                wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::row_major> a;
                wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::row_major> b;
                wmma::load_matrix_sync(a, dS16 + ri * 16 * ldP + kk * 16, ldP);
                wmma::load_matrix_sync(b, Ks + kk * 16 * LDH + j * 16, LDH);
                wmma::mma_sync(acc, a, b, acc);
            }
            float* stage = S + warp * 256;  // S is dead after the elementwise phase
            wmma::store_matrix_sync(stage, acc, 16, wmma::mem_row_major);
            __syncwarp();
            const int rr = lane >> 1, c0 = (lane & 1) * 8;
            const int qi = q0 + ri * 16 + rr;
            if (qi < L) {
                uint4 u;
                __half2* h2 = reinterpret_cast<__half2*>(&u);
                const float* sp = stage + rr * 16 + c0;
#pragma unroll
                for (int k = 0; k < 4; ++k) h2[k] = __floats2half2_rn(sp[2 * k], sp[2 * k + 1]);
                *reinterpret_cast<uint4*>(dbase + (size_t)qi * rs + j * 16 + c0) = u;
            }
        }
    }
    __syncthreads();
    // write dK, dV
    float* stage = S + warp * 256;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
        const int id = warp + 8 * t;
        if (id < ntile) {
            const int i = id >> 2, j = id & 3;
            const int rr = lane >> 1, c0 = (lane & 1) * 8;
            const int ki = i * 16 + rr;
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                wmma::store_matrix_sync(stage, which == 0 ? dk[t] : dv[t], 16, wmma::mem_row_major);
                __syncwarp();
                if (ki < L) {
                    uint4 u;
                    __half2* h2 = reinterpret_cast<__half2*>(&u);
                    const float* sp = stage + rr * 16 + c0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) h2[k] = __floats2half2_rn(sp[2 * k], sp[2 * k + 1]);
                    *reinterpret_cast<uint4*>(dbase + (size_t)ki * rs + (which + 1) * d + j * 16 + c0) = u;
                }
                __syncwarp();
            }
        }
    }
}

}  // namespace

static int g_force_legacy = 0;
// Test hook: 1 routes mvlpt_fmha_fwd / mvlpt_fmha_bwd to the legacy HMMA kernels (kept as an on-GPU cross-check of
// the tcgen05 kernels); returns the previous value.
extern "C" int mvlpt_fmha_force_legacy(int on) {
    const int old = g_force_legacy;
    g_force_legacy = on;
    return old;
}

extern "C" int mvlpt_fmha_fwd(const void* qkv, void* out, void* lse, int N, int L, int d, int heads, int causal,
                              mvlpt_stream_t stream) {
    if (!qkv || !out || !lse) return fail(MVLPT_EINVAL, "mvlpt_fmha_fwd: null argument");
    if (N <= 0 || L <= 0 || heads <= 0 || d != heads * HD)
        return fail(MVLPT_ESHAPE, "mvlpt_fmha_fwd: need d == heads*64 (got d=%d heads=%d)", d, heads);
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // longer than the single-pass kernels hold (ViT-L/14@336px: 577 tokens + prompts): streaming kernels
    if (!fmha_sm100_supported(L) && !(g_force_legacy && L <= 288))
        return fmha_long_fwd(qkv, out, lse, N, L, d, heads, causal, s);
    const float scale = 0.125f;  // 64^-1/2
    if (fmha_sm100_supported(L) && !g_force_legacy) return fmha_fwd_sm100(qkv, out, lse, N, L, d, heads, causal, s);
    const int Lp = (L + 15) / 16 * 16;
    const int ldS = (Lp + 8 > 72) ? Lp + 8 : 72;
    const size_t smem = (size_t)(64 + 2 * Lp) * LDH * 2 + (size_t)64 * ldS * 4;
    static DynSmemCache attr;
    if ((rc = ensure_dyn_smem(fmha_fwd_wmma_kernel, smem, attr))) return rc;
    dim3 grid((L + 63) / 64, heads, N);
    fmha_fwd_wmma_kernel<<<grid, 128, smem, s>>>(static_cast<const __half*>(qkv), static_cast<__half*>(out),
                                                 static_cast<float*>(lse), L, Lp, d, heads, causal, scale);
    return launched("fmha_fwd_wmma");
}

template <int MAXT>
static int launch_bwd(const void* qkv, const void* o, const void* d_o, const void* lse, void* dqkv, int N, int L, int Lp,
                      int d, int heads, int causal, cudaStream_t s) {
    const int ld = Lp + 8;
    const size_t smem = (size_t)(2 * Lp + 64) * LDH * 2 + (size_t)2 * 32 * ld * 4 + (size_t)2 * 32 * ld * 2 + 64 * 4;
    static DynSmemCache attr;
    if (int rc = ensure_dyn_smem(fmha_bwd_wmma_kernel<MAXT>, smem, attr)) return rc;
    dim3 grid(heads, N);
    fmha_bwd_wmma_kernel<MAXT><<<grid, 256, smem, s>>>(
        static_cast<const __half*>(qkv), static_cast<const __half*>(o), static_cast<const __half*>(d_o),
        static_cast<const float*>(lse), static_cast<__half*>(dqkv), L, Lp, d, heads, causal, 0.125f);
    return launched("fmha_bwd_wmma");
}

extern "C" int mvlpt_fmha_bwd(const void* qkv, const void* o, const void* d_o, const void* lse, void* dqkv, int N, int L,
                              int d, int heads, int causal, mvlpt_stream_t stream) {
    if (!qkv || !o || !d_o || !lse || !dqkv) return fail(MVLPT_EINVAL, "mvlpt_fmha_bwd: null argument");
    if (N <= 0 || L <= 0 || heads <= 0 || d != heads * HD)
        return fail(MVLPT_ESHAPE, "mvlpt_fmha_bwd: need d == heads*64 (got d=%d heads=%d)", d, heads);
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (L > 288) return fmha_long_bwd(qkv, o, d_o, lse, dqkv, N, L, d, heads, causal, s);  // see mvlpt_fmha_fwd
    if (fmha_bwd_sm100_supported(L) && !g_force_legacy)
        return fmha_bwd_sm100(qkv, o, d_o, lse, dqkv, N, L, d, heads, causal, s);
    const int Lp = (L + 15) / 16 * 16;
    if (Lp <= 128) return launch_bwd<4>(qkv, o, d_o, lse, dqkv, N, L, Lp, d, heads, causal, s);
    if (Lp <= 224) return launch_bwd<7>(qkv, o, d_o, lse, dqkv, N, L, Lp, d, heads, causal, s);
    return launch_bwd<9>(qkv, o, d_o, lse, dqkv, N, L, Lp, d, heads, causal, s);
}

#ifdef MVLPT_FMHA_DBG
// Debug builds only: copy out and reset the in-kernel timestamp trace of the backward kernel (not part of the ABI).
extern "C" int mvlpt_dbg_fmha_trace(unsigned long long* out, int* counts) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_fmha_dbg, sizeof(unsigned long long) * 2 * 2048);
    cudaMemcpyFromSymbol(counts, g_fmha_dbg_n, sizeof(int) * 2);
    int zero[2] = {0, 0};
    cudaMemcpyToSymbol(g_fmha_dbg_n, zero, sizeof(zero));
    return 0;
}
#endif
