// UPT shared-prompt projection, forward and backward INCLUDING weight gradients — the only trainable weights of the
// whole step (trainers/mvlpt.py:234-257 modules, :376-414 forward_mvlpt_proj; backward = its autograd).
//
//   x0 = [ ctx . Wc_pre^T + b ; [vpt; vpt_deep] . Wv_pre^T + b ]                 T = n_ctx + n_vpt tokens, width pd
//   x1 = x0 + (LN1(x0) . W_v^T + b_v) . W_o^T + b_o      <- the reference feeds a (1, T, pd) tensor to an (L, N, E)
//                                                            nn.MultiheadAttention: sequence length 1, softmax == 1,
//                                                            so attention is out_proj(v_proj(.)) per token and
//                                                            W_q / W_k receive zero gradient
//   x2 = x1 + quickgelu(LN2(x1) . W_fc^T + b_fc) . W_pr^T + b_pr
//   ctx' = x2[:n_ctx] . Wc_post^T + b ,  vpt' = x2[n_ctx:] . Wv_post^T + b
//
// The problem is tiny (T <= 16 + 24*8 tokens, pd = 128; ~0.1 GFLOP) and latency-bound: everything runs in fp32 on one
// generic strided SIMT GEMM plus three row kernels, all launched back to back on the caller's stream.  Activations
// needed by the backward live in the caller's workspace.
#include "common.cuh"
#include <cuda_fp16.h>

using namespace mvlpt;

namespace {

__device__ __forceinline__ float ld_any(const void* p, size_t i, int f16) {
    return f16 ? __half2float(static_cast<const __half*>(p)[i]) : static_cast<const float*>(p)[i];
}

struct SG {
    const void* A;      // element (m,k) at A[m*sam + k*sak]
    const void* B;      // element (k,n) at B[k*sbk + n*sbn]
    const void* bias;   // [N] or null
    const float* resid; // [M, ldc] or null
    float* C;           // [M, ldc]
    int M, N, K;
    long long sam, sak, sbk, sbn;
    int ldc;
    int a_f16, b_f16, bias_f16;
    int accumulate;     // C += ...
};

// 32x32 output tile per 256-thread block, 2x2 per thread... kept simple: 16x16 threads, 2x2 micro-tile, K step 16.
__global__ void __launch_bounds__(256) small_gemm_kernel(SG g) {
    __shared__ float As[16][33];
    __shared__ float Bs[16][33];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    for (int k0 = 0; k0 < g.K; k0 += 16) {
        for (int i = threadIdx.x; i < 512; i += 256) {
            int kk, mm;
            if (g.sak == 1) { kk = i & 15; mm = i >> 4; } else { mm = i & 31; kk = i >> 5; }
            const int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < g.M && k < g.K) ? ld_any(g.A, (size_t)(m * g.sam + k * g.sak), g.a_f16) : 0.f;
        }
        for (int i = threadIdx.x; i < 512; i += 256) {
            int kk, nn;
            if (g.sbk == 1) { kk = i & 15; nn = i >> 4; } else { nn = i & 31; kk = i >> 5; }
            const int n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < g.N && k < g.K) ? ld_any(g.B, (size_t)(k * g.sbk + n * g.sbn), g.b_f16) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float a0 = As[kk][ty], a1 = As[kk][ty + 16];
            const float b0 = Bs[kk][tx], b1 = Bs[kk][tx + 16];
            acc[0][0] += a0 * b0; acc[0][1] += a0 * b1;
            acc[1][0] += a1 * b0; acc[1][1] += a1 * b1;
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
            if (m < g.M && n < g.N) {
                float v = acc[i][j];
                if (g.bias) v += ld_any(g.bias, n, g.bias_f16);
                const size_t o = (size_t)m * g.ldc + n;
                if (g.resid) v += g.resid[o];
                if (g.accumulate) v += g.C[o];
                g.C[o] = v;
            }
        }
}

// out[n] (+)= sum_m X[m, n]
__global__ void colsum_kernel(const float* __restrict__ X, int M, int N, int ld, float* __restrict__ out, int accumulate) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += X[(size_t)m * ld + n];
    out[n] = accumulate ? out[n] + s : s;
}

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per row: xhat = (x - mean) * rstd, y = xhat * gamma + beta   (clip/model.py:153-159, fp32, eps 1e-5)
__global__ void upt_ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float* __restrict__ xhat, float* __restrict__ rstd,
                                  float* __restrict__ y, int rows, int d, float eps) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* xr = x + (size_t)r * d;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += xr[c];
    const float mu = wsum(s) / d;
    float q = 0.f;
    for (int c = lane; c < d; c += 32) { const float t = xr[c] - mu; q += t * t; }
    const float rs = rsqrtf(wsum(q) / d + eps);
    for (int c = lane; c < d; c += 32) {
        const float h = (xr[c] - mu) * rs;
        xhat[(size_t)r * d + c] = h;
        y[(size_t)r * d + c] = h * gamma[c] + beta[c];
    }
    if (lane == 0) rstd[r] = rs;
}

// dx_out = dx_in + rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma
__global__ void upt_ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ xhat,
                                  const float* __restrict__ rstd, const float* __restrict__ gamma,
                                  const float* __restrict__ dx_in, float* __restrict__ dx_out, int rows, int d) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= rows) return;
    const size_t o = (size_t)r * d;
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < d; c += 32) {
        const float g = dy[o + c] * gamma[c];
        s1 += g;
        s2 += g * xhat[o + c];
    }
    s1 = wsum(s1) / d;
    s2 = wsum(s2) / d;
    const float rs = rstd[r];
    for (int c = lane; c < d; c += 32) {
        const float g = dy[o + c] * gamma[c];
        dx_out[o + c] = dx_in[o + c] + rs * (g - s1 - xhat[o + c] * s2);
    }
}

// dgamma[c] = sum_r dy[r,c] * xhat[r,c] ; dbeta[c] = sum_r dy[r,c]
__global__ void upt_ln_param_grad_kernel(const float* __restrict__ dy, const float* __restrict__ xhat, int rows, int d,
                                         float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    float a = 0.f, b = 0.f;
    for (int r = 0; r < rows; ++r) {
        const float g = dy[(size_t)r * d + c];
        a += g * xhat[(size_t)r * d + c];
        b += g;
    }
    dgamma[c] = a;
    dbeta[c] = b;
}

__global__ void quickgelu_fwd_kernel(const float* __restrict__ t, float* __restrict__ g, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float v = t[i]; g[i] = v / (1.f + __expf(-1.702f * v)); }
}
__global__ void quickgelu_bwd_kernel(const float* __restrict__ t, float* __restrict__ dg_to_dt, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float v = t[i];
        const float s = 1.f / (1.f + __expf(-1.702f * v));
        dg_to_dt[i] *= s * (1.f + 1.702f * v * (1.f - s));
    }
}

struct Ctx {
    cudaStream_t s;
    int rc;
};

void gemm(Ctx& c, const void* A, int a_f16, long long sam, long long sak, const void* B, int b_f16, long long sbk,
          long long sbn, const void* bias, int bias_f16, const float* resid, float* C, int ldc, int M, int N, int K,
          int accumulate) {
    if (c.rc || M <= 0 || N <= 0) return;
    SG g{A, B, bias, resid, C, M, N, K, sam, sak, sbk, sbn, ldc, a_f16, b_f16, bias_f16, accumulate};
    dim3 grid(cdiv(N, 32), cdiv(M, 32));
    small_gemm_kernel<<<grid, 256, 0, c.s>>>(g);
    c.rc = launched("upt small_gemm");
}
// Y[M,N] = X[M,K] . W[N,K]^T + b (+ resid)
void linear(Ctx& c, const void* X, int x_f16, const void* W, const void* b, int w_f16, const float* resid, float* Y, int M,
            int N, int K) {
    gemm(c, X, x_f16, K, 1, W, w_f16, 1, K, b, w_f16, resid, Y, N, M, N, K, 0);
}
// dX[M,K] = dY[M,N] . W[N,K]
void dgrad(Ctx& c, const float* dY, const void* W, int w_f16, float* dX, int M, int N, int K) {
    gemm(c, dY, 0, N, 1, W, w_f16, K, 1, nullptr, 0, nullptr, dX, K, M, K, N, 0);
}
// dW[N,K] (+)= dY[M,N]^T . X[M,K] ; db[N] (+)= colsum(dY)
void wgrad(Ctx& c, const float* dY, const void* X, int x_f16, float* dW, float* db, int M, int N, int K, int accumulate) {
    gemm(c, dY, 0, 1, N, X, x_f16, K, 1, nullptr, 0, nullptr, dW, K, N, K, M, accumulate);
    if (c.rc || M <= 0) return;
    colsum_kernel<<<cdiv(N, 128), 128, 0, c.s>>>(dY, M, N, N, db, accumulate);
    c.rc = launched("upt colsum");
}

struct Ws {  // fp32 workspace carve-up, T tokens
    float *x0, *xh1, *rs1, *h1, *val, *x1, *xh2, *rs2, *h2, *t, *g, *x2;   // forward (saved)
    float *dx2, *dg, *dh, *dx1, *dval, *dx0;                               // backward scratch
};
size_t carve(Ws& w, float* base, int T, int pd) {
    size_t o = 0;
    auto take = [&](size_t n) { float* p = base ? base + o : nullptr; o += (n + 3) & ~size_t(3); return p; };
    const size_t a = (size_t)T * pd, b = (size_t)T * 4 * pd;
    w.x0 = take(a); w.xh1 = take(a); w.rs1 = take(T); w.h1 = take(a); w.val = take(a); w.x1 = take(a);
    w.xh2 = take(a); w.rs2 = take(T); w.h2 = take(a); w.t = take(b); w.g = take(b); w.x2 = take(a);
    w.dx2 = take(a); w.dg = take(b); w.dh = take(a); w.dx1 = take(a); w.dval = take(a); w.dx0 = take(a);
    return o * sizeof(float);
}

int check_desc(const mvlpt_upt_desc* d, const char* who) {
    if (!d) return fail(MVLPT_EINVAL, "%s: null descriptor", who);
    if (d->n_ctx <= 0 || d->v <= 0 || d->n_deep < 0 || d->dt <= 0 || d->dv <= 0 || d->pd <= 0)
        return fail(MVLPT_EINVAL, "%s: sizes must be positive", who);
    return MVLPT_OK;
}

}  // namespace

extern "C" {

size_t mvlpt_upt_workspace(const mvlpt_upt_desc* d) {
    if (!d) return 0;
    Ws w;
    return carve(w, nullptr, d->n_ctx + (1 + d->n_deep) * d->v, d->pd);
}

int mvlpt_upt_fwd(const mvlpt_upt_desc* d, const void* const* P, void* workspace, size_t ws_bytes, void* ctx_out,
                  void* vpt_out, mvlpt_stream_t stream) {
    int rc = check_desc(d, "mvlpt_upt_fwd");
    if (rc) return rc;
    if (!P || !workspace || !ctx_out || !vpt_out) return fail(MVLPT_EINVAL, "mvlpt_upt_fwd: null argument");
    for (int i = 0; i < MVLPT_UPT_NPARAM; ++i)
        if (!P[i] && !(i == MVLPT_UPT_VPT_DEEP && d->n_deep == 0))
            return fail(MVLPT_EINVAL, "mvlpt_upt_fwd: parameter %d is null", i);
    if (ws_bytes < mvlpt_upt_workspace(d)) return fail(MVLPT_EINVAL, "mvlpt_upt_fwd: workspace too small");
    rc = require_sm100();
    if (rc) return rc;
    const int n = d->n_ctx, v = d->v, nd = d->n_deep * d->v, T = n + v + nd, pd = d->pd, dt = d->dt, dv = d->dv;
    const int h = d->param_f16;
    Ws w;
    carve(w, static_cast<float*>(workspace), T, pd);
    Ctx c{static_cast<cudaStream_t>(stream), 0};
    linear(c, P[MVLPT_UPT_CTX], h, P[MVLPT_UPT_COOP_PRE_W], P[MVLPT_UPT_COOP_PRE_B], h, nullptr, w.x0, n, pd, dt);
    linear(c, P[MVLPT_UPT_VPT], h, P[MVLPT_UPT_VPT_PRE_W], P[MVLPT_UPT_VPT_PRE_B], h, nullptr, w.x0 + (size_t)n * pd, v, pd, dv);
    linear(c, P[MVLPT_UPT_VPT_DEEP], h, P[MVLPT_UPT_VPT_PRE_W], P[MVLPT_UPT_VPT_PRE_B], h, nullptr,
           w.x0 + (size_t)(n + v) * pd, nd, pd, dv);
    if (c.rc) return c.rc;
    upt_ln_fwd_kernel<<<cdiv(T, 4), 128, 0, c.s>>>(w.x0, (const float*)P[MVLPT_UPT_LN1_G], (const float*)P[MVLPT_UPT_LN1_B],
                                                  w.xh1, w.rs1, w.h1, T, pd, 1e-5f);
    if ((c.rc = launched("upt ln1"))) return c.rc;
    const float* Wv = static_cast<const float*>(P[MVLPT_UPT_IN_W]) + (size_t)2 * pd * pd;
    const float* bv = static_cast<const float*>(P[MVLPT_UPT_IN_B]) + 2 * pd;
    linear(c, w.h1, 0, Wv, bv, 0, nullptr, w.val, T, pd, pd);
    linear(c, w.val, 0, P[MVLPT_UPT_OUT_W], P[MVLPT_UPT_OUT_B], 0, w.x0, w.x1, T, pd, pd);
    if (c.rc) return c.rc;
    upt_ln_fwd_kernel<<<cdiv(T, 4), 128, 0, c.s>>>(w.x1, (const float*)P[MVLPT_UPT_LN2_G], (const float*)P[MVLPT_UPT_LN2_B],
                                                  w.xh2, w.rs2, w.h2, T, pd, 1e-5f);
    if ((c.rc = launched("upt ln2"))) return c.rc;
    linear(c, w.h2, 0, P[MVLPT_UPT_FC_W], P[MVLPT_UPT_FC_B], 0, nullptr, w.t, T, 4 * pd, pd);
    if (c.rc) return c.rc;
    quickgelu_fwd_kernel<<<cdiv(T * 4 * pd, 256), 256, 0, c.s>>>(w.t, w.g, T * 4 * pd);
    if ((c.rc = launched("upt quickgelu"))) return c.rc;
    linear(c, w.g, 0, P[MVLPT_UPT_PROJ_W], P[MVLPT_UPT_PROJ_B], 0, w.x1, w.x2, T, pd, 4 * pd);
    linear(c, w.x2, 0, P[MVLPT_UPT_COOP_POST_W], P[MVLPT_UPT_COOP_POST_B], h, nullptr, static_cast<float*>(ctx_out), n, dt, pd);
    linear(c, w.x2 + (size_t)n * pd, 0, P[MVLPT_UPT_VPT_POST_W], P[MVLPT_UPT_VPT_POST_B], h, nullptr,
           static_cast<float*>(vpt_out), v + nd, dv, pd);
    return c.rc;
}

int mvlpt_upt_bwd(const mvlpt_upt_desc* d, const void* const* P, void* workspace, size_t ws_bytes, const void* d_ctx_out,
                  const void* d_vpt_out, void* const* G, mvlpt_stream_t stream) {
    int rc = check_desc(d, "mvlpt_upt_bwd");
    if (rc) return rc;
    if (!P || !G || !workspace || !d_ctx_out || !d_vpt_out) return fail(MVLPT_EINVAL, "mvlpt_upt_bwd: null argument");
    for (int i = 0; i < MVLPT_UPT_NPARAM; ++i)
        if ((!P[i] || !G[i]) && !(i == MVLPT_UPT_VPT_DEEP && d->n_deep == 0))
            return fail(MVLPT_EINVAL, "mvlpt_upt_bwd: parameter/gradient %d is null", i);
    if (ws_bytes < mvlpt_upt_workspace(d)) return fail(MVLPT_EINVAL, "mvlpt_upt_bwd: workspace too small");
    rc = require_sm100();
    if (rc) return rc;
    const int n = d->n_ctx, v = d->v, nd = d->n_deep * d->v, T = n + v + nd, pd = d->pd, dt = d->dt, dv = d->dv;
    const int h = d->param_f16;
    Ws w;
    carve(w, static_cast<float*>(workspace), T, pd);
    Ctx c{static_cast<cudaStream_t>(stream), 0};
    auto Gf = [&](int i) { return static_cast<float*>(G[i]); };
    const float* dC = static_cast<const float*>(d_ctx_out);
    const float* dV = static_cast<const float*>(d_vpt_out);
    // post linears
    wgrad(c, dC, w.x2, 0, Gf(MVLPT_UPT_COOP_POST_W), Gf(MVLPT_UPT_COOP_POST_B), n, dt, pd, 0);
    wgrad(c, dV, w.x2 + (size_t)n * pd, 0, Gf(MVLPT_UPT_VPT_POST_W), Gf(MVLPT_UPT_VPT_POST_B), v + nd, dv, pd, 0);
    dgrad(c, dC, P[MVLPT_UPT_COOP_POST_W], h, w.dx2, n, dt, pd);
    dgrad(c, dV, P[MVLPT_UPT_VPT_POST_W], h, w.dx2 + (size_t)n * pd, v + nd, dv, pd);
    // MLP
    wgrad(c, w.dx2, w.g, 0, Gf(MVLPT_UPT_PROJ_W), Gf(MVLPT_UPT_PROJ_B), T, pd, 4 * pd, 0);
    dgrad(c, w.dx2, P[MVLPT_UPT_PROJ_W], 0, w.dg, T, pd, 4 * pd);
    if (c.rc) return c.rc;
    quickgelu_bwd_kernel<<<cdiv(T * 4 * pd, 256), 256, 0, c.s>>>(w.t, w.dg, T * 4 * pd);
    if ((c.rc = launched("upt quickgelu_bwd"))) return c.rc;
    wgrad(c, w.dg, w.h2, 0, Gf(MVLPT_UPT_FC_W), Gf(MVLPT_UPT_FC_B), T, 4 * pd, pd, 0);
    dgrad(c, w.dg, P[MVLPT_UPT_FC_W], 0, w.dh, T, 4 * pd, pd);
    if (c.rc) return c.rc;
    upt_ln_param_grad_kernel<<<cdiv(pd, 128), 128, 0, c.s>>>(w.dh, w.xh2, T, pd, Gf(MVLPT_UPT_LN2_G), Gf(MVLPT_UPT_LN2_B));
    if ((c.rc = launched("upt ln2 param grad"))) return c.rc;
    upt_ln_bwd_kernel<<<cdiv(T, 4), 128, 0, c.s>>>(w.dh, w.xh2, w.rs2, (const float*)P[MVLPT_UPT_LN2_G], w.dx2, w.dx1, T, pd);
    if ((c.rc = launched("upt ln2 bwd"))) return c.rc;
    // attention with sequence length 1: out_proj(v_proj(.))
    wgrad(c, w.dx1, w.val, 0, Gf(MVLPT_UPT_OUT_W), Gf(MVLPT_UPT_OUT_B), T, pd, pd, 0);
    dgrad(c, w.dx1, P[MVLPT_UPT_OUT_W], 0, w.dval, T, pd, pd);
    MVLPT_CUDA_OK(cudaMemsetAsync(G[MVLPT_UPT_IN_W], 0, sizeof(float) * 2 * pd * pd, c.s));  // W_q, W_k: no gradient
    MVLPT_CUDA_OK(cudaMemsetAsync(G[MVLPT_UPT_IN_B], 0, sizeof(float) * 2 * pd, c.s));
    wgrad(c, w.dval, w.h1, 0, Gf(MVLPT_UPT_IN_W) + (size_t)2 * pd * pd, Gf(MVLPT_UPT_IN_B) + 2 * pd, T, pd, pd, 0);
    const float* Wv = static_cast<const float*>(P[MVLPT_UPT_IN_W]) + (size_t)2 * pd * pd;
    dgrad(c, w.dval, Wv, 0, w.dh, T, pd, pd);
    if (c.rc) return c.rc;
    upt_ln_param_grad_kernel<<<cdiv(pd, 128), 128, 0, c.s>>>(w.dh, w.xh1, T, pd, Gf(MVLPT_UPT_LN1_G), Gf(MVLPT_UPT_LN1_B));
    if ((c.rc = launched("upt ln1 param grad"))) return c.rc;
    upt_ln_bwd_kernel<<<cdiv(T, 4), 128, 0, c.s>>>(w.dh, w.xh1, w.rs1, (const float*)P[MVLPT_UPT_LN1_G], w.dx1, w.dx0, T, pd);
    if ((c.rc = launched("upt ln1 bwd"))) return c.rc;
    // pre linears: weight grads (the vpt one sums over the shallow and the deep prompts) and the prompt grads
    const float* dxc = w.dx0;
    const float* dxv = w.dx0 + (size_t)n * pd;
    const float* dxd = w.dx0 + (size_t)(n + v) * pd;
    wgrad(c, dxc, P[MVLPT_UPT_CTX], h, Gf(MVLPT_UPT_COOP_PRE_W), Gf(MVLPT_UPT_COOP_PRE_B), n, pd, dt, 0);
    wgrad(c, dxv, P[MVLPT_UPT_VPT], h, Gf(MVLPT_UPT_VPT_PRE_W), Gf(MVLPT_UPT_VPT_PRE_B), v, pd, dv, 0);
    if (nd) wgrad(c, dxd, P[MVLPT_UPT_VPT_DEEP], h, Gf(MVLPT_UPT_VPT_PRE_W), Gf(MVLPT_UPT_VPT_PRE_B), nd, pd, dv, 1);
    dgrad(c, dxc, P[MVLPT_UPT_COOP_PRE_W], h, Gf(MVLPT_UPT_CTX), n, pd, dt);
    dgrad(c, dxv, P[MVLPT_UPT_VPT_PRE_W], h, Gf(MVLPT_UPT_VPT), v, pd, dv);
    if (nd) dgrad(c, dxd, P[MVLPT_UPT_VPT_PRE_W], h, Gf(MVLPT_UPT_VPT_DEEP), nd, pd, dv);
    return c.rc;
}

int mvlpt_vpt_proj_fwd(const void* emb, const void* W, const void* b, int param_f16, void* out, int rows, int d, int p,
                       mvlpt_stream_t stream) {
    if (!emb || !W || !b || !out) return fail(MVLPT_EINVAL, "mvlpt_vpt_proj_fwd: null argument");
    if (rows <= 0 || d <= 0 || p <= 0) return fail(MVLPT_EINVAL, "mvlpt_vpt_proj_fwd: sizes must be positive");
    int rc = require_sm100();
    if (rc) return rc;
    Ctx c{static_cast<cudaStream_t>(stream), 0};
    linear(c, emb, param_f16, W, b, param_f16, nullptr, static_cast<float*>(out), rows, d, p);
    return c.rc;
}

int mvlpt_vpt_proj_bwd(const void* d_out, const void* emb, const void* W, int param_f16, void* d_emb, void* dW, void* db,
                       int rows, int d, int p, int accumulate, mvlpt_stream_t stream) {
    if (!d_out || !emb || !W || !d_emb || !dW || !db) return fail(MVLPT_EINVAL, "mvlpt_vpt_proj_bwd: null argument");
    if (rows <= 0 || d <= 0 || p <= 0) return fail(MVLPT_EINVAL, "mvlpt_vpt_proj_bwd: sizes must be positive");
    int rc = require_sm100();
    if (rc) return rc;
    Ctx c{static_cast<cudaStream_t>(stream), 0};
    const float* dY = static_cast<const float*>(d_out);
    wgrad(c, dY, emb, param_f16, static_cast<float*>(dW), static_cast<float*>(db), rows, d, p, accumulate);
    dgrad(c, dY, W, param_f16, static_cast<float*>(d_emb), rows, d, p);
    return c.rc;
}

}  // extern "C"
