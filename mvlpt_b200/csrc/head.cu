// Logit-head and optimiser kernels: L2 normalisation of the image/text features (trainers/mvlpt.py:550-551),
// per-task logit masking (trainers/mvlpt.py:573-581), softmax cross-entropy with integer or soft targets and its
// gradient (trainers/mvlpt.py:914-916,922,931), small transposes feeding the head dgrad GEMMs, and the
// SGD-with-momentum update Dassl applies to the prompt tensors.
// The dense parts of the head (features @ proj, logit matmul and their dgrads) run on mvlpt_gemm.
#include "common.cuh"
#include <cuda_fp16.h>

using namespace mvlpt;

namespace {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// y = x / ||x||_2 per row (one warp per row)
// y16x3 (optional, [rows, 3e]): the fp16 split v = hi + lo laid out for ONE tensor-core GEMM of K = 3e that evaluates
// hi.hi' + hi.lo' + lo.hi' (the fp32 product up to the lo.lo' term, 2^-22 relative): pattern 0 writes [hi | hi | lo]
// (the A operand), pattern 1 writes [hi | lo | hi] (the W operand).
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, __half* __restrict__ y16, float* __restrict__ y32,
                                  float* __restrict__ inv_norm, __half* __restrict__ y16x3, int pattern, int rows, int e) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* xr = x + (size_t)r * e;
    float s = 0.f;
    for (int c = lane; c < e; c += 32) s += xr[c] * xr[c];
    const float inv = rsqrtf(wsum(s));
    for (int c = lane; c < e; c += 32) {
        const float v = xr[c] * inv;
        const __half hi = __float2half_rn(v);
        y32[(size_t)r * e + c] = v;
        y16[(size_t)r * e + c] = hi;
        if (y16x3) {
            const __half lo = __float2half_rn(v - __half2float(hi));
            __half* o = y16x3 + (size_t)r * 3 * e + c;
            o[0] = hi;
            o[e] = pattern ? lo : hi;
            o[2 * e] = pattern ? hi : lo;
        }
    }
    if (lane == 0) inv_norm[r] = inv;
}

// dx = (dy - y * (y . dy)) * inv_norm
__global__ void l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y32,
                                  const float* __restrict__ inv_norm, __half* __restrict__ dx16, int rows, int e) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* g = dy + (size_t)r * e;
    const float* y = y32 + (size_t)r * e;
    float s = 0.f;
    for (int c = lane; c < e; c += 32) s += g[c] * y[c];
    s = wsum(s);
    const float inv = inv_norm[r];
    for (int c = lane; c < e; c += 32) dx16[(size_t)r * e + c] = __float2half_rn((g[c] - y[c] * s) * inv);
}

// One block per batch row.  Applies the task mask in place (multiply by 0/1, as the reference does), then
// loss_i = -sum_c y_c log softmax(z)_c, dz = coef * (softmax(z) - y) * mask  (fp16, padded columns zeroed).
__global__ void ce_kernel(float* __restrict__ logits, int ldc, const long long* __restrict__ label,
                          const float* __restrict__ soft, const int* __restrict__ task, const int* __restrict__ ranges,
                          float* __restrict__ loss_rows, int* __restrict__ pred, int* __restrict__ hit,
                          __half* __restrict__ dz16, int C, float coef) {
    __shared__ float red[32];
    __shared__ int redi[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    float* z = logits + (size_t)b * ldc;
    int lo = 0, hi = C;
    if (task && ranges) {
        const int t = task[b];
        lo = ranges[2 * t];
        hi = ranges[2 * t + 1];
        for (int c = tid; c < C; c += blockDim.x)
            if (c < lo || c >= hi) z[c] = 0.f;
        __syncthreads();
    }
    // max + argmax (first index of the maximum, like torch.argmax on ties is unspecified; margins are checked upstream)
    float m = -INFINITY;
    int am = 0;
    for (int c = tid; c < C; c += blockDim.x) {
        const float v = z[c];
        if (v > m) { m = v; am = c; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oa = __shfl_xor_sync(0xffffffffu, am, o);
        if (om > m || (om == m && oa < am)) { m = om; am = oa; }
    }
    if (lane == 0) { red[warp] = m; redi[warp] = am; }
    __syncthreads();
    if (warp == 0) {
        float mm = lane < nw ? red[lane] : -INFINITY;
        int aa = lane < nw ? redi[lane] : 0x7fffffff;
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, mm, o);
            const int oa = __shfl_xor_sync(0xffffffffu, aa, o);
            if (om > mm || (om == mm && oa < aa)) { mm = om; aa = oa; }
        }
        if (lane == 0) { red[0] = mm; redi[0] = aa; }
    }
    __syncthreads();
    m = red[0];
    am = redi[0];
    __syncthreads();
    float s = 0.f, ysum = 0.f;
    float ym = -INFINITY;  // arg max of the soft target row = the "label" the reference scores accuracy against
    int ya = 0x7fffffff;   // (trainers/mvlpt.py:935-936)
    for (int c = tid; c < C; c += blockDim.x) {
        s += __expf(z[c] - m);
        if (soft) {
            const float y = soft[(size_t)b * C + c];
            ysum += y;
            if (y > ym) { ym = y; ya = c; }
        }
    }
    s = wsum(s);
    ysum = wsum(ysum);
    if (soft && hit) {
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, ym, o);
            const int oa = __shfl_xor_sync(0xffffffffu, ya, o);
            if (om > ym || (om == ym && oa < ya)) { ym = om; ya = oa; }
        }
        __syncthreads();
        if (lane == 0) { red[warp] = ym; redi[warp] = ya; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < nw; ++w)
                if (red[w] > ym || (red[w] == ym && redi[w] < ya)) { ym = red[w]; ya = redi[w]; }
            hit[b] = (ya == am);
        }
        __syncthreads();
    }
    if (lane == 0) { red[warp] = s; }
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < nw; ++w) tot += red[w];
    __syncthreads();
    if (lane == 0) red[warp] = ysum;
    __syncthreads();
    float ytot = 0.f;
    for (int w = 0; w < nw; ++w) ytot += red[w];
    __syncthreads();
    const float lse = m + __logf(tot);
    const float inv = 1.f / tot;
    const float yinv = soft ? 1.f / ytot : 0.f;
    const int lab = soft ? -1 : (int)label[b];
    float li = 0.f;
    for (int c = tid; c < ldc; c += blockDim.x) {
        float g = 0.f;
        if (c < C) {
            const float zc = z[c];
            const float p = __expf(zc - m) * inv;
            const float y = soft ? soft[(size_t)b * C + c] * yinv : (c == lab ? 1.f : 0.f);
            li -= y * (zc - lse);
            g = coef * (p - y);
            if (c < lo || c >= hi) g = 0.f;
        }
        if (dz16) dz16[(size_t)b * ldc + c] = __float2half_rn(g);
    }
    li = wsum(li);
    if (lane == 0) red[warp] = li;
    __syncthreads();
    if (tid == 0) {
        float t2 = 0.f;
        for (int w = 0; w < nw; ++w) t2 += red[w];
        loss_rows[b] = t2;
        pred[b] = am;
        if (hit && !soft) hit[b] = (am == lab);
    }
}

// out[0] = sum(loss_rows) * inv_div (the batch-mean loss, trainers/mvlpt.py:922/931), out[1] = 100 * mean(hit)
// (dassl compute_accuracy top-1, trainers/mvlpt.py:941).  One block.
__global__ void step_metrics_kernel(const float* __restrict__ loss_rows, const int* __restrict__ hit, int B, float inv_div,
                                    float* __restrict__ out) {
    __shared__ float red[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    float l = 0.f, h = 0.f;
    for (int i = tid; i < B; i += blockDim.x) {
        l += loss_rows[i];
        h += hit ? (float)hit[i] : 0.f;
    }
    l = wsum(l);
    h = wsum(h);
    if (lane == 0) { red[0][warp] = l; red[1][warp] = h; }
    __syncthreads();
    if (tid == 0) {
        float L = 0.f, H = 0.f;
        for (int w = 0; w < nw; ++w) { L += red[0][w]; H += red[1][w]; }
        out[0] = L * inv_div;
        out[1] = 100.f * H / (float)B;
    }
}

// dz16[b,c] = coef * dlogits[b,c] * mask(b,c)   (autograd path: the caller owns the loss)
__global__ void dlogits_kernel(const float* __restrict__ dl, int ld_in, const int* __restrict__ task,
                               const int* __restrict__ ranges, __half* __restrict__ dz16, int ldc, int C, float coef) {
    const int b = blockIdx.x;
    int lo = 0, hi = C;
    if (task && ranges) {
        lo = ranges[2 * task[b]];
        hi = ranges[2 * task[b] + 1];
    }
    for (int c = threadIdx.x; c < ldc; c += blockDim.x) {
        float g = 0.f;
        if (c >= lo && c < hi) g = coef * dl[(size_t)b * ld_in + c];
        dz16[(size_t)b * ldc + c] = __float2half_rn(g);
    }
}

// logits[b,c] *= 1[lo(task_b) <= c < hi(task_b)]   (trainers/mvlpt.py:573-581)
__global__ void task_mask_kernel(float* __restrict__ logits, int ldc, const int* __restrict__ task,
                                 const int* __restrict__ ranges, int C) {
    const int b = blockIdx.x;
    const int lo = ranges[2 * task[b]], hi = ranges[2 * task[b] + 1];
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        if (c < lo || c >= hi) logits[(size_t)b * ldc + c] = 0.f;
}

// out[c, r] = in[r, c]; out columns [R, ld_out) are zero-filled.  32x32 tiles through shared memory.
__global__ void transpose_f16_kernel(const __half* __restrict__ in, __half* __restrict__ out, int R, int Cc, int ld_in,
                                     int ld_out) {
    __shared__ __half tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < Cc) ? in[(size_t)r * ld_in + c] : __float2half(0.f);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < Cc && r < ld_out) out[(size_t)c * ld_out + r] = tile[threadIdx.x][i];
    }
}

// SGD with momentum and L2 weight decay on one tensor; math in fp32, storage in the tensor's own dtype.
template <typename T>
__global__ void sgd_kernel(T* __restrict__ p, T* __restrict__ buf, const float* __restrict__ g, int n, float lr, float mu,
                           float wd, int first) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float w = (float)p[i];
    const float grad = g[i] + wd * w;
    const float b = first ? grad : mu * (float)buf[i] + grad;
    buf[i] = (T)b;
    p[i] = (T)(w - lr * (float)(T)b);
}

}  // namespace

extern "C" {

int mvlpt_l2norm_fwd(const void* x, void* y16, void* y32, void* inv_norm, int rows, int e, mvlpt_stream_t stream) {
    return mvlpt_l2norm_fwd_split(x, y16, y32, inv_norm, nullptr, 0, rows, e, stream);
}

int mvlpt_l2norm_fwd_split(const void* x, void* y16, void* y32, void* inv_norm, void* y16x3, int pattern, int rows, int e,
                           mvlpt_stream_t stream) {
    if (!x || !y16 || !y32 || !inv_norm) return fail(MVLPT_EINVAL, "mvlpt_l2norm_fwd: null argument");
    if (rows <= 0 || e <= 0) return fail(MVLPT_EINVAL, "mvlpt_l2norm_fwd: bad sizes");
    int rc = require_sm100();
    if (rc) return rc;
    l2norm_fwd_kernel<<<cdiv(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float*>(x), static_cast<__half*>(y16), static_cast<float*>(y32),
        static_cast<float*>(inv_norm), static_cast<__half*>(y16x3), pattern, rows, e);
    return launched("l2norm_fwd");
}

int mvlpt_l2norm_bwd(const void* dy, const void* y32, const void* inv_norm, void* dx16, int rows, int e,
                     mvlpt_stream_t stream) {
    if (!dy || !y32 || !inv_norm || !dx16) return fail(MVLPT_EINVAL, "mvlpt_l2norm_bwd: null argument");
    if (rows <= 0 || e <= 0) return fail(MVLPT_EINVAL, "mvlpt_l2norm_bwd: bad sizes");
    int rc = require_sm100();
    if (rc) return rc;
    l2norm_bwd_kernel<<<cdiv(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float*>(dy), static_cast<const float*>(y32), static_cast<const float*>(inv_norm),
        static_cast<__half*>(dx16), rows, e);
    return launched("l2norm_bwd");
}

int mvlpt_ce_fwd_bwd(void* logits, int ldc, const void* label, const void* soft, const void* task, const void* ranges,
                     void* loss_rows, void* pred, void* hit, void* dz16, int B, int C, float coef,
                     mvlpt_stream_t stream) {
    if (!logits || !loss_rows || !pred || (!label && !soft)) return fail(MVLPT_EINVAL, "mvlpt_ce_fwd_bwd: null argument");
    if (B <= 0 || C <= 0 || ldc < C) return fail(MVLPT_EINVAL, "mvlpt_ce_fwd_bwd: bad sizes");
    int rc = require_sm100();
    if (rc) return rc;
    ce_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<float*>(logits), ldc, static_cast<const long long*>(label), static_cast<const float*>(soft),
        static_cast<const int*>(task), static_cast<const int*>(ranges), static_cast<float*>(loss_rows),
        static_cast<int*>(pred), static_cast<int*>(hit), static_cast<__half*>(dz16), C, coef);
    return launched("ce_fwd_bwd");
}

int mvlpt_step_metrics(const void* loss_rows, const void* hit, int B, float inv_div, void* out2, mvlpt_stream_t stream) {
    if (!loss_rows || !out2) return fail(MVLPT_EINVAL, "mvlpt_step_metrics: null argument");
    if (B <= 0) return fail(MVLPT_EINVAL, "mvlpt_step_metrics: B must be positive");
    int rc = require_sm100();
    if (rc) return rc;
    step_metrics_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float*>(loss_rows), static_cast<const int*>(hit), B, inv_div, static_cast<float*>(out2));
    return launched("step_metrics");
}

int mvlpt_dlogits_prepare(const void* dlogits, int ld_in, const void* task, const void* ranges, void* dz16, int ldc,
                          int B, int C, float coef, mvlpt_stream_t stream) {
    if (!dlogits || !dz16) return fail(MVLPT_EINVAL, "mvlpt_dlogits_prepare: null argument");
    if (B <= 0 || C <= 0 || ldc < C || ld_in < C) return fail(MVLPT_EINVAL, "mvlpt_dlogits_prepare: bad sizes");
    int rc = require_sm100();
    if (rc) return rc;
    dlogits_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float*>(dlogits), ld_in, static_cast<const int*>(task), static_cast<const int*>(ranges),
        static_cast<__half*>(dz16), ldc, C, coef);
    return launched("dlogits_prepare");
}

int mvlpt_task_mask(void* logits, int ldc, const void* task, const void* ranges, int B, int C, mvlpt_stream_t stream) {
    if (!logits || !task || !ranges) return fail(MVLPT_EINVAL, "mvlpt_task_mask: null argument");
    if (B <= 0 || C <= 0 || ldc < C) return fail(MVLPT_EINVAL, "mvlpt_task_mask: bad sizes");
    int rc = require_sm100();
    if (rc) return rc;
    task_mask_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<float*>(logits), ldc,
                                                                     static_cast<const int*>(task),
                                                                     static_cast<const int*>(ranges), C);
    return launched("task_mask");
}

int mvlpt_transpose_f16(const void* in, void* out, int R, int Cc, int ld_in, int ld_out, mvlpt_stream_t stream) {
    if (!in || !out) return fail(MVLPT_EINVAL, "mvlpt_transpose_f16: null argument");
    if (R <= 0 || Cc <= 0 || ld_in < Cc || ld_out < R) return fail(MVLPT_EINVAL, "mvlpt_transpose_f16: bad sizes");
    int rc = require_sm100();
    if (rc) return rc;
    dim3 grid(cdiv(Cc, 32), cdiv(ld_out, 32)), block(32, 8);
    transpose_f16_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __half*>(in), static_cast<__half*>(out), R, Cc, ld_in, ld_out);
    return launched("transpose_f16");
}

int mvlpt_sgd(void* p, void* buf, const void* g, int n, int is_f16, float lr, float momentum, float wd, int first_step,
              mvlpt_stream_t stream) {
    if (!p || !buf || !g) return fail(MVLPT_EINVAL, "mvlpt_sgd: null argument");
    if (n <= 0) return fail(MVLPT_EINVAL, "mvlpt_sgd: n must be positive");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (is_f16)
        sgd_kernel<__half><<<cdiv(n, 256), 256, 0, s>>>(static_cast<__half*>(p), static_cast<__half*>(buf),
                                                       static_cast<const float*>(g), n, lr, momentum, wd, first_step);
    else
        sgd_kernel<float><<<cdiv(n, 256), 256, 0, s>>>(static_cast<float*>(p), static_cast<float*>(buf),
                                                      static_cast<const float*>(g), n, lr, momentum, wd, first_step);
    return launched("sgd");
}

}  // extern "C"
