// HBM-bound row kernels of the hot path: LayerNorm forward/backward (fp32 statistics, clip/model.py:153-159),
// patch gather (im2col of the stride-p conv, clip/model.py:207), token assembly for both towers
// (trainers/mvlpt.py:53-58,416-437,455-510), deep-prompt row replacement (trainers/mvlpt.py:73-82) and the
// batch reductions that turn activation gradients into prompt gradients.
// One warp per row, 128-bit accesses, rows kept in registers (d <= 1024).
#include "common.cuh"
#include <cuda_fp16.h>

using namespace mvlpt;

namespace {

constexpr int kMaxV4 = 8;  // d <= 8 * 128

__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct Row {
    float4 v[kMaxV4];
};

__device__ __forceinline__ void row_load(Row& r, const float* p, int d, int lane) {
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i) {
        const int c = i * 128 + lane * 4;
        r.v[i] = (c < d) ? *reinterpret_cast<const float4*>(p + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
__device__ __forceinline__ void row_load_h(Row& r, const __half* p, int d, int lane) {
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i) {
        const int c = i * 128 + lane * 4;
        if (c < d) {
            const uint2 u = *reinterpret_cast<const uint2*>(p + c);
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
            r.v[i] = make_float4(a.x, a.y, b.x, b.y);
        } else {
            r.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}
__device__ __forceinline__ void row_store(const Row& r, float* p, int d, int lane) {
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i) {
        const int c = i * 128 + lane * 4;
        if (c < d) *reinterpret_cast<float4*>(p + c) = r.v[i];
    }
}
__device__ __forceinline__ void row_store_h(const Row& r, __half* p, int d, int lane) {
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i) {
        const int c = i * 128 + lane * 4;
        if (c < d) {
            uint2 u;
            *reinterpret_cast<__half2*>(&u.x) = __floats2half2_rn(r.v[i].x, r.v[i].y);
            *reinterpret_cast<__half2*>(&u.y) = __floats2half2_rn(r.v[i].z, r.v[i].w);
            *reinterpret_cast<uint2*>(p + c) = u;
        }
    }
}
// mean / rstd of a row held in registers (two-pass, biased variance)
__device__ __forceinline__ void row_stats(const Row& r, int d, int lane, float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i) s += r.v[i].x + r.v[i].y + r.v[i].z + r.v[i].w;  // padding lanes hold 0
    mean = wsum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i) {
        const int c = i * 128 + lane * 4;
        if (c < d) {
            const float a = r.v[i].x - mean, b = r.v[i].y - mean, e = r.v[i].z - mean, f = r.v[i].w - mean;
            q += a * a + b * b + e * e + f * f;
        }
    }
    rstd = rsqrtf(wsum(q) / d + eps);
}
__device__ __forceinline__ void row_affine(Row& r, const float* g, const float* b, int d, int lane, float mean,
                                           float rstd) {
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i) {
        const int c = i * 128 + lane * 4;
        if (c < d) {
            const float4 gg = *reinterpret_cast<const float4*>(g + c);
            const float4 bb = *reinterpret_cast<const float4*>(b + c);
            r.v[i].x = (r.v[i].x - mean) * rstd * gg.x + bb.x;
            r.v[i].y = (r.v[i].y - mean) * rstd * gg.y + bb.y;
            r.v[i].z = (r.v[i].z - mean) * rstd * gg.z + bb.z;
            r.v[i].w = (r.v[i].w - mean) * rstd * gg.w + bb.w;
        }
    }
}

// ---------------------------------------------------------------- LayerNorm forward
// hilo: y is [rows, 2d] = [hi | lo] with value = hi + lo in fp16 pairs (the operand of a K = 2d GEMM against [W | W]: the
// pooled CLS / EOT rows enter the feature projection without the fp16 rounding of their LayerNorm output).
__global__ void ln_fwd_kernel(const float* __restrict__ x, const int* __restrict__ row_index,
                              const float* __restrict__ gamma, const float* __restrict__ beta, __half* __restrict__ y,
                              int rows, int d, float eps, int hilo) {
    pdl_sync();
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const size_t src = row_index ? (size_t)row_index[r] : (size_t)r;
    Row row;
    row_load(row, x + src * d, d, lane);
    float mean, rstd;
    row_stats(row, d, lane, eps, mean, rstd);
    row_affine(row, gamma, beta, d, lane, mean, rstd);
    if (!hilo) {
        row_store_h(row, y + (size_t)r * d, d, lane);
        return;
    }
    __half* yr = y + (size_t)r * 2 * d;
    row_store_h(row, yr, d, lane);
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i) {  // lo = value - fp16(value)
        row.v[i].x -= __half2float(__float2half_rn(row.v[i].x));
        row.v[i].y -= __half2float(__float2half_rn(row.v[i].y));
        row.v[i].z -= __half2float(__float2half_rn(row.v[i].z));
        row.v[i].w -= __half2float(__float2half_rn(row.v[i].w));
    }
    row_store_h(row, yr + d, d, lane);
}

// ---------------------------------------------------------------- LayerNorm backward (gamma/beta frozen)
// g = dy*gamma ; dx = rstd * (g - mean(g) - xhat*mean(g*xhat))
__global__ void ln_bwd_kernel(const __half* __restrict__ dy, const float* __restrict__ x,
                              const int* __restrict__ row_index, const float* __restrict__ gamma,
                              float* __restrict__ dx_stream, __half* __restrict__ dx16, int rows, int d, float eps,
                              int accumulate) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const size_t dst = row_index ? (size_t)row_index[r] : (size_t)r;
    Row xr, g;
    row_load(xr, x + dst * d, d, lane);
    float mean, rstd;
    row_stats(xr, d, lane, eps, mean, rstd);
    row_load_h(g, dy + (size_t)r * d, d, lane);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i) {
        const int c = i * 128 + lane * 4;
        if (c < d) {
            const float4 gg = *reinterpret_cast<const float4*>(gamma + c);
            g.v[i].x *= gg.x; g.v[i].y *= gg.y; g.v[i].z *= gg.z; g.v[i].w *= gg.w;
            xr.v[i].x = (xr.v[i].x - mean) * rstd; xr.v[i].y = (xr.v[i].y - mean) * rstd;
            xr.v[i].z = (xr.v[i].z - mean) * rstd; xr.v[i].w = (xr.v[i].w - mean) * rstd;
            s1 += g.v[i].x + g.v[i].y + g.v[i].z + g.v[i].w;
            s2 += g.v[i].x * xr.v[i].x + g.v[i].y * xr.v[i].y + g.v[i].z * xr.v[i].z + g.v[i].w * xr.v[i].w;
        }
    }
    s1 = wsum(s1) / d;
    s2 = wsum(s2) / d;
    Row out;
    if (accumulate) {
        if (dx_stream) row_load(out, dx_stream + dst * d, d, lane);
        else row_load_h(out, dx16 + dst * d, d, lane);  // fp16 gradient stream: the running sum lives in dx16
    }
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i) {
        const float4 a = g.v[i], h = xr.v[i];
        float4 o = make_float4(rstd * (a.x - s1 - h.x * s2), rstd * (a.y - s1 - h.y * s2), rstd * (a.z - s1 - h.z * s2),
                               rstd * (a.w - s1 - h.w * s2));
        if (accumulate) { o.x += out.v[i].x; o.y += out.v[i].y; o.z += out.v[i].z; o.w += out.v[i].w; }
        out.v[i] = o;
    }
    if (dx_stream) row_store(out, dx_stream + dst * d, d, lane);
    if (dx16) row_store_h(out, dx16 + dst * d, d, lane);
}

// Fast path of the above for d == NV*128 without a row gather: every load of the row (x, dy and the running gradient)
// is issued before the first use, so one warp keeps up to 10*d bytes in flight instead of serialising three round
// trips to HBM behind the two warp reductions; statistics from sum / sum of squares in one pass over registers.
template <int NV, bool STREAM_F32>
__global__ void __launch_bounds__(256)
ln_bwd_fast_kernel(const __half* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                   float* __restrict__ dx_stream, __half* __restrict__ dx16, int rows, float eps, int accumulate) {
    pdl_sync();
    constexpr int d = NV * 128;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const size_t base = (size_t)r * d + lane * 4;
    float4 xv[NV], run[NV];
    uint2 gy[NV], run16[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) xv[i] = *reinterpret_cast<const float4*>(x + base + i * 128);
#pragma unroll
    for (int i = 0; i < NV; ++i) gy[i] = *reinterpret_cast<const uint2*>(dy + base + i * 128);
    if (accumulate) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (STREAM_F32) run[i] = *reinterpret_cast<const float4*>(dx_stream + base + i * 128);
            else run16[i] = *reinterpret_cast<const uint2*>(dx16 + base + i * 128);
        }
    }
    // statistics of x (fp32, biased variance around the mean)
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += xv[i].x + xv[i].y + xv[i].z + xv[i].w;
    const float mean = wsum(s) * (1.f / d);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
        q += xv[i].x * xv[i].x + xv[i].y * xv[i].y + xv[i].z * xv[i].z + xv[i].w * xv[i].w;
    }
    const float rstd = rsqrtf(wsum(q) * (1.f / d) + eps);
    float4 g[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 gg = *reinterpret_cast<const float4*>(gamma + lane * 4 + i * 128);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&gy[i].x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&gy[i].y));
        g[i] = make_float4(a.x * gg.x, a.y * gg.y, b.x * gg.z, b.y * gg.w);
        xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;  // xhat
        s1 += g[i].x + g[i].y + g[i].z + g[i].w;
        s2 += g[i].x * xv[i].x + g[i].y * xv[i].y + g[i].z * xv[i].z + g[i].w * xv[i].w;
    }
    s1 = wsum(s1) * (1.f / d);
    s2 = wsum(s2) * (1.f / d);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float4 o = make_float4(rstd * (g[i].x - s1 - xv[i].x * s2), rstd * (g[i].y - s1 - xv[i].y * s2),
                               rstd * (g[i].z - s1 - xv[i].z * s2), rstd * (g[i].w - s1 - xv[i].w * s2));
        if (accumulate) {
            if (STREAM_F32) {
                o.x += run[i].x; o.y += run[i].y; o.z += run[i].z; o.w += run[i].w;
            } else {
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&run16[i].x));
                const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&run16[i].y));
                o.x += a.x; o.y += a.y; o.z += b.x; o.w += b.y;
            }
        }
        if (STREAM_F32) *reinterpret_cast<float4*>(dx_stream + base + i * 128) = o;
        if (dx16) {
            uint2 u;
            *reinterpret_cast<__half2*>(&u.x) = __floats2half2_rn(o.x, o.y);
            *reinterpret_cast<__half2*>(&u.y) = __floats2half2_rn(o.z, o.w);
            *reinterpret_cast<uint2*>(dx16 + base + i * 128) = u;
        }
    }
}

template <int NV>
static void launch_ln_bwd_fast(const void* dy, const void* x, const void* gamma, void* dx_stream, void* dx16, int rows,
                               float eps, int accumulate, cudaStream_t s) {
    if (dx_stream)
        launch_pdl(ln_bwd_fast_kernel<NV, true>, dim3(cdiv(rows, 8)), dim3(256), 0, s, 1, static_cast<const __half*>(dy),
                   static_cast<const float*>(x), static_cast<const float*>(gamma), static_cast<float*>(dx_stream),
                   static_cast<__half*>(dx16), rows, eps, accumulate);
    else
        launch_pdl(ln_bwd_fast_kernel<NV, false>, dim3(cdiv(rows, 8)), dim3(256), 0, s, 1, static_cast<const __half*>(dy),
                   static_cast<const float*>(x), static_cast<const float*>(gamma), static_cast<float*>(nullptr),
                   static_cast<__half*>(dx16), rows, eps, accumulate);
}

// ---------------------------------------------------------------- im2col for the patch-embedding conv
// patches[(b*g + py)*g + px, c*p*p + ky*p + kx] = img[b, c, py*p+ky, px*p+kx]; columns [3pp, Kp) are zero.
template <typename T>
__global__ void im2col_kernel(const T* __restrict__ img, __half* __restrict__ patches, int B, int H, int W, int p, int Kp) {
    const int g = W / p, gh = H / p;
    const int patch = blockIdx.x;  // b*gh*g + py*g + px
    const int b = patch / (gh * g), rem = patch % (gh * g), py = rem / g, px = rem % g;
    __half* dst = patches + (size_t)patch * Kp;
    const int K = 3 * p * p;
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
        float v = 0.f;
        if (k < K) {
            const int c = k / (p * p), r2 = k % (p * p), ky = r2 / p, kx = r2 % p;
            v = (float)img[(((size_t)b * 3 + c) * H + py * p + ky) * W + px * p + kx];
        }
        dst[k] = __float2half_rn(v);
    }
}

// fp16 images with p % 8 == 0: one thread moves 8 consecutive pixels of a patch row (16 bytes in, 16 bytes out).
// grid.x = patches / 4, block = 4 patches x (3 * p * p / 8) threads padded to a multiple of 32.
__global__ void im2col_vec8_kernel(const __half* __restrict__ img, __half* __restrict__ patches, int n_patches, int H, int W,
                                   int p, int Kp, int per_patch) {
    const int gw = W / p, gh = H / p;
    const int local = threadIdx.x / per_patch, t = threadIdx.x % per_patch;
    const int patch = blockIdx.x * 4 + local;
    if (local >= 4 || patch >= n_patches) return;
    const int b = patch / (gh * gw), rem = patch % (gh * gw), py = rem / gw, px = rem % gw;
    const int K8 = 3 * p * p / 8;  // 16-byte units that carry pixels; [K8, Kp/8) are zero padding
    __half* dst = patches + (size_t)patch * Kp;
    for (int u = t; u < Kp / 8; u += per_patch) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (u < K8) {
            const int k = u * 8, c = k / (p * p), r2 = k % (p * p), ky = r2 / p, kx = r2 % p;
            v = *reinterpret_cast<const uint4*>(img + (((size_t)b * 3 + c) * H + py * p + ky) * W + px * p + kx);
        }
        *reinterpret_cast<uint4*>(dst + u * 8) = v;
    }
}

// ---------------------------------------------------------------- image token assembly
// x0[b,0] = LNpre(cls + pos[0]); x0[b,1..v] = prompt; x0[b,1+v+i] = LNpre(pe[b,i] + pos[1+i])
__global__ void embed_assemble_kernel(const __half* __restrict__ pe, const float* __restrict__ cls,
                                      const float* __restrict__ pos, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, const void* __restrict__ prompt, int prompt_f16,
                                      float* __restrict__ x0, int B, int G, int v, int d, float eps) {
    const int L = 1 + v + G;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= B * L) return;
    const int b = r / L, t = r % L;
    Row row;
    if (t >= 1 && t <= v) {
        if (prompt_f16) row_load_h(row, static_cast<const __half*>(prompt) + (size_t)(t - 1) * d, d, lane);
        else row_load(row, static_cast<const float*>(prompt) + (size_t)(t - 1) * d, d, lane);
    } else {
        Row p;
        const int pi = (t == 0) ? 0 : t - v;
        if (t == 0) row_load(row, cls, d, lane);
        else row_load_h(row, pe + ((size_t)b * G + (t - 1 - v)) * d, d, lane);
        row_load(p, pos + (size_t)pi * d, d, lane);
#pragma unroll
        for (int i = 0; i < kMaxV4; ++i) {
            row.v[i].x += p.v[i].x; row.v[i].y += p.v[i].y; row.v[i].z += p.v[i].z; row.v[i].w += p.v[i].w;
        }
        float mean, rstd;
        row_stats(row, d, lane, eps, mean, rstd);
        row_affine(row, gamma, beta, d, lane, mean, rstd);
    }
    row_store(row, x0 + (size_t)r * d, d, lane);
}

// vpt_dropout (trainers/mvlpt.py:165, applied :76 and :425 AFTER the expansion over the batch, so every (b, j, c) element
// has its own Bernoulli draw).  Counter-based: the keep bits of the 4 columns c..c+3 of row (slab, b, j) are the four
// 16-bit lanes of one 64-bit mix of (seed, element-group index); keep <=> lane >= thr, thr = round(p * 65536).  The same
// function is evaluated by the forward (set_prompt_rows), the backward (prompt_grad) and mvlpt_dropout_keep (tests).
__device__ __forceinline__ unsigned long long drop_bits(unsigned long long seed, int slab, int b, int j, int c4, int B,
                                                        int v, int d4) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (((((unsigned long long)slab * B + b) * v + j) * d4 + c4) + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float4 drop_scale4(unsigned long long bits, unsigned thr, float keep_scale) {
    float4 m;
    m.x = ((unsigned)(bits) & 0xFFFFu) >= thr ? keep_scale : 0.f;
    m.y = ((unsigned)(bits >> 16) & 0xFFFFu) >= thr ? keep_scale : 0.f;
    m.z = ((unsigned)(bits >> 32) & 0xFFFFu) >= thr ? keep_scale : 0.f;
    m.w = ((unsigned)(bits >> 48) & 0xFFFFu) >= thr ? keep_scale : 0.f;
    return m;
}

// x[b, 1+j] = prompt[j]  (deep prompt replacement before block l >= 1), times the dropout keep mask / (1-p) when thr > 0
// xt / record of a row held in registers, for the LayerNorm carried through the linears (gemm_sm100.cuh): xt = (x - mean) *
// gamma as fp16; record = first (sum, sum of squares) pair (0, M2), the other pairs 0, centring value = mean.
__device__ __forceinline__ void row_ln_carry(Row& row, const float* __restrict__ gamma, __half* __restrict__ xt_row,
                                             float* __restrict__ rec_row, int d, int lane) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i) s += row.v[i].x + row.v[i].y + row.v[i].z + row.v[i].w;
    const float mean = wsum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i) {
        const int c = i * 128 + lane * 4;
        if (c < d) {
            const float4 gg = *reinterpret_cast<const float4*>(gamma + c);
            const float a = row.v[i].x - mean, b = row.v[i].y - mean, e = row.v[i].z - mean, f = row.v[i].w - mean;
            q += a * a + b * b + e * e + f * f;
            row.v[i] = make_float4(a * gg.x, b * gg.y, e * gg.z, f * gg.w);
        }
    }
    q = wsum(q);
    row_store_h(row, xt_row, d, lane);
    if (lane < 5) {  // 20 floats: pairs 0..7, then c at index 16
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane == 0) o.y = q;
        if (lane == 4) o.x = mean;
        reinterpret_cast<float4*>(rec_row)[lane] = o;
    }
}

// Starts the chain: xt / records of plain fp32 rows.
__global__ void ln_prep_kernel(const float* __restrict__ x, const float* __restrict__ gamma, __half* __restrict__ xt,
                               float* __restrict__ rec, int rows, int d) {
    pdl_sync();
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    Row row;
    row_load(row, x + (size_t)r * d, d, lane);
    row_ln_carry(row, gamma, xt + (size_t)r * d, rec + (size_t)r * 20, d, lane);
}

// With xt != nullptr the xt / records of the new rows are written as well: the previous block's FC2 produced them for
// the rows this kernel replaces.
__global__ void set_prompt_rows_kernel(float* __restrict__ x, const void* __restrict__ prompt, int prompt_f16, int B,
                                       int L, int v, int d, unsigned thr, float keep_scale, unsigned long long seed,
                                       int slab, __half* __restrict__ xt, float* __restrict__ rec,
                                       const float* __restrict__ gamma) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= B * v) return;
    const int b = r / v, j = r % v;
    Row row;
    if (prompt_f16) row_load_h(row, static_cast<const __half*>(prompt) + (size_t)j * d, d, lane);
    else row_load(row, static_cast<const float*>(prompt) + (size_t)j * d, d, lane);
    if (thr) {
#pragma unroll
        for (int i = 0; i < kMaxV4; ++i) {
            const int c4 = lane + 32 * i;
            if (c4 * 4 < d) {
                const float4 m = drop_scale4(drop_bits(seed, slab, b, j, c4, B, v, d >> 2), thr, keep_scale);
                row.v[i].x *= m.x; row.v[i].y *= m.y; row.v[i].z *= m.z; row.v[i].w *= m.w;
            }
        }
    }
    const size_t ro = (size_t)b * L + 1 + j;
    row_store(row, x + ro * d, d, lane);
    if (xt) row_ln_carry(row, gamma, xt + ro * d, rec + ro * 20, d, lane);
}

// keep[b, j, c] in {0,1}: the mask the two kernels above/below apply (for tests and for replaying a step elsewhere)
__global__ void dropout_keep_kernel(unsigned char* __restrict__ keep, int B, int v, int d, unsigned thr,
                                    unsigned long long seed, int slab) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int d4 = d >> 2;
    if (i >= B * v * d4) return;
    const int c4 = i % d4, j = (i / d4) % v, b = i / (d4 * v);
    const unsigned long long bits = drop_bits(seed, slab, b, j, c4, B, v, d4);
    uchar4 o;
    o.x = ((unsigned)(bits) & 0xFFFFu) >= thr;
    o.y = ((unsigned)(bits >> 16) & 0xFFFFu) >= thr;
    o.z = ((unsigned)(bits >> 32) & 0xFFFFu) >= thr;
    o.w = ((unsigned)(bits >> 48) & 0xFFFFu) >= thr;
    reinterpret_cast<uchar4*>(keep)[i] = o;
}

// grad[j, :] = inv_scale * sum_b dx[b, 1+j, :]; optionally zero those rows (they do not flow further back)
// Block = (prompt row j, 64 columns): 16 column groups of 4 x 16 batch lanes; every batch lane walks b = lane, lane+16, ..
// and the 16 partial sums are added in a fixed order through shared memory (deterministic, no atomics).
__global__ void __launch_bounds__(256)
prompt_grad_kernel(float* __restrict__ dx, __half* __restrict__ dx16, float* __restrict__ grad, int B, int L, int v, int d,
                   float inv_scale, int zero_rows, unsigned thr, float keep_scale, unsigned long long seed, int slab) {
    __shared__ float4 part[16][17];
    const int j = blockIdx.y;
    const int cg = threadIdx.x & 15, bl = threadIdx.x >> 4;
    const int c = (blockIdx.x * 16 + cg) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < d) {
        // 8 rows per round: all loads first (the stores that clear the rows would otherwise fence every next load)
        for (int b0 = bl; b0 < B; b0 += 16 * 8) {
            float4 t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int b = b0 + 16 * k;
                t[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (b < B) {
                    const size_t off = ((size_t)b * L + 1 + j) * d + c;
                    if (dx) {
                        t[k] = *reinterpret_cast<const float4*>(dx + off);
                    } else {  // fp16 gradient stream
                        const uint2 u = *reinterpret_cast<const uint2*>(dx16 + off);
                        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
                        const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
                        t[k] = make_float4(lo.x, lo.y, hi.x, hi.y);
                    }
                    if (thr) {  // autograd of the dropout: the same keep mask / (1-p) the forward applied
                        const float4 m = drop_scale4(drop_bits(seed, slab, b, j, c >> 2, B, v, d >> 2), thr, keep_scale);
                        t[k].x *= m.x; t[k].y *= m.y; t[k].z *= m.z; t[k].w *= m.w;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int b = b0 + 16 * k;
                acc.x += t[k].x; acc.y += t[k].y; acc.z += t[k].z; acc.w += t[k].w;
                if (zero_rows && b < B) {
                    const size_t off = ((size_t)b * L + 1 + j) * d + c;
                    if (dx) *reinterpret_cast<float4*>(dx + off) = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (dx16) *reinterpret_cast<uint2*>(dx16 + off) = make_uint2(0u, 0u);
                }
            }
        }
    }
    part[bl][cg] = acc;
    __syncthreads();
    if (bl == 0 && c < d) {
        float4 s = part[0][cg];
#pragma unroll
        for (int k = 1; k < 16; ++k) {
            const float4 t = part[k][cg];
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        *reinterpret_cast<float4*>(grad + (size_t)j * d + c) =
            make_float4(s.x * inv_scale, s.y * inv_scale, s.z * inv_scale, s.w * inv_scale);
    }
}

// ---------------------------------------------------------------- text token assembly
// x0[c,t] = (slot[c,t] >= 0 ? ctx[(csc ? c*n : 0) + slot[c,t]] : emb[c,t]) + pos[t]
__global__ void text_assemble_kernel(const float* __restrict__ emb, const void* __restrict__ ctx, int ctx_f16,
                                     const int* __restrict__ slot, const float* __restrict__ pos, float* __restrict__ x0,
                                     int C, int Lt, int n_ctx, int d, int csc) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= C * Lt) return;
    const int c = r / Lt, t = r % Lt;
    const int s = (ctx != nullptr) ? slot[r] : -1;
    Row row, p;
    if (s >= 0) {
        const size_t off = ((size_t)(csc ? c * n_ctx : 0) + s) * d;
        if (ctx_f16) row_load_h(row, static_cast<const __half*>(ctx) + off, d, lane);
        else row_load(row, static_cast<const float*>(ctx) + off, d, lane);
    } else {
        row_load(row, emb + (size_t)r * d, d, lane);
    }
    row_load(p, pos + (size_t)t * d, d, lane);
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i) {
        row.v[i].x += p.v[i].x; row.v[i].y += p.v[i].y; row.v[i].z += p.v[i].z; row.v[i].w += p.v[i].w;
    }
    row_store(row, x0 + (size_t)r * d, d, lane);
}

// grad_ctx[j] = inv_scale * sum_c dx0[c, ctx_pos[c,j]]   (shared context)   or per class when csc
__device__ __forceinline__ float4 load4_any(const void* base, size_t off, int f16) {
    if (!f16) return *reinterpret_cast<const float4*>(static_cast<const float*>(base) + off);
    const uint2 u = *reinterpret_cast<const uint2*>(static_cast<const __half*>(base) + off);
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

__global__ void ctx_grad_kernel(const void* __restrict__ dx0, int dx_f16, const int* __restrict__ ctx_pos,
                                float* __restrict__ grad, int C, int Lt, int n_ctx, int d, int csc, float inv_scale) {
    const int j = blockIdx.y;
    const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (col >= d) return;
    if (csc) {
        const int c = blockIdx.z;
        const float4 t = load4_any(dx0, ((size_t)c * Lt + ctx_pos[c * n_ctx + j]) * d + col, dx_f16);
        *reinterpret_cast<float4*>(grad + ((size_t)c * n_ctx + j) * d + col) =
            make_float4(t.x * inv_scale, t.y * inv_scale, t.z * inv_scale, t.w * inv_scale);
        return;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < C; ++c) {
        const float4 t = load4_any(dx0, ((size_t)c * Lt + ctx_pos[c * n_ctx + j]) * d + col, dx_f16);
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    *reinterpret_cast<float4*>(grad + (size_t)j * d + col) =
        make_float4(acc.x * inv_scale, acc.y * inv_scale, acc.z * inv_scale, acc.w * inv_scale);
}

inline int check_d(int d, const char* who) {
    if (d <= 0 || d > kMaxV4 * 128 || (d % 4)) return fail(MVLPT_ESHAPE, "%s: d=%d must be a multiple of 4, <= 1024", who, d);
    return MVLPT_OK;
}

}  // namespace

extern "C" {

int mvlpt_ln_fwd(const void* x, const void* row_index, const void* gamma, const void* beta, void* y, int rows, int d,
                 float eps, mvlpt_stream_t stream) {
    return mvlpt_ln_fwd_hilo(x, row_index, gamma, beta, y, rows, d, eps, 0, stream);
}

int mvlpt_ln_fwd_hilo(const void* x, const void* row_index, const void* gamma, const void* beta, void* y, int rows, int d,
                      float eps, int hilo, mvlpt_stream_t stream) {
    if (!x || !gamma || !beta || !y) return fail(MVLPT_EINVAL, "mvlpt_ln_fwd: null argument");
    if (rows <= 0) return fail(MVLPT_EINVAL, "mvlpt_ln_fwd: rows must be positive");
    int rc = check_d(d, "mvlpt_ln_fwd");
    if (rc) return rc;
    if ((rc = require_sm100())) return rc;
    MVLPT_CUDA_OK(launch_pdl(ln_fwd_kernel, dim3(cdiv(rows, 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1,
                             static_cast<const float*>(x), static_cast<const int*>(row_index), static_cast<const float*>(gamma),
                             static_cast<const float*>(beta), static_cast<__half*>(y), rows, d, eps, hilo));
    return launched("ln_fwd");
}

int mvlpt_ln_bwd(const void* dy, const void* x, const void* row_index, const void* gamma, void* dx_stream, void* dx16,
                 int rows, int d, float eps, int accumulate, mvlpt_stream_t stream) {
    if (!dy || !x || !gamma || (!dx_stream && !dx16)) return fail(MVLPT_EINVAL, "mvlpt_ln_bwd: null argument");
    if (rows <= 0) return fail(MVLPT_EINVAL, "mvlpt_ln_bwd: rows must be positive");
    int rc = check_d(d, "mvlpt_ln_bwd");
    if (rc) return rc;
    if ((rc = require_sm100())) return rc;
    if (!row_index && (d == 512 || d == 768 || d == 1024)) {
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        if (d == 512) launch_ln_bwd_fast<4>(dy, x, gamma, dx_stream, dx16, rows, eps, accumulate, s);
        else if (d == 768) launch_ln_bwd_fast<6>(dy, x, gamma, dx_stream, dx16, rows, eps, accumulate, s);
        else launch_ln_bwd_fast<8>(dy, x, gamma, dx_stream, dx16, rows, eps, accumulate, s);
        return launched("ln_bwd");
    }
    ln_bwd_kernel<<<cdiv(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __half*>(dy), static_cast<const float*>(x), static_cast<const int*>(row_index),
        static_cast<const float*>(gamma), static_cast<float*>(dx_stream), static_cast<__half*>(dx16), rows, d, eps,
        accumulate);
    return launched("ln_bwd");
}

int mvlpt_im2col(const void* img, int img_f32, void* patches, int B, int H, int W, int p, int Kp,
                 mvlpt_stream_t stream) {
    if (!img || !patches) return fail(MVLPT_EINVAL, "mvlpt_im2col: null argument");
    if (B <= 0 || p <= 0 || H % p || W % p || Kp < 3 * p * p || (Kp % 8))
        return fail(MVLPT_ESHAPE, "mvlpt_im2col: bad shape B=%d H=%d W=%d p=%d Kp=%d", B, H, W, p, Kp);
    int rc = require_sm100();
    if (rc) return rc;
    const int patches_n = B * (H / p) * (W / p);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!img_f32 && (p % 8) == 0 && (W % 8) == 0 && (reinterpret_cast<uintptr_t>(img) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(patches) & 15) == 0) {
        int per_patch = 3 * p * p / 8;  // one 16-byte unit per thread
        if (per_patch > 64) per_patch = 64;
        im2col_vec8_kernel<<<cdiv(patches_n, 4), 4 * per_patch, 0, s>>>(static_cast<const __half*>(img),
                                                                        static_cast<__half*>(patches), patches_n, H, W, p,
                                                                        Kp, per_patch);
        return launched("im2col");
    }
    if (img_f32)
        im2col_kernel<float><<<patches_n, 256, 0, s>>>(static_cast<const float*>(img), static_cast<__half*>(patches), B,
                                                       H, W, p, Kp);
    else
        im2col_kernel<__half><<<patches_n, 256, 0, s>>>(static_cast<const __half*>(img), static_cast<__half*>(patches),
                                                        B, H, W, p, Kp);
    return launched("im2col");
}

int mvlpt_embed_assemble(const void* pe, const void* cls, const void* pos, const void* gamma, const void* beta,
                         const void* prompt, int prompt_f16, void* x0, int B, int G, int v, int d, float eps,
                         mvlpt_stream_t stream) {
    if (!pe || !cls || !pos || !gamma || !beta || !x0 || (v > 0 && !prompt))
        return fail(MVLPT_EINVAL, "mvlpt_embed_assemble: null argument");
    if (B <= 0 || G <= 0 || v < 0) return fail(MVLPT_EINVAL, "mvlpt_embed_assemble: bad sizes");
    int rc = check_d(d, "mvlpt_embed_assemble");
    if (rc) return rc;
    if ((rc = require_sm100())) return rc;
    const int rows = B * (1 + v + G);
    embed_assemble_kernel<<<cdiv(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __half*>(pe), static_cast<const float*>(cls), static_cast<const float*>(pos),
        static_cast<const float*>(gamma), static_cast<const float*>(beta), prompt, prompt_f16,
        static_cast<float*>(x0), B, G, v, d, eps);
    return launched("embed_assemble");
}

namespace {
// p in [0,1) -> (threshold on a 16-bit lane, 1/(1-p)); p = 0 -> (0, 1): no dropout
int drop_params(float p, const char* who, unsigned& thr, float& keep_scale) {
    if (!(p >= 0.f) || p >= 1.f) return fail(MVLPT_EINVAL, "%s: dropout probability must be in [0, 1)", who);
    thr = static_cast<unsigned>(p * 65536.f + 0.5f);
    keep_scale = 1.f / (1.f - p);
    return MVLPT_OK;
}
}  // namespace

int mvlpt_set_prompt_rows(void* x, const void* prompt, int prompt_f16, int B, int L, int v, int d, float drop_p,
                          uint64_t seed, int slab, mvlpt_stream_t stream) {
    return mvlpt_set_prompt_rows_ln(x, prompt, prompt_f16, B, L, v, d, drop_p, seed, slab, nullptr, nullptr, nullptr, stream);
}

int mvlpt_set_prompt_rows_ln(void* x, const void* prompt, int prompt_f16, int B, int L, int v, int d, float drop_p,
                             uint64_t seed, int slab, void* xt, void* rec, const void* gamma, mvlpt_stream_t stream) {
    if (!x || !prompt) return fail(MVLPT_EINVAL, "mvlpt_set_prompt_rows: null argument");
    if (xt && (!gamma || !rec)) return fail(MVLPT_EINVAL, "mvlpt_set_prompt_rows_ln: xt needs rec and gamma");
    if (B <= 0 || v <= 0 || L < 1 + v) return fail(MVLPT_EINVAL, "mvlpt_set_prompt_rows: bad sizes");
    int rc = check_d(d, "mvlpt_set_prompt_rows");
    if (rc) return rc;
    unsigned thr;
    float ks;
    if ((rc = drop_params(drop_p, "mvlpt_set_prompt_rows", thr, ks))) return rc;
    if ((rc = require_sm100())) return rc;
    set_prompt_rows_kernel<<<cdiv(B * v, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<float*>(x), prompt, prompt_f16, B, L, v, d, thr, ks, seed, slab, static_cast<__half*>(xt),
        static_cast<float*>(rec), static_cast<const float*>(gamma));
    return launched("set_prompt_rows");
}

int mvlpt_ln_prep(const void* x, const void* gamma, void* xt, void* rec, int rows, int d, mvlpt_stream_t stream) {
    if (!x || !gamma || !xt || !rec) return fail(MVLPT_EINVAL, "mvlpt_ln_prep: null argument");
    if (rows <= 0) return fail(MVLPT_EINVAL, "mvlpt_ln_prep: rows must be positive");
    int rc = check_d(d, "mvlpt_ln_prep");
    if (rc) return rc;
    if ((rc = require_sm100())) return rc;
    MVLPT_CUDA_OK(launch_pdl(ln_prep_kernel, dim3(cdiv(rows, 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1,
                             static_cast<const float*>(x), static_cast<const float*>(gamma), static_cast<__half*>(xt),
                             static_cast<float*>(rec), rows, d));
    return launched("ln_prep");
}

int mvlpt_prompt_grad(void* dx, void* dx16, void* grad, int B, int L, int v, int d, float inv_scale, int zero_rows,
                      float drop_p, uint64_t seed, int slab, mvlpt_stream_t stream) {
    if ((!dx && !dx16) || !grad) return fail(MVLPT_EINVAL, "mvlpt_prompt_grad: null argument");
    if (B <= 0 || v <= 0 || L < 1 + v) return fail(MVLPT_EINVAL, "mvlpt_prompt_grad: bad sizes");
    int rc = check_d(d, "mvlpt_prompt_grad");
    if (rc) return rc;
    unsigned thr;
    float ks;
    if ((rc = drop_params(drop_p, "mvlpt_prompt_grad", thr, ks))) return rc;
    if ((rc = require_sm100())) return rc;
    dim3 grid(cdiv(d, 64), v);
    prompt_grad_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<float*>(dx), static_cast<__half*>(dx16), static_cast<float*>(grad), B, L, v, d, inv_scale, zero_rows,
        thr, ks, seed, slab);
    return launched("prompt_grad");
}

int mvlpt_dropout_keep(void* keep, int B, int v, int d, float drop_p, uint64_t seed, int slab, mvlpt_stream_t stream) {
    if (!keep) return fail(MVLPT_EINVAL, "mvlpt_dropout_keep: null argument");
    if (B <= 0 || v <= 0) return fail(MVLPT_EINVAL, "mvlpt_dropout_keep: bad sizes");
    int rc = check_d(d, "mvlpt_dropout_keep");
    if (rc) return rc;
    unsigned thr;
    float ks;
    if ((rc = drop_params(drop_p, "mvlpt_dropout_keep", thr, ks))) return rc;
    if ((rc = require_sm100())) return rc;
    const int n = B * v * (d >> 2);
    dropout_keep_kernel<<<cdiv(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<unsigned char*>(keep), B, v, d, thr, seed, slab);
    return launched("dropout_keep");
}

int mvlpt_text_assemble(const void* emb, const void* ctx, int ctx_f16, const void* slot, const void* pos, void* x0,
                        int C, int Lt, int n_ctx, int d, int csc, mvlpt_stream_t stream) {
    if (!emb || !pos || !x0 || (ctx && !slot)) return fail(MVLPT_EINVAL, "mvlpt_text_assemble: null argument");
    if (C <= 0 || Lt <= 0) return fail(MVLPT_EINVAL, "mvlpt_text_assemble: bad sizes");
    int rc = check_d(d, "mvlpt_text_assemble");
    if (rc) return rc;
    if ((rc = require_sm100())) return rc;
    text_assemble_kernel<<<cdiv(C * Lt, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float*>(emb), ctx, ctx_f16, static_cast<const int*>(slot), static_cast<const float*>(pos),
        static_cast<float*>(x0), C, Lt, n_ctx, d, csc);
    return launched("text_assemble");
}

int mvlpt_ctx_grad(const void* dx0, int dx_f16, const void* ctx_pos, void* grad, int C, int Lt, int n_ctx, int d, int csc,
                   float inv_scale, mvlpt_stream_t stream) {
    if (!dx0 || !ctx_pos || !grad) return fail(MVLPT_EINVAL, "mvlpt_ctx_grad: null argument");
    if (C <= 0 || Lt <= 0 || n_ctx <= 0) return fail(MVLPT_EINVAL, "mvlpt_ctx_grad: bad sizes");
    int rc = check_d(d, "mvlpt_ctx_grad");
    if (rc) return rc;
    if ((rc = require_sm100())) return rc;
    dim3 grid(cdiv(d / 4, 64), n_ctx, csc ? C : 1);
    ctx_grad_kernel<<<grid, 64, 0, static_cast<cudaStream_t>(stream)>>>(
        dx0, dx_f16, static_cast<const int*>(ctx_pos), static_cast<float*>(grad), C, Lt, n_ctx, d, csc, inv_scale);
    return launched("ctx_grad");
}

int mvlpt_zero(void* p, size_t bytes, mvlpt_stream_t stream) {
    if (!p) return fail(MVLPT_EINVAL, "mvlpt_zero: null argument");
    MVLPT_CUDA_OK(cudaMemsetAsync(p, 0, bytes, static_cast<cudaStream_t>(stream)));
    return MVLPT_OK;
}

}  // extern "C"
