// Input pipeline in front of the image tower: decoded uint8 HWC images -> the normalised [B,3,S,S]
// batch, on the GPU, BIT-IDENTICAL to what the reference's CPU pipeline produces:
//   torchvision (Random)ResizedCrop / Resize [+ CenterCrop] with BICUBIC over Pillow -> RandomHorizontalFlip -> ToTensor ->
//   Normalize   (trainers/vision_benchmark/evaluation/feature.py:540-553; dassl build_transform for
//   configs/trainers/MVLPT/vit_b16.yaml:8-13).
// Pillow's Image.resize (src/libImaging/Resample.c, 8-bit path) is a separable two-pass convolution — horizontal first,
// uint8 intermediate — with double-precision bicubic (a = -0.5) coefficients whose support grows with the down-scaling
// factor, normalised, turned into 22-bit fixed point, accumulated in int32 with a rounding constant, shifted and clipped.
// Three kernels: (1) coefficient tables per image and axis, in fp64 with every operation individually rounded (no FMA
// contraction: the tables must equal what x86-64 Pillow computes), (2) horizontal pass box -> uint8 [bh, S_w, 3],
// (3) vertical pass + /255 + (x - mean)/std (IEEE fp32 division, as torch does) + flip + NCHW store (fp32 or fp16).
// HBM-bound byte work ("at 10 k+ img/s the CPU PIL/bicubic loader becomes the bottleneck").
#include "common.cuh"
#include <cuda_fp16.h>

using namespace mvlpt;

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ double bicubic_w(double x) {
    // ((a + 2) x - (a + 3)) x x + 1  |  (((x - 5) x + 8) x - 4) a   with a = -0.5, evaluated in Pillow's operation order
    if (x < 0.0) x = -x;
    if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(1.5, x), 2.5), x), x), 1.0);
    if (x < 2.0) return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), -0.5);
    return 0.0;
}

struct Axis {
    int in_size, out_size, win0, win;  // resize in_size -> out_size, keep outputs [win0, win0 + win)
};
__device__ __forceinline__ Axis axis_of(const mvlpt_image_desc& d, int axis, int out_h, int out_w) {
    return axis == 0 ? Axis{d.bw, d.rw, d.ox, out_w} : Axis{d.bh, d.rh, d.oy, out_h};
}

// tables for image b, axis a: bounds[(b*2+a)*S + i] = (first tap, tap count); kk[((b*2+a)*K + t)*S + i] (tap-major so that
// neighbouring outputs read neighbouring words).  S = max(out_h, out_w), K = max taps over the batch.
__global__ void coeff_kernel(const mvlpt_image_desc* __restrict__ descs, int B, int out_h, int out_w, int S, int K,
                             int2* __restrict__ bounds, int* __restrict__ kk) {
    const int b = blockIdx.y >> 1, a = blockIdx.y & 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const Axis ax = axis_of(descs[b], a, out_h, out_w);
    if (i >= ax.win) return;
    const int xx = ax.win0 + i;
    const double scale = __ddiv_rn((double)ax.in_size, (double)ax.out_size);
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = __dmul_rn(2.0, filterscale);
    const double ss = __ddiv_rn(1.0, filterscale);
    const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
    int xmin = __double2int_rz(__dadd_rn(__dsub_rn(center, support), 0.5));
    if (xmin < 0) xmin = 0;
    int xmax = __double2int_rz(__dadd_rn(__dadd_rn(center, support), 0.5));
    if (xmax > ax.in_size) xmax = ax.in_size;
    const int n = xmax - xmin;
    double ww = 0.0;
    for (int x = 0; x < n; ++x)
        ww = __dadd_rn(ww, bicubic_w(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss)));
    int* k = kk + (size_t)(b * 2 + a) * K * S + i;
    for (int x = 0; x < K; ++x) {
        int q = 0;
        if (x < n) {
            double w = bicubic_w(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
            if (ww != 0.0) w = __ddiv_rn(w, ww);
            const double f = __dmul_rn(w, (double)(1 << kPrecisionBits));
            q = __double2int_rz(w < 0.0 ? __dadd_rn(-0.5, f) : __dadd_rn(0.5, f));
        }
        k[(size_t)x * S] = q;
    }
    bounds[(size_t)(b * 2 + a) * S + i] = make_int2(xmin, n);
}

__device__ __forceinline__ int clip8(int acc) {
    const int v = acc >> kPrecisionBits;  // arithmetic shift, then clamp (Resample.c clip8 lookup)
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: tmp[b][y][i][c] for y in [0, bh), i in [0, out_w).  Block = kRowsPerBlock source rows of one image,
// staged in shared memory with aligned 16-byte loads (each row keeps its own misalignment); thread = one output column,
// all rows at once, so every tap weight is loaded once per block.
constexpr int kRowsPerBlock = 8;
__global__ void __launch_bounds__(256)
hpass_kernel(const unsigned char* __restrict__ src, const mvlpt_image_desc* __restrict__ descs, int out_w, int S, int K,
             const int2* __restrict__ bounds, const int* __restrict__ kk, unsigned char* __restrict__ tmp,
             size_t tmp_stride, int row_stride) {
    extern __shared__ __align__(16) unsigned char rows[];
    const int b = blockIdx.y;
    const mvlpt_image_desc d = descs[b];
    const int y0 = blockIdx.x * kRowsPerBlock;
    if (y0 >= d.bh) return;
    const int nrows = min(kRowsPerBlock, d.bh - y0);
    const int nbytes = d.bw * 3;
    int mis[kRowsPerBlock];
#pragma unroll
    for (int r = 0; r < kRowsPerBlock; ++r) {
        mis[r] = 0;
        if (r < nrows) {
            const unsigned char* g = src + d.src_off + ((size_t)(d.by + y0 + r) * d.W + d.bx) * 3;
            const int m = (int)(reinterpret_cast<size_t>(g) & 15);
            mis[r] = m;
            const uint4* ga = reinterpret_cast<const uint4*>(g - m);
            uint4* sa = reinterpret_cast<uint4*>(rows + (size_t)r * row_stride);
            for (int v = threadIdx.x; v * 16 < m + nbytes; v += blockDim.x) sa[v] = __ldg(ga + v);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < out_w; i += blockDim.x) {
    const int2 bd = bounds[(size_t)(b * 2) * S + i];
    const int* k = kk + (size_t)(b * 2) * K * S + i;
    int acc[kRowsPerBlock][3];
#pragma unroll
    for (int r = 0; r < kRowsPerBlock; ++r) acc[r][0] = acc[r][1] = acc[r][2] = 1 << (kPrecisionBits - 1);
    for (int t = 0; t < bd.y; ++t) {
        const int w = k[(size_t)t * S];
        const int o = (bd.x + t) * 3;
#pragma unroll
        for (int r = 0; r < kRowsPerBlock; ++r) {
            if (r < nrows) {
                const unsigned char* p = rows + (size_t)r * row_stride + mis[r] + o;
                acc[r][0] += p[0] * w; acc[r][1] += p[1] * w; acc[r][2] += p[2] * w;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kRowsPerBlock; ++r) {
        if (r < nrows) {
            unsigned char* o = tmp + (size_t)b * tmp_stride + ((size_t)(y0 + r) * out_w + i) * 3;
            o[0] = (unsigned char)clip8(acc[r][0]); o[1] = (unsigned char)clip8(acc[r][1]); o[2] = (unsigned char)clip8(acc[r][2]);
        }
    }
    }  // output columns
}

struct Norm {
    float mean[3], std[3];
};
template <typename T> __device__ __forceinline__ T cvt_out(float v);
template <> __device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __half cvt_out<__half>(float v) { return __float2half_rn(v); }

// vertical pass + ToTensor + Normalize + flip: out[b][c][yo][flip ? out_w-1-i : i].  Thread = 4 neighbouring output columns
// of one row (12 contiguous bytes of every intermediate row = three aligned words when out_w % 4 == 0).
template <typename OutT, bool VEC4>
__global__ void __launch_bounds__(256)
vpass_kernel(const mvlpt_image_desc* __restrict__ descs, int out_h, int out_w, int S, int K, const int2* __restrict__ bounds,
             const int* __restrict__ kk, const unsigned char* __restrict__ tmp, size_t tmp_stride, Norm nm,
             OutT* __restrict__ out) {
    constexpr int P = VEC4 ? 4 : 1;
    // ToTensor + Normalize: two IEEE divisions, as torch computes them (a 768-entry shared-memory table of all possible
    // results measured the same: 0.235 vs 0.231 ms per 256 photos)
    auto norm = [&](int c, int px) -> float {
        return __fdiv_rn(__fsub_rn(__fdiv_rn((float)px, 255.f), nm.mean[c]), nm.std[c]);
    };
    const int b = blockIdx.y;
    const int wq = out_w / P;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= out_h * wq) return;
    const int yo = idx / wq, i0 = (idx % wq) * P;
    const int2 bd = bounds[(size_t)(b * 2 + 1) * S + yo];
    const int* k = kk + (size_t)(b * 2 + 1) * K * S + yo;
    const unsigned char* p = tmp + (size_t)b * tmp_stride + ((size_t)bd.x * out_w + i0) * 3;
    int acc[P * 3];
#pragma unroll
    for (int j = 0; j < P * 3; ++j) acc[j] = 1 << (kPrecisionBits - 1);
    for (int t = 0; t < bd.y; ++t) {
        const int w = k[(size_t)t * S];
        const unsigned char* q = p + (size_t)t * out_w * 3;
        if (VEC4) {
            const uint32_t* q4 = reinterpret_cast<const uint32_t*>(q);
            const uint32_t u[3] = {__ldg(q4), __ldg(q4 + 1), __ldg(q4 + 2)};
#pragma unroll
            for (int j = 0; j < 12; ++j) acc[j] += (int)((u[j >> 2] >> (8 * (j & 3))) & 0xFFu) * w;
        } else {
#pragma unroll
            for (int j = 0; j < 3; ++j) acc[j] += q[j] * w;
        }
    }
    const int flip = descs[b].flip;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        OutT* o = out + (((size_t)b * 3 + c) * out_h + yo) * out_w;
        if (VEC4) {  // the 4 columns land in 4 neighbouring outputs (reversed when flipped): one vector store
            OutT w[4];
            __align__(16) OutT v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = cvt_out<OutT>(norm(c, clip8(acc[j * 3 + c])));
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = flip ? w[3 - j] : w[j];  // static indices: selects, no local memory
            OutT* dst = o + (flip ? out_w - 4 - i0 : i0);
            if (sizeof(OutT) == 2) *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(v);
            else *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
        } else {
            o[flip ? out_w - 1 - i0 : i0] = cvt_out<OutT>(norm(c, clip8(acc[c])));
        }
    }
}

struct Plan {
    int S, K, max_bh, max_bw;
    size_t off_bounds, off_kk, off_tmp, tmp_stride, total;
};

int taps(int in_size, int out_size) {  // ksize of Resample.c precompute_coeffs
    double fs = (double)in_size / out_size;
    if (fs < 1.0) fs = 1.0;
    return (int)ceil(2.0 * fs) * 2 + 1;
}

int make_plan(const mvlpt_image_desc* h, int B, int out_h, int out_w, Plan& p, const char* who) {
    if (!h) return fail(MVLPT_EINVAL, "%s: null descriptors", who);
    if (B <= 0 || out_h <= 0 || out_w <= 0 || out_h > 4096 || out_w > 4096)
        return fail(MVLPT_EINVAL, "%s: need B > 0 and output sizes in [1, 4096]", who);
    p.S = out_h > out_w ? out_h : out_w;
    p.K = 1; p.max_bh = 1; p.max_bw = 1;
    for (int b = 0; b < B; ++b) {
        const mvlpt_image_desc& d = h[b];
        if (d.H <= 0 || d.W <= 0 || d.bh <= 0 || d.bw <= 0 || d.by < 0 || d.bx < 0 || d.by + d.bh > d.H || d.bx + d.bw > d.W)
            return fail(MVLPT_ESHAPE, "%s: image %d: crop box outside the %dx%d image", who, b, d.H, d.W);
        if (d.rh <= 0 || d.rw <= 0 || d.oy < 0 || d.ox < 0 || d.oy + out_h > d.rh || d.ox + out_w > d.rw)
            return fail(MVLPT_ESHAPE, "%s: image %d: output window outside the resized %dx%d image", who, b, d.rh, d.rw);
        const int kx = taps(d.bw, d.rw), ky = taps(d.bh, d.rh);
        if (kx > p.K) p.K = kx;
        if (ky > p.K) p.K = ky;
        if (d.bh > p.max_bh) p.max_bh = d.bh;
        if (d.bw > p.max_bw) p.max_bw = d.bw;
    }
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    size_t o = 0;
    p.off_bounds = o; o = up(o + (size_t)B * 2 * p.S * sizeof(int2));
    p.off_kk = o;     o = up(o + (size_t)B * 2 * p.K * p.S * sizeof(int));
    p.tmp_stride = up((size_t)p.max_bh * out_w * 3);
    p.off_tmp = o;    o = up(o + (size_t)B * p.tmp_stride);
    p.total = o;
    return MVLPT_OK;
}

// ToTensor + Normalize of a batch that already has the model's size: uint8 [B, 3, H, W] -> (x / 255 - mean[c]) / std[c],
// the same two IEEE divisions as above (a 256-entry table per channel in shared memory), 16 pixels per thread.
template <typename T>
__global__ void __launch_bounds__(256) normalize_u8_kernel(const uint4* __restrict__ src, T* __restrict__ out, Norm nm,
                                                           int plane16, size_t total16) {
    __shared__ float lut[3][256];
    for (int i = threadIdx.x; i < 768; i += 256) {
        const int c = i >> 8, px = i & 255;
        lut[c][px] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)px, 255.f), nm.mean[c]), nm.std[c]);
    }
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total16; i += (size_t)gridDim.x * 256) {
        const int c = (int)((i / plane16) % 3);
        const uint4 v = __ldg(src + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        float f[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) f[k] = lut[c][(w[k >> 2] >> (8 * (k & 3))) & 255u];
        if constexpr (sizeof(T) == 2) {
            uint4 o[2];
            uint32_t* po = reinterpret_cast<uint32_t*>(o);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const __half2 h = __floats2half2_rn(f[2 * k], f[2 * k + 1]);
                po[k] = *reinterpret_cast<const uint32_t*>(&h);
            }
            reinterpret_cast<uint4*>(out)[2 * i] = o[0];
            reinterpret_cast<uint4*>(out)[2 * i + 1] = o[1];
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                reinterpret_cast<float4*>(out)[4 * i + k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
        }
    }
}

}  // namespace

extern "C" {

int mvlpt_normalize_u8(const void* src, void* out, int out_f16, int B, int H, int W, const float* mean3, const float* std3,
                       mvlpt_stream_t stream) {
    if (!src || !out || !mean3 || !std3) return fail(MVLPT_EINVAL, "mvlpt_normalize_u8: null argument");
    if (B <= 0 || H <= 0 || W <= 0 || ((size_t)H * W) % 16)
        return fail(MVLPT_ESHAPE, "mvlpt_normalize_u8: H*W must be a positive multiple of 16 (got %dx%d)", H, W);
    if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out)) & 15)
        return fail(MVLPT_EINVAL, "mvlpt_normalize_u8: src/out must be 16-byte aligned");
    int rc = require_sm100();
    if (rc) return rc;
    Norm nm;
    for (int c = 0; c < 3; ++c) { nm.mean[c] = mean3[c]; nm.std[c] = std3[c]; }
    const int plane16 = H * W / 16;
    const size_t total16 = (size_t)B * 3 * plane16;
    const int grid = (int)((total16 + 255) / 256 < (size_t)sm_count() * 8 ? (total16 + 255) / 256 : (size_t)sm_count() * 8);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (out_f16)
        normalize_u8_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const uint4*>(src), static_cast<__half*>(out), nm, plane16, total16);
    else
        normalize_u8_kernel<float><<<grid, 256, 0, s>>>(static_cast<const uint4*>(src), static_cast<float*>(out), nm, plane16, total16);
    return launched("normalize_u8");
}

size_t mvlpt_preprocess_workspace(const mvlpt_image_desc* descs_host, int B, int out_h, int out_w) {
    Plan p;
    if (make_plan(descs_host, B, out_h, out_w, p, "mvlpt_preprocess_workspace")) return 0;
    return p.total;
}

int mvlpt_preprocess(const void* src, const mvlpt_image_desc* descs_host, const mvlpt_image_desc* descs_dev, int B,
                     const float* mean3, const float* std3, void* out, int out_f16, int out_h, int out_w, void* workspace,
                     size_t ws_bytes, mvlpt_stream_t stream) {
    if (!src || !descs_dev || !mean3 || !std3 || !out || !workspace) return fail(MVLPT_EINVAL, "mvlpt_preprocess: null argument");
    Plan p;
    int rc = make_plan(descs_host, B, out_h, out_w, p, "mvlpt_preprocess");
    if (rc) return rc;
    if (ws_bytes < p.total) return fail(MVLPT_EINVAL, "mvlpt_preprocess: workspace too small (%zu < %zu)", ws_bytes, p.total);
    if ((rc = require_sm100())) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    int2* bounds = reinterpret_cast<int2*>(ws + p.off_bounds);
    int* kk = reinterpret_cast<int*>(ws + p.off_kk);
    unsigned char* tmp = reinterpret_cast<unsigned char*>(ws + p.off_tmp);
    coeff_kernel<<<dim3(cdiv(p.S, 128), 2 * B), 128, 0, s>>>(descs_dev, B, out_h, out_w, p.S, p.K, bounds, kk);
    if ((rc = launched("preprocess coeff"))) return rc;
    const int row_stride = (int)((((size_t)p.max_bw * 3 + 15) & ~size_t(15)) + 16);  // + the row's own misalignment
    const size_t smem = (size_t)row_stride * kRowsPerBlock;
    if (smem > 200 * 1024) return fail(MVLPT_ESHAPE, "mvlpt_preprocess: crop rows of %d pixels do not fit shared memory", p.max_bw);
    static DynSmemCache attr;
    if (smem > 48 * 1024 && (rc = ensure_dyn_smem(hpass_kernel, smem, attr))) return rc;
    hpass_kernel<<<dim3(cdiv(p.max_bh, kRowsPerBlock), B), 256, smem, s>>>(
        static_cast<const unsigned char*>(src), descs_dev, out_w, p.S, p.K, bounds, kk, tmp, p.tmp_stride, row_stride);
    if ((rc = launched("preprocess hpass"))) return rc;
    Norm nm;
    for (int c = 0; c < 3; ++c) { nm.mean[c] = mean3[c]; nm.std[c] = std3[c]; }
    const bool vec4 = out_w % 4 == 0;
    const dim3 grid(cdiv(out_h * (vec4 ? out_w / 4 : out_w), 256), B);
#define MVLPT_VPASS(T, V)                                                                                              \
    vpass_kernel<T, V><<<grid, 256, 0, s>>>(descs_dev, out_h, out_w, p.S, p.K, bounds, kk, tmp, p.tmp_stride, nm,    \
                                            static_cast<T*>(out))
    if (out_f16) { if (vec4) MVLPT_VPASS(__half, true); else MVLPT_VPASS(__half, false); }
    else         { if (vec4) MVLPT_VPASS(float, true);  else MVLPT_VPASS(float, false); }
#undef MVLPT_VPASS
    return launched("preprocess vpass");
}

}  // extern "C"
