// Persistent, warp-specialised tcgen05 GEMM for the MVLPT hot path.
//
//   OUT[M,N] = epilogue( alpha * A[M,K] . W[N,K]^T )           fp16 operands, fp32 accumulate in TMEM
//
// Replaces every nn.Linear / in_proj / out_proj / c_fc / c_proj call of the reference block
// (clip/model.py:171-177,183) and, with a transposed weight copy, every dgrad (autograd of the same).
// Both operands are K-major exactly as the reference stores them (activations [tokens, features],
// weights [out, in]).
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..5 = epilogue.
// Pipelines: smem ring (full/empty mbarriers, kStages deep) and a 2-deep TMEM accumulator ring so the
// epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include "ptx_sm100.cuh"

namespace mvlpt {

enum GemmAct : int {
    ACT_NONE = 0,
    ACT_QUICKGELU = 1,      // out = t * sigmoid(1.702 t), t = acc + bias     (clip/model.py:162-164)
    ACT_MUL_DQUICKGELU = 2  // out = acc * d/dt[t sigmoid(1.702 t)] at t = aux_in (backward of the above)
};

struct GemmEpilogue {
    const __half* bias;    // [N] or nullptr
    const __half* aux_in;  // [M, ld_aux] or nullptr   (ACT_MUL_DQUICKGELU)
    __half* aux_out;       // [M, ld_aux] or nullptr   (pre-activation t saved for backward)
    const float* resid;    // [M, ld_out] fp32 or nullptr; may alias out
    void* out;             // [M, ld_out] fp16 or fp32
    int ld_out;
    int ld_aux;
    int out_f32;
    int act;
    float alpha;
};

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;
constexpr int kGemmThreads = 192;

template <int BN>
struct GemmCfg {
    static constexpr int kStages = (BN == 256) ? 4 : 6;
    static constexpr int kABytes = kGemmBM * kGemmBK * 2;
    static constexpr int kBBytes = BN * kGemmBK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int kTmemCols = 2 * BN;
};

__device__ __forceinline__ float quickgelu_f(float t) {
    float s = 1.f / (1.f + __expf(-1.702f * t));
    return t * s;
}
__device__ __forceinline__ float dquickgelu_f(float t) {
    float s = 1.f / (1.f + __expf(-1.702f * t));
    return s * (1.f + 1.702f * t * (1.f - s));
}

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_f16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, int M, int N,
                   int K, GemmEpilogue ep) {
    using Cfg = GemmCfg<BN>;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * Cfg::kABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* tfull_bar = bars + 2 * kStages;
    uint64_t* tempty_bar = bars + 2 * kStages + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_tiles = (M + kGemmBM - 1) / kGemmBM;
    const int n_tiles = (N + BN - 1) / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int k_blocks = (K + kGemmBK - 1) / kGemmBK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 4);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles) * kGemmBM;
                const int n0 = (tile % n_tiles) * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                    tma_load_2d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * kGemmBK, m0);
                    tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_w, &full_bar[stage], kb * kGemmBK, n0);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(kGemmBM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t a_desc = umma_desc_k_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
                    const uint64_t b_desc = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
                    for (int k = 0; k < kGemmBK / 16; ++k) {
                        // advance 16 elements (32 B) along K inside the 128 B swizzle row: +2 in 16 B units
                        umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (4 warps, one accumulator row per thread) =====================
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access
        const int row_in_tile = quarter * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile / n_tiles) * kGemmBM;
            const int n0 = (tile % n_tiles) * BN;
            const int row = m0 + row_in_tile;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t raw[32];
                tmem_ld_32x32(t_row + c * 32, raw);
                tmem_ld_wait();
                const int col0 = n0 + c * 32;
                if (row < M && col0 < N) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * ep.alpha;
                    const bool full = (col0 + 32 <= N);
                    if (ep.bias) {
                        if (full) {
                            const uint4* bp = reinterpret_cast<const uint4*>(ep.bias + col0);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                uint4 u = __ldg(bp + q);
                                const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    float2 f = __half22float2(h2[j]);
                                    v[q * 8 + 2 * j] += f.x;
                                    v[q * 8 + 2 * j + 1] += f.y;
                                }
                            }
                        } else {
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < N) v[j] += __half2float(ep.bias[col0 + j]);
                        }
                    }
                    if (ep.act == ACT_QUICKGELU) {
                        if (ep.aux_out) {
                            __half* ap = ep.aux_out + (size_t)row * ep.ld_aux + col0;
                            if (full) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    uint4 u;
                                    __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        h2[j] = __floats2half2_rn(v[q * 8 + 2 * j], v[q * 8 + 2 * j + 1]);
                                    reinterpret_cast<uint4*>(ap)[q] = u;
                                }
                            } else {
                                for (int j = 0; j < 32; ++j)
                                    if (col0 + j < N) ap[j] = __float2half_rn(v[j]);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = quickgelu_f(v[j]);
                    } else if (ep.act == ACT_MUL_DQUICKGELU) {
                        const __half* ap = ep.aux_in + (size_t)row * ep.ld_aux + col0;
                        if (full) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                uint4 u = reinterpret_cast<const uint4*>(ap)[q];
                                const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    float2 f = __half22float2(h2[j]);
                                    v[q * 8 + 2 * j] *= dquickgelu_f(f.x);
                                    v[q * 8 + 2 * j + 1] *= dquickgelu_f(f.y);
                                }
                            }
                        } else {
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < N) v[j] *= dquickgelu_f(__half2float(ap[j]));
                        }
                    }
                    if (ep.resid) {
                        const float* rp = ep.resid + (size_t)row * ep.ld_out + col0;
                        if (full) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                float4 f = reinterpret_cast<const float4*>(rp)[q];
                                v[q * 4 + 0] += f.x;
                                v[q * 4 + 1] += f.y;
                                v[q * 4 + 2] += f.z;
                                v[q * 4 + 3] += f.w;
                            }
                        } else {
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < N) v[j] += rp[j];
                        }
                    }
                    if (ep.out_f32) {
                        float* op = reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ld_out + col0;
                        if (full) {
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                reinterpret_cast<float4*>(op)[q] =
                                    make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                        } else {
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < N) op[j] = v[j];
                        }
                    } else {
                        __half* op = reinterpret_cast<__half*>(ep.out) + (size_t)row * ep.ld_out + col0;
                        if (full) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                uint4 u;
                                __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    h2[j] = __floats2half2_rn(v[q * 8 + 2 * j], v[q * 8 + 2 * j + 1]);
                                reinterpret_cast<uint4*>(op)[q] = u;
                            }
                        } else {
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < N) op[j] = __float2half_rn(v[j]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

}  // namespace mvlpt
