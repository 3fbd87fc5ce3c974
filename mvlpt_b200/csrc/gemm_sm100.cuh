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
// Pipelines: operand ring in shared memory (full/empty mbarriers), a 2-deep TMEM accumulator ring so the epilogue of
// tile i overlaps the MMAs of tile i+1, and a ring of 16 KB output slabs.
//
// Epilogue: one accumulator row per thread, 128 bytes of output per step (64 fp16 or 32 fp32 columns): TMEM ->
// registers -> bias / QuickGELU / residual -> 128B-swizzled slab in shared memory -> TMA store (coalesced,
// asynchronous, clips the ragged edges).  A row-wise INPUT of the epilogue (the fp32 residual stream, or the saved
// fp16 pre-activation of the QuickGELU backward) is TMA-loaded into the very slab the result will leave from, a few
// steps ahead, and updated in place — per-thread row reads from global memory would cost one L1 wavefront per
// 16 bytes.
#pragma once
#include "ptx_sm100.cuh"

namespace mvlpt {

enum GemmAct : int {
    ACT_NONE = 0,
    ACT_QUICKGELU = 1,      // out = t * sigmoid(1.702 t), t = acc + bias     (clip/model.py:162-164)
    ACT_MUL_DQUICKGELU = 2  // out = acc * d/dt[t sigmoid(1.702 t)] at t = aux_in (backward of the above)
};

struct GemmEpilogue {
    const __half* bias;  // [N] or nullptr
    int has_in;          // a row-wise epilogue input arrives through tmap_in: the fp32 residual (fp32 output) or the
                         // saved fp16 pre-activation (ACT_MUL_DQUICKGELU)
    int has_aux_out;     // pre-activation t saved through tmap_aux (ACT_QUICKGELU while training)
    int act;
    float alpha;
    int stages;          // depth of the operand ring
    int ring;            // output ring slots (each 1 slab, or 2 when has_aux_out)
};

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;
constexpr int kGemmThreads = 192;
constexpr int kGemmSlab = 16384;  // 128 rows x 128 bytes
constexpr int kGemmMaxStages = 6;
constexpr int kGemmMaxRing = 6;
constexpr int kGemmSmemBudget = 227 * 1024 - 1024 /*align slack*/ - 256 /*barriers*/;

template <int BN>
struct GemmCfg {
    static constexpr int kABytes = kGemmBM * kGemmBK * 2;
    static constexpr int kBBytes = BN * kGemmBK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = 2 * BN;
    static constexpr int slabs_for(int stages) { return (kGemmSmemBudget - stages * kStageBytes) / kGemmSlab; }
    static constexpr int smem_bytes(int stages, int slabs) { return stages * kStageBytes + slabs * kGemmSlab + 1024 + 256; }
};

// sigmoid(z) = 0.5 + 0.5 tanh(z/2): one MUFU op instead of ex2 + rcp
__device__ __forceinline__ float sigmoid_fast(float z) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
    return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float quickgelu_f(float t) { return t * sigmoid_fast(1.702f * t); }
__device__ __forceinline__ float dquickgelu_f(float t) {
    const float s = sigmoid_fast(1.702f * t);
    return s * fmaf(1.702f * t, 1.f - s, 1.f);
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void gemm_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read_n(int n) {
    switch (n) {
        case 0: tma_store_wait_read<0>(); break;
        case 1: tma_store_wait_read<1>(); break;
        case 2: tma_store_wait_read<2>(); break;
        case 3: tma_store_wait_read<3>(); break;
        case 4: tma_store_wait_read<4>(); break;
        default: tma_store_wait_read<5>(); break;
    }
}

template <int BN, bool OUT_F32>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_f16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                   const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_aux,
                   const __grid_constant__ CUtensorMap tmap_in, int M, int N, int K, GemmEpilogue ep) {
    using Cfg = GemmCfg<BN>;
    constexpr int SW = OUT_F32 ? 32 : 64;  // output columns per 128-byte slab row
    constexpr int kSteps = BN / SW;
    const int kStages = ep.stages;
    const int per = ep.has_aux_out ? 2 : 1;  // slabs per ring slot
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * Cfg::kABytes;
    uint8_t* smem_e = smem + kStages * Cfg::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_e + ep.ring * per * kGemmSlab);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kGemmMaxStages;
    uint64_t* tfull_bar = bars + 2 * kGemmMaxStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* in_full = tempty_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_full + kGemmMaxRing);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_tiles = (M + kGemmBM - 1) / kGemmBM;
    const int n_tiles = (N + BN - 1) / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int k_blocks = (K + kGemmBK - 1) / kGemmBK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_out);
        if (ep.has_aux_out) tma_prefetch_desc(&tmap_aux);
        if (ep.has_in) tma_prefetch_desc(&tmap_in);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 4);
        }
        for (int s = 0; s < kGemmMaxRing; ++s) mbar_init(&in_full[s], 1);
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
        const int r = quarter * 32 + lane;
        const bool issuer = (threadIdx.x == 64);
        const int sw = r & 7;
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        const bool has_in = ep.has_in != 0;
        const bool two = ep.has_aux_out != 0;
        const int R = ep.ring;
        int acc = 0;
        uint32_t acc_phase = 0;
        int g = 0;  // steps done by this CTA
        const int my_tiles = (int)blockIdx.x < num_tiles ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        const int total_steps = my_tiles * kSteps;
        // TMA load of the epilogue input of step gg into its ring slot (issuer thread only)
        auto issue_in = [&](int gg) {
            if (gg >= total_steps) return;
            const int t = blockIdx.x + (gg / kSteps) * gridDim.x;
            const int mm = (t / n_tiles) * kGemmBM, nn = (t % n_tiles) * BN + (gg % kSteps) * SW;
            const int slot = gg % R;
            mbar_arrive_expect_tx(&in_full[slot], kGemmSlab);
            tma_load_2d(smem_e + slot * kGemmSlab, &tmap_in, &in_full[slot], nn, mm);
        };
        if (has_in && issuer)
            for (int gg = 0; gg < R - 1; ++gg) issue_in(gg);

        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile / n_tiles) * kGemmBM;
            const int n0 = (tile % n_tiles) * BN;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int s = 0; s < kSteps; ++s, ++g) {
                const int col0 = n0 + s * SW;
                const int slot = g % R;
                uint8_t* slab = smem_e + slot * per * kGemmSlab;
                if (has_in) {
                    mbar_wait(&in_full[slot], (uint32_t)(g / R) & 1);  // input landed (implies the slot was drained)
                } else {
                    // the slot about to be overwritten must have been drained by the TMA store of step g - R
                    if (issuer) tma_store_wait_read_n(R - 1);
                    gemm_bar_sync();
                }
                uint8_t* orow = slab + row_off;
#pragma unroll
                for (int hb = 0; hb < SW / 32; ++hb) {  // 32 accumulator columns at a time
                    uint32_t raw[32];
                    tmem_ld_32x32(t_row + s * SW + hb * 32, raw);
                    tmem_ld_wait();
                    if (s == kSteps - 1 && hb == SW / 32 - 1) {
                        // last TMEM read of this tile: hand the accumulator back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                    }
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * ep.alpha;
                    const int c0 = col0 + hb * 32;
                    if (ep.bias) {
                        if (c0 + 32 <= N) {
                            const uint4* bp = reinterpret_cast<const uint4*>(ep.bias + c0);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint4 u = __ldg(bp + q);
                                const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float2 f = __half22float2(h2[j]);
                                    v[q * 8 + 2 * j] += f.x;
                                    v[q * 8 + 2 * j + 1] += f.y;
                                }
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (c0 + j < N) v[j] += __half2float(ep.bias[c0 + j]);
                        }
                    }
                    if (OUT_F32) {
                        // fp32 slab row = 32 columns: unit u = columns 4u..4u+3; the residual is already in the slab
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            uint4* dst = reinterpret_cast<uint4*>(orow + ((u ^ sw) << 4));
                            if (has_in) {
                                const uint4 x = *dst;
                                v[4 * u + 0] += __uint_as_float(x.x);
                                v[4 * u + 1] += __uint_as_float(x.y);
                                v[4 * u + 2] += __uint_as_float(x.z);
                                v[4 * u + 3] += __uint_as_float(x.w);
                            }
                            *dst = make_uint4(__float_as_uint(v[4 * u]), __float_as_uint(v[4 * u + 1]),
                                              __float_as_uint(v[4 * u + 2]), __float_as_uint(v[4 * u + 3]));
                        }
                    } else {
                        // fp16 slab row = 64 columns: this half fills units 4*hb .. 4*hb+3 (8 columns each)
                        if (ep.act == ACT_QUICKGELU) {
                            if (two) {
                                uint8_t* arow = orow + kGemmSlab;
#pragma unroll
                                for (int u = 0; u < 4; ++u)
                                    *reinterpret_cast<uint4*>(arow + (((4 * hb + u) ^ sw) << 4)) =
                                        make_uint4(pack_h2(v[8 * u], v[8 * u + 1]), pack_h2(v[8 * u + 2], v[8 * u + 3]),
                                                   pack_h2(v[8 * u + 4], v[8 * u + 5]), pack_h2(v[8 * u + 6], v[8 * u + 7]));
                            }
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = quickgelu_f(v[j]);
                        } else if (ep.act == ACT_MUL_DQUICKGELU) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const uint4 t4 = *reinterpret_cast<const uint4*>(orow + (((4 * hb + u) ^ sw) << 4));
                                const __half2* h2 = reinterpret_cast<const __half2*>(&t4);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float2 f = __half22float2(h2[j]);
                                    v[8 * u + 2 * j] *= dquickgelu_f(f.x);
                                    v[8 * u + 2 * j + 1] *= dquickgelu_f(f.y);
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            *reinterpret_cast<uint4*>(orow + (((4 * hb + u) ^ sw) << 4)) =
                                make_uint4(pack_h2(v[8 * u], v[8 * u + 1]), pack_h2(v[8 * u + 2], v[8 * u + 3]),
                                           pack_h2(v[8 * u + 4], v[8 * u + 5]), pack_h2(v[8 * u + 6], v[8 * u + 7]));
                    }
                }
                fence_proxy_async_smem();
                gemm_bar_sync();
                if (issuer) {
                    if (col0 < N) {
                        tma_store_2d(&tmap_out, slab, col0, m0);
                        if (two) tma_store_2d(&tmap_aux, slab + kGemmSlab, col0, m0);
                    }
                    tma_store_commit();  // one bulk group per step, even when empty: wait_group counts groups
                    if (has_in) {
                        // the slot of step g-1 is drained once every group but the newest has been read:
                        // refill it with the input of step g-1+R
                        tma_store_wait_read<1>();
                        issue_in(g - 1 + R);
                    }
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (issuer) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

}  // namespace mvlpt
