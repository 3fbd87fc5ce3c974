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
    // ---- LayerNorm carried THROUGH the linears (CTA-pair kernel only; see the comment above gemm_f16_tn_2sm_kernel) ----
    // producer side (fp32 residual-stream output, N == row width): also emits xt = (out - c) * gamma as fp16 and the
    // row record of `out`
    const float* lnp_rec_in;   // records of the residual input rows (their mean becomes the centring value c) or nullptr
    float* lnp_rec_out;        // records of the output rows
    const float* lnp_gamma;    // [N] gamma of the LayerNorm that will consume `out`
    __half* lnp_xt;            // [M, N] fp16, row stride N
    // consumer side (A operand = xt of the rows to normalise): out = rstd * (acc - delta * sg) + bp, bp passed as `bias`
    const float* lnc_rec;      // records of the A rows, or nullptr
    const __half* lnc_sg;      // [N] fp16: sum_k gamma[k] W[n,k]   (bias = bp[n] = b[n] + sum_k beta[k] W[n,k])
    int ln_parts;              // partial (sum, sum of squares) pairs per record = 2 * (row width / 256)
    float ln_inv_d;            // 1 / row width
    float ln_eps;
};

// Row record of a residual-stream row x[0..d): kLnRec floats = up to 8 pairs (S1_j, S2_j) — sums of (x - c) and
// (x - c)^2 over the j-th 128-column slice, written by the epilogue thread that produced the slice — then the centring
// value c at index 16.  mean = c + sum S1 / d, var = sum S2 / d - (sum S1 / d)^2.
constexpr int kLnRec = 20;
constexpr int kLnRecC = 16;

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;
constexpr int kGemmThreads = 192;
constexpr int kGemm2Threads = 320;  // CTA-pair kernel: TMA warp, MMA warp, 8 epilogue warps
constexpr int kGemmSlab = 16384;  // 128 rows x 128 bytes
constexpr int kGemmMaxStages = 6;
constexpr int kGemmMaxStages2 = 7;  // CTA-pair kernel: 32 KB stages
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
__device__ __forceinline__ void gemm_bar_sync256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
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
    pdl_launch_dependents();  // the next kernel of the stream may start its prologue ...
    pdl_wait();               // ... and this one touches global memory only once its predecessor is complete

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

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 output tile with ONE
// tcgen05.mma of M = 256 per K step.  Each CTA stages only its 128 rows of A and its 128 of the 256 W rows (N columns)
// per stage — 32 KB instead of 48 KB — and the tensor cores of both SMs read the two W halves from both shared
// memories, so the shared-memory traffic per SM (TMA fill + operand reads) drops from 192 to 128 bytes per clock, which
// is what bounded the single-CTA kernel above (ncu: tensor pipe 85 % active at 60 % of its rate).
// The even CTA issues the MMAs; TMA loads of both CTAs count on its `full` barrier; tcgen05.commit multicasts to the
// `empty` / `tfull` barriers of both; the epilogue warps of both CTAs arrive on its `tempty` barrier.
// The epilogue is the one of the kernel above, per CTA on its own 128 accumulator rows, with twice the warps.
//
// LayerNorm without a LayerNorm kernel (clip/model.py:153-159 as used at :186-187).  For a row x with mean mu and
// rstd = (var + eps)^-1/2:   LN(x) . W^T + b  =  rstd * ( xt . W^T  -  (mu - c) * sg )  +  bp      with
//   xt[k] = (x[k] - c) * gamma[k]  (fp16),   sg[n] = sum_k gamma[k] W[n,k],   bp[n] = b[n] + sum_k beta[k] W[n,k],
// for ANY centring value c.  The residual-stream GEMM that produces x (out-proj, FC2) takes c = the mean of its residual
// input (known from that row's record; the block's update moves the mean only a little, so xt is centred to fp16
// accuracy and mu - c is small), writes xt straight from the registers that hold the fp32 row, and leaves the partial
// sums of (x - c), (x - c)^2 of its 128-column slice in the row record.  The GEMM that consumes LN(x) (QKV, FC1) reads xt
// as its A operand, finishes mean / rstd from the record and applies the identity above in its epilogue.  The rounding is
// that of the explicit LayerNorm path — one fp16 rounding of a centred, gamma-scaled row — the weights are untouched.
template <bool OUT_F32>
__global__ void __launch_bounds__(kGemm2Threads, 1)
gemm_f16_tn_2sm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                       const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_aux,
                       const __grid_constant__ CUtensorMap tmap_in, int M, int N, int K, GemmEpilogue ep) {
    constexpr int BN = 256;
    constexpr int kABytes = kGemmBM * kGemmBK * 2;       // this CTA's 128 rows of A
    constexpr int kBBytes = (BN / 2) * kGemmBK * 2;      // this CTA's 128 rows of W
    constexpr int kStageBytes = kABytes + kBBytes;
    constexpr int SW = OUT_F32 ? 32 : 64;  // output columns per 128-byte slab row
    constexpr int kSteps = BN / SW;
    const int kStages = ep.stages;
    const int per = ep.has_aux_out ? 2 : 1;  // slabs per ring slot
    extern __shared__ __align__(1024) uint8_t smem_gemm2[];
    uint8_t* smem = smem_gemm2;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * kABytes;
    uint8_t* smem_e = smem + kStages * kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_e + ep.ring * per * kGemmSlab);
    uint64_t* full_bar = bars;                          // used in the even CTA only
    uint64_t* empty_bar = bars + kGemmMaxStages2;
    uint64_t* tfull_bar = bars + 2 * kGemmMaxStages2;
    uint64_t* tempty_bar = tfull_bar + 2;               // used in the even CTA only
    uint64_t* in_full = tempty_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_full + kGemmMaxRing);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int m_tiles = (M + 2 * kGemmBM - 1) / (2 * kGemmBM);
    const int n_tiles = (N + BN - 1) / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int k_blocks = (K + kGemmBK - 1) / kGemmBK;

    if (warp == 0 && lane == 0) {
        if (smem_u32(smem) & 1023u) {
            printf("mvlpt: gemm dynamic shared memory is not 1024-byte aligned\n");
            __trap();
        }
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
            mbar_init(&tempty_bar[s], 16);  // 8 epilogue warps of each CTA
        }
        for (int s = 0; s < kGemmMaxRing; ++s) mbar_init(&in_full[s], 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc_2sm(tmem_slot, 512);
        tmem_relinquish_2sm();
    }
    tc_fence_before();
    cluster_sync_all();  // barriers of both CTAs are initialised before either touches the other's
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();  // the next kernel of the stream may start its prologue ...
    pdl_wait();               // ... and this one touches global memory only once its predecessor is complete

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                const int m0 = (tile / n_tiles) * (2 * kGemmBM) + (int)rank * kGemmBM;
                const int n0 = (tile % n_tiles) * BN + (int)rank * (BN / 2);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes);  // both CTAs' bytes
                    tma_load_2d_2sm(smem_a + stage * kABytes, &tmap_a, &full_bar[stage], kb * kGemmBK, m0);
                    tma_load_2d_2sm(smem_b + stage * kBBytes, &tmap_w, &full_bar[stage], kb * kGemmBK, n0);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (even CTA; the whole warp runs the loop, one lane issues) =====================
        if (rank == 0) {
            const bool leader = elect_one();
            constexpr uint32_t idesc = umma_idesc_f16(2 * kGemmBM, BN, 0, 0);
            const uint64_t a_base = umma_desc_k_sw128(smem_u32(smem_a));
            const uint64_t b_base = umma_desc_k_sw128(smem_u32(smem_b));
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t a_desc = a_base + (uint32_t)(stage * (kABytes >> 4));
                    const uint64_t b_desc = b_base + (uint32_t)(stage * (kBBytes >> 4));
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < kGemmBK / 16; ++k)
                            umma_f16_ss_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                        umma_commit_2sm(&empty_bar[stage]);
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if (leader) umma_commit_2sm(&tfull_bar[acc]);
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (8 warps = 2 per TMEM lane quarter) =====================
        // A thread owns one accumulator row; the two warps of a quarter split the columns of every 128-byte slab step
        // (fp16: 32 + 32 of 64 columns, fp32: 16 + 16 of 32), so every scheduler has two epilogue warps to interleave.
        constexpr int HC = SW / 2;             // accumulator columns per thread per step
        const int quarter = warp & 3;          // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;      // which half of the step's columns
        const int r = quarter * 32 + lane;
        const bool issuer = (threadIdx.x == 64);
        const int sw = r & 7;
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        const bool has_in = ep.has_in != 0;
        const bool two = ep.has_aux_out != 0;
        const bool scaled = ep.alpha != 1.f;
        const int R = ep.ring;
        int acc = 0;
        uint32_t acc_phase = 0;
        int g = 0;  // steps done by this CTA
        const int my_tiles = pair < num_tiles ? (num_tiles - 1 - pair) / num_pairs + 1 : 0;
        const int total_steps = my_tiles * kSteps;
        // TMA load of the epilogue input of step gg into its ring slot (issuer thread only)
        auto issue_in = [&](int gg) {
            if (gg >= total_steps) return;
            const int t = pair + (gg / kSteps) * num_pairs;
            const int mm = (t / n_tiles) * (2 * kGemmBM) + (int)rank * kGemmBM, nn = (t % n_tiles) * BN + (gg % kSteps) * SW;
            const int slot = gg % R;
            mbar_arrive_expect_tx(&in_full[slot], kGemmSlab);
            tma_load_2d(smem_e + slot * kGemmSlab, &tmap_in, &in_full[slot], nn, mm);
        };
        if (has_in && issuer)
            for (int gg = 0; gg < R - 1; ++gg) issue_in(gg);

        const bool lnc = !OUT_F32 && ep.lnc_rec != nullptr;
        const bool lnp = OUT_F32 && ep.lnp_rec_out != nullptr;
        for (int tile = pair; tile < num_tiles; tile += num_pairs) {
            const int m0 = (tile / n_tiles) * (2 * kGemmBM) + (int)rank * kGemmBM;
            const int n0 = (tile % n_tiles) * BN;
            // LayerNorm carried through the linears: per-row coefficients out of the row record (loads overlap the wait)
            const int gm = m0 + r;              // global row of this thread
            float ln_a = 1.f, ln_d = 0.f;       // consumer: out = ln_a * acc + ln_d * sg + bp
            float ln_c = 0.f, ln_s1 = 0.f, ln_s2 = 0.f;  // producer: centring value, partial sums of this thread's slice
            if ((lnc || (lnp && ep.lnp_rec_in)) && gm < M) {
                const float* rec = (lnc ? ep.lnc_rec : ep.lnp_rec_in) + (size_t)gm * kLnRec;
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (2 * q < ep.ln_parts) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(rec) + q);
                        s1 += t.x + t.z;
                        s2 += t.y + t.w;
                    }
                const float delta = s1 * ep.ln_inv_d;  // mean - c
                if (lnc) {
                    ln_a = rsqrtf(fmaxf(s2 * ep.ln_inv_d - delta * delta, 0.f) + ep.ln_eps);
                    ln_d = -delta * ln_a;
                } else {
                    ln_c = __ldg(rec + kLnRecC) + delta;  // the mean of the residual input row
                }
            }
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN + half * HC;
#pragma unroll 1
            for (int s = 0; s < kSteps; ++s, ++g) {
                const int col0 = n0 + s * SW;
                const int c0 = col0 + half * HC;  // first output column of this thread in this step
                const int slot = g % R;
                uint8_t* slab = smem_e + slot * per * kGemmSlab;
                // bias of this thread's columns: issued first, its latency hides behind the waits below
                uint4 braw[HC / 8];
                const bool bias_vec = ep.bias != nullptr && c0 + HC <= N;
                if (bias_vec) {
#pragma unroll
                    for (int q = 0; q < HC / 8; ++q) braw[q] = __ldg(reinterpret_cast<const uint4*>(ep.bias + c0) + q);
                }
                uint4 sgraw[HC / 8];
                if (lnc) {
#pragma unroll
                    for (int q = 0; q < HC / 8; ++q) sgraw[q] = __ldg(reinterpret_cast<const uint4*>(ep.lnc_sg + c0) + q);
                }
                // the slot about to be overwritten has been drained: with an epilogue input its arrival implies it; without,
                // the issuer made sure of it BEFORE the barrier that ended the previous step (one barrier per step, not two)
                if (has_in) mbar_wait(&in_full[slot], (uint32_t)(g / R) & 1);
                uint8_t* orow = slab + row_off;
                uint32_t raw[HC];
                if constexpr (HC == 32) tmem_ld_32x32(t_row + s * SW, raw);
                else tmem_ld_32x32b_x16(t_row + s * SW, raw);
                tmem_ld_wait();
                if (s == kSteps - 1) {
                    // last TMEM read of this tile: hand the accumulator back to the MMA warp of the even CTA
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cta0(&tempty_bar[acc]);
                }
                float v[HC];
#pragma unroll
                for (int j = 0; j < HC; ++j) v[j] = __uint_as_float(raw[j]);
                if (scaled) {
#pragma unroll
                    for (int j = 0; j < HC; ++j) v[j] *= ep.alpha;
                }
                if (lnc) {
                    // out = rstd * acc + (-(mean - c) * rstd * sg + bp): the bracket in half2 (a small correction plus the
                    // folded bias), N % 256 == 0 in this mode so bias_vec holds
                    const __half2 d2 = __float2half2_rn(ln_d);
#pragma unroll
                    for (int q = 0; q < HC / 8; ++q) {
                        const __half2* sg2 = reinterpret_cast<const __half2*>(&sgraw[q]);
                        const __half2* bp2 = reinterpret_cast<const __half2*>(&braw[q]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 f = __half22float2(__hfma2(d2, sg2[j], bp2[j]));
                            v[q * 8 + 2 * j] = fmaf(v[q * 8 + 2 * j], ln_a, f.x);
                            v[q * 8 + 2 * j + 1] = fmaf(v[q * 8 + 2 * j + 1], ln_a, f.y);
                        }
                    }
                } else if (bias_vec) {
#pragma unroll
                    for (int q = 0; q < HC / 8; ++q) {
                        const __half2* h2 = reinterpret_cast<const __half2*>(&braw[q]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 f = __half22float2(h2[j]);
                            v[q * 8 + 2 * j] += f.x;
                            v[q * 8 + 2 * j + 1] += f.y;
                        }
                    }
                } else if (ep.bias) {
#pragma unroll
                    for (int j = 0; j < HC; ++j)
                        if (c0 + j < N) v[j] += __half2float(ep.bias[c0 + j]);
                }
                if constexpr (OUT_F32) {
                    // fp32 slab row = 32 columns = 8 units of 4; this thread fills units 4*half .. 4*half+3; the
                    // residual is already in the slab
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint4* dst = reinterpret_cast<uint4*>(orow + (((4 * half + u) ^ sw) << 4));
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
                    if (lnp && gm < M) {
                        // xt = (x - c) * gamma for this thread's 16 columns, straight from the registers: 32 contiguous
                        // bytes of its row; partial sums of the centred values for the row record
                        uint32_t xt[HC / 2];
#pragma unroll
                        for (int q = 0; q < HC / 4; ++q) {
                            const float4 gmm = __ldg(reinterpret_cast<const float4*>(ep.lnp_gamma + c0) + q);
                            const float y0 = v[4 * q] - ln_c, y1 = v[4 * q + 1] - ln_c, y2 = v[4 * q + 2] - ln_c,
                                        y3 = v[4 * q + 3] - ln_c;
                            ln_s1 += (y0 + y1) + (y2 + y3);
                            ln_s2 = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, fmaf(y3, y3, ln_s2))));
                            xt[2 * q] = pack_h2(y0 * gmm.x, y1 * gmm.y);
                            xt[2 * q + 1] = pack_h2(y2 * gmm.z, y3 * gmm.w);
                        }
                        // one 256-bit store: the thread's 32 bytes are one full sector of its row
                        static_assert(HC == 16 || !OUT_F32, "xt slice = 16 columns per thread");
                        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ep.lnp_xt + (size_t)gm * N + c0),
                                     "r"(xt[0]), "r"(xt[1]), "r"(xt[2]), "r"(xt[3]), "r"(xt[4]), "r"(xt[5]), "r"(xt[6]), "r"(xt[7])
                                     : "memory");
                    }
                } else {
                    // fp16 slab row = 64 columns = 8 units of 8; this thread fills units 4*half .. 4*half+3
                    if (ep.act == ACT_QUICKGELU) {
                        if (two) {
                            uint8_t* arow = orow + kGemmSlab;
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                *reinterpret_cast<uint4*>(arow + (((4 * half + u) ^ sw) << 4)) =
                                    make_uint4(pack_h2(v[8 * u], v[8 * u + 1]), pack_h2(v[8 * u + 2], v[8 * u + 3]),
                                               pack_h2(v[8 * u + 4], v[8 * u + 5]), pack_h2(v[8 * u + 6], v[8 * u + 7]));
                        }
#pragma unroll
                        for (int j = 0; j < HC; ++j) v[j] = quickgelu_f(v[j]);
                    } else if (ep.act == ACT_MUL_DQUICKGELU) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint4 t4 = *reinterpret_cast<const uint4*>(orow + (((4 * half + u) ^ sw) << 4));
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
                    for (int u = 0; u < HC / 8; ++u)
                        *reinterpret_cast<uint4*>(orow + (((4 * half + u) ^ sw) << 4)) =
                            make_uint4(pack_h2(v[8 * u], v[8 * u + 1]), pack_h2(v[8 * u + 2], v[8 * u + 3]),
                                       pack_h2(v[8 * u + 4], v[8 * u + 5]), pack_h2(v[8 * u + 6], v[8 * u + 7]));
                }
                fence_proxy_async_smem();
                // the slot of step g+1 was last read by the store of step g+1-R: of the groups committed so far (up to
                // g-1) all but the newest R-2 must have been read before anybody passes this barrier
                if (!has_in && issuer) tma_store_wait_read_n(R - 2);
                gemm_bar_sync256();
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
            if (lnp && gm < M) {
                float* rec = ep.lnp_rec_out + (size_t)gm * kLnRec;
                *reinterpret_cast<float2*>(rec + 2 * (2 * (tile % n_tiles) + half)) = make_float2(ln_s1, ln_s2);
                if (n0 == 0 && half == 0) rec[kLnRecC] = ln_c;
            }
        }
        if (issuer) tma_store_wait_all();
    }

    tc_fence_before();
    cluster_sync_all();  // neither CTA may exit (or free TMEM) while the other still reads its memory / signals it
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, 512);
    }
}

}  // namespace mvlpt
