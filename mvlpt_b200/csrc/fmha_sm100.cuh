// tcgen05 attention FORWARD, head width 64, whole key range of a sequence in one pass (L <= 272).
//
// Replaces the inside of nn.MultiheadAttention as called at clip/model.py:181-183 (q*hd^-1/2, QK^T, causal -inf mask
// of the text tower clip/model.py:324-330, softmax, PV).  One work unit = (sequence n, head h, 128-query tile):
//
//   TMA   : Q tile [128 x 64], K [Lp x 64], V [Lp x 64] straight out of the packed in_proj output [N, L, 3d] through a
//           3-D tensor map (rows >= L are zero-filled by the TMA unit, so ragged L needs no padding in HBM)
//   MMA 1 : S[128 x Lp] = Q . K^T        tcgen05.mma, fp32 accumulator in TMEM columns [0, Lp)
//   warps : one thread per query row: row max, p = exp2(s*scale*log2e - max), row sum; P (fp16, unnormalised) goes
//           to shared memory in the 128B-swizzled K-major layout the next MMA reads (it overwrites the dead Q/K tiles)
//   MMA 2 : O[128 x 64] = P . V          V is consumed in place as an MN-major operand; O aliases S's TMEM columns
//   store : O / rowsum -> fp16 -> swizzled staging tile -> TMA store (clips rows >= L); LSE (for the backward) -> HBM
//
// Persistent CTAs (192 threads: TMA warp, MMA warp, 4 softmax/epilogue warps), two per SM so that one CTA's
// MUFU-bound softmax overlaps the other's loads and MMAs.  With head width 64 the kernel is exp-bound, not
// MMA-bound (26.6k exps per 6.8 MFLOP tile): see DESIGN.md.
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"
#include <stdlib.h>

namespace mvlpt {

struct FmhaFwdParams {
    int L, Lp, heads, q_tiles, num_tiles, causal;
    float scale_log2e;  // hd^-1/2 * log2(e)
    float scale;        // hd^-1/2
    float* lse;         // [N, heads, L]
    int box_h;          // rows per K/V TMA box (Lp / number of boxes)
    uint32_t tmem_cols;
    uint32_t off_v, off_o, off_bar;  // byte offsets inside the 1024-aligned dynamic shared memory
};

constexpr int kFmhaThreads = 192;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kFmhaThreads, 2)
fmha_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                   const __grid_constant__ CUtensorMap tmap_out, const FmhaFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                    // [128 x 64] K-major SW128          (dead after MMA 1)
    uint8_t* sK = smem + 16384;            // [Lp x 64]  K-major SW128          (dead after MMA 1)
    uint8_t* sP = smem;                    // ceil(Lp/64) chunks of [128 x 64]  (aliases Q and K)
    uint8_t* sV = smem + p.off_v;          // [Lp x 64]  rows = keys: MN-major B operand of MMA 2
    uint8_t* sO = smem + p.off_o;          // [128 x 64] staging for the TMA store
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
    uint64_t* qk_full = bars + 0;   // [2] by iteration parity
    uint64_t* v_full = bars + 2;
    uint64_t* s_full = bars + 4;
    uint64_t* p_full = bars + 6;
    uint64_t* o_full = bars + 8;
    uint64_t* o_read = bars + 10;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Lp = p.Lp, d = p.heads * 64;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        tma_prefetch_desc(&tmap_out);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&qk_full[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 128);
            mbar_init(&o_full[i], 1);
            mbar_init(&o_read[i], 128);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int qt = tile % p.q_tiles, nh = tile / p.q_tiles;
                const int h = nh % p.heads, n = nh / p.heads;
                // every barrier completes exactly once per tile: use k counts completions (parity (k)&1 of slot it&1)
                const int slot = it & 1;
                if (it > 0) mbar_wait(&o_full[(it - 1) & 1], ((it - 1) >> 1) & 1);  // P / K / V of the previous tile consumed
                mbar_arrive_expect_tx(&qk_full[slot], 16384u + (uint32_t)Lp * 128u);
                tma_load_3d(sQ, &tmap_q, &qk_full[slot], h * 64, qt * 128, n);
                for (int r0 = 0; r0 < Lp; r0 += p.box_h)
                    tma_load_3d(sK + r0 * 128, &tmap_kv, &qk_full[slot], d + h * 64, r0, n);
                mbar_arrive_expect_tx(&v_full[slot], (uint32_t)Lp * 128u);
                for (int r0 = 0; r0 < Lp; r0 += p.box_h)
                    tma_load_3d(sV + r0 * 128, &tmap_kv, &v_full[slot], 2 * d + h * 64, r0, n);
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int slot = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                mbar_wait(&qk_full[slot], ph);
                if (it > 0) mbar_wait(&o_read[(it - 1) & 1], ((it - 1) >> 1) & 1);  // O of the previous tile left TMEM
                tc_fence_after();
                // S = Q . K^T, N split in chunks of <= 256 columns
                const uint64_t q_desc = umma_desc_k_sw128(smem_u32(sQ));
                for (int n0 = 0; n0 < Lp; n0 += 256) {
                    const int nn = (Lp - n0) < 256 ? (Lp - n0) : 256;
                    const uint32_t idesc = umma_idesc_f16(128, (uint32_t)nn, 0, 0);
                    const uint64_t k_desc = umma_desc_k_sw128(smem_u32(sK + n0 * 128));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_ss(tmem_base + n0, q_desc + 2 * k, k_desc + 2 * k, idesc, k != 0);
                }
                umma_commit(&s_full[slot]);
                // O = P . V
                mbar_wait(&p_full[slot], ph);
                mbar_wait(&v_full[slot], ph);
                tc_fence_after();
                const uint32_t idesc_pv = umma_idesc_f16(128, 64, 0, 1);  // B (V) is MN-major
                const int ksteps = Lp / 16;
                for (int kk = 0; kk < ksteps; ++kk) {
                    const uint64_t a_desc = umma_desc_k_sw128(smem_u32(sP + (kk >> 2) * 16384 + (kk & 3) * 32));
                    const uint64_t b_desc = umma_desc_mn_sw128(smem_u32(sV + kk * 2048), 1024);
                    umma_f16_ss(tmem_base, a_desc, b_desc, idesc_pv, kk != 0);
                }
                umma_commit(&o_full[slot]);
            }
        }
    } else {
        // ============================== softmax + epilogue (one query row per thread) ==============================
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;  // row inside the tile == TMEM lane
        const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16);
        const int etid = threadIdx.x - 64;  // 0..127
        uint8_t* prow = sP + (r >> 3) * 1024 + (r & 7) * 128;
        uint8_t* orow = sO + (r >> 3) * 1024 + (r & 7) * 128;
        const int sw = r & 7;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int qt = tile % p.q_tiles, nh = tile / p.q_tiles;
            const int h = nh % p.heads, n = nh / p.heads;
            const int slot = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            const int q = qt * 128 + r;
            const bool warp_live = (qt * 128 + quarter * 32) < p.L;  // warp-uniform
            constexpr int lo = 0;  // this kernel never packs sequences
            const int lim = p.causal ? (q + 1 < p.L ? q + 1 : p.L) : p.L;  // valid keys: [0, lim)
            mbar_wait(&s_full[slot], ph);
            tc_fence_after();
            float inv_sum = 0.f;
            if (warp_live) {
                // pass 1: row max of the raw scores over the valid keys
                float m = -INFINITY;
                for (int c0 = 0; c0 < Lp; c0 += 16) {
                    uint32_t raw[16];
                    tmem_ld_32x32b_x16(t_row + c0, raw);
                    tmem_ld_wait();
                    if (c0 >= lo && c0 + 16 <= lim) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) m = fmaxf(m, __uint_as_float(raw[j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j >= lo && c0 + j < lim) m = fmaxf(m, __uint_as_float(raw[j]));
                    }
                }
                if (m == -INFINITY) m = 0.f;  // rows >= L (never stored)
                const float m2 = m * p.scale_log2e;
                // pass 2: p = exp2(s*scale*log2e - m2); unnormalised fp16 P -> smem; fp32 row sum
                float sum = 0.f;
                for (int c0 = 0; c0 < Lp; c0 += 16) {
                    uint32_t raw[16];
                    tmem_ld_32x32b_x16(t_row + c0, raw);
                    tmem_ld_wait();
                    float e[16];
                    if (c0 + 16 <= lim) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) e[j] = ex2_approx(fmaf(__uint_as_float(raw[j]), p.scale_log2e, -m2));
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            e[j] = (c0 + j < lim) ? ex2_approx(fmaf(__uint_as_float(raw[j]), p.scale_log2e, -m2)) : 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum += e[j];
                    uint8_t* chunk = prow + (c0 >> 6) * 16384;
                    const int u = (c0 & 63) >> 3;  // first of the two 16-byte units this group fills
                    *reinterpret_cast<uint4*>(chunk + ((u ^ sw) << 4)) =
                        make_uint4(pack_half2(e[0], e[1]), pack_half2(e[2], e[3]), pack_half2(e[4], e[5]), pack_half2(e[6], e[7]));
                    *reinterpret_cast<uint4*>(chunk + (((u + 1) ^ sw) << 4)) =
                        make_uint4(pack_half2(e[8], e[9]), pack_half2(e[10], e[11]), pack_half2(e[12], e[13]),
                                   pack_half2(e[14], e[15]));
                }
                inv_sum = 1.f / sum;
                if (q < p.L) p.lse[((size_t)n * p.heads + h) * p.L + q] = m * p.scale + __logf(sum);
            }
            fence_proxy_async_smem();  // P (generic-proxy stores) must be visible to the tensor core's async-proxy reads
            tc_fence_before();
            mbar_arrive(&p_full[slot]);

            mbar_wait(&o_full[slot], ph);
            tc_fence_after();
            uint32_t o0[32], o1[32];
            tmem_ld_32x32(t_row, o0);
            tmem_ld_32x32(t_row + 32, o1);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&o_read[slot]);
            // the previous tile's TMA store must have finished reading the staging tile
            if (etid == 0) tma_store_wait_read<0>();
            named_bar_sync(1, 128);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t* s = o0 + u * 8;
                *reinterpret_cast<uint4*>(orow + ((u ^ sw) << 4)) = make_uint4(
                    pack_half2(__uint_as_float(s[0]) * inv_sum, __uint_as_float(s[1]) * inv_sum),
                    pack_half2(__uint_as_float(s[2]) * inv_sum, __uint_as_float(s[3]) * inv_sum),
                    pack_half2(__uint_as_float(s[4]) * inv_sum, __uint_as_float(s[5]) * inv_sum),
                    pack_half2(__uint_as_float(s[6]) * inv_sum, __uint_as_float(s[7]) * inv_sum));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t* s = o1 + u * 8;
                *reinterpret_cast<uint4*>(orow + (((u + 4) ^ sw) << 4)) = make_uint4(
                    pack_half2(__uint_as_float(s[0]) * inv_sum, __uint_as_float(s[1]) * inv_sum),
                    pack_half2(__uint_as_float(s[2]) * inv_sum, __uint_as_float(s[3]) * inv_sum),
                    pack_half2(__uint_as_float(s[4]) * inv_sum, __uint_as_float(s[5]) * inv_sum),
                    pack_half2(__uint_as_float(s[6]) * inv_sum, __uint_as_float(s[7]) * inv_sum));
            }
            fence_proxy_async_smem();
            named_bar_sync(1, 128);
            if (etid == 0) {
                tma_store_3d(&tmap_out, sO, h * 64, qt * 128, n);
                tma_store_commit();
            }
        }
        if (etid == 0) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Forward, current generation (Lp <= 256): one work unit = one (sequence, head); its ceil(L/128) query tiles share
// one K / V load.  Differences from the kernel above:
//   * two S accumulators in TMEM (256 columns each) and two softmax warp groups: tile t lives in buffer t & 1, so the
//     tensor pipe computes S of the next tile / PV of the previous one while a group is in its MUFU-bound softmax;
//   * a group is 8 warps = 2 per TMEM lane quarter; the pair splits the key columns of a row and exchanges the partial
//     row maxima through shared memory, so 4 warps per scheduler keep the MUFU pipe busy;
//   * P never touches shared memory: it is written back in place over S as packed fp16 (tcgen05.st) and consumed by
//     the PV MMA as a tensor-memory A operand; O accumulates in free columns of the same buffer;
//   * the row sum is a 16-column MMA of P against a tile of ones, i.e. the sum of exactly the
//     fp16 values PV uses (ex2.approx.f16x2 was tried: it lowers to two MUFU.EX2.F16, no MUFU saving);
//   * Q / K / V of the next unit stream into a second shared-memory stage while this unit computes.
// 576 threads: TMA warp, MMA warp, 2 x 8 softmax/epilogue warps; all 512 TMEM columns, one CTA per SM.
struct FmhaFwd2Params {
    // L rows per unit.  seq_len < L: the unit PACKS L / seq_len short sequences (block-diagonal mask: a row only sees
    // the keys of its own sequence); seq_len == L: one sequence per unit.
    int L, Lp, heads, QT, num_units, causal, seq_len;
    float scale_log2e, scale;
    float* lse;  // [N, heads, L]
    uint32_t stage_bytes, off_k, off_v, off_o, off_ones, off_max, off_bar;
    // TMEM columns inside a 256-column buffer.  The two warps of a lane quarter split the key columns at `csplit`;
    // each writes its fp16 P in place over ITS OWN part of S (keys [0, csplit) -> columns [0, csplit/2), keys
    // [csplit, Lp) -> columns [csplit, csplit + (Lp-csplit)/2)), so neither overwrites scores the other still reads.
    // O (64 columns) and the row sum (16 columns) go where no P lives.
    int csplit, o_col, sum_col;
};

constexpr int kFmhaFwd2Threads = 576;
constexpr int kFmhaFwd2Group = 256;  // threads of one softmax group

template <bool PACKED>
__global__ void __launch_bounds__(kFmhaFwd2Threads, 1)
fmha_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const __grid_constant__ CUtensorMap tmap_out, const FmhaFwd2Params p) {
    extern __shared__ __align__(1024) uint8_t smem_fwd2[];
    uint8_t* smem = smem_fwd2;
    uint8_t* sOnes = smem + p.off_ones;                          // [16 x 64] fp16 ones (B operand of the row-sum MMA)
    float* sMax = reinterpret_cast<float*>(smem + p.off_max);    // [2 groups][2 halves][128 rows] partial row maxima
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
    uint64_t* qk_full = bars + 0;     // [2] per shared-memory stage
    uint64_t* v_full = bars + 2;      // [2]
    uint64_t* stage_free = bars + 4;  // [2] every MMA reading the stage has completed
    uint64_t* s_full = bars + 6;      // [2] per TMEM buffer
    uint64_t* p_full = bars + 8;      // [2]
    uint64_t* o_full = bars + 10;     // [2]
    uint64_t* o_read = bars + 12;     // [2] O has left TMEM: the buffer may be overwritten
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Lp = p.Lp, d = p.heads * 64, QT = p.QT;
    const int my_units = (int)blockIdx.x < p.num_units ? (p.num_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int T = my_units * QT;  // tiles this CTA processes; tile t = (unit t / QT, query tile t % QT)

    if (warp == 0 && lane == 0) {
        if (smem_u32(smem) & 1023u) {
            printf("mvlpt: fmha_fwd dynamic shared memory is not 1024-byte aligned\n");
            __trap();
        }
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        tma_prefetch_desc(&tmap_out);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&qk_full[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&stage_free[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], kFmhaFwd2Group);
            mbar_init(&o_full[i], 1);
            mbar_init(&o_read[i], kFmhaFwd2Group);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    if (warp >= 2 && threadIdx.x - 64 < 128)  // 2 KB of fp16 1.0
        reinterpret_cast<uint4*>(sOnes)[threadIdx.x - 64] = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (lane == 0) {
            for (int u = 0; u < my_units; ++u) {
                const int unit = blockIdx.x + u * gridDim.x;
                const int h = unit % p.heads, n = unit / p.heads;
                const int s = u & 1;
                uint8_t* st = smem + s * p.stage_bytes;
                if (u >= 2) mbar_wait(&stage_free[s], (uint32_t)((u >> 1) - 1) & 1);
                mbar_arrive_expect_tx(&qk_full[s], (uint32_t)QT * 16384u + (uint32_t)Lp * 128u);
                for (int qt = 0; qt < QT; ++qt) tma_load_3d(st + qt * 16384, &tmap_q, &qk_full[s], h * 64, qt * 128, n);
                tma_load_3d(st + p.off_k, &tmap_kv, &qk_full[s], d + h * 64, 0, n);
                mbar_arrive_expect_tx(&v_full[s], (uint32_t)Lp * 128u);
                tma_load_3d(st + p.off_v, &tmap_kv, &v_full[s], 2 * d + h * 64, 0, n);
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        // The whole warp runs the control flow (warp-uniform operands stay in uniform registers); one elected lane
        // issues the MMAs and commits.
        const bool leader = elect_one();
        const uint32_t idesc_s = umma_idesc_f16(128, (uint32_t)Lp, 0, 0);
        const uint32_t idesc_pv = umma_idesc_f16(128, 64, 0, 1);  // B (V) is MN-major
        const uint32_t idesc_sum = umma_idesc_f16(128, 16, 0, 0);
        const uint64_t ones_desc = umma_desc_k_sw128(smem_u32(sOnes));
        const uint64_t stage_k0 = umma_desc_k_sw128(smem_u32(smem));  // stage s: + s * (stage_bytes >> 4)
        const uint64_t stage_v0 = umma_desc_mn_sw128(smem_u32(smem + p.off_v), 1024);
        const int ksteps = Lp / 16, ks0 = p.csplit / 16;  // steps [0, ks0) read P of half 0, the rest P of half 1
        auto issue_s = [&](int t) {
            const int u = t / QT, qt = t - u * QT, s = u & 1, b = t & 1;
            if (qt == 0) mbar_wait(&qk_full[s], (uint32_t)(u >> 1) & 1);
            if (t >= 2) mbar_wait(&o_read[b], (uint32_t)((t >> 1) - 1) & 1);  // O of tile t-2 left the buffer
            tc_fence_after();
            const uint64_t st_desc = stage_k0 + (uint32_t)s * (p.stage_bytes >> 4);
            const uint64_t q_desc = st_desc + (uint32_t)(qt * (16384 >> 4));
            const uint64_t k_desc = st_desc + (p.off_k >> 4);
            if (leader) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16_ss(tmem_base + b * 256, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
                umma_commit(&s_full[b]);
            }
            __syncwarp();
        };
        if (T > 0) issue_s(0);
        for (int t = 0; t < T; ++t) {
            if (t + 1 < T) issue_s(t + 1);
            const int u = t / QT, qt = t - u * QT, s = u & 1, b = t & 1;
            mbar_wait(&p_full[b], (uint32_t)(t >> 1) & 1);
            if (qt == 0) mbar_wait(&v_full[s], (uint32_t)(u >> 1) & 1);
            tc_fence_after();
            const uint32_t tb = tmem_base + b * 256;
            const uint32_t o_acc = tb + p.o_col, sum_acc = tb + p.sum_col;
            const uint32_t pa1 = tb + p.csplit - (p.csplit >> 1);  // + 8 kk = P columns of step kk >= ks0
            const uint64_t v_desc = stage_v0 + (uint32_t)s * (p.stage_bytes >> 4);
            if (leader) {
                for (int kk = 0; kk < ks0; ++kk) {
                    umma_f16_ts(o_acc, tb + kk * 8, v_desc + kk * (2048 >> 4), idesc_pv, kk != 0);
                    umma_f16_ts(sum_acc, tb + kk * 8, ones_desc, idesc_sum, kk != 0);
                }
                for (int kk = ks0; kk < ksteps; ++kk) {
                    umma_f16_ts(o_acc, pa1 + kk * 8, v_desc + kk * (2048 >> 4), idesc_pv, 1);
                    umma_f16_ts(sum_acc, pa1 + kk * 8, ones_desc, idesc_sum, 1);
                }
                umma_commit(&o_full[b]);
                if (qt == QT - 1) umma_commit(&stage_free[s]);
            }
            __syncwarp();
        }
    } else {
        // ============================== softmax + epilogue: group g owns TMEM buffer g ==============================
        const int quarter = warp & 3;           // TMEM lane quarter this warp may access
        const int g = (warp - 2) >> 3;          // warps 2..9 -> group 0, 10..17 -> group 1
        const int half = ((warp - 2) >> 2) & 1; // which part of the key columns / of the O columns
        const int r = quarter * 32 + lane;      // row inside the tile == TMEM lane
        const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + g * 256;
        const bool leader = ((warp - 2) & 7) == 0 && lane == 0;  // issues this group's TMA stores
        uint8_t* sO = smem + p.off_o + g * 16384;
        uint8_t* orow = sO + (r >> 3) * 1024 + (r & 7) * 128;
        float* my_max = sMax + (g * 2 + half) * 128 + r;
        const float* other_max = sMax + (g * 2 + (half ^ 1)) * 128 + r;
        const int sw = r & 7;
        const float sl2 = p.scale_log2e;
        const int csplit = p.csplit;  // key columns [0, csplit) -> half 0, [csplit, Lp) -> half 1
        const int cb = half ? csplit : 0, ce = half ? Lp : csplit;
        const uint32_t p_row = t_row + (half ? csplit - (csplit >> 1) : 0);  // + (c0 >> 1) = where P of key c0 goes
        for (int t = g; t < T; t += 2) {
            const int u = t / QT, qt = t - u * QT;
            const int unit = blockIdx.x + u * gridDim.x;
            const int h = unit % p.heads, n = unit / p.heads;
            const uint32_t ph = (uint32_t)(t >> 1) & 1;
            const int q = qt * 128 + r;
            const bool warp_live = (qt * 128 + quarter * 32) < p.L;  // warp-uniform, same for both halves of a quarter
            // valid keys of this row: [lo, lim) — its own sequence, up to itself when causal.  Without packing lo is the
            // constant 0 and, for the image tower, lim is warp-uniform: the compiler keeps the branches below uniform.
            const int lo = PACKED ? (q / p.seq_len) * p.seq_len : 0;
            const int hi0 = p.causal ? q + 1 : (PACKED ? lo + p.seq_len : p.L);
            const int lim = hi0 < p.L ? hi0 : p.L;
            mbar_wait(&s_full[g], ph);
            tc_fence_after();
            float m = -INFINITY;
            if (warp_live) {
                // pass 1: max of the raw scores over this warp's share of the valid keys
                for (int c0 = cb; c0 < ce; c0 += 16) {
                    uint32_t raw[16];
                    tmem_ld_32x32b_x16(t_row + c0, raw);
                    tmem_ld_wait();
                    if (c0 >= lo && c0 + 16 <= lim) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) m = fmaxf(m, __uint_as_float(raw[j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j >= lo && c0 + j < lim) m = fmaxf(m, __uint_as_float(raw[j]));
                    }
                }
                *my_max = m;
            }
            named_bar_sync(1 + g, kFmhaFwd2Group);  // also orders the previous tile's staging reads (see below)
            if (warp_live) {
                m = fmaxf(m, *other_max);
                if (m == -INFINITY) m = 0.f;  // rows >= L (never stored)
                const float m2 = m * sl2;
                // pass 2: p = exp2(s*scale*log2e - m2) as packed fp16, back into TMEM over S
                for (int c0 = cb; c0 < ce; c0 += 16) {
                    uint32_t raw[16];
                    tmem_ld_32x32b_x16(t_row + c0, raw);
                    tmem_ld_wait();
                    uint32_t pk[8];
                    if (c0 >= lo && c0 + 16 <= lim) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            pk[j] = pack_half2(ex2_approx(fmaf(__uint_as_float(raw[2 * j]), sl2, -m2)),
                                               ex2_approx(fmaf(__uint_as_float(raw[2 * j + 1]), sl2, -m2)));
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int k0 = c0 + 2 * j, k1 = k0 + 1;
                            const float e0 = (k0 >= lo && k0 < lim) ? ex2_approx(fmaf(__uint_as_float(raw[2 * j]), sl2, -m2)) : 0.f;
                            const float e1 = (k1 >= lo && k1 < lim) ? ex2_approx(fmaf(__uint_as_float(raw[2 * j + 1]), sl2, -m2)) : 0.f;
                            pk[j] = pack_half2(e0, e1);
                        }
                    }
                    tmem_st_32x32b_x8(p_row + (c0 >> 1), pk);
                }
                tmem_st_wait();
            }
            tc_fence_before();
            mbar_arrive(&p_full[g]);

            mbar_wait(&o_full[g], ph);
            tc_fence_after();
            uint32_t o[32], sm[16];
            tmem_ld_32x32(t_row + p.o_col + 32 * half, o);
            tmem_ld_32x32b_x16(t_row + p.sum_col, sm);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&o_read[g]);
            const float sum = __uint_as_float(sm[0]);
            const float inv_sum = 1.f / sum;
            if (half == 0 && warp_live && q < p.L) {
                const int per = PACKED ? p.L / p.seq_len : 1, sq = PACKED ? q / p.seq_len : 0;  // sequences per unit, row's sequence
                p.lse[((size_t)(n * per + sq) * p.heads + h) * p.seq_len + (q - sq * p.seq_len)] = m * p.scale + __logf(sum);
            }
            // this group's previous TMA store must have finished reading the staging tile
            if (leader) tma_store_wait_read<0>();
            named_bar_sync(1 + g, kFmhaFwd2Group);
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
                const uint32_t* s = o + uu * 8;
                *reinterpret_cast<uint4*>(orow + (((uu + 4 * half) ^ sw) << 4)) = make_uint4(
                    pack_half2(__uint_as_float(s[0]) * inv_sum, __uint_as_float(s[1]) * inv_sum),
                    pack_half2(__uint_as_float(s[2]) * inv_sum, __uint_as_float(s[3]) * inv_sum),
                    pack_half2(__uint_as_float(s[4]) * inv_sum, __uint_as_float(s[5]) * inv_sum),
                    pack_half2(__uint_as_float(s[6]) * inv_sum, __uint_as_float(s[7]) * inv_sum));
            }
            fence_proxy_async_smem();
            named_bar_sync(1 + g, kFmhaFwd2Group);
            if (leader) {
                tma_store_3d(&tmap_out, sO, h * 64, qt * 128, n);
                tma_store_commit();
            }
        }
        if (leader) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// seq_len > 0: every unit packs L / seq_len sequences of seq_len rows (N counts the packed groups)
inline int fmha_fwd2_sm100(const void* qkv, void* out, void* lse, int N, int L, int d, int heads, int causal,
                           cudaStream_t stream, int seq_len = 0) {
    const int Lp = (L + 15) / 16 * 16;  // <= 256: K / V arrive in one TMA box each
    CUtensorMap tq, tkv, to;
    {
        uint64_t dims[3] = {(uint64_t)3 * d, (uint64_t)L, (uint64_t)N};
        uint64_t str[2] = {(uint64_t)3 * d * 2, (uint64_t)L * 3 * d * 2};
        uint32_t box_q[3] = {64u, 128u, 1u};
        uint32_t box_kv[3] = {64u, (uint32_t)Lp, 1u};
        int rc = make_tmap_f16(&tq, qkv, 3, dims, str, box_q);
        if (rc) return rc;
        rc = make_tmap_f16(&tkv, qkv, 3, dims, str, box_kv);
        if (rc) return rc;
    }
    {
        uint64_t dims[3] = {(uint64_t)d, (uint64_t)L, (uint64_t)N};
        uint64_t str[2] = {(uint64_t)d * 2, (uint64_t)L * d * 2};
        uint32_t box[3] = {64u, 128u, 1u};
        int rc = make_tmap_f16(&to, out, 3, dims, str, box);
        if (rc) return rc;
    }
    FmhaFwd2Params p;
    p.L = L;
    p.Lp = Lp;
    p.heads = heads;
    p.QT = (L + 127) / 128;
    p.num_units = N * heads;
    p.causal = causal;
    p.seq_len = seq_len > 0 ? seq_len : L;
    p.scale = 0.125f;
    p.scale_log2e = 0.125f * 1.4426950408889634f;
    p.lse = static_cast<float*>(lse);
    const uint32_t kv = (uint32_t)((Lp * 128 + 1023) / 1024 * 1024);
    p.off_k = (uint32_t)p.QT * 16384u;
    p.off_v = p.off_k + kv;
    p.stage_bytes = p.off_v + kv;
    p.off_o = 2u * p.stage_bytes;
    p.off_ones = p.off_o + 32768u;
    p.off_max = p.off_ones + 2048u;
    p.off_bar = p.off_max + 2048u;
    {
        p.csplit = ((Lp >> 4) + 1) / 2 * 16;
        const int p1_end = p.csplit + (Lp - p.csplit) / 2;
        const int tail0 = (p1_end + 31) / 32 * 32, gap0 = (p.csplit / 2 + 31) / 32 * 32;
        const int tail = 256 - tail0, gap = p.csplit - gap0;
        if (tail >= 80) { p.o_col = tail0; p.sum_col = tail0 + 64; }
        else if (tail >= 64 && gap >= 16) { p.o_col = tail0; p.sum_col = gap0; }
        else if (gap >= 64 && tail >= 16) { p.o_col = gap0; p.sum_col = tail0; }
        else if (gap >= 80) { p.o_col = gap0; p.sum_col = gap0 + 64; }
        else return fail(MVLPT_ESHAPE, "fmha_fwd2_sm100: no TMEM room for O at L=%d", L);
    }
    const size_t smem = (size_t)p.off_bar + 128;
    if (smem > 227 * 1024) return fail(MVLPT_ESHAPE, "fmha_fwd2_sm100: L=%d needs %zu bytes of shared memory", L, smem);
    static DynSmemCache attr_plain, attr_packed;
    if (int rc = ensure_dyn_smem(fmha_fwd_tc2_kernel<false>, smem, attr_plain)) return rc;
    if (int rc = ensure_dyn_smem(fmha_fwd_tc2_kernel<true>, smem, attr_packed)) return rc;
    const int grid = p.num_units < sm_count() ? p.num_units : sm_count();
    if (p.seq_len != p.L)
        MVLPT_CUDA_OK(launch_pdl(fmha_fwd_tc2_kernel<true>, dim3(grid), dim3(kFmhaFwd2Threads), smem, stream, 1, tq, tkv, to, p));
    else
        MVLPT_CUDA_OK(launch_pdl(fmha_fwd_tc2_kernel<false>, dim3(grid), dim3(kFmhaFwd2Threads), smem, stream, 1, tq, tkv, to, p));
    return launched("fmha_fwd_tc2");
}

inline bool fmha_sm100_supported(int L) { return L >= 1 && L <= 272; }

inline int fmha_fwd_sm100(const void* qkv, void* out, void* lse, int N, int L, int d, int heads, int causal,
                          cudaStream_t stream) {
    const int Lp = (L + 15) / 16 * 16;
    // short sequences (text prompts cut at the last EOT, CoCoOp's B*C prompts): G = 128 / L of them share one 128-row tile
    // under a block-diagonal mask.  Whole groups first, then the N % G left-over sequences as one smaller group.
    const char* pack_env = getenv("MVLPT_FMHA_PACK");  // read per call: tests switch it inside one process
    const bool pack = !(pack_env && pack_env[0] == '0');
    if (pack && L <= 64 && N >= 2 && !getenv("MVLPT_FMHA_FWD_V1")) {
        const int G = 128 / L, groups = N / G, rem = N - groups * G;
        if (groups > 0) {
            const int rc = fmha_fwd2_sm100(qkv, out, lse, groups, G * L, d, heads, causal, stream, L);
            if (rc) return rc;
        }
        if (rem > 0) {
            const size_t row0 = (size_t)groups * G * L;
            return fmha_fwd2_sm100(static_cast<const __half*>(qkv) + row0 * 3 * d, static_cast<__half*>(out) + row0 * d,
                                   static_cast<float*>(lse) + (size_t)groups * G * heads * L, 1, rem * L, d, heads, causal,
                                   stream, L);
        }
        return MVLPT_OK;
    }
    // two stages of Q/K/V fit shared memory up to 240 keys
    if (Lp <= 240 && !getenv("MVLPT_FMHA_FWD_V1")) return fmha_fwd2_sm100(qkv, out, lse, N, L, d, heads, causal, stream);
    const int nbox = (Lp + 255) / 256;  // K / V arrive in nbox TMA boxes of Lp/nbox rows (a multiple of 8)
    const int box_h = Lp / nbox;
    if (box_h * nbox != Lp || (box_h % 8)) return fail(MVLPT_ESHAPE, "fmha_fwd_sm100: unsupported L=%d", L);
    CUtensorMap tq, tkv, to;
    {
        uint64_t dims[3] = {(uint64_t)3 * d, (uint64_t)L, (uint64_t)N};
        uint64_t str[2] = {(uint64_t)3 * d * 2, (uint64_t)L * 3 * d * 2};
        uint32_t box_q[3] = {64u, 128u, 1u};
        uint32_t box_kv[3] = {64u, (uint32_t)box_h, 1u};
        int rc = make_tmap_f16(&tq, qkv, 3, dims, str, box_q);
        if (rc) return rc;
        rc = make_tmap_f16(&tkv, qkv, 3, dims, str, box_kv);
        if (rc) return rc;
    }
    {
        uint64_t dims[3] = {(uint64_t)d, (uint64_t)L, (uint64_t)N};
        uint64_t str[2] = {(uint64_t)d * 2, (uint64_t)L * d * 2};
        uint32_t box[3] = {64u, 128u, 1u};
        int rc = make_tmap_f16(&to, out, 3, dims, str, box);
        if (rc) return rc;
    }
    FmhaFwdParams p;
    p.L = L;
    p.Lp = Lp;
    p.heads = heads;
    p.q_tiles = (L + 127) / 128;
    p.num_tiles = N * heads * p.q_tiles;
    p.causal = causal;
    p.scale = 0.125f;
    p.scale_log2e = 0.125f * 1.4426950408889634f;
    p.lse = static_cast<float*>(lse);
    p.box_h = box_h;
    p.tmem_cols = Lp <= 64 ? 64 : Lp <= 128 ? 128 : Lp <= 256 ? 256 : 512;
    const uint32_t qk_bytes = 16384u + (uint32_t)Lp * 128u;
    const uint32_t p_bytes = (uint32_t)((Lp + 63) / 64) * 16384u;
    p.off_v = qk_bytes > p_bytes ? qk_bytes : p_bytes;
    p.off_o = p.off_v + (uint32_t)Lp * 128u;
    p.off_bar = p.off_o + 16384u;
    const size_t smem = (size_t)p.off_bar + 128 + 1024;
    static DynSmemCache attr;
    if (int rc = ensure_dyn_smem(fmha_fwd_tc_kernel, smem, attr)) return rc;
    const int per_sm = (smem * 2 <= 227 * 1024 && p.tmem_cols <= 256) ? 2 : 1;
    const int max_ctas = sm_count() * per_sm;
    const int grid = p.num_tiles < max_ctas ? p.num_tiles : max_ctas;
    fmha_fwd_tc_kernel<<<grid, kFmhaThreads, smem, stream>>>(tq, tkv, to, p);
    return launched("fmha_fwd_tc");
}

}  // namespace mvlpt
