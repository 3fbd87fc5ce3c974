// tcgen05 attention BACKWARD, head width 64, L <= 256 (autograd of F.multi_head_attention_forward as
// called at clip/model.py:181-183).  One work unit = one (sequence n, head h); the whole head lives on one SM:
//
//   TMA   : Q, dO (ceil(L/128) tiles of [128 x 64]) and K, V (same number of 128-row chunks) out of the packed
//           [N, L, 3d] / [N, L, d] tensors through 3-D maps (rows >= L zero-filled).  Every tile is a stack of
//           128-byte rows with the 128B swizzle, which serves BOTH operand majors: K-major when the contraction runs
//           over the head width, MN-major when it runs over the rows — so no transposed copy of anything is ever made.
//           The loads of a unit form two groups (tile/chunk 0 and tile/chunk 1), each with its own full/free
//           barriers: group 0 of the NEXT unit streams in while the last iteration of this unit still computes.
//   loop key chunk kc (128 keys) x query tile qt (128 queries):
//     MMA : S  = Q_qt . K_kc^T      -> TMEM [  0,128)         dP = dO_qt . V_kc^T  -> TMEM [128,256)
//     thr : 8 row warps = 2 per TMEM lane quarter; the pair splits the key columns of the chunk.  Per query row:
//           P = exp2(S*scale*log2e - lse*log2e),  dS = P o (dP - D),  D = rowsum(dO o O);
//           P and dS (fp16) -> swizzled shared memory
//     MMA : dV_kc += P^T . dO_qt   -> TMEM [256,320)    (A = P  read MN-major, contraction over queries)
//           dK_kc += dS^T . Q_qt   -> TMEM [320,384)    (A = dS read MN-major)
//           dQ_qt += dS . K_kc     -> TMEM [384+64 qt, ..)   (A = dS read K-major, B = K read MN-major)
//   after the last qt of a chunk: dK_kc * hd^-1/2, dV_kc -> fp16 -> staging -> TMA store; after the last chunk: dQ.
//
// 320 threads: TMA warp, MMA warp, 8 row warps.  TMEM: all 512 columns, hence one CTA per SM.
#pragma once
#include "fmha_sm100.cuh"

namespace mvlpt {

struct FmhaBwdParams {
    int L, Lp, heads, d, QT, KC, causal, num_units, box_h;
    float scale_log2e;   // hd^-1/2 * log2(e)
    const float* lse;    // [N, heads, L]
    const __half* o;     // [N, L, d]
    const __half* d_o;   // [N, L, d]
    uint32_t off_do, off_k, off_v, off_p, off_ds, off_stage, off_row, off_bar;
};

constexpr int kFmhaBwdThreads = 320;
constexpr int kFmhaBwdRowThreads = 256;

__global__ void __launch_bounds__(kFmhaBwdThreads, 1)
fmha_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                   const __grid_constant__ CUtensorMap tmap_do, const __grid_constant__ CUtensorMap tmap_dqkv,
                   const FmhaBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_bwd[];
    uint8_t* smem = smem_bwd;
    uint8_t* sQ = smem;                 // QT x [128 x 64]
    uint8_t* sdO = smem + p.off_do;     // QT x [128 x 64]
    uint8_t* sK = smem + p.off_k;       // KC x [box_h x 64]
    uint8_t* sV = smem + p.off_v;       // KC x [box_h x 64]
    uint8_t* sP = smem + p.off_p;       // 2 chunks (64 keys each) of [128 queries x 128 B]
    uint8_t* sdS = smem + p.off_ds;     // same shape
    uint8_t* sSt = smem + p.off_stage;  // 2 x [128 x 64] staging for the TMA stores
    float* sD = reinterpret_cast<float*>(smem + p.off_row);  // [256] rowsum(dO o O)
    float* sL2 = sD + 256;                                   // [256] lse * log2(e)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
    uint64_t* ld_full = bars + 0;  // [2] load group g landed
    uint64_t* g_free = bars + 2;   // [2] every MMA reading load group g of this unit has completed
    uint64_t* s_full = bars + 4;
    uint64_t* p_ready = bars + 5;
    uint64_t* acc_full = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Lp = p.Lp, d = p.d, QT = p.QT, KC = p.KC;
    const uint32_t kv_chunk = (uint32_t)p.box_h * 128u;  // bytes of one K (or V) chunk in shared memory
    constexpr uint32_t kColS = 0, kColDP = 128, kColDV = 256, kColDK = 320, kColDQ = 384;

    if (warp == 0 && lane == 0) {
        if (smem_u32(smem) & 1023u) {
            printf("mvlpt: fmha_bwd dynamic shared memory is not 1024-byte aligned\n");
            __trap();
        }
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        tma_prefetch_desc(&tmap_do);
        tma_prefetch_desc(&tmap_dqkv);
        for (int g = 0; g < 2; ++g) {
            mbar_init(&ld_full[g], 1);
            mbar_init(&g_free[g], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(p_ready, kFmhaBwdRowThreads);
        mbar_init(acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
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
            for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++it) {
                const int h = unit % p.heads, n = unit / p.heads;
                for (int g = 0; g < QT; ++g) {  // QT == KC: group g = {Q_g, dO_g, K_g, V_g}
                    if (it > 0) mbar_wait(&g_free[g], (uint32_t)(it - 1) & 1);
                    mbar_arrive_expect_tx(&ld_full[g], 32768u + 2u * kv_chunk);
                    tma_load_3d(sQ + g * 16384, &tmap_q, &ld_full[g], h * 64, g * 128, n);
                    tma_load_3d(sdO + g * 16384, &tmap_do, &ld_full[g], h * 64, g * 128, n);
                    tma_load_3d(sK + g * kv_chunk, &tmap_kv, &ld_full[g], d + h * 64, g * 128, n);
                    tma_load_3d(sV + g * kv_chunk, &tmap_kv, &ld_full[g], 2 * d + h * 64, g * 128, n);
                }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        if (lane == 0) {
            int it = 0;
            uint32_t gi = 0;  // iterations issued so far (parity of p_ready)
            auto issue_s_dp = [&](int kc, int qt) {
                const int nk = (Lp - kc * 128) < 128 ? (Lp - kc * 128) : 128;
                const uint32_t idesc = umma_idesc_f16(128, (uint32_t)nk, 0, 0);
                const uint64_t q_desc = umma_desc_k_sw128(smem_u32(sQ + qt * 16384));
                const uint64_t k_desc = umma_desc_k_sw128(smem_u32(sK + kc * kv_chunk));
                const uint64_t do_desc = umma_desc_k_sw128(smem_u32(sdO + qt * 16384));
                const uint64_t v_desc = umma_desc_k_sw128(smem_u32(sV + kc * kv_chunk));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + kColS, q_desc + 2 * k, k_desc + 2 * k, idesc, k != 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + kColDP, do_desc + 2 * k, v_desc + 2 * k, idesc, k != 0);
                umma_commit(s_full);
            };
            for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++it) {
                mbar_wait(&ld_full[0], (uint32_t)it & 1);
                tc_fence_after();
                issue_s_dp(0, 0);
                for (int kc = 0; kc < KC; ++kc) {
                    const int nk = (Lp - kc * 128) < 128 ? (Lp - kc * 128) : 128;
                    for (int qt = 0; qt < QT; ++qt) {
                        const int nq = (Lp - qt * 128) < 128 ? (Lp - qt * 128) : 128;  // live query rows (multiple of 16)
                        mbar_wait(p_ready, gi & 1);
                        ++gi;
                        tc_fence_after();
                        // dV_kc (+)= P^T . dO_qt ; dK_kc (+)= dS^T . Q_qt      contraction over the nq queries
                        const uint32_t idesc_t = umma_idesc_f16(128, 64, 1, 1);
                        for (int kk = 0; kk < nq / 16; ++kk) {
                            const uint64_t a_p = umma_desc_mn_sw128(smem_u32(sP + kk * 2048), 16384);
                            const uint64_t b_do = umma_desc_mn_sw128(smem_u32(sdO + qt * 16384 + kk * 2048), 1024);
                            umma_f16_ss(tmem_base + kColDV, a_p, b_do, idesc_t, (qt | kk) != 0);
                        }
                        for (int kk = 0; kk < nq / 16; ++kk) {
                            const uint64_t a_ds = umma_desc_mn_sw128(smem_u32(sdS + kk * 2048), 16384);
                            const uint64_t b_q = umma_desc_mn_sw128(smem_u32(sQ + qt * 16384 + kk * 2048), 1024);
                            umma_f16_ss(tmem_base + kColDK, a_ds, b_q, idesc_t, (qt | kk) != 0);
                        }
                        // dQ_qt (+)= dS . K_kc      contraction over the nk keys
                        const uint32_t idesc_q = umma_idesc_f16(128, 64, 0, 1);
                        for (int kk = 0; kk < nk / 16; ++kk) {
                            const uint64_t a_ds = umma_desc_k_sw128(smem_u32(sdS + (kk >> 2) * 16384 + (kk & 3) * 32));
                            const uint64_t b_k = umma_desc_mn_sw128(smem_u32(sK + kc * kv_chunk + kk * 2048), 1024);
                            umma_f16_ss(tmem_base + kColDQ + qt * 64, a_ds, b_k, idesc_q, (kc | kk) != 0);
                        }
                        if (qt == QT - 1) umma_commit(acc_full);
                        // load group 0 (Q_0, dO_0, K_0, V_0) is last read by iteration (KC-1, 0); group 1 by the last one
                        if (kc == KC - 1 && qt == 0) umma_commit(&g_free[0]);
                        if (kc == KC - 1 && qt == 1) umma_commit(&g_free[1]);
                        const int nqt = (qt + 1 == QT) ? 0 : qt + 1;
                        const int nkc = (qt + 1 == QT) ? kc + 1 : kc;
                        if (nkc < KC) {
                            if (kc == 0 && qt == 0) {  // first use of load group 1 (only reachable when QT == KC == 2)
                                mbar_wait(&ld_full[1], (uint32_t)it & 1);
                                tc_fence_after();
                            }
                            issue_s_dp(nkc, nqt);
                        }
                    }
                }
            }
        }
    } else {
        // ============================== row warps: softmax backward + epilogues ==============================
        const int quarter = warp & 3;        // TMEM lane quarter this warp may access
        const int group = (warp - 2) >> 2;   // the two warps of a quarter split the key columns / output columns
        const int r = quarter * 32 + lane;
        const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16);
        const int etid = threadIdx.x - 64;
        const int sw = r & 7;
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        uint32_t gi = 0, gacc = 0;

        // fp32 TMEM row, 32 columns [col + 32*group, ..) * mul -> fp16 -> units 4*group.. of row r of a swizzled
        // [128 x 64] staging tile
        auto stage_half = [&](uint32_t col, float mul, uint8_t* tile) {
            uint32_t a[32];
            tmem_ld_32x32(t_row + col + 32 * group, a);
            tmem_ld_wait();
            uint8_t* orow = tile + row_off;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t* s = a + u * 8;
                *reinterpret_cast<uint4*>(orow + (((u + 4 * group) ^ sw) << 4)) = make_uint4(
                    pack_half2(__uint_as_float(s[0]) * mul, __uint_as_float(s[1]) * mul),
                    pack_half2(__uint_as_float(s[2]) * mul, __uint_as_float(s[3]) * mul),
                    pack_half2(__uint_as_float(s[4]) * mul, __uint_as_float(s[5]) * mul),
                    pack_half2(__uint_as_float(s[6]) * mul, __uint_as_float(s[7]) * mul));
            }
        };

        for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
            const int h = unit % p.heads, n = unit / p.heads;
            // D = rowsum(dO o O) and lse*log2e: warp group g computes the rows of query tile g, shared through smem
            if (group < QT) {
                const int q = group * 128 + r;
                float acc = 0.f, l2 = 0.f;
                if (q < p.L) {
                    const uint4* po = reinterpret_cast<const uint4*>(p.o + ((size_t)n * p.L + q) * d + h * 64);
                    const uint4* pd = reinterpret_cast<const uint4*>(p.d_o + ((size_t)n * p.L + q) * d + h * 64);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint4 a = __ldg(po + i), b = __ldg(pd + i);
                        const __half2* ha = reinterpret_cast<const __half2*>(&a);
                        const __half2* hb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 fa = __half22float2(ha[j]), fb = __half22float2(hb[j]);
                            acc = fmaf(fa.x, fb.x, acc);
                            acc = fmaf(fa.y, fb.y, acc);
                        }
                    }
                    l2 = p.lse[((size_t)n * p.heads + h) * p.L + q] * 1.4426950408889634f;
                }
                sD[q] = acc;
                sL2[q] = l2;
            }
            named_bar_sync(1, kFmhaBwdRowThreads);
            float Dv[2], lse2[2];
            Dv[0] = sD[r];
            lse2[0] = sL2[r];
            Dv[1] = QT > 1 ? sD[128 + r] : 0.f;
            lse2[1] = QT > 1 ? sL2[128 + r] : 0.f;

            for (int kc = 0; kc < KC; ++kc) {
                const int nk = (Lp - kc * 128) < 128 ? (Lp - kc * 128) : 128;
                const int half0 = ((nk >> 4) + 1) / 2 * 16;  // columns [0, half0) -> group 0, [half0, nk) -> group 1
                const int cb = group ? half0 : 0, ce = group ? nk : half0;
                for (int qt = 0; qt < QT; ++qt) {
                    const int q = qt * 128 + r;
                    const bool qok = q < p.L;
                    int kmax = p.causal ? (q + 1 < p.L ? q + 1 : p.L) : p.L;  // valid keys: [0, kmax)
                    if (!qok) kmax = 0;
                    const float Dq = Dv[qt], l2 = lse2[qt];
                    mbar_wait(s_full, gi & 1);
                    ++gi;
                    tc_fence_after();
                    for (int c0 = cb; c0 < ce; c0 += 16) {
                        uint32_t s[16], dp[16];
                        tmem_ld_32x32b_x16(t_row + kColS + c0, s);
                        tmem_ld_32x32b_x16(t_row + kColDP + c0, dp);
                        tmem_ld_wait();
                        float pv[16], dsv[16];
                        const int key0 = kc * 128 + c0;
                        if (key0 + 16 <= kmax) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                pv[j] = ex2_approx(fmaf(__uint_as_float(s[j]), p.scale_log2e, -l2));
                                dsv[j] = pv[j] * (__uint_as_float(dp[j]) - Dq);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const bool ok = key0 + j < kmax;
                                pv[j] = ok ? ex2_approx(fmaf(__uint_as_float(s[j]), p.scale_log2e, -l2)) : 0.f;
                                dsv[j] = ok ? pv[j] * (__uint_as_float(dp[j]) - Dq) : 0.f;
                            }
                        }
                        const uint32_t off = (uint32_t)(c0 >> 6) * 16384u + row_off;
                        const int u = (c0 & 63) >> 3;
                        *reinterpret_cast<uint4*>(sP + off + ((u ^ sw) << 4)) = make_uint4(
                            pack_half2(pv[0], pv[1]), pack_half2(pv[2], pv[3]), pack_half2(pv[4], pv[5]), pack_half2(pv[6], pv[7]));
                        *reinterpret_cast<uint4*>(sP + off + (((u + 1) ^ sw) << 4)) = make_uint4(
                            pack_half2(pv[8], pv[9]), pack_half2(pv[10], pv[11]), pack_half2(pv[12], pv[13]),
                            pack_half2(pv[14], pv[15]));
                        *reinterpret_cast<uint4*>(sdS + off + ((u ^ sw) << 4)) = make_uint4(
                            pack_half2(dsv[0], dsv[1]), pack_half2(dsv[2], dsv[3]), pack_half2(dsv[4], dsv[5]),
                            pack_half2(dsv[6], dsv[7]));
                        *reinterpret_cast<uint4*>(sdS + off + (((u + 1) ^ sw) << 4)) = make_uint4(
                            pack_half2(dsv[8], dsv[9]), pack_half2(dsv[10], dsv[11]), pack_half2(dsv[12], dsv[13]),
                            pack_half2(dsv[14], dsv[15]));
                    }
                    fence_proxy_async_smem();
                    tc_fence_before();
                    mbar_arrive(p_ready);

                    if (qt == QT - 1) {
                        // dK_kc, dV_kc are complete: lanes = keys kc*128 + r
                        mbar_wait(acc_full, gacc & 1);
                        ++gacc;
                        tc_fence_after();
                        if (etid == 0) tma_store_wait_read<0>();
                        named_bar_sync(1, kFmhaBwdRowThreads);
                        stage_half(kColDK, 0.125f, sSt);
                        stage_half(kColDV, 1.0f, sSt + 16384);
                        fence_proxy_async_smem();
                        named_bar_sync(1, kFmhaBwdRowThreads);
                        if (etid == 0) {
                            tma_store_3d(&tmap_dqkv, sSt, d + h * 64, kc * 128, n);
                            tma_store_3d(&tmap_dqkv, sSt + 16384, 2 * d + h * 64, kc * 128, n);
                            tma_store_commit();
                        }
                    }
                }
            }
            // dQ tiles (the last acc_full covered every MMA of the unit)
            if (etid == 0) tma_store_wait_read<0>();
            named_bar_sync(1, kFmhaBwdRowThreads);
            for (int qt = 0; qt < QT; ++qt) stage_half(kColDQ + qt * 64, 0.125f, sSt + qt * 16384);
            tc_fence_before();
            fence_proxy_async_smem();
            named_bar_sync(1, kFmhaBwdRowThreads);
            if (etid == 0) {
                for (int qt = 0; qt < QT; ++qt) tma_store_3d(&tmap_dqkv, sSt + qt * 16384, h * 64, qt * 128, n);
                tma_store_commit();
            }
        }
        if (etid == 0) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

inline bool fmha_bwd_sm100_supported(int L) { return L >= 1 && L <= 256; }

inline int fmha_bwd_sm100(const void* qkv, const void* o, const void* d_o, const void* lse, void* dqkv, int N, int L,
                          int d, int heads, int causal, cudaStream_t stream) {
    const int Lp = (L + 15) / 16 * 16;
    const int KC = (Lp + 127) / 128;
    const int box_h = KC == 1 ? Lp : 128;  // rows of one K / V chunk (chunk 1 may run past L: zero-filled)
    CUtensorMap tq, tkv, tdo, tdq;
    {
        uint64_t dims[3] = {(uint64_t)3 * d, (uint64_t)L, (uint64_t)N};
        uint64_t str[2] = {(uint64_t)3 * d * 2, (uint64_t)L * 3 * d * 2};
        uint32_t box_q[3] = {64u, 128u, 1u};
        uint32_t box_kv[3] = {64u, (uint32_t)box_h, 1u};
        int rc = make_tmap_f16(&tq, qkv, 3, dims, str, box_q);
        if (rc) return rc;
        rc = make_tmap_f16(&tkv, qkv, 3, dims, str, box_kv);
        if (rc) return rc;
        rc = make_tmap_f16(&tdq, dqkv, 3, dims, str, box_q);
        if (rc) return rc;
    }
    {
        uint64_t dims[3] = {(uint64_t)d, (uint64_t)L, (uint64_t)N};
        uint64_t str[2] = {(uint64_t)d * 2, (uint64_t)L * d * 2};
        uint32_t box[3] = {64u, 128u, 1u};
        int rc = make_tmap_f16(&tdo, d_o, 3, dims, str, box);
        if (rc) return rc;
    }
    FmhaBwdParams p;
    p.L = L;
    p.Lp = Lp;
    p.heads = heads;
    p.d = d;
    p.QT = (L + 127) / 128;
    p.KC = KC;
    p.causal = causal;
    p.num_units = N * heads;
    p.box_h = box_h;
    p.scale_log2e = 0.125f * 1.4426950408889634f;
    p.lse = static_cast<const float*>(lse);
    p.o = static_cast<const __half*>(o);
    p.d_o = static_cast<const __half*>(d_o);
    if (p.QT != p.KC) return fail(MVLPT_ESHAPE, "fmha_bwd_sm100: internal: QT != KC for L=%d", L);
    p.off_do = (uint32_t)p.QT * 16384u;
    p.off_k = 2u * p.off_do;
    const uint32_t kv = (uint32_t)((KC * box_h * 128 + 1023) / 1024 * 1024);
    p.off_v = p.off_k + kv;
    p.off_p = p.off_v + kv;
    p.off_ds = p.off_p + 32768u;
    p.off_stage = p.off_ds + 32768u;
    p.off_row = p.off_stage + 32768u;
    p.off_bar = p.off_row + 2048u;
    const size_t smem = (size_t)p.off_bar + 128;
    if (smem > 227 * 1024) return fail(MVLPT_ESHAPE, "fmha_bwd_sm100: L=%d needs %zu bytes of shared memory", L, smem);
    static size_t attr = 0;
    if (smem > attr) {
        MVLPT_CUDA_OK(cudaFuncSetAttribute(fmha_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    const int grid = p.num_units < sm_count() ? p.num_units : sm_count();
    fmha_bwd_tc_kernel<<<grid, kFmhaBwdThreads, smem, stream>>>(tq, tkv, tdo, tdq, p);
    return launched("fmha_bwd_tc");
}

}  // namespace mvlpt
