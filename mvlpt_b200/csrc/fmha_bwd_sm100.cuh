// tcgen05 attention BACKWARD, head width 64, L <= 256 (autograd of F.multi_head_attention_forward as
// called at clip/model.py:181-183).  One work unit = one (sequence n, head h); the whole head lives on one SM.
//
// The score tile is computed TRANSPOSED (keys on the TMEM lanes, queries on the columns), in steps of
// (key chunk kc of 128 keys) x (query half-tile qh of 64 queries), double-buffered in TMEM:
//
//   MMA-1 : S^T  = K_kc . Q_qh^T   -> TMEM buffer b, columns [0,64)      dP^T = V_kc . dO_qh^T -> columns [64,128)
//   rows  : 8 warps = 2 per TMEM lane quarter (the pair splits the 64 query columns); thread = one key:
//             P^T = exp2(S^T*scale*log2e - lse[q]*log2e),   dS^T = P^T o (dP^T - D[q]),   D = rowsum(dO o O)
//           P^T and dS^T go back IN PLACE into TMEM as packed fp16 (tcgen05.st) — they are the A operands of
//             dV_kc += P^T  . dO_qh      -> TMEM [256,320)        dK_kc += dS^T . Q_qh  -> TMEM [320,384)
//           (tensor-memory A operand, no shared-memory traffic); dS^T also goes to shared memory (128-byte swizzled
//           rows = keys), from where a whole 128-query tile is consumed MN-major by
//             dQ_qt += dS . K_kc         -> TMEM [384+64 qt, ..)   once both of its halves are there.
//   While the row warps work on step j, the tensor pipe runs MMA-1 of step j+1 (other TMEM buffer) and the dV/dK/dQ
//   MMAs of step j-1; shared-memory dS is double-buffered per query tile.  No barrier is needed for either reuse: the
//   MMAs of one thread execute in issue order, and a later tcgen05.commit covers every earlier MMA.
//   after the last qh of a chunk: dK_kc * hd^-1/2, dV_kc -> fp16 -> staging -> TMA store; after the last chunk: dQ.
//
//   TMA   : Q, dO (128-row tiles) and K, V (128-row chunks) out of the packed [N, L, 3d] / [N, L, d] tensors through
//           3-D maps (rows >= L zero-filled).  Every tile is a stack of 128-byte rows with the 128B swizzle, which
//           serves BOTH operand majors (K-major when the contraction runs over the head width, MN-major when it runs
//           over the rows), so no transposed copy of anything is ever made.  The loads of a unit form two groups
//           (tile/chunk 0 and 1) with their own full/free barriers: group 0 of the NEXT unit streams in while the last
//           steps of this unit still compute.
//
// 448 threads: TMA warp, MMA warp, 8 row warps, 4 epilogue warps.  TMEM: all 512 columns, hence one CTA per SM.
#pragma once
#include "fmha_sm100.cuh"

namespace mvlpt {

struct FmhaBwdParams {
    // seq_len < L: every unit packs L / seq_len short sequences under a block-diagonal mask (see FmhaFwd2Params)
    int L, Lp, heads, d, QT, KC, NQH, causal, num_units, seq_len;
    float scale_log2e;   // hd^-1/2 * log2(e)
    const float* lse;    // [N, heads, L]
    const __half* o;     // [N, L, d]
    const __half* d_o;   // [N, L, d]
    uint32_t off_do, off_k, off_v, off_ds, off_stage, off_row, off_bar;
};

#ifdef MVLPT_FMHA_DBG
// Debug build only (tools/gpu_fmha_trace.py): CTA 0 records (tag, globaltimer) pairs of its MMA thread (stream 0) and of
// the first row thread (stream 1).
__device__ unsigned long long g_fmha_dbg[2][2048];
__device__ int g_fmha_dbg_n[2];
#define FMHA_DBG(stream, tag)                                                              \
    do {                                                                                   \
        if (blockIdx.x == 0) {                                                             \
            const int _i = g_fmha_dbg_n[stream];                                           \
            if (_i < 1024) {                                                               \
                g_fmha_dbg[stream][2 * _i] = (unsigned long long)(tag);                    \
                g_fmha_dbg[stream][2 * _i + 1] = globaltimer_ns();                         \
                g_fmha_dbg_n[stream] = _i + 1;                                             \
            }                                                                              \
        }                                                                                  \
    } while (0)
#else
#define FMHA_DBG(stream, tag) \
    do {                      \
    } while (0)
#endif

constexpr int kFmhaBwdThreads = 448;
constexpr int kFmhaBwdRowThreads = 256;

__global__ void __launch_bounds__(kFmhaBwdThreads, 1)
fmha_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                   const __grid_constant__ CUtensorMap tmap_do, const __grid_constant__ CUtensorMap tmap_dqkv,
                   const FmhaBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_bwd[];
    uint8_t* smem = smem_bwd;
    uint8_t* sQ = smem;                 // QT x [128 x 64]
    uint8_t* sdO = smem + p.off_do;     // QT x [128 x 64]
    uint8_t* sK = smem + p.off_k;       // KC x [128 x 64]
    uint8_t* sV = smem + p.off_v;       // KC x [128 x 64]
    uint8_t* sdS = smem + p.off_ds;     // 2 buffers x 2 chunks (64 queries each) of [128 keys x 128 B]
    uint8_t* sSt = smem + p.off_stage;  // 2 x [128 x 64] staging for the TMA stores
    float* sD = reinterpret_cast<float*>(smem + p.off_row);  // [256] rowsum(dO o O)
    float* sL2 = sD + 256;                                   // [256] lse * log2(e)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
    uint64_t* ld_full = bars + 0;   // [2] load group g landed
    uint64_t* g_free = bars + 2;    // [2] every MMA reading load group g of this unit has completed
    uint64_t* st_full = bars + 4;   // [2] S^T / dP^T of a step are in TMEM buffer b
    uint64_t* p_ready = bars + 6;   // [2] P^T / dS^T of a step are written (TMEM + shared memory)
    uint64_t* acc_full = bars + 8;  // dK / dV of a key chunk (and every earlier MMA) complete
    uint64_t* acc_free = bars + 9;  // the epilogue warps have read the accumulators of a key chunk out of TMEM
    uint64_t* d_full = bars + 10;   // per-query statistics (D, lse) of a unit are in shared memory
    uint64_t* kv1_full = bars + 11; // {K1,V1} landed (load group 1 is requested in two parts: {Q1,dO1} is needed first)
    uint64_t* stats_free = bars + 12;  // the row warps have read the per-query statistics of a unit for the last time
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Lp = p.Lp, d = p.d, QT = p.QT, KC = p.KC, NQH = p.NQH;
    constexpr uint32_t kColDV = 256, kColDK = 320, kColDQ = 384;

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
            mbar_init(&st_full[g], 1);
            mbar_init(&p_ready[g], kFmhaBwdRowThreads);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_free, 128);
        mbar_init(d_full, 128);
        mbar_init(kv1_full, 1);
        mbar_init(stats_free, kFmhaBwdRowThreads);
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
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (lane == 0) {
            int it = 0;
            for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++it) {
                const int h = unit % p.heads, n = unit / p.heads;
                for (int g = 0; g < QT; ++g) {  // QT == KC: group g = {Q_g, dO_g, K_g, V_g}
                    if (it > 0) mbar_wait(&g_free[g], (uint32_t)(it - 1) & 1);
                    // group 1 arrives right at the unit boundary: {Q1,dO1} (first used by step 3) goes first and has
                    // its own barrier, {K1,V1} (first used by step NQH+1) follows
                    uint64_t* kv_bar = g ? kv1_full : &ld_full[0];
                    mbar_arrive_expect_tx(&ld_full[g], g ? 32768u : 65536u);
                    tma_load_3d(sQ + g * 16384, &tmap_q, &ld_full[g], h * 64, g * 128, n);
                    tma_load_3d(sdO + g * 16384, &tmap_do, &ld_full[g], h * 64, g * 128, n);
                    if (g) mbar_arrive_expect_tx(kv1_full, 32768u);
                    tma_load_3d(sK + g * 16384, &tmap_q, kv_bar, d + h * 64, g * 128, n);
                    tma_load_3d(sV + g * 16384, &tmap_q, kv_bar, 2 * d + h * 64, g * 128, n);
                }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        // The WHOLE warp runs the control flow (warp-uniform values stay in uniform registers, which is what the
        // tcgen05 instructions take); one elected lane issues the MMAs and commits.
        const bool leader = elect_one();
        const uint32_t idesc_acc = umma_idesc_f16(128, 64, 0, 1);  // A from TMEM (K-major), B MN-major
        const uint32_t idesc_dq = umma_idesc_f16(128, 64, 1, 1);   // A = dS MN-major, B = K MN-major
        // descriptors of the tile bases; an operand at byte offset `off` inside a tile is desc + (off >> 4)
        const uint64_t dK_k = umma_desc_k_sw128(smem_u32(sK)), dQ_k = umma_desc_k_sw128(smem_u32(sQ));
        const uint64_t dV_k = umma_desc_k_sw128(smem_u32(sV)), ddO_k = umma_desc_k_sw128(smem_u32(sdO));
        const uint64_t ddO_mn = umma_desc_mn_sw128(smem_u32(sdO), 1024), dQ_mn = umma_desc_mn_sw128(smem_u32(sQ), 1024);
        const uint64_t dK_mn = umma_desc_mn_sw128(smem_u32(sK), 1024), ddS_mn = umma_desc_mn_sw128(smem_u32(sdS), 16384);
        uint32_t gi = 0;   // steps issued so far by this CTA (TMEM buffer = gi & 1)
        uint32_t tci = 0;  // (kc, query tile) pairs so far (dS shared-memory buffer = tci & 1)
        uint32_t gph = 0;  // key chunks started so far (parity of acc_free)
        // MMA-1 of step (kc, qh) into TMEM buffer b
        auto issue_st = [&](int kc, int qh, uint32_t b) {
            const int nq = (Lp - qh * 64) < 64 ? (Lp - qh * 64) : 64;
            const uint32_t idesc = umma_idesc_f16(128, (uint32_t)nq, 0, 0);
            const uint64_t k_desc = dK_k + (uint32_t)(kc * (16384 >> 4)), q_desc = dQ_k + (uint32_t)(qh * (8192 >> 4));
            const uint64_t v_desc = dV_k + (uint32_t)(kc * (16384 >> 4)), do_desc = ddO_k + (uint32_t)(qh * (8192 >> 4));
            const uint32_t tb = tmem_base + b * 128;
            if (leader) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tb, k_desc + 2 * k, q_desc + 2 * k, idesc, k != 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tb + 64, v_desc + 2 * k, do_desc + 2 * k, idesc, k != 0);
                umma_commit(&st_full[b]);
            }
            __syncwarp();
        };
        int it = 0;
        for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++it) {
            if (leader) FMHA_DBG(0, 1);
            mbar_wait(&ld_full[0], (uint32_t)it & 1);
            tc_fence_after();
            if (leader) FMHA_DBG(0, 2);
            issue_st(0, 0, gi & 1);
            bool g1_ready = false, kv1_ready = false;
            for (int kc = 0; kc < KC; ++kc) {
                const int nk = (Lp - kc * 128) < 128 ? (Lp - kc * 128) : 128;
                for (int qh = 0; qh < NQH; ++qh, ++gi) {
                    const uint32_t b = gi & 1;
                    const int nq = (Lp - qh * 64) < 64 ? (Lp - qh * 64) : 64;
                    // MMA-1 of the next step first: it overlaps the row warps' work on this one
                    {
                        const int nqh = (qh + 1 == NQH) ? 0 : qh + 1;
                        const int nkc = (qh + 1 == NQH) ? kc + 1 : kc;
                        if (nkc < KC) {
                            if (!g1_ready && nqh >= 2) {  // first touch of {Q1,dO1}
                                mbar_wait(&ld_full[1], (uint32_t)it & 1);
                                tc_fence_after();
                                g1_ready = true;
                            }
                            if (!kv1_ready && nkc > 0) {  // first touch of {K1,V1}
                                mbar_wait(kv1_full, (uint32_t)it & 1);
                                tc_fence_after();
                                kv1_ready = true;
                            }
                            issue_st(nkc, nqh, b ^ 1);
                        }
                    }
                    if (leader) FMHA_DBG(0, 3);
                    mbar_wait(&p_ready[b], (gi >> 1) & 1);
                    if (qh == 0) {
                        // the first dV / dK MMA of a chunk overwrites what the epilogue warps read (dQ likewise, later)
                        if (gph > 0) mbar_wait(acc_free, (gph - 1) & 1);
                        ++gph;
                    }
                    tc_fence_after();
                    if (leader) FMHA_DBG(0, 4);
                    // dV_kc (+)= P^T . dO_qh ; dK_kc (+)= dS^T . Q_qh      contraction over the nq queries
                    const uint32_t tb = tmem_base + b * 128;
                    const uint64_t b_do = ddO_mn + (uint32_t)(qh * (8192 >> 4)), b_q = dQ_mn + (uint32_t)(qh * (8192 >> 4));
                    const int nkk = nq >> 4;
                    if (leader) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            if (kk < nkk) {
                                const uint32_t pc = (uint32_t)(kk * 8 + (kk >= 2 ? 16 : 0));  // see the row warps' layout
                                umma_f16_ts(tmem_base + kColDV, tb + pc, b_do + kk * (2048 >> 4), idesc_acc, (qh | kk) != 0);
                                umma_f16_ts(tmem_base + kColDK, tb + 64 + pc, b_q + kk * (2048 >> 4), idesc_acc, (qh | kk) != 0);
                            }
                        }
                    }
                    if ((qh & 1) || qh == NQH - 1) {
                        // both halves of query tile qt are in shared memory: dQ_qt (+)= dS . K_kc over the nk keys
                        const int qt = qh >> 1;
                        const uint64_t a_ds = ddS_mn + (uint32_t)((tci & 1) * (32768 >> 4));
                        const uint64_t b_k = dK_mn + (uint32_t)(kc * (16384 >> 4));
                        const uint32_t dq = tmem_base + kColDQ + qt * 64;
                        const int nkk2 = nk >> 4;
                        if (leader) {
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk)
                                if (kk < nkk2)
                                    umma_f16_ss(dq, a_ds + kk * (2048 >> 4), b_k + kk * (2048 >> 4), idesc_dq, (kc | kk) != 0);
                        }
                        ++tci;
                    }
                    if (leader) {
                        if (qh == NQH - 1) umma_commit(acc_full);
                        // load group 0 (tile / chunk 0) is last read by the steps of query tile 0 in the last key chunk
                        if (kc == KC - 1 && qh == (NQH > 1 ? 1 : 0)) umma_commit(&g_free[0]);
                        if (kc == KC - 1 && qh == NQH - 1 && QT > 1) umma_commit(&g_free[1]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp < 10) {
        // ============================== row warps: softmax backward ==============================
        const int quarter = warp & 3;        // TMEM lane quarter this warp may access
        const int group = (warp - 2) >> 2;   // the two warps of a quarter split the query columns
        const int r = quarter * 32 + lane;   // key inside the chunk
        const uint32_t t_lane = tmem_base + (uint32_t(quarter * 32) << 16);
        const int sw = r & 7;
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        const float sl2 = p.scale_log2e;
        uint32_t gi = 0, tci = 0;
        int it = 0;
        for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++it) {
            if (threadIdx.x == 64) FMHA_DBG(1, 10);
            mbar_wait(d_full, (uint32_t)it & 1);  // D and lse of this unit are in shared memory
            if (threadIdx.x == 64) FMHA_DBG(1, 12);
            for (int kc = 0; kc < KC; ++kc) {
                const int nk = (Lp - kc * 128) < 128 ? (Lp - kc * 128) : 128;
                const bool live = quarter * 32 < nk;  // warp-uniform: some of this warp's keys exist
                const int key = kc * 128 + r;
                const bool key_ok = key < p.L;
                const int warp_key_hi = kc * 128 + quarter * 32 + 31;
                // queries that see this key: [q_first, q_end) — its own sequence, from itself on when causal
                const int key_seq0 = (key / p.seq_len) * p.seq_len;
                const int q_first = p.causal ? key : key_seq0;
                const int q_end = !key_ok ? q_first : (key_seq0 + p.seq_len < p.L ? key_seq0 + p.seq_len : p.L);
                const bool unpacked = p.seq_len == p.L;
                for (int qh = 0; qh < NQH; ++qh, ++gi) {
                    const uint32_t b = gi & 1;
                    const int nq = (Lp - qh * 64) < 64 ? (Lp - qh * 64) : 64;
                    // query columns [0, 32) -> group 0, [32, 64) -> group 1
                    const int cb = group * 32, ce = nq < cb + 32 ? nq : cb + 32;
                    if (threadIdx.x == 64) FMHA_DBG(1, 13);
                    mbar_wait(&st_full[b], (gi >> 1) & 1);
                    tc_fence_after();
                    if (threadIdx.x == 64) FMHA_DBG(1, 14);
                    if (live) {
                        const uint32_t tS = t_lane + b * 128, tDP = tS + 64;
                        uint8_t* ds_row = sdS + (tci & 1) * 32768u + (uint32_t)(qh & 1) * 16384u + row_off;
                        for (int c0 = cb; c0 < ce; c0 += 16) {
                            uint32_t s[16], dp[16];
                            tmem_ld_32x32b_x16(tS + c0, s);
                            tmem_ld_32x32b_x16(tDP + c0, dp);
                            const int q0 = qh * 64 + c0;
                            float l2[16], Dq[16];
#pragma unroll
                            for (int v = 0; v < 4; ++v) {
                                const float4 a = *reinterpret_cast<const float4*>(sL2 + q0 + 4 * v);
                                const float4 c = *reinterpret_cast<const float4*>(sD + q0 + 4 * v);
                                l2[4 * v] = a.x, l2[4 * v + 1] = a.y, l2[4 * v + 2] = a.z, l2[4 * v + 3] = a.w;
                                Dq[4 * v] = c.x, Dq[4 * v + 1] = c.y, Dq[4 * v + 2] = c.z, Dq[4 * v + 3] = c.w;
                            }
                            tmem_ld_wait();
                            float pv[16], dsv[16];
                            // warp-uniform: every (key of this warp, query of this group) pair is inside the mask,
                            // except for keys >= L, which are zeroed per thread afterwards
                            if (unpacked && q0 + 16 <= p.L && (!p.causal || warp_key_hi <= q0)) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    pv[j] = ex2_approx(fmaf(__uint_as_float(s[j]), sl2, -l2[j]));
                                    dsv[j] = pv[j] * (__uint_as_float(dp[j]) - Dq[j]);
                                }
                                if (!key_ok) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) pv[j] = 0.f, dsv[j] = 0.f;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    const bool ok = q0 + j >= q_first && q0 + j < q_end;
                                    pv[j] = ok ? ex2_approx(fmaf(__uint_as_float(s[j]), sl2, -l2[j])) : 0.f;
                                    dsv[j] = ok ? pv[j] * (__uint_as_float(dp[j]) - Dq[j]) : 0.f;
                                }
                            }
                            uint32_t pp[8], dd[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                pp[j] = pack_half2(pv[2 * j], pv[2 * j + 1]);
                                dd[j] = pack_half2(dsv[2 * j], dsv[2 * j + 1]);
                            }
                            // fp16 columns of queries [c0, c0+16): 8 TMEM columns inside this warp's own part of the tile
                            const uint32_t pc = (uint32_t)(c0 >> 1) + (c0 >= 32 ? 16u : 0u);
                            tmem_st_32x32b_x8(tS + pc, pp);
                            tmem_st_32x32b_x8(tDP + pc, dd);
                            const int u = c0 >> 3;
                            *reinterpret_cast<uint4*>(ds_row + ((u ^ sw) << 4)) = make_uint4(dd[0], dd[1], dd[2], dd[3]);
                            *reinterpret_cast<uint4*>(ds_row + (((u + 1) ^ sw) << 4)) = make_uint4(dd[4], dd[5], dd[6], dd[7]);
                        }
                        tmem_st_wait();
                    }
                    fence_proxy_async_smem();
                    tc_fence_before();
                    mbar_arrive(&p_ready[b]);
                    if (threadIdx.x == 64) FMHA_DBG(1, 15);
                    if ((qh & 1) || qh == NQH - 1) ++tci;
                }
            }
            mbar_arrive(stats_free);  // sD / sL2 of this unit may be overwritten
        }
    } else {
        // ============================== epilogue warps ==============================
        // Per unit: D = rowsum(dO o O) and lse*log2e of the NEXT unit (global loads, latency hidden behind this unit's
        // steps, published once this unit's steps are over); dK / dV after each key chunk and dQ after the last one:
        // TMEM -> fp16 -> swizzled staging tile -> TMA store.  The MMA warp waits for `acc_free` before it overwrites
        // an accumulator these warps read.
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;  // row of an output tile == TMEM lane
        const uint32_t t_lane = tmem_base + (uint32_t(quarter * 32) << 16);
        const bool leader = threadIdx.x == 320;
        const int sw = r & 7;
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        uint32_t gacc = 0;

        // D and lse*log2e of row r of query tile qt of a unit (0 for rows >= L)
        auto row_stats = [&](int unit, int qt, float& Dv, float& l2) {
            const int h = unit % p.heads, n = unit / p.heads;
            const int q = qt * 128 + r;
            Dv = 0.f, l2 = 0.f;
            if (q < p.L) {
                const uint4* po = reinterpret_cast<const uint4*>(p.o + ((size_t)n * p.L + q) * d + h * 64);
                const uint4* pd = reinterpret_cast<const uint4*>(p.d_o + ((size_t)n * p.L + q) * d + h * 64);
                float acc = 0.f;
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
                Dv = acc;
                const int per = p.L / p.seq_len, sq = q / p.seq_len;  // sequences per unit, this row's sequence
                l2 = p.lse[((size_t)(n * per + sq) * p.heads + h) * p.seq_len + (q - sq * p.seq_len)] * 1.4426950408889634f;
            }
        };
        auto publish_stats = [&](const float (&Dv)[2], const float (&l2)[2]) {
            for (int qt = 0; qt < QT; ++qt) {
                sD[qt * 128 + r] = Dv[qt];
                sL2[qt * 128 + r] = l2[qt];
            }
            mbar_arrive(d_full);
        };
        // 64 fp32 TMEM columns of row r, scaled -> 32 packed fp16 pairs
        auto load_row = [&](uint32_t col, float mul, uint32_t (&pk)[32]) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {  // 32 columns at a time keeps the register peak low
                uint32_t a[32];
                tmem_ld_32x32(t_lane + col + 32 * hh, a);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    pk[16 * hh + j] = pack_half2(__uint_as_float(a[2 * j]) * mul, __uint_as_float(a[2 * j + 1]) * mul);
            }
        };
        auto stage_row = [&](const uint32_t (&pk)[32], uint8_t* tile) {
            uint8_t* orow = tile + row_off;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                *reinterpret_cast<uint4*>(orow + ((u ^ sw) << 4)) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        };

        float Dn[2] = {0.f, 0.f}, Ln[2] = {0.f, 0.f};
        if ((int)blockIdx.x < p.num_units) {
            for (int qt = 0; qt < QT; ++qt) row_stats(blockIdx.x, qt, Dn[qt], Ln[qt]);
            publish_stats(Dn, Ln);
        }
        int it = 0;
        for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++it) {
            const int h = unit % p.heads, n = unit / p.heads;
            const int next = unit + gridDim.x;
            if (next < p.num_units)
                for (int qt = 0; qt < QT; ++qt) row_stats(next, qt, Dn[qt], Ln[qt]);
            for (int kc = 0; kc < KC; ++kc) {
                if (kc == KC - 1 && next < p.num_units) {
                    // the statistics of the NEXT unit go out as soon as the row warps are through with this unit's —
                    // before the accumulators of the last chunk are drained, so the next unit's first step does not wait
                    // for this epilogue (in-kernel trace: 2-3 us per unit of 14 before)
                    mbar_wait(stats_free, (uint32_t)it & 1);
                    publish_stats(Dn, Ln);
                }
                mbar_wait(acc_full, gacc & 1);  // dK_kc, dV_kc (and, for the last chunk, dQ) complete: lanes = rows
                ++gacc;
                tc_fence_after();
                if (leader) FMHA_DBG(1, 16);
                uint32_t pk[32];
                if (leader) tma_store_wait_read<0>();  // the previous stores have left the staging tiles
                named_bar_sync(2, 128);
                load_row(kColDK, 0.125f, pk);
                stage_row(pk, sSt);
                load_row(kColDV, 1.0f, pk);
                stage_row(pk, sSt + 16384);
                if (kc < KC - 1) {
                    tc_fence_before();
                    mbar_arrive(acc_free);
                    fence_proxy_async_smem();
                    named_bar_sync(2, 128);
                    if (leader) {
                        tma_store_3d(&tmap_dqkv, sSt, d + h * 64, kc * 128, n);
                        tma_store_3d(&tmap_dqkv, sSt + 16384, 2 * d + h * 64, kc * 128, n);
                        tma_store_commit();
                        FMHA_DBG(1, 17);
                    }
                } else {
                    // last chunk: the steps of this unit are over -> dQ is complete and the row warps are done with
                    // the per-query statistics
                    uint32_t pq[32];
                    load_row(kColDQ, 0.125f, pk);
                    if (QT > 1) load_row(kColDQ + 64, 0.125f, pq);
                    tc_fence_before();
                    mbar_arrive(acc_free);
                    fence_proxy_async_smem();
                    named_bar_sync(2, 128);
                    if (leader) {
                        tma_store_3d(&tmap_dqkv, sSt, d + h * 64, kc * 128, n);
                        tma_store_3d(&tmap_dqkv, sSt + 16384, 2 * d + h * 64, kc * 128, n);
                        tma_store_commit();
                        tma_store_wait_read<0>();
                    }
                    named_bar_sync(2, 128);
                    stage_row(pk, sSt);
                    if (QT > 1) stage_row(pq, sSt + 16384);
                    fence_proxy_async_smem();
                    named_bar_sync(2, 128);
                    if (leader) {
                        for (int qt = 0; qt < QT; ++qt) tma_store_3d(&tmap_dqkv, sSt + qt * 16384, h * 64, qt * 128, n);
                        tma_store_commit();
                        FMHA_DBG(1, 18);
                    }
                }
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

inline bool fmha_bwd_sm100_supported(int L) { return L >= 1 && L <= 256; }

inline int fmha_bwd_sm100(const void* qkv, const void* o, const void* d_o, const void* lse, void* dqkv, int N, int L,
                          int d, int heads, int causal, cudaStream_t stream, int seq_len = 0) {
    // short sequences: G = 128 / L of them per unit (same rule as fmha_fwd_sm100: both passes see the same lse layout)
    const char* pack_env = getenv("MVLPT_FMHA_PACK");  // read per call: tests switch it inside one process
    const bool pack = !(pack_env && pack_env[0] == '0');
    if (seq_len == 0 && pack && L <= 64 && N >= 2) {
        const int G = 128 / L, groups = N / G, rem = N - groups * G;
        if (groups > 0) {
            const int rc = fmha_bwd_sm100(qkv, o, d_o, lse, dqkv, groups, G * L, d, heads, causal, stream, L);
            if (rc) return rc;
        }
        if (rem > 0) {
            const size_t row0 = (size_t)groups * G * L;
            return fmha_bwd_sm100(static_cast<const __half*>(qkv) + row0 * 3 * d, static_cast<const __half*>(o) + row0 * d,
                                  static_cast<const __half*>(d_o) + row0 * d,
                                  static_cast<const float*>(lse) + (size_t)groups * G * heads * L,
                                  static_cast<__half*>(dqkv) + row0 * 3 * d, 1, rem * L, d, heads, causal, stream, L);
        }
        return MVLPT_OK;
    }
    const int Lp = (L + 15) / 16 * 16;
    CUtensorMap tq, tdo, tdq;
    {
        // one map serves Q, K and V (different column offsets): boxes of 128 rows, rows >= L zero-filled
        uint64_t dims[3] = {(uint64_t)3 * d, (uint64_t)L, (uint64_t)N};
        uint64_t str[2] = {(uint64_t)3 * d * 2, (uint64_t)L * 3 * d * 2};
        uint32_t box_q[3] = {64u, 128u, 1u};
        int rc = make_tmap_f16(&tq, qkv, 3, dims, str, box_q);
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
    p.KC = (Lp + 127) / 128;
    p.NQH = (Lp + 63) / 64;
    p.causal = causal;
    p.seq_len = seq_len > 0 ? seq_len : L;
    p.num_units = N * heads;
    p.scale_log2e = 0.125f * 1.4426950408889634f;
    p.lse = static_cast<const float*>(lse);
    p.o = static_cast<const __half*>(o);
    p.d_o = static_cast<const __half*>(d_o);
    if (p.QT != p.KC) return fail(MVLPT_ESHAPE, "fmha_bwd_sm100: internal: QT != KC for L=%d", L);
    p.off_do = (uint32_t)p.QT * 16384u;
    p.off_k = 2u * p.off_do;
    p.off_v = p.off_k + (uint32_t)p.KC * 16384u;
    p.off_ds = p.off_v + (uint32_t)p.KC * 16384u;
    p.off_stage = p.off_ds + 65536u;
    p.off_row = p.off_stage + 32768u;
    p.off_bar = p.off_row + 2048u;
    const size_t smem = (size_t)p.off_bar + 128;
    if (smem > 227 * 1024) return fail(MVLPT_ESHAPE, "fmha_bwd_sm100: L=%d needs %zu bytes of shared memory", L, smem);
    static DynSmemCache attr;
    if (int rc = ensure_dyn_smem(fmha_bwd_tc_kernel, smem, attr)) return rc;
    const int grid = p.num_units < sm_count() ? p.num_units : sm_count();
    MVLPT_CUDA_OK(launch_pdl(fmha_bwd_tc_kernel, dim3(grid), dim3(kFmhaBwdThreads), smem, stream, 1, tq, tq, tdo, tdq, p));
    return launched("fmha_bwd_tc");
}

}  // namespace mvlpt
