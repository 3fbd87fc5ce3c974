// C-ABI entry for the tcgen05 GEMM (see include/mvlpt_sm100.h: mvlpt_gemm).
#include <stdlib.h>

#include "common.cuh"
#include "gemm_sm100.cuh"

using namespace mvlpt;

// Operand-ring depth / output-ring slots.  Epilogues with a TMA-loaded input or a second output want a deep slab
// ring; plain epilogues want the deepest operand ring.  MVLPT_GEMM_STAGES overrides the depth (tuning only).
template <int BN>
static void pick_pipeline(GemmEpilogue& ep) {
    using Cfg = GemmCfg<BN>;
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("MVLPT_GEMM_STAGES");
        forced = e ? atoi(e) : 0;
    }
    int stages = (ep.has_in || ep.has_aux_out) ? (BN == 256 ? 3 : 4) : (BN == 256 ? 4 : 5);
    if (forced >= 2 && forced <= kGemmMaxStages && Cfg::slabs_for(forced) >= 2) stages = forced;
    int slabs = Cfg::slabs_for(stages);
    const int per = ep.has_aux_out ? 2 : 1;
    int ring = slabs / per;
    if (ring > kGemmMaxRing) ring = kGemmMaxRing;
    ep.stages = stages;
    ep.ring = ring;
}

template <int BN, bool F32>
static int launch_gemm(const mvlpt_gemm_desc* d, const void* A, const void* W, const void* in, void* aux_out, void* out,
                       GemmEpilogue ep, cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    pick_pipeline<BN>(ep);
    if (ep.ring < 2) return fail(MVLPT_ESHAPE, "mvlpt_gemm: no room for the output ring");
    const int smem_bytes = Cfg::smem_bytes(ep.stages, ep.ring * (ep.has_aux_out ? 2 : 1));
    CUtensorMap ta, tw, to, tx, ti;
    {
        uint64_t dims[2] = {(uint64_t)d->K, (uint64_t)d->M};
        uint64_t str[1] = {(uint64_t)d->lda * 2};
        uint32_t box[2] = {(uint32_t)kGemmBK, (uint32_t)kGemmBM};
        int rc = make_tmap_f16(&ta, A, 2, dims, str, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)d->K, (uint64_t)d->N};
        uint64_t str[1] = {(uint64_t)d->ldw * 2};
        uint32_t box[2] = {(uint32_t)kGemmBK, (uint32_t)BN};
        int rc = make_tmap_f16(&tw, W, 2, dims, str, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)d->N, (uint64_t)d->M};
        uint64_t str[1] = {(uint64_t)d->ld_out * (F32 ? 4 : 2)};
        uint32_t box[2] = {F32 ? 32u : 64u, (uint32_t)kGemmBM};
        int rc = make_tmap(&to, out, F32 ? 1 : 0, 2, dims, str, box);
        if (rc) return rc;
    }
    tx = to;
    ti = to;
    if (in) {
        uint64_t dims[2] = {(uint64_t)d->N, (uint64_t)d->M};
        uint64_t str[1] = {F32 ? (uint64_t)d->ld_out * 4 : (uint64_t)d->ld_aux * 2};
        uint32_t box[2] = {F32 ? 32u : 64u, (uint32_t)kGemmBM};
        int rc = make_tmap(&ti, in, F32 ? 1 : 0, 2, dims, str, box);
        if (rc) return rc;
    }
    if (aux_out) {
        uint64_t dims[2] = {(uint64_t)d->N, (uint64_t)d->M};
        uint64_t str[1] = {(uint64_t)d->ld_aux * 2};
        uint32_t box[2] = {64u, (uint32_t)kGemmBM};
        int rc = make_tmap_f16(&tx, aux_out, 2, dims, str, box);
        if (rc) return rc;
    }
    static DynSmemCache attr;
    if (int rc = ensure_dyn_smem(gemm_f16_tn_kernel<BN, F32>, (size_t)smem_bytes, attr)) return rc;
    const int tiles = cdiv(d->M, kGemmBM) * cdiv(d->N, BN);
    const int grid = tiles < sm_count() ? tiles : sm_count();
    MVLPT_CUDA_OK(launch_pdl(gemm_f16_tn_kernel<BN, F32>, dim3(grid), dim3(kGemmThreads), smem_bytes, stream, 1, ta, tw, to, tx, ti,
                             d->M, d->N, d->K, ep));
    return launched("gemm_f16_tn");
}

// CTA-pair kernel (256 x 256 tile per cluster of two CTAs): N a multiple of 256, at least one full 256-row tile.
static bool use_2sm(const mvlpt_gemm_desc* d) {
    static int off = -1;
    if (off < 0) off = getenv("MVLPT_GEMM_1SM") ? 1 : 0;
    return !off && d->N >= 256 && (d->N % 256) == 0 && d->M >= 256;
}

template <bool F32>
static int launch_gemm_2sm(const mvlpt_gemm_desc* d, const void* A, const void* W, const void* in, void* aux_out, void* out,
                           GemmEpilogue ep, cudaStream_t stream) {
    constexpr int kStage = 32768;
    // operand-ring depth / output-ring slots out of the 227 KB: deep operand ring for plain epilogues, deep slab ring
    // when the epilogue has a TMA-loaded input or a second output
    const int per = ep.has_aux_out ? 2 : 1;
    // measured (tools/gpu_bench_kernels.py, MVLPT_GEMM2_STAGES sweep): the fp32 residual epilogue is best with 5 operand
    // stages + 4 slabs, the QuickGELU ones with 4 + 6, plain ones with 6 + 2
    int stages = (ep.has_in || ep.has_aux_out) ? ((F32 && !ep.has_aux_out) ? 5 : 4) : 6;
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("MVLPT_GEMM2_STAGES");
        forced = e ? atoi(e) : 0;
    }
    if (forced >= 2 && forced <= kGemmMaxStages2) stages = forced;
    int slabs = (kGemmSmemBudget + 1024 - stages * kStage) / kGemmSlab;
    int ring = slabs / per;
    if (ring > kGemmMaxRing) ring = kGemmMaxRing;
    if (ring < 2) return fail(MVLPT_ESHAPE, "mvlpt_gemm: no room for the output ring");
    ep.stages = stages;
    ep.ring = ring;
    const int smem_bytes = stages * kStage + ring * per * kGemmSlab + 256;
    CUtensorMap ta, tw, to, tx, ti;
    {
        uint64_t dims[2] = {(uint64_t)d->K, (uint64_t)d->M};
        uint64_t str[1] = {(uint64_t)d->lda * 2};
        uint32_t box[2] = {(uint32_t)kGemmBK, (uint32_t)kGemmBM};
        int rc = make_tmap_f16(&ta, A, 2, dims, str, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)d->K, (uint64_t)d->N};
        uint64_t str[1] = {(uint64_t)d->ldw * 2};
        uint32_t box[2] = {(uint32_t)kGemmBK, 128u};  // each CTA of the pair stages half of the 256 W rows
        int rc = make_tmap_f16(&tw, W, 2, dims, str, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)d->N, (uint64_t)d->M};
        uint64_t str[1] = {(uint64_t)d->ld_out * (F32 ? 4 : 2)};
        uint32_t box[2] = {F32 ? 32u : 64u, (uint32_t)kGemmBM};
        int rc = make_tmap(&to, out, F32 ? 1 : 0, 2, dims, str, box);
        if (rc) return rc;
    }
    tx = to;
    ti = to;
    if (in) {
        uint64_t dims[2] = {(uint64_t)d->N, (uint64_t)d->M};
        uint64_t str[1] = {F32 ? (uint64_t)d->ld_out * 4 : (uint64_t)d->ld_aux * 2};
        uint32_t box[2] = {F32 ? 32u : 64u, (uint32_t)kGemmBM};
        int rc = make_tmap(&ti, in, F32 ? 1 : 0, 2, dims, str, box);
        if (rc) return rc;
    }
    if (aux_out) {
        uint64_t dims[2] = {(uint64_t)d->N, (uint64_t)d->M};
        uint64_t str[1] = {(uint64_t)d->ld_aux * 2};
        uint32_t box[2] = {64u, (uint32_t)kGemmBM};
        int rc = make_tmap_f16(&tx, aux_out, 2, dims, str, box);
        if (rc) return rc;
    }
    static DynSmemCache attr;
    if (int rc = ensure_dyn_smem(gemm_f16_tn_2sm_kernel<F32>, (size_t)smem_bytes, attr)) return rc;
    const int tiles = cdiv(d->M, 2 * kGemmBM) * cdiv(d->N, 256);
    const int pairs = tiles < sm_count() / 2 ? tiles : sm_count() / 2;
    MVLPT_CUDA_OK(launch_pdl(gemm_f16_tn_2sm_kernel<F32>, dim3(2 * pairs), dim3(kGemm2Threads), smem_bytes, stream, 2, ta, tw, to,
                             tx, ti, d->M, d->N, d->K, ep));
    return launched("gemm_f16_tn_2sm");
}

static int gemm_impl(const mvlpt_gemm_desc* d, const void* A, const void* W, const void* bias, const void* aux_in,
                     void* aux_out, const void* resid, void* out, const mvlpt_ln_carry* ln, mvlpt_stream_t stream);

extern "C" int mvlpt_gemm(const mvlpt_gemm_desc* d, const void* A, const void* W, const void* bias,
                          const void* aux_in, void* aux_out, const void* resid, void* out, mvlpt_stream_t stream) {
    return gemm_impl(d, A, W, bias, aux_in, aux_out, resid, out, nullptr, stream);
}

extern "C" int mvlpt_gemm_ln_supported(int M, int width) {
    static int off = -1;
    if (off < 0) off = (getenv("MVLPT_GEMM_1SM") || getenv("MVLPT_NO_FUSED_LN")) ? 1 : 0;
    return !off && M >= 256 && width >= 256 && (width % 256) == 0 && width <= 1024;
}

extern "C" int mvlpt_gemm_ln(const mvlpt_gemm_desc* d, const void* A, const void* W, const void* bias,
                             const void* aux_in, void* aux_out, const void* resid, void* out, const mvlpt_ln_carry* ln,
                             mvlpt_stream_t stream) {
    if (!d || !ln) return fail(MVLPT_EINVAL, "mvlpt_gemm_ln: null argument");
    const bool prod = ln->rec_out != nullptr, cons = ln->rec != nullptr;
    if (prod == cons) return fail(MVLPT_EINVAL, "mvlpt_gemm_ln: exactly one of the producer (rec_out) / consumer (rec) sides");
    if (!mvlpt_gemm_ln_supported(d->M, ln->width) || (d->N % 256))
        return fail(MVLPT_ESHAPE, "mvlpt_gemm_ln: needs M >= 256, N %% 256 == 0 and a row width that is a multiple of 256 up "
                                  "to 1024 (got M=%d N=%d width=%d)", d->M, d->N, ln->width);
    if (prod) {
        if (!d->out_f32 || d->act != ACT_NONE || d->N != ln->width || d->alpha != 1.f)
            return fail(MVLPT_ESHAPE, "mvlpt_gemm_ln producer: fp32 output of the row width, no activation, alpha 1");
        if (!ln->gamma || !ln->xt) return fail(MVLPT_EINVAL, "mvlpt_gemm_ln producer: gamma and xt are required");
        if ((reinterpret_cast<uintptr_t>(ln->gamma) | reinterpret_cast<uintptr_t>(ln->xt) |
             reinterpret_cast<uintptr_t>(ln->rec_out) | reinterpret_cast<uintptr_t>(ln->rec_in)) & 15)
            return fail(MVLPT_EINVAL, "mvlpt_gemm_ln: gamma / xt / records must be 16-byte aligned");
    } else {
        if (d->out_f32 || d->K != ln->width || d->alpha != 1.f || !bias)
            return fail(MVLPT_ESHAPE, "mvlpt_gemm_ln consumer: fp16 output, K == row width, alpha 1, bias = bp");
        if (!ln->sg) return fail(MVLPT_EINVAL, "mvlpt_gemm_ln consumer: sg is required");
        if ((reinterpret_cast<uintptr_t>(ln->sg) | reinterpret_cast<uintptr_t>(ln->rec)) & 15)
            return fail(MVLPT_EINVAL, "mvlpt_gemm_ln: sg / records must be 16-byte aligned");
    }
    return gemm_impl(d, A, W, bias, aux_in, aux_out, resid, out, ln, stream);
}

static int gemm_impl(const mvlpt_gemm_desc* d, const void* A, const void* W, const void* bias, const void* aux_in,
                     void* aux_out, const void* resid, void* out, const mvlpt_ln_carry* ln, mvlpt_stream_t stream) {
    if (!d || !A || !W || !out) return fail(MVLPT_EINVAL, "mvlpt_gemm: null argument");
    if (d->M <= 0 || d->N <= 0 || d->K <= 0) return fail(MVLPT_EINVAL, "mvlpt_gemm: M,N,K must be positive");
    if (d->lda < d->K || d->ldw < d->K || d->ld_out < d->N)
        return fail(MVLPT_EINVAL, "mvlpt_gemm: leading dimension smaller than extent");
    if ((d->lda % 8) || (d->ldw % 8)) return fail(MVLPT_ESHAPE, "mvlpt_gemm: lda/ldw must be multiples of 8");
    if (d->out_f32 ? (d->ld_out % 4) : (d->ld_out % 8))
        return fail(MVLPT_ESHAPE, "mvlpt_gemm: ld_out must be a multiple of %d", d->out_f32 ? 4 : 8);
    if ((reinterpret_cast<uintptr_t>(out) & 15) || (resid && (reinterpret_cast<uintptr_t>(resid) & 15)) ||
        (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) || (aux_in && (reinterpret_cast<uintptr_t>(aux_in) & 15)) ||
        (aux_out && (reinterpret_cast<uintptr_t>(aux_out) & 15)))
        return fail(MVLPT_EINVAL, "mvlpt_gemm: out/resid/bias/aux must be 16-byte aligned");
    if (d->act < 0 || d->act > 2) return fail(MVLPT_EINVAL, "mvlpt_gemm: unknown act %d", d->act);
    if (d->act != ACT_NONE && d->out_f32)
        return fail(MVLPT_ESHAPE, "mvlpt_gemm: the QuickGELU epilogues produce fp16 (clear out_f32)");
    if (d->act == ACT_MUL_DQUICKGELU && !aux_in) return fail(MVLPT_EINVAL, "mvlpt_gemm: act 2 needs aux_in");
    if (aux_out && d->act != ACT_QUICKGELU) return fail(MVLPT_EINVAL, "mvlpt_gemm: aux_out is the act-1 pre-activation");
    if ((aux_in || aux_out) && (d->ld_aux < d->N || (d->ld_aux % 8)))
        return fail(MVLPT_ESHAPE, "mvlpt_gemm: ld_aux must be >= N and a multiple of 8");
    if (resid && !d->out_f32) return fail(MVLPT_ESHAPE, "mvlpt_gemm: the residual stream is fp32; set out_f32");
    int rc = require_sm100();
    if (rc) return rc;

    GemmEpilogue ep;
    ep.bias = static_cast<const __half*>(bias);
    const void* in = d->out_f32 ? resid : (d->act == ACT_MUL_DQUICKGELU ? aux_in : nullptr);
    ep.has_in = in != nullptr;
    ep.has_aux_out = aux_out != nullptr;
    ep.act = d->act;
    ep.alpha = d->alpha;
    ep.stages = ep.ring = 0;
    ep.lnp_rec_in = ep.lnp_gamma = ep.lnc_rec = nullptr;
    ep.lnc_sg = nullptr;
    ep.lnp_rec_out = nullptr;
    ep.lnp_xt = nullptr;
    ep.ln_parts = 0;
    ep.ln_inv_d = ep.ln_eps = 0.f;
    if (ln) {
        ep.lnp_rec_in = static_cast<const float*>(ln->rec_in);
        ep.lnp_rec_out = static_cast<float*>(ln->rec_out);
        ep.lnp_gamma = static_cast<const float*>(ln->gamma);
        ep.lnp_xt = static_cast<__half*>(ln->xt);
        ep.lnc_rec = static_cast<const float*>(ln->rec);
        ep.lnc_sg = static_cast<const __half*>(ln->sg);
        ep.ln_parts = ln->width / 128;
        ep.ln_inv_d = 1.f / (float)ln->width;
        ep.ln_eps = ln->eps;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (use_2sm(d))
        return d->out_f32 ? launch_gemm_2sm<true>(d, A, W, in, aux_out, out, ep, s)
                          : launch_gemm_2sm<false>(d, A, W, in, aux_out, out, ep, s);
    if (d->N <= 128)
        return d->out_f32 ? launch_gemm<128, true>(d, A, W, in, aux_out, out, ep, s)
                          : launch_gemm<128, false>(d, A, W, in, aux_out, out, ep, s);
    return d->out_f32 ? launch_gemm<256, true>(d, A, W, in, aux_out, out, ep, s)
                      : launch_gemm<256, false>(d, A, W, in, aux_out, out, ep, s);
}
