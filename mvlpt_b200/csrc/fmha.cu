// Multi-head attention core, forward and backward, head width 64, whole key range resident in shared memory
// (L <= 288: the reference's sequences are 50..257 image tokens + prompts, <= 77 text tokens).
//
// Replaces the inside of nn.MultiheadAttention as called at clip/model.py:181-183
// (F.multi_head_attention_forward: scale q by hd^-1/2, QK^T, additive causal mask for the text tower
// clip/model.py:324-330, softmax over keys, PV) and its autograd backward.
// qkv is the packed in_proj output [N*L, 3d] (Q | K | V, head j = columns [64j, 64j+64) of each part).
//
// Kernels: csrc/fmha_sm100.cuh (tcgen05 forward, L <= 272), csrc/fmha_bwd_sm100.cuh (tcgen05 backward, L <= 256) and
// csrc/fmha_long.cuh (streaming kernels for everything longer: ViT-L/14@336px, and the ViT-L/14 + prompts backward).

#include "common.cuh"
#include "fmha_sm100.cuh"
#include "fmha_bwd_sm100.cuh"
#include "fmha_long.cuh"

using namespace mvlpt;

namespace {
constexpr int HD = 64;
}  // namespace

extern "C" int mvlpt_fmha_fwd(const void* qkv, void* out, void* lse, int N, int L, int d, int heads, int causal,
                              mvlpt_stream_t stream) {
    if (!qkv || !out || !lse) return fail(MVLPT_EINVAL, "mvlpt_fmha_fwd: null argument");
    if (N <= 0 || L <= 0 || heads <= 0 || d != heads * HD)
        return fail(MVLPT_ESHAPE, "mvlpt_fmha_fwd: need d == heads*64 (got d=%d heads=%d)", d, heads);
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // longer than the single-pass kernels hold (ViT-L/14@336px: 577 tokens + prompts): streaming kernels
    if (!fmha_sm100_supported(L)) return fmha_long_fwd(qkv, out, lse, N, L, d, heads, causal, s);
    return fmha_fwd_sm100(qkv, out, lse, N, L, d, heads, causal, s);
}

extern "C" int mvlpt_fmha_bwd(const void* qkv, const void* o, const void* d_o, const void* lse, void* dqkv, int N, int L,
                              int d, int heads, int causal, mvlpt_stream_t stream) {
    if (!qkv || !o || !d_o || !lse || !dqkv) return fail(MVLPT_EINVAL, "mvlpt_fmha_bwd: null argument");
    if (N <= 0 || L <= 0 || heads <= 0 || d != heads * HD)
        return fail(MVLPT_ESHAPE, "mvlpt_fmha_bwd: need d == heads*64 (got d=%d heads=%d)", d, heads);
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!fmha_bwd_sm100_supported(L)) return fmha_long_bwd(qkv, o, d_o, lse, dqkv, N, L, d, heads, causal, s);
    return fmha_bwd_sm100(qkv, o, d_o, lse, dqkv, N, L, d, heads, causal, s);
}

#ifdef MVLPT_FMHA_DBG
// Debug builds only: copy out and reset the in-kernel timestamp trace of the backward kernel (not part of the ABI).
extern "C" int mvlpt_dbg_fmha_trace(unsigned long long* out, int* counts) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_fmha_dbg, sizeof(unsigned long long) * 2 * 2048);
    cudaMemcpyFromSymbol(counts, g_fmha_dbg_n, sizeof(int) * 2);
    int zero[2] = {0, 0};
    cudaMemcpyToSymbol(g_fmha_dbg_n, zero, sizeof(zero));
    return 0;
}
#endif
