// tcgen05 attention forward for the image tower (non-causal, head width 64).  Placeholder until the TMEM
// single-pass kernel lands: reports "unsupported" so mvlpt_fmha_fwd uses the HMMA kernel in fmha.cu.
#pragma once
#include "common.cuh"

namespace mvlpt {
inline bool fmha_sm100_supported(int /*L*/) { return false; }
inline int fmha_fwd_sm100(const void*, void*, void*, int, int, int, int, float, cudaStream_t) {
    return fail(MVLPT_ESHAPE, "fmha_fwd_sm100: not built");
}
}  // namespace mvlpt
