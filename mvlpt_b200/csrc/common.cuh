// Host-side helpers shared by every C-ABI entry point: error slot, launch counter, TMA descriptor encode.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/mvlpt_sm100.h"

namespace mvlpt {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define MVLPT_CUDA_OK(expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) return fail(MVLPT_ECUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

// Check the launch that just happened and count it.
inline int launched(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(MVLPT_ECUDA, "%s launch: %s", what, cudaGetErrorString(e));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return MVLPT_OK;
}

// 0 when the current device is sm_100-class (cached per device).
int require_sm100();

// Encode a tiled fp16 TMA descriptor with 128B swizzle.  dims/box are innermost-first.
// strides_bytes has rank-1 entries (stride of dim 1, dim 2, ...).
int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);
// Same for fp32 (is_f32) or fp16 elements; the innermost box extent times the element size must be 128 bytes.
int make_tmap(CUtensorMap* out, const void* base, int is_f32, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box);

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
int sm_count();

}  // namespace mvlpt
