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

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: remember, per kernel (one cache
// object per launch site) and per device, the largest size already granted.
struct DynSmemCache {
    std::atomic<size_t> granted[64];
};
template <typename K>
inline int ensure_dyn_smem(K kernel, size_t bytes, DynSmemCache& c) {
    int dev = 0;
    MVLPT_CUDA_OK(cudaGetDevice(&dev));
    std::atomic<size_t>& g = c.granted[dev & 63];
    if (bytes > g.load(std::memory_order_relaxed)) {
        MVLPT_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        g.store(bytes, std::memory_order_relaxed);
    }
    return MVLPT_OK;
}

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

// Programmatic dependent launch (PDL).  A kernel launched with `launch_pdl` may begin while its predecessor in the
// stream is still draining: its CTAs are scheduled as soon as the predecessor's CTAs have all started and SM resources
// free up, run their prologue (barrier init, TMEM allocation, descriptor prefetch) and then block in `pdl_wait()` until
// the predecessor has completed and its memory is visible.  Every kernel launched this way executes
// `pdl_launch_dependents(); pdl_wait();` before its first global-memory access.  MVLPT_PDL=0 turns the attribute off.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              unsigned cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    unsigned n = 0;
    if (cluster_x > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = cluster_x;
        at[n].val.clusterDim.y = 1;
        at[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled()) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
int sm_count();

}  // namespace mvlpt
