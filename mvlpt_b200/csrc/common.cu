#include "common.cuh"
#include <stdlib.h>

#include <string.h>

namespace mvlpt {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::atomic<int> state{0};  // 0 unknown, 1 ok, 2 missing
    if (state.load() == 1) return fn;
    if (state.load() == 2) return nullptr;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        state.store(2);
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
    state.store(1);
    return fn;
}

int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
    return make_tmap(out, base, 0, rank, dims, strides_bytes, box);
}

int make_tmap(CUtensorMap* out, const void* base, int is_f32, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return fail(MVLPT_EARCH, "driver has no cuTensorMapEncodeTiled");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(MVLPT_EINVAL, "TMA base %p not 16-byte aligned", base);
    cuuint64_t gdim[5];
    cuuint64_t gstr[5];
    cuuint32_t bx[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        estr[i] = 1;
        if (i > 0) {
            gstr[i - 1] = strides_bytes[i - 1];
            if (gstr[i - 1] % 16) return fail(MVLPT_EINVAL, "TMA stride %llu not a multiple of 16 bytes",
                                              (unsigned long long)gstr[i - 1]);
        }
    }
    CUresult r = enc(out, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MVLPT_ECUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return MVLPT_OK;
}

static int g_dev_ok[64];  // 0 unknown, 1 ok, -1 bad
static int g_sms[64];

int require_sm100() {
    int dev = 0;
    MVLPT_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(MVLPT_EINVAL, "device index %d out of range", dev);
    if (g_dev_ok[dev] == 1) return MVLPT_OK;
    int major = 0, sms = 0;
    MVLPT_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    MVLPT_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (major != 10) {
        g_dev_ok[dev] = -1;
        return fail(MVLPT_EARCH, "device %d has compute capability %d.x; this library is sm_100a only", dev, major);
    }
    g_sms[dev] = sms;
    g_dev_ok[dev] = 1;
    return MVLPT_OK;
}

int sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || g_sms[dev] == 0) return 148;
    return g_sms[dev];
}

}  // namespace mvlpt

using namespace mvlpt;

extern "C" {

int mvlpt_version(void) { return MVLPT_ABI_VERSION; }
const char* mvlpt_last_error(void) { return g_err; }
uint64_t mvlpt_launch_count(void) { return g_launches.load(); }

int mvlpt_check_device(int dev) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(MVLPT_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    if (dev < 0 || dev >= n) return fail(MVLPT_EINVAL, "no CUDA device %d (count %d)", dev, n);
    int major = 0;
    MVLPT_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) return fail(MVLPT_EARCH, "device %d is sm_%d0, need sm_100", dev, major);
    return MVLPT_OK;
}

}  // extern "C"

namespace mvlpt {
bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("MVLPT_PDL");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on != 0;
}
}  // namespace mvlpt
