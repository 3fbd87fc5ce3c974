/*
 * mvlpt_sm100.h — C ABI of libmvlpt_sm100.so: the B200 (sm_100a) kernels behind the MVLPT prompt-tuning
 * hot path.  Plain pointers and sizes only; every pointer is a DEVICE pointer unless it says "host".
 * Every call is asynchronous on the CUDA stream passed in (0 = legacy default stream).
 *
 * The reference (sIncerass/MVLPT) has no FFI: its hot path is Python over torch ops.  Each entry point
 * below names the reference call site(s) whose arithmetic it replaces (file:line under the reference
 * tree).  The Python mirror of the reference classes (mvlpt_b200/trainers/mvlpt.py) binds these with
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Error convention: 0 = OK, <0 = one of MVLPT_E*; mvlpt_last_error() returns a thread-local message.
 * Nothing throws or aborts across the boundary; there is no CPU fallback — on a non-sm_100 device every
 * compute entry returns MVLPT_EARCH.
 */
#ifndef MVLPT_SM100_H
#define MVLPT_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVLPT_ABI_VERSION 1

#define MVLPT_OK 0
#define MVLPT_EINVAL (-1) /* bad argument (null pointer, misaligned, negative size) */
#define MVLPT_ESHAPE (-2) /* shape not supported by the kernel                      */
#define MVLPT_EARCH (-3)  /* device is not sm_100 / driver lacks a needed entry     */
#define MVLPT_ECUDA (-4)  /* a CUDA runtime / driver call failed                    */

typedef void* mvlpt_stream_t; /* cudaStream_t */

int mvlpt_version(void);
const char* mvlpt_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
uint64_t mvlpt_launch_count(void);
/* 0 if device `dev` can run this library (compute capability 10.x), else MVLPT_EARCH / MVLPT_ECUDA. */
int mvlpt_check_device(int dev);

/* ------------------------------------------------------------------------------------------------
 * Dense linear:  out[M,N] = epi( alpha * A[M,K] . W[N,K]^T )      fp16 operands, fp32 accumulation.
 * Replaces nn.Linear / MultiheadAttention in_proj,out_proj / mlp.c_fc,c_proj (clip/model.py:171-177,
 * 181-188), `x @ proj` (trainers/mvlpt.py:91,128), the logit matmul (trainers/mvlpt.py:554) and, with a
 * transposed weight copy, their autograd dgrads.  tcgen05 + TMA kernel (csrc/gemm_sm100.cuh).
 *   lda, ldw  : row strides in elements, multiples of 8
 *   bias      : fp16 [N] or NULL
 *   act       : 0 none | 1 QuickGELU (clip/model.py:162-164) | 2 multiply by QuickGELU'(aux_in)
 *   aux_in    : fp16 [M,ld_aux] (act 2)       aux_out: fp16 [M,ld_aux] pre-activation store (act 1) or NULL
 *   resid     : fp32 [M,ld_out] added after activation, or NULL (may alias out)
 *   out       : fp16 (out_f32=0) or fp32 (out_f32=1), row stride ld_out (multiple of 8 / 4)
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int M, N, K;
    int lda, ldw, ld_out, ld_aux;
    int act;
    int out_f32;
    float alpha;
} mvlpt_gemm_desc;

int mvlpt_gemm(const mvlpt_gemm_desc* d, const void* A, const void* W, const void* bias, const void* aux_in,
               void* aux_out, const void* resid, void* out, mvlpt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MVLPT_SM100_H */
