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
/* LayerNorm carried through the linears instead of run as a kernel (clip/model.py:185-188: `x + attn(ln_1(x))` feeds
 * ln_2, `x + mlp(ln_2(x))` feeds the next block's ln_1; LayerNorm itself clip/model.py:153-159).  With mean mu and
 * rstd of a row x:   LN(x).W^T + b = rstd * ( xt.W^T - (mu - c) * sg ) + bp,   xt[k] = (x[k] - c) * gamma[k] (fp16),
 * sg[n] = sum_k gamma[k] W[n,k],  bp[n] = b[n] + sum_k beta[k] W[n,k]  (fp32, precomputed per layer),  any centring c.
 *   producer side (rec_out != NULL): the residual-stream linear (out_f32, N == width) additionally writes xt of its
 *     output rows straight from the epilogue registers, centred on the mean of its residual rows (from rec_in; c = 0 when
 *     rec_in is NULL), and the partial sums of (x - c), (x - c)^2 into the row records rec_out;
 *   consumer side (rec != NULL): A is the xt of the rows to normalise (K == width); mean / rstd are finished from the
 *     records and the identity above is applied in the epilogue (then act / aux_out as in mvlpt_gemm); `bias` is bp
 *     (fp16 [N]) and `sg` fp16 [N].
 * Row record: MVLPT_LN_REC floats = 8 (sum, sum of squares) pairs of 128-column slices, the centring value at index 16.
 * Needs the CTA-pair kernel: M >= 256, N % 256 == 0, width % 256 == 0, width <= 1024 (mvlpt_gemm_ln_supported).
 * mvlpt_ln_prep starts the chain from a plain fp32 row block: xt = (x - mean) * gamma, record = (0, M2 | c = mean). */
#define MVLPT_LN_REC 20
typedef struct {
    const void* rec_in; /* producer: fp32 [M, MVLPT_LN_REC] records of the residual rows, or NULL */
    void* rec_out;      /* producer: fp32 [M, MVLPT_LN_REC] records of the output rows           */
    const void* gamma;  /* producer: fp32 [width] gamma of the LayerNorm that will consume `out`  */
    void* xt;           /* producer: fp16 [M, width]                                              */
    const void* rec;    /* consumer: fp32 [M, MVLPT_LN_REC] records of the A rows                 */
    const void* sg;     /* consumer: fp16 [N] (the folded bias bp goes in as `bias`)              */
    int width;          /* row width d of the normalised rows                                     */
    float eps;
} mvlpt_ln_carry;
int mvlpt_gemm_ln_supported(int M, int width);
int mvlpt_gemm_ln(const mvlpt_gemm_desc* d, const void* A, const void* W, const void* bias, const void* aux_in,
                  void* aux_out, const void* resid, void* out, const mvlpt_ln_carry* ln, mvlpt_stream_t stream);
int mvlpt_ln_prep(const void* x, const void* gamma, void* xt, void* rec, int rows, int d, mvlpt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Attention core, head width 64 (d == heads*64), any L (tcgen05 kernels up to 272 / 256 rows forward / backward,
 * streaming kernels beyond), packed qkv [N*L, 3d] fp16 (Q | K | V).
 * Replaces the inside of nn.MultiheadAttention (clip/model.py:171,181-183 -> F.multi_head_attention_forward:
 * q*hd^-1/2, QK^T, causal -inf mask for the text tower clip/model.py:324-330, softmax, PV) and its autograd.
 *   out  : fp16 [N*L, d]        lse : fp32 [N, heads, L]  (log-sum-exp of the scaled scores, saved for bwd)
 *   dqkv : fp16 [N*L, 3d]       d_o : fp16 [N*L, d]
 * ------------------------------------------------------------------------------------------------ */
int mvlpt_fmha_fwd(const void* qkv, void* out, void* lse, int N, int L, int d, int heads, int causal,
                   mvlpt_stream_t stream);
int mvlpt_fmha_bwd(const void* qkv, const void* o, const void* d_o, const void* lse, void* dqkv, int N, int L, int d,
                   int heads, int causal, mvlpt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm with fp32 statistics (clip/model.py:153-159).  x: fp32 residual stream [*, d]; y: fp16 [rows, d].
 * row_index (int32 [rows], may be NULL) gathers source rows: the CLS row (trainers/mvlpt.py:88) or the EOT
 * row (trainers/mvlpt.py:124-128).  d % 4 == 0, d <= 1024.
 * Backward (gamma/beta frozen, trainers/mvlpt.py:856-858): dx = rstd*(g - mean(g) - xhat*mean(g*xhat)),
 * g = dy*gamma; written (accumulate=0) or added (accumulate=1) to the fp32 gradient stream at the same rows,
 * with an optional fp16 copy dx16 of the result (operand of the next dgrad GEMM).  dx_stream NULL selects the
 * fp16 gradient stream: the running sum is read from and written to dx16 alone (the reference's own precision:
 * its residual gradients are fp16 tensors), which cuts the kernel's HBM traffic from 16 to 10 bytes per element.
 * ------------------------------------------------------------------------------------------------ */
int mvlpt_ln_fwd(const void* x, const void* row_index, const void* gamma, const void* beta, void* y, int rows, int d,
                 float eps, mvlpt_stream_t stream);
int mvlpt_ln_bwd(const void* dy, const void* x, const void* row_index, const void* gamma, void* dx_stream, void* dx16,
                 int rows, int d, float eps, int accumulate, mvlpt_stream_t stream);
/* hilo = 1: y is fp16 [rows, 2d] = [hi | lo], LN(x) = hi + lo to 2^-22: the A operand of a K = 2d mvlpt_gemm against
 * [W | W], used for the pooled CLS / EOT rows so that the feature projection (trainers/mvlpt.py:91,128) does not see
 * the fp16 rounding of its input.  hilo = 0 is mvlpt_ln_fwd. */
int mvlpt_ln_fwd_hilo(const void* x, const void* row_index, const void* gamma, const void* beta, void* y, int rows, int d,
                      float eps, int hilo, mvlpt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Image stem (trainers/mvlpt.py:53-58 = clip/model.py:220-225; conv1 clip/model.py:207).
 * mvlpt_im2col: img [B,3,H,W] (fp16, or fp32 when img_f32) -> patches fp16 [B*(H/p)*(W/p), Kp],
 *   column c*p*p + ky*p + kx (the conv weight's own flattening), columns [3pp, Kp) zero; Kp % 8 == 0.
 *   The conv itself is then mvlpt_gemm against conv1.weight viewed as [d, Kp].
 * mvlpt_embed_assemble: x0[b,0] = ln_pre(cls + pos[0]); x0[b,1..v] = prompt[0..v) (forward_vpt,
 *   trainers/mvlpt.py:416-437: prompts get neither pos-emb nor ln_pre); x0[b,1+v+i] = ln_pre(pe[b,i] + pos[1+i]).
 *   pe fp16 [B*G, d]; cls/pos/gamma/beta fp32; prompt fp16 or fp32 [v, d]; x0 fp32 [B, 1+v+G, d].
 * mvlpt_set_prompt_rows: x[b,1+j] = prompt[j] — the deep-prompt replacement before block l>=1
 *   (trainers/mvlpt.py:73-82).
 * mvlpt_prompt_grad: grad[j] = inv_scale * sum_b dx[b,1+j] (autograd of the expand over B); with zero_rows the
 *   rows are then cleared in dx (and dx16) because the replaced rows have no upstream.
 *   dx NULL = fp16 gradient stream: the rows are read from dx16.
 * vpt_dropout (nn.Dropout(VPT.DROPOUT), trainers/mvlpt.py:165; applied after the batch expansion at :76 and :425):
 *   with drop_p > 0, set_prompt_rows writes prompt[j,c] * keep(b,j,c) / (1-p) and prompt_grad sums
 *   dx[b,1+j,c] * keep(b,j,c) / (1-p), keep being a counter-based Bernoulli(1-p) draw (16-bit resolution in p) that
 *   depends only on (seed, slab, b, j, c, B, v, d) — the caller changes `seed` every step and passes the layer as
 *   `slab`; mvlpt_dropout_keep writes the same mask as uint8 [B, v, d] (tests; replaying a step).  drop_p = 0: no-op.
 *   Shallow prompts under dropout: embed_assemble, then set_prompt_rows on x0 with slab 0.
 * ------------------------------------------------------------------------------------------------ */
int mvlpt_im2col(const void* img, int img_f32, void* patches, int B, int H, int W, int p, int Kp, mvlpt_stream_t stream);
int mvlpt_embed_assemble(const void* pe, const void* cls, const void* pos, const void* gamma, const void* beta,
                         const void* prompt, int prompt_f16, void* x0, int B, int G, int v, int d, float eps,
                         mvlpt_stream_t stream);
int mvlpt_set_prompt_rows(void* x, const void* prompt, int prompt_f16, int B, int L, int v, int d, float drop_p,
                          uint64_t seed, int slab, mvlpt_stream_t stream);
/* Same, and for the rows just written xt[b,1+j] = (x - mean) * gamma (fp16 [B*L, d]) and their row records
 * (mvlpt_gemm_ln): the previous block's FC2 produced xt / records for the rows this call replaces.  xt NULL = plain
 * mvlpt_set_prompt_rows. */
int mvlpt_set_prompt_rows_ln(void* x, const void* prompt, int prompt_f16, int B, int L, int v, int d, float drop_p,
                             uint64_t seed, int slab, void* xt, void* rec, const void* gamma, mvlpt_stream_t stream);
int mvlpt_prompt_grad(void* dx, void* dx16, void* grad, int B, int L, int v, int d, float inv_scale, int zero_rows,
                      float drop_p, uint64_t seed, int slab, mvlpt_stream_t stream);
int mvlpt_dropout_keep(void* keep, int B, int v, int d, float drop_p, uint64_t seed, int slab, mvlpt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Text prompt assembly (forward_coop, trainers/mvlpt.py:439-515, fused with `+ positional_embedding`,
 * trainers/mvlpt.py:107/112).  emb fp32 [C, Lt, d] = token_embedding(tokenised "X .. X name.");
 * slot int32 [C, Lt]: >= 0 selects context vector `slot` (of class c when csc), -1 keeps emb.  The map encodes
 * end / middle / front placement, so no per-class Python loop is needed.  ctx NULL = fixed prompt.
 * mvlpt_ctx_grad: grad[j] = inv_scale * sum_c dx0[c, ctx_pos[c,j]]  (or per class when csc); ctx_pos int32 [C,n];
 *   dx0 is the fp32 gradient stream, or the fp16 one when dx_f16.
 * ------------------------------------------------------------------------------------------------ */
int mvlpt_text_assemble(const void* emb, const void* ctx, int ctx_f16, const void* slot, const void* pos, void* x0,
                        int C, int Lt, int n_ctx, int d, int csc, mvlpt_stream_t stream);
int mvlpt_ctx_grad(const void* dx0, int dx_f16, const void* ctx_pos, void* grad, int C, int Lt, int n_ctx, int d, int csc,
                   float inv_scale, mvlpt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Logit head (trainers/mvlpt.py:550-554, 573-581) and loss (trainers/mvlpt.py:914-916, 922/931).
 * mvlpt_l2norm_fwd: y = x/||x|| per row; x fp32 [rows,e] -> y16 fp16, y32 fp32, inv_norm fp32 [rows].
 * mvlpt_l2norm_bwd: dx16 = (dy - y (y.dy)) * inv_norm.
 * mvlpt_ce_fwd_bwd: logits fp32 [B, ldc] (C valid columns) are first multiplied in place by the task mask
 *   (task int32 [B], ranges int32 [T,2]; NULL = no mask), then loss_rows[b] = CE(row, target) with integer
 *   labels (int64 [B]) or soft targets (fp32 [B,C], rows normalised to sum 1 inside); pred[b] = argmax;
 *   hit[b] (int32, may be NULL) = pred[b] == label[b] (or == argmax of the soft row, trainers/mvlpt.py:935-936);
 *   dz16 fp16 [B, ldc] = coef * (softmax - y) * mask (may be NULL for evaluation).
 * mvlpt_step_metrics: out2[0] = inv_div * sum(loss_rows) (batch-mean loss), out2[1] = 100 * mean(hit) (top-1
 *   accuracy as dassl.metrics.compute_accuracy reports it, trainers/mvlpt.py:939-942); fp32 [2] on the device.
 * mvlpt_dlogits_prepare: dz16 = coef * dlogits * mask, for callers that compute their own loss (autograd).
 * mvlpt_transpose_f16: out[c,r] = in[r,c], zero padded to ld_out — operand layout for the head dgrads.
 * ------------------------------------------------------------------------------------------------ */
int mvlpt_l2norm_fwd(const void* x, void* y16, void* y32, void* inv_norm, int rows, int e, mvlpt_stream_t stream);
/* Same, plus y16x3 fp16 [rows, 3e] (may be NULL): the split y = hi + lo laid out for ONE mvlpt_gemm of K = 3e that sums
 * hi.hi' + hi.lo' + lo.hi' — the logit matmul (trainers/mvlpt.py:554) to fp32 accuracy on the fp16 tensor cores.
 * pattern 0 = [hi | hi | lo] (image side, the A operand), pattern 1 = [hi | lo | hi] (text side, the W operand). */
int mvlpt_l2norm_fwd_split(const void* x, void* y16, void* y32, void* inv_norm, void* y16x3, int pattern, int rows, int e,
                           mvlpt_stream_t stream);
int mvlpt_l2norm_bwd(const void* dy, const void* y32, const void* inv_norm, void* dx16, int rows, int e,
                     mvlpt_stream_t stream);
int mvlpt_ce_fwd_bwd(void* logits, int ldc, const void* label, const void* soft, const void* task, const void* ranges,
                     void* loss_rows, void* pred, void* hit, void* dz16, int B, int C, float coef,
                     mvlpt_stream_t stream);
int mvlpt_step_metrics(const void* loss_rows, const void* hit, int B, float inv_div, void* out2, mvlpt_stream_t stream);
int mvlpt_dlogits_prepare(const void* dlogits, int ld_in, const void* task, const void* ranges, void* dz16, int ldc,
                          int B, int C, float coef, mvlpt_stream_t stream);
int mvlpt_transpose_f16(const void* in, void* out, int R, int Cc, int ld_in, int ld_out, mvlpt_stream_t stream);
/* logits[b,c] *= 1[ranges[task[b]][0] <= c < ranges[task[b]][1]]  (trainers/mvlpt.py:573-581; multiplies by 0) */
int mvlpt_task_mask(void* logits, int ldc, const void* task, const void* ranges, int B, int C, mvlpt_stream_t stream);

/* SGD with momentum + L2 weight decay on one prompt tensor (torch.optim.SGD as Dassl builds it for
 * trainers/mvlpt.py:869: g += wd*p; buf = mu*buf + g (buf = g on the first step); p -= lr*buf).
 * p/buf: fp16 (is_f16) or fp32; g: fp32 unscaled gradient. */
int mvlpt_sgd(void* p, void* buf, const void* g, int n, int is_f16, float lr, float momentum, float wd, int first_step,
              mvlpt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * UPT shared-prompt projection (MultitaskVLPromptLearner.forward_mvlpt_proj, trainers/mvlpt.py:376-414, modules
 * :234-257, PROJECT_METHOD='transformer') — forward, and backward including the WEIGHT gradients of the projection
 * modules (the only trainable weights of the step).  The reference's 1-layer "transformer" sees sequence length 1
 * (SURVEY.md App. C): per token x += W_o(W_v LN1(x) + b_v) + b_o; x += W_pr quickgelu(W_fc LN2(x) + b_fc) + b_pr.
 *   params / grads: arrays of MVLPT_UPT_NPARAM device pointers in the order of the enum below.  ctx, vpt, vpt_deep and
 *   the four pre/post Linear tensors are fp16 when param_f16 (CLIP dtype) else fp32; the block's own tensors are always
 *   fp32 (trainers/mvlpt.py:256-259).  Every gradient is fp32, same shape as its parameter.
 *   ctx_out fp32 [n_ctx, dt]; vpt_out fp32 [(1+n_deep)*v, dv] (row block 0 = shallow prompts, then the deep ones).
 *   workspace: mvlpt_upt_workspace(d) bytes; the forward leaves the activations the backward needs in it.
 * ------------------------------------------------------------------------------------------------ */
enum {
    MVLPT_UPT_CTX = 0, MVLPT_UPT_VPT, MVLPT_UPT_VPT_DEEP,
    MVLPT_UPT_COOP_PRE_W, MVLPT_UPT_COOP_PRE_B, MVLPT_UPT_COOP_POST_W, MVLPT_UPT_COOP_POST_B,
    MVLPT_UPT_VPT_PRE_W, MVLPT_UPT_VPT_PRE_B, MVLPT_UPT_VPT_POST_W, MVLPT_UPT_VPT_POST_B,
    MVLPT_UPT_LN1_G, MVLPT_UPT_LN1_B, MVLPT_UPT_IN_W, MVLPT_UPT_IN_B, MVLPT_UPT_OUT_W, MVLPT_UPT_OUT_B,
    MVLPT_UPT_LN2_G, MVLPT_UPT_LN2_B, MVLPT_UPT_FC_W, MVLPT_UPT_FC_B, MVLPT_UPT_PROJ_W, MVLPT_UPT_PROJ_B,
    MVLPT_UPT_NPARAM
};
typedef struct {
    int n_ctx;     /* rows of ctx                                    */
    int v;         /* visual prompts per layer                       */
    int n_deep;    /* rows of vpt_embeddings_deep / v (0 = shallow)  */
    int dt, dv;    /* text / vision widths                           */
    int pd;        /* PROJECT_DIM                                    */
    int param_f16; /* dtype of ctx / vpt / pre / post tensors        */
} mvlpt_upt_desc;

size_t mvlpt_upt_workspace(const mvlpt_upt_desc* d);
int mvlpt_upt_fwd(const mvlpt_upt_desc* d, const void* const* params, void* workspace, size_t ws_bytes, void* ctx_out,
                  void* vpt_out, mvlpt_stream_t stream);
int mvlpt_upt_bwd(const mvlpt_upt_desc* d, const void* const* params, void* workspace, size_t ws_bytes,
                  const void* d_ctx_out, const void* d_vpt_out, void* const* grads, mvlpt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * vpt_proj: the Linear(VPT.PROJECT -> vision width) between the stored visual prompts and the rows the image tower sees
 * (trainers/mvlpt.py:170-175 construction; applied at :76-77 to every deep-prompt slab and at :425 to the shallow one).
 *   fwd: out[rows, d] fp32 = emb[rows, p] . W[d, p]^T + b[d]                 (emb, W, b fp16 when param_f16 else fp32)
 *   bwd: from d_out[rows, d] fp32 (what mvlpt_prompt_grad produced): d_emb[rows, p] = d_out . W  (overwritten),
 *        dW[d, p] (+)= d_out^T . emb, db[d] (+)= colsum(d_out); `accumulate` != 0 adds into dW / db (second slab).
 * ------------------------------------------------------------------------------------------------ */
int mvlpt_vpt_proj_fwd(const void* emb, const void* W, const void* b, int param_f16, void* out, int rows, int d, int p,
                       mvlpt_stream_t stream);
int mvlpt_vpt_proj_bwd(const void* d_out, const void* emb, const void* W, int param_f16, void* d_emb, void* dW, void* db,
                       int rows, int d, int p, int accumulate, mvlpt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Input pipeline in front of the image tower: decoded uint8 RGB images -> the normalised batch.
 * Replaces, bit for bit, the reference's CPU transforms: torchvision `Resize(SIZE, BICUBIC)` [+ `CenterCrop`] ->
 * `ToTensor` -> `Normalize(PIXEL_MEAN, PIXEL_STD)` (trainers/vision_benchmark/evaluation/feature.py:540-553) and Dassl's
 * `RandomResizedCrop` -> `RandomHorizontalFlip` -> ToTensor -> Normalize for configs/trainers/MVLPT/vit_b16.yaml:8-13,
 * i.e. Pillow's two-pass fixed-point bicubic `Image.resize` (src/libImaging/Resample.c) on the crop box.
 *   src: device buffer holding the images as uint8 [H, W, 3] rows packed, image b at byte offset descs[b].src_off;
 *     16-byte aligned and readable up to the next multiple of 16 behind every image (rows are fetched as aligned uint4).
 *   descs: one descriptor per image, given BOTH as a host array (launch planning) and as a device copy (kernels).
 *   out: [B, 3, out_h, out_w] fp32, or fp16 (fp32 result rounded to nearest) when out_f16.
 *   mean3 / std3: host pointers to 3 floats.  workspace: mvlpt_preprocess_workspace(...) bytes of device memory.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    uint64_t src_off;   /* byte offset of pixel (0,0) of this image in src                                   */
    int H, W;           /* decoded size                                                                      */
    int by, bx, bh, bw; /* crop box (top, left, height, width) taken BEFORE resizing; whole image = 0,0,H,W   */
    int rh, rw;         /* size the box is resized to                                                        */
    int oy, ox;         /* top-left of the out_h x out_w window kept from the resized box (CenterCrop); else 0 */
    int flip;           /* horizontal flip of the window                                                     */
} mvlpt_image_desc;

/* The tail of the same stack for a batch that already has the model's size (the workers cropped / resized, or the data is
 * stored that way): uint8 [B, 3, H, W] -> ToTensor -> Normalize, bit-identical to torchvision (two IEEE divisions); out is
 * [B, 3, H, W] fp16 (out_f16) or fp32.  H*W % 16 == 0.  Lets a loader ship 1 byte per value across PCIe instead of 2 or 4
 * (the reference's loaders ship fp32 tensors, trainers/mvlpt.py:959-960). */
int mvlpt_normalize_u8(const void* src, void* out, int out_f16, int B, int H, int W, const float* mean3, const float* std3,
                       mvlpt_stream_t stream);
size_t mvlpt_preprocess_workspace(const mvlpt_image_desc* descs_host, int B, int out_h, int out_w);
int mvlpt_preprocess(const void* src, const mvlpt_image_desc* descs_host, const mvlpt_image_desc* descs_dev, int B,
                     const float* mean3, const float* std3, void* out, int out_f16, int out_h, int out_w, void* workspace,
                     size_t ws_bytes, mvlpt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * CoCoOp branch (trainers/mvlpt.py:260-290 meta_net, :348-374 forward_cocoop, :556-571 per-image text features).
 * The text tower itself runs the usual kernels on B*C sequences; these are the pieces around it (fp32 arithmetic;
 * parameters fp16 or fp32 by param_f16 / ctx_f16).
 * mvlpt_metanet_fwd: h1[B,H] = relu(imf . W1^T + b1), bias[B,dt] = h1 . W2^T + b2   (imf = L2-normalised image features)
 * mvlpt_metanet_bwd: from d_bias[B,dt]: dW1[H,e], db1[H], dW2[dt,H], db2[dt] (overwritten) and
 *   d_imf[B,e] += d_imf_scale * d(loss)/d(imf); d_h1_ws = fp32 [B,H] scratch.
 * mvlpt_cocoop_assemble: x0[(b,c),t] = (slot[c,t] >= 0 ? ctx[slot] + bias[b] : emb[c,t]) + pos[t]
 *   (construct_prompts with the instance-shifted context, :362-372, fused with `+ positional_embedding`, :107).
 * mvlpt_pair_logits_fwd/bwd: logits[b,c] = scale * <img[b], txt[b*C+c]> (:566-569) and its gradients
 *   d_txt[(b,c)] = scale * dz[b,c] * img[b], d_img[b] = scale * sum_c dz[b,c] * txt[(b,c)]  (dz fp16 [B, ldc]).
 * mvlpt_cocoop_ctx_grad: d_ctx[j] = inv_scale * sum_{b,c} dx[(b,c), ctx_pos[c,j]] and
 *   d_bias[b] = inv_scale * sum_{c,j} dx[(b,c), ctx_pos[c,j]]; dx fp16 [B*C*Lk, d]; part_ws = fp32 [B, n_ctx, d] scratch.
 * ------------------------------------------------------------------------------------------------ */
int mvlpt_metanet_fwd(const void* imf, const void* W1, const void* b1, const void* W2, const void* b2, int param_f16,
                      void* h1, void* bias, int B, int e, int H, int dt, mvlpt_stream_t stream);
int mvlpt_metanet_bwd(const void* d_bias, const void* imf, const void* h1, const void* W1, const void* W2, int param_f16,
                      void* d_h1_ws, void* dW1, void* db1, void* dW2, void* db2, void* d_imf, float d_imf_scale, int B, int e,
                      int H, int dt, mvlpt_stream_t stream);
int mvlpt_cocoop_assemble(const void* emb, const void* ctx, int ctx_f16, const void* bias, const void* slot, const void* pos,
                          void* x0, int B, int C, int Lk, int d, mvlpt_stream_t stream);
int mvlpt_pair_logits_fwd(const void* img, const void* txt, float scale, void* logits, int ldc, int B, int C, int e,
                          mvlpt_stream_t stream);
int mvlpt_pair_logits_bwd(const void* dz16, int ldc, const void* img, const void* txt, float scale, void* d_txt, void* d_img,
                          int B, int C, int e, mvlpt_stream_t stream);
int mvlpt_cocoop_ctx_grad(const void* dx16, const void* ctx_pos, void* part_ws, void* d_ctx, void* d_bias, int B, int C, int Lk,
                          int n_ctx, int d, float inv_scale, mvlpt_stream_t stream);

/* cudaMemsetAsync(p, 0, bytes) on the stream. */
int mvlpt_zero(void* p, size_t bytes, mvlpt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MVLPT_SM100_H */
