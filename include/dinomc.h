/*
 * dinomc.h -- C ABI of libdinomc.so: the B200 (sm_100a) kernels behind the DINO-MC
 * head + loss + center + EMA step.
 *
 * The reference (HaykSahakyan11/Self-Supervised-Learning-for-Aerial-Image-Segmentation) has no
 * native layer: its hot path is eager PyTorch.  Each entry point below therefore cites the
 * reference *Python* lines whose arithmetic it replaces (paths relative to the reference root).
 * The Python drop-in modules (DINOHead / DINOLoss / ema_update_) bind these with ctypes; see
 * INTEGRATION.md for the stub a maintainer of the reference would add.
 *
 * Conventions (every entry point):
 *   - plain C types only; all tensor arguments are raw DEVICE pointers owned by the caller
 *     (the library never frees or retains them) unless the name ends in `_host`;
 *   - matrices are row-major with an explicit leading dimension in ELEMENTS;
 *   - `stream` is a cudaStream_t passed as void*; work is only enqueued, the library never
 *     synchronises the device and never allocates: scratch comes from a caller-supplied
 *     workspace whose size is returned by the matching *_workspace_bytes() query;
 *   - return value: 0 = ok, < 0 = invalid argument (see dmc_last_error_string()),
 *     > 0 = a cudaError_t raised while enqueuing;
 *   - re-entrant and stream-safe: no global mutable state except the thread-local error string.
 */
#ifndef DINOMC_H_
#define DINOMC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMC_VERSION 100 /* major*100 + minor */

/* element types of tensor arguments */
enum { DMC_F32 = 0, DMC_BF16 = 1 };

/* GEMM epilogue activation */
enum {
  DMC_ACT_NONE = 0,
  DMC_ACT_GELU = 1,     /* D = gelu(z); if aux != NULL also aux = z (pre-activation, saved for backward) */
  DMC_ACT_GELU_BWD = 2, /* D = z * gelu'(aux)   (aux = the saved pre-activation)                        */
  DMC_ACT_GELU_DG = 4,  /* D = gelu(z) and aux = gelu'(z) (one evaluation; aux required): the layer's backward then uses    */
  DMC_ACT_MUL_AUX = 5,  /* D = z * aux            -- a plain multiply in the dgrad epilogue instead of a GELU' evaluation */
  DMC_ACT_NORMALIZE_BWD = 3 /* backward of F.normalize (utils/vision_transformer.py:292) fused into the split-K reduction of
                               the last layer's dgrad: D[m,:] = (z[m,:] - (z[m,:] . aux[m,:]) aux[m,:]) * row_scale[m], with
                               aux = the normalised rows (fp32) and row_scale = 1/max(||row||, eps); rows clamped by eps
                               (row_scale * row_eps >= 1) get no projection term.  N <= 1024; the contraction is always split. */
};

int dmc_version(void);
/* Message for the last non-zero return on the calling thread ("" if none). */
const char* dmc_last_error_string(void);
/* 0 if device `device` is an sm_100 part this library can run on, else an error. */
int dmc_device_check(int device);
/* Programmatic dependent launch.  Every kernel of the library begins with griddepcontrol.launch_dependents +
 * griddepcontrol.wait, and by default is launched with the programmatic-stream-serialization attribute so that
 * its launch latency and prologue overlap the tail of the previous kernel on the stream (data dependencies are
 * unchanged: no kernel touches global memory before the wait).  dmc_set_pdl(0) launches without the attribute,
 * e.g. while NCCL kernels share the SMs; the environment variable DMC_PDL=0 sets the initial value.  Returns the
 * previous setting. */
int dmc_set_pdl(int enabled);
/* Grid cap for the row-streaming kernels (dmc_weightnorm_fwd, dmc_weightnorm_bwd[_bf16]) launched from the CALLING
 * THREAD until changed again; 0 = no cap (default).  These kernels are HBM-bound and normally launch thousands of CTAs,
 * which keep every SM full: a one-CTA-per-SM tcgen05 GEMM queued behind them cannot start before they drain.  Capped at
 * one CTA per SM (148) they walk their rows with a grid-stride loop and leave registers and shared memory for a GEMM CTA
 * on every SM, so an auxiliary-stream weight-norm pass (utils/vision_transformer.py:279) runs NEXT TO the MLP GEMMs.
 * Returns the previous cap. */
int dmc_set_streaming_ctas(int n);

/* ---------------------------------------------------------------------------------------------
 * GEMM:  D[M,N] = epilogue( sum_k A(m,k) * B(n,k) ),  fp32 accumulation.
 *
 * Replaces the nn.Linear calls of DINOHead.forward (utils/vision_transformer.py:291,293: the MLP
 * and the weight-normed last layer x @ W^T) and their autograd backward (dgrad / wgrad), reached
 * from main_dino_mc.py:386/393.
 *
 * Operand layout is described per operand:  *_mn_major == 0 : the contraction index k is the
 * contiguous one (A stored [M,K], B stored [N,K]);  == 1 : the M (resp. N) index is contiguous
 * (A stored [K,M], B stored [K,N]).  ld* is the row stride of the stored matrix, in elements.
 *   forward  x @ W^T ............ A=x [M,K] k-major,      B=W [N,K] k-major
 *   dgrad    dY @ W .............. A=dY [M,K] k-major,     B=W stored [K,N] -> b_mn_major=1
 *   wgrad    dY^T @ X ............ A=dY stored [K,M] -> a_mn_major=1, B=X stored [K,N] -> b_mn_major=1
 *
 * in_dtype == DMC_BF16 : one bf16 tcgen05 pass (kind::f16).
 * in_dtype == DMC_F32  : operands are fp32 read as TF32 (kind::tf32).  If A_lo and B_lo are given,
 *                        A = A + A_lo, B = B + B_lo (hi/lo split made by dmc_split_tf32) and three
 *                        passes hi*hi + hi*lo + lo*hi are accumulated ("3xTF32", ~fp32 accuracy).
 * Epilogue, in this order:  z = acc * col_scale[n] * alpha * (*alpha_dev) + bias[n];  activation;
 * store as out_dtype.  Every pointer in the epilogue may be NULL (= skipped).
 * ------------------------------------------------------------------------------------------- */
typedef struct dmc_gemm_args {
  int64_t M, N, K;
  const void* A;    int64_t lda; int32_t a_mn_major;
  const void* B;    int64_t ldb; int32_t b_mn_major;
  const void* A_lo; /* fp32 residuals for 3xTF32, or NULL */
  const void* B_lo;
  int32_t in_dtype;   /* DMC_BF16 | DMC_F32 */
  void* D;          int64_t ldd; int32_t out_dtype;
  const float* col_scale;  /* [N] or NULL : the weight-norm factor g_k/||v_k|| when folded in here */
  const float* bias;       /* [N] or NULL */
  const float* alpha_dev;  /* device scalar or NULL (upstream grad / loss scale) */
  float alpha;             /* host scalar, 1.0f for none */
  int32_t act;             /* DMC_ACT_* */
  void* aux;        int64_t ldaux; int32_t aux_dtype;
  int32_t split_k;         /* 0 = let the library choose, >= 1 = forced */
  void* workspace;  size_t workspace_bytes; /* needed when split_k != 1, see query */
  int32_t max_ctas;        /* 0 = one persistent CTA per SM (148); > 0 = cap, e.g. to leave SMs to a concurrent
                              NCCL all-reduce kernel (the persistent CTAs would otherwise queue behind it) */
  /* Optional statistics of the STORED output, fused into the epilogue (last-layer forward; plain epilogue,
   * K-major operands, N > 128).  With y2 = (D[m,n] - stat_center[n]) * stat_scale * log2(e):
   *   stat_row_partials[m][part] = { max_n y2, sum_n 2^(y2 - max) } over column part `part` (128 columns each,
   *   dmc_gemm_stats_parts(N) parts per row) -- merged by dmc_teacher_finalize / dmc_lse_finalize;
   *   stat_colsum_partials[g][n] = sum of D over the rows of 32-row group g (NULL = not wanted).
   * This replaces DINOLoss's separate statistics passes over the logits (main_dino_mc.py:446,456,468). */
  float stat_scale; const float* stat_center; float* stat_row_partials; float* stat_colsum_partials;
  const float* stat_bound; /* optional device scalar b >= max |D| (e.g. the largest weight-norm gain when the rows of A
                              are unit vectors): lets the epilogue skip the running max (used only without a center) */
  const float* stat_bound2; /* optional second device scalar ADDED to stat_bound; with stat_center set, stat_bound + stat_bound2 must
                               bound |D| + |center| (e.g. max gain + max |center|): the TEACHER statistics then use the same
                               fixed-shift form (bf16 output; while the shift stays below 55, else the running-maximum form) */
  const float* row_scale;  /* DMC_ACT_NORMALIZE_BWD: [M] */
  float row_eps;           /* DMC_ACT_NORMALIZE_BWD: the eps of F.normalize */
} dmc_gemm_args;

/* Number of 128-column parts per row that the fused statistics produce for an N-column output. */
int64_t dmc_gemm_stats_parts(int64_t N);

/* Upper bound of the split-K workspace dmc_gemm may need for this problem. */
size_t dmc_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int32_t in_dtype);
/* tcgen05 / TMEM / TMA tile kernel (sm_100a). */
int dmc_gemm(const dmc_gemm_args* args, void* stream);
/* Same contract on the fp32 FMA pipes (no tensor cores, A_lo/B_lo ignored, in_dtype must be F32):
 * the exact-fp32 arm used to cross-check the tensor-core kernel on the device. */
int dmc_gemm_simt(const dmc_gemm_args* args, void* stream);

/* Split fp32 `x` into hi = tf32-rounded x (low 13 mantissa bits zero) and lo = x - hi. */
int dmc_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream);
/* fp32 -> bf16 (round to nearest even). */
int dmc_cast_f32_to_bf16(const float* x, void* y_bf16, int64_t n, void* stream);
/* The same for up to 8 tensors in ONE launch (the bf16 operand copies of a head forward: features + Linear weights).
 * The three arrays are HOST arrays of `count` entries. */
int dmc_cast_f32_to_bf16_batch(const float* const* srcs_host, void* const* dsts_host, const int64_t* ns_host,
                               int32_t count, void* stream);
/* The way back (bf16 -> fp32, exact) for up to 8 tensors in one launch: gradients after a bf16 all-reduce
 * (the compressed form of the exchange at main_dino_mc.py:260).  HOST arrays of `count` entries. */
int dmc_cast_bf16_to_f32_batch(const void* const* srcs_host, float* const* dsts_host, const int64_t* ns_host,
                               int32_t count, void* stream);
/* out[n] = sum_m X[m,n]   (bias gradients of the MLP Linears).  X is F32 or BF16. */
size_t dmc_colsum_workspace_bytes(int64_t M, int64_t N);
int dmc_colsum(const void* X, int32_t dtype, int64_t M, int64_t N, int64_t ld, float* out,
               void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Row kernels of the head.
 * ------------------------------------------------------------------------------------------- */
/* F.normalize(x, dim=-1, p=2) (utils/vision_transformer.py:292): zhat = z / max(||z||_2, eps).
 * Writes any of zhat_f32 / zhat_bf16 / zhat_lo (tf32 residual of zhat_f32) that is non-NULL and
 * inv_den[N] = 1/max(||z||,eps) (saved for backward).  z is F32 or BF16 [N,dim], row stride ld. */
int dmc_normalize_rows_fwd(const void* z, int32_t z_dtype, int64_t n_rows, int64_t dim, int64_t ld, float eps,
                           float* zhat_f32, void* zhat_bf16, float* zhat_lo, float* inv_den, void* stream);
/* Backward of the above: dz = (dzhat - (dzhat . zhat) zhat) * inv_den (rows whose norm was clamped to
 * eps get dz = dzhat * inv_den).  dz is F32 or BF16. */
int dmc_normalize_rows_bwd(const float* dzhat, const float* zhat, const float* inv_den, int64_t n_rows,
                           int64_t dim, float eps, void* dz, int32_t dz_dtype, void* stream);
/* nn.utils.weight_norm pre-hook (utils/vision_transformer.py:279): w_k = g_k * v_k / ||v_k||_2 for every
 * output row k of v [K,dim].  Writes any non-NULL of w_f32 (tf32-rounded when w_lo is given) / w_lo /
 * w_bf16, and scale[K] = g_k/||v_k||, inv_vnorm[K] = 1/||v_k|| (saved for backward). */
int dmc_weightnorm_fwd(const float* v, const float* g, int64_t K, int64_t dim,
                       float* w_f32, float* w_lo, void* w_bf16, float* scale, float* inv_vnorm, float* gmax, void* stream);
/* gmax (optional device scalar) receives max_k |g_k|, the bound of every logit when the activations are unit rows. */
/* Backward: dv = scale * (dw - (dw . vhat) vhat), dg = dw . vhat (dg may be NULL: frozen gain,
 * utils/vision_transformer.py:281-282). */
int dmc_weightnorm_bwd(const float* dw, const float* v, const float* scale, const float* inv_vnorm,
                       int64_t K, int64_t dim, float* dv, float* dg, void* stream);
/* The same pass reading dW stored as BF16.  Used by the data-parallel bf16 gradient exchange (the all-reduce DDP performs
 * at main_dino_mc.py:260): the pass is linear in dW, so the ranks average dW in bf16 (half the bytes of dv) and every
 * rank then computes the already-averaged dv / dg from it. */
int dmc_weightnorm_bwd_bf16(const void* dw_bf16, const float* v, const float* scale, const float* inv_vnorm,
                            int64_t K, int64_t dim, float* dv, float* dg, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DINOLoss (main_dino_mc.py:419-473).  Rows are crop-major: row v*B + b = crop v of sample b
 * (the layout MultiCropWrapper guarantees, utils/utils.py:627-646).  Logits are F32 or BF16.
 * ------------------------------------------------------------------------------------------- */
/* One pass over the teacher logits t [Nt,K] (main_dino_mc.py:446 and :468): per-row statistics of
 * softmax((t - center) * inv_temp), kept in the base-2 domain the loss kernels compute in:
 * row_stats[Nt][2] = {max_k y2, 1 / sum_k 2^(y2 - max)} with y2 = (t - center) * inv_temp * log2(e);
 * and the per-GPU batch column sum colsum[K] = sum_rows t (input of the center all-reduce). */
size_t dmc_teacher_workspace_bytes(int64_t Nt, int64_t K);
int dmc_teacher_stats_colsum(const void* t, int32_t dtype, int64_t Nt, int64_t K, int64_t ld,
                             const float* center, float inv_temp, float* row_stats, float* colsum,
                             void* workspace, size_t workspace_bytes, void* stream);
/* Same pass with caller-supplied bounds: bounds_dev[0] >= max |t| (for logits of unit rows against weight-normed rows: the
 * largest gain, utils/vision_transformer.py:279,292), bounds_dev[1] >= max |center| (device scalars, e.g. from dmc_absmax).
 * While (b0 + b1) * inv_temp * log2(e) stays below 55 the row sums use that fixed shift instead of the row maximum -- one
 * pass over the registers, about half the instructions of the general form (the pass is instruction-bound) -- and the
 * reported "maximum" of a row is the shift; otherwise identical to dmc_teacher_stats_colsum. */
int dmc_teacher_stats_colsum_bounded(const void* t, int32_t dtype, int64_t Nt, int64_t K, int64_t ld,
                                     const float* center, float inv_temp, const float* bounds_dev, float* row_stats,
                                     float* colsum, void* workspace, size_t workspace_bytes, void* stream);
/* out[0] = max_i |x[i]| (fp32, single CTA: for vectors of at most a few hundred thousand entries). */
int dmc_absmax(const float* x, int64_t n, float* out, void* stream);
/* Same outputs as dmc_teacher_stats_colsum, but from the partials the last-layer GEMM's epilogue already wrote
 * (dmc_gemm_args.stat_row_partials [Nt][parts] and stat_colsum_partials [row_groups][K]): no pass over the logits.
 * colsum_partials may be NULL (then colsum is not written: row statistics only, see dmc_rowdot for the column sums). */
int dmc_teacher_finalize(const float* row_partials, const float* colsum_partials, int64_t Nt, int64_t K, int64_t parts,
                         int64_t row_groups, float* row_stats, float* colsum, void* stream);
/* out[k] = sum_j W[k, j] * x[j] (W row-major [K, dim], F32 or BF16; x, out fp32).  With W = the weight-normed last layer's
 * operand g v/||v|| and x = the column sum of the teacher's normalised bottleneck rows, this IS torch.sum(teacher_output,
 * dim=0) of main_dino_mc.py:468 (sum_rows zhat_r . W_k = (sum_rows zhat_r) . W_k), from 2*K*dim bytes instead of a pass
 * over the [Nt, K] logits.  Honours dmc_set_streaming_ctas. */
int dmc_rowdot(const void* W, int32_t dtype, int64_t K, int64_t dim, const float* x, float* out, void* stream);

/* center_out = center_in * momentum + (colsum / count) * one_minus_momentum  (main_dino_mc.py:470-473)
 * with the reference's fp32 roundings: true division by count (= Nt * world_size), separate multiplies and
 * add (no FMA).  center_out may alias center_in.  `momentum` and `one_minus_momentum` are passed
 * separately because the reference forms (1 - momentum) in float64 before the cast to fp32. */
int dmc_center_update(const float* center_in, float* center_out, const float* colsum, int64_t K, float count,
                      float momentum, float one_minus_momentum, void* stream);

/* Crop-pair cross-entropy forward (main_dino_mc.py:441-459) in closed form: every student and every
 * teacher logit is read once.  Outputs s_lse[C*B] (log-sum-exp of s/tau_s per student row, saved for
 * backward) and loss[1]. */
size_t dmc_ce_workspace_bytes(int64_t B, int32_t C, int32_t G, int64_t K);
int dmc_ce_fwd(const void* s, int32_t s_dtype, int64_t lds, const void* t, int32_t t_dtype, int64_t ldt,
               const float* center, const float* t_row_stats, int64_t B, int32_t C, int32_t G, int64_t K,
               float inv_student_temp, float inv_teacher_temp, float* s_lse, float* loss,
               void* workspace, size_t workspace_bytes, void* stream);
/* Backward: ds[v*B+b,k] = grad_out * (n_v p_v - sum_{i != v} q_i) / (n B tau_s), written once as
 * ds_dtype.  grad_out is a DEVICE scalar (the upstream gradient of the 0-dim loss). */
int dmc_ce_bwd(const void* s, int32_t s_dtype, int64_t lds, const void* t, int32_t t_dtype, int64_t ldt,
               const float* center, const float* t_row_stats, const float* s_lse, const float* grad_out,
               int64_t B, int32_t C, int32_t G, int64_t K, float inv_student_temp, float inv_teacher_temp,
               void* ds, int32_t ds_dtype, int64_t ldds, void* stream);

/* lse[m] = logsumexp_n (D[m,n] * scale) from the GEMM epilogue's stat_row_partials [M][parts] (computed with
 * stat_scale = scale = 1/tau_s and no center). */
int dmc_lse_finalize(const float* row_partials, int64_t M, int64_t parts, float* lse, void* stream);
/* Fused forward + backward of the crop-pair cross-entropy for rows whose log-sum-exp is already known: ONE pass
 * over the logits writes loss[1] and ds = dL/ds for an upstream gradient of exactly 1 (dmc_scale_inplace_if
 * rescales when autograd later delivers something else).  Workspace as for dmc_ce_fwd. */
int dmc_ce_fused(const void* s, int32_t s_dtype, int64_t lds, const void* t, int32_t t_dtype, int64_t ldt,
                 const float* center, const float* t_row_stats, const float* s_lse, int64_t B, int32_t C, int32_t G,
                 int64_t K, float inv_student_temp, float inv_teacher_temp, void* ds, int32_t ds_dtype, int64_t ldds,
                 float* loss, void* workspace, size_t workspace_bytes, void* stream);
/* x *= (*scale / expected) unless the device scalar *scale == expected (then it costs one 4-byte read per CTA). */
int dmc_scale_inplace_if(void* x, int32_t dtype, int64_t n, const float* scale, float expected, void* stream);

/* ---------------------------------------------------------------------------------------------
 * EMA teacher update (main_dino_mc.py:403-406):  p_k = fp32(p_k * m) + fp32((1-m) * p_q) for every
 * (student, teacher) parameter pair, as ONE launch over a chunk table ("plan").
 * The plan is built once on the host from the parameter pointer lists, copied to the device by the
 * caller, and reused every step while the parameter storages do not move.
 * ------------------------------------------------------------------------------------------- */
size_t dmc_ema_plan_bytes(const int64_t* numels_host, int64_t n_tensors);
/* Fills plan_host (capacity from dmc_ema_plan_bytes) and *n_chunks_out.  Host-only; no CUDA calls. */
int dmc_ema_build_plan(const void* const* teacher_ptrs_host, const void* const* student_ptrs_host,
                       const int64_t* numels_host, int64_t n_tensors, void* plan_host, size_t plan_bytes,
                       int64_t* n_chunks_out);
int dmc_ema_multi_tensor(const void* plan_dev, int64_t n_chunks, float m, float one_minus_m, void* stream);

/* Plan v2: the same update (main_dino_mc.py:403-406, bit-exact), with
 *   - (m, 1-m) read from DEVICE memory (scalars_dev[0..1], fp32): a captured CUDA graph then follows the momentum
 *     schedule (main_dino_mc.py:404 `m = momentum_schedule[it]`) instead of freezing the captured value;
 *   - optional bf16 "shadow" copies of the NEW teacher values (shadow_bf16_ptrs_host[i], may be NULL per tensor or
 *     as a whole): the teacher head's MLP GEMM operands for the next step (utils/vision_transformer.py:291);
 *   - optionally the weight-normed last layer (utils/vision_transformer.py:279) handled row-wise: tensor wn_v_index is
 *     weight_v [rows, wn_dim], tensor wn_g_index is weight_g [rows]; besides both EMA updates the kernel writes
 *     wn_w_bf16[rows, wn_dim] = g v/||v|| (bf16), wn_scale[rows] = g/||v||, wn_inv_norm[rows] = 1/||v|| of the NEW
 *     values -- what dmc_weightnorm_fwd would compute next step.  wn_v_index = -1 disables it. */
size_t dmc_ema_plan2_bytes(const int64_t* numels_host, int64_t n_tensors, int64_t wn_v_index, int64_t wn_dim);
int dmc_ema_build_plan2(const void* const* teacher_ptrs_host, const void* const* student_ptrs_host,
                        const int64_t* numels_host, void* const* shadow_bf16_ptrs_host, int64_t n_tensors,
                        int64_t wn_v_index, int64_t wn_g_index, int64_t wn_dim, void* wn_w_bf16, float* wn_scale,
                        float* wn_inv_norm, void* plan_host, size_t plan_bytes, int64_t* n_chunks_out);
int dmc_ema_multi_tensor2(const void* plan_dev, int64_t n_chunks, const float* scalars_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Per-parameter gradient clipping: utils/utils.py:145-154 `clip_gradients(model, clip)`
 *   n = ||grad||_2 ; c = clip / (n + 1e-6) ; if c < 1: grad *= c        for every gradient tensor
 * as two multi-tensor launches over a host-built chunk plan (cf. dmc_ema_*).  norms[n_tensors] receives the
 * (pre-clip) norms on the device; nothing synchronises.  workspace: n_chunks floats.
 * --------------------------------------------------------------------------------------------- */
size_t dmc_clip_plan_bytes(const int64_t* numels_host, int64_t n_tensors);
/* Fills plan_host (capacity from dmc_clip_plan_bytes) and *n_chunks_out.  Host-only; no CUDA calls. */
int dmc_clip_build_plan(const void* const* grad_ptrs_host, const int64_t* numels_host, int64_t n_tensors,
                        void* plan_host, size_t plan_bytes, int64_t* n_chunks_out);
int dmc_clip_grads(const void* plan_dev, int64_t n_chunks, float clip, float* norms, float* workspace,
                   size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * torch.optim.AdamW step (main_dino_mc.py:281-282, :391, :399) for one parameter group as one multi-tensor launch:
 * p *= 1 - lr*wd; m.lerp_(g, 1-beta1); v = v*beta2 + (1-beta2) g g; p -= lr/(1-beta1^t) * m / (sqrt(v)/sqrt(1-beta2^t) + eps).
 * Scalars are doubles (the Python floats of torch) and are rounded to fp32 exactly where torch's kernels round them.
 * --------------------------------------------------------------------------------------------- */
size_t dmc_adamw_plan_bytes(const int64_t* numels_host, int64_t n_tensors);
/* Fills plan_host (capacity from dmc_adamw_plan_bytes) and *n_chunks_out.  Host-only; no CUDA calls. */
int dmc_adamw_build_plan(const void* const* param_ptrs_host, const void* const* grad_ptrs_host,
                         const void* const* exp_avg_ptrs_host, const void* const* exp_avg_sq_ptrs_host,
                         const int64_t* numels_host, int64_t n_tensors, void* plan_host, size_t plan_bytes,
                         int64_t* n_chunks_out);
int dmc_adamw_multi_tensor(const void* plan_dev, int64_t n_chunks, double lr, double beta1, double beta2, double eps,
                           double weight_decay, int64_t step, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The reference's LARS step (utils/utils.py:570-608; `--optimizer lars`, main_dino_mc.py:285-286) for one parameter group
 * as two multi-tensor launches.  Per parameter: adapt = (ndim != 1); d = adapt ? g + wd*p : g;
 * q = adapt && ||p|| > 0 && ||d|| > 0 ? eta*||p||/||d|| : 1; mu = mu*momentum + d*q; p -= lr*mu.
 * adapt_host[i] != 0 marks the tensors with ndim != 1.  workspace: 2 floats per chunk.
 * --------------------------------------------------------------------------------------------- */
size_t dmc_lars_plan_bytes(const int64_t* numels_host, int64_t n_tensors);
/* Fills plan_host (capacity from dmc_lars_plan_bytes) and *n_chunks_out.  Host-only; no CUDA calls. */
int dmc_lars_build_plan(const void* const* param_ptrs_host, const void* const* grad_ptrs_host,
                        const void* const* mu_ptrs_host, const int64_t* numels_host, const int32_t* adapt_host,
                        int64_t n_tensors, void* plan_host, size_t plan_bytes, int64_t* n_chunks_out);
int dmc_lars_multi_tensor(const void* plan_dev, int64_t n_chunks, double lr, double weight_decay, double momentum,
                          double eta, float* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Cross-rank exchange over NVLink / NVSwitch peer memory (one process per GPU, one node).
 * Replaces, for buffers in symmetric memory, the NCCL all-reduces behind DistributedDataParallel's gradient averaging
 * (main_dino_mc.py:260) and dist.all_reduce(batch_center) (main_dino_mc.py:469).
 *
 * Every rank holds the buffer at peer_ptrs_host[rank]; peer_ptrs_host[p] is rank p's buffer mapped into this process
 * and multicast_ptr (may be NULL) the NVSwitch multicast address bound to all of them (torch.distributed
 * ._symmetric_memory provides all three, plus one zero-initialised signal pad per rank of at least
 * dmc_xrank_signal_bytes(world, ctas) bytes).  In place, on every rank: buffer <- scale * sum_ranks buffer (two-shot:
 * rank r reduces slice r -- one multimem.ld_reduce per 16 bytes with a multicast address, else world-1 peer loads --
 * and broadcasts it -- one multimem.st, else world-1 peer stores), framed by two barriers over all CTAs of all ranks (rank-local
 * counter, then one signal per rank pair on the signal pads).
 * BF16 buffers are summed in fp32 and rounded once.  Optional epilogue (n_out > 0, BF16 only): widen element ranges of
 * the finished buffer into fp32 tensors (out_ptrs_host[i][0..n) <- buffer[off..off+n), off % 8 == 0).
 * All ranks must call it with the same numel / dtype / ctas, in the same order per signal pad.  ctas x 256 threads,
 * 32 registers, 4 bytes of shared memory: with ctas = 148 one CTA per SM, co-resident with a tcgen05 GEMM CTA.  Every CTA of a
 * launch must be resident at the same time (the barrier is grid-wide per rank): keep ctas <= 148. */
size_t dmc_xrank_signal_bytes(int32_t world, int32_t ctas);
int dmc_xrank_allreduce(void* multicast_ptr, void* const* peer_ptrs_host, void* const* signal_pads_host, int64_t numel,
                        int32_t dtype, int32_t rank, int32_t world, float scale, int32_t ctas, int32_t n_out,
                        float* const* out_ptrs_host, const int64_t* out_offsets_host, const int64_t* out_numels_host,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DINOMC_H_ */
