/*
 * precond_b200 -- C ABI of the B200-native Shampoo preconditioner hot path.
 *
 * Drop-in boundary for the device work of precondition.distributed_shampoo
 * (reference: /root/reference/precondition/distributed_shampoo.py, "DS" below;
 * quantization_utils.py = "QU").  The reference has no FFI of its own -- every
 * entry point below replaces the jnp/XLA computation at the cited seam and is
 * what a jax.ffi custom call (or ctypes, as precondition_b200/_lib.py does)
 * binds.  See INTEGRATION.md for the reference-side stubs.
 *
 * Conventions
 *   - plain C, raw DEVICE pointers, sizes as int/int64_t, a cudaStream_t (passed
 *     as void*) on which all work is enqueued;
 *   - return value: 0 = ok, <0 = invalid argument / CUDA error detected on the
 *     host (pc_last_error() gives the message).  Numerical failure is NOT an
 *     error: it is reported through the metrics rows exactly like the
 *     reference's TrainingMetrics (DS:902-907) and acted on by the caller
 *     (DS:2936-2950);
 *   - the library owns no user memory: callers pass a workspace whose size comes
 *     from the matching *_workspace_bytes query;
 *   - matrices are row-major, contiguous, 16-byte aligned;
 *   - every entry point enqueues on `stream` and returns; pc_inverse_pth_root_* do so through a
 *     cached CUDA graph with a device-driven WHILE loop (PC_ROOT_MODE=poll restores the
 *     host-polled loop, which waits on events while the device iterates).
 */
#ifndef PRECOND_B200_H_
#define PRECOND_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PC_OK 0
#define PC_ERR_INVALID (-1)
#define PC_ERR_CUDA (-2)
#define PC_ERR_WORKSPACE (-3)
#define PC_ERR_UNSUPPORTED (-4)

/* metrics row layout == TrainingMetrics scalar fields, DS:338-351 */
#define PC_METRIC_ERROR 0       /* inverse_pth_root_errors */
#define PC_METRIC_ITERS 1       /* inverse_pth_root_iters  */
#define PC_METRIC_ERROR_RATIO 2 /* final_error_ratio       */
#define PC_METRIC_MAX_EV 3      /* max_eigen_value         */
#define PC_METRIC_RETRIES 4     /* total_retries           */
#define PC_NUM_METRICS 5

/* GEMM engine used by the Newton chain */
#define PC_ENGINE_AUTO 0
#define PC_ENGINE_SIMT_FP32 1 /* CUDA-core fp32 FFMA tiles (any n)              */
#define PC_ENGINE_TC_BF16X6 2 /* tcgen05, bf16 3-way split, 6 products (~fp32)  */
#define PC_ENGINE_TC_BF16X3 3 /* tcgen05, bf16 2-way split, 3 products (~2^-16) */
#define PC_ENGINE_TC_FP16X3 4 /* tcgen05, fp16 + 2^11-scaled fp16 residual, 3 products
                                 (22-bit operands, like 3xTF32 at the f16 MMA rate)  */
#define PC_ENGINE_TC_SMALL 5  /* tcgen05, n <= 128: one persistent CTA per matrix runs the whole
                                 solve (power iteration, Newton loop, retries) with the iterates
                                 in shared / tensor memory; exact bf16x6 products; exponents 2^s
                                 given on the host (pc_inverse_pth_root_enqueue).  PC_ENGINE_AUTO
                                 picks it whenever it applies.                          */

/* quantised storage of statistics / preconditioners, QU:49-113 */
#define PC_QDTYPE_F32 0
#define PC_QDTYPE_INT16 1
#define PC_QDTYPE_INT8 2
#define PC_QDTYPE_BF16 3

/* GraftingType, DS:499-506 */
#define PC_GRAFT_NONE 0
#define PC_GRAFT_SGD 1
#define PC_GRAFT_ADAGRAD 2
#define PC_GRAFT_RMSPROP 3
#define PC_GRAFT_RMSPROP_NORMALIZED 4
#define PC_GRAFT_SQRT_N 5
#define PC_GRAFT_ADAGRAD_NORMALIZED 6

int pc_version(void);
const char* pc_last_error(void);
/* 1 if the running device is sm_100 (tcgen05 engines usable). */
int pc_device_supports_tcgen05(void);

/* ------------------------------------------------------------------------
 * instrumentation (bench.py): kernel-launch counter and, when enabled, CUDA-event
 * timing of the Newton-chain GEMM launches on the caller's stream.
 * ------------------------------------------------------------------------ */
typedef struct {
  int64_t kernel_launches;   /* every kernel this library launched since reset  */
  int64_t gemm_launches;     /* Newton-chain GEMM phase launches                 */
  double gemm_ms;            /* sum of their durations (only if timing enabled)  */
  double gemm_flops;         /* algorithmic flops of the executed GEMM tiles:
                                2 n^3 per (matrix, step), summed on the host from
                                the final iteration counts                       */
} pc_stats;
void pc_stats_reset(int enable_gemm_timing);
void pc_stats_get(pc_stats* out);

/* ------------------------------------------------------------------------
 * (2) batched matrix_inverse_pth_root
 * replaces: _matrix_inverse_pth_root_vmap (DS:2742-2744) =
 *           jax.vmap(matrix_inverse_pth_root) (DS:702-940), incl. power_iteration
 *           (DS:595-652), ridge damping (DS:830), coupled Newton (DS:836-885),
 *           retry loop (DS:858-885), padding mask (DS:777-783, DS:930-937).
 *   xs              [batch, n, n] f32   statistics (symmetric PSD)
 *   ps              [batch] i32         exponents p in [1, 16] (larger p: zero root, NaN error)
 *   padding_starts  [batch] i32         rows/cols >= padding_start are padding
 *                                       (pass n for "no padding"; may be NULL)
 *   roots           [batch, n, n] f32   out: (A + eps I)^(-1/p), zeros in padding
 *   metrics         [batch, 5] f32      out: PC_METRIC_* rows
 * ------------------------------------------------------------------------ */
typedef struct {
  float ridge_epsilon;         /* matrix_epsilon, DS:1855 (default 1e-6)       */
  float error_tolerance;       /* DS:707 (default 1e-6)                        */
  int num_iters;               /* DS:705 (default 100)                         */
  int relative_matrix_epsilon; /* DS:1884 (default 1)                          */
  int engine;                  /* PC_ENGINE_*                                  */
  int reserved;
} pc_root_options;

void pc_root_options_default(pc_root_options* opt);
/* engine PC_ENGINE_AUTO resolves to for statistics of size n on the current device */
int pc_resolve_engine(int n, int engine);

size_t pc_inverse_pth_root_workspace_bytes(int batch, int n, int engine);

int pc_inverse_pth_root_batched(const float* xs, const int32_t* ps,
                                const int32_t* padding_starts, int batch, int n,
                                const pc_root_options* opt, float* roots,
                                float* metrics, void* workspace,
                                size_t workspace_bytes, void* stream);

/* The same solver with the exponents also given as a HOST array (ps_host, may be NULL): the
 * host then knows how many GEMM launches one Newton iteration needs without reading `ps`
 * back, and exponents outside [1, 16] are rejected up front.  Both entry points ENQUEUE AND
 * RETURN in the default "graph" mode: the call is one CUDA graph whose Newton loop is a
 * conditional WHILE node driven by a device-side convergence check (no host polling; the
 * executable graph is cached per argument set, so a training loop only re-launches it).
 * PC_ROOT_MODE=poll selects the host-polled loop of round 1 (it waits on events while the
 * device iterates); GEMM-timing runs (pc_stats_reset(1)) use it too. */
int pc_inverse_pth_root_enqueue(const float* xs, const int32_t* ps, const int32_t* ps_host,
                                const int32_t* padding_starts, int batch, int n,
                                const pc_root_options* opt, float* roots, float* metrics,
                                void* workspace, size_t workspace_bytes, void* stream);
/* 1 if pc_inverse_pth_root_* currently run in graph mode (enqueue only), 0 if host-polled. */
int pc_root_mode(void);

/* Test hook for the tcgen05 engine: C[b] = A[b] * B[b]^T (fp32 in/out, computed as
 * split-bf16 products, `passes` = 6 or 3, or scaled-fp16 products, `passes` = -3), n % 128 == 0.  Workspace of at least
 * pc_inverse_pth_root_workspace_bytes(batch, n, PC_ENGINE_TC_BF16X6) bytes. */
int pc_debug_tc_gemm(const float* a, const float* b, float* c, int batch, int n, int passes,
                     void* workspace, size_t workspace_bytes, void* stream);

/* power_iteration alone (DS:595-652): lambda[b] = Rayleigh quotient of the last
 * executed step, iters[b] = steps taken (may be NULL). */
int pc_power_iteration_batched(const float* xs, const int32_t* padding_starts,
                               int batch, int n, int num_iters,
                               float error_tolerance, float* lambdas,
                               int32_t* iters, void* stream);

/* ------------------------------------------------------------------------
 * grouped GEMM with fused epilogue -- the building block behind (1) and (4):
 *   C = alpha * op(A) op(B) + beta * C_in          (fp32, CUDA cores or tcgen05)
 * replaces: jnp.tensordot in gram_weighted_update (DS:1468-1470) and in
 *           _precondition_block (DS:1707).  One descriptor per block; operands
 *           are addressed as  X(i, k) = base[i*s_i + (k / k_inner)*s_ko +
 *           (k % k_inner)*s_ki]  so that every mode-k unfolding of a rank<=3
 *           gradient block is a view (no jnp.split copies, DS:1412-1422).
 * ------------------------------------------------------------------------ */
/*   A(i,k) = a[(i / a_iinner)*a_sio + (i % a_iinner)*a_si + (k / a_kinner)*a_sko + (k % a_kinner)*a_ski] */
typedef struct {
  const float* a;    /* A(i,k), i in [0,M), k in [0,K) */
  const float* b;    /* B(j,k), j in [0,N), k in [0,K)  (i.e. op(B)^T) */
  const float* c_in; /* optional, addressed like c */
  float* c;          /* C(i,j) = c[(i / c_iinner)*c_sio + (i % c_iinner)*c_sii + j] */
  int64_t a_sio, a_si, a_sko, a_ski;
  int64_t b_sj, b_sko, b_ski;
  int64_t c_sio, c_sii;
  int32_t a_iinner, a_kinner, b_kinner, c_iinner;
  int32_t m, n, k;
  float alpha, beta;
  int32_t reserved;
  const float* beta_dev; /* optional DEVICE scalar: if not NULL it replaces `beta` (a weight that
                          * only exists on the device, e.g. the `const` of a packed sketch) */
} pc_gemm_desc;

/* descs: DEVICE array of `count` descriptors; max_m/max_n bound the tile grid. */
int pc_grouped_gemm(const pc_gemm_desc* descs, int count, int max_m, int max_n,
                    void* stream);

/* Split-K form of pc_grouped_gemm for descriptors with a small output and a very long
 * contraction (e.g. the 9 x 9 statistic of a 3 x 3 x 512 x 512 kernel: k = 262144 in a single
 * tile): `splits` CTAs per tile, partial products summed in a fixed order (deterministic). */
size_t pc_grouped_gemm_splitk_workspace_bytes(int count, int max_m, int max_n, int splits);
int pc_grouped_gemm_splitk(const pc_gemm_desc* descs, int count, int max_m, int max_n, int splits,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Thin products on CUDA cores, HBM-bound streaming kernels for the shapes a 64 x 64 tile wastes
 * (same descriptors, DEVICE array, as pc_grouped_gemm; any sizes are computed correctly, the
 * kinds say what each kernel is built for):
 *   PC_THIN_GEMV    m <= 4: a rank-1 parameter times its preconditioner (DS:1707 on a [1, n]
 *                   block) -- one CTA per 32 columns, the matrix is read once;
 *   PC_THIN_ROWMAP  n <= 16 and k <= 16 over very many rows: the mode product of a [9, c, c]
 *                   convolution kernel with its 9 x 9 preconditioner -- one thread per row.
 *   PC_THIN_OUTER   k <= 4: the statistic of a rank-1
 *                   parameter, S <- w1 S + w2 g g^T (DS:1468-1470 on a [n] block) -- a streaming
 *                   pass over the output, bitwise symmetric for A == B and a symmetric C_in.
 * pc_grouped_gemm_splitk picks its thin form by itself when max_m, max_n <= 16 (the 9 x 9
 * statistic of the same kernel, DS:1468-1470). */
#define PC_THIN_GEMV 0
#define PC_THIN_ROWMAP 1
#define PC_THIN_OUTER 2
int pc_grouped_gemm_thin(const pc_gemm_desc* descs, int count, int max_m, int max_n, int kind,
                         void* stream);

/* tcgen05 path of the grouped GEMM (any m and k, n a multiple of 4; c / c_in 16-byte aligned with
 * c_sii, c_sio multiples of 4; edge tiles of sizes that are no multiples of 128 -- the 1000 x 1000
 * statistic of a classifier, 576-row convolution kernels -- are zero-filled when the operand is
 * packed and masked when the result is stored): every operand view is
 * packed once into scaled-fp16 plane tiles (22 mantissa bits, per-operand power-of-two scale)
 * and the products run as three kind::f16 MMAs per k-step with fp32 accumulation.  A descriptor
 * with identical A and B views (the Gram update of DS:1468-1470) is computed as a symmetric
 * rank-k update: lower tiles only, mirrored, so the result is bitwise symmetric.
 * descs_host: HOST array (the host plans the tile list and uploads it into `workspace`, which
 * costs one synchronisation of `stream`).  reuse_plan bit 0: the caller guarantees that
 * `workspace` still holds the plan of the previous call with identical descriptors (a
 * training loop's static launch list) -- nothing is uploaded and the call only enqueues.
 * reuse_plan == 3: additionally the B views are unchanged since that call, their packed planes
 * are reused (a fixed matrix applied to changing blocks). */
size_t pc_grouped_gemm_tc_workspace_bytes(const pc_gemm_desc* descs_host, int count);
int pc_grouped_gemm_tc(const pc_gemm_desc* descs_host, int count, void* workspace,
                       size_t workspace_bytes, int reuse_plan, void* stream);

/* The same call with QuantizedValue (QU:49-113) fused in, for the statistics update of
 * quantised second moments (DS:1588-1590 around gram_weighted_update): per descriptor (a
 * symmetric product into a contiguous [n, n] matrix, n a multiple of 128)
 *   q_in != NULL      C_in is read as to_float(q_in, diag_in, bucket_in) = q * bucket[col] + diag
 *                     on the diagonal -- the dequantised matrix is never materialised;
 *   colmax_out != NULL  [n] uint32, zero on entry: receives the bit patterns of the per-column
 *                     max |off-diagonal| of the RESULT (the reduction QU:86 needs), so that
 *                     pc_quantize_from_colmax_batched requantises in a single pass.
 * The fp32 result is still written to desc.c (scratch for the requantisation). */
typedef struct {
  const void* q_in;
  const float* diag_in;
  const float* bucket_in;
  uint32_t* colmax_out;
  int32_t qdtype; /* PC_QDTYPE_INT16 or PC_QDTYPE_INT8 */
  int32_t reserved;
  /* b_q != NULL: the B operand is a QuantizedValue with extracted diagonal -- a square
   * [b_ld, b_ld] matrix q (b_qdtype), diag [b_ld], bucket [b_ld] -- read through desc.b's
   * strides with desc.b ignored: to_float(q)[r][c] = q[r][c] * bucket[c] + diag[r] (r == c) is
   * formed while the operand is packed, the dequantised preconditioner is never materialised
   * (DS:3556 _maybe_dequantize_preconditioners + DS:1707).  One-level B addressing only. */
  const void* b_q;
  const float* b_diag;
  const float* b_bucket;
  int32_t b_ld;
  int32_t b_qdtype;
} pc_gemm_quant;
int pc_grouped_gemm_tc_quant(const pc_gemm_desc* descs_host, const pc_gemm_quant* quant_host,
                             int count, void* workspace, size_t workspace_bytes, int reuse_plan,
                             void* stream);
/* from_float with extract_diagonal (QU:49-95) given the column maxima: x [batch, n, n] f32,
 * colmax [batch, n] -> q, diag [batch, n], bucket [batch, n]; one pass over x. */
int pc_quantize_from_colmax_batched(const float* x, const uint32_t* colmax, int batch, int n,
                                    int qdtype, void* q, float* diag, float* bucket,
                                    void* stream);

/* Failure fallback of DS:2936-2950 without a host round trip: for every matrix b,
 * dst[b] <- src[b] unless metrics[b][PC_METRIC_ERROR] is NaN or >= threshold
 * (then dst keeps the previous preconditioner).  src rows are [src_rows, src_cols]
 * (the padded root); the top-left [rows, cols] corner is copied (DS:2950). */
int pc_select_preconditioners(const float* src, const float* metrics, float threshold,
                              float* dst, int batch, int src_rows, int src_cols, int rows,
                              int cols, void* stream);

/* Scatter form for block-sharded roots after the all-gather (DS:2876-2879 + DS:2936-2950):
 * row j of a gathered buffer goes to row dst_index[j] of the state (dst_index[j] < 0: filler
 * row of the padded batch, DS:2844-2850, skipped) unless its metrics row -- at
 * metrics_base + metrics_offset[j] -- reports NaN or an error >= threshold.  Rows are raw
 * bytes at src + src_offset_bytes[j], so fp32 roots and the int16 / int8, diagonal and bucket
 * parts of quantised preconditioners (DS:3116-3122) use the same call; metrics_dst (optional)
 * receives the metrics rows in state order.  All index arrays are DEVICE arrays of `count`. */
int pc_select_scatter(const void* src, const int64_t* src_offset_bytes,
                      const float* metrics_base, const int64_t* metrics_offset,
                      const int32_t* dst_index, float threshold, void* dst, int64_t row_bytes,
                      float* metrics_dst, int count, void* stream);

/* ------------------------------------------------------------------------
 * (3) Sketchy / frequent-directions sketch update
 * replaces: _fd_update_root (DS:1123-1290) as vmapped by new_mi_pth_root
 *           (DS:2706-2738), incl. _fd_low_rank_unpack / _fd_low_rank_pack (DS:555-592).
 *   new_grad  input_is_gram = 0: [batch, d, m] f32, any factor F with F F^T = x x^T -- the
 *             reference's zero-padded QR factor from frequent_directions_update
 *             (DS:1473-1505) has m = d; the gradient block unfolding itself works too.
 *             input_is_gram = 1: [batch, d, d] f32, x x^T itself (m ignored).
 *   prev      [batch, d, rank+2] f32  previous packed sketch (DS:563-568)
 *   ps        [batch] i32             exponents
 *   padding_starts [batch] i32        may be NULL (= d)
 *   out       [batch, d, rank+2] f32  new packed sketch
 *   metrics   [batch, 5] f32          may be NULL; error 0 like DS:1263-1264
 * The left singular vectors / values of [sqrt(beta2) U sqrt(lambda+eps) | G] (DS:1180-1193)
 * are computed as eigenpairs of the d x d covariance: exactly (cyclic Jacobi on all of it)
 * for d <= full_eigh_max_dim, else by block subspace iteration warm-started from the
 * previous sketch with a (rank+1+oversample)^2 Rayleigh-Ritz eigenproblem.
 * ------------------------------------------------------------------------ */
typedef struct {
  float ridge_epsilon;         /* matrix_epsilon                                  */
  float error_tolerance;       /* floor of the relative damping, DS:1159 (1e-6)   */
  int relative_matrix_epsilon; /* DS:1155-1158                                    */
  float decay;                 /* beta2, DS:1180, DS:1201                         */
  int input_is_gram;
  int subspace_iters;          /* block iterations of the large-d path (6)        */
  int oversample;              /* extra basis vectors of the large-d path (32)    */
  int full_eigh_max_dim;       /* <= 512: d up to here is solved exactly (512)    */
  /* tearfree's Sketchy (TF/sketchy.py:380-470) on the same eigen-solver: the eigenvalue slots
   * of prev / out hold SINGULAR values, the tail decays by sqrt(decay), no ridge enters the
   * sketch (ridge_epsilon is ignored), the inverted values are (undeflated + eps)^(-1/p) with
   * eps = tearfree_epsilon (* the largest undeflated eigenvalue if relative), the const slot
   * is (tail + eps)^(-1/p), and the has_zeros skip flag is never set. */
  int tearfree;
  float tearfree_epsilon;
  int tearfree_relative_epsilon;
} pc_fd_options;

void pc_fd_options_default(pc_fd_options* opt);
size_t pc_fd_update_workspace_bytes(int batch, int d, int m, int rank, const pc_fd_options* opt);
int pc_fd_update_batched(const float* new_grad, const float* prev, const int32_t* ps,
                         const int32_t* padding_starts, int batch, int d, int m, int rank,
                         const pc_fd_options* opt, float* out, float* metrics, void* workspace,
                         size_t workspace_bytes, void* stream);

/* eigh-based low-rank root for compression_rank != 0 without frequent_directions:
 * replaces _low_rank_root (DS:1033-1120) as vmapped by new_mi_pth_root (DS:2706-2738).
 *   xs [batch, d, d] f32 (lower triangle authoritative), ps [batch] i32, padding_starts [batch]
 *   i32 or NULL, compression_rank: > 0 keeps the largest eigenvalues, < 0 the smallest
 *   out [batch, d, |rank|+2] packed (eigvecs, inverted eigenvalues, const; DS:548-552),
 *   metrics [batch, 5]: error = max |U^T reg U - diag(e)| (DS:1076-1081).  d <= 2048 (the
 *   Jacobi solve runs in one thread-block cluster per matrix: fast up to 512, slow above). */
size_t pc_low_rank_root_workspace_bytes(int batch, int d);
int pc_low_rank_root_batched(const float* xs, const int32_t* ps, const int32_t* padding_starts,
                             int batch, int d, int compression_rank, float ridge_epsilon,
                             float error_tolerance, int relative_matrix_epsilon, float* out,
                             float* metrics, void* workspace, size_t workspace_bytes,
                             void* stream);

/* eigh-based full root, `eigh=True`: replaces matrix_inverse_pth_root_eigh (DS:943-1030).
 * Same inputs as pc_inverse_pth_root_batched, workspace from pc_low_rank_root_workspace_bytes;
 * roots [batch, d, d] = U diag(max(e, ridge)^(-1/p)) U^T, metrics error = max|U^T reg U - diag(e)|.
 * d <= 2048. */
int pc_inverse_pth_root_eigh_batched(const float* xs, const int32_t* ps,
                                     const int32_t* padding_starts, int batch, int d,
                                     float ridge_epsilon, float error_tolerance,
                                     int relative_matrix_epsilon, float* roots, float* metrics,
                                     void* workspace, size_t workspace_bytes, void* stream);

/* Pseudo-inverse p-th root by eigendecomposition, as tearfree's blocked Shampoo computes its
 * preconditioners: replaces _pth_inv_root (TF/shampoo.py:440-448).
 *   roots[b] = V diag(w_i^(-1/p), or 0 where w_i <= rel_cutoff * max(w)) V^T,  (w, V) = eigh(xs[b])
 * No ridge is added (rank-deficient statistics are expected); rel_cutoff = 1e-6 in the reference.
 * xs [batch, d, d] f32 symmetric, ps [batch] i32 DEVICE, workspace from
 * pc_low_rank_root_workspace_bytes(batch, d).  d <= 2048. */
int pc_pinv_pth_root_eigh_batched(const float* xs, const int32_t* ps, int batch, int d,
                                  float rel_cutoff, float* roots, void* workspace,
                                  size_t workspace_bytes, void* stream);
/* The same with a warm start: eigvecs [batch, d, d] (rows = eigenvectors, descending) receives
 * this call's eigenvectors and, when eigvecs_valid != 0, holds those of the previous call for
 * the same statistics -- the eigen-solve then runs in that basis and needs a fraction of the
 * Jacobi sweeps while the statistics drift slowly (every step of a training run). */
int pc_pinv_pth_root_eigh_warm_batched(const float* xs, const int32_t* ps, int batch, int d,
                                       float rel_cutoff, float* roots, float* eigvecs,
                                       int eigvecs_valid, void* workspace,
                                       size_t workspace_bytes, void* stream);

/* Dense form of the operator a packed low-rank preconditioner applies in
 * _precondition_block (DS:1690-1705, _low_rank_unpack DS:540-545):
 *   dense[b] = c I + V diag(lambda^- - c) V^T   (identity if the has_zeros flag is set),
 * so that g -> c (g - g V V^T) + (g V lambda^-) V^T is one product g * dense[b] and the
 * low-rank blocks go through the same grouped GEMM as the full ones.
 *   packed [batch, d, rank+2] f32 -> dense [batch, d, d] f32 */
size_t pc_low_rank_to_dense_workspace_bytes(int batch, int d, int rank);
/* The same operator in factored form, for applying it as  g -> c g + (g V) W^T  (two thin
 * products, DS:1690-1705 as written) instead of through its d x d matrix:
 *   w [batch, d, rank] = V diag(lambda^- - c)  (zero if has_zeros),  c [batch] (1 if has_zeros). */
int pc_low_rank_factors(const float* packed, int batch, int d, int rank, float* w, float* c,
                        void* stream);
int pc_low_rank_to_dense(const float* packed, int batch, int d, int rank, float* dense,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * (1b) QuantizedValue (QU:49-113) for square statistics / preconditioners with
 *      extract_diagonal=True (DS:2087-2095) and for momenta (DS:2111-2114).
 *   quantize:   x [rows, cols] f32 -> q (int16/int8/bf16), diag [rows] (if
 *               extract_diagonal), bucket [cols]; per-COLUMN max-abs (QU:86).
 *   dequantize: inverse (QU:97-113).
 *   `batch` independent matrices, contiguous.
 * ------------------------------------------------------------------------ */
int pc_quantize_batched(const float* x, int batch, int rows, int cols, int qdtype,
                        int extract_diagonal, void* q, float* diag, float* bucket,
                        void* stream);
int pc_dequantize_batched(const void* q, const float* diag, const float* bucket,
                          int batch, int rows, int cols, int qdtype,
                          int extract_diagonal, float* x, void* stream);

/* Grouped form for the int8 momenta of a whole model (the reference maps to_float /
 * from_float over the parameter tree around _transform_grad, DS:3582-3586, DS:3620-3621): one
 * segment per momentum, work cut into chunks of pc_quant_group_chunk_elems() elements.
 *   segments       DEVICE [num_segments]: int8 data q [rows, cols], bucket [cols], the fp32 view
 *                  x (e.g. a slice of a flat buffer), colmax [cols] uint32 scratch of the
 *                  quantiser; first_chunk = running sum of nchunks = ceil(rows*cols / chunk)
 *                  the column maxima are reduced over tiles of pc_quant_group_tile_rows() rows x
 *                  128 columns: col_tiles = ceil(cols / 128), first_tile = running sum of
 *                  ceil(rows / tile_rows) * col_tiles
 *   chunk_segment  DEVICE [total_chunks] i32;  tile_segment  DEVICE [total_tiles] i32
 *   pc_dequantize_grouped  x = q * bucket[col]                                   (QU:97-113)
 *   pc_quantize_grouped    bucket = max_rows |x| / 127, q = round(x / bucket)    (QU:49-95);
 *                          colmax_all / colmax_bytes: the scratch all segments' colmax point into
 *                          (zeroed here). */
typedef struct {
  void* q;
  float* bucket;
  float* x;
  uint32_t* colmax;
  int32_t rows, cols;
  int32_t first_chunk, nchunks;
  int32_t first_tile, col_tiles;
} pc_quant_segment;
int64_t pc_quant_group_chunk_elems(void);
int pc_quant_group_tile_rows(void);
int pc_dequantize_grouped(const pc_quant_segment* segments, const int32_t* chunk_segment,
                          int num_segments, int64_t total_chunks, void* stream);
int pc_quantize_grouped(const pc_quant_segment* segments, const int32_t* chunk_segment,
                        int num_segments, int64_t total_chunks, const int32_t* tile_segment,
                        int64_t total_tiles, uint32_t* colmax_all, size_t colmax_bytes,
                        void* stream);

/* ------------------------------------------------------------------------
 * (4) grafting + momentum tail of _transform_grad (DS:3496-3625) for one
 *     parameter tensor of `numel` elements.  precond_grad is the output of the
 *     preconditioner application (DS:3553-3556) or NULL when the parameter is
 *     skipped (DS:3557-3561).  Momenta / diagonal statistics are f32 here; int8
 *     momenta go through pc_(de)quantize_batched.
 * ------------------------------------------------------------------------ */
typedef struct {
  double beta1, beta2; /* as Python floats: 1 - beta is formed in double, DS:3522, DS:3579 */
  int graft_type;
  float diagonal_epsilon;
  float weight_decay;
  float learning_rate; /* already evaluated at this step, DS:3545-3547 */
  int nesterov;
  int moving_average_for_momentum;
  int decoupled_learning_rate;
  int decoupled_weight_decay;
  int run_shampoo;                  /* step >= start_preconditioning_step, DS:3588 */
  float clip_by_scaled_gradient_norm; /* <=0: disabled, DS:3530 */
} pc_graft_options;

size_t pc_graft_momentum_workspace_bytes(int64_t numel);

int pc_graft_momentum(const float* grad, const float* param,
                      const float* precond_grad, float* diagonal_statistics,
                      float* diagonal_momentum, float* momentum, float* update,
                      int64_t numel, const pc_graft_options* opt, void* workspace,
                      size_t workspace_bytes, void* stream);

/* Grouped form of (4): the tail of _transform_grad for EVERY parameter of the model in a fixed
 * number of launches (the reference maps _transform_grad over the parameter tree,
 * DS:3650-3657).  All per-parameter arrays live in flat buffers with one segment per parameter
 * at the same element offset in each of them (offsets multiples of 32 elements, so every
 * segment is 16-byte aligned); precond_grad may be NULL, and a segment with has_precond == 0
 * behaves like precond_grad == NULL in pc_graft_momentum (skipped parameter, DS:3557-3561).
 * Work is cut into chunks of pc_graft_group_chunk_elems() elements:
 *   segments       DEVICE [num_segments]; first_chunk = running sum of nchunks,
 *                  nchunks = ceil(numel / chunk)
 *   chunk_segment  DEVICE [total_chunks] i32: the segment each chunk belongs to
 * Norms are reduced per segment with fixed-order two-level sums (deterministic). */
typedef struct {
  int64_t offset;      /* first element of the segment in every flat buffer */
  int64_t numel;
  int32_t first_chunk;
  int32_t nchunks;
  int32_t has_precond;
  int32_t reserved;
} pc_graft_segment;

int64_t pc_graft_group_chunk_elems(void);
size_t pc_graft_momentum_grouped_workspace_bytes(int num_segments, int64_t total_chunks);
int pc_graft_momentum_grouped(const pc_graft_segment* segments, const int32_t* chunk_segment,
                              int num_segments, int64_t total_chunks, const float* grad,
                              const float* param, const float* precond_grad,
                              float* diagonal_statistics, float* diagonal_momentum,
                              float* momentum, float* update, const pc_graft_options* opt,
                              void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * Top-k deflation around the inverse p-th root: the `lobpcg_topk_precondition` branch of
 * matrix_inverse_pth_root (DS:789-812, DS:889-928) and the diagnostics of DS:109-195.
 * The top-k eigenpairs come from pc_fd_update_batched (the matrix passed as a Gram, an empty
 * previous sketch, decay 1): `packed` below is that call's output [batch, n, k + 2].
 *   pc_lobpcg_deflate_prep    eigvals [batch, k] out; scalars [batch, 4] = {max eigenvalue, min
 *                             eigenvalue, absolute ridge = ridge_epsilon * max(max_ev, 1e-25),
 *                             1 / max_ev}; s1 [batch, n, k] = V sqrt((l - l_min) / max_ev);
 *                             a_scaled = A / max_ev.  The caller forms A' = a_scaled - s1 s1^T
 *                             (grouped GEMM) and solves it with an ABSOLUTE epsilon ridge_epsilon.
 *   pc_lobpcg_redeflate_prep  roots <- roots * max_ev^(-1/p); s2 = V sqrt(_pth_root_difference)
 *                             (DS:681-699); the caller forms root - s2 s2^T (DS:896-900).
 *   pc_root_diagnostics       InversePthRootDiagnostics (DS:109-142) of mat_m = B^p A:
 *                             out [batch, 4] = {max, mean |diag - 1|, max, mean |offdiag|}.
 *   pc_lobpcg_diagnostics     LOBPCGDiagnostics (DS:149-195) from av = A V [batch, n, k] and
 *                             gram = V^T V [batch, k, k]: out [batch, 7] in field order.
 * ------------------------------------------------------------------------ */
int pc_lobpcg_deflate_prep(const float* packed, const float* a, int batch, int n, int k,
                           float ridge_epsilon, int relative_matrix_epsilon, float* scalars,
                           float* eigvals, float* s1, float* a_scaled, void* stream);
int pc_lobpcg_redeflate_prep(const float* packed, const float* scalars, const float* eigvals,
                             const int32_t* ps, int batch, int n, int k, float* roots, float* s2,
                             void* stream);
int pc_root_diagnostics(const float* mat_m, int batch, int n, float* out, void* stream);
int pc_lobpcg_diagnostics(const float* packed, const float* av, const float* gram,
                          const float* eigvals, int batch, int n, int k, float iters, float* out,
                          void* stream);

/* ------------------------------------------------------------------------
 * SM3 (precondition/sm3.py:40-168): the diagonal-accumulator optimizer that shares the int8
 * momentum storage with distributed_shampoo.  One parameter tensor of rank 1..4 per call:
 *   nu = beta2 * min_axis(acc_in[axis][index_axis]) + w2 * g^2        (rank 1: acc_in[0])
 *   update = -lr * (m' + weight_decay * param),  m' = beta1 * to_float(m) + w1 * g / sqrt(nu + eps)
 *   acc_out[axis][i] = max of nu over the other axes                  (rank 1: nu)
 * w = 1 - beta (or 1 if beta == 1); normalize_grads: g <- g / (|g| + 1e-16).
 * momentum_q / momentum_bucket: the int8 QuantizedValue of the momentum (bucket over
 * shape[1:]; both may be NULL = zero momentum); momentum_f receives m' in fp32 -- requantise it
 * with pc_quantize_batched(rows = dims[0], cols = numel / dims[0], PC_QDTYPE_INT8).
 * acc_out must not alias acc_in.
 * ------------------------------------------------------------------------ */
typedef struct {
  double beta1, beta2;
  float diagonal_epsilon;
  float weight_decay;
  float learning_rate;
  int normalize_grads;
} pc_sm3_options;
size_t pc_sm3_workspace_bytes(int64_t numel);
int pc_sm3_update(const float* grad, const float* param, const float* const* acc_in,
                  float* const* acc_out, const int8_t* momentum_q, const float* momentum_bucket,
                  float* momentum_f, float* update, int rank, const int32_t* dims,
                  const pc_sm3_options* opt, void* workspace, size_t workspace_bytes,
                  void* stream);

/* ------------------------------------------------------------------------
 * tearfree (precondition/tearfree): everything after the second-order direction, for every
 * parameter of the model in three launches.  Per parameter (segment), in the reference's order:
 *   grafting.graft   (TF/grafting.py:224-300)  u = graft update of g: g itself (SGD) or
 *                    g * rsqrt(acc' + eps) with acc' = decay * acc + (1 - decay) * g^2
 *                    (acc + g^2 if decay == 1) (RMSPROP, TF/grafting.py:190-222);
 *                    x = precond * (|u| / |precond|, 0 if |precond| == 0) once use_precond is set
 *                    (count >= start_preconditioning_step), u before; parameters without a
 *                    direction (precond == NULL: masked by skip_preconditioning_*) always take u;
 *                    graft_type PC_TF_GRAFT_NONE passes precond (or g if NULL) through.
 *   momentum.apply   (TF/momentum.py:81-139)   x *= 1 - decay if ema; v' = x + decay * v;
 *                    x = x + decay * v' (nesterov) or v'; skipped if decay == 0; weight decay
 *                    x += weight_decay * param before or after it.
 *   learning rate    (TF/optimizer.py:91-99)   update = scale * x, scale = -learning_rate.
 * Norms are two-level sums in a fixed order.  Arrays are addressed per segment (no flat copy
 * of the model is needed); pointers 16-byte aligned.  acc / velocity are updated in place and
 * may be NULL when their stage is off; update may alias grad.
 *   segments       DEVICE [num_segments]; first_chunk = running sum of nchunks, nchunks =
 *                  ceil(numel / pc_graft_group_chunk_elems())
 *   chunk_segment  DEVICE [total_chunks] i32: the segment each chunk belongs to
 * ------------------------------------------------------------------------ */
#define PC_TF_GRAFT_NONE 0
#define PC_TF_GRAFT_SGD 1
#define PC_TF_GRAFT_RMSPROP 2
typedef struct {
  const float* grad;
  const float* param;   /* NULL allowed when weight_decay == 0 */
  const float* precond; /* NULL: parameter without a second-order direction */
  float* acc;
  float* velocity;
  float* update;
  int64_t numel;
  int32_t first_chunk;
  int32_t nchunks;
} pc_tearfree_segment;
typedef struct {
  int graft_type;
  float graft_decay;
  float graft_epsilon;
  int use_precond;
  int ema;
  int nesterov;
  float momentum_decay;
  float weight_decay;
  int weight_decay_after_momentum;
  float scale;
} pc_tearfree_options;
size_t pc_tearfree_transform_workspace_bytes(int num_segments, int64_t total_chunks);
int pc_tearfree_transform(const pc_tearfree_segment* segments, const int32_t* chunk_segment,
                          int num_segments, int64_t total_chunks, const pc_tearfree_options* opt,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * (5) all-gather of the block-sharded roots over NVLink peer memory
 * replaces: jax.lax.all_gather of the preconditioners and metrics (DS:2876-2877; the four
 *           gathers of the quantised path, DS:3116-3122) when every rank of the batch axis
 *           is a process on the same NVSwitch box.
 * Copy engines only -- no SM is used, so the exchange overlaps the persistent GEMM kernels of
 * the next sub-batch / bucket (an NCCL kernel would compete with them for SMs).  One process
 * per GPU: every rank allocates a receive buffer [world][slot_bytes] and a flag block of
 * PC_PEER_FLAG_WORDS uint32 (zeroed), exports both with pc_ipc_export, exchanges the handles
 * through its host-side launcher (torch.distributed in this repo), opens the peers' with
 * pc_ipc_open and fills a pc_peer_group (its own pointers at index `rank`).
 *   pc_peer_all_gather  stream-ordered: waits until every peer has released the previous
 *                       epoch, pushes `bytes` of `send` into slot `rank` of every receive
 *                       buffer (cudaMemcpyAsync peer-to-peer) followed by the epoch flag, then
 *                       makes `stream` wait (cuStreamWaitValue32) until all peers' payloads of
 *                       this epoch have landed here.  Epochs start at 1 and increase by 1.
 *   pc_peer_release     enqueue after the last consumer of the receive buffer: tells the peers
 *                       that this rank's buffer may be overwritten by epoch + 1.
 * ------------------------------------------------------------------------ */
#define PC_MAX_PEERS 16
#define PC_PEER_FLAG_WORDS (2 * PC_MAX_PEERS + 16)
typedef struct {
  unsigned char handle[64]; /* cudaIpcMemHandle_t of the allocation that contains the pointer */
  int64_t offset;           /* of the pointer inside that allocation */
  int64_t size;
  int32_t device;
  int32_t reserved;
} pc_ipc_handle;
typedef struct {
  int32_t world, rank;
  int64_t slot_bytes;
  void* recv[PC_MAX_PEERS];  /* receive buffers: own pointer at [rank], peers' mapped pointers */
  void* flags[PC_MAX_PEERS]; /* flag blocks, same convention */
} pc_peer_group;
int pc_ipc_export(const void* dev_ptr, pc_ipc_handle* out);
int pc_ipc_open(const pc_ipc_handle* handle, void** out_ptr);
int pc_peer_all_gather(const pc_peer_group* group, const void* send, size_t bytes, uint32_t epoch,
                       void* stream);
int pc_peer_release(const pc_peer_group* group, uint32_t epoch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PRECOND_B200_H_ */
