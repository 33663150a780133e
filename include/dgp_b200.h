/*
 * dgp_b200.h -- C-ABI of the B200-native DSDGP hot path (libdgp_b200.so).
 *
 * The reference (aboustati/dgplib) has no FFI of its own: its hot path is a chain of GPflow/TensorFlow graph ops
 * built by Python.  Each entry point below replaces the TF sub-graph emitted by the cited reference
 * call site; the Python classes in dgplib_b200/ (same names as the reference's) bind them through ctypes.
 *
 * Conventions
 *   - every pointer named in a "device" field/argument is a CUDA device pointer to fp64 (row-major) owned by the
 *     caller; the library never allocates; scratch lives in caller-provided workspaces sized by the *_bytes calls
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and the call returns without syncing
 *   - return value: 0 = enqueued, <0 = argument / launch error (dgp_strerror).  Numerical failure (Cholesky pivot
 *     <= 0) is reported asynchronously in the int32 device word `status` of the layer workspace
 *     (0 ok, k>0 = failed at pivot k-1), read with dgp_layer_status() after a stream sync
 *   - stateless and re-entrant per stream
 */
#ifndef DGP_B200_H
#define DGP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGP_KIND_RBF 0       /* gpflow.kernels.RBF      : s2 * exp(-r2/2)                                   */
#define DGP_KIND_MATERN52 1  /* gpflow.kernels.Matern52 : s2 * (1+sqrt5 r+5/3 r^2) exp(-sqrt5 r), r=sqrt(r2+1e-12) */
#define DGP_MAX_TASKS 16     /* kernels inside one SwitchedKernel                                             */
#define DGP_MAX_D 64         /* kernel input dimension                                                       */
#define DGP_MAX_R 64         /* outputs of one conditional                                                   */

#define DGP_OK 0
#define DGP_ERR_ARG (-1)
#define DGP_ERR_LAUNCH (-2)
#define DGP_ERR_WORKSPACE (-3)
#define DGP_ERR_UNSUPPORTED (-4)

/*
 * One whitened sparse-GP conditional = one `gpflow.conditionals.conditional(Xnew, feature, kernel, q_mu,
 * q_sqrt=q_sqrt, full_cov=False, white=True)` call: dgplib/layers.py:66-69 (Layer: one per layer) and
 * dgplib/multikernel_layers.py:75-78 (MultikernelLayer: one per kernel, on a column slice of q_mu / q_sqrt).
 */
typedef struct dgp_cond_desc {
    int32_t M;        /* inducing points (Layer.num_inducing, dgplib/layers.py:34)                              */
    int32_t D;        /* kernel input_dim: the leading D columns of X / Z rows are the kernel's active dims      */
    int32_t Dx;       /* row stride of X and Z in doubles: D, or D+1 when a task-id column is carried            */
    int32_t R;        /* outputs of this conditional (output_dim, or offset = output_dim/num_kernels)             */
    int32_t ldr;      /* row stride of q_mu / mean / var / eps / gradients in doubles (the layer's output_dim)     */
    int32_t col0;     /* first output column of this conditional inside the layer: i*offset for kernel i of a      */
                      /*     MultikernelLayer (dgplib/multikernel_layers.py:75-76), 0 for a Layer                  */
    int32_t T;        /* 1 = stationary kernel (+White); >1 = SwitchedKernel over T kernels selected by the        */
                      /*     integer task id in column Dx-1 (dgplib/specialized_kernels.py:23-72)                 */
    int32_t kind[DGP_MAX_TASKS]; /* DGP_KIND_* of each of the T kernels                                          */
    double jitter;    /* gpflow.settings.numerics.jitter_level = 1e-6 (dgplib/utils.py:12)                        */
    /* device pointers */
    const double* Z;      /* [M, Dx]   feature.Z (dgplib/layers.py:36-41)                                         */
    const double* theta;  /* [T, 2+D]  per kernel: variance, White variance (0 if none), lengthscales[D]          */
    const double* q_mu;   /* [M, ldr]    the layer's q_mu; this conditional uses columns col0 .. col0+R-1        */
    const double* q_sqrt; /* [ldr, M, M] the layer's q_sqrt; rows col0 .. col0+R-1; only tril is read (band_part) */
    const double* mean_A; /* [Dx, ldr]   Linear mean weights (gpflow Linear.A), or NULL for Zero                  */
    const double* mean_b; /* [ldr]       Linear mean offset, or NULL                                              */
} dgp_cond_desc;

/* ---- parameter-only stage: Kuu(+jitter) -> Cholesky -> inverse factor, tril(q_sqrt)^T, KL ------------------- */
/* bytes of the per-conditional workspace; need_bwd != 0 adds the buffers of the backward pass; n_rows bounds the
 * number of rows (S*B) of any forward/backward call made with this workspace */
size_t dgp_cond_ws_bytes(const dgp_cond_desc* d, int64_t n_rows, int need_bwd);
/* Replaces: features.Kuu + tf.cholesky inside base_conditional, band_part(q_sqrt), and gauss_kl(q_mu, q_sqrt,
 * K=None) of dgplib/layers.py:58-60.  kl_out (device, 1 double) receives this conditional's KL if not NULL.
 * Stream semantics (also dgp_cond_backward_params): everything is ordered after the work already queued on `stream`
 * and complete, in stream order, for whatever is queued on it afterwards; internally the q_sqrt-only chain may run on
 * a library-owned helper stream that is forked from and joined back into `stream` with events before the call returns
 * (a capture on `stream` simply gains a parallel branch). */
int dgp_cond_prepare(const dgp_cond_desc* d, void* ws, size_t ws_bytes, int need_bwd, double* kl_out, void* stream);
/* copies the Cholesky status word to the host (synchronises the stream) */
int dgp_cond_status(const void* ws, void* stream, int32_t* status_host);
/* the same copy without the synchronisation: status_pinned (page-locked host memory) is valid once the work queued on
 * `stream` up to here has finished (the caller records an event behind it and looks one step later) */
int dgp_cond_status_async(const void* ws, void* stream, int32_t* status_pinned);

/* All per-row arrays below (eps, mean, var, F, g_*) are the LAYER's arrays with row stride ldr (F: ldf); a
 * conditional reads / writes only its own columns col0 .. col0+R-1. */
/* ---- per-row stage, forward ----------------------------------------------------------------------------------
 * Replaces: Kuf, matrix_triangular_solve, A^T q_mu, fvar = Knn - sum A^2 + sum (L^T A)^2, + mean_function(X)
 * (dgplib/layers.py:65-71, flattened over samples as in :84-87) and normal_sample's diag branch
 * (dgplib/utils.py:25-27).
 *   X       [x_rows, Dx]; row n of the flattened [S*B] batch reads X row (n % x_rows): tile_over_samples
 *           (dgplib/utils.py:15-18) without the copies
 *   eps     [n_rows, ldr] N(0,1) draws, or NULL to draw Philox4x32-10 normals keyed by (seed, global row id, column)
 *           with global row id = row_offset + (n / rows_per_sample) * sample_stride + n % rows_per_sample, so that a
 *           shard holding rows_per_sample of sample_stride minibatch rows draws what the unsharded run draws
 *           (rows_per_sample <= 0: n_rows; sample_stride <= 0: rows_per_sample).  seed_dev (device, nullable) is added to
 *           seed on the device, so that a captured CUDA graph can be replayed with fresh noise
 *   mean,var[n_rows, ldr] outputs (this conditional's R columns)
 *   F       [n_rows, ldf] sample mean + eps*sqrt(var) (NULL: skip).  When ldf > ldr the trailing column Dx-1 of X
 *           is copied into column ldf-1 (task-label propagation, dgplib/multitask_dsdgp.py:20)
 *   A_save  [n_rows, Mpad] Lm^-1 Kuf per row, kept for the backward pass (NULL: skip)
 *   U_save  dgp_usave_bytes(d, n_rows) bytes: U_r = tril(q_sqrt_r)^T A for every output r, kept for the backward pass
 *           (the reference's [R, M, N'] LTA tensor of base_conditional, which TF keeps for its own reverse mode) in
 *           16 x 8 blocks [r][m / 16][n / 8][m % 16][n % 8]; NULL: not kept (prediction)
 *   K_save  [n_rows, Mpad] Kuf = k(Z, x_n) per row (nullable): for an RBF kernel, or a SwitchedKernel whose T kernels are
 *           all RBFs, the backward pass then needs no exp (k' = -k / 2); ignored there for other kernels, which recompute
 *           the covariance function
 */
int dgp_cond_forward(const dgp_cond_desc* d, const void* ws, const double* X, int64_t x_rows, int64_t n_rows,
                     const double* eps, uint64_t seed, const uint64_t* seed_dev, uint64_t row_offset,
                     int64_t rows_per_sample, int64_t sample_stride, double* mean, double* var, double* F, int32_t ldf,
                     double* A_save, double* U_save, double* K_save, void* stream);
/* MultikernelLayer (dgplib/multikernel_layers.py:74-85: one conditional per kernel, built in a Python loop): the n
 * conditionals of one layer on the same rows in ONE launch over (conditional, row tile) pairs.  d / ws / A_save / U_save /
 * K_save are host arrays of n pointers (the saved-tensor arrays may be NULL as a whole); everything else as in
 * dgp_cond_forward.  Members must agree in M, D, Dx, R, T, ldr and n <= 8, otherwise DGP_ERR_UNSUPPORTED (launch them one
 * by one).  dgp_cond_backward_group is the same for dgp_cond_backward when no input gradient is wanted (gX accumulates
 * over the members); dgp_cond_backward_gram / dgp_cond_backward_params stay per member. */
int dgp_cond_forward_group(int32_t n, const dgp_cond_desc* const* d, const void* const* ws, const double* X, int64_t x_rows,
                           int64_t n_rows, const double* eps, uint64_t seed, const uint64_t* seed_dev, uint64_t row_offset,
                           int64_t rows_per_sample, int64_t sample_stride, double* mean, double* var, double* F,
                           int32_t ldf, double* const* A_save, double* const* U_save, double* const* K_save, void* stream);
int dgp_cond_backward_group(int32_t n, const dgp_cond_desc* const* d, void* const* ws, const double* X, int64_t x_rows,
                            int64_t n_rows, const double* const* A_save, const double* const* U_save,
                            const double* const* K_save, const double* mean, const double* var, const double* F,
                            int32_t ldf, const double* g_mean_in, const double* g_var_in, const double* g_F, int32_t ldgf,
                            int32_t mean_grad, void* stream);
/* bytes of the U_save buffer for n_rows rows of this conditional (R * Mpad * 8 * ceil(n_rows / 8) * 8) */
size_t dgp_usave_bytes(const dgp_cond_desc* d, int64_t n_rows);

/* ---- first-layer sample sharing -------------------------------------------------------------------------------
 * tile_over_samples (dgplib/utils.py:15-18, used at dgplib/dsdgp.py:89) hands the first layer S identical copies of
 * every data row, so its conditional is the same for all of them: call dgp_cond_forward on the B data rows only
 * (F = NULL), then draw the S samples per row with dgp_sample_expand; in the backward pass dgp_sample_reduce folds the
 * S sample gradients of a row into (g_mean, g_var) [B, ldr], which go to dgp_cond_backward as g_mean_in / g_var_in.
 *   F[s*B + b, c] = mean[b, c] + eps[s*B + b, c] * sqrt(var[b, c]),  eps == NULL: Philox keyed (seed, row_offset +
 *   s * sample_stride + b, c) -- the draws dgp_cond_forward would make.  ldf > ldr: column ldf-1 = X[b, Dx-1]. */
int dgp_sample_expand(const double* mean, const double* var, const double* eps, uint64_t seed, const uint64_t* seed_dev,
                      uint64_t row_offset, int64_t sample_stride, int32_t S, int64_t B, int32_t R, int32_t ldr, double* F, int32_t ldf,
                      const double* X, int32_t Dx, void* stream);
int dgp_sample_reduce(const double* g_F, int32_t ldgf, const double* F, int32_t ldf, const double* mean,
                      const double* var, int32_t S, int64_t B, int32_t R, int32_t ldr, double* g_mean, double* g_var,
                      void* stream);

/* ---- likelihood ------------------------------------------------------------------------------------------------
 * Gaussian / SwitchedLikelihood variational expectations, mean over S, sum over rows and outputs
 * (dgplib/dsdgp.py:115-122), and their gradients.
 *   Fmu, Fvar [n_rows, R];  Y [y_rows, ldy] (row n reads Y row n % y_rows; ldy = R, or R+1 with the task column)
 *   lik_var   [T] device; T = 1 Gaussian, T > 1 SwitchedLikelihood
 *   weight    = (num_data / B) / S   (dgplib/dsdgp.py:121, :128-129)
 *   ve_sum    device double, += weight * sum VE      g_mu, g_var [n_rows, R] (NULL: forward only)
 *   g_lik_var [T] device, += d/d lik_var
 */
int dgp_lik_gaussian(const double* Fmu, const double* Fvar, const double* Y, int64_t y_rows, int32_t ldy,
                     int64_t n_rows, int32_t R, const double* lik_var, int32_t T, double weight, double* ve_sum,
                     double* g_mu, double* g_var, double* g_lik_var, void* ws, size_t ws_bytes, void* stream);
size_t dgp_lik_ws_bytes(int64_t n_rows, int32_t R);

/* ---- per-row stage, backward ---------------------------------------------------------------------------------
 * Reverse mode of dgp_cond_forward (what tf.gradients emits for dgplib/layers.py:65-71 and dgplib/utils.py:27).
 *   g_mean_in, g_var_in [n_rows, ldr]  direct gradients of mean / var (last layer: from the likelihood), or NULL
 *   g_F       [n_rows, ldgf] gradient of the sample F (from the next layer's gX), or NULL;  then
 *             g_mean = g_mean_in + g_F,  g_var = g_var_in + g_F * (F - mean) / (2 var)
 *   A_save, U_save, K_save  what dgp_cond_forward kept for these rows (K_save may be NULL)
 *   gX        [n_rows, Dx] gradient w.r.t. the input rows (NULL for the data layer); accumulate != 0 adds to it
 *   mean_grad != 0: the Linear mean (d->mean_A, d->mean_b) is trainable -- an OutputLayer built with its own Linear mean,
 *             which dgplib/layers.py:211-218 does not freeze: also accumulate dA = X^T g_mean, db = sum_n g_mean
 * dgp_cond_backward leaves dKuf and the combined g_mean / g_var in the workspace; dgp_cond_backward_gram (call it next,
 * same n_rows / A_save) reduces them over the rows into the Gram accumulators G_r = sum_n g_var[n,r] a_n a_n^T and
 * P = sum_n dKuf_n a_n^T.  Row-summed parameter gradients are accumulated inside the workspace (several row chunks
 * may be pushed between one dgp_cond_prepare and dgp_cond_backward_params, which finishes them).
 */
int dgp_cond_backward(const dgp_cond_desc* d, void* ws, const double* X, int64_t x_rows, int64_t n_rows,
                      const double* A_save, const double* U_save, const double* K_save, const double* mean,
                      const double* var, const double* F, int32_t ldf, const double* g_mean_in, const double* g_var_in, const double* g_F, int32_t ldgf,
                      double* gX, int32_t accumulate_gX, int32_t mean_grad, void* stream);
int dgp_cond_backward_gram(const dgp_cond_desc* d, void* ws, const double* A_save, int64_t n_rows, int32_t accumulate,
                           void* stream);   /* accumulate: 0 for the first row chunk after dgp_cond_prepare, 1 after */
/* Parameter-only tail: Cholesky / Kuu / KL reverse mode.  Outputs (device, this conditional's part OVERWRITTEN):
 *   gZ [M, Dx], gtheta [T, 2+D], gq_mu [M, ldr] (columns col0..), gq_sqrt [ldr, M, M] (rows col0..),
 *   gmean_A [Dx, ldr], gmean_b [ldr] (columns col0..; NULL unless dgp_cond_backward ran with mean_grad)
 *   kl_weight: this rank's share of the KL term (1/world_size) so that an all-reduce(sum) over ranks is exact */
int dgp_cond_backward_params(const dgp_cond_desc* d, void* ws, double kl_weight, double* gZ, double* gtheta,
                             double* gq_mu, double* gq_sqrt, double* gmean_A, double* gmean_b, void* stream);

/* ---- full-covariance branch (small N: plots, posterior draws) ---------------------------------------------------
 * dgp_cond_fullcov replaces base_conditional(full_cov=True) as called from dgplib/layers.py:74-81:
 *   fvar[r] = k(X, X) - A^T A + (Lq_r^T A)^T (Lq_r^T A)   for the conditional's R outputs, fvar [R, n, n] (n <= 16384:
 *   the *_ws_bytes calls return 0 beyond; above 1024 points the covariance is factored by a blocked multi-launch Cholesky);
 *   A_save [n, Mpad] comes from dgp_cond_forward on the same rows.
 * dgp_chol_sample replaces normal_sample's full-covariance branch (dgplib/utils.py:28-36) and the sampling step of
 * predict_f_samples (dgplib/dsdgp.py:170-185): for b < batch
 *   out[b, :, :] = mean[b] + chol(var[b] + jitter I) z[b],  var[b](i, j) = var[b*stride_b + i*stride_i + j*stride_j],
 *   z [batch, n, nz], out [batch, n, nz], mean[b](i) = mean[b*mean_stride_b + i*mean_stride_i] (NULL: zero).
 *   status_dev (nullable) [batch]: failing pivot + 1 of each factorisation. */
size_t dgp_cond_fullcov_ws_bytes(const dgp_cond_desc* d, int64_t n);
int dgp_cond_fullcov(const dgp_cond_desc* d, const void* ws, const double* X, int64_t n, const double* A_save,
                     double* fvar, void* scratch, size_t scratch_bytes, void* stream);
size_t dgp_chol_sample_ws_bytes(int64_t n, int32_t batch);
int dgp_chol_sample(const double* var, int64_t stride_b, int64_t stride_i, int64_t stride_j, int64_t n, int32_t batch,
                    double jitter, const double* z, int32_t nz, const double* mean, int64_t mean_stride_b,
                    int64_t mean_stride_i, double* out, void* scratch, size_t scratch_bytes, int32_t* status_dev,
                    void* stream);

/* ---- layer initialisation: PCA weights of a narrowing layer ------------------------------------------------------
 * dgp_pca_weights replaces the `np.linalg.svd(X)` of find_weights (dgplib/layers.py:104-109, input_dim > output_dim):
 *   W [D, R] (row-major) = the R leading right-singular vectors of X [n_rows, D] (row pitch ldx; the caller drops the
 *   task-id column of a multitask X by passing D = columns - 1), computed as eigenvectors of X^T X (one pass over X,
 *   then a D x D Jacobi).  sing (nullable) [D]: the singular values, descending.  Sign convention: the component of
 *   largest magnitude of each vector is negative (an SVD leaves the sign open; this choice reproduces the reference's
 *   golden vector, tests/test_layer.py:33-38).  D <= DGP_MAX_D. */
size_t dgp_pca_ws_bytes(int64_t n_rows, int32_t D);
int dgp_pca_weights(const double* X, int64_t n_rows, int32_t ldx, int32_t D, int32_t R, void* ws, size_t ws_bytes,
                    double* W, double* sing, void* stream);

/* ---- training step around the evaluation ------------------------------------------------------
 * The reference trains with gpflow.train.AdamOptimizer(lr).minimize(model, maxiter) (tests/test_dsdgp.py:56-60): TF's Adam
 * on the unconstrained variables of -ELBO, positive parameters through gpflow's Log1pe transform softplus(u) + 1e-6.
 * All of a model's unconstrained parameters live in one flat buffer u [n] (device).
 * dgp_transform_params: the composite constrained arrays the kernels read (per conditional theta [T, 2+D], likelihood
 *   variances) as c[i] = sum_{e in crow[i] .. crow[i+1]} (cpos[e] ? softplus(u[cidx[e]]) + lower : u[cidx[e]]).
 * dgp_adam_step: for every trainable element j (flags bit 1; bit 0: positive transform)
 *   g = -(gidx[j] >= 0 ? grad[gidx[j]] : sum_{e in grow[r] .. grow[r+1]} grad[gimg[e]], r = -2 - gidx[j]; -1: none)
 *       * (positive ? sigmoid(u[j]) : 1);   t = *step_dev + 1;   lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t)
 *   m = beta1 m + (1 - beta1) g;  v = beta2 v + (1 - beta2) g^2;  u -= lr_t m / (sqrt(v) + eps)      (tf.train.AdamOptimizer)
 *   then *step_dev += 1 (device counter: the launch pair can be replayed from a CUDA graph).  grad is the flat
 *   [elbo | gradients] buffer the evaluation filled (dgp_cond_backward_params / dgp_lik_gaussian outputs).
 * dgp_set_u64: stream-ordered device write of one 64-bit word (the Philox seed offset of a graph replay). */
int dgp_transform_params(const double* u, const int32_t* crow, const int32_t* cidx, const uint8_t* cpos, int32_t n_c,
                         double lower, double* c, void* stream);
int dgp_adam_step(double* u, double* m, double* v, int64_t n, const double* grad, const int32_t* gidx,
                  const int32_t* grow, const int32_t* gimg, const uint8_t* flags, double lr, double beta1, double beta2,
                  double eps, int64_t* step_dev, void* stream);
int dgp_set_u64(uint64_t* dst, uint64_t value, void* stream);
/* closes an evaluation (dgplib/dsdgp.py:124-130): *elbo -= kl_weight * sum_c kl[c] over the ncond per-conditional KL terms
 * dgp_cond_prepare wrote (summed in index order); dgp_zero: cudaMemsetAsync of the flat [elbo | gradients | kl] buffer */
int dgp_finish_elbo(double* elbo, const double* kl, int32_t ncond, double kl_weight, void* stream);
int dgp_zero(void* dst, size_t bytes, void* stream);

/* ---- stand-alone kernel-matrix entry points (dgplib_b200.specialized_kernels; gpflow Kernel.K / Kdiag) -------
 *   K [n1, n2] = k(X1, X2) (White contributes nothing, as in gpflow when X2 is given);  symmetric != 0 computes
 *   k(X1, X1) + White on the diagonal.  Kdiag [n1]. */
int dgp_kernel_K(const dgp_cond_desc* d, const double* X1, int64_t n1, const double* X2, int64_t n2,
                 int32_t symmetric, double* K, void* stream);
int dgp_kernel_Kdiag(const dgp_cond_desc* d, const double* X1, int64_t n1, double* Kdiag, void* stream);

/* ---- utilities ------------------------------------------------------------------------------------------------ */
int32_t dgp_mpad(int32_t M);                 /* padded inducing count used for A_save rows and the workspaces      */
int dgp_philox_normal(double* out, int64_t n_rows, int32_t R, int32_t ld, uint64_t seed, uint64_t row_offset,
                      int64_t rows_per_sample, int64_t sample_stride,
                      void* stream);         /* the same draws dgp_cond_forward makes when eps == NULL             */
const char* dgp_strerror(int code);
const char* dgp_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t dgp_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DGP_B200_H */
