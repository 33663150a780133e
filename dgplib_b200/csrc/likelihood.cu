// Gaussian / SwitchedLikelihood variational expectations with the sample mean and the data sum folded in, forward and
// reverse mode in one pass.  Replaces the tf.map_fn over samples + reduce_mean + reduce_sum of
// dgplib/dsdgp.py:115-122 around gpflow's Gaussian.variational_expectations
//   VE = -0.5 log(2 pi) - 0.5 log s2 - 0.5 ((y - mu)^2 + v) / s2        
// and SwitchedLikelihood's row-wise selection by the integer in Y's last column.  HBM-bound elementwise work.
#include "common.cuh"

namespace dgp {

constexpr int LIK_THREADS = 256;
constexpr int LIK_MAX_BLOCKS = 1024;

__global__ void __launch_bounds__(LIK_THREADS)
k_lik_gaussian(const double* __restrict__ Fmu, const double* __restrict__ Fvar, const double* __restrict__ Y,
               long long y_rows, int ldy, long long n_rows, int R, const double* __restrict__ lik_var, int T,
               double weight, double* __restrict__ g_mu, double* __restrict__ g_var, double* __restrict__ partial) {
    __shared__ double red[32];
    const long long total = n_rows * R;
    const long long stride = (long long)gridDim.x * blockDim.x;
    double ve = 0.0;
    for (int tk = 0; tk < T; ++tk) {
        const double s2 = lik_var[tk];
        const double inv = 1.0 / s2;
        const double c0 = -0.5 * 1.8378770664093454836 - 0.5 * log(s2);  // -0.5 log(2 pi) - 0.5 log s2
        double gs = 0.0;
        for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
            const long long n = e / R;
            const int r = (int)(e % R);
            const double* yr = Y + (size_t)(n % y_rows) * ldy;
            if (T > 1) {
                const int id = (int)yr[ldy - 1];
                if (id != tk) {
                    // a task id outside [0, T) matches no likelihood (the reference's dynamic_partition would raise): the
                    // row contributes nothing, and its gradient slots are written (zeros) rather than left stale
                    if (tk == 0 && (id < 0 || id >= T) && g_mu) { g_mu[e] = 0.0; g_var[e] = 0.0; }
                    continue;
                }
            }
            const double dlt = yr[r] - Fmu[e];
            const double q = fma(dlt, dlt, Fvar[e]);
            ve += c0 - 0.5 * q * inv;
            gs += -0.5 * inv + 0.5 * q * inv * inv;
            if (g_mu) {
                g_mu[e] = weight * dlt * inv;
                g_var[e] = -0.5 * weight * inv;
            }
        }
        const double tot = block_sum(gs, red);
        if (threadIdx.x == 0) partial[(size_t)blockIdx.x * (T + 1) + 1 + tk] = tot;
    }
    const double tot = block_sum(ve, red);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * (T + 1)] = tot;
}

__global__ void k_lik_finish(const double* __restrict__ partial, int nblocks, int T, double weight,
                             double* __restrict__ ve_sum, double* __restrict__ g_lik_var) {
    const int c = threadIdx.x;
    if (c > T) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * (T + 1) + c];
    if (c == 0) { if (ve_sum) ve_sum[0] += weight * s; }
    else if (g_lik_var) g_lik_var[c - 1] += weight * s;
}

__global__ void k_philox_normal(double* __restrict__ out, long long n_rows, int R, int ld, unsigned long long seed,
                                unsigned long long row_offset, long long rps, long long sstride) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_rows * R) return;
    const long long n = e / R;
    const int r = (int)(e % R);
    const unsigned long long rid = row_offset + (unsigned long long)(n / rps) * sstride + (n % rps);
    out[(size_t)n * ld + r] = philox_normal(seed, rid, (uint32_t)r);
}

// ---- layer-0 sample sharing.  tile_over_samples (dgplib/utils.py:15-18) makes the S copies of a data row identical, so
// the first layer's conditional is evaluated once per data row and only the sampling step runs per (sample, row):
//   expand : F[s*B + b, c] = mean[b, c] + eps[s*B + b, c] * sqrt(var[b, c])   (+ task-id column copy when ldf > ldr)
//   reduce : g_mean[b, c] = sum_s g_F[s*B+b, c] ;  g_var[b, c] = sum_s g_F[s*B+b, c] (F[s*B+b, c] - mean[b, c]) / (2 var[b, c])
__global__ void k_sample_expand(const double* __restrict__ mean, const double* __restrict__ var, const double* __restrict__ eps,
                                unsigned long long seed, const unsigned long long* seed_dev,
                                unsigned long long row_offset, long long sstride, int S,
                                long long B, int R, int ldr, double* __restrict__ F, int ldf, const double* __restrict__ X,
                                int Dx) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)S * B * R) return;
    const int c = (int)(e % R);
    const long long n = e / R, b = n % B, s = n / B;
    const double mu = mean[(size_t)b * ldr + c], v = var[(size_t)b * ldr + c];
    const double z = eps ? eps[(size_t)n * ldr + c]
                         : philox_normal(seed + (seed_dev ? *seed_dev : 0ull),
                                         row_offset + (unsigned long long)s * sstride + b, (uint32_t)c);
    F[(size_t)n * ldf + c] = mu + z * sqrt(v);
    if (c == 0 && ldf > ldr) F[(size_t)n * ldf + ldf - 1] = X[(size_t)b * Dx + Dx - 1];
}
__global__ void k_sample_reduce(const double* __restrict__ gF, int ldgf, const double* __restrict__ F, int ldf,
                                const double* __restrict__ mean, const double* __restrict__ var, int S, long long B, int R,
                                int ldr, double* __restrict__ g_mean, double* __restrict__ g_var) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * R) return;
    const int c = (int)(e % R);
    const long long b = e / R;
    const double mu = mean[(size_t)b * ldr + c], inv = 1.0 / (2.0 * var[(size_t)b * ldr + c]);
    double gm = 0.0, gv = 0.0;
    for (int s = 0; s < S; ++s) {
        const size_t n = (size_t)s * B + b;
        const double g = gF[n * ldgf + c];
        gm += g;
        gv = fma(g * (F[n * ldf + c] - mu), inv, gv);
    }
    g_mean[(size_t)b * ldr + c] = gm;
    g_var[(size_t)b * ldr + c] = gv;
}

// gpflow Kernel.K / Kdiag for the descriptor's kernel (stand-alone entry points; not on the ELBO path)
__global__ void k_kernel_K(dgp_cond_desc d, const double* __restrict__ X1, long long n1, const double* __restrict__ X2,
                           long long n2, int symmetric, double* __restrict__ K) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = blockIdx.y;
    if (j >= n2) return;
    const double* xi = X1 + (size_t)i * d.Dx;
    const double* xj = X2 + (size_t)j * d.Dx;
    int ti = 0, tj = 0;
    if (d.T > 1) { ti = (int)xi[d.Dx - 1]; tj = (int)xj[d.Dx - 1]; }
    double v = 0.0;
    if (ti == tj && ti >= 0 && ti < d.T) {
        const double* th = d.theta + (size_t)ti * (2 + d.D);
        double ni = 0.0, nj = 0.0, c = 0.0;
        for (int k = 0; k < d.D; ++k) {
            const double a = xi[k] / th[2 + k], b = xj[k] / th[2 + k];
            ni = fma(a, a, ni); nj = fma(b, b, nj); c = fma(a, b, c);
        }
        v = th[0] * kern_val(d.kind[ti], ni + nj - 2.0 * c);
        if (symmetric && i == j && d.T == 1) v += th[1];
    }
    K[(size_t)i * n2 + j] = v;
}
__global__ void k_kernel_Kdiag(dgp_cond_desc d, const double* __restrict__ X1, long long n1, double* __restrict__ Kd) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    int ti = 0;
    if (d.T > 1) ti = (int)X1[(size_t)i * d.Dx + d.Dx - 1];
    double v = 0.0;
    if (ti >= 0 && ti < d.T) v = d.theta[(size_t)ti * (2 + d.D)] + d.theta[(size_t)ti * (2 + d.D) + 1];
    Kd[i] = v;
}

}  // namespace dgp

using namespace dgp;

extern "C" size_t dgp_lik_ws_bytes(int64_t, int32_t) { return (size_t)LIK_MAX_BLOCKS * (DGP_MAX_TASKS + 1) * sizeof(double); }

extern "C" int dgp_lik_gaussian(const double* Fmu, const double* Fvar, const double* Y, int64_t y_rows, int32_t ldy,
                                int64_t n_rows, int32_t R, const double* lik_var, int32_t T, double weight, double* ve_sum,
                                double* g_mu, double* g_var, double* g_lik_var, void* ws, size_t ws_bytes, void* stream) {
    if (!Fmu || !Fvar || !Y || !lik_var || !ws || n_rows < 0 || R < 1 || T < 1 || T > DGP_MAX_TASKS || y_rows < 1) return DGP_ERR_ARG;
    if (ldy < R + (T > 1 ? 1 : 0)) return DGP_ERR_ARG;
    if ((g_mu == nullptr) != (g_var == nullptr)) return DGP_ERR_ARG;
    if (ws_bytes < dgp_lik_ws_bytes(n_rows, R)) return DGP_ERR_WORKSPACE;
    if (n_rows == 0) return DGP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    long long total = (long long)n_rows * R;
    int blocks = (int)((total + LIK_THREADS * 4 - 1) / (LIK_THREADS * 4));
    if (blocks > LIK_MAX_BLOCKS) blocks = LIK_MAX_BLOCKS;
    if (blocks < 1) blocks = 1;
    k_lik_gaussian<<<blocks, LIK_THREADS, 0, st>>>(Fmu, Fvar, Y, y_rows, ldy, n_rows, R, lik_var, T, weight, g_mu, g_var, (double*)ws);
    DGP_COUNT_LAUNCH();
    k_lik_finish<<<1, 32, 0, st>>>((const double*)ws, blocks, T, weight, ve_sum, g_lik_var);
    DGP_COUNT_LAUNCH();
    DGP_CHECK_LAUNCH();
    return DGP_OK;
}

extern "C" int dgp_philox_normal(double* out, int64_t n_rows, int32_t R, int32_t ld, uint64_t seed, uint64_t row_offset,
                                 int64_t rows_per_sample, int64_t sample_stride, void* stream) {
    if (!out || n_rows < 0 || R < 1 || ld < R) return DGP_ERR_ARG;
    if (n_rows == 0) return DGP_OK;
    const long long rps = rows_per_sample > 0 ? rows_per_sample : n_rows;
    const long long ss = sample_stride > 0 ? sample_stride : rps;
    const long long total = (long long)n_rows * R;
    k_philox_normal<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out, n_rows, R, ld, seed, row_offset, rps, ss);
    DGP_COUNT_LAUNCH();
    DGP_CHECK_LAUNCH();
    return DGP_OK;
}

extern "C" int dgp_sample_expand(const double* mean, const double* var, const double* eps, uint64_t seed,
                                 const uint64_t* seed_dev, uint64_t row_offset, int64_t sample_stride, int32_t S, int64_t B, int32_t R, int32_t ldr, double* F, int32_t ldf,
                                 const double* X, int32_t Dx, void* stream) {
    if (!mean || !var || !F || S < 1 || B < 0 || R < 1 || ldr < R || ldf < ldr) return DGP_ERR_ARG;
    if (ldf > ldr && (!X || Dx < 1)) return DGP_ERR_ARG;
    if (B == 0) return DGP_OK;
    const long long total = (long long)S * B * R;
    k_sample_expand<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        mean, var, eps, seed, reinterpret_cast<const unsigned long long*>(seed_dev), row_offset,
        sample_stride > 0 ? sample_stride : B, S, B, R, ldr, F, ldf, X, Dx);
    DGP_COUNT_LAUNCH();
    DGP_CHECK_LAUNCH();
    return DGP_OK;
}
extern "C" int dgp_sample_reduce(const double* g_F, int32_t ldgf, const double* F, int32_t ldf, const double* mean,
                                 const double* var, int32_t S, int64_t B, int32_t R, int32_t ldr, double* g_mean,
                                 double* g_var, void* stream) {
    if (!g_F || !F || !mean || !var || !g_mean || !g_var || S < 1 || B < 0 || R < 1 || ldr < R || ldf < ldr || ldgf < R) return DGP_ERR_ARG;
    if (B == 0) return DGP_OK;
    k_sample_reduce<<<(unsigned)((B * R + 127) / 128), 128, 0, (cudaStream_t)stream>>>(g_F, ldgf, F, ldf, mean, var, S, B, R, ldr, g_mean, g_var);
    DGP_COUNT_LAUNCH();
    DGP_CHECK_LAUNCH();
    return DGP_OK;
}

extern "C" int dgp_kernel_K(const dgp_cond_desc* d, const double* X1, int64_t n1, const double* X2, int64_t n2,
                            int32_t symmetric, double* K, void* stream) {
    if (!d || !X1 || !K || !d->theta || n1 < 0 || d->D < 1 || d->Dx < d->D || d->T < 1 || d->T > DGP_MAX_TASKS) return DGP_ERR_ARG;
    if (symmetric) { X2 = X1; n2 = n1; }
    if (!X2 || n2 < 0) return DGP_ERR_ARG;
    if (n1 == 0 || n2 == 0) return DGP_OK;
    if (n1 > 65535) return DGP_ERR_UNSUPPORTED;
    k_kernel_K<<<dim3((unsigned)((n2 + 127) / 128), (unsigned)n1), 128, 0, (cudaStream_t)stream>>>(*d, X1, n1, X2, n2, symmetric, K);
    DGP_COUNT_LAUNCH();
    DGP_CHECK_LAUNCH();
    return DGP_OK;
}
extern "C" int dgp_kernel_Kdiag(const dgp_cond_desc* d, const double* X1, int64_t n1, double* Kdiag, void* stream) {
    if (!d || !X1 || !Kdiag || !d->theta || n1 < 0) return DGP_ERR_ARG;
    if (n1 == 0) return DGP_OK;
    k_kernel_Kdiag<<<(unsigned)((n1 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*d, X1, n1, Kdiag);
    DGP_COUNT_LAUNCH();
    DGP_CHECK_LAUNCH();
    return DGP_OK;
}
