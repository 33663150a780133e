// Full-covariance branch of the conditional and of the sampler (small N: plots and posterior
// draws).  Replaces gpflow's base_conditional(full_cov=True) as called from dgplib/layers.py:74-81
//     fvar_r = Knn - A^T A + (Lq_r^T A)^T (Lq_r^T A)          [R, N, N],  Knn = k(X, X) incl. White
// and normal_sample's full-covariance branch (dgplib/utils.py:28-36) / predict_f_samples (dgplib/dsdgp.py:170-185)
//     f = mean + chol(var + jitter I) z.
// O(R N^2 M + R N^3): parameter-stage class work, done with the small strided GEMM.  Up to 1024 points the covariance
// is factored by the one-CTA Cholesky of the parameter stage; above (to FULLCOV_MAX_N points) by a blocked right-looking
// factorisation whose 256-wide diagonal blocks go through that same kernel and whose panel solves and trailing updates
// are GEMMs over all SMs.
#include "common.cuh"
#include "gemm_small.cuh"

namespace dgp {

__global__ void k_pad_rows(const double* __restrict__ src, int n, int ld, double* __restrict__ dst, int np) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= ld) return;
    dst[(size_t)i * ld + j] = (i < n) ? src[(size_t)i * ld + j] : 0.0;
}

// out[r][i][j] = k(x_i, x_j) (+ White on the diagonal) - G[i][j] + H[r][i][j]
__global__ void k_fullcov_finish(dgp_cond_desc d, const double* __restrict__ X, int n, int np, const double* __restrict__ G,
                                 const double* __restrict__ H, double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, r = blockIdx.z;
    if (j >= n) return;
    const double* xi = X + (size_t)i * d.Dx;
    const double* xj = X + (size_t)j * d.Dx;
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
        // White: the symmetric call K(X) of a plain kernel; inside a SwitchedKernel sub-kernels see an explicit X2
        if (i == j && d.T == 1) v += th[1];
    }
    out[((size_t)r * n + i) * n + j] = v - G[(size_t)i * np + j] + H[((size_t)r * np + i) * np + j];
}

// padded copy of `batch` n x n matrices (strides given in doubles) with + jitter on the diagonal, identity on the pad
__global__ void k_pad_cov(const double* __restrict__ var, long long sb, long long si, long long sj, int n, int np,
                          double jitter, double* __restrict__ dst) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    if (j >= np) return;
    double v = (i == j) ? 1.0 : 0.0;
    if (i < n && j < n) {
        v = var[b * sb + i * si + j * sj];
        if (i == j) v += jitter;
    }
    dst[((size_t)b * np + i) * np + j] = v;
}
// out[b][i][c] = mean[b][i] (strided) + sum_{k <= i} L[b][i][k] z[b][k][c]
__global__ void k_chol_apply(const double* __restrict__ L, int n, int np, const double* __restrict__ z, int nz,
                             const double* __restrict__ mean, long long mb, long long mi, double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    if (c >= nz) return;
    const double* Lr = L + ((size_t)b * np + i) * np;
    const double* zb = z + (size_t)b * n * nz;
    double s = mean ? mean[b * mb + i * mi] : 0.0;
    for (int k = 0; k <= i; ++k) s = fma(Lr[k], zb[(size_t)k * nz + c], s);
    out[((size_t)b * n + i) * nz + c] = s;
}

// ---- blocked Cholesky for np > 1024 (np a multiple of 64): K [batch][np][np] lower factor in place
constexpr int CB = 256;            // diagonal block width
constexpr int FULLCOV_MAX_N = 16384;

// D[b] = the CB x CB block of K[b] at (j0, j0), cb <= CB wide; pad rows / columns of D form an identity
__global__ void k_block_get(const double* __restrict__ K, int np, int j0, int cb, double* __restrict__ D) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    if (j >= CB) return;
    double v = (i == j) ? 1.0 : 0.0;
    if (i < cb && j < cb) v = K[((size_t)b * np + j0 + i) * np + j0 + j];
    D[((size_t)b * CB + i) * CB + j] = v;
}
__global__ void k_block_put(double* __restrict__ K, int np, int j0, int cb, const double* __restrict__ D) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    if (j >= cb || i >= cb) return;
    K[((size_t)b * np + j0 + i) * np + j0 + j] = (j <= i) ? D[((size_t)b * CB + i) * CB + j] : 0.0;
}
// the solved panel P [b][rows][CB] back into K[b][j0 + cb .., j0 .. j0 + cb)
__global__ void k_panel_put(double* __restrict__ K, int np, int j0, int cb, int rows, const double* __restrict__ P,
                            long long pstride) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    if (j >= cb || i >= rows) return;
    K[((size_t)b * np + j0 + cb + i) * np + j0 + j] = P[(size_t)b * pstride + (size_t)i * CB + j];
}
__global__ void k_status_merge(const int32_t* __restrict__ blk, int j0, int batch, int32_t* __restrict__ status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch && status[b] == 0 && blk[b] != 0) status[b] = j0 + blk[b];
}

size_t chol_blocked_scratch_doubles(int np, int batch) {
    return (size_t)batch * (2 * CB * CB + (size_t)np * CB) + 64;        // D, Dinv, panel P (+ block status words)
}

int cholesky_blocked(double* K, int np, int batch, int32_t* status, double* scratch, cudaStream_t st) {
    double* D = scratch;
    double* Dinv = D + (size_t)batch * CB * CB;
    double* P = Dinv + (size_t)batch * CB * CB;
    int32_t* blk = reinterpret_cast<int32_t*>(P + (size_t)batch * np * CB);
    if (batch > 128) return DGP_ERR_UNSUPPORTED;                        // 64 doubles of status words
    if (cudaMemsetAsync(status, 0, sizeof(int32_t) * batch, st) != cudaSuccess) return DGP_ERR_LAUNCH;
    const long long pstride = (long long)np * CB;
    for (int j0 = 0; j0 < np; j0 += CB) {
        const int cb = (np - j0 < CB) ? np - j0 : CB;                    // np is a multiple of 64, so is cb
        const int rows = np - j0 - cb;
        k_block_get<<<dim3(CB / 128, CB, batch), 128, 0, st>>>(K, np, j0, cb, D); DGP_COUNT_LAUNCH();
        int rc = cholesky_batched(D, CB, batch, blk, st, Dinv);
        if (rc != DGP_OK) return rc;
        k_status_merge<<<(batch + 127) / 128, 128, 0, st>>>(blk, j0, batch, status); DGP_COUNT_LAUNCH();
        k_block_put<<<dim3((cb + 127) / 128, cb, batch), 128, 0, st>>>(K, np, j0, cb, D); DGP_COUNT_LAUNCH();
        if (rows == 0) break;
        for (int b = 0; b < batch; ++b)
            if ((rc = tri_inv_cols(D + (size_t)b * CB * CB, Dinv + (size_t)b * CB * CB, CB, st)) != DGP_OK) return rc;
        // panel: P = K[j0 + cb .., j0 .. j0 + cb) D^-T   (B(k, j) = Dinv[j][k]; the identity pad of D keeps cb < CB exact)
        GemmArgs g{};
        g.A = K + (size_t)(j0 + cb) * np + j0; g.B = Dinv; g.C = P;
        g.I = rows; g.J = CB; g.K = cb; g.batch = batch;
        g.sAi = np; g.sAk = 1; g.sBk = 1; g.sBj = CB; g.sCi = CB; g.sCj = 1;
        g.bA = (long long)np * np; g.bB = (long long)CB * CB; g.bC = pstride; g.alpha = 1.0;
        if ((rc = gemm_small(g, st)) != DGP_OK) return rc;
        // trailing update (lower tiles): K[j0 + cb .., j0 + cb ..] -= P P^T
        g = GemmArgs{};
        g.A = P; g.B = P; g.C = K + (size_t)(j0 + cb) * np + j0 + cb;
        g.I = rows; g.J = rows; g.K = cb; g.batch = batch;
        g.sAi = CB; g.sAk = 1; g.sBk = 1; g.sBj = CB; g.sCi = np; g.sCj = 1;
        g.bA = g.bB = pstride; g.bC = (long long)np * np; g.alpha = -1.0; g.beta = 1.0; g.lower_tiles_only = 1;
        if ((rc = gemm_small(g, st)) != DGP_OK) return rc;
        k_panel_put<<<dim3((cb + 127) / 128, rows, batch), 128, 0, st>>>(K, np, j0, cb, rows, P, pstride); DGP_COUNT_LAUNCH();
    }
    return cudaGetLastError() == cudaSuccess ? DGP_OK : DGP_ERR_LAUNCH;
}

}  // namespace dgp

using namespace dgp;

static inline int np_of(int64_t n) { return (int)((n + 63) / 64 * 64); }

extern "C" size_t dgp_cond_fullcov_ws_bytes(const dgp_cond_desc* d, int64_t n) {
    if (!d || n < 1 || n > FULLCOV_MAX_N) return 0;
    const size_t np = np_of(n), Mp = dgp_mpad(d->M);
    return (np * Mp * (1 + d->R) + np * np * (1 + d->R)) * sizeof(double) + 256;
}

extern "C" int dgp_cond_fullcov(const dgp_cond_desc* d, const void* ws, const double* X, int64_t n, const double* A_save,
                                double* fvar, void* scratch, size_t scratch_bytes, void* stream) {
    int rc = check_desc(d);
    if (rc != DGP_OK) return rc;
    if (!ws || !X || !A_save || !fvar || !scratch || n < 1) return DGP_ERR_ARG;
    if (n > FULLCOV_MAX_N) return DGP_ERR_UNSUPPORTED;
    if (scratch_bytes < dgp_cond_fullcov_ws_bytes(d, n)) return DGP_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const WsLayout w = ws_layout(*d, 0, 0);
    const int Mp = w.Mpad, np = np_of(n), R = d->R;
    const double* LqT = ws_ptr<double>(ws, w.LqT);
    double* Ap = static_cast<double*>(scratch);              // [np][Mp]
    double* Up = Ap + (size_t)np * Mp;                       // [R][np][Mp]
    double* G = Up + (size_t)R * np * Mp;                    // [np][np]
    double* H = G + (size_t)np * np;                         // [R][np][np]
    k_pad_rows<<<dim3((Mp + 127) / 128, np), 128, 0, st>>>(A_save, (int)n, Mp, Ap, np); DGP_COUNT_LAUNCH();
    GemmArgs g{};
    // G = A A^T
    g.A = Ap; g.B = Ap; g.C = G; g.I = np; g.J = np; g.K = Mp; g.batch = 1;
    g.sAi = Mp; g.sAk = 1; g.sBk = 1; g.sBj = Mp; g.sCi = np; g.sCj = 1; g.alpha = 1.0;
    if ((rc = gemm_small(g, st)) != DGP_OK) return rc;
    // U_r = A Lq_r :  B(k, m) = Lq_r[k][m] = LqT_r[m][k]
    g = GemmArgs{};
    g.A = Ap; g.B = LqT; g.C = Up; g.I = np; g.J = Mp; g.K = Mp; g.batch = R;
    g.sAi = Mp; g.sAk = 1; g.sBk = 1; g.sBj = Mp; g.sCi = Mp; g.sCj = 1;
    g.bA = 0; g.bB = (long long)Mp * Mp; g.bC = (long long)np * Mp; g.alpha = 1.0;
    if ((rc = gemm_small(g, st)) != DGP_OK) return rc;
    // H_r = U_r U_r^T
    g = GemmArgs{};
    g.A = Up; g.B = Up; g.C = H; g.I = np; g.J = np; g.K = Mp; g.batch = R;
    g.sAi = Mp; g.sAk = 1; g.sBk = 1; g.sBj = Mp; g.sCi = np; g.sCj = 1;
    g.bA = g.bB = (long long)np * Mp; g.bC = (long long)np * np; g.alpha = 1.0;
    if ((rc = gemm_small(g, st)) != DGP_OK) return rc;
    k_fullcov_finish<<<dim3((unsigned)((n + 127) / 128), (unsigned)n, R), 128, 0, st>>>(*d, X, (int)n, np, G, H, fvar);
    DGP_COUNT_LAUNCH();
    DGP_CHECK_LAUNCH();
    return DGP_OK;
}

extern "C" size_t dgp_chol_sample_ws_bytes(int64_t n, int32_t batch) {
    if (n < 1 || n > FULLCOV_MAX_N || batch < 1 || (n > 1024 && batch > 128)) return 0;
    const size_t np = np_of(n);
    size_t doubles = (size_t)batch * np * np;
    if (np > 1024) doubles += chol_blocked_scratch_doubles((int)np, batch);
    return doubles * sizeof(double) + (size_t)batch * sizeof(int32_t) + 256;
}

extern "C" int dgp_chol_sample(const double* var, int64_t stride_b, int64_t stride_i, int64_t stride_j, int64_t n,
                               int32_t batch, double jitter, const double* z, int32_t nz, const double* mean,
                               int64_t mean_stride_b, int64_t mean_stride_i, double* out, void* scratch,
                               size_t scratch_bytes, int32_t* status_dev, void* stream) {
    if (!var || !z || !out || !scratch || n < 1 || batch < 1 || nz < 1) return DGP_ERR_ARG;
    if (n > FULLCOV_MAX_N || (n > 1024 && batch > 128)) return DGP_ERR_UNSUPPORTED;
    if (scratch_bytes < dgp_chol_sample_ws_bytes(n, batch)) return DGP_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int np = np_of(n);
    double* L = static_cast<double*>(scratch);
    double* extra = L + (size_t)batch * np * np;                      // blocked-factorisation scratch (np > 1024)
    if (np > 1024) extra += chol_blocked_scratch_doubles(np, batch);
    int32_t* status = status_dev ? status_dev : reinterpret_cast<int32_t*>(extra);
    k_pad_cov<<<dim3((np + 127) / 128, np, batch), 128, 0, st>>>(var, stride_b, stride_i, stride_j, (int)n, np, jitter, L);
    DGP_COUNT_LAUNCH();
    int rc = (np <= 1024) ? cholesky_batched(L, np, batch, status, st)
                          : cholesky_blocked(L, np, batch, status, L + (size_t)batch * np * np, st);
    if (rc != DGP_OK) return rc;
    k_chol_apply<<<dim3((nz + 31) / 32, (unsigned)n, batch), 32, 0, st>>>(L, (int)n, np, z, nz, mean, mean_stride_b,
                                                                          mean_stride_i, out);
    DGP_COUNT_LAUNCH();
    DGP_CHECK_LAUNCH();
    return DGP_OK;
}
