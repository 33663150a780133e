// Initial Linear-mean weights of a narrowing layer: the leading right-singular vectors of X (dgplib/layers.py:97-121,
// `find_weights`, the input_dim > output_dim branch: `_, _, V = np.linalg.svd(X); W = V[:output_dim].T`).
//
// The right-singular vectors of X are the eigenvectors of G = X^T X, so the N x D SVD becomes
//   k_xtx      one pass over X (HBM-bound for D <= ~24, fp64 tensor pipe above): G partials per CTA on DMMA.8x8x4,
//   k_pca_eig  one CTA: fixed-order sum of the partials, cyclic two-sided Jacobi on the D x D matrix (round-robin
//              ordering, D/2 disjoint rotations per round), sort, sign convention, W[D, R] out.
// Sign: an SVD fixes each vector only up to sign; LAPACK's choice is an implementation detail.  Here the component of
// largest magnitude of every vector is made negative, which reproduces the reference's own golden vector
// (tests/test_layer.py:33-38: X = [[2,4],[1,3],[0,0],[0,0]] -> [-0.4, -0.91]).
#include "common.cuh"

namespace dgp {

constexpr int XTX_THREADS = 256;

// NTI = number of 8-column tiles (D <= 8 NTI).  Each warp walks 4-row groups of X: lane (g, t) loads X[row0 + t][8 i + g],
// which is at once the A fragment of tile row i and the B fragment of tile column i of the 8x8x4 product.
template <int NTI>
__global__ void __launch_bounds__(XTX_THREADS) k_xtx(const double* __restrict__ X, int64_t n_rows, int ldx, int D,
                                                     double* __restrict__ part) {
    constexpr int DP = 8 * NTI;
    constexpr int U = NTI >= 8 ? 2 : (NTI >= 4 ? 4 : 8);   // row groups in flight per warp
    __shared__ double G[DP * DP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    double acc[NTI][NTI][2];
#pragma unroll
    for (int i = 0; i < NTI; ++i)
#pragma unroll
        for (int j = 0; j < NTI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int64_t ngroups = (n_rows + 3) / 4;
    const int64_t nwarps = (int64_t)gridDim.x * (XTX_THREADS / 32);
    for (int64_t grp = (int64_t)blockIdx.x * (XTX_THREADS / 32) + warp; grp < ngroups; grp += nwarps * U) {
        double f[U][NTI];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t row = (grp + u * nwarps) * 4 + t;
#pragma unroll
            for (int i = 0; i < NTI; ++i) {
                const int col = 8 * i + g;
                f[u][i] = (row < n_rows && col < D) ? __ldg(X + row * ldx + col) : 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int i = 0; i < NTI; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) dmma884(acc[i][j], f[u][i], f[u][j]);
    }
    // fixed-order reduction over the warps of the CTA (lower tiles only), then one partial per CTA
    for (int e = threadIdx.x; e < DP * DP; e += XTX_THREADS) G[e] = 0.0;
    __syncthreads();
    for (int w = 0; w < XTX_THREADS / 32; ++w) {
        if (warp == w) {
#pragma unroll
            for (int i = 0; i < NTI; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    G[(8 * i + g) * DP + 8 * j + 2 * t] += acc[i][j][0];
                    G[(8 * i + g) * DP + 8 * j + 2 * t + 1] += acc[i][j][1];
                }
        }
        __syncthreads();
    }
    double* out = part + (size_t)blockIdx.x * DP * DP;
    for (int e = threadIdx.x; e < DP * DP; e += XTX_THREADS) out[e] = G[e];
}

constexpr int EIG_THREADS = 1024;   // upper bound; small problems launch fewer
constexpr int EIG_MAX_SWEEPS = 40;

// One CTA.  G and V live in shared memory with pitch D + 1.
__global__ void __launch_bounds__(EIG_THREADS) k_pca_eig(const double* __restrict__ part, int nparts, int DP, int D,
                                                         int R, double* __restrict__ W, double* __restrict__ sing) {
    extern __shared__ double sm[];
    const int ld = D + 1;
    double* G = sm;                  // [D][ld]
    double* V = G + D * ld;          // [D][ld]
    double* cs = V + D * ld;         // [D/2 + 1][2]
    double* red = cs + 2 * (D / 2 + 1);   // [32]
    int* pq = reinterpret_cast<int*>(red + 32);   // [D/2 + 1][2]
    int* rank = pq + 2 * (D / 2 + 1);             // [D]
    const int tid = threadIdx.x, nthr = blockDim.x;

    // G = sum of the partials in CTA order; the Gram kernel filled tiles on and below the tile diagonal
    for (int e = tid; e < D * D; e += nthr) {
        const int i = e / D, j = e % D;
        const int a = i >= j ? i : j, b = i >= j ? j : i;   // (a, b) in the lower triangle
        double s = 0.0;
#pragma unroll 8
        for (int p = 0; p < nparts; ++p) s += __ldg(part + (size_t)p * DP * DP + a * DP + b);
        G[i * ld + j] = s;
        V[i * ld + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();

    const int n = D + (D & 1);       // players of the round-robin; index D (if any) is a bye
    const int npairs = n / 2;
    for (int sweep = 0; sweep < EIG_MAX_SWEEPS; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int e = tid; e < D * D; e += nthr) {
            const int i = e / D, j = e % D;
            const double v = G[i * ld + j];
            if (i == j) dg += v * v; else off += v * v;
        }
        off = block_sum(off, red);
        dg = block_sum(dg, red);
        if (off <= 1e-30 * dg || dg == 0.0) break;
        for (int round = 0; round < n - 1; ++round) {
            if (tid < npairs) {
                if (round + sweep > 0 && pq[2 * tid + 1] < D) {   // the element the previous rotation annihilated
                    G[pq[2 * tid] * ld + pq[2 * tid + 1]] = 0.0;
                    G[pq[2 * tid + 1] * ld + pq[2 * tid]] = 0.0;
                }
                int a, b;
                if (tid == 0) { a = n - 1; b = round; }
                else { a = (round + tid) % (n - 1); b = (round - tid + (n - 1)) % (n - 1); }
                const int p = a < b ? a : b, q = a < b ? b : a;
                double c = 1.0, s = 0.0;
                if (q < D) {
                    const double apq = G[p * ld + q];
                    if (apq != 0.0) {
                        const double tau = (G[q * ld + q] - G[p * ld + p]) / (2.0 * apq);
                        const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + tt * tt);
                        s = tt * c;
                    }
                }
                pq[2 * tid] = p; pq[2 * tid + 1] = q;
                cs[2 * tid] = c; cs[2 * tid + 1] = s;
            }
            __syncthreads();
            // columns: G <- G J, V <- V J
            for (int e = tid; e < npairs * D; e += nthr) {
                const int k = e / D, i = e % D;
                const int p = pq[2 * k], q = pq[2 * k + 1];
                if (q >= D) continue;
                const double c = cs[2 * k], s = cs[2 * k + 1];
                const double gp = G[i * ld + p], gq = G[i * ld + q];
                G[i * ld + p] = c * gp - s * gq;
                G[i * ld + q] = s * gp + c * gq;
                const double vp = V[i * ld + p], vq = V[i * ld + q];
                V[i * ld + p] = c * vp - s * vq;
                V[i * ld + q] = s * vp + c * vq;
            }
            __syncthreads();
            // rows: G <- J^T G
            for (int e = tid; e < npairs * D; e += nthr) {
                const int k = e / D, j = e % D;
                const int p = pq[2 * k], q = pq[2 * k + 1];
                if (q >= D) continue;
                const double c = cs[2 * k], s = cs[2 * k + 1];
                const double gp = G[p * ld + j], gq = G[q * ld + j];
                G[p * ld + j] = c * gp - s * gq;
                G[q * ld + j] = s * gp + c * gq;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    // descending order of the eigenvalues (ties: lower index first), sign convention, output
    for (int i = tid; i < D; i += nthr) {
        const double li = G[i * ld + i];
        int r = 0;
        for (int j = 0; j < D; ++j) {
            const double lj = G[j * ld + j];
            r += (lj > li) || (lj == li && j < i);
        }
        rank[i] = r;
        if (sing) sing[r] = sqrt(li > 0.0 ? li : 0.0);
    }
    __syncthreads();
    for (int i = tid; i < D; i += nthr) {
        const int r = rank[i];
        if (r >= R) continue;
        double best = 0.0, val = 0.0;
        for (int d = 0; d < D; ++d) {
            const double v = V[d * ld + i];
            if (fabs(v) > best) { best = fabs(v); val = v; }
        }
        const double sgn = val > 0.0 ? -1.0 : 1.0;
        for (int d = 0; d < D; ++d) W[d * R + r] = sgn * V[d * ld + i];
    }
}

static int xtx_grid(int64_t n_rows) {
    const int64_t ngroups = (n_rows + 3) / 4;
    const int64_t want = (ngroups + (XTX_THREADS / 32) - 1) / (XTX_THREADS / 32);
    const int64_t cap = 2 * (int64_t)device_sm_count();
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

// padded width of the Gram partials: the k_xtx instantiation (1, 2, 4 or 8 column tiles) that covers D
static int xtx_dp(int D) { return D <= 8 ? 8 : (D <= 16 ? 16 : (D <= 32 ? 32 : 64)); }

}  // namespace dgp

using namespace dgp;

extern "C" size_t dgp_pca_ws_bytes(int64_t n_rows, int32_t D) {
    if (D < 1 || D > DGP_MAX_D || n_rows < 1) return 0;
    const int DP = xtx_dp(D);
    return align256((size_t)xtx_grid(n_rows) * DP * DP * sizeof(double));
}

extern "C" int dgp_pca_weights(const double* X, int64_t n_rows, int32_t ldx, int32_t D, int32_t R, void* ws,
                               size_t ws_bytes, double* W, double* sing, void* stream) {
    if (!X || !ws || !W || n_rows < 1 || D < 1 || D > DGP_MAX_D || R < 1 || R > D || ldx < D) return DGP_ERR_ARG;
    if (ws_bytes < dgp_pca_ws_bytes(n_rows, D)) return DGP_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = xtx_grid(n_rows);
    double* part = static_cast<double*>(ws);
    const int DP = xtx_dp(D);
    if (DP == 8) k_xtx<1><<<grid, XTX_THREADS, 0, st>>>(X, n_rows, ldx, D, part);
    else if (DP == 16) k_xtx<2><<<grid, XTX_THREADS, 0, st>>>(X, n_rows, ldx, D, part);
    else if (DP == 32) k_xtx<4><<<grid, XTX_THREADS, 0, st>>>(X, n_rows, ldx, D, part);
    else k_xtx<8><<<grid, XTX_THREADS, 0, st>>>(X, n_rows, ldx, D, part);
    DGP_COUNT_LAUNCH();
    DGP_CHECK_LAUNCH();
    const size_t smem = (size_t)(2 * D * (D + 1) + 2 * (D / 2 + 1) + 32) * sizeof(double) +
                        (size_t)(2 * (D / 2 + 1) + D) * sizeof(int);
    static bool configured_dev[DGP_MAX_DEVICES] = {};
    bool& configured = configured_dev[current_device()];
    if (!configured) {
        if (cudaFuncSetAttribute(k_pca_eig, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess) return DGP_ERR_LAUNCH;
        configured = true;
    }
    k_pca_eig<<<1, D <= 16 ? 256 : EIG_THREADS, smem, st>>>(part, grid, DP, D, R, W, sing);
    DGP_COUNT_LAUNCH();
    DGP_CHECK_LAUNCH();
    return DGP_OK;
}
