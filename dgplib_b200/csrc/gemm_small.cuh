// Small strided batched fp64 GEMM for the parameter-only stages (O(M^3) work, M <= 1024; <0.01% of an evaluation's
// flops).  C[b](i,j) = alpha * sum_k A[b](i,k) B[b](k,j) + beta * C[b](i,j), arbitrary element strides so
// that transposes are free.  I, J multiples of 64 and K a multiple of 16 (all operands are Mpad-padded).
#pragma once
#include "common.cuh"

namespace dgp {

struct GemmArgs {
    const double* A; const double* B; double* C;
    int I, J, K, batch;
    long long sAi, sAk, sBk, sBj, sCi, sCj, bA, bB, bC;
    double alpha, beta;
    int lower_tiles_only;  // skip 64x64 output tiles strictly above the diagonal
};

int gemm_small(const GemmArgs& g, cudaStream_t st);

#ifdef DGP_GEMM_SMALL_IMPL
__global__ void __launch_bounds__(256) k_gemm_small(GemmArgs g) {
    __shared__ double As[16][65];
    __shared__ double Bs[16][68];
    const int bi = blockIdx.y, bj = blockIdx.x, b = blockIdx.z;
    if (g.lower_tiles_only && bj > bi) return;
    const double* A = g.A + (long long)b * g.bA + (long long)bi * 64 * g.sAi;
    const double* B = g.B + (long long)b * g.bB + (long long)bj * 64 * g.sBj;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.0;
    // the next 16-deep chunk travels global -> registers while the current one is multiplied out of shared memory:
    // these GEMMs sit on the serial parameter chains, where the exposed load latency of every chunk was the whole cost
    double ra[4], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = tid + q * 256;
            int i, k;
            if (g.sAk == 1) { k = e & 15; i = e >> 4; } else { i = e & 63; k = e >> 6; }
            ra[q] = A[(long long)i * g.sAi + (long long)(k0 + k) * g.sAk];
            int j, kb;
            if (g.sBk == 1) { kb = e & 15; j = e >> 4; } else { j = e & 63; kb = e >> 6; }
            rb[q] = B[(long long)(k0 + kb) * g.sBk + (long long)j * g.sBj];
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < g.K; k0 += 16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = tid + q * 256;
            int i, k;
            if (g.sAk == 1) { k = e & 15; i = e >> 4; } else { i = e & 63; k = e >> 6; }
            As[k][i] = ra[q];
            int j, kb;
            if (g.sBk == 1) { kb = e & 15; j = e >> 4; } else { j = e & 63; kb = e >> 6; }
            Bs[kb][j] = rb[q];
        }
        __syncthreads();
        if (k0 + 16 < g.K) fetch(k0 + 16);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            double a[4], c[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q] = As[k][ty * 4 + q];
#pragma unroll
            for (int q = 0; q < 4; ++q) c[q] = Bs[k][tx + 16 * q];   // interleaved columns: conflict-free across tx
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = fma(a[p], c[q], acc[p][q]);
        }
        __syncthreads();
    }
    double* C = g.C + (long long)b * g.bC;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long long off = (long long)(bi * 64 + ty * 4 + p) * g.sCi + (long long)(bj * 64 + tx + 16 * q) * g.sCj;
            double v = g.alpha * acc[p][q];
            if (g.beta != 0.0) v = fma(g.beta, C[off], v);
            C[off] = v;
        }
}

// 32 x 32 output tiles (2 x 2 per thread) for launches that would otherwise be a handful of CTAs: the single-matrix
// products of the Cholesky reverse mode sit on a serial chain, where the latency of one CTA's K loop is the whole cost.
__global__ void __launch_bounds__(256) k_gemm_small32(GemmArgs g) {
    __shared__ double As[16][33];
    __shared__ double Bs[16][33];
    const int bi = blockIdx.y, bj = blockIdx.x, b = blockIdx.z;
    if (g.lower_tiles_only && (bj >> 1) > (bi >> 1)) return;      // same 64 x 64 granularity as the large-tile kernel
    const double* A = g.A + (long long)b * g.bA + (long long)bi * 32 * g.sAi;
    const double* B = g.B + (long long)b * g.bB + (long long)bj * 32 * g.sBj;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    double ra[2], rb[2];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int e = tid + q * 256;
            int i, k;
            if (g.sAk == 1) { k = e & 15; i = e >> 4; } else { i = e & 31; k = e >> 5; }
            ra[q] = A[(long long)i * g.sAi + (long long)(k0 + k) * g.sAk];
            int j, kb;
            if (g.sBk == 1) { kb = e & 15; j = e >> 4; } else { j = e & 31; kb = e >> 5; }
            rb[q] = B[(long long)(k0 + kb) * g.sBk + (long long)j * g.sBj];
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < g.K; k0 += 16) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int e = tid + q * 256;
            int i, k;
            if (g.sAk == 1) { k = e & 15; i = e >> 4; } else { i = e & 31; k = e >> 5; }
            As[k][i] = ra[q];
            int j, kb;
            if (g.sBk == 1) { kb = e & 15; j = e >> 4; } else { j = e & 31; kb = e >> 5; }
            Bs[kb][j] = rb[q];
        }
        __syncthreads();
        if (k0 + 16 < g.K) fetch(k0 + 16);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const double a0 = As[k][ty], a1 = As[k][ty + 16];
            const double c0 = Bs[k][tx], c1 = Bs[k][tx + 16];
            acc[0][0] = fma(a0, c0, acc[0][0]); acc[0][1] = fma(a0, c1, acc[0][1]);
            acc[1][0] = fma(a1, c0, acc[1][0]); acc[1][1] = fma(a1, c1, acc[1][1]);
        }
        __syncthreads();
    }
    double* C = g.C + (long long)b * g.bC;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const long long off = (long long)(bi * 32 + ty + 16 * p) * g.sCi + (long long)(bj * 32 + tx + 16 * q) * g.sCj;
            double v = g.alpha * acc[p][q];
            if (g.beta != 0.0) v = fma(g.beta, C[off], v);
            C[off] = v;
        }
}

int gemm_small(const GemmArgs& g, cudaStream_t st) {
    if (g.I % 64 || g.J % 64 || g.K % 16 || g.batch < 1) return DGP_ERR_ARG;
    const long long ctas64 = (long long)(g.I / 64) * (g.J / 64) * g.batch;
    if (ctas64 < 64) k_gemm_small32<<<dim3(g.J / 32, g.I / 32, g.batch), 256, 0, st>>>(g);
    else k_gemm_small<<<dim3(g.J / 64, g.I / 64, g.batch), 256, 0, st>>>(g);
    DGP_COUNT_LAUNCH();
    return cudaGetLastError() == cudaSuccess ? DGP_OK : DGP_ERR_LAUNCH;
}
#endif

}  // namespace dgp
