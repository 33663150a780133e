// Parameter-only stage of one whitened conditional (runs once per ELBO evaluation, independent of the data rows):
//   Kuu = k(Z,Z) + jitter I  ->  Lm = chol(Kuu)  ->  Linv = Lm^-1,   LqT_r = tril(q_sqrt_r)^T,   KL(q||p) whitened.
// Replaces features.Kuu / tf.cholesky inside gpflow's base_conditional (called at dgplib/layers.py:66-69) and
// gauss_kl(q_mu, q_sqrt, K=None) (dgplib/layers.py:58-60).  SURVEY.md Appendix A.
//
// Design note: the reference does A = triangular_solve(Lm, Kuf) per evaluation.  Here the solve is turned into a
// tensor-core GEMM by forming Lm^-1 once (O(M^3), parameter-only) so that the per-row stage is A = Linv * Kuf with
// block skipping; measured against 50-digit arithmetic the explicit inverse stays within 3x of LAPACK's substitution
// error on the worst-conditioned reference configuration (D=1, M=50).
#include <atomic>
#include <cstdlib>

#include "common.cuh"
#define DGP_GEMM_SMALL_IMPL
#include "gemm_small.cuh"
#include "stream_gemm.cuh"

namespace dgp {

std::atomic<long long> g_launches{0};
#ifdef DGP_TEST_HOOKS
int g_force_fwd_nt = 0, g_force_bwd_nt = 0, g_force_gram_splits = 0, g_debug_mode = 0, g_force_stages = 0;
#endif

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    return dev < 0 ? 0 : (dev >= DGP_MAX_DEVICES ? DGP_MAX_DEVICES - 1 : dev);
}

int device_sm_count() {
    static std::atomic<int> n[DGP_MAX_DEVICES];
    const int dev = current_device();
    int v = n[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) {
            cudaGetLastError();
            v = 148;  // B200; also used when cross-checking layouts on a GPU-less host
        }
        n[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

int check_desc(const dgp_cond_desc* d) {
    if (!d) return DGP_ERR_ARG;
    if (d->M < 1 || d->M > 1024) return DGP_ERR_UNSUPPORTED;
    if (d->D < 1 || d->D > DGP_MAX_D || d->Dx < d->D) return DGP_ERR_ARG;
    if (d->R < 1 || d->R > DGP_MAX_R || d->ldr < d->R || d->col0 < 0 || d->col0 + d->R > d->ldr) return DGP_ERR_ARG;
    if (d->T < 1 || d->T > DGP_MAX_TASKS) return DGP_ERR_ARG;
    if (d->T > 1 && d->Dx < d->D + 1) return DGP_ERR_ARG;
    for (int t = 0; t < d->T; ++t)
        if (d->kind[t] != DGP_KIND_RBF && d->kind[t] != DGP_KIND_MATERN52) return DGP_ERR_UNSUPPORTED;
    if (!d->Z || !d->theta || !d->q_mu || !d->q_sqrt) return DGP_ERR_ARG;
    return DGP_OK;
}

// ------------------------------------------------------------------------------------------------ scaled inducing inputs
__global__ void k_prep_z(dgp_cond_desc d, int Mpad, int Dp, double* __restrict__ Zs, double* __restrict__ zn,
                         int32_t* __restrict__ ztask) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= Mpad) return;
    if (m >= d.M) {
        for (int k = 0; k < Dp; ++k) Zs[(size_t)m * Dp + k] = 0.0;
        zn[m] = 0.0;
        ztask[m] = -1;
        return;
    }
    int t = 0;
    if (d.T > 1) {
        t = (int)d.Z[(size_t)m * d.Dx + d.Dx - 1];
        if (t < 0 || t >= d.T) t = -1;  // out-of-range id: matches no kernel (row of zeros, like an empty partition)
    }
    ztask[m] = t;
    const double* th = d.theta + (size_t)(t < 0 ? 0 : t) * (2 + d.D);
    double s = 0.0;
    for (int k = 0; k < Dp; ++k) {
        double v = 0.0;
        if (k < d.D) v = d.Z[(size_t)m * d.Dx + k] / th[2 + k];
        Zs[(size_t)m * Dp + k] = v;
        s = fma(v, v, s);
    }
    zn[m] = s;
}

// Kuu(+jitter) into the Lm buffer, padded with the identity.  r2 = |zi|^2 + |zj|^2 - 2 zi.zj on scaled inputs with no
// clamp, as gpflow's Stationary.scaled_square_dist does.  White enters the symmetric call only (T == 1); inside a
// SwitchedKernel the sub-kernels are always called with an explicit X2 (dgplib/specialized_kernels.py:47).
__global__ void k_kuu(dgp_cond_desc d, int Mpad, int Dp, const double* __restrict__ Zs, const double* __restrict__ zn,
                      const int32_t* __restrict__ ztask, double* __restrict__ K) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= Mpad) return;
    double v;
    if (i >= d.M || j >= d.M) {
        v = (i == j) ? 1.0 : 0.0;
    } else {
        const int ti = ztask[i], tj = ztask[j];
        v = 0.0;
        if (ti == tj && ti >= 0) {
            const double* th = d.theta + (size_t)ti * (2 + d.D);
            double c = 0.0;
            for (int k = 0; k < Dp; ++k) c = fma(Zs[(size_t)i * Dp + k], Zs[(size_t)j * Dp + k], c);
            const double r2 = zn[i] + zn[j] - 2.0 * c;
            v = th[0] * kern_val(d.kind[ti], r2);
        }
        if (i == j) {
            v += d.jitter;
            if (d.T == 1) v += d.theta[1];
        }
    }
    K[(size_t)i * Mpad + j] = v;
}

// ------------------------------------------------------------------------------------------------ Cholesky (one CTA)
// Right-looking blocked Cholesky of the Mpad x Mpad matrix in global memory (L2-resident), lower factor in place,
// strict upper triangle zeroed.  NB = 32 (Mpad <= 512) or 16.  512 threads.
constexpr int CHOL_THREADS = 512;

__global__ void __launch_bounds__(CHOL_THREADS, 1)
k_cholesky(double* __restrict__ K, int Mpad, int NB, int32_t* __restrict__ status) {
    K += (size_t)blockIdx.x * Mpad * Mpad;      // batched: one CTA per matrix
    status += blockIdx.x;
    extern __shared__ double sm[];
    double* sD = sm;                 // [NB][33]
    double* sDinv = sD + 32 * 33;    // [32]
    double* sP = sDinv + 32;         // [(Mpad-NB)][NB+1]
    __shared__ int s_fail;
    const int tid = threadIdx.x;
    const int ldp = NB + 1;
    if (tid == 0) s_fail = 0;
    __syncthreads();

    for (int j0 = 0; j0 < Mpad; j0 += NB) {
        // 1. diagonal block -> smem
        for (int e = tid; e < NB * NB; e += CHOL_THREADS) {
            const int i = e / NB, j = e % NB;
            sD[i * 33 + j] = K[(size_t)(j0 + i) * Mpad + j0 + j];
        }
        __syncthreads();
        // 2. unblocked factorisation of the NB x NB block by the whole CTA
        for (int k = 0; k < NB; ++k) {
            if (tid == 0) {
                double p = sD[k * 33 + k];
                if (!(p > 0.0)) {
                    if (s_fail == 0) s_fail = j0 + k + 1;
                    p = 1.0;
                }
                const double dk = sqrt(p);
                sD[k * 33 + k] = dk;
                sDinv[k] = 1.0 / dk;
            }
            __syncthreads();
            if (tid > k && tid < NB) sD[tid * 33 + k] *= sDinv[k];
            __syncthreads();
            for (int e = tid; e < NB * NB; e += CHOL_THREADS) {
                const int i = e / NB, j = e % NB;
                if (j > k && j <= i) sD[i * 33 + j] = fma(-sD[i * 33 + k], sD[j * 33 + k], sD[i * 33 + j]);
            }
            __syncthreads();
        }
        // write the factored block back (upper part zero)
        for (int e = tid; e < NB * NB; e += CHOL_THREADS) {
            const int i = e / NB, j = e % NB;
            K[(size_t)(j0 + i) * Mpad + j0 + j] = (j <= i) ? sD[i * 33 + j] : 0.0;
        }
        const int rem = Mpad - j0 - NB;
        // 3. panel: rows below the block, one thread per row: x = a * Ljj^-T by forward substitution
        for (int ri = tid; ri < rem; ri += CHOL_THREADS) {
            const int i = j0 + NB + ri;
            double x[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                if (k < NB) {
                    double s = K[(size_t)i * Mpad + j0 + k];
#pragma unroll
                    for (int p = 0; p < k; ++p) s = fma(-x[p], sD[k * 33 + p], s);
                    x[k] = s * sDinv[k];
                }
            }
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                if (k < NB) {
                    K[(size_t)i * Mpad + j0 + k] = x[k];
                    sP[ri * ldp + k] = x[k];
                }
            }
        }
        // zero the strict upper part right of the diagonal block
        for (int e = tid; e < NB * rem; e += CHOL_THREADS) {
            const int i = e / rem, j = e % rem;
            K[(size_t)(j0 + i) * Mpad + j0 + NB + j] = 0.0;
        }
        __syncthreads();
        // 4. trailing update (lower part only), 4x4 micro-tiles over the triangular tile set
        const int nt = rem / 4;
        const int ntile = nt * (nt + 1) / 2;
        for (int p = tid; p < ntile; p += CHOL_THREADS) {
            int ti = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
            while ((ti + 1) * (ti + 2) / 2 <= p) ++ti;
            while (ti * (ti + 1) / 2 > p) --ti;
            const int tj = p - ti * (ti + 1) / 2;
            double acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
            for (int k = 0; k < NB; ++k) {
                double pa[4], pb[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) pa[a] = sP[(ti * 4 + a) * ldp + k];
#pragma unroll
                for (int b = 0; b < 4; ++b) pb[b] = sP[(tj * 4 + b) * ldp + k];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = fma(pa[a], pb[b], acc[a][b]);
            }
            const int i0 = j0 + NB + ti * 4, c0 = j0 + NB + tj * 4;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (c0 + b <= i0 + a) K[(size_t)(i0 + a) * Mpad + c0 + b] -= acc[a][b];
        }
        __syncthreads();
    }
    if (tid == 0) status[0] = s_fail;
}

// Same factorisation for Mpad <= 512 with the three stages of a 32-column panel laid out for latency, the quantity that
// matters here (the factorisation is a serial chain at the head of every evaluation):
//   1. warp 0 factors the 32 x 32 diagonal block in registers (lane = row, column broadcasts by shuffle) with no CTA
//      barrier inside, while the other 480 threads already fetch their panel rows;
//   2. one thread per row below the block solves x L_jj^T = a by substitution (L_jj broadcast from shared memory);
//   3. the trailing update runs on DMMA.8x8x4, 8 x 8 tiles of the lower triangle dealt to the 16 warps.
// Inverse of the factored 32 x 32 diagonal block in sD (pitch 33, reciprocal pivots in sDinv) by one warp: lane = column
// of the inverse, held in registers; forward substitution with the rows of the block broadcast from shared memory.
__device__ __noinline__ void diag_block_inverse(const double* sD, const double* sDinv, double* __restrict__ out,
                                                   int ld, int lane) {
    double x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {              // column-oriented: independent updates once x[k] is final
        x[k] *= sDinv[k];                       // rows above the lane's column stay exactly zero
        out[(size_t)k * ld + lane] = x[k];
#pragma unroll
        for (int i = k + 1; i < 32; ++i) x[i] = fma(-sD[i * 33 + k], x[k], x[i]);
    }
}

constexpr int CHOL_LDP = 36;   // panel pitch in shared memory: 36 mod 16 = 4 keeps the DMMA fragment loads conflict-free

__global__ void __launch_bounds__(CHOL_THREADS, 1)
k_cholesky32(double* __restrict__ K, int Mpad, int32_t* __restrict__ status, double* __restrict__ Linv) {
    K += (size_t)blockIdx.x * Mpad * Mpad;      // batched: one CTA per matrix
    status += blockIdx.x;
    if (Linv) Linv += (size_t)blockIdx.x * Mpad * Mpad;
    extern __shared__ double sm[];
    double* sD = sm;                 // [32][33] factored diagonal block
    double* sDinv = sD + 32 * 33;    // [32] reciprocal pivots
    double* sP = sDinv + 32;         // [Mpad - 32][CHOL_LDP] panel below the block
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    int fail = 0;                    // warp 0 only

    for (int j0 = 0; j0 < Mpad; j0 += 32) {
        const int rem = Mpad - j0 - 32;          // rows below the diagonal block (<= 480)
        const int ri = tid - 32;                 // panel row of this thread (warps 1..15)
        double x[32];
        if (warp == 0) {
            const double* src = K + (size_t)(j0 + lane) * Mpad + j0;
#pragma unroll
            for (int k = 0; k < 32; ++k) x[k] = src[k];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                double dj = __shfl_sync(0xffffffffu, x[j], j);
                if (!(dj > 0.0)) {
                    if (fail == 0) fail = j0 + j + 1;
                    dj = 1.0;
                }
                // reciprocal square root (<= 1 ulp) instead of sqrt followed by a division: this chain of 32 dependent
                // pivots is the critical path of the whole evaluation's parameter stage
                const double inv = rsqrt(dj), dk = dj * inv;
                x[j] = (lane == j) ? dk : (lane > j ? x[j] * inv : 0.0);
                if (lane == j) sDinv[j] = inv;
#pragma unroll
                for (int k = j + 1; k < 32; ++k) {
                    const double lkj = __shfl_sync(0xffffffffu, x[j], k);
                    x[k] = fma(-x[j], lkj, x[k]);     // rows above k carry don't-care values until column k zeroes them
                }
            }
            double* dst = K + (size_t)(j0 + lane) * Mpad + j0;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                sD[lane * 33 + k] = x[k];
                dst[k] = x[k];
            }
        } else if (ri < rem) {
            const double* src = K + (size_t)(j0 + 32 + ri) * Mpad + j0;
#pragma unroll
            for (int k = 0; k < 32; ++k) x[k] = src[k];
        }
        if (rem == 0) {
            if (warp == 0 && Linv) {
                __syncwarp();
                diag_block_inverse(sD, sDinv, Linv + (size_t)j0 * Mpad + j0, Mpad, lane);
            }
            break;
        }
        // zero the strict upper part right of the diagonal block
        for (int e = tid; e < 32 * rem; e += CHOL_THREADS) {
            const int i = e / rem, j = e % rem;
            K[(size_t)(j0 + i) * Mpad + j0 + 32 + j] = 0.0;
        }
        __syncthreads();
        if (warp == 0 && Linv) diag_block_inverse(sD, sDinv, Linv + (size_t)j0 * Mpad + j0, Mpad, lane);
        if (warp > 0 && ri < rem) {
            // column-oriented substitution: once x[k] is final, the updates of x[k+1 ..] are independent FMAs (the
            // row-oriented form is one dependent chain per entry and ran 3x slower)
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                x[k] *= sDinv[k];
#pragma unroll
                for (int j = k + 1; j < 32; ++j) x[j] = fma(-x[k], sD[j * 33 + k], x[j]);
            }
            double* dst = K + (size_t)(j0 + 32 + ri) * Mpad + j0;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                dst[k] = x[k];
                sP[ri * CHOL_LDP + k] = x[k];
            }
        }
        __syncthreads();
        // trailing update of the lower triangle: C(ti, tj) -= P(ti) P(tj)^T on 8 x 8 tiles
        const int nt = rem / 8;
        const int ntile = nt * (nt + 1) / 2;
        // four tiles per warp and pass, so that four read-modify-write round trips to L2 are in flight at once
        constexpr int TU = 4;
        for (int p0 = warp * TU; p0 < ntile; p0 += (CHOL_THREADS / 32) * TU) {
            double* cptr[TU];
            double2 c0[TU];
            const double *pa[TU], *pb[TU];
            bool lo[TU], hi[TU];
#pragma unroll
            for (int u = 0; u < TU; ++u) {
                const int p = (p0 + u < ntile) ? p0 + u : ntile - 1;   // the clamped duplicates are not stored
                int ti = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
                while ((ti + 1) * (ti + 2) / 2 <= p) ++ti;
                while (ti * (ti + 1) / 2 > p) --ti;
                const int tj = p - ti * (ti + 1) / 2;
                const int row = j0 + 32 + ti * 8 + g, col = j0 + 32 + tj * 8 + 2 * t;
                cptr[u] = K + (size_t)row * Mpad + col;
                c0[u] = *reinterpret_cast<const double2*>(cptr[u]);
                pa[u] = sP + (ti * 8 + g) * CHOL_LDP + t;
                pb[u] = sP + (tj * 8 + g) * CHOL_LDP + t;
                lo[u] = (p0 + u < ntile) && col <= row;
                hi[u] = (p0 + u < ntile) && col + 1 <= row;
            }
            double c[TU][2];
#pragma unroll
            for (int u = 0; u < TU; ++u) c[u][0] = c[u][1] = 0.0;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
#pragma unroll
                for (int u = 0; u < TU; ++u) dmma884(c[u], pa[u][kk * 4], pb[u][kk * 4]);
#pragma unroll
            for (int u = 0; u < TU; ++u) {
                if (lo[u]) cptr[u][0] = c0[u].x - c[u][0];
                if (hi[u]) cptr[u][1] = c0[u].y - c[u][1];
            }
        }
        __syncthreads();
    }
    if (tid == 0) status[0] = fail;
}

// ------------------------------------------------------------------------------------------------ triangular inverse
// Linv = Lm^-1 by block forward substitution on 32 x 32 blocks.  The inverses of the diagonal blocks come out of the
// Cholesky kernel (k_diag_inv does the same for the legacy path); k_tri_inv_cols fills the blocks below them, one CTA
// per 32-wide column block c:   X_Ic = -Linv_II * sum_{K=c}^{I-1} L_IK X_Kc,   I = c+1 .. nb-1,
// both products on DMMA.8x8x4 with the fragments of the long one read straight from global memory (L1 / L2 resident).
// The blocks above the diagonal are zeroed here (the workspace arrives uninitialised).
__global__ void __launch_bounds__(256)
k_tri_inv_cols(const double* __restrict__ L, double* __restrict__ Linv, int Mpad) {
    __shared__ double sT[32 * 36];
    __shared__ double sDi[32 * 36];
    const int c = blockIdx.x, nb = Mpad / 32, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int c0 = c * 32;
    const int ta = warp >> 1, tb = (warp & 1) * 2;       // this warp's tiles: (ta, tb) and (ta, tb + 1) of the 4 x 4 grid
    for (int e = tid; e < c0 * 32; e += 256) Linv[(size_t)(e >> 5) * Mpad + c0 + (e & 31)] = 0.0;
    for (int I = c + 1; I < nb; ++I) {
        const int i0 = I * 32;
        for (int e = tid; e < 32 * 32; e += 256)
            sDi[(e >> 5) * 36 + (e & 31)] = Linv[(size_t)(i0 + (e >> 5)) * Mpad + i0 + (e & 31)];
        // T = L[I rows][c0 .. i0) * X[c0 .. i0)[c block]
        double acc0[2] = {0.0, 0.0}, acc1[2] = {0.0, 0.0};
        const double* pa = L + (size_t)(i0 + ta * 8 + g) * Mpad + t;
        const double* pb = Linv + (size_t)t * Mpad + c0 + tb * 8 + g;
        for (int kb = c0; kb < i0; kb += 32) {      // one 32-deep block per pass: 24 loads in flight, then 16 DMMAs
            double a[8], b0[8], b1[8];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const int k = kb + 4 * kk;
                a[kk] = pa[k];
                b0[kk] = pb[(size_t)k * Mpad];
                b1[kk] = pb[(size_t)k * Mpad + 8];
            }
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                dmma884(acc0, a[kk], b0[kk]);
                dmma884(acc1, a[kk], b1[kk]);
            }
        }
        sT[(ta * 8 + g) * 36 + tb * 8 + 2 * t] = acc0[0];
        sT[(ta * 8 + g) * 36 + tb * 8 + 2 * t + 1] = acc0[1];
        sT[(ta * 8 + g) * 36 + tb * 8 + 8 + 2 * t] = acc1[0];
        sT[(ta * 8 + g) * 36 + tb * 8 + 8 + 2 * t + 1] = acc1[1];
        __syncthreads();
        // X_Ic = -Linv_II * T
        double o0[2] = {0.0, 0.0}, o1[2] = {0.0, 0.0};
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const double a = sDi[(ta * 8 + g) * 36 + kk * 4 + t];
            dmma884(o0, a, sT[(kk * 4 + t) * 36 + tb * 8 + g]);
            dmma884(o1, a, sT[(kk * 4 + t) * 36 + tb * 8 + 8 + g]);
        }
        double* dst = Linv + (size_t)(i0 + ta * 8 + g) * Mpad + c0 + tb * 8 + 2 * t;
        dst[0] = -o0[0]; dst[1] = -o0[1];
        dst[8] = -o1[0]; dst[9] = -o1[1];
        __syncthreads();  // X_Ic visible to the whole CTA before the next block row reads it (global, same CTA)
    }
}

// Same substitution for Mpad <= 256 with everything the long product touches staged in shared memory: the column panel
// X_Kc stays resident (this CTA produces it), the block row of L and Linv_II arrive by cp.async in one batch per step, so
// a step costs one L2 round trip instead of one per 32-deep slice (53 -> ~16 us at M = 256 on the serial parameter chain).
constexpr int TRI_LDX = 36;     // pitch of the 32-wide panels: 36 mod 16 = 4 keeps the DMMA fragment loads conflict-free
__global__ void __launch_bounds__(256)
k_tri_inv_cols_smem(const double* __restrict__ L, double* __restrict__ Linv, int Mpad) {
    extern __shared__ __align__(16) double tsm[];
    const int ldl = Mpad + 4;
    double* sL = tsm;                       // [32][ldl]      block row I of L, columns c0 .. i0
    double* sX = sL + 32 * ldl;             // [Mpad][TRI_LDX] rows k - c0 of the column panel X[., c block]
    double* sDi = sX + Mpad * TRI_LDX;      // [32][TRI_LDX]  Linv_II
    double* sT = sDi + 32 * TRI_LDX;        // [32][TRI_LDX]  T = L_I. X
    const int c = blockIdx.x, nb = Mpad / 32, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int c0 = c * 32;
    const int ta = warp >> 1, tb = (warp & 1) * 2;
    for (int e = tid; e < c0 * 32; e += 256) Linv[(size_t)(e >> 5) * Mpad + c0 + (e & 31)] = 0.0;
    for (int e = tid; e < 32 * 32; e += 256)   // X_cc: the diagonal block's inverse (written by the Cholesky kernel)
        sX[(e >> 5) * TRI_LDX + (e & 31)] = Linv[(size_t)(c0 + (e >> 5)) * Mpad + c0 + (e & 31)];
    for (int I = c + 1; I < nb; ++I) {
        const int i0 = I * 32, kw = i0 - c0;             // kw = depth of the long product
        // stage L[i0 .. i0+32)[c0 .. i0) and Linv_II: 16-byte cp.async, all in flight at once
        const int per_row = kw / 2;                      // 16-byte pieces per row
        for (int e = tid; e < 32 * per_row; e += 256) {
            const int r = e / per_row, q = e % per_row;
            cp_async16(sL + r * ldl + 2 * q, L + (size_t)(i0 + r) * Mpad + c0 + 2 * q);
        }
        for (int e = tid; e < 32 * 16; e += 256) {
            const int r = e >> 4, q = e & 15;
            cp_async16(sDi + r * TRI_LDX + 2 * q, Linv + (size_t)(i0 + r) * Mpad + i0 + 2 * q);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        double acc0[2] = {0.0, 0.0}, acc1[2] = {0.0, 0.0};
        const double* pa = sL + (ta * 8 + g) * ldl + t;
        const double* pb = sX + t * TRI_LDX + tb * 8 + g;
#pragma unroll 4
        for (int k = 0; k < kw; k += 4) {
            const double a = pa[k];
            dmma884(acc0, a, pb[k * TRI_LDX]);
            dmma884(acc1, a, pb[k * TRI_LDX + 8]);
        }
        sT[(ta * 8 + g) * TRI_LDX + tb * 8 + 2 * t] = acc0[0];
        sT[(ta * 8 + g) * TRI_LDX + tb * 8 + 2 * t + 1] = acc0[1];
        sT[(ta * 8 + g) * TRI_LDX + tb * 8 + 8 + 2 * t] = acc1[0];
        sT[(ta * 8 + g) * TRI_LDX + tb * 8 + 8 + 2 * t + 1] = acc1[1];
        __syncthreads();
        // X_Ic = -Linv_II * T  -> global and the resident panel
        double o0[2] = {0.0, 0.0}, o1[2] = {0.0, 0.0};
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const double a = sDi[(ta * 8 + g) * TRI_LDX + kk * 4 + t];
            dmma884(o0, a, sT[(kk * 4 + t) * TRI_LDX + tb * 8 + g]);
            dmma884(o1, a, sT[(kk * 4 + t) * TRI_LDX + tb * 8 + 8 + g]);
        }
        double* dst = Linv + (size_t)(i0 + ta * 8 + g) * Mpad + c0 + tb * 8 + 2 * t;
        dst[0] = -o0[0]; dst[1] = -o0[1];
        dst[8] = -o1[0]; dst[9] = -o1[1];
        double* xs = sX + (kw + ta * 8 + g) * TRI_LDX + tb * 8 + 2 * t;
        xs[0] = -o0[0]; xs[1] = -o0[1];
        xs[8] = -o1[0]; xs[9] = -o1[1];
        __syncthreads();   // panel rows of block I and the reuse of sL / sDi / sT by the next step
    }
}

static size_t tri_inv_smem_bytes(int Mpad) {
    return (size_t)(32 * (Mpad + 4) + Mpad * TRI_LDX + 2 * 32 * TRI_LDX) * sizeof(double);
}

int tri_inv_cols(const double* L, double* Linv, int Mpad, cudaStream_t st) {
    if (Mpad <= 256) {
        static DevOnceSize attr;
        const int dev = current_device();
        if (!attr.covers(dev, 1)) {
            if (cudaFuncSetAttribute(k_tri_inv_cols_smem, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)tri_inv_smem_bytes(256)) != cudaSuccess) return DGP_ERR_LAUNCH;
            attr.raise(dev, 1);
        }
        k_tri_inv_cols_smem<<<Mpad / 32, 256, tri_inv_smem_bytes(Mpad), st>>>(L, Linv, Mpad);
    } else {
        k_tri_inv_cols<<<Mpad / 32, 256, 0, st>>>(L, Linv, Mpad);
    }
    DGP_COUNT_LAUNCH();
    return cudaGetLastError() == cudaSuccess ? DGP_OK : DGP_ERR_LAUNCH;
}

// Legacy path (Mpad > 512): inverses of the 32 x 32 diagonal blocks, one warp per block.
__global__ void k_diag_inv(const double* __restrict__ L, double* __restrict__ Linv, int Mpad) {
    __shared__ double sL[32 * 33];
    __shared__ double sInv[32];
    L += (size_t)blockIdx.y * Mpad * Mpad;
    Linv += (size_t)blockIdx.y * Mpad * Mpad;
    const int b0 = blockIdx.x * 32, lane = threadIdx.x;
    for (int i = 0; i < 32; ++i) sL[i * 33 + lane] = L[(size_t)(b0 + i) * Mpad + b0 + lane];
    __syncwarp();
    sInv[lane] = 1.0 / sL[lane * 33 + lane];
    __syncwarp();
    diag_block_inverse(sL, sInv, Linv + (size_t)b0 * Mpad + b0, Mpad, lane);
}

// strict upper triangle of an Mpad x Mpad matrix <- 0 (the blocked factorisation leaves the input there)
__global__ void k_zero_upper(double* __restrict__ K, int Mpad) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j < Mpad && j > i) K[(size_t)i * Mpad + j] = 0.0;
}

int cholesky_batched(double* K, int Mpad, int batch, int32_t* status, cudaStream_t st, double* Linv) {
    static DevOnceSize attr_set;
    const int dev = current_device();
    if (!attr_set.covers(dev, 1)) {
        if (cudaFuncSetAttribute(k_cholesky, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return DGP_ERR_LAUNCH;
        if (cudaFuncSetAttribute(k_cholesky32, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return DGP_ERR_LAUNCH;
        attr_set.raise(dev, 1);
    }
    if (Mpad <= 512) {
        const size_t smem = (32 * 33 + 32 + (size_t)(Mpad - 32) * CHOL_LDP) * sizeof(double);
        k_cholesky32<<<batch, CHOL_THREADS, smem, st>>>(K, Mpad, status, Linv);
    } else {
        const int NB = (Mpad <= 512) ? 32 : 16;
        const size_t chol_smem = (32 * 33 + 32 + (size_t)(Mpad - NB) * (NB + 1)) * sizeof(double);
        k_cholesky<<<batch, CHOL_THREADS, chol_smem, st>>>(K, Mpad, NB, status);
        if (Linv) { k_diag_inv<<<dim3(Mpad / 32, batch), 32, 0, st>>>(K, Linv, Mpad); DGP_COUNT_LAUNCH(); }
    }
    DGP_COUNT_LAUNCH();
    return cudaGetLastError() == cudaSuccess ? DGP_OK : DGP_ERR_LAUNCH;
}

// ------------------------------------------------------------------------------------------------ q_sqrt, KL
// LqT[r][i][j] = tril(q_sqrt[col0+r])[j][i]  (upper triangular), zero padded.  The same pass accumulates this tile's
// share of the whitened KL term: sum over the lower triangle of q^2, minus log(q_ii^2) on the diagonal.
// Lq (nullable, backward pass only) receives the untransposed lower factor [r][j][i], zero padded.
__global__ void k_lq_prep(dgp_cond_desc d, int Mpad, double* __restrict__ LqT, double* __restrict__ Lq,
                          double* __restrict__ klpart) {
    __shared__ double tile[32][33];
    __shared__ double red[32];
    const int r = blockIdx.z;
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;  // output tile rows i0.., cols j0..
    const double* q = d.q_sqrt + (size_t)(d.col0 + r) * d.M * d.M;
    double s = 0.0;
    // read q[j][i] for j in j0.., i in i0.. (coalesced along i)
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
        const int j = j0 + jj, i = i0 + threadIdx.x;
        double v = 0.0;
        if (j < d.M && i < d.M && i <= j) {
            v = q[(size_t)j * d.M + i];
            s = fma(v, v, s);
            if (i == j) s -= log(v * v);
        }
        tile[jj][threadIdx.x] = v;
        if (Lq) Lq[((size_t)r * Mpad + j) * Mpad + i] = v;
    }
    __syncthreads();
    for (int ii = threadIdx.y; ii < 32; ii += blockDim.y)
        LqT[((size_t)r * Mpad + i0 + ii) * Mpad + j0 + threadIdx.x] = tile[threadIdx.x][ii];
    // deterministic block sum (thread id = y * 32 + x)
    s = warp_sum(s);
    if (threadIdx.x == 0) red[threadIdx.y] = s;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        double tot = 0.0;
        for (int w = 0; w < (int)blockDim.y; ++w) tot += red[w];
        klpart[((size_t)r * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = tot;
    }
}

// KL = 0.5 (sum q_mu^2 - M R - sum log diag(Lq)^2 + sum Lq^2) for this conditional's columns: sums the per-tile partials
// of k_lq_prep and the q_mu term (one CTA, fixed order)
__global__ void __launch_bounds__(256) k_kl(dgp_cond_desc d, const double* __restrict__ klpart, int npartial,
                                            double* __restrict__ kl_out) {
    __shared__ double red[32];
    const int M = d.M, R = d.R;
    double s = 0.0;
    for (int e = threadIdx.x; e < npartial; e += blockDim.x) s += klpart[e];
    for (int e = threadIdx.x; e < M * R; e += blockDim.x) {
        const double v = d.q_mu[(size_t)(e / R) * d.ldr + d.col0 + (e % R)];
        s = fma(v, v, s);
    }
    const double tot = block_sum(s, red);
    if (threadIdx.x == 0) kl_out[0] = 0.5 * (tot - (double)M * R);
}

__global__ void k_transpose(const double* __restrict__ in, double* __restrict__ out, int n) {
    __shared__ double tile[32][33];
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) tile[ii][threadIdx.x] = in[(size_t)(i0 + ii) * n + j0 + threadIdx.x];
    __syncthreads();
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) out[(size_t)(j0 + jj) * n + i0 + threadIdx.x] = tile[threadIdx.x][jj];
}

// row-major [batch][Mpad][Mpad] -> tiled [batch][block row][chunk][128][16], 32-byte groups of a row permuted by
// (row & 3) (common.cuh tile_swizzle); rows / columns beyond Mpad are zero
__global__ void k_retile(const double* __restrict__ src, double* __restrict__ dst, int Mpad, size_t tl) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= tl) return;
    const int row = (int)((e / TILE_WLD) % TILE_MB);
    const int c = tile_swizzle(row, (int)(e % TILE_WLD));    // the permutation is an involution: stored slot -> column
    const size_t tile = e / TILE_DOUBLES;
    const int nch = tiled_mpad(Mpad) / TILE_KC;
    const int kc = (int)(tile % nch), bi = (int)(tile / nch);
    const int i = bi * TILE_MB + row, k = kc * TILE_KC + c;
    double v = 0.0;
    if (i < Mpad && k < Mpad) v = src[(size_t)blockIdx.y * Mpad * Mpad + (size_t)i * Mpad + k];
    dst[(size_t)blockIdx.y * tl + e] = v;
}
int retile(const double* src, double* dst, int Mpad, int batch, cudaStream_t st) {
    const size_t tl = tiled_doubles(Mpad);
    k_retile<<<dim3((unsigned)((tl + 255) / 256), batch), 256, 0, st>>>(src, dst, Mpad, tl);
    DGP_COUNT_LAUNCH();
    return cudaGetLastError() == cudaSuccess ? DGP_OK : DGP_ERR_LAUNCH;
}

}  // namespace dgp

using namespace dgp;

// ------------------------------------------------------------------------------------------------ helper streams
// dgp_cond_prepare and dgp_cond_backward_params fork an independent chain onto a helper stream.  A small pool per process, created on first use
// outside any stream capture; a lane (stream + fork/join events) is held under its mutex from fork to join, so that
// concurrent callers never share a pair of events.  No lane available (first call arrives inside a capture, or stream
// creation fails): the caller runs both chains on its own stream.
namespace dgp {
constexpr int N_SIDE_LANES = 16;
static SideLane g_lanes[DGP_MAX_DEVICES][N_SIDE_LANES];
static std::once_flag g_lanes_once[DGP_MAX_DEVICES];
static std::atomic<bool> g_lanes_ok[DGP_MAX_DEVICES];
static std::atomic<unsigned> g_lane_next{0};

SideLane* side_lane_acquire(cudaStream_t caller) {
    const int dev = current_device();
    if (!g_lanes_ok[dev].load(std::memory_order_acquire)) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(caller, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
            cudaGetLastError();
            return nullptr;                       // never create streams while a capture is in progress
        }
        std::call_once(g_lanes_once[dev], [dev] {
            bool ok = true;
            for (int i = 0; i < N_SIDE_LANES && ok; ++i) {
                ok = cudaStreamCreateWithFlags(&g_lanes[dev][i].stream, cudaStreamNonBlocking) == cudaSuccess &&
                     cudaEventCreateWithFlags(&g_lanes[dev][i].fork, cudaEventDisableTiming) == cudaSuccess &&
                     cudaEventCreateWithFlags(&g_lanes[dev][i].join, cudaEventDisableTiming) == cudaSuccess;
            }
            if (!ok) cudaGetLastError();
            g_lanes_ok[dev].store(ok, std::memory_order_release);
        });
        if (!g_lanes_ok[dev].load(std::memory_order_acquire)) return nullptr;
    }
    SideLane* l = &g_lanes[dev][g_lane_next.fetch_add(1) % N_SIDE_LANES];
    l->mu.lock();
    return l;
}
void side_lane_release(SideLane* l) { l->mu.unlock(); }
}  // namespace dgp

extern "C" int32_t dgp_mpad(int32_t M) { return (M + 63) / 64 * 64; }

extern "C" size_t dgp_cond_ws_bytes(const dgp_cond_desc* d, int64_t n_rows, int need_bwd) {
    if (!d || d->M < 1 || d->M > 1024 || d->R < 1 || n_rows < 0) return 0;
    return ws_layout(*d, n_rows, need_bwd).total;
}

extern "C" int dgp_cond_prepare(const dgp_cond_desc* d, void* ws, size_t ws_bytes, int need_bwd, double* kl_out, void* stream_) {
    int rc = check_desc(d);
    if (rc != DGP_OK) return rc;
    if (!ws) return DGP_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream_;
    // the caller sized the workspace for some n_rows; only the parameter part (independent of n_rows) is touched here
    const WsLayout w = ws_layout(*d, 0, 0);
    if (ws_bytes < w.total) return DGP_ERR_WORKSPACE;
    const int Mpad = w.Mpad, Dp = w.Dp;
    double* Zs = ws_ptr<double>(ws, w.Zs);
    double* zn = ws_ptr<double>(ws, w.zn);
    int32_t* ztask = ws_ptr<int32_t>(ws, w.ztask);
    double* Lm = ws_ptr<double>(ws, w.Lm);
    double* Linv = ws_ptr<double>(ws, w.Linv);
    double* LqT = ws_ptr<double>(ws, w.LqT);
    int32_t* status = ws_ptr<int32_t>(ws, w.status);

    // Two independent chains: (A) Z -> Kuu -> Cholesky -> Linv (one SM, a serial chain) and (B) everything that depends
    // on q_sqrt only (LqT, Lq, KL and their tiled copies).  B runs on a helper stream forked from and joined back
    // into `st` with events, so that it overlaps the Cholesky instead of queueing behind it on the critical path of the
    // forward pass.  (Inside a stream capture the helper stream simply becomes a parallel branch of the graph.)
    const WsLayout wb = ws_layout(*d, 0, 1);      // LinvT and SmI offsets do not depend on n_rows
    if (need_bwd && ws_bytes < wb.tail) return DGP_ERR_WORKSPACE;
    SideLane* lane = side_lane_acquire(st);
    cudaStream_t sb = lane ? lane->stream : st;
    if (lane) {
        cudaEventRecord(lane->fork, st);
        cudaStreamWaitEvent(sb, lane->fork, 0);
    }
    // ---- chain B
    double* klpart = ws_ptr<double>(ws, w.klpart);
    double* Lq = need_bwd ? ws_ptr<double>(ws, wb.Lq) : nullptr;
    k_lq_prep<<<dim3(Mpad / 32, Mpad / 32, d->R), dim3(32, 8), 0, sb>>>(*d, Mpad, LqT, Lq, klpart); DGP_COUNT_LAUNCH();
    if (kl_out) { k_kl<<<1, 256, 0, sb>>>(*d, klpart, d->R * (Mpad / 32) * (Mpad / 32), kl_out); DGP_COUNT_LAUNCH(); }
    rc = retile(LqT, ws_ptr<double>(ws, w.LqTTl), Mpad, d->R, sb);
    if (rc == DGP_OK && need_bwd) {
        // the backward kernel streams tril(q_sqrt_r) itself: d A = ... + 2 sum_r Lq_r (U_r . g_var_r) with the U_r kept from
        // the forward pass
        rc = retile(Lq, ws_ptr<double>(ws, wb.LqTl), Mpad, d->R, sb);
        // the per-CTA row-sum partials of the backward pass start from zero; dgp_cond_backward adds to them (row chunks).
        // (The Gram partials are not cleared: the first dgp_cond_backward_gram call of an evaluation stores into them.)
        if (rc == DGP_OK && cudaMemsetAsync(ws_ptr<char>(ws, wb.part), 0, wb.tail - wb.part, sb) != cudaSuccess) rc = DGP_ERR_LAUNCH;
    }
    if (lane) cudaEventRecord(lane->join, sb);
    // ---- chain A
    if (rc == DGP_OK) {
        k_prep_z<<<(Mpad + 127) / 128, 128, 0, st>>>(*d, Mpad, Dp, Zs, zn, ztask); DGP_COUNT_LAUNCH();
        k_kuu<<<dim3((Mpad + 127) / 128, Mpad), 128, 0, st>>>(*d, Mpad, Dp, Zs, zn, ztask, Lm); DGP_COUNT_LAUNCH();
        if (Mpad > 512) {
            // above the reach of the latency-tuned one-CTA kernel the single-CTA fallback took 11 ms at M = 1024: the
            // blocked factorisation (256-wide diagonal blocks through that kernel, panel solves and trailing updates as
            // GEMMs over all SMs) takes the serial chain down to four small factorisations
            rc = cholesky_blocked(Lm, Mpad, 1, status, ws_ptr<double>(ws, w.cholws), st);
            if (rc == DGP_OK) {
                k_zero_upper<<<dim3((Mpad + 127) / 128, Mpad), 128, 0, st>>>(Lm, Mpad); DGP_COUNT_LAUNCH();
                k_diag_inv<<<dim3(Mpad / 32, 1), 32, 0, st>>>(Lm, Linv, Mpad); DGP_COUNT_LAUNCH();
            }
        } else {
            rc = cholesky_batched(Lm, Mpad, 1, status, st, Linv);
        }
    }
    if (rc == DGP_OK) {
        rc = tri_inv_cols(Lm, Linv, Mpad, st);
        if (rc == DGP_OK) rc = retile(Linv, ws_ptr<double>(ws, w.LinvTl), Mpad, 1, st);
    }
    if (rc == DGP_OK && need_bwd) {
        double* LinvT = ws_ptr<double>(ws, wb.LinvT);
        k_transpose<<<dim3(Mpad / 32, Mpad / 32), dim3(32, 8), 0, st>>>(Linv, LinvT, Mpad); DGP_COUNT_LAUNCH();
        rc = retile(LinvT, ws_ptr<double>(ws, wb.LinvTTl), Mpad, 1, st);
    }
    if (lane) {                                   // join, also on the error paths (a capture must not be left forked)
        cudaStreamWaitEvent(st, lane->join, 0);
        side_lane_release(lane);
    }
    if (rc != DGP_OK) return rc;
    DGP_CHECK_LAUNCH();
    return DGP_OK;
}

extern "C" int dgp_cond_status(const void* ws, void* stream_, int32_t* status_host) {
    if (!ws || !status_host) return DGP_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream_;
    if (cudaMemcpyAsync(status_host, ws, sizeof(int32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess) return DGP_ERR_LAUNCH;
    if (cudaStreamSynchronize(st) != cudaSuccess) return DGP_ERR_LAUNCH;
    return DGP_OK;
}

extern "C" int dgp_cond_status_async(const void* ws, void* stream_, int32_t* status_pinned) {
    if (!ws || !status_pinned) return DGP_ERR_ARG;
    return cudaMemcpyAsync(status_pinned, ws, sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream_) == cudaSuccess
               ? DGP_OK : DGP_ERR_LAUNCH;
}

extern "C" const char* dgp_strerror(int code) {
    switch (code) {
        case DGP_OK: return "ok";
        case DGP_ERR_ARG: return "invalid argument";
        case DGP_ERR_LAUNCH: return "CUDA launch / runtime error";
        case DGP_ERR_WORKSPACE: return "workspace too small";
        case DGP_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown error";
    }
}
extern "C" const char* dgp_version(void) { return "dgp_b200 0.1 (sm_100a)"; }
extern "C" int64_t dgp_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" size_t dgp_usave_bytes(const dgp_cond_desc* d, int64_t n_rows) {
    if (!d || d->M < 1 || d->M > 1024 || d->R < 1 || n_rows < 0) return 0;
    return usave_doubles(dgp_mpad(d->M), d->R, n_rows) * sizeof(double);
}
#ifdef DGP_TEST_HOOKS
extern "C" void dgp_test_force_tiles(int fwd_nt, int bwd_nt) { g_force_fwd_nt = fwd_nt; g_force_bwd_nt = bwd_nt; }
extern "C" void dgp_test_force_gram_splits(int n) { g_force_gram_splits = n; }
extern "C" void dgp_test_debug_mode(int m) { g_debug_mode = m; }
extern "C" void dgp_test_force_stages(int n) { g_force_stages = n; }
#endif
