// Parameter-only tail of the backward pass of one whitened conditional (O(R M^3), once per evaluation):
//   d q_sqrt_r = 2 tril(G_r Lq_r) - d KL ,  d q_mu = sum_n a_n g_mean_n^T - d KL
//   d Lm = -tril(P)  ->  Cholesky reverse mode  d Kuu = sym(Lm^-T Phi(Lm^T dLm) Lm^-1)   
//   d Kuu, plus the row-summed Kuf / Kdiag pieces  ->  d Z, d variance, d lengthscales, d White
// This is what tf.gradients emits for CholeskyGrad, the kernel matrices and gauss_kl in the reference graph
// (dgplib/layers.py:58-69).  Deterministic: all cross-CTA sums are taken in a fixed order.
#include "common.cuh"
#include "gemm_small.cuh"

namespace dgp {

// sum of n values `stride` apart in four interleaved partial sums (fixed association): the loads are independent, only the
// adds chain -- these reductions sit on the latency path of the parameter tail
__device__ __forceinline__ double split_sum(const double* __restrict__ p, int n, size_t stride) {
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    int c = 0;
    for (; c + 8 <= n; c += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = p[(size_t)(c + u) * stride];
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u & 3] += v[u];
    }
    for (; c < n; ++c) s[c & 3] += p[(size_t)c * stride];
    return (s[0] + s[1]) + (s[2] + s[3]);
}

// (a) reduce the split-K Gram partials.  mat < R: full symmetric G (lower tiles mirrored); mat == R: Lbar = -tril(P).
//     blockIdx.z + mat0 = matrix index, so that the G_r and the Lbar parts can be launched separately.
__global__ void k_gram_reduce(const double* __restrict__ gram, int nsplit, int R, int Mpad, int mat0,
                              double* __restrict__ Gfull, double* __restrict__ Lbar) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, mat = mat0 + blockIdx.z;
    if (j >= Mpad) return;
    const size_t mm = (size_t)Mpad * Mpad;
    if (mat < R) {
        const bool lower = (i / 8) >= (j / 8);     // k_gram fills the 8x8 tiles on and below the tile diagonal
        const size_t off = lower ? (size_t)i * Mpad + j : (size_t)j * Mpad + i;
        Gfull[(size_t)mat * mm + (size_t)i * Mpad + j] = split_sum(gram + (size_t)mat * mm + off, nsplit, (size_t)(R + 1) * mm);
    } else {
        double s = 0.0;
        if (j <= i) s = split_sum(gram + (size_t)R * mm + (size_t)i * Mpad + j, nsplit, (size_t)(R + 1) * mm);
        Lbar[(size_t)i * Mpad + j] = -s;
    }
}

// (b) d q_sqrt (this conditional's rows) = 2 tril(G_r Lq_r) - klw (Lq - diag(1/diag Lq)); strict upper triangle: 0
__global__ void k_gq_sqrt(dgp_cond_desc d, int Mpad, const double* __restrict__ Tq, double klw, double* __restrict__ gq_sqrt) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, r = blockIdx.z;
    if (j >= d.M) return;
    const size_t o = ((size_t)(d.col0 + r) * d.M + i) * d.M + j;
    double v = 0.0;
    if (j <= i) {
        const double q = d.q_sqrt[o];
        v = 2.0 * Tq[((size_t)r * Mpad + i) * Mpad + j] - klw * (q - (i == j ? 1.0 / q : 0.0));
    }
    gq_sqrt[o] = v;
}

// sum the per-CTA partial buffers of k_cond_bwd over the CTAs, in CTA order (deterministic)
__global__ void k_part_reduce(const double* __restrict__ part, int ncta, int npart, double* __restrict__ psum) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= npart) return;
    psum[e] = split_sum(part + e, ncta, (size_t)npart);
}

// (c) d q_mu = row-summed partials - klw q_mu
__global__ void k_gq_mu(dgp_cond_desc d, int Mpad, int Dp, const double* __restrict__ psum, double klw,
                        double* __restrict__ gq_mu) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= d.M * d.R) return;
    const size_t o = (size_t)(idx / d.R) * d.ldr + d.col0 + idx % d.R;
    gq_mu[o] = psum[(size_t)Mpad * Dp + Mpad + idx] - klw * d.q_mu[o];
}

// (c') gradient of a trainable Linear mean: this conditional's columns of dA [Dx, ldr] and db [ldr]
__global__ void k_gmean(dgp_cond_desc d, const double* __restrict__ pm, double* __restrict__ gA, double* __restrict__ gb) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (d.Dx + 1) * d.R) return;
    const int k = idx / d.R, r = idx % d.R;
    if (k < d.Dx) gA[(size_t)k * d.ldr + d.col0 + r] = pm[idx];
    else if (gb) gb[d.col0 + r] = pm[idx];
}

// Phi: lower triangle with halved diagonal, in place
__global__ void k_phi(double* __restrict__ T, int Mpad) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= Mpad) return;
    double v = T[(size_t)i * Mpad + j];
    if (j > i) v = 0.0;
    else if (j == i) v *= 0.5;
    T[(size_t)i * Mpad + j] = v;
}

// (e) reverse mode of Kuu = k(Z,Z) + (white + jitter) I with dKuu = (T3 + T3^T)/2; one CTA per inducing row i.
//     rowp[i][0] = sum_j Kbar_ij k_ij/variance ; rowp[i][1] = Kbar_ii ; rowp[i][2+k] = sum_j Gr2_ij (zs_ik - zs_jk)^2
//     dzs[i][k]  = 4 sum_j Gr2_ij (zs_ik - zs_jk)
__global__ void __launch_bounds__(128)
k_kuu_bwd(dgp_cond_desc d, int Mpad, int Dp, const double* __restrict__ Zs, const double* __restrict__ zn,
          const int32_t* __restrict__ ztask, const double* __restrict__ T3, double* __restrict__ rowp,
          double* __restrict__ dzs) {
    __shared__ double red[32];
    __shared__ double zi[DGP_MAX_D];
    const int i = blockIdx.x, tid = threadIdx.x;
    const int nq = 2 + Dp;
    const int ti = ztask[i];
    for (int k = tid; k < Dp; k += blockDim.x) zi[k] = Zs[(size_t)i * Dp + k];
    __syncthreads();
    double acc[2 + 2 * DGP_MAX_D];  // [0]=dvar, [1..Dp]=ls terms, [1+Dp ..]=dzs  (indexed with compile-time-unknown Dp)
    for (int k = 0; k < 1 + 2 * Dp; ++k) acc[k] = 0.0;
    if (ti >= 0) {
        const double var = d.theta[(size_t)ti * (2 + d.D)];
        const int kind = d.kind[ti];
        for (int j = tid; j < d.M; j += blockDim.x) {
            if (ztask[j] != ti) continue;
            const double kb = 0.5 * (T3[(size_t)i * Mpad + j] + T3[(size_t)j * Mpad + i]);
            const double* zj = Zs + (size_t)j * Dp;
            double c = 0.0;
            for (int k = 0; k < Dp; ++k) c = fma(zi[k], zj[k], c);
            const double r2 = zn[i] + zn[j] - 2.0 * c;
            double kv, dk;
            kern_val_grad(kind, r2, kv, dk);
            acc[0] = fma(kb, kv, acc[0]);
            const double gr2 = kb * var * dk;
            for (int k = 0; k < Dp; ++k) {
                const double df = zi[k] - zj[k];
                acc[1 + k] = fma(gr2 * df, df, acc[1 + k]);
                acc[1 + Dp + k] = fma(gr2, df, acc[1 + Dp + k]);
            }
        }
    }
    for (int k = 0; k < 1 + 2 * Dp; ++k) {
        const double s = block_sum(acc[k], red);
        if (tid == 0) {
            if (k == 0) rowp[(size_t)i * nq + 0] = s;
            else if (k <= Dp) rowp[(size_t)i * nq + 1 + k] = s;
            else dzs[(size_t)i * Dp + (k - 1 - Dp)] = 4.0 * s;
        }
    }
    if (tid == 0) rowp[(size_t)i * nq + 1] = T3[(size_t)i * Mpad + i];
}

// (f) final assembly of d Z ...
__global__ void k_finish_Z(dgp_cond_desc d, int Mpad, int Dp, const double* __restrict__ Zs, const int32_t* __restrict__ ztask,
                           const double* __restrict__ psum, const double* __restrict__ dzs, double* __restrict__ gZ) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= d.M * d.Dx) return;
    const int m = idx / d.Dx, k = idx % d.Dx;
    double v = 0.0;
    const int tm = ztask[m];
    if (k < d.D && tm >= 0) {
        const double t2 = psum[(size_t)m * Dp + k], cs = psum[(size_t)Mpad * Dp + m];
        const double zs = Zs[(size_t)m * Dp + k];
        v = (dzs[(size_t)m * Dp + k] + 2.0 * (zs * cs - t2)) / d.theta[(size_t)tm * (2 + d.D) + 2 + k];
    }
    gZ[idx] = v;
}
// ... and of d theta[t][c] (one warp per entry, lanes stride over the inducing rows, fixed-order shuffle sum)
__global__ void k_finish_theta(dgp_cond_desc d, int Mpad, int Dp, const double* __restrict__ Zs,
                               const int32_t* __restrict__ ztask, const double* __restrict__ psum,
                               const double* __restrict__ rowp, double* __restrict__ gtheta) {
    const int nq = 2 + Dp, thd = 2 + d.D;
    const int idx = blockIdx.x, lane = threadIdx.x;
    const int tk = idx / thd, c = idx % thd;
    const double* cs = psum + (size_t)Mpad * Dp;
    double s = 0.0;
    for (int m = lane; m < d.M; m += 32) {
        if (c == 0) {
            if (ztask[m] == tk) s += rowp[(size_t)m * nq + 0];
        } else if (c == 1) {
            if (d.T == 1) s += rowp[(size_t)m * nq + 1];
        } else if (ztask[m] == tk) {
            const int k = c - 2;
            const double zs = Zs[(size_t)m * Dp + k];
            s += zs * zs * cs[m] + rowp[(size_t)m * nq + 2 + k];
        }
    }
    s = warp_sum(s);
    s += psum[(size_t)Mpad * Dp + Mpad + (size_t)Mpad * d.R + tk * nq + c];
    if (c >= 2) s *= -2.0 / d.theta[(size_t)tk * thd + c];
    if (lane == 0) gtheta[idx] = s;
}

}  // namespace dgp

using namespace dgp;

extern "C" int dgp_cond_backward_params(const dgp_cond_desc* d, void* ws, double kl_weight, double* gZ, double* gtheta,
                                        double* gq_mu, double* gq_sqrt, double* gmean_A, double* gmean_b, void* stream) {
    int rc = check_desc(d);
    if (rc != DGP_OK) return rc;
    if (!ws || !gZ || !gtheta || !gq_mu || !gq_sqrt) return DGP_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const WsLayout w = ws_layout(*d, 0, 1);
    const int Mpad = w.Mpad, Dp = w.Dp, R = d->R;
    const size_t mm = (size_t)Mpad * Mpad;
    const double* Zs = ws_ptr<double>(ws, w.Zs);
    const double* zn = ws_ptr<double>(ws, w.zn);
    const int32_t* ztask = ws_ptr<int32_t>(ws, w.ztask);
    const double* Lm = ws_ptr<double>(ws, w.Lm);
    const double* Linv = ws_ptr<double>(ws, w.Linv);
    const double* LinvT = ws_ptr<double>(ws, w.LinvT);
    const double* LqT = ws_ptr<double>(ws, w.LqT);
    const double* gram = ws_ptr<double>(ws, w.gram);
    const double* part = ws_ptr<double>(ws, w.part);
    double* tail = ws_ptr<double>(ws, w.tail);
    double* Gfull = tail;                 // [R]
    double* Tq = tail + mm * R;           // [R]
    double* Lbar = tail + mm * 2 * R;     // then T1, T2, T3
    double* T1 = Lbar + mm;
    double* T2 = T1 + mm;
    double* T3 = T2 + mm;
    double* scal = ws_ptr<double>(ws, w.scal);  // unused here; small vectors live at the front of T2 after it is consumed
    (void)scal;

    // Two chains that share only the Gram partials: (A, on `st`) Lbar -> Cholesky reverse mode -> Kuu reverse mode -> gZ,
    // gtheta, a serial chain of small GEMMs; (B, helper stream) G_r -> gq_sqrt and the row-sum partials -> gq_mu.  B joins
    // before the finishing kernels, which read the reduced partials.
    double* psum = ws_ptr<double>(ws, w.psum);
    SideLane* lane = side_lane_acquire(st);
    cudaStream_t sb = lane ? lane->stream : st;
    if (lane) {
        cudaEventRecord(lane->fork, st);
        cudaStreamWaitEvent(sb, lane->fork, 0);
    }
    // ---- chain B
    k_gram_reduce<<<dim3((Mpad + 127) / 128, Mpad, R), 128, 0, sb>>>(gram, w.nsplit, R, Mpad, 0, Gfull, Lbar); DGP_COUNT_LAUNCH();
    // Tq_r = G_r Lq_r ;  B(k,j) = Lq_r[k][j] = LqT_r[j][k]
    GemmArgs g{};
    g.A = Gfull; g.B = LqT; g.C = Tq; g.I = g.J = g.K = Mpad; g.batch = R;
    g.sAi = Mpad; g.sAk = 1; g.sBk = 1; g.sBj = Mpad; g.sCi = Mpad; g.sCj = 1;
    g.bA = g.bB = g.bC = (long long)mm; g.alpha = 1.0; g.beta = 0.0; g.lower_tiles_only = 1;
    rc = gemm_small(g, sb);
    if (rc == DGP_OK) {
        k_gq_sqrt<<<dim3((d->M + 127) / 128, d->M, R), 128, 0, sb>>>(*d, Mpad, Tq, kl_weight, gq_sqrt); DGP_COUNT_LAUNCH();
        k_part_reduce<<<(w.npart + 127) / 128, 128, 0, sb>>>(part, w.ncta, w.npart, psum); DGP_COUNT_LAUNCH();
        k_gq_mu<<<(d->M * R + 127) / 128, 128, 0, sb>>>(*d, Mpad, Dp, psum, kl_weight, gq_mu); DGP_COUNT_LAUNCH();
        if (gmean_A) {
            const double* pm = psum + (size_t)Mpad * Dp + Mpad + (size_t)Mpad * R + (size_t)d->T * (2 + Dp);
            k_gmean<<<((d->Dx + 1) * R + 127) / 128, 128, 0, sb>>>(*d, pm, gmean_A, gmean_b); DGP_COUNT_LAUNCH();
        }
    }
    if (lane) cudaEventRecord(lane->join, sb);
    // ---- chain A:  T1 = Lm^T Lbar ; Phi ; T2 = Linv^T T1 ; T3 = T2 Linv
    // rowp [Mpad][2 + Dp] and dzs [Mpad][Dp] reuse Lbar | T1 | T2 (contiguous, 3 Mpad^2 doubles, all consumed before the Kuu
    // reverse mode runs); 2 + 2 Dp <= 130 <= 3 Mpad always fits -- T1 | T2 alone would not for M <= 64 with D > 60
    double* rowp = Lbar;
    double* dzs = rowp + (size_t)Mpad * (2 + Dp);
    if (rc == DGP_OK) {
        k_gram_reduce<<<dim3((Mpad + 127) / 128, Mpad, 1), 128, 0, st>>>(gram, w.nsplit, R, Mpad, R, Gfull, Lbar); DGP_COUNT_LAUNCH();
        g = GemmArgs{};
        g.A = Lm; g.B = Lbar; g.C = T1; g.I = g.J = g.K = Mpad; g.batch = 1;
        g.sAi = 1; g.sAk = Mpad; g.sBk = Mpad; g.sBj = 1; g.sCi = Mpad; g.sCj = 1; g.alpha = 1.0;
        rc = gemm_small(g, st);
    }
    if (rc == DGP_OK) {
        k_phi<<<dim3((Mpad + 127) / 128, Mpad), 128, 0, st>>>(T1, Mpad); DGP_COUNT_LAUNCH();
        g.A = LinvT; g.B = T1; g.C = T2; g.sAi = Mpad; g.sAk = 1;
        rc = gemm_small(g, st);
    }
    if (rc == DGP_OK) {
        g.A = T2; g.B = Linv; g.C = T3;
        rc = gemm_small(g, st);
    }
    if (rc == DGP_OK) {   // Kuu reverse mode
        k_kuu_bwd<<<d->M, 128, 0, st>>>(*d, Mpad, Dp, Zs, zn, ztask, T3, rowp, dzs); DGP_COUNT_LAUNCH();
    }
    if (lane) {                                   // join, also on the error paths (a capture must not be left forked)
        cudaStreamWaitEvent(st, lane->join, 0);
        side_lane_release(lane);
    }
    if (rc != DGP_OK) return rc;
    k_finish_Z<<<(d->M * d->Dx + 127) / 128, 128, 0, st>>>(*d, Mpad, Dp, Zs, ztask, psum, dzs, gZ); DGP_COUNT_LAUNCH();
    k_finish_theta<<<d->T * (2 + d->D), 32, 0, st>>>(*d, Mpad, Dp, Zs, ztask, psum, rowp, gtheta); DGP_COUNT_LAUNCH();
    DGP_CHECK_LAUNCH();
    return DGP_OK;
}
