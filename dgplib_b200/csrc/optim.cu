// Training-step kernels around the ELBO + gradient evaluation: the reference trains with
// gpflow.train.AdamOptimizer(lr).minimize(model, maxiter) (tests/test_dsdgp.py:56-60, dgplib/dsdgp.py:74-81), i.e.
// tf.train.AdamOptimizer on the UNCONSTRAINED variables of -ELBO, positive parameters going through gpflow's Log1pe
// transform theta = softplus(u) + 1e-6.
//
// All parameters of a model live in ONE flat unconstrained buffer u (dgplib_b200/flatparams.py).  Two launches bracket
// the evaluation:
//   dgp_transform_params : the small composite constrained arrays the kernels read -- per conditional theta[T][2 + D]
//                          (variance, summed White variances, lengthscales broadcast when not ARD) and the likelihood
//                          variances -- as sums of transformed u entries (CSR).  Z, q_mu, q_sqrt and a Linear mean are read
//                          by the kernels straight from u (identity transform).
//   dgp_adam_step        : one pass over u: gathers d ELBO / d constrained from the evaluation's flat gradient buffer
//                          (one image per element, or a CSR row for elements with several images: a scalar lengthscale,
//                          a shared Z), applies the softplus chain factor, and takes tf.train.AdamOptimizer's step on
//                          -ELBO; the step counter lives on the device so the launch can be replayed from a CUDA graph.
#include "common.cuh"

namespace dgp {

__device__ __forceinline__ double softplus(double x) {
    // log(1 + exp(x)), stable on both sides (gpflow Log1pe forward)
    return x > 0.0 ? x + log1p(exp(-x)) : log1p(exp(x));
}
__device__ __forceinline__ double sigmoid(double x) {
    if (x >= 0.0) return 1.0 / (1.0 + exp(-x));
    const double e = exp(x);
    return e / (1.0 + e);
}

__global__ void k_transform(const double* __restrict__ u, const int32_t* __restrict__ crow, const int32_t* __restrict__ cidx,
                            const uint8_t* __restrict__ cpos, int n_c, double lower, double* __restrict__ c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_c) return;
    double s = 0.0;
    for (int e = crow[i]; e < crow[i + 1]; ++e) {
        const double x = u[cidx[e]];
        s += cpos[e] ? softplus(x) + lower : x;
    }
    c[i] = s;
}

__global__ void k_adam(double* __restrict__ u, double* __restrict__ m, double* __restrict__ v, long long n,
                       const double* __restrict__ grad, const int32_t* __restrict__ gidx, const int32_t* __restrict__ grow,
                       const int32_t* __restrict__ gimg, const uint8_t* __restrict__ flags, double lr, double beta1,
                       double beta2, double eps, const long long* __restrict__ step) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint8_t f = flags[j];
    if (!(f & 2)) return;                      // not trainable
    const int32_t gi = gidx[j];
    double g = 0.0;
    if (gi >= 0) {
        g = grad[gi];
    } else if (gi <= -2) {
        const int row = -2 - gi;
        for (int e = grow[row]; e < grow[row + 1]; ++e) g += grad[gimg[e]];    // fixed order: deterministic
    }
    const double x = u[j];
    if (f & 1) g *= sigmoid(x);                // d softplus(u) / d u
    g = -g;                                    // the optimiser minimises -ELBO
    const double t = (double)(*step + 1);
    const double lr_t = lr * sqrt(1.0 - pow(beta2, t)) / (1.0 - pow(beta1, t));
    const double mj = beta1 * m[j] + (1.0 - beta1) * g;
    const double vj = beta2 * v[j] + (1.0 - beta2) * g * g;
    m[j] = mj;
    v[j] = vj;
    u[j] = x - lr_t * mj / (sqrt(vj) + eps);   // tf.train.AdamOptimizer: epsilon is "epsilon hat" (added to sqrt(v), not to sqrt(v hat))
}

// elbo -= w * sum_c kl[c], summed in index order by one thread (deterministic; a handful of conditionals)
__global__ void k_finish_elbo(double* __restrict__ elbo, const double* __restrict__ kl, int ncond, double w) {
    double s = 0.0;
    for (int c = 0; c < ncond; ++c) s += kl[c];
    *elbo -= w * s;
}
__global__ void k_step_inc(long long* step) { *step += 1; }
__global__ void k_set_u64(unsigned long long* p, unsigned long long v) { *p = v; }

}  // namespace dgp

using namespace dgp;

extern "C" int dgp_transform_params(const double* u, const int32_t* crow, const int32_t* cidx, const uint8_t* cpos,
                                    int32_t n_c, double lower, double* c, void* stream) {
    if (n_c < 0 || (n_c > 0 && (!u || !crow || !cidx || !cpos || !c))) return DGP_ERR_ARG;
    if (n_c == 0) return DGP_OK;
    k_transform<<<(n_c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(u, crow, cidx, cpos, n_c, lower, c);
    DGP_COUNT_LAUNCH();
    DGP_CHECK_LAUNCH();
    return DGP_OK;
}

extern "C" int dgp_adam_step(double* u, double* m, double* v, int64_t n, const double* grad, const int32_t* gidx,
                             const int32_t* grow, const int32_t* gimg, const uint8_t* flags, double lr, double beta1,
                             double beta2, double eps, int64_t* step_dev, void* stream) {
    if (n < 0 || (n > 0 && (!u || !m || !v || !grad || !gidx || !flags || !step_dev))) return DGP_ERR_ARG;
    if (n == 0) return DGP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    k_adam<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(u, m, v, (long long)n, grad, gidx, grow, gimg, flags, lr, beta1, beta2,
                                                          eps, reinterpret_cast<const long long*>(step_dev));
    DGP_COUNT_LAUNCH();
    k_step_inc<<<1, 1, 0, st>>>(reinterpret_cast<long long*>(step_dev));
    DGP_COUNT_LAUNCH();
    DGP_CHECK_LAUNCH();
    return DGP_OK;
}

extern "C" int dgp_finish_elbo(double* elbo, const double* kl, int32_t ncond, double kl_weight, void* stream) {
    if (!elbo || ncond < 0 || (ncond > 0 && !kl)) return DGP_ERR_ARG;
    k_finish_elbo<<<1, 1, 0, (cudaStream_t)stream>>>(elbo, kl, ncond, kl_weight);
    DGP_COUNT_LAUNCH();
    DGP_CHECK_LAUNCH();
    return DGP_OK;
}

extern "C" int dgp_zero(void* dst, size_t bytes, void* stream) {
    if (!dst && bytes) return DGP_ERR_ARG;
    return cudaMemsetAsync(dst, 0, bytes, (cudaStream_t)stream) == cudaSuccess ? DGP_OK : DGP_ERR_LAUNCH;
}

extern "C" int dgp_set_u64(uint64_t* dst, uint64_t value, void* stream) {
    if (!dst) return DGP_ERR_ARG;
    k_set_u64<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)value);
    DGP_COUNT_LAUNCH();
    DGP_CHECK_LAUNCH();
    return DGP_OK;
}
