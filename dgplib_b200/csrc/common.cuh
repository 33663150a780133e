// Shared device/host helpers for libdgp_b200 (sm_100a only).
#pragma once
#include <atomic>
#include <mutex>
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdlib.h>
#include "../../include/dgp_b200.h"

namespace dgp {

// ------------------------------------------------------------------------------------------------ launch counter
extern std::atomic<long long> g_launches;  // bench.py's gpu_launches (host-side count of our kernel launches)
#define DGP_COUNT_LAUNCH() (::dgp::g_launches.fetch_add(1, std::memory_order_relaxed))

// Per-device "already configured" flags of the launchers (dynamic shared-memory opt-ins): monotonic values behind
// atomics, so that host threads driving the library concurrently never race on them (the attribute call itself is
// idempotent, so two threads that both see "not yet" simply both make it).
struct DevOnceSize {
    std::atomic<size_t> v[16];
    bool covers(int dev, size_t bytes) const { return v[dev].load(std::memory_order_acquire) >= bytes; }
    void raise(int dev, size_t bytes) {
        size_t cur = v[dev].load(std::memory_order_relaxed);
        while (cur < bytes && !v[dev].compare_exchange_weak(cur, bytes, std::memory_order_release)) {}
    }
};

#ifdef DGP_TEST_HOOKS
// Test-only build (libdgp_b200_test.so, `make test_lib`): lets the parity tests force every row-tile instantiation
// and the tuning experiments of tools/kbench.py override the Gram split count.  None of this exists in the product
// library.
extern int g_force_fwd_nt, g_force_bwd_nt, g_force_gram_splits, g_debug_mode, g_force_stages;
#endif

#define DGP_CHECK_LAUNCH()                                   \
    do {                                                     \
        if (cudaGetLastError() != cudaSuccess) return DGP_ERR_LAUNCH; \
    } while (0)

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Tiled operand layout streamed by the row-tile kernels (stream_gemm.cuh): [block row][chunk][128 rows][16 doubles],
// with the four 32-byte groups of a row permuted by (row & 3): element (row, k) of a chunk sits at
// row * 16 + (((k >> 2) ^ (row & 3)) << 2) + (k & 3).  The eight rows x four columns a DMMA A fragment reads then
// cover all 32 banks twice (conflict-free) without padding the 128-byte rows: a chunk is 16 KB instead of the 20 KB of
// a padded pitch, which buys a deeper ring (or a wider row tile) out of the 227 KB of shared memory.
constexpr int TILE_MB = 128;   // rows of a staged chunk
constexpr int TILE_KC = 16;    // depth of a staged chunk
constexpr int TILE_WLD = 16;   // row pitch of a staged chunk (doubles)
constexpr int TILE_DOUBLES = TILE_MB * TILE_WLD;
__host__ __device__ inline int tile_swizzle(int row, int k) { return (((k >> 2) ^ (row & 3)) << 2) | (k & 3); }

// U_r = tril(q_sqrt_r)^T A kept from the forward pass for the backward pass (dgp_cond_forward's U_save):
//   [r][16-deep chunk of the inducing index][group of 8 rows][16][8]
// i.e. one (output, chunk) slab holds, for every group of 8 consecutive rows n, a 16 x 8 block U_r[16 kc + k][8 g + n]
// of 1 KB.  A backward row tile of NT rows reads NT/8 consecutive blocks = ONE contiguous bulk copy per (output, chunk),
// for any tile width that is a multiple of 8, and the forward kernel's accumulator fragments store into it as 512-byte
// contiguous pieces per warp instruction.  Inside a block, element (k, n) sits at k * 8 + (n ^ (((k >> 1) & 1) << 2)):
// the two 32-byte halves of rows 2, 3 (mod 4) are swapped, so that the 4 x 4 elements a half-warp reads as a DMMA B
// fragment (k = 4 kk + t, n = g) fall into 16 distinct 8-byte bank pairs (unswizzled, rows t and t + 2 collide: measured
// 35 % excess shared-memory wavefronts).
constexpr int USAVE_BLOCK = 128;   // doubles of one 16 x 8 block
#ifdef DGP_NO_USWZ
__host__ __device__ inline int usave_swz(int k, int n) { return k * 8 + n; }
#else
__host__ __device__ inline int usave_swz(int k, int n) { return k * 8 + (n ^ (((k >> 1) & 1) << 2)); }
#endif
__host__ __device__ inline long long usave_groups(long long n_rows) { return (n_rows + 7) / 8; }
__host__ __device__ inline size_t usave_doubles(int Mpad, int R, long long n_rows) {
    return (size_t)R * (Mpad / TILE_KC) * (size_t)usave_groups(n_rows) * USAVE_BLOCK;
}
__host__ __device__ inline int tiled_mpad(int Mpad) { return (Mpad + TILE_MB - 1) / TILE_MB * TILE_MB; }
__host__ __device__ inline size_t tiled_doubles(int Mpad) {
    const int mt = tiled_mpad(Mpad);
    return (size_t)(mt / TILE_MB) * (mt / TILE_KC) * TILE_DOUBLES;
}
static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// ------------------------------------------------------------------------------------------------ workspace layout
// One conditional's workspace.  All offsets in bytes from the workspace base; computed identically by
// dgp_cond_ws_bytes and every entry point (pure function of the descriptor, n_rows and need_bwd).
struct WsLayout {
    int Mpad, Dp, ncta, nsplit, npart;
    size_t status;     // int32[64]
    size_t scal;       // double[256] small reductions (KL partials, ...)
    size_t Zs;         // [Mpad][Dp]   z / lengthscale(task(z)); pad rows / cols zero
    size_t zn;         // [Mpad]       |Zs row|^2
    size_t ztask;      // int32[Mpad]  task id of each inducing point (-1 on pad rows; 0 when T == 1)
    size_t Lm;         // [Mpad][Mpad] Kuu(+jitter) -> Cholesky factor in place (pad = identity)
    size_t Linv;       // [Mpad][Mpad] Lm^-1 (lower)
    size_t LqT;        // [R][Mpad][Mpad] tril(q_sqrt_r)^T (upper), pad zero
    size_t LinvTl;     // tiled copy of Linv        (streamed by the forward kernel)
    size_t LqTTl;      // [R] tiled copies of LqT_r (streamed by the forward kernel)
    size_t tl;         // doubles of one tiled matrix
    size_t klpart;     // [R][(Mpad/32)^2] per-tile partial sums of the KL term (written by k_lq_prep)
    size_t cholws;     // scratch of the blocked Cholesky (Mpad > 512 only)
    // backward only
    size_t LinvT;      // [Mpad][Mpad]
    size_t Lq;         // [R][Mpad][Mpad] tril(q_sqrt_r) (lower), pad zero
    size_t LinvTTl;    // tiled copy of LinvT       (streamed by the backward kernel)
    size_t LqTl;       // [R] tiled copies of Lq_r  (streamed by the backward kernel)
    size_t dKuf;       // [n_rows][Mpad]
    size_t gmv;        // [n_rows][2R] combined g_mean | g_var of this conditional
    size_t gram;       // [nsplit][R+1][Mpad][Mpad]  split-K partials of G_r (r<R) and P (r==R)
    size_t part;       // [ncta][npart] per-CTA partials: T2z[Mpad][Dp], cs[Mpad], dq_mu[Mpad][R], th[T][2+Dp], dA_lin[Dx][R], db[R]
    size_t psum;       // [npart] the per-CTA partials summed over CTAs (fixed order)
    size_t tail;       // [2R+4][Mpad][Mpad] scratch of the parameter tail (Gfull, Tq, Lbar, T1, T2, T3)
    size_t total;
};

// Host-side caches (shared-memory opt-ins, SM counts, helper streams) are kept per device, so that one process may
// drive several GPUs through the library (the usual deployment is still one process per GPU).
constexpr int DGP_MAX_DEVICES = 16;
int current_device();   // cudaGetDevice, clamped to [0, DGP_MAX_DEVICES)
int device_sm_count();

// Helper stream + fork/join events for the independent chains of the parameter stages (defined in prep.cu).  acquire
// returns nullptr when no lane can be had (then the caller keeps everything on its own stream); a lane is held under its
// mutex from fork to join.
struct SideLane {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    std::mutex mu;
};
SideLane* side_lane_acquire(cudaStream_t caller);
void side_lane_release(SideLane* lane);

static inline WsLayout ws_layout(const dgp_cond_desc& d, int64_t n_rows, int need_bwd) {
    WsLayout w{};
    w.Mpad = dgp_mpad(d.M);
    w.Dp = round_up(d.D, 4);
    w.ncta = device_sm_count();
    const size_t mm = size_t(w.Mpad) * w.Mpad * sizeof(double);
    size_t o = 0;
    w.status = o; o += align256(64 * sizeof(int32_t));
    w.scal = o;   o += align256(256 * sizeof(double));
    w.Zs = o;     o += align256(size_t(w.Mpad) * w.Dp * sizeof(double));
    w.zn = o;     o += align256(size_t(w.Mpad) * sizeof(double));
    w.ztask = o;  o += align256(size_t(w.Mpad) * sizeof(int32_t));
    w.Lm = o;     o += align256(mm);
    w.Linv = o;   o += align256(mm);
    w.LqT = o;    o += align256(mm * d.R);
    w.tl = tiled_doubles(w.Mpad);
    w.LinvTl = o; o += align256(w.tl * sizeof(double));
    w.LqTTl = o;  o += align256(w.tl * sizeof(double) * d.R);
    w.klpart = o; o += align256(size_t(d.R) * (w.Mpad / 32) * (w.Mpad / 32) * sizeof(double));
    w.cholws = o; if (w.Mpad > 512) o += align256((2 * 256 * 256 + size_t(w.Mpad) * 256 + 64) * sizeof(double));
    if (need_bwd) {
        // split-K factor of the Gram kernel: enough CTAs to fill the machine ~3x (independent of n_rows so that every
        // offset except the two row-sized buffers at the end is a function of the descriptor only)
        const int nb = w.Mpad / 64;
        const int tiles = (d.R + 1) * nb * (nb + 1) / 2;
        // k_gram runs 4 CTAs per SM and its CTAs are uneven (diagonal tiles do 9/16 of the work): about five waves of
        // CTAs balance well (measured: 13 -> 32 splits took the T-shape Gram kernel from 2.33 to 2.09 ms); the partial
        // buffers are bounded to 512 MB
        const int slots = 4 * w.ncta;
        int ns = (5 * slots + tiles / 2) / tiles;
        const size_t per_split = mm * (d.R + 1);
        if ((size_t)ns * per_split > (size_t)512 << 20) ns = (int)(((size_t)512 << 20) / per_split);
        if (ns < 1) ns = 1;
        if (ns > 64) ns = 64;
#ifdef DGP_TEST_HOOKS
        if (g_force_gram_splits >= 1 && g_force_gram_splits <= 64) ns = g_force_gram_splits;   // tuning experiments
#endif
        w.nsplit = ns;
        w.npart = round_up(w.Mpad * w.Dp + w.Mpad + w.Mpad * d.R + d.T * (2 + w.Dp) + d.Dx * d.R + d.R, 32);
        w.LinvT = o; o += align256(mm);
        w.Lq = o;    o += align256(mm * d.R);
        w.LinvTTl = o; o += align256(w.tl * sizeof(double));
        w.LqTl = o;    o += align256(w.tl * sizeof(double) * d.R);
        w.gram = o;  o += align256(mm * (d.R + 1) * w.nsplit);
        w.part = o;  o += align256(size_t(w.ncta) * w.npart * sizeof(double));
        w.psum = o;  o += align256(size_t(w.npart) * sizeof(double));
        w.tail = o;  o += align256(mm * (2 * d.R + 4));
        w.gmv = o;   o += align256(size_t(n_rows) * 2 * d.R * sizeof(double));
        w.dKuf = o;  o += align256(size_t(n_rows) * w.Mpad * sizeof(double));
    }
    w.total = o;
    return w;
}

// Row-tile plan of a persistent row kernel over n rows on `sms` SMs (Mpad <= 256): one tile width for the whole launch,
// or 64-row tiles for the full waves plus ONE wave of narrower tiles for the rows a last, partly empty wave would hold
// (100 000 rows on 148 SMs: 10 waves of 64-row tiles + 132 tiles of 40 rows instead of 11 waves: -3 % time).
// Costs are rows per SM weighted by the measured per-row penalty of each width (pen72: the backward kernel's 72-row
// variant has only a 2-slot ring and is 15 % slower per row, the forward kernel's 4 %).
struct TilePlan { int nt_main; long long rows_main; int nt_tail; };
static inline TilePlan plan_tiles(long long n, long long sms, double pen72) {
    auto waves = [&](long long rows, long long nt) { return ((rows + nt - 1) / nt + sms - 1) / sms; };
    TilePlan best{64, n, 0};
    double cost = (double)waves(n, 64) * 64.0;
    const double c72 = (double)waves(n, 72) * 72.0 * pen72, c32 = (double)waves(n, 32) * 32.0 * 1.08;
    if (c72 < cost) { cost = c72; best = TilePlan{72, n, 0}; }
    if (c32 < cost) { cost = c32; best = TilePlan{32, n, 0}; }
    const long long w = n / (64 * sms), rem = n - w * 64 * sms;
    if (w >= 1 && rem > 0) {
        const int cand[3] = {16, 32, 40};
        const double pen[3] = {1.6, 1.08, 1.05};
        for (int i = 0; i < 3; ++i)
            if ((rem + cand[i] - 1) / cand[i] <= sms) {
                const double c = (double)w * 64.0 + cand[i] * pen[i];
                if (c < cost) { cost = c; best = TilePlan{64, w * 64 * sms, cand[i]}; }
                break;
            }
    }
    return best;
}

template <typename T>
static inline T* ws_ptr(void* ws, size_t off) { return reinterpret_cast<T*>(static_cast<char*>(ws) + off); }
template <typename T>
static inline const T* ws_ptr(const void* ws, size_t off) { return reinterpret_cast<const T*>(static_cast<const char*>(ws) + off); }

int check_desc(const dgp_cond_desc* d);
// batched in-place lower Cholesky of `batch` Mpad x Mpad matrices (Mpad multiple of 64, <= 1024); status[b] = failing pivot + 1
// Linv (nullable): receives the inverses of the 32 x 32 diagonal blocks of the factor (input of k_tri_inv_cols)
int cholesky_batched(double* K, int Mpad, int batch, int32_t* status, cudaStream_t st, double* Linv = nullptr);
// Linv = L^-1 for one lower-triangular Mpad x Mpad factor whose 32 x 32 diagonal-block inverses are already in Linv
// (written by cholesky_batched); prep.cu
int tri_inv_cols(const double* L, double* Linv, int Mpad, cudaStream_t st);
// Blocked right-looking Cholesky over all SMs for np > 512 (np a multiple of 64; fullcov.cu): 256-wide diagonal blocks
// through the one-CTA kernel, panel solves and trailing updates as GEMMs.  Lower factor in place (the strict upper
// triangle outside the diagonal blocks keeps its input values); scratch: chol_blocked_scratch_doubles(np, batch).
size_t chol_blocked_scratch_doubles(int np, int batch);
int cholesky_blocked(double* K, int np, int batch, int32_t* status, double* scratch, cudaStream_t st);

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------ PTX wrappers
// fp64 tensor-core MMA, m8n8k4 (SASS: DMMA.8x8x4).  Fragment layout (g = lane>>2, t = lane&3):
//   a = A[g][t]     b = B[t][g]     c[i] = C[g][2t+i]
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

// ---- bulk async copies (TMA engine, SASS UBLKCP): 16-byte aligned addresses, size multiple of 16
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "DGP_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DGP_DONE_%=;\n"
        "bra DGP_WAIT_%=;\n"
        "DGP_DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// the same copy with an L2 eviction-priority hint (createpolicy): evict_last for lines that will be read again soon,
// evict_first for streaming data that will not
__device__ __forceinline__ uint64_t l2_policy(bool keep) {
    uint64_t p;
    if (keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
    else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
                 :: "l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void* gdst, const void* smem_src, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;\n"
                 :: "l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes), "l"(policy) : "memory");
}
// L2 prefetch of a contiguous global range (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" :: "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ covariance functions
// value k(r2)/variance and d k/d r2 / variance  (gpflow RBF / Matern52)
__device__ __forceinline__ double kern_val(int kind, double r2) {
    if (kind == DGP_KIND_RBF) return exp(-0.5 * r2);
    const double s5 = 2.23606797749978969641;
    const double r = sqrt(r2 + 1e-12);
    return (1.0 + s5 * r + (5.0 / 3.0) * r * r) * exp(-s5 * r);
}
__device__ __forceinline__ void kern_val_grad(int kind, double r2, double& k, double& dk_dr2) {
    if (kind == DGP_KIND_RBF) {
        k = exp(-0.5 * r2);
        dk_dr2 = -0.5 * k;
        return;
    }
    const double s5 = 2.23606797749978969641;
    const double r = sqrt(r2 + 1e-12);
    const double e = exp(-s5 * r);
    k = (1.0 + s5 * r + (5.0 / 3.0) * r * r) * e;
    dk_dr2 = -(5.0 / 6.0) * (1.0 + s5 * r) * e;  // (dk/dr) / (2 r); finite at r -> 0
}

// ------------------------------------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
// One N(0,1) draw per (global row, output column): counter = (row, col), key = seed.  Independent of how rows are
// sharded over GPUs or tiles, so 1/2/4/8-GPU runs see the same noise.
__device__ __forceinline__ double philox_normal(uint64_t seed, uint64_t row, uint32_t col) {
    uint32_t c[4] = {(uint32_t)row, (uint32_t)(row >> 32), col, 0x44475042u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u1 = ((double)(((uint64_t)c[0] << 21) | (c[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
    const double u2 = ((double)(((uint64_t)c[2] << 21) | (c[3] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// ------------------------------------------------------------------------------------------------ block reduce
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// deterministic block sum; `red` is shared scratch of >= 32 doubles; result valid in every thread
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < nw; ++i) s += red[i];
    return s;
}
#endif  // __CUDACC__

}  // namespace dgp
