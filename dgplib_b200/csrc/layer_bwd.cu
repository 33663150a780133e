// Reverse mode of the per-row stage of one whitened conditional (what tf.gradients emits for the graph of
// gpflow's base_conditional(white=True) called at dgplib/layers.py:65-71, and for normal_sample, dgplib/utils.py:27).
// Formulas: SURVEY.md Appendix B, with A = Lm^-1 Kuf and U_r = tril(q_sqrt_r)^T A kept from the forward pass:
//
//   g_mean = g_mean_in + g_F ;  g_var = g_var_in + g_F (F - mean) / (2 var)
//   dA   = q_mu g_mean^T - 2 A . (sum_r g_var_r) + 2 sum_r Lq_r (U_r . g_var_r)     [lower-triangular streamed GEMMs, K = R M / 2]
//   dKuf = Lm^-T dA                                                                   [upper-triangular streamed GEMM]
//   dX, dZ, d lengthscale, d variance from dKuf . k'(r2)                              [per tile, small]
//   G_r  = sum_n g_var[n,r] a_n a_n^T  (-> d q_sqrt_r = 2 tril(G_r Lq_r)),  P = sum_n dKuf_n a_n^T (-> d Lm = -tril P)
//                                                                                     [split-K Gram kernel over the rows]
// This executes the R M^2 + M^2 flops per row of the reference's own reverse mode (backward = 2 F with the
// Gram kernel); the first version of this kernel recomputed through the dense S_r - I instead (2 R M^2 + M^2) in order to
// keep only A.  U_r costs 8 R Mpad bytes per row each way in HBM, which the tensor-pipe-bound kernels hide.
// k_cond_bwd : one CTA per SM walks row tiles.  BOTH operands of the dA phase are streamed through the same ring by the
//              producer warp (TMA engine): the tiled tril(q_sqrt_r) chunk from L2 and the matching [NT/8][16][8] chunk of
//              the saved U_r from HBM; only dA / dKuf (one row tile) is resident in shared memory, so the row tile is twice
//              as wide as a design that also keeps the A tile resident (64-72 rows at Mpad <= 256, 32 at 512, 16 at 1024).
//              A itself is needed only elementwise (epilogue of dA, d q_mu) and is read straight from L2 after a bulk L2
//              prefetch of the tile.  dKuf leaves by bulk async store; row-summed pieces go to per-CTA partial buffers
//              (no atomics: bitwise reproducible).
// k_gram     : 64x64 output tiles x split-K over the rows, fp64 DMMA, partials per split (no atomics).
#include "stream_gemm.cuh"

#ifndef DGP_BWD_WN
#define DGP_BWD_WN 2   // consumer warp grid is 4 x WN
#endif

namespace dgp {

#ifdef DGP_PROFILE_PHASES
#define PH(i) do { if (tid == 0) { long long now_ = clock64(); ph[i] += now_ - tlast; tlast = now_; } } while (0)
#else
#define PH(i) do {} while (0)
#endif

struct BwdArgs {
    dgp_cond_desc d;
    int Mpad, Dp;
    const double* Zs; const double* zn; const int32_t* ztask;
    const double* LinvT; const double* LqL; size_t tl;   // tiled copies (stream_gemm.cuh): Lm^-T and tril(q_sqrt_r)
    const double* X; long long x_rows, n_rows;
    const double* A_save; const double* U_save; long long NG;
    const double* K_save;    // Kuf per row kept by the forward pass (nullptr: recompute the covariance function)
    int k_rbf;               // K_save is given and every kernel of the conditional is an RBF: k'(r2) = -k / 2
    const double* mean; const double* var; const double* F; int ldf;
    const double* g_mean_in; const double* g_var_in; const double* g_F; int ldgf;
    double* gX; int accumulate_gX;
    int mean_grad;       // != 0: also accumulate the gradient of a trainable Linear mean (A, b) in the per-CTA partials
    double* dKuf; double* gmv; double* part; int npart; double* scal;
    long long row0;      // first row of this launch (a call may be split into a main and a tail launch of another width)
    long long row_end;   // one past the last row of this launch
    int ntiles;
    int debug;           // test-hook builds only
};

template <class C>
struct BwdSmem {
    int ldb, xld;
    int Dt, Wst, xs, xn, gm, gv, gs, rs, dv, rowq, xtask, bars, total_doubles;
    __host__ __device__ BwdSmem(int Mpad, int Dp, int R) {
        ldb = Mpad + 4;
        xld = Dp + 4;
        int o = 0;
        Dt = o; o += C::NT * ldb;
        Wst = o; o += C::STAGES * C::STAGE_DOUBLES;
        xs = o; o += C::NT * xld;
        xn = o; o += C::NT;
        gm = o; o += C::NT * R;
        gv = o; o += C::NT * R;
        gs = o; o += C::NT;
        rs = o; o += C::NT;
        dv = o; o += C::NT;
        rowq = o; o += C::NT * (2 + Dp) > 32 ? C::NT * (2 + Dp) : 32;   // also: scratch of >= NWARPS doubles
        xtask = o; o += C::NT / 2 + 1;
        o += o & 1;
        bars = o; o += 2 * C::STAGES;
        total_doubles = o;
    }
};

// One launch over the (conditional, row tile) pairs of a MultikernelLayer (see layer_fwd.cu): members share the rows and
// the shapes, differ in workspaces (operands, per-CTA partials, dKuf), descriptors and saved tensors.  Used when no input
// gradient is wanted (a MultikernelInputLayer): gX is accumulated over the members and two of them in flight would race.
constexpr int DGP_GROUP_MAX = 8;
struct BwdGroup {
    int n;
    BwdArgs m[DGP_GROUP_MAX];
};
__device__ __forceinline__ const BwdArgs& bwd_member(const BwdArgs& p, int) { return p; }
__device__ __forceinline__ const BwdArgs& bwd_member(const BwdGroup& p, int c) { return p.m[c]; }
__device__ __forceinline__ int bwd_members(const BwdArgs&) { return 1; }
__device__ __forceinline__ int bwd_members(const BwdGroup& p) { return p.n; }

template <class C, class P>
__global__ void __launch_bounds__(C::NTHREADS, 1) k_cond_bwd(const __grid_constant__ P p) {
    extern __shared__ __align__(128) double smem[];
    const BwdArgs& a0 = bwd_member(p, 0);      // shapes, tile count and shared-memory layout are common to all members
    const int ntiles = a0.ntiles, total_tiles = ntiles * bwd_members(p);
    const BwdSmem<C> L(a0.Mpad, a0.Dp, a0.d.R);
    double* Dt = smem + L.Dt;
    double* Wst = smem + L.Wst;
    double* xs = smem + L.xs;
    double* xn = smem + L.xn;
    double* gm = smem + L.gm;
    double* gv = smem + L.gv;
    double* gs = smem + L.gs;
    double* rs = smem + L.rs;
    double* dv = smem + L.dv;
    double* rowq = smem + L.rowq;
    int32_t* xtask = reinterpret_cast<int32_t*>(smem + L.xtask);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int Mpad = a0.Mpad, Dp = a0.Dp, ldb = L.ldb, xld = L.xld, R = a0.d.R;
    const int nb = (Mpad + C::MB - 1) / C::MB;
    const int nkc = Mpad / C::KC;
    const int nq = 2 + Dp;
    const BwdArgs* cur = &a0;       // member of the tile being produced / consumed (the segment lists read it)

    // segment lists of the two streamed phases (shared by the producer warp and the consumers)
    auto seg_dA = [&](int s) {      // dA block row i accumulates Lq_r (U_r . g_r) over all r: lower-triangular Lq_r
        Seg sg;
        sg.i = s / R;
        sg.r = s % R;
        sg.W = cur->LqL + (size_t)sg.r * cur->tl;
        sg.kc0 = 0;
        int kend = (sg.i + 1) * C::MB;
        if (kend > Mpad) kend = Mpad;
        sg.kc1 = kend / C::KC;
        sg.flush = (sg.r == R - 1); // one accumulation over all r: the column weights g_var[n, r] scale the B fragments
        return sg;
    };
    auto seg_dK = [&](int s) {      // dKuf = LinvT * dA, block rows top-down, in place
        Seg sg;
        sg.W = cur->LinvT;
        sg.i = s;
        sg.kc0 = s * C::MB / C::KC;
        sg.kc1 = Mpad / C::KC;
        sg.r = 0;
        sg.flush = true;
        return sg;
    };

    Pipe pipe;
    pipe_init<C>(pipe, bars, Wst);
    // the U halves of the ring slots start as zeros: columns beyond the last row of a ragged tile are never copied and
    // enter the GEMM with weight 0, which must not meet a NaN bit pattern left in shared memory by an earlier kernel
    for (int e = tid; e < C::STAGES * C::UCH_DOUBLES; e += C::NTHREADS)
        Wst[(e / C::UCH_DOUBLES) * C::STAGE_DOUBLES + TILE_DOUBLES + e % C::UCH_DOUBLES] = 0.0;
    fence_async_smem();
    __syncthreads();   // the only CTA-wide barrier
    if (warp == C::NWARPS) {
        // ---- producer warp: lane 0 issues the bulk copies of every ring slot (W chunk + the matching U chunk).
        // (Pulling the next segment's U chunks into L2 ahead of the ring -- bulk prefetches by lane 0, or plain prefetch
        // instructions by the idle lanes -- was measured slower both ways, 3.10 -> 3.31 / 3.25 ms at the T hidden-layer
        // shape: the ring's look-ahead already covers the HBM latency.)
        if (lane == 0) {
            for (int gt = blockIdx.x; gt < total_tiles; gt += gridDim.x) {
                const BwdArgs& a = bwd_member(p, gt / ntiles);
                cur = &a;
                const int tile = gt % ntiles;
                const long long n0 = a.row0 + (long long)tile * C::NT;
                const long long left = a.n_rows - n0;
                const int nvalid = (int)(left < C::NT ? left : C::NT);
                // the A tile is only read elementwise (dA epilogue, d q_mu): pull it into L2 ahead of those reads.
                // (Also prefetching the saved Kuf tile and the next tile's row data -- upstream gradients, F, mean, var,
                // inputs -- was measured: no gain when issued behind the previous tile's last chunk, a loss when issued
                // here, where 2 x 131 KB per SM queue up in front of the prologue's own loads.)
                bulk_prefetch_l2(a.A_save + (size_t)n0 * Mpad, (uint32_t)((size_t)nvalid * Mpad * sizeof(double)));
                long long g0 = n0 / 8;
                const long long gl = a.NG - g0;
                uint32_t ubytes = (uint32_t)((gl < C::NT / 8 ? gl : C::NT / 8) * USAVE_BLOCK * sizeof(double));
#ifdef DGP_TEST_HOOKS
                if (a.debug == 1 || a.debug == 4) ubytes = 0;   // timing experiments only: no U traffic / U always from L2
                if (a.debug == 2) g0 = (long long)blockIdx.x * (C::NT / 8);
#endif
                // a U chunk is read once per block row that reaches it: every block row but the last leaves it behind for a
                // later one (~1 MB of other chunks per CTA in between), so those reads ask L2 to keep the lines and the last
                // read of each chunk lets them go first (without the hints the re-reads came from HBM: 3.2 GB per launch at
                // the T hidden-layer shape against 2.1 GB algorithmic)
                const uint64_t keep = l2_policy(true), drop = l2_policy(false);
                auto ufn = [&](const Seg& sg, int kc, uint32_t& bytes, uint64_t& policy) {
                    bytes = ubytes;
                    policy = (sg.i < nb - 1) ? keep : drop;
                    return a.U_save + (((size_t)sg.r * nkc + kc) * (size_t)a.NG + (size_t)g0) * USAVE_BLOCK;
                };
                produce_phase<C>(pipe, nb * R, seg_dA, Mpad, ufn);
                produce_phase<C>(pipe, nb, seg_dK, Mpad);
            }
        }
        return;
    }
#ifdef DGP_PROFILE_PHASES
    long long ph[12] = {0}; long long tlast = clock64();
#endif

    for (int gt = blockIdx.x; gt < total_tiles; gt += gridDim.x) {
        const BwdArgs& a = bwd_member(p, gt / ntiles);
        cur = &a;
        const dgp_cond_desc& d = a.d;
        const int thd = 2 + d.D;
        const int tile = gt % ntiles;
        // this CTA's partial buffers of this member (accumulated across tiles and across calls; zeroed by dgp_cond_prepare)
        double* pT2z = a.part + (size_t)blockIdx.x * a.npart;   // [Mpad][Dp]
        double* pcs = pT2z + (size_t)Mpad * Dp;                  // [Mpad]
        double* pqmu = pcs + Mpad;                               // [Mpad][R]
        double* pth = pqmu + (size_t)Mpad * R;                   // [T][2+Dp]
        double* pmA = pth + (size_t)d.T * (2 + Dp);              // [Dx][R] Linear mean weights: X^T g_mean
        double* pmb = pmA + (size_t)d.Dx * R;                    // [R]     Linear mean offset: sum_n g_mean
        const long long n0 = a.row0 + (long long)tile * C::NT;
        const int nvalid = (int)((a.n_rows - n0 < C::NT) ? (a.n_rows - n0) : C::NT);
        bulk_wait_read0();   // previous tile's dKuf store has finished reading Dt
        consumer_sync<C>();  // every warp is done with the previous tile
        PH(0);

        // ---- per-row setup: scaled inputs (one thread per (row, input column): the fp64 divisions of a row do not queue up
        //      behind each other in one thread), combined upstream gradients
        for (int idx = tid; idx < C::NT * Dp; idx += C::NCONS) {
            const int n = idx / Dp, k = idx - n * Dp;
            long long gn = n0 + n;
            if (gn >= a.n_rows) gn = a.n_rows - 1;
            const double* xr = a.X + (size_t)(gn % a.x_rows) * d.Dx;
            int tk = 0;
            if (d.T > 1) {
                tk = (int)xr[d.Dx - 1];
                if (tk < 0 || tk >= d.T) tk = -1;
            }
            if (k == 0) xtask[n] = tk;
            const double* th = d.theta + (size_t)(tk < 0 ? 0 : tk) * thd;
            xs[n * xld + k] = (k < d.D) ? xr[k] / th[2 + k] : 0.0;
        }
        for (int idx = tid; idx < C::NT * R; idx += C::NCONS) {
            const int n = idx / R, r = idx % R;
            const long long gn = n0 + n;
            double gmean = 0.0, gvar = 0.0;
            if (gn < a.n_rows) {
                const size_t o = (size_t)gn * d.ldr + d.col0 + r;
                if (a.g_mean_in) gmean = a.g_mean_in[o];
                if (a.g_var_in) gvar = a.g_var_in[o];
                if (a.g_F) {
                    const double gf = a.g_F[(size_t)gn * a.ldgf + d.col0 + r];
                    gmean += gf;
                    gvar += gf * (a.F[(size_t)gn * a.ldf + d.col0 + r] - a.mean[o]) / (2.0 * a.var[o]);
                }
                a.gmv[(size_t)gn * 2 * R + r] = gmean;
                a.gmv[(size_t)gn * 2 * R + R + r] = gvar;
            }
            gm[idx] = gmean;
            gv[idx] = gvar;
        }
        consumer_sync<C>();
        for (int n = tid; n < C::NT; n += C::NCONS) {      // sum_r g_var[n, r]: weight of the -2 A term (and of Kdiag)
            double s = 0.0;
            for (int r = 0; r < R; ++r) s += gv[n * R + r];
            gs[n] = s;
            double s2 = 0.0;                                // squared norm of the scaled input row
            for (int k = 0; k < Dp; ++k) s2 = fma(xs[n * xld + k], xs[n * xld + k], s2);
            xn[n] = s2;
        }
        // ---- Dt <- q_mu g_mean^T (the mean path of dA) on the tensor cores, K = R: the dA tile is free until the epilogue
        //      of the streamed phase adds the variance path to it.  (As a scalar loop in that epilogue -- q_mu rows from
        //      global memory behind every shared-memory store -- this term cost 9 % of the kernel at R = 8, 13 % at R = 1.)
        for (int mt = warp; mt < Mpad / 8; mt += C::NWARPS) {
            const int m = mt * 8 + g;
            const double* qrow = d.q_mu + (size_t)m * d.ldr + d.col0;
            for (int r0 = 0; r0 < R; r0 += 16) {
                double aq[4];
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const int k = r0 + ks * 4 + t;
                    aq[ks] = (m < d.M && k < R) ? qrow[k] : 0.0;
                }
#pragma unroll
                for (int nt = 0; nt < C::NT / 8; ++nt) {
                    double c[2] = {0.0, 0.0};
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const int k = r0 + ks * 4 + t;
                        if (r0 + ks * 4 < R) dmma884(c, aq[ks], k < R ? gm[(nt * 8 + g) * R + k] : 0.0);
                    }
                    double* dst = Dt + (nt * 8 + 2 * t) * ldb + m;
                    if (r0 == 0) { dst[0] = c[0]; dst[ldb] = c[1]; }
                    else { dst[0] += c[0]; dst[ldb] += c[1]; }
                }
            }
        }
        PH(1);
        // (run_phase starts with a consumer barrier)

        // ---- phase dA: Dt = q_mu g_mean^T - 2 A . gs + 2 sum_r Lq_r (U_r . g_var_r)
        {
            // The weights g_var[n, r] scale the streamed U fragments (one DMUL per B fragment, placed behind the previous
            // DMMA group: chunk_mma), so that ONE accumulator set runs over all r of a block row.  Keeping the GEMM
            // unweighted and folding w * acc into a second register set per (block row, r) left too few of the 96 registers
            // of a 17-warp CTA for the software pipeline of the streamed loop (ptxas put every LDS right in front of its
            // DMMA: DMMA pipe 59 % active); accumulating that sum in the shared-memory tile was slower still.
            // Columns beyond the last valid row get weight 0; their ring bytes are stale but finite (zeroed at start).
            auto epi = [&](auto& acc, const Seg& sg) {
                PH(3);
                // A[n][row] for every element this thread owns, straight from L2 (prefetched by the producer warp): all
                // 8 NJ loads are issued before the first use, one exposed round trip per block row
                double av[4][C::NJ][2];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int row = acc_row<C>(sg.i, q);
#pragma unroll
                    for (int j = 0; j < C::NJ; ++j)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int n = acc_col<C>(j, e);
                            av[q][j][e] = (n < nvalid && row < Mpad) ? __ldg(a.A_save + (size_t)(n0 + n) * Mpad + row) : 0.0;
                        }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int row = acc_row<C>(sg.i, q);
                    if (row < Mpad) {
#pragma unroll
                        for (int j = 0; j < C::NJ; ++j)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int n = acc_col<C>(j, e);
                                double* dst = Dt + n * ldb + row;       // holds (q_mu g_mean^T)[row][n] from the prologue
                                *dst = fma(2.0, acc.v[q][j][e] - av[q][j][e] * gs[n], *dst);
                            }
                    }
                }
                PH(10);
            };
#ifdef DGP_PROFILE_PHASES
            consumer_sync<C>();
            PH(11);
#endif
#ifdef DGP_TEST_HOOKS
            if (a.debug == 4)        // timing experiment: B fragments from the resident tile (the forward kernel's pattern)
                run_phase<C, MASK_LOWER, false, true>(pipe, nb * R, seg_dA, epi, Dt, ldb, Mpad, gv, R, nvalid);
            else
#endif
            run_phase<C, MASK_LOWER, true, true>(pipe, nb * R, seg_dA, epi, nullptr, 0, Mpad, gv, R, nvalid);
        }
        PH(3);

        // ---- phase dK: Dt <- Lm^-T Dt, block rows top-down, in place
        {
            auto epi = [&](auto& acc, const Seg& sg) {
                consumer_sync<C>();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int row = acc_row<C>(sg.i, q);
                    if (row < Mpad) {
#pragma unroll
                        for (int j = 0; j < C::NJ; ++j) {
                            Dt[acc_col<C>(j, 0) * ldb + row] = acc.v[q][j][0];
                            Dt[acc_col<C>(j, 1) * ldb + row] = acc.v[q][j][1];
                        }
                    }
                }
            };
            run_phase<C, MASK_UPPER, false>(pipe, nb, seg_dK, epi, Dt, ldb, Mpad);
        }
        consumer_sync<C>();
        PH(4);
#ifdef DGP_TEST_HOOKS
        if (a.debug == 3) continue;      // timing experiment: GEMM phases only
#endif

        // ---- dKuf tile -> HBM (for the Gram kernel)
        fence_async_smem();
        if (tid < nvalid) {
            bulk_s2g(a.dKuf + (size_t)(n0 + tid) * Mpad, Dt + tid * ldb, (uint32_t)(Mpad * sizeof(double)));
            bulk_commit();
        }
        // ---- d q_mu partial: sum_n A[n][m] g_mean[n][r] on the tensor cores while the store drains.  Output tile rows =
        //      inducing index m, columns = output r, K = the NT rows of the tile; the A fragments come straight from L2.
        for (int mt = warp; mt < Mpad / 8; mt += C::NWARPS) {
            const int m0 = mt * 8;
            double af[C::NT / 4];
#pragma unroll
            for (int ks = 0; ks < C::NT / 4; ++ks) {
                const int n = ks * 4 + t;
                af[ks] = (n < nvalid) ? a.A_save[(size_t)(n0 + n) * Mpad + m0 + g] : 0.0;
            }
            for (int r0 = 0; r0 < R; r0 += 8) {
                double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
                // this CTA's running partial (global, L2): fetched before the DMMA chain so that its round trip is hidden
                double pold[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int m = m0 + g, r = r0 + 2 * t + e;
                    pold[e] = (m < d.M && r < R) ? pqmu[m * R + r] : 0.0;
                }
#pragma unroll
                for (int ks = 0; ks < C::NT / 4; ++ks) {
                    const int n = ks * 4 + t;
                    const double bv = (r0 + g < R) ? gm[n * R + r0 + g] : 0.0;
                    if (ks & 1) dmma884(c1, af[ks], bv); else dmma884(c0, af[ks], bv);
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int m = m0 + g, r = r0 + 2 * t + e;
                    if (m < d.M && r < R) pqmu[m * R + r] = pold[e] + (c0[e] + c1[e]);
                }
            }
        }
        PH(5);
        bulk_wait_read0();
        consumer_sync<C>();
        PH(6);

        // ---- Dt <- Gr2 = dKuf . variance . k'(r2);  per-row sums: rs = sum_m Gr2, dv = sum_m dKuf k(r2)/variance
        const bool saved_k = a.k_rbf != 0;
        if (saved_k) {
            // RBF with Kuf kept by the forward pass: k'(r2) = -k/2, so Gr2 = -dKuf Kuf / 2 and dv = sum dKuf Kuf / variance
            // need no exp, no cross term and no inducing inputs: one coalesced pass, a warp per row, 32 independent loads
            // per lane in flight.  (Recomputing exp(-r2/2) for the 16 k elements of a tile was 7 % of the kernel at R = 8
            // and 17 % at R = 1: latency-bound chains, five times the fp64-pipe time of the arithmetic.)
            // (a SwitchedKernel of RBFs: Kuf is already zero where the tasks of row and inducing point differ)
            for (int n = warp; n < C::NT; n += C::NWARPS) {
                double srs = 0.0, sdv = 0.0;
                const int tx = xtask[n];
                const double ivar = (tx >= 0) ? 1.0 / d.theta[(size_t)tx * thd] : 0.0;
                if (n < nvalid) {
                    const double* krow = a.K_save + (size_t)(n0 + n) * Mpad;
                    for (int m0 = 0; m0 < Mpad; m0 += 256) {
                        double kv[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int m = m0 + u * 32 + lane;
                            kv[u] = (m < Mpad) ? __ldcs(krow + m) : 0.0;      // read once: streaming load
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int m = m0 + u * 32 + lane;
                            if (m < Mpad) {
                                const double gk = Dt[n * ldb + m];
                                const double out = -0.5 * gk * kv[u];
                                sdv = fma(gk, kv[u], sdv);
                                srs += out;
                                Dt[n * ldb + m] = out;
                            }
                        }
                    }
                } else {
                    for (int m = lane; m < Mpad; m += 32) Dt[n * ldb + m] = 0.0;
                }
                srs = warp_sum(srs);
                sdv = warp_sum(sdv);
                if (lane == 0) { rs[n] = srs; dv[n] = sdv * ivar; }
            }
        } else if (d.T == 1) {
            // fast path (one kernel for every row): cross term z.x on the tensor cores as in the forward pass, the
            // elementwise part in the accumulator layout (8 independent exp chains per thread); rs comes out of the T1
            // contraction below (ones column), and only the tile total of dv is needed (all rows share the task).
            const double varn = d.theta[0];
            const int kind = d.kind[0];
            double sdv = 0.0;
            for (int mt = warp; mt < Mpad / 8; mt += C::NWARPS) {
                const int m = mt * 8 + g;
                const double znm = a.zn[m];
                const double* zrow = a.Zs + (size_t)m * Dp;
                double az[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) az[kk] = (kk * 4 < Dp) ? zrow[kk * 4 + t] : 0.0;
#pragma unroll
                for (int nt = 0; nt < C::NT / 8; ++nt) {
                    double c[2] = {0.0, 0.0};
                    if (Dp <= 16) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            if (kk * 4 < Dp) dmma884(c, az[kk], xs[(nt * 8 + g) * xld + kk * 4 + t]);
                    } else {
                        for (int kk = 0; kk < Dp / 4; ++kk) dmma884(c, zrow[kk * 4 + t], xs[(nt * 8 + g) * xld + kk * 4 + t]);
                    }
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int n = nt * 8 + 2 * t + e;
                        double out = 0.0;
                        if (m < d.M && n < nvalid) {
                            const double r2 = znm + xn[n] - 2.0 * c[e];
                            double kv, dk;
                            kern_val_grad(kind, r2, kv, dk);
                            const double gk = Dt[n * ldb + m];
                            sdv = fma(gk, kv, sdv);
                            out = gk * varn * dk;
                        }
                        Dt[n * ldb + m] = out;
                    }
                }
            }
            sdv = warp_sum(sdv);
            if (lane == 0) rowq[warp] = sdv;          // rowq is free here; reduced in warp order below
            consumer_sync<C>();
            if (tid < C::NT) {
                double tot = 0.0;
                if (tid == 0)
                    for (int w = 0; w < C::NWARPS; ++w) tot += rowq[w];
                dv[tid] = tot;                          // the whole tile's sum is booked on row 0
            }
        } else {
            for (int n = warp; n < C::NT; n += C::NWARPS) {
                const int tx = xtask[n];
                double srs = 0.0, sdv = 0.0;
                const double varn = (tx >= 0) ? d.theta[(size_t)tx * thd] : 0.0;
                const int kind = d.kind[tx < 0 ? 0 : tx];
                for (int m = lane; m < Mpad; m += 32) {
                    double out = 0.0;
                    if (n < nvalid && m < d.M && tx >= 0 && a.ztask[m] == tx) {
                        const double* zr = a.Zs + (size_t)m * Dp;
                        double c = 0.0;
                        for (int k = 0; k < Dp; ++k) c = fma(zr[k], xs[n * xld + k], c);
                        const double r2 = a.zn[m] + xn[n] - 2.0 * c;
                        double kv, dk;
                        kern_val_grad(kind, r2, kv, dk);
                        const double gk = Dt[n * ldb + m];
                        sdv = fma(gk, kv, sdv);
                        out = gk * varn * dk;
                        srs += out;
                    }
                    Dt[n * ldb + m] = out;
                }
                srs = warp_sum(srs);
                sdv = warp_sum(sdv);
                if (lane == 0) { rs[n] = srs; dv[n] = sdv; }
            }
        }
        consumer_sync<C>();

        PH(7);
        // ---- T1[n][k] = sum_m Gr2[n][m] zs[m][k] on the tensor cores, one 8-row tile of the row tile per warp, four
        //      interleaved accumulators over the inducing index; parked in rowq[n][2 + k].  Column Dp of the product is a
        //      column of ones: the row sums of Gr2, parked in rowq[n][0].
        const int t1_ones = (d.T == 1 && !saved_k) ? 1 : 0;     // the other paths have the row sums already
        // (T1 has NT/8 units of Mpad/4 DMMAs each, T2z Mpad/8 units of NT/4: the first NT/8 warps take T1, the others
        //  share the column sums and T2z, so that both finish at about the same time)
        constexpr int T1W = (C::NT / 8 < C::NWARPS) ? C::NT / 8 : 0;      // warps that do T1 only (0: everyone does both)
        constexpr int T2W = C::NWARPS - T1W;
        for (int nt4 = warp; nt4 < C::NT / 8; nt4 += C::NWARPS) {
            for (int d0 = 0; d0 < Dp + t1_ones; d0 += 8) {
                double c4[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
                const int kcol = d0 + g;
                for (int m = 0; m < Mpad; m += 64) {          // Mpad is a multiple of 64; 16 global loads in flight
                    double bv[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const int mk = m + 4 * u + t;
                        bv[u] = (kcol < Dp) ? __ldg(a.Zs + (size_t)mk * Dp + kcol) : (kcol == Dp ? 1.0 : 0.0);
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u) dmma884(c4[u & 3], Dt[(nt4 * 8 + g) * ldb + m + 4 * u + t], bv[u]);
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = d0 + 2 * t + e;
                    const double v = (c4[0][e] + c4[1][e]) + (c4[2][e] + c4[3][e]);
                    if (k < Dp) rowq[(nt4 * 8 + g) * nq + 2 + k] = v;
                    else if (k == Dp) rowq[(nt4 * 8 + g) * nq] = v;
                }
            }
        }
        // ---- cs[m] = sum_n Gr2[n][m]: plain column sums, one thread per inducing point (as a ninth column of the T2z
        //      product below it cost a whole extra pass of DMMAs whenever Dp is a multiple of 8)
        for (int m = tid - T1W * 32; m < Mpad; m += T2W * 32) {
            if (m < 0) break;                               // a T1 warp
            const double pold = (m < d.M) ? pcs[m] : 0.0;
            double s4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
            for (int n = 0; n < C::NT; n += 4) {
                s4[0] += Dt[n * ldb + m];
                s4[1] += Dt[(n + 1) * ldb + m];
                s4[2] += Dt[(n + 2) * ldb + m];
                s4[3] += Dt[(n + 3) * ldb + m];
            }
            if (m < d.M) pcs[m] = pold + ((s4[0] + s4[1]) + (s4[2] + s4[3]));
        }
        // ---- T2z[m][k] = sum_n Gr2[n][m] xs[n][k]
        for (int mt = warp - T1W; mt < Mpad / 8; mt += T2W) {
            if (mt < 0) break;                              // a T1 warp
            const int m0 = mt * 8;
            for (int d0 = 0; d0 < Dp; d0 += 8) {
                double c[2] = {0.0, 0.0}, c2[2] = {0.0, 0.0};
                const int kcol = d0 + g;
                double pold[2];          // running partials of this CTA (global, L2): fetched ahead of the DMMA chain
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int m = m0 + g, k = d0 + 2 * t + e;
                    pold[e] = (m < d.M && k < Dp) ? pT2z[m * Dp + k] : 0.0;
                }
#pragma unroll
                for (int ks = 0; ks < C::NT / 4; ++ks) {
                    const int n = ks * 4 + t;
                    const double bv = (kcol < Dp) ? xs[n * xld + kcol] : 0.0;
                    if (ks & 1) dmma884(c2, Dt[n * ldb + m0 + g], bv); else dmma884(c, Dt[n * ldb + m0 + g], bv);
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int m = m0 + g, k = d0 + 2 * t + e;
                    if (m < d.M && k < Dp) pT2z[m * Dp + k] = pold[e] + (c[e] + c2[e]);
                }
            }
        }
        consumer_sync<C>();
        PH(8);
        // ---- row sums of Gr2 from the ones column (fast path; the other paths computed them above)
        if (d.T == 1 && !saved_k) {
            if (tid < C::NT) rs[tid] = rowq[tid * nq];
            consumer_sync<C>();
        }
        // ---- finish T1 -> gX and the lengthscale row terms (in place in rowq[n][2 + k])
        for (int idx = tid; idx < C::NT * Dp; idx += C::NCONS) {
            const int n = idx / Dp, k = idx % Dp;
            const double t1 = rowq[n * nq + 2 + k];
            const double xv = xs[n * xld + k];
            rowq[n * nq + 2 + k] = xv * xv * rs[n] - 2.0 * xv * t1;
            const long long gn = n0 + n;
            if (a.gX && gn < a.n_rows && k < d.D) {
                const int tx = xtask[n];
                double gx = 0.0;
                if (tx >= 0) gx = 2.0 * (xv * rs[n] - t1) / d.theta[(size_t)tx * thd + 2 + k];
                if (d.mean_A)
                    for (int r = 0; r < R; ++r) gx = fma(gm[n * R + r], d.mean_A[(size_t)k * d.ldr + d.col0 + r], gx);
                double* o = a.gX + (size_t)gn * d.Dx + k;
                *o = a.accumulate_gX ? (*o + gx) : gx;
            }
        }
        // columns of X the kernel does not see (task id): only the Linear mean reaches them
        if (a.gX && d.Dx > d.D) {
            for (int idx = tid; idx < C::NT * (d.Dx - d.D); idx += C::NCONS) {
                const int n = idx / (d.Dx - d.D), k = d.D + idx % (d.Dx - d.D);
                const long long gn = n0 + n;
                if (gn < a.n_rows) {
                    double gx = 0.0;
                    if (d.mean_A)
                        for (int r = 0; r < R; ++r) gx = fma(gm[n * R + r], d.mean_A[(size_t)k * d.ldr + d.col0 + r], gx);
                    double* o = a.gX + (size_t)gn * d.Dx + k;
                    *o = a.accumulate_gX ? (*o + gx) : gx;
                }
            }
        }
        // trainable Linear mean (an OutputLayer built with its own Linear mean, dgplib/layers.py:211-218): mean += X A + b, so
        // dA = X^T g_mean over the raw input columns and db = sum_n g_mean; a few thousand FMAs per tile, rows in order
        if (a.mean_grad) {
            for (int idx = tid; idx < (d.Dx + 1) * R; idx += C::NCONS) {
                const int k = idx / R, r = idx % R;
                double s2 = 0.0;
                for (int n = 0; n < nvalid; ++n) {
                    const double xv = (k < d.Dx) ? a.X[(size_t)((n0 + n) % a.x_rows) * d.Dx + k] : 1.0;
                    s2 = fma(xv, gm[n * R + r], s2);
                }
                if (k < d.Dx) pmA[k * R + r] += s2;
                else pmb[r] += s2;
            }
        }
        // Kdiag terms: d variance += sum_r g_var, d white += sum_r g_var
        for (int n = tid; n < C::NT; n += C::NCONS) {
            const double s = (n < nvalid) ? gs[n] : 0.0;
            rowq[n * nq + 0] = dv[n] + s;
            rowq[n * nq + 1] = s;
        }
        consumer_sync<C>();
        // per-task hyper-parameter row terms: one warp per (task, column), lanes over the rows, fixed shuffle tree
        // (deterministic); the global partial sees one read-modify-write per tile
        for (int idx = warp; idx < d.T * nq; idx += C::NWARPS) {
            const int tk = idx / nq, c = idx % nq;
            const double pold = (lane == 0) ? pth[idx] : 0.0;
            double s = 0.0;
            for (int n = lane; n < nvalid; n += 32)
                if (xtask[n] == tk) s += rowq[n * nq + c];
            s = warp_sum(s);
            if (lane == 0) pth[idx] = pold + s;
        }
        PH(9);
    }
    bulk_wait_read0();
#ifdef DGP_PROFILE_PHASES
    if (blockIdx.x == 0 && tid == 0) for (int i = 0; i < 12; ++i) a0.scal[i] = (double)ph[i];
#endif
}

template <class C>
static int launch_bwd(const BwdArgs& a, cudaStream_t st) {
    const BwdSmem<C> L(a.Mpad, a.Dp, a.d.R);
    const size_t bytes = (size_t)L.total_doubles * sizeof(double);
    if (bytes > 227 * 1024) return DGP_ERR_UNSUPPORTED;
    static DevOnceSize configured;
    const int dev = current_device();
    if (!configured.covers(dev, bytes)) {
        if (cudaFuncSetAttribute(k_cond_bwd<C, BwdArgs>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return DGP_ERR_LAUNCH;
        configured.raise(dev, bytes);
    }
    const int grid = a.ntiles < device_sm_count() ? a.ntiles : device_sm_count();
    k_cond_bwd<C, BwdArgs><<<grid, C::NTHREADS, bytes, st>>>(a);
    DGP_COUNT_LAUNCH();
    return cudaGetLastError() == cudaSuccess ? DGP_OK : DGP_ERR_LAUNCH;
}

template <class C>
static int launch_bwd_group(const BwdGroup& grp, cudaStream_t st) {
    const BwdSmem<C> L(grp.m[0].Mpad, grp.m[0].Dp, grp.m[0].d.R);
    const size_t bytes = (size_t)L.total_doubles * sizeof(double);
    if (bytes > 227 * 1024) return DGP_ERR_UNSUPPORTED;
    static DevOnceSize configured;
    const int dev = current_device();
    if (!configured.covers(dev, bytes)) {
        if (cudaFuncSetAttribute(k_cond_bwd<C, BwdGroup>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return DGP_ERR_LAUNCH;
        configured.raise(dev, bytes);
    }
    const long long total = (long long)grp.m[0].ntiles * grp.n;
    const int grid = (int)(total < device_sm_count() ? total : device_sm_count());
    k_cond_bwd<C, BwdGroup><<<grid, C::NTHREADS, bytes, st>>>(grp);
    DGP_COUNT_LAUNCH();
    return cudaGetLastError() == cudaSuccess ? DGP_OK : DGP_ERR_LAUNCH;
}

// ------------------------------------------------------------------------------------------------ Gram kernel
// out[s][mat][m1][m2] += sum_{n in split s} w_mat[n] P_mat[n][m1] A[n][m2]   for 64x64 tiles with tile(m1) >= tile(m2)
//   (inside diagonal tiles only the 8x8 tiles on and below the tile diagonal are computed and consumed)
//   mat <  R : P = A,    w = g_var[:, mat]      (G_r, symmetric)
//   mat == R : P = dKuf, w = 1                  (P = dKuf^T A)
struct GramArgs {
    int Mpad, R, nsplit;
    long long n_rows, rows_per_split;
    const double* A; const double* dKuf; const double* gmv;
    double* gram;
    int accumulate;   // 0: this call owns the partial buffers (store), != 0: add to them (further row chunks)
};
#ifndef DGP_GR_KC
#define DGP_GR_KC 16
#endif
#ifndef DGP_GR_STAGES
#define DGP_GR_STAGES 3
#endif
constexpr int GR_THREADS = 128, GR_STAGES = DGP_GR_STAGES, GR_KC = DGP_GR_KC, GR_LD = 68;

// Diagonal 64x64 tile: only the 36 lower 8x8 tiles are needed.  Warp W takes tile rows W and 7-W (W+1 and 8-W tiles:
// nine per warp, balanced), with compile-time bounds so that nothing is predicated off (a predicated-off DMMA still
// occupies the fp64 pipe).  9/16 of the work of a full tile.
template <int W>
__device__ __forceinline__ void gram_diag_chunk(double (&accA)[8][2], double (&accB)[8][2], const double* __restrict__ P,
                                                const double* __restrict__ Q, const double* __restrict__ Wt, int g, int t) {
#pragma unroll
    for (int kk = 0; kk < GR_KC / 4; ++kk) {
        const double w = Wt[kk * 4 + t];
        const double a0 = P[(kk * 4 + t) * GR_LD + W * 8 + g] * w;
        const double a1 = P[(kk * 4 + t) * GR_LD + (7 - W) * 8 + g] * w;
#pragma unroll
        for (int j = 0; j <= 7 - W; ++j) {
            const double b = Q[(kk * 4 + t) * GR_LD + j * 8 + g];
            if (j <= W) dmma884(accA[j], a0, b);
            dmma884(accB[j], a1, b);
        }
    }
}
template <int W>
__device__ __forceinline__ void gram_diag_store(const double (&accA)[8][2], const double (&accB)[8][2], double* out, int Mpad,
                                                int base, int g, int t, bool add) {
#pragma unroll
    for (int j = 0; j <= 7 - W; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int m2 = base + j * 8 + 2 * t + e;
            if (j <= W) {
                double* o = out + (size_t)(base + W * 8 + g) * Mpad + m2;
                *o = add ? (*o + accA[j][e]) : accA[j][e];
            }
            double* o = out + (size_t)(base + (7 - W) * 8 + g) * Mpad + m2;
            *o = add ? (*o + accB[j][e]) : accB[j][e];
        }
}

__global__ void __launch_bounds__(GR_THREADS, 4) k_gram(GramArgs a) {
    extern __shared__ __align__(128) double gsm[];
    double (*Ps)[GR_KC * GR_LD] = reinterpret_cast<double (*)[GR_KC * GR_LD]>(gsm);
    double (*Qs)[GR_KC * GR_LD] = reinterpret_cast<double (*)[GR_KC * GR_LD]>(gsm + GR_STAGES * GR_KC * GR_LD);
    double (*wsm)[GR_KC] = reinterpret_cast<double (*)[GR_KC]>(gsm + 2 * GR_STAGES * GR_KC * GR_LD);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    // warp w owns rows 16 w .. 16 w + 15 of the 64x64 tile and all 64 columns: two A fragments (scaled by the row
    // weight: 2 DMUL per 16 DMMA) against eight B fragments per k-step
    const int wm = warp;
    // tile pair from the linear index over the lower-triangular tile set
    int p = blockIdx.x, bi = 0;
    while ((bi + 1) * (bi + 2) / 2 <= p) ++bi;
    const int bj = p - bi * (bi + 1) / 2;
    const int mat = blockIdx.y, split = blockIdx.z;
    const int Mpad = a.Mpad, R = a.R;
    const double* Psrc = (mat < R ? a.A : a.dKuf) + bi * 64;
    const double* Qsrc = a.A + bj * 64;
    const long long nbeg = (long long)split * a.rows_per_split;
    long long nend = nbeg + a.rows_per_split;
    if (nend > a.n_rows) nend = a.n_rows;
    const int nchunks = nend > nbeg ? (int)((nend - nbeg + GR_KC - 1) / GR_KC) : 0;
    const bool diag = (bi == bj);     // diagonal tiles compute the lower 8x8 tiles only (gram_diag_chunk)
    constexpr int jmax = 8;

    double acc[2][8][2];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[q][j][0] = 0.0; acc[q][j][1] = 0.0; }

    auto issue = [&](int c, int stage) {
        if (c < nchunks) {
            const long long r0 = nbeg + (long long)c * GR_KC;
#pragma unroll
            for (int q = 0; q < GR_KC / 4; ++q) {
                const int e = tid + q * GR_THREADS;  // row = e/32, 16-byte column = e%32
                const int row = e >> 5, cc = e & 31;
                long long n = r0 + row;
                if (n >= nend) n = nend - 1;  // clamped rows get weight 0
                cp_async16(&Ps[stage][row * GR_LD + cc * 2], Psrc + (size_t)n * Mpad + cc * 2);
                cp_async16(&Qs[stage][row * GR_LD + cc * 2], Qsrc + (size_t)n * Mpad + cc * 2);
            }
            if (tid < GR_KC) {
                const long long n = r0 + tid;
                double w = 0.0;
                if (n < nend) w = (mat < R) ? a.gmv[(size_t)n * 2 * R + R + mat] : 1.0;
                wsm[stage][tid] = w;
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s0 = 0; s0 < GR_STAGES - 1; ++s0) issue(s0, s0);
    int stage = 0, lstage = GR_STAGES - 1;
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<GR_STAGES - 2>();
        __syncthreads();
        issue(c + GR_STAGES - 1, lstage);
        lstage = (lstage + 1 == GR_STAGES) ? 0 : lstage + 1;
        const double* P = Ps[stage];
        const double* Q = Qs[stage];
        const double* W = wsm[stage];
        stage = (stage + 1 == GR_STAGES) ? 0 : stage + 1;
        if (diag) {
            switch (wm) {      // acc[0] = tile row W, acc[1] = tile row 7-W
                case 0: gram_diag_chunk<0>(acc[0], acc[1], P, Q, W, g, t); break;
                case 1: gram_diag_chunk<1>(acc[0], acc[1], P, Q, W, g, t); break;
                case 2: gram_diag_chunk<2>(acc[0], acc[1], P, Q, W, g, t); break;
                default: gram_diag_chunk<3>(acc[0], acc[1], P, Q, W, g, t); break;
            }
            continue;
        }
#pragma unroll
        for (int kk = 0; kk < GR_KC / 4; ++kk) {
            const double w = W[kk * 4 + t];
            double af[2], bf[8];
#pragma unroll
            for (int q = 0; q < 2; ++q) af[q] = P[(kk * 4 + t) * GR_LD + wm * 16 + q * 8 + g] * w;
            if (jmax == 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) bf[j] = Q[(kk * 4 + t) * GR_LD + j * 8 + g];
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int j = 0; j < 8; ++j) dmma884(acc[q][j], af[q], bf[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = Q[(kk * 4 + t) * GR_LD + j * 8 + g];
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[q][j], af[q], bf[j]);
            }
        }
    }
    cp_async_wait<0>();
    const bool add = a.accumulate != 0;
    if (nchunks == 0 && add) return;      // an empty split still has to store zeros when it owns the buffer
    double* out = a.gram + ((size_t)split * (R + 1) + mat) * Mpad * Mpad;
    if (diag) {
        switch (wm) {
            case 0: gram_diag_store<0>(acc[0], acc[1], out, Mpad, bi * 64, g, t, add); break;
            case 1: gram_diag_store<1>(acc[0], acc[1], out, Mpad, bi * 64, g, t, add); break;
            case 2: gram_diag_store<2>(acc[0], acc[1], out, Mpad, bi * 64, g, t, add); break;
            default: gram_diag_store<3>(acc[0], acc[1], out, Mpad, bi * 64, g, t, add); break;
        }
        return;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (j >= jmax) continue;
                const int m1 = bi * 64 + wm * 16 + q * 8 + g;
                const int m2 = bj * 64 + j * 8 + 2 * t + e;
                double* o = out + (size_t)m1 * Mpad + m2;
                *o = add ? (*o + acc[q][j][e]) : acc[q][j][e];
            }
}

#ifdef DGP_TEST_HOOKS
#define DGP_STAGE_OK(STG) (g_force_stages == 0 || g_force_stages == (STG))
#else
#define DGP_STAGE_OK(STG) true
#endif

// rows [row0, row_end) of the call with NT-row tiles (row0 a multiple of 8: the U blocks hold 8 rows)
static int bwd_width(BwdArgs a, int nt, long long row0, long long row_end, cudaStream_t st) {
    int rc = DGP_ERR_UNSUPPORTED;
    a.row0 = row0;
    a.row_end = row_end;
#define DGP_TRY_BWD(WNV, NTV, STG)                                                        \
    if (DGP_STAGE_OK(STG)) {                                                          \
        using C = TileCfg<4, WNV, NTV, STG, true>;                                        \
        a.ntiles = (int)((row_end - row0 + C::NT - 1) / C::NT);                       \
        rc = launch_bwd<C>(a, st);                                                    \
        if (rc != DGP_ERR_UNSUPPORTED) return rc;                                     \
    }
    switch (nt) {
        case 72: DGP_TRY_BWD(3, 72, 2) break;
        case 64: DGP_TRY_BWD(DGP_BWD_WN, 64, 3) DGP_TRY_BWD(DGP_BWD_WN, 64, 2) break;
        case 40: DGP_TRY_BWD(5, 40, 4) DGP_TRY_BWD(5, 40, 3) break;
        case 32: DGP_TRY_BWD(DGP_BWD_WN, 32, 4) DGP_TRY_BWD(DGP_BWD_WN, 32, 3) DGP_TRY_BWD(DGP_BWD_WN, 32, 2) break;
        case 16: DGP_TRY_BWD(2, 16, 4) DGP_TRY_BWD(2, 16, 3) DGP_TRY_BWD(2, 16, 2) break;
        case 8: DGP_TRY_BWD(1, 8, 3) DGP_TRY_BWD(1, 8, 2) break;
        default: break;
    }
#undef DGP_TRY_BWD
    return rc;
}

// the grouped launch with NT-row tiles (every member has the same tile count)
static int bwd_group_width(BwdGroup grp, int nt, cudaStream_t st) {
    int rc = DGP_ERR_UNSUPPORTED;
#define DGP_TRY_BWDG(WNV, NTV, STG)                                                       \
    if (DGP_STAGE_OK(STG)) {                                                          \
        using C = TileCfg<4, WNV, NTV, STG, true>;                                        \
        for (int c = 0; c < grp.n; ++c) {                                             \
            grp.m[c].row0 = 0;                                                        \
            grp.m[c].row_end = grp.m[c].n_rows;                                       \
            grp.m[c].ntiles = (int)((grp.m[c].n_rows + C::NT - 1) / C::NT);           \
        }                                                                             \
        rc = launch_bwd_group<C>(grp, st);                                            \
        if (rc != DGP_ERR_UNSUPPORTED) return rc;                                     \
    }
    switch (nt) {
        case 72: DGP_TRY_BWDG(3, 72, 2) break;
        case 64: DGP_TRY_BWDG(DGP_BWD_WN, 64, 3) DGP_TRY_BWDG(DGP_BWD_WN, 64, 2) break;
        case 32: DGP_TRY_BWDG(DGP_BWD_WN, 32, 4) DGP_TRY_BWDG(DGP_BWD_WN, 32, 3) break;
        case 16: DGP_TRY_BWDG(2, 16, 4) DGP_TRY_BWDG(2, 16, 3) break;
        default: break;
    }
#undef DGP_TRY_BWDG
    return rc;
}

}  // namespace dgp

using namespace dgp;

static int fill_bwd_args(BwdArgs& a, const dgp_cond_desc* d, void* ws, const double* X, int64_t x_rows, int64_t n_rows,
                         const double* A_save, const double* U_save, const double* K_save, const double* mean,
                         const double* var, const double* F, int32_t ldf, const double* g_mean_in, const double* g_var_in,
                         const double* g_F, int32_t ldgf, double* gX, int32_t accumulate_gX, int32_t mean_grad) {
    int rc = check_desc(d);
    if (rc != DGP_OK) return rc;
    if (!ws || !X || !A_save || !U_save || x_rows < 1 || n_rows < 0) return DGP_ERR_ARG;
    if (g_F && (!F || !mean || !var || ldgf < d->ldr || ldf < d->ldr)) return DGP_ERR_ARG;
    if (!g_F && !g_mean_in && !g_var_in) return DGP_ERR_ARG;
    const WsLayout w = ws_layout(*d, n_rows, 1);
    a = BwdArgs{};
    a.d = *d;
    a.Mpad = w.Mpad; a.Dp = w.Dp;
    a.Zs = ws_ptr<double>(ws, w.Zs); a.zn = ws_ptr<double>(ws, w.zn); a.ztask = ws_ptr<int32_t>(ws, w.ztask);
    a.LinvT = ws_ptr<double>(ws, w.LinvTTl); a.LqL = ws_ptr<double>(ws, w.LqTl); a.tl = w.tl;
    a.X = X; a.x_rows = x_rows; a.n_rows = n_rows;
    a.A_save = A_save; a.U_save = U_save; a.NG = usave_groups(n_rows);
    a.K_save = K_save;
    a.k_rbf = K_save != nullptr;
    for (int tk = 0; tk < d->T; ++tk)
        if (d->kind[tk] != DGP_KIND_RBF) a.k_rbf = 0;
    a.mean = mean; a.var = var; a.F = F; a.ldf = ldf;
    a.g_mean_in = g_mean_in; a.g_var_in = g_var_in; a.g_F = g_F; a.ldgf = ldgf;
    a.gX = gX; a.accumulate_gX = accumulate_gX;
    a.mean_grad = (mean_grad && d->mean_A) ? 1 : 0;
    a.dKuf = ws_ptr<double>(ws, w.dKuf); a.gmv = ws_ptr<double>(ws, w.gmv);
    a.part = ws_ptr<double>(ws, w.part); a.npart = w.npart; a.scal = ws_ptr<double>(ws, w.scal);
#ifdef DGP_TEST_HOOKS
    a.debug = g_debug_mode;
#endif
    return DGP_OK;
}

// MultikernelLayer: the backward row kernels of the n conditionals of one layer in ONE launch over (conditional, row tile)
// pairs.  Only without an input gradient (gX accumulates over the members); members must agree in M, D, Dx, R, T, n <= 8;
// otherwise DGP_ERR_UNSUPPORTED (the caller launches them one by one).  dgp_cond_backward_gram / _params stay per member.
extern "C" int dgp_cond_backward_group(int32_t n, const dgp_cond_desc* const* d, void* const* ws, const double* X,
                                       int64_t x_rows, int64_t n_rows, const double* const* A_save,
                                       const double* const* U_save, const double* const* K_save, const double* mean,
                                       const double* var, const double* F, int32_t ldf, const double* g_mean_in,
                                       const double* g_var_in, const double* g_F, int32_t ldgf, int32_t mean_grad,
                                       void* stream) {
    if (n < 1 || !d || !ws || !A_save || !U_save) return DGP_ERR_ARG;
    if (n > DGP_GROUP_MAX) return DGP_ERR_UNSUPPORTED;
    BwdGroup grp{};
    grp.n = n;
    for (int c = 0; c < n; ++c) {
        if (!d[c]) return DGP_ERR_ARG;
        if (d[c]->M != d[0]->M || d[c]->D != d[0]->D || d[c]->Dx != d[0]->Dx || d[c]->R != d[0]->R || d[c]->T != d[0]->T ||
            d[c]->ldr != d[0]->ldr)
            return DGP_ERR_UNSUPPORTED;
    }
    for (int c = 0; c < n; ++c) {
        const int rc = fill_bwd_args(grp.m[c], d[c], ws[c], X, x_rows, n_rows, A_save[c], U_save[c],
                                     K_save ? K_save[c] : nullptr, mean, var, F, ldf, g_mean_in, g_var_in, g_F, ldgf,
                                     nullptr, 0, mean_grad);
        if (rc != DGP_OK) return rc;
    }
    if (n_rows == 0) return DGP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const long long sms = device_sm_count();
    const int Mpad = grp.m[0].Mpad;
    const int cand[4] = {64, 72, 32, 16};
    const double pen[4] = {1.0, 1.15, 1.08, 1.6};
    int order[4] = {0, 1, 2, 3};
    double cost[4];
    for (int i = 0; i < 4; ++i) {
        const long long tiles = ((n_rows + cand[i] - 1) / cand[i]) * n;
        cost[i] = (double)((tiles + sms - 1) / sms) * cand[i] * pen[i];
        if ((Mpad > 256 && cand[i] > 32) || (Mpad > 512 && cand[i] > 16)) cost[i] = 1e300;
    }
    for (int i = 0; i < 4; ++i)
        for (int j = i + 1; j < 4; ++j)
            if (cost[order[j]] < cost[order[i]]) { const int tmp = order[i]; order[i] = order[j]; order[j] = tmp; }
    for (int i = 0; i < 4; ++i) {
        if (cost[order[i]] >= 1e300) break;
        const int rc = bwd_group_width(grp, cand[order[i]], st);
        if (rc != DGP_ERR_UNSUPPORTED) return rc;
    }
    return DGP_ERR_UNSUPPORTED;
}

extern "C" int dgp_cond_backward(const dgp_cond_desc* d, void* ws, const double* X, int64_t x_rows, int64_t n_rows,
                                 const double* A_save, const double* U_save, const double* K_save, const double* mean,
                                 const double* var,
                                 const double* F, int32_t ldf, const double* g_mean_in, const double* g_var_in,
                                 const double* g_F, int32_t ldgf, double* gX, int32_t accumulate_gX, int32_t mean_grad,
                                 void* stream) {
    BwdArgs a{};
    int rc = fill_bwd_args(a, d, ws, X, x_rows, n_rows, A_save, U_save, K_save, mean, var, F, ldf, g_mean_in, g_var_in,
                           g_F, ldgf, gX, accumulate_gX, mean_grad);
    if (rc != DGP_OK) return rc;
    if (n_rows == 0) return DGP_OK;
    const WsLayout w = ws_layout(*d, n_rows, 1);
    cudaStream_t st = (cudaStream_t)stream;
    int nt_force = 0;
#ifdef DGP_TEST_HOOKS
    nt_force = g_force_bwd_nt;      // parity tests pin every instantiation (tests/test_gpu_conditional.py)
#endif
    if (nt_force) return bwd_width(a, nt_force, 0, n_rows, st);
    if (w.Mpad <= 256) {
        // Only dA / dKuf is resident, so the row tile is as wide as the forward kernel's: 64 rows (4 x 4 consumer warps,
        // two column tiles per warp, 3-slot ring), 72 rows (4 x 3 warps, three column tiles, 2-slot ring) when that saves a
        // wave -- 10 000 rows are one wave of 72-row tiles on 148 SMs -- or 64-row tiles for the full waves and one wave of
        // 40- / 32- / 16-row tiles for the rest (common.cuh plan_tiles).
        const TilePlan p = plan_tiles(n_rows, device_sm_count(), 1.15);
        if (p.nt_tail) {
            rc = bwd_width(a, p.nt_main, 0, p.rows_main, st);
            if (rc == DGP_OK) {
                rc = bwd_width(a, p.nt_tail, p.rows_main, n_rows, st);
                if (rc == DGP_ERR_UNSUPPORTED) rc = bwd_width(a, 64, p.rows_main, n_rows, st);
                if (rc == DGP_ERR_UNSUPPORTED) rc = bwd_width(a, 32, p.rows_main, n_rows, st);
                if (rc == DGP_ERR_UNSUPPORTED) rc = bwd_width(a, 16, p.rows_main, n_rows, st);
                return rc;
            }
            if (rc != DGP_ERR_UNSUPPORTED) return rc;
        } else {
            rc = bwd_width(a, p.nt_main, 0, n_rows, st);
            if (rc != DGP_ERR_UNSUPPORTED) return rc;
        }
        rc = bwd_width(a, 64, 0, n_rows, st);
        if (rc != DGP_ERR_UNSUPPORTED) return rc;
    }
    if (w.Mpad <= 512) {
        rc = bwd_width(a, 32, 0, n_rows, st);
        if (rc != DGP_ERR_UNSUPPORTED) return rc;
    }
    rc = bwd_width(a, 16, 0, n_rows, st);
    if (rc != DGP_ERR_UNSUPPORTED) return rc;
    return bwd_width(a, 8, 0, n_rows, st);       // very wide inputs / many outputs next to Mpad = 1024
}

extern "C" int dgp_cond_backward_gram(const dgp_cond_desc* d, void* ws, const double* A_save, int64_t n_rows,
                                      int32_t accumulate, void* stream) {
    int rc = check_desc(d);
    if (rc != DGP_OK) return rc;
    if (!ws || !A_save || n_rows < 0) return DGP_ERR_ARG;
    if (n_rows == 0 && accumulate) return DGP_OK;
    const WsLayout w = ws_layout(*d, n_rows, 1);
    cudaStream_t st = (cudaStream_t)stream;
    GramArgs ga{};
    ga.Mpad = w.Mpad; ga.R = d->R; ga.nsplit = w.nsplit; ga.n_rows = n_rows;
    long long rps = (n_rows + w.nsplit - 1) / w.nsplit;
    ga.rows_per_split = (rps + GR_KC - 1) / GR_KC * GR_KC;
    ga.A = A_save; ga.dKuf = ws_ptr<double>(ws, w.dKuf); ga.gmv = ws_ptr<double>(ws, w.gmv);
    ga.gram = ws_ptr<double>(ws, w.gram);
    ga.accumulate = accumulate;
    const int nbt = w.Mpad / 64;
    const size_t gsm_bytes = (size_t)(2 * GR_STAGES * GR_KC * GR_LD + GR_STAGES * GR_KC) * sizeof(double);
    static DevOnceSize gram_attr;
    const int gdev = current_device();
    if (!gram_attr.covers(gdev, gsm_bytes)) {
        if (cudaFuncSetAttribute(k_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm_bytes) != cudaSuccess) return DGP_ERR_LAUNCH;
        gram_attr.raise(gdev, gsm_bytes);
    }
    k_gram<<<dim3(nbt * (nbt + 1) / 2, d->R + 1, w.nsplit), GR_THREADS, gsm_bytes, st>>>(ga);
    DGP_COUNT_LAUNCH();
    DGP_CHECK_LAUNCH();
    return DGP_OK;
}
