// CTA-level building block shared by the fused forward and backward row-tile kernels:
//
//     C[block-row i][NT columns]  (+)=  W[block-row i][k-range]  x  Bt[NT columns][k-range]^T
//
// W  : an Mpad x Mpad parameter matrix in global memory (L2-resident: Lm^-1, tril(q_sqrt_r)^T, ...), streamed in
//      MB x 16 chunks through a STAGES-deep cp.async ring in shared memory (row pitch 20 doubles = conflict-free
//      A-fragment loads for DMMA.8x8x4);
// Bt : the resident row tile in shared memory, stored [n][k] with pitch Mpad+4 (conflict-free B-fragment loads);
// C  : fp64 tensor-core accumulators in registers.  Warp grid WM x WN; a warp owns 4 interleaved 8-row tiles
//      (rows (q*WM+wm)*8 ..) so that triangular skipping stays balanced across warps, and NJ = NT/(8 WN) column tiles.
//
// A phase is a list of segments (matrix, block row, chunk range); the cp.async pipeline runs across segment
// boundaries, and an epilogue functor consumes the accumulators whenever a segment is flagged `flush`.
#pragma once
#include "common.cuh"

namespace dgp {

enum { MASK_LOWER = 0, MASK_UPPER = 1, MASK_DENSE = 2 };

template <int WM_, int WN_, int NT_, int STAGES_>
struct TileCfg {
    static constexpr int WM = WM_, WN = WN_, NT = NT_, STAGES = STAGES_;
    static constexpr int NWARPS = WM * WN, NTHREADS = NWARPS * 32;
    static constexpr int MB = 32 * WM;         // rows per block row
    static constexpr int NJ = NT / (8 * WN);   // 8-column tiles per warp
    static constexpr int KC = 16, WLD = 20;    // chunk depth and padded pitch of a staged W chunk
    static constexpr int STAGE_DOUBLES = MB * WLD;
    static_assert(NT % (8 * WN) == 0, "NT must be a multiple of 8*WN");
};

struct Seg {
    const double* W;  // matrix base (row-major, pitch Mpad)
    int i;            // block row
    int kc0, kc1;     // chunk range [kc0, kc1)
    int r;            // output index (weights / epilogue bookkeeping)
    bool flush;       // run the epilogue (and clear the accumulators) after this segment
};

template <class C>
struct Acc {
    double v[4][C::NJ][2];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < C::NJ; ++j) { v[q][j][0] = 0.0; v[q][j][1] = 0.0; }
    }
};

// Runs one phase.  WEIGHTED: the B fragment of column n is multiplied by wts[n * wld + seg.r] (backward pass).
template <class C, int MASK, bool WEIGHTED, class SegFn, class EpiFn>
__device__ __forceinline__ void run_phase(int nseg, SegFn segfn, EpiFn epi, double* __restrict__ Wst,
                                          const double* __restrict__ Bt, int ldb, int Mpad,
                                          const double* __restrict__ wts, int wld) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / C::WN, wn = warp % C::WN;

    Acc<C> acc;
    acc.clear();
    __syncthreads();  // every warp is done with the staging ring (previous phase) and Bt is up to date

    // ---- loader state (runs STAGES-1 chunks ahead of the consumer, across segment boundaries)
    int ls = 0;
    Seg lseg = segfn(0);
    int lkc = lseg.kc0;
    auto issue_load = [&](int stage) {
        if (ls < nseg) {
            const int row0 = lseg.i * C::MB;
            const double* src = lseg.W + (size_t)row0 * Mpad + (size_t)lkc * C::KC;
            double* dst = Wst + stage * C::STAGE_DOUBLES;
            for (int e = tid; e < C::MB * 8; e += C::NTHREADS) {
                const int row = e >> 3, c = e & 7;
                if (row0 + row < Mpad) cp_async16(dst + row * C::WLD + c * 2, src + (size_t)row * Mpad + c * 2);
            }
            if (++lkc == lseg.kc1) {
                if (++ls < nseg) {
                    lseg = segfn(ls);
                    lkc = lseg.kc0;
                }
            }
        }
        cp_async_commit();  // committed even when empty so that group counting stays uniform
    };
#pragma unroll
    for (int s = 0; s < C::STAGES - 1; ++s) issue_load(s);

    int stage = 0, lstage = C::STAGES - 1;
    for (int cs = 0; cs < nseg; ++cs) {
        const Seg seg = segfn(cs);
        const int row0 = seg.i * C::MB;
        double wv[C::NJ];
        if (WEIGHTED) {
#pragma unroll
            for (int j = 0; j < C::NJ; ++j) wv[j] = wts[(wn * C::NJ * 8 + j * 8 + g) * wld + seg.r];
        }
        for (int kc = seg.kc0; kc < seg.kc1; ++kc) {
            cp_async_wait<C::STAGES - 2>();
            __syncthreads();
            issue_load(lstage);
            lstage = (lstage + 1 == C::STAGES) ? 0 : lstage + 1;
            const double* Ws = Wst + stage * C::STAGE_DOUBLES;
            stage = (stage + 1 == C::STAGES) ? 0 : stage + 1;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int kbase = kc * C::KC + kk * 4;
                double b[C::NJ];
#pragma unroll
                for (int j = 0; j < C::NJ; ++j) {
                    b[j] = Bt[(wn * C::NJ * 8 + j * 8 + g) * ldb + kbase + t];
                    if (WEIGHTED) b[j] *= wv[j];
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int mt = q * C::WM + wm;
                    const int mrow0 = row0 + mt * 8;
                    bool active = mrow0 < Mpad;
                    if (MASK == MASK_LOWER) active = active && (mrow0 + 7 >= kbase);
                    if (MASK == MASK_UPPER) active = active && (mrow0 <= kbase + 3);
                    if (active) {
                        const double a = Ws[(mt * 8 + g) * C::WLD + kk * 4 + t];
#pragma unroll
                        for (int j = 0; j < C::NJ; ++j) dmma884(acc.v[q][j], a, b[j]);
                    }
                }
            }
        }
        if (seg.flush) {
            epi(acc, seg);
            acc.clear();
        }
    }
    cp_async_wait<0>();
}

// Accumulator element (q, j, e) of the calling thread maps to
//   row    = seg.i * MB + (q * WM + wm) * 8 + g
//   column = wn * NJ * 8 + j * 8 + 2 t + e
template <class C>
__device__ __forceinline__ int acc_row(int i, int q) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    return i * C::MB + (q * C::WM + warp / C::WN) * 8 + (lane >> 2);
}
template <class C>
__device__ __forceinline__ int acc_col(int j, int e) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    return (warp % C::WN) * C::NJ * 8 + j * 8 + 2 * (lane & 3) + e;
}

}  // namespace dgp
