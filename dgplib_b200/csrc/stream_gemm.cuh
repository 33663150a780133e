// CTA-level building block shared by the fused forward and backward row-tile kernels:
//
//     C[block-row i][NT columns]  (+)=  W[block-row i][k-range]  x  Bt[NT columns][k-range]^T
//
// W  : an Mpad x Mpad parameter matrix (Lm^-1, tril(q_sqrt_r)^T, S_r - I, Lm^-T), pre-tiled in global memory by the
//      parameter stage as [block row][16-deep chunk][128 rows][20 doubles] (dgp::retile), so that one chunk is ONE
//      contiguous 20 KB block.  A dedicated producer warp streams the chunks with bulk async copies (TMA engine,
//      cp.async.bulk + mbarrier complete_tx) into a STAGES-deep shared-memory ring; the 20-double pitch makes the
//      DMMA A-fragment loads conflict-free.
// Bt : the resident row tile in shared memory, stored [n][k] with pitch Mpad+4 (conflict-free B-fragment loads);
// C  : fp64 tensor-core accumulators (DMMA.8x8x4) in registers.  Consumer warp grid WM x WN; a warp owns 4
//      interleaved 8-row tiles (rows (q*WM+wm)*8 ..) so that triangular skipping stays balanced across warps, and
//      NJ = NT/(8 WN) column tiles.
//
// Synchronisation is mbarrier-only inside a phase: full[stage] (producer -> consumers, transaction bytes) and
// empty[stage] (one arrival per consumer warp).  There is no CTA-wide barrier in the main loop, so warps drift freely
// and the tensor pipe never drains at chunk boundaries.  A phase is a list of segments (matrix, block row, chunk
// range); the ring position runs across segments, phases and tiles, and the producer runs ahead across all of them.
#pragma once
#include "common.cuh"

namespace dgp {

enum { MASK_LOWER = 0, MASK_UPPER = 1, MASK_DENSE = 2 };

// host: row-major [Mpad][Mpad] (batched) -> tiled layout (defined in prep.cu)
int retile(const double* src, double* dst, int Mpad, int batch, cudaStream_t st);

template <int WM_, int WN_, int NT_, int STAGES_>
struct TileCfg {
    static constexpr int WM = WM_, WN = WN_, NT = NT_, STAGES = STAGES_;
    static constexpr int NWARPS = WM * WN;              // consumer warps
    static constexpr int NCONS = NWARPS * 32;           // consumer threads
    static constexpr int NTHREADS = NCONS + 32;         // + one producer warp
    static constexpr int MB = 32 * WM;                  // rows per block row
    static constexpr int NJ = NT / (8 * WN);            // 8-column tiles per warp
    static constexpr int KC = TILE_KC, WLD = TILE_WLD;
    static constexpr int STAGE_DOUBLES = TILE_DOUBLES;
    static_assert(MB == TILE_MB, "the tiled operand layout assumes 128-row block rows (WM = 4)");
    static_assert(NT % (8 * WN) == 0, "NT must be a multiple of 8*WN");
};

struct Seg {
    const double* W;  // tiled matrix base
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

// ring state; every thread (producer and consumers) advances its own copy identically
struct Pipe {
    uint64_t* full;    // [STAGES]
    uint64_t* empty;   // [STAGES]
    double* stages;    // [STAGES][TILE_DOUBLES]
    uint32_t stage;    // slot of the next chunk
    uint32_t round;    // how many times the ring has wrapped (parity source)
};

template <class C>
__device__ __forceinline__ void pipe_init(Pipe& p, uint64_t* bars, double* stages) {
    p.full = bars;
    p.empty = bars + C::STAGES;
    p.stages = stages;
    p.stage = 0;
    p.round = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&p.full[s], 1);
            mbar_init(&p.empty[s], C::NWARPS);
        }
        mbar_fence_init();
    }
}
template <class C>
__device__ __forceinline__ void pipe_advance(Pipe& p) {
    if (++p.stage == C::STAGES) { p.stage = 0; ++p.round; }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}
// barrier among the consumer threads only (the producer warp never joins it)
template <class C>
__device__ __forceinline__ void consumer_sync() {
    asm volatile("bar.sync 1, %0;\n" :: "n"(C::NCONS) : "memory");
}

// ---- producer side: one elected lane enqueues every chunk of the phase
template <class C, class SegFn>
__device__ __forceinline__ void produce_phase(Pipe& p, int nseg, SegFn segfn, int Mpad) {
    const int nch = tiled_mpad(Mpad) / C::KC;   // chunks per block row in the tiled layout
    for (int s = 0; s < nseg; ++s) {
        const Seg sg = segfn(s);
        for (int kc = sg.kc0; kc < sg.kc1; ++kc) {
            if (p.round > 0) mbar_wait(&p.empty[p.stage], (p.round - 1) & 1);
            mbar_expect_tx(&p.full[p.stage], (uint32_t)(C::STAGE_DOUBLES * sizeof(double)));
            bulk_g2s(p.stages + (size_t)p.stage * C::STAGE_DOUBLES,
                     sg.W + ((size_t)sg.i * nch + kc) * C::STAGE_DOUBLES,
                     (uint32_t)(C::STAGE_DOUBLES * sizeof(double)), &p.full[p.stage]);
            pipe_advance<C>(p);
        }
    }
}

// One staged W chunk (16 deep) against the resident tile for the warp's tiles q in [QLO, QHI).
//   Aw = &Ws[(wm*8 + g) * WLD + t]          tile q adds q*WM*8 rows
//   Bw = &Bt[(wn*NJ*8 + g) * ldb + k0 + t]  column tile j adds j*8 rows of Bt
template <class C, int QLO, int QHI, bool WEIGHTED>
__device__ __forceinline__ void chunk_mma(Acc<C>& acc, const double* __restrict__ Aw, const double* __restrict__ Bw,
                                          int ldb, const double (&wv)[C::NJ]) {
    // Software-pipelined over the four k-steps of the chunk: the fragments of step kk+1 are loaded (into their own
    // registers) before the DMMAs of step kk are issued.  Left to itself ptxas reused one register pair for every A
    // fragment and put each LDS directly in front of its DMMA (30 % short-scoreboard stalls on the DMMAs in ncu).
    double af[2][4], bf[2][C::NJ];
    auto load = [&](int kk, int buf) {
#pragma unroll
        for (int j = 0; j < C::NJ; ++j) {
            bf[buf][j] = Bw[j * 8 * ldb + kk * 4];
            if (WEIGHTED) bf[buf][j] *= wv[j];
        }
#pragma unroll
        for (int q = QLO; q < QHI; ++q) af[buf][q] = Aw[q * C::WM * 8 * C::WLD + kk * 4];
    };
    load(0, 0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        if (kk + 1 < 4) load(kk + 1, (kk + 1) & 1);
        asm volatile("" ::: "memory");   // keep the prefetch above the DMMA group
#pragma unroll
        for (int q = QLO; q < QHI; ++q)
#pragma unroll
            for (int j = 0; j < C::NJ; ++j) dmma884(acc.v[q][j], af[kk & 1][q], bf[kk & 1][j]);
    }
}

// ---- consumer side.  WEIGHTED: the B fragment of column n is multiplied by wts[n * wld + seg.r] (backward pass).
// Starts with a consumer barrier (Bt is complete, previous phase's epilogue writes are visible).
template <class C, int MASK, bool WEIGHTED, class SegFn, class EpiFn>
__device__ __forceinline__ void run_phase(Pipe& p, int nseg, SegFn segfn, EpiFn epi, const double* __restrict__ Bt,
                                          int ldb, int Mpad, const double* __restrict__ wts, int wld) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / C::WN, wn = warp % C::WN;

    Acc<C> acc;
    acc.clear();
    consumer_sync<C>();

    for (int cs = 0; cs < nseg; ++cs) {
        const Seg seg = segfn(cs);
        const int row0 = seg.i * C::MB;
        double wv[C::NJ] = {};
        if (WEIGHTED) {
#pragma unroll
            for (int j = 0; j < C::NJ; ++j) wv[j] = wts[(wn * C::NJ * 8 + j * 8 + g) * wld + seg.r];
        }
        // Active 8-row tiles of this warp per chunk: a contiguous range [qlo, qhi) of its 4 interleaved tiles (tile q
        // starts at row rbase + 32 q).  Skipping uses real (warp-uniform) branches: a predicated-off DMMA still occupies
        // the fp64 tensor pipe for its full 16 cycles (measured with ncu), so predication would save nothing.
        const int rbase = row0 + wm * 8;
        int qmax = (Mpad - rbase + 31) >> 5;          // tiles with first row < Mpad
        qmax = qmax < 0 ? 0 : (qmax > 4 ? 4 : qmax);
        const double* Bw0 = Bt + (wn * C::NJ * 8 + g) * ldb + t;
        const int aoff = (wm * 8 + g) * C::WLD + t;
        for (int kc = seg.kc0; kc < seg.kc1; ++kc) {
            const int k0 = kc * C::KC;
            const double* Bw = Bw0 + k0;
            const double* Aw = p.stages + (size_t)p.stage * C::STAGE_DOUBLES + aoff;
            if (MASK == MASK_DENSE && qmax == 4) {
                // dense phase, full block row: no range bookkeeping at all
                mbar_wait(&p.full[p.stage], p.round & 1);
                chunk_mma<C, 0, 4, WEIGHTED>(acc, Aw, Bw, ldb, wv);
            } else {
                int qlo = 0, qhi = qmax;
                if (MASK == MASK_LOWER) {                 // tile active iff rbase + 32 q + 7 >= k0
                    qlo = (k0 - rbase + 24) >> 5;
                    qlo = qlo < 0 ? 0 : qlo;
                }
                if (MASK == MASK_UPPER) {                 // tile active iff rbase + 32 q <= k0 + 15
                    const int h = ((k0 + 15 - rbase) >> 5) + 1;
                    qhi = h < qmax ? (h < 0 ? 0 : h) : qmax;
                }
                mbar_wait(&p.full[p.stage], p.round & 1);
                if (qlo == 0 && qhi == 4) {
                    chunk_mma<C, 0, 4, WEIGHTED>(acc, Aw, Bw, ldb, wv);
                } else if (qlo < qhi) switch ((qlo << 3) | qhi) {
                    case (1 << 3) | 4: chunk_mma<C, 1, 4, WEIGHTED>(acc, Aw, Bw, ldb, wv); break;
                    case (2 << 3) | 4: chunk_mma<C, 2, 4, WEIGHTED>(acc, Aw, Bw, ldb, wv); break;
                    case (3 << 3) | 4: chunk_mma<C, 3, 4, WEIGHTED>(acc, Aw, Bw, ldb, wv); break;
                    case (0 << 3) | 3: chunk_mma<C, 0, 3, WEIGHTED>(acc, Aw, Bw, ldb, wv); break;
                    case (0 << 3) | 2: chunk_mma<C, 0, 2, WEIGHTED>(acc, Aw, Bw, ldb, wv); break;
                    case (0 << 3) | 1: chunk_mma<C, 0, 1, WEIGHTED>(acc, Aw, Bw, ldb, wv); break;
                    case (1 << 3) | 3: chunk_mma<C, 1, 3, WEIGHTED>(acc, Aw, Bw, ldb, wv); break;
                    case (1 << 3) | 2: chunk_mma<C, 1, 2, WEIGHTED>(acc, Aw, Bw, ldb, wv); break;
                    case (2 << 3) | 3: chunk_mma<C, 2, 3, WEIGHTED>(acc, Aw, Bw, ldb, wv); break;
                    default: break;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&p.empty[p.stage]);   // this warp is done reading the slot
            pipe_advance<C>(p);
        }
        if (seg.flush) {
            epi(acc, seg);
            acc.clear();
        }
    }
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
