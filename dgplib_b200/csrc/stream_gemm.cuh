// CTA-level building block shared by the fused forward and backward row-tile kernels:
//
//     C[block-row i][NT columns]  (+)=  W[block-row i][k-range]  x  Bt[NT columns][k-range]^T
//
// W  : an Mpad x Mpad parameter matrix (Lm^-1, tril(q_sqrt_r)^T, tril(q_sqrt_r), Lm^-T), pre-tiled in global memory by
//      the parameter stage as [block row][16-deep chunk][128 rows][16 doubles] with the 32-byte groups of a row
//      XOR-permuted by (row & 3) (dgp::retile, common.cuh), so that one chunk is ONE contiguous 16 KB block and the
//      DMMA A-fragment loads are conflict-free without padding.  A dedicated producer warp streams the chunks with bulk
//      async copies (TMA engine, cp.async.bulk + mbarrier complete_tx) into a STAGES-deep shared-memory ring.
// Bt : the resident row tile in shared memory, stored [n][k] with pitch Mpad+4 (conflict-free B-fragment loads) -- or,
//      for the backward pass's dA phase (BSTAGE), a second streamed operand: the chunk [NT/8][16][8] of the saved
//      U_r = tril(q_sqrt_r)^T A that travels in the same ring slot as the W chunk it is multiplied with;
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

template <int WM_, int WN_, int NT_, int STAGES_, bool USTREAM_ = false>
struct TileCfg {
    static constexpr int WM = WM_, WN = WN_, NT = NT_, STAGES = STAGES_;
    static constexpr bool USTREAM = USTREAM_;            // ring slots carry a [NT/8][16][8] chunk of the second operand
    static constexpr int NWARPS = WM * WN;              // consumer warps
    static constexpr int NCONS = NWARPS * 32;           // consumer threads
    static constexpr int NTHREADS = NCONS + 32;         // + one producer warp
    static constexpr int MB = 32 * WM;                  // rows per block row
    static constexpr int NJ = NT / (8 * WN);            // 8-column tiles per warp
    static constexpr int KC = TILE_KC, WLD = TILE_WLD;
    static constexpr int UCH_DOUBLES = USTREAM ? NT * TILE_KC : 0;
    static constexpr int STAGE_DOUBLES = TILE_DOUBLES + UCH_DOUBLES;
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

// Accumulators of one warp: 4 row tiles x NJ column tiles.  With a single column tile per warp (NJ = 1: the 16- and 32-row
// configurations of Mpad > 256) the same four accumulators would be hit again every fourth DMMA; there the odd k-steps
// go to a second set (SPLIT), which doubles the distance between dependent DMMAs, and fold() adds the two sets before the
// epilogue reads v.  (Used where the B operand is the resident tile: M = 1024 forward -3 % with the saved tensors, -5 % without;
// with the streamed, weighted B operand of the backward dA phase it measured 2 % slower and is left off.)
template <class C, bool SPLIT_>
struct AccSecond { double w[4][C::NJ][2]; };
template <class C>
struct AccSecond<C, false> {};

template <class C, bool SPLIT_ = false>
struct Acc : AccSecond<C, SPLIT_> {
    static constexpr bool SPLIT = SPLIT_;
    double v[4][C::NJ][2];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < C::NJ; ++j) {
                v[q][j][0] = 0.0; v[q][j][1] = 0.0;
                if constexpr (SPLIT_) { this->w[q][j][0] = 0.0; this->w[q][j][1] = 0.0; }
            }
    }
    __device__ __forceinline__ void fold() {
        if constexpr (SPLIT_) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < C::NJ; ++j) { v[q][j][0] += this->w[q][j][0]; v[q][j][1] += this->w[q][j][1]; }
        }
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

// ---- producer side: one elected lane enqueues every chunk of the phase.  ufn(seg, kc, bytes&, policy&) returns the
//      global address of the second operand's chunk that shares the slot (bytes = 0: none) and its L2 eviction policy.
struct NoU {
    __device__ __forceinline__ const double* operator()(const Seg&, int, uint32_t& bytes, uint64_t&) const {
        bytes = 0;
        return nullptr;
    }
};
template <class C, class SegFn, class UFn = NoU>
__device__ __forceinline__ void produce_phase(Pipe& p, int nseg, SegFn segfn, int Mpad, UFn ufn = UFn()) {
    const int nch = tiled_mpad(Mpad) / C::KC;   // chunks per block row in the tiled layout
    for (int s = 0; s < nseg; ++s) {
        const Seg sg = segfn(s);
        for (int kc = sg.kc0; kc < sg.kc1; ++kc) {
            if (p.round > 0) mbar_wait(&p.empty[p.stage], (p.round - 1) & 1);
            uint32_t ub = 0;
            uint64_t upol = 0;
            const double* usrc = ufn(sg, kc, ub, upol);
            double* slot = p.stages + (size_t)p.stage * C::STAGE_DOUBLES;
            mbar_expect_tx(&p.full[p.stage], (uint32_t)(TILE_DOUBLES * sizeof(double)) + ub);
            bulk_g2s(slot, sg.W + ((size_t)sg.i * nch + kc) * TILE_DOUBLES, (uint32_t)(TILE_DOUBLES * sizeof(double)),
                     &p.full[p.stage]);
            if (ub) bulk_g2s_hint(slot + TILE_DOUBLES, usrc, ub, &p.full[p.stage], upol);
            pipe_advance<C>(p);
        }
    }
}

// One staged W chunk (16 deep) against the B operand for the warp's tiles q in [QLO, QHI).
//   Aw = &Ws[(wm*8 + g) * 16 + t]            tile q adds q*WM*8 rows; k-step kk sits at column group kk ^ (g & 3) (ko[kk])
//   resident B : Bw = &Bt[(wn*NJ*8 + g) * ldb + k0 + t]     column tile j adds j*8 rows of Bt, k-step kk 4 columns
//   streamed B : Bw = &Us[wn*NJ*128 + usave_swz(t, g)]      column tile j adds one 16 x 8 block, k-step kk 4 rows of it
// WEIGHTED: the B fragment of column n is multiplied by wv[j] (the backward pass's g_var[n, r]).  The multiply is a DMUL
// on the same fp64 pipe as the DMMAs and warps issue in order, so it is placed AFTER the DMMA group of the previous
// k-step (as volatile asm, which keeps its place between the volatile DMMAs): by then the LDS it depends on has long
// returned.  (Multiplying right at the load put a stalled DMUL in front of every DMMA group: 24 % of all warp stalls.)
__device__ __forceinline__ void dmul_inplace(double& x, double w) {
    asm volatile("mul.f64 %0, %0, %1;\n" : "+d"(x) : "d"(w));
}
// DIAG: the chunk starts exactly at the first row of one of the warp's 8-row tiles, which therefore meets only half of its
// 16 columns: 1 = lower-triangular operand, tile QLO uses k-steps 0..1 (columns k0 .. k0+7 <= its last row);
// 2 = upper-triangular operand whose tile QHI-1 starts 8 rows into the chunk and uses k-steps 2..3.  Resolved at compile
// time in the unrolled loops: the skipped DMMAs (and their A fragments) are simply not there.
template <int DIAG, int QLO, int QHI>
__device__ __forceinline__ constexpr bool kstep_active(int q, int kk) {
    return !((DIAG == 1 && q == QLO && kk >= 2) || (DIAG == 2 && q == QHI - 1 && kk < 2));
}
template <class C, int QLO, int QHI, bool BSTAGE, bool WEIGHTED, int DIAG = 0, class AccT>
__device__ __forceinline__ void chunk_mma(AccT& acc, const double* __restrict__ Aw, const double* __restrict__ Bw,
                                          int ldb, const int (&ko)[4], const double (&wv)[C::NJ]) {
    // Software-pipelined over the four k-steps of the chunk: the fragments of step kk+1 are loaded (into their own
    // registers) before the DMMAs of step kk are issued.  Left to itself ptxas reused one register pair for every A
    // fragment and put each LDS directly in front of its DMMA (30 % short-scoreboard stalls on the DMMAs in ncu).
    double af[2][4], bf[2][C::NJ];
    auto load = [&](int kk, int buf) {
#pragma unroll
        for (int j = 0; j < C::NJ; ++j)
            bf[buf][j] = BSTAGE ? Bw[j * USAVE_BLOCK + kk * 32] : Bw[j * 8 * ldb + kk * 4];
#pragma unroll
        for (int q = QLO; q < QHI; ++q)
            if (kstep_active<DIAG, QLO, QHI>(q, kk)) af[buf][q] = Aw[q * C::WM * 8 * C::WLD + ko[kk]];
    };
    auto scale = [&](int buf) {
#pragma unroll
        for (int j = 0; j < C::NJ; ++j) dmul_inplace(bf[buf][j], wv[j]);
    };
    load(0, 0);
    if (WEIGHTED) scale(0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        if (kk + 1 < 4) load(kk + 1, (kk + 1) & 1);
        asm volatile("" ::: "memory");   // keep the prefetch above the DMMA group
#pragma unroll
        for (int q = QLO; q < QHI; ++q)
#pragma unroll
            for (int j = 0; j < C::NJ; ++j)
                if (kstep_active<DIAG, QLO, QHI>(q, kk)) {
                    if constexpr (AccT::SPLIT) {
                        if (kk & 1) dmma884(acc.w[q][j], af[kk & 1][q], bf[kk & 1][j]);
                        else dmma884(acc.v[q][j], af[kk & 1][q], bf[kk & 1][j]);
                    } else {
                        dmma884(acc.v[q][j], af[kk & 1][q], bf[kk & 1][j]);
                    }
                }
        if (WEIGHTED && kk + 1 < 4) scale((kk + 1) & 1);
    }
}

// ---- consumer side.  BSTAGE: the B operand is the chunk streamed in the same ring slot (Bt / ldb unused).
// WEIGHTED: B column n is scaled by wts[n * wld + seg.r] for n < nvalid and by 0 beyond (backward pass).
// Starts with a consumer barrier (Bt is complete, previous phase's epilogue writes are visible).
template <class C, int MASK, bool BSTAGE, bool WEIGHTED = false, class SegFn, class EpiFn>
__device__ __forceinline__ void run_phase(Pipe& p, int nseg, SegFn segfn, EpiFn epi, const double* __restrict__ Bt,
                                          int ldb, int Mpad, const double* __restrict__ wts = nullptr, int wld = 0,
                                          int nvalid = 0) {
    static_assert(!BSTAGE || C::USTREAM, "BSTAGE needs a tile configuration whose ring slots carry the second operand");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / C::WN, wn = warp % C::WN;
    const int ko[4] = {((0 ^ (g & 3)) << 2), ((1 ^ (g & 3)) << 2), ((2 ^ (g & 3)) << 2), ((3 ^ (g & 3)) << 2)};

    Acc<C, (C::NJ == 1 && !BSTAGE)> acc;
    acc.clear();
    consumer_sync<C>();

    for (int cs = 0; cs < nseg; ++cs) {
        const Seg seg = segfn(cs);
        const int row0 = seg.i * C::MB;
        double wv[C::NJ] = {};
        if (WEIGHTED) {
#pragma unroll
            for (int j = 0; j < C::NJ; ++j) {
                const int n = wn * C::NJ * 8 + j * 8 + g;      // B-fragment column of this lane
                wv[j] = n < nvalid ? wts[n * wld + seg.r] : 0.0;
            }
        }
        // Active 8-row tiles of this warp per chunk: a contiguous range [qlo, qhi) of its 4 interleaved tiles (tile q
        // starts at row rbase + 32 q).  Skipping uses real (warp-uniform) branches: a predicated-off DMMA still occupies
        // the fp64 tensor pipe for its full 16 cycles (measured with ncu), so predication would save nothing.
        const int rbase = row0 + wm * 8;
        int qmax = (Mpad - rbase + 31) >> 5;          // tiles with first row < Mpad
        qmax = qmax < 0 ? 0 : (qmax > 4 ? 4 : qmax);
        const double* Bw0 = BSTAGE ? (const double*)nullptr : Bt + (wn * C::NJ * 8 + g) * ldb + t;
        const int boff = TILE_DOUBLES + wn * C::NJ * USAVE_BLOCK + usave_swz(t, g);   // + kk * 32: same swizzle bit
        const int aoff = (wm * 8 + g) * C::WLD + t;
        for (int kc = seg.kc0; kc < seg.kc1; ++kc) {
            const int k0 = kc * C::KC;
            const double* slot = p.stages + (size_t)p.stage * C::STAGE_DOUBLES;
            const double* Bw = BSTAGE ? slot + boff : Bw0 + k0;
            const double* Aw = slot + aoff;
            if (MASK == MASK_DENSE && qmax == 4) {
                // dense phase, full block row: no range bookkeeping at all
                mbar_wait(&p.full[p.stage], p.round & 1);
                chunk_mma<C, 0, 4, BSTAGE, WEIGHTED>(acc, Aw, Bw, ldb, ko, wv);
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
                // the chunk that starts on a tile's own first row (lower) / 8 columns before it (upper): half of that
                // tile's k-steps hit the zero triangle (3 % of all DMMAs of a 256-row operand)
                // Only in the wide-tile configurations (Mpad <= 256): with 16- / 32-row tiles, i.e. 512 or 1024 operand rows,
                // the share of diagonal chunks is small and the extra code measured slower (M = 1024 forward + 2.5 %).
                constexpr bool use_diag = C::NT >= 40;
                const bool diag_lo = use_diag && MASK == MASK_LOWER && qhi == 4 && qlo < 4 && rbase + 32 * qlo == k0;
                const bool diag_up = use_diag && MASK == MASK_UPPER && qlo == 0 && qhi >= 1 && rbase + 32 * (qhi - 1) == k0 + 8;
                mbar_wait(&p.full[p.stage], p.round & 1);
                if (diag_lo) switch (qlo) {
                    case 0: chunk_mma<C, 0, 4, BSTAGE, WEIGHTED, 1>(acc, Aw, Bw, ldb, ko, wv); break;
                    case 1: chunk_mma<C, 1, 4, BSTAGE, WEIGHTED, 1>(acc, Aw, Bw, ldb, ko, wv); break;
                    case 2: chunk_mma<C, 2, 4, BSTAGE, WEIGHTED, 1>(acc, Aw, Bw, ldb, ko, wv); break;
                    default: chunk_mma<C, 3, 4, BSTAGE, WEIGHTED, 1>(acc, Aw, Bw, ldb, ko, wv); break;
                } else if (diag_up) switch (qhi) {
                    case 1: chunk_mma<C, 0, 1, BSTAGE, WEIGHTED, 2>(acc, Aw, Bw, ldb, ko, wv); break;
                    case 2: chunk_mma<C, 0, 2, BSTAGE, WEIGHTED, 2>(acc, Aw, Bw, ldb, ko, wv); break;
                    case 3: chunk_mma<C, 0, 3, BSTAGE, WEIGHTED, 2>(acc, Aw, Bw, ldb, ko, wv); break;
                    default: chunk_mma<C, 0, 4, BSTAGE, WEIGHTED, 2>(acc, Aw, Bw, ldb, ko, wv); break;
                } else if (qlo == 0 && qhi == 4) {
                    chunk_mma<C, 0, 4, BSTAGE, WEIGHTED>(acc, Aw, Bw, ldb, ko, wv);
                } else if (qlo < qhi) switch ((qlo << 3) | qhi) {
                    case (1 << 3) | 4: chunk_mma<C, 1, 4, BSTAGE, WEIGHTED>(acc, Aw, Bw, ldb, ko, wv); break;
                    case (2 << 3) | 4: chunk_mma<C, 2, 4, BSTAGE, WEIGHTED>(acc, Aw, Bw, ldb, ko, wv); break;
                    case (3 << 3) | 4: chunk_mma<C, 3, 4, BSTAGE, WEIGHTED>(acc, Aw, Bw, ldb, ko, wv); break;
                    case (0 << 3) | 3: chunk_mma<C, 0, 3, BSTAGE, WEIGHTED>(acc, Aw, Bw, ldb, ko, wv); break;
                    case (0 << 3) | 2: chunk_mma<C, 0, 2, BSTAGE, WEIGHTED>(acc, Aw, Bw, ldb, ko, wv); break;
                    case (0 << 3) | 1: chunk_mma<C, 0, 1, BSTAGE, WEIGHTED>(acc, Aw, Bw, ldb, ko, wv); break;
                    case (1 << 3) | 3: chunk_mma<C, 1, 3, BSTAGE, WEIGHTED>(acc, Aw, Bw, ldb, ko, wv); break;
                    case (1 << 3) | 2: chunk_mma<C, 1, 2, BSTAGE, WEIGHTED>(acc, Aw, Bw, ldb, ko, wv); break;
                    case (2 << 3) | 3: chunk_mma<C, 2, 3, BSTAGE, WEIGHTED>(acc, Aw, Bw, ldb, ko, wv); break;
                    default: break;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&p.empty[p.stage]);   // this warp is done reading the slot
            pipe_advance<C>(p);
        }
        if (seg.flush) {
            acc.fold();
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
