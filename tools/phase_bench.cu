// Micro-benchmark of the product's own streamed-GEMM phase code (stream_gemm.cuh: produce_phase / run_phase) in
// isolation: one dense phase over many segments, nothing else in the kernel.  Used to separate the cost of the phase
// machinery from the rest of the row-tile kernels (profiles/r01_notes.md).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../dgplib_b200/csrc/stream_gemm.cuh"

namespace dgp { long long g_launches = 0; int device_sm_count() { return 148; } }
extern "C" int32_t dgp_mpad(int32_t M) { return (M + 63) / 64 * 64; }
using namespace dgp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <class C, int MASK>
__global__ void __launch_bounds__(C::NTHREADS, 1) k_phase(const double* W, size_t tl, int R, int Mpad, int reps, double* out) {
    extern __shared__ __align__(128) double smem[];
    const int ldb = Mpad + 4;
    double* Bt = smem;
    double* Wst = smem + C::NT * ldb;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Wst + C::STAGES * C::STAGE_DOUBLES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = (Mpad + C::MB - 1) / C::MB;
    for (int i = tid; i < C::NT * ldb; i += C::NTHREADS) Bt[i] = 1e-3 * (i % 97) + 1e-7 * ((i * 2654435761u) % 1000003u);
    auto seg = [&](int s) {
        Seg sg;
        sg.i = (s / R) % nb;
        sg.r = s % R;
        sg.W = W + (size_t)sg.r * tl;
        if (MASK == MASK_DENSE) { sg.kc0 = 0; sg.kc1 = Mpad / C::KC; }
        else if (MASK == MASK_UPPER) { sg.kc0 = sg.i * C::MB / C::KC; sg.kc1 = Mpad / C::KC; }
        else { sg.kc0 = 0; int kend = (sg.i + 1) * C::MB; if (kend > Mpad) kend = Mpad; sg.kc1 = kend / C::KC; }
        sg.flush = true;
        return sg;
    };
    Pipe pipe;
    pipe_init<C>(pipe, bars, Wst);
    __syncthreads();
    const int nseg = nb * R * reps;
    if (warp == C::NWARPS) {
        if (lane == 0) produce_phase<C>(pipe, nseg, seg, Mpad);
        return;
    }
    Acc<C> tot;
    tot.clear();
    auto epi = [&](Acc<C>& acc, const Seg& sg) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < C::NJ; ++j) { tot.v[q][j][0] += acc.v[q][j][0]; tot.v[q][j][1] += acc.v[q][j][1]; }
    };
    run_phase<C, MASK, false>(pipe, nseg, seg, epi, Bt, ldb, Mpad, nullptr, 0);
    double s = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < C::NJ; ++j) s += tot.v[q][j][0] + tot.v[q][j][1];
    out[blockIdx.x * C::NTHREADS + tid] = s;
}

template <class C, int MASK>
void run(const char* name, const double* W, size_t tl, int R, int Mpad, double* out) {
    const int reps = 40;
    size_t bytes = (size_t)(C::NT * (Mpad + 4) + C::STAGES * C::STAGE_DOUBLES + 2 * C::STAGES) * sizeof(double);
    CK(cudaFuncSetAttribute(k_phase<C, MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
        CK(cudaEventRecord(e0));
        k_phase<C, MASK><<<148, C::NTHREADS, bytes>>>(W, tl, R, Mpad, reps, out);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (it > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const int nb = Mpad / 128;
    double dm = 0;   // DMMAs per CTA at 8-row / 16-deep granularity
    for (int i = 0; i < nb; ++i)
        for (int mt = 0; mt < 16; ++mt)
            for (int kc = 0; kc < Mpad / 16; ++kc) {
                int r0 = i * 128 + mt * 8, k0 = kc * 16;
                bool act = MASK == MASK_DENSE || (MASK == MASK_LOWER ? (kc < (i + 1) * 8 && r0 + 7 >= k0) : (kc >= i * 8 && r0 <= k0 + 15));
                if (act) dm += 4.0 * (C::NT / 8);
            }
    dm *= (double)R * reps * 148;
    printf("%-58s %8.3f ms  %6.2f TFLOP/s (executed DMMAs)\n", name, best, dm * 512 / best * 1e-9);
}

int main() {
    const int Mpad = 256, R = 8;
    const size_t tl = tiled_doubles(Mpad);
    std::vector<double> h(tl * R);
    for (auto& v : h) v = (double)rand() / RAND_MAX - 0.5;
    double *W, *out;
    CK(cudaMalloc(&W, h.size() * sizeof(double))); CK(cudaMemcpy(W, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&out, sizeof(double) * 148 * 1024));
    run<TileCfg<4, 4, 32, 4>, MASK_DENSE>("run_phase dense  4x4 warps NT=32 (backward dA)", W, tl, R, Mpad, out);
    run<TileCfg<4, 2, 32, 4>, MASK_DENSE>("run_phase dense  4x2 warps NT=32", W, tl, R, Mpad, out);
    run<TileCfg<4, 4, 64, 4>, MASK_DENSE>("run_phase dense  4x4 warps NT=64", W, tl, R, Mpad, out);
    run<TileCfg<4, 4, 64, 4>, MASK_UPPER>("run_phase upper  4x4 warps NT=64 (forward U)", W, tl, R, Mpad, out);
    run<TileCfg<4, 2, 64, 4>, MASK_UPPER>("run_phase upper  4x2 warps NT=64", W, tl, R, Mpad, out);
    run<TileCfg<4, 4, 64, 4>, MASK_LOWER>("run_phase lower  4x4 warps NT=64 (forward T)", W, tl, R, Mpad, out);
    return 0;
}
