// FP64 peak microbenchmark for B200 (sm_100a): register-resident DMMA m8n8k4 and DFMA issue rates.
// Used once to establish the "FP64 tensor peak" roofline denominator; not on the product path.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters, double a0, double b0) {
    double acc[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA with shared-memory operand loads each k-step (32x32 warp tile: 4 A + 4 B LDS.64 per 16 DMMA)
__global__ void k_dmma_lds(double* out, int iters) {
    __shared__ double sA[64 * 20];
    __shared__ double sB[64 * 20];
    for (int i = threadIdx.x; i < 64 * 20; i += blockDim.x) { sA[i] = 1e-3 * i; sB[i] = 1e-4 * i; }
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0; acc[i][j][1] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[(i * 8 + g) * 20 + kk * 4 + t];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[(j * 8 + g) * 20 + kk * 4 + t];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// The streamed-GEMM inner loop of the product kernels without any synchronisation: A fragments from a 128x(16+4)
// staged chunk (4 ring slots), B fragments from a resident NT x (256+4) tile; warp grid 4 x WN, 4 interleaved 8-row
// tiles per warp.  Shows the ceiling of the smem-fed DMMA loop itself.
template <int WN, int NT>
__global__ void k_dmma_stream(double* out, int nchunks) {
    extern __shared__ double sm[];
    constexpr int NJ = NT / (8 * WN), LDB = 260;
    double* Bt = sm;
    double* Ws = sm + NT * LDB;
    for (int i = threadIdx.x; i < NT * LDB + 4 * 128 * 20; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int wm = warp / WN, wn = warp % WN;
    double acc[4][NJ][2];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < NJ; ++j) { acc[q][j][0] = 0; acc[q][j][1] = 0; }
    for (int c = 0; c < nchunks; ++c) {
        const double* Aw = Ws + (c & 3) * 128 * 20 + (wm * 8 + g) * 20 + t;
        const double* Bw = Bt + (wn * NJ * 8 + g) * LDB + (c & 15) * 16 + t;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            double b[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) b[j] = Bw[j * 8 * LDB + kk * 4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double a = Aw[q * 32 * 20 + kk * 4];
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[q][j][0], acc[q][j][1], a, b[j]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < NJ; ++j) s += acc[q][j][0] + acc[q][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// The same loop WITH the product's pipeline: a producer warp streams 20 KB chunks from global memory with cp.async.bulk
// into a 4-slot ring, full/empty mbarriers, consumers wait/arrive per chunk (stream_gemm.cuh, dense phase).
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int WN, int NT, int EPI>
__global__ void k_dmma_stream_pipe(double* out, const double* W, int ntiles_src, int nchunks) {
    extern __shared__ __align__(128) double sm[];
    constexpr int NJ = NT / (8 * WN), LDB = 260, STAGES = 4, NW = 4 * WN;
    double* Bt = sm;
    double* Ws = sm + NT * LDB;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(Ws + STAGES * 128 * 20);
    for (int i = threadIdx.x; i < NT * LDB; i += blockDim.x) Bt[i] = 1e-3 * (i % 97) + 1.2345678901234e-7 * ((i * 2654435761u) % 1000003u);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(&bars[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(&bars[STAGES + s])), "r"(NW));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    auto wait = [&](unsigned long long* b, unsigned par) {
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
                     :: "r"(s32(b)), "r"(par) : "memory");
    };
    if (warp == NW) {
        if (lane == 0) {
            for (int c = 0; c < nchunks; ++c) {
                const int st = c % STAGES, rd = c / STAGES;
                if (rd > 0) wait(&bars[STAGES + st], (rd - 1) & 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(&bars[st])), "r"(128 * 20 * 8) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(s32(Ws + st * 128 * 20)), "l"(W + (size_t)(c % ntiles_src) * 128 * 20), "r"(128 * 20 * 8), "r"(s32(&bars[st])) : "memory");
            }
        }
        return;
    }
    const int wm = warp / WN, wn = warp % WN;
    double acc[4][NJ][2], tot[4][NJ][2];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < NJ; ++j) { acc[q][j][0] = 0; acc[q][j][1] = 0; tot[q][j][0] = 0; tot[q][j][1] = 0; }
    for (int c = 0; c < nchunks; ++c) {
        const int st = c % STAGES, rd = c / STAGES;
        if (EPI && (c & 15) == 0 && c > 0) {
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double w = Bt[(wn * NJ * 8 + j * 8 + 2 * t + e) * LDB + (c >> 4 & 7)];
#pragma unroll
                    for (int q = 0; q < 4; ++q) { tot[q][j][e] = fma(w, acc[q][j][e], tot[q][j][e]); acc[q][j][e] = 0; }
                }
        }
        const double* Aw = Ws + st * 128 * 20 + (wm * 8 + g) * 20 + t;
        const double* Bw = Bt + (wn * NJ * 8 + g) * LDB + (c & 15) * 16 + t;
        wait(&bars[st], rd & 1);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            double b[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) b[j] = Bw[j * 8 * LDB + kk * 4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double a = Aw[q * 32 * 20 + kk * 4];
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[q][j][0], acc[q][j][1], a, b[j]);
            }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(s32(&bars[STAGES + st])) : "memory");
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < NJ; ++j) s += acc[q][j][0] + acc[q][j][1] + tot[q][j][0] + tot[q][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f();  // warm
    if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) {      // e.g. too many registers for 1024 threads: no number
        printf("  (launch failed: %s -- configuration skipped)\n", cudaGetErrorString(e));
        return -1.0f;
    }
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int ctas = 1; ctas <= 2; ++ctas) {
            int threads = warps * 32 / ctas;
            if (threads > 1024 || threads < 32) continue;
            {
                float ms = time_ms([&] { k_dmma<8><<<sms * ctas, threads>>>(out, iters, 1.0, 1.0); });
                double fl = 2.0 * 256 * 8 * (double)iters * warps * sms;
                if (ms > 0) printf("DMMA884 acc=8  warps/SM=%2d ctas/SM=%d : %.3f ms  %.2f TFLOP/s\n", warps, ctas, ms, fl / ms * 1e-9);
            }
            {
                float ms = time_ms([&] { k_dmma<16><<<sms * ctas, threads>>>(out, iters, 1.0, 1.0); });
                double fl = 2.0 * 256 * 16 * (double)iters * warps * sms;
                if (ms > 0) printf("DMMA884 acc=16 warps/SM=%2d ctas/SM=%d : %.3f ms  %.2f TFLOP/s\n", warps, ctas, ms, fl / ms * 1e-9);
            }
            {
                float ms = time_ms([&] { k_dfma<16><<<sms * ctas, threads>>>(out, iters, 1.0000001, 1e-9); });
                double fl = 2.0 * 32 * 16 * (double)iters * warps * sms;
                if (ms > 0) printf("DFMA    acc=16 warps/SM=%2d ctas/SM=%d : %.3f ms  %.2f TFLOP/s\n", warps, ctas, ms, fl / ms * 1e-9);
            }
        }
    }
    for (int warps = 4; warps <= 16; warps *= 2) {
        float ms = time_ms([&] { k_dmma_lds<<<sms, warps * 32>>>(out, 2000); });
        double fl = 2.0 * 256 * 64 * 2000.0 * warps * sms;
        printf("DMMA884+LDS 32x32 warp tile warps/SM=%2d : %.3f ms  %.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
    }
    {
        const int nch = 20000;
        auto run = [&](auto kern, int wn, int nt, const char* name) {
            size_t bytes = (size_t)(nt * 260 + 4 * 128 * 20) * sizeof(double);
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            float ms = time_ms([&] { kern<<<sms, 4 * wn * 32, bytes>>>(out, nch); });
            double fl = 2.0 * 256 * 16.0 * (nt / 8) * 4 * (double)nch * sms;   // 16 m-tiles x nt/8 n-tiles x 4 k-steps per chunk
            printf("DMMA stream loop %s: %.3f ms  %.2f TFLOP/s\n", name, ms, fl / ms * 1e-9);
        };
        run(k_dmma_stream<2, 64>, 2, 64, "4x2 warps, NT=64 (NJ=4)");
        run(k_dmma_stream<4, 64>, 4, 64, "4x4 warps, NT=64 (NJ=2)");
        run(k_dmma_stream<2, 32>, 2, 32, "4x2 warps, NT=32 (NJ=2)");
        run(k_dmma_stream<4, 32>, 4, 32, "4x4 warps, NT=32 (NJ=1)");
    }
    {
        const int nch = 20000, nsrc = 512;
        double* W; CK(cudaMalloc(&W, sizeof(double) * 128 * 20 * nsrc)); CK(cudaMemset(W, 0, sizeof(double) * 128 * 20 * nsrc));
        if (getenv("DGP_RANDOM_W")) {
            std::vector<double> h((size_t)128 * 20 * nsrc);
            for (size_t i = 0; i < h.size(); ++i) h[i] = (double)rand() / RAND_MAX - 0.5;
            CK(cudaMemcpy(W, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
        auto run = [&](auto kern, int wn, int nt, const char* name) {
            size_t bytes = (size_t)(nt * 260 + 4 * 128 * 20 + 16) * sizeof(double);
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            float ms = time_ms([&] { kern<<<sms, 4 * wn * 32 + 32, bytes>>>(out, W, nsrc, nch); });
            double fl = 2.0 * 256 * 16.0 * (nt / 8) * 4 * (double)nch * sms;
            printf("DMMA stream loop + TMA/mbarrier pipeline %s: %.3f ms  %.2f TFLOP/s\n", name, ms, fl / ms * 1e-9);
        };
        run(k_dmma_stream_pipe<2, 64, 0>, 2, 64, "4x2 warps, NT=64");
        run(k_dmma_stream_pipe<4, 64, 0>, 4, 64, "4x4 warps, NT=64");
        run(k_dmma_stream_pipe<2, 32, 0>, 2, 32, "4x2 warps, NT=32");
        run(k_dmma_stream_pipe<4, 32, 0>, 4, 32, "4x4 warps, NT=32");
        run(k_dmma_stream_pipe<4, 32, 1>, 4, 32, "4x4 warps, NT=32, scale-accumulate epilogue every 16 chunks");
        run(k_dmma_stream_pipe<4, 64, 1>, 4, 64, "4x4 warps, NT=64, scale-accumulate epilogue every 16 chunks");
        CK(cudaFree(W));
    }
    CK(cudaFree(out));
    return 0;
}
