// How many non-FP64 instructions can ride along with a DFMA before they cost FP64 throughput?  The FP64 pipe of an SM
// sub-partition takes a warp-wide DFMA every 2 cycles.  If the second cycle is a free issue slot, one extra instruction per
// DFMA is free; if the DFMA blocks the dispatch port for both cycles, every extra instruction costs time:
//   pipe utilisation = 2 / (2 + N)   for N extra instructions per DFMA.
// Kernel: 8 independent DFMA chains per thread (no latency limit), NX integer multiply-adds (IMAD, FMA-lite/ALU pipes)
// or moves interleaved per 8 DFMA.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_mix_bench tools/issue_mix_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NX>
__global__ void __launch_bounds__(128) mix(double* out, int iters, double a, double b, int seed) {
    double f[8];
    int x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { f[i] = threadIdx.x * 1e-3 + i; x[i] = seed + i * 7 + threadIdx.x; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = fma(f[i], a, b);
#pragma unroll
            for (int q = 0; q < NX; ++q) x[q & 7] = x[q & 7] * 3 + seed;       // IMAD, independent chains
        }
    }
    double s = 0; int t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += f[i]; t ^= x[i]; }
    if (s == 12345.678 || t == 0x7fffffff) out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}

template <int NX>
double run(int nsm, int bps, int iters) {
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<NX><<<nsm * bps, 128>>>(d, iters, 1.0000001, 1e-9, 3);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        mix<NX><<<nsm * bps, 128>>>(d, iters, 1.0000001, 1e-9, 3);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFree(d);
    return 2.0 * 16 * 8 * (double)iters * 128 * bps * nsm / (best * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount, iters = 2000;
    for (int bps = 2; bps <= 4; ++bps)
        printf("{\"warps_per_sm\":%d,\"dfma_tflops_by_extra_instr_per_dfma\":{\"0\":%.2f,\"0.25\":%.2f,\"0.5\":%.2f,\"1\":%.2f,\"1.5\":%.2f,\"2\":%.2f}}\n", 4 * bps,
               run<0>(nsm, bps, iters), run<2>(nsm, bps, iters), run<4>(nsm, bps, iters), run<8>(nsm, bps, iters), run<12>(nsm, bps, iters), run<16>(nsm, bps, iters));
    return 0;
}
