// FP64 tensor-core (DMMA) peak next to the DFMA peak of tools/fp64_peak.cu, and whether the two pipes overlap.
// north_star allows a dense [K.nlm x nlm] x [nlm x N] FP64 contraction on tensor cores "only where ncu shows it beats
// CUDA-core FP64"; this tool supplies the denominator of that comparison on sm_100a:
//   dmma_m8n8k4   : mma.sync.aligned.m8n8k4.row.col.f64   (512 flop per warp instruction)
//   dmma_m16n8k16 : mma.sync.aligned.m16n8k16.row.col.f64 (4096 flop per warp instruction)
//   dfma          : 8 independent DFMA chains per thread
//   mixed         : both instruction streams interleaved in one warp (do the pipes run concurrently?)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_peak tools/dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// MODE 0: m8n8k4, 1: m16n8k16, 2: DFMA, 3: m16n8k16 + DFMA interleaved (NF DFMAs per DMMA)
template <int MODE, int NF>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a0, double b0) {
    double c2[8][2], c4[4][4], f[8], a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c2[i][0] = c2[i][1] = 0.0; f[i] = threadIdx.x * 1e-3 + i; a[i] = a0 + i * 1e-3; }
#pragma unroll
    for (int i = 0; i < 4; ++i) { b[i] = b0 + i * 1e-3; c4[i][0] = c4[i][1] = c4[i][2] = c4[i][3] = 0.0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) mma884(c2[i], a[i], b[i & 3]);
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 4; ++i) mma16816(c4[i], a, b);
            } else if (MODE == 2) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] = fma(f[i], a0, b0);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    mma16816(c4[i], a, b);
#pragma unroll
                    for (int q = 0; q < NF; ++q) f[(i * NF + q) & 7] = fma(f[(i * NF + q) & 7], a0, b0);
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c2[i][0] + c2[i][1] + f[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c4[i][0] + c4[i][1] + c4[i][2] + c4[i][3];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NF>
float run(int nsm, int bps, int iters) {
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, NF><<<nsm * bps, 256>>>(d, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<MODE, NF><<<nsm * bps, 256>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFree(d);
    return best * 1e-3f;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount, iters = 2048;
    printf("{\"gpu\":\"%s\",\"sms\":%d,\"points\":[", p.name, nsm);
    bool first = true;
    double best884 = 0, best16816 = 0, bestf = 0;
    for (int bps = 1; bps <= 4; bps *= 2) {
        const double warps = (double)nsm * bps * 8;
        const double t0 = run<0, 0>(nsm, bps, iters), t1 = run<1, 0>(nsm, bps, iters), t2 = run<2, 0>(nsm, bps, iters);
        const double f0 = warps * iters * 8 * 8 * 512.0 / t0 / 1e12;         // m8n8k4: 2*8*8*4 flop
        const double f1 = warps * iters * 8 * 4 * 4096.0 / t1 / 1e12;        // m16n8k16: 2*16*8*16 flop
        const double f2 = warps * 32 * iters * 8 * 32 * 2.0 / t2 / 1e12;
        printf("%s{\"warps_per_sm\":%d,\"dmma_m8n8k4_tflops\":%.2f,\"dmma_m16n8k16_tflops\":%.2f,\"dfma_tflops\":%.2f", first ? "" : ",", bps * 8, f0, f1, f2);
        // mixed: 4 DMMA (m16n8k16) + 4*NF DFMA per inner step
        const double tm8 = run<3, 8>(nsm, bps, iters), tm32 = run<3, 32>(nsm, bps, iters), tm64 = run<3, 64>(nsm, bps, iters);
        auto mixed = [&](double t, int nf) { return warps * iters * 8 * 4 * (4096.0 + 32.0 * nf * 2.0) / t / 1e12; };
        printf(",\"mixed_nf8_tflops\":%.2f,\"mixed_nf32_tflops\":%.2f,\"mixed_nf64_tflops\":%.2f}", mixed(tm8, 8), mixed(tm32, 32), mixed(tm64, 64));
        first = false;
        if (f0 > best884) best884 = f0;
        if (f1 > best16816) best16816 = f1;
        if (f2 > bestf) bestf = f2;
    }
    printf("],\"dmma_m8n8k4_tflops\":%.2f,\"dmma_m16n8k16_tflops\":%.2f,\"dfma_tflops\":%.2f}\n", best884, best16816, bestf);
    return 0;
}
