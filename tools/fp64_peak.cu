// FP64 DFMA-chain peak micro-benchmark for the roofline denominator (MEASURED_PEAKS.json has no
// FP64 entry; SURVEY.md 8d asks the builder to measure one).  Reports TFLOP/s (2 flop per DFMA)
// for several (warps per SM, independent chains per thread) points and the sustained figure of a
// multi-second loop.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_chain(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
double run(int blocks_per_sm, int threads, int iters, int reps, int nsm) {
    double* d;
    cudaMalloc(&d, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_chain<ILP><<<nsm * blocks_per_sm, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        dfma_chain<ILP><<<nsm * blocks_per_sm, threads>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFree(d);
    double flops = 2.0 * 16.0 * ILP * (double)iters * threads * blocks_per_sm * nsm;
    return flops / (best * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    printf("{\"gpu\":\"%s\",\"sms\":%d,\"points\":[", p.name, nsm);
    const int iters = 4096;
    struct { int bps, thr; } cfg[] = {{1, 128}, {1, 256}, {2, 256}, {4, 256}, {8, 256}};
    double best = 0;
    bool first = true;
    for (auto c : cfg) {
        double t4 = run<4>(c.bps, c.thr, iters, 5, nsm);
        double t8 = run<8>(c.bps, c.thr, iters, 5, nsm);
        double t2 = run<2>(c.bps, c.thr, iters, 5, nsm);
        printf("%s{\"warps_per_sm\":%d,\"ilp2\":%.2f,\"ilp4\":%.2f,\"ilp8\":%.2f}", first ? "" : ",", c.bps * c.thr / 32, t2, t4, t8);
        first = false;
        if (t8 > best) best = t8;
        if (t4 > best) best = t4;
    }
    // sustained: ~3 s back to back at the best shape
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int launches = 0;
    cudaEventRecord(e0);
    float ms = 0;
    do {
        for (int k = 0; k < 20; ++k) dfma_chain<8><<<nsm * 4, 256>>>(d, iters, 1.0000001, 1e-9);
        launches += 20;
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    } while (ms < 3000.f);
    double sustained = 2.0 * 16.0 * 8 * (double)iters * 256 * 4 * nsm * launches / (ms * 1e-3) / 1e12;
    printf("],\"fp64_tflops_burst\":%.2f,\"fp64_tflops_sustained\":%.2f}\n", best, sustained);
    return 0;
}
