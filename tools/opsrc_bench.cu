// Where should a table-driven FP64 loop take its warp-uniform operator constants from?  Micro-benchmark of the windowed
// loop kernel's instruction mix: a loop body of K DFMAs, every constant feeding two DFMAs (re, im of a register-resident
// column value), with the constants read per iteration from
//   mode 0: shared memory, broadcast LDS.128 (two constants per load)
//   mode 1: __constant__ memory with a warp-uniform running index, two constants per load (LDCU.128)
//   mode 2: global memory through the read-only path (warp-uniform LDG.128)
//   mode 3: __constant__ memory, one 64-bit load per constant (LDCU.64)
// Reports TFLOP/s (2 flop per DFMA) for several warps-per-SM points.  The body (K DFMAs + K/4 loads, < 12 KB) stays in the
// instruction cache, so the number isolates the operand path.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/opsrc_bench tools/opsrc_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int K = 512;          // DFMAs per loop iteration
constexpr int NIT = 9;          // iterations per sweep (one per m block)
constexpr int NTAB = NIT * K / 2;   // doubles
__constant__ double c_tab[NTAB];

template <int MODE>
__global__ void __launch_bounds__(128) body(const double* __restrict__ gtab, double* out, int sweeps, double seed) {
    extern __shared__ __align__(16) double stab[];
    if (MODE == 0) {
        for (int i = threadIdx.x; i < NTAB; i += blockDim.x) stab[i] = gtab[i];
        __syncthreads();
    }
    double yr[16], yi[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { yr[i] = seed + i * 0.01 + threadIdx.x * 1e-3; yi[i] = seed - i * 0.02 + threadIdx.x * 1e-3; }
    double ar[5], ai[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) { ar[i] = 0.0; ai[i] = 0.0; }
    for (int s = 0; s < sweeps; ++s) {
        for (int it = 0; it < NIT; ++it) {
            const int base = it * (K / 2);
#pragma unroll
            for (int i = 0; i < K / 2; i += 2) {
                double c0, c1;
                if (MODE == 0) { const double2 p = *reinterpret_cast<const double2*>(stab + base + i); c0 = p.x; c1 = p.y; }
                else if (MODE == 1) { const double2 p = reinterpret_cast<const double2*>(c_tab)[(base + i) / 2]; c0 = p.x; c1 = p.y; }
                else if (MODE == 2) { const double2 p = __ldg(reinterpret_cast<const double2*>(gtab + base + i)); c0 = p.x; c1 = p.y; }
                else { c0 = c_tab[base + i]; c1 = c_tab[base + i + 1]; }
                const int a = (i / 2) % 5, j0 = i % 16, j1 = (i + 1) % 16;
                ar[a] = fma(c0, yr[j0], ar[a]); ai[a] = fma(c0, yi[j0], ai[a]);
                ar[a] = fma(c1, yr[j1], ar[a]); ai[a] = fma(c1, yi[j1], ai[a]);
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) { yr[i] += 1e-9 * ar[i % 5]; yi[i] -= 1e-9 * ai[i % 5]; }
    }
    double t = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) t += ar[i] + ai[i];
    if (t == 1.2345) out[threadIdx.x] = t;
}


// LDCU throughput: S DFMAs per 64-bit constant (constant bank, warp-uniform running index)
template <int S>
__global__ void __launch_bounds__(128) ratio(double* out, int sweeps, double seed) {
    double y[8], a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { y[i] = seed + i * 0.01 + threadIdx.x * 1e-3; a[i] = 0.0; }
    for (int s = 0; s < sweeps; ++s) {
        for (int it = 0; it < NIT; ++it) {
            const int base = it * (K / 2);
#pragma unroll
            for (int i = 0; i < K / 2; ++i) {
                const double c0 = c_tab[base + i];
#pragma unroll
                for (int q = 0; q < S; ++q) a[(i * S + q) & 7] = fma(c0, y[(i + q) & 7], a[(i * S + q) & 7]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] += 1e-9 * a[i];
    }
    double t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += a[i];
    if (t == 1.2345) out[threadIdx.x] = t;
}
template <int S>
double run_ratio(double* d, int nsm, int cps, int sweeps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    ratio<S><<<nsm * cps, 128>>>(d, sweeps, 1.0);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        ratio<S><<<nsm * cps, 128>>>(d, sweeps, 1.0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return 2.0 * S * (K / 2) * NIT * (double)sweeps * 128 * cps * nsm / (best * 1e-3) / 1e12;
}

template <int MODE>
double run(const double* gtab, double* d, int nsm, int ctas_per_sm, int sweeps) {
    const size_t smem = MODE == 0 ? NTAB * 8 : 0;
    cudaFuncSetAttribute(body<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    body<MODE><<<nsm * ctas_per_sm, 128, smem>>>(gtab, d, sweeps, 1.0);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        body<MODE><<<nsm * ctas_per_sm, 128, smem>>>(gtab, d, sweeps, 1.0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    if (cudaGetLastError() != cudaSuccess) return -1.0;
    return 2.0 * K * NIT * (double)sweeps * 128 * ctas_per_sm * nsm / (best * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount;
    std::vector<double> h(NTAB);
    for (int i = 0; i < NTAB; ++i) h[i] = 0.5 + (i % 97) * 1e-3;
    double *gtab, *d;
    cudaMalloc(&gtab, NTAB * 8); cudaMalloc(&d, 4096);
    cudaMemcpy(gtab, h.data(), NTAB * 8, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(c_tab, h.data(), NTAB * 8);
    const char* names[4] = {"lds128_broadcast", "constant_bank_ldcu128", "ldg128_uniform", "constant_bank_ldcu64"};
    for (int cps = 1; cps <= 4; ++cps) {
        const int sweeps = 400;
        double r[4] = {run<0>(gtab, d, nsm, cps, sweeps), run<1>(gtab, d, nsm, cps, sweeps), run<2>(gtab, d, nsm, cps, sweeps),
                       run<3>(gtab, d, nsm, cps, sweeps)};
        for (int m = 0; m < 4; ++m)
            printf("{\"source\":\"%s\",\"warps_per_sm\":%d,\"dfma_per_iter\":%d,\"table_kb\":%.1f,\"tflops\":%.2f}\n", names[m], 4 * cps, K,
                   NTAB * 8 / 1024.0, r[m]);
    }
    for (int cps = 2; cps <= 4; ++cps) {
        const int sweeps = 400;
        printf("{\"source\":\"constant_bank_ldcu64\",\"warps_per_sm\":%d,\"tflops_by_dfma_per_constant\":{\"1\":%.2f,\"2\":%.2f,\"3\":%.2f,\"4\":%.2f}}\n", 4 * cps,
               run_ratio<1>(d, nsm, cps, sweeps), run_ratio<2>(d, nsm, cps, sweeps), run_ratio<3>(d, nsm, cps, sweeps), run_ratio<4>(d, nsm, cps, sweeps));
    }
    return 0;
}
