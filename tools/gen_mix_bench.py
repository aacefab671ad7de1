"""Generate a micro-benchmark of the step kernel's instruction mix: straight-line DFMAs whose multiplicand is a
distinct FP64 immediate (materialised by ptxas as UMOV pairs), K DFMAs per loop iteration, to measure
(a) the FP64 rate of the DFMA+UMOV mix and (b) the instruction-fetch ceiling when the body exceeds the I-cache."""
import random, sys
random.seed(1)
out = ["#include <cstdio>", "#include <cuda_runtime.h>"]
cfgs = []
for K in (256, 1024, 4096, 16384):
    for share in (2, 1):
        name = "k%d_s%d" % (K, share)
        cfgs.append((name, K, share))
        out.append("__global__ void __launch_bounds__(128) %s(double* o, int iters, double seed) {" % name)
        out.append("  double a[8]; for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x * 1e-3;")
        out.append("  double x[4]; for (int i = 0; i < 4; ++i) x[i] = seed * (i + 1);")
        out.append("  for (int it = 0; it < iters; ++it) {")
        n = 0
        while n < K:
            c = random.uniform(0.1, 2.0)
            for s in range(share):
                out.append("    a[%d] = fma(%s, x[%d], a[%d]);" % (n % 8, float(c).hex(), (n // 8) % 4, n % 8))
                n += 1
        out.append("  }")
        out.append("  double s = 0; for (int i = 0; i < 8; ++i) s += a[i]; if (s == 1.2345) o[threadIdx.x] = s;")
        out.append("}")
out.append("""
template <class F> double run(F f, int K, int blocks, int threads, int iters) {
  double* d; cudaMalloc(&d, 4096);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f<<<blocks, threads>>>(d, iters, 1.0); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); f<<<blocks, threads>>>(d, iters, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  cudaFree(d);
  return 2.0 * K * (double)iters * threads * blocks / (best * 1e-3) / 1e12;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int nsm = p.multiProcessorCount;
""")
for name, K, share in cfgs:
    for wps in (4, 8, 16):
        out.append('  printf("{\\"K\\":%d,\\"dfma_per_const\\":%d,\\"warps_per_sm\\":%d,\\"tflops\\":%%.2f}\\n", run(%s, %d, nsm * %d, 128, %d));'
                   % (K, share, wps, name, K, wps // 4, max(1, 4000000 // K // 8)))
out.append("  return 0; }")
open(sys.argv[1], "w").write("\n".join(out) + "\n")
