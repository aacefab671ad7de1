// DRAM efficiency of the step kernels' access pattern: a field stored row-major [rows][N] (complex(8), 16 B per node),
// every CTA copies ONE tile of `tile` consecutive nodes of all `rows` rows (chunks of tile*16 bytes, N*16 bytes apart)
// and writes them to a second array.  Compares chunk sizes 256 B .. 8 KB with a flat copy of the same bytes.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/rowcopy_bench tools/rowcopy_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void tile_copy(const double2* __restrict__ in, double2* __restrict__ out, long long N, int rows, int tile) {
    const long long node0 = (long long)blockIdx.x * tile;
    for (int t = threadIdx.x; t < tile; t += blockDim.x) {
        const long long p = node0 + t;
        if (p >= N) break;
#pragma unroll 5
        for (int r = 0; r < rows; ++r) out[(long long)r * N + p] = in[(long long)r * N + p];
    }
}
__global__ void flat_copy(const double2* __restrict__ in, double2* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = in[i];
}

int main() {
    const long long N = 1000000;
    const int rows = 45;
    double2 *a, *b;
    cudaMalloc(&a, sizeof(double2) * N * rows);
    cudaMalloc(&b, sizeof(double2) * N * rows);
    cudaMemset(a, 0, sizeof(double2) * N * rows);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = 2.0 * 16 * N * rows;
    printf("{\"nodes\": %lld, \"rows\": %d, \"bytes_per_pass\": %.0f", N, rows, bytes);
    for (int it = 0; it < 2; ++it) flat_copy<<<148 * 16, 256>>>(a, b, N * rows);
    cudaEventRecord(e0);
    for (int it = 0; it < 10; ++it) flat_copy<<<148 * 16, 256>>>(a, b, N * rows);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf(", \"flat_GBs\": %.0f, \"tile_GBs\": {", bytes * 10 / ms / 1e6);
    const int tiles[] = {16, 32, 64, 128, 256, 512};
    for (int k = 0; k < 6; ++k) {
        const int tile = tiles[k], thr = tile < 32 ? 32 : (tile > 256 ? 256 : tile);
        const unsigned grid = (unsigned)((N + tile - 1) / tile);
        for (int it = 0; it < 2; ++it) tile_copy<<<grid, thr>>>(a, b, N, rows, tile);
        cudaEventRecord(e0);
        for (int it = 0; it < 10; ++it) tile_copy<<<grid, thr>>>(a, b, N, rows, tile);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%s\"%d\": %.0f", k ? ", " : "", tile * 16, bytes * 10 / ms / 1e6);
    }
    printf("}}\n");
    return cudaGetLastError() != cudaSuccess;
}
