// Microbenchmark: HBM bandwidth of the column-pass access pattern.
// Every CTA reads a tile of ROWS rows x CHUNK bytes (row stride = pitch bytes) and writes it back in place,
// exactly like an in-place column FFT pass, for different chunk widths.  Build: nvcc -O3 -arch=sm_100a.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int CHUNK16>   // chunk width in 16-byte units
__global__ void tile_rw(float4* data, int rows, long long pitch16, int tiles_x, long long plane16, int nplanes) {
    const int tile = blockIdx.x;
    const int tx = tile % tiles_x;
    const int pl = tile / tiles_x;
    if (pl >= nplanes) return;
    float4* base = data + pl * plane16 + (long long)tx * CHUNK16;
    const int c = threadIdx.x % CHUNK16;
    const int rstep = blockDim.x / CHUNK16;
    if (rstep == 0 || threadIdx.x >= rstep * CHUNK16) return;
    float4 acc[8];
    for (int r0 = threadIdx.x / CHUNK16; r0 < rows; r0 += 8 * rstep) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = r0 + u * rstep;
            if (r < rows) acc[u] = __ldg(base + (long long)r * pitch16 + c);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = r0 + u * rstep;
            if (r < rows) { float4 v = acc[u]; v.x += 1.f; base[(long long)r * pitch16 + c] = v; }
        }
    }
}

__global__ void flat_rw(float4* data, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float4 v = __ldg(data + i); v.x += 1.f; data[i] = v;
    }
}

template <int C>
float run(float4* d, int rows, long long pitch16, long long plane16, int nplanes, int threads) {
    const int tiles_x = (int)(pitch16 / C);
    const int grid = tiles_x * nplanes;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) tile_rw<C><<<grid, threads>>>(d, rows, pitch16, tiles_x, plane16, nplanes);
    cudaEventRecord(a);
    const int reps = 10;
    for (int i = 0; i < reps; ++i) tile_rw<C><<<grid, threads>>>(d, rows, pitch16, tiles_x, plane16, nplanes);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    const int rows = 560, nplanes = 286;
    const long long pitch16 = 144;               // 288 float2 = 2304 B per row
    const long long plane16 = pitch16 * rows;
    const long long total16 = plane16 * nplanes;
    float4* d; cudaMalloc(&d, total16 * 16); cudaMemset(d, 0, total16 * 16);
    const double bytes = 2.0 * total16 * 16;
    for (int threads : {128, 256, 512}) {
        printf("threads %d:", threads);
        printf("  64B %.0f", bytes / run<4>(d, rows, pitch16, plane16, nplanes, threads) / 1e6);
        printf("  128B %.0f", bytes / run<8>(d, rows, pitch16, plane16, nplanes, threads) / 1e6);
        printf("  256B %.0f", bytes / run<16>(d, rows, pitch16, plane16, nplanes, threads) / 1e6);
        printf("  768B %.0f", bytes / run<48>(d, rows, pitch16, plane16, nplanes, threads) / 1e6);
        if (threads >= 144) printf("  2304B %.0f GB/s", bytes / run<144>(d, rows, pitch16, plane16, nplanes, threads) / 1e6);
        printf("\n"); fflush(stdout);
    }
    // z-pass like pattern: rows strided by a whole plane (288 rows of 128 B, stride 560*2304 B)
    {
        const int zrows = 286;
        const int tiles_x = 18;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        for (int threads : {128, 256}) {
            const int grid = tiles_x * rows;
            // reuse tile_rw with "plane" = one y row, pitch = plane
            for (int i = 0; i < 3; ++i) tile_rw<8><<<grid, threads>>>(d, zrows, plane16, tiles_x, pitch16, rows);
            cudaEventRecord(a);
            for (int i = 0; i < 10; ++i) tile_rw<8><<<grid, threads>>>(d, zrows, plane16, tiles_x, pitch16, rows);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("z-pattern threads %d: 128B chunks stride %lld B: %.0f GB/s\n", threads, plane16 * 16, bytes / (ms / 10) / 1e6);
        }
    }
    {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        for (int i = 0; i < 3; ++i) flat_rw<<<148 * 8, 256>>>(d, total16);
        cudaEventRecord(a);
        for (int i = 0; i < 10; ++i) flat_rw<<<148 * 8, 256>>>(d, total16);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("flat in-place rw: %.0f GB/s\n", bytes / (ms / 10) / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
