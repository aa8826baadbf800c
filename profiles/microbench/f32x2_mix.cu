#include <cuda_runtime.h>
#include <cstdio>
// FP : INT mix test: per element-pair NF fp ops (scalar: 2*NF instr, packed: NF instr) + NI integer ops
template <int PACKED, int NI>
__global__ void __launch_bounds__(256) k(float2* p, int* q, int n, float2 c, float2 s) {
    float2 a[8]; int z[8];
    const int t = threadIdx.x + blockIdx.x * blockDim.x;
    for (int i = 0; i < 8; ++i) { a[i] = p[t + i * 1024]; z[i] = q[t + i * 1024]; }
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (PACKED) { a[i] = __ffma2_rn(a[i], c, s); a[i] = __fadd2_rn(a[i], a[(i + 1) & 7]); }
            else { a[i].x = fmaf(a[i].x, c.x, s.x); a[i].y = fmaf(a[i].y, c.y, s.y); a[i].x += a[(i + 1) & 7].x; a[i].y += a[(i + 1) & 7].y; }
#pragma unroll
            for (int j = 0; j < NI; ++j) z[i] = (z[i] ^ z[(i + j + 1) & 7]) + it;
        }
    }
    for (int i = 0; i < 8; ++i) { p[t + i * 1024] = a[i]; q[t + i * 1024] = z[i]; }
}
template <int PACKED, int NI> void run(float2* p, int* q, const char* name) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int n = 2048, blocks = 148 * 8, th = 256;
    float ms;
    k<PACKED, NI><<<blocks, th>>>(p, q, 16, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f));
    cudaEventRecord(e0); k<PACKED, NI><<<blocks, th>>>(p, q, n, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f)); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double cyc = ms * 1e-3 * 1.965e9;
    double per_iter = cyc / n / 8 / (blocks * th / 32.0 / (148 * 4));     // SMSP cycles per (warp, element-pair)
    printf("%-28s %.3f ms   %.2f SMSP-cycles per warp per element (fp: %d instr, int: %d instr)\n", name, ms, per_iter, PACKED ? 2 : 4, 2 * NI);
}
int main() {
    float2* p; int* q; cudaMalloc(&p, 1 << 26); cudaMemset(p, 0, 1 << 26); cudaMalloc(&q, 1 << 26); cudaMemset(q, 0, 1 << 26);
    run<0, 0>(p, q, "scalar fp only"); run<1, 0>(p, q, "packed fp only");
    run<0, 1>(p, q, "scalar fp + 2 int"); run<1, 1>(p, q, "packed fp + 2 int");
    run<0, 2>(p, q, "scalar fp + 4 int"); run<1, 2>(p, q, "packed fp + 4 int");
    run<0, 4>(p, q, "scalar fp + 8 int"); run<1, 4>(p, q, "packed fp + 8 int");
    return 0;
}
