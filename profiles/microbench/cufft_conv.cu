// COMPARISON POINT ONLY -- never linked into the product library (BASELINE.json north_star: "cuFFT appears only as an ncu
// comparison point").  The reference's GPU path is the external FourierConvolutionCUDALib: pad kernel + cuFFT R2C +
// modulate + cuFFT C2R around host copies.  This program times the device-resident core of that
// design on the same volume, the same padded FFT size and the same conv1 work as one of our FFT-convolution passes:
//     extend (mirror) + pad  ->  cufftExecR2C  ->  spectrum * cached kernel spectrum  ->  cufftExecC2R  ->  crop + ratio
// which is MORE favourable to cuFFT than the legacy library was (kernel spectrum cached, no PCIe, no plan creation, no
// allocation inside the timed region).
//   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a cufft_conv.cu -lcufft -o cufft_conv
//   usage: cufft_conv nz ny nx k Pz Py Px iters      (defaults: 256 512 512 31 288 560 560 20)
// Output: one JSON line (ms per convolution, ms of the two FFTs alone, GB/s against the 44*Np accounting of SURVEY 8d).
#include <cuda_runtime.h>
#include <cufft.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s: %s\"}\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define CF(x) do { cufftResult r_ = (x); if (r_ != CUFFT_SUCCESS) { printf("{\"error\": \"%s: cufft error %d\"}\n", #x, (int)r_); return 1; } } while (0)

__device__ __forceinline__ int mirror_single(int a, int n) {
    if ((unsigned)a < (unsigned)n) return a;
    if (n == 1) return 0;
    const int p = 2 * (n - 1);
    int m = a % p;
    if (m < 0) m += p;
    return m < n ? m : p - m;
}
// position u on a circular axis of length P holding [0, n + h) then the gap then [-h, 0)
__device__ __forceinline__ int pad_coord(int u, int n, int h, int P, bool& gap) {
    gap = false;
    if (u < n + h) return u;
    if (u >= P - h) return u - P;
    gap = true;
    return 0;
}

__global__ void pad_mirror(const float* __restrict__ src, float* __restrict__ dst, int nz, int ny, int nx, int h, int Pz, int Py, int Px) {
    const long long total = (long long)Pz * Py * Px;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % Px);
        const long long r = i / Px;
        const int y = (int)(r % Py), z = (int)(r / Py);
        bool gx, gy, gz;
        const int ax = pad_coord(x, nx, h, Px, gx), ay = pad_coord(y, ny, h, Py, gy), az = pad_coord(z, nz, h, Pz, gz);
        float v = 0.f;
        if (!(gx || gy || gz))
            v = src[((long long)mirror_single(az, nz) * ny + mirror_single(ay, ny)) * nx + mirror_single(ax, nx)];
        dst[i] = v;
    }
}

__global__ void pad_kernel(const float* __restrict__ k, float* __restrict__ dst, int ks, int Pz, int Py, int Px, float scale) {
    const int c = ks / 2;
    const long long total = (long long)ks * ks * ks;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % ks);
        const long long r = i / ks;
        const int y = (int)(r % ks), z = (int)(r / ks);
        const int ux = (x - c + Px) % Px, uy = (y - c + Py) % Py, uz = (z - c + Pz) % Pz;
        dst[((long long)uz * Py + uy) * Px + ux] = k[i] * scale;
    }
}

__global__ void modulate(float2* __restrict__ s, const float2* __restrict__ kh, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float2 a = s[i], b = kh[i];
        s[i] = make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
    }
}

__global__ void crop_ratio(const float* __restrict__ blur, const float* __restrict__ img, float* __restrict__ out, int nz, int ny, int nx, int Py, int Px) {
    const long long total = (long long)nz * ny * nx;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % nx);
        const long long r = i / nx;
        const int y = (int)(r % ny), z = (int)(r / ny);
        const float b = blur[((long long)z * Py + y) * Px + x];
        const float v = img[i];
        out[i] = v > 0.f ? v / b : 1.f;
    }
}

int main(int argc, char** argv) {
    const int nz = argc > 1 ? atoi(argv[1]) : 256, ny = argc > 2 ? atoi(argv[2]) : 512, nx = argc > 3 ? atoi(argv[3]) : 512;
    const int ks = argc > 4 ? atoi(argv[4]) : 31;
    const int Pz = argc > 5 ? atoi(argv[5]) : 288, Py = argc > 6 ? atoi(argv[6]) : 560, Px = argc > 7 ? atoi(argv[7]) : 560;
    const int iters = argc > 8 ? atoi(argv[8]) : 20;
    const int h = ks / 2;
    if (nz < 1 || ny < 1 || nx < 1 || ks < 1 || Pz < nz + ks - 1 || Py < ny + ks - 1 || Px < nx + ks - 1 || iters < 1) {
        printf("{\"error\": \"bad arguments\"}\n");
        return 1;
    }
    const long long N = (long long)nz * ny * nx, Pn = (long long)Pz * Py * Px, Cn = (long long)Pz * Py * (Px / 2 + 1);
    float *d_img, *d_psi, *d_out, *d_pad, *d_k;
    float2 *d_spec, *d_kh;
    CK(cudaMalloc(&d_img, N * 4)); CK(cudaMalloc(&d_psi, N * 4)); CK(cudaMalloc(&d_out, N * 4));
    CK(cudaMalloc(&d_pad, Pn * 4)); CK(cudaMalloc(&d_spec, Cn * 8)); CK(cudaMalloc(&d_kh, Cn * 8));
    CK(cudaMalloc(&d_k, (size_t)ks * ks * ks * 4));
    {
        std::vector<float> hv((size_t)N);
        unsigned s = 12345u;
        for (long long i = 0; i < N; ++i) { s = s * 1664525u + 1013904223u; hv[(size_t)i] = 0.05f + (float)(s >> 8) * (1.0f / 16777216.0f); }
        CK(cudaMemcpy(d_img, hv.data(), N * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_psi, hv.data(), N * 4, cudaMemcpyHostToDevice));
        std::vector<float> hk((size_t)ks * ks * ks, 1.0f / ((float)ks * ks * ks));
        CK(cudaMemcpy(d_k, hk.data(), hk.size() * 4, cudaMemcpyHostToDevice));
    }
    cufftHandle r2c, c2r;
    CF(cufftPlan3d(&r2c, Pz, Py, Px, CUFFT_R2C));
    CF(cufftPlan3d(&c2r, Pz, Py, Px, CUFFT_C2R));
    const int T = 256, B = 148 * 8;
    // cached kernel spectrum, pre-scaled by 1 / (Px Py Pz)
    CK(cudaMemset(d_pad, 0, Pn * 4));
    pad_kernel<<<B, T>>>(d_k, d_pad, ks, Pz, Py, Px, 1.0f / (float)Pn);
    CF(cufftExecR2C(r2c, d_pad, (cufftComplex*)d_kh));
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto conv = [&]() -> int {
        pad_mirror<<<B, T>>>(d_psi, d_pad, nz, ny, nx, h, Pz, Py, Px);
        CF(cufftExecR2C(r2c, d_pad, (cufftComplex*)d_spec));
        modulate<<<B, T>>>(d_spec, d_kh, Cn);
        CF(cufftExecC2R(c2r, (cufftComplex*)d_spec, d_pad));
        crop_ratio<<<B, T>>>(d_pad, d_img, d_out, nz, ny, nx, Py, Px);
        return 0;
    };
    for (int i = 0; i < 3; ++i) if (conv()) return 1;
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) if (conv()) return 1;
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms_conv = 0.f;
    CK(cudaEventElapsedTime(&ms_conv, e0, e1));
    ms_conv /= (float)iters;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) {
        CF(cufftExecR2C(r2c, d_pad, (cufftComplex*)d_spec));
        CF(cufftExecC2R(c2r, (cufftComplex*)d_spec, d_pad));
    }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms_fft = 0.f;
    CK(cudaEventElapsedTime(&ms_fft, e0, e1));
    ms_fft /= (float)iters;
    CK(cudaGetLastError());
    // sanity: a normalised box kernel over data in (0.05, 1.05) gives blur > 0 and a finite ratio
    float probe[4];
    CK(cudaMemcpy(probe, d_out + N / 2, sizeof(probe), cudaMemcpyDeviceToHost));
    const double np_min = (double)(nz + ks - 1) * (ny + ks - 1) * (nx + ks - 1);
    size_t ws_r2c = 0, ws_c2r = 0;
    cufftGetSize(r2c, &ws_r2c); cufftGetSize(c2r, &ws_c2r);
    printf("{\"what\": \"cuFFT-based convolution, comparison only\", \"dims_zyx\": [%d, %d, %d], \"psf\": %d, \"fft_dims_zyx\": [%d, %d, %d], "
           "\"ms_per_conv\": %.4f, \"ms_fft_pair_only\": %.4f, \"gbs_at_44np\": %.1f, \"gbs_fft_pair_at_44np\": %.1f, "
           "\"cufft_workspace_mb\": %.0f, \"iters\": %d, \"ratio_probe\": [%.4f, %.4f, %.4f, %.4f]}\n",
           nz, ny, nx, ks, Pz, Py, Px, ms_conv, ms_fft, 44.0 * np_min / (ms_conv * 1e-3) / 1e9, 44.0 * np_min / (ms_fft * 1e-3) / 1e9,
           (double)(ws_r2c + ws_c2r) / 1048576.0, iters, probe[0], probe[1], probe[2], probe[3]);
    return 0;
}
