// Thin runtime shim: device memory, copies, streams and kernel launch.
// CUDA build: real device calls.  SPIM_HOST_EMU build (tests only): host memory + serial blocks.
#pragma once
#include "hd.h"
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include <stdexcept>
#include <type_traits>
#if defined(SPIM_HOST_EMU)
#include <thread>
#endif

namespace spim { namespace rt {

struct Error : public std::runtime_error {
    explicit Error(const std::string& s) : std::runtime_error(s) {}
};

#if defined(SPIM_HOST_EMU)

typedef int Stream;
inline int device_count() { return 1; }
inline void set_device(int) {}
inline void* dmalloc(size_t n) { void* p = calloc(1, n ? n : 1); if (!p) throw Error("emu: out of memory"); return p; }
inline void dfree(void* p) { free(p); }
inline void h2d(void* d, const void* h, size_t n, Stream) { memcpy(d, h, n); }
inline void d2h(void* h, const void* d, size_t n, Stream) { memcpy(h, d, n); }
inline void d2d(void* d, const void* s, size_t n, Stream) { memmove(d, s, n); }
inline void dzero(void* d, size_t n, Stream) { memset(d, 0, n); }
// copy a host box [ext] (tightly packed floats) into a device volume [ddims] at offset lo; all triples are (z, y, x)
inline void h2d_box(float* d, const int ddims[3], const int lo[3], const float* h, const int ext[3], Stream) {
    for (int z = 0; z < ext[0]; ++z)
        for (int y = 0; y < ext[1]; ++y)
            memcpy(d + ((size_t)(z + lo[0]) * ddims[1] + (y + lo[1])) * ddims[2] + lo[2],
                   h + ((size_t)z * ext[1] + y) * ext[2], (size_t)ext[2] * sizeof(float));
}
inline Stream stream_create() { return 0; }
inline void stream_destroy(Stream) {}
inline void stream_sync(Stream) {}
inline size_t max_smem() { return 227 * 1024; }
inline int sm_count() { return 148; }

// every kernel can run a block as several real threads (hd.h) unless it declares `static constexpr bool kEmuThreads = false`
// (bodies whose emulator branch is written for one thread per block: the TMA pipeline, warp-private columns, push / wait)
template <class B, class = void> struct emu_threaded { static constexpr bool value = true; };
template <class B> struct emu_threaded<B, std::void_t<decltype(B::kEmuThreads)>> { static constexpr bool value = B::kEmuThreads; };
inline int emu_threads() { const char* e = getenv("SPIM_EMU_THREADS"); const int t = e ? atoi(e) : 1; return (t >= 1 && t <= 64) ? t : 1; }

template <class Body, int MAXT = 256, int MINB = 1>
inline void launch(const typename Body::Params& p, long long grid, int /*block*/, size_t smem, Stream) {
    std::vector<double> buf((smem + 7) / 8 + 1);
    float2* sm = reinterpret_cast<float2*>(buf.data());
    const int T = emu_threaded<Body>::value ? emu_threads() : 1;
    if (T <= 1) {
        for (long long b = 0; b < grid; ++b) Body::run(p, (int)b, sm);
        return;
    }
    for (long long b = 0; b < grid; ++b) {      // one block at a time, T threads sharing its "shared memory"
        SpimEmuBlock blk(T);
        std::vector<std::thread> ts;
        for (int t = 0; t < T; ++t)
            ts.emplace_back([&, t]() {
                spim_emu_tid = t; spim_emu_nthr = T; spim_emu_blk = &blk;
                Body::run(p, (int)b, sm);
                spim_emu_tid = 0; spim_emu_nthr = 1; spim_emu_blk = nullptr;
            });
        for (auto& th : ts) th.join();
    }
}

inline bool encode_tensor_map(SpimTensorMap*, void*, int, const unsigned long long*, const unsigned long long*, const unsigned int*) { return false; }

struct KernelTimer {
    void enable(bool) {}
    void begin(int, Stream) {}
    void end(int, Stream) {}
    void collect(double*, long long*, int n) { (void)n; }
    void reset() {}
};

#else

#define SPIM_CUDA_CHECK(expr)                                                                        \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            throw ::spim::rt::Error(std::string(#expr) + ": " + cudaGetErrorString(e__));            \
    } while (0)

typedef cudaStream_t Stream;
inline int device_count() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e == cudaErrorNoDevice) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) { cudaGetLastError(); return -1; }
    return n;
}
inline void set_device(int d) { SPIM_CUDA_CHECK(cudaSetDevice(d)); }
inline void* dmalloc(size_t n) { void* p = nullptr; SPIM_CUDA_CHECK(cudaMalloc(&p, n ? n : 1)); return p; }
inline void dfree(void* p) { if (p) cudaFree(p); }
// the source may be pageable / pinned host memory or (unified addressing) memory of any device: views that already live on
// a GPU -- generated there, or produced by the fusion pre-step of another session -- are handed over through the same entry points
inline void h2d(void* d, const void* h, size_t n, Stream s) { SPIM_CUDA_CHECK(cudaMemcpyAsync(d, h, n, cudaMemcpyDefault, s)); }
inline void d2h(void* h, const void* d, size_t n, Stream s) { SPIM_CUDA_CHECK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s)); }
inline void d2d(void* d, const void* s_, size_t n, Stream s) { SPIM_CUDA_CHECK(cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s)); }
inline void dzero(void* d, size_t n, Stream s) { SPIM_CUDA_CHECK(cudaMemsetAsync(d, 0, n, s)); }
// copy a host box [ext] (tightly packed floats) into a device volume [ddims] at offset lo; all triples are (z, y, x)
inline void h2d_box(float* d, const int ddims[3], const int lo[3], const float* h, const int ext[3], Stream s) {
    cudaMemcpy3DParms q;
    memset(&q, 0, sizeof(q));
    q.srcPtr = make_cudaPitchedPtr(const_cast<float*>(h), (size_t)ext[2] * sizeof(float), (size_t)ext[2], (size_t)ext[1]);
    q.dstPtr = make_cudaPitchedPtr(d, (size_t)ddims[2] * sizeof(float), (size_t)ddims[2], (size_t)ddims[1]);
    q.dstPos = make_cudaPos((size_t)lo[2] * sizeof(float), (size_t)lo[1], (size_t)lo[0]);
    q.extent = make_cudaExtent((size_t)ext[2] * sizeof(float), (size_t)ext[1], (size_t)ext[0]);
    q.kind = cudaMemcpyDefault;
    SPIM_CUDA_CHECK(cudaMemcpy3DAsync(&q, s));
}
inline Stream stream_create() { Stream s; SPIM_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); return s; }
inline void stream_destroy(Stream s) { cudaStreamDestroy(s); }
inline void stream_sync(Stream s) { SPIM_CUDA_CHECK(cudaStreamSynchronize(s)); }
inline size_t max_smem() {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    return (size_t)v;
}
inline int sm_count() {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
}

// (Programmatic dependent launch was measured in round 2 in two forms -- every block releasing the successor at its start,
// and only the blocks of the last wave doing so -- and removed: per-kernel times were unchanged and whole iterations 2-7 %
// slower than with plain launches, because the successor's parked blocks take slots from the predecessor's own tail.)
template <class Body, int MAXT>
__global__ void __launch_bounds__(MAXT) kernel_entry(const __grid_constant__ typename Body::Params p) {
    extern __shared__ __align__(1024) unsigned char spim_smem[];
    Body::run(p, (int)blockIdx.x, reinterpret_cast<float2*>(spim_smem));
}
// MINB blocks of MAXT threads per SM must fit (caps the registers per thread)
template <class Body, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) kernel_entry_capped(const __grid_constant__ typename Body::Params p) {
    extern __shared__ __align__(1024) unsigned char spim_smem[];
    Body::run(p, (int)blockIdx.x, reinterpret_cast<float2*>(spim_smem));
}

template <class K, class P>
inline void launch_kernel(K kernel, const P& p, unsigned grid, int block, size_t smem, Stream s) {
    kernel<<<grid, block, smem, s>>>(p);
}

// not `inline`: in the split build (instances.h) the engine kernels' launches are declared `extern template`, which only
// suppresses the implicit instantiation of non-inline templates
template <class Body, int MAXT = 256, int MINB = 1>
void launch(const typename Body::Params& p, long long grid, int block, size_t smem, Stream s) {
    if (grid <= 0) return;
    static thread_local size_t configured[64] = {0};   // per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (block > MAXT) block = MAXT;
    if constexpr (MINB > 1) {
        if (dev < 64 && smem > configured[dev]) {
            SPIM_CUDA_CHECK(cudaFuncSetAttribute(kernel_entry_capped<Body, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured[dev] = smem;
        }
        launch_kernel(kernel_entry_capped<Body, MAXT, MINB>, p, (unsigned)grid, block, smem, s);
    } else {
        if (dev < 64 && smem > configured[dev]) {
            SPIM_CUDA_CHECK(cudaFuncSetAttribute(kernel_entry<Body, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured[dev] = smem;
        }
        launch_kernel(kernel_entry<Body, MAXT>, p, (unsigned)grid, block, smem, s);
    }
    SPIM_CUDA_CHECK(cudaGetLastError());
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
inline bool encode_tensor_map(SpimTensorMap* out, void* base, int rank, const unsigned long long* dims,
                              const unsigned long long* strides_bytes /* rank-1 */, const unsigned int* box) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr) {
            cudaGetLastError();
            return false;
        }
        fn = (EncodeFn)ptr;
    }
    cuuint64_t gdim[5]; cuuint64_t gstr[4]; cuuint32_t bx[5]; cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, base, gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// optional per-kernel-class timing with CUDA events on the launching stream
struct KernelTimer {
    bool on = false;
    struct Rec { int id; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t cur = nullptr;
    void enable(bool v) { on = v; }
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; SPIM_CUDA_CHECK(cudaEventCreate(&e)); return e;
    }
    void begin(int, Stream s) { if (!on) return; cur = get(); SPIM_CUDA_CHECK(cudaEventRecord(cur, s)); }
    void end(int id, Stream s) {
        if (!on) return;
        cudaEvent_t b = get();
        SPIM_CUDA_CHECK(cudaEventRecord(b, s));
        recs.push_back(Rec{id, cur, b});
    }
    // accumulate milliseconds and launch counts per id (caller synchronised the stream)
    void collect(double* ms, long long* cnt, int n) {
        for (auto& r : recs) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && r.id < n) { ms[r.id] += t; cnt[r.id] += 1; }
            pool.push_back(r.a); pool.push_back(r.b);
        }
        recs.clear();
    }
    void reset() { for (auto& r : recs) { pool.push_back(r.a); pool.push_back(r.b); } recs.clear(); }
    ~KernelTimer() { for (auto e : pool) cudaEventDestroy(e); for (auto& r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); } }
};

#endif

}}  // namespace spim::rt
