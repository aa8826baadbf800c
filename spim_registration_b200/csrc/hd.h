// Portability shim: the kernel bodies in this directory are written once and compiled
//  (a) by nvcc for sm_100a  -> the product library, and
//  (b) by g++ with -DSPIM_HOST_EMU -> a *test-only* emulator (tests/emu) that executes each
//      thread block's phases serially on the CPU so index math can be debugged without a GPU.
// The emulator is never loaded by the package; the product path has no CPU fallback.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(SPIM_HOST_EMU)

#include <condition_variable>
#include <mutex>

struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
struct alignas(8) int2 { int x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#define SPIM_DEV inline
#define SPIM_NOINLINE_DEV static __attribute__((noinline))
#define SPIM_HD inline
// Work items and barriers.  Normally the emulator runs a block as ONE thread (tid 0 of 1): the loop below visits every item
// and a barrier is nothing.  With SPIM_EMU_THREADS=T every kernel that does not opt out (kEmuThreads = false) runs each
// block as T real threads that split the items like the threads of a CUDA block and meet at real
// barriers -- the mode tests/test_tsan_kernels.py runs under ThreadSanitizer to check barrier placement without a GPU.
struct SpimEmuBlock {
    int nthr, count = 0, gen = 0;
    std::mutex m;
    std::condition_variable cv;
    explicit SpimEmuBlock(int n) : nthr(n) {}
    void wait() {
        std::unique_lock<std::mutex> l(m);
        const int g = gen;
        if (++count == nthr) { count = 0; ++gen; cv.notify_all(); }
        else cv.wait(l, [&] { return gen != g; });
    }
};
inline thread_local int spim_emu_tid = 0;
inline thread_local int spim_emu_nthr = 1;
inline thread_local SpimEmuBlock* spim_emu_blk = nullptr;
static inline void spim_emu_barrier() { if (spim_emu_blk) spim_emu_blk->wait(); }
#define SPIM_FOR_ITEMS(i, n) for (int i = spim_emu_tid; i < (int)(n); i += spim_emu_nthr)
#define SPIM_BARRIER() spim_emu_barrier()
// a thread group = the threads that cooperate on one tile (the whole CTA, or one consumer group of a
// warp-specialised kernel); the emulator runs every group as a single serial thread
struct alignas(64) SpimTensorMap { unsigned long long opaque[16]; };   // CUtensorMap stand-in (unused by the emulator)
struct TG { int tid, n, bar; };
static inline TG tg_cta() { TG t; t.tid = spim_emu_tid; t.n = spim_emu_nthr; t.bar = 0; return t; }
static inline void tg_barrier(const TG&) { spim_emu_barrier(); }
#define SPIM_FOR_ITEMS_TG(tg, i, cnt) for (int i = (tg).tid; i < (int)(cnt); i += (tg).n)
static inline void spim_syncwarp() {}
#define SPIM_NTHREADS spim_emu_nthr
#define SPIM_TID spim_emu_tid
template <class T> static inline T spim_ldg(const T* p) { return *p; }
static inline float4 ldg_stream(const float4* p) { return *p; }
static inline float2 ldg_stream(const float2* p) { return *p; }
static inline void stg_stream(float4* p, float4 v) { *p = v; }
static inline void stg_stream_if(float4* p, float4 v, bool ok) { if (ok) *p = v; }
static inline uint32_t spim_umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
// cp.async (LDGSTS) shims: the emulator copies immediately
static inline void cp_async16(void* smem_dst, const void* gsrc) { memcpy(smem_dst, gsrc, 16); }
static inline void cp_async_commit() {}
template <int N> static inline void cp_async_wait() {}
static inline float spim_fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float spim_fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float spim_fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float spim_fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float spim_fsqrt_rn(float a) { volatile float r = sqrtf(a); return r; }
static inline float spim_fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
// un-fused fp64 ops (the Java double arithmetic of the fusion pre-step has no FMA contraction)
static inline double spim_dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double spim_dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double spim_dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double spim_ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline float spim_d2f_rn(double a) { volatile float r = (float)a; return r; }
// seeds of the fast epilogue (csrc/fast_math.h): the emulator uses correctly rounded values, the GPU MUFU approximations
static inline float spim_rcp_seed(float b) { volatile float r = 1.0f / b; return r; }
static inline float spim_rsqrt_seed(float x) { volatile float r = (float)(1.0 / sqrt((double)x)); return r; }
#define SPIM_FM_HD static inline
#define SPIM_FM_MUL(a, b) spim_fmul_rn(a, b)

#else

#include <cuda_runtime.h>
#include <cuda.h>
typedef CUtensorMap SpimTensorMap;
#define SPIM_DEV __device__ __forceinline__
#define SPIM_NOINLINE_DEV static __device__ __noinline__
#define SPIM_HD __host__ __device__ __forceinline__
#define SPIM_FOR_ITEMS(i, n) for (int i = (int)threadIdx.x; i < (int)(n); i += (int)blockDim.x)
#define SPIM_BARRIER() __syncthreads()
struct TG { int tid, n, bar; };   // thread index in the group, group size, named barrier id (0 = whole CTA)
__device__ __forceinline__ TG tg_cta() { TG t; t.tid = (int)threadIdx.x; t.n = (int)blockDim.x; t.bar = 0; return t; }
__device__ __forceinline__ void tg_barrier(const TG& tg) {
    if (tg.bar == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(tg.bar), "r"(tg.n) : "memory");
}
#define SPIM_FOR_ITEMS_TG(tg, i, cnt) for (int i = (tg).tid; i < (int)(cnt); i += (tg).n)
__device__ __forceinline__ void spim_syncwarp() { __syncwarp(); }
#define SPIM_NTHREADS ((int)blockDim.x)
#define SPIM_TID ((int)threadIdx.x)
template <class T> __device__ __forceinline__ T spim_ldg(const T* p) { return __ldg(p); }
// streaming accesses for data touched exactly once per kernel: do not allocate in the (tiny, smem-carved) L1
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ldg_stream(const float2* p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// the same store under a predicate that lives inside the instruction (@p st...) instead of a branch around it:
// rows beyond the kept range are skipped without reconvergence scaffolding
__device__ __forceinline__ void stg_stream_if(float4* p, float4 v, bool ok) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %5, 0;\n\t@q st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};\n\t}"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"((int)ok) : "memory");
}
__device__ __forceinline__ uint32_t spim_umulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
// asynchronous 16-byte global -> shared copies (LDGSTS), grouped and awaited per thread
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// ---- mbarrier + TMA bulk-copy helpers (sm_90+/sm_100a) ------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) { while (!mbar_try_wait(bar, parity)) {} }
// TMA bulk copy global -> shared (UBLKCP), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// TMA tensor copies global -> shared (UTMALDG): one request moves a whole [rows][128 B] box
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const SpimTensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst_smem)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst_smem, const SpimTensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst_smem)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
// pull the 16-byte granules that lie entirely inside [p, p + n floats) into L2 (no data is returned, nothing is ordered)
__device__ __forceinline__ void bulk_prefetch_l2(const float* p, int n) {
    const unsigned long long a = (reinterpret_cast<unsigned long long>(p) + 15ull) & ~15ull;
    const unsigned long long e = (reinterpret_cast<unsigned long long>(p) + 4ull * (unsigned long long)n) & ~15ull;
    if (e > a) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"((unsigned)(e - a)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// explicitly un-fused fp32 ops: the reference's Java float arithmetic has no FMA contraction
__device__ __forceinline__ float spim_fadd_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float spim_fsub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float spim_fmul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float spim_fdiv_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float spim_fsqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ float spim_fmaf_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
// un-fused fp64 ops (the Java double arithmetic of the fusion pre-step has no FMA contraction)
__device__ __forceinline__ double spim_dadd_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double spim_dsub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double spim_dmul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double spim_ddiv_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float spim_d2f_rn(double a) { return __double2float_rn(a); }
// seeds of the fast epilogue (csrc/fast_math.h): MUFU.RCP (<= 1 ulp) and MUFU.RSQ (<= 2 ulp)
__device__ __forceinline__ float spim_rcp_seed(float b) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b)); return r; }
__device__ __forceinline__ float spim_rsqrt_seed(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#define SPIM_FM_HD __device__ __forceinline__
#define SPIM_FM_MUL(a, b) __fmul_rn(a, b)

#endif
