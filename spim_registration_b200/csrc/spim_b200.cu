// B200-native multi-view deconvolution library: C ABI (include/spim_fftconv.h, include/spim_mvdecon.h)
// over the FFT-convolution engine (engine.h / kernels.h).
//
// Built two ways (see hd.h): nvcc -gencode arch=compute_100a,code=sm_100a -> the product .so;
// g++ -x c++ -DSPIM_HOST_EMU -> tests/emu/libspim_emu.so, a test-only CPU emulator of the kernels.
#include "engine.h"
#include "../../include/spim_fftconv.h"
#include "../../include/spim_mvdecon.h"

#include <string>
#include <vector>
#include <mutex>
#include <time.h>
#include <unistd.h>
#include <map>
#include <array>

using namespace spim;

// ================================================================================================
// element-wise kernels (initialisation work; not on the per-iteration path)
// ================================================================================================
namespace spim {

constexpr int kChunk = 2048;

struct EwGrid { long long n; int nblocks; };

struct FillK {
    struct Params { float* p; long long n; float v; int nblocks; };
    SPIM_DEV static void run(const Params& p, int bid, float2*) {
        for (long long base = (long long)bid * kChunk; base < p.n; base += (long long)p.nblocks * kChunk) {
            SPIM_FOR_ITEMS(i, kChunk) { const long long idx = base + i; if (idx < p.n) p.p[idx] = p.v; }
        }
    }
};

struct ScaleClampK {   // adjustOSEMspeedup: w <- min(1, w * (float)osem)
    struct Params { float* p; long long n; float f; int nblocks; };
    SPIM_DEV static void run(const Params& p, int bid, float2*) {
        for (long long base = (long long)bid * kChunk; base < p.n; base += (long long)p.nblocks * kChunk) {
            SPIM_FOR_ITEMS(i, kChunk) {
                const long long idx = base + i;
                if (idx < p.n) {
                    const float x = spim_fmul_rn(p.p[idx], p.f);
                    p.p[idx] = (x < 1.f || x != x) ? x : 1.f;     // Math.min(1, x): NaN stays NaN (fminf would drop it)
                }
            }
        }
    }
};

#if defined(SPIM_HOST_EMU)
// the emulator's stand-ins for the device atomics: one lock (a block may run as several real threads, hd.h)
static std::mutex g_emu_atomic_mutex;
SPIM_DEV void atomic_add_d(double* p, double v) { std::lock_guard<std::mutex> l(g_emu_atomic_mutex); *p += v; }
SPIM_DEV void atomic_add_u64(unsigned long long* p, unsigned long long v) { std::lock_guard<std::mutex> l(g_emu_atomic_mutex); *p += v; }
SPIM_DEV void atomic_min_u32(unsigned int* p, unsigned int v) { std::lock_guard<std::mutex> l(g_emu_atomic_mutex); if (v < *p) *p = v; }
SPIM_DEV double warp_sum_d(double v) { return v; }
SPIM_DEV unsigned long long warp_sum_u64(unsigned long long v) { return v; }
SPIM_DEV unsigned int warp_min_u32(unsigned int v) { return v; }
SPIM_DEV bool warp_leader() { return true; }
#else
SPIM_DEV void atomic_add_d(double* p, double v) { atomicAdd(p, v); }
SPIM_DEV void atomic_add_u64(unsigned long long* p, unsigned long long v) { atomicAdd(p, v); }
SPIM_DEV void atomic_min_u32(unsigned int* p, unsigned int v) { atomicMin(p, v); }
SPIM_DEV double warp_sum_d(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
SPIM_DEV unsigned long long warp_sum_u64(unsigned long long v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
SPIM_DEV unsigned int warp_min_u32(unsigned int v) {
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
SPIM_DEV bool warp_leader() { return (threadIdx.x & 31) == 0; }
#endif

// psi initialisation statistics.
//  gen-2 (FirstIteration.java:122-154): sum over voxels with >=1 view img>0 of the mean of those views.
//  gen-1 (D2/AdjustInput.java:136-167): views counted where weight != 0; intensity sum where >= 2 views.
struct InitStatsK {
    struct Params { ViewPtrs v; long long n; int gen; int nblocks; double* sum; unsigned long long* cnt; unsigned int* min_overlap; };
    SPIM_DEV static void run(const Params& p, int bid, float2*) {
        double s = 0.0;
        unsigned long long c0 = 0, c1 = 0, c2 = 0;
        unsigned int mn = 0xffffffffu;
        for (long long base = (long long)bid * kChunk; base < p.n; base += (long long)p.nblocks * kChunk) {
            SPIM_FOR_ITEMS(i, kChunk) {
                const long long idx = base + i;
                if (idx >= p.n) continue;
                double loc = 0.0;
                int cnt = 0;
                for (int v = 0; v < p.v.nviews; ++v) {
                    const float im = spim_ldg(p.v.img[v] + idx);
                    bool use;
                    if (p.gen == 2) use = im > 0.f;
                    else use = p.v.w[v] ? (spim_ldg(p.v.w[v] + idx) != 0.f) : true;
                    if (use) { loc += (double)im; ++cnt; }
                }
                if (p.gen == 2) {
                    if (cnt > 0) { s += loc / (double)cnt; ++c0; }
                } else {
                    if (cnt > 1) { s += loc; c0 += (unsigned long long)cnt; }
                    if (cnt > 0) { c1 += (unsigned long long)cnt; ++c2; if ((unsigned)cnt < mn) mn = (unsigned)cnt; }
                }
            }
        }
        s = warp_sum_d(s); c0 = warp_sum_u64(c0); c1 = warp_sum_u64(c1); c2 = warp_sum_u64(c2); mn = warp_min_u32(mn);
        if (warp_leader()) {
            atomic_add_d(p.sum, s);
            atomic_add_u64(p.cnt + 0, c0); atomic_add_u64(p.cnt + 1, c1); atomic_add_u64(p.cnt + 2, c2);
            atomic_min_u32(p.min_overlap, mn);
        }
    }
};

struct MaskK {   // MVDeconvolution.java:201-208: psi <- 0 where no view has img > 0
    struct Params { ViewPtrs v; float* psi; int pdims[3]; int porigin[3]; int n[3]; int nblocks; };
    SPIM_DEV static void run(const Params& p, int bid, float2*) {
        const long long total = (long long)p.n[0] * p.n[1] * p.n[2];
        for (long long base = (long long)bid * kChunk; base < total; base += (long long)p.nblocks * kChunk) {
            SPIM_FOR_ITEMS(i, kChunk) {
                const long long idx = base + i;
                if (idx >= total) continue;
                bool any = false;
                for (int v = 0; v < p.v.nviews; ++v) any = any || (spim_ldg(p.v.img[v] + idx) > 0.f);
                if (!any) {
                    const int x = (int)(idx % p.n[2]);
                    const long long r = idx / p.n[2];
                    const int y = (int)(r % p.n[1]);
                    const int z = (int)(r / p.n[1]);
                    p.psi[((long long)(z + p.porigin[0]) * p.pdims[1] + (y + p.porigin[1])) * p.pdims[2] + x + p.porigin[2]] = 0.f;
                }
            }
        }
    }
};

// copy between an unpadded [n] volume and the interior of a haloed buffer
struct CopyRegionK {
    struct Params { const float* src; float* dst; int sdims[3], sorigin[3], ddims[3], dorigin[3], n[3]; int nblocks; };
    SPIM_DEV static void run(const Params& p, int bid, float2*) {
        const long long total = (long long)p.n[0] * p.n[1] * p.n[2];
        for (long long base = (long long)bid * kChunk; base < total; base += (long long)p.nblocks * kChunk) {
            SPIM_FOR_ITEMS(i, kChunk) {
                const long long idx = base + i;
                if (idx >= total) continue;
                const int x = (int)(idx % p.n[2]);
                const long long r = idx / p.n[2];
                const int y = (int)(r % p.n[1]);
                const int z = (int)(r / p.n[1]);
                const float v = p.src[((long long)(z + p.sorigin[0]) * p.sdims[1] + (y + p.sorigin[1])) * p.sdims[2] + x + p.sorigin[2]];
                p.dst[((long long)(z + p.dorigin[0]) * p.ddims[1] + (y + p.dorigin[1])) * p.ddims[2] + x + p.dorigin[2]] = v;
            }
        }
    }
};

// brick mode: fill one halo face of a haloed buffer from its interior by the out-of-bounds rule
struct HaloFillK {
    struct Params { float* buf; int dims[3], origin[3], n[3]; int axis, side, width; int ext; float value; int nblocks; };
    SPIM_DEV static void run(const Params& p, int bid, float2*) {
        int ext_[3] = {p.dims[0], p.dims[1], p.dims[2]};
        ext_[p.axis] = p.width;
        const long long total = (long long)ext_[0] * ext_[1] * ext_[2];
        for (long long base = (long long)bid * kChunk; base < total; base += (long long)p.nblocks * kChunk) {
            SPIM_FOR_ITEMS(i, kChunk) {
                const long long idx = base + i;
                if (idx >= total) continue;
                int c[3];
                c[2] = (int)(idx % ext_[2]);
                const long long r = idx / ext_[2];
                c[1] = (int)(r % ext_[1]);
                c[0] = (int)(r / ext_[1]);
                // coordinate along the axis (array index), lo side: [origin-width, origin), hi side: [origin+n, origin+n+width)
                const int ai = p.side == 0 ? p.origin[p.axis] - p.width + c[p.axis] : p.origin[p.axis] + p.n[p.axis] + c[p.axis];
                const int a = ai - p.origin[p.axis];
                const int e = ext_map(a, p.n[p.axis], p.ext);
                int d[3] = {c[0], c[1], c[2]};
                d[p.axis] = ai;
                float v = p.value;
                if (e >= 0) {
                    int s[3] = {c[0], c[1], c[2]};
                    s[p.axis] = e + p.origin[p.axis];
                    v = p.buf[((long long)s[0] * p.dims[1] + s[1]) * p.dims[2] + s[2]];
                }
                p.buf[((long long)d[0] * p.dims[1] + d[1]) * p.dims[2] + d[2]] = v;
            }
        }
    }
};

// brick mode: gather / scatter up to 26 box-shaped pieces of a haloed buffer to / from one flat staging buffer
constexpr int MAX_PIECES = 26;
struct HaloPackK {
    struct Params {
        float* buf; float* flat; int dims[3]; int npieces; int unpack; int nblocks;
        int lo[MAX_PIECES][3]; int ext[MAX_PIECES][3]; long long off[MAX_PIECES + 1];
    };
    SPIM_DEV static void run(const Params& p, int bid, float2*) {
        const long long total = p.off[p.npieces];
        for (long long base = (long long)bid * kChunk; base < total; base += (long long)p.nblocks * kChunk) {
            SPIM_FOR_ITEMS(i, kChunk) {
                const long long idx = base + i;
                if (idx >= total) continue;
                int pc = 0;
                while (pc + 1 < p.npieces && idx >= p.off[pc + 1]) ++pc;
                long long r = idx - p.off[pc];
                const int x = (int)(r % p.ext[pc][2]); r /= p.ext[pc][2];
                const int y = (int)(r % p.ext[pc][1]);
                const int z = (int)(r / p.ext[pc][1]);
                const long long a = ((long long)(z + p.lo[pc][0]) * p.dims[1] + (y + p.lo[pc][1])) * p.dims[2] + x + p.lo[pc][2];
                if (p.unpack) p.buf[a] = p.flat[idx]; else p.flat[idx] = p.buf[a];
            }
        }
    }
};

// emulator only: memory orders of the epoch flags.  tests/cpp/p2p_tsan_driver.cpp builds a negative control with
// -DSPIM_EMU_RELAXED_FLAGS, in which ThreadSanitizer must find the races the release / acquire pair prevents.
#if defined(SPIM_EMU_RELAXED_FLAGS)
#define SPIM_EMU_FLAG_STORE_ORDER __ATOMIC_RELAXED
#define SPIM_EMU_FLAG_LOAD_ORDER __ATOMIC_RELAXED
#else
#define SPIM_EMU_FLAG_STORE_ORDER __ATOMIC_RELEASE
#define SPIM_EMU_FLAG_LOAD_ORDER __ATOMIC_ACQUIRE
#endif

// brick mode, direct halo push (fused copy + signal over peer memory, no staging buffer and no NCCL on the data path):
// up to 26 box-shaped pieces of MY buffer are stored straight into the neighbours' halos through their peer-mapped
// (CUDA IPC / NVLink) buffers; the last block to finish raises this buffer's epoch flag at every neighbour with a
// system-scope release store.  All bricks share one geometry, so a destination is (peer base pointer, box origin).
struct HaloPushK {
    static constexpr bool kEmuThreads = false;     // the emulated signal step assumes one thread per block
    struct Params {
        const float* buf; int dims[3]; int npieces; int nblocks;
        float* dst[MAX_PIECES];            // base of the neighbour's buffer
        unsigned int* flag[MAX_PIECES];    // the neighbour's flag word for (this buffer, my direction)
        int lo[MAX_PIECES][3]; int dlo[MAX_PIECES][3]; int ext[MAX_PIECES][3]; long long off[MAX_PIECES + 1];
        int vsh[MAX_PIECES];               // 2: the piece's rows are moved as float4 (x origin, extent and row length are
                                           // multiples of 4 -- the y / z faces, i.e. almost all bytes), 0: float by float
        unsigned int* done;                // my block counter
        unsigned int* epoch;               // my push counter for this buffer
    };
    SPIM_DEV static void run(const Params& p, int bid, float2*) {
        const long long total = p.off[p.npieces];      // in items: float4 for vectorised pieces, float otherwise
        for (long long base = (long long)bid * kChunk; base < total; base += (long long)p.nblocks * kChunk) {
            SPIM_FOR_ITEMS(i, kChunk) {
                const long long idx = base + i;
                if (idx >= total) continue;
                int pc = 0;
                while (pc + 1 < p.npieces && idx >= p.off[pc + 1]) ++pc;
                const int sh = p.vsh[pc];
                const int ex = p.ext[pc][2] >> sh;
                long long r = idx - p.off[pc];
                const int x = (int)(r % ex) << sh; r /= ex;
                const int y = (int)(r % p.ext[pc][1]);
                const int z = (int)(r / p.ext[pc][1]);
                const long long a = ((long long)(z + p.lo[pc][0]) * p.dims[1] + (y + p.lo[pc][1])) * p.dims[2] + x + p.lo[pc][2];
                const long long b = ((long long)(z + p.dlo[pc][0]) * p.dims[1] + (y + p.dlo[pc][1])) * p.dims[2] + x + p.dlo[pc][2];
                if (sh) *reinterpret_cast<float4*>(p.dst[pc] + b) = *reinterpret_cast<const float4*>(p.buf + a);
                else p.dst[pc][b] = p.buf[a];
            }
        }
#if defined(SPIM_HOST_EMU)
        if (bid == p.nblocks - 1) {        // blocks run in order: every piece has been written
            const unsigned int e = *p.epoch + 1u;
            *p.epoch = e;
            for (int i = 0; i < p.npieces; ++i) __atomic_store_n(p.flag[i], e, SPIM_EMU_FLAG_STORE_ORDER);
        }
#else
        __threadfence_system();            // my stores are visible system-wide ...
        __syncthreads();                   // ... before thread 0 counts this block as done
        if (threadIdx.x == 0) {
            const unsigned int t = atomicAdd(p.done, 1u);
            if (t == (unsigned int)p.nblocks - 1u) {
                *p.done = 0u;              // ready for the next launch (stream order)
                __threadfence_system();
                const unsigned int e = *p.epoch + 1u;
                *p.epoch = e;
                for (int i = 0; i < p.npieces; ++i)
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.flag[i]), "r"(e) : "memory");
            }
        }
#endif
    }
};

// the matching wait: one thread per neighbour spins (system-scope acquire loads) until that neighbour's flag has reached
// my own epoch for this buffer -- every rank pushes the same sequence, so equal epochs pair up.  A timeout raises an
// error word instead of hanging the GPU.
struct HaloWaitK {
    static constexpr bool kEmuThreads = false;
    struct Params {
        unsigned int* flags; int nslots; int slot[MAX_PIECES]; const unsigned int* epoch; unsigned int* err;
        unsigned long long timeout_ns;
    };
    SPIM_DEV static void run(const Params& p, int, float2*) {
#if defined(SPIM_HOST_EMU)
        const unsigned int want = *p.epoch;
        for (int i = 0; i < p.nslots; ++i) {
            struct timespec t0, t1;
            clock_gettime(CLOCK_MONOTONIC, &t0);
            while ((int)(__atomic_load_n(p.flags + p.slot[i], SPIM_EMU_FLAG_LOAD_ORDER) - want) < 0) {
                clock_gettime(CLOCK_MONOTONIC, &t1);
                const double ns = (double)(t1.tv_sec - t0.tv_sec) * 1e9 + (double)(t1.tv_nsec - t0.tv_nsec);
                if (ns > (double)p.timeout_ns) { __atomic_fetch_or(p.err, 1u, __ATOMIC_RELAXED); break; }
                struct timespec nap = {0, 50000};
                nanosleep(&nap, nullptr);
            }
        }
#else
        const int i = (int)threadIdx.x;
        if (i < p.nslots) {
            const unsigned int want = *p.epoch;
            unsigned long long t0, t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            for (;;) {
                unsigned int f;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(p.flags + p.slot[i]) : "memory");
                if ((int)(f - want) >= 0) break;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > p.timeout_ns) { atomicOr(p.err, 1u); break; }
                __nanosleep(200);
            }
        }
        __syncthreads();
        __threadfence_system();
#endif
    }
};

inline int ew_blocks(long long n) {
    long long b = (n + kChunk - 1) / kChunk;
    const long long cap = 148LL * 16;
    return (int)std::max<long long>(1, std::min(b, cap));
}

}  // namespace spim

#include "fusion.h"

// ================================================================================================
// error plumbing
// ================================================================================================
static thread_local std::string g_last_error;
static int fail(const std::string& msg) {
    g_last_error = msg;
    return 1;
}
#define SPIM_API_BEGIN try {
#define SPIM_API_END                                                                    \
    }                                                                                   \
    catch (const std::exception& e) { return fail(e.what()); }                          \
    catch (...) { return fail("unknown error"); }

// ================================================================================================
// host-side kernel (PSF) helpers: LRFFT.init / MVDeconFFT.init
// ================================================================================================
namespace {

struct HostVol {
    int d[3] = {0, 0, 0};
    std::vector<float> v;
    size_t size() const { return (size_t)d[0] * d[1] * d[2]; }
};

// RealSum (mpicbg.util.RealSum / net.imglib2.util.RealSum): compensated fp64 summation
double real_sum(const std::vector<float>& v) {
    double sum = 0.0, comp = 0.0;
    for (float f : v) {   // Neumaier variant
        const double x = (double)f;
        const double t = sum + x;
        if (fabs(sum) >= fabs(x)) comp += (sum - t) + x; else comp += (x - t) + sum;
        sum = t;
    }
    return sum + comp;
}

// AdjustInput.normImage D2/AdjustInput.java:53-59: t <- (float)((double)t / sum)
void norm_image(HostVol& k) {
    const double s = real_sum(k.v);
    for (float& f : k.v) f = (float)((double)f / s);
}

// computeInvertedKernel: Mirror.mirror along every axis (FD/Mirror.java:52-129).  Exact flip for
// odd sizes; for even sizes the reference swaps positions <= size/2 which un-swaps the middle pair.
HostVol invert_kernel(const HostVol& k) {
    HostVol o = k;
    for (int ax = 0; ax < 3; ++ax) {
        const int n = o.d[ax];
        std::vector<int> idx(n);
        for (int i = 0; i < n; ++i) idx[i] = i;
        if (n % 2 == 1) { for (int i = 0; i < n; ++i) idx[i] = n - 1 - i; }
        else { for (int pos = 0; pos <= n / 2; ++pos) std::swap(idx[pos], idx[n - 1 - pos]); }
        HostVol t = o;
        for (int z = 0; z < o.d[0]; ++z)
            for (int y = 0; y < o.d[1]; ++y)
                for (int x = 0; x < o.d[2]; ++x) {
                    int c[3] = {z, y, x};
                    int s[3] = {z, y, x};
                    s[ax] = idx[c[ax]];
                    t.v[((size_t)z * o.d[1] + y) * o.d[2] + x] = o.v[((size_t)s[0] * o.d[1] + s[1]) * o.d[2] + s[2]];
                }
        o = t;
    }
    return o;
}

// computeExponentialKernel + pow (LRFFT.java:361-391): repeated fp32 multiplication
HostVol exponential_kernel(const HostVol& k, int num_views) {
    HostVol o = k;
    for (size_t i = 0; i < o.v.size(); ++i) {
        volatile float r = k.v[i];
        for (int j = 1; j < num_views; ++j) r = r * k.v[i];
        o.v[i] = r;
    }
    return o;
}

struct DeviceGuard {
    explicit DeviceGuard(int dev) { rt::set_device(dev); }
};

// stand-alone convolution on host buffers through the engine (used for the PSF-sized kernel-building
// convolutions, by mvd_convolve and by the legacy JNA entry)
struct ConvWorkspace {
    ConvPlan plan;
    bool valid = false;
    int n[3] = {0, 0, 0}, k[3] = {0, 0, 0};
    bool exact = false;
    float* d_img = nullptr;
    float2* d_khat = nullptr;
    rt::Stream stream = 0;
    bool have_stream = false;

    void ensure(const int n_[3], const int k_[3], bool periodic_exact) {
        if (valid && exact == periodic_exact && !memcmp(n, n_, sizeof(n)) && !memcmp(k, k_, sizeof(k))) return;
        release();
        if (!have_stream) { stream = rt::stream_create(); have_stream = true; }
        plan.create(n_, k_, periodic_exact, true);
        memcpy(n, n_, sizeof(n)); memcpy(k, k_, sizeof(k));
        exact = periodic_exact;
        d_img = (float*)rt::dmalloc((size_t)n[0] * n[1] * n[2] * sizeof(float));
        d_khat = (float2*)rt::dmalloc(plan.spec_bytes());
        valid = true;
    }
    void release() {
        if (!valid) return;
        plan.destroy();
        rt::dfree(d_img); rt::dfree(d_khat);
        d_img = nullptr; d_khat = nullptr;
        valid = false;
    }
    // out may alias img
    void run(const float* img, const float* kernel, int ext, float value, float* out) {
        const size_t bytes = (size_t)n[0] * n[1] * n[2] * sizeof(float);
        rt::h2d(d_img, img, bytes, stream);
        plan.kernel_spectrum(kernel, d_khat, stream);
        SrcDesc src;
        src.p = d_img;
        for (int d = 0; d < 3; ++d) { src.dims[d] = n[d]; src.origin[d] = 0; }
        src.ext = ext; src.ext_value = value;
        EpiDesc e;
        e.epi = EPI_STORE; e.dst = d_img;
        for (int d = 0; d < 3; ++d) { e.dst_dims[d] = n[d]; e.dst_origin[d] = 0; }
        plan.convolve(src, d_khat, e, stream);
        rt::d2h(out, d_img, bytes, stream);
        rt::stream_sync(stream);
    }
    ~ConvWorkspace() {}
};

}  // namespace

// ================================================================================================
// session
// ================================================================================================
struct mvd_session {
    mvd_params prm;
    int n[3];
    int kmax[3] = {0, 0, 0};
    long long N = 0;
    std::vector<HostVol> psf, k1, k2;
    std::vector<double> k2sum;          // sum of each compound kernel (the response of conv2 to the constant 1)
    std::vector<float*> d_img, d_w;
    std::vector<float2*> d_kh1, d_kh2;
    float* d_psi = nullptr;    // dims pdims, origin porigin
    float* d_tmp = nullptr;    // ratio buffer, same geometry
    int pdims[3], porigin[3];
    int halo_lo_mask = 0, halo_hi_mask = 0;   // brick mode: sides whose halo is filled by a neighbour
    long long pelems = 0;
    ConvPlan plan;
    bool plan_ok = false, inited = false;
    rt::Stream stream = 0;
    rt::KernelTimer timer;
    double* d_stat_sum = nullptr;
    unsigned int* d_stat_max = nullptr;
    size_t stat_cap = 0;
    double avg = 0.0, osem = 1.0, avg_overlap = 0.0;
    int min_overlap = 0;
    long long dev_bytes = 0;
    double t_ms[8] = {0};
    long long t_cnt[8] = {0};
    std::vector<std::unique_ptr<ConvWorkspace>> small;   // PSF-sized plans for kernel building
    // fusion pre-step (include/spim_fusion.h)
    float* d_stack = nullptr;  // raw stack of the view being transformed
    size_t stack_cap = 0;
    int stack_dims[3] = {0, 0, 0};
    double* d_lut = nullptr;   // blending cosine table
    float* d_sumw = nullptr;   // virtual weights: sum image of WeightNormalizer.ComputeSumImage
    bool virtual_weights = false, wn_valid = false;
    int wn_min = 0;
    double wn_avg = 0.0;
    // brick mode, direct halo push over peer memory (mvd_p2p_*)
    struct P2P {
        unsigned int* d_ctl = nullptr;     // words [0,64) flags[buffer][slot], [64] block counter, [65,66] epochs, [67] error
        bool connected = false;
        int npieces = 0;
        float* peer_buf[2][MAX_PIECES];
        unsigned int* peer_ctl[MAX_PIECES];
        int lo[MAX_PIECES][3], dlo[MAX_PIECES][3], ext[MAX_PIECES][3];
        int slot_there[MAX_PIECES], slot_here[MAX_PIECES];
        std::vector<void*> opened;         // cudaIpcOpenMemHandle mappings to close
        long long pushes[2] = {0, 0}, waits[2] = {0, 0};
        // halo push fused into the x-inverse epilogue (HaloFuse, kernels.h): one descriptor per buffer on the device;
        // fused_valid[b]: buffer b was last written by a fused epilogue, i.e. the neighbours' halos already hold its faces
        // and the next push of b only has to raise the flags
        HaloFuse* d_fuse[2] = {nullptr, nullptr};
        bool fuse_ok = false;
        bool fused_valid[2] = {false, false};
    } p2p;

    int conv1_ext() const { return prm.conv1_ext >= 0 ? prm.conv1_ext : EXT_MIRROR_SINGLE; }
    int conv2_ext() const { return prm.conv2_ext >= 0 ? prm.conv2_ext : (prm.generation == 2 ? EXT_CONSTANT : EXT_MIRROR_SINGLE); }
    bool shift_active() const { return env_int("SPIM_CONST_SHIFT", 1) != 0 && conv2_ext() == EXT_CONSTANT; }

    size_t psi_bytes = 0, kh_bytes = 0;       // sizes the psi / ratio buffers and the kernel spectra were allocated with

    void* dalloc(size_t bytes) { dev_bytes += (long long)bytes; return rt::dmalloc(bytes); }
    void dfree_counted(void* p, size_t bytes) { if (p) { rt::dfree(p); dev_bytes -= (long long)bytes; } }

    HostVol small_conv(const HostVol& a, const HostVol& b) {   // zero-extended PSF-sized convolution
        ConvWorkspace* ws = nullptr;
        for (auto& w : small)
            if (!memcmp(w->n, a.d, sizeof(a.d)) && !memcmp(w->k, b.d, sizeof(b.d))) ws = w.get();
        if (!ws) {
            small.emplace_back(new ConvWorkspace());
            ws = small.back().get();
            ws->ensure(a.d, b.d, false);
        }
        HostVol o = a;
        ws->run(a.v.data(), b.v.data(), EXT_ZERO, 0.f, o.v.data());
        return o;
    }

    // LRInput.init -> LRFFT.init per view in list order (LRFFT.java:214-325, MVDeconFFT.java:183-323)
    void init_kernels() {
        const int V = prm.num_views;
        k1 = psf;
        k2.assign(V, HostVol());
        for (int v = 0; v < V; ++v) {
            norm_image(k1[v]);
            if (V == 1 || prm.iteration_type == MVD_INDEPENDENT) {
                k2[v] = invert_kernel(k1[v]);
            } else if (prm.iteration_type == MVD_EFFICIENT_BAYESIAN) {
                HostVol tmp = invert_kernel(k1[v]);
                for (int w = 0; w < V; ++w) {
                    if (w == v) continue;
                    HostVol c1 = small_conv(invert_kernel(k1[v]), k1[w]);
                    HostVol c2 = small_conv(c1, invert_kernel(k1[w]));
                    for (size_t i = 0; i < tmp.v.size(); ++i) { volatile float m = c2.v[i] * tmp.v[i]; tmp.v[i] = m; }
                }
                norm_image(tmp);
                k2[v] = tmp;
            } else if (prm.iteration_type == MVD_OPTIMIZATION_I) {
                HostVol tmp = k1[v];
                for (int w = 0; w < V; ++w) {
                    if (w == v) continue;
                    HostVol c = small_conv(k1[v], invert_kernel(k1[w]));
                    for (size_t i = 0; i < tmp.v.size(); ++i) { volatile float m = c.v[i] * tmp.v[i]; tmp.v[i] = m; }
                }
                norm_image(tmp);
                k2[v] = invert_kernel(tmp);
            } else {
                HostVol e = exponential_kernel(k1[v], V);
                norm_image(e);
                k2[v] = invert_kernel(e);
            }
        }
        for (auto& w : small) w->release();
        small.clear();
        k2sum.assign(V, 0.0);
        for (int v = 0; v < V; ++v) { double t = 0.0; for (float x : k2[v].v) t += (double)x; k2sum[v] = t; }
    }

    ViewPtrs view_ptrs() const {
        ViewPtrs vp;
        memset(&vp, 0, sizeof(vp));
        vp.nviews = prm.num_views;
        for (int v = 0; v < prm.num_views; ++v) { vp.img[v] = d_img[v]; vp.w[v] = d_w[v]; }
        return vp;
    }

    void partials(double out[6]) {
        double* d_sum = (double*)rt::dmalloc(sizeof(double));
        unsigned long long* d_cnt = (unsigned long long*)rt::dmalloc(3 * sizeof(unsigned long long));
        unsigned int* d_min = (unsigned int*)rt::dmalloc(sizeof(unsigned int));
        rt::dzero(d_sum, sizeof(double), stream);
        rt::dzero(d_cnt, 3 * sizeof(unsigned long long), stream);
        const unsigned int big = 0xffffffffu;
        rt::h2d(d_min, &big, sizeof(big), stream);
        rt::stream_sync(stream);
        InitStatsK::Params p;
        p.v = view_ptrs(); p.n = N; p.gen = prm.generation; p.nblocks = ew_blocks(N);
        p.sum = d_sum; p.cnt = d_cnt; p.min_overlap = d_min;
        rt::launch<InitStatsK>(p, p.nblocks, kThreads, 0, stream);
        double s = 0; unsigned long long c[3] = {0, 0, 0}; unsigned int mn = 0;
        rt::d2h(&s, d_sum, sizeof(s), stream);
        rt::d2h(c, d_cnt, sizeof(c), stream);
        rt::d2h(&mn, d_min, sizeof(mn), stream);
        rt::stream_sync(stream);
        rt::dfree(d_sum); rt::dfree(d_cnt); rt::dfree(d_min);
        out[0] = s; out[1] = (double)c[0]; out[2] = (double)c[1]; out[3] = (double)c[2];
        out[4] = (mn == 0xffffffffu) ? 2147483647.0 : (double)mn; out[5] = 0.0;
    }

    // avg / overlap statistics from (all-reduced) partial sums
    static void reduce_partials(int gen, const double p[6], double& avg, int& min_ov, double& avg_ov) {
        if (gen == 2) {
            avg = p[0] / p[1];            // NaN when no data at all (reference: falls back to 0.5)
            if (avg != avg) avg = 0.5;
            min_ov = 0; avg_ov = 0.0;
        } else {
            min_ov = (int)p[4];
            avg_ov = p[2] / p[3];
            avg = (p[1] == 0.0) ? 1.0 : (double)(float)(p[0] / p[1]);   // this.avg = (float)result[0]
        }
    }

    void apply_avg(double avg_, double osem_) {
        avg = avg_; osem = osem_;
        if (virtual_weights && d_sumw) {
            // NormalizingRandomAccess.get(): w <- (float)min(1, (w / sumWeights) * osem), materialised once
            WeightNormParams p;
            memset(&p, 0, sizeof(p));
            p.v.nviews = prm.num_views;
            for (int v = 0; v < prm.num_views; ++v) p.v.w[v] = d_w[v];
            p.n = N; p.mode = 2; p.sumw = d_sumw; p.osem = osem; p.nportions = 0; p.nblocks = ew_blocks(N);
            rt::launch<WeightNormK>(p, p.nblocks, kThreads, 16, stream);
            rt::stream_sync(stream);
            rt::dfree(d_sumw); d_sumw = nullptr; dev_bytes -= (long long)N * (long long)sizeof(float);
            virtual_weights = false;
        } else if (osem != 1.0) {
            for (int v = 0; v < prm.num_views; ++v) {
                if (!d_w[v]) {   // constant-1 weight: min(1, 1*osem) = 1 for osem >= 1; materialise otherwise
                    if (osem >= 1.0) continue;
                    d_w[v] = (float*)dalloc((size_t)N * sizeof(float));
                    FillK::Params f; f.p = d_w[v]; f.n = N; f.v = 1.f; f.nblocks = ew_blocks(N);
                    rt::launch<FillK>(f, f.nblocks, kThreads, 0, stream);
                }
                ScaleClampK::Params p; p.p = d_w[v]; p.n = N; p.f = (float)osem; p.nblocks = ew_blocks(N);
                rt::launch<ScaleClampK>(p, p.nblocks, kThreads, 0, stream);
            }
        }
        FillK::Params f; f.p = d_psi; f.n = pelems; f.v = (float)avg; f.nblocks = ew_blocks(pelems);
        rt::launch<FillK>(f, f.nblocks, kThreads, 0, stream);
        p2p.fused_valid[0] = false;
        rt::stream_sync(stream);
        inited = true;
    }

    void phase(int v, int ph, double* d_sum, unsigned int* d_max) {
        SrcDesc src;
        for (int d = 0; d < 3; ++d) { src.dims[d] = pdims[d]; src.origin[d] = porigin[d]; }
        src.halo_lo = halo_lo_mask; src.halo_hi = halo_hi_mask;
        EpiDesc e;
        for (int d = 0; d < 3; ++d) { e.dst_dims[d] = pdims[d]; e.dst_origin[d] = porigin[d]; }
        e.min_value = prm.min_value;
        // Constant extension of the quotient (gen-2 conv2, value c = 1) without transforming a single halo line: by linearity
        // conv(ext_c(r), K) = conv(ext_0(r - c), K) + c * sum(K), so conv1's epilogue stores r - c, conv2 runs with ZERO
        // extension -- the halo outside the volume joins the zero gap of the padded transform -- and the update epilogue adds
        // c * sum(K2) back.  Neighbour-provided halos (brick mode) carry r - c like the interior.  SPIM_CONST_SHIFT=0 keeps
        // the literal constant extension (A/B and parity runs).
        const bool shift = shift_active();
        const float cext = 1.f;
        if (ph == 0) {
            src.p = d_psi; src.ext = conv1_ext(); src.ext_value = 0.f;
            e.epi = EPI_RATIO; e.dst = d_tmp; e.img = d_img[v]; e.gen2_quotient = (prm.generation == 2); e.fast_epilogue = prm.fast_epilogue;
            if (shift) e.ratio_offset = -cext;
            if (p2p.connected && p2p.fuse_ok) e.fuse = p2p.d_fuse[1];
            p2p.fused_valid[1] = plan.convolve(src, d_kh1[v], e, stream);
        } else {
            src.p = d_tmp; src.ext = conv2_ext(); src.ext_value = cext;
            if (shift) { src.ext = EXT_ZERO; src.ext_value = 0.f; e.blur_offset = (float)((double)cext * k2sum[v]); }
            e.epi = EPI_UPDATE; e.dst = d_psi; e.weight = d_w[v]; e.const_weight = 1.f;
            e.lambda = prm.lambda; e.stat_sum = d_sum; e.stat_max = d_max; e.exact_tikhonov = prm.exact_tikhonov; e.fast_epilogue = prm.fast_epilogue;
            if (p2p.connected && p2p.fuse_ok) e.fuse = p2p.d_fuse[0];
            p2p.fused_valid[0] = plan.convolve(src, d_kh2[v], e, stream);
        }
    }

    void ensure_stats(size_t slots) {
        if (slots <= stat_cap) return;
        rt::dfree(d_stat_sum); rt::dfree(d_stat_max);
        d_stat_sum = (double*)rt::dmalloc(slots * sizeof(double));
        d_stat_max = (unsigned int*)rt::dmalloc(slots * sizeof(unsigned int));
        stat_cap = slots;
    }
};

static void p2p_disconnect(mvd_session* s) {
#if !defined(SPIM_HOST_EMU)
    for (void* q : s->p2p.opened) cudaIpcCloseMemHandle(q);
#endif
    s->p2p.opened.clear();
    s->p2p.connected = false;
    s->p2p.npieces = 0;
    s->p2p.fuse_ok = false;
    s->p2p.fused_valid[0] = s->p2p.fused_valid[1] = false;
    for (int b = 0; b < 2; ++b) { rt::dfree(s->p2p.d_fuse[b]); s->p2p.d_fuse[b] = nullptr; }
}

// ================================================================================================
// session C ABI
// ================================================================================================
extern "C" {

const char* mvd_last_error(void) { return g_last_error.c_str(); }
const char* mvd_version(void) {
#if defined(SPIM_HOST_EMU)
    return "spim_registration_b200 0.1 (HOST EMULATOR - tests only)";
#else
    return "spim_registration_b200 0.1 (sm_100a)";
#endif
}

void mvd_params_default(mvd_params* p) {
    memset(p, 0, sizeof(*p));
    p->struct_size = (int)sizeof(mvd_params);
    p->num_views = 1;
    p->iteration_type = MVD_EFFICIENT_BAYESIAN;
    p->generation = 2;
    p->lambda = 0.006;
    p->min_value = 0.0001f;
    p->osem_speedup = 1.0;
    p->osem_index = 0;
    p->conv1_ext = -1;
    p->conv2_ext = -1;
    p->device = 0;
    p->haloed = 0;
    p->fast_epilogue = 1;
}

int mvd_session_create(const mvd_params* p, mvd_session** out) {
    SPIM_API_BEGIN
    if (!p || !out) return fail("mvd_session_create: null argument");
    if (p->struct_size != (int)sizeof(mvd_params)) return fail("mvd_session_create: struct_size mismatch");
    if (p->num_views < 1 || p->num_views > MAX_VIEWS) return fail("mvd_session_create: num_views out of range");
    if (p->iteration_type < 0 || p->iteration_type > 3) return fail("mvd_session_create: bad iteration_type");
    if (p->generation != 1 && p->generation != 2) return fail("mvd_session_create: generation must be 1 or 2");
    for (int d = 0; d < 3; ++d) if (p->dims[d] < 1) return fail("mvd_session_create: bad dims");
    const int ndev = rt::device_count();
    if (ndev <= 0) return fail("mvd_session_create: no CUDA device available (this library has no CPU fallback)");
    if (p->device < 0 || p->device >= ndev) return fail("mvd_session_create: bad device ordinal");
    rt::set_device(p->device);
    std::unique_ptr<mvd_session> s(new mvd_session());
    s->prm = *p;
    for (int d = 0; d < 3; ++d) s->n[d] = p->dims[d];
    s->N = (long long)s->n[0] * s->n[1] * s->n[2];
    const int V = p->num_views;
    s->psf.assign(V, HostVol());
    s->d_img.assign(V, nullptr); s->d_w.assign(V, nullptr);
    s->d_kh1.assign(V, nullptr); s->d_kh2.assign(V, nullptr);
    s->stream = rt::stream_create();
    *out = s.release();
    return 0;
    SPIM_API_END
}

void mvd_session_destroy(mvd_session* s) {
    if (!s) return;
    try {
        rt::set_device(s->prm.device);
        rt::stream_sync(s->stream);
        for (auto p : s->d_img) rt::dfree(p);
        for (auto p : s->d_w) rt::dfree(p);
        for (auto p : s->d_kh1) rt::dfree(p);
        for (auto p : s->d_kh2) rt::dfree(p);
        rt::dfree(s->d_psi); rt::dfree(s->d_tmp);
        rt::dfree(s->d_stat_sum); rt::dfree(s->d_stat_max);
        rt::dfree(s->d_stack); rt::dfree(s->d_lut); rt::dfree(s->d_sumw);
        p2p_disconnect(s);
        rt::dfree(s->p2p.d_ctl);
        if (s->plan_ok) s->plan.destroy();
        rt::stream_destroy(s->stream);
    } catch (...) {}
    delete s;
}

int mvd_set_view(mvd_session* s, int v, const float* img, const float* weight, const float* psf, const int psf_dims[3]) {
    SPIM_API_BEGIN
    if (!s || !img || !psf || !psf_dims) return fail("mvd_set_view: null argument");
    if (v < 0 || v >= s->prm.num_views) return fail("mvd_set_view: view index out of range");
    for (int d = 0; d < 3; ++d) if (psf_dims[d] < 1) return fail("mvd_set_view: bad psf dims");
    rt::set_device(s->prm.device);
    const size_t bytes = (size_t)s->N * sizeof(float);
    if (!s->d_img[v]) s->d_img[v] = (float*)s->dalloc(bytes);
    rt::h2d(s->d_img[v], img, bytes, s->stream);
    if (weight) {
        if (!s->d_w[v]) s->d_w[v] = (float*)s->dalloc(bytes);
        rt::h2d(s->d_w[v], weight, bytes, s->stream);
    } else if (s->d_w[v]) {
        rt::dfree(s->d_w[v]); s->d_w[v] = nullptr;
    }
    HostVol& k = s->psf[v];
    for (int d = 0; d < 3; ++d) k.d[d] = psf_dims[d];
    k.v.assign(psf, psf + k.size());
    rt::stream_sync(s->stream);
    s->inited = false;
    return 0;
    SPIM_API_END
}

int mvd_upload_region(mvd_session* s, int view, int which, const float* data, const int lo[3], const int ext[3]) {
    SPIM_API_BEGIN
    if (!s || !data || !lo || !ext) return fail("mvd_upload_region: null argument");
    if (view < 0 || view >= s->prm.num_views) return fail("mvd_upload_region: view index out of range");
    if (which != 0 && which != 1) return fail("mvd_upload_region: which must be 0 (image) or 1 (weight)");
    for (int d = 0; d < 3; ++d)
        if (lo[d] < 0 || ext[d] < 1 || (long long)lo[d] + ext[d] > s->n[d]) return fail("mvd_upload_region: region out of range");
    rt::set_device(s->prm.device);
    float*& dst = which == 0 ? s->d_img[view] : s->d_w[view];
    if (!dst) {
        // a freshly created buffer starts as "no data" (image 0) / "no contribution" (weight 0)
        dst = (float*)s->dalloc((size_t)s->N * sizeof(float));
        rt::dzero(dst, (size_t)s->N * sizeof(float), s->stream);
    }
    rt::h2d_box(dst, s->n, lo, data, ext, s->stream);
    rt::stream_sync(s->stream);
    s->inited = false;
    return 0;
    SPIM_API_END
}

static int session_prepare(mvd_session* s) {
    const int V = s->prm.num_views;
    for (int v = 0; v < V; ++v) if (!s->d_img[v] || s->psf[v].v.empty()) return fail("mvd_init: view " + std::to_string(v) + " not set");
    for (int d = 0; d < 3; ++d) {
        s->kmax[d] = 0;
        for (int v = 0; v < V; ++v) s->kmax[d] = std::max(s->kmax[d], s->psf[v].d[d]);
    }
    if (s->plan_ok) { s->dev_bytes -= (long long)s->plan.spec_bytes(); s->plan.destroy(); s->plan_ok = false; }
    s->plan.create(s->n, s->kmax, false, true);
    s->plan.timer = &s->timer;
    s->plan_ok = true;
    s->dev_bytes += (long long)s->plan.spec_bytes();
    for (int d = 0; d < 3; ++d) {
        if (s->prm.haloed) {
            s->porigin[d] = s->plan.hm[d];
            s->pdims[d] = s->porigin[d] + s->n[d] + s->plan.hp[d];
            if (d == 2) {
                // x: keep the brick's first voxel and the row length 8-byte aligned so that the x passes run their
                // vectorised paths in brick mode too (an odd PSF/2 halo would shift every row by one float)
                s->porigin[d] += s->porigin[d] & 1;
                s->pdims[d] = s->porigin[d] + s->n[d] + s->plan.hp[d];
                s->pdims[d] += s->pdims[d] & 1;
            }
        } else { s->pdims[d] = s->n[d]; s->porigin[d] = 0; }
    }
    s->pelems = (long long)s->pdims[0] * s->pdims[1] * s->pdims[2];
    // a second mvd_init after mvd_set_view changed a PSF re-plans: the padded FFT size, and with it the size of every
    // spectrum and (brick mode) of the haloed psi / ratio buffers, may have changed -- buffers are re-created whenever
    // their size did (a larger PSF would otherwise write past the allocations of the first plan)
    const size_t pbytes = (size_t)s->pelems * sizeof(float);
    if (pbytes != s->psi_bytes) {
        if (s->p2p.connected) return fail("mvd_init: the haloed buffers change size while peers are connected (mvd_p2p_disconnect first)");
        s->dfree_counted(s->d_psi, s->psi_bytes); s->dfree_counted(s->d_tmp, s->psi_bytes);
        s->d_psi = (float*)s->dalloc(pbytes);
        s->d_tmp = (float*)s->dalloc(pbytes);
        s->psi_bytes = pbytes;
    }
    s->init_kernels();
    const size_t sbytes = s->plan.spec_bytes();
    if (sbytes != s->kh_bytes) {
        for (int v = 0; v < V; ++v) {
            s->dfree_counted(s->d_kh1[v], s->kh_bytes); s->d_kh1[v] = nullptr;
            s->dfree_counted(s->d_kh2[v], s->kh_bytes); s->d_kh2[v] = nullptr;
        }
        s->kh_bytes = sbytes;
    }
    for (int v = 0; v < V; ++v) {
        if (!s->d_kh1[v]) s->d_kh1[v] = (float2*)s->dalloc(sbytes);
        if (!s->d_kh2[v]) s->d_kh2[v] = (float2*)s->dalloc(sbytes);
        // kernels smaller than kmax are spectrum-transformed with their own dims
        ConvPlan& pl = s->plan;
        int save[3] = {pl.k[0], pl.k[1], pl.k[2]};
        for (int d = 0; d < 3; ++d) pl.k[d] = s->k1[v].d[d];
        pl.kernel_spectrum(s->k1[v].v.data(), s->d_kh1[v], s->stream);
        pl.kernel_spectrum(s->k2[v].v.data(), s->d_kh2[v], s->stream);
        for (int d = 0; d < 3; ++d) pl.k[d] = save[d];
    }
    rt::stream_sync(s->stream);
    return 0;
}

int mvd_init(mvd_session* s) {
    SPIM_API_BEGIN
    if (!s) return fail("mvd_init: null session");
    rt::set_device(s->prm.device);
    if (int rc = session_prepare(s)) return rc;
    if (s->prm.haloed) return 0;   // brick mode: caller all-reduces mvd_init_partials and calls mvd_set_avg
    double part[6];
    s->partials(part);
    double avg; int mn; double av;
    mvd_session::reduce_partials(s->prm.generation, part, avg, mn, av);
    s->min_overlap = mn; s->avg_overlap = av;
    double osem = s->prm.osem_speedup;
    if (s->prm.generation == 1) {
        if (s->prm.osem_index == 1) osem = std::max(1.0, (double)mn);
        else if (s->prm.osem_index == 2) osem = std::max(1.0, av);
    } else if (s->wn_valid) {
        // ProcessForDeconvolution.java:346-358: overlap statistics of the WeightNormalizer, each max(1, .)
        s->min_overlap = s->wn_min; s->avg_overlap = s->wn_avg;
        if (s->prm.osem_index == 1) osem = (double)std::max(1, s->wn_min);
        else if (s->prm.osem_index == 2) osem = std::max(1.0, s->wn_avg);
    }
    s->apply_avg(avg, osem);
    return 0;
    SPIM_API_END
}

int mvd_init_partials(mvd_session* s, double partial[6]) {
    SPIM_API_BEGIN
    if (!s || !partial) return fail("mvd_init_partials: null argument");
    if (!s->plan_ok) return fail("mvd_init_partials: call mvd_init first");
    rt::set_device(s->prm.device);
    s->partials(partial);
    return 0;
    SPIM_API_END
}

int mvd_set_avg(mvd_session* s, double avg, double osem) {
    SPIM_API_BEGIN
    if (!s) return fail("mvd_set_avg: null session");
    if (!s->plan_ok) return fail("mvd_set_avg: call mvd_init first");
    rt::set_device(s->prm.device);
    s->apply_avg(avg, osem);
    return 0;
    SPIM_API_END
}

int mvd_run(mvd_session* s, int n_iterations, double* sum_change, double* max_change) {
    SPIM_API_BEGIN
    if (!s) return fail("mvd_run: null session");
    if (!s->inited) return fail("mvd_run: session not initialised (mvd_init)");
    if (s->prm.haloed) return fail("mvd_run: brick-mode sessions are driven with mvd_view_phase");
    if (n_iterations < 0) return fail("mvd_run: negative iteration count");
    rt::set_device(s->prm.device);
    const int V = s->prm.num_views;
    const size_t slots = (size_t)n_iterations * V;
    if (slots == 0) return 0;
    s->ensure_stats(slots);
    rt::dzero(s->d_stat_sum, slots * sizeof(double), s->stream);
    rt::dzero(s->d_stat_max, slots * sizeof(unsigned int), s->stream);
    for (int it = 0; it < n_iterations; ++it)
        for (int v = 0; v < V; ++v) {
            const size_t slot = (size_t)it * V + v;
            s->phase(v, 0, nullptr, nullptr);
            s->phase(v, 1, s->d_stat_sum + slot, s->d_stat_max + slot);
        }
    if (sum_change || max_change) {
        std::vector<double> hs(slots);
        std::vector<unsigned int> hm(slots);
        rt::d2h(hs.data(), s->d_stat_sum, slots * sizeof(double), s->stream);
        rt::d2h(hm.data(), s->d_stat_max, slots * sizeof(unsigned int), s->stream);
        rt::stream_sync(s->stream);
        for (size_t i = 0; i < slots; ++i) {
            if (sum_change) sum_change[i] = hs[i];
            if (max_change) { float f; memcpy(&f, &hm[i], 4); max_change[i] = (double)f; }
        }
    } else {
        rt::stream_sync(s->stream);
    }
    s->timer.collect(s->t_ms, s->t_cnt, 8);
    return 0;
    SPIM_API_END
}

int mvd_view_phase(mvd_session* s, int view, int phase, double* stats) {
    SPIM_API_BEGIN
    if (!s) return fail("mvd_view_phase: null session");
    if (!s->inited) return fail("mvd_view_phase: session not initialised");
    if (view < 0 || view >= s->prm.num_views || (phase != 0 && phase != 1)) return fail("mvd_view_phase: bad view/phase");
    rt::set_device(s->prm.device);
    if (phase == 1) {
        s->ensure_stats(1);
        rt::dzero(s->d_stat_sum, sizeof(double), s->stream);
        rt::dzero(s->d_stat_max, sizeof(unsigned int), s->stream);
        s->phase(view, 1, s->d_stat_sum, s->d_stat_max);
        if (stats) {
            double hs; unsigned int hm;
            rt::d2h(&hs, s->d_stat_sum, sizeof(double), s->stream);
            rt::d2h(&hm, s->d_stat_max, sizeof(unsigned int), s->stream);
            rt::stream_sync(s->stream);
            float f; memcpy(&f, &hm, 4);
            stats[0] = hs; stats[1] = (double)f;
        }
    } else {
        s->phase(view, 0, nullptr, nullptr);
    }
    return 0;
    SPIM_API_END
}

int mvd_finish(mvd_session* s) {
    SPIM_API_BEGIN
    if (!s) return fail("mvd_finish: null session");
    if (!s->inited) return fail("mvd_finish: session not initialised");
    rt::set_device(s->prm.device);
    if (s->prm.generation == 2) {
        MaskK::Params p;
        p.v = s->view_ptrs(); p.psi = s->d_psi;
        for (int d = 0; d < 3; ++d) { p.pdims[d] = s->pdims[d]; p.porigin[d] = s->porigin[d]; p.n[d] = s->n[d]; }
        p.nblocks = ew_blocks(s->N);
        rt::launch<MaskK>(p, p.nblocks, kThreads, 0, s->stream);
        s->p2p.fused_valid[0] = false;       // psi changed without the neighbours' halos following
    }
    rt::stream_sync(s->stream);
    return 0;
    SPIM_API_END
}

static void copy_region(mvd_session* s, const float* src, const int sd[3], const int so[3], float* dst, const int dd[3], const int dof[3]) {
    CopyRegionK::Params p;
    p.src = src; p.dst = dst;
    for (int d = 0; d < 3; ++d) { p.sdims[d] = sd[d]; p.sorigin[d] = so[d]; p.ddims[d] = dd[d]; p.dorigin[d] = dof[d]; p.n[d] = s->n[d]; }
    p.nblocks = ew_blocks(s->N);
    rt::launch<CopyRegionK>(p, p.nblocks, kThreads, 0, s->stream);
}

int mvd_get_psi(mvd_session* s, float* out) {
    SPIM_API_BEGIN
    if (!s || !out) return fail("mvd_get_psi: null argument");
    if (!s->d_psi) return fail("mvd_get_psi: session not initialised");
    rt::set_device(s->prm.device);
    const size_t bytes = (size_t)s->N * sizeof(float);
    if (!s->prm.haloed) {
        rt::d2h(out, s->d_psi, bytes, s->stream);
        rt::stream_sync(s->stream);
    } else {
        float* t = (float*)rt::dmalloc(bytes);
        const int zero[3] = {0, 0, 0};
        copy_region(s, s->d_psi, s->pdims, s->porigin, t, s->n, zero);
        rt::d2h(out, t, bytes, s->stream);
        rt::stream_sync(s->stream);
        rt::dfree(t);
    }
    return 0;
    SPIM_API_END
}

int mvd_set_psi(mvd_session* s, const float* in) {
    SPIM_API_BEGIN
    if (!s || !in) return fail("mvd_set_psi: null argument");
    if (!s->d_psi) return fail("mvd_set_psi: session not initialised");
    rt::set_device(s->prm.device);
    s->p2p.fused_valid[0] = false;
    const size_t bytes = (size_t)s->N * sizeof(float);
    if (!s->prm.haloed) {
        rt::h2d(s->d_psi, in, bytes, s->stream);
        rt::stream_sync(s->stream);
    } else {
        float* t = (float*)rt::dmalloc(bytes);
        rt::h2d(t, in, bytes, s->stream);
        const int zero[3] = {0, 0, 0};
        copy_region(s, t, s->n, zero, s->d_psi, s->pdims, s->porigin);
        rt::stream_sync(s->stream);
        rt::dfree(t);
    }
    return 0;
    SPIM_API_END
}

int mvd_get_kernel(mvd_session* s, int view, int which, float* out) {
    SPIM_API_BEGIN
    if (!s || !out) return fail("mvd_get_kernel: null argument");
    if (view < 0 || view >= s->prm.num_views || (which != 1 && which != 2)) return fail("mvd_get_kernel: bad view/which");
    if (s->k1.empty()) return fail("mvd_get_kernel: call mvd_init first");
    const HostVol& k = which == 1 ? s->k1[view] : s->k2[view];
    memcpy(out, k.v.data(), k.v.size() * sizeof(float));
    return 0;
    SPIM_API_END
}

int mvd_get_info(mvd_session* s, mvd_info* o) {
    SPIM_API_BEGIN
    if (!s || !o) return fail("mvd_get_info: null argument");
    memset(o, 0, sizeof(*o));
    o->avg = s->avg; o->osem = s->osem; o->min_overlap = s->min_overlap; o->avg_overlap = s->avg_overlap;
    o->n_voxels = s->N;
    o->device_bytes = s->dev_bytes;
    if (s->plan_ok) {
        for (int d = 0; d < 3; ++d) { o->fft_dims[d] = s->plan.P[d]; o->halo_lo[d] = s->plan.hm[d]; o->halo_hi[d] = s->plan.hp[d]; }
        o->pitch = s->plan.pitch;
        o->np_voxels = s->plan.padded_min_voxels();
    }
    return 0;
    SPIM_API_END
}

int mvd_sync(mvd_session* s) {
    SPIM_API_BEGIN
    if (!s) return fail("mvd_sync: null session");
    rt::set_device(s->prm.device);
    rt::stream_sync(s->stream);
    return 0;
    SPIM_API_END
}

int mvd_get_stream(mvd_session* s, void** stream) {
    SPIM_API_BEGIN
    if (!s || !stream) return fail("mvd_get_stream: null argument");
    *stream = (void*)(uintptr_t)s->stream;
    return 0;
    SPIM_API_END
}

int mvd_set_timing(mvd_session* s, int on) {
    SPIM_API_BEGIN
    if (!s) return fail("mvd_set_timing: null session");
    rt::set_device(s->prm.device);
    rt::stream_sync(s->stream);
    s->timer.reset();
    s->timer.enable(on != 0);
    for (int i = 0; i < 8; ++i) { s->t_ms[i] = 0; s->t_cnt[i] = 0; }
    return 0;
    SPIM_API_END
}

int mvd_get_timing(mvd_session* s, double ms[8], long long launches[8]) {
    SPIM_API_BEGIN
    if (!s || !ms || !launches) return fail("mvd_get_timing: null argument");
    rt::set_device(s->prm.device);
    rt::stream_sync(s->stream);
    s->timer.collect(s->t_ms, s->t_cnt, 8);
    for (int i = 0; i < 8; ++i) { ms[i] = s->t_ms[i]; launches[i] = s->t_cnt[i]; }
    return 0;
    SPIM_API_END
}

int mvd_get_device_buffer(mvd_session* s, int which, void** dptr, int dims[3], int origin[3]) {
    SPIM_API_BEGIN
    if (!s || !dptr || !dims || !origin) return fail("mvd_get_device_buffer: null argument");
    if (!s->d_psi) return fail("mvd_get_device_buffer: call mvd_init first");
    *dptr = which == 0 ? (void*)s->d_psi : (void*)s->d_tmp;
    for (int d = 0; d < 3; ++d) { dims[d] = s->pdims[d]; origin[d] = s->porigin[d]; }
    return 0;
    SPIM_API_END
}

int mvd_set_halo_mask(mvd_session* s, int lo_mask, int hi_mask) {
    SPIM_API_BEGIN
    if (!s) return fail("mvd_set_halo_mask: null session");
    if (!s->prm.haloed) return fail("mvd_set_halo_mask: not a brick-mode session");
    s->halo_lo_mask = lo_mask & 7;
    s->halo_hi_mask = hi_mask & 7;
    return 0;
    SPIM_API_END
}

static int halo_pack_common(mvd_session* s, int which, int npieces, const int* regions, void* flat, int unpack) {
    if (!s || !regions || !flat) return fail("mvd_halo_pack: null argument");
    if (!s->prm.haloed || !s->d_psi) return fail("mvd_halo_pack: not a brick-mode session / not initialised");
    if (npieces < 0 || npieces > MAX_PIECES) return fail("mvd_halo_pack: too many pieces");
    if (npieces == 0) return 0;
    rt::set_device(s->prm.device);
    HaloPackK::Params p;
    memset(&p, 0, sizeof(p));
    p.buf = which == 0 ? s->d_psi : s->d_tmp;
    p.flat = (float*)flat;
    p.npieces = npieces; p.unpack = unpack;
    for (int d = 0; d < 3; ++d) p.dims[d] = s->pdims[d];
    long long off = 0;
    for (int i = 0; i < npieces; ++i) {
        p.off[i] = off;
        long long n = 1;
        for (int d = 0; d < 3; ++d) {
            p.lo[i][d] = regions[i * 6 + d];
            p.ext[i][d] = regions[i * 6 + 3 + d];
            if (p.lo[i][d] < 0 || p.ext[i][d] < 1 || p.lo[i][d] + p.ext[i][d] > s->pdims[d]) return fail("mvd_halo_pack: region out of range");
            n *= p.ext[i][d];
        }
        off += n;
    }
    p.off[npieces] = off;
    p.nblocks = ew_blocks(off);
    rt::launch<HaloPackK>(p, p.nblocks, kThreads, 0, s->stream);
    return 0;
}

int mvd_halo_pack(mvd_session* s, int which, int npieces, const int* regions, void* flat) {
    SPIM_API_BEGIN
    return halo_pack_common(s, which, npieces, regions, flat, 0);
    SPIM_API_END
}

int mvd_halo_unpack(mvd_session* s, int which, int npieces, const int* regions, void* flat) {
    SPIM_API_BEGIN
    return halo_pack_common(s, which, npieces, regions, flat, 1);
    SPIM_API_END
}

int mvd_fill_halo(mvd_session* s, int which, int lo_mask, int hi_mask) {
    SPIM_API_BEGIN
    if (!s) return fail("mvd_fill_halo: null session");
    if (!s->prm.haloed || !s->d_psi) return fail("mvd_fill_halo: not a brick-mode session / not initialised");
    rt::set_device(s->prm.device);
    const int ext = which == 0 ? s->conv1_ext() : s->conv2_ext();
    // the ratio buffer holds r - 1 when the constant extension is realised by shift (mvd_session::phase): its constant is 0
    const float value = which == 0 ? 0.f : (s->shift_active() ? 0.f : 1.f);
    for (int axis = 2; axis >= 0; --axis) {
        for (int side = 0; side < 2; ++side) {
            const int mask = side == 0 ? lo_mask : hi_mask;
            if (!(mask & (1 << axis))) continue;
            const int width = side == 0 ? s->plan.hm[axis] : s->plan.hp[axis];
            if (width <= 0) continue;
            HaloFillK::Params p;
            p.buf = which == 0 ? s->d_psi : s->d_tmp;
            for (int d = 0; d < 3; ++d) { p.dims[d] = s->pdims[d]; p.origin[d] = s->porigin[d]; p.n[d] = s->n[d]; }
            p.axis = axis; p.side = side; p.width = width; p.ext = ext; p.value = value;
            long long total = (long long)width;
            for (int d = 0; d < 3; ++d) if (d != axis) total *= s->pdims[d];
            p.nblocks = ew_blocks(total);
            rt::launch<HaloFillK>(p, p.nblocks, kThreads, 0, s->stream);
        }
    }
    return 0;
    SPIM_API_END
}

// ---- direct halo push over peer memory (NVLink / CUDA IPC) -------------------------------------
struct P2PBlob {               // 96 bytes: how another rank reaches one of my device buffers
    char magic[8];             // "SPIMP2P1"
    int pid, device;
    unsigned long long ptr;    // valid inside the exporting process
    unsigned char ipc[64];     // cudaIpcMemHandle_t, valid in every other process of the node
    unsigned char pad[8];
};
static_assert(sizeof(P2PBlob) == 96, "P2PBlob layout");
constexpr size_t kCtlBytes = 2u << 20;   // a whole 2 MiB page: an IPC mapping never exposes anything else

static void p2p_fill_blob(P2PBlob& b, int device, void* ptr) {
    memset(&b, 0, sizeof(b));
    memcpy(b.magic, "SPIMP2P1", 8);
    b.pid = (int)getpid();
    b.device = device;
    b.ptr = (unsigned long long)(uintptr_t)ptr;
#if !defined(SPIM_HOST_EMU)
    cudaIpcMemHandle_t h;
    SPIM_CUDA_CHECK(cudaIpcGetMemHandle(&h, ptr));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t size");
    memcpy(b.ipc, &h, 64);
#endif
}

static void* p2p_resolve(mvd_session* s, const P2PBlob& b) {
    if (memcmp(b.magic, "SPIMP2P1", 8) != 0) throw rt::Error("mvd_p2p_connect: bad peer record");
    if (b.pid == (int)getpid()) {
#if !defined(SPIM_HOST_EMU)
        if (b.device != s->prm.device) {
            int can = 0;
            SPIM_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, s->prm.device, b.device));
            if (!can) throw rt::Error("mvd_p2p_connect: no peer access between devices " + std::to_string(s->prm.device) + " and " + std::to_string(b.device));
            const cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) throw rt::Error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            cudaGetLastError();
        }
#endif
        return (void*)(uintptr_t)b.ptr;
    }
#if defined(SPIM_HOST_EMU)
    (void)s;
    throw rt::Error("mvd_p2p_connect: the peer lives in another process (the emulator has no shared device memory)");
#else
    cudaIpcMemHandle_t h;
    memcpy(&h, b.ipc, 64);
    void* q = nullptr;
    SPIM_CUDA_CHECK(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
    s->p2p.opened.push_back(q);
    return q;
#endif
}

int mvd_p2p_export(mvd_session* s, unsigned char record[MVD_P2P_RECORD_BYTES]) {
    SPIM_API_BEGIN
    if (!s || !record) return fail("mvd_p2p_export: null argument");
    if (!s->prm.haloed || !s->d_psi || !s->d_tmp) return fail("mvd_p2p_export: not a brick-mode session / not initialised (mvd_init)");
    rt::set_device(s->prm.device);
    if (!s->p2p.d_ctl) {
        s->p2p.d_ctl = (unsigned int*)rt::dmalloc(kCtlBytes);
        rt::dzero(s->p2p.d_ctl, kCtlBytes, s->stream);
        rt::stream_sync(s->stream);
    }
    P2PBlob b[3];
    p2p_fill_blob(b[0], s->prm.device, s->d_psi);
    p2p_fill_blob(b[1], s->prm.device, s->d_tmp);
    p2p_fill_blob(b[2], s->prm.device, s->p2p.d_ctl);
    static_assert(sizeof(b) == MVD_P2P_RECORD_BYTES, "record size");
    memcpy(record, b, sizeof(b));
    return 0;
    SPIM_API_END
}

int mvd_p2p_connect(mvd_session* s, int npieces, const unsigned char* records, const int* boxes, const int* slots) {
    SPIM_API_BEGIN
    if (!s || !records || !boxes || !slots) return fail("mvd_p2p_connect: null argument");
    if (!s->p2p.d_ctl) return fail("mvd_p2p_connect: call mvd_p2p_export first");
    if (npieces < 1 || npieces > MAX_PIECES) return fail("mvd_p2p_connect: piece count out of range");
    rt::set_device(s->prm.device);
    p2p_disconnect(s);
    mvd_session::P2P& q = s->p2p;
    try {
        for (int i = 0; i < npieces; ++i) {
            P2PBlob b[3];
            memcpy(b, records + (size_t)i * MVD_P2P_RECORD_BYTES, sizeof(b));
            q.peer_buf[0][i] = (float*)p2p_resolve(s, b[0]);
            q.peer_buf[1][i] = (float*)p2p_resolve(s, b[1]);
            q.peer_ctl[i] = (unsigned int*)p2p_resolve(s, b[2]);
            for (int d = 0; d < 3; ++d) {
                q.lo[i][d] = boxes[i * 9 + d]; q.ext[i][d] = boxes[i * 9 + 3 + d]; q.dlo[i][d] = boxes[i * 9 + 6 + d];
                if (q.ext[i][d] < 1 || q.lo[i][d] < 0 || q.dlo[i][d] < 0 || q.lo[i][d] + q.ext[i][d] > s->pdims[d] ||
                    q.dlo[i][d] + q.ext[i][d] > s->pdims[d])
                    throw rt::Error("mvd_p2p_connect: box out of range");
            }
            q.slot_there[i] = slots[i * 2]; q.slot_here[i] = slots[i * 2 + 1];
            if ((unsigned)q.slot_there[i] >= 27u || (unsigned)q.slot_here[i] >= 27u) throw rt::Error("mvd_p2p_connect: flag slot out of range");
        }
    } catch (...) {
        p2p_disconnect(s);
        throw;
    }
    q.npieces = npieces;
    q.connected = true;
    q.pushes[0] = q.pushes[1] = q.waits[0] = q.waits[1] = 0;
    // Descriptors of the fused push: the x-inverse epilogue stores halo voxels straight into the neighbours' buffers and the
    // push only raises the flags.  Opt-in (SPIM_BRICK_FUSE=1): correct on hardware (2 B200: oracle parity and the full-size
    // check pass), but measured SLOWER there -- 13.78 vs 13.30 ms per iteration, the fused epilogue costs the x-inverse
    // kernels 6 % while one z-face exchange is only 30 us; it can pay only where the exchange is larger (8 bricks: 86 us).  The direction of piece i follows from where its box lands in the
    // neighbour's buffer: beyond the brick = the neighbour is below me on that axis, before it = above.
    q.fuse_ok = false;
    if (env_int("SPIM_BRICK_FUSE", 0) != 0) {
        HaloFuse h[2];
        memset(h, 0, sizeof(h));
        bool ok = (s->porigin[2] % 2 == 0) && (s->pdims[2] % 2 == 0) && (s->n[2] % 2 == 0);
        int seen_lo = 0, seen_hi = 0;
        for (int b = 0; b < 2; ++b)
            for (int d = 0; d < 3; ++d) { h[b].n[d] = s->n[d]; h[b].wlo[d] = s->plan.hm[d]; h[b].whi[d] = s->plan.hp[d]; }
        for (int i = 0; i < npieces && ok; ++i) {
            int off[3];
            for (int d = 0; d < 3; ++d)
                off[d] = q.dlo[i][d] >= s->porigin[d] + s->n[d] ? -1 : (q.dlo[i][d] < s->porigin[d] ? 1 : 0);
            const int dir = (off[0] + 1) * 9 + (off[1] + 1) * 3 + (off[2] + 1);
            if (dir == 13) { ok = false; break; }
            const long long shift = ((long long)off[0] * s->n[0] * s->pdims[1] + (long long)off[1] * s->n[1]) * s->pdims[2] + (long long)off[2] * s->n[2];
            for (int b = 0; b < 2; ++b) { h[b].peer[dir] = q.peer_buf[b][i]; h[b].shift[dir] = shift; }
            for (int d = 0; d < 3; ++d) { if (off[d] < 0) seen_lo |= 1 << d; if (off[d] > 0) seen_hi |= 1 << d; }
        }
        // every combination of the axes that have a neighbour must be mapped (a regular grid of bricks provides them)
        for (int dz = -1; dz <= 1 && ok; ++dz)
            for (int dy = -1; dy <= 1 && ok; ++dy)
                for (int dx = -1; dx <= 1 && ok; ++dx) {
                    if (!dz && !dy && !dx) continue;
                    const int o[3] = {dz, dy, dx};
                    bool need = true;
                    for (int d = 0; d < 3; ++d) if ((o[d] < 0 && !(seen_lo & (1 << d))) || (o[d] > 0 && !(seen_hi & (1 << d)))) need = false;
                    if (need && !h[0].peer[(dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)]) ok = false;
                }
        // the x pieces travel as pairs (2n, 2n+1): the widened ranges must exist in the neighbours' rows
        const int lo2 = 2 * ((s->plan.hp[2] + 1) >> 1), hi2 = s->n[2] - 2 * ((s->n[2] - s->plan.hm[2]) >> 1);
        if ((seen_lo & 4) && s->porigin[2] + s->n[2] + lo2 > s->pdims[2]) ok = false;
        if ((seen_hi & 4) && s->porigin[2] < hi2) ok = false;
        for (int d = 0; d < 3; ++d) if (s->n[d] < s->plan.hm[d] + s->plan.hp[d]) ok = false;     // a line is in at most one face region per axis
        if (ok && (seen_lo | seen_hi)) {
            for (int b = 0; b < 2; ++b) {
                h[b].has_lo = seen_lo; h[b].has_hi = seen_hi;
                q.d_fuse[b] = (HaloFuse*)rt::dmalloc(sizeof(HaloFuse));
                rt::h2d(q.d_fuse[b], &h[b], sizeof(HaloFuse), s->stream);
            }
            rt::stream_sync(s->stream);
            q.fuse_ok = true;
        }
    }
    q.fused_valid[0] = q.fused_valid[1] = false;
    return 0;
    SPIM_API_END
}

int mvd_p2p_push(mvd_session* s, int which) {
    SPIM_API_BEGIN
    if (!s || (which != 0 && which != 1)) return fail("mvd_p2p_push: bad argument");
    if (!s->p2p.connected) return fail("mvd_p2p_push: not connected (mvd_p2p_connect)");
    if (s->p2p.pushes[which] != s->p2p.waits[which]) return fail("mvd_p2p_push: the previous push of this buffer has not been waited for");
    rt::set_device(s->prm.device);
    const mvd_session::P2P& q = s->p2p;
    HaloPushK::Params p;
    memset(&p, 0, sizeof(p));
    p.buf = which == 0 ? s->d_psi : s->d_tmp;
    for (int d = 0; d < 3; ++d) p.dims[d] = s->pdims[d];
    p.npieces = q.npieces;
    long long off = 0;
    for (int i = 0; i < q.npieces; ++i) {
        p.dst[i] = q.peer_buf[which][i];
        p.flag[i] = q.peer_ctl[i] + which * 32 + q.slot_there[i];
        p.off[i] = off;
        long long n = 1;
        for (int d = 0; d < 3; ++d) { p.lo[i][d] = q.lo[i][d]; p.dlo[i][d] = q.dlo[i][d]; p.ext[i][d] = q.ext[i][d]; n *= q.ext[i][d]; }
        const bool vec = (q.ext[i][2] % 4 == 0) && (q.lo[i][2] % 4 == 0) && (q.dlo[i][2] % 4 == 0) && (s->pdims[2] % 4 == 0) &&
                         ((reinterpret_cast<uintptr_t>(p.buf) | reinterpret_cast<uintptr_t>(p.dst[i])) & 15) == 0;
        p.vsh[i] = vec ? 2 : 0;
        off += vec ? n / 4 : n;
    }
    if (q.fused_valid[which]) {
        // the x-inverse epilogue that produced this buffer has already stored every piece at the neighbours: only the flags
        for (int i = 0; i <= q.npieces; ++i) p.off[i] = 0;
        off = 0;
        s->p2p.fused_valid[which] = false;
    }
    p.off[q.npieces] = off;
    p.nblocks = off > 0 ? ew_blocks(off) : 1;
    p.done = s->p2p.d_ctl + 64;
    p.epoch = s->p2p.d_ctl + 65 + which;
    rt::launch<HaloPushK>(p, p.nblocks, kThreads, 0, s->stream);
    s->p2p.pushes[which] += 1;
    return 0;
    SPIM_API_END
}

int mvd_p2p_wait(mvd_session* s, int which) {
    SPIM_API_BEGIN
    if (!s || (which != 0 && which != 1)) return fail("mvd_p2p_wait: bad argument");
    if (!s->p2p.connected) return fail("mvd_p2p_wait: not connected (mvd_p2p_connect)");
    if (s->p2p.pushes[which] != s->p2p.waits[which] + 1) return fail("mvd_p2p_wait: no outstanding push of this buffer");
    rt::set_device(s->prm.device);
    const mvd_session::P2P& q = s->p2p;
    HaloWaitK::Params p;
    memset(&p, 0, sizeof(p));
    p.flags = s->p2p.d_ctl + which * 32;
    p.nslots = q.npieces;
    for (int i = 0; i < q.npieces; ++i) p.slot[i] = q.slot_here[i];
    p.epoch = s->p2p.d_ctl + 65 + which;
    p.err = s->p2p.d_ctl + 67;
    static int timeout_ms = env_int("SPIM_P2P_TIMEOUT_S", 30) * 1000;
    p.timeout_ns = (unsigned long long)timeout_ms * 1000000ull;
    rt::launch<HaloWaitK>(p, 1, 32, 0, s->stream);
    s->p2p.waits[which] += 1;
    return 0;
    SPIM_API_END
}

int mvd_p2p_status(mvd_session* s, int* timed_out) {
    SPIM_API_BEGIN
    if (!s || !timed_out) return fail("mvd_p2p_status: null argument");
    *timed_out = 0;
    if (!s->p2p.d_ctl) return 0;
    rt::set_device(s->prm.device);
    unsigned int e = 0;
    rt::d2h(&e, s->p2p.d_ctl + 67, sizeof(e), s->stream);
    rt::stream_sync(s->stream);
    *timed_out = e ? 1 : 0;
    return 0;
    SPIM_API_END
}

int mvd_p2p_disconnect(mvd_session* s) {
    SPIM_API_BEGIN
    if (!s) return fail("mvd_p2p_disconnect: null session");
    rt::set_device(s->prm.device);
    rt::stream_sync(s->stream);
    p2p_disconnect(s);
    return 0;
    SPIM_API_END
}

// ---- stand-alone helpers ---------------------------------------------------------------------
static std::mutex g_ws_mutex[64];
static ConvWorkspace* g_ws[64][2] = {{nullptr}};   // per device: [0] general, [1] legacy (periodic-exact)

static int convolve_common(const float* img, const int im_dims[3], const float* kernel, const int kdims[3],
                           int ext, float value, float* out, int device, bool legacy) {
    if (!img || !im_dims || !kernel || !kdims || !out) return fail("convolve: null argument");
    for (int d = 0; d < 3; ++d) if (im_dims[d] < 1 || kdims[d] < 1) return fail("convolve: bad dims");
    const int ndev = rt::device_count();
    if (ndev <= 0) return fail("convolve: no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev || device >= 64) return fail("convolve: bad device ordinal");
    std::lock_guard<std::mutex> lock(g_ws_mutex[device]);
    rt::set_device(device);
    ConvWorkspace*& ws = g_ws[device][legacy ? 1 : 0];
    if (!ws) ws = new ConvWorkspace();
    ws->ensure(im_dims, kdims, legacy);
    ws->run(img, kernel, ext, value, out);
    return 0;
}

int mvd_convolve(const float* img, const int im_dims[3], const float* kernel, const int kernel_dims[3],
                 int ext, float ext_value, float* out, int device) {
    SPIM_API_BEGIN
    if (ext < 0 || ext > 4) return fail("mvd_convolve: bad extension mode");
    return convolve_common(img, im_dims, kernel, kernel_dims, ext, ext_value, out, device, false);
    SPIM_API_END
}

long long mvd_debug_counter(int which) { return debug_counter(which).load(); }
int mvd_fft_size(int min_n, int need_even) { return choose_fft_size(min_n, need_even != 0); }

// ================================================================================================
// legacy JNA boundary (include/spim_fftconv.h)
// ================================================================================================
const char* spim_fftconv_last_error(void) { return g_last_error.c_str(); }

int getNumDevicesCUDA(void) { return rt::device_count(); }

#if defined(SPIM_HOST_EMU)
int getCUDAcomputeCapabilityMajorVersion(int) { return 10; }
int getCUDAcomputeCapabilityMinorVersion(int) { return 0; }
void getNameDeviceCUDA(int, char* name) { if (name) { memset(name, 0, 256); strcpy(name, "HOST EMULATOR (tests only)"); } }
long long getMemDeviceCUDA(int) { return 1LL << 34; }
long long getFreeMemDeviceCUDA(int) { return 1LL << 34; }
#else
static bool dev_prop(int dev, cudaDeviceProp& p) {
    const int n = rt::device_count();
    if (dev < 0 || dev >= n) return false;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { cudaGetLastError(); return false; }
    return true;
}
int getCUDAcomputeCapabilityMajorVersion(int dev) { cudaDeviceProp p; return dev_prop(dev, p) ? p.major : -1; }
int getCUDAcomputeCapabilityMinorVersion(int dev) { cudaDeviceProp p; return dev_prop(dev, p) ? p.minor : -1; }
void getNameDeviceCUDA(int dev, char* name) {
    if (!name) return;
    memset(name, 0, 256);
    cudaDeviceProp p;
    if (dev_prop(dev, p)) strncpy(name, p.name, 255);
}
long long getMemDeviceCUDA(int dev) { cudaDeviceProp p; return dev_prop(dev, p) ? (long long)p.totalGlobalMem : -1; }
long long getFreeMemDeviceCUDA(int dev) {
    const int n = rt::device_count();
    if (dev < 0 || dev >= n) return -1;
    int cur = 0;
    cudaGetDevice(&cur);
    size_t fr = 0, tot = 0;
    if (cudaSetDevice(dev) != cudaSuccess || cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); return -1; }
    cudaSetDevice(cur);
    return (long long)fr;
}
#endif

void convolution3DfftCUDAInPlace(float* im, int* imDim, float* kernel, int* kernelDim, int devCUDA) {
    int rc = 1;
    try {
        rc = convolve_common(im, imDim, kernel, kernelDim, EXT_PERIODIC, 0.f, im, devCUDA, true);
    } catch (const std::exception& e) { fail(e.what()); }
    catch (...) { fail("unknown error"); }
    if (rc) fprintf(stderr, "convolution3DfftCUDAInPlace failed: %s (buffer left untouched)\n", g_last_error.c_str());
    else g_last_error.clear();      // a void entry: callers that can read spim_fftconv_last_error() see "" after a success
}

float* convolution3DfftCUDA(float* im, int* imDim, float* kernel, int* kernelDim, int devCUDA) {
    if (!im || !imDim) { fail("convolution3DfftCUDA: null argument"); return nullptr; }
    const size_t nvox = (size_t)imDim[0] * imDim[1] * imDim[2];
    float* out = (float*)malloc(nvox * sizeof(float));
    if (!out) { fail("convolution3DfftCUDA: out of host memory"); return nullptr; }
    int rc = 1;
    try {
        rc = convolve_common(im, imDim, kernel, kernelDim, EXT_PERIODIC, 0.f, out, devCUDA, true);
    } catch (const std::exception& e) { fail(e.what()); }
    catch (...) { fail("unknown error"); }
    if (rc) {
        fprintf(stderr, "convolution3DfftCUDA failed: %s\n", g_last_error.c_str());
        free(out);
        return nullptr;
    }
    g_last_error.clear();
    return out;
}

}  // extern "C"

#include "fusion_api.h"
