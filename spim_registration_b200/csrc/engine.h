// Host-side driver of the FFT-convolution engine: FFT size selection, stage planning, twiddle
// tables, and the five-sweep convolution built from the kernel bodies in kernels.h.
#pragma once
#include "kernels.h"
#include "runtime.h"
#include "instances.h"
#include <algorithm>
#include <cmath>
#include <map>
#include <memory>
#include <mutex>
#include <tuple>
#include <atomic>

namespace spim {

// host-side launch counters (mvd_debug_counter): [0] column passes launched with narrow tiles, [1] convolutions whose
// forward sweeps ran de-duplicated, [2] convolutions with a dropped (zero) halo, [3] x-inverse launches with the fused halo push
inline std::atomic<long long>& debug_counter(int i) { static std::atomic<long long> c[4]; return c[i & 3]; }

// x-inverse launches are timed per epilogue: K_XINV = ratio (conv1), K_XINV_UPDATE = update (conv2), K_XINV_STORE = plain store
enum KernelId { K_XFWD = 0, K_YFWD, K_ZMID, K_YINV, K_XINV, K_ZFWD, K_XINV_STORE, K_XINV_UPDATE, K_COUNT };

// ------------------------------------------------------------------------------------------
// stage planning
// ------------------------------------------------------------------------------------------
inline bool is_smooth(int n, int maxp) {
    if (n < 1) return false;
    for (int p : {2, 3, 5, 7, 11, 13}) {
        if (p > maxp) break;
        while (n % p == 0) n /= p;
    }
    return n == 1;
}

// factor n into <= MAX_STAGES radices from {2..16}: fewest stages, then smallest radix sum
inline bool plan_radices(int n, std::vector<int>& best) {
    static const int allowed[] = {16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2};
    std::vector<int> cur;
    int best_sum = 1 << 30;
    best.clear();
    struct Rec {
        static void go(int rem, int start, std::vector<int>& cur, std::vector<int>& best, int& best_sum) {
            if (rem == 1) {
                int sum = 0;
                for (int r : cur) sum += r;
                if (best.empty() || cur.size() < best.size() || (cur.size() == best.size() && sum < best_sum)) {
                    best = cur;
                    best_sum = sum;
                }
                return;
            }
            if ((int)cur.size() >= MAX_STAGES) return;
            if (!best.empty() && cur.size() + 1 > best.size()) return;
            for (int i = start; i < 15; ++i) {
                int r = allowed[i];
                if (r > SPIM_MAX_RADIX || rem % r) continue;
                cur.push_back(r);
                go(rem / r, i, cur, best, best_sum);
                cur.pop_back();
            }
        }
    };
    if (n == 1) return false;
    Rec::go(n, 0, cur, best, best_sum);
    return !best.empty();
}

// smallest supported FFT length >= min_n (even if need_even).  7-smooth sizes are preferred; an
// 11/13-smooth size is taken only when it is >3% shorter.
inline int choose_fft_size(int min_n, bool need_even) {
    if (min_n < 2) min_n = 2;
    if (need_even && min_n < 4) min_n = 4;
    int best7 = -1, best13 = -1;
    for (int n = min_n; n < 4 * min_n + 64; ++n) {
        if (need_even && (n & 1)) continue;
        const int m = need_even ? n / 2 : n;
        std::vector<int> r;
        if (best13 < 0 && is_smooth(n, 13) && plan_radices(m, r)) best13 = n;
        if (is_smooth(n, 7) && plan_radices(m, r)) { best7 = n; break; }
    }
    if (best7 < 0) return best13;
    if (best13 > 0 && best13 < 0.97 * best7) return best13;
    return best7;
}

struct FftPlanHost {
    FftPlanDev dev;
    float2* d_tw = nullptr;
    std::vector<int> radices;
    std::vector<unsigned short> pos;   // pos[k]: row that holds frequency k after the DIF stages

    void create(int n) {
        if (!plan_radices(n, radices)) throw rt::Error("unsupported FFT length " + std::to_string(n));
        memset(&dev, 0, sizeof(dev));
        dev.n = n;
        dev.nstages = (int)radices.size();
        int prod = 1;
        for (int s = 0; s < dev.nstages; ++s) {
            dev.radix[s] = radices[s];
            prod *= radices[s];
            dev.M[s] = n / prod;
            dev.magicM[s] = dev.M[s] > 1 ? (uint32_t)((0x100000000ull / (uint64_t)dev.M[s]) + 1ull) : 0u;
        }
        // per-stage twiddle tables, contiguous per butterfly: [j][p-1] = exp(-2 pi i j p / (M*R))
        std::vector<float2> tw;
        for (int s = 0; s < dev.nstages; ++s) {
            const int R = dev.radix[s], M = dev.M[s], L = M * R;
            dev.tw_off[s] = (int)tw.size();
            // M == 1 (the last stage): one butterfly's worth of ones, so that the register-resident stage of the x kernels can
            // load and apply its twiddles unconditionally when it is the only stage of a plan
            for (int j = 0; j < M; ++j)
                for (int p = 1; p < R; ++p) {
                    const double a = -2.0 * 3.14159265358979323846 * (double)j * (double)p / (double)L;
                    tw.push_back(make_float2((float)cos(a), (float)sin(a)));
                }
        }
        if (tw.empty()) tw.push_back(make_float2(1.f, 0.f));
        d_tw = (float2*)rt::dmalloc(sizeof(float2) * tw.size());
        rt::h2d(d_tw, tw.data(), sizeof(float2) * tw.size(), 0);
        rt::stream_sync(0);
        dev.tws = d_tw;
        pos.resize(n);
        for (int k = 0; k < n; ++k) {
            int rem = k, p = 0;
            for (int s = 0; s < dev.nstages; ++s) {
                const int d = rem % dev.radix[s];
                rem /= dev.radix[s];
                p += d * dev.M[s];
            }
            pos[k] = (unsigned short)p;
        }
    }
    void destroy() { rt::dfree(d_tw); d_tw = nullptr; }
};

// The few switches that remain after the round-2 A/B runs (profiles/README.md has the numbers behind every default):
//   SPIM_XFWD_TMA=0   x-forward from the plain-load kernel (XFwd) instead of the TMA-fed pipeline (XFwdT)
//   SPIM_COLP=0|2|3   column passes: 0 first stage straight from global memory, 2 one-shot cp.async staging, 3 persistent
//                     TMA / mbarrier pipeline; unset = 3 for large tiles (y passes of the bench volume), 2 for small ones
//   SPIM_COL_NARROW=0|1  force 16- / 8-column tiles (unset: narrow where a 16-column tile leaves one block per SM);
//                        SPIM_XFWD_LINES=16|8 the same for the lines per x-forward tile
//   SPIM_FAST_EPI=0|1    override mvd_params.fast_epilogue
//   SPIM_CONST_SHIFT=0   gen-2 conv2 with the literal constant extension instead of zero extension of (ratio - 1) (spim_b200.cu)
//   SPIM_DEDUP=0         forward sweeps transform every padded line / plane instead of only those that exist as data
inline int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    const int r = atoi(v);
    return (r >= 0 && r <= 1024) ? r : dflt;
}
// Column tiles larger than a third of the shared memory leave room for only two / one block per SM: scale the block
// so that ~384 threads stay resident (2 x 192, 1 x 384).
inline int threads_col_for(size_t smem_bytes, size_t smem_limit) {
    const size_t blocks = smem_limit / (smem_bytes + 1024);   // 1 KB per block is reserved by the driver
    // three blocks (72 KB tiles): 160 threads each -- 126 registers x 480 threads still fit the register file (1.82 vs 1.88 ms on
    // the z pass of the 1024^2 x 512 volume)
    return blocks >= 4 ? 128 : (blocks == 3 ? 160 : (blocks == 2 ? 192 : 384));
}

inline uint32_t magic_for(int d) { return d > 1 ? (uint32_t)((0x100000000ull / (uint64_t)d) + 1ull) : 0u; }

// ------------------------------------------------------------------------------------------
// geometry of one convolution: image n, kernel k, padded circular size P per axis ([z,y,x])
// ------------------------------------------------------------------------------------------
struct SrcDesc {            // real input volume
    const float* p = nullptr;
    int dims[3] = {0, 0, 0};    // array dims  [z,y,x]
    int origin[3] = {0, 0, 0};  // index = logical coordinate + origin
    int ext = EXT_ZERO;
    float ext_value = 0.f;
    int halo_lo = 7, halo_hi = 7;   // bit d: array cells outside [0,n) on that side of axis d are valid data
};

struct EpiDesc {            // what XInv does with the result
    int epi = EPI_STORE;
    float* dst = nullptr;
    int dst_dims[3] = {0, 0, 0};
    int dst_origin[3] = {0, 0, 0};
    const float* img = nullptr;
    const float* weight = nullptr;
    float const_weight = 1.f;
    double lambda = 0.0;
    float min_value = 1e-4f;
    int gen2_quotient = 1;
    float ratio_offset = 0.f;     // EPI_RATIO: added to the stored quotient
    float blur_offset = 0.f;      // EPI_UPDATE: added to the convolution result
    const HaloFuse* fuse = nullptr;   // brick mode with mapped peers: the epilogue also stores the neighbours' halo voxels (device pointer)
    int exact_tikhonov = 0;
    int fast_epilogue = 1;
    double* stat_sum = nullptr;
    unsigned int* stat_max = nullptr;
};

class ConvPlan {
public:
    int n[3], k[3], hp[3], hm[3], P[3];
    int pitch = 0, N2 = 0;
    FftPlanHost fx, fy, fz;
    unsigned short* d_pos = nullptr;
    float2* d_wx = nullptr;
    float2* spec = nullptr;       // work spectrum
    size_t spec_elems = 0;
    float* d_kernel = nullptr;    // staging for kernel uploads
    size_t kernel_cap = 0;
    rt::KernelTimer* timer = nullptr;

    // periodic_exact: when the image dims themselves are supported FFT sizes, use P = n with no halo
    // (pure circular convolution, the legacy JNA semantics on FFT-friendly block sizes)
    void create(const int n_[3], const int k_[3], bool periodic_exact = false, bool alloc_spec = true) {
        for (int d = 0; d < 3; ++d) {
            n[d] = n_[d]; k[d] = k_[d];
            hp[d] = k[d] / 2;
            hm[d] = k[d] - 1 - k[d] / 2;
            const bool even = (d == 2);
            bool exact = false;
            if (periodic_exact && n[d] >= 2 && n[d] >= k[d]) {
                std::vector<int> r;
                if ((!even || (n[d] % 2 == 0)) && is_smooth(n[d], 13) && plan_radices(even ? std::max(1, n[d] / 2) : n[d], r) ) exact = true;
                if (even && n[d] == 2) exact = false;
            }
            if (exact) { P[d] = n[d]; hp[d] = hm[d] = 0; }
            else P[d] = choose_fft_size(std::max(n[d] + k[d] - 1, even ? 4 : 2), even);
        }
        N2 = P[2] / 2;
        pitch = ((N2 + 1 + TC - 1) / TC) * TC;
        fx.create(N2); fy.create(P[1]); fz.create(P[0]);
        // one tile per block: [N2][16] float2 along x, [P][16] -- or, for long axes, the narrow [P][8] -- along y / z
        const size_t need_x = (size_t)N2 * TC * sizeof(float2) + 512;
        const size_t need_c = (size_t)std::max(P[1], P[0]) * (TC / 2) * sizeof(float2) + 512;
        if (std::max(need_x, need_c) > rt::max_smem())
            throw rt::Error("FFT axis too long for one shared-memory tile (padded x <= 3624, padded y / z <= 3624 voxels): "
                            "split the volume into blocks (blockSize) or bricks");
        d_pos = (unsigned short*)rt::dmalloc(sizeof(unsigned short) * N2);
        rt::h2d(d_pos, fx.pos.data(), sizeof(unsigned short) * N2, 0);
        const int nk = N2 / 2 + 1;
        std::vector<float2> wx(nk);
        for (int t = 0; t < nk; ++t) {
            const double a = -2.0 * 3.14159265358979323846 * (double)t / (double)P[2];
            wx[t] = make_float2((float)cos(a), (float)sin(a));
        }
        d_wx = (float2*)rt::dmalloc(sizeof(float2) * nk);
        rt::h2d(d_wx, wx.data(), sizeof(float2) * nk, 0);
        rt::stream_sync(0);
        spec_elems = (size_t)pitch * P[1] * P[0];
        if (alloc_spec) spec = (float2*)rt::dmalloc(spec_elems * sizeof(float2));
    }
    void destroy() {
        fx.destroy(); fy.destroy(); fz.destroy();
        rt::dfree(d_pos); rt::dfree(d_wx); rt::dfree(spec); rt::dfree(d_kernel);
        for (auto& kv : xtables) rt::dfree(kv.second);
        xtables.clear();
        for (auto& kv : xfixes) rt::dfree(kv.second.d);
        xfixes.clear();
        for (auto& kv : const_rows) rt::dfree(kv.second);
        const_rows.clear();
        for (auto& kv : dedups) rt::dfree(const_cast<int*>(kv.second.d_dup));
        dedups.clear();
        d_pos = nullptr; d_wx = nullptr; spec = nullptr; d_kernel = nullptr; kernel_cap = 0;
    }
    size_t spec_bytes() const { return spec_elems * sizeof(float2); }
    double kernel_scale() const { return 1.0 / (4.0 * (double)P[0] * (double)P[1] * (double)P[2]); }
    long long padded_min_voxels() const {   // Np of SURVEY section 8d
        return (long long)(n[0] + k[0] - 1) * (n[1] + k[1] - 1) * (n[2] + k[2] - 1);
    }

    // x-position -> source-index tables of the x-forward loader, cached per geometry
    struct XKey { int nx, hp, hm, sx, ox, ext, vlo, vhi; bool operator<(const XKey& o) const {
        return std::tie(nx, hp, hm, sx, ox, ext, vlo, vhi) < std::tie(o.nx, o.hp, o.hm, o.sx, o.ox, o.ext, o.vlo, o.vhi); } };
    std::map<XKey, int*> xtables;
    const int* x_index_table(int nx, int hp, int hm, int sx, int ox, int ext, int vlo, int vhi, rt::Stream st) {
        const XKey key{nx, hp, hm, sx, ox, ext, vlo, vhi};
        auto it = xtables.find(key);
        if (it != xtables.end()) return it->second;
        std::vector<int> t(P[2]);
        for (int u = 0; u < P[2]; ++u) {
            const int a = pad_to_coord(u, nx, hp, hm, P[2]);
            if (a == kGap) { t[u] = -1; continue; }
            int i = a + ox;
            if ((unsigned)i >= (unsigned)sx || (a < 0 && !vlo) || (a >= nx && !vhi)) {
                const int e = ext_map(a, nx, ext);
                i = e < 0 ? -2 : e + ox;
            }
            t[u] = i;
        }
        int* d = (int*)rt::dmalloc(sizeof(int) * P[2]);
        rt::h2d(d, t.data(), sizeof(int) * P[2], st);
        rt::stream_sync(st);
        xtables[key] = d;
        return d;
    }

    // fix-up list of the TMA-fed x-forward kernel (XFwdT): every padded position whose value is not already where the bulk
    // copy of the array row puts it (staging index ox + u): {ox + u, source index | -1 zero gap | -2 constant}
    struct XFix { int2* d = nullptr; int n = 0; int ndyn = 0; };
    std::map<XKey, XFix> xfixes;
    XFix x_fix_table(int nx, int hp, int hm, int sx, int ox, int ext, int vlo, int vhi, rt::Stream st) {
        const XKey key{nx, hp, hm, sx, ox, ext, vlo, vhi};
        auto it = xfixes.find(key);
        if (it != xfixes.end()) return it->second;
        std::vector<int2> f;
        for (int u = 0; u < P[2]; ++u) {
            const int a = pad_to_coord(u, nx, hp, hm, P[2]);
            int i;
            if (a == kGap) i = -1;
            else {
                i = a + ox;
                if ((unsigned)i >= (unsigned)sx || (a < 0 && !vlo) || (a >= nx && !vhi)) {
                    const int e = ext_map(a, nx, ext);
                    i = e < 0 ? -2 : e + ox;
                }
            }
            if (i != ox + u) { int2 e; e.x = ox + u; e.y = i; f.push_back(e); }
        }
        // entries that copy data first; constants written to cells beyond the row copy (index >= sx) last: those survive from
        // tile to tile in the staging slot and are applied once
        std::stable_partition(f.begin(), f.end(), [sx](const int2& e) { return !(e.y < 0 && e.x >= sx); });
        XFix r;
        r.ndyn = (int)std::count_if(f.begin(), f.end(), [sx](const int2& e) { return !(e.y < 0 && e.x >= sx); });
        r.n = (int)f.size();
        r.d = (int2*)rt::dmalloc(sizeof(int2) * std::max<size_t>(1, f.size()));
        if (!f.empty()) rt::h2d(r.d, f.data(), sizeof(int2) * f.size(), st);
        rt::stream_sync(st);
        xfixes[key] = r;
        return r;
    }
    // De-duplicated forward sweeps (mirror / periodic extension): along y and z the halo lines / planes that the source array
    // does not provide as data are copies of image lines / planes, so the x-forward pass transforms only the lines that exist
    // and stores each spectrum to its own row and to the halo row that mirrors it, and the forward y pass does the same with
    // whole planes.  Per axis: U unique coordinates, enumerated [0, split) then the neighbour-provided halo before the image;
    // dup[u] = padded position of the halo coordinate that maps to u, or -1.
    struct AxisDedup { bool ok = false; int U = 0, split = 0; const int* d_dup = nullptr; bool any = false; };
    struct DKey { int n, hp, hm, P, ext, lo, hi; bool operator<(const DKey& o) const {
        return std::tie(n, hp, hm, P, ext, lo, hi) < std::tie(o.n, o.hp, o.hm, o.P, o.ext, o.lo, o.hi); } };
    std::map<DKey, AxisDedup> dedups;
    AxisDedup axis_dedup(int n_, int hp_, int hm_, int P_, int ext, bool data_lo, bool data_hi, rt::Stream st) {
        const DKey key{n_, hp_, hm_, P_, ext, data_lo ? 1 : 0, data_hi ? 1 : 0};
        auto it = dedups.find(key);
        if (it != dedups.end()) return it->second;
        AxisDedup r;
        r.split = n_ + (data_hi ? hp_ : 0);
        r.U = r.split + (data_lo ? hm_ : 0);
        std::vector<int> dup((size_t)r.U, -1);
        bool ok = (ext == EXT_MIRROR_SINGLE || ext == EXT_MIRROR_DOUBLE || ext == EXT_PERIODIC);
        auto add = [&](int b) {           // halo coordinate b is not data: which image coordinate is it a copy of?
            const int a = ext_map(b, n_, ext);
            if (a < 0 || a >= n_ || dup[(size_t)a] >= 0) { ok = false; return; }
            dup[(size_t)a] = b >= 0 ? b : P_ + b;
            r.any = true;
        };
        if (!data_lo) for (int b = -hm_; b < 0 && ok; ++b) add(b);
        if (!data_hi) for (int b = n_; b < n_ + hp_ && ok; ++b) add(b);
        r.ok = ok;
        if (ok) {
            int* d = (int*)rt::dmalloc(sizeof(int) * (size_t)std::max(1, r.U));
            rt::h2d(d, dup.data(), sizeof(int) * (size_t)r.U, st);
            rt::stream_sync(st);
            r.d_dup = d;
        }
        dedups[key] = r;
        return r;
    }
    struct Dedup { bool on = false; AxisDedup y, z; };

    // a row of the out-of-bounds constant in global memory: lines that are constant along y / z are bulk-copied from it
    std::map<std::pair<int, unsigned>, float*> const_rows;
    const float* const_row(int sx, float value, rt::Stream st) {
        unsigned bits;
        memcpy(&bits, &value, 4);
        const auto key = std::make_pair(sx, bits);
        auto it = const_rows.find(key);
        if (it != const_rows.end()) return it->second;
        std::vector<float> h((size_t)sx, value);
        float* d = (float*)rt::dmalloc(sizeof(float) * (size_t)sx);
        rt::h2d(d, h.data(), sizeof(float) * (size_t)sx, st);
        rt::stream_sync(st);
        const_rows[key] = d;
        return d;
    }

    // ---- sweeps -------------------------------------------------------------------------
    struct Geom { int n[3], hp[3], hm[3]; };   // logical size + halos of whatever is being transformed

    // dd (optional): in = de-duplication wanted, out = whether this launch used it (only the TMA-fed kernel does)
    void x_forward(const SrcDesc& src, const Geom& g, float2* out, rt::Stream st, Dedup* dd = nullptr) {
        XFwdParams p;
        memset(&p, 0, sizeof(p));
        p.src = src.p;
        p.sz = src.dims[0]; p.sy = src.dims[1]; p.sx = src.dims[2];
        p.oz = src.origin[0]; p.oy = src.origin[1]; p.ox = src.origin[2];
        p.nz = g.n[0]; p.ny = g.n[1]; p.nx = g.n[2];
        p.hpz = g.hp[0]; p.hpy = g.hp[1]; p.hpx = g.hp[2];
        p.hmz = g.hm[0]; p.hmy = g.hm[1]; p.hmx = g.hm[2];
        p.ext = src.ext; p.ext_value = src.ext_value;
        p.halo_lo = src.halo_lo; p.halo_hi = src.halo_hi;
        p.Pz = P[0]; p.Py = P[1]; p.Px = P[2]; p.pitch = pitch;
        p.spec = out;
        p.plan = fx.dev; p.pos = d_pos; p.wx = d_wx;
        p.LY = g.n[1] + g.hp[1] + g.hm[1];
        p.LZ = g.n[0] + g.hp[0] + g.hm[0];
        p.nlines = (long long)p.LY * p.LZ;
        p.magic_m0 = magic_for(fx.dev.M[0]);
        p.nk = N2 / 2 + 1;
        p.magic_nk = magic_for(p.nk);
        p.src_vec_ok = ((reinterpret_cast<uintptr_t>(src.p) & 7) == 0) && (p.sx % 2 == 0) && (p.ox % 2 == 0);
        p.xidx = x_index_table(p.nx, p.hpx, p.hmx, p.sx, p.ox, p.ext, (src.halo_lo >> 2) & 1, (src.halo_hi >> 2) & 1, st);
        const long long grid = (p.nlines + TC - 1) / TC;
        const size_t smem = (size_t)N2 * TC * sizeof(float2) + 2 * TC * sizeof(long long);
        // TMA-fed persistent pipeline (XFwdT, the default): needs 16-byte aligned source rows and an even x origin
        if (env_int("SPIM_XFWD_TMA", 1) && (reinterpret_cast<uintptr_t>(src.p) & 15) == 0 && p.sx % 4 == 0 && p.ox % 2 == 0 && grid <= 0x7fffffff) {
            XFwdTParams q;
            memset(&q, 0, sizeof(q));
            q.x = p;
            q.LS = (std::max(p.sx, p.ox + P[2]) + 3) & ~3;
            q.row_bytes = (unsigned)p.sx * 4u;
            const XFix fx_ = x_fix_table(p.nx, p.hpx, p.hmx, p.sx, p.ox, p.ext, (src.halo_lo >> 2) & 1, (src.halo_hi >> 2) & 1, st);
            const size_t lim = rt::max_smem();
            // shared memory of a block with L lines per tile: packed tile + one staging slot + descriptors + fix-up list
            auto smem_for = [&](int L) {
                return (size_t)N2 * L * sizeof(float2) + (size_t)L * q.LS * sizeof(float) + (XFwdT::MAXSLOT + 1) * L * 2 * sizeof(long long) +
                       XFwdT::MAXSLOT * sizeof(uint64_t) + (XFwdT::MAXSLOT + 1) * sizeof(int) + (size_t)std::max(1, fx_.n) * sizeof(int2);
            };
            // 16 lines per tile; lines so long that this leaves one block per SM (1080 voxels and more) run 8-line tiles, three
            // blocks per SM again (SPIM_XFWD_LINES=16 / 8 forces either for A/B runs and tests)
            int L = TC;
            const int want = env_int("SPIM_XFWD_LINES", 0);
            if (want == 8 || (want != 16 && 2 * (smem_for(TC) + 1024) > lim)) L = TC / 2;
            const size_t sm = smem_for(L);
            if (sm + 1024 <= lim) {
                const int bps = (int)std::min<size_t>(3, lim / (sm + 1024));
                q.fix = fx_.d; q.nfix = fx_.n;
                q.magic_nfix = magic_for(std::max(1, fx_.n));
                q.nfix_dyn = fx_.ndyn;
                q.magic_nfix_dyn = magic_for(std::max(1, fx_.ndyn));
                q.cval = p.ext == EXT_CONSTANT ? p.ext_value : 0.f;
                q.const_row = const_row(p.sx, q.cval, st);
                q.nslot = 1;
                if (dd && dd->on) {
                    q.dedup = 1;
                    q.UY = dd->y.U; q.UZ = dd->z.U; q.splity = dd->y.split; q.splitz = dd->z.split;
                    q.dupy = dd->y.d_dup;
                    q.x.nlines = (long long)q.UY * q.UZ;
                }
                const long long tiles = (q.x.nlines + L - 1) / L;
                q.ntiles = (int)tiles;
                q.nctas = (int)std::min<long long>(tiles, (long long)rt::sm_count() * bps);
                if (timer) timer->begin(K_XFWD, st);
                // three blocks per SM: 160 threads each -- the phases of a tile hold (N2 / R) * 8 items (280 / 320 / 448 for the
                // 280-point plan 8 * 7 * 5, 282 for the split step), which 160 threads cover in 2 + 2 + 3 + 2 rounds at 92 % lane
                // use where 256 threads need 2 + 2 + 2 + 2 at 65 % (measured 0.197 vs 0.208 ms, profiles/r2)
                if (L == TC / 2) {
                    if (bps >= 3) rt::launch<XFwdTNarrow, 256, 3>(q, q.nctas, 160, sm, st);      // 216 / 240 / 360 / 271 items per phase: 1.35 ms with 160 threads, 1.39 with 256
                    else rt::launch<XFwdTNarrow, 384, 2>(q, q.nctas, bps == 2 ? 384 : 512, sm, st);
                }
                else if (bps >= 3) rt::launch<XFwdT, 256, 3>(q, q.nctas, 160, sm, st);
                else if (bps == 2) rt::launch<XFwdT, 384, 2>(q, q.nctas, 384, sm, st);
                // one block per SM: 768 threads at <= 85 registers -- the 432 / 480 / 720 / 542 items of a 540-point tile then take
                // one round per phase
                else rt::launch<XFwdT, 768, 1>(q, q.nctas, 768, sm, st);
                if (timer) timer->end(K_XFWD, st);
                return;
            }
        }
        if (dd) dd->on = false;        // the plain-load kernel transforms every padded line
        if (timer) timer->begin(K_XFWD, st);
        rt::launch<XFwd, 192, 4>(p, grid, 192, smem, st);       // four 192-thread blocks per SM (<= 85 registers)
        if (timer) timer->end(K_XFWD, st);
    }

    void col_pass(int id, float2* data, const float2* khat, int axis /*1=y,0=z*/, int mode, const Geom& g,
                  int out_rows, int outer_valid_lo, int outer_count, int outer_P, rt::Stream st, const int* dup_outer = nullptr) {
        ColPassParams p;
        memset(&p, 0, sizeof(p));
        p.dup_outer = dup_outer;
        p.data = data; p.khat = khat;
        p.plan = (axis == 1) ? fy.dev : fz.dev;
        const int Pa = P[axis];
        // narrow tiles (8 columns, 64-byte rows): automatically where a 16-column tile would leave one block per SM
        // (FFT lengths above ~880, e.g. the 1080-long axes of a 1024^2 x 512 volume on one GPU); SPIM_COL_NARROW=0/1 forces
        const size_t lim = rt::max_smem();
        const int narrow_env = env_int("SPIM_COL_NARROW", -1);
        // staging: the persistent TMA / mbarrier pipeline for the plain forward / inverse passes where three tiles fit (FFT
        // lengths up to ~590: the 72 KB y tiles of the bench volume run 0.136 / 0.126 ms instead of 0.161 / 0.148), but not for
        // small tiles (<= 40 KB: five or six blocks per SM with one-shot cp.async staging are faster, 0.247 vs 0.300 ms on the
        // z pass) and not for the fused forward-multiply-inverse pass (72 KB z tiles of the 1024^2 x 512 volume: 1.84 vs 1.94 ms)
        const size_t tile16 = (size_t)Pa * TC * sizeof(float2);
        int colp = env_int("SPIM_COLP", -1);
        if (colp < 0) colp = (tile16 > 40 * 1024 && mode != COL_MID) ? 3 : 2;
        if (colp == 3 && 3 * tile16 + 64 > lim) colp = 2;
        const bool narrow = colp == 2 && (narrow_env >= 0 ? narrow_env != 0 : 2 * (tile16 + 1024) > lim);
        const int tcols = narrow ? TC / 2 : TC;
        p.ntx = pitch / tcols;
        if (axis == 1) { p.row_stride = pitch; p.outer_stride = (long long)pitch * P[1]; }
        else { p.row_stride = (long long)pitch * P[1]; p.outer_stride = pitch; }
        // outer index map: first outer_valid_lo indices map to themselves, the rest to the top of the axis
        p.outer_split = outer_valid_lo;
        p.outer_shift = outer_P - outer_count;
        if (mode == COL_INV) { p.va = Pa; p.vb = Pa; }
        else { p.va = g.n[axis] + g.hp[axis]; p.vb = Pa - g.hm[axis]; }
        p.sa = out_rows;
        p.mode = mode;
        const long long grid = (long long)p.ntx * outer_count;
        const size_t smem = (size_t)Pa * tcols * sizeof(float2);
        if (timer) timer->begin(id, st);
        if (narrow) {
            debug_counter(0) += 1;
            p.ntiles = -1;    // async mode flag
            const int T = threads_col_for(smem, lim);
            // 36 KB tiles and smaller: keep five 128-thread blocks per SM, like the 16-column small-tile instantiation
            if (smem <= 40 * 1024 && T <= 128) rt::launch<ColPassNarrow, 128, 5>(p, grid, T, smem, st);
            else rt::launch<ColPassNarrow>(p, grid, T, smem, st);
        } else if (colp == 3 && grid <= 0x7fffffff) {
            // persistent, warp-specialised: one CTA per SM, a producer warp and two consumer groups over a ring of three tiles
            p.ntiles = (int)grid;
            p.nctas = (int)std::min<long long>(grid, (long long)rt::sm_count());
            // tensor-map producer: one request per box of box_rows rows instead of one bulk copy per row
            p.use_tmap = 0;
            int br = 0;
            for (int d = std::min(Pa, 256); d >= 1; --d) if (Pa % d == 0) { br = d; break; }
            const unsigned long long fl = 2ull * pitch;                      // floats per row
            bool ok = br >= 8 && Pa / br <= 32;
            if (ok && axis == 1) {
                const unsigned long long dims[2] = {fl, (unsigned long long)P[1] * P[0]};
                const unsigned long long str[1] = {fl * 4};
                const unsigned int box[2] = {2 * TC, (unsigned)br};
                ok = rt::encode_tensor_map(&p.tmap, data, 2, dims, str, box);
                p.tmap_rank = 2;
            } else if (ok) {
                const unsigned long long dims[3] = {fl, (unsigned long long)P[1], (unsigned long long)P[0]};
                const unsigned long long str[2] = {fl * 4, fl * 4 * P[1]};
                const unsigned int box[3] = {2 * TC, 1, (unsigned)br};
                ok = rt::encode_tensor_map(&p.tmap, data, 3, dims, str, box);
                p.tmap_rank = 3;
            }
            if (ok) { p.use_tmap = 1; p.box_rows = br; }
            rt::launch<ColPassT, 512>(p, p.nctas, 480, 3 * smem + 64, st);
        } else {
            p.ntiles = colp >= 2 ? -1 : 0;    // -1: async mode (whole tile staged by cp.async)
            int rmax = 0;
            for (int s_ = 0; s_ < p.plan.nstages; ++s_) rmax = std::max(rmax, p.plan.radix[s_]);
            if (smem <= 37 * 1024 && rmax <= 8)
                // small tiles of plans without radices 9 / 10 (288 = 8 * 6 * 6, the z axis of the bench volume): the instantiation
                // compiled for radices <= 8 needs 80 registers -- six 128-thread blocks per SM (0.2445 vs 0.2467 ms)
                rt::launch<ColPassR8, 128, 6>(p, grid, 128, smem, st);
            else if (smem <= 40 * 1024)
                // small tiles: registers, not shared memory, limit the resident blocks -- five blocks of 128 threads (96 registers)
                rt::launch<ColPass, 128, 5>(p, grid, 128, smem, st);
            else {
                const int T = threads_col_for(smem, lim);
                if (T > 256) rt::launch<ColPass, 384>(p, grid, T, smem, st);     // one 1080-row tile per SM
                else rt::launch<ColPass>(p, grid, T, smem, st);
            }
        }
        if (timer) timer->end(id, st);
    }

    // returns true when the epilogue also pushed the neighbours' halos (e.fuse given and the fused instantiation applies)
    bool x_inverse(const float2* in, const EpiDesc& e, rt::Stream st) {
        XInvParams p;
        memset(&p, 0, sizeof(p));
        p.spec = in; p.pitch = pitch; p.Px = P[2]; p.Py = P[1];
        p.nz = n[0]; p.ny = n[1]; p.nx = n[2];
        p.plan = fx.dev; p.pos = d_pos; p.wx = d_wx;
        p.magic_m0 = magic_for(fx.dev.M[0]);
        p.nk = N2 / 2 + 1;
        p.magic_nk = magic_for(p.nk);
        p.nlines = (long long)n[0] * n[1];
        p.dst = e.dst;
        p.dsy = e.dst_dims[1]; p.dsx = e.dst_dims[2];
        p.doz = e.dst_origin[0]; p.doy = e.dst_origin[1]; p.dox = e.dst_origin[2];
        p.epi = e.epi; p.img = e.img; p.weight = e.weight; p.const_weight = e.const_weight;
        p.lambda = e.lambda; p.min_value = e.min_value; p.gen2_quotient = e.gen2_quotient;
        p.ratio_offset = e.ratio_offset; p.blur_offset = e.blur_offset;
        p.two_lambda = (float)(2.0 * e.lambda);
        p.exact_tikhonov = e.exact_tikhonov;
        static int fast_env = env_int("SPIM_FAST_EPI", -1);      // A/B switch for the benchmarks: 0 / 1 override the parameter
        p.fast_epilogue = fast_env >= 0 ? (fast_env ? 1 : 0) : (e.fast_epilogue ? 1 : 0);
        p.stat_sum = e.stat_sum; p.stat_max = e.stat_max;
        auto al8 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 7) == 0; };
        p.vec_ok = al8(e.dst) && (p.dsx % 2 == 0) && (p.dox % 2 == 0) && (n[2] % 2 == 0) && al8(e.img) && al8(e.weight);
        const long long grid = (p.nlines + TC - 1) / TC;
        const int xinv_id = e.epi == EPI_UPDATE ? K_XINV_UPDATE : (e.epi == EPI_RATIO ? K_XINV : K_XINV_STORE);
        const size_t smem = (size_t)N2 * TC * sizeof(float2) + 3 * TC * sizeof(long long);
        if (timer) timer->begin(xinv_id, st);
        // Block size: small tiles (<= 36 KB: five / six blocks per SM) run 128 threads, larger ones (three or fewer blocks per
        // SM) 256 -- measured on the 540-point lines of the 1024^2 x 512 volume: 2.16 ms with 256 threads, 2.41 with 128.
        // The register-capped instantiations keep four 192-thread (ratio, 80 registers) / five 128-thread (update, 96 registers)
        // blocks resident on the small tiles (update: 0.300 vs 0.336 ms uncapped).  A TMA-fed persistent variant of this kernel was
        // built and measured in round 2 (0.325 vs 0.300 ms average) and removed again: its staging slot halves the resident blocks.
        const bool small = smem <= 37 * 1024;
        const int T = small ? 128 : 256;
        if (e.fuse && p.vec_ok && p.fast_epilogue && !e.exact_tikhonov && (e.epi == EPI_RATIO || e.epi == EPI_UPDATE)) {
            p.fuse = e.fuse;
            const size_t sm = smem + kFuseSmemBytes;
            if (e.epi == EPI_RATIO) {
                if (small) rt::launch<XInvRatioFastFuse, 192, 4>(p, grid, 192, sm, st);
                else rt::launch<XInvRatioFastFuse>(p, grid, T, sm, st);
            } else {
                if (small) rt::launch<XInvUpdateFastFuse, 128, 5>(p, grid, T, sm, st);
                else rt::launch<XInvUpdateFastFuse>(p, grid, T, sm, st);
            }
            if (timer) timer->end(xinv_id, st);
            debug_counter(3) += 1;
            return true;
        }
        if (e.epi == EPI_STORE) rt::launch<XInvStore>(p, grid, T, smem, st);
        else if (e.epi == EPI_RATIO) {
            if (p.fast_epilogue) {
                // four 192-thread blocks per SM: 8192 tiles of the bench volume are 13.8 waves of 592 blocks (0.213 ms) where six
                // 128-thread blocks are 9.2 waves of 888 with a nearly empty tenth (0.236 ms)
                if (small) rt::launch<XInvRatioFast, 192, 4>(p, grid, 192, smem, st);
                else rt::launch<XInvRatioFast>(p, grid, T, smem, st);
            } else {
                if (small) rt::launch<XInvRatioIeee, 128, 6>(p, grid, T, smem, st);
                else rt::launch<XInvRatioIeee>(p, grid, T, smem, st);
            }
        }
        else if (e.exact_tikhonov) rt::launch<XInvUpdateExact64>(p, grid, T, smem, st);
        else if (p.fast_epilogue) {
            if (small) rt::launch<XInvUpdateFast, 128, 5>(p, grid, T, smem, st);
            else rt::launch<XInvUpdateFast>(p, grid, T, smem, st);
        }
        else rt::launch<XInvUpdateIeee>(p, grid, T, smem, st);
        if (timer) timer->end(xinv_id, st);
        return false;
    }

    // spectrum of a (host) kernel, pre-scaled by 1/(4 Px Py Pz), in the layout the mid pass expects
    void kernel_spectrum(const float* h_kernel, float2* khat, rt::Stream st) {
        const size_t kn = (size_t)k[0] * k[1] * k[2];
        std::vector<float> scaled(kn);
        const double s = kernel_scale();
        for (size_t i = 0; i < kn; ++i) scaled[i] = (float)((double)h_kernel[i] * s);
        if (kn > kernel_cap) {
            rt::dfree(d_kernel);
            d_kernel = (float*)rt::dmalloc(kn * sizeof(float));
            kernel_cap = kn;
        }
        rt::h2d(d_kernel, scaled.data(), kn * sizeof(float), st);
        rt::stream_sync(st);   // 'scaled' goes out of scope
        SrcDesc src;
        src.p = d_kernel;
        Geom g;
        for (int d = 0; d < 3; ++d) {
            const int c = k[d] / 2;
            src.dims[d] = k[d];
            src.origin[d] = c;          // logical coordinate a in [-c, k-1-c] -> index a + c
            g.n[d] = k[d] - c; g.hp[d] = 0; g.hm[d] = c;
        }
        src.ext = EXT_ZERO;
        x_forward(src, g, khat, st);
        const int LZ = g.n[0] + g.hm[0];
        col_pass(K_YFWD, khat, nullptr, 1, COL_FWD, g, P[1], g.n[0], LZ, P[0], st);
        col_pass(K_ZFWD, khat, nullptr, 0, COL_FWD, g, P[0], P[1], P[1], P[1], st);
    }

    // out = ext(src) (*) kernel, cropped to the logical image, through the epilogue
    bool convolve(const SrcDesc& src, const float2* khat, const EpiDesc& e, rt::Stream st) {
        Geom g;
        for (int d = 0; d < 3; ++d) { g.n[d] = n[d]; g.hp[d] = hp[d]; g.hm[d] = hm[d]; }
        if (src.ext == EXT_ZERO) {
            // zero extension: a halo that the source array does not provide as data is all zeros -- it joins the gap, and the
            // forward sweeps neither transform its lines and planes nor load its rows
            for (int d = 0; d < 3; ++d) {
                const int bit = 1 << d;
                const bool lo = (src.halo_lo & bit) && src.origin[d] >= hm[d];
                const bool hi = (src.halo_hi & bit) && src.dims[d] - src.origin[d] - n[d] >= hp[d];
                if (!lo) g.hm[d] = 0;
                if (!hi) g.hp[d] = 0;
            }
            debug_counter(2) += 1;
        }
        // mirror / periodic extension: transform every line and plane once (SPIM_DEDUP=0: every padded line, as in round 1)
        Dedup dd;
        const bool dedup_on = env_int("SPIM_DEDUP", 1) != 0;
        if (dedup_on && (src.ext == EXT_MIRROR_SINGLE || src.ext == EXT_MIRROR_DOUBLE || src.ext == EXT_PERIODIC)) {
            auto data_side = [&](int d, bool hi) {
                const int bit = 1 << d;
                return hi ? ((src.halo_hi & bit) && g.hp[d] > 0 && src.dims[d] - src.origin[d] - n[d] >= g.hp[d])
                          : ((src.halo_lo & bit) && g.hm[d] > 0 && src.origin[d] >= g.hm[d]);
            };
            dd.y = axis_dedup(n[1], g.hp[1], g.hm[1], P[1], src.ext, data_side(1, false), data_side(1, true), st);
            dd.z = axis_dedup(n[0], g.hp[0], g.hm[0], P[0], src.ext, data_side(0, false), data_side(0, true), st);
            dd.on = dd.y.ok && dd.z.ok && (dd.y.any || dd.z.any);
        }
        x_forward(src, g, spec, st, &dd);
        if (dd.on) debug_counter(1) += 1;
        const int LZ = n[0] + g.hp[0] + g.hm[0];
        if (dd.on) col_pass(K_YFWD, spec, nullptr, 1, COL_FWD, g, P[1], dd.z.split, dd.z.U, P[0], st, dd.z.any ? dd.z.d_dup : nullptr);
        else col_pass(K_YFWD, spec, nullptr, 1, COL_FWD, g, P[1], n[0] + g.hp[0], LZ, P[0], st);
        col_pass(K_ZMID, spec, khat, 0, COL_MID, g, n[0], P[1], P[1], P[1], st);
        col_pass(K_YINV, spec, nullptr, 1, COL_INV, g, n[1], n[0], n[0], P[0], st);
        return x_inverse(spec, e, st);
    }
};

}  // namespace spim
