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

// host-side launch counters (mvd_debug_counter): [0] column passes launched with narrow tiles
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

    // ascending: smallest radix first.  Stage 0 is the register-resident stage of the x kernels (first from global memory in
    // x-forward, last with the fused epilogue in x-inverse); a small radix there keeps the epilogue's per-item state small
    void create(int n, bool ascending = false) {
        if (!plan_radices(n, radices)) throw rt::Error("unsupported FFT length " + std::to_string(n));
        if (ascending) std::reverse(radices.begin(), radices.end());
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

// block sizes (tunable through the environment for experiments; defaults chosen from ncu runs)
inline int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    const int r = atoi(v);
    return (r >= 0 && r <= 1024) ? r : dflt;
}
// the same, read on every call (switches the tests flip inside one process)
inline int env_int_now(const char* name, int dflt) { return env_int(name, dflt); }
inline int threads_xfwd() { static int t = env_int("SPIM_THREADS_XFWD", 192); return t; }
inline int threads_col() { static int t = env_int("SPIM_THREADS_COL", 128); return t; }
// Column tiles larger than a third of the shared memory leave room for only two / one block per SM: scale the block
// so that ~384 threads stay resident (2 x 192, 1 x 384).  SPIM_THREADS_COL, when set, wins.
inline int threads_col_for(size_t smem_bytes, size_t smem_limit) {
    static int user = env_int("SPIM_THREADS_COL", 0);
    if (user > 0) return user;
    const size_t blocks = smem_limit / (smem_bytes + 1024);   // 1 KB per block is reserved by the driver
    return blocks >= 3 ? 128 : (blocks == 2 ? 192 : 384);
}
inline int threads_xinv() { static int t = env_int("SPIM_THREADS_XINV", 128); return t; }
inline int threads_colt() { static int t = env_int("SPIM_THREADS_COLT", 480); return t; }
// SPIM_REGCAP (experiments): 1 = column passes from an instantiation capped at 85 registers (3 x 256 threads per SM),
// 2 = y tiles with 3 x 192 threads (<= 113 registers), 3 = z tiles with 6 x 128 threads (<= 85 registers)
inline int use_regcap() { static int t = env_int("SPIM_REGCAP", 0); return t; }
// SPIM_SERPENTINE=1 (experiment): the y-forward pass and the x-inverse pass walk their tiles from the last to the first, so
// that each starts on the ~100 MB its predecessor (x-forward / y-inverse, which end at the high planes) has just left in
// the 126 MB L2, and the next x-forward pass (ascending) starts on what the x-inverse pass wrote last
inline int use_serpentine() { return env_int_now("SPIM_SERPENTINE", 0); }
inline int use_colp() { static int t = env_int("SPIM_COLP", 2); return t; }   // 0 direct loads in the first stage, 2 one-shot cp.async tile staging (default), 3 experimental TMA pipeline, 4 experimental warp-private columns

// per-axis override for A/B runs: SPIM_COLP_Y / SPIM_COLP_Z (e.g. the TMA pipeline for the 72 KB y tiles only)
// Defaults (measured on B200, profiles/): the y passes (72 KB tiles at the bench size) run the persistent TMA / mbarrier
// pipeline, the z pass (36 KB tiles, five blocks per SM) the one-shot cp.async staging.
inline int use_colp_for(int axis) {
    static int y = env_int("SPIM_COLP_Y", -1), z = env_int("SPIM_COLP_Z", -1);
    static int all = env_int("SPIM_COLP", -1);
    const int v = axis == 1 ? y : z;
    if (v >= 0) return v;
    if (all >= 0) return all;
    return axis == 1 ? 3 : 2;
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
        fx.create(N2, env_int_now("SPIM_XPLAN_ASC", 0) != 0); fy.create(P[1]); fz.create(P[0]);
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
    struct XFix { int2* d = nullptr; int n = 0; };
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
        XFix r;
        r.n = (int)f.size();
        r.d = (int2*)rt::dmalloc(sizeof(int2) * std::max<size_t>(1, f.size()));
        if (!f.empty()) rt::h2d(r.d, f.data(), sizeof(int2) * f.size(), st);
        rt::stream_sync(st);
        xfixes[key] = r;
        return r;
    }
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

    void x_forward(const SrcDesc& src, const Geom& g, float2* out, rt::Stream st) {
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
        const int tma = env_int_now("SPIM_XFWD_TMA", 1);
        if (tma && (reinterpret_cast<uintptr_t>(src.p) & 15) == 0 && p.sx % 4 == 0 && p.ox % 2 == 0 && grid <= 0x7fffffff) {
            XFwdTParams q;
            memset(&q, 0, sizeof(q));
            q.x = p;
            q.LS = (std::max(p.sx, p.ox + P[2]) + 3) & ~3;
            q.row_bytes = (unsigned)p.sx * 4u;
            const size_t slot_bytes = (size_t)TC * q.LS * sizeof(float);
            const XFix fx_ = x_fix_table(p.nx, p.hpx, p.hmx, p.sx, p.ox, p.ext, (src.halo_lo >> 2) & 1, (src.halo_hi >> 2) & 1, st);
            const size_t fixed = (size_t)N2 * TC * sizeof(float2) + (XFwdT::MAXSLOT + 1) * TC * sizeof(long long) + XFwdT::MAXSLOT * sizeof(uint64_t) +
                                 (size_t)std::max(1, fx_.n) * sizeof(int2);
            const size_t lim = rt::max_smem();
            // blocks per SM / slots per block: three blocks with one slot each where that fits (tiles up to ~36 KB), else two
            // blocks with one slot, else one block with as many slots as fit
            int nslot = env_int_now("SPIM_XFWD_SLOTS", 0), bps = 0;
            if (nslot <= 0) nslot = 1;
            nslot = std::min(nslot, (int)XFwdT::MAXSLOT);
            while (nslot > 1 && fixed + nslot * slot_bytes + 1024 > lim) --nslot;
            const size_t sm = fixed + nslot * slot_bytes;
            if (sm + 1024 <= lim) {
                bps = (int)std::min<size_t>(3, lim / (sm + 1024));
                q.fix = fx_.d; q.nfix = fx_.n;
                q.magic_nfix = magic_for(std::max(1, fx_.n));
                q.cval = p.ext == EXT_CONSTANT ? p.ext_value : 0.f;
                q.const_row = const_row(p.sx, q.cval, st);
                q.nslot = nslot;
                q.ntiles = (int)grid;
                q.nctas = (int)std::min<long long>(grid, (long long)rt::sm_count() * bps);
                if (timer) timer->begin(K_XFWD, st);
                const int T = env_int_now("SPIM_THREADS_XFWDT", 0);
                if (bps >= 3) rt::launch<XFwdT, 256, 3>(q, q.nctas, T > 0 ? T : 256, sm, st);
                else if (bps == 2) rt::launch<XFwdT, 384, 2>(q, q.nctas, T > 0 ? T : 384, sm, st);
                else rt::launch<XFwdT, 512, 1>(q, q.nctas, T > 0 ? T : 512, sm, st);
                if (timer) timer->end(K_XFWD, st);
                return;
            }
        }
        if (timer) timer->begin(K_XFWD, st);
        // four 192-thread blocks per SM (the measured round-1 configuration) need <= 85 registers
        // SPIM_THREADS_XFWD=160 (experiment): the phases of a 280-point tile hold 280 / 320 / 448 / 282 items -- 160 threads need
        // the same 2 + 2 + 3 + 2 rounds as 192 do, so five blocks of 160 (72 registers) replace four of 192
        if (threads_xfwd() == 160 && 5 * (smem + 1024) <= rt::max_smem()) rt::launch<XFwd, 160, 5>(p, grid, 160, smem, st);
        else if (threads_xfwd() <= 192) rt::launch<XFwd, 192, 4>(p, grid, threads_xfwd(), smem, st);
        else rt::launch<XFwd>(p, grid, threads_xfwd(), smem, st);
        if (timer) timer->end(K_XFWD, st);
    }

    void col_pass(int id, float2* data, const float2* khat, int axis /*1=y,0=z*/, int mode, const Geom& g,
                  int out_rows, int outer_valid_lo, int outer_count, int outer_P, rt::Stream st) {
        ColPassParams p;
        memset(&p, 0, sizeof(p));
        p.data = data; p.khat = khat;
        p.plan = (axis == 1) ? fy.dev : fz.dev;
        const int Pa = P[axis];
        // narrow tiles (8 columns, 64-byte rows): automatically where a 16-column tile would leave one block per SM
        // (FFT lengths above ~880, e.g. the 1080-long axes of a 1024^2 x 512 volume on one GPU); SPIM_COL_NARROW=0/1 forces
        const size_t lim = rt::max_smem();
        const int narrow_env = env_int_now("SPIM_COL_NARROW", -1);
        int colp = use_colp_for(axis);
        // the TMA pipeline keeps three tiles per block: axes too long for that (FFT lengths above ~590) fall back to cp.async staging
        if (colp == 3 && 3 * ((size_t)Pa * TC * sizeof(float2)) + 64 > lim) colp = 2;
        const bool narrow = colp == 2 &&
                            (narrow_env >= 0 ? narrow_env != 0 : 2 * ((size_t)Pa * TC * sizeof(float2) + 1024) > lim);
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
        p.nblocks = (int)grid;
        p.reverse = (use_serpentine() && axis == 1 && mode == COL_FWD && colp == 2 && grid <= 0x7fffffff) ? 1 : 0;
        const size_t smem = (size_t)Pa * tcols * sizeof(float2);
        if (timer) timer->begin(id, st);
        if (narrow) {
            debug_counter(0) += 1;
            p.ntiles = -1;    // async mode flag
            p.kstage = 0;
            const int T = threads_col_for(smem, lim);
            // 36 KB tiles and smaller: keep five 128-thread blocks per SM, like the 16-column small-tile instantiation
            if (smem <= 40 * 1024 && T <= 128) rt::launch<ColPassNarrow, 128, 5>(p, grid, T, smem, st);
            else rt::launch<ColPassNarrow>(p, grid, T, smem, st);
        } else if (colp == 3 && 3 * smem + 64 <= lim && grid <= 0x7fffffff) {
            // experimental TMA / mbarrier pipeline: correct, but slower than the default in round 1 (see kernels.h)
            p.kstage = 0;
            p.ntiles = (int)grid;
            p.nctas = (int)std::min<long long>(grid, (long long)rt::sm_count());
            int T = threads_colt();
            if (T < 96) T = 96;
            // tensor-map producer (SPIM_TMAP=1): one request per box of box_rows rows instead of one per row
            static int want_tmap = env_int("SPIM_TMAP", 1);
            p.use_tmap = 0;
            if (want_tmap) {
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
            }
            rt::launch<ColPassT, 512>(p, p.nctas, T, 3 * smem + 64, st);
        } else if (colp == 4) {
            // experimental warp-private-column variant: 4 warps per tile, no CTA barriers between stages
            rt::launch<ColPassW>(p, grid, 128, smem, st);
        } else if (colp >= 2) {
            p.ntiles = -1;    // async mode flag
            static int ks = env_int("SPIM_KSTAGE", 0);
            p.kstage = (ks && mode == COL_MID && 2 * smem <= 76 * 1024) ? 1 : 0;
            const int rc = use_regcap();
            // SPIM_COL_LEAN=1 (experiment): plans without radices 9 / 10 (288 = 8*6*6, the z axis of the bench volume) run from
            // an instantiation compiled for radices <= 8 only: 80 registers without spills instead of 127 (96 when capped), so
            // six 128-thread blocks of a 36 KB tile are resident per SM instead of five
            int rmax = 0;
            for (int s_ = 0; s_ < p.plan.nstages; ++s_) rmax = std::max(rmax, p.plan.radix[s_]);
            if (env_int_now("SPIM_COL_LEAN", 0) && rmax <= 8 && !p.kstage) {
                if (smem <= 37 * 1024) rt::launch<ColPassR8, 128, 6>(p, grid, 128, smem, st);
                else rt::launch<ColPassR8, 256, 1>(p, grid, threads_col_for(smem, lim), smem, st);
            }
            else if (rc == 1) rt::launch<ColPass, 256, 3>(p, grid, threads_col(), (p.kstage ? 2 : 1) * smem, st);
            // SPIM_REGCAP=2: large tiles (three 72 KB y tiles per SM) with 192 threads each, <= 113 registers: 18 warps
            else if (rc == 2 && smem > 40 * 1024 && 3 * (smem + 1024) <= lim && !p.kstage) rt::launch<ColPass, 192, 3>(p, grid, 192, smem, st);
            // SPIM_REGCAP=3: small tiles (36 KB z tiles) as six blocks of 128 threads per SM, <= 85 registers: 24 warps
            else if (rc == 3 && smem <= 37 * 1024 && !p.kstage) rt::launch<ColPass, 128, 6>(p, grid, 128, smem, st);
            else if (smem <= 40 * 1024 && threads_col() <= 128 && !p.kstage)
                // small tiles (z pass of the 512x512x256 brick: 36 KB): registers, not shared memory, limit the resident
                // blocks -- keep 5 blocks of 128 threads per SM (<= 102 registers) as in the measured round-1 binary
                rt::launch<ColPass, 128, 5>(p, grid, threads_col(), smem, st);
            else {
                const size_t sm = (p.kstage ? 2 : 1) * smem;
                const int T = threads_col_for(sm, lim);
                if (T > 256) rt::launch<ColPass, 384>(p, grid, T, sm, st);     // one 1080-row tile per SM
                else rt::launch<ColPass>(p, grid, T, sm, st);
            }
        } else {
            rt::launch<ColPass>(p, grid, threads_col(), smem, st);
        }
        if (timer) timer->end(id, st);
    }

    void x_inverse(const float2* in, const EpiDesc& e, rt::Stream st) {
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
        p.two_lambda = (float)(2.0 * e.lambda);
        p.exact_tikhonov = e.exact_tikhonov;
        static int fast_env = env_int("SPIM_FAST_EPI", -1);      // A/B switch for the benchmarks: 0 / 1 override the parameter
        p.fast_epilogue = fast_env >= 0 ? (fast_env ? 1 : 0) : (e.fast_epilogue ? 1 : 0);
        p.stat_sum = e.stat_sum; p.stat_max = e.stat_max;
        auto al8 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 7) == 0; };
        p.vec_ok = al8(e.dst) && (p.dsx % 2 == 0) && (p.dox % 2 == 0) && (n[2] % 2 == 0) && al8(e.img) && al8(e.weight);
        const long long grid = (p.nlines + TC - 1) / TC;
        const int xinv_id = e.epi == EPI_UPDATE ? K_XINV_UPDATE : (e.epi == EPI_RATIO ? K_XINV : K_XINV_STORE);
        p.nblocks = (int)grid;
        p.reverse = (use_serpentine() && e.epi != EPI_STORE && grid <= 0x7fffffff) ? 1 : 0;
        const size_t smem = (size_t)N2 * TC * sizeof(float2) + 3 * TC * sizeof(long long);
        // TMA-fed persistent pipeline (XInvP, the default): spectrum rows are always 16-byte aligned
        if (env_int_now("SPIM_XINV_TMA", 1) && grid <= 0x7fffffff) {
            XInvPParams q;
            memset(&q, 0, sizeof(q));
            q.x = p;
            q.row_bytes = (unsigned)pitch * (unsigned)sizeof(float2);
            q.prefetch = env_int_now("SPIM_XINV_PREFETCH", 1);
            const size_t slot_bytes = (size_t)TC * pitch * sizeof(float2);
            const size_t fixed = (size_t)N2 * TC * sizeof(float2) + 2 * TC * sizeof(long long) + 3 * sizeof(uint64_t);
            const size_t lim = rt::max_smem();
            int nslot = std::max(1, std::min(3, env_int_now("SPIM_XINV_SLOTS", 1)));
            while (nslot > 1 && fixed + nslot * slot_bytes + 1024 > lim) --nslot;
            const size_t sm = fixed + nslot * slot_bytes;
            if (sm + 1024 <= lim) {
                const int bps = (int)std::min<size_t>(3, lim / (sm + 1024));
                q.nslot = nslot;
                q.ntiles = (int)grid;
                const int T = env_int_now("SPIM_THREADS_XINVP", 0);
                if (timer) timer->begin(xinv_id, st);
                const bool hot = p.fast_epilogue && !e.exact_tikhonov && e.epi != EPI_STORE;
                if (hot && bps >= 3) {
                    q.nctas = (int)std::min<long long>(grid, 3LL * rt::sm_count());
                    // ratio: 80 registers, three blocks of 256 threads; update: 128 registers uncapped -- 160 threads run it
                    // without spills, 192 threads (one round less per tile) cap it at 96 registers with ~130 bytes of spills
                    if (e.epi == EPI_RATIO) rt::launch<XInvPRatioFast, 256, 3>(q, q.nctas, T > 0 ? T : 256, sm, st);
                    else if (T > 0 && T <= 160) rt::launch<XInvPUpdateFast, 160, 3>(q, q.nctas, T, sm, st);
                    else rt::launch<XInvPUpdateFast, 192, 3>(q, q.nctas, T > 0 ? T : 192, sm, st);
                } else if (hot && bps == 1) {
                    q.nctas = (int)std::min<long long>(grid, (long long)rt::sm_count());
                    if (e.epi == EPI_RATIO) rt::launch<XInvPRatioFast, 512, 1>(q, q.nctas, T > 0 ? T : 512, sm, st);
                    else rt::launch<XInvPUpdateFast, 512, 1>(q, q.nctas, T > 0 ? T : 512, sm, st);
                } else {
                    const int b2 = std::min(bps, 2);
                    q.nctas = (int)std::min<long long>(grid, (long long)b2 * rt::sm_count());
                    const int T2 = T > 0 ? T : 256;
                    if (e.epi == EPI_STORE) rt::launch<XInvPStore, 256, 2>(q, q.nctas, T2, sm, st);
                    else if (e.epi == EPI_RATIO) {
                        if (p.fast_epilogue) rt::launch<XInvPRatioFast, 256, 2>(q, q.nctas, T2, sm, st);
                        else rt::launch<XInvPRatioIeee, 256, 2>(q, q.nctas, T2, sm, st);
                    }
                    else if (e.exact_tikhonov) rt::launch<XInvPUpdateExact64, 256, 2>(q, q.nctas, T2, sm, st);
                    else if (p.fast_epilogue) rt::launch<XInvPUpdateFast, 256, 2>(q, q.nctas, T2, sm, st);
                    else rt::launch<XInvPUpdateIeee, 256, 2>(q, q.nctas, T2, sm, st);
                }
                if (timer) timer->end(xinv_id, st);
                return;
            }
        }
        if (timer) timer->begin(xinv_id, st);
        const int T = threads_xinv();
        // 36 KB tiles: six 128-thread blocks fit per SM as long as the ratio kernel stays within 85 registers
        const bool cap6 = T <= 128 && smem <= 37 * 1024;
        if (e.epi == EPI_STORE) rt::launch<XInvT<EPI_STORE, MATH_IEEE>>(p, grid, T, smem, st);
        else if (e.epi == EPI_RATIO) {
            if (p.fast_epilogue) {
                // SPIM_THREADS_XINV=160 (experiment): 11 rounds per tile instead of 15 with 128 threads, five blocks per SM
                if (T == 160 && 5 * (smem + 1024) <= rt::max_smem()) rt::launch<XInvRatioFast, 160, 5>(p, grid, 160, smem, st);
                else if (cap6) rt::launch<XInvT<EPI_RATIO, MATH_FAST>, 128, 6>(p, grid, T, smem, st);
                else rt::launch<XInvT<EPI_RATIO, MATH_FAST>>(p, grid, T, smem, st);
            } else {
                if (cap6) rt::launch<XInvT<EPI_RATIO, MATH_IEEE>, 128, 6>(p, grid, T, smem, st);
                else rt::launch<XInvT<EPI_RATIO, MATH_IEEE>>(p, grid, T, smem, st);
            }
        }
        else if (e.exact_tikhonov) rt::launch<XInvT<EPI_UPDATE, MATH_EXACT64>>(p, grid, T, smem, st);
        else if (p.fast_epilogue) {
            // SPIM_XINV_CAP=5 (experiment): five 128-thread blocks of the update kernel per SM (96 registers, ~270 bytes of
            // spills) instead of four at 128 registers
            static int cap = env_int("SPIM_XINV_CAP", 0);
            // SPIM_XINV_R0=1 (experiment, with SPIM_XPLAN_ASC=1): when the x plan starts with a small radix, run the update
            // from an instantiation compiled for stage-0 radices <= 5 (80 registers, six 128-thread blocks per SM, no spills)
            // or <= 7 (96 registers, five blocks) instead of the general one (128 registers, four blocks)
            const int lean = env_int_now("SPIM_XINV_R0", 0);
            const int r0 = fx.dev.radix[0];
            if (lean && r0 <= 5 && T == 160 && 5 * (smem + 1024) <= rt::max_smem()) rt::launch<XInvUpdateFastR5, 160, 5>(p, grid, 160, smem, st);
            else if (lean && r0 <= 5 && T <= 128 && 6 * (smem + 1024) <= rt::max_smem()) rt::launch<XInvUpdateFastR5, 128, 6>(p, grid, T, smem, st);
            else if (lean && r0 <= 7 && T <= 128 && 5 * (smem + 1024) <= rt::max_smem()) rt::launch<XInvUpdateFastR7, 128, 5>(p, grid, T, smem, st);
            else if (cap == 5 && T <= 128 && 5 * (smem + 1024) <= rt::max_smem()) rt::launch<XInvUpdateFast, 128, 5>(p, grid, T, smem, st);
            else rt::launch<XInvT<EPI_UPDATE, MATH_FAST>>(p, grid, T, smem, st);
        }
        else rt::launch<XInvT<EPI_UPDATE, MATH_IEEE>>(p, grid, T, smem, st);
        if (timer) timer->end(xinv_id, st);
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
    void convolve(const SrcDesc& src, const float2* khat, const EpiDesc& e, rt::Stream st) {
        Geom g;
        for (int d = 0; d < 3; ++d) { g.n[d] = n[d]; g.hp[d] = hp[d]; g.hm[d] = hm[d]; }
        x_forward(src, g, spec, st);
        const int LZ = n[0] + hp[0] + hm[0];
        col_pass(K_YFWD, spec, nullptr, 1, COL_FWD, g, P[1], n[0] + hp[0], LZ, P[0], st);
        col_pass(K_ZMID, spec, khat, 0, COL_MID, g, n[0], P[1], P[1], P[1], st);
        col_pass(K_YINV, spec, nullptr, 1, COL_INV, g, n[1], n[0], n[0], P[0], st);
        x_inverse(spec, e, st);
    }
};

}  // namespace spim
