// Kernel bodies of the FFT-convolution hot path.
//
// Data layout in HBM:
//   real volumes      [z][y][x] fp32, x contiguous (exactly what the JNA boundary passes)
//   spectrum / work   [zp][yp][pitch] float2, pitch = roundup(Px/2+1, 16); x-frequencies in natural
//                     order, y- and z-frequencies in the digit-reversed order the in-place
//                     decimation-in-frequency stages leave them in (the inverse passes are
//                     decimation-in-time and consume exactly that order, so no permutation pass
//                     ever touches HBM; the kernel spectrum is produced by the same forward passes
//                     and therefore sits in the same order).
//
// One FFT-convolution = five sweeps, each one read + one write of the volume:
//   XFwdT  real -> half spectrum along x; out-of-bounds extension (mirror / constant / periodic /
//          zero) is applied on load, the padded volume never exists in memory
//   ColPass(fwd, y)   in-place column FFT over tiles of 16 x-frequencies
//   ColPass(mid, z)   forward z-FFT, multiply with the kernel spectrum, inverse z-FFT in one kernel
//   ColPass(inv, y)
//   XInv   half spectrum -> real along x with the fused epilogue: plain store, ratio img/blur
//          (computeQuotient) or the weighted Tikhonov update of psi + change statistics
//          (computeFinalValues)
//
// Every block works on a shared-memory tile [rows][16] of float2 (8-column / 8-line tiles for long axes).  A stage is a set
// of independent radix-R butterflies on TWO adjacent columns executed in place on packed pairs (fft_math.h: FADD2 / FMUL2 /
// FFMA2 -- both columns share every butterfly instruction); tiles are staged by TMA (x-forward lines, y tiles) or cp.async
// (z tiles) and the last stage writes straight back from registers.  Bodies are written as "phases" of independent items
// separated by barriers so the same source runs under the CPU emulator used by the unit tests (see hd.h).
#pragma once
#include "fft_math.h"
#include "fast_math.h"

namespace spim {

constexpr int TC = 16;            // tile columns (float2) = one 128-byte segment per row
constexpr int TP = 8;             // the same row as float4 column pairs
constexpr int MAX_STAGES = 6;
constexpr int kThreads = 256;

enum ExtMode { EXT_ZERO = 0, EXT_CONSTANT = 1, EXT_MIRROR_SINGLE = 2, EXT_MIRROR_DOUBLE = 3, EXT_PERIODIC = 4 };
enum EpiMode { EPI_STORE = 0, EPI_RATIO = 1, EPI_UPDATE = 2 };
enum ColMode { COL_FWD = 0, COL_INV = 1, COL_MID = 2 };

struct FftPlanDev {
    int n;
    int nstages;
    int radix[MAX_STAGES];
    int M[MAX_STAGES];            // sub-transform length after stage s: n / (radix[0]*...*radix[s])
    uint32_t magicM[MAX_STAGES];  // floor(2^32 / M) + 1, for m / M with m < 2^16
    int tw_off[MAX_STAGES];       // offset of stage s in tws
    const float2* tws;            // per-stage twiddles: tws[tw_off[s] + j*(R-1) + (p-1)] = exp(-2 pi i j p / (M*R))
};

SPIM_DEV int fastdiv(int m, uint32_t magic) { return (int)spim_umulhi((uint32_t)m, magic); }

// ---------------------------------------------------------------------------------------------
// out-of-bounds rules
// ---------------------------------------------------------------------------------------------
// logical coordinate a (any integer) -> index in [0,n) or -1 when the rule yields a constant
SPIM_HD int ext_map(int a, int n, int mode) {
    if ((unsigned)a < (unsigned)n) return a;
    switch (mode) {
        case EXT_PERIODIC: {
            int m = a % n;
            return m < 0 ? m + n : m;
        }
        case EXT_MIRROR_SINGLE: {
            if (n == 1) return 0;
            int p = 2 * (n - 1);
            int m = a % p;
            if (m < 0) m += p;
            return m < n ? m : p - m;
        }
        case EXT_MIRROR_DOUBLE: {
            int p = 2 * n;
            int m = a % p;
            if (m < 0) m += p;
            return m < n ? m : p - 1 - m;
        }
        default: return -1;
    }
}

constexpr int kGap = -0x40000000;
// position u in the circular padded axis of length P -> logical coordinate, or kGap in the zero gap.
// [0, n+hp) holds coordinates 0..n+hp-1, [P-hm, P) holds coordinates -hm..-1.
SPIM_HD int pad_to_coord(int u, int n, int hp, int hm, int P) {
    if (u < n + hp) return u;
    if (u >= P - hm) return u - P;
    return kGap;
}

// ---------------------------------------------------------------------------------------------
// generic in-place radix stage on a [rows][8] tile of float4 (= 16 float2 columns).
// One work item = one radix-R butterfly on TWO adjacent columns (one 128-bit access per row).
// ---------------------------------------------------------------------------------------------
struct GRows {          // the global-memory side of a column tile
    float2* p;          // element (row 0, col 0), 16-byte aligned
    long long stride;   // float2 units between consecutive rows (even)
    int va, vb;         // rows that hold data on load: [0, va) U [vb, n); others read as zero
    int sa;             // rows written on store: [0, sa)
    long long dup4;     // != 0: every stored row also goes to the same tile dup4 float4 further (the mirrored plane, see ColPassParams::dup_outer)
};

SPIM_HD float2 lo2(float4 v) { return make_float2(v.x, v.y); }
SPIM_HD float2 hi2(float4 v) { return make_float2(v.z, v.w); }

// twiddles of one butterfly, fetched into registers at the very start of the item so that their L1 / L2
// latency overlaps the shared-memory loads and the butterfly instead of being exposed right before use
template <int R>
SPIM_DEV void load_twiddles(float2 (&w)[R], const float2* tw) {
#pragma unroll
    for (int p = 1; p < R; ++p) w[p] = spim_ldg(tw + (p - 1));
}
template <int R, bool INV>
SPIM_DEV void mul_twiddles1(float2 (&a)[R], const float2 (&w)[R]) {
#pragma unroll
    for (int p = 1; p < R; ++p) a[p] = INV ? cmulc(a[p], w[p]) : cmul(a[p], w[p]);
}

// physical float4 slot of column pair c2 in row `row`: the x kernels rotate the pairs so that their transposing first / last
// phases are conflict-free.  W = 8 pairs (128-byte rows) rotate by the row, W = 4 (64-byte rows, two rows per 128 bytes of
// banks) by every second row, so that 8 consecutive rows of one pair still cover all eight 16-byte bank groups
template <int W> SPIM_HD int xslot(int c2, int row) { return W == 8 ? ((c2 + row) & 7) : ((c2 + (row >> 1)) & 3); }

// twiddle multiply of a packed pair: x[p] *= w[p] (forward) or conj(w[p]) (inverse), w broadcast to both columns
template <int R, bool INV>
SPIM_DEV void mul_twiddles_c2(C2 (&x)[R], const float2 (&w)[R]) {
#pragma unroll
    for (int p = 1; p < R; ++p) x[p] = INV ? cmulc_s(x[p], w[p].x, w[p].y) : cmul_s(x[p], w[p].x, w[p].y);
}

template <int R, bool INV>
SPIM_DEV void apply_twiddles_c2(C2 (&x)[R], const float2* tw) {
#pragma unroll
    for (int p = 1; p < R; ++p) {
        const float2 w = spim_ldg(tw + (p - 1));
        x[p] = INV ? cmulc_s(x[p], w.x, w.y) : cmul_s(x[p], w.x, w.y);
    }
}

// W = float4 column pairs per tile row: TP (16 columns, 128-byte rows) or TP / 2 (narrow tiles: 8 columns, 64-byte rows, for
// axes so long that a 16-column tile would leave one block per SM).
// The butterflies run on packed pairs (fft_math.h, C2).  Global memory and freshly staged tiles hold the two columns
// interleaved (re_a, im_a, re_b, im_b); between the stages of one kernel the shared-memory tile holds them packed
// (re_a, re_b, im_a, im_b), so only a kernel's first load and last store pay the register shuffle.
// src_p / dst_p: the shared-memory source / destination of this stage is in the packed layout.
template <int R, bool INV, bool TW, int W = TP>
SPIM_DEV void stage_tile(const TG& tg, const FftPlanDev& pl, int s, float4* tile, int swz, int src_g, int dst_g, const GRows& g,
                         int src_p = 0, int dst_p = 0) {
    const int M = pl.M[s];
    const int L = M * R;
    const int nb = pl.n / R;
    const uint32_t magic = pl.magicM[s];
    const float2* twp = pl.tws + pl.tw_off[s];
    const long long gs4 = g.stride >> 1;
    const long long gstep = (long long)M * gs4;
    constexpr int LW = (W == 8) ? 3 : 2;
    static_assert(W == 8 || W == 4, "tile rows hold 8 or 4 column pairs");
    SPIM_FOR_ITEMS_TG(tg, i, nb * W) {
        const int c2 = i & (W - 1);
        const int m = i >> LW;
        const int blk = (M == 1) ? m : fastdiv(m, magic);
        const int j = m - blk * M;
        const int base = blk * L + j;
        C2 x[R];
        float2 w[R];
        if (TW) load_twiddles<R>(w, twp + j * (R - 1));
        if (src_g) {
            const float4* gp = reinterpret_cast<const float4*>(g.p) + (long long)base * gs4 + c2;
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int row = base + q * M;
                float4 v = ldg_stream(gp + q * gstep);   // gap rows hold stale data: load anyway, select zero
                const bool ok = (row < g.va) || (row >= g.vb);
                if (!ok) v = make_float4(0.f, 0.f, 0.f, 0.f);
                x[q] = c2_from_il(v);
            }
        } else {
            float4 v[R];
            if (!swz) {
                const float4* sp = tile + base * W + c2;
#pragma unroll
                for (int q = 0; q < R; ++q) v[q] = sp[q * M * W];
            } else {
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const int row = base + q * M;
                    v[q] = tile[row * W + xslot<W>(c2, row)];
                }
            }
            if (src_p) {
#pragma unroll
                for (int q = 0; q < R; ++q) x[q] = c2_from_pk(v[q]);
            } else {
#pragma unroll
                for (int q = 0; q < R; ++q) x[q] = c2_from_il(v[q]);
            }
        }
        if (!INV) {
            dft<R, false>(x);
            if (TW) mul_twiddles_c2<R, false>(x, w);
        } else {
            if (TW) mul_twiddles_c2<R, true>(x, w);
            dft<R, true>(x);
        }
        if (dst_g) {
            // predicated streaming stores, the row pointer advanced by one 64-bit add per row
            float4* gp = reinterpret_cast<float4*>(g.p) + (long long)base * gs4 + c2;
            if (g.dup4 == 0) {
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    stg_stream_if(gp, c2_to_il(x[q]), base + q * M < g.sa);
                    gp += gstep;
                }
            } else {
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const float4 v = c2_to_il(x[q]);
                    const bool ok = base + q * M < g.sa;
                    stg_stream_if(gp, v, ok);
                    stg_stream_if(gp + g.dup4, v, ok);
                    gp += gstep;
                }
            }
        } else {
            float4 v[R];
            if (dst_p) {
#pragma unroll
                for (int q = 0; q < R; ++q) v[q] = c2_to_pk(x[q]);
            } else {
#pragma unroll
                for (int q = 0; q < R; ++q) v[q] = c2_to_il(x[q]);
            }
            if (!swz) {
                float4* sp = tile + base * W + c2;
#pragma unroll
                for (int q = 0; q < R; ++q) sp[q * M * W] = v[q];
            } else {
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const int row = base + q * M;
                    tile[row * W + xslot<W>(c2, row)] = v[q];
                }
            }
        }
    }
#if !defined(SPIM_EMU_NO_STAGE_BARRIER)      // negative control of tests/test_tsan_kernels.py: without it stages race
    tg_barrier(tg);
#endif
}

// last forward stage + kernel-spectrum multiply + first inverse stage, fused in registers (packed pairs; the kernel
// spectrum arrives interleaved from global memory or its staged copy).  src_p / dst_p as in stage_tile.
template <int R, int W = TP>
SPIM_DEV void mid_tile(const TG& tg, const FftPlanDev& pl, float4* tile, int src_g, int dst_g, const GRows& g, const float2* kh, long long ks4,
                       int src_p = 0, int dst_p = 0) {
    const int nb = pl.n / R;
    const long long gs4 = g.stride >> 1;
    constexpr int LW = (W == 8) ? 3 : 2;
    SPIM_FOR_ITEMS_TG(tg, i, nb * W) {
        const int c2 = i & (W - 1);
        const int base = (i >> LW) * R;
        C2 x[R];
        float4 kv[R];
        {
            const float4* kp = reinterpret_cast<const float4*>(kh) + (long long)base * ks4 + c2;
#pragma unroll
            for (int q = 0; q < R; ++q) kv[q] = ldg_stream(kp + q * ks4);
        }
        if (src_g) {
            const float4* gp = reinterpret_cast<const float4*>(g.p) + (long long)base * gs4 + c2;
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int row = base + q;
                float4 v = ldg_stream(gp + q * gs4);
                const bool ok = (row < g.va) || (row >= g.vb);
                if (!ok) v = make_float4(0.f, 0.f, 0.f, 0.f);
                x[q] = c2_from_il(v);
            }
        } else {
            const float4* sp = tile + base * W + c2;
            if (src_p) {
#pragma unroll
                for (int q = 0; q < R; ++q) x[q] = c2_from_pk(sp[q * W]);
            } else {
#pragma unroll
                for (int q = 0; q < R; ++q) x[q] = c2_from_il(sp[q * W]);
            }
        }
        dft<R, false>(x);
#pragma unroll
        for (int q = 0; q < R; ++q) x[q] = cmul(x[q], c2_from_il(kv[q]));
        dft<R, true>(x);
        if (dst_g) {
            float4* gp = reinterpret_cast<float4*>(g.p) + (long long)base * gs4 + c2;
#pragma unroll
            for (int q = 0; q < R; ++q) {
                stg_stream_if(gp, c2_to_il(x[q]), base + q < g.sa);
                gp += gs4;
            }
        } else {
            float4* sp = tile + base * W + c2;
            if (dst_p) {
#pragma unroll
                for (int q = 0; q < R; ++q) sp[q * W] = c2_to_pk(x[q]);
            } else {
#pragma unroll
                for (int q = 0; q < R; ++q) sp[q * W] = c2_to_il(x[q]);
            }
        }
    }
    tg_barrier(tg);
}

// radix dispatch.  SPIM_MAX_RADIX bounds the radices the planner may use (and therefore the code paths
// ptxas has to allocate registers for: a two-column radix-16 butterfly needs ~180 registers).
#ifndef SPIM_MAX_RADIX
#define SPIM_MAX_RADIX 10
#endif
#define SPIM_RADIX_CASE(N, CALL) case N: if constexpr (N <= SPIM_MAX_RADIX) { constexpr int RR = N; CALL; } break;
// the same switch restricted to radices <= MAXR (register-lean instantiations of a kernel phase)
#define SPIM_RADIX_CASE_MAX(N, MAXR, CALL) case N: if constexpr (N <= SPIM_MAX_RADIX && N <= (MAXR)) { constexpr int RR = N; CALL; } break;
#define SPIM_RADIX_SWITCH_MAX(R_, MAXR, CALL)                                                    \
    switch (R_) {                                                                                \
        SPIM_RADIX_CASE_MAX(2, MAXR, CALL) SPIM_RADIX_CASE_MAX(3, MAXR, CALL) SPIM_RADIX_CASE_MAX(4, MAXR, CALL)   \
        SPIM_RADIX_CASE_MAX(5, MAXR, CALL) SPIM_RADIX_CASE_MAX(6, MAXR, CALL) SPIM_RADIX_CASE_MAX(7, MAXR, CALL)   \
        SPIM_RADIX_CASE_MAX(8, MAXR, CALL) SPIM_RADIX_CASE_MAX(9, MAXR, CALL) SPIM_RADIX_CASE_MAX(10, MAXR, CALL)  \
        SPIM_RADIX_CASE_MAX(11, MAXR, CALL) SPIM_RADIX_CASE_MAX(12, MAXR, CALL) SPIM_RADIX_CASE_MAX(13, MAXR, CALL) \
        SPIM_RADIX_CASE_MAX(14, MAXR, CALL) SPIM_RADIX_CASE_MAX(15, MAXR, CALL) SPIM_RADIX_CASE_MAX(16, MAXR, CALL) \
        default: break;                                                                          \
    }
#define SPIM_RADIX_SWITCH(R_, CALL)                                                              \
    switch (R_) {                                                                                \
        SPIM_RADIX_CASE(2, CALL) SPIM_RADIX_CASE(3, CALL) SPIM_RADIX_CASE(4, CALL)               \
        SPIM_RADIX_CASE(5, CALL) SPIM_RADIX_CASE(6, CALL) SPIM_RADIX_CASE(7, CALL)               \
        SPIM_RADIX_CASE(8, CALL) SPIM_RADIX_CASE(9, CALL) SPIM_RADIX_CASE(10, CALL)              \
        SPIM_RADIX_CASE(11, CALL) SPIM_RADIX_CASE(12, CALL) SPIM_RADIX_CASE(13, CALL)            \
        SPIM_RADIX_CASE(14, CALL) SPIM_RADIX_CASE(15, CALL) SPIM_RADIX_CASE(16, CALL)            \
        default: break;                                                                          \
    }

// RMAX: largest radix the instantiation is compiled for (register-lean variants for plans without large radices)
template <bool INV, int W = TP, int RMAX = 16>
SPIM_DEV void stage_dispatch(const TG& tg, const FftPlanDev& pl, int s, float4* tile, int swz, int src_g, int dst_g, const GRows& g,
                             int src_p = 0, int dst_p = 0) {
    if (pl.M[s] > 1) { SPIM_RADIX_SWITCH_MAX(pl.radix[s], RMAX, (stage_tile<RR, INV, true, W>(tg, pl, s, tile, swz, src_g, dst_g, g, src_p, dst_p))) }
    else { SPIM_RADIX_SWITCH_MAX(pl.radix[s], RMAX, (stage_tile<RR, INV, false, W>(tg, pl, s, tile, swz, src_g, dst_g, g, src_p, dst_p))) }
}
template <int W = TP, int RMAX = 16>
SPIM_DEV void mid_dispatch(const TG& tg, const FftPlanDev& pl, float4* tile, int src_g, int dst_g, const GRows& g, const float2* kh, long long ks4,
                           int src_p = 0, int dst_p = 0) {
    SPIM_RADIX_SWITCH_MAX(pl.radix[pl.nstages - 1], RMAX, (mid_tile<RR, W>(tg, pl, tile, src_g, dst_g, g, kh, ks4, src_p, dst_p)))
}

// ---------------------------------------------------------------------------------------------
// ColPass: column FFT along y or z over tiles of 16 x-frequencies
// ---------------------------------------------------------------------------------------------
struct ColPassParams {
    float2* data;
    const float2* khat;        // COL_MID only
    FftPlanDev plan;
    int ntx;                   // number of 16-column tiles along x (pitch / 16)
    long long row_stride;      // between consecutive FFT rows
    long long outer_stride;    // between consecutive outer indices
    int outer_split, outer_shift;  // outer = o < split ? o : o + shift  (skips the zero gap)
    int va, vb, sa;
    int mode;
    const int* dup_outer;      // forward y pass over de-duplicated planes: dup_outer[o] = plane that is a mirror image of plane o
                               // (stored too) or -1; nullptr = none
    int ntiles, nctas;         // ColPassT (persistent): tiles in total / CTAs launched; ntiles < 0 selects the async mode of ColPass
    // ColPassT with tensor maps (use_tmap): y pass = 2-D map {2*pitch floats, Py*Pz rows}, box {32, box_rows};
    // z pass = 3-D map {2*pitch, Py, Pz}, box {32, 1, box_rows}; box_rows divides the FFT length
    int use_tmap, box_rows, tmap_rank;
    SpimTensorMap tmap;
};

// float4 distance from the tile of outer index o to its duplicate destination (0 = none)
SPIM_DEV long long col_dup4(const ColPassParams& p, int o) {
    if (p.dup_outer == nullptr) return 0;
    const int d = spim_ldg(p.dup_outer + o);
    if (d < 0) return 0;
    const int outer = o < p.outer_split ? o : o + p.outer_shift;
    return ((long long)(d - outer) * p.outer_stride) >> 1;
}

// asynchronous copy of tile rows [row_lo, row_hi) (W float4 = W x 16 B each) from global to shared memory.
// Each thread keeps its column pair and walks rows with a constant stride: ~4 instructions per 16-byte chunk.
template <int W>
SPIM_DEV void async_rows(float4* buf, const float4* gp, long long gs4, int row_lo, int row_hi) {
#if defined(SPIM_HOST_EMU)
    for (int row = row_lo + spim_emu_tid; row < row_hi; row += spim_emu_nthr)
        for (int c2 = 0; c2 < W; ++c2) cp_async16(buf + row * W + c2, gp + (long long)row * gs4 + c2);
#else
    constexpr int LW = (W == 8) ? 3 : 2;
    const int c2 = threadIdx.x & (W - 1);
    const int rstep = blockDim.x >> LW;
    int row = row_lo + (threadIdx.x >> LW);
    float4* d = buf + row * W + c2;
    const float4* g = gp + (long long)row * gs4 + c2;
    const long long gstep = (long long)rstep * gs4;
    const int dstep = rstep * W;
    // four rows per trip: one bounds check and one pointer update per four 16-byte copies
    for (; row + 3 * rstep < row_hi; row += 4 * rstep, d += 4 * dstep, g += 4 * gstep) {
        cp_async16(d, g);
        cp_async16(d + dstep, g + gstep);
        cp_async16(d + 2 * dstep, g + 2 * gstep);
        cp_async16(d + 3 * dstep, g + 3 * gstep);
    }
    for (; row < row_hi; row += rstep, d += dstep, g += gstep) cp_async16(d, g);
#endif
}

// W = TP: tiles of 16 x-frequencies (the default); W = TP / 2: narrow tiles of 8 (ntx = pitch / 8), half the shared memory
// per block -- for FFT lengths above ~880, where a 16-column tile would leave a single block per SM and nothing to overlap
// its load / compute / store phases with
template <int W, int RMAX = 16>
struct ColPassN {
    typedef ColPassParams Params;
    static constexpr bool kEmuThreads = true;
    SPIM_DEV static void run(const Params& p, int bid, float2* tile2) {
        const TG tg = tg_cta();
        float4* tile = reinterpret_cast<float4*>(tile2);
        const int o = bid / p.ntx;
        const int tx = bid - o * p.ntx;
        const int outer = o < p.outer_split ? o : o + p.outer_shift;
        const long long base = (long long)outer * p.outer_stride + (long long)tx * (2 * W);
        GRows g;
        g.p = p.data + base;
        g.stride = p.row_stride;
        g.va = p.va; g.vb = p.vb; g.sa = p.sa;
        g.dup4 = col_dup4(p, o);
        const FftPlanDev& pl = p.plan;
        const int S = pl.nstages;
        int sg = 1;                       // first stage reads global memory directly ...
        if (p.ntiles < 0) {               // ... or (async mode) the whole tile is staged with one burst of cp.async
            const int P = pl.n;
            const long long gs4 = p.row_stride >> 1;
            const float4* gp = reinterpret_cast<const float4*>(g.p);
            if (p.va < p.vb) {
                async_rows<W>(tile, gp, gs4, 0, p.va);
                async_rows<W>(tile, gp, gs4, p.vb, P);
            } else {
                async_rows<W>(tile, gp, gs4, 0, P);
            }
            cp_async_commit();
            if (p.va < p.vb) {
                SPIM_FOR_ITEMS(i, (p.vb - p.va) * W) tile[p.va * W + i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            cp_async_wait<0>();
            SPIM_BARRIER();
            sg = 0;
            g.va = P; g.vb = P;
        }
        // the tile is interleaved as loaded and packed between the stages (stage_tile)
        if (p.mode == COL_FWD) {
            for (int s = 0; s < S; ++s) stage_dispatch<false, W, RMAX>(tg, pl, s, tile, 0, sg && s == 0, s == S - 1, g, s > 0, 1);
        } else if (p.mode == COL_INV) {
            for (int s = S - 1; s >= 0; --s) stage_dispatch<true, W, RMAX>(tg, pl, s, tile, 0, sg && s == S - 1, s == 0, g, s < S - 1, 1);
        } else {
            for (int s = 0; s < S - 1; ++s) stage_dispatch<false, W, RMAX>(tg, pl, s, tile, 0, sg && s == 0, 0, g, s > 0, 1);
            mid_dispatch<W, RMAX>(tg, pl, tile, sg && S == 1, S == 1, g, p.khat + base, p.row_stride >> 1, S > 1, 1);
            for (int s = S - 2; s >= 0; --s) stage_dispatch<true, W, RMAX>(tg, pl, s, tile, 0, 0, s == 0, g, 1, 1);
        }
    }
};
typedef ColPassN<TP> ColPass;
typedef ColPassN<TP / 2> ColPassNarrow;
typedef ColPassN<TP, 8> ColPassR8;          // plans without radices 9 / 10 (e.g. 288 = 8 * 6 * 6): leaner in registers

// tile index -> offset of its first element (x tiles fastest, outer index mapped over the zero gap)
SPIM_DEV void col_tile_base(const ColPassParams& p, int t, long long& base) {
    const int o = t / p.ntx;
    const int tx = t - o * p.ntx;
    const int outer = o < p.outer_split ? o : o + p.outer_shift;
    base = (long long)outer * p.outer_stride + (long long)tx * TC;
}

// Persistent, warp-specialised variant with a TMA / mbarrier pipeline (EXPERIMENTAL, SPIM_COLP=3): correct on
// B200 (parity tests pass) but 3.7x slower than ColPass in round 1 -- 560 separate 128-byte bulk copies per
// tile saturate the TMA unit; the next step is 2-D tensor-map boxes (one request per 256 rows).  One CTA per SM; the last warp is the producer and streams whole tile rows into a ring
// of NSLOT shared-memory tiles with cp.async.bulk (UBLKCP, completion counted on an mbarrier); two consumer
// groups alternate over the tiles with their own named barriers, so one group's barrier / latency stalls are
// filled by the other; results leave straight from the registers of the last stage.
struct ColPassT {
    typedef ColPassParams Params;
    static constexpr bool kEmuThreads = false;
    static constexpr int NSLOT = 3;
    SPIM_DEV static void consume(const TG& tg, const Params& p, int t, float4* tile) {
        const FftPlanDev& pl = p.plan;
        const int P = pl.n, S = pl.nstages;
        if (p.va < p.vb) {   // rows of the zero gap are never loaded
            SPIM_FOR_ITEMS_TG(tg, i, (p.vb - p.va) * TP) tile[p.va * TP + i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tg_barrier(tg);
        long long base;
        col_tile_base(p, t, base);
        GRows g;
        g.p = p.data + base;
        g.stride = p.row_stride;
        g.va = P; g.vb = P; g.sa = p.sa;
        g.dup4 = col_dup4(p, t / p.ntx);
        if (p.mode == COL_FWD) {
            for (int s = 0; s < S; ++s) stage_dispatch<false>(tg, pl, s, tile, 0, 0, s == S - 1, g, s > 0, 1);
        } else if (p.mode == COL_INV) {
            for (int s = S - 1; s >= 0; --s) stage_dispatch<true>(tg, pl, s, tile, 0, 0, s == 0, g, s < S - 1, 1);
        } else {
            for (int s = 0; s < S - 1; ++s) stage_dispatch<false>(tg, pl, s, tile, 0, 0, 0, g, s > 0, 1);
            mid_dispatch(tg, pl, tile, 0, S == 1, g, p.khat + base, p.row_stride >> 1, S > 1, 1);
            for (int s = S - 2; s >= 0; --s) stage_dispatch<true>(tg, pl, s, tile, 0, 0, s == 0, g, 1, 1);
        }
    }
    SPIM_DEV static void run(const Params& p, int bid, float2* smem2) {
        const int P = p.plan.n;
        float4* smem = reinterpret_cast<float4*>(smem2);
        const long long gs4 = p.row_stride >> 1;
        const bool gap = p.va < p.vb;
        const int nvalid = gap ? P - (p.vb - p.va) : P;
#if defined(SPIM_HOST_EMU)
        const TG tg = tg_cta();
        for (int t = bid; t < p.ntiles; t += p.nctas) {
            long long base;
            col_tile_base(p, t, base);
            const float4* gp = reinterpret_cast<const float4*>(p.data + base);
            for (int r = 0; r < nvalid; ++r) {
                const int row = (gap && r >= p.va) ? r + (p.vb - p.va) : r;
                memcpy(smem + row * TP, gp + (long long)row * gs4, TP * sizeof(float4));
            }
            consume(tg, p, t, smem);
        }
#else
        uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)NSLOT * P * TP);
        uint64_t* empty = full + NSLOT;
        const int tid = threadIdx.x;
        const int nthr = blockDim.x;
        const int gsize = ((nthr - 32) / 64) * 32;      // two consumer groups of whole warps
        const int ntl = (p.ntiles - bid + p.nctas - 1) / p.nctas;
        if (tid == 0) {
            for (int s = 0; s < NSLOT; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
            mbar_fence_init();
        }
        __syncthreads();
        if (tid >= 2 * gsize) {
            if (tid >= 2 * gsize + 32) return;          // spare threads of a partial warp layout
            // ---- producer warp -------------------------------------------------------------------------
            const int lane = tid & 31;
            for (int i = 0; i < ntl; ++i) {
                const int slot = i % NSLOT, use = i / NSLOT;
                if (use > 0) mbar_wait(empty + slot, (unsigned)((use - 1) & 1));
                const int t = bid + i * p.nctas;
                long long base;
                col_tile_base(p, t, base);
                const float4* gp = reinterpret_cast<const float4*>(p.data + base);
                float4* dst = smem + (size_t)slot * P * TP;
                if (p.use_tmap) {
                    // whole tile (gap rows included: they are zeroed by the consumers after arrival) in P / box_rows requests
                    const int o = t / p.ntx;
                    const int tx = t - o * p.ntx;
                    const int outer = o < p.outer_split ? o : o + p.outer_shift;
                    const int nbox = P / p.box_rows;
                    if (lane == 0) mbar_expect_tx(full + slot, (unsigned)(P * TP * sizeof(float4)));
                    __syncwarp();
                    if (lane < nbox) {
                        float4* d = dst + (size_t)lane * p.box_rows * TP;
                        if (p.tmap_rank == 2) tma_load_2d(d, &p.tmap, tx * 2 * TC, outer * P + lane * p.box_rows, full + slot);
                        else tma_load_3d(d, &p.tmap, tx * 2 * TC, outer, lane * p.box_rows, full + slot);
                    }
                } else {
                    if (lane == 0) mbar_expect_tx(full + slot, (unsigned)(nvalid * TP * sizeof(float4)));
                    __syncwarp();
                    for (int r = lane; r < nvalid; r += 32) {
                        const int row = (gap && r >= p.va) ? r + (p.vb - p.va) : r;
                        bulk_g2s(dst + row * TP, gp + (long long)row * gs4, TP * sizeof(float4), full + slot);
                    }
                }
            }
        } else {
            // ---- two consumer groups, alternating tiles -------------------------------------------------
            const int group = tid / gsize;
            TG tg;
            tg.tid = tid - group * gsize; tg.n = gsize; tg.bar = 1 + group;
            for (int i = group; i < ntl; i += 2) {
                const int slot = i % NSLOT, use = i / NSLOT;
                mbar_wait(full + slot, (unsigned)(use & 1));
                float4* tile = smem + (size_t)slot * P * TP;
                consume(tg, p, bid + i * p.nctas, tile);
                // generic-proxy accesses to this slot are done; order them before the next async-proxy refill
                fence_proxy_async();
                tg_barrier(tg);
                if (tg.tid == 0) mbar_arrive(empty + slot);
            }
        }
#endif
    }
};

// ---------------------------------------------------------------------------------------------
// XFwd: real lines -> half spectrum (R2C via one complex FFT of length Px/2 + split step)
// Tile = [Px/2 rows][16 lines]; line pair bp of row r lives in float4 slot (bp + r) & 7.
// ---------------------------------------------------------------------------------------------
struct XFwdParams {
    const float* src;
    int sx, sy, sz;            // dims of the source array
    int ox, oy, oz;            // array index = logical coordinate + origin
    int nx, ny, nz;            // logical image size (coordinates [0,n) are 'inside' for the ext rule)
    int hpx, hmx, hpy, hmy, hpz, hmz;  // halo after (+) / before (-) the image on each axis
    int ext;
    float ext_value;
    int halo_lo, halo_hi;      // bit d (0=z,1=y,2=x): the source array holds valid neighbour data before / after the
                               // image on axis d (brick mode); otherwise the out-of-bounds rule applies there
    int Px, Py, Pz, pitch;
    float2* spec;
    FftPlanDev plan;           // n = Px / 2
    const unsigned short* pos; // pos[k] = tile row holding frequency k after the DIF stages
    const float2* wx;          // exp(-2 pi i k / Px), k in [0, Px/4]
    int LY, LZ;                // lines per axis that are not in the gap: n + hp + hm
    long long nlines;          // LY * LZ
    uint32_t magic_m0;         // for i / M[0]
    int nk;                    // Px/4 + 1 (pairs handled by the split step)
    uint32_t magic_nk;
    int src_vec_ok;            // float2 loads allowed (alignment)
    const int* xidx;           // [Px] x-position -> source index / -1 gap / -2 constant
};

constexpr long long kLineInvalid = -2;
constexpr long long kLineConst = -1;

// x-forward loader.  Interior pairs are one 8-byte load; everything else (halo, gap, constant or invalid
// lines) goes through a per-geometry index table: xidx[u] >= 0 source index, -1 zero gap, -2 ext constant.
SPIM_DEV float xfwd_val(const XFwdParams& p, long long so, int u) {
    const int i = spim_ldg(p.xidx + u);
    if (so >= 0 && i >= 0) return spim_ldg(p.src + so + i);
    if (i == -1 || so == kLineInvalid) return 0.f;
    return (p.ext == EXT_CONSTANT) ? p.ext_value : 0.f;
}

// everything that is not an interior pair of two real lines: halo, gap, constant or invalid lines, the last odd sample.
// Kept out of line (one call site per row, both lines of the pair) so that the first stage's inner loop is just the
// interior loads.
SPIM_NOINLINE_DEV float4 xfwd_pair_slow(const XFwdParams& p, long long so0, long long so1, int n) {
    return make_float4(xfwd_val(p, so0, 2 * n), xfwd_val(p, so0, 2 * n + 1), xfwd_val(p, so1, 2 * n), xfwd_val(p, so1, 2 * n + 1));
}

template <int R, bool VEC>
SPIM_DEV void xfwd_stage0(const XFwdParams& p, float4* tile, const long long* srcoff) {
    const FftPlanDev& pl = p.plan;
    const int M = pl.M[0];
    const float2* twp = pl.tws + pl.tw_off[0];
    const int nh = p.nx >> 1;                      // pairs (2n, 2n+1) that lie entirely inside the image: n < nx / 2
    SPIM_FOR_ITEMS(i, M * TP) {
        const int bp = (M == 1) ? i : fastdiv(i, p.magic_m0);
        const int m = i - bp * M;
        const long long so0 = srcoff[2 * bp], so1 = srcoff[2 * bp + 1];
        // x = 0 of the two lines (only dereferenced when both are real lines: so >= 0)
        const float* l0 = p.src + (so0 + p.ox);
        const float* l1 = p.src + (so1 + p.ox);
        const bool real2 = (so0 >= 0) && (so1 >= 0);
        float2 a[R], b[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int n = m + q * M;
            if (real2 && n < nh) {
                if (VEC) {
                    a[q] = ldg_stream(reinterpret_cast<const float2*>(l0) + n);
                    b[q] = ldg_stream(reinterpret_cast<const float2*>(l1) + n);
                } else {
                    a[q] = make_float2(spim_ldg(l0 + 2 * n), spim_ldg(l0 + 2 * n + 1));
                    b[q] = make_float2(spim_ldg(l1 + 2 * n), spim_ldg(l1 + 2 * n + 1));
                }
            } else {
                const float4 v = xfwd_pair_slow(p, so0, so1, n);
                a[q] = lo2(v); b[q] = hi2(v);
            }
        }
        C2 x[R];
#pragma unroll
        for (int q = 0; q < R; ++q) x[q] = c2_from_ab(a[q], b[q]);
        dft<R, false>(x);
        // twiddles are fetched on use here: holding them across the loads and the butterfly costs the fourth resident block
        apply_twiddles_c2<R, false>(x, twp + m * (R - 1));     // M == 1: a table of ones
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int row = m + q * M;
            tile[row * TP + ((bp + row) & (TP - 1))] = c2_to_pk(x[q]);      // packed tile (stage_tile)
        }
    }
    SPIM_BARRIER();
}

// float2 view of line b in row `row` of a rotated tile
SPIM_HD int xelem(int b, int row) { return row * TC + ((((b >> 1) + row) & (TP - 1)) << 1) + (b & 1); }

// forward split of one line: X[k] = (Zk + conj Zm) - i w (Zk - conj Zm); X[N2-k] = (Zm + conj Zk) + i conj(w) (Zm - conj Zk)
SPIM_DEV void split_fwd(float2 zk, float2 zm, float2 w, float2& xk, float2& xm) {
    const float2 s = make_float2(zk.x + zm.x, zk.y - zm.y);
    const float2 d = make_float2(zk.x - zm.x, zk.y + zm.y);
    const float2 t = cmul(w, d);
    xk = make_float2(s.x + t.y, s.y - t.x);
    // second output: s2 = conj(s), d2 = -conj(d)  ->  t2 = d2 * conj(w) = -conj(d w) = -conj(t)
    xm = make_float2(s.x - t.y, -s.y - t.x);
}
// inverse pre-step: Z'[k] = (A + conj B) + i conj(w)(A - conj B); Z'[N2-k] = (B + conj A) - i w (B - conj A)
SPIM_DEV void split_inv(float2 A, float2 B, float2 w, float2& zk, float2& zm) {
    const float2 s = make_float2(A.x + B.x, A.y - B.y);
    const float2 d = make_float2(A.x - B.x, A.y + B.y);
    const float2 t = cmulc(d, w);
    zk = make_float2(s.x - t.y, s.y + t.x);
    // s2 = conj(s), d2 = -conj(d), t2 = d2 * w = -conj(d conj(w)) = -conj(t):  zm = s2 + (-i) t2... evaluated directly:
    zm = make_float2(s.x + t.y, -s.y + t.x);
}

// the same two steps on a packed pair (both lines of a pair share w)
SPIM_DEV void split_fwd(C2 zk, C2 zm, float2 w, C2& xk, C2& xm) {
    C2 s, d;
    s.re = p2add(zk.re, zm.re); s.im = p2sub(zk.im, zm.im);
    d.re = p2sub(zk.re, zm.re); d.im = p2add(zk.im, zm.im);
    const C2 t = cmul_s(d, w.x, w.y);
    xk.re = p2add(s.re, t.im); xk.im = p2sub(s.im, t.re);
    xm.re = p2sub(s.re, t.im); xm.im = p2sub(p2neg(s.im), t.re);
}
SPIM_DEV void split_inv(C2 A, C2 B, float2 w, C2& zk, C2& zm) {
    C2 s, d;
    s.re = p2add(A.re, B.re); s.im = p2sub(A.im, B.im);
    d.re = p2sub(A.re, B.re); d.im = p2add(A.im, B.im);
    const C2 t = cmulc_s(d, w.x, w.y);
    zk.re = p2sub(s.re, t.im); zk.im = p2add(s.im, t.re);
    zm.re = p2add(s.re, t.im); zm.im = p2sub(t.re, s.im);
}

// forward split step of one tile: half spectrum of 16 real lines from the 8 complex transforms of their pairs (factor
// 1/2 folded into the kernel scale).  One item = one frequency pair (k, N2 - k) of FOUR line pairs, see xinv_presplit.
// dstoff[line * DS]: destination row of a line; DS = 2 (XFwdT): dstoff[line * 2 + 1] = a second row that receives the same
// spectrum (the mirrored halo row) or -1 -- visited only when `anydup` says the tile has one
template <int DS, int W = TP>
SPIM_DEV void xfwd_split(const XFwdParams& p, const float4* tile, const long long* dstoff, int N2, int anydup = 0) {
    const int nk = p.nk;
    SPIM_FOR_ITEMS(i, nk * (W / 4)) {
        const int h = fastdiv(i, p.magic_nk);
        const int k = i - h * nk;
        const int km = N2 - k;
        const float2 w = spim_ldg(p.wx + k);
        const int rk = spim_ldg(p.pos + k);
        const int rm = spim_ldg(p.pos + (k == 0 ? 0 : km));
        const bool two = km != k;
        float4 zk[4], zm[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int bp = h * 4 + g;
            zk[g] = tile[rk * W + xslot<W>(bp, rk)];
            zm[g] = tile[rm * W + xslot<W>(bp, rm)];
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int bp = h * 4 + g;
            const long long d0 = dstoff[(2 * bp) * DS], d1 = dstoff[(2 * bp + 1) * DS];
            C2 xk, xm;
            split_fwd(c2_from_pk(zk[g]), c2_from_pk(zm[g]), w, xk, xm);      // the tile is packed (stage_tile)
            if (d0 >= 0) {
                p.spec[d0 + k] = c2_a(xk);
                if (two) p.spec[d0 + km] = c2_a(xm);
            }
            if (d1 >= 0) {
                p.spec[d1 + k] = c2_b(xk);
                if (two) p.spec[d1 + km] = c2_b(xm);
            }
            if (DS > 1 && anydup) {
                const long long e0 = dstoff[(2 * bp) * DS + 1], e1 = dstoff[(2 * bp + 1) * DS + 1];
                if (e0 >= 0) {
                    p.spec[e0 + k] = c2_a(xk);
                    if (two) p.spec[e0 + km] = c2_a(xm);
                }
                if (e1 >= 0) {
                    p.spec[e1 + k] = c2_b(xk);
                    if (two) p.spec[e1 + km] = c2_b(xm);
                }
            }
        }
    }
}

struct XFwd {
    typedef XFwdParams Params;
    static constexpr bool kEmuThreads = true;
    SPIM_DEV static void run(const Params& p, int bid, float2* tile2) {
        const TG tg = tg_cta();
        float4* tile = reinterpret_cast<float4*>(tile2);
        const FftPlanDev& pl = p.plan;
        const int N2 = pl.n;
        long long* srcoff = reinterpret_cast<long long*>(tile2 + (size_t)N2 * TC);
        long long* dstoff = srcoff + TC;
        // line descriptors for the 16 lines of this tile
        SPIM_FOR_ITEMS(b, TC) {
            const long long l = (long long)bid * TC + b;
            long long so = kLineInvalid, d_o = -1;
            if (l < p.nlines) {
                const int iz = (int)(l / p.LY);
                const int iy = (int)(l - (long long)iz * p.LY);
                const int yp = iy < p.ny + p.hpy ? iy : iy + (p.Py - p.LY);
                const int zp = iz < p.nz + p.hpz ? iz : iz + (p.Pz - p.LZ);
                const int ay = iy < p.ny + p.hpy ? iy : iy - p.LY;
                const int az = iz < p.nz + p.hpz ? iz : iz - p.LZ;
                int jy = ay + p.oy, jz = az + p.oz;
                bool cst = false;
                if ((unsigned)jy >= (unsigned)p.sy || (ay < 0 && !(p.halo_lo & 2)) || (ay >= p.ny && !(p.halo_hi & 2))) {
                    const int e = ext_map(ay, p.ny, p.ext);
                    if (e < 0) cst = true; else jy = e + p.oy;
                }
                if ((unsigned)jz >= (unsigned)p.sz || (az < 0 && !(p.halo_lo & 1)) || (az >= p.nz && !(p.halo_hi & 1))) {
                    const int e = ext_map(az, p.nz, p.ext);
                    if (e < 0) cst = true; else jz = e + p.oz;
                }
                so = cst ? kLineConst : ((long long)jz * p.sy + jy) * (long long)p.sx;
                d_o = ((long long)zp * p.Py + yp) * (long long)p.pitch;
            }
            srcoff[b] = so;
            dstoff[b] = d_o;
        }
        SPIM_BARRIER();
        if (p.src_vec_ok) { SPIM_RADIX_SWITCH(pl.radix[0], (xfwd_stage0<RR, true>(p, tile, srcoff))) }
        else { SPIM_RADIX_SWITCH(pl.radix[0], (xfwd_stage0<RR, false>(p, tile, srcoff))) }
        GRows g;
        g.p = nullptr; g.stride = 0; g.va = g.vb = g.sa = 0; g.dup4 = 0;
        for (int s = 1; s < pl.nstages; ++s) stage_dispatch<false>(tg, pl, s, tile, 1, 0, 0, g, 1, 1);
        xfwd_split<1>(p, tile, dstoff, N2);
        // zero the pad columns [N2+1, pitch)
        const int npad = p.pitch - (N2 + 1);
        SPIM_FOR_ITEMS(i, npad * TC) {
            const int b = i / (npad > 0 ? npad : 1);
            const int j = i - b * npad;
            const long long d_o = dstoff[b];
            if (d_o >= 0) p.spec[d_o + N2 + 1 + j] = make_float2(0.f, 0.f);
        }
    }
};

// ---------------------------------------------------------------------------------------------
// XFwdT: the x-forward pass as a persistent, TMA-fed pipeline (the default wherever the source rows are 16-byte aligned).
//
// A tile's 16 source rows are whole contiguous lines of the real volume, so each is fetched by ONE bulk copy
// (cp.async.bulk, UBLKCP) into a line-major staging slot in shared memory, completion counted on an mbarrier.  The CTA is
// persistent (tiles bid, bid + nctas, ...): while it runs the butterflies of tile i, the rows of tile i + nslot are
// already in flight, so no warp ever waits on a global load, and the first stage is straight-line code -- every operand of
// the padded line comes from the staging slot:
//   * staging index ox + u holds padded position u.  The bulk copy drops the array row at index 0, which puts the image
//     (and, in brick mode, the neighbour-provided halo after it) exactly there;
//   * everything else -- the halo before the image (padded positions [P - hm, P)), mirrored / constant out-of-bounds
//     values, the zero gap -- is patched in by a short fix-up list (dst index, src index | -1 zero | -2 constant) built on
//     the host per geometry, ~3 entries per line-pair item;
//   * lines that are constant by the out-of-bounds rule along y / z are copies of a constant row in global memory, so they
//     take the same path as every other line.
// The transposed tile, the remaining stages and the R2C split step are those of XFwd.
// ---------------------------------------------------------------------------------------------
struct XFwdTParams {
    XFwdParams x;
    int ntiles, nctas, nslot;
    int LS;                    // floats per staging line (multiple of 4, >= max(sx, ox + Px))
    unsigned row_bytes;        // sx * 4, multiple of 16
    const float* const_row;    // sx floats of the out-of-bounds constant (0 for rules that never yield one)
    float cval;                // that constant
    const int2* fix;           // fix-up list: staging[fix.x] = fix.y >= 0 ? staging[fix.y] : (fix.y == -1 ? 0 : cval)
    int nfix;
    uint32_t magic_nfix;       // for item / nfix
    // De-duplicated lines (mirror / periodic extension): the halo lines along y and z are copies of image lines, so only the
    // UY x UZ lines that exist as data are transformed; line index u along an axis has coordinate u < split ? u : u - U
    // (neighbour-provided halo before the image last), and its spectrum is also stored to padded row dupy[u] when a halo row
    // mirrors it.  Mirrored PLANES are left to the forward y pass (ColPassParams::dup_outer).  dedup = 0: every padded line.
    int dedup, UY, UZ, splity, splitz;
    const int* dupy;
    int nfix_dyn;              // the first nfix_dyn entries copy data (mirrored / halo values): applied to every tile.  The rest
    uint32_t magic_nfix_dyn;   // write constants (zero gap, out-of-bounds constant) to cells no row copy ever touches: once per slot
};

// source row of padded line l (nullptr beyond the last line) and its destination offset in the spectrum
SPIM_DEV const float* xfwd_line(const XFwdTParams& q, long long l, long long& d_o, long long& d_dup) {
    const XFwdParams& p = q.x;
    d_o = -1; d_dup = -1;
    if (l >= p.nlines) return nullptr;
    if (q.dedup) {
        const int uz = (int)(l / q.UY);
        const int uy = (int)(l - (long long)uz * q.UY);
        const int ay = uy < q.splity ? uy : uy - q.UY;
        const int az = uz < q.splitz ? uz : uz - q.UZ;
        const int yp = ay >= 0 ? ay : p.Py + ay;
        const int zp = az >= 0 ? az : p.Pz + az;
        d_o = ((long long)zp * p.Py + yp) * (long long)p.pitch;
        const int yd = spim_ldg(q.dupy + uy);
        if (yd >= 0) d_dup = ((long long)zp * p.Py + yd) * (long long)p.pitch;
        return p.src + ((long long)(az + p.oz) * p.sy + (ay + p.oy)) * (long long)p.sx;
    }
    const int iz = (int)(l / p.LY);
    const int iy = (int)(l - (long long)iz * p.LY);
    const int yp = iy < p.ny + p.hpy ? iy : iy + (p.Py - p.LY);
    const int zp = iz < p.nz + p.hpz ? iz : iz + (p.Pz - p.LZ);
    const int ay = iy < p.ny + p.hpy ? iy : iy - p.LY;
    const int az = iz < p.nz + p.hpz ? iz : iz - p.LZ;
    int jy = ay + p.oy, jz = az + p.oz;
    bool cst = false;
    if ((unsigned)jy >= (unsigned)p.sy || (ay < 0 && !(p.halo_lo & 2)) || (ay >= p.ny && !(p.halo_hi & 2))) {
        const int e = ext_map(ay, p.ny, p.ext);
        if (e < 0) cst = true; else jy = e + p.oy;
    }
    if ((unsigned)jz >= (unsigned)p.sz || (az < 0 && !(p.halo_lo & 1)) || (az >= p.nz && !(p.halo_hi & 1))) {
        const int e = ext_map(az, p.nz, p.ext);
        if (e < 0) cst = true; else jz = e + p.oz;
    }
    d_o = ((long long)zp * p.Py + yp) * (long long)p.pitch;
    return cst ? q.const_row : p.src + ((long long)jz * p.sy + jy) * (long long)p.sx;
}

template <int R, int W>
SPIM_DEV void xfwdt_stage0(const XFwdParams& p, float4* tile, const float* stg, int LS) {
    const FftPlanDev& pl = p.plan;
    const int M = pl.M[0];
    const float2* twp = pl.tws + pl.tw_off[0];
    SPIM_FOR_ITEMS(i, M * W) {
        const int bp = (M == 1) ? i : fastdiv(i, p.magic_m0);
        const int m = i - bp * M;
        const float2* l0 = reinterpret_cast<const float2*>(stg + (2 * bp) * LS + p.ox) + m;
        const float2* l1 = reinterpret_cast<const float2*>(stg + (2 * bp + 1) * LS + p.ox) + m;
        float2 a[R], b[R];
#pragma unroll
        for (int q = 0; q < R; ++q) { a[q] = l0[q * M]; b[q] = l1[q * M]; }
        C2 x[R];
#pragma unroll
        for (int q = 0; q < R; ++q) x[q] = c2_from_ab(a[q], b[q]);
        dft<R, false>(x);
        apply_twiddles_c2<R, false>(x, twp + m * (R - 1));     // M == 1: a table of ones
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int row = m + q * M;
            tile[row * W + xslot<W>(bp, row)] = c2_to_pk(x[q]);      // packed tile (stage_tile)
        }
    }
}

template <int W>
struct XFwdTW {
    static constexpr int TCW = 2 * W;           // lines per tile
    typedef XFwdTParams Params;
    static constexpr bool kEmuThreads = true;
    static constexpr int MAXSLOT = 3;
    // descriptors of tile t into slot `slot`, and its 16 row copies.  GPU: the first warp; emulator: every thread its lines.
    SPIM_DEV static void issue(const Params& q, int t, int slot, float* stg, long long* dsto, int* anyd, uint64_t* full) {
#if defined(SPIM_HOST_EMU)
        (void)full;
        SPIM_FOR_ITEMS(b, TCW) {
            long long d_o, d_dup;
            const float* sp = xfwd_line(q, (long long)t * TCW + b, d_o, d_dup);
            dsto[(slot * TCW + b) * 2] = d_o;
            dsto[(slot * TCW + b) * 2 + 1] = d_dup;
            if (sp) memcpy(stg + ((size_t)slot * TCW + b) * q.LS, sp, q.row_bytes);
        }
        if (SPIM_TID == 0) anyd[slot] = q.dedup;
#else
        if (threadIdx.x < 32) {
            const int b = (int)threadIdx.x;
            long long d_o = -1, d_dup = -1;
            const float* sp = nullptr;
            if (b < TCW) {
                sp = xfwd_line(q, (long long)t * TCW + b, d_o, d_dup);
                dsto[(slot * TCW + b) * 2] = d_o;
                dsto[(slot * TCW + b) * 2 + 1] = d_dup;
            }
            const unsigned live = __ballot_sync(0xffffffffu, sp != nullptr);
            const unsigned dups = __ballot_sync(0xffffffffu, d_dup >= 0);
            if (b == 0) { anyd[slot] = dups != 0u; mbar_expect_tx(full + slot, (unsigned)__popc(live) * q.row_bytes); }
            __syncwarp();
            if (sp) bulk_g2s(stg + ((size_t)slot * TCW + b) * q.LS, sp, q.row_bytes, full + slot);
        }
#endif
    }
    SPIM_DEV static void run(const Params& q, int bid, float2* smem2) {
        const XFwdParams& p = q.x;
        const TG tg = tg_cta();
        const FftPlanDev& pl = p.plan;
        const int N2 = pl.n;
        float4* tile = reinterpret_cast<float4*>(smem2);
        float* stg = reinterpret_cast<float*>(smem2 + (size_t)N2 * TCW);
        long long* dsto = reinterpret_cast<long long*>(stg + (size_t)q.nslot * TCW * q.LS);     // [nslot][TCW][2]
        long long* dcur = dsto + MAXSLOT * TCW * 2;                                             // [TCW][2], the tile in work
        uint64_t* full = reinterpret_cast<uint64_t*>(dcur + TCW * 2);                           // [MAXSLOT]
        int* anyd = reinterpret_cast<int*>(full + MAXSLOT);                                    // [MAXSLOT + 1]: tile has duplicate rows
        int2* fixs = reinterpret_cast<int2*>(anyd + MAXSLOT + 1);                              // [nfix], the fix-up list
        const int ntl = (q.ntiles - bid + q.nctas - 1) / q.nctas;
#if !defined(SPIM_HOST_EMU)
        if (threadIdx.x == 0) {
            for (int s = 0; s < q.nslot; ++s) mbar_init(full + s, 1);
            mbar_fence_init();
        }
        __syncthreads();
#endif
        for (int j = 0; j < q.nslot && j < ntl; ++j) issue(q, bid + j * q.nctas, j, stg, dsto, anyd, full);
        SPIM_FOR_ITEMS(j, q.nfix) fixs[j] = spim_ldg(q.fix + j);
        SPIM_BARRIER();
        GRows g;
        g.p = nullptr; g.stride = 0; g.va = g.vb = g.sa = 0; g.dup4 = 0;
        for (int i = 0; i < ntl; ++i) {
            const int slot = i % q.nslot;
            float* sl = stg + (size_t)slot * TCW * q.LS;
#if !defined(SPIM_HOST_EMU)
            mbar_wait(full + slot, (unsigned)((i / q.nslot) & 1));
#endif
            // fix-ups: halo before the image, out-of-bounds values, zero gap (the constant ones only on a slot's first use)
            {
                const int nf = i < q.nslot ? q.nfix : q.nfix_dyn;
                const uint32_t mg = i < q.nslot ? q.magic_nfix : q.magic_nfix_dyn;
                SPIM_FOR_ITEMS(it, TCW * nf) {
                    const int b = nf > 1 ? fastdiv(it, mg) : it;
                    const int j = it - b * nf;
                    const int2 f = fixs[j];
                    float* ln = sl + b * q.LS;
                    ln[f.x] = f.y >= 0 ? ln[f.y] : (f.y == -1 ? 0.f : q.cval);
                }
            }
            SPIM_FOR_ITEMS(b, TCW * 2) dcur[b] = dsto[slot * TCW * 2 + b];
            if (SPIM_TID == 0) anyd[MAXSLOT] = anyd[slot];
            SPIM_BARRIER();
            SPIM_RADIX_SWITCH(pl.radix[0], (xfwdt_stage0<RR, W>(p, tile, sl, q.LS)))
#if !defined(SPIM_HOST_EMU)
            fence_proxy_async();       // this slot's generic-proxy accesses are ordered before the async-proxy refill below
#endif
            SPIM_BARRIER();
            if (i + q.nslot < ntl) issue(q, bid + (i + q.nslot) * q.nctas, slot, stg, dsto, anyd, full);
            for (int s = 1; s < pl.nstages; ++s) stage_dispatch<false, W>(tg, pl, s, tile, 1, 0, 0, g, 1, 1);
            const int anydup = anyd[MAXSLOT];
            xfwd_split<2, W>(p, tile, dcur, N2, anydup);
            const int npad = p.pitch - (N2 + 1);
            SPIM_FOR_ITEMS(k, npad * TCW) {
                const int b = k / (npad > 0 ? npad : 1);
                const int j = k - b * npad;
                const long long d_o = dcur[b * 2], d_d = dcur[b * 2 + 1];
                if (d_o >= 0) p.spec[d_o + N2 + 1 + j] = make_float2(0.f, 0.f);
                if (d_d >= 0) p.spec[d_d + N2 + 1 + j] = make_float2(0.f, 0.f);
            }
            SPIM_BARRIER();            // the tile and dcur are free for the next round
        }
    }
};
typedef XFwdTW<TP> XFwdT;
typedef XFwdTW<TP / 2> XFwdTNarrow;      // 8-line tiles: lines so long (1080 voxels and more) that a 16-line tile leaves one block per SM

// ---------------------------------------------------------------------------------------------
// XInv: half spectrum -> real lines + fused epilogue (specialised at compile time per epilogue mode)
// ---------------------------------------------------------------------------------------------
// Brick mode: the halo push fused into the x-inverse epilogue.  Every voxel a neighbour needs for its halo is stored
// straight into that neighbour's (peer-mapped, NVLink) buffer by the thread that has just computed it, next to the local
// store -- the transfer overlaps the sweep tile by tile and no separate copy pass ever re-reads the faces.  All bricks share
// one geometry, so the copy of local element i at the neighbour in direction (dz, dy, dx) is element i - shift[dir] of its
// buffer, dir = (dz+1)*9 + (dy+1)*3 + (dx+1).  A neighbour below needs my first `whi` cells along that axis (its upper
// halo), a neighbour above my last `wlo` cells.  Flags are raised afterwards by HaloSignalWaitK (spim_b200.cu).
struct HaloFuse {
    float* peer[27];           // base of the neighbour's buffer per direction, nullptr = no neighbour there
    long long shift[27];
    int n[3], wlo[3], whi[3];  // (z, y, x): brick size, halo before / after the brick
    int has_lo, has_hi;        // bit d (0 = z, 1 = y, 2 = x): a neighbour exists below / above along axis d
};
constexpr int kFuseCombos = 4;         // (dz, dy) in {0, dzs} x {0, dys} per line
constexpr int kFuseSlots = 3;          // per combo: whole line, x-low part, x-high part
// bytes of shared memory a FUSE instantiation needs behind its other arrays: [TC][kFuseCombos][kFuseSlots] pointers + [TC] counts
constexpr size_t kFuseSmemBytes = TC * kFuseCombos * kFuseSlots * sizeof(float2*) + TC * sizeof(int);

// where line (z, y) of the brick goes besides the local buffer: for (dz, dy) in {0, dzs} x {0, dys} the whole line (except
// (0, 0)), its first cells for the x neighbour below and its last cells for the one above.  d_o = offset of the line's first
// voxel in the local buffer; f = this line's [kFuseCombos][kFuseSlots] pointers.  Returns the number of combos in use.
SPIM_DEV int fuse_line_targets(const HaloFuse& h, int z, int y, long long d_o, float2** f) {
    const int dzs = (z < h.whi[0] && (h.has_lo & 1)) ? -1 : ((z >= h.n[0] - h.wlo[0] && (h.has_hi & 1)) ? 1 : 0);
    const int dys = (y < h.whi[1] && (h.has_lo & 2)) ? -1 : ((y >= h.n[1] - h.wlo[1] && (h.has_hi & 2)) ? 1 : 0);
    int nc = 0;
    for (int iz = 0; iz < (dzs ? 2 : 1); ++iz)
        for (int iy = 0; iy < (dys ? 2 : 1); ++iy) {
            const int cz = iz ? dzs : 0, cy = iy ? dys : 0;
            const int dir = (cz + 1) * 9 + (cy + 1) * 3 + 1;
            f[0] = (cz || cy) ? reinterpret_cast<float2*>(h.peer[dir] + (d_o - h.shift[dir])) : nullptr;
            f[1] = ((h.has_lo & 4) && h.peer[dir - 1]) ? reinterpret_cast<float2*>(h.peer[dir - 1] + (d_o - h.shift[dir - 1])) : nullptr;
            f[2] = ((h.has_hi & 4) && h.peer[dir + 1]) ? reinterpret_cast<float2*>(h.peer[dir + 1] + (d_o - h.shift[dir + 1])) : nullptr;
            f += kFuseSlots;
            ++nc;
        }
    return nc;
}

struct XInvParams {
    const HaloFuse* fuse;      // device pointer, FUSE instantiations only
    const float2* spec;
    int pitch, Px, Py;
    int nx, ny, nz;            // output region (logical image size)
    FftPlanDev plan;           // n = Px / 2
    const unsigned short* pos;
    const float2* wx;
    uint32_t magic_m0;
    int nk;
    uint32_t magic_nk;
    long long nlines;          // ny * nz
    // destination real array (psi for EPI_UPDATE)
    float* dst;
    int dsx, dsy;              // dst row / plane dims
    int dox, doy, doz;         // dst index = logical coordinate + origin
    int epi;
    const float* img;          // EPI_RATIO: observed view, unpadded [nz][ny][nx]
    const float* weight;       // EPI_UPDATE: per-voxel weight (unpadded) or nullptr
    float const_weight;
    double lambda;
    float two_lambda;          // (float)(2*lambda)
    float min_value;
    int gen2_quotient;         // 1: q = img > 0 ? img / blur : 1 ; 0: q = img / blur
    float ratio_offset;        // EPI_RATIO stores q + ratio_offset (-c when the next convolution extends by the constant c, see engine.h)
    float blur_offset;         // EPI_UPDATE: added to the convolution result before the update (c * sum of the kernel)
    int exact_tikhonov;        // 1: evaluate (sqrt(1+2 lambda v)-1)/lambda in fp64 exactly like the Java code
    int fast_epilogue;         // 1: MATH_FAST division / square root in the ratio and update epilogues
    double* stat_sum;          // EPI_UPDATE statistics (may be nullptr)
    unsigned int* stat_max;    // max |change| as float bits
    int vec_ok;                // float2 accesses to dst / img / weight allowed and nx even
};

struct EpiAcc { double sum; float mx; };

SPIM_DEV float tikhonov_fp64(float value, double lambda) {
    // (float)((Math.sqrt(1.0 + 2.0*lambda*value) - 1.0) / lambda), FD/MVDeconvolution.java:726
#if defined(SPIM_HOST_EMU)
    volatile double t = 2.0 * lambda; t = t * (double)value; t = 1.0 + t;
    volatile double r = sqrt(t) - 1.0; r = r / lambda;
    return (float)r;
#else
    const double t = __dadd_rn(1.0, __dmul_rn(__dmul_rn(2.0, lambda), (double)value));
    return (float)__ddiv_rn(__dadd_rn(__dsqrt_rn(t), -1.0), lambda);
#endif
}

// next psi value, computeNextValue FD/MVDeconvolution.java:692-724 (= D2/BayesMVDeconvolution.java:452-476).
// Default: the algebraically identical cancellation-free fp32 form 2v / (1 + sqrt(1 + 2 lambda v))
// (<= 2 ulp from the fp64 expression for lambda >= 0.006 on v in [1e-5, 10]); EXACT selects the fp64 path.
// Written with selects instead of branches (the Tikhonov value is computed unconditionally and discarded where the
// reference does not use it): the result is the same for every input, NaN included, and the fused epilogue stays
// straight-line code.  `lam_pos` = (lambda > 0), hoisted by the caller.
// MATH: 0 = IEEE fp32 division / square root intrinsics (default), 1 = the reference's fp64 Tikhonov expression,
//       2 = fast epilogue: branch-free refinement of the hardware approximations (fast_math.h; identical results for
//           operands in the normal range)
enum { MATH_IEEE = 0, MATH_EXACT64 = 1, MATH_FAST = 2 };
template <int MATH> SPIM_DEV float epi_div(float a, float b) {
    return MATH == MATH_FAST ? spim_div_from_seed(a, b, spim_rcp_seed(b)) : spim_fdiv_rn(a, b);
}
template <int MATH> SPIM_DEV float epi_sqrt(float x) {
    return MATH == MATH_FAST ? spim_sqrt_from_seed(x, spim_rsqrt_seed(x)) : spim_fsqrt_rn(x);
}

template <int MATH>
SPIM_DEV float next_value(const XInvParams& p, bool lam_pos, float last, float integral) {
    const float value = spim_fmul_rn(last, integral);
    float tik;
    if (MATH == MATH_EXACT64) tik = (value > 0.f && lam_pos) ? tikhonov_fp64(value, p.lambda) : value;
    else tik = epi_div<MATH>(value + value, 1.f + epi_sqrt<MATH>(spim_fmaf_rn(p.two_lambda, value, 1.f)));
    float adj = lam_pos ? tik : value;
    adj = (value > 0.f) ? adj : p.min_value;      // value NaN -> minValue, like the reference's else branch
    return fmaxf(p.min_value, adj);               // fmaxf returns the non-NaN operand: NaN -> minValue
}

struct EpiFlags { bool lam_pos, gen2q, has_w; };

template <int EPI, int MATH>
SPIM_DEV float epi_one(const XInvParams& p, const EpiFlags& f, float v, float x1, float x2, float& csum, float& cmax) {
    if (EPI == EPI_RATIO) {            // x1 = observed image value; gen-2: q = img > 0 ? img / blur : 1
        const float q = epi_div<MATH>(x1, v);
        return ((f.gen2q && !(x1 > 0.f)) ? 1.f : q) + p.ratio_offset;
    }
    if (EPI == EPI_UPDATE) {           // x1 = psi (last), x2 = weight
        const float next = next_value<MATH>(p, f.lam_pos, x1, v + p.blur_offset);
        const float nw = spim_fadd_rn(x1, spim_fmul_rn(spim_fsub_rn(next, x1), x2));
        const float ch = fabsf(spim_fsub_rn(nw, x1));
        csum += ch;
        cmax = fmaxf(cmax, ch);
        return nw;
    }
    return v;
}

// The fast epilogue on the two adjacent voxels of a row pair at once: the refinement sequences of fast_math.h on packed
// pairs (FMUL2 / FFMA2 / FADD2) -- the same operations per component, so the results are bit-identical to epi_one<., MATH_FAST>.
SPIM_DEV float2 p2_div_from_seed(float2 a, float2 b, float2 r0) {
    const float2 nb = p2neg(b);
    const float2 r = p2fma(p2fma(nb, r0, p2bc(1.f)), r0, r0);
    const float2 q0 = p2mul(a, r);
    const float2 q1 = p2fma(p2fma(nb, q0, a), r, q0);
    return p2fma(p2fma(nb, q1, a), r, q1);
}
SPIM_DEV float2 p2_sqrt_from_seed(float2 x, float2 y0) {
    float2 g = p2mul(x, y0), h = p2mul(p2bc(0.5f), y0);
    const float2 r = p2fma(p2neg(h), g, p2bc(0.5f));
    g = p2fma(g, r, g);
    h = p2fma(h, r, h);
    const float2 d = p2fma(p2neg(g), g, x);
    return p2fma(d, h, g);
}
// v = blurred values, x1 = image (ratio) / psi (update), x2 = weights; s0, s1 = |change| of the two voxels, mx = their maximum
template <int EPI>
SPIM_DEV float2 epi_pair_fast(const XInvParams& p, const EpiFlags& f, float2 v, float2 x1, float2 x2, float& s0, float& s1, float& mx) {
    if (EPI == EPI_RATIO) {
        const float2 q = p2_div_from_seed(x1, v, make_float2(spim_rcp_seed(v.x), spim_rcp_seed(v.y)));
        return p2add(make_float2((f.gen2q && !(x1.x > 0.f)) ? 1.f : q.x, (f.gen2q && !(x1.y > 0.f)) ? 1.f : q.y), p2bc(p.ratio_offset));
    }
    if (EPI == EPI_UPDATE) {
        const float2 value = p2mul(x1, p2add(v, p2bc(p.blur_offset)));
        const float2 t = p2fma(p2bc(p.two_lambda), value, p2bc(1.f));
        const float2 sq = p2_sqrt_from_seed(t, make_float2(spim_rsqrt_seed(t.x), spim_rsqrt_seed(t.y)));
        const float2 den = p2add(p2bc(1.f), sq);
        const float2 tik = p2_div_from_seed(p2add(value, value), den, make_float2(spim_rcp_seed(den.x), spim_rcp_seed(den.y)));
        float ax = f.lam_pos ? tik.x : value.x, ay = f.lam_pos ? tik.y : value.y;
        ax = (value.x > 0.f) ? ax : p.min_value;
        ay = (value.y > 0.f) ? ay : p.min_value;
        const float2 next = make_float2(fmaxf(p.min_value, ax), fmaxf(p.min_value, ay));
        const float2 nw = p2add(x1, p2mul(p2sub(next, x1), x2));
        const float2 ch = p2sub(nw, x1);
        s0 = fabsf(ch.x); s1 = fabsf(ch.y);
        mx = fmaxf(s0, s1);
        return nw;
    }
    return v;
}

// epilogue inputs for output samples x = 2n, 2n+1 of one line, fetched BEFORE the butterfly so that the
// global-memory latency overlaps the shared-memory stage (scalar path: odd nx / unaligned buffers)
template <int EPI>
SPIM_DEV void epi_fetch_scalar(const XInvParams& p, long long aux0, long long dst0, int n, float2& x1, float2& x2) {
    const int u0 = 2 * n;
    x1 = make_float2(0.f, 0.f);
    x2 = make_float2(p.const_weight, p.const_weight);
    if (EPI == EPI_STORE || u0 >= p.nx) return;
    const long long ai = aux0 + u0, di = dst0 + u0;
    const bool two = (u0 + 1 < p.nx);
    if (EPI == EPI_RATIO) { x1.x = spim_ldg(p.img + ai); if (two) x1.y = spim_ldg(p.img + ai + 1); }
    else {
        if (p.weight) { x2.x = spim_ldg(p.weight + ai); if (two) x2.y = spim_ldg(p.weight + ai + 1); }
        x1.x = p.dst[di]; if (two) x1.y = p.dst[di + 1];
    }
}

template <int EPI, int MATH>
SPIM_DEV void epi_store_scalar(const XInvParams& p, const EpiFlags& f, long long dst0, int n, float2 v, float2 x1, float2 x2,
                               float& csum, float& cmax) {
    const int u0 = 2 * n;
    if (u0 >= p.nx) return;
    const long long di = dst0 + u0;
    p.dst[di] = epi_one<EPI, MATH>(p, f, v.x, x1.x, x2.x, csum, cmax);
    if (u0 + 1 < p.nx) p.dst[di + 1] = epi_one<EPI, MATH>(p, f, v.y, x1.y, x2.y, csum, cmax);
}

template <int R, int EPI, int MATH, bool VEC, bool FUSE = false>
SPIM_DEV void xinv_stage0(const XInvParams& p, float2* tile2, const long long* auxoff, const long long* dstoff, EpiAcc& acc,
                          float2* const* fptr = nullptr, const int* fnc = nullptr, int nlo2 = 0, int nhi2 = 0) {
    const FftPlanDev& pl = p.plan;
    const int M = pl.M[0];
    const float2* twp = pl.tws + pl.tw_off[0];
    EpiFlags f;
    f.lam_pos = p.lambda > 0.0; f.gen2q = p.gen2_quotient != 0; f.has_w = p.weight != nullptr;
    SPIM_FOR_ITEMS(i, M * TC) {
        const int b = (M == 1) ? i : fastdiv(i, p.magic_m0);
        const int m = i - b * M;
        const long long d_o = dstoff[b];
        if (d_o < 0) continue;
        const long long a_o = auxoff[b];
        float2 x1[R], x2[R], w[R];
        load_twiddles<R>(w, twp + m * (R - 1));     // single-stage plans (M == 1) have a table of ones: no conditional definition
        if (VEC) {
            // 8-byte accesses (nx even, aligned buffers): one base pointer per array and item, predicated accesses per row
            // pair instead of early exits, the weight's null check once per item
            const float2* in1 = reinterpret_cast<const float2*>(EPI == EPI_RATIO ? p.img + a_o : p.dst + d_o);
            const float2* in2 = reinterpret_cast<const float2*>(p.weight + a_o);
            const int nh = p.nx >> 1;
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int n = m + q * M;
                x1[q] = make_float2(0.f, 0.f);
                x2[q] = make_float2(p.const_weight, p.const_weight);
                if (EPI != EPI_STORE && n < nh) x1[q] = (EPI == EPI_RATIO) ? ldg_stream(in1 + n) : in1[n];
            }
            if (EPI == EPI_UPDATE && f.has_w) {
#pragma unroll
                for (int q = 0; q < R; ++q) { const int n = m + q * M; if (n < nh) x2[q] = ldg_stream(in2 + n); }
            }
        } else {
#pragma unroll
            for (int q = 0; q < R; ++q) epi_fetch_scalar<EPI>(p, a_o, d_o, m + q * M, x1[q], x2[q]);
        }
        float2 a[R];
#pragma unroll
        for (int q = 0; q < R; ++q) a[q] = tile2[xelem(b, m + q * M)];
        mul_twiddles1<R, true>(a, w);
        dft<R, true>(a);
        float csum = 0.f, cmax = 0.f;
        if (VEC) {
            float2* out = reinterpret_cast<float2*>(p.dst + d_o);
            const int nh = p.nx >> 1;
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int n = m + q * M;
                // rows beyond the image take part in the arithmetic with benign inputs and are simply not stored / counted
                float s0 = 0.f, s1 = 0.f, mx = 0.f;
                float2 v2;
                if (MATH == MATH_FAST && EPI != EPI_STORE) v2 = epi_pair_fast<EPI>(p, f, a[q], x1[q], x2[q], s0, s1, mx);
                else {
                    v2.x = epi_one<EPI, MATH>(p, f, a[q].x, x1[q].x, x2[q].x, s0, mx);
                    v2.y = epi_one<EPI, MATH>(p, f, a[q].y, x1[q].y, x2[q].y, s1, mx);
                }
                if (n < nh) {
                    out[n] = v2;
                    csum += s0; csum += s1;
                    cmax = fmaxf(cmax, mx);
                    if (FUSE) {
                        // combo 0 = this line at the x neighbours only; further combos (lines inside a y / z face region) also
                        // send the whole line.  Pointers are pre-offset per line, so the pair index n applies unchanged.
                        float2* const* fp = fptr + b * (kFuseCombos * kFuseSlots);
                        const int nc = fnc[b];
                        for (int c = 0; c < nc; ++c, fp += kFuseSlots) {
                            if (fp[0]) fp[0][n] = v2;
                            if (n < nlo2 && fp[1]) fp[1][n] = v2;
                            if (n >= nhi2 && fp[2]) fp[2][n] = v2;
                        }
                    }
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < R; ++q) epi_store_scalar<EPI, MATH>(p, f, d_o, m + q * M, a[q], x1[q], x2[q], csum, cmax);
        }
        if (EPI == EPI_UPDATE) { acc.sum += (double)csum; acc.mx = fmaxf(acc.mx, cmax); }
    }
}

SPIM_DEV void stats_commit(const XInvParams& p, float2* tile, EpiAcc& acc) {
    if (p.stat_sum == nullptr) return;
#if defined(SPIM_HOST_EMU)
    (void)tile;
    static std::mutex stat_mutex;          // the emulator's stand-in for the two atomics (blocks may run as several threads)
    std::lock_guard<std::mutex> lock(stat_mutex);
    *p.stat_sum += acc.sum;
    unsigned int bits; memcpy(&bits, &acc.mx, 4);
    if (bits > *p.stat_max) *p.stat_max = bits;
#else
    double s = acc.sum; float m = acc.mx;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    __syncthreads();   // tile no longer in use
    double* ssum = reinterpret_cast<double*>(tile);
    float* smax = reinterpret_cast<float*>(ssum + 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { ssum[warp] = s; smax[warp] = m; }
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        s = lane < nw ? ssum[lane] : 0.0;
        m = lane < nw ? smax[lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        }
        if (lane == 0) {
            atomicAdd(p.stat_sum, s);
            atomicMax(p.stat_max, __float_as_uint(m));
        }
    }
#endif
}

// inverse pre-step of one tile: the 8 complex sequences of the 16 lines' pairs from their half spectra.
// One item = one frequency pair (k, N2 - k) of FOUR line pairs: the table lookups (pos[], twiddle) and the index split
// are paid once per item, and 16 spectrum loads per thread are in flight before the first is consumed.
constexpr int XG = 4;     // line pairs per split-step item
// dst_p: the tile is written packed (further stages follow) or interleaved (single-stage plan: stage 0 reads it directly)
SPIM_DEV void xinv_presplit(const XInvParams& p, float4* tile, const long long* srcoff, int N2, int dst_p) {
    const int nk = p.nk;
    SPIM_FOR_ITEMS(i, nk * (TP / XG)) {
        const int h = fastdiv(i, p.magic_nk);
        const int k = i - h * nk;
        const int km = N2 - k;
        const float2 w = spim_ldg(p.wx + k);
        const int rk = spim_ldg(p.pos + k);
        const int rm = spim_ldg(p.pos + (k == 0 ? 0 : km));
        const bool two = (k != 0) && (km != k);
        float2 A0[XG], B0[XG], A1[XG], B1[XG];
#pragma unroll
        for (int g = 0; g < XG; ++g) {
            const int bp = h * XG + g;
            const long long s0 = srcoff[2 * bp], s1 = srcoff[2 * bp + 1];
            A0[g] = B0[g] = A1[g] = B1[g] = make_float2(0.f, 0.f);
            if (s0 >= 0) { A0[g] = ldg_stream(p.spec + s0 + k); B0[g] = ldg_stream(p.spec + s0 + km); }
            if (s1 >= 0) { A1[g] = ldg_stream(p.spec + s1 + k); B1[g] = ldg_stream(p.spec + s1 + km); }
        }
#pragma unroll
        for (int g = 0; g < XG; ++g) {
            const int bp = h * XG + g;
            C2 zk, zm;
            split_inv(c2_from_ab(A0[g], A1[g]), c2_from_ab(B0[g], B1[g]), w, zk, zm);
            tile[rk * TP + ((bp + rk) & (TP - 1))] = dst_p ? c2_to_pk(zk) : c2_to_il(zk);
            if (two) tile[rm * TP + ((bp + rm) & (TP - 1))] = dst_p ? c2_to_pk(zm) : c2_to_il(zm);
        }
    }
}

// FUSE: brick mode with connected peers -- the epilogue also stores every voxel a neighbour's halo needs straight into that
// neighbour's buffer (HaloFuse above); the host selects it only when the vectorised epilogue applies.
template <int EPI, int MATH, bool FUSE = false>
struct XInvT {
    typedef XInvParams Params;
    static constexpr bool kEmuThreads = true;
    SPIM_DEV static void run(const Params& p, int bid, float2* tile2) {
        const TG tg = tg_cta();
        float4* tile = reinterpret_cast<float4*>(tile2);
        const FftPlanDev& pl = p.plan;
        const int N2 = pl.n;
        long long* srcoff = reinterpret_cast<long long*>(tile2 + (size_t)N2 * TC);
        long long* dstoff = srcoff + TC;
        long long* auxoff = dstoff + TC;
        float2** fptr = reinterpret_cast<float2**>(auxoff + TC);                   // FUSE: [TC][kFuseCombos][kFuseSlots]
        int* fnc = reinterpret_cast<int*>(fptr + TC * kFuseCombos * kFuseSlots);   // FUSE: combos in use per line
        SPIM_FOR_ITEMS(b, TC) {
            const long long l = (long long)bid * TC + b;
            long long so = -1, d_o = -1, a_o = -1;
            int nc = 0;
            if (l < p.nlines) {
                const int z = (int)(l / p.ny);
                const int y = (int)(l - (long long)z * p.ny);
                so = ((long long)z * p.Py + y) * (long long)p.pitch;
                d_o = ((long long)(z + p.doz) * p.dsy + (y + p.doy)) * (long long)p.dsx + p.dox;
                a_o = ((long long)z * p.ny + y) * (long long)p.nx;
                if (FUSE) nc = fuse_line_targets(*p.fuse, z, y, d_o, fptr + b * kFuseCombos * kFuseSlots);
            }
            srcoff[b] = so; dstoff[b] = d_o; auxoff[b] = a_o;
            if (FUSE) fnc[b] = nc;
        }
        SPIM_BARRIER();
        xinv_presplit(p, tile, srcoff, N2, pl.nstages > 1);
        SPIM_BARRIER();
        GRows g;
        g.p = nullptr; g.stride = 0; g.va = g.vb = g.sa = 0; g.dup4 = 0;
        for (int s = pl.nstages - 1; s >= 1; --s) stage_dispatch<true>(tg, pl, s, tile, 1, 0, 0, g, 1, s > 1);   // stage 0 reads single lines: interleaved
        EpiAcc acc;
        acc.sum = 0.0; acc.mx = 0.f;
        if (FUSE) {
            const int nlo2 = (p.fuse->whi[2] + 1) >> 1;              // pairs (2n, 2n+1) that reach into x < whi
            const int nhi2 = (p.nx - p.fuse->wlo[2]) >> 1;           // ... into x >= nx - wlo
            SPIM_RADIX_SWITCH(pl.radix[0], (xinv_stage0<RR, EPI, MATH, true, true>(p, tile2, auxoff, dstoff, acc, fptr, fnc, nlo2, nhi2)))
        }
        else if (p.vec_ok) { SPIM_RADIX_SWITCH(pl.radix[0], (xinv_stage0<RR, EPI, MATH, true>(p, tile2, auxoff, dstoff, acc))) }
        else { SPIM_RADIX_SWITCH(pl.radix[0], (xinv_stage0<RR, EPI, MATH, false>(p, tile2, auxoff, dstoff, acc))) }
        if (EPI == EPI_UPDATE) stats_commit(p, tile2, acc);
    }
};

// ---------------------------------------------------------------------------------------------
// element-wise helpers (one-off initialisation work, a6-a9 of SURVEY section 8)
// ---------------------------------------------------------------------------------------------
constexpr int MAX_VIEWS = 64;

struct FillParams { float* p; long long n; float v; };
struct ScaleClampParams { float* p; long long n; float f; };           // w <- min(1, w*f)   (OSEM)
struct ViewPtrs { const float* img[MAX_VIEWS]; const float* w[MAX_VIEWS]; int nviews; };
struct InitStatsParams {     // psi initialisation statistics
    ViewPtrs v; long long n; int gen;
    double* sum;             // gen-2: sum over voxels of mean over views with img>0 ; gen-1: sum of intensities where >= 2 views
    unsigned long long* cnt; // [0] gen-2: #voxels with data ; gen-1: sum of counts where >= 2 views
                             // [1] gen-1: sum of counts where >= 1 view ; [2] gen-1: #voxels with >= 1 view
    unsigned int* min_overlap;
};
struct MaskParams { ViewPtrs v; float* psi; long long n; };            // psi <- 0 where no view has img > 0

}  // namespace spim
