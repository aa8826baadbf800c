// Fusion pre-step kernels (the callers on the input side of the hot path):
//   ResampleK     affine resampling of a raw stack into the bounding box + cosine blending weight
//                 (FD/TransformInput.java:70-116, FD/TransformInputAndWeights.java:76-135,
//                  FD/TransformWeights.java:73-111, FD/ExtractPSF.java:417-457)
//   ExtractPsfK   bead-averaged PSF extraction (FD/ExtractPSF.java:374-415)
//   WeightNormK   sum-of-weights normalisation + overlap statistics (FD/WeightNormalizer.java:132-253,
//                 FW/NormalizingRandomAccess.java:57-66)
//   MinMaxK / NormalizeK   loader normalisation (spim/fiji/spimdata/imgloaders/AbstractImgLoader.java:164-184)
// FD/ = spim/process/fusion/deconvolution/, FW/ = spim/process/fusion/weights/ under /root/reference/src/main/java/.
//
// Streaming work: one coalesced write per output voxel, the gathers of the tri-linear taps walk a straight line
// through the source stack and are served by L2.  The Java arithmetic is reproduced operation by operation (double
// position math and interpolation weights, every tap rounded to float, float accumulation in the interpolator's
// Gray-code order, no FMA contraction anywhere), which makes ResampleK bound by fp64 / integer issue (~100 un-fused
// double operations per voxel) rather than by its ~10 bytes per voxel of HBM traffic; WeightNormK and the loader
// normalisation are plain HBM-bound passes.
// Included by spim_b200.cu after its element-wise helpers (atomic_* / warp_* shims, kChunk, ew_blocks).
#pragma once

namespace spim {

// out-of-bounds rule on 64-bit coordinates (positions come from an arbitrary affine)
SPIM_HD int ext_map64(long long a, int n, int mode) {
    if (a >= 0 && a < (long long)n) return (int)a;
    switch (mode) {
        case EXT_PERIODIC: {
            long long m = a % n;
            return (int)(m < 0 ? m + n : m);
        }
        case EXT_MIRROR_SINGLE: {
            if (n == 1) return 0;
            const long long p = 2LL * (n - 1);
            long long m = a % p;
            if (m < 0) m += p;
            return (int)(m < n ? m : p - m);
        }
        case EXT_MIRROR_DOUBLE: {
            const long long p = 2LL * n;
            long long m = a % p;
            if (m < 0) m += p;
            return (int)(m < n ? m : p - 1 - m);
        }
        default: return -1;
    }
}

struct SrcVol { const float* p; int sx, sy, sz; };

SPIM_DEV float sample_ext(const SrcVol& s, long long ix, long long iy, long long iz, int ext, float cvalue) {
    const int jx = ext_map64(ix, s.sx, ext), jy = ext_map64(iy, s.sy, ext), jz = ext_map64(iz, s.sz, ext);
    if ((jx | jy | jz) < 0) return cvalue;
    return spim_ldg(s.p + ((long long)jz * s.sy + jy) * (long long)s.sx + jx);
}

// net.imglib2 NLinearInterpolator3D.get() on FloatType: weights in double, every tap rounded to float by
// FloatType.mul(double), accumulated in float in the order 000,100,110,010,011,111,101,001 (x = first bit)
SPIM_DEV float nlinear3d(const SrcVol& s, double px, double py, double pz, int ext, float cvalue) {
    if (!(fabs(px) < 1e15 && fabs(py) < 1e15 && fabs(pz) < 1e15)) return cvalue;   // NaN / absurd positions
    const double fx = floor(px), fy = floor(py), fz = floor(pz);
    const double w0 = spim_dsub_rn(px, fx), w1 = spim_dsub_rn(py, fy), w2 = spim_dsub_rn(pz, fz);
    const double w0n = spim_dsub_rn(1.0, w0), w1n = spim_dsub_rn(1.0, w1), w2n = spim_dsub_rn(1.0, w2);
    const long long ix = (long long)fx, iy = (long long)fy, iz = (long long)fz;
    float v[8];   // Gray-code order
    if (ix >= 0 && iy >= 0 && iz >= 0 && ix + 1 < s.sx && iy + 1 < s.sy && iz + 1 < s.sz) {
        const float* q = s.p + ((long long)iz * s.sy + iy) * (long long)s.sx + ix;
        const long long ys = s.sx, zs = (long long)s.sx * s.sy;
        v[0] = spim_ldg(q);            v[1] = spim_ldg(q + 1);
        v[2] = spim_ldg(q + ys + 1);   v[3] = spim_ldg(q + ys);
        v[4] = spim_ldg(q + zs + ys);  v[5] = spim_ldg(q + zs + ys + 1);
        v[6] = spim_ldg(q + zs + 1);   v[7] = spim_ldg(q + zs);
    } else {
        v[0] = sample_ext(s, ix, iy, iz, ext, cvalue);
        v[1] = sample_ext(s, ix + 1, iy, iz, ext, cvalue);
        v[2] = sample_ext(s, ix + 1, iy + 1, iz, ext, cvalue);
        v[3] = sample_ext(s, ix, iy + 1, iz, ext, cvalue);
        v[4] = sample_ext(s, ix, iy + 1, iz + 1, ext, cvalue);
        v[5] = sample_ext(s, ix + 1, iy + 1, iz + 1, ext, cvalue);
        v[6] = sample_ext(s, ix + 1, iy, iz + 1, ext, cvalue);
        v[7] = sample_ext(s, ix, iy, iz + 1, ext, cvalue);
    }
    const double c00 = spim_dmul_rn(w0n, w1n), c10 = spim_dmul_rn(w0, w1n), c11 = spim_dmul_rn(w0, w1), c01 = spim_dmul_rn(w0n, w1);
    float acc = spim_d2f_rn(spim_dmul_rn((double)v[0], spim_dmul_rn(c00, w2n)));
    acc = spim_fadd_rn(acc, spim_d2f_rn(spim_dmul_rn((double)v[1], spim_dmul_rn(c10, w2n))));
    acc = spim_fadd_rn(acc, spim_d2f_rn(spim_dmul_rn((double)v[2], spim_dmul_rn(c11, w2n))));
    acc = spim_fadd_rn(acc, spim_d2f_rn(spim_dmul_rn((double)v[3], spim_dmul_rn(c01, w2n))));
    acc = spim_fadd_rn(acc, spim_d2f_rn(spim_dmul_rn((double)v[4], spim_dmul_rn(c01, w2))));
    acc = spim_fadd_rn(acc, spim_d2f_rn(spim_dmul_rn((double)v[5], spim_dmul_rn(c11, w2))));
    acc = spim_fadd_rn(acc, spim_d2f_rn(spim_dmul_rn((double)v[6], spim_dmul_rn(c10, w2))));
    acc = spim_fadd_rn(acc, spim_d2f_rn(spim_dmul_rn((double)v[7], spim_dmul_rn(c00, w2))));
    return acc;
}

// AffineTransform3D.apply: ((x*m0 + y*m1) + z*m2) + m3 in double, row-packed (x, y, z) matrix
struct Affine12 { double m[12]; };
SPIM_DEV void affine_apply(const Affine12& a, double x, double y, double z, double& t0, double& t1, double& t2) {
    t0 = spim_dadd_rn(spim_dadd_rn(spim_dadd_rn(spim_dmul_rn(x, a.m[0]), spim_dmul_rn(y, a.m[1])), spim_dmul_rn(z, a.m[2])), a.m[3]);
    t1 = spim_dadd_rn(spim_dadd_rn(spim_dadd_rn(spim_dmul_rn(x, a.m[4]), spim_dmul_rn(y, a.m[5])), spim_dmul_rn(z, a.m[6])), a.m[7]);
    t2 = spim_dadd_rn(spim_dadd_rn(spim_dadd_rn(spim_dmul_rn(x, a.m[8]), spim_dmul_rn(y, a.m[9])), spim_dmul_rn(z, a.m[10])), a.m[11]);
}

// computeWeight, FW/BlendingRealRandomAccess.java:91-121.  Coordinates in (x, y, z) order.
struct BlendDesc {
    float border[3], blending[3];
    int imin[3], dim_minus1[3];
    const double* lut;          // lookUp[1001], FW/BlendingRealRandomAccess.java:44-54
};
SPIM_DEV float blend_weight(const BlendDesc& b, float t0, float t1, float t2) {
    const float loc[3] = {t0, t1, t2};
    float md = 1.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float l = spim_fsub_rn(loc[d], (float)b.imin[d]);
        const float a1 = spim_fsub_rn(l, b.border[d]);
        const float a2 = spim_fsub_rn(spim_fsub_rn((float)b.dim_minus1[d], l), b.border[d]);
        const float dist = fmaxf(0.f, fminf(a1, a2));
        if (dist == 0.f) return 0.f;
        const float rel = spim_fdiv_rn(dist, b.blending[d]);
        if (rel < 1.f) {
            int idx = (int)floor(spim_dadd_rn(spim_dmul_rn((double)rel, 1000.0), 0.5));   // Math.round
            idx = idx < 0 ? 0 : (idx > 1000 ? 1000 : idx);
            md = spim_d2f_rn(spim_dmul_rn((double)md, spim_ldg(b.lut + idx)));            // float *= double
        }
    }
    return md;
}

// (x, y, z) of a linear voxel index without a 64-bit division per voxel: the items a thread visits inside one chunk
// are `step` apart, so the coordinates are decomposed once per chunk and then advanced with carries.
struct VoxelWalker {
    int x, y, z, nx, ny;
    SPIM_DEV void start(long long idx, int nx_, int ny_) {
        nx = nx_; ny = ny_;
        x = (int)(idx % nx);
        const long long r = idx / nx;
        y = (int)(r % ny);
        z = (int)(r / ny);
    }
    SPIM_DEV void advance(int step) {
        x += step;
        while (x >= nx) { x -= nx; ++y; }
        while (y >= ny) { y -= ny; ++z; }
    }
};
#if defined(SPIM_HOST_EMU)
#define SPIM_ITEM_STEP spim_emu_nthr
#else
#define SPIM_ITEM_STEP ((int)blockDim.x)
#endif

// ---------------------------------------------------------------------------------------------
// ResampleK: one output voxel per item, x fastest (coalesced stores).
//   pos_mode 0  TransformInput arithmetic: s = (float)voxel + (float)offset, t = (float)(inverse * s)
//   pos_mode 1  ExtractPSF.transform arithmetic: everything in double, offset is a double
// ---------------------------------------------------------------------------------------------
struct ResampleParams {
    SrcVol src;
    Affine12 inv;
    double off[3];              // (x, y, z)
    int pos_mode;
    int ext; float ext_value;
    int on[3];                  // output dims (z, y, x)
    float* out_img;             // nullptr: no image (WEIGHTS_ONLY)
    float* out_w;               // nullptr: no weight
    int clamp_inside;           // 1: value = max(min_value, v) where the position intersects the stack, 0 elsewhere
    float min_value;
    BlendDesc blend;
    int nblocks;
};

struct ResampleK {
    typedef ResampleParams Params;
    SPIM_DEV static void run(const Params& p, int bid, float2*) {
        const long long total = (long long)p.on[0] * p.on[1] * p.on[2];
        for (long long base = (long long)bid * kChunk; base < total; base += (long long)p.nblocks * kChunk) {
            VoxelWalker w;
            w.start(base + SPIM_TID, p.on[2], p.on[1]);
            SPIM_FOR_ITEMS(i, kChunk) {
                const long long idx = base + i;
                const int x = w.x, y = w.y, z = w.z;
                w.advance(SPIM_ITEM_STEP);
                if (idx >= total) continue;
                double t0, t1, t2;
                if (p.pos_mode == 0) {
                    const float s0 = spim_fadd_rn((float)x, (float)p.off[0]);
                    const float s1 = spim_fadd_rn((float)y, (float)p.off[1]);
                    const float s2 = spim_fadd_rn((float)z, (float)p.off[2]);
                    affine_apply(p.inv, (double)s0, (double)s1, (double)s2, t0, t1, t2);
                    t0 = (double)spim_d2f_rn(t0); t1 = (double)spim_d2f_rn(t1); t2 = (double)spim_d2f_rn(t2);
                } else {
                    affine_apply(p.inv, spim_dadd_rn((double)x, p.off[0]), spim_dadd_rn((double)y, p.off[1]),
                                 spim_dadd_rn((double)z, p.off[2]), t0, t1, t2);
                }
                if (p.out_img) {
                    float v = 0.f;
                    if (p.clamp_inside) {
                        // FusionHelper.intersects + Math.max(minValue, v): NaN propagates in Java
                        if (t0 >= 0 && t1 >= 0 && t2 >= 0 && t0 < p.src.sx && t1 < p.src.sy && t2 < p.src.sz) {
                            const float q = nlinear3d(p.src, t0, t1, t2, p.ext, p.ext_value);
                            v = (q != q) ? q : fmaxf(p.min_value, q);
                        }
                    } else {
                        v = nlinear3d(p.src, t0, t1, t2, p.ext, p.ext_value);
                    }
                    p.out_img[idx] = v;
                }
                if (p.out_w) p.out_w[idx] = blend_weight(p.blend, (float)t0, (float)t1, (float)t2);
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// ExtractPsfK: psf[v] = sum over beads (in list order, float accumulation) of the tri-linear sample at
// voxel - size/2 + bead position, periodic extension.  One PSF voxel per item.
// ---------------------------------------------------------------------------------------------
struct ExtractPsfParams {
    SrcVol src;
    const double* loc;          // n_beads x (x, y, z)
    int n_beads;
    int size[3];                // (z, y, x)
    float* out;
    int nblocks;
};
struct ExtractPsfK {
    typedef ExtractPsfParams Params;
    SPIM_DEV static void run(const Params& p, int bid, float2*) {
        const long long total = (long long)p.size[0] * p.size[1] * p.size[2];
        for (long long base = (long long)bid * kChunk; base < total; base += (long long)p.nblocks * kChunk) {
            VoxelWalker w;
            w.start(base + SPIM_TID, p.size[2], p.size[1]);
            SPIM_FOR_ITEMS(i, kChunk) {
                const long long idx = base + i;
                const int x = w.x, y = w.y, z = w.z;
                w.advance(SPIM_ITEM_STEP);
                if (idx >= total) continue;
                const double dx = (double)(x - p.size[2] / 2), dy = (double)(y - p.size[1] / 2), dz = (double)(z - p.size[0] / 2);
                float acc = 0.f;
                for (int b = 0; b < p.n_beads; ++b) {
                    const double px = spim_dadd_rn(dx, spim_ldg(p.loc + 3 * b));
                    const double py = spim_dadd_rn(dy, spim_ldg(p.loc + 3 * b + 1));
                    const double pz = spim_dadd_rn(dz, spim_ldg(p.loc + 3 * b + 2));
                    acc = spim_fadd_rn(acc, nlinear3d(p.src, px, py, pz, EXT_PERIODIC, 0.f));
                }
                p.out[idx] = acc;
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// WeightNormK
//   mode 0 (ApplyDirectly)   w_v <- (float)(w_v / sumW) for every voxel, sumW = double sum over views in order
//   mode 1 (ComputeSumImage) sumw <- sumW > 1 ? (float)sumW : 1
//   mode 2 (NormalizingRandomAccess) w_v <- (float)min(1, (w_v / sumw) * osem) in double
// Modes 0 and 1 also count the views with w > 0 per voxel; the counts are kept per ImagePortion
// (F/FusionHelper.java:257-280) in a shared-memory histogram so that the host can form the reference's
// "mean of the portion means" exactly from integers.
// ---------------------------------------------------------------------------------------------
struct WeightPtrs { float* w[MAX_VIEWS]; int nviews; };
struct WeightNormParams {
    WeightPtrs v;
    long long n;
    int mode;
    float* sumw;                        // mode 1: out, mode 2: in
    double osem;
    int nportions; long long chunk;     // portion of voxel idx = min(idx / chunk, nportions - 1)
    unsigned long long* cnt;            // [nportions] sum of counts
    unsigned int* pmin;                 // [nportions] minimum count
    int nblocks;
};
struct WeightNormK {
    typedef WeightNormParams Params;
    SPIM_DEV static void flush(unsigned long long* scnt, unsigned int* smin, int portion, unsigned long long c, unsigned int mn) {
        if (portion < 0) return;
#if defined(SPIM_HOST_EMU)
        std::lock_guard<std::mutex> l(g_emu_atomic_mutex);
        scnt[portion] += c;
        if (mn < smin[portion]) smin[portion] = mn;
#else
        atomicAdd(scnt + portion, c);
        atomicMin(smin + portion, mn);
#endif
    }
    SPIM_DEV static void run(const Params& p, int bid, float2* smem) {
        unsigned long long* scnt = reinterpret_cast<unsigned long long*>(smem);
        unsigned int* smin = reinterpret_cast<unsigned int*>(scnt + p.nportions);
        const bool stats = p.mode != 2;
        if (stats) {
            SPIM_FOR_ITEMS(i, p.nportions) { scnt[i] = 0ull; smin[i] = 0xffffffffu; }
            SPIM_BARRIER();
        }
        int cur = -1;
        unsigned long long c = 0;
        unsigned int mn = 0xffffffffu;
        for (long long base = (long long)bid * kChunk; base < p.n; base += (long long)p.nblocks * kChunk) {
            SPIM_FOR_ITEMS(i, kChunk) {
                const long long idx = base + i;
                if (idx >= p.n) continue;
                if (p.mode == 2) {
                    const double s = (double)p.sumw[idx];
                    for (int v = 0; v < p.v.nviews; ++v) {
                        const double r = spim_dmul_rn(spim_ddiv_rn((double)p.v.w[v][idx], s), p.osem);
                        p.v.w[v][idx] = spim_d2f_rn((r < 1.0 || r != r) ? r : 1.0);   // Math.min(1, r) keeps NaN
                    }
                    continue;
                }
                double sum = 0.0;
                unsigned int count = 0;
                for (int v = 0; v < p.v.nviews; ++v) {
                    const float w = p.v.w[v][idx];
                    sum = spim_dadd_rn(sum, (double)w);
                    if (w > 0.f) ++count;
                }
                if (p.mode == 0) {
                    for (int v = 0; v < p.v.nviews; ++v) p.v.w[v][idx] = spim_d2f_rn(spim_ddiv_rn((double)p.v.w[v][idx], sum));
                } else {
                    p.sumw[idx] = sum > 1.0 ? spim_d2f_rn(sum) : 1.f;
                }
                long long pl = p.chunk > 0 ? idx / p.chunk : 0;
                const int portion = (int)(pl < p.nportions - 1 ? pl : p.nportions - 1);
                if (portion != cur) { flush(scnt, smin, cur, c, mn); cur = portion; c = 0; mn = 0xffffffffu; }
                c += count;
                if (count < mn) mn = count;
            }
        }
        if (stats) {
            flush(scnt, smin, cur, c, mn);
            SPIM_BARRIER();
            SPIM_FOR_ITEMS(i, p.nportions) {
                if (smin[i] != 0xffffffffu) { atomic_add_u64(p.cnt + i, scnt[i]); atomic_min_u32(p.pmin + i, smin[i]); }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// loader normalisation: (v - min) / (max - min) in float
// ---------------------------------------------------------------------------------------------
// order-preserving map float -> uint32 so that integer atomics give float min / max
SPIM_HD unsigned int float_order_bits(float f) {
    unsigned int b;
    memcpy(&b, &f, 4);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
SPIM_HD float float_from_order_bits(unsigned int u) {
    const unsigned int b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
    memcpy(&f, &b, 4);
    return f;
}
struct MinMaxK {
    struct Params { const float* p; long long n; unsigned int* mm; int nblocks; };   // mm[0] = min, mm[1] = max (order bits)
    SPIM_DEV static void run(const Params& p, int bid, float2*) {
        unsigned int lo = 0xffffffffu, hi = 0u;
        for (long long base = (long long)bid * kChunk; base < p.n; base += (long long)p.nblocks * kChunk) {
            SPIM_FOR_ITEMS(i, kChunk) {
                const long long idx = base + i;
                if (idx >= p.n) continue;
                const float v = p.p[idx];
                if (v != v) continue;          // "v < min" / "v > max" are false for NaN in the Java loop
                const unsigned int u = float_order_bits(v);
                lo = u < lo ? u : lo;
                hi = u > hi ? u : hi;
            }
        }
#if defined(SPIM_HOST_EMU)
        std::lock_guard<std::mutex> l(g_emu_atomic_mutex);
        if (lo < p.mm[0]) p.mm[0] = lo;
        if (hi > p.mm[1]) p.mm[1] = hi;
#else
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(p.mm, lo); atomicMax(p.mm + 1, hi); }
#endif
    }
};
struct NormalizeK {
    struct Params { float* p; long long n; float mn, diff; int nblocks; };
    SPIM_DEV static void run(const Params& p, int bid, float2*) {
        for (long long base = (long long)bid * kChunk; base < p.n; base += (long long)p.nblocks * kChunk) {
            SPIM_FOR_ITEMS(i, kChunk) {
                const long long idx = base + i;
                if (idx < p.n) p.p[idx] = spim_fdiv_rn(spim_fsub_rn(p.p[idx], p.mn), p.diff);
            }
        }
    }
};

}  // namespace spim
