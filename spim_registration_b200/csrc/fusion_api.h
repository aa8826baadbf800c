// C ABI of the device-side fusion pre-step (include/spim_fusion.h) over the kernels in fusion.h.
// Included at the end of spim_b200.cu (needs mvd_session, fail(), SPIM_API_BEGIN / END).
#pragma once
#include "../../include/spim_fusion.h"

namespace {

// static lookUp[1001], FW/BlendingRealRandomAccess.java:44-54: d accumulates in double, index = Math.round(d * 1000)
void blending_lookup(double* lut) {
    for (int i = 0; i < 1001; ++i) lut[i] = 0.0;
    for (volatile double d = 0; d <= 1.0001; d = d + 0.001) {
        const int idx = (int)floor(d * 1000.0 + 0.5);
        if (idx >= 0 && idx <= 1000) lut[idx] = (cos((1 - d) * 3.14159265358979323846) + 1) / 2;
    }
}

const double* session_lut(mvd_session* s) {
    if (!s->d_lut) {
        double lut[1001];
        blending_lookup(lut);
        s->d_lut = (double*)rt::dmalloc(sizeof(lut));
        rt::h2d(s->d_lut, lut, sizeof(lut), s->stream);
        rt::stream_sync(s->stream);
    }
    return s->d_lut;
}

SrcVol src_vol(const float* p, const int dims[3]) {
    SrcVol v;
    v.p = p; v.sz = dims[0]; v.sy = dims[1]; v.sx = dims[2];
    return v;
}

// AffineTransform3D.apply(double[], double[])
void affine_apply_host(const double* m, const double* s, double* t) {
    for (int r = 0; r < 3; ++r) {
        volatile double a = s[0] * m[4 * r];
        volatile double b = s[1] * m[4 * r + 1];
        volatile double c = s[2] * m[4 * r + 2];
        volatile double u = a + b;
        u = u + c;
        u = u + m[4 * r + 3];
        t[r] = u;
    }
}

}  // namespace

extern "C" {

int mvd_blending_lookup(double out[1001]) {
    SPIM_API_BEGIN
    if (!out) return fail("mvd_blending_lookup: null argument");
    blending_lookup(out);
    return 0;
    SPIM_API_END
}

int mvd_load_stack(mvd_session* s, const float* stack, const int dims[3], int normalize) {
    SPIM_API_BEGIN
    if (!s) return fail("mvd_load_stack: null session");
    rt::set_device(s->prm.device);
    if (!stack) {
        rt::stream_sync(s->stream);
        rt::dfree(s->d_stack);
        s->d_stack = nullptr; s->stack_cap = 0;
        s->stack_dims[0] = s->stack_dims[1] = s->stack_dims[2] = 0;
        return 0;
    }
    if (!dims) return fail("mvd_load_stack: null dims");
    for (int d = 0; d < 3; ++d) if (dims[d] < 1) return fail("mvd_load_stack: bad dims");
    const long long n = (long long)dims[0] * dims[1] * dims[2];
    if ((size_t)n > s->stack_cap) {
        rt::stream_sync(s->stream);
        rt::dfree(s->d_stack);
        s->d_stack = nullptr; s->stack_cap = 0;
        s->d_stack = (float*)rt::dmalloc((size_t)n * sizeof(float));
        s->stack_cap = (size_t)n;
    }
    for (int d = 0; d < 3; ++d) s->stack_dims[d] = dims[d];
    rt::h2d(s->d_stack, stack, (size_t)n * sizeof(float), s->stream);
    if (normalize) {
        unsigned int* d_mm = (unsigned int*)rt::dmalloc(2 * sizeof(unsigned int));
        const unsigned int init[2] = {0xffffffffu, 0u};
        rt::h2d(d_mm, init, sizeof(init), s->stream);
        MinMaxK::Params mp; mp.p = s->d_stack; mp.n = n; mp.mm = d_mm; mp.nblocks = ew_blocks(n);
        rt::launch<MinMaxK>(mp, mp.nblocks, kThreads, 0, s->stream);
        unsigned int mm[2];
        rt::d2h(mm, d_mm, sizeof(mm), s->stream);
        rt::stream_sync(s->stream);
        rt::dfree(d_mm);
        // float min = Float.MAX_VALUE, max = -Float.MAX_VALUE when nothing compares (all NaN)
        const float mn = mm[0] == 0xffffffffu ? 3.402823466e+38f : float_from_order_bits(mm[0]);
        const float mx = mm[1] == 0u ? -3.402823466e+38f : float_from_order_bits(mm[1]);
        NormalizeK::Params np_; np_.p = s->d_stack; np_.n = n; np_.mn = mn;
        { volatile float d = mx - mn; np_.diff = d; }
        np_.nblocks = ew_blocks(n);
        rt::launch<NormalizeK>(np_, np_.nblocks, kThreads, 0, s->stream);
    }
    rt::stream_sync(s->stream);   // the caller may reuse its host buffer
    return 0;
    SPIM_API_END
}

int mvd_transform_view(mvd_session* s, int view, const mvd_transform* t) {
    SPIM_API_BEGIN
    if (!s || !t) return fail("mvd_transform_view: null argument");
    if (t->struct_size != (int)sizeof(mvd_transform)) return fail("mvd_transform_view: struct_size mismatch");
    if (view < 0 || view >= s->prm.num_views) return fail("mvd_transform_view: view index out of range");
    if (!t->want_image && !t->want_weight) return fail("mvd_transform_view: nothing requested");
    if (t->want_image && !s->d_stack) return fail("mvd_transform_view: no stack loaded (mvd_load_stack)");
    if (s->stack_dims[0] < 1) return fail("mvd_transform_view: stack dimensions unknown (mvd_load_stack)");
    rt::set_device(s->prm.device);
    const size_t bytes = (size_t)s->N * sizeof(float);
    if (t->want_image && !s->d_img[view]) s->d_img[view] = (float*)s->dalloc(bytes);
    if (t->want_weight && !s->d_w[view]) s->d_w[view] = (float*)s->dalloc(bytes);
    ResampleParams p;
    memset(&p, 0, sizeof(p));
    p.src = src_vol(s->d_stack, s->stack_dims);
    for (int i = 0; i < 12; ++i) p.inv.m[i] = t->inverse[i];
    for (int d = 0; d < 3; ++d) { p.off[d] = (double)t->offset[d]; p.on[d] = s->n[d]; }
    p.pos_mode = 0;
    p.ext = EXT_MIRROR_SINGLE; p.ext_value = 0.f;
    p.out_img = t->want_image ? s->d_img[view] : nullptr;
    p.out_w = t->want_weight ? s->d_w[view] : nullptr;
    p.clamp_inside = 1;
    p.min_value = s->prm.min_value;
    if (t->want_weight) {
        for (int d = 0; d < 3; ++d) {
            p.blend.border[d] = t->border[d]; p.blend.blending[d] = t->range[d];
            p.blend.imin[d] = 0; p.blend.dim_minus1[d] = s->stack_dims[2 - d] - 1;
        }
        p.blend.lut = session_lut(s);
    }
    p.nblocks = ew_blocks(s->N);
    rt::launch<ResampleK>(p, p.nblocks, kThreads, 0, s->stream);
    rt::stream_sync(s->stream);
    s->inited = false;
    if (t->want_weight) { s->wn_valid = false; s->virtual_weights = false; }
    return 0;
    SPIM_API_END
}

int mvd_set_psf(mvd_session* s, int view, const float* psf, const int psf_dims[3]) {
    SPIM_API_BEGIN
    if (!s || !psf || !psf_dims) return fail("mvd_set_psf: null argument");
    if (view < 0 || view >= s->prm.num_views) return fail("mvd_set_psf: view index out of range");
    for (int d = 0; d < 3; ++d) if (psf_dims[d] < 1) return fail("mvd_set_psf: bad psf dims");
    HostVol& k = s->psf[view];
    for (int d = 0; d < 3; ++d) k.d[d] = psf_dims[d];
    k.v.assign(psf, psf + k.size());
    s->inited = false;
    return 0;
    SPIM_API_END
}

int mvd_normalize_weights(mvd_session* s, int mode, int num_portions, int* min_views, double* avg_views) {
    SPIM_API_BEGIN
    if (!s) return fail("mvd_normalize_weights: null session");
    if (mode != MVD_WEIGHTS_PRECOMPUTED && mode != MVD_WEIGHTS_VIRTUAL) return fail("mvd_normalize_weights: bad mode");
    if (num_portions < 1 || num_portions > 4096) return fail("mvd_normalize_weights: num_portions out of range [1, 4096]");
    const int V = s->prm.num_views;
    for (int v = 0; v < V; ++v) if (!s->d_w[v]) return fail("mvd_normalize_weights: view " + std::to_string(v) + " has no weight image");
    rt::set_device(s->prm.device);
    WeightNormParams p;
    memset(&p, 0, sizeof(p));
    p.v.nviews = V;
    for (int v = 0; v < V; ++v) p.v.w[v] = s->d_w[v];
    p.n = s->N;
    p.mode = mode == MVD_WEIGHTS_PRECOMPUTED ? 0 : 1;
    if (mode == MVD_WEIGHTS_VIRTUAL) {
        if (!s->d_sumw) s->d_sumw = (float*)s->dalloc((size_t)s->N * sizeof(float));
        p.sumw = s->d_sumw;
    }
    p.osem = 1.0;
    p.nportions = num_portions;
    p.chunk = s->N / num_portions;          // FusionHelper.divideIntoPortions: the last portion takes the remainder
    p.cnt = (unsigned long long*)rt::dmalloc(num_portions * sizeof(unsigned long long));
    p.pmin = (unsigned int*)rt::dmalloc(num_portions * sizeof(unsigned int));
    rt::dzero(p.cnt, num_portions * sizeof(unsigned long long), s->stream);
    std::vector<unsigned int> hmin(num_portions, 0xffffffffu);
    rt::h2d(p.pmin, hmin.data(), num_portions * sizeof(unsigned int), s->stream);
    p.nblocks = ew_blocks(s->N);
    rt::launch<WeightNormK>(p, p.nblocks, kThreads, (size_t)num_portions * 12 + 16, s->stream);
    std::vector<unsigned long long> hcnt(num_portions);
    rt::d2h(hcnt.data(), p.cnt, num_portions * sizeof(unsigned long long), s->stream);
    rt::d2h(hmin.data(), p.pmin, num_portions * sizeof(unsigned int), s->stream);
    rt::stream_sync(s->stream);
    rt::dfree(p.cnt); rt::dfree(p.pmin);
    // WeightNormalizer.java:94-108: min over portions of (int)Math.round(min), mean over portions of the portion means
    int mn = V;
    double avg = 0.0;
    for (int i = 0; i < num_portions; ++i) {
        const long long loop = (i == num_portions - 1) ? p.chunk + s->N % num_portions : p.chunk;
        const int pm = hmin[i] == 0xffffffffu ? V : std::min<int>(V, (int)hmin[i]);
        mn = std::min(mn, pm);
        avg += (double)hcnt[i] / (double)loop;     // 0 / 0 = NaN for an empty portion, as in the reference
    }
    avg /= (double)num_portions;
    s->wn_valid = true; s->wn_min = mn; s->wn_avg = avg;
    s->virtual_weights = (mode == MVD_WEIGHTS_VIRTUAL);
    s->inited = false;
    if (min_views) *min_views = mn;
    if (avg_views) *avg_views = avg;
    return 0;
    SPIM_API_END
}

int mvd_get_view(mvd_session* s, int view, int which, float* out) {
    SPIM_API_BEGIN
    if (!s || !out) return fail("mvd_get_view: null argument");
    if (view < 0 || view >= s->prm.num_views || (which != 0 && which != 1)) return fail("mvd_get_view: bad view/which");
    const float* src = which == 0 ? s->d_img[view] : s->d_w[view];
    if (!src) return fail("mvd_get_view: buffer not set (constant weight or missing image)");
    rt::set_device(s->prm.device);
    rt::d2h(out, src, (size_t)s->N * sizeof(float), s->stream);
    rt::stream_sync(s->stream);
    return 0;
    SPIM_API_END
}

int mvd_extract_psf(mvd_session* s, int n_beads, const double* locations_xyz, const int size[3], int normalize, float* out) {
    SPIM_API_BEGIN
    if (!s || !size || !out || (n_beads > 0 && !locations_xyz)) return fail("mvd_extract_psf: null argument");
    if (n_beads < 0) return fail("mvd_extract_psf: negative bead count");
    if (!s->d_stack) return fail("mvd_extract_psf: no stack loaded (mvd_load_stack)");
    for (int d = 0; d < 3; ++d) if (size[d] < 1) return fail("mvd_extract_psf: bad size");
    rt::set_device(s->prm.device);
    const long long n = (long long)size[0] * size[1] * size[2];
    float* d_out = (float*)rt::dmalloc((size_t)n * sizeof(float));
    double* d_loc = (double*)rt::dmalloc((size_t)std::max(1, n_beads) * 3 * sizeof(double));
    if (n_beads > 0) rt::h2d(d_loc, locations_xyz, (size_t)n_beads * 3 * sizeof(double), s->stream);
    ExtractPsfParams p;
    memset(&p, 0, sizeof(p));
    p.src = src_vol(s->d_stack, s->stack_dims);
    p.loc = d_loc; p.n_beads = n_beads;
    for (int d = 0; d < 3; ++d) p.size[d] = size[d];
    p.out = d_out;
    p.nblocks = ew_blocks(n);
    rt::launch<ExtractPsfK>(p, p.nblocks, kThreads, 0, s->stream);
    rt::d2h(out, d_out, (size_t)n * sizeof(float), s->stream);
    rt::stream_sync(s->stream);
    rt::dfree(d_out); rt::dfree(d_loc);
    if (normalize) {
        // ExtractPSF.normalize (:298-316): PSF-sized, in double on the host like LRFFT.init's normalisation
        double mn = 1.7976931348623157e308, mx = -1.7976931348623157e308;
        for (long long i = 0; i < n; ++i) { const double v = (double)out[i]; if (v < mn) mn = v; if (v > mx) mx = v; }
        for (long long i = 0; i < n; ++i) { volatile double a = (double)out[i] - mn; volatile double b = mx - mn; out[i] = (float)(a / b); }
    }
    return 0;
    SPIM_API_END
}

int mvd_transform_psf_size(const int dims[3], const double model[12], int out_dims[3], double offset_xyz[3]) {
    SPIM_API_BEGIN
    if (!dims || !model || !out_dims || !offset_xyz) return fail("mvd_transform_psf_size: null argument");
    for (int d = 0; d < 3; ++d) if (dims[d] < 1) return fail("mvd_transform_psf_size: bad dims");
    const int dx[3] = {dims[2], dims[1], dims[0]};           // (x, y, z)
    // AffineTransform3D.estimateBounds: min / max over the 8 transformed corners of [0, dim - 1]
    double lo[3] = {1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308};
    double hi[3] = {-lo[0], -lo[0], -lo[0]};
    for (int c = 0; c < 8; ++c) {
        const double sc[3] = {(c & 1) ? (double)(dx[0] - 1) : 0.0, (c & 2) ? (double)(dx[1] - 1) : 0.0, (c & 4) ? (double)(dx[2] - 1) : 0.0};
        double t[3];
        affine_apply_host(model, sc, t);
        for (int d = 0; d < 3; ++d) { if (t[d] < lo[d]) lo[d] = t[d]; if (t[d] > hi[d]) hi[d] = t[d]; }
    }
    const double center[3] = {(double)(dx[0] / 2), (double)(dx[1] / 2), (double)(dx[2] / 2)};
    double tc[3];
    affine_apply_host(model, center, tc);
    for (int d = 0; d < 3; ++d) {
        volatile double size = hi[d] - lo[d];
        if (!(size >= 0 && size < 1e6)) return fail("mvd_transform_psf_size: transformed PSF extent out of range");
        int ns = (int)size + 1;
        if (ns % 2 == 0) ++ns;
        out_dims[2 - d] = ns;
        volatile double off = tc[d] - (double)(ns / 2);
        offset_xyz[d] = off;
    }
    return 0;
    SPIM_API_END
}

int mvd_transform_psf(const float* psf, const int dims[3], const double model[12], const double inverse[12],
                      float* out, const int out_dims[3], int device) {
    SPIM_API_BEGIN
    if (!psf || !dims || !model || !inverse || !out || !out_dims) return fail("mvd_transform_psf: null argument");
    int want[3]; double off[3];
    if (int rc = mvd_transform_psf_size(dims, model, want, off)) return rc;
    for (int d = 0; d < 3; ++d) if (want[d] != out_dims[d]) return fail("mvd_transform_psf: out_dims do not match mvd_transform_psf_size");
    const int ndev = rt::device_count();
    if (ndev <= 0) return fail("mvd_transform_psf: no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail("mvd_transform_psf: bad device ordinal");
    rt::set_device(device);
    const long long nin = (long long)dims[0] * dims[1] * dims[2];
    const long long nout = (long long)out_dims[0] * out_dims[1] * out_dims[2];
    rt::Stream st = rt::stream_create();
    float* d_in = (float*)rt::dmalloc((size_t)nin * sizeof(float));
    float* d_out = (float*)rt::dmalloc((size_t)nout * sizeof(float));
    rt::h2d(d_in, psf, (size_t)nin * sizeof(float), st);
    ResampleParams p;
    memset(&p, 0, sizeof(p));
    p.src = src_vol(d_in, dims);
    for (int i = 0; i < 12; ++i) p.inv.m[i] = inverse[i];
    for (int d = 0; d < 3; ++d) { p.off[d] = off[d]; p.on[d] = out_dims[d]; }
    p.pos_mode = 1;
    p.ext = EXT_ZERO; p.ext_value = 0.f;
    p.out_img = d_out; p.out_w = nullptr;
    p.clamp_inside = 0;
    p.nblocks = ew_blocks(nout);
    rt::launch<ResampleK>(p, p.nblocks, kThreads, 0, st);
    rt::d2h(out, d_out, (size_t)nout * sizeof(float), st);
    rt::stream_sync(st);
    rt::dfree(d_in); rt::dfree(d_out);
    rt::stream_destroy(st);
    return 0;
    SPIM_API_END
}

}  // extern "C"
