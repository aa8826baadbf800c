"""Parity cases of the fusion pre-step, shared by the emulator tests (CPU) and the
GPU tests: each takes the loaded library (emulator or CUDA) and compares the C-ABI path with oracle/fusion_oracle.py.

The arithmetic is reproduced operation by operation (double position math, float taps, float accumulation), so the
bar here is BIT-EXACT equality (NaNs included), not a tolerance."""
import math

import numpy as np

from oracle import fusion_oracle as F
from oracle import mvdecon_oracle as O
from spim_registration_b200 import fusion
from spim_registration_b200.deconvolution import Session


def view_model(angle_deg, stack_shape_zyx, z_scale=2.5, shift=(0.0, 0.0, 0.0)):
    """A SPIM-like registration: anisotropic z calibration, rotation about the y axis through the stack centre, shift.
    Row-packed (x, y, z) affine mapping raw-stack coordinates to global coordinates."""
    nz, ny, nx = stack_shape_zyx
    a = math.radians(angle_deg)
    c, s = math.cos(a), math.sin(a)
    cx, cy, cz = (nx - 1) / 2.0, (ny - 1) / 2.0, (nz - 1) / 2.0 * z_scale
    # p' = R (S p - centre) + centre + shift,  S = diag(1, 1, z_scale)
    m = [c, 0.0, s * z_scale, 0.0,
         0.0, 1.0, 0.0, 0.0,
         -s, 0.0, c * z_scale, 0.0]
    m[3] = -(c * cx + s * cz) + cx + shift[0]
    m[7] = shift[1]
    m[11] = -(-s * cx + c * cz) + cz + shift[2]
    return fusion.AffineTransform3D(m)


def make_stack(shape, seed):
    rng = np.random.default_rng(seed)
    return (rng.random(shape, dtype=np.float32) * np.float32(900.0) + np.float32(100.0)).astype(np.float32)


def assert_same(a, b, what):
    np.testing.assert_array_equal(a, b, err_msg=what)   # NaN == NaN at the same positions


def transform_case(lib, stack_shape, out_dims, angle, offset_xyz, border, blend_range, weights=True, image=True,
                   normalize=False, seed=0, z_scale=2.5):
    stack = make_stack(stack_shape, seed)
    model = view_model(angle, stack_shape, z_scale)
    inv = F.invert_affine(model.getRowPackedCopy())
    np.testing.assert_array_equal(inv, model.inverse().getRowPackedCopy())
    src = F.loader_normalize(stack) if normalize else stack
    with Session(out_dims, 1, O.INDEPENDENT, lib=lib) as s:
        fusion.load_stack(s, stack, normalize=normalize)
        bl = fusion.Blending(stack_shape[::-1], border, blend_range) if weights else None
        fusion.transform_view(s, 0, model, offset_xyz, bl, want_image=image)
        if image:
            assert_same(fusion.get_view(s, 0, 0), F.transform_input(src, inv, out_dims, offset_xyz), "transformed image")
        if weights:
            assert_same(fusion.get_view(s, 0, 1),
                        F.transform_weights(stack_shape, inv, out_dims, offset_xyz, border, blend_range), "blending weight")


def make_view_set(V, stack_shape, out_dims, seed=0):
    stacks = [make_stack(stack_shape, seed + v) for v in range(V)]
    models = [view_model(360.0 * v / V, stack_shape, shift=(1.5 * v, -0.75 * v, 0.25 * v)) for v in range(V)]
    return stacks, models


def oracle_views(stacks, models, out_dims, offset_xyz, border, blend_range, normalize=True):
    imgs, ws = [], []
    for st, m in zip(stacks, models):
        src = F.loader_normalize(st) if normalize else st
        inv = F.invert_affine(m.getRowPackedCopy())
        i, w = F.transform_input_and_weights(src, inv, out_dims, offset_xyz, border, blend_range)
        imgs.append(i); ws.append(w)
    return imgs, ws


def normalize_case(lib, virtual, num_portions, osem_index=0, osem=1.0, V=3, stack_shape=(9, 20, 22), out_dims=(14, 18, 24),
                   offset_xyz=(-2, 1, 3), border=(2, 2, 1), blend_range=(6, 6, 3)):
    """WeightNormalizer + OSEM clamp: device weights after mvd_init must equal the oracle's bit for bit."""
    stacks, models = make_view_set(V, stack_shape, out_dims)
    imgs, ws = oracle_views(stacks, models, out_dims, offset_xyz, border, blend_range)
    with Session(out_dims, V, O.INDEPENDENT, generation=2, osem_speedup=osem, osem_index=osem_index, lib=lib) as s:
        pfd = fusion.ProcessForDeconvolution(s, offset_xyz, border, blend_range, numThreads=max(1, num_portions // 2))
        psfs = [np.ones((3, 3, 3), dtype=np.float32)] * V
        wt = fusion.WeightType.VIRTUAL_WEIGHTS if virtual else fusion.WeightType.PRECOMPUTED_WEIGHTS
        assert pfd.fuseStacksAndGetPSFs(stacks, models, osem_index, osem, wt, psfs=psfs)
        nport = max(1, num_portions // 2) * 2
        if virtual:
            sumw, mn, avg = F.weight_normalizer_virtual(ws, nport)
        else:
            wn, mn, avg = F.weight_normalizer_direct(ws, nport)
        assert pfd.getMinOverlappingViews() == max(1, mn)
        assert pfd.getAvgOverlappingViews() == max(1.0, avg)
        eff = float(osem)
        if osem_index == 1:
            eff = float(max(1, mn))
        elif osem_index == 2:
            eff = max(1.0, avg)
        assert pfd.osemspeedup == eff
        for v in range(V):
            assert_same(fusion.get_view(s, v, 0), imgs[v], f"image {v}")
        s.init()
        assert s.info().osem == eff
        if virtual:
            want = [F.normalizing_access(w, sumw, eff) for w in ws]
        else:
            want = F.adjust_for_osem(wn, eff)
        for v in range(V):
            assert_same(fusion.get_view(s, v, 1), want[v], f"weight {v}")


def pipeline_case(lib, typ=O.EFFICIENT_BAYESIAN, V=3, stack_shape=(10, 22, 24), out_dims=(16, 20, 26), iters=3,
                  offset_xyz=(-1, 1, 2), border=(1, 1, 0), blend_range=(5, 5, 2), psf_size=5):
    """stacks -> device transform + weights + normalisation -> deconvolution, against the oracle end to end.
    The fused bounding box is chosen inside the union of the views so that the reference's 0/0 weights do not occur."""
    from spim_registration_b200 import synthetic
    stacks, models = make_view_set(V, stack_shape, out_dims, seed=11)
    imgs, ws = oracle_views(stacks, models, out_dims, offset_xyz, border, blend_range)
    sumw, mn, avg = F.weight_normalizer_virtual(ws, 2)
    wv = [F.normalizing_access(w, sumw, 1.0) for w in ws]
    psfs = synthetic.make_psfs(V, psf_size)
    with Session(out_dims, V, typ, generation=2, lam=0.006, lib=lib) as s:
        pfd = fusion.ProcessForDeconvolution(s, offset_xyz, border, blend_range, numThreads=1)
        assert pfd.fuseStacksAndGetPSFs(stacks, models, 0, 1.0, fusion.WeightType.VIRTUAL_WEIGHTS, psfs=psfs)
        s.init()
        s.run(iters)
        s.finish()
        psi = s.get_psi()
    ref = O.deconvolve(imgs, wv, psfs, O.DeconParams(iteration_type=typ, num_iterations=iters, lam=0.006, gen=O.GEN2))
    per, l2 = O.parity_errors(psi, ref.psi)
    assert per <= 1e-3 and l2 <= 1e-4, (per, l2)


def psf_case(lib, stack_shape=(12, 26, 28), n_beads=7, psf_size_xyz=(9, 7, 5), angle=30.0, seed=3):
    stack = make_stack(stack_shape, seed)
    rng = np.random.default_rng(seed + 1)
    nz, ny, nx = stack_shape
    # sub-pixel bead locations, some close to the faces so that the periodic extension is exercised
    loc = np.stack([rng.uniform(-1.0, nx + 1.0, n_beads), rng.uniform(-1.0, ny + 1.0, n_beads), rng.uniform(-1.0, nz + 1.0, n_beads)], axis=1)
    model = view_model(angle, stack_shape)
    with Session((4, 4, 4), 1, O.INDEPENDENT, lib=lib) as s:
        fusion.load_stack(s, stack, normalize=True)
        src = F.loader_normalize(stack)
        raw = fusion.ExtractPSF.extractPSFLocal(s, loc, psf_size_xyz, normalize=False)
        assert_same(raw, F.extract_psf_local(src, loc, psf_size_xyz), "extractPSFLocal")
        e = fusion.ExtractPSF(lib=lib)
        e.extractNextImg(s, "view0", model, loc, psf_size_xyz)
    want_t, want_o = F.extract_next_img(src, model.getRowPackedCopy(), loc, psf_size_xyz)
    assert_same(e.getInputCalibrationPSFs()["view0"], want_o, "normalised original PSF")
    got = e.getTransformedPSF("view0")
    assert got.shape == want_t.shape and all(d % 2 == 1 for d in got.shape)
    assert_same(got, want_t, "transformed PSF")


def identity_properties_case(lib, shape, blend_range=(12, 12, 12)):
    """Size-independent properties: the identity registration reproduces the stack bit for bit (clamped to minValue), its
    blending weight is separable, and a pure integer shift reproduces the shifted stack with zeros outside."""
    rng = np.random.default_rng(5)
    stack = rng.random(shape, dtype=np.float32)
    ident = fusion.AffineTransform3D()
    cz, cy, cx = shape[0] // 2, shape[1] // 2, shape[2] // 2
    with Session(shape, 1, O.INDEPENDENT, lib=lib) as s:
        fusion.load_stack(s, stack)
        fusion.transform_view(s, 0, ident, (0, 0, 0), fusion.Blending(shape[::-1], (0, 0, 0), blend_range))
        img = fusion.get_view(s, 0, 0)
        np.testing.assert_array_equal(img, np.maximum(np.float32(1e-4), stack))
        w = fusion.get_view(s, 0, 1)
        assert w.min() == 0.0 and w.max() == 1.0 and w[cz, cy, cx] == 1.0 and np.all(w[0] == 0)
        wz, wy, wx = w[:, cy, cx], w[cz, :, cx], w[cz, cy, :]
        np.testing.assert_allclose(w[::7, ::5, ::3], (wz[::7, None, None] * wy[None, ::5, None] * wx[None, None, ::3]), rtol=3e-7)
        fusion.transform_view(s, 0, ident, (3, -2, 1), None)
        sh = fusion.get_view(s, 0, 0)
        np.testing.assert_array_equal(sh[:-1, 2:, :-3], np.maximum(np.float32(1e-4), stack[1:, :-2, 3:]))
        assert np.all(sh[-1] == 0) and np.all(sh[:, :2] == 0) and np.all(sh[:, :, -3:] == 0)
