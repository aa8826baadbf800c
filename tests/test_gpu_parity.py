"""GPU parity tests: the CUDA library (sm_100a) through the C-ABI against the oracle.
Run on the B200 box with  python -m pytest tests -m gpu."""
import numpy as np
import pytest

import parity_cases as P
from oracle import mvdecon_oracle as O
from spim_registration_b200 import native, synthetic

pytestmark = pytest.mark.gpu


def test_device_queries(gpu):
    c = native.CUDAFourierConvolution()
    n = c.getNumDevicesCUDA()
    assert n >= 1
    name = bytearray(256)
    c.getNameDeviceCUDA(0, name)
    assert b"NVIDIA" in bytes(name)
    assert c.getCUDAcomputeCapabilityMajorVersion(0) == 10
    assert c.getMemDeviceCUDA(0) > 100 * (1 << 30)
    assert 0 < c.getFreeMemDeviceCUDA(0) <= c.getMemDeviceCUDA(0)
    assert c.getCUDAcomputeCapabilityMajorVersion(n + 3) == -1
    from spim_registration_b200 import CUDATools
    devs = CUDATools.queryCUDADetails(c, askForMultipleDevices=True)
    assert len(devs) == n and devs[0].getMajorComputeVersion() == 10


@pytest.mark.parametrize("ext", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("shape,kshape", [((8, 8, 8), (3, 3, 3)), ((9, 7, 11), (3, 5, 3)), ((6, 6, 6), (4, 2, 6)),
                                          ((33, 47, 65), (7, 9, 11)), ((64, 64, 64), (15, 15, 15)),
                                          ((1, 200, 300), (1, 31, 31)), ((100, 90, 110), (31, 31, 31))])
def test_conv_all_extensions(gpu, shape, kshape, ext):
    P.conv_case(gpu, shape, kshape, ext)


def test_conv_halo_wider_than_image(gpu):
    for ext in (2, 3, 4):
        P.conv_case(gpu, (3, 4, 6), (7, 9, 5), ext)


@pytest.mark.parametrize("n", [16, 18, 20, 22, 24, 26, 28, 30, 36, 40, 42, 44, 48, 50, 52, 54, 56, 60, 64, 66, 70, 72, 78, 80,
                               84, 88, 90, 96, 98, 100, 104, 108, 110, 112, 120, 126, 128, 130, 132, 140, 144, 256, 560, 1080])
def test_every_radix_path_along_each_axis(gpu, n):
    for shape in ((n, 4, 8), (4, n, 8), (4, 4, n)):
        P.legacy_case(gpu, shape, (3, 3, 3), seed=n)


@pytest.mark.parametrize("shape,kshape", [((16, 16, 16), (5, 5, 5)), ((12, 20, 18), (3, 7, 5)), ((10, 9, 7), (3, 3, 3)),
                                          ((8, 8, 8), (8, 8, 8)), ((128, 128, 128), (15, 15, 15)),
                                          ((142, 142, 142), (15, 15, 15)), ((256, 256, 256), (31, 31, 31))])
def test_legacy_entry_is_circular_convolution(gpu, shape, kshape):
    P.legacy_case(gpu, shape, kshape)


def test_legacy_entry_never_called_variant(gpu):
    c = native.CUDAFourierConvolution()
    rng = np.random.default_rng(3)
    im = rng.random((12, 14, 16), dtype=np.float32)
    k = rng.random((3, 3, 5), dtype=np.float32)
    out = c.convolution3DfftCUDA(im, im.shape, k, k.shape, 0)
    ref = O.circular_convolve(im, k, dtype=np.float64)
    assert np.abs(out - ref).max() / np.abs(ref).max() < P.TOL_CONV


def test_legacy_entry_bad_device_leaves_buffer_untouched(gpu):
    c = native.CUDAFourierConvolution()
    im = np.ones((8, 8, 8), np.float32)
    c.convolution3DfftCUDAInPlace(im, im.shape, np.ones((3, 3, 3), np.float32), (3, 3, 3), 99)
    assert np.all(im == 1.0) and "device" in c.last_error()


def test_golden_conv(gpu):
    P.golden_conv_case(gpu)


@pytest.mark.parametrize("gen", [1, 2])
@pytest.mark.parametrize("typ", [0, 1, 2, 3])
def test_golden_deconvolution(gpu, gen, typ):
    P.golden_case(gpu, gen, typ)


@pytest.mark.parametrize("gen", [1, 2])
@pytest.mark.parametrize("typ", [0, 1, 2, 3])
def test_deconvolution_all_types(gpu, gen, typ):
    P.decon_case(gpu, (40, 48, 56), 3, 9, typ, gen, 4)


def test_config1_beads_128_efficient_bayesian_10_iterations(gpu):
    """BASELINE.json configs[0]: 3 views, 128^3 beads, 15^3 PSFs, Efficient-Bayesian, 10 iterations."""
    per, l2 = P.decon_case(gpu, (128, 128, 128), 3, 15, O.EFFICIENT_BAYESIAN, 2, 10, kind="beads")
    print(f"config1 parity: per-voxel {per:.3e}  L2 {l2:.3e}")


def test_config2_quarter_scale_7_views(gpu):
    """BASELINE.json configs[1] at quarter scale per axis: 7 views 128x128x64, 31^3 PSFs, EB, 20 iterations."""
    P.decon_case(gpu, (64, 128, 128), 7, 31, O.EFFICIENT_BAYESIAN, 2, 20, kind="specimen")


def test_config3_quarter_scale_optimization2(gpu):
    """configs[2] at quarter scale: 6 views 256x256x128, 31^3 PSFs, OPTIMIZATION_II."""
    P.decon_case(gpu, (128, 256, 256), 6, 31, O.OPTIMIZATION_II, 2, 3, kind="specimen")


def test_config4_quarter_scale_independent_tikhonov_blending(gpu):
    """configs[3] at quarter scale: INDEPENDENT + Tikhonov + per-view blending weights."""
    P.decon_case(gpu, (128, 256, 256), 6, 31, O.INDEPENDENT, 2, 3, kind="specimen", weight_mode="blending", lam=0.006)


def test_odd_dims_no_weights_no_tikhonov(gpu):
    P.decon_case(gpu, (37, 41, 43), 2, 7, O.INDEPENDENT, 2, 3, lam=0.0, use_weights=False)


def test_gen1_osem_variants(gpu):
    P.decon_case(gpu, (24, 28, 32), 3, 5, O.OPTIMIZATION_II, 1, 2, osem_index=2)
    P.decon_case(gpu, (24, 28, 32), 3, 5, O.OPTIMIZATION_I, 1, 2, osem_index=1)
    P.decon_case(gpu, (24, 28, 32), 3, 5, O.EFFICIENT_BAYESIAN, 1, 2, osem=2.0, osem_index=0)


def test_second_init_with_larger_psf_recreates_buffers(gpu):
    P.reinit_case(gpu, (24, 28, 32))


def test_exact_tikhonov_switch(gpu):
    P.exact_tikhonov_case(gpu, (30, 34, 38))


def test_views_with_different_psf_sizes(gpu):
    shape = (30, 34, 38)
    _, imgs, ws, _ = synthetic.make_dataset(shape, 3, 7)
    psfs = [synthetic.make_psf(7, 0, 3), synthetic.make_psf(9, 1, 3), synthetic.make_psf(5, 2, 3)]
    ref = O.deconvolve(imgs, ws, psfs, O.DeconParams(iteration_type=O.INDEPENDENT, num_iterations=3))
    psi, *_ = P.run_session(gpu, imgs, ws, psfs, O.INDEPENDENT, 2, 3)
    per, l2 = O.parity_errors(psi, ref.psi)
    assert per <= P.TOL_PER_VOXEL and l2 <= P.TOL_L2


def test_set_psi_resume_equals_uninterrupted_run(gpu):
    """checkpoint/resume through get_psi/set_psi (the reference's 'initialImage' hook)."""
    shape = (24, 28, 32)
    _, imgs, ws, psfs = synthetic.make_dataset(shape, 2, 5)
    full, *_ = P.run_session(gpu, imgs, ws, psfs, 2, 2, 4)
    from spim_registration_b200.deconvolution import Session
    with Session(shape, 2, 2, lib=gpu) as s:
        for v in range(2):
            s.set_view(v, imgs[v], ws[v], psfs[v])
        s.init()
        s.run(2)
        half = s.get_psi()
    with Session(shape, 2, 2, lib=gpu) as s:
        for v in range(2):
            s.set_view(v, imgs[v], ws[v], psfs[v])
        s.init()
        s.set_psi(half)
        s.run(2)
        s.finish()
        resumed = s.get_psi()
    assert np.array_equal(full, resumed)          # deterministic: bit-identical


def test_repeatability_bitwise(gpu):
    shape = (20, 24, 28)
    _, imgs, ws, psfs = synthetic.make_dataset(shape, 3, 5)
    a, *_ = P.run_session(gpu, imgs, ws, psfs, 2, 2, 3)
    b, *_ = P.run_session(gpu, imgs, ws, psfs, 2, 2, 3)
    assert np.array_equal(a, b)


# ---- the reference's Java class surface --------------------------------------------------------------
def test_host_mirror_gen2_classes(gpu):
    from spim_registration_b200 import MVDeconFFT, MVDeconInput, MVDeconvolution, PSFTYPE
    shape = (28, 32, 36)
    _, imgs, ws, psfs = synthetic.make_dataset(shape, 3, 7)
    views = MVDeconInput()
    for v in range(3):
        views.add(MVDeconFFT(imgs[v], ws[v], psfs[v], None, [0], False, None, False))
    d = MVDeconvolution(views, PSFTYPE.EFFICIENT_BAYESIAN, 3, 0.006, 1.0, 0, "deconvolved")
    ref = O.deconvolve(imgs, ws, psfs, O.DeconParams(iteration_type=O.EFFICIENT_BAYESIAN, num_iterations=3, gen=O.GEN2))
    per, l2 = O.parity_errors(d.getPsi(), ref.psi)
    assert per <= P.TOL_PER_VOXEL and l2 <= P.TOL_L2
    assert d.getCurrentIteration() == 3 and d.getName() == "deconvolved" and len(d.stats) == 9
    assert np.abs(views.getViews()[1].getKernel2() - ref.kernel2[1]).max() <= 5e-5 * ref.kernel2[1].max()
    d.close()


def test_host_mirror_gen1_classes(gpu):
    from spim_registration_b200 import LRFFT, LRInput, BayesMVDeconvolution, PSFTYPE
    shape = (28, 32, 36)
    _, imgs, ws, psfs = synthetic.make_dataset(shape, 3, 7)
    views = LRInput()
    for v in range(3):
        views.add(LRFFT(imgs[v], ws[v], psfs[v], [0], False, None))
    d = BayesMVDeconvolution(views, PSFTYPE.OPTIMIZATION_I, 3, 0.006, 1.0, 2, "deconvolved")
    ref = O.deconvolve(imgs, ws, psfs, O.DeconParams(iteration_type=O.OPTIMIZATION_I, num_iterations=3, gen=O.GEN1, osem_index=2))
    per, l2 = O.parity_errors(d.getPsi(), ref.psi)
    assert per <= P.TOL_PER_VOXEL and l2 <= P.TOL_L2
    assert np.isclose(d.getAvg(), ref.avg, rtol=1e-6)
    d.close()


def test_view_convolve1_convolve2_blocked_equals_unblocked(gpu):
    """LRFFT/MVDeconFFT.convolve1/2 through the legacy JNA entry, block-wise (copyBlock -> JNA ->
    pasteBlock) and as one single block: identical to the oracle's whole-image convolution."""
    from spim_registration_b200 import MVDeconFFT, MVDeconInput, PSFTYPE
    shape = (40, 44, 52)
    _, imgs, ws, psfs = synthetic.make_dataset(shape, 2, 7)
    rng = np.random.default_rng(0)
    psi = rng.random(shape, dtype=np.float32)
    for use_blocks, bs in ((False, None), (True, (24, 20, 16))):
        views = MVDeconInput()
        for v in range(2):
            views.add(MVDeconFFT(imgs[v], ws[v], psfs[v], None, [0], use_blocks, bs, False))
        views.init(PSFTYPE.EFFICIENT_BAYESIAN)
        v0 = views.getViews()[0]
        if use_blocks:
            assert len(v0.blocks) > 8
        c1 = v0.convolve1(psi)
        r1 = O.convolve(psi, v0.getKernel1(), O.EXT_MIRROR_SINGLE, dtype=np.float64)
        assert np.abs(c1 - r1).max() / np.abs(r1).max() < P.TOL_CONV
        c2 = v0.convolve2(psi)
        r2 = O.convolve(psi, v0.getKernel2(), O.EXT_CONSTANT, value=1.0, dtype=np.float64)
        assert np.abs(c2 - r2).max() / np.abs(r2).max() < P.TOL_CONV


# ---- full-size properties (BASELINE configs[1] size: 512 x 512 x 256, 31^3 PSF) --------------------------
FULL = (256, 512, 512)


def test_full_size_delta_kernel_is_identity_and_linearity(gpu):
    rng = np.random.default_rng(1)
    a = rng.random(FULL, dtype=np.float32)
    delta = np.zeros((31, 31, 31), np.float32)
    delta[15, 15, 15] = 1.0
    out = native.convolve(a, delta, O.EXT_MIRROR_SINGLE, lib=gpu)
    assert np.abs(out - a).max() < 2e-6
    # shifted delta == shift with mirror boundary
    sh = np.zeros((31, 31, 31), np.float32)
    sh[15, 15, 20] = 1.0
    out = native.convolve(a, sh, O.EXT_MIRROR_SINGLE, lib=gpu)
    assert np.abs(out[:, :, 5:] - a[:, :, :-5]).max() < 2e-6
    assert np.abs(out[:, :, 0] - a[:, :, 5]).max() < 2e-6          # mirror-single: x=-5 -> x=5
    # linearity
    b = rng.random(FULL, dtype=np.float32)
    k = synthetic.make_psf(31, 1, 7, 2.0)
    ca = native.convolve(a, k, O.EXT_MIRROR_SINGLE, lib=gpu)
    cb = native.convolve(b, k, O.EXT_MIRROR_SINGLE, lib=gpu)
    cab = native.convolve(a + 2 * b, k, O.EXT_MIRROR_SINGLE, lib=gpu)
    assert np.abs(cab - (ca + 2 * cb)).max() < 2e-5
    # periodic extension preserves the sum: sum(conv) = sum(img) * sum(k)
    cp = native.convolve(a, k, O.EXT_PERIODIC, lib=gpu)
    assert abs(cp.sum(dtype=np.float64) / (a.sum(dtype=np.float64) * k.sum(dtype=np.float64)) - 1) < 1e-6


def test_full_size_view_step_spot_check_against_oracle(gpu):
    """One iteration of 2 views at 512x512x256; sub-bricks of psi recomputed by the oracle from a
    crop with a 2*(31//2)-voxel halo (two convolutions deep), at the interior and at a corner."""
    V = 2
    _, imgs, ws, psfs = synthetic.make_dataset(FULL, V, 31, kind="specimen")
    psi, k1, k2, st, (avg, *_r) = P.run_session(gpu, imgs, ws, psfs, O.EFFICIENT_BAYESIAN, 2, 1)
    h = 30 * V      # each view-step reaches 2 * (31 // 2) voxels further
    for (z0, y0, x0), interior in (((100, 200, 300), True), ((0, 0, 0), False)):
        n = 48
        lo = [max(0, c - h) for c in (z0, y0, x0)]
        hi = [min(s, c + n + h) for c, s in zip((z0, y0, x0), FULL)]
        sl = tuple(slice(a, b) for a, b in zip(lo, hi))
        ci = [im[sl] for im in imgs]
        cw = [w[sl] for w in ws]
        p = O.DeconParams(iteration_type=O.EFFICIENT_BAYESIAN, num_iterations=1, gen=O.GEN2, mask_at_end=True,
                          psi_init=np.full(ci[0].shape, np.float32(avg), np.float32))
        ref = O.deconvolve(ci, cw, psfs, p)
        off = [c - l for c, l in zip((z0, y0, x0), lo)]
        rs = tuple(slice(o, o + n) for o in off)
        gs = tuple(slice(c, c + n) for c in (z0, y0, x0))
        per, l2 = O.parity_errors(psi[gs], ref.psi[rs])
        assert per <= P.TOL_PER_VOXEL and l2 <= P.TOL_L2, (interior, per, l2)


# ---- brick mode on one GPU ---------------------------------------------------------------------------
def test_brick_mode_without_neighbours_equals_plain_session(gpu):
    """A haloed (brick-mode) session whose every face is a volume face must reproduce the plain session:
    the loader applies the boundary rule itself on faces without a neighbour."""
    from spim_registration_b200.deconvolution import Session
    shape, V = (20, 26, 30), 2
    _, imgs, ws, psfs = synthetic.make_dataset(shape, V, 7)
    plain, *_ = P.run_session(gpu, imgs, ws, psfs, 2, 2, 2)
    with Session(shape, V, 2, generation=2, haloed=True, lib=gpu) as s:
        for v in range(V):
            s.set_view(v, imgs[v], ws[v], psfs[v])
        s.init()
        part = s.init_partials()
        s.set_avg(part[0] / part[1], 1.0)
        s.set_halo_mask(0, 0)
        for _ in range(2):
            for v in range(V):
                s.view_phase(v, 0)
                s.view_phase(v, 1)
        s.finish()
        psi = s.get_psi()
    assert np.array_equal(plain, psi)


def test_halo_pack_unpack_roundtrip(gpu):
    import torch
    from spim_registration_b200.deconvolution import Session
    shape, V = (12, 14, 16), 1
    _, imgs, ws, psfs = synthetic.make_dataset(shape, V, 5)
    with Session(shape, V, 3, generation=2, haloed=True, lib=gpu) as s:
        s.set_view(0, imgs[0], ws[0], psfs[0])
        s.init()
        s.set_avg(0.5, 1.0)
        ptr, dims, origin = s.device_buffer(0)

        class _CAI:
            pass
        o = _CAI()
        o.__cuda_array_interface__ = {"shape": tuple(dims), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
        t = torch.as_tensor(o, device="cuda:0")
        rng = torch.Generator(device="cuda:0").manual_seed(1)
        t.copy_(torch.rand(tuple(dims), generator=rng, device="cuda:0"))
        regions = [(2, 0, 3, 4, dims[1], 2), (0, 5, 0, dims[0], 3, dims[2]), (1, 1, 1, 1, 1, 1)]
        n = sum(r[3] * r[4] * r[5] for r in regions)
        flat = torch.zeros(n, device="cuda:0")
        s.halo_pack(0, regions, flat.data_ptr())
        s.sync()
        want = torch.cat([t[r[0]:r[0] + r[3], r[1]:r[1] + r[4], r[2]:r[2] + r[5]].reshape(-1) for r in regions])
        assert torch.equal(flat, want)
        before = t.clone()
        flat2 = torch.rand(n, device="cuda:0")
        s.halo_pack(0, regions[:1], flat2.data_ptr(), unpack=True)
        s.sync()
        r = regions[0]
        assert torch.equal(t[r[0]:r[0] + r[3], r[1]:r[1] + r[4], r[2]:r[2] + r[5]].reshape(-1), flat2[:r[3] * r[4] * r[5]])
        mask = torch.ones_like(t, dtype=torch.bool)
        mask[r[0]:r[0] + r[3], r[1]:r[1] + r[4], r[2]:r[2] + r[5]] = False
        assert torch.equal(t[mask], before[mask])


# ---- randomized sweeps on the GPU (same seeds as the emulator tests) ------------------------------------------
def test_randomized_convolutions(gpu):
    rng = np.random.default_rng(12345)
    for _ in range(120):
        shape = tuple(int(rng.integers(1, 24)) for _ in range(3))
        ks = tuple(int(rng.integers(1, 10)) for _ in range(3))
        ext = int(rng.integers(0, 5))
        P.conv_case(gpu, shape, ks, ext, seed=int(rng.integers(0, 1 << 30)))
        if all(k <= s for k, s in zip(ks, shape)):
            P.legacy_case(gpu, shape, ks, seed=int(rng.integers(0, 1 << 30)))


def test_randomized_deconvolutions(gpu):
    rng = np.random.default_rng(777)
    for _ in range(12):
        shape = tuple(int(rng.integers(5, 16)) for _ in range(3))
        V = int(rng.integers(1, 5))
        ks = int(rng.choice([3, 5, 7]))
        typ = int(rng.integers(0, 4))
        gen = int(rng.integers(1, 3))
        P.decon_case(gpu, shape, V, ks, typ, gen, int(rng.integers(1, 4)), lam=float(rng.choice([0.0, 0.006, 0.06])),
                     weight_mode=str(rng.choice(["normalized", "blending", "ones"])), use_weights=bool(rng.integers(0, 2)),
                     osem_index=int(rng.integers(0, 3)) if gen == 1 else 0, seed=int(rng.integers(0, 1000)))
