"""Oracle self-checks: the reference holds no golden vectors for this path
(parity unpinned), so the oracle is pinned by mathematical identities and known closed forms."""
import numpy as np
import pytest

from oracle import mvdecon_oracle as O

RNG = np.random.default_rng(42)


def _rand(shape, dtype=np.float64):
    return RNG.random(shape).astype(dtype)


@pytest.mark.parametrize("ext", [O.EXT_ZERO, O.EXT_CONSTANT, O.EXT_MIRROR_SINGLE, O.EXT_MIRROR_DOUBLE, O.EXT_PERIODIC])
@pytest.mark.parametrize("kshape", [(3, 3, 3), (5, 3, 1), (2, 4, 3)])
def test_fft_conv_equals_direct_sum(ext, kshape):
    img = _rand((7, 6, 9))
    k = _rand(kshape)
    a = O.convolve(img, k, ext, value=1.0, dtype=np.float64)
    b = O.convolve_direct(img, k, ext, value=1.0)
    assert np.abs(a - b).max() < 1e-12


def test_convolution_not_correlation_and_centre():
    img = np.zeros((9, 9, 9))
    img[4, 4, 4] = 1.0
    k = _rand((3, 5, 3))
    out = O.convolve(img, k, O.EXT_ZERO, dtype=np.float64)
    # a delta at p0 reproduces the kernel with its centre (dim//2) at p0, un-flipped
    assert np.allclose(out[3:6, 2:7, 3:6], k, atol=1e-13)


def test_mirror_single_does_not_repeat_edge():
    a = np.arange(5.0).reshape(1, 1, 5)
    e = O.extend(a, (0, 0, 2), (0, 0, 2), O.EXT_MIRROR_SINGLE)
    assert e.ravel().tolist() == [2, 1, 0, 1, 2, 3, 4, 3, 2]
    e = O.extend(a, (0, 0, 2), (0, 0, 2), O.EXT_MIRROR_DOUBLE)
    assert e.ravel().tolist() == [1, 0, 0, 1, 2, 3, 4, 4, 3]
    e = O.extend(a, (0, 0, 2), (0, 0, 1), O.EXT_CONSTANT, 7.0)
    assert e.ravel().tolist() == [7, 7, 0, 1, 2, 3, 4, 7]


@pytest.mark.parametrize("ext,val", [(O.EXT_MIRROR_SINGLE, 0.0), (O.EXT_CONSTANT, 1.0)])
def test_blocked_equals_unblocked(ext, val):
    img = _rand((20, 17, 23))
    k = _rand((5, 5, 7))
    whole = O.convolve(img, k, ext, value=val, dtype=np.float64)
    blocked = O.convolve_blocked(img, k, (12, 11, 16), ext, value=val, dtype=np.float64)
    assert np.abs(whole - blocked).max() < 1e-12


def test_periodic_pad_equals_circular():
    img = _rand((10, 12, 14))
    k = _rand((3, 5, 3))
    circ = O.circular_convolve(img, k, dtype=np.float64)
    per = O.convolve(img, k, O.EXT_PERIODIC, dtype=np.float64)
    assert np.abs(circ - per).max() < 1e-12


def test_block_geometry_matches_reference_rule():
    # C3 with 512^3 blocks: eff 482^3 -> 3x3x2 = 18 blocks
    blocks = O.divide_into_blocks((512, 1024, 1024), (512, 512, 512), (31, 31, 31))
    assert len(blocks) == 2 * 3 * 3
    b0 = blocks[0]
    assert b0.offset == (-15, -15, -15) and b0.effective_size == (482, 482, 482) and b0.effective_local_offset == (15, 15, 15)
    last = blocks[-1]
    assert last.effective_offset == (482, 964, 964) and last.effective_size == (30, 60, 60)
    assert O.divide_into_blocks((64, 64, 64), (16, 16, 16), (31, 31, 31), gen=O.GEN2) is None
    g1 = O.divide_into_blocks((64, 64, 64), (16, 16, 16), (31, 31, 31), gen=O.GEN1)
    assert g1[0].block_size == (32, 32, 32)


def _psfs(V, size):
    from spim_registration_b200 import synthetic
    return synthetic.make_psfs(V, size)


def test_kernel2_identities():
    psfs = _psfs(3, 7)
    k1, k2 = O.init_kernels(psfs, O.INDEPENDENT)
    for a, b in zip(k1, k2):
        assert np.array_equal(b, a[::-1, ::-1, ::-1])
        assert abs(float(a.sum(dtype=np.float64)) - 1) < 1e-6
    # one view: every type equals INDEPENDENT
    for t in range(4):
        _, k2s = O.init_kernels(psfs[:1], t)
        assert np.array_equal(k2s[0], O.init_kernels(psfs[:1], O.INDEPENDENT)[1][0])
    # OPTIMIZATION_II: K2 = flip(K1^V / sum)
    k1, k2 = O.init_kernels(psfs, O.OPTIMIZATION_II)
    e = k1[1].astype(np.float64) ** 3
    e /= e.sum()
    assert np.allclose(k2[1], e[::-1, ::-1, ::-1], rtol=1e-5, atol=1e-12)
    for t in (O.EFFICIENT_BAYESIAN, O.OPTIMIZATION_I):
        _, k2t = O.init_kernels(psfs, t)
        for k in k2t:
            assert abs(float(k.sum(dtype=np.float64)) - 1) < 1e-5
            assert k.min() >= 0


def test_efficient_bayesian_compound_kernel_formula():
    # K2_v = norm( flip(K1_v) * prod_w [ (flip(K1_v) conv K1_w) conv flip(K1_w) ] ), zero-extended, K1_v-sized
    psfs = _psfs(2, 5)
    k1, k2 = O.init_kernels(psfs, O.EFFICIENT_BAYESIAN)
    f = lambda a: a[::-1, ::-1, ::-1]
    c = O.convolve_direct(O.convolve_direct(f(k1[0]), k1[1], O.EXT_ZERO), f(k1[1]), O.EXT_ZERO)
    t = f(k1[0]).astype(np.float64) * c
    t /= t.sum()
    assert np.allclose(k2[0], t, rtol=2e-4, atol=1e-9)


def test_mirror_quirk_even_sizes():
    a = np.arange(4.0).reshape(1, 1, 4)
    assert O.mirror_quirk(a).ravel().tolist() == [3, 1, 2, 0]     # FD/Mirror.java:93 double-swaps the middle pair
    b = np.arange(5.0).reshape(1, 1, 5)
    assert O.mirror_quirk(b).ravel().tolist() == [4, 3, 2, 1, 0]


def test_tikhonov_closed_form_table():
    # MVDeconvolution.main (FD/MVDeconvolution.java:728-735): tikhonov(d, 0.0006)
    for d in np.arange(0, 10, 0.1):
        for v in (d, d * 10000):
            assert np.isclose(O.tikhonov(v, 0.0006), (np.sqrt(1 + 2 * 0.0006 * v) - 1) / 0.0006)
    # small-lambda limit is the identity
    assert np.isclose(O.tikhonov(3.0, 1e-9), 3.0, rtol=1e-6)


def test_update_rules():
    psi = np.array([1.0, 2.0, 0.5, 1.0], dtype=np.float32)
    integ = np.array([2.0, -1.0, np.nan, 1.0], dtype=np.float32)
    w = np.array([1.0, 1.0, 1.0, 0.25], dtype=np.float32)
    new, s, m = O.compute_final_values(psi, integ, w, 0.0)
    assert new[0] == 2.0                     # plain multiplicative update
    # value <= 0 / NaN -> minValue; psi + (minValue - psi) * 1 is evaluated in fp32 like the Java code
    assert new[1] == np.float32(2.0) + (O.MIN_VALUE - np.float32(2.0)) and np.isclose(new[1], 1e-4, rtol=1e-3)
    assert np.isclose(new[2], 1e-4, rtol=1e-3)
    assert new[3] == 1.0                     # value == psi: no change
    new2, _, _ = O.compute_final_values(psi[:1], integ[:1], np.float32(0.5), 0.0)
    assert new2[0] == 1.5                    # weight scales the change
    q1 = O.compute_quotient(np.array([0.0, 2.0], np.float32), np.array([4.0, 4.0], np.float32), O.GEN1)
    q2 = O.compute_quotient(np.array([0.0, 2.0], np.float32), np.array([4.0, 4.0], np.float32), O.GEN2)
    assert q1.tolist() == [0.0, 0.5] and q2.tolist() == [1.0, 0.5]


def test_single_view_is_classic_richardson_lucy():
    from spim_registration_b200 import synthetic
    shape = (12, 14, 16)
    truth = synthetic.bead_truth(shape, 5, seed=3).astype(np.float64)
    psf = synthetic.make_psf(5, 0, 1).astype(np.float64)
    img = O.convolve(truth, psf, O.EXT_MIRROR_SINGLE, dtype=np.float64)
    p = O.DeconParams(iteration_type=O.INDEPENDENT, num_iterations=3, lam=0.0, gen=O.GEN1, dtype=np.float64,
                      conv2_ext=O.EXT_MIRROR_SINGLE)
    res = O.deconvolve([img], [np.ones(shape, np.float32)], [psf], p)
    # hand-rolled RL
    psi = np.full(shape, np.float32(1.0), dtype=np.float64)   # gen-1 avg with <2 views = 1
    k = (psf / psf.sum()).astype(np.float32).astype(np.float64)
    for _ in range(3):
        blur = O.convolve(psi, k, O.EXT_MIRROR_SINGLE, dtype=np.float64)
        psi = np.maximum(1e-4, psi * O.convolve(img / blur, k[::-1, ::-1, ::-1], O.EXT_MIRROR_SINGLE, dtype=np.float64))
    assert np.allclose(res.psi, psi, rtol=1e-5)


def test_fixed_point_noise_free():
    from spim_registration_b200 import synthetic
    shape = (14, 14, 14)
    truth = (synthetic.bead_truth(shape, 6, seed=5) / 100.0).astype(np.float64)
    psf = synthetic.make_psf(5, 0, 1).astype(np.float64)
    k1 = O.norm_image(psf).astype(np.float64)
    img = O.convolve(truth, k1, O.EXT_MIRROR_SINGLE, dtype=np.float64)
    p = O.DeconParams(iteration_type=O.INDEPENDENT, num_iterations=2, lam=0.0, gen=O.GEN2, dtype=np.float64,
                      psi_init=truth, mask_at_end=False)
    res = O.deconvolve([img], [np.ones(shape, np.float32)], [psf], p)
    assert np.abs(res.psi - truth).max() / truth.max() < 1e-6


def test_psi_init_rules():
    a = np.array([[[0.0, 2.0, 4.0]]], dtype=np.float32)
    b = np.array([[[0.0, 0.0, 2.0]]], dtype=np.float32)
    avg, cnt = O.fuse_first_iteration_gen2([a, b])
    assert cnt.ravel().tolist() == [0, 1, 2] and np.isclose(avg, (2.0 + 3.0) / 2)
    w1 = np.array([[[1.0, 1.0, 1.0]]], dtype=np.float32)
    w2 = np.array([[[0.0, 0.0, 1.0]]], dtype=np.float32)
    avg1, mn, av = O.norm_all_images_gen1([a, b], [w1, w2])
    assert np.isclose(avg1, (4.0 + 2.0) / 2) and mn == 1 and np.isclose(av, 4 / 3)


def test_fp32_mode_tracks_fp64_truth():
    from spim_registration_b200 import synthetic
    shape = (20, 20, 20)
    _, imgs, ws, psfs = synthetic.make_dataset(shape, 3, 7, kind="beads")
    r32 = O.deconvolve(imgs, ws, psfs, O.DeconParams(num_iterations=5, dtype=np.float32))
    r64 = O.deconvolve(imgs, ws, psfs, O.DeconParams(num_iterations=5, dtype=np.float64))
    per, l2 = O.parity_errors(r32.psi, r64.psi)
    assert per < 1e-3 and l2 < 1e-4


@pytest.mark.parametrize("kshape", [(3, 3, 3), (5, 3, 7), (4, 2, 6)])
def test_convolution_against_independent_library(kshape):
    """The oracle's convolution definition (kernel origin at dim//2, out-of-bounds rules) against
    scipy.ndimage.convolve, an independent implementation: ndimage 'mirror' = mirror-single, 'reflect' =
    mirror-double, 'wrap' = periodic, 'constant' = constant value; ndimage.convolve with origin 0 places the
    kernel origin at dim//2 for odd and even sizes alike, which is the reference's convention."""
    import scipy.ndimage as ndi
    img = _rand((9, 8, 11))
    k = _rand(kshape)
    origin = 0
    for ext, mode, cval in [(O.EXT_MIRROR_SINGLE, "mirror", 0.0), (O.EXT_MIRROR_DOUBLE, "reflect", 0.0),
                            (O.EXT_PERIODIC, "wrap", 0.0), (O.EXT_CONSTANT, "constant", 1.0), (O.EXT_ZERO, "constant", 0.0)]:
        want = ndi.convolve(img, k, mode=mode, cval=cval, origin=origin)
        got = O.convolve(img, k, ext, value=cval, dtype=np.float64)
        assert np.abs(got - want).max() < 1e-12, (ext, kshape)


def test_richardson_lucy_against_independent_formula():
    """One gen-1 view-step with w = 1, lambda = 0 written out with scipy.ndimage only."""
    import scipy.ndimage as ndi
    rng = np.random.default_rng(3)
    img = rng.random((8, 9, 10)) + 0.1
    psf = rng.random((3, 3, 3)).astype(np.float32)     # kernels are float images in the reference
    psi0 = np.full(img.shape, 0.7)
    p = O.DeconParams(iteration_type=O.INDEPENDENT, num_iterations=1, lam=0.0, gen=O.GEN1, dtype=np.float64,
                      psi_init=psi0)
    k1f = O.norm_image(psf).astype(np.float64)        # the oracle normalises in fp64 and stores fp32
    blur = ndi.convolve(psi0, k1f, mode="mirror")
    want = np.maximum(1e-4, psi0 * ndi.convolve(img / blur, k1f[::-1, ::-1, ::-1], mode="mirror"))
    got = O.deconvolve([img], [np.ones(img.shape, np.float32)], [psf], p).psi
    assert np.abs(got - want).max() < 1e-10
