"""Index-math tests of the kernel bodies under the CPU emulator (tests/emu, built from the very
same sources with -DSPIM_HOST_EMU).  These are host-logic tests: they validate tiling, digit
reversal, extension-on-load, the R2C/C2R split steps and the fused epilogues before any GPU time is
spent.  The emulator is test infrastructure only and is never loaded by the package."""
import numpy as np
import pytest

import parity_cases as P
from oracle import mvdecon_oracle as O


def test_emulator_identifies_itself(emu_lib):
    assert b"EMULATOR" in emu_lib.mvd_version()


@pytest.mark.parametrize("ext", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("shape,kshape", [((8, 8, 8), (3, 3, 3)), ((9, 7, 11), (3, 5, 3)), ((6, 6, 6), (4, 2, 6)),
                                          ((5, 30, 33), (1, 7, 9)), ((1, 16, 18), (1, 5, 5))])
def test_conv_all_extensions(emu_lib, shape, kshape, ext):
    P.conv_case(emu_lib, shape, kshape, ext)


def test_conv_halo_wider_than_image(emu_lib):
    # multiple reflections / wraps: kernel larger than the image along an axis
    for ext in (2, 3, 4):
        P.conv_case(emu_lib, (3, 4, 6), (7, 9, 5), ext)


@pytest.mark.parametrize("n", [16, 18, 20, 22, 24, 26, 28, 30, 36, 40, 42, 44, 48, 50, 52, 54, 56, 60, 64, 66, 70, 72, 78, 80,
                               84, 88, 90, 96, 98, 100, 104, 108, 110, 112, 120, 126, 128, 130, 132, 140, 144])
def test_every_radix_path_along_each_axis(emu_lib, n):
    # exact periodic mode keeps P = n, so the stage plan of n (and n/2 on x) is what runs
    rng = np.random.default_rng(n)
    k = rng.random((3, 3, 3), dtype=np.float32)
    for shape in ((n, 4, 8), (4, n, 8), (4, 4, n)):
        P.legacy_case(emu_lib, shape, (3, 3, 3), seed=n)


@pytest.mark.parametrize("shape,kshape", [((16, 16, 16), (5, 5, 5)), ((12, 20, 18), (3, 7, 5)), ((10, 9, 7), (3, 3, 3)),
                                          ((8, 8, 8), (8, 8, 8))])
def test_legacy_entry_is_circular_convolution(emu_lib, shape, kshape):
    P.legacy_case(emu_lib, shape, kshape)


def test_golden_conv(emu_lib):
    P.golden_conv_case(emu_lib)


@pytest.mark.parametrize("gen", [1, 2])
@pytest.mark.parametrize("typ", [0, 1, 2, 3])
def test_golden_deconvolution(emu_lib, gen, typ):
    P.golden_case(emu_lib, gen, typ)


@pytest.mark.parametrize("gen,typ", [(2, O.EFFICIENT_BAYESIAN), (1, O.OPTIMIZATION_I)])
def test_deconvolution_vs_oracle(emu_lib, gen, typ):
    P.decon_case(emu_lib, (14, 18, 22), 3, 5, typ, gen, 3)


def test_deconvolution_odd_dims_no_weights_no_tikhonov(emu_lib):
    P.decon_case(emu_lib, (9, 11, 13), 2, 3, O.INDEPENDENT, 2, 2, lam=0.0, use_weights=False)


def test_gen1_osem_from_overlap(emu_lib):
    P.decon_case(emu_lib, (10, 12, 14), 3, 3, O.OPTIMIZATION_II, 1, 2, osem_index=2)
    P.decon_case(emu_lib, (10, 12, 14), 3, 3, O.OPTIMIZATION_II, 1, 2, osem=2.0, osem_index=0)


def test_views_uploaded_in_cells(emu_lib):
    P.cells_case(emu_lib)


def test_second_init_with_larger_psf_recreates_buffers(emu_lib):
    P.reinit_case(emu_lib)


def test_fast_epilogue_switch(emu_lib):
    P.fast_epilogue_case(emu_lib, bit_identical=True)


def test_exact_tikhonov_switch(emu_lib):
    P.exact_tikhonov_case(emu_lib)


def test_error_paths(emu_lib):
    import ctypes
    from spim_registration_b200 import native
    from spim_registration_b200.deconvolution import Session
    with pytest.raises(native.NativeError):
        Session((4, 4, 4), 0, 2, lib=emu_lib)
    with pytest.raises(native.NativeError):
        Session((4, 4, 4), 1, 7, lib=emu_lib)
    with Session((4, 4, 4), 2, 2, lib=emu_lib) as s:
        s.set_view(0, np.ones((4, 4, 4), np.float32), None, np.ones((3, 3, 3), np.float32))
        with pytest.raises(native.NativeError, match="view 1 not set"):
            s.init()
        with pytest.raises(native.NativeError, match="not initialised"):
            s.run(1)
        with pytest.raises(ValueError):
            s.set_view(1, np.ones((4, 4, 5), np.float32), None, np.ones((3, 3, 3), np.float32))


# ---- edge cases and error behaviour (host logic, no GPU) -------------------------------------------------
def test_degenerate_and_ragged_shapes(emu_lib):
    # 2-D data (one plane), one-voxel-thick axes, kernels with unit axes, tiny images
    for shape, kshape in [((1, 1, 8), (1, 1, 3)), ((1, 9, 1), (1, 3, 1)), ((2, 2, 2), (1, 1, 1)), ((1, 1, 2), (1, 1, 1)),
                          ((3, 1, 5), (3, 1, 3)), ((2, 3, 4), (2, 3, 4))]:
        for ext in (0, 1, 2, 3, 4):
            P.conv_case(emu_lib, shape, kshape, ext)


def test_single_view_and_many_views(emu_lib):
    P.decon_case(emu_lib, (8, 9, 10), 1, 3, O.EFFICIENT_BAYESIAN, 2, 2)      # V = 1: every type degenerates to INDEPENDENT
    P.decon_case(emu_lib, (6, 7, 8), 9, 3, O.OPTIMIZATION_II, 2, 1)           # K1^9 by repeated fp32 multiplication


def test_views_without_any_data_and_psi_nan_rule(emu_lib):
    """all-zero images: gen-2 average is NaN -> 0.5 (MVDeconvolution.java:138-142) and psi is masked to 0."""
    from spim_registration_b200.deconvolution import Session
    shape = (6, 6, 6)
    z = np.zeros(shape, np.float32)
    k = np.ones((3, 3, 3), np.float32)
    with Session(shape, 2, 2, generation=2, lib=emu_lib) as s:
        s.set_view(0, z, None, k)
        s.set_view(1, z, None, k)
        s.init()
        assert s.info().avg == 0.5
        s.run(1)
        s.finish()
        assert np.all(s.get_psi() == 0)


def test_limits_and_argument_validation(emu_lib):
    import ctypes
    from spim_registration_b200 import native
    from spim_registration_b200.deconvolution import Session
    with pytest.raises(native.NativeError, match="num_views"):
        Session((4, 4, 4), 65, 2, lib=emu_lib)                       # MAX_VIEWS = 64
    with pytest.raises(native.NativeError, match="generation"):
        Session((4, 4, 4), 1, 2, generation=3, lib=emu_lib)
    with pytest.raises(native.NativeError, match="dims"):
        Session((0, 4, 4), 1, 2, lib=emu_lib)
    # an FFT axis that does not fit one shared-memory tile is refused with a message, not mis-computed
    with Session((1, 1, 3000), 1, 3, lib=emu_lib) as s:
        s.set_view(0, np.ones((1, 1, 3000), np.float32), None, np.ones((1, 1, 3), np.float32))
        s.init()                                                     # x axis: Px/2 ~ 1512 rows of 128 B -> fits in 227 KB
    with Session((1, 2000, 4), 1, 3, lib=emu_lib) as s:
        s.set_view(0, np.ones((1, 2000, 4), np.float32), None, np.ones((1, 3, 1), np.float32))
        s.init()                                                     # y axis: 2016 rows -> narrow (8-column) tiles of 129 KB
        s.run(1, stats=False)
        want = 2.0 / (1.0 + np.sqrt(1.0 + 2 * 0.006))               # all-ones views: psi stays 1 up to the Tikhonov step
        assert np.abs(s.get_psi() - want).max() < 1e-5
    with Session((1, 4000, 4), 1, 3, lib=emu_lib) as s:
        s.set_view(0, np.ones((1, 4000, 4), np.float32), None, np.ones((1, 3, 1), np.float32))
        with pytest.raises(native.NativeError, match="too long"):
            s.init()                                                 # not even a narrow tile fits: refused with a message
    # struct_size guards the ABI
    p = native.MvdParams()
    emu_lib.mvd_params_default(ctypes.byref(p))
    p.struct_size = 4
    h = ctypes.c_void_p()
    assert emu_lib.mvd_session_create(ctypes.byref(p), ctypes.byref(h)) != 0
    assert b"struct_size" in emu_lib.mvd_last_error()
    # brick-only entry points refuse plain sessions
    with Session((4, 4, 4), 1, 3, lib=emu_lib) as s:
        s.set_view(0, np.ones((4, 4, 4), np.float32), None, np.ones((3, 3, 3), np.float32))
        s.init()
        with pytest.raises(native.NativeError, match="brick"):
            s.set_halo_mask(1, 1)
        with pytest.raises(native.NativeError):
            s.view_phase(5, 0)


def test_legacy_entry_argument_checks(emu_lib):
    from spim_registration_b200 import native
    im = np.ones((4, 4, 4), np.float32)
    k = np.ones((3, 3, 3), np.float32)
    emu_lib.convolution3DfftCUDAInPlace(im.ctypes.data_as(native.c_float_p), native.int3((4, 4, 0)),
                                        k.ctypes.data_as(native.c_float_p), native.int3((3, 3, 3)), 0)
    assert np.all(im == 1) and b"dims" in emu_lib.spim_fftconv_last_error()     # void entry: buffer untouched + message


def test_block_mirror_classes_match_oracle():
    """Block / BlockGeneratorFixedSizePrecise (host logic) against the oracle's restatement."""
    from spim_registration_b200 import BlockGeneratorFixedSizePrecise
    from spim_registration_b200.blocks import divide_into_blocks
    img_xyz, blk_xyz, k_xyz = (23, 17, 20), (16, 11, 12), (7, 5, 5)
    blocks = BlockGeneratorFixedSizePrecise(blk_xyz).divideIntoBlocks(img_xyz, k_xyz)
    ref = O.divide_into_blocks(img_xyz[::-1], blk_xyz[::-1], k_xyz[::-1])
    assert len(blocks) == len(ref)
    key = lambda t: tuple(t)
    assert sorted(key(b.getOffset()[::-1]) for b in blocks) == sorted(key(r.offset) for r in ref)
    assert sorted(key(b.getEffectiveSize()[::-1]) for b in blocks) == sorted(key(r.effective_size) for r in ref)
    assert BlockGeneratorFixedSizePrecise((4, 4, 4)).divideIntoBlocks(img_xyz, k_xyz) is None        # gen-2: too small -> null
    assert divide_into_blocks(img_xyz, (4, 4, 4), k_xyz, double_too_small=True)[0].blockSize == (8, 8, 8)   # gen-1 doubles
    rng = np.random.default_rng(0)
    src = rng.random(img_xyz[::-1], dtype=np.float32)
    out = np.zeros_like(src)
    buf = np.empty(blk_xyz[::-1], np.float32)
    for b in blocks:                       # copy (mirror extension) + paste of the effective region = identity
        b.copyBlock(src, buf, 2)
        b.pasteBlock(out, buf)
    assert np.array_equal(out, src)


def test_device_query_mirrors_without_gpu(cuda_lib):
    from spim_registration_b200 import CUDATools, NativeLibraryTools, native
    c = native.CUDAFourierConvolution()
    if c.getNumDevicesCUDA() <= 0:
        assert CUDATools.queryCUDADetails(c) is None                  # -1 / 0 devices -> null like the reference
    lib = NativeLibraryTools.loadNativeLibrary()
    assert lib is not None and callable(lib.convolution3DfftCUDAInPlace)
    assert NativeLibraryTools.loadNativeLibrary(directory="/nonexistent") is None


# ---- randomized sweeps (fixed seeds): ragged shapes, kernels larger than the image, every extension rule ------
def test_randomized_convolutions(emu_lib):
    from spim_registration_b200 import native
    rng = np.random.default_rng(12345)
    for _ in range(120):
        shape = tuple(int(rng.integers(1, 24)) for _ in range(3))
        ks = tuple(int(rng.integers(1, 10)) for _ in range(3))
        ext = int(rng.integers(0, 5))
        P.conv_case(emu_lib, shape, ks, ext, seed=int(rng.integers(0, 1 << 30)))
        if all(k <= s for k, s in zip(ks, shape)):
            P.legacy_case(emu_lib, shape, ks, seed=int(rng.integers(0, 1 << 30)))


def test_randomized_deconvolutions(emu_lib):
    rng = np.random.default_rng(777)
    for _ in range(12):
        shape = tuple(int(rng.integers(5, 16)) for _ in range(3))
        V = int(rng.integers(1, 5))
        ks = int(rng.choice([3, 5, 7]))
        typ = int(rng.integers(0, 4))
        gen = int(rng.integers(1, 3))
        P.decon_case(emu_lib, shape, V, ks, typ, gen, int(rng.integers(1, 4)), lam=float(rng.choice([0.0, 0.006, 0.06])),
                     weight_mode=str(rng.choice(["normalized", "blending", "ones"])), use_weights=bool(rng.integers(0, 2)),
                     osem_index=int(rng.integers(0, 3)) if gen == 1 else 0, seed=int(rng.integers(0, 1000)))


@pytest.mark.parametrize("ks", [3, 5, 7])
def test_brick_mode_without_neighbours_equals_plain_session(emu_lib, ks):
    """A haloed (brick-mode) session whose every face is a volume face reproduces the plain session bit for bit; odd
    PSF/2 halos (ks = 3, 7) exercise the even-x-origin padding of the haloed buffers."""
    from spim_registration_b200 import synthetic
    from spim_registration_b200.deconvolution import Session
    shape, V = (10, 12, 14), 2
    _, imgs, ws, psfs = synthetic.make_dataset(shape, V, ks)
    plain, *_ = P.run_session(emu_lib, imgs, ws, psfs, 2, 2, 2)
    with Session(shape, V, 2, generation=2, haloed=True, lib=emu_lib) as s:
        for v in range(V):
            s.set_view(v, imgs[v], ws[v], psfs[v])
        s.init()
        part = s.init_partials()
        s.set_avg(part[0] / part[1], 1.0)
        s.set_halo_mask(0, 0)
        ptr, dims, origin = s.device_buffer(0)
        assert origin[2] % 2 == 0 and dims[2] % 2 == 0 and origin[2] >= s.info().halo_lo[2]
        for _ in range(2):
            for v in range(V):
                s.view_phase(v, 0)
                s.view_phase(v, 1)
        s.finish()
        psi = s.get_psi()
    assert np.array_equal(plain, psi)


# ---- narrow column tiles (8 columns / 64-byte rows; automatic for FFT lengths above ~880) ----------------------
@pytest.mark.parametrize("n", [16, 18, 20, 28, 30, 36, 42, 48, 50, 54, 56, 60, 70, 72, 80, 90, 96, 100, 112, 126, 140, 144])
def test_narrow_tiles_every_radix_path(emu_lib, monkeypatch, n):
    monkeypatch.setenv("SPIM_COL_NARROW", "1")
    for shape in ((n, 4, 8), (4, n, 8)):
        P.legacy_case(emu_lib, shape, (3, 3, 3), seed=n)


@pytest.mark.parametrize("ext", [0, 1, 2, 3, 4])
def test_narrow_tiles_conv_all_extensions(emu_lib, monkeypatch, ext):
    monkeypatch.setenv("SPIM_COL_NARROW", "1")
    P.conv_case(emu_lib, (9, 7, 11), (3, 5, 3), ext)
    P.conv_case(emu_lib, (5, 30, 33), (1, 7, 9), ext)      # pitch 32: two 16-column tiles = four narrow ones


def test_narrow_tiles_deconvolution(emu_lib, monkeypatch):
    monkeypatch.setenv("SPIM_COL_NARROW", "1")
    c0 = emu_lib.mvd_debug_counter(0)
    P.decon_case(emu_lib, (14, 18, 22), 3, 5, O.EFFICIENT_BAYESIAN, 2, 3)
    P.golden_case(emu_lib, 1, 1)
    P.golden_conv_case(emu_lib)
    assert emu_lib.mvd_debug_counter(0) > c0


def test_long_axes_select_narrow_tiles_automatically(emu_lib):
    # 1080-long y / z axes (a 1024^2 x 512 volume with a 31^3 PSF on one GPU): a 16-column tile would be 138 KB
    c0 = emu_lib.mvd_debug_counter(0)
    P.conv_case(emu_lib, (3, 1050, 6), (3, 31, 3), 2)
    c1 = emu_lib.mvd_debug_counter(0)
    assert c1 > c0                       # the long y passes ran on narrow tiles ...
    P.conv_case(emu_lib, (1040, 2, 6), (15, 1, 3), 1)
    assert emu_lib.mvd_debug_counter(0) > c1
    c2 = emu_lib.mvd_debug_counter(0)
    P.conv_case(emu_lib, (8, 8, 8), (3, 3, 3), 2)
    assert emu_lib.mvd_debug_counter(0) == c2     # ... and short axes stay on the 16-column tiles


def test_multi_device_block_queue_of_the_view_classes(emu_lib):
    """MVDeconFFT.convolve1 in the reference's multi-device mode (MVDeconFFT.java:447-469): one host thread per deviceList
    entry pulling blocks from a shared counter.  Three threads on the emulator's single device exercise the queue and the
    native side's per-device serialisation; the result must equal the single-thread block loop and the oracle."""
    import __graft_entry__ as g
    from spim_registration_b200 import MVDeconFFT, native, synthetic
    from spim_registration_b200.deconvolution import _ViewFFT
    saved = _ViewFFT.cuda
    _ViewFFT.cuda = native.CUDAFourierConvolution(g.build_emulator())
    try:
        shape = (20, 22, 26)
        _, imgs, ws, psfs = synthetic.make_dataset(shape, 1, 5)
        psi = np.random.default_rng(0).random(shape, dtype=np.float32)
        one = MVDeconFFT(imgs[0], ws[0], psfs[0], None, [0], True, (12, 10, 8), False)
        many = MVDeconFFT(imgs[0], ws[0], psfs[0], None, [0, 0, 0], True, (12, 10, 8), False)
        assert len(many.blocks) > 8
        a, b = one.convolve1(psi), many.convolve1(psi)
        assert np.array_equal(a, b)
        r = O.convolve(psi, psfs[0], O.EXT_MIRROR_SINGLE, dtype=np.float64)
        assert np.abs(b - r).max() / np.abs(r).max() < P.TOL_CONV
    finally:
        _ViewFFT.cuda = saved


def test_column_staging_modes_are_bit_identical(emu_lib, monkeypatch):
    """SPIM_COLP selects how a column tile reaches shared memory (0 first stage from global memory, 2 one-shot staging,
    3 persistent pipeline) and SPIM_COL_NARROW its width: none of them may change a single bit of the result."""
    shape = (14, 18, 22)
    _, imgs, ws, psfs = __import__("spim_registration_b200").synthetic.make_dataset(shape, 3, 5, kind="beads")
    a, *_ = P.run_session(emu_lib, imgs, ws, psfs, O.EFFICIENT_BAYESIAN, 2, 3)
    for colp in ("0", "2", "3"):
        monkeypatch.setenv("SPIM_COLP", colp)
        b, *_ = P.run_session(emu_lib, imgs, ws, psfs, O.EFFICIENT_BAYESIAN, 2, 3)
        assert np.array_equal(a, b), colp
    monkeypatch.delenv("SPIM_COLP")
    P.conv_case(emu_lib, (9, 7, 11), (3, 5, 3), 2)
    P.golden_case(emu_lib, 1, 1)
    monkeypatch.setenv("SPIM_COL_NARROW", "1")
    c, *_ = P.run_session(emu_lib, imgs, ws, psfs, O.EFFICIENT_BAYESIAN, 2, 3)
    assert np.array_equal(a, c)


def test_lean_column_pass_for_small_radix_plans(emu_lib):
    """Small column tiles of plans without radices 9 / 10 run from the instantiation compiled for radices <= 8 (same
    butterflies, fewer registers); plans that do need them keep the general kernel."""
    for n in (16, 24, 28, 36, 48, 56, 64, 72, 84, 96, 112, 128, 144, 288,      # radices <= 8 only
              18, 20, 30, 50, 90, 100, 560):                                   # plans with 9 / 10: general kernel
        for shp in ((n, 4, 8), (4, n, 8)):
            P.legacy_case(emu_lib, shp, (3, 3, 3), seed=n)


def test_deduplicated_forward_sweeps_are_bit_identical(emu_lib, monkeypatch):
    """Mirror / periodic extension: the halo lines and planes are copies of image lines, so the forward x and y sweeps
    transform each only once and store it twice (SPIM_DEDUP=0: every padded line).  Same data through the same
    arithmetic: not a bit may differ -- single convolutions with every rule, gen-1 (mirror in both convolutions) and gen-2."""
    syn = __import__("spim_registration_b200").synthetic
    for shape, k in (((14, 18, 24), 5), ((33, 20, 40), 7), ((12, 40, 16), 9)):
        _, imgs, ws, psfs = syn.make_dataset(shape, 2, k, kind="beads")
        for gen in (1, 2):
            monkeypatch.setenv("SPIM_DEDUP", "1")
            c1 = emu_lib.mvd_debug_counter(1)
            a, *_ = P.run_session(emu_lib, imgs, ws, psfs, O.EFFICIENT_BAYESIAN, gen, 2)
            assert emu_lib.mvd_debug_counter(1) >= c1 + (8 if gen == 1 else 4)      # mirror: both convolutions (gen-1) / conv1 (gen-2)
            monkeypatch.setenv("SPIM_DEDUP", "0")
            b, *_ = P.run_session(emu_lib, imgs, ws, psfs, O.EFFICIENT_BAYESIAN, gen, 2)
            assert np.array_equal(a, b), (shape, gen)
    monkeypatch.setenv("SPIM_DEDUP", "1")
    for ext in range(5):
        P.conv_case(emu_lib, (10, 12, 16), (5, 5, 5), ext)
        P.conv_case(emu_lib, (6, 9, 8), (5, 7, 3), ext)         # halo wider than half the image: several rows per source row -> falls back


def test_constant_extension_by_shift(emu_lib, monkeypatch):
    """gen-2 conv2 extends the quotient by the constant 1.  conv(ext_1(r), K) = conv(ext_0(r - 1), K) + sum(K): the session
    stores r - 1, convolves zero-extended and adds sum(K2) in the update (SPIM_CONST_SHIFT=0: the literal extension).
    Identical in exact arithmetic; in fp32 the two agree to round-off and both meet the parity bar."""
    syn = __import__("spim_registration_b200").synthetic
    shape = (14, 18, 24)
    _, imgs, ws, psfs = syn.make_dataset(shape, 3, 5, kind="beads")
    for typ in (O.EFFICIENT_BAYESIAN, O.INDEPENDENT, O.OPTIMIZATION_I, O.OPTIMIZATION_II):
        monkeypatch.setenv("SPIM_CONST_SHIFT", "1")
        c2 = emu_lib.mvd_debug_counter(2)
        a, *_ = P.run_session(emu_lib, imgs, ws, psfs, typ, 2, 3)
        assert emu_lib.mvd_debug_counter(2) >= c2 + 9
        monkeypatch.setenv("SPIM_CONST_SHIFT", "0")
        b, *_ = P.run_session(emu_lib, imgs, ws, psfs, typ, 2, 3)
        per, l2 = O.parity_errors(a, b)
        assert per <= 2e-5 and l2 <= 2e-6, (typ, per, l2)
    monkeypatch.setenv("SPIM_CONST_SHIFT", "1")
    P.decon_case(emu_lib, shape, 3, 5, O.EFFICIENT_BAYESIAN, 2, 3)


def test_single_brick_with_rule_filled_halos_equals_plain_session(emu_lib, monkeypatch):
    """Brick mode on one rank with EVERY halo declared neighbour data and filled by the out-of-bounds rule
    (mvd_fill_halo) must reproduce the plain session -- also with conv2's constant realised by shift, where the ratio
    buffer holds r - 1 and its rule constant therefore is 0."""
    from spim_registration_b200 import synthetic
    from spim_registration_b200.deconvolution import Session
    shape, V = (12, 16, 20), 2
    _, imgs, ws, psfs = synthetic.make_dataset(shape, V, 5)
    for shift in ("1", "0"):
        monkeypatch.setenv("SPIM_CONST_SHIFT", shift)
        with Session(shape, V, 2, generation=2, haloed=True, lib=emu_lib) as s:
            for v in range(V):
                s.set_view(v, imgs[v], ws[v], psfs[v])
            s.init()
            part = s.init_partials()
            s.set_avg(part[0] / part[1], 1.0)
            s.set_halo_mask(7, 7)
            for v in range(V):
                s.fill_halo(0, 7, 7)
                s.view_phase(v, 0)
                s.fill_halo(1, 7, 7)
                s.view_phase(v, 1)
            s.finish()
            brick = s.get_psi()
        plain, *_ = P.run_session(emu_lib, imgs, ws, psfs, 2, 2, 1)
        assert np.abs(brick - plain).max() <= 1e-5 * np.abs(plain).max(), shift


def test_eight_line_x_forward_tiles(emu_lib, monkeypatch):
    """Lines of 1080 voxels and more run 8-line x-forward tiles (three blocks per SM instead of one); forced here on small
    volumes: same arithmetic per line pair, so not a bit may differ from the 16-line tiles -- every extension rule, odd line
    counts (partial tiles), de-duplicated and full sweeps, brick-style haloed sources."""
    syn = __import__("spim_registration_b200").synthetic
    for shape, k in (((14, 18, 24), 5), ((7, 13, 40), 7), ((12, 40, 16), 9)):
        _, imgs, ws, psfs = syn.make_dataset(shape, 2, k, kind="beads")
        for gen in (1, 2):
            monkeypatch.setenv("SPIM_XFWD_LINES", "16")
            a, *_ = P.run_session(emu_lib, imgs, ws, psfs, O.EFFICIENT_BAYESIAN, gen, 2)
            monkeypatch.setenv("SPIM_XFWD_LINES", "8")
            b, *_ = P.run_session(emu_lib, imgs, ws, psfs, O.EFFICIENT_BAYESIAN, gen, 2)
            assert np.array_equal(a, b), (shape, gen)
    monkeypatch.setenv("SPIM_XFWD_LINES", "8")
    for ext in range(5):
        P.conv_case(emu_lib, (10, 12, 16), (5, 5, 5), ext)
    for n in (20, 30, 42, 56, 60, 70, 84, 100, 120, 140):
        P.legacy_case(emu_lib, (4, 4, n), (3, 3, 3), seed=n)
    P.decon_case(emu_lib, (14, 18, 24), 3, 5, O.EFFICIENT_BAYESIAN, 2, 3)


def test_randomized_work_saving_switches(emu_lib, monkeypatch):
    """Random shapes (rows of 4 k voxels, so the TMA-fed x-forward kernel runs), kernels up to and beyond the image size, every
    extension rule: the de-duplicated sweeps and the 8-line tiles must reproduce the full 16-line sweeps bit for bit, and all
    must agree with the oracle."""
    from spim_registration_b200 import native
    rng = np.random.default_rng(20241018)
    for _ in range(60):
        shape = (int(rng.integers(1, 20)), int(rng.integers(1, 20)), 4 * int(rng.integers(1, 7)))
        ks = tuple(int(rng.integers(1, 12)) for _ in range(3))
        ext = int(rng.integers(0, 5))
        img = rng.random(shape, dtype=np.float32)
        k = rng.random(ks, dtype=np.float32)
        outs = []
        for dedup, lines in (("1", "16"), ("0", "16"), ("1", "8"), ("0", "8")):
            monkeypatch.setenv("SPIM_DEDUP", dedup)
            monkeypatch.setenv("SPIM_XFWD_LINES", lines)
            outs.append(native.convolve(img, k, ext, 1.0, lib=emu_lib))
        for o in outs[1:]:
            assert np.array_equal(outs[0], o), (shape, ks, ext)
        ref = O.convolve(img, k, ext, value=1.0, dtype=np.float64)
        assert np.abs(outs[0] - ref).max() <= P.TOL_CONV * np.abs(ref).max(), (shape, ks, ext)
