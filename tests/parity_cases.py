"""Parity checks shared by the emulator tests (CPU, tiny sizes: index math of the kernel bodies) and
the GPU tests (the real CUDA library through the C-ABI).  Every check compares the library against
the oracle on the same seeded inputs."""
import os

import numpy as np
import pytest

from oracle import mvdecon_oracle as O
from spim_registration_b200 import native, synthetic
from spim_registration_b200.deconvolution import Session

# BASELINE.md section 5 / north_star: per-voxel relative error <= 1e-3 (denominator floored at
# minValue = 1e-4) and relative L2 <= 1e-4 against the fp32-mode oracle
TOL_PER_VOXEL = 1e-3
TOL_L2 = 1e-4
# a single FFT convolution in fp32 (round-off of two FFT implementations with different radices)
TOL_CONV = 2e-5

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def conv_case(lib, shape, kshape, ext, seed=0, value=1.0, tol=TOL_CONV):
    rng = np.random.default_rng(seed)
    img = rng.random(shape, dtype=np.float32)
    k = rng.random(kshape, dtype=np.float32)
    ref = O.convolve(img, k, ext, value=value, dtype=np.float64)
    out = native.convolve(img, k, ext, value, lib=lib)
    err = float(np.abs(out - ref).max() / np.abs(ref).max())
    assert err <= tol, f"conv {shape} {kshape} ext {ext}: rel err {err:.3e}"
    return err


def legacy_case(lib, shape, kshape, seed=0, tol=TOL_CONV):
    """convolution3DfftCUDAInPlace == circular convolution on exactly imDim."""
    rng = np.random.default_rng(seed)
    im = rng.random(shape, dtype=np.float32)
    k = rng.random(kshape, dtype=np.float32)
    ref = O.circular_convolve(im, k, dtype=np.float64)
    got = im.copy()
    kk = k.copy()
    lib.convolution3DfftCUDAInPlace(got.ctypes.data_as(native.c_float_p), native.int3(shape),
                                    kk.ctypes.data_as(native.c_float_p), native.int3(kshape), 0)
    assert np.array_equal(kk, k), "kernel must not be modified"
    err = float(np.abs(got - ref).max() / np.abs(ref).max())
    assert err <= tol, f"legacy {shape} {kshape}: rel err {err:.3e} ({lib.spim_fftconv_last_error()})"
    return err


def run_session(lib, imgs, ws, psfs, typ, gen, iters, lam=0.006, osem=1.0, osem_index=0, psi0=None, **kw):
    shape = imgs[0].shape
    V = len(imgs)
    with Session(shape, V, typ, generation=gen, lam=lam, osem_speedup=osem, osem_index=osem_index, lib=lib, **kw) as s:
        for v in range(V):
            s.set_view(v, imgs[v], None if ws is None else ws[v], psfs[v])
        s.init()
        if psi0 is not None:
            s.set_psi(psi0)
        k1 = [s.get_kernel(v, 1) for v in range(V)]
        k2 = [s.get_kernel(v, 2) for v in range(V)]
        st = s.run(iters)
        s.finish()
        psi = s.get_psi()
        info = s.info()
        return psi, k1, k2, st, (info.avg, info.osem, info.min_overlap, info.avg_overlap)


def decon_case(lib, shape, V, ks, typ, gen, iters, kind="beads", lam=0.006, weight_mode="normalized",
               osem=1.0, osem_index=0, use_weights=True, seed=7):
    _, imgs, ws, psfs = synthetic.make_dataset(shape, V, ks, kind=kind, seed=seed, weight_mode=weight_mode)
    if not use_weights:
        ws_o = [np.ones(shape, np.float32) for _ in range(V)]
        ws_l = None
    else:
        ws_o = ws_l = ws
    p = O.DeconParams(iteration_type=typ, num_iterations=iters, lam=lam, gen=gen, osem_speedup=osem, osem_index=osem_index)
    ref = O.deconvolve(imgs, ws_o, psfs, p)
    psi, k1, k2, st, (avg, osem_used, mn, av) = run_session(lib, imgs, ws_l, psfs, typ, gen, iters, lam, osem, osem_index)
    for v in range(V):
        assert np.abs(k1[v] - ref.kernel1[v]).max() <= 1e-6 * ref.kernel1[v].max()
        assert np.abs(k2[v] - ref.kernel2[v]).max() <= 5e-5 * ref.kernel2[v].max(), f"kernel2 view {v}"
    assert np.isclose(avg, ref.avg, rtol=1e-6), (avg, ref.avg)
    assert np.isclose(osem_used, ref.osem, rtol=1e-6), (osem_used, ref.osem)
    per, l2 = O.parity_errors(psi, ref.psi)
    assert per <= TOL_PER_VOXEL and l2 <= TOL_L2, f"psi parity: per-voxel {per:.3e}, L2 {l2:.3e}"
    # the per view-step statistics the reference logs (MVDeconvolution.java:441-457)
    s, m = st
    for (it, v, rs, rm) in ref.stats:
        assert np.isclose(s[it, v], rs, rtol=2e-3, atol=1e-6), (it, v, s[it, v], rs)
        assert np.isclose(m[it, v], rm, rtol=5e-3, atol=1e-6), (it, v, m[it, v], rm)
    return per, l2


def exact_tikhonov_case(lib, shape=(14, 16, 18)):
    """exact_tikhonov=1 evaluates the reference's fp64 expression; the default fp32 form must agree with
    it to a few ulp and both must meet the parity bar."""
    _, imgs, ws, psfs = synthetic.make_dataset(shape, 2, 5, kind="beads")
    ref = O.deconvolve(imgs, ws, psfs, O.DeconParams(iteration_type=3, num_iterations=3, lam=0.006, gen=2))
    a, *_ = run_session(lib, imgs, ws, psfs, 3, 2, 3, exact_tikhonov=True)
    b, *_ = run_session(lib, imgs, ws, psfs, 3, 2, 3, exact_tikhonov=False)
    for psi in (a, b):
        per, l2 = O.parity_errors(psi, ref.psi)
        assert per <= TOL_PER_VOXEL and l2 <= TOL_L2
    per, l2 = O.parity_errors(a, b)
    assert per <= 2e-5 and l2 <= 2e-6, (per, l2)


def golden_case(lib, gen, typ):
    d = np.load(os.path.join(G, "decon_small.npz"))
    V = int(d["num_views"])
    imgs = [d[f"img{v}"] for v in range(V)]
    ws = [d[f"w{v}"] for v in range(V)]
    psfs = [d[f"psf{v}"] for v in range(V)]
    psi, _, k2, _, (avg, *_rest) = run_session(lib, imgs, ws, psfs, typ, gen, 2)
    per, l2 = O.parity_errors(psi, d[f"psi_g{gen}_t{typ}"])
    assert per <= TOL_PER_VOXEL and l2 <= TOL_L2, f"golden psi g{gen} t{typ}: {per:.3e} {l2:.3e}"
    assert np.isclose(avg, float(d[f"avg_g{gen}_t{typ}"]), rtol=1e-6)
    if gen == 2:
        for v in range(V):
            g = d[f"k2_t{typ}_v{v}"]
            assert np.abs(k2[v] - g).max() <= 5e-5 * g.max()


def golden_conv_case(lib):
    d = np.load(os.path.join(G, "conv_small.npz"))
    for ext in range(5):
        out = native.convolve(d["img"], d["kernel"], ext, 1.0, lib=lib)
        assert np.abs(out - d[f"ext{ext}"]).max() / np.abs(d[f"ext{ext}"]).max() < TOL_CONV
    got = d["img"].copy()
    k = d["kernel"].copy()
    lib.convolution3DfftCUDAInPlace(got.ctypes.data_as(native.c_float_p), native.int3(got.shape),
                                    k.ctypes.data_as(native.c_float_p), native.int3(k.shape), 0)
    assert np.abs(got - d["circular"]).max() / np.abs(d["circular"]).max() < TOL_CONV


def cells_case(lib, shape=(10, 12, 14), cell=(4, 5, 6), V=2, ks=3):
    """Views handed over cell by cell (mvd_upload_region, the CellImg / > 2^31-element path) give bit-identical
    device buffers -- and therefore a bit-identical deconvolution -- to whole-array uploads."""
    from spim_registration_b200 import fusion
    _, imgs, ws, psfs = synthetic.make_dataset(shape, V, ks, kind="beads")
    whole, *_ = run_session(lib, imgs, ws, psfs, O.EFFICIENT_BAYESIAN, 2, 2)
    with Session(shape, V, O.EFFICIENT_BAYESIAN, generation=2, lib=lib) as s:
        for v in range(V):
            for z in range(0, shape[0], cell[0]):
                for y in range(0, shape[1], cell[1]):
                    for x in range(0, shape[2], cell[2]):
                        sl = (slice(z, z + cell[0]), slice(y, y + cell[1]), slice(x, x + cell[2]))
                        s.upload_region(v, 0, imgs[v][sl], (z, y, x))
                        s.upload_region(v, 1, ws[v][sl], (z, y, x))
            fusion.set_psf(s, v, psfs[v])
            assert np.array_equal(fusion.get_view(s, v, 0), imgs[v]) and np.array_equal(fusion.get_view(s, v, 1), ws[v])
        with pytest.raises(native.NativeError, match="out of range"):
            s.upload_region(0, 0, np.zeros((2, 2, 2), np.float32), (shape[0] - 1, 0, 0))
        s.init()
        s.run(2)
        s.finish()
        assert np.array_equal(s.get_psi(), whole)


def fast_epilogue_case(lib, shape=(14, 16, 18), bit_identical=False):
    """fast_epilogue = 1 (branch-free division / square root, csrc/fast_math.h): for operands in the normal range the values
    are the correctly rounded ones, so the deconvolution must agree with the IEEE epilogue to rounding noise and meet the
    same parity bar; under the emulator (correctly rounded seeds) it is bit-identical."""
    for gen, typ, lam in ((2, O.EFFICIENT_BAYESIAN, 0.006), (1, O.OPTIMIZATION_I, 0.06), (2, O.INDEPENDENT, 0.0)):
        _, imgs, ws, psfs = synthetic.make_dataset(shape, 3, 5, kind="beads")
        ref = O.deconvolve(imgs, ws, psfs, O.DeconParams(iteration_type=typ, num_iterations=3, lam=lam, gen=gen))
        a, *_ = run_session(lib, imgs, ws, psfs, typ, gen, 3, lam=lam, fast_epilogue=True)
        b, *_ = run_session(lib, imgs, ws, psfs, typ, gen, 3, lam=lam, fast_epilogue=False)
        per, l2 = O.parity_errors(a, ref.psi)
        assert per <= TOL_PER_VOXEL and l2 <= TOL_L2, (gen, typ, per, l2)
        if bit_identical:
            assert np.array_equal(a, b)
        else:
            per, l2 = O.parity_errors(a, b)
            assert per <= 2e-5 and l2 <= 2e-6, (gen, typ, per, l2)


def reinit_case(lib, shape=(16, 16, 16)):
    """mvd_init a second time after mvd_set_view replaced the PSFs by larger ones (3^3 -> 15^3): the session re-plans
    (larger padded FFT size) and must re-create every buffer whose size changed; the second run equals a fresh session."""
    _, imgs, ws, small = synthetic.make_dataset(shape, 2, 3, kind="beads")
    _, _, _, big = synthetic.make_dataset(shape, 2, 15, kind="beads")
    fresh, *_ = run_session(lib, imgs, ws, big, O.EFFICIENT_BAYESIAN, 2, 2)
    for haloed in (False, True):
        with Session(shape, 2, O.EFFICIENT_BAYESIAN, generation=2, lib=lib, haloed=haloed) as s:
            for v in range(2):
                s.set_view(v, imgs[v], ws[v], small[v])
            s.init()
            b0 = s.info().device_bytes
            if haloed:
                s.set_avg(0.5, 1.0)
                s.view_phase(0, 0); s.view_phase(0, 1)
            else:
                s.run(1)
            for v in range(2):
                s.set_view(v, imgs[v], ws[v], big[v])
            s.init()
            i = s.info()
            assert tuple(i.fft_dims) == tuple(lib.mvd_fft_size(n + 14, 1 if d == 2 else 0) for d, n in enumerate(shape))
            assert i.device_bytes > b0
            if haloed:
                _, dims, origin = s.device_buffer(0)
                assert tuple(dims)[:2] == tuple(n + 14 for n in shape[:2]) and tuple(origin)[:2] == (7, 7)
                continue
            s.run(2)
            s.finish()
            assert np.array_equal(s.get_psi(), fresh)
            # and back to the small PSFs: the accounting shrinks again
            for v in range(2):
                s.set_view(v, imgs[v], ws[v], small[v])
            s.init()
            assert s.info().device_bytes == b0
