"""The C-ABI library loads on a CPU-only box, exports every symbol the headers declare, and fails
loudly (no CPU fallback) when asked to compute without a GPU.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

from spim_registration_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src)
    return sorted(set(names))


def test_headers_and_binding_agree():
    legacy = declared_symbols("spim_fftconv.h")
    session = declared_symbols("spim_mvdecon.h")
    assert sorted(native.LEGACY_SYMBOLS) == legacy
    assert sorted(native.SESSION_SYMBOLS) == session
    assert sorted(native.FUSION_SYMBOLS) == declared_symbols("spim_fusion.h")


def test_library_exports_every_declared_symbol(cuda_lib):
    for name in declared_symbols("spim_fftconv.h") + declared_symbols("spim_mvdecon.h") + declared_symbols("spim_fusion.h"):
        assert hasattr(cuda_lib, name), name


def test_reference_library_names_are_shipped(cuda_lib):
    d = os.path.join(ROOT, "spim_registration_b200")
    for n in native.LIB_NAMES:
        assert os.path.exists(os.path.join(d, n)), n


def test_jna_interface_signatures():
    # the eight symbols of CUDAStandardFunctions.java:36-44 / CUDAFourierConvolution.java:30-31
    c = native.CUDAFourierConvolution()
    for m in ("getCUDAcomputeCapabilityMinorVersion", "getCUDAcomputeCapabilityMajorVersion", "getNumDevicesCUDA",
              "getNameDeviceCUDA", "getMemDeviceCUDA", "getFreeMemDeviceCUDA", "convolution3DfftCUDA",
              "convolution3DfftCUDAInPlace"):
        assert callable(getattr(c, m))
    assert c.lib.getMemDeviceCUDA.restype is ctypes.c_longlong     # Java long


def test_fft_size_helper(cuda_lib):
    assert cuda_lib.mvd_fft_size(542, 1) == 560
    assert cuda_lib.mvd_fft_size(286, 0) == 288
    assert cuda_lib.mvd_fft_size(1054, 1) in (1056, 1080)
    assert cuda_lib.mvd_fft_size(142, 1) == 144
    for n in (2, 3, 17, 100, 257, 1000):
        m = cuda_lib.mvd_fft_size(n, 0)
        assert m >= n
        r = m
        for p in (2, 3, 5, 7, 11, 13):
            while r % p == 0:
                r //= p
        assert r == 1


def test_no_gpu_means_loud_failure_not_cpu_fallback(cuda_lib):
    n = cuda_lib.getNumDevicesCUDA()
    if n > 0:
        pytest.skip("a GPU is present")
    assert n in (0, -1)
    p = native.MvdParams()
    cuda_lib.mvd_params_default(ctypes.byref(p))
    p.dims[:] = [4, 4, 4]
    h = ctypes.c_void_p()
    rc = cuda_lib.mvd_session_create(ctypes.byref(p), ctypes.byref(h))
    assert rc != 0 and b"no CUDA device" in cuda_lib.mvd_last_error()
    c = native.CUDAFourierConvolution()
    im = np.ones((4, 4, 4), np.float32)
    before = im.copy()
    c.convolution3DfftCUDAInPlace(im, im.shape, np.ones((3, 3, 3), np.float32), (3, 3, 3), 0)
    assert np.array_equal(im, before)                    # buffer left untouched on failure
    assert "no CUDA device" in c.last_error()
    assert c.getCUDAcomputeCapabilityMajorVersion(0) == -1


def test_missing_library_raises():
    with pytest.raises(OSError):
        native.load_library("/nonexistent/libConvolution3D_fftCUDAlib.so")


def test_params_default_values(cuda_lib):
    p = native.MvdParams()
    cuda_lib.mvd_params_default(ctypes.byref(p))
    assert p.struct_size == ctypes.sizeof(native.MvdParams)
    assert p.generation == 2 and p.iteration_type == 2 and abs(p.lambda_ - 0.006) < 1e-12
    assert abs(p.min_value - 1e-4) < 1e-10 and p.conv1_ext == -1 and p.conv2_ext == -1


def test_host_only_fusion_helpers_work_without_gpu(cuda_lib):
    # geometry of ExtractPSF.transformPSF and the blending table are host logic: no device needed
    lut = (ctypes.c_double * 1001)()
    assert cuda_lib.mvd_blending_lookup(lut) == 0 and lut[0] == 0.0 and abs(lut[1000] - 1.0) < 1e-12
    od, off = (ctypes.c_int * 3)(), (ctypes.c_double * 3)()
    scal = (ctypes.c_double * 12)(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 2.5, 0)
    assert cuda_lib.mvd_transform_psf_size(native.int3((5, 7, 9)), scal, od, off) == 0
    assert tuple(od) == (11, 7, 9) and tuple(off) == (0.0, 0.0, 0.0)


def test_fusion_compute_fails_loudly_without_gpu(cuda_lib):
    if cuda_lib.getNumDevicesCUDA() > 0:
        pytest.skip("a GPU is present")
    psf = np.ones((3, 3, 3), np.float32)
    ident = (ctypes.c_double * 12)(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0)
    out = np.zeros((3, 3, 3), np.float32)
    rc = cuda_lib.mvd_transform_psf(psf.ctypes.data, native.int3((3, 3, 3)), ident, ident, out.ctypes.data, native.int3((3, 3, 3)), 0)
    assert rc != 0 and b"no CUDA device" in cuda_lib.mvd_last_error()
