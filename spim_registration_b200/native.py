"""ctypes binding of the C-ABI shared library -- the stand-in for the reference's JNA proxies.

``CUDAFourierConvolution`` mirrors the JNA interface of the same name
(/root/reference/src/main/java/spim/process/cuda/CUDAFourierConvolution.java:25-32 and
CUDAStandardFunctions.java:26-45) method for method, using JNA's marshalling conventions
(primitive arrays in/out, ``long`` = 64 bit, ``byte[256]`` name buffer).

There is no CPU fallback: if the CUDA library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
#: names the reference looks for (fiji/plugin/Multi_View_Deconvolution.java:858,
#: spim/process/fusion/deconvolution/EfficientBayesianBased.java:1127-1131)
LIB_NAMES = ("libConvolution3D_fftCUDAlib.so", "libFourierConvolutionCUDALib.so")

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
c_double_p = C.POINTER(C.c_double)


class MvdParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int),
        ("dims", C.c_int * 3),
        ("num_views", C.c_int),
        ("iteration_type", C.c_int),
        ("generation", C.c_int),
        ("lambda_", C.c_double),
        ("min_value", C.c_float),
        ("osem_speedup", C.c_double),
        ("osem_index", C.c_int),
        ("conv1_ext", C.c_int),
        ("conv2_ext", C.c_int),
        ("device", C.c_int),
        ("haloed", C.c_int),
        ("exact_tikhonov", C.c_int),
        ("fast_epilogue", C.c_int),
        ("reserved", C.c_int * 6),
    ]


class MvdInfo(C.Structure):
    _fields_ = [
        ("avg", C.c_double),
        ("osem", C.c_double),
        ("min_overlap", C.c_int),
        ("avg_overlap", C.c_double),
        ("fft_dims", C.c_int * 3),
        ("pitch", C.c_int),
        ("n_voxels", C.c_longlong),
        ("np_voxels", C.c_longlong),
        ("device_bytes", C.c_longlong),
        ("halo_lo", C.c_int * 3),
        ("halo_hi", C.c_int * 3),
    ]


#: every symbol declared in include/spim_fftconv.h and include/spim_mvdecon.h
LEGACY_SYMBOLS = (
    "getCUDAcomputeCapabilityMinorVersion", "getCUDAcomputeCapabilityMajorVersion", "getNumDevicesCUDA",
    "getNameDeviceCUDA", "getMemDeviceCUDA", "getFreeMemDeviceCUDA", "convolution3DfftCUDAInPlace",
    "convolution3DfftCUDA", "spim_fftconv_last_error",
)
SESSION_SYMBOLS = (
    "mvd_params_default", "mvd_session_create", "mvd_session_destroy", "mvd_set_view", "mvd_upload_region", "mvd_init", "mvd_run",
    "mvd_finish", "mvd_get_psi", "mvd_set_psi", "mvd_get_kernel", "mvd_get_info", "mvd_sync", "mvd_get_stream", "mvd_set_timing",
    "mvd_get_timing", "mvd_get_device_buffer", "mvd_set_halo_mask", "mvd_halo_pack", "mvd_halo_unpack", "mvd_fill_halo", "mvd_view_phase", "mvd_init_partials",
    "mvd_set_avg", "mvd_p2p_export", "mvd_p2p_connect", "mvd_p2p_push", "mvd_p2p_wait", "mvd_p2p_status", "mvd_p2p_disconnect", "mvd_convolve", "mvd_fft_size", "mvd_debug_counter", "mvd_last_error", "mvd_version",
)

#: include/spim_fusion.h
FUSION_SYMBOLS = (
    "mvd_load_stack", "mvd_transform_view", "mvd_set_psf", "mvd_normalize_weights", "mvd_get_view", "mvd_extract_psf",
    "mvd_transform_psf_size", "mvd_transform_psf", "mvd_blending_lookup",
)


class MvdTransform(C.Structure):
    """``mvd_transform`` of include/spim_fusion.h."""
    _fields_ = [
        ("struct_size", C.c_int),
        ("inverse", C.c_double * 12),
        ("offset", C.c_longlong * 3),
        ("want_image", C.c_int),
        ("want_weight", C.c_int),
        ("border", C.c_float * 3),
        ("range", C.c_float * 3),
        ("reserved", C.c_int * 8),
    ]


def default_library_path() -> str:
    env = os.environ.get("SPIM_B200_LIBRARY")
    if env:
        return env
    return os.path.join(_HERE, LIB_NAMES[0])


def _declare(lib: C.CDLL) -> None:
    lib.getCUDAcomputeCapabilityMinorVersion.argtypes = [C.c_int]
    lib.getCUDAcomputeCapabilityMinorVersion.restype = C.c_int
    lib.getCUDAcomputeCapabilityMajorVersion.argtypes = [C.c_int]
    lib.getCUDAcomputeCapabilityMajorVersion.restype = C.c_int
    lib.getNumDevicesCUDA.argtypes = []
    lib.getNumDevicesCUDA.restype = C.c_int
    lib.getNameDeviceCUDA.argtypes = [C.c_int, C.c_char_p]
    lib.getNameDeviceCUDA.restype = None
    lib.getMemDeviceCUDA.argtypes = [C.c_int]
    lib.getMemDeviceCUDA.restype = C.c_longlong
    lib.getFreeMemDeviceCUDA.argtypes = [C.c_int]
    lib.getFreeMemDeviceCUDA.restype = C.c_longlong
    lib.convolution3DfftCUDAInPlace.argtypes = [c_float_p, c_int_p, c_float_p, c_int_p, C.c_int]
    lib.convolution3DfftCUDAInPlace.restype = None
    lib.convolution3DfftCUDA.argtypes = [c_float_p, c_int_p, c_float_p, c_int_p, C.c_int]
    lib.convolution3DfftCUDA.restype = C.c_void_p
    lib.spim_fftconv_last_error.argtypes = []
    lib.spim_fftconv_last_error.restype = C.c_char_p

    P = C.POINTER(MvdParams)
    S = C.c_void_p
    lib.mvd_params_default.argtypes = [P]
    lib.mvd_params_default.restype = None
    lib.mvd_session_create.argtypes = [P, C.POINTER(S)]
    lib.mvd_session_create.restype = C.c_int
    lib.mvd_session_destroy.argtypes = [S]
    lib.mvd_session_destroy.restype = None
    lib.mvd_set_view.argtypes = [S, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, c_int_p]
    lib.mvd_set_view.restype = C.c_int
    lib.mvd_upload_region.argtypes = [S, C.c_int, C.c_int, C.c_void_p, c_int_p, c_int_p]
    lib.mvd_upload_region.restype = C.c_int
    lib.mvd_init.argtypes = [S]
    lib.mvd_init.restype = C.c_int
    lib.mvd_run.argtypes = [S, C.c_int, c_double_p, c_double_p]
    lib.mvd_run.restype = C.c_int
    lib.mvd_finish.argtypes = [S]
    lib.mvd_finish.restype = C.c_int
    lib.mvd_get_psi.argtypes = [S, C.c_void_p]
    lib.mvd_get_psi.restype = C.c_int
    lib.mvd_set_psi.argtypes = [S, C.c_void_p]
    lib.mvd_set_psi.restype = C.c_int
    lib.mvd_get_kernel.argtypes = [S, C.c_int, C.c_int, C.c_void_p]
    lib.mvd_get_kernel.restype = C.c_int
    lib.mvd_get_info.argtypes = [S, C.POINTER(MvdInfo)]
    lib.mvd_get_info.restype = C.c_int
    lib.mvd_sync.argtypes = [S]
    lib.mvd_sync.restype = C.c_int
    lib.mvd_get_stream.argtypes = [S, C.POINTER(C.c_void_p)]
    lib.mvd_get_stream.restype = C.c_int
    lib.mvd_set_timing.argtypes = [S, C.c_int]
    lib.mvd_set_timing.restype = C.c_int
    lib.mvd_get_timing.argtypes = [S, c_double_p, C.POINTER(C.c_longlong)]
    lib.mvd_get_timing.restype = C.c_int
    lib.mvd_get_device_buffer.argtypes = [S, C.c_int, C.POINTER(C.c_void_p), c_int_p, c_int_p]
    lib.mvd_get_device_buffer.restype = C.c_int
    lib.mvd_set_halo_mask.argtypes = [S, C.c_int, C.c_int]
    lib.mvd_set_halo_mask.restype = C.c_int
    lib.mvd_halo_pack.argtypes = [S, C.c_int, C.c_int, c_int_p, C.c_void_p]
    lib.mvd_halo_pack.restype = C.c_int
    lib.mvd_halo_unpack.argtypes = [S, C.c_int, C.c_int, c_int_p, C.c_void_p]
    lib.mvd_halo_unpack.restype = C.c_int
    lib.mvd_fill_halo.argtypes = [S, C.c_int, C.c_int, C.c_int]
    lib.mvd_fill_halo.restype = C.c_int
    lib.mvd_p2p_export.argtypes = [S, C.c_void_p]
    lib.mvd_p2p_export.restype = C.c_int
    lib.mvd_p2p_connect.argtypes = [S, C.c_int, C.c_void_p, c_int_p, c_int_p]
    lib.mvd_p2p_connect.restype = C.c_int
    lib.mvd_p2p_push.argtypes = [S, C.c_int]
    lib.mvd_p2p_push.restype = C.c_int
    lib.mvd_p2p_wait.argtypes = [S, C.c_int]
    lib.mvd_p2p_wait.restype = C.c_int
    lib.mvd_p2p_status.argtypes = [S, c_int_p]
    lib.mvd_p2p_status.restype = C.c_int
    lib.mvd_p2p_disconnect.argtypes = [S]
    lib.mvd_p2p_disconnect.restype = C.c_int
    lib.mvd_view_phase.argtypes = [S, C.c_int, C.c_int, c_double_p]
    lib.mvd_view_phase.restype = C.c_int
    lib.mvd_init_partials.argtypes = [S, c_double_p]
    lib.mvd_init_partials.restype = C.c_int
    lib.mvd_set_avg.argtypes = [S, C.c_double, C.c_double]
    lib.mvd_set_avg.restype = C.c_int
    lib.mvd_convolve.argtypes = [C.c_void_p, c_int_p, C.c_void_p, c_int_p, C.c_int, C.c_float, C.c_void_p, C.c_int]
    lib.mvd_convolve.restype = C.c_int
    lib.mvd_debug_counter.argtypes = [C.c_int]
    lib.mvd_debug_counter.restype = C.c_longlong
    lib.mvd_fft_size.argtypes = [C.c_int, C.c_int]
    lib.mvd_fft_size.restype = C.c_int
    lib.mvd_last_error.argtypes = []
    lib.mvd_last_error.restype = C.c_char_p
    lib.mvd_version.argtypes = []
    lib.mvd_version.restype = C.c_char_p

    T = C.POINTER(MvdTransform)
    lib.mvd_load_stack.argtypes = [S, C.c_void_p, c_int_p, C.c_int]
    lib.mvd_load_stack.restype = C.c_int
    lib.mvd_transform_view.argtypes = [S, C.c_int, T]
    lib.mvd_transform_view.restype = C.c_int
    lib.mvd_set_psf.argtypes = [S, C.c_int, C.c_void_p, c_int_p]
    lib.mvd_set_psf.restype = C.c_int
    lib.mvd_normalize_weights.argtypes = [S, C.c_int, C.c_int, c_int_p, c_double_p]
    lib.mvd_normalize_weights.restype = C.c_int
    lib.mvd_get_view.argtypes = [S, C.c_int, C.c_int, C.c_void_p]
    lib.mvd_get_view.restype = C.c_int
    lib.mvd_extract_psf.argtypes = [S, C.c_int, c_double_p, c_int_p, C.c_int, C.c_void_p]
    lib.mvd_extract_psf.restype = C.c_int
    lib.mvd_transform_psf_size.argtypes = [c_int_p, c_double_p, c_int_p, c_double_p]
    lib.mvd_transform_psf_size.restype = C.c_int
    lib.mvd_transform_psf.argtypes = [C.c_void_p, c_int_p, c_double_p, c_double_p, C.c_void_p, c_int_p, C.c_int]
    lib.mvd_transform_psf.restype = C.c_int
    lib.mvd_blending_lookup.argtypes = [c_double_p]
    lib.mvd_blending_lookup.restype = C.c_int


_LIB_CACHE = {}


def load_library(path: Optional[str] = None) -> C.CDLL:
    """``Native.load(...)``: load the shared library and declare every entry point.
    Raises ``OSError`` (the analogue of ``UnsatisfiedLinkError``) when the library is missing --
    there is deliberately no fallback implementation."""
    path = path or default_library_path()
    if path in _LIB_CACHE:
        return _LIB_CACHE[path]
    if not os.path.exists(path):
        raise OSError(
            f"CUDA library not found: {path}. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback).")
    lib = C.CDLL(path)
    _declare(lib)
    _LIB_CACHE[path] = lib
    return lib


class NativeError(RuntimeError):
    pass


def check(lib: C.CDLL, rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.mvd_last_error()
        raise NativeError(f"{what}: {msg.decode() if msg else 'error'}")


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def int3(v: Sequence[int]):
    return (C.c_int * 3)(int(v[0]), int(v[1]), int(v[2]))


class CUDAFourierConvolution:
    """Python twin of the JNA interface ``spim.process.cuda.CUDAFourierConvolution``."""

    def __init__(self, path: Optional[str] = None):
        self.lib = load_library(path)

    # -- CUDAStandardFunctions ------------------------------------------------------------------
    def getCUDAcomputeCapabilityMinorVersion(self, devCUDA: int) -> int:
        return self.lib.getCUDAcomputeCapabilityMinorVersion(devCUDA)

    def getCUDAcomputeCapabilityMajorVersion(self, devCUDA: int) -> int:
        return self.lib.getCUDAcomputeCapabilityMajorVersion(devCUDA)

    def getNumDevicesCUDA(self) -> int:
        return self.lib.getNumDevicesCUDA()

    def getNameDeviceCUDA(self, devCUDA: int, name: bytearray) -> None:
        buf = C.create_string_buffer(256)
        self.lib.getNameDeviceCUDA(devCUDA, buf)
        name[:256] = buf.raw[:len(name)]

    def getMemDeviceCUDA(self, devCUDA: int) -> int:
        return self.lib.getMemDeviceCUDA(devCUDA)

    def getFreeMemDeviceCUDA(self, devCUDA: int) -> int:
        return self.lib.getFreeMemDeviceCUDA(devCUDA)

    # -- CUDAFourierConvolution ---------------------------------------------------------------------
    def convolution3DfftCUDAInPlace(self, im: np.ndarray, imDim: Sequence[int], kernel: np.ndarray,
                                    kernelDim: Sequence[int], devCUDA: int) -> None:
        """``im`` (float32, C-contiguous, flat or [z,y,x]) is overwritten; dims are (z, y, x)."""
        if im.dtype != np.float32 or not im.flags.c_contiguous:
            raise TypeError("im must be a C-contiguous float32 array (JNA float[])")
        k = _f32(kernel)
        self.lib.convolution3DfftCUDAInPlace(im.ctypes.data_as(c_float_p), int3(imDim),
                                             k.ctypes.data_as(c_float_p), int3(kernelDim), devCUDA)

    def convolution3DfftCUDA(self, im: np.ndarray, imDim: Sequence[int], kernel: np.ndarray,
                             kernelDim: Sequence[int], devCUDA: int) -> Optional[np.ndarray]:
        a = _f32(im)
        k = _f32(kernel)
        ptr = self.lib.convolution3DfftCUDA(a.ctypes.data_as(c_float_p), int3(imDim),
                                            k.ctypes.data_as(c_float_p), int3(kernelDim), devCUDA)
        if not ptr:
            return None
        n = int(np.prod(imDim))
        out = np.ctypeslib.as_array(C.cast(ptr, c_float_p), shape=(n,)).copy().reshape(tuple(imDim))
        C.CDLL(None).free(C.c_void_p(ptr))
        return out

    def last_error(self) -> str:
        m = self.lib.spim_fftconv_last_error()
        return m.decode() if m else ""


def convolve(img: np.ndarray, kernel: np.ndarray, ext: int, value: float = 0.0, device: int = 0,
             lib: Optional[C.CDLL] = None) -> np.ndarray:
    """out = ext(img) (*) kernel on the GPU (``mvd_convolve``)."""
    lib = lib or load_library()
    a = _f32(img)
    k = _f32(kernel)
    out = np.empty_like(a)
    rc = lib.mvd_convolve(a.ctypes.data, int3(a.shape), k.ctypes.data, int3(k.shape), int(ext), float(value),
                          out.ctypes.data, device)
    check(lib, rc, "mvd_convolve")
    return out
