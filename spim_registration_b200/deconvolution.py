"""Host-side mirror of the reference's multi-view deconvolution classes.

The reference is Java and no JVM exists in this environment, so the host layer above the C-ABI
is written in Python with the same class names, constructor arguments, argument meaning and
error behaviour as the Java classes it mirrors (paths under /root/reference/src/main/java/):

  gen-1  mpicbg/spim/postprocessing/deconvolution2/{LRFFT,LRInput,BayesMVDeconvolution,Deconvolver}.java
  gen-2  spim/process/fusion/deconvolution/{MVDeconFFT,MVDeconInput,MVDeconvolution}.java

Volumes are numpy ``[z, y, x]`` float32 arrays (what the Java code hands to JNA after reversing
its (x,y,z) dims); ``blockSize`` arguments keep the reference's (x, y, z) order.

All arithmetic runs on the GPU through the shared library (session API for the iteration, the
legacy JNA entry for ``convolve1`` / ``convolve2``).  There is no CPU implementation here.
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import List, Optional, Sequence

import numpy as np

from . import native
from .blocks import (Block, BlockGeneratorFixedSizePrecise, divide_into_blocks, EXT_CONSTANT, EXT_MIRROR_SINGLE)


class PSFTYPE(enum.IntEnum):
    """LRFFT.java:54 / MVDeconFFT.java:49 -- ordinal order matters (dialog index 0..3)."""
    OPTIMIZATION_II = 0
    OPTIMIZATION_I = 1
    EFFICIENT_BAYESIAN = 2
    INDEPENDENT = 3


minValue = np.float32(0.0001)   # LRInput.java:30, MVDeconvolution.java:70


# --------------------------------------------------------------------------------------------------
# thin RAII wrapper of the session C-ABI
# --------------------------------------------------------------------------------------------------
class Session:
    def __init__(self, dims: Sequence[int], num_views: int, iteration_type: int, generation: int = 2,
                 lam: float = 0.006, osem_speedup: float = 1.0, osem_index: int = 0, device: int = 0,
                 conv1_ext: int = -1, conv2_ext: int = -1, haloed: bool = False, min_value: float = 0.0001,
                 exact_tikhonov: bool = False, fast_epilogue: Optional[bool] = None, lib: Optional[C.CDLL] = None):
        self.lib = lib or native.load_library()
        p = native.MvdParams()
        self.lib.mvd_params_default(C.byref(p))
        p.dims[:] = [int(d) for d in dims]
        p.num_views = int(num_views)
        p.iteration_type = int(iteration_type)
        p.generation = int(generation)
        p.lambda_ = float(lam)
        p.min_value = float(min_value)
        p.osem_speedup = float(osem_speedup)
        p.osem_index = int(osem_index)
        p.conv1_ext = int(conv1_ext)
        p.conv2_ext = int(conv2_ext)
        p.device = int(device)
        p.haloed = 1 if haloed else 0
        p.exact_tikhonov = 1 if exact_tikhonov else 0
        if fast_epilogue is not None:           # None = the library default (on, see spim_mvdecon.h)
            p.fast_epilogue = 1 if fast_epilogue else 0
        self.dims = tuple(int(d) for d in dims)
        self.num_views = int(num_views)
        self.psf_dims: List[tuple] = [()] * self.num_views
        self._h = C.c_void_p()
        native.check(self.lib, self.lib.mvd_session_create(C.byref(p), C.byref(self._h)), "mvd_session_create")

    def close(self):
        if getattr(self, "_h", None) and self._h:
            self.lib.mvd_session_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_view(self, v: int, img: np.ndarray, weight: Optional[np.ndarray], psf: np.ndarray):
        img = np.ascontiguousarray(img, dtype=np.float32)
        if img.shape != self.dims:
            raise ValueError(f"view {v}: image shape {img.shape} != session dims {self.dims}")
        w = None if weight is None else np.ascontiguousarray(weight, dtype=np.float32)
        if w is not None and w.shape != self.dims:
            raise ValueError(f"view {v}: weight shape {w.shape} != session dims {self.dims}")
        k = np.ascontiguousarray(psf, dtype=np.float32)
        self.psf_dims[v] = k.shape
        native.check(self.lib, self.lib.mvd_set_view(self._h, v, img.ctypes.data, None if w is None else w.ctypes.data,
                                                     k.ctypes.data, native.int3(k.shape)), "mvd_set_view")

    def set_view_ptr(self, v: int, img_ptr: int, weight_ptr: Optional[int], psf: np.ndarray):
        """Same with raw host pointers (e.g. pinned torch tensors' ``data_ptr()``)."""
        k = np.ascontiguousarray(psf, dtype=np.float32)
        self.psf_dims[v] = k.shape
        native.check(self.lib, self.lib.mvd_set_view(self._h, v, C.c_void_p(img_ptr),
                                                     C.c_void_p(weight_ptr) if weight_ptr else None,
                                                     k.ctypes.data, native.int3(k.shape)), "mvd_set_view")

    def upload_region(self, v: int, which: int, data: np.ndarray, lo: Sequence[int]):
        """Upload one cell ``data`` ([z, y, x]) of the view's image (which = 0) / weight (which = 1) at offset ``lo``."""
        a = np.ascontiguousarray(data, dtype=np.float32)
        native.check(self.lib, self.lib.mvd_upload_region(self._h, int(v), int(which), a.ctypes.data, native.int3(lo),
                                                          native.int3(a.shape)), "mvd_upload_region")

    def upload_region_ptr(self, v: int, which: int, ptr: int, lo: Sequence[int], ext: Sequence[int]):
        """Same from a raw pointer to a tightly packed [z, y, x] cell -- host memory, or device memory (a cell generated or
        fused on a GPU; mvd_upload_region copies with unified addressing)."""
        native.check(self.lib, self.lib.mvd_upload_region(self._h, int(v), int(which), C.c_void_p(int(ptr)), native.int3(lo),
                                                          native.int3(ext)), "mvd_upload_region")

    def init(self):
        native.check(self.lib, self.lib.mvd_init(self._h), "mvd_init")

    def run(self, n_iterations: int, stats: bool = True):
        n = n_iterations * self.num_views
        if stats and n > 0:
            s = np.zeros(n, dtype=np.float64)
            m = np.zeros(n, dtype=np.float64)
            native.check(self.lib, self.lib.mvd_run(self._h, n_iterations, s.ctypes.data_as(native.c_double_p),
                                                    m.ctypes.data_as(native.c_double_p)), "mvd_run")
            return s.reshape(n_iterations, self.num_views), m.reshape(n_iterations, self.num_views)
        native.check(self.lib, self.lib.mvd_run(self._h, n_iterations, None, None), "mvd_run")
        return None

    def finish(self):
        native.check(self.lib, self.lib.mvd_finish(self._h), "mvd_finish")

    def get_psi(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.dims, dtype=np.float32)
        native.check(self.lib, self.lib.mvd_get_psi(self._h, out.ctypes.data), "mvd_get_psi")
        return out

    def get_psi_ptr(self, host_ptr: int):
        native.check(self.lib, self.lib.mvd_get_psi(self._h, C.c_void_p(host_ptr)), "mvd_get_psi")

    def set_psi(self, psi: np.ndarray):
        a = np.ascontiguousarray(psi, dtype=np.float32)
        native.check(self.lib, self.lib.mvd_set_psi(self._h, a.ctypes.data), "mvd_set_psi")

    def get_kernel(self, v: int, which: int) -> np.ndarray:
        out = np.empty(self.psf_dims[v], dtype=np.float32)
        native.check(self.lib, self.lib.mvd_get_kernel(self._h, v, which, out.ctypes.data), "mvd_get_kernel")
        return out

    def info(self) -> native.MvdInfo:
        i = native.MvdInfo()
        native.check(self.lib, self.lib.mvd_get_info(self._h, C.byref(i)), "mvd_get_info")
        return i

    def sync(self):
        native.check(self.lib, self.lib.mvd_sync(self._h), "mvd_sync")

    def stream(self) -> int:
        st = C.c_void_p()
        native.check(self.lib, self.lib.mvd_get_stream(self._h, C.byref(st)), "mvd_get_stream")
        return st.value or 0

    def set_timing(self, on: bool):
        native.check(self.lib, self.lib.mvd_set_timing(self._h, 1 if on else 0), "mvd_set_timing")

    def get_timing(self):
        ms = (C.c_double * 8)()
        cnt = (C.c_longlong * 8)()
        native.check(self.lib, self.lib.mvd_get_timing(self._h, ms, cnt), "mvd_get_timing")
        return list(ms), list(cnt)

    # -- brick mode ---------------------------------------------------------------------------------
    def device_buffer(self, which: int):
        ptr = C.c_void_p()
        dims = (C.c_int * 3)()
        origin = (C.c_int * 3)()
        native.check(self.lib, self.lib.mvd_get_device_buffer(self._h, which, C.byref(ptr), dims, origin),
                     "mvd_get_device_buffer")
        return ptr.value, tuple(dims), tuple(origin)

    def set_halo_mask(self, lo_mask: int, hi_mask: int):
        native.check(self.lib, self.lib.mvd_set_halo_mask(self._h, lo_mask, hi_mask), "mvd_set_halo_mask")

    def halo_pack(self, which: int, regions, flat_ptr: int, unpack: bool = False):
        """regions: sequence of (z0, y0, x0, nz, ny, nx); flat_ptr: device pointer of the staging buffer."""
        n = len(regions)
        arr = (C.c_int * (6 * max(n, 1)))(*[int(v) for r in regions for v in r])
        fn = self.lib.mvd_halo_unpack if unpack else self.lib.mvd_halo_pack
        native.check(self.lib, fn(self._h, which, n, arr, C.c_void_p(flat_ptr)), "mvd_halo_unpack" if unpack else "mvd_halo_pack")

    # direct halo push over peer memory (spim_mvdecon.h, mvd_p2p_*)
    P2P_RECORD_BYTES = 288

    def p2p_export(self) -> bytes:
        rec = (C.c_ubyte * self.P2P_RECORD_BYTES)()
        native.check(self.lib, self.lib.mvd_p2p_export(self._h, rec), "mvd_p2p_export")
        return bytes(rec)

    def p2p_connect(self, records: Sequence[bytes], boxes, slots):
        """records[i]: the neighbour's export record; boxes[i] = (z0, y0, x0, nz, ny, nx, dz0, dy0, dx0); slots[i] =
        (flag raised at the neighbour, flag the neighbour raises here)."""
        n = len(records)
        blob = b"".join(records)
        assert len(blob) == n * self.P2P_RECORD_BYTES
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        barr = (C.c_int * (9 * n))(*[int(v) for b in boxes for v in b])
        sarr = (C.c_int * (2 * n))(*[int(v) for s_ in slots for v in s_])
        native.check(self.lib, self.lib.mvd_p2p_connect(self._h, n, buf, barr, sarr), "mvd_p2p_connect")

    def p2p_push(self, which: int):
        native.check(self.lib, self.lib.mvd_p2p_push(self._h, which), "mvd_p2p_push")

    def p2p_wait(self, which: int):
        native.check(self.lib, self.lib.mvd_p2p_wait(self._h, which), "mvd_p2p_wait")

    def p2p_timed_out(self) -> bool:
        t = C.c_int(0)
        native.check(self.lib, self.lib.mvd_p2p_status(self._h, C.byref(t)), "mvd_p2p_status")
        return bool(t.value)

    def p2p_disconnect(self):
        native.check(self.lib, self.lib.mvd_p2p_disconnect(self._h), "mvd_p2p_disconnect")

    def fill_halo(self, which: int, lo_mask: int, hi_mask: int):
        native.check(self.lib, self.lib.mvd_fill_halo(self._h, which, lo_mask, hi_mask), "mvd_fill_halo")

    def view_phase(self, view: int, phase: int, want_stats: bool = False):
        if want_stats:
            st = (C.c_double * 2)()
            native.check(self.lib, self.lib.mvd_view_phase(self._h, view, phase, st), "mvd_view_phase")
            return st[0], st[1]
        native.check(self.lib, self.lib.mvd_view_phase(self._h, view, phase, None), "mvd_view_phase")
        return None

    def init_partials(self) -> np.ndarray:
        p = np.zeros(6, dtype=np.float64)
        native.check(self.lib, self.lib.mvd_init_partials(self._h, p.ctypes.data_as(native.c_double_p)),
                     "mvd_init_partials")
        return p

    def set_avg(self, avg: float, osem: float = 1.0):
        native.check(self.lib, self.lib.mvd_set_avg(self._h, float(avg), float(osem)), "mvd_set_avg")


# --------------------------------------------------------------------------------------------------
# per-view operator: LRFFT (gen-1) / MVDeconFFT (gen-2)
# --------------------------------------------------------------------------------------------------
class _ViewFFT:
    """Common part of ``LRFFT`` (LRFFT.java:58-204) and ``MVDeconFFT`` (MVDeconFFT.java:53-172)."""

    #: ``public static CUDAFourierConvolution cuda`` (LRFFT.java:56, MVDeconFFT.java:51); assigned by
    #: the plugin before any view is constructed.  Lazily bound to the in-tree library.
    cuda: Optional[native.CUDAFourierConvolution] = None

    _generation = 2

    def __init__(self, image: np.ndarray, weight: Optional[np.ndarray], kernel: np.ndarray,
                 deviceList: Sequence[int] = (0,), useBlocks: bool = False,
                 blockSize: Optional[Sequence[int]] = None):
        self.image = np.ascontiguousarray(image, dtype=np.float32)
        self.weight = None if weight is None else np.ascontiguousarray(weight, dtype=np.float32)
        self.kernel1 = np.ascontiguousarray(kernel, dtype=np.float32)
        self.kernel2: Optional[np.ndarray] = None
        self.deviceList = [int(d) for d in deviceList]
        if any(d < 0 for d in self.deviceList):
            raise ValueError("device id -1 (CPU) is not supported: this implementation has no CPU path")
        self.device0 = self.deviceList[0]
        self.numViews = 0
        self.iterationType: Optional[PSFTYPE] = None
        self.views: Optional[List["_ViewFFT"]] = None
        self.i = -1
        n = self.image.shape                       # [z,y,x]
        k = self.kernel1.shape
        img_xyz = (n[2], n[1], n[0])
        k_xyz = (k[2], k[1], k[0])
        if useBlocks:
            if blockSize is None:
                raise ValueError("useBlocks requires blockSize")
            self.blockSize = tuple(int(b) for b in blockSize)
        else:
            # one single block for CUDA processing (LRFFT.java:179-190, MVDeconFFT.java:144-164)
            self.blockSize = tuple(img_xyz[d] + k_xyz[d] - 1 for d in range(3))
        self.useBlocks = True
        self.blocks = divide_into_blocks(img_xyz, self.blockSize, k_xyz, double_too_small=(self._generation == 1))
        if self.blocks is None:
            raise ValueError("Blocksize is smaller than the kernel (BlockGeneratorFixedSizePrecise.java:59-63)")
        self.blockSize = self.blocks[0].blockSize

    # LRFFT.java:206 / MVDeconFFT.java:174
    def setNumViews(self, numViews: int):
        self.numViews = int(numViews)

    def init(self, iterationType: PSFTYPE, views: List["_ViewFFT"]):
        """Records the iteration type; kernel2 is computed on the device when the deconvolution
        session initialises (LRFFT.java:214-325 / MVDeconFFT.java:183-323 run there)."""
        self.iterationType = PSFTYPE(iterationType)
        self.views = views
        if self.numViews == 0:
            print("Warning, numViews was not set.")
            self.numViews = 1

    def setImage(self, image):
        self.image = np.ascontiguousarray(image, dtype=np.float32)
        self.setCurrentIteration(-1)

    def setWeight(self, weight):
        self.weight = None if weight is None else np.ascontiguousarray(weight, dtype=np.float32)

    def setKernel(self, kernel):
        self.kernel1 = np.ascontiguousarray(kernel, dtype=np.float32)
        if self.views is not None and self.iterationType is not None:
            _init_views(self.views, self.iterationType, self._generation, self.device0)
        self.setCurrentIteration(-1)

    def getImage(self):
        return self.image

    def getWeight(self):
        return self.weight

    def getKernel1(self):
        return self.kernel1

    def getKernel2(self):
        return self.kernel2

    def setCurrentIteration(self, i: int):
        self.i = i

    def getCurrentIteration(self) -> int:
        return self.i

    @classmethod
    def _cuda(cls) -> native.CUDAFourierConvolution:
        if _ViewFFT.cuda is None:
            _ViewFFT.cuda = native.CUDAFourierConvolution()
        return _ViewFFT.cuda

    def _convolve_blocks(self, image: np.ndarray, kernel: np.ndarray, ext: int, value: float) -> np.ndarray:
        """The CUDA single-device path: per block copyBlock -> JNA -> pasteBlock
        (LRFFTThreads.java:61-92, MVDeconFFTThreads.java:73-114)."""
        cuda = self._cuda()
        image = np.ascontiguousarray(image, dtype=np.float32)
        result = np.empty_like(image)
        bs = self.blockSize
        block = np.empty((bs[2], bs[1], bs[0]), dtype=np.float32)
        kdim = kernel.shape
        def conv_block(buf, dev):
            # the JNA entry is void and leaves the buffer untouched on failure (the reference's contract); this host layer
            # must not paste an un-convolved block back as if it were the result, so it reads the error string the call left
            # behind (cleared by the library on success, thread-local like the call)
            cuda.convolution3DfftCUDAInPlace(buf, buf.shape, kernel, kdim, dev)
            err = cuda.last_error()
            if err:
                raise RuntimeError(f"convolution3DfftCUDAInPlace failed on device {dev}: {err}")

        if len(self.deviceList) == 1:
            for b in self.blocks:
                b.copyBlock(image, block, ext, value)
                conv_block(block, self.device0)
                b.pasteBlock(result, block)
            return result
        # multi-device mode (LRFFT.java:499-522, MVDeconFFT.java:447-469): one host thread per entry of deviceList, each
        # taking the next block index from a shared counter (the reference's AtomicInteger) until none are left; blocks
        # paste disjoint effective regions, and the native side serialises per device, not globally
        import itertools
        import threading
        counter, lock, errors = itertools.count(), threading.Lock(), []

        def worker(dev: int):
            try:
                mine = np.empty((bs[2], bs[1], bs[0]), dtype=np.float32)
                while True:
                    with lock:
                        i = next(counter)
                    if i >= len(self.blocks):
                        return
                    b = self.blocks[i]
                    b.copyBlock(image, mine, ext, value)
                    conv_block(mine, dev)
                    b.pasteBlock(result, mine)
            except BaseException as e:      # noqa: BLE001
                errors.append(e)

        threads = [threading.Thread(target=worker, args=(d,)) for d in self.deviceList]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return result

    def convolve1(self, image: np.ndarray, result: Optional[np.ndarray] = None) -> np.ndarray:
        """psi (*) kernel1 with mirror extension (LRFFT.java:423-526, MVDeconFFT.java:384-470)."""
        out = self._convolve_blocks(image, self.kernel1, EXT_MIRROR_SINGLE, 0.0)
        if result is not None:
            result[...] = out
            return result
        return out

    def convolve2(self, image: np.ndarray, result: Optional[np.ndarray] = None) -> np.ndarray:
        """ratio (*) kernel2 (LRFFT.java:550-626, MVDeconFFT.java:477-560); gen-2 extends with 1.0."""
        if self.kernel2 is None:
            raise RuntimeError("kernel2 not initialised: call LRInput.init / MVDeconInput.init first")
        if self._generation == 2:
            out = self._convolve_blocks(image, self.kernel2, EXT_CONSTANT, 1.0)
        else:
            out = self._convolve_blocks(image, self.kernel2, EXT_MIRROR_SINGLE, 0.0)
        if result is not None:
            result[...] = out
            return result
        return out


class LRFFT(_ViewFFT):
    """gen-1 view: ``LRFFT(image, weight, kernel, deviceList, useBlocks, blockSize)`` (LRFFT.java:131-199)."""
    _generation = 1
    PSFTYPE = PSFTYPE

    def clone(self) -> "LRFFT":
        v = LRFFT(self.image.copy(), None if self.weight is None else self.weight.copy(), self.kernel1.copy(),
                  self.deviceList, True, self.blockSize)
        return v


class MVDeconFFT(_ViewFFT):
    """gen-2 view: ``MVDeconFFT(image, weight, kernel, blockFactory, deviceList, useBlocks, blockSize,
    saveMemory)`` (MVDeconFFT.java:79-85); the ImgLib factory arguments have no meaning here and are
    accepted for signature compatibility."""
    _generation = 2
    PSFTYPE = PSFTYPE

    def __init__(self, image, weight, kernel, blockFactory=None, deviceList: Sequence[int] = (0,),
                 useBlocks: bool = False, blockSize: Optional[Sequence[int]] = None, saveMemory: bool = False):
        super().__init__(image, weight, kernel, deviceList, useBlocks, blockSize)
        self.saveMemory = saveMemory


def _init_views(views: List[_ViewFFT], iterationType: PSFTYPE, generation: int, device: int,
                lib: Optional[C.CDLL] = None) -> None:
    """``views.init(iterationType)``: normalise kernel1 and build kernel2 for every view in list order on
    the device (a tiny session that only runs the kernel construction)."""
    dims = views[0].image.shape
    with Session(dims, len(views), int(iterationType), generation=generation, device=device, lib=lib) as s:
        for i, v in enumerate(views):
            s.set_view(i, v.image, v.weight, v.kernel1)
        s.init()
        for i, v in enumerate(views):
            v.kernel1 = s.get_kernel(i, 1)
            v.kernel2 = s.get_kernel(i, 2)


class _Input:
    _generation = 2
    minValue = minValue

    def __init__(self):
        self.views: List[_ViewFFT] = []

    def add(self, view: _ViewFFT):
        """LRInput.java:33-39 -- re-broadcasts numViews to every view."""
        self.views.append(view)
        for v in self.views:
            v.setNumViews(self.getNumViews())

    def init(self, iterationType: PSFTYPE):
        """LRInput.java:47-53 / MVDeconInput.java:57-63."""
        for v in self.views:
            v.init(iterationType, self.views)
        _init_views(self.views, PSFTYPE(iterationType), self._generation, self.views[0].device0)
        return self

    def getViews(self) -> List[_ViewFFT]:
        return self.views

    def getNumViews(self) -> int:
        return len(self.views)


class LRInput(_Input):
    """gen-1 ``LRInput`` (LRInput.java:28-76)."""
    _generation = 1

    def clone(self) -> "LRInput":
        c = LRInput()
        for v in self.views:
            c.add(v.clone())
        return c


class MVDeconInput(_Input):
    """gen-2 ``MVDeconInput`` (MVDeconInput.java:31-80); ``imgFactory`` kept for signature parity."""
    _generation = 2

    def __init__(self, imgFactory=None):
        super().__init__()
        self._imgFactory = imgFactory

    def imgFactory(self):
        return self._imgFactory


# --------------------------------------------------------------------------------------------------
# the iteration: BayesMVDeconvolution (gen-1) / MVDeconvolution (gen-2)
# --------------------------------------------------------------------------------------------------
class Deconvolver:
    """``Deconvolver`` interface (Deconvolver.java:27-35)."""

    def getName(self) -> str:
        raise NotImplementedError

    def getAvg(self) -> float:
        raise NotImplementedError

    def getData(self):
        raise NotImplementedError

    def getPsi(self) -> np.ndarray:
        raise NotImplementedError

    def runIteration(self) -> None:
        raise NotImplementedError


class _Deconvolution(Deconvolver):
    _generation = 2
    #: static hooks of the Java classes (BayesMVDeconvolution.java:50-56, MVDeconvolution.java:62-69)
    initialImage: Optional[np.ndarray] = None
    checkNumbers = True
    debug = False
    debugInterval = 1
    collectStatistics = True
    minValue = minValue

    def __init__(self, views: _Input, iterationType: PSFTYPE, numIterations: int, lambda_: float,
                 osemspeedup: float = 1.0, osemspeedupindex: int = 0, name: str = "deconvolved"):
        self.name = name
        self.views = views
        self.data = views.getViews()
        self.numViews = len(self.data)
        if self.numViews == 0:
            raise ValueError("no views")
        self.lambda_ = float(lambda_)
        self.i = 0
        self.stats: List[tuple] = []
        dims = self.data[0].getImage().shape
        dev = self.data[0].device0
        if self._generation == 2:
            # gen-2 clamps the weights upstream (ProcessForDeconvolution.adjustForOSEM, :372-405); the
            # constructor's osemspeedup arguments are unused in MVDeconvolution.java
            osemspeedup, osemspeedupindex = 1.0, 0
        self._session = Session(dims, self.numViews, int(iterationType), generation=self._generation,
                                lam=self.lambda_, osem_speedup=float(osemspeedup),
                                osem_index=int(osemspeedupindex), device=dev)
        for v, view in enumerate(self.data):
            view.init(PSFTYPE(iterationType), self.data)
            self._session.set_view(v, view.getImage(), view.getWeight(), view.getKernel1())
        # views.init(iterationType) + psi initialisation happen on the device
        self._session.init()
        for v, view in enumerate(self.data):
            view.kernel1 = self._session.get_kernel(v, 1)
            view.kernel2 = self._session.get_kernel(v, 2)
        self.avg = float(self._session.info().avg)
        init_img = type(self).initialImage
        if init_img is not None:
            psi0 = np.array(init_img, dtype=np.float32, copy=True)
            if psi0.shape != tuple(dims):
                raise ValueError("initialImage has the wrong dimensions")
            if type(self).checkNumbers:
                # loadInitialImage: values <= 0 / NaN are replaced by minValue
                bad = ~(psi0 > 0)
                psi0[bad] = minValue
            self._session.set_psi(psi0)
        # run the deconvolution (both Java constructors iterate to completion)
        while self.i < numIterations:
            self.runIteration()
        if self._generation == 2:
            self._session.finish()          # "Masking never updated pixels." MVDeconvolution.java:201-208
        self._psi: Optional[np.ndarray] = None

    def getData(self):
        return self.views

    def getName(self) -> str:
        return self.name

    def getAvg(self) -> float:
        return self.avg

    def getCurrentIteration(self) -> int:
        return self.i

    def getPsi(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """psi as a host array; `out` (optional, C-contiguous float32 of the volume's shape -- e.g. pinned memory) receives
        the download instead of a freshly allocated array"""
        return self._session.get_psi(out)

    def runIteration(self) -> None:
        r = self._session.run(1, stats=type(self).collectStatistics)
        if r is not None:
            s, m = r
            for v in range(self.numViews):
                self.stats.append((self.i, v, float(s[0, v]), float(m[0, v])))
        self.i += 1

    def close(self):
        self._session.close()


class BayesMVDeconvolution(_Deconvolution):
    """gen-1 ``BayesMVDeconvolution(LRInput views, PSFTYPE iterationType, int numIterations, double lambda,
    double osemspeedup, int osemspeedupindex, String name)`` -- BayesMVDeconvolution.java:79-178."""
    _generation = 1


class MVDeconvolution(_Deconvolution):
    """gen-2 ``MVDeconvolution(MVDeconInput views, PSFTYPE iterationType, int numIterations, double lambda,
    double osemspeedup, int osemspeedupindex, String name)`` -- MVDeconvolution.java:94-211."""
    _generation = 2
    setBackgroundToAvg = False
