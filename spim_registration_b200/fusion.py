"""Host-side mirror of the reference's fusion pre-step for the multi-view deconvolution
(SURVEY.md section 8f ranks 1-3), over the device-side C-ABI of include/spim_fusion.h.

Mirrored classes (paths under /root/reference/src/main/java/):
  spim/process/fusion/deconvolution/ProcessForDeconvolution.java   (fuseStacksAndGetPSFs, WeightType, adjustForOSEM)
  spim/process/fusion/deconvolution/ExtractPSF.java                (extractNextImg, transformPSF, makeSameSize, commonSize)
  spim/process/fusion/weights/Blending.java                        (interval + border + range)
  net.imglib2.realtransform.AffineTransform3D                      (only set / inverse / getRowPackedCopy / apply, host logic)

Volumes are numpy [z, y, x] float32 arrays; coordinates, offsets, borders and affine matrices keep the
reference's (x, y, z) order.  All per-voxel arithmetic (affine resampling, tri-linear interpolation,
blending, weight normalisation, bead averaging) runs on the GPU inside the session; this module only holds
the reference's control flow and its PSF-sized bookkeeping.  There is no CPU implementation here.
"""
from __future__ import annotations

import ctypes as C
import enum
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import native
from .deconvolution import Session


class WeightType(enum.IntEnum):
    """ProcessForDeconvolution.java:81 (ordinal order)."""
    WEIGHTS_ONLY = 0
    NO_WEIGHTS = 1
    VIRTUAL_WEIGHTS = 2
    PRECOMPUTED_WEIGHTS = 3
    LOAD_WEIGHTS = 4


class AffineTransform3D:
    """The few methods of net.imglib2.realtransform.AffineTransform3D the fusion pre-step uses."""

    def __init__(self, rowPacked: Optional[Sequence[float]] = None):
        self.m = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float64)
        if rowPacked is not None:
            self.set(*rowPacked)

    def set(self, *values: float) -> None:
        if len(values) != 12:
            raise ValueError("AffineTransform3D.set needs 12 row-packed values")
        self.m = np.array([float(v) for v in values], dtype=np.float64)

    def getRowPackedCopy(self) -> np.ndarray:
        return self.m.copy()

    def inverse(self) -> "AffineTransform3D":
        m00, m01, m02, m03, m10, m11, m12, m13, m20, m21, m22, m23 = [float(v) for v in self.m]
        det = (m00 * m11 * m22 + m10 * m21 * m02 + m20 * m01 * m12
               - m02 * m11 * m20 - m12 * m21 * m00 - m22 * m01 * m10)
        if det == 0:
            raise RuntimeError("Matrix is singular.")
        idet = 1.0 / det
        i00 = (m11 * m22 - m12 * m21) * idet
        i01 = (m02 * m21 - m01 * m22) * idet
        i02 = (m01 * m12 - m02 * m11) * idet
        i10 = (m12 * m20 - m10 * m22) * idet
        i11 = (m00 * m22 - m02 * m20) * idet
        i12 = (m02 * m10 - m00 * m12) * idet
        i20 = (m10 * m21 - m11 * m20) * idet
        i21 = (m01 * m20 - m00 * m21) * idet
        i22 = (m00 * m11 - m01 * m10) * idet
        i03 = -i00 * m03 - i01 * m13 - i02 * m23
        i13 = -i10 * m03 - i11 * m13 - i12 * m23
        i23 = -i20 * m03 - i21 * m13 - i22 * m23
        return AffineTransform3D([i00, i01, i02, i03, i10, i11, i12, i13, i20, i21, i22, i23])

    def apply(self, source: Sequence[float]) -> List[float]:
        m, s = self.m, [float(v) for v in source]
        return [s[0] * m[4 * r] + s[1] * m[4 * r + 1] + s[2] * m[4 * r + 2] + m[4 * r + 3] for r in range(3)]


class Blending:
    """spim/process/fusion/weights/Blending.java: the interval is the raw stack, (x, y, z) sizes."""

    def __init__(self, interval_xyz: Sequence[int], border: Sequence[float], blending: Sequence[float]):
        self.interval = tuple(int(v) for v in interval_xyz)
        self.border = [float(v) for v in border]
        self.blending = [float(v) for v in blending]


def _dbl12(values) -> "C.Array":
    return (C.c_double * 12)(*[float(v) for v in values])


# -- thin wrappers of the C-ABI, bound to a Session -----------------------------------------------------

def load_stack(s: Session, stack: Optional[np.ndarray], normalize: bool = False) -> None:
    if stack is None:
        native.check(s.lib, s.lib.mvd_load_stack(s._h, None, None, 0), "mvd_load_stack")
        return
    a = np.ascontiguousarray(stack, dtype=np.float32)
    if a.ndim != 3:
        raise ValueError("stack must be a [z, y, x] volume")
    native.check(s.lib, s.lib.mvd_load_stack(s._h, a.ctypes.data, native.int3(a.shape), 1 if normalize else 0), "mvd_load_stack")


def transform_view(s: Session, view: int, transform: AffineTransform3D, offset_xyz: Sequence[int],
                   blending: Optional[Blending] = None, want_image: bool = True) -> None:
    t = native.MvdTransform()
    t.struct_size = C.sizeof(native.MvdTransform)
    t.inverse[:] = list(transform.inverse().getRowPackedCopy())
    t.offset[:] = [int(v) for v in offset_xyz]
    t.want_image = 1 if want_image else 0
    t.want_weight = 1 if blending is not None else 0
    if blending is not None:
        t.border[:] = blending.border
        t.range[:] = blending.blending
    native.check(s.lib, s.lib.mvd_transform_view(s._h, int(view), C.byref(t)), "mvd_transform_view")


def set_psf(s: Session, view: int, psf: np.ndarray) -> None:
    k = np.ascontiguousarray(psf, dtype=np.float32)
    s.psf_dims[view] = k.shape
    native.check(s.lib, s.lib.mvd_set_psf(s._h, int(view), k.ctypes.data, native.int3(k.shape)), "mvd_set_psf")


def normalize_weights(s: Session, virtual: bool, num_portions: int = 1):
    mn = C.c_int()
    avg = C.c_double()
    native.check(s.lib, s.lib.mvd_normalize_weights(s._h, 1 if virtual else 0, int(num_portions), C.byref(mn), C.byref(avg)),
                 "mvd_normalize_weights")
    return mn.value, avg.value


def get_view(s: Session, view: int, which: int) -> np.ndarray:
    out = np.empty(s.dims, dtype=np.float32)
    native.check(s.lib, s.lib.mvd_get_view(s._h, int(view), int(which), out.ctypes.data), "mvd_get_view")
    return out


def blending_lookup(lib: Optional[C.CDLL] = None) -> np.ndarray:
    lib = lib or native.load_library()
    out = np.zeros(1001, dtype=np.float64)
    native.check(lib, lib.mvd_blending_lookup(out.ctypes.data_as(native.c_double_p)), "mvd_blending_lookup")
    return out


# -- ExtractPSF -------------------------------------------------------------------------------------------

class ExtractPSF:
    """spim/process/fusion/deconvolution/ExtractPSF.java.  View ids are any hashable key."""

    def __init__(self, lib: Optional[C.CDLL] = None, device: int = 0):
        self.lib = lib or native.load_library()
        self.device = device
        self.pointSpreadFunctions: Dict[object, np.ndarray] = {}
        self.originalPSFs: Dict[object, np.ndarray] = {}
        self.viewIds: List[object] = []
        self.mapViewIds: Dict[object, object] = {}

    def getViewIdMapping(self):
        return self.mapViewIds

    def getPSFMap(self):
        return self.pointSpreadFunctions

    def getTransformedPSF(self, viewId) -> np.ndarray:
        if viewId in self.pointSpreadFunctions:
            return self.pointSpreadFunctions[viewId]
        if viewId in self.mapViewIds:
            return self.pointSpreadFunctions[self.mapViewIds[viewId]]
        raise RuntimeError(f"Cannot find PSF for view {viewId!r}")

    def getInputCalibrationPSFs(self):
        return self.originalPSFs

    def getViewIdsForPSFs(self):
        return self.viewIds

    def extractNextImg(self, session: Session, viewId, model: AffineTransform3D, locations: Sequence[Sequence[float]],
                       psfSize: Sequence[int]) -> None:
        """ExtractPSF.java:277-296 on the stack currently loaded in ``session``; psfSize is (x, y, z)."""
        original = self.extractPSFLocal(session, locations, psfSize, normalize=True)
        psf = self.transformPSF(original, model, lib=self.lib, device=self.device)
        self.viewIds.append(viewId)
        self.pointSpreadFunctions[viewId] = psf
        self.originalPSFs[viewId] = original

    @staticmethod
    def extractPSFLocal(session: Session, locations: Sequence[Sequence[float]], size: Sequence[int], normalize: bool = False) -> np.ndarray:
        loc = np.ascontiguousarray(np.asarray(locations, dtype=np.float64).reshape(-1, 3))
        size_zyx = (int(size[2]), int(size[1]), int(size[0]))
        out = np.empty(size_zyx, dtype=np.float32)
        native.check(session.lib, session.lib.mvd_extract_psf(session._h, loc.shape[0], loc.ctypes.data_as(native.c_double_p),
                                                              native.int3(size_zyx), 1 if normalize else 0, out.ctypes.data),
                     "mvd_extract_psf")
        return out

    @staticmethod
    def transformPSF(psf: np.ndarray, model: AffineTransform3D, lib: Optional[C.CDLL] = None, device: int = 0) -> np.ndarray:
        """ExtractPSF.java:325-367: odd-sized output whose centre is the transformed centre of ``psf``."""
        lib = lib or native.load_library()
        a = np.ascontiguousarray(psf, dtype=np.float32)
        od = (C.c_int * 3)()
        off = (C.c_double * 3)()
        m = _dbl12(model.getRowPackedCopy())
        native.check(lib, lib.mvd_transform_psf_size(native.int3(a.shape), m, od, off), "mvd_transform_psf_size")
        out = np.empty(tuple(od), dtype=np.float32)
        native.check(lib, lib.mvd_transform_psf(a.ctypes.data, native.int3(a.shape), m, _dbl12(model.inverse().getRowPackedCopy()),
                                                out.ctypes.data, od, device), "mvd_transform_psf")
        return out

    @classmethod
    def loadAndTransformPSFs(cls, filenames: Dict[object, str], viewIds: Sequence[object],
                             models: Optional[Dict[object, AffineTransform3D]] = None, lib: Optional[C.CDLL] = None,
                             device: int = 0) -> "ExtractPSF":
        """ExtractPSF.java:541-598: one TIFF per view (``filenames[viewId]``), transformed with the view's model when
        ``models`` is given, used as it is otherwise."""
        from .export import read_tiff_stack
        e = cls(lib=lib, device=device)
        for vd in viewIds:
            if vd not in filenames or not os.path.exists(filenames[vd]):
                raise RuntimeError(f"Could not load '{filenames.get(vd)}' (should be a TIFF file).")
            psfImage = read_tiff_stack(filenames[vd])
            psf = cls.transformPSF(psfImage, models[vd], lib=e.lib, device=device) if models is not None else psfImage.copy()
            e.viewIds.append(vd)
            e.pointSpreadFunctions[vd] = psf
            e.originalPSFs[vd] = psfImage
        return e

    @staticmethod
    def makeSameSize(img: np.ndarray, sizeIn: Sequence[int]) -> np.ndarray:
        """ExtractPSF.java:466-496 (PSF-sized copy, host; the gen-1 plugin applies it with commonSize,
        fiji/plugin/Multi_View_Deconvolution.java:133-143): centre ``img`` in an (x, y, z)-sized array padded with its minimum."""
        sx, sy, sz = [int(v) for v in sizeIn]
        nz, ny, nx = img.shape
        out = np.full((sz, sy, sx), np.float32(img.astype(np.float64).min()), dtype=np.float32)
        # square position q reads input position q - size/2 + dim/2
        (z0, z1, a0), (y0, y1, b0), (x0, x1, c0) = ExtractPSF._overlap(sz, nz), ExtractPSF._overlap(sy, ny), ExtractPSF._overlap(sx, nx)
        if z1 > z0 and y1 > y0 and x1 > x0:
            out[z0:z1, y0:y1, x0:x1] = img[a0:a0 + (z1 - z0), b0:b0 + (y1 - y0), c0:c0 + (x1 - x0)]
        return out

    @staticmethod
    def _overlap(s: int, n: int):
        shift = -(s // 2) + n // 2              # input index = q + shift
        lo = max(0, -shift)
        hi = min(s, n - shift)
        return lo, hi, lo + shift

    @staticmethod
    def commonSize(images: Sequence[np.ndarray]) -> Optional[List[int]]:
        """ExtractPSF.java:505-517 -> (x, y, z)."""
        if not images:
            return None
        size = [0, 0, 0]
        for im in images:
            for d in range(3):
                size[d] = max(size[d], im.shape[2 - d])
        return size

    def computeMaxDimTransformedPSF(self) -> List[int]:
        return self.commonSize(list(self.pointSpreadFunctions.values())) or [0, 0, 0]


# -- ProcessForDeconvolution --------------------------------------------------------------------------------

class ProcessForDeconvolution:
    """spim/process/fusion/deconvolution/ProcessForDeconvolution.java, driving a device-resident Session.

    ``bb_min`` is the bounding box minimum (x, y, z); its dimensions are the session's dims.  Instead of SpimData
    view descriptions, ``fuseStacksAndGetPSFs`` takes the raw stacks and their registrations directly."""

    def __init__(self, session: Session, bb_min: Sequence[int], blendingBorder: Sequence[int], blendingRange: Sequence[int],
                 numThreads: int = 1):
        self.session = session
        self.bb_min = [int(v) for v in bb_min]
        self.blendingBorder = [int(v) for v in blendingBorder]
        self.blendingRange = [int(v) for v in blendingRange]
        self.numThreads = int(numThreads)
        self.ePSF: Optional[ExtractPSF] = None
        self.minOverlappingViews = 0
        self.avgOverlappingViews = 0.0
        self.osemspeedup = 1.0

    def getExtractPSF(self):
        return self.ePSF

    def getMinOverlappingViews(self):
        return self.minOverlappingViews

    def getAvgOverlappingViews(self):
        return self.avgOverlappingViews

    def getBlending(self, interval_xyz: Sequence[int]) -> Blending:
        """ProcessForDeconvolution.java:566-580."""
        return Blending(interval_xyz, [float(v) for v in self.blendingBorder], [float(v) for v in self.blendingRange])

    def fuseStacksAndGetPSFs(self, stacks: Sequence[np.ndarray], transforms: Sequence[AffineTransform3D], osemIndex: int,
                             osemspeedup: float, weightType: WeightType, psfs: Optional[Sequence[np.ndarray]] = None,
                             beadLocations: Optional[Sequence[Sequence[Sequence[float]]]] = None,
                             psfSize: Optional[Sequence[int]] = None, normalizeStacks: bool = True) -> bool:
        """ProcessForDeconvolution.java:129-366.  Either ``psfs`` (already transformed, one per view) or
        ``beadLocations`` + ``psfSize`` (extraction from the stacks) supplies the kernels."""
        s = self.session
        V = len(stacks)
        if V == 0 or V != s.num_views or len(transforms) != V:
            return False
        if weightType == WeightType.LOAD_WEIGHTS:
            raise RuntimeError(weightType.name + " not implemented yet.")
        extract = beadLocations is not None
        if extract:
            if psfSize is None:
                raise ValueError("psfSize is required to extract PSFs")
            self.ePSF = ExtractPSF(lib=s.lib)
        elif psfs is None and weightType != WeightType.WEIGHTS_ONLY:
            return False
        for i in range(V):
            stack = np.ascontiguousarray(stacks[i], dtype=np.float32)
            load_stack(s, stack, normalize=normalizeStacks)          # ProcessFusion.getImage(..., normalize = true)
            interval_xyz = stack.shape[::-1]
            if weightType in (WeightType.PRECOMPUTED_WEIGHTS, WeightType.VIRTUAL_WEIGHTS):
                transform_view(s, i, transforms[i], self.bb_min, self.getBlending(interval_xyz), want_image=True)
            elif weightType == WeightType.WEIGHTS_ONLY:
                transform_view(s, i, transforms[i], self.bb_min, self.getBlending(interval_xyz), want_image=False)
            else:
                transform_view(s, i, transforms[i], self.bb_min, None, want_image=True)
            if extract:
                self.ePSF.extractNextImg(s, i, transforms[i], beadLocations[i], psfSize)
        load_stack(s, None)
        if extract:
            # EfficientBayesianBased.java:271-275: pfd.getExtractPSF().getTransformedPSF( vd ) goes to MVDeconFFT as it is
            for i in range(V):
                set_psf(s, i, self.ePSF.getTransformedPSF(i))
        elif psfs is not None:
            for i in range(V):
                set_psf(s, i, psfs[i])
        if weightType in (WeightType.PRECOMPUTED_WEIGHTS, WeightType.WEIGHTS_ONLY, WeightType.VIRTUAL_WEIGHTS):
            mn, avg = normalize_weights(s, virtual=(weightType == WeightType.VIRTUAL_WEIGHTS), num_portions=self.numThreads * 2)
            self.minOverlappingViews = max(1, mn)
            self.avgOverlappingViews = max(1.0, avg)
        if osemIndex == 1:
            osemspeedup = self.getMinOverlappingViews()
        elif osemIndex == 2:
            osemspeedup = self.getAvgOverlappingViews()
        self.osemspeedup = float(osemspeedup)
        return True
