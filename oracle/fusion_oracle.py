"""CPU oracle for the fusion pre-step of the multi-view deconvolution -- TEST INFRASTRUCTURE ONLY.

SURVEY.md section 8(f) rows 1-3 ("next" rows: the callers on the input side of the hot path):
  rank 1  weight construction: cosine blending, sum-of-weights normalisation, OSEM clamp
  rank 2  view transformation: affine resampling of a raw stack into the bounding box
  rank 3  PSF pipeline: bead-averaged extraction, min-max normalisation, centre-preserving transform
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module; the
product package never does.

PARITY UNPINNED (same situation as ``mvdecon_oracle``): the reference holds no tests or golden vectors
for these functions and cannot run here (no JVM).  Everything whose source is in the reference tree is
restated line by line with citations.  Two pieces of arithmetic live in un-vendored jars and are restated
from their published source:
  * ``net.imglib2.interpolation.randomaccess.NLinearInterpolator3D.get()`` (imglib2 core, BOM
    pom-scijava 34.1.0, imglib2 6.1.0): weights in double, each of the 8 terms rounded to float by
    ``FloatType.mul(double)`` and accumulated in float in Gray-code order 000,100,110,010,011,111,101,001
    (x fastest bit first);
  * ``net.imglib2.realtransform.AffineTransform3D``: ``apply`` = ``s0*m00 + s1*m01 + s2*m02 + m03`` in
    double (the reference's own copy of this formula is at
    FW/TransformedInterpolatedRealRandomAccess.java:96-117), ``invert`` = adjugate / determinant,
    ``estimateBounds`` = min / max over the 8 transformed corners.

Conventions: volumes are ``numpy`` arrays indexed [z, y, x]; coordinates, offsets, borders and affine
matrices are in the reference's (x, y, z) order; an affine is the 12 row-packed doubles of
``AffineTransform3D.getRowPackedCopy()``.

Citations: FD/ = spim/process/fusion/deconvolution/, FW/ = spim/process/fusion/weights/,
F/ = spim/process/fusion/ (all under /root/reference/src/main/java/).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

MIN_VALUE = np.float32(0.0001)     # FD/MVDeconvolution.java:70

EXT_ZERO, EXT_CONSTANT, EXT_MIRROR_SINGLE, EXT_MIRROR_DOUBLE, EXT_PERIODIC = 0, 1, 2, 3, 4

f32 = np.float32
f64 = np.float64


# --------------------------------------------------------------------------------------
# affine helpers (imglib2-realtransform AffineTransform3D, restated)
# --------------------------------------------------------------------------------------

def invert_affine(m: Sequence[float]) -> np.ndarray:
    """``AffineTransform3D.invert()``: adjugate times 1/det, translation = -(R^-1 t)."""
    a = [float(v) for v in m]
    m00, m01, m02, m03, m10, m11, m12, m13, m20, m21, m22, m23 = a
    det = (m00 * m11 * m22 + m10 * m21 * m02 + m20 * m01 * m12
           - m02 * m11 * m20 - m12 * m21 * m00 - m22 * m01 * m10)
    if det == 0:
        raise ValueError("Matrix is singular.")
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
    return np.array([i00, i01, i02, i03, i10, i11, i12, i13, i20, i21, i22, i23], dtype=f64)


def apply_affine(m: Sequence[float], x, y, z):
    """``AffineTransform3D.apply(double[], double[])``: left-to-right double sums."""
    m = np.asarray(m, dtype=f64)
    x = np.asarray(x, dtype=f64); y = np.asarray(y, dtype=f64); z = np.asarray(z, dtype=f64)
    t0 = x * m[0] + y * m[1] + z * m[2] + m[3]
    t1 = x * m[4] + y * m[5] + z * m[6] + m[7]
    t2 = x * m[8] + y * m[9] + z * m[10] + m[11]
    return t0, t1, t2


def estimate_bounds(m: Sequence[float], dims_xyz: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
    """``AffineTransform3D.estimateBounds(interval)`` for the interval [0, dim-1]^3."""
    lo = np.full(3, np.finfo(f64).max)
    hi = -lo.copy()
    for cz in (0.0, float(dims_xyz[2] - 1)):
        for cy in (0.0, float(dims_xyz[1] - 1)):
            for cx in (0.0, float(dims_xyz[0] - 1)):
                t = np.array([float(v) for v in apply_affine(m, cx, cy, cz)])
                lo = np.minimum(lo, t)
                hi = np.maximum(hi, t)
    return lo, hi


# --------------------------------------------------------------------------------------
# out-of-bounds index rules (Views.extendMirrorSingle / extendPeriodic / extendZero / extendValue)
# --------------------------------------------------------------------------------------

def ext_index(a: np.ndarray, n: int, mode: int) -> np.ndarray:
    a = np.asarray(a, dtype=np.int64)
    if mode in (EXT_ZERO, EXT_CONSTANT):
        return np.where((a >= 0) & (a < n), a, -1)
    if mode == EXT_PERIODIC:
        return np.mod(a, n)
    if mode == EXT_MIRROR_SINGLE:
        if n == 1:
            return np.zeros_like(a)
        p = 2 * (n - 1)
        m = np.mod(a, p)
        return np.where(m < n, m, p - m)
    if mode == EXT_MIRROR_DOUBLE:
        p = 2 * n
        m = np.mod(a, p)
        return np.where(m < n, m, p - 1 - m)
    raise ValueError(mode)


def _sample(src: np.ndarray, ix, iy, iz, ext: int, value: float) -> np.ndarray:
    nz, ny, nx = src.shape
    jx, jy, jz = ext_index(ix, nx, ext), ext_index(iy, ny, ext), ext_index(iz, nz, ext)
    out = (jx < 0) | (jy < 0) | (jz < 0)
    v = src[np.where(out, 0, jz), np.where(out, 0, jy), np.where(out, 0, jx)].astype(f32)
    c = f32(0.0) if ext == EXT_ZERO else f32(value)
    return np.where(out, c, v)


def nlinear3d(src: np.ndarray, px, py, pz, ext: int, value: float = 0.0) -> np.ndarray:
    """``NLinearInterpolator3D.get()`` on ``Views.extend*(src)`` at double positions (px, py, pz)."""
    px = np.asarray(px, dtype=f64); py = np.asarray(py, dtype=f64); pz = np.asarray(pz, dtype=f64)
    fx, fy, fz = np.floor(px), np.floor(py), np.floor(pz)
    w0, w1, w2 = px - fx, py - fy, pz - fz
    w0n, w1n, w2n = 1.0 - w0, 1.0 - w1, 1.0 - w2
    ix, iy, iz = fx.astype(np.int64), fy.astype(np.int64), fz.astype(np.int64)

    def term(dx, dy, dz, w):
        v = _sample(src, ix + dx, iy + dy, iz + dz, ext, value)
        return (v.astype(f64) * w).astype(f32)           # FloatType.mul(double): (float)(get() * c)

    acc = term(0, 0, 0, w0n * w1n * w2n)
    for dx, dy, dz, w in ((1, 0, 0, w0 * w1n * w2n), (1, 1, 0, w0 * w1 * w2n), (0, 1, 0, w0n * w1 * w2n),
                          (0, 1, 1, w0n * w1 * w2), (1, 1, 1, w0 * w1 * w2), (1, 0, 1, w0 * w1n * w2),
                          (0, 0, 1, w0n * w1n * w2)):
        acc = (acc + term(dx, dy, dz, w)).astype(f32)    # FloatType.add: float + float
    return acc


# --------------------------------------------------------------------------------------
# rank 1: blending weights (FW/BlendingRealRandomAccess.java)
# --------------------------------------------------------------------------------------

def blending_lookup() -> np.ndarray:
    """static lookUp[1001], FW/BlendingRealRandomAccess.java:44-54 (d accumulates in double)."""
    lut = np.zeros(1001, dtype=f64)
    d = 0.0
    while d <= 1.0001:
        lut[int(math.floor(d * 1000.0 + 0.5))] = (math.cos((1 - d) * math.pi) + 1) / 2
        d = d + 0.001
    return lut


_LUT = blending_lookup()


def blending_weight(t0, t1, t2, dims_xyz: Sequence[int], border: Sequence[float], blending: Sequence[float],
                    min_xyz: Sequence[int] = (0, 0, 0)) -> np.ndarray:
    """``computeWeight`` FW/BlendingRealRandomAccess.java:91-121 at float locations (t0, t1, t2)."""
    loc = [np.asarray(t0, dtype=f32), np.asarray(t1, dtype=f32), np.asarray(t2, dtype=f32)]
    w = np.ones(loc[0].shape, dtype=f32)
    zero = np.zeros(loc[0].shape, dtype=bool)
    for d in range(3):
        l = (loc[d] - f32(min_xyz[d])).astype(f32)
        b = f32(border[d])
        a1 = (l - b).astype(f32)
        a2 = ((f32(dims_xyz[d] - 1) - l).astype(f32) - b).astype(f32)
        dist = np.maximum(f32(0), np.minimum(a1, a2)).astype(f32)
        zero |= (dist == 0)
        rel = (dist / f32(blending[d])).astype(f32)
        idx = np.floor(rel.astype(f64) * 1000.0 + 0.5)
        use = (rel < 1) & ~np.isnan(rel)
        idx = np.where(use, idx, 0).astype(np.int64)
        idx = np.clip(idx, 0, 1000)
        w = np.where(use, (w.astype(f64) * _LUT[idx]).astype(f32), w)   # float *= double
    return np.where(zero, f32(0), w).astype(f32)


def inverse_positions(out_dims_zyx: Sequence[int], offset_xyz: Sequence[int], inv: Sequence[float]):
    """cursor.localize(s); s += offset; transform.applyInverse(t, s) -- FD/TransformInput.java:93-100
    (= FW/TransformedInterpolatedRealRandomAccess.java:96-117): double sums, cast to float."""
    nz, ny, nx = out_dims_zyx
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    s0 = (x.astype(f32) + f32(offset_xyz[0])).astype(f32)
    s1 = (y.astype(f32) + f32(offset_xyz[1])).astype(f32)
    s2 = (z.astype(f32) + f32(offset_xyz[2])).astype(f32)
    t0, t1, t2 = apply_affine(inv, s0, s1, s2)
    return t0.astype(f32), t1.astype(f32), t2.astype(f32)


def intersects(t0, t1, t2, dims_xyz) -> np.ndarray:
    """F/FusionHelper.java:75-81."""
    return (t0 >= 0) & (t1 >= 0) & (t2 >= 0) & (t0 < dims_xyz[0]) & (t1 < dims_xyz[1]) & (t2 < dims_xyz[2])


# --------------------------------------------------------------------------------------
# rank 2: view transformation (FD/TransformInput.java, TransformInputAndWeights.java, TransformWeights.java)
# --------------------------------------------------------------------------------------

def transform_input(stack: np.ndarray, inv: Sequence[float], out_dims_zyx: Sequence[int],
                    offset_xyz: Sequence[int]) -> np.ndarray:
    """FD/TransformInput.java:70-116: tri-linear on Views.extendMirrorSingle, max(minValue, .) inside the
    stack, 0 (untouched) outside."""
    t0, t1, t2 = inverse_positions(out_dims_zyx, offset_xyz, inv)
    dims_xyz = stack.shape[::-1]
    inside = intersects(t0, t1, t2, dims_xyz)
    v = nlinear3d(stack, t0, t1, t2, EXT_MIRROR_SINGLE)
    v = np.maximum(MIN_VALUE, v)            # Math.max(minValue, v): NaN propagates in Java and in np.maximum
    return np.where(inside, v, f32(0)).astype(f32)


def transform_weights(stack_dims_zyx: Sequence[int], inv: Sequence[float], out_dims_zyx: Sequence[int],
                      offset_xyz: Sequence[int], border: Sequence[float], blending: Sequence[float]) -> np.ndarray:
    """FD/TransformWeights.java:73-111 / the weight half of TransformInputAndWeights.java:127-134: blending
    weight at the back-projected position of every output voxel ("the border can be negative")."""
    t0, t1, t2 = inverse_positions(out_dims_zyx, offset_xyz, inv)
    return blending_weight(t0, t1, t2, tuple(stack_dims_zyx[::-1]), border, blending)


def transform_input_and_weights(stack, inv, out_dims_zyx, offset_xyz, border, blending):
    """FD/TransformInputAndWeights.java:76-135."""
    return (transform_input(stack, inv, out_dims_zyx, offset_xyz),
            transform_weights(stack.shape, inv, out_dims_zyx, offset_xyz, border, blending))


def loader_normalize(img: np.ndarray) -> np.ndarray:
    """spim/fiji/spimdata/imgloaders/AbstractImgLoader.java:164-184: (v - min) / (max - min) in float."""
    a = np.asarray(img, dtype=f32)
    mn, mx = a.min(), a.max()
    return ((a - mn).astype(f32) / f32(mx - mn)).astype(f32)


# --------------------------------------------------------------------------------------
# rank 1 (continued): FD/WeightNormalizer.java, FW/NormalizingRandomAccess.java, OSEM clamp
# --------------------------------------------------------------------------------------

def divide_into_portions(size: int, num_portions: int) -> List[Tuple[int, int]]:
    """F/FusionHelper.java:257-280 -> [(start, loop_size)]."""
    chunk, mod = size // num_portions, size % num_portions
    return [(p * chunk, chunk + (mod if p == num_portions - 1 else 0)) for p in range(num_portions)]


def _overlap_stats(weights: Sequence[np.ndarray], num_portions: int) -> Tuple[int, float]:
    """min / avg number of views with w > 0, combined over portions as FD/WeightNormalizer.java:94-108
    (minimum of the portion minima, mean of the portion means)."""
    count = np.zeros(weights[0].shape, dtype=np.int64)
    for w in weights:
        count += (w > 0)
    flat = count.reshape(-1)
    mn, avg = len(weights), 0.0
    portions = divide_into_portions(flat.size, num_portions)
    for start, loop in portions:
        c = flat[start:start + loop]
        pmin = min(len(weights), int(c.min())) if loop > 0 else len(weights)
        mn = min(mn, int(round(float(pmin))))
        avg += float(c.sum()) / float(loop) if loop > 0 else float("nan")
    return mn, avg / len(portions)


def weight_normalizer_direct(weights: Sequence[np.ndarray], num_portions: int = 1):
    """``new WeightNormalizer(weights).process()`` -> ApplyDirectly, FD/WeightNormalizer.java:132-178:
    w_v <- (float)(w_v / sumW) for EVERY voxel (the ``if (sumW > 1)`` is commented out; 0/0 = NaN where no
    view has weight).  Returns (weights, minOverlappingViews, avgOverlappingViews)."""
    sum_w = np.zeros(weights[0].shape, dtype=f64)
    for w in weights:
        sum_w = sum_w + w.astype(f64)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = [(w.astype(f64) / sum_w).astype(f32) for w in weights]
    mn, avg = _overlap_stats(weights, num_portions)
    return out, mn, avg


def weight_normalizer_virtual(weights: Sequence[np.ndarray], num_portions: int = 1):
    """``new WeightNormalizer(weights, factory).process()`` -> ComputeSumImage, FD/WeightNormalizer.java:180-253:
    sumWeights = sumW > 1 ? (float)sumW : 1.  Returns (sumWeights, min, avg)."""
    sum_w = np.zeros(weights[0].shape, dtype=f64)
    for w in weights:
        sum_w = sum_w + w.astype(f64)
    s = np.where(sum_w > 1, sum_w.astype(f32), f32(1)).astype(f32)
    mn, avg = _overlap_stats(weights, num_portions)
    return s, mn, avg


def normalizing_access(w: np.ndarray, sum_weights: np.ndarray, osem: float = 1.0) -> np.ndarray:
    """FW/NormalizingRandomAccess.java:57-66: (float) min(1, (w / sumWeights) * osemspeedup) in double."""
    v = w.astype(f64) / sum_weights.astype(f64)
    return np.minimum(1.0, v * float(osem)).astype(f32)


def adjust_for_osem(weights: Sequence[np.ndarray], osem: float) -> List[np.ndarray]:
    """FD/ProcessForDeconvolution.java:372-385 (precomputed weights): min(1, w * (float)osem) in float."""
    if osem == 1.0:
        return [w.copy() for w in weights]
    with np.errstate(invalid="ignore"):
        # Math.min(1, NaN) = NaN in Java; np.minimum propagates NaN as well
        return [np.minimum(f32(1), (w * f32(osem)).astype(f32)).astype(f32) for w in weights]


# --------------------------------------------------------------------------------------
# rank 3: PSF pipeline (FD/ExtractPSF.java)
# --------------------------------------------------------------------------------------

def extract_psf_local(img: np.ndarray, locations_xyz: Sequence[Sequence[float]], size_xyz: Sequence[int]) -> np.ndarray:
    """FD/ExtractPSF.java:374-415: sum over beads of the tri-linearly interpolated neighbourhood
    (Views.extendPeriodic), accumulated in float in list order."""
    sx, sy, sz = [int(v) for v in size_xyz]
    z, y, x = np.meshgrid(np.arange(sz), np.arange(sy), np.arange(sx), indexing="ij")
    psf = np.zeros((sz, sy, sx), dtype=f32)
    for p in locations_xyz:
        px = (x - sx // 2).astype(f64) + float(p[0])
        py = (y - sy // 2).astype(f64) + float(p[1])
        pz = (z - sz // 2).astype(f64) + float(p[2])
        psf = (psf + nlinear3d(img, px, py, pz, EXT_PERIODIC)).astype(f32)
    return psf


def psf_normalize(psf: np.ndarray) -> np.ndarray:
    """FD/ExtractPSF.java:298-316: (v - min) / (max - min) in double, stored as float."""
    a = psf.astype(f64)
    mn, mx = a.min(), a.max()
    with np.errstate(divide="ignore", invalid="ignore"):
        return ((a - mn) / (mx - mn)).astype(f32)


def transform(image: np.ndarray, model: Sequence[float], new_dim_xyz: Sequence[int], offset_xyz: Sequence[float]) -> np.ndarray:
    """FD/ExtractPSF.java:417-457: resample ``image`` under ``model`` into a new array (zero extension,
    tri-linear, all position arithmetic in double)."""
    inv = invert_affine(model)
    nx, ny, nz = [int(v) for v in new_dim_xyz]
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    t0, t1, t2 = apply_affine(inv, x.astype(f64) + float(offset_xyz[0]), y.astype(f64) + float(offset_xyz[1]),
                              z.astype(f64) + float(offset_xyz[2]))
    return nlinear3d(image, t0, t1, t2, EXT_ZERO)


def transform_psf_geometry(psf_dims_zyx: Sequence[int], model: Sequence[float]):
    """Size and offset of the transformed PSF, FD/ExtractPSF.java:325-367: odd size (int)extent + 1, the
    transformed centre voxel stays the centre."""
    dims_xyz = tuple(int(v) for v in psf_dims_zyx[::-1])
    lo, hi = estimate_bounds(model, dims_xyz)
    center = [float(d // 2) for d in dims_xyz]
    tmp = [float(v) for v in apply_affine(model, *center)]
    new_size, offset = [], []
    for d in range(3):
        size = hi[d] - lo[d]
        ns = int(size) + 1
        if ns % 2 == 0:
            ns += 1
        new_size.append(ns)
        offset.append(tmp[d] - float(ns // 2))
    return new_size, offset


def transform_psf(psf: np.ndarray, model: Sequence[float]) -> np.ndarray:
    new_size, offset = transform_psf_geometry(psf.shape, model)
    return transform(psf, model, new_size, offset)


def make_same_size(img: np.ndarray, size_xyz: Sequence[int]) -> np.ndarray:
    """FD/ExtractPSF.java:466-496: centre ``img`` in an array of ``size``, padding with its minimum."""
    sx, sy, sz = [int(v) for v in size_xyz]
    nz, ny, nx = img.shape
    mn = f32(img.astype(f64).min())
    z, y, x = np.meshgrid(np.arange(sz), np.arange(sy), np.arange(sx), indexing="ij")
    return _sample(img, x - sx // 2 + nx // 2, y - sy // 2 + ny // 2, z - sz // 2 + nz // 2, EXT_CONSTANT, mn)


def common_size(images: Sequence[np.ndarray]) -> List[int]:
    """FD/ExtractPSF.java:505-517 -> (x, y, z) maxima."""
    s = [0, 0, 0]
    for im in images:
        for d in range(3):
            s[d] = max(s[d], im.shape[2 - d])
    return s


def extract_next_img(img: np.ndarray, model: Sequence[float], locations_xyz, psf_size_xyz):
    """FD/ExtractPSF.java:277-296 -> (transformed PSF, original PSF)."""
    original = psf_normalize(extract_psf_local(img, locations_xyz, psf_size_xyz))
    return transform_psf(original, model), original
