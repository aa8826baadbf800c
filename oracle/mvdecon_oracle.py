"""CPU oracle for the multiview Bayesian deconvolution hot path -- TEST INFRASTRUCTURE ONLY.

This module restates, in NumPy/SciPy, the algorithm of fiji/SPIM_Registration's
multi-view deconvolution.  It is the checker the
CUDA path is compared against.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the product
package ``spim_registration_b200`` never does.

PARITY UNPINNED: the reference repository holds no tests, fixtures, golden vectors or
known-answer values for this path, the reference is Java and no
JVM exists in this environment, and its FFT convolution arithmetic lives in un-vendored
third-party jars (ImgLib1 ``FourierConvolution``, ImgLib2 ``FFTConvolution`` from
imglib2-algorithm via pom-scijava 34.1.0 / imglib2 6.1.0, Mines-JTK ``FftReal``).  The
convolution is therefore restated from its published mathematical definition (linear
convolution with a stated out-of-bounds rule, kernel origin at ``dim/2``) and pinned by
self-consistency checks in ``tests/test_oracle.py`` (direct-sum vs FFT, blocked vs
unblocked, K2 identities, Richardson-Lucy fixed point, Tikhonov closed form).

Array convention: every volume is a C-ordered ``numpy`` array indexed ``[z, y, x]``
(x fastest), the layout the reference hands to its JNA boundary
(``spim/process/fusion/deconvolution/MVDeconFFTThreads.java:157-165`` reverses
(x,y,z) -> (z,y,x)).  All per-axis rules in the reference are symmetric in the axes, so the
reversed index order changes nothing.

Reference path prefixes used in citations (all under /root/reference/src/main/java/):
  D2/ = mpicbg/spim/postprocessing/deconvolution2/   (gen-1: BayesMVDeconvolution, LRFFT, LRInput, Block)
  FD/ = spim/process/fusion/deconvolution/           (gen-2: MVDeconvolution, MVDeconFFT, MVDeconInput)
  PC/ = spim/process/cuda/                           (JNA interface, Block, BlockGeneratorFixedSizePrecise)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import scipy.fft as sfft

# --------------------------------------------------------------------------------------
# constants / enums
# --------------------------------------------------------------------------------------

#: ``LRInput.minValue`` D2/LRInput.java:30, ``MVDeconvolution.minValue`` FD/MVDeconvolution.java:70
MIN_VALUE = np.float32(0.0001)

#: ``enum PSFTYPE`` ordinal order, D2/LRFFT.java:54, FD/MVDeconFFT.java:49
OPTIMIZATION_II = 0
OPTIMIZATION_I = 1
EFFICIENT_BAYESIAN = 2
INDEPENDENT = 3
PSFTYPE_NAMES = ("OPTIMIZATION_II", "OPTIMIZATION_I", "EFFICIENT_BAYESIAN", "INDEPENDENT")

#: out-of-bounds rules (values shared with include/spim_mvdecon.h ``MVD_EXT_*``)
EXT_ZERO = 0            # Views.extendZero / OutOfBoundsStrategyValueFactory (kernel-building convs)
EXT_CONSTANT = 1        # Views.extendValue(image, c)  (gen-2 conv2 uses c = 1)
EXT_MIRROR_SINGLE = 2   # Views.extendMirrorSingle: edge voxel not repeated (gen-2 conv1)
EXT_MIRROR_DOUBLE = 3   # edge voxel repeated (offered because ImgLib1's default is not visible in tree)
EXT_PERIODIC = 4        # circular: what the JNA entry convolution3DfftCUDAInPlace computes on a block

GEN1 = 1  # D2/BayesMVDeconvolution + D2/LRFFT semantics
GEN2 = 2  # FD/MVDeconvolution + FD/MVDeconFFT semantics


# --------------------------------------------------------------------------------------
# out-of-bounds extension
# --------------------------------------------------------------------------------------

def ext_index(a: np.ndarray, n: int, mode: int) -> np.ndarray:
    """Map (possibly out-of-range) coordinates ``a`` to source indices in [0, n) for the
    index-mapping extension rules; returns -1 where the rule is a constant (ZERO/CONSTANT).
    Multiple reflections are handled (halo wider than the image)."""
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
    raise ValueError(f"unknown extension mode {mode}")


def extend(img: np.ndarray, lo: Sequence[int], hi: Sequence[int], mode: int, value: float = 0.0) -> np.ndarray:
    """Return ``img`` extended by ``lo[d]`` voxels before and ``hi[d]`` after each axis d."""
    out = img
    for d in range(img.ndim):
        n = img.shape[d]
        coords = np.arange(-lo[d], n + hi[d])
        idx = ext_index(coords, n, mode)
        taken = np.take(out, np.where(idx < 0, 0, idx), axis=d)
        if mode in (EXT_ZERO, EXT_CONSTANT):
            c = 0.0 if mode == EXT_ZERO else value
            shape = [1] * img.ndim
            shape[d] = len(idx)
            mask = (idx < 0).reshape(shape)
            taken = np.where(mask, np.asarray(c, dtype=img.dtype), taken)
        out = taken
    return np.ascontiguousarray(out)


# --------------------------------------------------------------------------------------
# convolution (out[p] = sum_q ext(A)[p - (q - c)] K[q], c = dim(K)/2)
# --------------------------------------------------------------------------------------

def _halos(kshape: Sequence[int]) -> Tuple[List[int], List[int]]:
    """left / right halo per axis for kernel origin at k//2: lo = k-1-k//2, hi = k//2."""
    lo = [k - 1 - k // 2 for k in kshape]
    hi = [k // 2 for k in kshape]
    return lo, hi


WORKERS = -1        # pocketfft threads when a call does not say (-1 = os.cpu_count()); bench.py pins it to the usable cores


def convolve(img: np.ndarray, kernel: np.ndarray, ext: int, value: float = 0.0,
             dtype=np.float32, workers: Optional[int] = None) -> np.ndarray:
    """Linear convolution of ``img`` (extended by rule ``ext``) with ``kernel`` whose origin is
    element ``dim//2``; output has ``img``'s shape.  True convolution, not correlation
    (``setComputeComplexConjugate(false)``, FD/MVDeconFFT.java:395,416,488,509).

    Restates ImgLib2 ``FFTConvolution`` / ImgLib1 ``FourierConvolution`` (third party, not in
    tree; see module docstring): pad by the kernel, real-to-complex FFT in ``dtype`` precision,
    multiply, inverse, crop.  The result is independent of the padded FFT size."""
    dtype = np.dtype(dtype)
    if workers is None:
        workers = WORKERS
    img = np.asarray(img, dtype=dtype)
    kernel = np.asarray(kernel, dtype=dtype)
    lo, hi = _halos(kernel.shape)
    big = extend(img, lo, hi, ext, value)                      # shape n + k - 1
    fshape = [sfft.next_fast_len(s, real=True) for s in big.shape]
    fi = sfft.rfftn(big, s=fshape, workers=workers)
    fk = sfft.rfftn(kernel, s=fshape, workers=workers)
    fi *= fk
    full = sfft.irfftn(fi, s=fshape, workers=workers)
    # 'valid' region of the padded image: starts at k-1 in the full linear convolution
    sl = tuple(slice(k - 1, k - 1 + n) for k, n in zip(kernel.shape, img.shape))
    return np.ascontiguousarray(full[sl]).astype(dtype, copy=False)


def convolve_direct(img: np.ndarray, kernel: np.ndarray, ext: int, value: float = 0.0) -> np.ndarray:
    """Direct-sum evaluation of the same definition in float64 (tiny volumes only)."""
    img = np.asarray(img, dtype=np.float64)
    kernel = np.asarray(kernel, dtype=np.float64)
    lo, hi = _halos(kernel.shape)
    big = extend(img, lo, hi, ext, value)
    out = np.zeros(img.shape, dtype=np.float64)
    n = img.shape
    for q in np.ndindex(*kernel.shape):
        # out[p] += big[p + lo - (q - c)] * K[q]; with lo = k-1-c this is big[p + (k-1-q)]
        sl = tuple(slice(k - 1 - qi, k - 1 - qi + ni) for k, qi, ni in zip(kernel.shape, q, n))
        out += big[sl] * kernel[q]
    return out


def circular_convolve(im: np.ndarray, kernel: np.ndarray, dtype=np.float32) -> np.ndarray:
    """What ``convolution3DfftCUDAInPlace`` (PC/CUDAFourierConvolution.java:31) returns:
    circular convolution over exactly ``im.shape`` with the kernel zero-padded to that shape
    and shifted so element ``kernelDim/2`` sits at the origin."""
    dtype = np.dtype(dtype)
    im = np.asarray(im, dtype=dtype)
    kp = np.zeros(im.shape, dtype=dtype)
    kp[tuple(slice(0, k) for k in kernel.shape)] = kernel
    kp = np.roll(kp, [-(k // 2) for k in kernel.shape], axis=tuple(range(im.ndim)))
    out = sfft.irfftn(sfft.rfftn(im) * sfft.rfftn(kp), s=im.shape)
    return out.astype(dtype, copy=False)


# --------------------------------------------------------------------------------------
# kernel helpers (LRFFT.init / MVDeconFFT.init)
# --------------------------------------------------------------------------------------

def sum_image(img: np.ndarray) -> float:
    """``AdjustInput.sumImage`` D2/AdjustInput.java:65-73 (RealSum = compensated fp64 sum)."""
    return math.fsum(np.asarray(img, dtype=np.float64).ravel().tolist()) if img.size <= (1 << 20) \
        else float(np.sum(img, dtype=np.float64))


def norm_image(img: np.ndarray) -> np.ndarray:
    """``AdjustInput.normImage`` D2/AdjustInput.java:53-59 / ``normImg`` FD/AdjustInput.java:50-56:
    each element becomes ``(float)((double)t / sum)``.  (The gen-2 ``sumImg`` double-count bug,
    FD/AdjustInput.java:114-118, is thread-count dependent and deliberately not reproduced;
    see ``sum_image_gen2_bug`` for a forensic emulation.)"""
    s = sum_image(img)
    return (np.asarray(img, dtype=np.float64) / s).astype(np.float32)


def sum_image_gen2_bug(img: np.ndarray, n_threads: int) -> float:
    """Forensic emulation of FD/AdjustInput.java:62-121 assuming tasks start in submission order:
    portions = 2*nThreads chunks (FusionHelper.divideIntoPortions), ``sums[0]`` added twice."""
    flat = np.asarray(img, dtype=np.float64).ravel()
    n_portions = 2 * n_threads
    size = flat.size
    chunk = size // n_portions
    mod = size % n_portions
    sums = []
    start = 0
    for i in range(n_portions):
        ln = chunk + mod if i == n_portions - 1 else chunk
        sums.append(math.fsum(flat[start:start + ln].tolist()))
        start += ln
    return math.fsum([sums[0]] + sums)


def mirror_quirk(img: np.ndarray) -> np.ndarray:
    """``Mirror.mirror`` applied along every axis, FD/Mirror.java:52-129 as called by
    ``computeInvertedKernel`` FD/MVDeconFFT.java:336-344 (gen-1: ImgLib1 MirrorImage,
    D2/LRFFT.java:373-381).  For odd sizes this is an exact flip.  For even sizes the reference
    swaps positions ``<= size/2`` (FD/Mirror.java:93), which swaps the middle pair twice, leaving it
    unflipped; restated literally."""
    out = np.array(img, copy=True)
    for d in range(out.ndim):
        n = out.shape[d]
        if n % 2 == 1:
            out = np.flip(out, axis=d)
        else:
            idx = np.arange(n)
            # sequential swaps for pos = 0..n/2 with partner n-1-pos
            for pos in range(0, n // 2 + 1):
                a, b = pos, n - 1 - pos
                idx[a], idx[b] = idx[b], idx[a]
            out = np.take(out, idx, axis=d)
    return np.ascontiguousarray(out)


def invert_kernel(k: np.ndarray) -> np.ndarray:
    return mirror_quirk(k)


def exponential_kernel(k: np.ndarray, num_views: int) -> np.ndarray:
    """``computeExponentialKernel`` + ``pow`` D2/LRFFT.java:361-391, FD/MVDeconFFT.java:326-354:
    repeated fp32 multiplication."""
    k = np.asarray(k, dtype=np.float32)
    r = k.copy()
    for _ in range(1, num_views):
        r = (r * k).astype(np.float32)
    return r


def _kconv(a: np.ndarray, b: np.ndarray, dtype) -> np.ndarray:
    """PSF-sized convolution with zero extension used while building kernel2
    (D2/LRFFT.java:249-260,293-297; FD/MVDeconFFT.java:222-248,285-295)."""
    return convolve(a, b, EXT_ZERO, dtype=dtype)


def init_kernels(psfs: Sequence[np.ndarray], iteration_type: int, dtype=np.float32
                 ) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """``LRInput.init`` -> ``LRFFT.init`` per view in list order (D2/LRInput.java:47-53,
    D2/LRFFT.java:214-325; FD/MVDeconInput.java:57-63, FD/MVDeconFFT.java:183-323).

    Faithful detail: view v normalises its own kernel1 at the start of its ``init``; when it
    then reads another view w's kernel1, that kernel is already normalised if w < v and still
    raw if w > v.  (The compound kernel is renormalised at the end, so the effect is a common
    scale removed again up to fp32 rounding.)"""
    dtype = np.dtype(dtype)
    k1 = [np.asarray(p, dtype=np.float32).copy() for p in psfs]
    num_views = len(k1)
    k2: List[Optional[np.ndarray]] = [None] * num_views
    for v in range(num_views):
        k1[v] = norm_image(k1[v])
        if num_views == 1 or iteration_type == INDEPENDENT:
            k2[v] = invert_kernel(k1[v])
        elif iteration_type == EFFICIENT_BAYESIAN:
            tmp = invert_kernel(k1[v]).astype(np.float32)
            for w in range(num_views):
                if w == v:
                    continue
                c1 = _kconv(invert_kernel(k1[v]), k1[w], dtype)
                c2 = _kconv(c1, invert_kernel(k1[w]), dtype)
                tmp = (c2.astype(np.float32) * tmp).astype(np.float32)
            k2[v] = norm_image(tmp)
        elif iteration_type == OPTIMIZATION_I:
            tmp = k1[v].copy()
            for w in range(num_views):
                if w == v:
                    continue
                c = _kconv(k1[v], invert_kernel(k1[w]), dtype)
                tmp = (c.astype(np.float32) * tmp).astype(np.float32)
            tmp = norm_image(tmp)
            k2[v] = invert_kernel(tmp)
        elif iteration_type == OPTIMIZATION_II:
            e = exponential_kernel(k1[v], num_views)
            e = norm_image(e)
            k2[v] = invert_kernel(e)
        else:
            raise ValueError(f"unknown iteration type {iteration_type}")
    return k1, [np.asarray(k, dtype=np.float32) for k in k2]


# --------------------------------------------------------------------------------------
# psi initialisation
# --------------------------------------------------------------------------------------

def norm_all_images_gen1(imgs: Sequence[np.ndarray], weights: Sequence[np.ndarray]) -> Tuple[float, int, float]:
    """``AdjustInput.normAllImages`` D2/AdjustInput.java:75-267 -> (avg, minOverlap, avgOverlap)."""
    nz = [np.asarray(w) != 0 for w in weights]
    count_local = np.zeros(imgs[0].shape, dtype=np.int64)
    sum_local = np.zeros(imgs[0].shape, dtype=np.float64)
    for im, m in zip(imgs, nz):
        count_local += m
        sum_local += np.where(m, np.asarray(im, dtype=np.float64), 0.0)
    two = count_local > 1
    any_ = count_local > 0
    count = int(count_local[two].sum())
    total = float(sum_local[two].sum(dtype=np.float64))
    min_overlap = int(count_local[any_].min()) if any_.any() else np.iinfo(np.int32).max
    n_any = int(any_.sum())
    avg_overlap = float(count_local[any_].sum()) / n_any if n_any else float("nan")
    if count == 0:
        return 1.0, min_overlap, avg_overlap
    return total / count, min_overlap, avg_overlap


def fuse_first_iteration_gen2(imgs: Sequence[np.ndarray]) -> Tuple[float, np.ndarray]:
    """``fuseFirstIteration`` FD/MVDeconvolution.java:213-256 + FD/FirstIteration.java:83-154:
    returns (avg, per-voxel count of views with img > 0)."""
    count = np.zeros(imgs[0].shape, dtype=np.int64)
    s = np.zeros(imgs[0].shape, dtype=np.float64)
    for im in imgs:
        m = np.asarray(im) > 0
        count += m
        s += np.where(m, np.asarray(im, dtype=np.float64), 0.0)
    has = count > 0
    n = int(has.sum())
    if n == 0:
        return float("nan"), count
    per_voxel = s[has] / count[has]
    return float(per_voxel.sum(dtype=np.float64)) / n, count


def adjust_osem(weights: Sequence[np.ndarray], osem: float) -> List[np.ndarray]:
    """``adjustOSEMspeedup`` D2/BayesMVDeconvolution.java:181-191 /
    ``adjustForOSEM`` FD/ProcessForDeconvolution.java:372-405: w <- min(1, w * (float)osem)."""
    if osem == 1.0:
        return [np.asarray(w, dtype=np.float32) for w in weights]
    f = np.float32(osem)
    return [np.minimum(np.float32(1), (np.asarray(w, dtype=np.float32) * f).astype(np.float32)) for w in weights]


# --------------------------------------------------------------------------------------
# per-voxel steps
# --------------------------------------------------------------------------------------

def compute_quotient(img: np.ndarray, blurred: np.ndarray, gen: int) -> np.ndarray:
    """gen-1 D2/BayesMVDeconvolution.java:410-432: q = img / blur.
    gen-2 FD/MVDeconvolution.java:494-546: q = img > 0 ? img / blur : 1."""
    with np.errstate(divide="ignore", invalid="ignore"):
        q = (img / blurred).astype(img.dtype)
    if gen == GEN2:
        q = np.where(img > 0, q, np.asarray(1, dtype=img.dtype))
    return q


def tikhonov(value: np.ndarray, lam: float) -> np.ndarray:
    """FD/MVDeconvolution.java:726: (sqrt(1 + 2*lambda*v) - 1) / lambda in fp64."""
    v = np.asarray(value, dtype=np.float64)
    return (np.sqrt(1.0 + 2.0 * lam * v) - 1.0) / lam


def compute_final_values(psi: np.ndarray, integral: np.ndarray, weight, lam: float,
                         min_value=MIN_VALUE) -> Tuple[np.ndarray, float, float]:
    """``computeFinalValues``/``computeNextValue`` D2/BayesMVDeconvolution.java:434-486,
    FD/MVDeconvolution.java:603-724.  Returns (new psi, sum |change|, max |change|).
    Works in psi's dtype (fp32 mirrors Java's float arithmetic; fp64 is the 'truth' mode)."""
    dt = psi.dtype
    mv = np.asarray(min_value, dtype=dt)
    value = (psi * integral).astype(dt)
    with np.errstate(invalid="ignore"):
        if lam > 0:
            adj = np.where(value > 0, tikhonov(np.where(value > 0, value, 0), lam).astype(dt), mv)
        else:
            adj = np.where(value > 0, value, mv)
    nxt = np.where(np.isnan(adj), mv, np.maximum(mv, adj)).astype(dt)
    w = np.asarray(weight, dtype=dt)
    new = (psi + ((nxt - psi).astype(dt) * w).astype(dt)).astype(dt)
    change = np.abs((new - psi).astype(dt))
    return new, float(change.sum(dtype=np.float64)), float(change.max()) if change.size else -1.0


# --------------------------------------------------------------------------------------
# block decomposition (a10)
# --------------------------------------------------------------------------------------

@dataclass
class Block:
    """One block of PC/Block.java:34-131 / D2/Block.java (fields in [z,y,x] order here)."""
    block_size: Tuple[int, ...]
    offset: Tuple[int, ...]             # global coordinate of block voxel 0 (may be negative)
    effective_size: Tuple[int, ...]
    effective_offset: Tuple[int, ...]   # global coordinate of the kept region
    effective_local_offset: Tuple[int, ...]

    def copy_block(self, source: np.ndarray, ext: int, value: float = 0.0) -> np.ndarray:
        """``copyBlock`` PC/Block.java:134-215: read through the out-of-bounds extension."""
        out = source
        for d in range(source.ndim):
            coords = np.arange(self.offset[d], self.offset[d] + self.block_size[d])
            idx = ext_index(coords, source.shape[d], ext)
            taken = np.take(out, np.where(idx < 0, 0, idx), axis=d)
            if ext in (EXT_ZERO, EXT_CONSTANT):
                c = 0.0 if ext == EXT_ZERO else value
                shape = [1] * source.ndim
                shape[d] = len(idx)
                taken = np.where((idx < 0).reshape(shape), np.asarray(c, dtype=source.dtype), taken)
            out = taken
        return np.ascontiguousarray(out)

    def paste_block(self, target: np.ndarray, block: np.ndarray) -> None:
        """``pasteBlock`` PC/Block.java:217-251: only the effective region is written back."""
        src = tuple(slice(lo, lo + s) for lo, s in zip(self.effective_local_offset, self.effective_size))
        dst = tuple(slice(o, o + s) for o, s in zip(self.effective_offset, self.effective_size))
        target[dst] = block[src]


def divide_into_blocks(img_size: Sequence[int], block_size: Sequence[int], kernel_size: Sequence[int],
                       gen: int = GEN2) -> Optional[List[Block]]:
    """``BlockGeneratorFixedSizePrecise.divideIntoBlocks`` PC/BlockGeneratorFixedSizePrecise.java:46-122
    (gen-2: returns None when a block is smaller than the kernel) and ``Block.divideIntoBlocks``
    D2/Block.java:376-458 (gen-1: doubles a too-small block size and retries).
    Block order: first listed axis fastest is irrelevant to results; we iterate x fastest like the
    reference's LocalizingZeroMinIntervalIterator over (x,y,z)."""
    nd = len(img_size)
    block_size = list(block_size)
    while True:
        eff = [block_size[d] - kernel_size[d] + 1 for d in range(nd)]
        if all(e > 0 for e in eff):
            break
        if gen == GEN2:
            return None
        for d in range(nd):
            if eff[d] <= 0:
                block_size[d] *= 2
    local = [kernel_size[d] // 2 for d in range(nd)]
    nblocks = [img_size[d] // eff[d] + (1 if img_size[d] % eff[d] else 0) for d in range(nd)]
    blocks: List[Block] = []
    # arrays are [z,y,x]; reference iterates dim 0 (= x = our last axis) fastest
    for cur in np.ndindex(*nblocks):
        eff_off = [cur[d] * eff[d] for d in range(nd)]
        off = [eff_off[d] - kernel_size[d] // 2 for d in range(nd)]
        eff_sz = [min(eff[d], img_size[d] - eff_off[d]) for d in range(nd)]
        blocks.append(Block(tuple(block_size), tuple(off), tuple(eff_sz), tuple(eff_off), tuple(local)))
    return blocks


def convolve_blocked(img: np.ndarray, kernel: np.ndarray, block_size: Sequence[int], ext: int,
                     value: float = 0.0, dtype=np.float32, gen: int = GEN2) -> np.ndarray:
    """Block-wise evaluation exactly as the CUDA/JNA path does it (FD/MVDeconFFTThreads.java:73-114):
    copyBlock through the extension, *circular* convolution on the block, paste the effective region."""
    blocks = divide_into_blocks(img.shape, block_size, kernel.shape, gen)
    if blocks is None:
        raise ValueError("block size smaller than kernel")
    out = np.empty_like(img)
    for b in blocks:
        blk = b.copy_block(img, ext, value)
        blk = circular_convolve(blk, kernel, dtype=dtype)
        b.paste_block(out, blk)
    return out


# --------------------------------------------------------------------------------------
# the iteration
# --------------------------------------------------------------------------------------

@dataclass
class DeconParams:
    iteration_type: int = EFFICIENT_BAYESIAN
    num_iterations: int = 10
    lam: float = 0.006
    gen: int = GEN2
    osem_speedup: float = 1.0
    osem_index: int = 0            # gen-1 only: 0 given, 1 min overlap, 2 avg overlap, 3 manual
    conv1_ext: Optional[int] = None   # default mirror-single
    conv2_ext: Optional[int] = None   # default gen-2 constant 1.0, gen-1 mirror-single
    dtype: type = np.float32       # np.float64 = 'truth' mode
    mask_at_end: Optional[bool] = None  # default: gen-2 True (FD/MVDeconvolution.java:201-208)
    psi_init: Optional[np.ndarray] = None  # 'initialImage' hook (D2/...:88-89, FD/...:116-125)


@dataclass
class DeconResult:
    psi: np.ndarray
    avg: float
    kernel1: List[np.ndarray]
    kernel2: List[np.ndarray]
    stats: List[Tuple[int, int, float, float]] = field(default_factory=list)  # (iter, view, sum, max)
    osem: float = 1.0


def view_step(psi, img, weight, k1, k2, p: DeconParams):
    """One view of ``runIteration`` (D2/BayesMVDeconvolution.java:278-342,
    FD/MVDeconvolution.java:374-462): conv1 -> quotient -> conv2 -> update."""
    e1 = p.conv1_ext if p.conv1_ext is not None else EXT_MIRROR_SINGLE
    e2 = p.conv2_ext if p.conv2_ext is not None else (EXT_CONSTANT if p.gen == GEN2 else EXT_MIRROR_SINGLE)
    blurred = convolve(psi, k1, e1, dtype=p.dtype)
    q = compute_quotient(img, blurred, p.gen)
    integral = convolve(q, k2, e2, value=1.0, dtype=p.dtype)
    return compute_final_values(psi, integral, weight, p.lam)


def deconvolve(imgs: Sequence[np.ndarray], weights: Sequence[np.ndarray], psfs: Sequence[np.ndarray],
               p: DeconParams) -> DeconResult:
    """``new BayesMVDeconvolution(...)`` (D2/BayesMVDeconvolution.java:79-178) or
    ``new MVDeconvolution(...)`` (FD/MVDeconvolution.java:94-211): both run every iteration inside
    the constructor."""
    dt = np.dtype(p.dtype)
    imgs = [np.asarray(i, dtype=dt) for i in imgs]
    weights = [np.asarray(w, dtype=np.float32) for w in weights]
    osem = p.osem_speedup
    if p.gen == GEN1:
        avg, min_ov, avg_ov = norm_all_images_gen1(imgs, weights)
        if p.osem_index == 1:
            osem = max(1.0, float(min_ov))
        elif p.osem_index == 2:
            osem = max(1.0, avg_ov)
        weights = adjust_osem(weights, osem)
        k1, k2 = init_kernels(psfs, p.iteration_type, dtype=np.float32)
        avg = float(np.float32(avg))                       # this.avg = (float)result[0]
    else:
        # gen-2: OSEM clamp happens upstream (FD/ProcessForDeconvolution.java:372-405); the
        # constructor's osem arguments are unused.  Apply it here so both generations accept
        # the same inputs.
        weights = adjust_osem(weights, osem)
        k1, k2 = init_kernels(psfs, p.iteration_type, dtype=np.float32)
        avg, _ = fuse_first_iteration_gen2(imgs)
        if math.isnan(avg):
            avg = 0.5
    weights = [w.astype(dt) for w in weights]
    if p.psi_init is not None:
        psi = np.asarray(p.psi_init, dtype=dt).copy()
    else:
        psi = np.full(imgs[0].shape, np.float32(avg), dtype=dt)
    res = DeconResult(psi=psi, avg=avg, kernel1=k1, kernel2=k2, osem=osem)
    k1c = [k.astype(dt) for k in k1]
    k2c = [k.astype(dt) for k in k2]
    for it in range(p.num_iterations):
        for v in range(len(imgs)):
            psi, s, m = view_step(psi, imgs[v], weights[v], k1c[v], k2c[v], p)
            res.stats.append((it, v, s, m))
    mask = p.mask_at_end if p.mask_at_end is not None else (p.gen == GEN2)
    if mask:
        _, count = fuse_first_iteration_gen2(imgs)
        psi = np.where(count == 0, np.asarray(0, dtype=dt), psi)
    res.psi = psi
    return res


# --------------------------------------------------------------------------------------
# parity metric
# --------------------------------------------------------------------------------------

def parity_errors(a: np.ndarray, b: np.ndarray, floor: float = float(MIN_VALUE)) -> Tuple[float, float]:
    """(max per-voxel |a-b| / max(|b|, floor), relative L2 ||a-b|| / ||b||) in fp64."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    per = np.abs(a - b) / np.maximum(np.abs(b), floor)
    l2 = float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))
    return float(per.max()) if per.size else 0.0, l2
