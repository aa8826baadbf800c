"""Deterministic synthetic multiview data: bead / specimen ground truth,
rotated anisotropic Gaussian PSFs, noisy min-max-normalised views and blending-style weights.

This is input generation, not the hot path: it uses NumPy (and SciPy's FFT for the forward
blur of the ground truth) on the host.  Array convention is ``[z, y, x]`` (x fastest), as
handed to the JNA boundary (FD/MVDeconFFTThreads.java:157-165).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

MIN_VALUE = np.float32(0.0001)


def make_psf(size: int, view: int, num_views: int, scale: float = 1.0) -> np.ndarray:
    """Anisotropic Gaussian (sigma_axial = 3*scale along z, sigma_lateral = 1.2*scale) rotated about
    the y axis by 360*view/num_views degrees, sampled on the odd grid centred at size//2, sum 1."""
    c = size // 2
    z, y, x = np.meshgrid(np.arange(size) - c, np.arange(size) - c, np.arange(size) - c, indexing="ij")
    ang = 2.0 * np.pi * view / max(1, num_views)
    ca, sa = np.cos(ang), np.sin(ang)
    xr = ca * x + sa * z
    zr = -sa * x + ca * z
    sl, sz = 1.2 * scale, 3.0 * scale
    g = np.exp(-0.5 * ((xr / sl) ** 2 + (y / sl) ** 2 + (zr / sz) ** 2))
    g /= g.sum()
    return g.astype(np.float32)


def make_psfs(num_views: int, size: int) -> List[np.ndarray]:
    scale = 1.0 if size <= 15 else 2.0
    return [make_psf(size, v, num_views, scale) for v in range(num_views)]


def bead_truth(shape: Sequence[int], n_beads: int = 200, seed: int = 20140613) -> np.ndarray:
    """Background 1.0 + delta beads at uniform-random integer positions >= 8 voxels from faces,
    amplitudes U(200, 1000)."""
    rng = np.random.default_rng(seed)
    t = np.ones(shape, dtype=np.float32)
    margin = [min(8, max(0, (s - 1) // 2)) for s in shape]
    for _ in range(n_beads):
        pos = tuple(int(rng.integers(m, s - m)) for m, s in zip(margin, shape))
        t[pos] += np.float32(rng.uniform(200.0, 1000.0))
    return t


def specimen_truth(shape: Sequence[int], n_blobs: int = 64, n_beads: int = 2000, seed: int = 2929) -> np.ndarray:
    """Sum of random 3-D Gaussian blobs (sigma U(6,40) voxels per axis scaled to the volume,
    amplitude U(50,500)) + beads + background 5.0.  Blobs are built separably so that the
    2 GiB volumes of the large configs stay cheap."""
    rng = np.random.default_rng(seed)
    nz, ny, nx = shape
    t = np.full(shape, 5.0, dtype=np.float32)
    s_scale = min(1.0, min(shape) / 256.0)
    for _ in range(n_blobs):
        c = [rng.uniform(0, s) for s in shape]
        sig = [max(1.5, rng.uniform(6.0, 40.0) * s_scale) for _ in range(3)]
        amp = rng.uniform(50.0, 500.0)
        gz = np.exp(-0.5 * ((np.arange(nz) - c[0]) / sig[0]) ** 2).astype(np.float32)
        gy = np.exp(-0.5 * ((np.arange(ny) - c[1]) / sig[1]) ** 2).astype(np.float32)
        gx = np.exp(-0.5 * ((np.arange(nx) - c[2]) / sig[2]) ** 2).astype(np.float32)
        # restrict to the 4-sigma box
        lo = [max(0, int(ci - 4 * si)) for ci, si in zip(c, sig)]
        hi = [min(s, int(ci + 4 * si) + 1) for ci, si, s in zip(c, sig, shape)]
        if any(h <= l for l, h in zip(lo, hi)):
            continue
        sub = (gz[lo[0]:hi[0], None, None] * gy[None, lo[1]:hi[1], None]) * gx[None, None, lo[2]:hi[2]]
        t[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] += np.float32(amp) * sub
    margin = [min(8, max(0, (s - 1) // 2)) for s in shape]
    n_beads = int(n_beads * min(1.0, np.prod(shape) / float(512 * 512 * 256)))
    for _ in range(n_beads):
        pos = tuple(int(rng.integers(m, s - m)) for m, s in zip(margin, shape))
        t[pos] += np.float32(rng.uniform(200.0, 1000.0))
    return t


def _mirror_blur_numpy(truth: np.ndarray, psf: np.ndarray) -> np.ndarray:
    import scipy.fft as sfft
    lo = [k - 1 - k // 2 for k in psf.shape]
    hi = [k // 2 for k in psf.shape]
    big = np.pad(truth, list(zip(lo, hi)), mode="reflect")
    fs = [sfft.next_fast_len(s, real=True) for s in big.shape]
    out = sfft.irfftn(sfft.rfftn(big, s=fs, workers=-1) * sfft.rfftn(psf, s=fs, workers=-1), s=fs, workers=-1)
    sl = tuple(slice(k - 1, k - 1 + n) for k, n in zip(psf.shape, truth.shape))
    return np.ascontiguousarray(out[sl]).astype(np.float32)


def view_footprint(shape: Sequence[int], view: int, num_views: int, cover: float = 0.8) -> Tuple[slice, ...]:
    """Per-view axis-aligned box covering ``cover`` of each axis, shifted per view so that the
    number of overlapping views varies 1..V across the volume."""
    sl = []
    for d, s in enumerate(shape):
        ln = max(1, int(round(cover * s)))
        free = s - ln
        phase = ((view * (d + 1)) % max(1, num_views)) / max(1, num_views - 1) if num_views > 1 else 0.5
        start = int(round(free * phase))
        sl.append(slice(start, start + ln))
    return tuple(sl)


def cosine_blend(shape: Sequence[int], box: Tuple[slice, ...], border: Sequence[float], rng_: Sequence[float]) -> np.ndarray:
    """Cosine blending weights from the distance to the box border (the shape of
    spim/process/fusion/weights/BlendingRealRandomAccess.java:91-121): 0 outside box+border,
    ramping 0->1 over ``rng_`` voxels via 0.5*(cos((1-d/range)*pi)+1), 1 inside."""
    w = np.ones(shape, dtype=np.float32)
    for d, s in enumerate(shape):
        c = np.arange(s, dtype=np.float64)
        lo = box[d].start + border[d]
        hi = box[d].stop - 1 - border[d]
        dist = np.minimum(c - lo, hi - c)
        f = np.where(dist < 0, 0.0,
                     np.where(dist < rng_[d], 0.5 * (np.cos((1.0 - dist / max(rng_[d], 1e-9)) * np.pi) + 1.0), 1.0))
        sh = [1, 1, 1]
        sh[d] = s
        w = w * f.reshape(sh).astype(np.float32)
    return w


def make_views(truth: np.ndarray, psfs: Sequence[np.ndarray], noise_sigma: float = 0.5, seed: int = 7,
               blur: Optional[Callable[[np.ndarray, np.ndarray], np.ndarray]] = None,
               weight_mode: str = "normalized", cover: float = 0.8
               ) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """raw_v = mirror-conv(truth, PSF_v) + N(0, sigma^2); min-max normalise to [0,1] as the
    reference's loaders do (spim/fiji/spimdata/imgloaders/AbstractImgLoader.java:164-184); clamp to
    >= 1e-4 inside the view's footprint and 0 outside (FD/TransformInput.java:108-115).

    weight_mode: 'normalized' -> cosine blending divided by the sum over views (sum_v w_v <= 1);
                 'blending'   -> cosine blending normalised only where sum > 1
                                 (FD/WeightNormalizer.java:243-246);
                 'ones'       -> 1 inside the footprint."""
    blur = blur or _mirror_blur_numpy
    rng = np.random.default_rng(seed)
    V = len(psfs)
    shape = truth.shape
    imgs, ws = [], []
    for v in range(V):
        raw = blur(truth, psfs[v]).astype(np.float32)
        raw = raw + rng.normal(0.0, noise_sigma, size=shape).astype(np.float32)
        mn, mx = float(raw.min()), float(raw.max())
        raw = ((raw - mn) / max(mx - mn, 1e-20)).astype(np.float32)
        box = view_footprint(shape, v, V, cover)
        img = np.zeros(shape, dtype=np.float32)
        img[box] = np.maximum(MIN_VALUE, raw[box])
        imgs.append(img)
        if weight_mode == "ones":
            w = np.zeros(shape, dtype=np.float32)
            w[box] = 1.0
        else:
            border = [0.0, 0.0, 0.0]
            rr = [min(12.0, s / 8.0) for s in shape]
            w = cosine_blend(shape, box, border, rr)
            w[img == 0] = 0.0
        ws.append(w.astype(np.float32))
    if weight_mode == "normalized":
        tot = np.sum(ws, axis=0, dtype=np.float32)
        ws = [np.where(tot > 0, w / np.maximum(tot, 1e-20), 0).astype(np.float32) for w in ws]
    elif weight_mode == "blending":
        tot = np.sum(ws, axis=0, dtype=np.float32)
        ws = [np.where(tot > 1, w / np.maximum(tot, 1e-20), w).astype(np.float32) for w in ws]
    return imgs, ws


def make_dataset(shape: Sequence[int], num_views: int, psf_size: int, kind: str = "specimen",
                 seed: int = 7, weight_mode: str = "normalized",
                 blur: Optional[Callable[[np.ndarray, np.ndarray], np.ndarray]] = None):
    """Convenience: (truth, imgs, weights, psfs) for one of the BASELINE.json configurations."""
    truth = bead_truth(shape) if kind == "beads" else specimen_truth(shape)
    psfs = make_psfs(num_views, psf_size)
    imgs, ws = make_views(truth, psfs, seed=seed, weight_mode=weight_mode, blur=blur)
    return truth, imgs, ws, psfs


# --------------------------------------------------------------------------------------------------
# position-deterministic noise volumes for the full-size multi-GPU runs
# --------------------------------------------------------------------------------------------------
# Every voxel's value is a hash of its GLOBAL linear index and a seed, so that (a) each rank generates its own brick on
# its own GPU (torch twin below) without any host memory or PCIe traffic, and (b) a test can regenerate any crop of the
# global volume on the CPU (numpy twin), bit for bit, and hand it to the oracle.
_M32 = 0xFFFFFFFF


def _hash_u24(idx, seed: int, xp):
    """idx: int64 array (numpy or torch) of global linear indices -> 24-bit integers, identical in both libraries
    (64-bit two's-complement products, masked to 32 bits after every step)."""
    h = (idx * 0x9E3779B1 + (int(seed) * 0x85EBCA6B & _M32)) & _M32
    h = h ^ (h >> 15)
    h = (h * 0x2C1B3C6D) & _M32
    h = h ^ (h >> 12)
    h = (h * 0x297A2D39) & _M32
    h = h ^ (h >> 15)
    return h >> 8


def hash_volume_numpy(gshape: Sequence[int], lo: Sequence[int], ext: Sequence[int], seed: int, a: float, b: float) -> np.ndarray:
    """a + b * u, u in [0, 1) hashed from the global index, for the box [lo, lo + ext) of a volume of shape gshape."""
    z = np.arange(lo[0], lo[0] + ext[0], dtype=np.int64)[:, None, None]
    y = np.arange(lo[1], lo[1] + ext[1], dtype=np.int64)[None, :, None]
    x = np.arange(lo[2], lo[2] + ext[2], dtype=np.int64)[None, None, :]
    idx = (z * int(gshape[1]) + y) * int(gshape[2]) + x
    u = _hash_u24(idx, seed, np).astype(np.float32) * np.float32(2.0 ** -24)
    return (np.float32(a) + np.float32(b) * u).astype(np.float32)


def hash_volume_torch(gshape: Sequence[int], lo: Sequence[int], ext: Sequence[int], seed: int, a: float, b: float, device):
    """The same values as hash_volume_numpy, computed on `device` (a float32 torch tensor of shape ext)."""
    import torch
    z = torch.arange(lo[0], lo[0] + ext[0], dtype=torch.int64, device=device)[:, None, None]
    y = torch.arange(lo[1], lo[1] + ext[1], dtype=torch.int64, device=device)[None, :, None]
    x = torch.arange(lo[2], lo[2] + ext[2], dtype=torch.int64, device=device)[None, None, :]
    idx = (z * int(gshape[1]) + y) * int(gshape[2]) + x
    u = _hash_u24(idx, seed, torch).to(torch.float32) * (2.0 ** -24)
    return (u * b + a).contiguous()      # one rounding for the product, one for the sum, like the numpy twin


def hash_view(gshape, lo, ext, view: int, num_views: int, xp="numpy", device=None):
    """(image, weight) of one view of the hash dataset: image in [0.05, 1), weight in [0.25, 1) / num_views."""
    f = hash_volume_numpy if xp == "numpy" else (lambda *a_: hash_volume_torch(*a_, device))
    img = f(gshape, lo, ext, 1000 + view, 0.05, 0.95)
    w = f(gshape, lo, ext, 5000 + view, 0.25 / num_views, 0.75 / num_views)
    return img, w
