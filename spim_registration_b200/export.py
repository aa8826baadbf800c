"""Output side of the path: the deconvolved volume as a 32-bit TIFF stack, and the
TIFF reader the PSF loader needs.  Host-side IO, like the reference's (which delegates to ImageJ).

Mirrored (paths under /root/reference/src/main/java/):
  spim/process/fusion/export/Save3dTIFF.java:72-130         exportImage: title + ".tif", display range = min / max of the image,
                                                             calibration origin = -bb.min / downsampling, pixel size = downsampling
  spim/process/fusion/export/{DefaultImgTitler,FixedNameImgTitler}.java
  spim/process/fusion/FusionHelper.java:55-73                getIllumName / getAngleName
  spim/process/fusion/deconvolution/EfficientBayesianBased.java:297   "TP<t>_Ch<c>_Ill..._Ang..."
  fiji/plugin/Multi_View_Deconvolution.java:272-293          gen-1 name "DC(l=<lambda>)_t<tp>_ch<ch>"
  spim/process/fusion/deconvolution/ExtractPSF.java:541-598  loadAndTransformPSFs: PSFs from TIFF files

The TIFF layout follows what ImageJ's FileSaver.saveAsTiffStack writes (ImageJ is an un-vendored dependency, so this is
"readable as the same stack by ImageJ", not byte parity): big-endian, one IFD per plane, planes stored back to back after
the first IFD, SampleFormat = IEEE float, and the ImageJ description block (images / slices / min / max / origin / spacing)
in the first IFD.  The reader handles uncompressed strips of 8 / 16 / 32-bit integer and 32-bit float samples in either
byte order, which covers what ImageJ's Opener hands to ExtractPSF.loadAndTransformPSFs.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

IMAGEJ_VERSION = "1.53t"


def java_double_to_string(v: float) -> str:
    """Java's Double.toString for the magnitudes a lambda takes: plain decimal in [1e-3, 1e7), else d.dddE<exp>."""
    v = float(v)
    if v == 0:
        return "0.0"
    a = abs(v)
    if 1e-3 <= a < 1e7:
        s = repr(a)
        if "e" in s or "E" in s:
            s = f"{a:.17f}".rstrip("0")
        if "." not in s:
            s += ".0"
        if s.endswith("."):
            s += "0"
    else:
        from decimal import Decimal
        sign, digits, exp = Decimal(repr(a)).as_tuple()      # shortest digits that round-trip, like Java's
        ds = "".join(str(d) for d in digits).rstrip("0") or "0"
        e10 = len(digits) - 1 + exp
        s = ds[0] + "." + (ds[1:] or "0") + "E" + str(e10)
    return ("-" if v < 0 else "") + s


def gen1_output_name(lam: float, timepoint: int, channel: int, tikhonov: bool = True) -> str:
    """fiji/plugin/Multi_View_Deconvolution.java:281, 291-293."""
    return "DC(l=" + (java_double_to_string(lam) if tikhonov else "0") + ")_t" + str(timepoint) + "_ch" + str(channel)


def illum_name(names: Sequence[str]) -> str:
    return "_Ill" + ",".join(str(n) for n in names)


def angle_name(names: Sequence[str]) -> str:
    return "_Ang" + ",".join(str(n) for n in names)


def gen2_output_title(timepoint: str, channel: str, illums: Sequence[str], angles: Sequence[str]) -> str:
    """EfficientBayesianBased.java:297."""
    return "TP" + str(timepoint) + "_Ch" + str(channel) + illum_name(illums) + angle_name(angles)


class DefaultImgTitler:
    def getImageTitle(self, tp, vs) -> str:
        """tp: timepoint id; vs: dict with 'channel', 'illumination', 'angle' names (DefaultImgTitler.java:31)."""
        return f"Timepoint{tp}_Channel{vs['channel']}_Illum{vs['illumination']}_Angle{vs['angle']}"


class FixedNameImgTitler:
    def __init__(self, title: str):
        self.title = title

    def setTitle(self, title: str) -> None:
        self.title = title

    def getImageTitle(self, tp=None, vs=None) -> str:
        return self.title


# --------------------------------------------------------------------------------------------------------------
# TIFF writer
# --------------------------------------------------------------------------------------------------------------
_TYPES = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II"}


def _ifd_entry(tag: int, typ: int, count: int, value: int) -> bytes:
    if typ == 3 and count == 1:
        return struct.pack(">HHIHH", tag, typ, count, value, 0)
    return struct.pack(">HHII", tag, typ, count, value)


def write_tiff_stack(path: str, vol: np.ndarray, display_range: Optional[Tuple[float, float]] = None,
                     origin_xyz: Optional[Sequence[float]] = None, spacing: Optional[float] = None) -> None:
    """Write a [z, y, x] float32 volume as a multi-page big-endian TIFF in ImageJ's stack layout."""
    a = np.ascontiguousarray(vol, dtype=np.float32)
    if a.ndim == 2:
        a = a[None]
    if a.ndim != 3:
        raise ValueError("volume must be [z, y, x]")
    nz, ny, nx = a.shape
    plane = nx * ny * 4
    if 8 + nz * plane + nz * 256 >= 2 ** 32:
        raise ValueError("volume too large for a classic TIFF (ImageJ switches to its raw > 4 GB layout there); export in parts")
    mn, mx = display_range if display_range is not None else (float(a.min()), float(a.max()))
    desc = f"ImageJ={IMAGEJ_VERSION}\n"
    if nz > 1:
        desc += f"images={nz}\nslices={nz}\n"
    if spacing is not None:
        desc += "unit=pixel\n" if spacing == 1 else "unit=px\n"
        if nz > 1:
            desc += f"spacing={java_double_to_string(spacing)}\n"
    if nz > 1:
        desc += "loop=false\n"
    desc += f"min={java_double_to_string(mn)}\nmax={java_double_to_string(mx)}\n"
    if origin_xyz is not None:
        for n_, v in zip(("xorigin", "yorigin", "zorigin"), origin_xyz):
            if v != 0:
                desc += f"{n_}={java_double_to_string(v)}\n"
    desc_b = desc.encode("latin-1") + b"\0"

    def ifd_bytes(first: bool, data_off: int, next_off: int, desc_off: int) -> bytes:
        e = [_ifd_entry(254, 4, 1, 0), _ifd_entry(256, 4, 1, nx), _ifd_entry(257, 4, 1, ny), _ifd_entry(258, 3, 1, 32),
             _ifd_entry(262, 3, 1, 1)]
        if first:
            e.append(_ifd_entry(270, 2, len(desc_b), desc_off))
        e += [_ifd_entry(273, 4, 1, data_off), _ifd_entry(277, 3, 1, 1), _ifd_entry(278, 3, 1, ny), _ifd_entry(279, 4, 1, plane),
              _ifd_entry(339, 3, 1, 3)]
        return struct.pack(">H", len(e)) + b"".join(e) + struct.pack(">I", next_off)

    n_first = 2 + 11 * 12 + 4
    n_rest = 2 + 10 * 12 + 4
    desc_off = 8 + n_first
    data_off = desc_off + len(desc_b)
    data_off += data_off & 1
    rest_off = data_off + nz * plane
    with open(path, "wb") as f:
        f.write(b"MM\0*" + struct.pack(">I", 8))
        f.write(ifd_bytes(True, data_off, rest_off if nz > 1 else 0, desc_off))
        f.write(desc_b)
        f.write(b"\0" * (data_off - (desc_off + len(desc_b))))
        f.write(a.astype(">f4").tobytes())
        for z in range(1, nz):
            nxt = rest_off + z * n_rest if z + 1 < nz else 0
            f.write(ifd_bytes(False, data_off + z * plane, nxt, 0))


# --------------------------------------------------------------------------------------------------------------
# TIFF reader (uncompressed strips; 8 / 16 / 32-bit unsigned or signed integer, 32-bit float)
# --------------------------------------------------------------------------------------------------------------
def read_tiff_stack(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:2] == b"II":
        bo = "<"
    elif raw[:2] == b"MM":
        bo = ">"
    else:
        raise ValueError(f"{path}: not a TIFF file")
    if struct.unpack(bo + "H", raw[2:4])[0] != 42:
        raise ValueError(f"{path}: not a classic TIFF (BigTIFF is not supported)")
    off = struct.unpack(bo + "I", raw[4:8])[0]
    size = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 8: 2, 9: 4, 11: 4, 12: 8, 16: 8}
    planes: List[np.ndarray] = []
    first_desc = ""
    seen = set()
    while off and off not in seen:
        seen.add(off)
        n = struct.unpack(bo + "H", raw[off:off + 2])[0]
        tags: Dict[int, list] = {}
        for i in range(n):
            tag, typ, cnt = struct.unpack(bo + "HHI", raw[off + 2 + 12 * i: off + 10 + 12 * i])
            field = raw[off + 10 + 12 * i: off + 14 + 12 * i]
            nbytes = size.get(typ, 1) * cnt
            data = field[:nbytes] if nbytes <= 4 else raw[struct.unpack(bo + "I", field)[0]:][:nbytes]
            if typ == 2:
                tags[tag] = [data.split(b"\0")[0].decode("latin-1")]
            elif typ in (3, 4, 1):
                tags[tag] = list(struct.unpack(bo + {1: "B", 3: "H", 4: "I"}[typ] * cnt, data))
            else:
                tags[tag] = [data]
        nxt = struct.unpack(bo + "I", raw[off + 2 + 12 * n: off + 6 + 12 * n])[0]
        w, h = tags[256][0], tags[257][0]
        bits = tags.get(258, [1])[0]
        fmt = tags.get(339, [1])[0]
        if tags.get(259, [1])[0] != 1:
            raise ValueError(f"{path}: compressed TIFFs are not supported")
        if tags.get(277, [1])[0] != 1:
            raise ValueError(f"{path}: only single-channel images are supported")
        if fmt == 3 and bits == 32:
            dt = np.dtype(bo + "f4")
        elif bits in (8, 16, 32) and fmt in (1, 2):
            dt = np.dtype(bo + ("u" if fmt == 1 else "i") + str(bits // 8))
        else:
            raise ValueError(f"{path}: unsupported sample type (bits {bits}, format {fmt})")
        offs, cnts = tags[273], tags.get(279, [w * h * dt.itemsize])
        buf = b"".join(raw[o:o + c] for o, c in zip(offs, cnts))
        planes.append(np.frombuffer(buf, dtype=dt, count=w * h).reshape(h, w).astype(np.float32))
        if not planes[1:]:
            first_desc = tags.get(270, [""])[0]
        off = nxt
    # ImageJ stores big stacks with a single IFD and "images=N": the planes follow the first one back to back
    if len(planes) == 1 and "images=" in first_desc:
        try:
            nimg = int(first_desc.split("images=")[1].split("\n")[0])
        except ValueError:
            nimg = 1
        if nimg > 1:
            start = offs[0]
            h, w = planes[0].shape
            arr = np.frombuffer(raw, dtype=dt, count=w * h * nimg, offset=start).reshape(nimg, h, w)
            return arr.astype(np.float32)
    return np.stack(planes, axis=0)


# --------------------------------------------------------------------------------------------------------------
class Save3dTIFF:
    """spim/process/fusion/export/Save3dTIFF.java."""

    def __init__(self, path: str, compress: bool = False):
        if compress:
            raise NotImplementedError("zip-compressed export is not supported")
        self.path = path
        self.imgTitler = DefaultImgTitler()

    def setImgTitler(self, t) -> None:
        self.imgTitler = t

    def getImgTitler(self):
        return self.imgTitler

    def getDescription(self) -> str:
        return "Save as TIFF stack"

    def exportImage(self, img: Optional[np.ndarray], bb_min: Optional[Sequence[int]] = None, downsampling: int = 1,
                    tp=None, vs=None, min: float = float("nan"), max: float = float("nan"), title: Optional[str] = None) -> bool:
        """Save3dTIFF.java:72-130.  ``title`` = exportImage( img, title ) of :61-69."""
        if img is None:
            return False
        a = np.ascontiguousarray(img, dtype=np.float32)
        if np.isnan(min) or np.isnan(max):
            rng = (float(a.min()), float(a.max()))               # FusionHelper.minMax
        else:
            rng = (float(np.float32(min)), float(np.float32(max)))
        name = title if title is not None else self.imgTitler.getImageTitle(tp, vs)
        fn = os.path.join(self.path, name if name.endswith(".tif") else name + ".tif")
        origin = spacing = None
        if bb_min is not None:
            ds = int(downsampling)
            jdiv = lambda v: abs(int(v)) // ds * (1 if v >= 0 else -1)      # Java integer division truncates towards zero
            origin = [-jdiv(bb_min[d]) for d in range(3)]
            spacing = float(downsampling)
        write_tiff_stack(fn, a, display_range=rng, origin_xyz=origin, spacing=spacing)
        return True

    def finish(self) -> bool:
        return False
