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


# =================================================================================================================
# SpimData2 XML projects around the exported stacks (spim/process/fusion/export/ExportSpimData2TIFF.java,
# AppendSpimData2.java, XMLTIFFImgTitler.java; loader tags from spim/fiji/spimdata/imgloaders/XmlIoStackImgLoader.java:44-66).
# The XML writer of the reference is the un-vendored spimdata library (mpicbg.spim.data.XmlIoSpimData): the element
# layout below is its published "SpimData version 0.2" schema -- readable by BigDataViewer / Multiview-Reconstruction,
# not a byte copy of what JDOM would emit.  The HDF5 variants (ExportSpimData2HDF5 / AppendSpimData2HDF5) need an HDF5
# library, which this image does not have: they raise NotImplementedError with that message.
# =================================================================================================================
import xml.etree.ElementTree as _ET
from dataclasses import dataclass, field


@dataclass(frozen=True)
class Entity:
    """Angle / Channel / Illumination of the spimdata model: an id and a display name (equality by id, like
    mpicbg.spim.data.generic.base.Entity)."""
    id: int
    name: str = field(default="", compare=False)

    def getName(self) -> str:
        return self.name if self.name else str(self.id)


@dataclass(frozen=True)
class TimePoint:
    id: int

    def getId(self) -> int:
        return self.id

    def getName(self) -> str:
        return str(self.id)


@dataclass(frozen=True)
class ViewSetup:
    id: int
    angle: Entity = Entity(0)
    channel: Entity = Entity(0)
    illumination: Entity = Entity(0)
    name: str = ""
    size_xyz: Optional[Tuple[int, int, int]] = None
    voxel_size_xyz: Tuple[float, float, float] = (1.0, 1.0, 1.0)
    voxel_unit: str = "um"

    def getId(self) -> int:
        return self.id

    def getAngle(self) -> Entity:
        return self.angle

    def getChannel(self) -> Entity:
        return self.channel

    def getIllumination(self) -> Entity:
        return self.illumination


class XMLTIFFImgTitler:
    """XMLTIFFImgTitler.java:46-63: "img" + _TL<t> / _Ch<c> / _Ill<i> / _Angle<a>, each only when more than one exists."""

    def __init__(self, newTimepoints: Sequence[TimePoint], newViewSetups: Sequence[ViewSetup]):
        self.timepoints, self.viewSetups = list(newTimepoints), list(newViewSetups)

    def getImageTitle(self, tp: TimePoint, vs: ViewSetup) -> str:
        fn = "img"
        if len(self.timepoints) > 1:
            fn += "_TL" + str(tp.getId())
        if len({v.getChannel() for v in self.viewSetups}) > 1:
            fn += "_Ch" + vs.getChannel().getName()
        if len({v.getIllumination() for v in self.viewSetups}) > 1:
            fn += "_Ill" + vs.getIllumination().getName()
        if len({v.getAngle() for v in self.viewSetups}) > 1:
            fn += "_Angle" + vs.getAngle().getName()
        return fn


@dataclass
class FileNamePattern:
    layoutTP: int = 0
    layoutChannels: int = 0
    layoutIllum: int = 0
    layoutAngles: int = 0
    fileNamePattern: str = "img"


def getFileNamePattern(timepoints: Sequence[TimePoint], viewSetups: Sequence[ViewSetup], compress: bool = False) -> FileNamePattern:
    """ExportSpimData2TIFF.java:186-224."""
    f = FileNamePattern()
    if len(timepoints) > 1:
        f.fileNamePattern += "_TL{t}"; f.layoutTP = 1
    if len({v.getChannel() for v in viewSetups}) > 1:
        f.fileNamePattern += "_Ch{c}"; f.layoutChannels = 1
    if len({v.getIllumination() for v in viewSetups}) > 1:
        f.fileNamePattern += "_Ill{i}"; f.layoutIllum = 1
    if len({v.getAngle() for v in viewSetups}) > 1:
        f.fileNamePattern += "_Angle{a}"; f.layoutAngles = 1
    f.fileNamePattern += ".tif"
    if compress:
        f.fileNamePattern += ".zip"
    return f


def integer_pattern(ids: Sequence[int]) -> str:
    """Resave_TIFF.listAllTimePoints: the ids as a comma separated list (ranges are not collapsed by the reference either)."""
    return ",".join(str(int(i)) for i in ids)


@dataclass
class SpimData2:
    """The part of spim.fiji.spimdata.SpimData2 the exporters touch: sequence description (time points, view setups, stack
    loader), one transform list per (timepoint, setup), empty interest points / bounding boxes."""
    basePath: str
    timepoints: List[TimePoint]
    viewSetups: List[ViewSetup]
    loader: Optional[FileNamePattern] = None
    registrations: Dict[Tuple[int, int], List[Tuple[str, Tuple[float, ...]]]] = field(default_factory=dict)

    def to_xml(self) -> "_ET.Element":
        root = _ET.Element("SpimData", {"version": "0.2"})
        _ET.SubElement(root, "BasePath", {"type": "relative"}).text = "."
        seq = _ET.SubElement(root, "SequenceDescription")
        if self.loader is not None:
            il = _ET.SubElement(seq, "ImageLoader", {"format": "spimreconstruction.stack.ij"})
            _ET.SubElement(il, "imagedirectory", {"type": "relative"}).text = "."
            _ET.SubElement(il, "filePattern").text = self.loader.fileNamePattern
            _ET.SubElement(il, "layoutTimepoints").text = str(self.loader.layoutTP)
            _ET.SubElement(il, "layoutChannels").text = str(self.loader.layoutChannels)
            _ET.SubElement(il, "layoutIlluminations").text = str(self.loader.layoutIllum)
            _ET.SubElement(il, "layoutAngles").text = str(self.loader.layoutAngles)
            _ET.SubElement(il, "imglib2container").text = "ArrayImgFactory"
        vss = _ET.SubElement(seq, "ViewSetups")
        for vs in self.viewSetups:
            e = _ET.SubElement(vss, "ViewSetup")
            _ET.SubElement(e, "id").text = str(vs.id)
            if vs.name:
                _ET.SubElement(e, "name").text = vs.name
            if vs.size_xyz is not None:
                _ET.SubElement(e, "size").text = " ".join(str(int(x)) for x in vs.size_xyz)
            vx = _ET.SubElement(e, "voxelSize")
            _ET.SubElement(vx, "unit").text = vs.voxel_unit
            _ET.SubElement(vx, "size").text = " ".join(repr(float(x)) for x in vs.voxel_size_xyz)
            at = _ET.SubElement(e, "attributes")
            _ET.SubElement(at, "illumination").text = str(vs.illumination.id)
            _ET.SubElement(at, "channel").text = str(vs.channel.id)
            _ET.SubElement(at, "angle").text = str(vs.angle.id)
        for tag, cls, get in (("illumination", "Illumination", ViewSetup.getIllumination), ("channel", "Channel", ViewSetup.getChannel),
                              ("angle", "Angle", ViewSetup.getAngle)):
            a = _ET.SubElement(vss, "Attributes", {"name": tag})
            for ent in sorted({get(v) for v in self.viewSetups}, key=lambda q: q.id):
                e = _ET.SubElement(a, cls)
                _ET.SubElement(e, "id").text = str(ent.id)
                _ET.SubElement(e, "name").text = ent.getName()
        tps = _ET.SubElement(seq, "Timepoints", {"type": "pattern"})
        _ET.SubElement(tps, "integerpattern").text = integer_pattern([t.id for t in self.timepoints])
        _ET.SubElement(seq, "MissingViews")
        regs = _ET.SubElement(root, "ViewRegistrations")
        for tp in self.timepoints:
            for vs in self.viewSetups:
                r = _ET.SubElement(regs, "ViewRegistration", {"timepoint": str(tp.id), "setup": str(vs.id)})
                for name, m in self.registrations.get((tp.id, vs.id), [("identity", (1., 0., 0., 0., 0., 1., 0., 0., 0., 0., 1., 0.))]):
                    t = _ET.SubElement(r, "ViewTransform", {"type": "affine"})
                    _ET.SubElement(t, "Name").text = name
                    _ET.SubElement(t, "affine").text = " ".join(repr(float(x)) for x in m)
        _ET.SubElement(root, "ViewInterestPoints")
        _ET.SubElement(root, "BoundingBoxes")
        return root

    def save(self, xml_path: str) -> None:
        root = self.to_xml()
        _ET.indent(root, space="  ")
        _ET.ElementTree(root).write(xml_path, encoding="UTF-8", xml_declaration=True)


def load_spimdata_xml(xml_path: str) -> SpimData2:
    """Reader for what SpimData2.save wrote (and for hand-written projects of the same schema)."""
    root = _ET.parse(xml_path).getroot()
    seq = root.find("SequenceDescription")
    ents = {}
    for a in seq.find("ViewSetups").findall("Attributes"):
        ents[a.get("name")] = {int(e.find("id").text): Entity(int(e.find("id").text), (e.find("name").text or "")) for e in a}
    setups = []
    for e in seq.find("ViewSetups").findall("ViewSetup"):
        at = e.find("attributes")
        size = e.find("size")
        vx = e.find("voxelSize")
        setups.append(ViewSetup(int(e.find("id").text),
                                angle=ents["angle"][int(at.find("angle").text)],
                                channel=ents["channel"][int(at.find("channel").text)],
                                illumination=ents["illumination"][int(at.find("illumination").text)],
                                name=(e.find("name").text if e.find("name") is not None else ""),
                                size_xyz=tuple(int(x) for x in size.text.split()) if size is not None else None,
                                voxel_size_xyz=tuple(float(x) for x in vx.find("size").text.split()),
                                voxel_unit=vx.find("unit").text))
    tps = [TimePoint(int(x)) for x in seq.find("Timepoints").find("integerpattern").text.split(",") if x.strip()]
    il = seq.find("ImageLoader")
    loader = None
    if il is not None:
        loader = FileNamePattern(int(il.find("layoutTimepoints").text), int(il.find("layoutChannels").text),
                                 int(il.find("layoutIlluminations").text), int(il.find("layoutAngles").text), il.find("filePattern").text)
    regs = {}
    for r in root.find("ViewRegistrations").findall("ViewRegistration"):
        regs[(int(r.get("timepoint")), int(r.get("setup")))] = [
            (t.find("Name").text, tuple(float(x) for x in t.find("affine").text.split())) for t in r.findall("ViewTransform")]
    return SpimData2(os.path.dirname(os.path.abspath(xml_path)), tps, setups, loader, regs)


def fusion_bounding_box_transform(bb_min: Sequence[int], downsampling: float = 1.0) -> Tuple[str, Tuple[float, ...]]:
    """ExportSpimData2TIFF.java:91-99 / AppendSpimData2.java:93-101: the registration of an exported volume is its bounding
    box -- scale = downsampling, translation = bb.min -- replacing whatever transforms the view had."""
    s = float(downsampling)
    return ("fusion bounding box", (s, 0.0, 0.0, float(bb_min[0]), 0.0, s, 0.0, float(bb_min[1]), 0.0, 0.0, s, float(bb_min[2])))


class ExportSpimData2TIFF:
    """"Save as new XML Project (TIFF)", ExportSpimData2TIFF.java: stacks named by XMLTIFFImgTitler next to a new XML whose
    loader pattern finds them again; finish() writes the XML and returns False (the caller's project was not modified)."""

    def __init__(self, xml_path: str, compress: bool = False):
        self.xml_path, self.compress = xml_path, compress
        self.newTimepoints = self.newViewSetups = None
        self.saver = self.spimData = None

    def setXMLData(self, newTimepoints: Sequence[TimePoint], newViewSetups: Sequence[ViewSetup]) -> None:
        self.newTimepoints, self.newViewSetups = list(newTimepoints), list(newViewSetups)

    def queryParameters(self) -> bool:
        if self.newTimepoints is None or self.newViewSetups is None:
            return False      # "new timepoints and new viewsetup list not set yet ... cannot continue"
        base = os.path.dirname(os.path.abspath(self.xml_path))
        self.saver = Save3dTIFF(base, self.compress)
        self.saver.setImgTitler(XMLTIFFImgTitler(self.newTimepoints, self.newViewSetups))
        self.spimData = SpimData2(base, self.newTimepoints, self.newViewSetups,
                                  getFileNamePattern(self.newTimepoints, self.newViewSetups, self.compress))
        return True

    def exportImage(self, img, bb_min: Sequence[int], tp: TimePoint, vs: ViewSetup, downsampling: int = 1,
                    min: float = float("nan"), max: float = float("nan")) -> bool:
        if not self.saver.exportImage(img, bb_min, downsampling, tp, vs, min, max):
            return False
        self.spimData.registrations[(tp.getId(), vs.getId())] = [fusion_bounding_box_transform(bb_min, downsampling)]
        return True

    def finish(self) -> bool:
        self.spimData.save(self.xml_path)
        return False

    def getDescription(self) -> str:
        return "Save as new XML Project (TIFF)"


class AppendSpimData2(ExportSpimData2TIFF):
    """"Append to current XML Project", AppendSpimData2.java (stack-loader branch): the new view setups / time points join an
    EXISTING project -- its loader pattern must be able to name them (:219-262 checks exactly that) -- the stacks are written
    next to it, and finish() returns True: the project object was modified and the caller saves it."""

    def __init__(self, spimData: SpimData2, xml_path: str):
        super().__init__(xml_path, False)
        self.existing = spimData

    def queryParameters(self) -> bool:
        if self.newTimepoints is None or self.newViewSetups is None:
            return False
        if self.existing.loader is None:
            raise NotImplementedError("AppendSpimData2: only projects with a stack image loader (TIFF) can be appended to; the HDF5 "
                                      "branch (AppendSpimData2HDF5) needs an HDF5 library that this image does not provide")
        ids = {v.id for v in self.existing.viewSetups}
        if any(v.id in ids for v in self.newViewSetups):
            raise ValueError("AppendSpimData2: new view setup ids collide with existing ones")
        all_tps = sorted({t.id for t in self.existing.timepoints} | {t.id for t in self.newTimepoints})
        all_vs = list(self.existing.viewSetups) + list(self.newViewSetups)
        need = getFileNamePattern([TimePoint(t) for t in all_tps], all_vs)
        have = self.existing.loader
        for a, b, what in ((need.layoutTP, have.layoutTP, "timepoints"), (need.layoutChannels, have.layoutChannels, "channels"),
                           (need.layoutIllum, have.layoutIllum, "illuminations"), (need.layoutAngles, have.layoutAngles, "angles")):
            if a and not b:
                raise ValueError(f"AppendSpimData2: the project's file pattern '{have.fileNamePattern}' cannot distinguish several {what}")
        self.saver = Save3dTIFF(self.existing.basePath, False)
        titler = _PatternTitler(have)
        self.saver.setImgTitler(titler)
        self.existing.viewSetups = all_vs
        self.existing.timepoints = [TimePoint(t) for t in all_tps]
        self.spimData = self.existing
        return True

    def finish(self) -> bool:
        return True

    def getDescription(self) -> str:
        return "Append to current XML Project"


class _PatternTitler:
    """file name of (timepoint, setup) under a stack loader's pattern: {t} {c} {i} {a} replaced (StackImgLoader.getFileName)"""

    def __init__(self, pattern: FileNamePattern):
        self.p = pattern

    def getImageTitle(self, tp: TimePoint, vs: ViewSetup) -> str:
        return (self.p.fileNamePattern.replace("{t}", str(tp.getId())).replace("{c}", vs.getChannel().getName())
                .replace("{i}", vs.getIllumination().getName()).replace("{a}", vs.getAngle().getName()))


class ExportSpimData2HDF5:
    """ExportSpimData2HDF5.java writes BigDataViewer HDF5 (multi-resolution chunked int16) through the un-vendored
    bdv / jhdf5 libraries.  No HDF5 library exists in this image (h5py absent), so this exporter is declared, not built."""

    def __init__(self, *a, **k):
        raise NotImplementedError("HDF5 export needs an HDF5 library (h5py / jhdf5); use ExportSpimData2TIFF")


AppendSpimData2HDF5 = ExportSpimData2HDF5
