"""Output side: 32-bit TIFF stacks in ImageJ's layout, naming rules, PSF loading from TIFF.
The written files are cross-read with two independent TIFF decoders (Pillow, OpenCV)."""
import os

import numpy as np
import pytest

from oracle import fusion_oracle as F
from spim_registration_b200 import export, fusion


def _vol(shape, seed=0):
    return (np.random.default_rng(seed).random(shape, dtype=np.float32) * 5 - 1).astype(np.float32)


@pytest.mark.parametrize("shape", [(5, 7, 9), (1, 4, 6), (12, 33, 17)])
def test_tiff_round_trip_and_independent_readers(tmp_path, shape):
    a = _vol(shape, 1)
    fn = str(tmp_path / "stack.tif")
    export.write_tiff_stack(fn, a, origin_xyz=(2, -3, 0), spacing=1.0)
    np.testing.assert_array_equal(export.read_tiff_stack(fn), a)
    from PIL import Image
    im = Image.open(fn)
    assert getattr(im, "n_frames", 1) == shape[0] and im.size == (shape[2], shape[1]) and im.mode == "F"
    for z in range(shape[0]):
        im.seek(z)
        np.testing.assert_array_equal(np.asarray(im, dtype=np.float32), a[z])
    desc = im.tag_v2[270] if shape[0] == 1 else Image.open(fn).tag_v2[270]
    assert desc.startswith("ImageJ=") and ("images=%d" % shape[0] in desc) == (shape[0] > 1)
    assert "xorigin=2.0" in desc and "yorigin=-3.0" in desc and "zorigin" not in desc
    import cv2
    ok, pages = cv2.imreadmulti(fn, flags=cv2.IMREAD_UNCHANGED)
    assert ok and len(pages) == shape[0]
    np.testing.assert_array_equal(np.stack(pages), a)


def test_reader_handles_little_endian_integer_tiffs(tmp_path):
    from PIL import Image
    a8 = (np.random.default_rng(2).random((6, 8)) * 255).astype(np.uint8)
    a16 = (np.random.default_rng(3).random((3, 6, 8)) * 60000).astype(np.uint16)
    f8, f16 = str(tmp_path / "a8.tif"), str(tmp_path / "a16.tif")
    Image.fromarray(a8).save(f8)
    pages = [Image.fromarray(p) for p in a16]
    pages[0].save(f16, save_all=True, append_images=pages[1:])
    np.testing.assert_array_equal(export.read_tiff_stack(f8), a8[None].astype(np.float32))
    np.testing.assert_array_equal(export.read_tiff_stack(f16), a16.astype(np.float32))
    with pytest.raises(ValueError):
        open(str(tmp_path / "x.tif"), "wb").write(b"not a tiff")
        export.read_tiff_stack(str(tmp_path / "x.tif"))


def test_naming_rules():
    assert export.java_double_to_string(0.006) == "0.006" and export.java_double_to_string(6e-4) == "6.0E-4"
    assert export.java_double_to_string(1.0) == "1.0" and export.java_double_to_string(12345678.0) == "1.2345678E7"
    assert export.java_double_to_string(0.0) == "0.0" and export.java_double_to_string(-2.5) == "-2.5"
    assert export.gen1_output_name(0.006, 18, 0) == "DC(l=0.006)_t18_ch0"
    assert export.gen1_output_name(0.006, 18, 0, tikhonov=False) == "DC(l=0)_t18_ch0"
    assert export.gen2_output_title("18", "0", ["0", "1"], ["0", "45", "90"]) == "TP18_Ch0_Ill0,1_Ang0,45,90"
    t = export.DefaultImgTitler().getImageTitle(3, {"channel": "c", "illumination": "i", "angle": "a"})
    assert t == "Timepoint3_Channelc_Illumi_Anglea"


def test_save3dtiff_export(tmp_path):
    a = _vol((4, 6, 8), 4)
    ex = export.Save3dTIFF(str(tmp_path))
    assert ex.exportImage(None) is False
    ex.setImgTitler(export.FixedNameImgTitler("TP18_Ch0_Ill0_Ang0,45"))
    assert ex.exportImage(a, bb_min=(-5, 7, 0), downsampling=2)
    fn = str(tmp_path / "TP18_Ch0_Ill0_Ang0,45.tif")
    assert os.path.exists(fn)
    np.testing.assert_array_equal(export.read_tiff_stack(fn), a)
    from PIL import Image
    desc = Image.open(fn).tag_v2[270]
    # origin = -(bb.min / downsampling) with Java's truncating integer division; display range = image min / max
    assert "xorigin=2.0" in desc and "yorigin=-3.0" in desc and "spacing=2.0" in desc
    assert f"min={export.java_double_to_string(float(a.min()))}" in desc
    assert ex.exportImage(a, title="named.tif") and os.path.exists(str(tmp_path / "named.tif"))


def test_load_and_transform_psfs_from_tiff(tmp_path, emu_lib):
    import fusion_cases as FC
    psf = np.random.default_rng(5).random((5, 7, 7), dtype=np.float32)
    fn = str(tmp_path / "psf.tif")
    export.write_tiff_stack(fn, psf)
    model = FC.view_model(40.0, (5, 7, 7), z_scale=2.0)
    e = fusion.ExtractPSF.loadAndTransformPSFs({"v": fn}, ["v"], {"v": model}, lib=emu_lib)
    np.testing.assert_array_equal(e.getInputCalibrationPSFs()["v"], psf)
    np.testing.assert_array_equal(e.getTransformedPSF("v"), F.transform_psf(psf, model.getRowPackedCopy()))
    e2 = fusion.ExtractPSF.loadAndTransformPSFs({"v": fn}, ["v"], None, lib=emu_lib)
    np.testing.assert_array_equal(e2.getTransformedPSF("v"), psf)
    with pytest.raises(RuntimeError):
        fusion.ExtractPSF.loadAndTransformPSFs({}, ["v"], None, lib=emu_lib)


def test_xml_project_export_and_append(tmp_path):
    """ExportSpimData2TIFF / AppendSpimData2 (ExportSpimData2TIFF.java:79-224, AppendSpimData2.java:75-262): file names by
    XMLTIFFImgTitler, loader pattern with the layout flags, 'fusion bounding box' registrations, XML round trip."""
    E = export
    tps = [E.TimePoint(0), E.TimePoint(3)]
    vss = [E.ViewSetup(0, channel=E.Entity(0, "488"), size_xyz=(9, 7, 5)), E.ViewSetup(1, channel=E.Entity(1, "561"), size_xyz=(9, 7, 5))]
    pat = E.getFileNamePattern(tps, vss)
    assert pat.fileNamePattern == "img_TL{t}_Ch{c}.tif" and (pat.layoutTP, pat.layoutChannels, pat.layoutIllum, pat.layoutAngles) == (1, 1, 0, 0)
    assert E.getFileNamePattern(tps[:1], vss[:1]).fileNamePattern == "img.tif"
    assert E.XMLTIFFImgTitler(tps, vss).getImageTitle(tps[1], vss[1]) == "img_TL3_Ch561"
    xml = str(tmp_path / "dataset.xml")
    ex = E.ExportSpimData2TIFF(xml)
    assert ex.queryParameters() is False           # setXMLData first, like the reference
    ex.setXMLData(tps, vss)
    assert ex.queryParameters()
    vols = {}
    for tp in tps:
        for vs in vss:
            vols[(tp.id, vs.id)] = _vol((5, 7, 9), seed=10 * tp.id + vs.id)
            assert ex.exportImage(vols[(tp.id, vs.id)], (10, -20, 30), tp, vs, downsampling=2)
    assert ex.finish() is False
    for (t, v), a in vols.items():
        np.testing.assert_array_equal(E.read_tiff_stack(str(tmp_path / f"img_TL{t}_Ch{vss[v].channel.name}.tif")), a)
    sd = E.load_spimdata_xml(xml)
    assert [t.id for t in sd.timepoints] == [0, 3] and [v.id for v in sd.viewSetups] == [0, 1]
    assert sd.viewSetups[1].channel == E.Entity(1, "561") and sd.viewSetups[0].size_xyz == (9, 7, 5)
    assert sd.loader.fileNamePattern == "img_TL{t}_Ch{c}.tif" and sd.loader.layoutChannels == 1
    name, m = sd.registrations[(3, 1)][0]
    assert name == "fusion bounding box" and m == (2.0, 0, 0, 10.0, 0, 2.0, 0, -20.0, 0, 0, 2.0, 30.0)
    txt = open(xml).read()
    assert '<SpimData version="0.2">' in txt and 'format="spimreconstruction.stack.ij"' in txt and "<integerpattern>0,3</integerpattern>" in txt
    # append a deconvolved channel to the project just written
    new_vs = [E.ViewSetup(2, channel=E.Entity(2, "dc"), size_xyz=(9, 7, 5))]
    ap = E.AppendSpimData2(sd, xml)
    ap.setXMLData(tps, new_vs)
    assert ap.queryParameters()
    d = _vol((5, 7, 9), seed=99)
    assert ap.exportImage(d, (0, 0, 0), tps[0], new_vs[0])
    assert ap.finish() is True                     # the project object was modified: the caller saves it
    sd.save(xml)
    sd2 = E.load_spimdata_xml(xml)
    assert [v.id for v in sd2.viewSetups] == [0, 1, 2] and sd2.registrations[(0, 2)][0][0] == "fusion bounding box"
    np.testing.assert_array_equal(E.read_tiff_stack(str(tmp_path / "img_TL0_Chdc.tif")), d)
    # a project whose pattern has no {c} cannot take a second channel
    one = E.SpimData2(str(tmp_path), tps[:1], vss[:1], E.getFileNamePattern(tps[:1], vss[:1]))
    ap2 = E.AppendSpimData2(one, xml)
    ap2.setXMLData(tps[:1], new_vs)
    with pytest.raises(ValueError):
        ap2.queryParameters()
    with pytest.raises(NotImplementedError):
        E.ExportSpimData2HDF5()
