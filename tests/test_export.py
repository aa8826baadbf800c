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
