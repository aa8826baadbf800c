"""Self-checks that pin oracle/fusion_oracle.py (the reference has no tests or golden vectors for this path):
every piece is compared with an independent implementation (scipy.ndimage, numpy.linalg, closed forms)."""
import math

import numpy as np
import pytest
import scipy.ndimage as ndi

from oracle import fusion_oracle as F


def _rand(shape, seed=0):
    return np.random.default_rng(seed).random(shape, dtype=np.float32)


@pytest.mark.parametrize("ext,mode", [(F.EXT_MIRROR_SINGLE, "mirror"), (F.EXT_PERIODIC, "grid-wrap"), (F.EXT_ZERO, "grid-constant")])
def test_nlinear_matches_ndimage_order1(ext, mode):
    src = _rand((6, 7, 8), 1)
    rng = np.random.default_rng(2)
    pz, py, px = rng.uniform(-9, 15, 500), rng.uniform(-9, 16, 500), rng.uniform(-9, 17, 500)
    got = F.nlinear3d(src, px, py, pz, ext)
    want = ndi.map_coordinates(src.astype(np.float64), [pz, py, px], order=1, mode=mode, cval=0.0)
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)


def test_nlinear_is_exact_on_grid_points():
    src = _rand((4, 5, 6), 3)
    z, y, x = np.meshgrid(np.arange(4), np.arange(5), np.arange(6), indexing="ij")
    np.testing.assert_array_equal(F.nlinear3d(src, x, y, z, F.EXT_ZERO), src)


def test_invert_affine_matches_linalg():
    rng = np.random.default_rng(4)
    for _ in range(20):
        m = rng.normal(size=12)
        inv = F.invert_affine(m)
        a = np.vstack([m.reshape(3, 4), [0, 0, 0, 1]])
        np.testing.assert_allclose(np.vstack([inv.reshape(3, 4), [0, 0, 0, 1]]), np.linalg.inv(a), rtol=1e-9, atol=1e-9)
    with pytest.raises(ValueError):
        F.invert_affine([0.0] * 12)


def test_blending_lookup_table():
    lut = F.blending_lookup()
    assert lut[0] == 0.0 and abs(lut[1000] - 1.0) < 1e-12
    i = np.arange(1001)
    np.testing.assert_allclose(lut, (np.cos((1 - i / 1000.0) * math.pi) + 1) / 2, atol=1e-12)
    assert np.all(np.diff(lut) > 0)


def test_blending_weight_closed_form():
    dims = (30, 20, 10)
    border, rng_ = (2.0, 1.0, 0.0), (6.0, 5.0, 3.0)
    r = np.random.default_rng(5)
    t = [r.uniform(-3, d + 3, 2000).astype(np.float32) for d in dims]
    got = F.blending_weight(t[0], t[1], t[2], dims, border, rng_)
    want = np.ones(2000)
    for d in range(3):
        dist = np.maximum(0, np.minimum(t[d] - border[d], (dims[d] - 1) - t[d] - border[d]))
        rel = dist / rng_[d]
        want *= np.where(rel < 1, (np.cos((1 - rel) * math.pi) + 1) / 2, 1.0) * (dist > 0)
    # the table quantises rel to 1/1000: |d/drel| <= pi/2 per axis
    np.testing.assert_allclose(got, want, atol=3 * 0.5e-3 * math.pi / 2 + 1e-6)
    assert got.min() == 0.0 and got.max() == 1.0


def test_identity_transform_reproduces_the_stack():
    st = _rand((5, 6, 7), 6) + np.float32(0.5)
    ident = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]
    np.testing.assert_array_equal(F.transform_input(st, ident, st.shape, (0, 0, 0)), st)
    # shifted bounding box: outside the stack the image is 0, inside it is clamped to minValue
    out = F.transform_input(np.zeros_like(st), ident, (5, 6, 9), (-1, 0, 0))
    assert np.all(out[:, :, 0] == 0) and np.all(out[:, :, 8] == 0) and np.all(out[:, :, 1:8] == F.MIN_VALUE)


def test_transform_matches_ndimage_affine():
    st = _rand((8, 9, 10), 7)
    a = math.radians(25.0)
    m = [math.cos(a), 0, math.sin(a), 1.0, 0, 1, 0, 0.5, -math.sin(a), 0, math.cos(a), 2.0]
    inv = F.invert_affine(m)
    out_dims, off = (10, 9, 12), (-1, 0, 1)
    got = F.transform_input(st, inv, out_dims, off)
    t0, t1, t2 = F.inverse_positions(out_dims, off, inv)
    want = ndi.map_coordinates(st.astype(np.float64), [t2, t1, t0], order=1, mode="mirror")
    inside = F.intersects(t0, t1, t2, st.shape[::-1])
    np.testing.assert_allclose(got[inside], np.maximum(1e-4, want[inside]), atol=2e-6)
    assert np.all(got[~inside] == 0) and inside.any() and (~inside).any()


def test_weight_normalizer_rules():
    rng = np.random.default_rng(8)
    ws = [np.where(rng.random((4, 5, 6)) < 0.3, 0, rng.random((4, 5, 6))).astype(np.float32) for _ in range(4)]
    ws[0][0, 0, :] = ws[1][0, 0, :] = ws[2][0, 0, :] = ws[3][0, 0, :] = 0     # nobody covers this line
    wn, mn, avg = F.weight_normalizer_direct(ws, 1)
    s = sum(w.astype(np.float64) for w in wn)
    covered = sum(w for w in ws) > 0
    np.testing.assert_allclose(s[covered], 1.0, atol=3e-7)
    assert np.isnan(wn[0][0, 0, :]).all() and mn == 0
    cnt = sum((w > 0).astype(int) for w in ws)
    assert avg == cnt.sum() / cnt.size
    # portions: mean of portion means, last portion takes the remainder
    _, mn2, avg2 = F.weight_normalizer_direct(ws, 7)
    flat = cnt.reshape(-1)
    chunk = flat.size // 7
    means = [flat[i * chunk:(i + 1) * chunk].mean() for i in range(6)] + [flat[6 * chunk:].mean()]
    assert mn2 == 0 and abs(avg2 - np.mean(means)) < 1e-12
    sumw, _, _ = F.weight_normalizer_virtual(ws, 1)
    assert sumw.min() == 1.0
    v = [F.normalizing_access(w, sumw, 1.0) for w in ws]
    tot = sum(x.astype(np.float64) for x in v)
    assert tot.max() <= 1.0 + 3e-7 and not np.isnan(tot).any()
    assert F.normalizing_access(ws[1], sumw, 50.0).max() == 1.0


def test_loader_normalize_range():
    a = _rand((3, 4, 5), 9) * 7 - 2
    n = F.loader_normalize(a)
    assert n.min() == 0.0 and n.max() == 1.0


def test_extract_psf_single_integer_bead_is_a_crop():
    img = _rand((9, 10, 11), 10)
    psf = F.extract_psf_local(img, [(5.0, 4.0, 4.0)], (3, 5, 3))
    np.testing.assert_array_equal(psf, img[3:6, 2:7, 4:7])
    # two beads add up; periodic wrap at the faces
    psf2 = F.extract_psf_local(img, [(5.0, 4.0, 4.0), (0.0, 0.0, 0.0)], (3, 5, 3))
    wrap = np.take(np.take(np.take(img, [-1, 0, 1], axis=0), [-2, -1, 0, 1, 2], axis=1), [-1, 0, 1], axis=2)
    np.testing.assert_array_equal(psf2, (psf + wrap).astype(np.float32))


def test_transform_psf_geometry_and_identity():
    psf = _rand((5, 7, 9), 11)
    ident = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]
    size, off = F.transform_psf_geometry(psf.shape, ident)
    assert size == [9, 7, 5] and off == [0.0, 0.0, 0.0]
    np.testing.assert_array_equal(F.transform_psf(psf, ident), psf)
    # anisotropic calibration: z scaled by 2.5 -> (int)(4 * 2.5) + 1 = 11, centre voxel stays the centre
    scal = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 2.5, 0]
    size, off = F.transform_psf_geometry(psf.shape, scal)
    assert size == [9, 7, 11]
    t = F.transform_psf(psf, scal)
    assert t.shape == (11, 7, 9) and t[5, 3, 4] == psf[2, 3, 4]
    # rotation by 90 degrees about y swaps the x and z extents
    rot = [0, 0, 1, 0, 0, 1, 0, 0, -1, 0, 0, 0]
    size, _ = F.transform_psf_geometry(psf.shape, rot)
    assert size == [5, 7, 9]
    r = F.transform_psf(psf, rot)
    assert r[4, 3, 2] == psf[2, 3, 4]
    np.testing.assert_allclose(r, np.transpose(psf, (2, 1, 0))[::-1, :, :], atol=1e-6)


def test_make_same_size_and_common_size():
    a, b = _rand((3, 5, 7), 12), _rand((5, 3, 9), 13)
    assert F.common_size([a, b]) == [9, 5, 5]
    m = F.make_same_size(a, (9, 5, 5))
    assert m.shape == (5, 5, 9)
    np.testing.assert_array_equal(m[1:4, :, 1:8], a)
    assert np.all(m[0] == a.min()) and np.all(m[:, :, 0] == a.min())
    c = F.make_same_size(a, (3, 3, 1))       # cropping keeps the centre
    np.testing.assert_array_equal(c, a[1:2, 1:4, 2:5])
