"""Fusion pre-step under the CPU emulator: the very kernel bodies of csrc/fusion.h,
compiled with -DSPIM_HOST_EMU, against oracle/fusion_oracle.py -- bit-exact.  Host-logic tests; the GPU versions of the
same cases are in test_gpu_fusion.py."""
import numpy as np
import pytest

import fusion_cases as FC
from oracle import fusion_oracle as F
from oracle import mvdecon_oracle as O
from spim_registration_b200 import fusion, native
from spim_registration_b200.deconvolution import Session


def test_blending_table_matches_reference_loop(emu_lib):
    np.testing.assert_array_equal(fusion.blending_lookup(emu_lib), F.blending_lookup())


@pytest.mark.parametrize("angle", [0.0, 17.0, 90.0, 200.0])
def test_transform_image_and_weights(emu_lib, angle):
    FC.transform_case(emu_lib, (9, 20, 22), (14, 18, 24), angle, (-2, 1, 3), (2, 2, 1), (6, 6, 3))


def test_transform_negative_border_and_offsets(emu_lib):
    # "the border can be negative": weights are > 0 outside the stack, the image is 0 there
    FC.transform_case(emu_lib, (7, 12, 14), (12, 16, 20), 33.0, (-5, -3, -4), (-8, -8, -3), (12, 12, 12))


def test_transform_weights_only_and_image_only(emu_lib):
    FC.transform_case(emu_lib, (6, 10, 12), (8, 12, 14), 45.0, (0, 0, 0), (1, 1, 1), (4, 4, 2), weights=True, image=False)
    FC.transform_case(emu_lib, (6, 10, 12), (8, 12, 14), 45.0, (0, 0, 0), (1, 1, 1), (4, 4, 2), weights=False, image=True)


def test_transform_with_loader_normalisation(emu_lib):
    FC.transform_case(emu_lib, (8, 12, 10), (10, 12, 12), 10.0, (0, -1, 0), (0, 0, 0), (3, 3, 3), normalize=True)


def test_transform_degenerate_stack_dims(emu_lib):
    # one-voxel-thick stack: mirror-single of n == 1, every tap maps to index 0
    FC.transform_case(emu_lib, (1, 8, 9), (3, 8, 9), 0.0, (0, 0, -1), (0, 0, 0), (2, 2, 2), z_scale=1.0)


@pytest.mark.parametrize("virtual", [False, True])
@pytest.mark.parametrize("num_portions,osem_index,osem", [(2, 0, 1.0), (8, 0, 2.0), (6, 1, 1.0), (4, 2, 1.0), (2, 3, 1.5)])
def test_weight_normalizer_and_osem(emu_lib, virtual, num_portions, osem_index, osem):
    FC.normalize_case(emu_lib, virtual, num_portions, osem_index, osem)


def test_precomputed_weights_keep_reference_nan(emu_lib):
    # ApplyDirectly divides everywhere: voxels no view covers get 0/0 = NaN (WeightNormalizer.java:168-170)
    V, sd, od = 2, (5, 8, 8), (6, 10, 30)
    stacks, models = FC.make_view_set(V, sd, od)
    with Session(od, V, O.INDEPENDENT, lib=emu_lib) as s:
        for v in range(V):
            fusion.load_stack(s, stacks[v])
            fusion.transform_view(s, v, models[v], (0, 0, 0), fusion.Blending(sd[::-1], (0, 0, 0), (2, 2, 2)))
        fusion.normalize_weights(s, virtual=False, num_portions=3)
        w = fusion.get_view(s, 0, 1)
    assert np.isnan(w).any() and np.isfinite(w).any()


@pytest.mark.parametrize("typ", [O.EFFICIENT_BAYESIAN, O.OPTIMIZATION_II])
def test_pipeline_stacks_to_deconvolution(emu_lib, typ):
    FC.pipeline_case(emu_lib, typ=typ)


def test_psf_extraction_and_transform(emu_lib):
    FC.psf_case(emu_lib)
    FC.psf_case(emu_lib, stack_shape=(8, 9, 10), n_beads=3, psf_size_xyz=(5, 5, 3), angle=120.0, seed=9)


def test_identity_properties(emu_lib):
    FC.identity_properties_case(emu_lib, (30, 34, 40))


def test_fusion_error_paths(emu_lib):
    with Session((4, 4, 4), 2, O.INDEPENDENT, lib=emu_lib) as s:
        m = fusion.AffineTransform3D()
        with pytest.raises(native.NativeError, match="no stack loaded"):
            fusion.transform_view(s, 0, m, (0, 0, 0))
        fusion.load_stack(s, np.ones((2, 2, 2), np.float32))
        with pytest.raises(native.NativeError, match="out of range"):
            fusion.transform_view(s, 5, m, (0, 0, 0))
        fusion.transform_view(s, 0, m, (0, 0, 0), fusion.Blending((2, 2, 2), (0, 0, 0), (1, 1, 1)))
        with pytest.raises(native.NativeError, match="has no weight image"):
            fusion.normalize_weights(s, virtual=False)
        with pytest.raises(native.NativeError, match="num_portions"):
            native.check(s.lib, s.lib.mvd_normalize_weights(s._h, 0, 0, None, None), "mvd_normalize_weights")
        with pytest.raises(native.NativeError, match="buffer not set"):
            fusion.get_view(s, 1, 0)
    with pytest.raises(RuntimeError, match="singular"):
        fusion.AffineTransform3D([0] * 12).inverse()
