"""GPU parity tests of the fusion pre-step: the CUDA kernels of csrc/fusion.h through
the C-ABI of include/spim_fusion.h against oracle/fusion_oracle.py -- bit-exact (the Java arithmetic is reproduced
operation by operation, no FMA contraction).  Run on the B200 box with  python -m pytest tests -m gpu.

The file name sorts last on purpose: these rows were written after the round's GPU budget was spent, so under `pytest -x`
the hot-path parity tests (test_gpu_parity.py), which have run on hardware, are reported before anything here can stop the run."""
import numpy as np
import pytest

import fusion_cases as FC
import parity_cases as P
from oracle import fusion_oracle as F
from oracle import mvdecon_oracle as O
from spim_registration_b200 import fusion
from spim_registration_b200.deconvolution import Session

pytestmark = pytest.mark.gpu


def test_blending_table(gpu):
    np.testing.assert_array_equal(fusion.blending_lookup(gpu), F.blending_lookup())


@pytest.mark.parametrize("angle", [0.0, 17.0, 90.0, 200.0])
def test_transform_image_and_weights(gpu, angle):
    FC.transform_case(gpu, (9, 20, 22), (14, 18, 24), angle, (-2, 1, 3), (2, 2, 1), (6, 6, 3))


def test_transform_larger_volume(gpu):
    # reference defaults: border -8,-8,round(-8/2.5), range 12 (FD/EfficientBayesianBased.java:96-97, 694-696)
    FC.transform_case(gpu, (40, 96, 100), (90, 100, 110), 45.0, (-4, -2, 3), (-8, -8, -3), (12, 12, 12), normalize=True)


def test_transform_negative_border_and_offsets(gpu):
    FC.transform_case(gpu, (7, 12, 14), (12, 16, 20), 33.0, (-5, -3, -4), (-8, -8, -3), (12, 12, 12))


def test_transform_weights_only_and_image_only(gpu):
    FC.transform_case(gpu, (6, 10, 12), (8, 12, 14), 45.0, (0, 0, 0), (1, 1, 1), (4, 4, 2), weights=True, image=False)
    FC.transform_case(gpu, (6, 10, 12), (8, 12, 14), 45.0, (0, 0, 0), (1, 1, 1), (4, 4, 2), weights=False, image=True)


def test_transform_degenerate_stack_dims(gpu):
    FC.transform_case(gpu, (1, 8, 9), (3, 8, 9), 0.0, (0, 0, -1), (0, 0, 0), (2, 2, 2), z_scale=1.0)


@pytest.mark.parametrize("virtual", [False, True])
@pytest.mark.parametrize("num_portions,osem_index,osem", [(2, 0, 1.0), (8, 0, 2.0), (6, 1, 1.0), (4, 2, 1.0), (2, 3, 1.5)])
def test_weight_normalizer_and_osem(gpu, virtual, num_portions, osem_index, osem):
    FC.normalize_case(gpu, virtual, num_portions, osem_index, osem)


def test_weight_normalizer_many_portions_larger_volume(gpu):
    FC.normalize_case(gpu, True, 64, 2, 1.0, V=4, stack_shape=(30, 70, 72), out_dims=(60, 64, 80))
    FC.normalize_case(gpu, False, 48, 1, 1.0, V=4, stack_shape=(30, 70, 72), out_dims=(60, 64, 80))


@pytest.mark.parametrize("typ", [O.EFFICIENT_BAYESIAN, O.OPTIMIZATION_II, O.INDEPENDENT])
def test_pipeline_stacks_to_deconvolution(gpu, typ):
    FC.pipeline_case(gpu, typ=typ)


def test_psf_extraction_and_transform(gpu):
    FC.psf_case(gpu)
    FC.psf_case(gpu, stack_shape=(8, 9, 10), n_beads=3, psf_size_xyz=(5, 5, 3), angle=120.0, seed=9)
    FC.psf_case(gpu, stack_shape=(40, 80, 90), n_beads=40, psf_size_xyz=(19, 19, 25), angle=45.0, seed=21)


def test_full_size_identity_properties(gpu):
    """C2-sized view (512 x 512 x 256): size-independent properties instead of an oracle run."""
    FC.identity_properties_case(gpu, (256, 512, 512))


def test_views_uploaded_in_cells(gpu):
    P.cells_case(gpu)
    P.cells_case(gpu, shape=(40, 50, 60), cell=(16, 32, 24), V=3, ks=7)


def test_cpp_fusion_mirror_on_gpu(gpu, tmp_path):
    from spim_registration_b200 import native
    from test_cpp_fusion_mirror import check
    check(native.default_library_path(), tmp_path)


def test_fast_epilogue_switch(gpu):
    """opt-in MUFU-seeded division / square root in the fused epilogues (mvd_params.fast_epilogue)"""
    P.fast_epilogue_case(gpu)
    P.fast_epilogue_case(gpu, shape=(40, 48, 56))
