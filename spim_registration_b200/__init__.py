"""spim_registration_b200 -- B200-native multi-view Bayesian deconvolution (the hot path of
fiji/SPIM_Registration) behind the reference's own native boundary.

Layout:
  csrc/              CUDA kernels for sm_100a + the C-ABI (include/*.h)
  native.py          ctypes twin of the JNA interface ``CUDAFourierConvolution``
  deconvolution.py   host-side mirror of LRFFT / LRInput / BayesMVDeconvolution and
                     MVDeconFFT / MVDeconInput / MVDeconvolution
  blocks.py          Block / BlockGeneratorFixedSizePrecise
  cuda.py            CUDADevice / CUDATools / NativeLibraryTools
  bricks.py          multi-GPU brick partition + halo exchange (one process per GPU)
  synthetic.py       deterministic synthetic bead / specimen datasets

The CUDA library is mandatory; nothing here computes on the CPU.
"""
from .deconvolution import (PSFTYPE, LRFFT, LRInput, BayesMVDeconvolution, MVDeconFFT, MVDeconInput,
                            MVDeconvolution, Deconvolver, Session, minValue)
from .blocks import Block, BlockGeneratorFixedSizePrecise
from .cuda import CUDADevice, CUDATools, NativeLibraryTools
from .native import CUDAFourierConvolution, load_library

__all__ = [
    "PSFTYPE", "LRFFT", "LRInput", "BayesMVDeconvolution", "MVDeconFFT", "MVDeconInput", "MVDeconvolution",
    "Deconvolver", "Session", "minValue", "Block", "BlockGeneratorFixedSizePrecise", "CUDADevice", "CUDATools",
    "NativeLibraryTools", "CUDAFourierConvolution", "load_library",
]
