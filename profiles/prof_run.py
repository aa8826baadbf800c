"""Profiling driver: one device-resident session at the bench brick size, a warm iteration, then
`cudaProfilerStart` .. one iteration .. `cudaProfilerStop` so that ncu (--profile-from-start off)
captures exactly the hot-path kernels.  Usage (on the GPU box):
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof \
      python profiles/prof_run.py [views] [z y x] [psf]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from spim_registration_b200 import synthetic
from spim_registration_b200.deconvolution import Session

V = int(sys.argv[1]) if len(sys.argv) > 1 else 1
shape = tuple(int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (256, 512, 512)
ks = int(sys.argv[5]) if len(sys.argv) > 5 else 31
rng = np.random.default_rng(0)
psfs = synthetic.make_psfs(V, ks)
with Session(shape, V, 3, generation=2, lam=0.006) as s:
    for v in range(V):
        img = rng.random(shape, dtype=np.float32) + 0.01
        s.set_view(v, img, rng.random(shape, dtype=np.float32), psfs[v])
    s.init()
    s.run(1, stats=False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    s.run(1, stats=False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("psi mean", float(s.get_psi().mean()))
