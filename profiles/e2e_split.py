"""Where the end-to-end time of the plugin call goes (configs[1], pinned host arrays):   python profiles/e2e_split.py"""
import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from spim_registration_b200.deconvolution import MVDeconFFT, MVDeconInput, MVDeconvolution, PSFTYPE, Session
imgs, ws, psfs = bench.make_inputs(bench.BRICK, fast=True)
pin_img = [torch.from_numpy(a).pin_memory() for a in imgs]
pin_w = [torch.from_numpy(a).pin_memory() for a in ws]
out = torch.empty(bench.BRICK, dtype=torch.float32).pin_memory()
for rep in range(2):
    t0 = time.perf_counter()
    s = Session(bench.BRICK, 7, 2, generation=2, lam=0.006)
    t1 = time.perf_counter()
    for v in range(7):
        s.set_view_ptr(v, pin_img[v].data_ptr(), pin_w[v].data_ptr(), psfs[v])
    t2 = time.perf_counter()
    s.init()
    t3 = time.perf_counter()
    s.run(20, stats=True)
    t4 = time.perf_counter()
    s.finish(); s.get_psi_ptr(out.data_ptr())
    t5 = time.perf_counter()
    s.close()
    print(f"create {t1-t0:.4f} upload {t2-t1:.4f} ({3.76/(t2-t1):.1f} GB/s) init {t3-t2:.4f} run20 {t4-t3:.4f} finish+download {t5-t4:.4f} total {t5-t0:.4f}")
# the plugin call itself (what bench.py times as e2e)
views_np = [(pin_img[v].numpy(), pin_w[v].numpy()) for v in range(7)]
for rep in range(3):
    t0 = time.perf_counter()
    views = MVDeconInput()
    for v in range(7):
        views.add(MVDeconFFT(views_np[v][0], views_np[v][1], psfs[v], None, (0,), False, None, False))
    decon = MVDeconvolution(views, PSFTYPE(2), 20, 0.006, 1.0, 0, "bench")
    decon.getPsi(out.numpy())
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    del decon, views
    print(f"MVDeconvolution(...) 20 iterations + getPsi: {t1-t0:.4f} s = {67108864*7*20/(t1-t0)/1e9:.2f} G vvi/s")
