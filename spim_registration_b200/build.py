"""Build the CUDA shared library in-tree with nvcc for sm_100a (no JIT cache: the .so travels with
the repository snapshot to the GPU box).

The library is compiled from several translation units in parallel: csrc/spim_b200.cu (C ABI, session, element-wise
and fusion kernels) and csrc/inst.cu once per group of csrc/instances.h (the FFT kernels' instantiations) -- the heavy
templates cost minutes of compiler front-end time each, so one file per core turns a 5.5-minute build into about two.
SPIM_SINGLE_TU=1 selects the plain one-file build (same code, implicit instantiation)."""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SRC = os.path.join(CSRC, "spim_b200.cu")
INST = os.path.join(CSRC, "inst.cu")
DEPS = [os.path.join(CSRC, f) for f in ("spim_b200.cu", "inst.cu", "instances.h", "engine.h", "kernels.h", "fft_math.h", "fast_math.h",
                                        "hd.h", "runtime.h", "fusion.h", "fusion_api.h")] + \
       [os.path.join(HERE, "..", "include", f) for f in ("spim_fftconv.h", "spim_mvdecon.h", "spim_fusion.h")]
OUT = os.path.join(HERE, "libConvolution3D_fftCUDAlib.so")
ALIAS = os.path.join(HERE, "libFourierConvolutionCUDALib.so")
OBJDIR = os.path.join(HERE, "build")

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "--expt-relaxed-constexpr", "-Xfatbin=-compress-all", "-Xcompiler", "-fPIC"]


def is_stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def instance_groups():
    """the group names of csrc/instances.h (SPIM_INSTANCE_GROUPS)"""
    txt = open(os.path.join(CSRC, "instances.h")).read()
    m = re.search(r"#define SPIM_INSTANCE_GROUPS (.*)", txt)
    return re.findall(r'"([A-Z_0-9]+)"', m.group(1))


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    return r.returncode, r.stdout + r.stderr


def build_cuda_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    extra = ["-Xptxas", "-v"] if verbose else []
    if os.environ.get("SPIM_SINGLE_TU") == "1":
        jobs, objs = [[nvcc] + NVCC_FLAGS + extra + ["-shared", SRC, "-o", OUT]], []
    else:
        os.makedirs(OBJDIR, exist_ok=True)
        main_o = os.path.join(OBJDIR, "spim_b200.o")
        jobs = [[nvcc] + NVCC_FLAGS + extra + ["-DSPIM_SPLIT_BUILD", "-c", SRC, "-o", main_o]]
        objs = [main_o]
        for g in instance_groups():
            o = os.path.join(OBJDIR, f"inst_{g}.o")
            jobs.append([nvcc] + NVCC_FLAGS + extra + ["-DSPIM_SPLIT_BUILD", f"-DSPIM_INST_GROUP={g}", "-c", INST, "-o", o])
            objs.append(o)
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
        results = list(ex.map(_run, jobs))
    log = "".join(out for _, out in results)
    if any(rc != 0 for rc, _ in results):
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building " + OUT)
    if objs:
        # link under a temporary name and rename: a snapshot of the tree taken meanwhile never sees a half-written library
        tmp = OUT + ".tmp"
        rc, out = _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", tmp])
        log += out
        if rc != 0:
            sys.stderr.write(log)
            raise RuntimeError("nvcc failed linking " + OUT)
        os.replace(tmp, OUT)
    if verbose:
        sys.stderr.write(log)
    # the second name the reference's library picker pre-selects (EfficientBayesianBased.java:1127-1131)
    shutil.copyfile(OUT, ALIAS + ".tmp")
    os.replace(ALIAS + ".tmp", ALIAS)
    return OUT


def build_cufft_comparison() -> str:
    """profiles/microbench/cufft_conv: the cuFFT-based convolution bench.py times next to ours (comparison only -- it is a
    separate executable and nothing of it is linked into the library).  Returns "" when it cannot be built."""
    src = os.path.join(HERE, "..", "profiles", "microbench", "cufft_conv.cu")
    out = os.path.join(HERE, "..", "profiles", "microbench", "cufft_conv")
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    rc, log = _run([nvcc, "-O3", "-gencode", "arch=compute_100a,code=sm_100a", src, "-lcufft", "-o", out])
    if rc != 0:
        sys.stderr.write("cufft comparison binary not built:\n" + log)
        return ""
    return out


if __name__ == "__main__":
    print(build_cuda_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
