"""Build the in-tree CUDA library (libv2e2v_b200.so) for sm_100a with nvcc.

    python -m v2e2v_b200.build [--force] [--verbose]

No torch headers are involved (the library is plain C ABI), so a full build takes seconds.
The .so is written next to this file: it is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "libv2e2v_b200.so")
BUILD = os.path.join(PKG, "_build")

CU_SOURCES = ["v2e_api.cu", "v2e_stepwise.cu", "v2e_persistent.cu", "v2e_runahead.cu", "v2e_fixpoint.cu", "v2e_fixpoint_spec.cu",
              "v2e_voxel.cu"]
C_SOURCES = ["host_scalars.c"]
HEADERS = ["v2e_device.cuh", "v2e_engine.cuh", "v2e_fixpoint.cuh", os.path.join("..", "..", "include", "v2e2v_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",              # belt and braces: exact sections also use __f*_rn intrinsics
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _gcc() -> str:
    for c in ("/usr/bin/gcc", shutil.which("gcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("gcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs = []
    log = []
    jobs = []
    for src in CU_SOURCES:
        s, o = os.path.join(CSRC, src), os.path.join(BUILD, src + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append((src, s, o))

    def _compile(job):
        src, s, o = job
        extra = os.environ.get("V2E_NVCC_EXTRA", "").split()   # development aid
        r = subprocess.run([_nvcc(), *NVCC_FLAGS, *extra, "-c", s, "-o", o],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stderr}")
        return r.stderr

    if jobs:   # the translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            log.extend(ex.map(_compile, jobs))
    for src in C_SOURCES:
        s, o = os.path.join(CSRC, src), os.path.join(BUILD, src + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            r = subprocess.run([_gcc(), "-O2", "-fPIC", "-std=c11", "-ffp-contract=off",
                                "-fno-fast-math", "-c", s, "-o", o], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"gcc failed for {src}:\n{r.stderr}")
    if force or _stale(OUT, objs):
        r = subprocess.run([_nvcc(), "-shared", "-o", OUT, *objs, "-gencode",
                            "arch=compute_100a,code=sm_100a", "-cudart", "static"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stderr}")
    if verbose:
        print("\n".join(log))
    with open(os.path.join(BUILD, "ptxas.log"), "a" if not force else "w") as f:
        f.write("\n".join(log))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
