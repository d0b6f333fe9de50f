"""Builds gpc_b200/libgpc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m gpc_b200.build [--force] [--verbose]

The library links the static CUDA runtime so it can be loaded next to PyTorch (or without it, from C++).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgpc_b200.so")
SHIM = os.path.join(HERE, "libgpc_lapack_shim.so")
SOURCES = ["dense.cu", "ozaki.cu", "gpkern.cu", "api.cu", "lapack_api.cu", "host.cu", "modelio.cu", "dist.cu", "sparse.cu", "smpart.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(os.path.dirname(HERE), "include", "gpc_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + HEADERS):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            print(" ".join(cmd))
            print(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if force or procs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl", "-lpthread"]
        subprocess.check_call(cmd)
    # the Fortran-ABI shim (dpotrf_, dpotri_, dtrsm_, dsyrk_, dgemm_ -> gpc_d*): link or LD_PRELOAD it in front of a
    # BLAS and the unmodified reference objects run those calls on the GPU (level 0)
    shim_src = os.path.join(CSRC, "lapack_shim.cpp")
    if force or _stale(SHIM, [shim_src, LIB] + HEADERS):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-o", SHIM, shim_src,
                               "-L" + HERE, "-lgpc_b200", "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
