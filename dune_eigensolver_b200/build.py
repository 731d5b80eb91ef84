"""Builds libdune_eigensolver_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

`python -m dune_eigensolver_b200.build` or `build_library()`; nvcc cross-compiles without a GPU.
The host compiler is pinned to /usr/bin/g++: this image exports CXX=/opt/gcc/bin/g++, a relocated compiler that
links libstdc++ statically and breaks shared objects loaded next to another libstdc++.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libdune_eigensolver_b200.so")
METIS = "/usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a"

SOURCES = ["de_capi.cu"]
HEADERS = ["kernels_sparse.cuh", "kernels_dense.cuh", "kernels_trsv.cuh", "kernels_tallskinny.cuh",
           "kernels_spmm_blocked.cuh", "brb_format.hpp", "kernels_tallskinny2.cuh", "kernels_tail.cuh", "kernels_peer.cuh",
           "kernels_brb_build.cuh", "kernels_lobpcg.cuh", "lobpcg_core.hpp", "host_eig.hpp",
           os.path.join("..", "..", "include", "dune_eigensolver_b200.h"),
           os.path.join("..", "..", "include", "dune", "eigensolver", "sparse_lu.hh")]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force=False, verbose=False):
    """Compile the shared library if it is missing or older than its sources. Returns its path."""
    if not force and not _stale():
        return LIB
    hostcxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [_nvcc(), "-std=c++17", "-O3", "-lineinfo",
           "-gencode", "arch=compute_100a,code=sm_100a",
           "-ccbin", hostcxx,
           "-Xcompiler", "-fPIC,-O3,-march=x86-64-v3,-Wno-sign-compare,-pthread",
           "-shared", "-o", LIB]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    if os.path.exists(METIS):
        cmd += ["-DDE_B200_HAVE_METIS", METIS]
    cmd += ["-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building " + LIB)
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
