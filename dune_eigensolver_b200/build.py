"""Builds libdune_eigensolver_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

`python -m dune_eigensolver_b200.build` or `build_library()`; nvcc cross-compiles without a GPU.
Every translation unit is compiled to an object file (in parallel, rebuilt only when it or a header is newer) and
the objects are linked into one shared library.
The host compiler is pinned to /usr/bin/g++: this image exports CXX=/opt/gcc/bin/g++, a relocated compiler that
links libstdc++ statically and breaks shared objects loaded next to another libstdc++.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libdune_eigensolver_b200.so")
METIS = "/usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a"

SOURCES = ["de_runtime.cu", "de_spmm.cu", "de_dense.cu", "de_dense64.cu", "de_trsv.cu", "de_snode.cu", "de_drivers.cu",
           "de_multi.cu"]
HEADER_DIRS = [CSRC, os.path.join(HERE, "..", "include"), os.path.join(HERE, "..", "include", "dune", "eigensolver")]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _headers():
    out = [os.path.abspath(__file__)]
    for d in HEADER_DIRS:
        for f in os.listdir(d):
            if f.endswith((".cuh", ".hpp", ".hh", ".h")):
                out.append(os.path.join(d, f))
    return out


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _compile(src, obj, verbose):
    hostcxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [_nvcc(), "-std=c++17", "-O3", "-lineinfo",
           "-gencode", "arch=compute_100a,code=sm_100a",
           "-ccbin", hostcxx, "-diag-suppress", "177",
           "-Xcompiler", "-fPIC,-O3,-march=x86-64-v3,-Wno-sign-compare,-Wno-unused-function,-pthread",
           "-c", "-o", obj, src]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    if os.path.exists(METIS):
        cmd += ["-DDE_B200_HAVE_METIS"]
    return subprocess.run(cmd, capture_output=True, text=True)


def build_library(force=False, verbose=False):
    """Compile the shared library if it is missing or older than its sources. Returns its path."""
    os.makedirs(OBJ, exist_ok=True)
    hdr_time = max(os.path.getmtime(h) for h in _headers())
    jobs = []
    for s in _sources():
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            jobs.append((src, obj))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
            results = list(ex.map(lambda j: (j, _compile(j[0], j[1], verbose)), jobs))
        failed = False
        for (src, obj), res in results:
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                failed = True
                if os.path.exists(obj):
                    os.remove(obj)
            elif verbose:
                sys.stderr.write(res.stdout + res.stderr)
        if failed:
            raise RuntimeError("nvcc failed building " + LIB)
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in _sources()]
    if jobs or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs):
        hostcxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        cmd = [_nvcc(), "-shared", "-ccbin", hostcxx, "-o", LIB] + objs
        if os.path.exists(METIS):
            cmd += [METIS]
        cmd += ["-ldl", "-lpthread"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("linking failed: " + LIB)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
