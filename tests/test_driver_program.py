"""The C++ driver program (tools/driver/dune_eigensolver.cc): the reference's executable (src/dune-eigensolver.cc:448-787)
re-created on the drop-in headers -- ini file, command-line overrides, the three tests with their output tables, the
mgs performance line, the thread-replica harness and the new parallel.numgpus key."""
import os
import re
import subprocess

import numpy as np
import pytest

from dune_eigensolver_b200 import build as B, matrices as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "tools", "driver")
EXE = os.path.join(DRV, "dune_eigensolver")
INI = os.path.join(DRV, "dune-eigensolver.ini")


def build_exe():
    lib = B.build_library()
    srcs = [os.path.join(DRV, f) for f in ("dune_eigensolver.cc", "simple_bcrs.hh")]
    hdrs = [os.path.join(ROOT, "include", "dune", "eigensolver", f) for f in os.listdir(os.path.join(ROOT, "include", "dune", "eigensolver"))]
    if os.path.exists(EXE) and os.path.getmtime(EXE) > max(os.path.getmtime(p) for p in srcs + hdrs + [lib]):
        return EXE
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++17", "-O2", "-march=x86-64-v3", "-I", os.path.join(ROOT, "include"), "-I", DRV, srcs[0], "-o", EXE, lib,
           "-Wl,-rpath," + os.path.dirname(lib), "-lpthread"]
    if os.path.exists(B.METIS):
        cmd += ["-DDE_B200_HAVE_METIS", B.METIS]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return EXE


def run(*args):
    out = subprocess.run([build_exe(), "-ini", INI, *[str(a) for a in args]], capture_output=True, text=True, timeout=300)
    return out.returncode, out.stdout, out.stderr


def test_driver_program_builds_and_fails_loudly_without_gpu():
    assert os.path.exists(build_exe())
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        rc, out, err = run("-ev.N", 8)
        assert rc == 3 and "no CPU fallback" in err


def _evs(text):
    return np.array([float(line.split()[2]) for line in text.splitlines() if line.startswith("EV ")])


@pytest.mark.gpu
def test_driver_largest_table_and_overrides():
    """what the reference's shipped main() runs (largest_eigenvalues_convergence_test, :620-730)"""
    rc, out, err = run("-test", "largest", "-ev.N", 20, "-ev.m", 8, "-ev.tol", 1e-10)
    assert rc == 0, out + err
    an = M.eigenvalues_laplace_dirichlet_2d(20)[::-1][:8]
    assert np.abs(_evs(out) - an).max() < 2e-8
    assert "N_M_TOL_ESARERROR_ARPERROR_ESANERROR_TIMERATIO_ARPACKITER" in out
    row = [ln for ln in out.splitlines() if " & " in ln][-1].split("&")
    assert int(row[0]) == 400 and int(row[1]) == 8 and float(row[2]) == 1e-10 and float(row[3]) < 2e-8


@pytest.mark.gpu
def test_driver_eigenvalues_raes_and_summary_line(golden):
    """eigenvalues_test with ev.method = raes (:448-525): GeneralizedInverse on (Neumann, PU-masked) -- the golden run of
    SURVEY.md §4 (N = 16, overlap 3, shift 1e-3, 8 pairs, tol 1e-12: 121 iterations)"""
    rc, out, err = run("-test", "eigenvalues", "-ev.N", 16, "-ev.m", 8, "-ev.tol", 1e-12, "-ev.verbose", 1)
    assert rc == 0, out + err
    assert np.abs(_evs(out) - golden["d_geninv_eval"]).max() <= 1e-10 * np.abs(golden["d_geninv_eval"]).max()
    m = re.search(r"GeneralizedInverse:\s+time_total=\S+ time_factorization=\S+ iterations=(\d+) relerror=\S+", out)
    assert m and abs(int(m.group(1)) - 121) <= 1
    rc, out, err = run("-test", "smallest", "-ev.N", 16, "-ev.m", 8, "-ev.tol", 1e-6)
    assert rc == 0 and "N_M_TOL_RASERROR_ARPERROR_TIMERATIO_ARPACKITER" in out
    row = [ln for ln in out.splitlines() if " & " in ln][-1].split("&")
    assert float(row[3]) < 1e-4  # error of the tol = 1e-6 run against the tol = 1e-13 run


@pytest.mark.gpu
def test_driver_lobpcg_method_mgs_line_and_replicas():
    rc, out, err = run("-test", "eigenvalues", "-ev.N", 16, "-ev.m", 8, "-ev.tol", 1e-9, "-ev.method", "lobpcg")
    assert rc == 0, out + err
    dense = np.linalg.eigvalsh(M.to_scipy(M.laplacian_neumann_2d(16)).toarray())[:8]
    assert np.abs(_evs(out) - dense).max() < 1e-8
    rc, out, err = run("-test", "mgs", "-mgs.n", 20000, "-mgs.m", 24, "-mgs.n_iter", 20)
    assert rc == 0, out + err
    t = [ln for ln in out.splitlines() if ln.startswith("P_n_m_i_iblocked_perfn_perfb_perfv")][0].split()
    assert t[1:4] == ["1", "20000", "24"] and float(t[7]) > 0.0
    # the replica harness: two host threads, each with its own GPU context, behind the barrier (:756-773)
    rc, out, err = run("-test", "largest", "-ev.N", 16, "-ev.m", 8, "-ev.tol", 1e-8, "-parallel.numthreads", 2)
    assert rc == 0, out + err
    assert np.abs(_evs(out) - M.eigenvalues_laplace_dirichlet_2d(16)[::-1][:8]).max() < 1e-6


@pytest.mark.gpu
def test_driver_numgpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("parallel.numgpus = 2 needs two GPUs")
    rc, out, err = run("-test", "largest", "-ev.N", 24, "-ev.m", 8, "-ev.tol", 1e-10, "-parallel.numgpus", 2)
    assert rc == 0, out + err
    assert np.abs(_evs(out) - M.eigenvalues_laplace_dirichlet_2d(24)[::-1][:8]).max() < 2e-8
