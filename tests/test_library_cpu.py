"""CPU-side checks of the product library (no GPU): the C-ABI library loads, exports every symbol the header
declares, its host-only entry points are right, and compute entry points FAIL (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from dune_eigensolver_b200 import capi, eigensolver as E, matrices as M, parallel as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu():
    try:
        import torch

        return not torch.cuda.is_available()
    except Exception:
        return True


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "dune_eigensolver_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char \*)\s*(de_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert len(declared) >= 40
    L = C.CDLL(capi._build.build_library())
    for name in sorted(declared):
        assert hasattr(L, name), "library does not export " + name
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)
    assert capi.lib().de_version() >= 100


def test_start_block_is_the_reference_stream(oracle, golden):
    n, m = int(golden["k_N"]) ** 2, int(golden["k_m"])
    mine = E.from_panels(E.start_block(n, m, 123), n, m)
    assert np.array_equal(mine, golden["k_X"])
    assert np.array_equal(mine, oracle.start_block(n, m, 123))


def test_layout_helpers_roundtrip():
    X = np.arange(24 * 16, dtype=float).reshape(24, 16)
    p = E.to_panels(X)
    # reference indexing formula, multivector.hh:130-133
    for (i, j) in [(0, 0), (5, 7), (5, 8), (23, 15)]:
        assert p[((j // 8) * 24 + i) * 8 + j % 8] == X[i, j]
    assert np.array_equal(E.from_panels(p, 24, 16), X)
    with pytest.raises(capi.DeError, match="multiple of block size"):
        E.to_panels(np.zeros((4, 12)))
    assert [E.padded_cols(k) for k in (1, 8, 9, 16, 17)] == [8, 8, 16, 16, 24]


@pytest.mark.parametrize("ordering", [0, 1, 2])
@pytest.mark.parametrize("scale", [False, True])
def test_host_factorization_contract(oracle, ordering, scale):
    """the provider fills the UMFPACK field contract (umfpacktools.hh:26-44): check the layout promises and that the
    REFERENCE apply inverts A with these factors."""
    N = 9
    rp, ci, v = M.laplacian_neumann_2d(N)
    v = v.copy()
    E._add_to_diagonal(rp, ci, v, 0.25)
    hf = E.HostFactorization((rp, ci, v), ordering, scale)
    F = hf.arrays()
    n = N * N
    assert hf.n == n and len(F["Lp"]) == n + 1 and F["Lp"][-1] == hf.lnz and F["Up"][-1] == hf.unz
    for i in range(n):
        row = F["Lj"][F["Lp"][i]:F["Lp"][i + 1]]
        assert row[-1] == i and np.all(np.diff(row) > 0) and F["Lx"][F["Lp"][i + 1] - 1] == 1.0
        colm = F["Ui"][F["Up"][i]:F["Up"][i + 1]]
        assert colm[-1] == i and np.all(np.diff(colm) > 0)
    assert sorted(F["P"]) == list(range(n)) and sorted(F["Q"]) == list(range(n))
    assert (F["do_recip"] == 0) == scale
    A = M.to_scipy((rp, ci, v))
    X = np.random.default_rng(1).standard_normal((n, 8))
    sol, _ = oracle.factor_apply(F, A @ X)
    np.testing.assert_allclose(sol, X, rtol=0, atol=1e-11)


def test_host_factorization_reports_singular():
    rp, ci, v = M.laplacian_neumann_2d(6)  # singular: constants are in the kernel
    with pytest.raises(capi.DeError, match="singular") as ei:
        E.HostFactorization((rp, ci, v), 0)
    assert ei.value.status == capi.DE_ERR_SINGULAR


def test_nested_dissection_reduces_fill():
    A = M.laplacian_dirichlet_2d(40)
    nat, nd = E.HostFactorization(A, 0), E.HostFactorization(A, 1)
    assert nd.lnz < 0.6 * nat.lnz


def test_halo_plan_single_process():
    N = 6
    rp, ci, v = M.laplacian_dirichlet_2d(N)
    part = P.partition_rows(N * N, 3, align=N)
    assert list(part) == [0, 12, 24, 36]
    s, e = part[1], part[2]
    lrp = rp[s:e + 1] - rp[s]
    lci = ci[rp[s]:rp[e]]
    col_local, halo, recv = P.halo_plan_local(lrp, lci, part, 1)
    assert list(recv) == [N, 0, N] and len(halo) == 2 * N
    assert list(halo) == list(range(6, 12)) + list(range(24, 30))
    owned = (lci >= s) & (lci < e)
    assert np.array_equal(col_local[owned], lci[owned] - s)
    assert np.array_equal(halo[col_local[~owned] - (e - s)], lci[~owned])
    with pytest.raises(capi.DeError):
        P.halo_plan_local(lrp, lci, np.array([0, 10, 24, 36]), 1)  # partition does not match n_owned


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(capi.DeError) as ei:
        E.Context(0)
    assert ei.value.status == capi.DE_ERR_CUDA and "no CPU fallback" in str(ei.value)


def test_generators_row_ranges_and_symmetry():
    for gen, args in [(M.laplacian_fd, ((5, 4, 3),)), (M.q1_stiffness, ((4, 3, 5),)), (M.q1_mass, ((4, 5),))]:
        full = gen(*args)
        n = len(full[0]) - 1
        a, b = n // 3, 2 * n // 3 + 1
        part = gen(*args, rows=(a, b))
        s, e = full[0][a], full[0][b]
        assert np.array_equal(part[1], full[1][s:e]) and np.array_equal(part[2], full[2][s:e])
        D = M.to_scipy(full).toarray()
        assert np.abs(D - D.T).max() == 0.0
    K = M.to_scipy(M.q1_stiffness((4, 4, 4))).toarray()
    Mm = M.to_scipy(M.q1_mass((4, 4, 4))).toarray()
    import scipy.linalg as sl

    np.testing.assert_allclose(np.linalg.eigvalsh(K), M.eigenvalues_q1_stiffness((4, 4, 4)), atol=1e-12)
    np.testing.assert_allclose(sl.eigh(K, Mm, eigvals_only=True), M.eigenvalues_q1_pencil((4, 4, 4)), atol=1e-11)
    np.testing.assert_allclose(np.linalg.eigvalsh(M.to_scipy(M.laplacian_fd((5, 4, 3))).toarray()),
                               M.eigenvalues_laplacian_fd((5, 4, 3)), atol=1e-12)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path: one replica of the reference solve per host core) prints
    ONE JSON line with the contract's keys; run here on a small grid with a bounded sample"""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid", "12", "--steps", "1",
                          "--warmup", "0", "--cpu-sample-iters", "2"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "time-to-m-eigenpairs" and d["unit"] == "s"
    assert d["higher_is_better"] is False and d["value"] > 0
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"] == {"value": d["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_lobpcg_legs_cannot_break_the_main_line():
    """the LOBPCG legs of bench.py run in child processes: whatever happens there (here: no GPU at all) comes back as an
    {"error": ...} object instead of an exception"""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    res = mod.lobpcg_leg(["--grid", "6", "--nev", "4", "--steps", "1"], 120)
    assert isinstance(res, dict)
    if not has_gpu:
        assert "no CPU fallback" in res["error"]
    assert "error" in mod.lobpcg_leg(["--no-such-option"], 60)


def test_c_abi_from_plain_c99(oracle):
    """include/dune_eigensolver_b200.h is a C header: compiled with gcc -std=c99 -pedantic -Werror into a program
    without any C++, linked against the library; host-only entry points work, a context needs a device"""
    import subprocess

    from dune_eigensolver_b200 import build as B

    lib = B.build_library()
    exe = os.path.join(ROOT, "tests", "c", "cabi_smoke")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "cabi_smoke.c"), "-o", exe, lib, "-Wl,-rpath," + os.path.dirname(lib)],
                   check=True, capture_output=True, text=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    f = out.stdout.split()
    assert abs(float(f[3]) - 1.0) <= 1e-14 and abs(float(f[4]) - 3.0) <= 1e-14
    assert float(f[6]) == float(np.asarray(oracle.start_block(2, 8, 123)).reshape(-1)[0])  # the reference's stream
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        assert f[8] == "3" and "no CPU fallback" in out.stdout


def test_lu_provider_rejects_what_static_pivoting_cannot_factor():
    """the LU provider has no numerical pivoting; an indefinite matrix (shift inside the spectrum) must
    raise instead of silently returning an inaccurate factor. [[eps, 1], [1, eps]]-like 2 x 2 blocks need pivoting."""
    import scipy.sparse as sp

    from dune_eigensolver_b200 import eigensolver as E

    n = 40
    blocks = [sp.coo_array(np.array([[1e-14, 1.0], [1.0, 1e-14]]))] * (n // 2)
    S = sp.block_diag(blocks, format="csr") + 1e-3 * sp.diags([np.ones(n - 1), np.ones(n - 1)], [-1, 1], format="csr")
    S = S.tocsr()
    S.sort_indices()
    with pytest.raises(E.DeError) as e:
        E.HostFactorization((S.indptr, S.indices, S.data), ordering=0)
    assert "backward-error" in str(e.value) or "singular" in str(e.value)
    # a well-posed shifted Laplacian still factors, and its backward error is at round-off
    A = M.laplacian_dirichlet_2d(12)
    hF = E.HostFactorization(A)
    assert hF.lnz > 0
    hF.close()
