"""The C++ drop-in headers (include/dune/eigensolver/*.hh) compile as a user's translation unit against a
BCRSMatrix-like type, and (on the GPU) reproduce the oracle's eigenvalues through the reference's own signatures."""
import os
import subprocess

import numpy as np
import pytest

from dune_eigensolver_b200 import build as B, matrices as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "dropin_test")


def build_exe():
    lib = B.build_library()
    src = os.path.join(ROOT, "tests", "cpp", "dropin_main.cc")
    if os.path.exists(EXE) and os.path.getmtime(EXE) > max(os.path.getmtime(src), os.path.getmtime(lib)):
        return EXE
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++17", "-O2", "-march=x86-64-v3", "-I", os.path.join(ROOT, "include"), "-I",
           os.path.join(ROOT, "oracle", "shim"), src, "-o", EXE, lib, "-Wl,-rpath," + os.path.dirname(lib)]
    metis = B.METIS
    if os.path.exists(metis):
        cmd += ["-DDE_B200_HAVE_METIS", metis]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return EXE


def run(*args):
    out = subprocess.run([build_exe(), *[str(a) for a in args]], capture_output=True, text=True, timeout=300)
    vals = {}
    for line in out.stdout.splitlines():
        k, _, rest = line.partition(" ")
        vals.setdefault(k, rest)
    return out.returncode, vals, out.stdout


def test_dropin_headers_compile_and_fail_loudly_without_gpu():
    exe = build_exe()
    assert os.path.exists(exe)
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        rc, vals, text = run("largest", 8, 8)
        assert rc == 3 and "no CPU fallback" in text


@pytest.mark.gpu
def test_dropin_drivers_match_oracle(oracle):
    rc, vals, text = run("largest", 20, 8, 1e-10)
    assert rc == 0, text
    ev = np.array([float(x) for x in vals["eval"].split()])
    ref, V, k = oracle.standard_largest(M.laplacian_dirichlet_2d(20), 0.0, 1e-10, 4000, 8)
    assert np.abs(ev - ref).max() <= 1e-10 * np.abs(ref).max()

    rc, vals, text = run("inverse", 20, 8, 1e-10)
    assert rc == 0, text
    ev = np.array([float(x) for x in vals["eval"].split()])
    ref, V, k = oracle.standard_inverse(M.laplacian_dirichlet_2d(20), 1e-3, 1e-10, 4000, 8)
    assert np.abs(ev - ref).max() <= 1e-10 * np.abs(ref).max()
    assert abs(float(vals["diag0"]) - 4.001) < 1e-15  # caller's matrix shifted in place

    rc, vals, text = run("generalized", 16, 8, 1e-12)
    assert rc == 0, text
    ev = np.array([float(x) for x in vals["eval"].split()])
    ref, V, it = oracle.generalized_inverse(M.laplacian_neumann_2d(16), M.laplacian_B_2d(16, 3), 1e-3, 0.0, 1e-12, 4000, 8)
    assert np.abs(ev - ref).max() <= 1e-10 * np.abs(ref).max()
    assert "iterations=%d" % it in text or "iterations=%d" % (it + 1) in text or "iterations=%d" % (it - 1) in text
    # the reference's machine-greppable summary line, all five fields in its order (eigensolver.hh:343-350)
    import re

    assert re.search(r"GeneralizedInverse:\s+time_total=\S+ time_factorization=\S+ iterations=\d+ relerror=\S+", text)


@pytest.mark.gpu
def test_dropin_kernels(oracle):
    rc, vals, text = run("kernels", 12, 8)
    assert rc == 0, text
    A = M.laplacian_dirichlet_2d(12)
    X = oracle.start_block(144, 16, 123)
    ref = oracle.diag_dot(X, oracle.spmm(A, X))
    dp = np.array([float(x) for x in vals["diagdot"].split()])
    assert np.abs(dp - ref).max() <= 1e-12 * np.abs(ref).max()
    assert float(vals["ortho_defect"]) < 1e-13
    assert "number of cols must be a multiple of block size" in vals["caught"]


@pytest.mark.gpu
def test_dropin_multi_gpu_from_cpp(oracle):
    """`StandardLargest(A, ...)` of the C++ drop-in header on several ranks without Python / MPI / NCCL (C ABI
    de_multi_*). On a one-GPU box the ranks share ordinal 0."""
    import torch

    have = max(torch.cuda.device_count(), 1)
    for ranks in (2, 4):
        devs = ",".join(str(r % have) for r in range(ranks))
        rc, vals, text = run("largest", 24, 8, 1e-10, devs)
        assert rc == 0, text
        assert vals["gpus"] == str(ranks)
        ev = np.array([float(x) for x in vals["eval"].split()])
        ref, V, k = oracle.standard_largest(M.laplacian_dirichlet_2d(24), 0.0, 1e-10, 4000, 8)
        assert np.abs(ev - ref).max() <= 1e-10 * np.abs(ref).max()
        if have >= ranks:  # LOBPCG needs one GPU per rank (tests/test_multi_inprocess_gpu.py::test_partitioned_lobpcg)
            rc, vals, text = run("lobpcg", 24, 8, 1e-9, devs)
            assert rc == 0, text
            ev = np.array([float(x) for x in vals["eval"].split()])
            assert np.abs(ev - M.eigenvalues_laplace_dirichlet_2d(24)[:8]).max() <= 1e-9 * 8.0


def test_dropin_cost_models_match_the_reference(oracles, golden):
    """flops_orthonormalize / bytes_orthonormalize_{naive,blocked} of the PRODUCT header (include/dune/eigensolver/
    kernels_b200.hh) against the golden values generated from the reference (kernels_cpp.hh:98-116, :157-175) and, on
    more shapes, against the reference itself (oracle/_ref) or the port pinned to it."""
    shapes = [(1000, 24, 8), (40000, 16, 8), (1000000, 32, 8), (2097152, 64, 8), (17, 8, 8), (123, 40, 8)]
    out = subprocess.run([build_exe(), "costmodel", *[str(x) for sh in shapes for x in sh]], capture_output=True, text=True,
                         timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    got = {}
    for line in out.stdout.splitlines():
        t = line.split()
        if t and t[0] == "cost":
            got[(int(t[1]), int(t[2]), int(t[3]))] = tuple(float(x) for x in t[4:7])
    assert set(got) == set(shapes)
    assert got[(1000, 24, 8)] == (float(golden["k_cost_flops"]), float(golden["k_cost_bytes_naive"]),
                                  float(golden["k_cost_bytes_blocked"]))
    ref = oracles[0] if oracles[0] is not None else oracles[1]
    for (n, m, b), (fl, bn, bb) in got.items():
        assert fl == ref.flops_orthonormalize(n, m)
        assert bn == ref.bytes_orthonormalize_naive(n, m)
        assert bb == ref.bytes_orthonormalize_blocked(n, m, b)


def _block_laplacian(N, k):
    """kron(L, C) as scalar CSR and as (rowptr, col, blocks): the matrix tests/cpp/dropin_main.cc builds for bcsr<k>"""
    import scipy.sparse as sp

    L = M.to_scipy(M.laplacian_dirichlet_2d(N)).tocsr()
    Cm = np.array([[2.0 + r if r == c else (0.5 if abs(r - c) == 1 else 0.0) for c in range(k)] for r in range(k)])
    S = sp.kron(L, Cm, format="csr")
    S.sort_indices()
    blocks = L.data[:, None, None] * Cm[None, :, :]
    return (S.indptr.astype(np.int64), S.indices.astype(np.int64), S.data.copy()), (L.indptr, L.indices, blocks)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [2, 3])
def test_dropin_bcsr_blocks(oracle, k):
    """BCRSMatrix<FieldMatrix<double,k,k>>, k > 1 (every reference kernel throws for it,
    kernels_cpp.hh:632-633): the drop-in runs it as the scalar matrix the blocks denote, so the reference's own driver on
    that scalar matrix is the oracle."""
    N, nev = 14, 8
    rc, vals, text = run("bcsr%d" % k, N, nev, 1e-10)
    assert rc == 0, text
    scalar, _ = _block_laplacian(N, k)
    ev, V, it = oracle.standard_largest(scalar, 0.0, 1e-10, 4000, nev)
    got = np.array([float(x) for x in vals["eval"].split()])
    assert np.abs(got - ev).max() <= 1e-10 * np.abs(ev).max()
    blk = np.array([float(x) for x in vals["evec_block1"].split()]).reshape(nev, k)
    assert np.abs(np.abs(blk) - np.abs(V[:, k:2 * k])).max() <= 1e-7
    X = oracle.start_block(N * N * k, 8, 7)
    ref = oracle.spmm(scalar, X)
    row = np.array([float(x) for x in vals["spmm_row"].split()])
    assert np.abs(row - ref[k + 1]).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [2, 4])
def test_bcsr_matrix_through_the_c_abi(ctx, oracle, k):
    from dune_eigensolver_b200 import eigensolver as E

    N, m = 20, 16
    scalar, (rp, ci, blocks) = _block_laplacian(N, k)
    dA = E.Matrix.bcsr(ctx, rp, ci, blocks)
    assert dA.n == N * N * k
    X = oracle.start_block(N * N * k, m, 123)
    dX, dY = E.MultiVector.from_array(ctx, X), E.MultiVector(ctx, N * N * k, m)
    E.matmul_sparse_tallskinny_blocked(dY, dA, dX)
    ref = oracle.spmm(scalar, X)
    assert np.abs(dY.download() - ref).max() <= 1e-13 * np.abs(ref).max()
    for h in (dA, dX, dY):
        h.close()


@pytest.mark.gpu
def test_dropin_cholesky_provider():
    """UMFPackFactorizedMatrix with Provider::cholesky (explicit) and Provider::automatic (n >= 10 000, symmetric: N = 110)"""
    rc, vals, text = run("cholesky", 40, 8, 1e-10)
    assert rc == 0, text
    assert vals["supernodal"] == "1" and vals["contract"].split()[0] == "1" and int(vals["contract"].split()[1]) == 1600
    assert float(vals["solve_error"]) < 1e-11
    ev = np.array([float(x) for x in vals["eval"].split()])
    assert np.abs(ev - M.eigenvalues_laplace_dirichlet_2d(40)[:8]).max() < 1e-9
    rc, vals, text = run("inverse", 110, 8, 1e-9)  # automatic (n = 12 100 >= 10 000): supernodal Cholesky of the shifted Laplacian
    assert rc == 0, text
    ev = np.array([float(x) for x in vals["eval"].split()])
    assert np.abs(ev - M.eigenvalues_laplace_dirichlet_2d(110)[:8]).max() < 1e-8
