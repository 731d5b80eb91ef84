"""LOBPCG drivers, CPU side (no GPU): the host Rayleigh-Ritz solver against numpy / scipy, and the orchestration
template of csrc/lobpcg_core.hpp -- the same code the library runs with its device kernels -- instantiated with plain
host loops (tests/cpp/lobpcg_host_test.cc, test infrastructure) against analytic spectra.

The reference has no LOBPCG, so there is no reference vector to pin iteration counts to: what is
checked is the converged eigenpairs."""
import os
import subprocess

import numpy as np
import pytest

from dune_eigensolver_b200 import eigensolver as E, matrices as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "lobpcg_host_test.cc")
EXE = os.path.join(ROOT, "tests", "cpp", "lobpcg_host_test")
CORE = [os.path.join(ROOT, "dune_eigensolver_b200", "csrc", f) for f in ("lobpcg_core.hpp", "host_eig.hpp")]


def build_exe():
    if os.path.exists(EXE) and os.path.getmtime(EXE) > max(os.path.getmtime(f) for f in [SRC] + CORE):
        return EXE
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++17", "-O2", SRC, "-o", EXE], check=True, capture_output=True, text=True)
    return EXE


def run(*args):
    out = subprocess.run([build_exe(), *[str(a) for a in args]], capture_output=True, text=True, timeout=300)
    vals = {}
    for line in out.stdout.splitlines():
        k, _, rest = line.partition(" ")
        vals[k] = rest
    return out.returncode, vals, out.stdout


@pytest.mark.parametrize("n", [1, 2, 3, 8, 24, 96, 192])
def test_host_sym_eig_matches_numpy(n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    A = A + A.T
    w, V = E.host_sym_eig(A)
    assert np.abs(w - np.linalg.eigvalsh(A)).max() <= 1e-13 * max(1.0, np.abs(A).max() * n)
    assert np.abs(V.T @ V - np.eye(n)).max() <= 1e-13
    assert np.abs(A @ V - V * w).max() <= 1e-12 * max(1.0, np.abs(w).max())


def test_host_sym_eig_degenerate():
    for A in (np.eye(7), np.zeros((5, 5)), np.diag([3.0, 1, 2, 2, 2, 5])):
        w, V = E.host_sym_eig(A)
        assert np.allclose(w, np.sort(np.diag(A)))
        assert np.abs(V.T @ V - np.eye(len(w))).max() <= 1e-14


@pytest.mark.parametrize("n", [2, 16, 48, 96, 192])
def test_host_sym_gen_eig_matches_scipy(n):
    import scipy.linalg as sl

    rng = np.random.default_rng(100 + n)
    A = rng.standard_normal((n, n))
    A = A + A.T
    Bm = rng.standard_normal((n, n + 3))
    Bm = Bm @ Bm.T + 0.1 * np.eye(n)
    D = np.diag(10.0 ** rng.uniform(-6, 2, n))  # badly scaled basis vectors: the unit-diagonal scaling must cope
    Bm = D @ Bm @ D
    A = D @ A @ D
    w, C, piv = E.host_sym_gen_eig(A, Bm)
    wr = sl.eigh(A, Bm, eigvals_only=True)
    assert np.abs(w - wr).max() <= 1e-11 * np.abs(wr).max()
    assert np.abs(C.T @ Bm @ C - np.eye(n)).max() <= 1e-11
    assert 0.0 < piv <= 1.0


def test_host_sym_gen_eig_rejects_indefinite():
    GA = np.eye(4)
    GB = np.diag([1.0, 1.0, -1.0, 1.0])
    with pytest.raises(E.DeError) as e:
        E.host_sym_gen_eig(GA, GB)
    assert e.value.status == E.capi.DE_ERR_SINGULAR
    v = np.ones((4, 1))
    with pytest.raises(E.DeError):
        E.host_sym_gen_eig(GA, v @ v.T)  # rank one


@pytest.mark.parametrize("N,nev,tol", [(20, 8, 1e-8), (24, 32, 1e-10), (40, 24, 1e-6)])
def test_lobpcg_orchestration_standard_vs_analytic(N, nev, tol):
    """2D Dirichlet Laplacian of the reference (src/dune-eigensolver.cc:98-103) against its analytic spectrum
    (:437-446): eigenvalue error of a Ritz pair is <= residual^2 / gap, far below tol * lambda here."""
    rc, vals, text = run(N, nev, tol, 0, 0)
    assert rc == 0, text
    ev = np.array([float(x) for x in vals["eval"].split()])
    an = M.eigenvalues_laplace_dirichlet_2d(N)[:nev]
    assert np.abs(ev - an).max() <= 10 * tol * np.abs(an).max()
    assert float(vals["maxres"]) <= tol and float(vals["orth"]) <= 1e-12


def test_lobpcg_orchestration_largest():
    rc, vals, text = run(20, 8, 1e-8, 0, 1)
    assert rc == 0, text
    ev = np.array([float(x) for x in vals["eval"].split()])
    an = M.eigenvalues_laplace_dirichlet_2d(20)[::-1][:8]
    assert np.abs(ev - an).max() <= 1e-7


@pytest.mark.parametrize("N,nev,tol", [(16, 12, 1e-9), (30, 16, 1e-12)])
def test_lobpcg_orchestration_generalized(N, nev, tol):
    """A x = lambda B x with the 5-point Laplacian and an SPD matrix on the same pattern: both are polynomials in the
    1D second-difference matrices, so the pencil's spectrum is analytic."""
    rc, vals, text = run(N, nev, tol, 1, 0)
    assert rc == 0, text
    ev = np.array([float(x) for x in vals["eval"].split()])
    c = np.cos(np.pi * np.arange(1, N + 1) / (N + 1.0))
    lam = (4.0 - 2.0 * (c[:, None] + c[None, :])) / (4.0 + 1.0 * (c[:, None] + c[None, :]))
    an = np.sort(lam.reshape(-1))[:nev]
    assert np.abs(ev - an).max() <= 10 * tol * np.abs(an).max()
    assert float(vals["maxres"]) <= tol and float(vals["orth"]) <= 1e-12


def test_lobpcg_orchestration_restarts_when_basis_is_singular():
    """3 m = 192 columns in a 144-dimensional space: S^T B S is singular every iteration, the Rayleigh-Ritz step falls
    back to [X W] each time and the iteration still converges."""
    rc, vals, text = run(12, 64, 1e-8, 1, 0, 0, 1)  # Gram-Schmidt orthonormalisation of W
    assert rc == 0, text
    assert int(vals["restarts"]) >= int(vals["iterations"]) - 2
    assert float(vals["maxres"]) <= 1e-8


def test_lobpcg_orchestration_passes_rank_deficiency_through():
    """same problem with the CholQR2 of the device path (its pivot test included): 64 residual columns squeezed into
    the 80-dimensional complement of X are numerically dependent -- the error code of the orthonormalisation must
    come back unchanged (DE_ERR_SINGULAR = 5 on the device) instead of NaNs"""
    rc, vals, text = run(12, 64, 1e-8, 1, 0)
    assert rc == 1 and int(vals["rc"]) == 5, text


@pytest.mark.parametrize("generalized", [0, 1])
def test_lobpcg_orchestration_chebyshev_preconditioner(generalized):
    """the Chebyshev polynomial preconditioner (the drivers' default): same eigenpairs, several times fewer
    iterations and fewer applications of A in total"""
    runs = {}
    for deg in (0, 8):
        rc, vals, text = run(40, 12, 1e-9, generalized, 0, 0, 0, deg)
        assert rc == 0, text
        runs[deg] = (int(vals["iterations"]), int(vals["spmm"]), np.array([float(x) for x in vals["eval"].split()]))
        assert float(vals["maxres"]) <= 1e-9
    assert np.abs(runs[0][2] - runs[8][2]).max() <= 1e-12
    assert runs[8][0] * 3 <= runs[0][0] and runs[8][1] < runs[0][1]


def _dump_csr(A, path):
    rp, ci, v = A
    with open(path, "wb") as f:
        np.array([len(rp) - 1, len(ci)], dtype=np.int64).tofile(f)
        np.asarray(rp, dtype=np.int64).tofile(f)
        np.asarray(ci, dtype=np.int64).tofile(f)
        np.asarray(v, dtype=np.float64).tofile(f)


@pytest.mark.parametrize("contrast", [1e3, 1e6])
def test_lobpcg_orchestration_high_contrast(tmp_path, contrast):
    """configs[3]-type matrix (Q1 diffusion, kappa in {1, contrast} in a block pattern): the Jacobi-scaled Chebyshev
    preconditioner converges in a few dozen iterations where the plain iteration does not converge at all; the
    eigenvalues agree with a shift-invert Lanczos solve (scipy ARPACK)"""
    import scipy.sparse.linalg as spl

    shape = (16, 16, 16)
    A = M.q1_stiffness(shape, kappa=M.high_contrast_kappa(contrast, 8))
    path = str(tmp_path / "A.bin")
    _dump_csr(A, path)
    env = dict(os.environ, LOBPCG_TEST_A=path)
    out = subprocess.run([build_exe(), "16", "8", "1e-7", "0", "0", "0", "0", "8"], env=env, capture_output=True,
                         text=True, timeout=600)
    vals = dict(l.split(" ", 1) for l in out.stdout.splitlines() if " " in l)
    assert out.returncode == 0, out.stdout
    assert int(vals["iterations"]) <= 120
    ev = np.array([float(x) for x in vals["eval"].split()])
    ref = np.sort(spl.eigsh(M.to_scipy(A).tocsc(), k=8, sigma=0.0, which="LM", tol=1e-13)[0])
    assert np.abs(ev - ref).max() <= 1e-9 * np.abs(ref).max()
    assert float(vals["maxres"]) <= 1e-7


def test_host_sym_eig_clustered_and_graded():
    """multiple eigenvalues (2D Laplacian: lambda_ij = lambda_ji) and a strongly graded matrix: the QL iteration must
    keep the eigenvectors of a cluster orthonormal and the small eigenvalues of a graded matrix accurate"""
    A = M.to_scipy(M.laplacian_dirichlet_2d(9)).toarray()
    w, V = E.host_sym_eig(A)
    assert np.abs(w - M.eigenvalues_laplace_dirichlet_2d(9)).max() <= 1e-13
    assert np.abs(V.T @ V - np.eye(81)).max() <= 1e-13 and np.abs(A @ V - V * w).max() <= 1e-13
    rng = np.random.default_rng(5)
    Q, _ = np.linalg.qr(rng.standard_normal((40, 40)))
    lam = 10.0 ** np.linspace(-8, 4, 40)
    G = (Q * lam) @ Q.T
    w, V = E.host_sym_eig(G)
    assert np.abs(w - lam).max() <= 1e-12 * lam.max()
    assert np.abs(V.T @ V - np.eye(40)).max() <= 1e-13


def test_lobpcg_orchestration_maxiter_and_padding():
    """the loop falls through silently at maxiter like the reference's drivers (eigensolver.hh:191, :327): exactly
    maxiter basis updates, no error; maxiter = 0 returns the Rayleigh-Ritz vectors of the start block; nev = 3 runs a
    block of 8 and tests convergence on the first 3 only"""
    rc, vals, text = run(20, 8, 1e-12, 0, 0, 0, 0, 0, 0, 3)
    assert int(vals["rc"]) == 0 and int(vals["iterations"]) == 3 and int(vals["converged"]) == 0, text
    rc, vals, text = run(20, 8, 1e-12, 0, 0, 0, 0, 8, 0, 0)
    assert int(vals["rc"]) == 0 and int(vals["iterations"]) == 0 and float(vals["orth"]) <= 1e-13, text
    rc, vals, text = run(17, 3, 1e-8, 0, 0)
    assert rc == 0 and len(vals["eval"].split()) == 3, text
    ev = np.array([float(x) for x in vals["eval"].split()])
    assert np.abs(ev - M.eigenvalues_laplace_dirichlet_2d(17)[:3]).max() <= 1e-10


def _stencil(N, c, o):
    import scipy.sparse as sp

    T = sp.diags([np.ones(N - 1), np.ones(N - 1)], [-1, 1])
    I = sp.identity(N)
    return (c * sp.identity(N * N) + o * (sp.kron(I, T) + sp.kron(T, I))).tocsr()


@pytest.mark.parametrize("N,nev,tol,generalized,degree", [(20, 8, 1e-8, 0, 0), (16, 12, 1e-9, 1, 0), (40, 12, 1e-9, 0, 8),
                                                          (30, 13, 1e-9, 1, 8), (40, 22, 1e-7, 0, 4)])
def test_lobpcg_orchestration_matches_independent_oracle(N, nev, tol, generalized, degree):
    """oracle/lobpcg_oracle.py (numpy / LAPACK, no code shared with the product) and the product's orchestration take
    the same number of iterations (+-3: the two round differently) from the same start block and find the same
    eigenvalues. A wrong search direction or coefficient block would still converge -- but not in the same number of
    iterations. (Block boundaries that cut a multiple eigenvalue are avoided: there the count depends on rounding.)"""
    from oracle import lobpcg_oracle as LO

    m = E.padded_cols(nev)
    n = N * N
    # the test program fills its start block row-major from std::mt19937{123}: the same stream as one 8-wide panel
    X0 = E.from_panels(E.start_block(n * m // 8, 8, 123), n * m // 8, 8).reshape(n, m)
    A = _stencil(N, 4.0, -1.0)
    B = _stencil(N, 4.0, 0.5) if generalized else None
    theta, X, it, restarts, conv = LO.lobpcg(A, B, X0, nev, tol, 2000, degree)
    rc, vals, text = run(N, nev, tol, generalized, 0, 0, 0, degree)
    assert rc == 0 and conv, text
    assert abs(int(vals["iterations"]) - it) <= 3, (vals["iterations"], it)
    ev = np.array([float(x) for x in vals["eval"].split()])
    assert np.abs(ev - theta[:nev]).max() <= 1e-12


def test_lobpcg_orchestration_without_positive_diagonal():
    """a negative definite matrix has no Jacobi scale: the Chebyshev preconditioner is dropped and the plain iteration
    runs (the drivers' default degree is 8, so this is what a caller of StandardLOBPCG gets)"""
    out = subprocess.run([build_exe(), "12", "8", "1e-8", "0", "0", "1", "0", "8"], capture_output=True, text=True,
                         timeout=300, env=dict(os.environ, LOBPCG_TEST_NEGATE_A="1"))
    vals = dict(l.split(" ", 1) for l in out.stdout.splitlines() if " " in l)
    assert out.returncode == 0 and "running without the Chebyshev preconditioner" in out.stdout, out.stdout
    ev = np.array([float(x) for x in vals["eval"].split()])
    an = -M.eigenvalues_laplace_dirichlet_2d(12)[::-1][:8]
    assert np.abs(ev - an).max() <= 1e-10 * np.abs(an).max()
    assert int(vals["spmm"]) == 3 * int(vals["iterations"]) + 1  # A W, A X, A P per iteration: no Chebyshev products
