"""Pins the oracle (CPU, no GPU): the port against the reference compiled verbatim, both against the committed
golden vectors and against the analytic spectra the reference's own test uses
(reference src/dune-eigensolver.cc:437-446)."""
import numpy as np

from dune_eigensolver_b200 import matrices as M


def both(oracles):
    return [o for o in oracles if o is not None]


def test_reference_build_is_present_where_the_mount_exists(oracles):
    import os

    if os.path.isdir("/root/reference"):
        assert oracles[0] is not None and oracles[0].kind == "reference"
    assert oracles[1].kind == "port"


def test_kernels_match_golden(oracles, golden):
    N, m = int(golden["k_N"]), int(golden["k_m"])
    A, B = M.laplacian_dirichlet_2d(N), M.laplacian_B_2d(N, 1)
    X, Y = golden["k_X"], golden["k_Y"]
    for o in both(oracles):
        assert np.array_equal(o.start_block(N * N, m, 123), X)  # libstdc++ stream is the same on this toolchain
        np.testing.assert_allclose(o.spmm(A, X), golden["k_spmm"], rtol=0, atol=1e-14)
        np.testing.assert_allclose(o.diag_dot(X, Y), golden["k_diag_dot"], rtol=1e-15)
        np.testing.assert_allclose(o.gram(X, Y), golden["k_gram"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(o.orthonormalize(X), golden["k_ortho"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(o.orthonormalize_naive(X), golden["k_ortho_naive"], rtol=0, atol=1e-13)
        q, nrm = o.b_orthonormalize(B, X)
        np.testing.assert_allclose(q, golden["k_bortho"], rtol=0, atol=1e-12)
        assert abs(nrm - float(golden["k_bortho_norm"])) < 1e-10
        F = {k[2:]: golden[k] for k in golden.files if k.startswith("f_") and k != "f_apply"}
        sol, _ = o.factor_apply(F, X)
        np.testing.assert_allclose(sol, golden["f_apply"], rtol=0, atol=1e-13)
        assert o.flops_orthonormalize(1000, 24) == float(golden["k_cost_flops"])
        assert o.bytes_orthonormalize_naive(1000, 24) == float(golden["k_cost_bytes_naive"])
        assert o.bytes_orthonormalize_blocked(1000, 24, 8) == float(golden["k_cost_bytes_blocked"])


def test_factor_apply_inverts(oracles, golden):
    """apply pinned by the identity F^-1 (A X) = X on the exported factors of A + 0.5 I."""
    N = int(golden["k_N"])
    A = M.to_scipy(M.laplacian_dirichlet_2d(N)).toarray() + 0.5 * np.eye(N * N)
    F = {k[2:]: golden[k] for k in golden.files if k.startswith("f_") and k != "f_apply"}
    X = golden["k_X"]
    for o in both(oracles):
        sol, _ = o.factor_apply(F, A @ X)
        np.testing.assert_allclose(sol, X, rtol=0, atol=1e-12)


def test_drivers_match_golden(oracles, golden):
    for o in both(oracles):
        ev, V, k = o.standard_largest(M.laplacian_dirichlet_2d(20), 0.0, 1e-10, 4000, 8)
        assert k == int(golden["d_largest_iter"]) == 1456
        np.testing.assert_allclose(ev, golden["d_largest_eval"], rtol=1e-12)
        ev, V, it = o.generalized_inverse(M.laplacian_neumann_2d(16), M.laplacian_B_2d(16, 3), 1e-3, 0.0, 1e-12, 4000, 8)
        assert it == int(golden["d_geninv_iter"]) == 121
        np.testing.assert_allclose(ev, golden["d_geninv_eval"], rtol=0, atol=1e-12)
        ev, V, k = o.standard_inverse(M.laplacian_dirichlet_2d(20), 1e-3, 1e-10, 4000, 8)
        assert k == int(golden["d_inverse_iter"])
        np.testing.assert_allclose(ev, golden["d_inverse_eval"], rtol=1e-12)


def test_shipped_ini_configuration(oracles, golden):
    """reference src/dune-eigensolver.ini (tol 2e-3, shift 1e-3, overlap 3) with ev.m = 16 at N = 40."""
    for o in both(oracles):
        ev, V, it = o.generalized_inverse(M.laplacian_neumann_2d(40), M.laplacian_B_2d(40, 3), 1e-3, 0.0, 2e-3, 4000, 16)
        assert it == int(golden["d_ini_iter"])
        np.testing.assert_allclose(ev, golden["d_ini_eval"], rtol=0, atol=1e-12)
        ev, V, k = o.standard_largest(M.laplacian_dirichlet_2d(40), 0.0, 2e-3, 4000, 16)
        assert k == int(golden["d_ini_largest_iter"])
        np.testing.assert_allclose(ev, golden["d_ini_largest_eval"], rtol=1e-12)


def test_analytic_spectrum_largest(oracle, golden):
    """the reference's known answer: top eigenvalues of the 2D 5-point Laplacian (src/dune-eigensolver.cc:437-446)."""
    an = M.eigenvalues_laplace_dirichlet_2d(20)[::-1][:8]
    assert np.abs(golden["d_largest_eval"] - an).max() < 2e-8
    ev, V, k = oracle.standard_largest(M.laplacian_dirichlet_2d(12), 0.0, 1e-12, 4000, 8)
    an = M.eigenvalues_laplace_dirichlet_2d(12)[::-1][:8]
    assert np.abs(ev - an).max() < 1e-9


def test_analytic_spectrum_inverse(oracle):
    N = 12
    ev, V, k = oracle.standard_inverse(M.laplacian_dirichlet_2d(N), 1e-3, 1e-12, 4000, 8)
    an = M.eigenvalues_laplace_dirichlet_2d(N)[:8]
    assert np.abs(ev - an).max() < 1e-9
    A = M.to_scipy(M.laplacian_dirichlet_2d(N))
    for j in range(6):  # residuals of well-separated or degenerate pairs alike
        r = A @ V[j] - ev[j] * V[j]
        assert np.linalg.norm(r) < 1e-5


def test_generalized_against_dense_solve(oracle, golden):
    import scipy.linalg as sl

    A = M.to_scipy(M.laplacian_neumann_2d(16)).toarray()
    B = M.to_scipy(M.laplacian_B_2d(16, 3)).toarray()
    w = sl.eigh(B, A + 1e-3 * B, eigvals_only=True)
    lam = 1.0 / w[::-1][:8] - 1e-3
    assert np.abs(lam - golden["d_geninv_eval"]).max() < 5e-12


def test_edge_cases(oracles):
    """a single panel; nev that is not a multiple of 8 (columns are padded, eigensolver.hh:43)."""
    for o in both(oracles):
        q = o.orthonormalize(np.arange(80, dtype=float).reshape(10, 8) ** 2 % 7 + np.eye(10, 8))
        assert np.abs(q.T @ q - np.eye(8)).max() < 1e-12
        ev, V, k = o.standard_largest(M.laplacian_dirichlet_2d(8), 0.0, 1e-10, 3000, 5)
        assert ev.shape == (5,) and V.shape == (5, 64)
        an = M.eigenvalues_laplace_dirichlet_2d(8)[::-1][:5]
        assert np.abs(ev - an).max() < 1e-7
