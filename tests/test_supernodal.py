"""The second factorisation provider -- supernodal multifrontal Cholesky (include/dune/eigensolver/supernodal_cholesky.hh,
C ABI de_host_factorize_spd) -- and the supernodal factored apply on the GPU (csrc/kernels_snode.cuh).

The factorisation itself is parity-unpinned like the LU provider (UMFPACK is absent); what IS pinned:
  * the factor arrays, expanded into the reference's UMFPACK field contract, drive the REFERENCE's own apply
    (matmul_inverse_tallskinny_blocked, kernels_cpp.hh:660-755) to the exact inverse -- CPU test
  * the GPU supernodal apply equals the reference's apply on the same factorisation -- GPU test
  * GeneralizedInverse with this provider: iteration count +-1 and eigenvalues against the reference run with the same
    factorisation, and against the analytic spectrum of the Q1 pencil -- GPU test"""
import numpy as np
import pytest

from dune_eigensolver_b200 import eigensolver as E, matrices as M


def _pencil(shape, shift=1e-3):
    K, Mm = M.q1_stiffness(shape), M.q1_mass(shape)
    A = (K[0], K[1], K[2] + shift * Mm[2])
    return K, Mm, A


@pytest.mark.parametrize("shape,ordering", [((7, 6, 5), 1), ((9, 8, 7), 1), ((12, 12), 0), ((10, 9, 8), 2)])
def test_supernodal_factor_through_the_reference_apply(oracle, shape, ordering):
    K, Mm, A = _pencil(shape)
    n = len(A[0]) - 1
    hF = E.HostFactorization(A, ordering=ordering, spd=True, nthreads=2)
    assert hF.supernodal and hF.info["lnz"] >= n and hF.info["stored"] >= hF.info["lnz"]
    F = hF.arrays()
    # the contract: L unit lower by rows with the diagonal last, U upper by columns with the diagonal last, P = Q
    assert np.array_equal(F["P"], F["Q"]) and sorted(F["P"]) == list(range(n)) and np.all(F["Rs"] == 1.0)
    assert np.all(F["Lj"][F["Lp"][1:] - 1] == np.arange(n)) and np.all(F["Lx"][F["Lp"][1:] - 1] == 1.0)
    assert np.all(F["Ui"][F["Up"][1:] - 1] == np.arange(n)) and np.all(F["Ux"][F["Up"][1:] - 1] > 0.0)
    X = oracle.start_block(n, 16, 123)
    S = M.to_scipy(A)
    sol, _ = oracle.factor_apply(F, S @ X)      # the reference's own apply on these arrays
    assert np.abs(sol - X).max() <= 1e-10 * np.abs(X).max()
    hF.close()


@pytest.mark.parametrize("shape,kind", [((9, 8, 7), "q1"), ((13, 5, 4), "fd"), ((17, 9), "q1"), ((3, 3, 40), "q1"), ((23, 23), "fd")])
def test_geometric_nested_dissection_on_structured_grids(oracle, shape, kind):
    """ordering 4: plane separators on the detected grid (the default from 50 000 rows for stencils with diagonal neighbours;
    the Q1 Laplace stiffness matrix has exact zeros on the faces, which the provider drops: the detection must cope). Both
    providers, checked through the reference's own apply; ordering 3 forces METIS for comparison."""
    import scipy.sparse as sp

    A = M.q1_stiffness(shape) if kind == "q1" else (M.laplacian_fd(shape) if len(shape) == 3 else M.laplacian_dirichlet_2d(shape[0]))
    n = len(A[0]) - 1
    S = (sp.csr_matrix((A[2].copy(), A[1], A[0])) + 1e-3 * sp.identity(n, format="csr")).tocsr()
    S.sort_indices()
    csr = (S.indptr, S.indices, S.data)
    X = oracle.start_block(n, 8, 123)
    fills = {}
    for ordering in (4, 3):
        for spd in (True, False):
            hF = E.HostFactorization(csr, ordering=ordering, spd=spd, nthreads=2) if spd else E.HostFactorization(csr, ordering=ordering)
            F = hF.arrays()
            assert sorted(F["P"]) == list(range(n))
            sol, _ = oracle.factor_apply(F, S @ X)
            assert np.abs(sol - X).max() <= 1e-10 * np.abs(X).max()
            fills[(ordering, spd)] = hF.info["lnz"] if spd else hF.lnz
            hF.close()
    # plane separators are within a factor of two of METIS's fill on these small grids (equal on large 27-point grids)
    assert fills[(4, True)] <= 2.0 * fills[(3, True)]


def test_supernodal_rejects_indefinite_matrices():
    A = M.laplacian_dirichlet_2d(8)
    v = A[2].copy()
    v[A[1] == np.repeat(np.arange(64), np.diff(A[0]))] = -1.0   # negative diagonal
    with pytest.raises(E.DeError) as e:
        E.HostFactorization((A[0], A[1], v), spd=True)
    assert e.value.status == E.capi.DE_ERR_SINGULAR and "positive definite" in str(e.value)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,m", [((7, 6, 5), 8), ((12, 11, 10), 16), ((20, 20, 20), 32), ((16, 15, 14), 64), ((40, 40), 24)])
def test_supernodal_apply_matches_the_reference_apply(ctx, oracle, shape, m):
    K, Mm, A = _pencil(shape)
    n = len(A[0]) - 1
    hF = E.HostFactorization(A, ordering=1, spd=True, nthreads=4)
    dF = E.Factor(ctx, hF)
    X = oracle.start_block(n, m, 123)
    B = M.to_scipy(A) @ X
    dX, dY = E.MultiVector.from_array(ctx, B), E.MultiVector(ctx, n, m)
    E.matmul_inverse_tallskinny_blocked(dY, dF, dX)
    got = dY.download()
    ref, _ = oracle.factor_apply(hF.arrays(), B)
    assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max()
    assert np.abs(got - X).max() <= 1e-9 * np.abs(X).max()
    # a second apply (the captured graph of the sweeps is replayed) gives the same answer
    dX.upload(B)
    E.matmul_inverse_tallskinny_blocked(dY, dF, dX)
    assert np.abs(dY.download() - ref).max() <= 1e-11 * np.abs(ref).max()
    for h in (dX, dY, dF, hF):
        h.close()


@pytest.mark.gpu
def test_generalized_inverse_with_the_cholesky_provider(ctx, oracle):
    """configs[2] in small: Q1 stiffness + mass pencil, shift-invert with the factored solve"""
    shape, nev, shift, tol = (14, 13, 12), 16, 1e-3, 1e-10
    K, Mm = M.q1_stiffness(shape), M.q1_mass(shape)
    r = E.GeneralizedInverse(ctx, K, Mm, shift, 0.0, tol, 4000, nev, factorization="cholesky", nthreads=4)
    an = M.eigenvalues_q1_pencil(shape)[:nev]
    assert np.abs(np.sort(r.eval) - an).max() <= 1e-8 * an.max()
    assert r.factor_info["lnz"] > 0 and r.factor_info["flops"] > 0
    # the reference's driver with the scalar-LU provider: same algorithm, different (exact) factorisation
    ev, V, it = oracle.generalized_inverse(K, Mm, shift, 0.0, tol, 4000, nev)
    assert abs(r.iterations - it) <= 1, (r.iterations, it)
    assert np.abs(r.eval - ev).max() <= 1e-9 * np.abs(ev).max()
