"""BASELINE.json configuration 2 AT FULL SIZE (what bench.py measures): 3D Laplace on 100^3 nodes, 32 eigenpairs,
StandardLargest with the shipped ini parameters, through the reference-facing call with host buffers -- against the
reference itself run at that size (tests/golden/reference_fullsize.npz, made by make_golden_fullsize.py from the
reference's headers compiled verbatim; ~50 s of CPU per case, which is why it is a fixture and not an oracle call).

north_star bar: iteration count within +-1, eigenvalues within 1e-10 relative when the counts agree (otherwise within
the run's own tolerance), eigenvector residuals ||A x - lambda x|| no worse than the reference's."""
import os

import numpy as np
import pytest

from dune_eigensolver_b200 import eigensolver as E, matrices as M

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def full():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_fullsize.npz"))


@pytest.mark.parametrize("tag", ["q1", "fd"])
def test_bench_workload_matches_the_reference_at_full_size(ctx, full, tag):
    N, nev, tol = int(full[tag + "_N"]), int(full[tag + "_nev"]), float(full[tag + "_tol"])
    gen = M.q1_stiffness if tag == "q1" else M.laplacian_fd
    A = gen((N, N, N))
    r = E.StandardLargest(ctx, (A[0], A[1], A[2].copy()), 0.0, tol, 4000, nev)
    k_ref, ev_ref = int(full[tag + "_iterations"]), full[tag + "_eval"]
    assert abs(r.iterations - k_ref) <= 1, (r.iterations, k_ref)
    scale = np.abs(ev_ref).max()
    if r.iterations == k_ref:
        assert np.abs(r.eval - ev_ref).max() <= 1e-10 * scale
        idx = full[tag + "_sample_idx"]
        V = np.asarray(r.evec)
        assert np.abs(V[:, idx] - full[tag + "_sample"]).max() <= 1e-8  # unit vectors; same iterate up to rounding
    else:
        assert np.abs(r.eval - ev_ref).max() <= tol * scale
    S = M.to_scipy(A)
    V = np.asarray(r.evec)
    res = np.array([np.linalg.norm(S @ V[j] - r.eval[j] * V[j]) for j in range(nev)])
    assert (res <= 2.0 * full[tag + "_residual"] + 1e-9).all()
    # the converged block is orthonormal
    G = V @ V.T
    assert np.abs(G - np.eye(nev)).max() <= 1e-12
