"""Generates tests/golden/reference_vectors.npz from the reference compiled VERBATIM (oracle/_ref).

Run in the CPU container where /root/reference is mounted:

    make -C oracle ref && python tests/golden/make_golden.py

The reference has no golden vectors of its own (no test directory; results are eyeballed), so these
fixtures are outputs of the reference's own functions (reference dune/eigensolver/kernels_cpp.hh and
eigensolver.hh, included from the read-only mount by oracle/ref_capi.cc) on seeded inputs. They travel with the
repo; /root/reference itself is never read at test time.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from dune_eigensolver_b200 import matrices as M  # noqa: E402


def main():
    ref = O.load_reference()
    if ref is None or ref.kind != "reference":
        raise SystemExit("oracle/_ref/libde_reference.so missing: run `make -C oracle ref` where /root/reference exists")
    out = {}
    # ---- kernel level: 2D 5-point matrices of the reference's own generators, N = 6 (n = 36), m = 16 -------------
    N, m = 6, 16
    n = N * N
    A = M.laplacian_dirichlet_2d(N)
    B = M.laplacian_B_2d(N, 1)
    X = ref.start_block(n, m, 123)   # the reference's own start block stream (seed 123)
    Y = ref.start_block(n, m, 7)
    out["k_N"], out["k_m"] = N, m
    out["k_X"], out["k_Y"] = X, Y
    out["k_spmm"] = ref.spmm(A, X)
    out["k_diag_dot"] = ref.diag_dot(X, Y)
    out["k_gram"] = ref.gram(X, Y)
    out["k_ortho"] = ref.orthonormalize(X)
    out["k_ortho_naive"] = ref.orthonormalize_naive(X)
    q, nrm = ref.b_orthonormalize(B, X)
    out["k_bortho"], out["k_bortho_norm"] = q, nrm
    # factored apply: factors of A + 0.5 I from the repo's host provider, exported so the test needs no provider
    from dune_eigensolver_b200 import eigensolver as E
    As = (A[0], A[1], A[2].copy())
    E._add_to_diagonal(As[0], As[1], As[2], 0.5)
    hf = E.HostFactorization(As, ordering=1, scale_rows=True)
    F = hf.arrays()
    for k, v in F.items():
        out["f_" + k] = np.asarray(v)
    sol, clob = ref.factor_apply(F, X)
    out["f_apply"] = sol
    out["k_cost_flops"] = ref.flops_orthonormalize(1000, 24)
    out["k_cost_bytes_naive"] = ref.bytes_orthonormalize_naive(1000, 24)
    out["k_cost_bytes_blocked"] = ref.bytes_orthonormalize_blocked(1000, 24, 8)

    # ---- driver level (eigenvalues + the reference's own iteration counters) ---------------------------------------
    # (a) SURVEY.md §4 item 4: StandardLargest, N = 20, nev = 8, tol = 1e-10
    ev, V, k = ref.standard_largest(M.laplacian_dirichlet_2d(20), 0.0, 1e-10, 4000, 8)
    out["d_largest_eval"], out["d_largest_iter"] = ev, k
    # (b) SURVEY.md §4 item 4: GeneralizedInverse, N = 16, overlap = 3, shift = 1e-3, nev = 8, tol = 1e-12 -> 121 iterations
    ev, V, it = ref.generalized_inverse(M.laplacian_neumann_2d(16), M.laplacian_B_2d(16, 3), 1e-3, 0.0, 1e-12, 4000, 8)
    out["d_geninv_eval"], out["d_geninv_iter"] = ev, it
    # (c) StandardInverse, N = 20, shift 1e-3, nev = 8, tol 1e-10
    ev, V, k = ref.standard_inverse(M.laplacian_dirichlet_2d(20), 1e-3, 1e-10, 4000, 8)
    out["d_inverse_eval"], out["d_inverse_iter"] = ev, k
    # (d) the shipped ini (src/dune-eigensolver.ini) at reduced N: tol 2e-3, shift 1e-3, overlap 3, nev 16 (BASELINE C1)
    ev, V, it = ref.generalized_inverse(M.laplacian_neumann_2d(40), M.laplacian_B_2d(40, 3), 1e-3, 0.0, 2e-3, 4000, 16)
    out["d_ini_eval"], out["d_ini_iter"] = ev, it
    ev, V, k = ref.standard_largest(M.laplacian_dirichlet_2d(40), 0.0, 2e-3, 4000, 16)
    out["d_ini_largest_eval"], out["d_ini_largest_iter"] = ev, k

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;",
          "largest k=%d geninv it=%d inverse k=%d ini it=%d ini_largest k=%d" %
          (out["d_largest_iter"], out["d_geninv_iter"], out["d_inverse_iter"], out["d_ini_iter"],
           out["d_ini_largest_iter"]))


if __name__ == "__main__":
    main()
