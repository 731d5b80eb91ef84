#!/usr/bin/env python
"""BASELINE.json configuration 0 -- the reference's shipped ini (2D 200 x 200 Neumann Laplacian, partition-of-unity B,
GeneralizedInverse, tol 2e-3, shift 1e-3) with 16 eigenpairs -- on the B200 path next to the reference itself
(oracle/_ref, one host core), same factorisation provider for both so that only the iteration differs."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dune_eigensolver_b200 import eigensolver as E, matrices as M
from oracle import oracle as O

def main():
    N, nev, shift, reg, tol, maxiter = 200, 16, 1e-3, 0.0, 2e-3, 4000
    A, B = M.laplacian_neumann_2d(N), M.laplacian_B_2d(N, 3)
    ctx = E.Context(0)
    for prov in ("lu", "cholesky"):
        for rep in range(3):
            t0 = time.perf_counter()
            r = E.GeneralizedInverse(ctx, A, B, shift, reg, tol, maxiter, nev, factorization=prov)
            dt = time.perf_counter() - t0
            print("B200 %-8s: %.3f s total (%.3f s host factorisation), %d iterations, %.2f ms per iteration incl. setup" %
                  (prov, dt, r.time_factorization or 0.0, r.iterations, (dt - (r.time_factorization or 0.0)) / max(r.iterations, 1) * 1e3))
    orc = O.load_best()
    t0 = time.perf_counter()
    ev, V, it = orc.generalized_inverse(A, B, shift, reg, tol, maxiter, nev)
    dt = time.perf_counter() - t0
    print("reference : %.3f s total on one host core (%s), %d iterations" % (dt, orc.kind, it))
    print("max |eval_b200 - eval_ref| = %.2e (relative to max |eval|: %.2e)" %
          (np.abs(r.eval - ev).max(), np.abs(r.eval - ev).max() / np.abs(ev).max()))

if __name__ == "__main__":
    main()
