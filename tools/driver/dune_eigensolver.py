#!/usr/bin/env python
"""The reference's driver program (reference src/dune-eigensolver.cc:448-787) on the B200 path.

Reads `dune-eigensolver.ini` (same [ev] keys; `-ev.N 100`-style command-line overrides as dune-common's
ParameterTreeParser::readOptions accepts) and runs one of the reference's three tests with the same matrices, the same
driver calls and the same output tables:

    --test largest      largest_eigenvalues_convergence_test (:620-730, what the shipped main() runs):
                        StandardLargest on the 2D Dirichlet Laplacian against the analytic spectrum and ARPACK
    --test smallest     smallest_eigenvalues_convergence_test (:528-617): GeneralizedInverse on (Neumann Laplacian,
                        partition-of-unity B) against ARPACK shift-invert
    --test eigenvalues  eigenvalues_test (:448-525), method = raes | arpack
    --test lobpcg       NEW (the reference has no LOBPCG): StandardLOBPCG, smallest ev.m eigenpairs of the 2D Dirichlet
                        Laplacian without a factorisation, against the analytic spectrum (:437-446); also reached
                        with `--test eigenvalues -ev.method lobpcg`

ARPACK++ (absent here) is replaced by scipy.sparse.linalg.eigsh, which wraps the same ARPACK routines
(shift-invert, smallest magnitude). The thread-replica harness (`parallel.numthreads`, a CPU bandwidth benchmark)
is out of scope: one GPU context, one solve.
"""
import argparse
import configparser
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def read_parameters(path, overrides):
    cp = configparser.ConfigParser(inline_comment_prefixes=("#",))
    cp.optionxform = str  # keys are case-sensitive (ev.N)
    cp.read(path)
    tree = {"%s.%s" % (s, k): v.strip() for s in cp.sections() for k, v in cp[s].items()}
    it = iter(overrides)
    for key in it:  # "-ev.N 100"
        tree[key.lstrip("-")] = next(it)
    return tree


def arpack_shift_invert(A, B, sigma, tol, m):
    """smallest-magnitude eigenvalues of A x = lambda B x by ARPACK in shift-invert mode around sigma
    (what computeGenSymShiftInvertMinMagnitude does, reference arpack_geneo_wrapper.hh); returns (values, iterations)"""
    import scipy.sparse.linalg as sla

    count = [0]

    class CountingSolve(sla.LinearOperator):
        def __init__(self, lu, n):
            super().__init__(dtype=np.float64, shape=(n, n))
            self.lu = lu

        def _matvec(self, x):
            count[0] += 1
            return self.lu.solve(np.asarray(x).ravel())

    n = A.shape[0]
    lu = sla.splu((A - sigma * B).tocsc())
    vals = sla.eigsh(A, k=m, M=B, sigma=sigma, which="LM", tol=tol, OPinv=CountingSolve(lu, n),
                     return_eigenvectors=False)
    return np.sort(vals), count[0]


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--ini", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "dune-eigensolver.ini"))
    ap.add_argument("--test", default="largest", choices=["largest", "smallest", "eigenvalues", "lobpcg"])
    ap.add_argument("--no-arpack", action="store_true", help="skip the ARPACK comparison columns")
    args, rest = ap.parse_known_args()
    p = read_parameters(args.ini, rest)

    from dune_eigensolver_b200 import eigensolver as E, matrices as M

    print("Hello World! This is dune-eigensolver.")
    N, overlap, m = int(p["ev.N"]), int(p["ev.overlap"]), int(p["ev.m"])
    maxiter, tol, verbose = int(p["ev.maxiter"]), float(p["ev.tol"]), int(p["ev.verbose"])
    shift, reg, seed = float(p["ev.shift"]), float(p["ev.regularization"]), int(p.get("ev.seed", "123"))
    method = p.get("ev.method", "raes")
    n = N * N
    ctx = E.Context(0)
    sci = M.to_scipy

    if args.test == "lobpcg" or (args.test == "eigenvalues" and method == "lobpcg"):
        # the pencil of eigenvalues_test has a singular B (rows masked by the partition of unity), which no
        # B-orthonormal iteration without shift-invert can use; the new driver is shown on the Dirichlet Laplacian
        A = M.laplacian_dirichlet_2d(N)
        t0 = time.perf_counter()
        r = E.StandardLOBPCG(ctx, A, tol, maxiter, m, verbose=verbose, seed=seed)
        dt = time.perf_counter() - t0
        exact = M.eigenvalues_laplace_dirichlet_2d(N)[:m]
        for i, ev in enumerate(r.eval):
            print("eval[%3d]=%20.12e %.2e" % (i, ev, abs(ev - exact[i])))
        print(": eigensolver elapsed time %g" % dt)
        print("N_M_TOL_LOBPCGERROR_ITER %d & %d & %g & %g & %d \\\\" % (n, m, tol, float(np.max(np.abs(r.eval - exact))),
                                                                   r.iterations))
        return 0

    if args.test == "eigenvalues":
        A, B = M.laplacian_neumann_2d(N), M.laplacian_B_2d(N, overlap)
        if method == "raes":
            t0 = time.perf_counter()
            r = E.GeneralizedInverse(ctx, A, B, shift, reg, tol, maxiter, m, verbose=verbose)
            dt = time.perf_counter() - t0
            for i, ev in enumerate(r.eval):
                print("eval[%3d]=%20.12e" % (i, ev))
            print("0: eigensolver elapsed time %g" % dt)
        elif method == "arpack":
            t0 = time.perf_counter()
            vals, _ = arpack_shift_invert(sci(A), sci(B), -shift, tol, m)
            dt = time.perf_counter() - t0
            for i, ev in enumerate(vals):
                print("eval[%3d]=%20.12e" % (i, ev))
            print("arpack elapsed time %g" % dt)
        return 0

    if args.test == "smallest":
        A, B = M.laplacian_neumann_2d(N), M.laplacian_B_2d(N, overlap)
        acc = acc2 = np.full(m, np.nan)
        time_arpack, iters = float("nan"), 0
        if not args.no_arpack:
            acc, _ = arpack_shift_invert(sci(A), sci(B), -shift, 1e-14, m)
            t0 = time.perf_counter()
            acc2, iters = arpack_shift_invert(sci(A), sci(B), -shift, tol, m)
            time_arpack = time.perf_counter() - t0
            print(": arpack elapsed time %g" % time_arpack)
        maxerror2 = float(np.max(np.abs(acc2 - acc))) if not args.no_arpack else float("nan")
        t0 = time.perf_counter()
        r = E.GeneralizedInverse(ctx, A, B, shift, reg, tol, maxiter, m, verbose=verbose, seed=seed)
        time_es = time.perf_counter() - t0
        for i, ev in enumerate(r.eval):
            print("eval[%3d]=%10.2e %.2e" % (i, ev, abs(ev - acc2[i])))
        print(": eigensolver elapsed time %g" % time_es)
        maxerror = float(np.max(np.abs(r.eval - acc2))) if not args.no_arpack else float("nan")
        print("N_M_TOL_RASERROR_ARPERROR_TIMERATIO_ARPACKITER %d & %d & %g & %g & %g & %g & %d \\\\" %
              (n, m, tol, maxerror, maxerror2, time_es / time_arpack, iters))
        return 0

    # largest (the shipped main): 2D Dirichlet Laplacian, B = identity, shift forced to 0 (:643)
    A = M.laplacian_dirichlet_2d(N)
    shift = 0.0
    acc = acc2 = np.full(m, np.nan)
    time_arpack, iters = float("nan"), 0
    if not args.no_arpack:
        Id = M.to_scipy(M.identity_on_laplacian_pattern_2d(N))
        acc, _ = arpack_shift_invert(sci(A), Id, -shift - 1e-9, 1e-14, m)
        t0 = time.perf_counter()
        acc2, iters = arpack_shift_invert(sci(A), Id, -shift - 1e-9, tol, m)
        time_arpack = time.perf_counter() - t0
        print(": arpack elapsed time %g" % time_arpack)
    maxerror2 = float(np.max(np.abs(acc2 - acc))) if not args.no_arpack else float("nan")
    t0 = time.perf_counter()
    r = E.StandardLargest(ctx, (A[0], A[1], A[2].copy()), shift, tol, maxiter, m, verbose=verbose, seed=seed)
    time_es = time.perf_counter() - t0
    ana = M.eigenvalues_laplace_dirichlet_2d(N)
    ana = np.sort(ana)[::-1][:m]  # the m largest, as StandardLargest converges to them (descending)
    ev = np.asarray(r.eval)
    order = np.argsort(ev)[::-1]
    print("eval_num__EIGENSOLVER_ANALYTICAL_ARPACKACR_ARPACKTOL_ESANERROR_ESARERR")
    for i in range(m):
        print("eval[%3d]=%10.2e  %.2e  %.2e  %.2e  %.2e  %.2e" % (i, ev[i], ana[i], acc[i], acc2[i], abs(ev[order][i] - ana[i]),
                                                                   abs(ev[i] - acc[i])))
    print(": eigensolver elapsed time %g (%d iterations)" % (time_es, r.iterations))
    maxerror3 = float(np.max(np.abs(ev[order] - ana)))
    maxerror = float(np.max(np.abs(ev - acc))) if not args.no_arpack else float("nan")
    print("N_M_TOL_ESARERROR_ARPERROR_ESANERROR_TIMERATIO_ARPACKITER ")
    print("%d & %d & %g & %g & %g & %g & %g & %d \\\\" % (n, m, tol, maxerror, maxerror2, maxerror3, time_es / time_arpack, iters))
    return 0


if __name__ == "__main__":
    sys.exit(main())
