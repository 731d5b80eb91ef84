"""StandardLOBPCG on one B200: time-to-nev-smallest-eigenpairs of a 3D Laplacian (BASELINE.json configs[1]:
"3D Q1 Laplace 100^3, 32 eigenpairs via StandardLOBPCG on 1 B200"), printed as ONE JSON line.

    python tools/lobpcg_probe.py [--grid 100] [--stencil q1|fd] [--mass] [--nev 32] [--tol 2e-3] [--steps 3] [--verify]

The reference has no LOBPCG and its shift-invert drivers cannot reach the smallest eigenpairs of this matrix without a
3D factorisation, so there is no reference arm for this number; what pins the result is the analytic spectrum of the
matrix (matrices.eigenvalues_q1_stiffness) and, with --verify, residuals recomputed on the host with scipy.
bench.py runs this as a child process and attaches the line as its "lobpcg" object.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=100)
    ap.add_argument("--stencil", default="q1", choices=["q1", "fd"])
    ap.add_argument("--nev", type=int, default=32)
    ap.add_argument("--tol", type=float, default=2e-3)
    ap.add_argument("--maxiter", type=int, default=4000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cheb", default="8", help="degree of the Chebyshev preconditioner (0: none); 8 is the default of "
                                                "the StandardLOBPCG driver. A comma-separated list prints one line each.")
    ap.add_argument("--mass", action="store_true", help="generalized problem A x = lambda B x with the consistent Q1 mass "
                                                        "matrix (GeneralizedLOBPCG; BASELINE.json configs[2])")
    ap.add_argument("--contrast", type=float, default=0.0,
                    help="> 0: Q1 diffusion with coefficient `contrast` (1 elsewhere) in the block pattern of "
                         "matrices.high_contrast_kappa (BASELINE.json configs[3] on one GPU); no analytic spectrum")
    ap.add_argument("--e2e", action="store_true", help="also time the host-buffer driver call (host CSR in, host "
                                                       "eigenvectors out); standard problem, driver default degree")
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--verbose", type=int, default=0)
    args = ap.parse_args()

    from dune_eigensolver_b200 import eigensolver as E, matrices as M

    shape = (args.grid,) * 3
    n = args.grid ** 3
    m = E.padded_cols(args.nev)
    if args.mass:
        args.stencil = "q1"
    if args.contrast > 0.0:
        args.stencil = "q1"
        A = M.q1_stiffness(shape, kappa=M.high_contrast_kappa(args.contrast, 8))
    else:
        A = M.q1_stiffness(shape) if args.stencil == "q1" else M.laplacian_fd(shape)
    B = M.q1_mass(shape) if args.mass else None
    if args.contrast > 0.0:
        analytic = None
    elif args.mass:
        analytic = M.eigenvalues_q1_pencil(shape)[:m]
    else:
        analytic = (M.eigenvalues_q1_stiffness(shape) if args.stencil == "q1" else M.eigenvalues_laplacian_fd(shape))[:m]
    ctx = E.Context(0)
    dA = E.Matrix(ctx, A)
    dB = E.Matrix(ctx, B) if B is not None else None
    start = E.start_block(n, m, 123)
    Q = E.MultiVector(ctx, n, m)

    degrees = [int(x) for x in str(args.cheb).split(",")]
    for deg in degrees:
        args.cheb = deg
        probe_one(args, E, ctx, dA, dB, Q, start, A, B, analytic, n, m)


def probe_one(args, E, ctx, dA, dB, Q, start, A, B, analytic, n, m):
    def solve():
        Q.upload_panels(start)
        ctx.synchronize()
        t0 = time.perf_counter()
        out = E.lobpcg_mv(ctx, dA, Q, args.tol, args.maxiter, nev=args.nev, dB=dB, verbose=args.verbose,
                          cheb_degree=args.cheb)
        ctx.synchronize()
        return time.perf_counter() - t0, out

    solve()  # warm-up (kernel attributes, allocator cache)
    times, out = [], None
    for _ in range(args.steps):
        t, out = solve()
        times.append(t)
    lam, rn, it, restarts, conv = out
    # one more solve with every kernel category timed: where the time goes
    ctx.profile(reset=True)
    ctx.set_profiling(True)
    t_prof, _ = solve()
    prof = ctx.profile(reset=True)
    ctx.set_profiling(False)
    kernel_ms = sum(v[0] for v in prof.values())
    line = {
        "driver": "%s (de_lobpcg_mv, blocks resident in HBM)" % ("GeneralizedLOBPCG" if args.mass else "StandardLOBPCG"),
        "workload": "3D %s %d^3 (n=%d), %d smallest eigenpairs, m=%d, relative residual tol=%g, seed=123" %
                    (("Q1 diffusion, coefficient %g in a pattern of 8-cell blocks and 1 elsewhere (27-point)" % args.contrast
                      if args.contrast > 0.0 else "") +
                     (" + consistent mass pencil" if args.mass and args.contrast > 0.0 else
                      "Q1 stiffness + consistent mass pencil (27-point)" if args.mass else
                      "" if args.contrast > 0.0 else
                      "Q1 27-point FE stiffness Laplace" if args.stencil == "q1" else "7-point FD Laplace"), args.grid, n,
                     args.nev, m, args.tol),
        "chebyshev_degree": args.cheb, "seconds": float(np.median(times)), "seconds_all": [round(t, 5) for t in times],
        "iterations": it, "restarts": restarts, "converged": conv, "ms_per_iteration": 1e3 * float(np.median(times)) / max(it, 1),
        "max_rel_residual": float((rn[:args.nev] / np.abs(lam[:args.nev])).max()),
        "max_rel_eigenvalue_error_vs_analytic":
            float((np.abs(lam[:args.nev] - analytic[:args.nev]) / analytic[:args.nev]).max()) if analytic is not None else None,
        "eigenvalues_head": [float(x) for x in lam[:4]],
        "kernel_ms_per_solve": {k: round(v[0], 3) for k, v in prof.items() if v[1] > 0},
        "kernel_launches_per_solve": {k: int(v[1]) for k, v in prof.items() if v[1] > 0},
        "host_share": max(0.0, 1.0 - kernel_ms * 1e-3 / t_prof) if t_prof > 0 else None,
        "reference_arm": None,
        "note": "no reference arm: the reference has no LOBPCG and no factorisation-free route to these eigenpairs",
    }
    if args.e2e and B is None and args.cheb == 8:
        t_e2e = []
        for _ in range(2):  # the second call is the measurement (allocator cache, pinned buffers warm)
            t0 = time.perf_counter()
            r = E.StandardLOBPCG(ctx, A, args.tol, args.maxiter, args.nev)
            t_e2e.append(time.perf_counter() - t0)
        line["e2e_seconds"] = t_e2e[-1]
        line["e2e_iterations"] = r.iterations
        line["e2e_api"] = "de_matrix_create_csr + de_standard_lobpcg (host CSR in, host eigenvectors out)"
    if args.verify:
        import scipy.sparse as sp

        As = sp.csr_matrix((A[2], A[1], A[0]), shape=(n, n))
        X = Q.download_rowmajor()[:, :args.nev]
        BX = X if B is None else sp.csr_matrix((B[2], B[1], B[0]), shape=(n, n)) @ X
        R = As @ X - BX * lam[:args.nev]
        line["verified_max_rel_residual"] = float((np.linalg.norm(R, axis=0) / np.abs(lam[:args.nev])).max())
        line["verified_orthonormality_defect"] = float(np.abs(X.T @ BX - np.eye(args.nev)).max())
    print(json.dumps(line))


if __name__ == "__main__":
    main()
