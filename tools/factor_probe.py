#!/usr/bin/env python
"""configs[2] probe: Q1 stiffness + mass pencil on grid^3 nodes, supernodal Cholesky of A + shift B on the host, factored
apply (K6) and GeneralizedInverse on the GPU. Prints one JSON line."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from dune_eigensolver_b200 import eigensolver as E, matrices as M

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=64)
ap.add_argument("--nev", type=int, default=64)
ap.add_argument("--tol", type=float, default=2e-3)
ap.add_argument("--shift", type=float, default=1e-3)
ap.add_argument("--maxiter", type=int, default=400)
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--no-solve", action="store_true")
args = ap.parse_args()
G, m = args.grid, E.padded_cols(args.nev)
shape = (G, G, G)
t0 = time.perf_counter()
K, Mm = M.q1_stiffness(shape), M.q1_mass(shape)
n = G ** 3
A = (K[0], K[1], K[2] + args.shift * Mm[2])
t_gen = time.perf_counter() - t0
t0 = time.perf_counter()
hF = E.HostFactorization(A, ordering=1, spd=True, nthreads=args.threads, arrays=False)
t_fact = time.perf_counter() - t0
ctx = E.Context(0)
t0 = time.perf_counter()
dF = E.Factor(ctx, hF)
t_up = time.perf_counter() - t0
rng = np.random.default_rng(1)
X = rng.standard_normal((n, m))
import scipy.sparse as sp
S = M.to_scipy(A)
B = S @ X
dX, dY = E.MultiVector.from_array(ctx, B), E.MultiVector(ctx, n, m)
E.matmul_inverse_tallskinny_blocked(dY, dF, dX)
err = float(np.abs(dY.download() - X).max())
ctx.synchronize()
ts = []
for _ in range(args.reps):
    dX.upload(B)
    ctx.synchronize()
    t0 = time.perf_counter()
    E.matmul_inverse_tallskinny_blocked(dY, dF, dX)
    ctx.synchronize()
    ts.append(time.perf_counter() - t0)
apply_s = float(np.median(ts))
info = hF.info
# algorithmic bytes of the apply (SURVEY.md §8d form, with the factor as stored: 8 B per entry, read once per sweep)
bytes_apply = 2 * 8.0 * info["stored"] + 16.0 * n + 2 * 16.0 * n * m
flops_apply = 4.0 * info["lnz"] * m
out = {"workload": "Q1 stiffness + mass pencil %d^3 (n=%d), A + %g B, m=%d" % (G, n, args.shift, m),
       "factorization": {"provider": "supernodal multifrontal Cholesky, METIS nested dissection, host (%d threads)" % (args.threads or os.cpu_count()),
                         "seconds_total": t_fact, **info, "GFLOPs": info["flops"] / max(info["seconds_numeric"], 1e-9) / 1e9},
       "upload_s": t_up, "generate_s": t_gen,
       "apply": {"seconds": apply_s, "max_abs_error_vs_known_solution": err, "algorithmic_bytes": bytes_apply,
                 "GBps": bytes_apply / apply_s / 1e9, "flops": flops_apply, "TFLOPs": flops_apply / apply_s / 1e12}}
if not args.no_solve:
    t0 = time.perf_counter()
    start = E.start_block(n, m, 123)
    dA, dB = E.Matrix(ctx, A), E.Matrix(ctx, Mm)
    ev, V, it, rel = np.zeros(args.nev), np.zeros((args.nev, n)), E.C.c_int(0), E.C.c_double(0.0)
    ctx.synchronize()
    t1 = time.perf_counter()
    E.check(E.capi.lib().de_generalized_inverse(ctx._h, dA._h, dB._h, dF._h, args.shift, args.tol, args.maxiter, args.nev,
                                                E.dptr(start), E.dptr(ev), E.dptr(V), 0, E.C.byref(it), E.C.byref(rel)), ctx._h)
    t_solve = time.perf_counter() - t1
    Sk, Sm = M.to_scipy(K), M.to_scipy(Mm)
    res = [float(np.linalg.norm(Sk @ V[j] - ev[j] * (Sm @ V[j])) / max(abs(ev[j]) * np.linalg.norm(Sm @ V[j]), 1e-300)) for j in range(0, args.nev, max(1, args.nev // 8))]
    an = M.eigenvalues_q1_pencil(shape)[:args.nev]
    out["generalized_inverse"] = {"driver": "GeneralizedInverse (reference eigensolver.hh:204-351), tol %g, shift %g" % (args.tol, args.shift),
                                  "seconds": t_solve, "iterations": it.value, "relerror": rel.value,
                                  "eigenvalues_head": [float(x) for x in ev[:4]],
                                  "max_rel_error_vs_analytic": float(np.abs(np.sort(ev) - an).max() / an.max()),
                                  "relative_residuals_sample": res}
print(json.dumps(out))
