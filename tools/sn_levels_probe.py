#!/usr/bin/env python
"""Per-level launch list of the supernodal factored apply (run under `ncu --metrics gpu__time_duration.sum`): the sweeps are
launched one level at a time (no CUDA graph while the library's timers are on)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from dune_eigensolver_b200 import eigensolver as E, matrices as M

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=64)
ap.add_argument("--m", type=int, default=64)
args = ap.parse_args()
G, m = args.grid, args.m
K, Mm = M.q1_stiffness((G, G, G)), M.q1_mass((G, G, G))
A = (K[0], K[1], K[2] + 1e-3 * Mm[2])
hF = E.HostFactorization(A, ordering=1, spd=True, nthreads=0, arrays=False)
print(hF.info, flush=True)
ctx = E.Context(0)
dF = E.Factor(ctx, hF)
n = G ** 3
dX, dY = E.MultiVector.from_array(ctx, np.random.default_rng(1).standard_normal((n, m))), E.MultiVector(ctx, n, m)
ctx.set_profiling(True)
E.matmul_inverse_tallskinny_blocked(dY, dF, dX)
ctx.synchronize()
prof = ctx.profile(reset=True)
print("trsv category: %.3f ms in %d launches" % prof["trsv"], flush=True)
