#!/usr/bin/env python
"""Diagnostic: per-solve wall time of the bench workload and the host-side enqueue time of the asynchronous loop."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dune_eigensolver_b200 import eigensolver as E, matrices as M

def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    ctx = E.Context(0)
    dA = E.Matrix(ctx, M.q1_stiffness((N,) * 3))
    n = N ** 3
    Q0 = E.MultiVector(ctx, n, m)
    Q0.upload_panels(E.start_block(n, m, 123))
    Q = E.MultiVector(ctx, n, m)
    for i in range(6):
        Q.copy_from(Q0)
        ctx.synchronize()
        t0 = time.perf_counter()
        ev, it = E.standard_largest_mv(ctx, dA, 0.0, 2e-3, 4000, Q, verbose=2 if i >= 3 else 0)
        dt = time.perf_counter() - t0
        print("solve %2d: %.2f ms (%d iterations)" % (i, dt * 1e3, it), flush=True)

if __name__ == "__main__":
    main()
