#!/usr/bin/env python
"""Diagnostic: where the end-to-end time of the bench workload goes (matrix creation / solve / copy-out)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dune_eigensolver_b200 import eigensolver as E, matrices as M

def main():
    N, m, nev = 100, 32, 32
    ctx = E.Context(0)
    rp, ci, v = M.q1_stiffness((N,) * 3)
    rp, ci, v = np.ascontiguousarray(rp, dtype=np.int64), np.ascontiguousarray(ci, dtype=np.int64), np.ascontiguousarray(v)
    n = N ** 3
    start = E.start_block(n, m, 123)
    for rep in range(4):
        t0 = time.perf_counter()
        dA = E.Matrix(ctx, (rp, ci, v))
        ctx.synchronize()
        t1 = time.perf_counter()
        evl, V, it = np.zeros(nev), np.zeros((nev, n)), E.C.c_int(0)
        E.check(E.capi.lib().de_standard_largest(ctx._h, dA._h, 0.0, 2e-3, 4000, nev, E.dptr(start), E.dptr(evl), E.dptr(V), 0,
                                                 E.C.byref(it)), ctx._h)
        t2 = time.perf_counter()
        dA.close()
        t3 = time.perf_counter()
        print("rep %d: create %.1f ms, solve+copies %.1f ms (%d it), destroy %.1f ms, info %s" %
              (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, it.value, (t3 - t2) * 1e3, ""))
    dA = E.Matrix(ctx, (rp, ci, v))
    print(dA.spmm_info())

if __name__ == "__main__":
    main()
