#!/usr/bin/env python
"""Diagnostic: per-solve wall and device times of the bench workload, plus host CPU limits (for timing variance)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dune_eigensolver_b200 import eigensolver as E, matrices as M

def main():
    for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu.stat"):
        try:
            print(f, open(f).read().replace("\n", " | "))
        except Exception as e:
            print(f, "n/a", e)
    print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "loadavg", os.getloadavg())
    N, m = 100, 32
    ctx = E.Context(0)
    dA = E.Matrix(ctx, M.q1_stiffness((N,) * 3))
    n = N ** 3
    Q0 = E.MultiVector(ctx, n, m)
    Q0.upload_panels(E.start_block(n, m, 123))
    Q = E.MultiVector(ctx, n, m)
    for i in range(16):
        Q.copy_from(Q0)
        t0 = time.perf_counter()
        ev, it = E.standard_largest_mv(ctx, dA, 0.0, 2e-3, 4000, Q)
        dt = time.perf_counter() - t0
        print("solve %2d: %.2f ms (%d iterations)" % (i, dt * 1e3, it))
    try:
        print("/sys/fs/cgroup/cpu.stat", open("/sys/fs/cgroup/cpu.stat").read().replace("\n", " | "))
    except Exception:
        pass

if __name__ == "__main__":
    main()
