import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from dune_eigensolver_b200 import eigensolver as E, matrices as M
ctx = E.Context(0)
# C1: 2D 5-point Laplace 200x200, 16 eigenpairs, StandardLargest, tol 1e-10 (parity run) and 2e-3 (shipped ini)
A = M.laplacian_dirichlet_2d(200)
n, m = 200 * 200, 16
dA = E.Matrix(ctx, A)
print(dA.spmm_info())
Q0 = E.MultiVector(ctx, n, m); Q0.upload_panels(E.start_block(n, m, 123)); Q = E.MultiVector(ctx, n, m)
for tol in (2e-3, 1e-10):
    for rep in range(3):
        Q.copy_from(Q0); ctx.synchronize(); t0 = time.perf_counter()
        ev, it = E.standard_largest_mv(ctx, dA, 0.0, tol, 4000, Q)
        dt = time.perf_counter() - t0
        print("C1 tol=%g: %.2f ms, %d iterations, %.1f us/iteration" % (tol, dt * 1e3, it, dt * 1e6 / it))
