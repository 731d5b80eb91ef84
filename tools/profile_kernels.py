#!/usr/bin/env python
"""Short driver for ncu: launches each hot-path kernel a few times at the bench size (3D Q1 100^3, m = 32 by default).

    python tools/profile_kernels.py && ncu --set full --clock-control none --import-source on \
        -k regex:'spmm|gram|update' -c 12 -o gpurun_out/prof python tools/profile_kernels.py
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=100)
    ap.add_argument("--stencil", default="q1")
    ap.add_argument("--m", type=int, default=32)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    from dune_eigensolver_b200 import eigensolver as E, matrices as M

    shape = (args.grid,) * 3
    A = M.q1_stiffness(shape) if args.stencil == "q1" else M.laplacian_fd(shape)
    n = len(A[0]) - 1
    ctx = E.Context(0)
    dA = E.Matrix(ctx, A)
    rng = np.random.default_rng(0)
    X = E.MultiVector(ctx, n, args.m)
    X.upload_rowmajor(rng.standard_normal((n, args.m)))
    Y = E.MultiVector(ctx, n, args.m)
    R = np.triu(rng.standard_normal((args.m, args.m))) / args.m + np.eye(args.m)
    for _ in range(args.reps):
        E.matmul_sparse_tallskinny_blocked(Y, dA, X)
        E.matmul_sparse_tallskinny_with_dots(Y, dA, X)
        E.dot_products_all_blocked(X, X)
        E.dot_products_all_blocked(X, Y)
        E.block_update(Y, R)
        E.orthonormalize_blocked(Y)
    ctx.synchronize()
    print("ok launches", ctx.launch_count())


if __name__ == "__main__":
    main()
