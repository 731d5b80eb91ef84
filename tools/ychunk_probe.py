#!/usr/bin/env python
"""Which y-chunk of the BRB tile order (brb::grid_order) keeps the halo planes of X in the L2? Times the SpMM kernel on one
matrix for several values of the brb_plane_points option (the matrix is re-uploaded for each: the order is fixed at build time).

    python tools/ychunk_probe.py --grid 256 --stencil q1 --cols 32
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--stencil", default="q1", choices=["fd", "q1"])
    ap.add_argument("--cols", default="32")
    ap.add_argument("--points", default="1000000000,32768,16384,8192,4096,2048")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    from dune_eigensolver_b200 import eigensolver as E, matrices as M

    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    shape = (args.grid,) * 3
    A = M.laplacian_fd(shape) if args.stencil == "fd" else M.q1_stiffness(shape)
    n, nnz = len(A[0]) - 1, len(A[1])
    ctx = E.Context(0)
    for m in [int(c) for c in args.cols.split(",")]:
        X = E.MultiVector(ctx, n, m)
        X.upload_rowmajor(np.random.default_rng(m).standard_normal((n, m)))
        Y = E.MultiVector(ctx, n, m)
        for pts in args.points.split(","):
            ctx.set_option("brb_plane_points", int(pts))
            dA = E.Matrix(ctx, A)
            for _ in range(3):
                E.matmul_sparse_tallskinny_with_dots(Y, dA, X)
            ctx.profile(reset=True)
            ctx.set_profiling(True)
            for _ in range(args.reps):
                E.matmul_sparse_tallskinny_with_dots(Y, dA, X)
            prof = ctx.profile(reset=True)
            ctx.set_profiling(False)
            ms = prof["spmm"][0] / max(prof["spmm"][1], 1)
            gbs = (12.0 * nnz + 4.0 * (n + 1) + 16.0 * n * m) / (ms * 1e-3) / 1e9
            print("grid %d %s m=%d plane_points=%s: spmm+dot %.4f ms %.0f GB/s %.3f of peak" %
                  (args.grid, args.stencil, m, pts, ms, gbs, gbs / peak), flush=True)
            dA.close()
        X.close()
        Y.close()


if __name__ == "__main__":
    main()
