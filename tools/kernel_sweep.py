#!/usr/bin/env python
"""Block-width sweep of the hot-path kernels against the HBM roofline (BASELINE.json config 5):
SpMM, SpMM + fused dots, Gram, block update, diag-dot and the complete orthonormalisation for m = 8/16/32/64 on
a 3D Laplacian. Per-kernel durations come from the library's CUDA-event timers (events on the launching stream).

    python tools/kernel_sweep.py --grid 200 --stencil fd [--csv profiles/sweep.csv]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


FP64_TENSOR_TFLOPS = 37.0  # measured DMMA rate of one B200 (tools/micro/fp64_peak.cu)


def tensor_flops(kernel, n, m):
    """flops the kernel issues on the FP64 tensor pipe (what bounds the wide-block kernels): Gram X^T X computes the 8 x 8 tiles
    on or above the diagonal only, a triangular factor skips the zero half; None for the streaming / sparse kernels"""
    nb = m // 8
    full = 2.0 * n * m * m
    tri = full * (nb * (nb + 1) / 2) / (nb * nb)
    # (the "update" row calls de_block_update with a general factor: full product; the orthonormalisation's updates are triangular)
    return {"gram_xx": tri, "gram_xy": full, "update": full, "ortho": 4.0 * tri, "lincomb3": 3.0 * full}.get(kernel)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=200)
    ap.add_argument("--stencil", default="fd", choices=["fd", "q1"])
    ap.add_argument("--cols", default="8,16,32,64")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--csv", default=None)
    args = ap.parse_args()

    from dune_eigensolver_b200 import eigensolver as E, matrices as M

    peak = 6650.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    shape = (args.grid,) * 3
    A = M.laplacian_fd(shape) if args.stencil == "fd" else M.q1_stiffness(shape)
    n, nnz = len(A[0]) - 1, len(A[1])
    ctx = E.Context(0)
    dA = E.Matrix(ctx, A)
    del A
    rows = []
    print("grid %d^3 %s n=%d nnz=%d peak=%.0f GB/s" % (args.grid, args.stencil, n, nnz, peak))
    print("%4s %-10s %10s %10s %8s" % ("m", "kernel", "avg ms", "GB/s", "frac"))
    for m in [int(c) for c in args.cols.split(",")]:
        rng = np.random.default_rng(m)
        X = E.MultiVector(ctx, n, m)
        X.upload_rowmajor(rng.standard_normal((n, m)))
        Y = E.MultiVector(ctx, n, m)
        R = np.triu(rng.standard_normal((m, m))) / m + np.eye(m)
        bytes_ = {
            "spmm": 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n * m,
            "spmm+dot": 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n * m,
            "dot": 16.0 * n * m,
            "gram_xx": 8.0 * n * m,
            "gram_xy": 16.0 * n * m,
            "update": 16.0 * n * m,
            "ortho": 24.0 * n * m,  # single-pass CholQR minimum; extra passes count against it
            "lincomb3": 40.0 * n * m,  # LOBPCG: 3 sources read once, 2 results written (kernels_lobpcg.cuh)
        }
        Z = E.MultiVector(ctx, n, m)
        Z.copy_from(X)
        C3 = [rng.standard_normal((m, m)) / m for _ in range(3)]
        calls = {
            "spmm": (lambda: E.matmul_sparse_tallskinny_blocked(Y, dA, X), "spmm"),
            "spmm+dot": (lambda: E.matmul_sparse_tallskinny_with_dots(Y, dA, X), "spmm"),
            "dot": (lambda: E.dot_products_diagonal_blocked(X, Y), "dot"),
            "gram_xx": (lambda: E.dot_products_all_blocked(X, X), "gram"),
            "gram_xy": (lambda: E.dot_products_all_blocked(X, Y), "gram"),
            "update": (lambda: E.block_update(Y, R), "update"),
            "ortho": (lambda: E.orthonormalize_blocked(Y), None),
            "lincomb3": (lambda: E.block_lincomb(Z, [Z, X, Y], C3, Y), "update"),
        }
        E.matmul_sparse_tallskinny_blocked(Y, dA, X)
        for name, (fn, cat) in calls.items():
            for _ in range(3):
                fn()
            ctx.profile(reset=True)
            ctx.set_profiling(True)
            for _ in range(args.reps):
                fn()
            prof = ctx.profile(reset=True)
            ctx.set_profiling(False)
            if cat is None:
                ms = sum(v[0] for v in prof.values()) / args.reps
            else:
                ms = prof[cat][0] / max(prof[cat][1], 1)
            gbs = bytes_[name] / (ms * 1e-3) / 1e9
            rows.append((m, name, ms, gbs, gbs / peak))
            print("%4d %-10s %10.4f %10.1f %8.3f" % rows[-1])
        X.close()
        Y.close()
        Z.close()
    if args.csv:
        with open(args.csv, "w") as f:
            f.write("grid,stencil,n,nnz,m,kernel,avg_ms,GBps,frac_of_measured_hbm_peak,frac_of_fp64_tensor_peak\n")
            for r in rows:
                fl = tensor_flops(r[1], n, r[0])
                ft = "%.4f" % (fl / (r[2] * 1e-3) / 1e12 / FP64_TENSOR_TFLOPS) if fl else ""
                f.write("%d,%s,%d,%d,%d,%s,%.5f,%.1f,%.4f,%s\n" % (args.grid, args.stencil, n, nnz, *r, ft))


if __name__ == "__main__":
    main()
