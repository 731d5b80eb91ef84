#!/usr/bin/env python
"""Block-width sweep of the hot-path kernels on a (row-partitioned) matrix -- BASELINE.json configs[4]: "p = 8/16/32/64
SpMM + Gram kernels on 3D 200^3 at 1/2/4/8 GPUs vs HBM roofline". Distributed SpMM (halo rows as NVLink peer stores +
interior / boundary launches), all-reduced Gram, two-operand Gram, fused SpMM + Rayleigh quotients and the complete
orthonormalisation, as aggregate ALGORITHMIC GB/s over all ranks (SURVEY.md §8d bytes / max-over-ranks time) against
N x the measured HBM peak, with the kernel time (CUDA events on the launching stream) and the halo wait kept apart from
the wall time of the call (which also holds launch gaps and the wait for the slowest rank).

Stand-alone:  python tools/mg_sweep.py --grid 200 --stencil q1                       (one GPU)
              python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/mg_sweep.py --grid 200
bench.py attaches the same table as the `c5` object of its JSON line at every N.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sweep(ctx, dist, rank, world, grid, stencil, cols=(8, 16, 32, 64), reps=10, peak=6529.1, log=None):
    """-> list of row dicts (identical on every rank). `dist` is torch.distributed (None when world == 1)."""
    import torch

    from dune_eigensolver_b200 import eigensolver as E, matrices as M, parallel as P

    N = grid
    n = N ** 3
    part = P.partition_rows(n, world, align=N * N)
    r0, r1 = int(part[rank]), int(part[rank + 1])
    gen = M.laplacian_fd if stencil == "fd" else M.q1_stiffness
    rp, cg, v = gen((N, N, N), rows=(r0, r1)) if world > 1 else gen((N, N, N))
    nnz_local = len(cg)
    dA = P.build_distributed_matrix(ctx, rp, cg, v, part, rank, dist) if world > 1 else E.Matrix(ctx, (rp, cg, v))
    del rp, cg, v
    nnz = nnz_local
    if world > 1:
        t = torch.tensor([nnz_local], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        nnz = int(t.item())
    nl = r1 - r0
    rows = []

    def maxr(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for m in cols:
        X = E.MultiVector(ctx, nl, m)
        X.upload_rowmajor(np.random.default_rng(100 * m + rank).standard_normal((nl, m)))
        Y = E.MultiVector(ctx, nl, m)
        spmm_bytes = 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n * m
        calls = [
            ("spmm", lambda: E.matmul_sparse_tallskinny_blocked(Y, dA, X), spmm_bytes, ("spmm", "spmm_boundary")),
            ("spmm+dot", lambda: E.matmul_sparse_tallskinny_with_dots(Y, dA, X), spmm_bytes, ("spmm", "spmm_boundary")),
            ("gram_xx", lambda: E.dot_products_all_blocked(X, X), 8.0 * n * m, ("gram",)),
            ("gram_xy", lambda: E.dot_products_all_blocked(X, Y), 16.0 * n * m, ("gram",)),
            ("ortho", lambda: E.orthonormalize_blocked(Y), 24.0 * n * m, ("gram", "update")),
        ]
        E.matmul_sparse_tallskinny_blocked(Y, dA, X)
        for name, fn, nbytes, cats in calls:
            for _ in range(2):
                fn()
            if world > 1:
                dist.barrier()
            ctx.synchronize()
            ctx.profile(reset=True)
            ctx.set_profiling(True)
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            ctx.synchronize()
            wall = (time.perf_counter() - t0) / reps
            prof = ctx.profile(reset=True)
            ctx.set_profiling(False)
            kern = sum(prof[c][0] for c in cats) / reps * 1e-3
            halo_wait = prof["halo_wait"][0] / reps * 1e-3
            halo_push = prof["halo_push"][0] / reps * 1e-3
            wall, kern, halo_wait, halo_push = maxr(wall), maxr(kern), maxr(halo_wait), maxr(halo_push)
            row = {"gpus": world, "grid": N, "stencil": stencil, "m": m, "kernel": name,
                   "kernel_ms": kern * 1e3, "halo_wait_ms": halo_wait * 1e3, "halo_push_ms": halo_push * 1e3,
                   "wall_ms": wall * 1e3, "algorithmic_bytes": nbytes,
                   "GBps_kernel": nbytes / kern / 1e9 if kern > 0 else 0.0,
                   "frac_kernel": nbytes / kern / 1e9 / (world * peak) if kern > 0 else 0.0,
                   "frac_wall": nbytes / wall / 1e9 / (world * peak)}
            rows.append(row)
            if log and rank == 0:
                log("%d GPUs %s %d^3 m=%2d %-9s kernel %8.4f ms (%.3f of %d x %.0f GB/s)  halo wait %.4f ms  wall %8.4f ms (%.3f)" %
                    (world, stencil, N, m, name, kern * 1e3, row["frac_kernel"], world, peak, halo_wait * 1e3, wall * 1e3,
                     row["frac_wall"]))
        X.close()
        Y.close()
    dA.close()
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=200)
    ap.add_argument("--stencil", default="q1", choices=["fd", "q1"])
    ap.add_argument("--cols", default="8,16,32,64")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--csv", default=None)
    args = ap.parse_args()
    import torch

    from dune_eigensolver_b200 import eigensolver as E, parallel as P

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = E.Context(local)
    if world > 1:
        P.init_comm(ctx, dist)
    peak = 6650.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    rows = sweep(ctx, dist, rank, world, args.grid, args.stencil, [int(c) for c in args.cols.split(",")], args.reps, peak,
                 log=lambda s: print(s, flush=True))
    if rank == 0 and args.csv:
        keys = ["gpus", "grid", "stencil", "m", "kernel", "kernel_ms", "halo_wait_ms", "halo_push_ms", "wall_ms",
                "algorithmic_bytes", "GBps_kernel", "frac_kernel", "frac_wall"]
        with open(args.csv, "w") as f:
            f.write(",".join(keys) + "\n")
            for r in rows:
                f.write(",".join(("%.5g" % r[k]) if isinstance(r[k], float) else str(r[k]) for k in keys) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
