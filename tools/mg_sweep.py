#!/usr/bin/env python
"""Block-width sweep of the hot-path kernels on a ROW-PARTITIONED matrix (BASELINE.json config 5 at 2/4/8 GPUs):
distributed SpMM (halo exchange + interior/boundary launches), all-reduced Gram and the complete orthonormalisation,
as aggregate algorithmic GB/s over all ranks against N x the measured HBM peak. Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/mg_sweep.py --grid 200 --stencil q1
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=200)
    ap.add_argument("--stencil", default="q1", choices=["fd", "q1"])
    ap.add_argument("--cols", default="8,16,32,64")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--csv", default=None)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from dune_eigensolver_b200 import eigensolver as E, matrices as M, parallel as P

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = E.Context(local)
    P.init_comm(ctx, dist)
    peak = 6650.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    N = args.grid
    n = N ** 3
    part = P.partition_rows(n, world, align=N * N)
    r0, r1 = int(part[rank]), int(part[rank + 1])
    gen = M.laplacian_fd if args.stencil == "fd" else M.q1_stiffness
    rp, cg, v = gen((N, N, N), rows=(r0, r1))
    nnz_local = len(cg)
    dA = P.build_distributed_matrix(ctx, rp, cg, v, part, rank, dist)
    del rp, cg, v
    t = torch.tensor([nnz_local], dtype=torch.int64, device="cuda")
    dist.all_reduce(t)
    nnz = int(t.item())
    rows = []
    if rank == 0:
        print("grid %d^3 %s n=%d nnz=%d on %d GPUs (%s), peak %d x %.0f GB/s" %
              (N, args.stencil, n, nnz, world, "NVLink peer memory" if ctx.peer_ready() else "NCCL", world, peak), flush=True)
    for m in [int(c) for c in args.cols.split(",")]:
        X = E.MultiVector(ctx, r1 - r0, m)
        X.upload_rowmajor(np.random.default_rng(100 * m + rank).standard_normal((r1 - r0, m)))
        Y = E.MultiVector(ctx, r1 - r0, m)
        calls = {
            "spmm": (lambda: E.matmul_sparse_tallskinny_blocked(Y, dA, X), 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n * m),
            "spmm+dot": (lambda: E.matmul_sparse_tallskinny_with_dots(Y, dA, X), 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n * m),
            "gram_xx": (lambda: E.dot_products_all_blocked(X, X), 8.0 * n * m),
            "ortho": (lambda: E.orthonormalize_blocked(Y), 24.0 * n * m),
        }
        E.matmul_sparse_tallskinny_blocked(Y, dA, X)
        for name, (fn, nbytes) in calls.items():
            for _ in range(3):
                fn()
            dist.barrier()
            torch.cuda.synchronize()
            ctx.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.reps):
                fn()
            ctx.synchronize()
            dt = torch.tensor([(time.perf_counter() - t0) / args.reps], dtype=torch.float64, device="cuda")
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            ms = float(dt.item()) * 1e3
            gbs = nbytes / (ms * 1e-3) / 1e9
            rows.append((m, name, ms, gbs, gbs / (world * peak)))
            if rank == 0:
                print("%4d %-10s %10.4f ms %10.1f GB/s aggregate  %.3f of %d x peak" % (m, name, ms, gbs, gbs / (world * peak), world),
                      flush=True)
        X.close()
        Y.close()
    if rank == 0 and args.csv:
        with open(args.csv, "w") as f:
            f.write("gpus,grid,stencil,n,nnz,m,kernel,wall_ms_per_call_max_over_ranks,aggregate_GBps,frac_of_N_x_measured_hbm_peak\n")
            for r in rows:
                f.write("%d,%d,%s,%d,%d,%d,%s,%.5f,%.1f,%.4f\n" % (world, N, args.stencil, n, nnz, *r))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
