#!/usr/bin/env python
"""Diagnostic (torchrun, one rank per GPU): per-launch SpMM time of a row-partitioned matrix, and of the same local
grid as a stand-alone single-GPU problem."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def main():
    import torch, torch.distributed as dist
    from dune_eigensolver_b200 import eigensolver as E, matrices as M, parallel as P
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = E.Context(local)
    P.init_comm(ctx, dist)
    N, m = 100, 32
    n = N ** 3
    part = P.partition_rows(n, world, align=N * N)
    r0, r1 = int(part[rank]), int(part[rank + 1])
    rp, cg, v = M.q1_stiffness((N, N, N), rows=(r0, r1))
    dA = P.build_distributed_matrix(ctx, rp, cg, v, part, rank, dist)
    X = E.MultiVector(ctx, r1 - r0, m); X.upload_rowmajor(np.random.default_rng(rank).standard_normal((r1 - r0, m)))
    Y = E.MultiVector(ctx, r1 - r0, m)
    def timed(fn, reps=20):
        for _ in range(3): fn()
        ctx.profile(reset=True); ctx.set_profiling(True)
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): fn()
        ctx.synchronize(); wall = (time.perf_counter() - t0) / reps
        prof = ctx.profile(reset=True); ctx.set_profiling(False)
        return wall, {k: (round(a / max(c, 1) * 1e3, 1), c) for k, (a, c) in prof.items() if c}
    for fmt in ("brb", "csr"):
        dA.set_spmm_format(fmt)
        w, p = timed(lambda: E.matmul_sparse_tallskinny_blocked(Y, dA, X))
        print("rank %d distributed %s: wall %.1f us per SpMM; per-launch us %s; %s" % (rank, fmt, w * 1e6, p, dA.spmm_info()), flush=True)
    dist.barrier()
    # the same number of rows as a stand-alone problem on this GPU (no halo)
    A1 = M.q1_stiffness((N, N, (r1 - r0) // (N * N)))
    d1 = E.Matrix(ctx, A1)
    w, p = timed(lambda: E.matmul_sparse_tallskinny_blocked(Y, d1, X))
    print("rank %d stand-alone %dx%dx%d: wall %.1f us per SpMM; per-launch us %s" % (rank, N, N, (r1 - r0) // (N * N), w * 1e6, p), flush=True)
    G = lambda: E.dot_products_all_blocked(X, X)
    w, p = timed(G)
    print("rank %d gram_xx (all-reduced): wall %.1f us; %s" % (rank, w * 1e6, p), flush=True)
    dist.barrier()
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
