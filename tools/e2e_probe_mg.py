#!/usr/bin/env python
"""Where does the end-to-end time of a row-partitioned StandardLargest call go? (torchrun, one rank per GPU;
DE_TRACE_SETUP=1 prints the library's own setup laps)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from dune_eigensolver_b200 import eigensolver as E, matrices as M, parallel as P

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = E.Context(local)
if world > 1:
    P.init_comm(ctx, dist)
N, nev = 100, 32
n, m = N ** 3, 32
part = P.partition_rows(n, world, align=N * N)
r0, r1 = int(part[rank]), int(part[rank + 1])
rp, ci, v = M.q1_stiffness((N, N, N), rows=(r0, r1)) if world > 1 else M.q1_stiffness((N, N, N))
start = E.to_panels(np.ascontiguousarray(E.from_panels(E.start_block(n, m, 123), n, m)[r0:r1]))
for it in range(4):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mat = P.build_distributed_matrix(ctx, rp, ci, v, part, rank, dist) if world > 1 else E.Matrix(ctx, (rp, ci, v))
    t1 = time.perf_counter()
    evl, V, k = np.zeros(nev), np.zeros((nev, r1 - r0)), E.C.c_int(0)
    E.check(E.capi.lib().de_standard_largest(ctx._h, mat._h, 0.0, 2e-3, 4000, nev, E.dptr(start), E.dptr(evl), E.dptr(V), 0, E.C.byref(k)), ctx._h)
    t2 = time.perf_counter()
    mat.close()
    t3 = time.perf_counter()
    print("rank %d step %d: matrix %.1f ms, solve call %.1f ms, close %.1f ms, total %.1f ms" %
          (rank, it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3), flush=True)
if world > 1:
    dist.destroy_process_group()
