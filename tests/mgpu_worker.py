"""Worker for tests/test_multigpu_gpu.py: one process per GPU (torchrun). Row-partitioned SpMM with halo exchange,
all-reduced Gram / dots, CholQR2 and a complete StandardLargest solve, each compared with the single-process oracle;
LOBPCG (new driver) against a dense solve of the global matrix."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from dune_eigensolver_b200 import eigensolver as E, matrices as M, parallel as P
    from oracle import oracle as O

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = E.Context(local)
    P.init_comm(ctx, dist)
    orc = O.load_best()
    fails = []

    def gather_rows(local_rows, part):
        """all ranks' row blocks -> global array on every rank"""
        n = int(part[-1])
        out = torch.zeros((n, local_rows.shape[1]), dtype=torch.float64, device="cuda")
        out[int(part[rank]):int(part[rank + 1])] = torch.from_numpy(local_rows).cuda()
        dist.all_reduce(out)
        return out.cpu().numpy()

    for shape, kind, m in [((6, 5, 8), "fd", 16), ((7, 6, 9), "q1", 32), ((5, 5, 6), "q1", 8)]:
        gen = M.laplacian_fd if kind == "fd" else M.q1_stiffness
        n = int(np.prod(shape))
        plane = int(np.prod(shape[:-1]))
        part = P.partition_rows(n, world, align=plane)
        r0, r1 = int(part[rank]), int(part[rank + 1])
        rp, cg, v = gen(shape, rows=(r0, r1))
        dA = P.build_distributed_matrix(ctx, rp, cg, v, part, rank, dist)
        Aglob = gen(shape)
        Xg = np.random.default_rng(3).standard_normal((n, m))
        dX = E.MultiVector.from_array(ctx, Xg[r0:r1])
        dY = E.MultiVector(ctx, r1 - r0, m)
        # SpMM with halo exchange + fused all-reduced dots, both kernel families (tensor-core BRB tiles / CSR rows)
        ref = orc.spmm(Aglob, Xg)
        dref = orc.diag_dot(Xg, ref)
        # every rank must make the same sequence of collective calls: ranks whose local block has no BRB form (e.g.
        # no rows at all when there are fewer grid planes than ranks) run the second round with their only kernel family
        for fmt in ("csr", "brb"):
            dA.set_spmm_format(fmt if (fmt == "csr" or dA.spmm_info()["tiles"] > 0) else "auto")
            dY.upload(np.zeros((r1 - r0, m)))
            dp = E.matmul_sparse_tallskinny_with_dots(dY, dA, dX)
            Yg = gather_rows(dY.download(), part)
            if np.abs(Yg - ref).max() > 1e-12 * np.abs(ref).max():
                fails.append(("spmm", fmt, shape, float(np.abs(Yg - ref).max())))
            if np.abs(dp - dref).max() > 1e-11 * np.abs(dref).max():
                fails.append(("dots", fmt, shape, float(np.abs(dp - dref).max())))
        dA.set_spmm_format("auto")
        # Gram with all-reduce
        G = E.dot_products_all_blocked(dX, dY)
        if np.abs(G - Xg.T @ ref).max() > 1e-11 * np.abs(Xg.T @ ref).max():
            fails.append(("gram", shape))
        # CholQR2 across ranks
        E.orthonormalize_blocked(dX)
        Qg = gather_rows(dX.download(), part)
        qref = orc.orthonormalize(Xg)
        if np.abs(Qg - qref).max() > 1e-10 * np.linalg.cond(Xg):
            fails.append(("ortho", shape, float(np.abs(Qg - qref).max())))
        # complete solve, device resident
        nev = m
        start = E.from_panels(E.start_block(n, m, 123), n, m)
        Q = E.MultiVector(ctx, r1 - r0, m)
        Q.upload_rowmajor(np.ascontiguousarray(start[r0:r1]))
        ev, it = E.standard_largest_mv(ctx, dA, 0.0, 1e-9, 3000, Q)
        evr, Vr, k = orc.standard_largest(Aglob, 0.0, 1e-9, 3000, nev)
        if abs(it - k) > 1 or np.abs(ev - evr).max() > 1e-8 * np.abs(evr).max():
            fails.append(("largest", shape, it, k, float(np.abs(ev - evr).max())))
        # LOBPCG across ranks (plain and with the Chebyshev polynomial preconditioner): the Rayleigh-Ritz problem is
        # solved redundantly on every rank from bit-identical all-reduced Gram matrices, so all ranks take the same
        # decisions; smallest eigenvalues against a dense solve of the global matrix
        dense = np.linalg.eigvalsh(M.to_scipy(Aglob).toarray())[:m]
        for deg in ((0, 6) if world <= shape[-1] else ()):  # every rank owns at least one grid plane
            Q.upload_rowmajor(np.ascontiguousarray(start[r0:r1]))
            lam, rn, it2, rs, conv = E.lobpcg_mv(ctx, dA, Q, 1e-9, 2000, nev=m, cheb_degree=deg)
            if not conv or np.abs(lam - dense).max() > 1e-10 * np.abs(dense).max():
                fails.append(("lobpcg", deg, shape, it2, bool(conv), float(np.abs(lam - dense).max())))
    t = torch.tensor([len(fails)], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("MGPU_RESULT world=%d fails=%d %s" % (world, int(t.item()), fails))
    dist.destroy_process_group()
    sys.exit(1 if int(t.item()) else 0)


if __name__ == "__main__":
    main()
