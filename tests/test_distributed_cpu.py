"""World-size-2 `gloo` test of the multi-GPU HOST logic on CPU (no GPU): row partition, halo plan from the CSR
column indices, exchange of the send lists over torch.distributed, and a numpy emulation of the resulting
halo exchange + local SpMM that must reproduce the global product. The device data path itself (NCCL over
NVLink) is exercised by tests/test_multigpu_gpu.py on the GPU box."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, kind, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from dune_eigensolver_b200 import matrices as M, parallel as P

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = int(np.prod(shape))
        plane = int(np.prod(shape[:-1]))
        part = P.partition_rows(n, world, align=plane)
        gen = M.laplacian_fd if kind == "fd" else M.q1_stiffness
        rp, cg, v = gen(shape, rows=(part[rank], part[rank + 1]))   # each rank builds only its slab
        col_local, halo_global, recv_counts = P.halo_plan_local(rp, cg, part, rank)
        send_lists = P.exchange_send_lists(halo_global, recv_counts, part, rank, dist)
        n_owned = len(rp) - 1
        # --- emulate the device data path with numpy + gloo: pack -> send/recv -> local SpMM over [owned | halo]
        m = 8
        Xglob = np.random.default_rng(5).standard_normal((n, m))
        X = Xglob[part[rank]:part[rank + 1]]
        halo = np.zeros((len(halo_global), m))
        offs = np.concatenate([[0], np.cumsum(recv_counts)])
        reqs, bufs = [], {}
        for p in range(world):
            if p == rank:
                continue
            if p in send_lists:
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(X[send_lists[p]])), p))
            if recv_counts[p] > 0:
                bufs[p] = torch.empty((int(recv_counts[p]), m), dtype=torch.float64)
                reqs.append(dist.irecv(bufs[p], p))
        for r in reqs:
            r.wait()
        for p, b in bufs.items():
            halo[offs[p]:offs[p + 1]] = b.numpy()
        # --- the NVLink peer path deposits rows at offsets every sender computes itself (parallel.peer_deposit_offsets,
        # csrc/kernels_peer.cuh halo_push_kernel): emulate the stores and compare with the halo block built above
        peers = sorted(set(send_lists.keys()) | {p for p in range(world) if recv_counts[p] > 0})
        deposits, max_halo = P.peer_deposit_offsets(recv_counts, peers, rank, dist)
        msgs = [(rank, p, int(deposits[i]), np.ascontiguousarray(X[send_lists[p]])) for i, p in enumerate(peers)
                if p in send_lists]
        allmsgs = [None] * world
        dist.all_gather_object(allmsgs, msgs)
        halo_peer = np.full((len(halo_global), m), np.nan)
        for lst in allmsgs:
            for src, dst, off, rows in lst:
                if dst == rank:
                    halo_peer[off:off + len(rows)] = rows
        peer_ok = bool(np.array_equal(halo_peer, halo)) and max_halo >= len(halo_global)
        # --- the same plan from the C ABI's host-only planner (de_halo_plan_peers: what de_matrix_create_rowblock and
        # the single-process de_multi_* front end use) fed with all-gathered counts and halo lists
        import ctypes as C

        from dune_eigensolver_b200 import capi

        allc = [None] * world
        dist.all_gather_object(allc, [int(c) for c in recv_counts])
        alll = [None] * world
        dist.all_gather_object(alll, [int(g) for g in halo_global])
        stride = max(1, max(len(x) for x in alll))
        counts_all = capi.i64(np.asarray(allc).reshape(-1))
        lists_all = np.full((world, stride), -1, dtype=np.int64)
        for q, lst in enumerate(alll):
            lists_all[q, :len(lst)] = lst
        npeers, sym, mx = C.c_int(0), C.c_int(0), C.c_int64(0)
        pr = np.zeros(world, dtype=np.int32)
        rc, so, dep = (np.zeros(world + 1, dtype=np.int64) for _ in range(3))
        sr = np.zeros(max(1, int(sum(allc[q][rank] for q in range(world) if q != rank))), dtype=np.int64)
        capi.check(capi.lib().de_halo_plan_peers(world, rank, capi.i64ptr(capi.i64(part)), capi.i64ptr(counts_all),
                                                 capi.i64ptr(lists_all), stride, C.byref(npeers), capi.i32ptr(pr),
                                                 capi.i64ptr(rc), capi.i64ptr(so), capi.i64ptr(sr), capi.i64ptr(dep),
                                                 C.byref(mx), C.byref(sym)))
        k = npeers.value
        c_ok = (list(pr[:k]) == peers and list(dep[:k]) == [int(d) for d in deposits] and mx.value == max_halo
                and sym.value == 1 and list(rc[:k]) == [int(recv_counts[p]) for p in peers]
                and all(np.array_equal(sr[so[i]:so[i + 1]], send_lists.get(p, np.zeros(0, dtype=np.int64)))
                        for i, p in enumerate(peers)))
        peer_ok = peer_ok and bool(c_ok)
        import scipy.sparse as sp

        Aloc = sp.csr_matrix((v, col_local, rp), shape=(n_owned, n_owned + len(halo_global)))
        Y = Aloc @ np.vstack([X, halo])
        Aglob = M.to_scipy(gen(shape))
        err = np.abs(Y - (Aglob @ Xglob)[part[rank]:part[rank + 1]]).max()
        # reductions: partial Gram + all-reduce equals the global Gram
        G = torch.from_numpy(X.T @ X)
        dist.all_reduce(G)
        gerr = np.abs(G.numpy() - Xglob.T @ Xglob).max()
        ret[rank] = (float(err), float(gerr), int(len(halo_global)), [int(c) for c in recv_counts], peer_ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape,kind,world", [((6, 5, 8), "fd", 2), ((5, 4, 6), "q1", 2), ((4, 5, 7), "q1", 3)])
def test_row_partition_halo_exchange_gloo(shape, kind, world):
    import torch.multiprocessing as mp

    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, shape, kind, ret), nprocs=world, join=True)
    plane = int(np.prod(shape[:-1]))
    for rank in range(world):
        err, gerr, nhalo, recv, peer_ok = ret[rank]
        assert err < 1e-12 and gerr < 1e-10
        assert peer_ok                              # peer-store deposit offsets reproduce the halo block layout
        neighbours = [p for p in (rank - 1, rank + 1) if 0 <= p < world]
        assert nhalo == plane * len(neighbours)     # a z-slab needs one plane per neighbouring slab
        assert all(recv[p] == plane for p in neighbours) and recv[rank] == 0


def test_partition_rows():
    from dune_eigensolver_b200 import parallel as P

    assert list(P.partition_rows(10, 3)) == [0, 4, 7, 10]
    assert list(P.partition_rows(100, 4, align=25)) == [0, 25, 50, 75, 100]
    p = P.partition_rows(7 * 9, 4, align=9)
    assert p[0] == 0 and p[-1] == 63 and all((b - a) % 9 == 0 for a, b in zip(p[:-1], p[1:]))
