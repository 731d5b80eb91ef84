"""Row-partitioned multi-GPU plumbing (new; the reference is single-threaded).

One process per GPU. The global matrix is split into contiguous 1-D row blocks (a z-slab for lexicographic 3D
grids); each rank keeps its CSR rows with columns renumbered to [owned | halo]. torch.distributed is used ONLY
for setup plumbing: exchanging the halo index lists and broadcasting the NCCL unique id. The data path --
halo rows of the vector block over NVLink, all-reduce of the m and m x m reductions -- runs inside the C library
on its own NCCL communicator.
"""
import numpy as np

from . import capi
from .capi import check, i64, i64ptr


def partition_rows(n, nranks, align=1):
    """Contiguous row partition offsets part[0..nranks]; `align` keeps cuts on multiples (e.g. a grid plane)."""
    units = n // align
    base, extra = divmod(units, nranks)
    part = [0]
    for r in range(nranks):
        part.append(part[-1] + (base + (1 if r < extra else 0)) * align)
    part[-1] = n
    return np.asarray(part, dtype=np.int64)


def halo_plan_local(rowptr, col_global, part, rank):
    """Host-only: renumber this rank's columns to [owned | halo] and list the halo rows per owner.
    Returns (col_local, halo_global, recv_counts[nranks])."""
    rp, cg, part = i64(rowptr), i64(col_global), i64(part)
    nranks = len(part) - 1
    n_owned = len(rp) - 1
    col_local = np.empty(len(cg), dtype=np.int64)
    halo = np.empty(max(len(cg), 1), dtype=np.int64)
    n_halo = np.zeros(1, dtype=np.int64)
    recv = np.zeros(nranks, dtype=np.int64)
    check(capi.lib().de_halo_plan_local(n_owned, i64ptr(rp), i64ptr(cg), nranks, rank, i64ptr(part), i64ptr(col_local),
                                        i64ptr(halo), i64ptr(n_halo), i64ptr(recv)))
    return col_local, halo[:int(n_halo[0])].copy(), recv


def exchange_send_lists(halo_global, recv_counts, part, rank, dist=None):
    """Tell every owner which of its rows this rank needs (torch.distributed all_to_all of index lists).
    Returns send_rows_per_peer: dict owner_rank -> local row indices (in the requester's halo order) that THIS
    rank must send to that peer."""
    import torch
    if dist is None:
        import torch.distributed as dist
    nranks = len(part) - 1
    counts_out = torch.tensor([int(c) for c in recv_counts], dtype=torch.int64)
    counts_in = torch.zeros(nranks, dtype=torch.int64)
    dev = None
    if dist.get_backend() == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        counts_out, counts_in = counts_out.to(dev), counts_in.to(dev)
    dist.all_to_all_single(counts_in, counts_out)
    counts_in_l = [int(c) for c in counts_in.cpu()]
    offs = np.concatenate([[0], np.cumsum(recv_counts)]).astype(np.int64)
    out_lists = [torch.from_numpy(np.ascontiguousarray(halo_global[offs[p]:offs[p + 1]] - part[p])) for p in range(nranks)]
    in_lists = [torch.empty(c, dtype=torch.int64) for c in counts_in_l]
    if dev is not None:
        out_lists = [t.to(dev) for t in out_lists]
        in_lists = [t.to(dev) for t in in_lists]
    if dist.get_backend() == "gloo":
        # gloo has no all_to_all for ragged lists on every build: use pairwise send/recv
        reqs = []
        for p in range(nranks):
            if p == rank:
                continue
            if len(out_lists[p]) > 0:
                reqs.append(dist.isend(out_lists[p], p))
            if counts_in_l[p] > 0:
                reqs.append(dist.irecv(in_lists[p], p))
        for r in reqs:
            r.wait()
    else:
        dist.all_to_all(in_lists, out_lists)
    return {p: in_lists[p].cpu().numpy().astype(np.int64) for p in range(nranks) if p != rank and counts_in_l[p] > 0}


def build_distributed_matrix(ctx, rowptr, col_global, val, part, rank, dist=None):
    """Create this rank's device matrix from its global-index CSR row block."""
    from .eigensolver import Matrix

    col_local, halo_global, recv_counts = halo_plan_local(rowptr, col_global, part, rank)
    send_lists = exchange_send_lists(halo_global, recv_counts, part, rank, dist)
    nranks = len(part) - 1
    peers = sorted(set(send_lists.keys()) | {p for p in range(nranks) if recv_counts[p] > 0})
    recv = [int(recv_counts[p]) for p in peers]
    send_offsets, send_rows = [0], []
    for p in peers:
        lst = send_lists.get(p, np.zeros(0, dtype=np.int64))
        send_rows.append(lst)
        send_offsets.append(send_offsets[-1] + len(lst))
    send_rows = np.concatenate(send_rows) if send_rows else np.zeros(0, dtype=np.int64)
    n_owned = len(rowptr) - 1
    return Matrix.distributed(ctx, n_owned, len(halo_global), rowptr, col_local, val, peers, recv, send_offsets,
                              send_rows)


def init_comm(ctx, dist=None):
    """Create the library's NCCL communicator: rank 0 makes the unique id, torch.distributed broadcasts it."""
    import torch
    if dist is None:
        import torch.distributed as dist
    from .eigensolver import comm_unique_id

    rank, nranks = dist.get_rank(), dist.get_world_size()
    if nranks == 1:
        return
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        buf = buf.to(torch.device("cuda", torch.cuda.current_device()))
    dist.broadcast(buf, 0)
    ctx.init_comm(rank, nranks, bytes(buf.cpu().numpy().tobytes()))
