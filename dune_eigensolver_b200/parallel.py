"""Row-partitioned multi-GPU plumbing (new; the reference is single-threaded).

One process per GPU. The global matrix is split into contiguous 1-D row blocks (a z-slab for lexicographic 3D
grids); each rank keeps its CSR rows with columns renumbered to [owned | halo]. torch.distributed is used ONLY
for setup plumbing: exchanging the halo index lists and broadcasting the NCCL unique id. The data path --
halo rows of the vector block over NVLink, all-reduce of the m and m x m reductions -- runs inside the C library:
over NVLink peer memory (every rank's window is mapped into every peer through CUDA IPC handles exchanged here) and,
where that is not available, on the library's own NCCL communicator.
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


def allgather_bytes(dist=None):
    """all-gather of equal-length byte strings over torch.distributed, in the form Matrix.rowblock wants"""
    import torch
    if dist is None:
        import torch.distributed as dist

    def ag(send):
        t = torch.frombuffer(bytearray(send), dtype=torch.uint8)
        dev = None
        if dist.get_backend() == "nccl":
            dev = torch.device("cuda", torch.cuda.current_device())
            t = t.to(dev)
        out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        return b"".join(bytes(o.cpu().numpy().tobytes()) for o in out)

    return ag


def build_distributed_matrix(ctx, rowptr, col_global, val, part, rank, dist=None):
    """Create this rank's device matrix from its global-index CSR row block. Halo planning, the exchange of the halo
    lists (two all-gathers) and the peer-deposit offsets are computed inside the library (de_matrix_create_rowblock);
    torch.distributed only carries the all-gather."""
    from .eigensolver import Matrix

    assert ctx.rank()[0] == rank or ctx.rank()[1] == 1
    return Matrix.rowblock(ctx, rowptr, col_global, val, part, allgather_bytes(dist))


def peer_deposit_offsets(recv_counts, peers, rank, dist=None):
    """all-gather every rank's recv_counts[nranks]; deposit[p] = sum_{q < rank} recv_counts_of_peer_p[q].
    Also returns the largest halo block (rows) over all ranks."""
    import torch
    if dist is None:
        import torch.distributed as dist
    nranks = dist.get_world_size()
    mine = torch.tensor([int(c) for c in recv_counts], dtype=torch.int64)
    dev = None
    if dist.get_backend() == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        mine = mine.to(dev)
    allc = [torch.zeros_like(mine) for _ in range(nranks)]
    dist.all_gather(allc, mine)
    allc = [t.cpu().numpy() for t in allc]
    return (np.asarray([int(allc[p][:rank].sum()) for p in peers], dtype=np.int64),
            max(int(c.sum()) for c in allc))


def init_comm(ctx, dist=None, peer_window_bytes=128 << 20):
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
    if peer_window_bytes and dist.get_backend() == "nccl" and 2 <= nranks <= 8:
        init_peer_window(ctx, dist, peer_window_bytes)


def init_peer_window(ctx, dist, halo_bytes):
    """NVLink fast path: create this rank's window, all-gather the CUDA IPC handles, map the peers' windows.
    Every rank must succeed, otherwise all stay on NCCL (the decision is all-reduced)."""
    import torch

    dev = torch.device("cuda", torch.cuda.current_device())
    nranks = dist.get_world_size()
    ok = 1
    try:
        handle = ctx.peer_window_create(halo_bytes)
    except capi.DeError:
        handle, ok = bytes(64), 0
    mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).clone().to(dev)
    allh = [torch.zeros_like(mine) for _ in range(nranks)]
    dist.all_gather(allh, mine)
    flag = torch.tensor([ok], dtype=torch.int64, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        return False
    handles = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh)
    # opening must succeed everywhere before anyone uses the windows
    try:
        ctx_ok = 1
        ctx._peer_handles = handles
        capi.check(capi.lib().de_context_peer_window_open(ctx._h, (capi.C.c_char * len(handles)).from_buffer_copy(handles)),
                   ctx._h)
    except capi.DeError:
        ctx_ok = 0
    flag = torch.tensor([ctx_ok], dtype=torch.int64, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    if int(flag.item()) == 0 and ctx_ok:
        raise RuntimeError("NVLink peer window: mapped on this rank but not on all ranks; cannot continue consistently")
    return bool(int(flag.item()))
