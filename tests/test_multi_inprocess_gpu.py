"""The row-partitioned data path through the single-process multi-GPU front end of the C ABI (de_multi_*): halo rows as
peer stores, one-shot peer all-reduce fused into the reduction tails, replicated Cholesky / convergence decision.

A box with fewer GPUs than ranks lists ordinal 0 several times: the ranks then are independent contexts (own streams,
own windows) on the same device, which exercises exactly the same kernels and flow control -- so these tests do NOT skip
on a one-GPU box. Checked against the reference itself: small grids through the oracle, the bench workload (100^3,
m = 32) against tests/golden/reference_fullsize.npz (41 iterations +-1, eigenvalues to 1e-10)."""
import os

import numpy as np
import pytest

from dune_eigensolver_b200 import eigensolver as E, matrices as M

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _devices(ranks):
    import torch

    have = torch.cuda.device_count()
    return [r % max(have, 1) for r in range(ranks)]


@pytest.mark.parametrize("ranks", [2, 3, 4])
@pytest.mark.parametrize("shape,kind,nev", [((6, 5, 8), "fd", 16), ((7, 6, 9), "q1", 32), ((5, 5, 6), "q1", 8)])
def test_partitioned_standard_largest_matches_the_reference(oracle, ranks, shape, kind, nev):
    gen = M.laplacian_fd if kind == "fd" else M.q1_stiffness
    A = gen(shape)
    plane = int(np.prod(shape[:-1]))
    mg = E.Multi(_devices(ranks), timeout_s=5)
    try:
        r = mg.StandardLargest((A[0], A[1], A[2].copy()), 0.0, 1e-9, 3000, nev, row_align=plane)
    finally:
        mg.close()
    evr, Vr, k = oracle.standard_largest((A[0], A[1], A[2].copy()), 0.0, 1e-9, 3000, nev)
    assert abs(r.iterations - k) <= 1, (r.iterations, k)
    assert np.abs(r.eval - evr).max() <= 1e-8 * np.abs(evr).max()
    S = M.to_scipy(A)
    V = np.asarray(r.evec)
    assert np.abs(V @ V.T - np.eye(nev)).max() <= 1e-11
    res = np.array([np.linalg.norm(S @ V[j] - r.eval[j] * V[j]) for j in range(nev)])
    res_ref = np.array([np.linalg.norm(S @ Vr[j] - evr[j] * Vr[j]) for j in range(nev)])
    assert (res <= 2.0 * res_ref + 1e-8).all()


def test_halo_rows_stored_by_the_update_kernel(oracle, monkeypatch):
    """DE_B200_FUSED_PUSH=1 (off by default, see de_internal.hpp): the block-update kernels of the orthonormalisation store
    the halo rows of the next SpMM into the neighbours' windows and the last of them raises the flags -- same result"""
    monkeypatch.setenv("DE_B200_FUSED_PUSH", "1")
    shape = (7, 6, 9)
    A = M.q1_stiffness(shape)
    mg = E.Multi(_devices(3), timeout_s=5)
    try:
        r = mg.StandardLargest((A[0], A[1], A[2].copy()), 0.0, 1e-9, 3000, 32, row_align=shape[0] * shape[1])
    finally:
        mg.close()
    evr, _, k = oracle.standard_largest((A[0], A[1], A[2].copy()), 0.0, 1e-9, 3000, 32)
    assert abs(r.iterations - k) <= 1
    assert np.abs(r.eval - evr).max() <= 1e-8 * np.abs(evr).max()


def test_unaligned_partition_and_shift(oracle):
    """cuts inside grid planes (row_align = 1): halo lists are no longer whole planes; shift != 0 (eigensolver.hh:57-66)"""
    A = M.q1_stiffness((5, 6, 7))
    mg = E.Multi(_devices(3), timeout_s=5)
    try:
        r = mg.StandardLargest((A[0], A[1], A[2].copy()), 0.25, 1e-9, 3000, 8)
    finally:
        mg.close()
    evr, _, k = oracle.standard_largest((A[0], A[1], A[2].copy()), 0.25, 1e-9, 3000, 8)
    assert abs(r.iterations - k) <= 1
    assert np.abs(r.eval - evr).max() <= 1e-8 * np.abs(evr).max()


@pytest.mark.parametrize("ranks", [2, 4])
def test_partitioned_lobpcg(ranks):
    """LOBPCG copies small coefficient matrices host -> device in front of its kernels every iteration. With several ranks
    on ONE device those copies share the device's copy-engine queues: a copy queued behind another rank's device -> host
    copy, which itself waits for a kernel spinning on this rank's contribution, never runs (measured with
    tools/micro/peer_ar_probe.cu: 3+ ranks on one device deadlock, one rank per device does not). So this test needs one
    GPU per rank; the StandardLargest loop has no copy in front of its kernels and is tested on shared devices above."""
    import torch

    if torch.cuda.device_count() < ranks:
        pytest.skip("needs one GPU per rank (copy-engine queues are shared between ranks on one device)")
    shape = (7, 6, 9)
    A = M.q1_stiffness(shape)
    dense = np.linalg.eigvalsh(M.to_scipy(A).toarray())[:8]
    mg = E.Multi(_devices(ranks), timeout_s=5)
    try:
        r = mg.StandardLOBPCG(A, 1e-9, 2000, 8, row_align=shape[0] * shape[1])
    finally:
        mg.close()
    assert np.abs(r.eval - dense).max() <= 1e-9 * np.abs(dense).max()


@pytest.mark.parametrize("ranks", [2, 4])
def test_bench_workload_partitioned_matches_the_reference_at_full_size(ranks):
    """configs[1] at full size, row-partitioned: the reference's own run is the fixture (41 iterations)"""
    full = np.load(os.path.join(ROOT, "tests", "golden", "reference_fullsize.npz"))
    N, nev, tol = int(full["q1_N"]), int(full["q1_nev"]), float(full["q1_tol"])
    A = M.q1_stiffness((N, N, N))
    mg = E.Multi(_devices(ranks), timeout_s=5)
    try:
        r = mg.StandardLargest((A[0], A[1], A[2].copy()), 0.0, tol, 4000, nev, row_align=N * N)
        launches = mg.launch_count()
    finally:
        mg.close()
    assert launches > 0
    k_ref, ev_ref = int(full["q1_iterations"]), full["q1_eval"]
    assert abs(r.iterations - k_ref) <= 1, (r.iterations, k_ref)
    scale = np.abs(ev_ref).max()
    if r.iterations == k_ref:
        assert np.abs(r.eval - ev_ref).max() <= 1e-10 * scale
        V = np.asarray(r.evec)
        assert np.abs(V[:, full["q1_sample_idx"]] - full["q1_sample"]).max() <= 1e-8
    else:
        assert np.abs(r.eval - ev_ref).max() <= tol * scale
    S = M.to_scipy(A)
    V = np.asarray(r.evec)
    res = np.array([np.linalg.norm(S @ V[j] - r.eval[j] * V[j]) for j in range(nev)])
    assert (res <= 2.0 * full["q1_residual"] + 1e-9).all()
    assert np.abs(V @ V.T - np.eye(nev)).max() <= 1e-12
