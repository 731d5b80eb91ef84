"""Host-side checks (no GPU) of the BRB SpMM format builder (dune_eigensolver_b200/csrc/brb_format.hpp) through the
C ABI's de_brb_format_check: every tile is decoded the way the kernel decodes it and applied to a probe vector; the
result must equal the CSR product (reference kernels_cpp.hh:626-657 semantics) up to rounding, every row must be
covered exactly once, and the structured-grid detection must pick box tiles for the BASELINE.json matrices."""
import ctypes as C

import numpy as np
import pytest

from dune_eigensolver_b200 import capi, matrices as M


def check(A, ncols=None, n_owned=None, nthreads=0):
    rp, ci, v = (np.ascontiguousarray(A[0], dtype=np.int64), np.ascontiguousarray(A[1], dtype=np.int64),
                 np.ascontiguousarray(A[2], dtype=np.float64))
    n = len(rp) - 1
    ncols = n if ncols is None else ncols
    n_owned = n if n_owned is None else n_owned
    info = (C.c_int64 * 8)()
    diff = C.c_double(-1.0)
    s = capi.lib().de_brb_format_check(n, ncols, n_owned, capi.i64ptr(rp), capi.i64ptr(ci), capi.dptr(v), nthreads, info,
                                       C.byref(diff))
    assert s == 0, capi.lib().de_last_error_string(None)
    keys = ["ok", "grid", "tiles", "interior", "blocks", "steps", "max_u", "shape"]
    d = dict(zip(keys, [int(x) for x in info]))
    d["shape"] = (d["shape"] & 0xffff, (d["shape"] >> 16) & 0xffff, (d["shape"] >> 32) & 0xffff)
    d["diff"] = diff.value
    d["scale"] = float(np.abs(v).max()) * 2.0 * max(1, int(np.diff(rp).max())) if len(v) else 1.0
    return d


GRIDS = {
    "lap2d_17": (lambda: M.laplacian_dirichlet_2d(17), 2),
    "lap2d_200": (lambda: M.laplacian_dirichlet_2d(200), 2),        # BASELINE config 1
    "fd3d_9x7x5": (lambda: M.laplacian_fd((9, 7, 5)), 3),
    "fd3d_24": (lambda: M.laplacian_fd((24, 24, 24)), 3),
    "q1_3d_8": (lambda: M.q1_stiffness((8, 8, 8)), 3),
    "q1_3d_21x13x10": (lambda: M.q1_stiffness((21, 13, 10)), 3),
    "q1_hc_12": (lambda: M.q1_stiffness((12, 12, 12), kappa=M.high_contrast_kappa(1e6, 2)), 3),
    "q1_mass_9": (lambda: M.q1_mass((9, 9, 9)), 3),
    # a plane of more than 16384 points: the tile order sweeps z inside chunks of y (brb::grid_order), ragged last chunk
    "fd3d_150x131x5": (lambda: M.laplacian_fd((150, 131, 5)), 3),
    "q1_3d_140x122x6": (lambda: M.q1_stiffness((140, 122, 6)), 3),
}


@pytest.mark.parametrize("name", sorted(GRIDS))
def test_grid_matrices_get_box_tiles(name):
    gen, dim = GRIDS[name]
    A = gen()
    d = check(A)
    assert d["ok"] == 1 and d["grid"] == 1, d
    assert d["diff"] <= 1e-13 * d["scale"], d
    assert d["interior"] == d["tiles"]
    tw, th, td = d["shape"]
    assert tw >= 2 and th >= 2 and (td >= 2 if dim == 3 else td == 1), d
    assert d["max_u"] <= 512
    n = len(A[0]) - 1
    assert d["blocks"] >= (n + 7) // 8


def test_threads_do_not_change_the_format():
    A = M.q1_stiffness((13, 11, 9))
    a, b = check(A, nthreads=1), check(A, nthreads=5)
    assert a == b


def _random_csr(n, ncols, per_row, seed, empty_every=0, dup=False):
    rng = np.random.default_rng(seed)
    rp, ci, v = [0], [], []
    for i in range(n):
        if empty_every and i % empty_every == 0:
            rp.append(len(ci))
            continue
        k = int(rng.integers(1, per_row + 1))
        cols = np.sort(rng.choice(ncols, size=k, replace=False))
        if dup and k > 1:
            cols[1] = cols[0]  # duplicate entry: must accumulate
        ci.extend(cols.tolist())
        v.extend(rng.standard_normal(k).tolist())
        rp.append(len(ci))
    return np.array(rp), np.array(ci), np.array(v)


def test_unstructured_matrix_uses_consecutive_rows():
    A = _random_csr(700, 700, 9, 1, empty_every=13)
    d = check(A)
    assert d["ok"] == 1 and d["grid"] == 0, d
    assert d["diff"] <= 1e-13 * d["scale"]
    assert d["blocks"] == (700 + 7) // 8


def test_duplicate_entries_accumulate():
    A = _random_csr(300, 300, 6, 2, dup=True)
    d = check(A)
    assert d["ok"] == 1
    assert d["diff"] <= 1e-13 * d["scale"]


def test_banded_matrix():
    n, bw = 1000, 11
    rp, ci, v = [0], [], []
    for i in range(n):
        for j in range(max(0, i - bw), min(n, i + bw + 1)):
            ci.append(j)
            v.append(1.0 / (1 + abs(i - j)))
        rp.append(len(ci))
    d = check((np.array(rp), np.array(ci), np.array(v)))
    assert d["ok"] == 1 and d["diff"] <= 1e-13 * d["scale"], d


def test_rows_too_wide_for_a_tile_have_no_brb_form():
    # 8 rows x 200 random columns out of 20000: the union of one row block exceeds the 512 rows a tile can stage
    A = _random_csr(64, 20000, 200, 3)
    A = (A[0], A[1], A[2])
    rp = A[0]
    if np.diff(rp).max() < 150:
        pytest.skip("generator produced short rows")
    d = check(A, ncols=20000)
    assert d["ok"] == 0


def test_local_block_of_a_distributed_matrix_separates_boundary_tiles():
    """rows of a z-slab of a 12 x 12 x 20 grid with columns renumbered [owned | halo] (what de_halo_plan_local does)"""
    N = 12
    rp, ci, v = M.q1_stiffness((N, N, 20))
    r0, r1 = 2 * N * N, 18 * N * N
    n_owned = r1 - r0
    lrp, lci, lv, halo = [0], [], [], {}
    for r in range(r0, r1):
        for k in range(rp[r], rp[r + 1]):
            c = int(ci[k])
            if r0 <= c < r1:
                lci.append(c - r0)
            else:
                lci.append(n_owned + halo.setdefault(c, len(halo)))
            lv.append(v[k])
        lrp.append(len(lci))
    d = check((np.array(lrp), np.array(lci), np.array(lv)), ncols=n_owned + len(halo), n_owned=n_owned)
    assert d["ok"] == 1 and d["grid"] == 1, d
    assert d["diff"] <= 1e-13 * d["scale"]
    assert 0 < d["interior"] < d["tiles"]  # tiles holding the first or last plane touch halo columns


def test_empty_matrix():
    d = check((np.array([0]), np.array([], dtype=np.int64), np.array([])))
    assert d["ok"] == 0


@pytest.mark.parametrize("n", [1, 5, 8, 9, 63, 65])
def test_tiny_matrices(n):
    """fewer rows than one row block / than the grid detector looks at: consecutive-row tiles, partly empty blocks"""
    rp, ci, v = [0], [], []
    for i in range(n):
        for j in (i - 1, i, i + 1):
            if 0 <= j < n:
                ci.append(j)
                v.append(2.0 if j == i else -1.0)
        rp.append(len(ci))
    d = check((np.array(rp), np.array(ci), np.array(v)))
    assert d["ok"] == 1 and d["diff"] <= 1e-14 * d["scale"], d
    assert d["blocks"] == (n + 7) // 8


def test_matrix_without_diagonal_entries():
    """rows whose own column is absent: the dot epilogue's self-column id is 0xffff (checked by de_brb_format_check)"""
    n = 200
    rp, ci, v = [0], [], []
    for i in range(n):
        for j in ((i + 1) % n, (i + 7) % n):
            ci.append(j)
            v.append(1.0 + 0.1 * j)
        order = np.argsort(ci[-2:])
        ci[-2:] = [ci[-2:][k] for k in order]
        v[-2:] = [v[-2:][k] for k in order]
        rp.append(len(ci))
    d = check((np.array(rp), np.array(ci), np.array(v)))
    assert d["ok"] == 1 and d["diff"] <= 1e-13 * d["scale"], d
