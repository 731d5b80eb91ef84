"""Synthetic matrix generators and analytic spectra for the eigensolver hot path (host side, numpy).

The first group restates the reference's own generators (reference src/dune-eigensolver.cc:98-156, built on
dune-istl's `setupLaplacian`, which the in-repo analytic spectrum pins as the 5-point stencil with
diagonal 4, off-diagonals -1, no h^2 scaling, lexicographic ordering `x = idx % N, y = idx / N`,
src/dune-eigensolver.cc:132-133, :437-446).

The second group is NEW (the reference has no 3D / Q1 / mass / coefficient code): 3D 7-point
finite differences, Q1 (bi/tri-linear) finite-element stiffness and consistent mass on structured grids with
Dirichlet elimination, optionally with a cellwise constant high-contrast coefficient. All generators can emit
a contiguous ROW RANGE of the global matrix so that each rank of a row-partitioned run builds only its own
slab (BASELINE.json configs 4/5).

Every function returns `(rowptr int64, col int64, val float64)` with ascending columns inside a row, and
keeps structural zeros (as FE assembly and the reference's masked/identity matrices do).
"""
import itertools

import numpy as np


# ------------------------------------------------------------------------------------------------
# generic structured stencil -> CSR
# ------------------------------------------------------------------------------------------------
def _row_range(n, rows):
    if rows is None:
        return 0, n
    r0, r1 = int(rows[0]), int(rows[1])
    if not (0 <= r0 <= r1 <= n):
        raise ValueError("row range out of bounds")
    return r0, r1


def _coords(idx, shape):
    """lexicographic index -> coordinates, first coordinate fastest (x = idx % Nx, ...)."""
    out = []
    for s in shape:
        out.append(idx % s)
        idx = idx // s
    return out


def _assemble(shape, offsets, values, rows=None, chunk=1 << 21):
    """Build CSR rows [r0, r1) of a structured-grid stencil matrix.

    offsets : list of integer offset tuples (dx, dy[, dz]) in ASCENDING lexicographic-index order
    values  : callable(coords list, k) -> array of the k-th offset's entries for those rows
    """
    dim = len(shape)
    n = int(np.prod(shape))
    r0, r1 = _row_range(n, rows)
    strides = [1]
    for s in shape[:-1]:
        strides.append(strides[-1] * s)
    lin = [sum(o[d] * strides[d] for d in range(dim)) for o in offsets]
    if any(lin[k] >= lin[k + 1] for k in range(len(lin) - 1)):
        raise ValueError("stencil offsets must be in ascending index order")
    counts = np.empty(r1 - r0, dtype=np.int64)
    cols_parts, vals_parts = [], []
    for c0 in range(r0, r1, chunk):
        c1 = min(r1, c0 + chunk)
        idx = np.arange(c0, c1, dtype=np.int64)
        xyz = _coords(idx, shape)
        nk = len(offsets)
        valid = np.empty((c1 - c0, nk), dtype=bool)
        col = np.empty((c1 - c0, nk), dtype=np.int64)
        val = np.empty((c1 - c0, nk), dtype=np.float64)
        for k, o in enumerate(offsets):
            ok = np.ones(c1 - c0, dtype=bool)
            for d in range(dim):
                if o[d] < 0:
                    ok &= xyz[d] >= -o[d]
                elif o[d] > 0:
                    ok &= xyz[d] < shape[d] - o[d]
            valid[:, k] = ok
            col[:, k] = idx + lin[k]
            val[:, k] = values(xyz, k)
        counts[c0 - r0:c1 - r0] = valid.sum(axis=1)
        cols_parts.append(col[valid])
        vals_parts.append(val[valid])
    rowptr = np.zeros(r1 - r0 + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    col = np.concatenate(cols_parts) if cols_parts else np.zeros(0, dtype=np.int64)
    val = np.concatenate(vals_parts) if vals_parts else np.zeros(0, dtype=np.float64)
    return rowptr, col, val


def _box_offsets(dim, radius_one_norm=None):
    """all offsets in {-1,0,1}^dim in ascending index order (last coordinate slowest)."""
    offs = []
    for rev in itertools.product((-1, 0, 1), repeat=dim):
        o = tuple(reversed(rev))  # product varies the LAST factor fastest -> make x fastest
        if radius_one_norm is None or sum(abs(v) for v in o) <= radius_one_norm:
            offs.append(o)
    return offs


# ------------------------------------------------------------------------------------------------
# group 1: the reference's generators (2D, 5-point)
# ------------------------------------------------------------------------------------------------
def laplacian_dirichlet_2d(N, rows=None):
    """reference get_laplacian_dirichlet (src/dune-eigensolver.cc:98-103)."""
    offs = _box_offsets(2, 1)

    def v(xyz, k):
        return np.full(xyz[0].shape, 4.0 if offs[k] == (0, 0) else -1.0)

    return _assemble((N, N), offs, v, rows)


def laplacian_neumann_2d(N, rows=None):
    """reference get_laplacian_neumann (src/dune-eigensolver.cc:105-121): diagonal = |sum of off-diagonals|."""
    offs = _box_offsets(2, 1)

    def v(xyz, k):
        if offs[k] != (0, 0):
            return np.full(xyz[0].shape, -1.0)
        x, y = xyz
        nb = (x > 0).astype(np.float64) + (x < N - 1) + (y > 0) + (y < N - 1)
        return nb

    return _assemble((N, N), offs, v, rows)


def laplacian_B_2d(N, overlap, rows=None):
    """reference get_laplacian_B (src/dune-eigensolver.cc:124-143): Dirichlet Laplacian masked by a 0/1
    partition of unity that vanishes within `overlap` nodes of the boundary; zero entries are kept."""
    offs = _box_offsets(2, 1)

    def pu(x, y):
        out = (x < overlap) | (x > N - 1 - overlap) | (y < overlap) | (y > N - 1 - overlap)
        return np.where(out, 0.0, 1.0)

    def v(xyz, k):
        x, y = xyz
        base = 4.0 if offs[k] == (0, 0) else -1.0
        return base * pu(x, y) * pu(x + offs[k][0], y + offs[k][1])

    return _assemble((N, N), offs, v, rows)


def identity_on_laplacian_pattern_2d(N, rows=None):
    """reference get_identity (src/dune-eigensolver.cc:145-156)."""
    offs = _box_offsets(2, 1)

    def v(xyz, k):
        return np.full(xyz[0].shape, 1.0 if offs[k] == (0, 0) else 0.0)

    return _assemble((N, N), offs, v, rows)


def eigenvalues_laplace_dirichlet_2d(N):
    """reference eigenvalues_laplace_dirichlet_2d (src/dune-eigensolver.cc:437-446), ascending."""
    h = 1.0 / (N + 1.0)
    s = np.sin(0.5 * h * np.pi * np.arange(1, N + 1)) ** 2
    return np.sort((4.0 * (s[:, None] + s[None, :])).reshape(-1))


# ------------------------------------------------------------------------------------------------
# group 2: new generators (finite differences in 3D, Q1 finite elements in 2D/3D)
# ------------------------------------------------------------------------------------------------
def laplacian_fd(shape, rows=None):
    """(2d+1)-point finite-difference Dirichlet Laplacian, diagonal 2d, off-diagonals -1, no h scaling
    (the reference's convention extended to any dimension)."""
    shape = tuple(int(s) for s in shape)
    dim = len(shape)
    offs = _box_offsets(dim, 1)
    zero = (0,) * dim

    def v(xyz, k):
        return np.full(xyz[0].shape, 2.0 * dim if offs[k] == zero else -1.0)

    return _assemble(shape, offs, v, rows)


def eigenvalues_laplacian_fd(shape):
    """analytic spectrum of laplacian_fd, ascending: sum_d 4 sin^2(pi i_d / (2 (N_d + 1)))."""
    lam = np.zeros(1)
    for s in shape:
        l1 = 4.0 * np.sin(0.5 * np.pi * np.arange(1, s + 1) / (s + 1.0)) ** 2
        lam = (lam[:, None] + l1[None, :]).reshape(-1)
    return np.sort(lam)


_K1 = np.array([[1.0, -1.0], [-1.0, 1.0]])
_M1 = np.array([[2.0, 1.0], [1.0, 2.0]]) / 6.0


def _q1_element(dim, kind):
    """Q1 element matrix on the unit cube (h = 1): index [p..., p'...] with p in {0,1}^dim, x fastest."""
    def kron(mats):
        out = np.ones((1, 1))
        for mtx in reversed(mats):  # last dimension slowest
            out = np.kron(out, mtx)
        return out

    if kind == "mass":
        return kron([_M1] * dim)
    E = np.zeros((2 ** dim, 2 ** dim))
    for d in range(dim):
        E += kron([_K1 if e == d else _M1 for e in range(dim)])
    return E


def _q1(shape, kind, kappa=None, rows=None):
    shape = tuple(int(s) for s in shape)
    dim = len(shape)
    offs = _box_offsets(dim)
    E = _q1_element(dim, kind)
    slot = {o: k for k, o in enumerate(offs)}
    corner = list(itertools.product((0, 1), repeat=dim))  # tuples in (slowest..fastest) order
    corner = [tuple(reversed(c)) for c in corner]         # -> (x, y, z) order, x fastest in local index

    def lidx(p):
        return sum(p[d] << d for d in range(dim))

    # contributions: for the cell at offset c in {0,1}^dim (cell coordinate = node coordinate + c) the node is
    # the cell's local corner p = 1 - c, and the cell's corner p' sits at node offset c - 1 + p'.
    contrib = {k: [] for k in range(len(offs))}
    for c in corner:
        p = tuple(1 - cd for cd in c)
        for pp in corner:
            o = tuple(c[d] - 1 + pp[d] for d in range(dim))
            contrib[slot[o]].append((c, E[lidx(p), lidx(pp)]))

    def v(xyz, k):
        out = np.zeros(xyz[0].shape)
        for c, e in contrib[k]:
            if kappa is None:
                out += e
            else:
                out += e * kappa(*[xyz[d] + c[d] for d in range(dim)])
        return out

    return _assemble(shape, offs, v, rows)


def q1_stiffness(shape, kappa=None, rows=None):
    """Q1 stiffness matrix (9-point in 2D, 27-point in 3D) on prod(shape) interior nodes, Dirichlet boundary
    eliminated, h = 1. `kappa(cx, cy[, cz])` optionally gives a cellwise constant coefficient on the
    (N_d + 1)^dim cells; cell c spans nodes c-1 .. c."""
    return _q1(shape, "stiffness", kappa, rows)


def q1_mass(shape, rows=None):
    """consistent Q1 mass matrix on the same pattern as q1_stiffness."""
    return _q1(shape, "mass", None, rows)


def high_contrast_kappa(contrast=1e6, period=8):
    """deterministic channel/inclusion pattern: kappa = contrast where ((cx/period)+(cy/period)+(cz/period)) % 4 == 0
    (SURVEY.md §8d), 1 elsewhere."""
    def kappa(*c):
        s = sum(ci // period for ci in c)
        return np.where(s % 4 == 0, float(contrast), 1.0)

    return kappa


def _q1_1d(N):
    th = np.pi * np.arange(1, N + 1) / (N + 1.0)
    return 2.0 - 2.0 * np.cos(th), (4.0 + 2.0 * np.cos(th)) / 6.0  # stiffness, mass eigenvalues (h = 1)


def eigenvalues_q1_stiffness(shape):
    """analytic spectrum of q1_stiffness (constant coefficient), ascending."""
    ks, ms = zip(*[_q1_1d(s) for s in shape])
    dim = len(shape)
    lam = np.zeros(1)
    grids = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    lam = np.zeros(grids[0].shape)
    for d in range(dim):
        term = np.ones(grids[0].shape)
        for e in range(dim):
            term = term * (ks[e] if e == d else ms[e])[grids[e]]
        lam += term
    return np.sort(lam.reshape(-1))


def eigenvalues_q1_pencil(shape):
    """analytic spectrum of the pencil (q1_stiffness, q1_mass), ascending: sum_d kappa_d / mu_d."""
    lam = np.zeros(1)
    for s in shape:
        k, mu = _q1_1d(s)
        lam = (lam[:, None] + (k / mu)[None, :]).reshape(-1)
    return np.sort(lam)


def to_scipy(csr, ncols=None):
    import scipy.sparse as sp

    rp, ci, v = csr
    n = len(rp) - 1
    return sp.csr_matrix((v, ci, rp), shape=(n, ncols if ncols is not None else n))
