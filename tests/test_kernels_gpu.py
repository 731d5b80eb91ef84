"""Parity of every CUDA kernel against the oracle, through the C ABI (ctypes), on the same seeded inputs.

Tolerances (fp64): SpMM / dots / Gram 1e-13 relative to the magnitude of the terms (different summation trees);
orthonormalisation: |Q^T Q - I| <= 1e-13 and |Q_gpu - Q_ref| <= 1e-10 * cond."""
import numpy as np
import pytest

from dune_eigensolver_b200 import capi, eigensolver as E, matrices as M

pytestmark = pytest.mark.gpu


def rnd(n, m, seed=0):
    return np.random.default_rng(seed).standard_normal((n, m))


MATS = {
    "lap2d_17": lambda: M.laplacian_dirichlet_2d(17),
    "lapB_17": lambda: M.laplacian_B_2d(17, 3),
    "fd3d_9x7x5": lambda: M.laplacian_fd((9, 7, 5)),
    "q1_3d_8": lambda: M.q1_stiffness((8, 8, 8)),
    "q1_hc_7": lambda: M.q1_stiffness((7, 7, 7), kappa=M.high_contrast_kappa(1e6, 2)),
}


def test_layout_roundtrip(ctx, golden):
    for n, m in [(1, 8), (37, 16), (1000, 64), (5, 24)]:
        X = rnd(n, m, n)
        mv = E.MultiVector.from_array(ctx, X)
        assert np.array_equal(mv.download(), X)
        assert np.array_equal(mv.download_rowmajor(), X)
        mv2 = E.MultiVector(ctx, n, m)
        assert np.array_equal(mv2.download(), np.zeros((n, m)))  # zero-initialised like multivector.hh:52
        mv2.copy_from(mv)
        assert np.array_equal(mv2.download(), X)
    with pytest.raises(capi.DeError, match="number of cols must be a multiple of block size"):
        E.MultiVector(ctx, 10, 12)  # multivector.hh:48-49


@pytest.mark.parametrize("name", sorted(MATS))
@pytest.mark.parametrize("m", [8, 16, 24, 32, 48, 64])
def test_spmm_matches_oracle(ctx, oracle, name, m):
    A = MATS[name]()
    n = len(A[0]) - 1
    X = rnd(n, m, m)
    dA, dX, dY = E.Matrix(ctx, A), E.MultiVector.from_array(ctx, X), E.MultiVector(ctx, n, m)
    E.matmul_sparse_tallskinny_blocked(dY, dA, dX)
    ref = oracle.spmm(A, X)
    scale = np.abs(M.to_scipy(A)).dot(np.abs(X)).max()
    assert np.abs(dY.download() - ref).max() <= 1e-14 * scale
    # fused Rayleigh-quotient epilogue
    dp = E.matmul_sparse_tallskinny_with_dots(dY, dA, dX)
    assert np.abs(dY.download() - ref).max() <= 1e-14 * scale
    dref = oracle.diag_dot(X, ref)
    assert np.abs(dp - dref).max() <= 1e-13 * (np.abs(X) * np.abs(ref)).sum(0).max()


BRB_MATS = {
    "lap2d_40": lambda: M.laplacian_dirichlet_2d(40),
    "fd3d_19x11x7": lambda: M.laplacian_fd((19, 11, 7)),
    "fd3d_24": lambda: M.laplacian_fd((24, 24, 24)),
    "q1_3d_13": lambda: M.q1_stiffness((13, 13, 13)),
    "q1_3d_21x9x10": lambda: M.q1_stiffness((21, 9, 10)),
    "q1_hc_12": lambda: M.q1_stiffness((12, 12, 12), kappa=M.high_contrast_kappa(1e6, 2)),
}


def _tridiag(n, diagonal=True):
    rp, ci, v = [0], [], []
    for i in range(n):
        for j in (i - 1, i, i + 1):
            if 0 <= j < n and (diagonal or j != i):
                ci.append(j)
                v.append(2.0 if j == i else -1.0 - 0.01 * j)
        rp.append(len(ci))
    return np.array(rp), np.array(ci), np.array(v)


@pytest.mark.parametrize("n", [1, 5, 9, 70])
@pytest.mark.parametrize("diagonal", [True, False])
def test_spmm_tiny_and_diagonal_free_matrices(ctx, oracle, n, diagonal):
    """partly empty row blocks; rows without their own column (the dot epilogue then reads X from global memory)"""
    A = _tridiag(n, diagonal)
    if len(A[1]) == 0:
        pytest.skip("empty matrix")
    m = 16
    X = rnd(n, m, n)
    dA, dX, dY = E.Matrix(ctx, A), E.MultiVector.from_array(ctx, X), E.MultiVector(ctx, n, m)
    ref = oracle.spmm(A, X)
    for fmt in (["brb"] if dA.spmm_info()["tiles"] > 0 else []) + ["csr"]:
        dA.set_spmm_format(fmt)
        dp = E.matmul_sparse_tallskinny_with_dots(dY, dA, dX)
        assert np.abs(dY.download() - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max()), fmt
        assert np.abs(dp - oracle.diag_dot(X, ref)).max() <= 1e-12 * max(1.0, (np.abs(X) * np.abs(ref)).sum(0).max()), fmt


def _random_csr(n, per_row, seed):
    rng = np.random.default_rng(seed)
    rp, ci, v = [0], [], []
    for i in range(n):
        k = 0 if i % 17 == 3 else int(rng.integers(1, per_row + 1))  # some empty rows
        ci.extend(np.sort(rng.choice(n, size=k, replace=False)).tolist())
        v.extend(rng.standard_normal(k).tolist())
        rp.append(len(ci))
    return np.array(rp), np.array(ci), np.array(v)


@pytest.mark.parametrize("name", sorted(BRB_MATS) + ["random_900"])
@pytest.mark.parametrize("m", [8, 16, 24, 32, 40, 56, 64])
def test_spmm_brb_and_csr_kernels_agree_with_oracle(ctx, oracle, name, m):
    """both SpMM kernel families (tensor-core BRB tiles, CSR rows) against the reference kernel (kernels_cpp.hh:626-657)"""
    A = BRB_MATS[name]() if name in BRB_MATS else _random_csr(900, 12, 5)
    n = len(A[0]) - 1
    X = rnd(n, m, m + 1)
    dA, dX, dY = E.Matrix(ctx, A), E.MultiVector.from_array(ctx, X), E.MultiVector(ctx, n, m)
    ref = oracle.spmm(A, X)
    scale = np.abs(M.to_scipy(A)).dot(np.abs(X)).max()
    dref = oracle.diag_dot(X, ref)
    dscale = (np.abs(X) * np.abs(ref)).sum(0).max()
    info = dA.spmm_info()
    assert info["tiles"] > 0, "no BRB form was built"
    if name in BRB_MATS:
        assert min(info["tile_shape"][:2]) >= 2, info  # structured grid detected
    out = {}
    for fmt in ("brb", "csr"):
        dA.set_spmm_format(fmt)
        assert dA.spmm_info()["format"] == fmt
        dY.upload(np.zeros((n, m)))
        E.matmul_sparse_tallskinny_blocked(dY, dA, dX)
        out[fmt] = dY.download()
        assert np.abs(out[fmt] - ref).max() <= 1e-14 * scale, fmt
        dY.upload(np.zeros((n, m)))
        dp = E.matmul_sparse_tallskinny_with_dots(dY, dA, dX)
        assert np.abs(dY.download() - ref).max() <= 1e-14 * scale, fmt
        assert np.abs(dp - dref).max() <= 1e-13 * dscale, fmt
    dA.set_spmm_format("auto")


@pytest.mark.parametrize("name", sorted(BRB_MATS) + ["random_900", "lap2d_200"])
def test_device_built_brb_arrays_equal_the_host_builder(ctx, name):
    """the BRB form is built by CUDA kernels from the uploaded CSR; the host builder (exercised on the CPU by
    tests/test_brb_format_cpu.py) must produce the same words"""
    if name == "lap2d_200":
        A = M.laplacian_dirichlet_2d(200)
    else:
        A = BRB_MATS[name]() if name in BRB_MATS else _random_csr(900, 12, 5)
    dA = E.Matrix(ctx, A)
    assert dA.spmm_info()["tiles"] > 0
    assert dA.brb_selfcheck(A) == 0


@pytest.mark.parametrize("name", ["q1_3d_13", "fd3d_24", "lap2d_40", "random_900"])
@pytest.mark.parametrize("m", [8, 16, 32, 48, 64])
def test_spmm_with_gram_epilogue(ctx, oracle, name, m):
    """Y = A X, diag(X^T Y) and Y^T Y from one pass (the Gram epilogue of the tensor-core kernel for m = 8/16/32, a
    separate Gram pass otherwise) against the reference kernels (kernels_cpp.hh:626-657, :24-55, :58-96)"""
    A = BRB_MATS[name]() if name in BRB_MATS else _random_csr(900, 12, 5)
    n = len(A[0]) - 1
    X = rnd(n, m, 7 * m)
    dA, dX, dY = E.Matrix(ctx, A), E.MultiVector.from_array(ctx, X), E.MultiVector(ctx, n, m)
    ref = oracle.spmm(A, X)
    scale = np.abs(M.to_scipy(A)).dot(np.abs(X)).max()
    for fmt in ("brb", "csr"):
        dA.set_spmm_format(fmt)
        dp, G = E.matmul_sparse_tallskinny_with_dots_and_gram(dY, dA, dX)
        assert np.abs(dY.download() - ref).max() <= 1e-14 * scale, fmt
        assert np.abs(dp - oracle.diag_dot(X, ref)).max() <= 1e-13 * (np.abs(X) * np.abs(ref)).sum(0).max(), fmt
        Gref = oracle.gram(ref, ref)
        assert np.abs(G - Gref).max() <= 1e-13 * np.abs(Gref).max() * np.sqrt(n), fmt
        assert np.abs(G - G.T).max() <= 1e-12 * np.abs(G).max()


def test_spmm_brb_is_deterministic_and_linear(ctx):
    """size-independent properties at a size the oracle would not finish quickly: run-to-run bit identity, and
    A(aX + bZ) = a AX + b AZ to rounding, on the 27-point 64^3 matrix (m = 32)."""
    A = M.q1_stiffness((64, 64, 64))
    n, m = 64 ** 3, 32
    dA = E.Matrix(ctx, A)
    dA.set_spmm_format("brb")
    X, Z = rnd(n, m, 1), rnd(n, m, 2)
    dX, dZ, dY = E.MultiVector.from_array(ctx, X), E.MultiVector.from_array(ctx, Z), E.MultiVector(ctx, n, m)
    E.matmul_sparse_tallskinny_blocked(dY, dA, dX)
    y1 = dY.download()
    E.matmul_sparse_tallskinny_blocked(dY, dA, dX)
    assert np.array_equal(y1, dY.download())
    E.matmul_sparse_tallskinny_blocked(dY, dA, dZ)
    yz = dY.download()
    dX.upload(2.0 * X - 0.5 * Z)
    E.matmul_sparse_tallskinny_blocked(dY, dA, dX)
    assert np.abs(dY.download() - (2.0 * y1 - 0.5 * yz)).max() <= 1e-13 * np.abs(y1).max()
    # against the CSR kernel of the same library at full size
    dA.set_spmm_format("csr")
    E.matmul_sparse_tallskinny_blocked(dY, dA, dX)
    ycsr = dY.download()
    dA.set_spmm_format("brb")
    E.matmul_sparse_tallskinny_blocked(dY, dA, dX)
    assert np.abs(dY.download() - ycsr).max() <= 1e-13 * np.abs(ycsr).max()


def test_spmm_golden(ctx, golden):
    N, m = int(golden["k_N"]), int(golden["k_m"])
    A = M.laplacian_dirichlet_2d(N)
    dA, dX, dY = E.Matrix(ctx, A), E.MultiVector.from_array(ctx, golden["k_X"]), E.MultiVector(ctx, N * N, m)
    E.matmul_sparse_tallskinny_blocked(dY, dA, dX)
    np.testing.assert_allclose(dY.download(), golden["k_spmm"], rtol=0, atol=1e-13)


def test_spmm_shape_errors(ctx):
    A = M.laplacian_dirichlet_2d(5)
    dA = E.Matrix(ctx, A)
    with pytest.raises(capi.DeError, match="rows does not match"):
        E.matmul_sparse_tallskinny_blocked(E.MultiVector(ctx, 25, 8), dA, E.MultiVector(ctx, 24, 8))
    with pytest.raises(capi.DeError, match="columns does not match"):
        E.matmul_sparse_tallskinny_blocked(E.MultiVector(ctx, 25, 16), dA, E.MultiVector(ctx, 25, 8))
    with pytest.raises(capi.DeError, match="column index out of range"):
        E.Matrix(ctx, (np.array([0, 1]), np.array([3]), np.array([1.0])))


@pytest.mark.parametrize("n,m", [(1, 8), (36, 16), (1000, 8), (4099, 24), (20000, 32), (70001, 64), (300000, 40)])
def test_dots_and_gram_match_oracle(ctx, oracle, n, m):
    X, Y = rnd(n, m, 1), rnd(n, m, 2)
    dX, dY = E.MultiVector.from_array(ctx, X), E.MultiVector.from_array(ctx, Y)
    tol = 1e-13 * np.sqrt(n) * 10 + 1e-13
    dp = E.dot_products_diagonal_blocked(dX, dY)
    assert np.abs(dp - oracle.diag_dot(X, Y)).max() <= tol * max(1.0, np.abs(dp).max())
    G = E.dot_products_all_blocked(dX, dY)
    Gref = oracle.gram(X, Y) if n * m * m < 3e8 else X.T @ Y
    assert np.abs(G - Gref).max() <= tol * max(1.0, np.abs(Gref).max())
    Gs = E.dot_products_all_blocked(dX, dX)  # Y aliases X: single-operand staging
    assert np.abs(Gs - X.T @ X).max() <= tol * n
    assert np.array_equal(Gs, Gs.T) or np.abs(Gs - Gs.T).max() < 1e-9


def test_dot_golden_and_errors(ctx, golden):
    dX, dY = E.MultiVector.from_array(ctx, golden["k_X"]), E.MultiVector.from_array(ctx, golden["k_Y"])
    np.testing.assert_allclose(E.dot_products_diagonal_blocked(dX, dY), golden["k_diag_dot"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(E.dot_products_all_blocked(dX, dY), golden["k_gram"], rtol=0, atol=1e-12)
    with pytest.raises(capi.DeError, match="number of rows does not match"):  # kernels_cpp.hh:29-30
        E.dot_products_diagonal_blocked(dX, E.MultiVector(ctx, 35, 16))
    with pytest.raises(capi.DeError, match="number of columns does not match"):  # kernels_cpp.hh:31-32
        E.dot_products_diagonal_blocked(dX, E.MultiVector(ctx, 36, 8))


@pytest.mark.parametrize("n,m", [(1, 8), (130, 16), (5000, 8), (5000, 24), (33333, 32), (9000, 40), (9000, 56), (40001, 64)])
def test_block_update_and_project(ctx, n, m):
    X = rnd(n, m, 3)
    R = np.triu(rnd(m, m, 4)) + 3 * np.eye(m)
    dX = E.MultiVector.from_array(ctx, X)
    E.block_update(dX, R)  # the V <- V U of kernels_cpp.hh:293-305 with a general factor
    np.testing.assert_allclose(dX.download(), X @ R, rtol=0, atol=1e-12 * m)
    Q = rnd(m, m, 5)
    dX.upload(X)
    E.block_update(dX, Q)
    np.testing.assert_allclose(dX.download(), X @ Q, rtol=0, atol=1e-12 * m)
    if m >= 16:  # projection Q_j -= Q_k S of kernels_cpp.hh:335-348
        S = rnd(8, 8, 6)
        dX.upload(X)
        E.block_project(dX, 8, 0, S)
        ref = X.copy()
        ref[:, 8:16] -= X[:, 0:8] @ S
        np.testing.assert_allclose(dX.download(), ref, rtol=0, atol=1e-12)
        with pytest.raises(capi.DeError, match="disjoint"):
            E.block_project(dX, 0, 0, S)


@pytest.mark.parametrize("n,m", [(20000, 32), (9000, 64), (5000, 8)])
def test_orthonormalize_one_sweep_decision(ctx, oracle, monkeypatch, n, m):
    """a block whose scaled Gram matrix is close to I (the steady state of the subspace iteration: X = A Q): the tail of the
    first Gram reduction decides that ONE CholQR sweep is enough (kernels_dense.cuh, kWellCond). The result must be as
    orthonormal as the two-sweep one, equal to it to rounding, and equal to the reference's Gram-Schmidt"""
    Q0, _ = np.linalg.qr(rnd(n, m, 11))
    X = Q0 * np.linspace(1.0, 1.3, m) + 1e-3 * rnd(n, m, 12) / np.sqrt(n)
    dX = E.MultiVector.from_array(ctx, X)
    c0 = ctx.launch_count()
    E.orthonormalize_blocked(dX)
    launches_one = ctx.launch_count() - c0
    Q1 = dX.download()
    dX.close()
    assert np.abs(Q1.T @ Q1 - np.eye(m)).max() <= 1e-13
    ref = oracle.orthonormalize(X)
    assert np.abs(Q1 - ref).max() <= 1e-10
    monkeypatch.setenv("DE_B200_ONE_SWEEP", "0")
    ctx2 = E.Context(0)
    try:
        dX2 = E.MultiVector.from_array(ctx2, X)
        c0 = ctx2.launch_count()
        E.orthonormalize_blocked(dX2)
        launches_two = ctx2.launch_count() - c0
        Q2 = dX2.download()
        dX2.close()
    finally:
        ctx2.close()
    assert np.abs(Q2.T @ Q2 - np.eye(m)).max() <= 1e-13
    assert np.abs(Q1 - Q2).max() <= 1e-12
    assert launches_one > launches_two  # the extra launch is the plain update that replaces the fused update + Gram


@pytest.mark.parametrize("n,m", [(36, 16), (64, 64), (3000, 8), (3000, 24), (50000, 32), (20000, 48), (100003, 64)])
def test_orthonormalize_matches_oracle(ctx, oracle, n, m):
    X = rnd(n, m, 7)
    X[:, 1] += 0.9 * X[:, 0]  # some correlation between columns
    dX = E.MultiVector.from_array(ctx, X)
    E.orthonormalize_blocked(dX)
    Q = dX.download()
    assert np.abs(Q.T @ Q - np.eye(m)).max() <= 1e-13
    ref = oracle.orthonormalize(X)
    cond = np.linalg.cond(X)
    assert np.abs(Q - ref).max() <= 1e-10 * cond
    # triangular factor with positive diagonal: Q^T X upper triangular, diag > 0 (column order preserved)
    Rf = Q.T @ X
    assert np.abs(np.tril(Rf, -1)).max() <= 1e-10 * np.abs(Rf).max()
    assert np.all(np.diag(Rf) > 0)


@pytest.mark.parametrize("m", [8, 16, 32, 64])
@pytest.mark.parametrize("cond", [1.0, 1e3, 1e6])
def test_orthonormalize_conditioning(ctx, oracle, m, cond):
    """CholQR2 across conditioning regimes: cond = 1 exercises the device-side skip of the second sweep (the fused
    Gram of the first sweep is already I), cond = 1e6 needs the second sweep to restore orthogonality."""
    n = 20011
    rng = np.random.default_rng(m)
    Q0, _ = np.linalg.qr(rng.standard_normal((n, m)))
    V, _ = np.linalg.qr(rng.standard_normal((m, m)))
    X = (Q0 * np.logspace(0, -np.log10(cond), m)) @ V.T if cond > 1.0 else Q0  # singular values 1 .. 1/cond
    dX = E.MultiVector.from_array(ctx, X)
    E.orthonormalize_blocked(dX)
    Q = dX.download()
    assert np.abs(Q.T @ Q - np.eye(m)).max() <= 5e-14
    ref = oracle.orthonormalize(X)
    assert np.abs(Q - ref).max() <= 1e-10 * max(cond, 1.0)


def test_orthonormalize_golden_and_rank_deficient(ctx, golden):
    dX = E.MultiVector.from_array(ctx, golden["k_X"])
    E.orthonormalize_blocked(dX)
    np.testing.assert_allclose(dX.download(), golden["k_ortho"], rtol=0, atol=1e-12)
    X = rnd(500, 16, 8)
    X[:, 5] = X[:, 2]  # exactly dependent columns: the reference would produce NaN; we fail loudly
    dX = E.MultiVector.from_array(ctx, X)
    with pytest.raises(capi.DeError) as ei:
        E.orthonormalize_blocked(dX)
    assert ei.value.status == capi.DE_ERR_SINGULAR


@pytest.mark.parametrize("name,m", [("lapB_17", 16), ("lapB_17", 32), ("q1_3d_8", 24), ("q1_3d_8", 64)])
def test_b_orthonormalize_matches_oracle(ctx, oracle, name, m):
    B = MATS[name]() if name != "q1_3d_8" else M.q1_mass((8, 8, 8))
    n = len(B[0]) - 1
    X = rnd(n, m, 9)
    dB, dX, dBX = E.Matrix(ctx, B), E.MultiVector.from_array(ctx, X), E.MultiVector(ctx, n, m)
    E.B_orthonormalize_blocked(dB, dX, dBX)
    Q = dX.download()
    Bs = M.to_scipy(B)
    assert np.abs(Q.T @ (Bs @ Q) - np.eye(m)).max() <= 1e-12
    ref, _ = oracle.b_orthonormalize(B, X)
    assert np.abs(Q - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
    np.testing.assert_allclose(dBX.download(), Bs @ Q, rtol=0, atol=1e-11 * max(1.0, np.abs(Bs @ Q).max()))


def test_b_orthonormalize_golden(ctx, golden):
    N = int(golden["k_N"])
    dB = E.Matrix(ctx, M.laplacian_B_2d(N, 1))
    dX = E.MultiVector.from_array(ctx, golden["k_X"])
    E.B_orthonormalize_blocked(dB, dX)
    np.testing.assert_allclose(dX.download(), golden["k_bortho"], rtol=0, atol=1e-10)


@pytest.mark.parametrize("ordering", [0, 1])
@pytest.mark.parametrize("m", [8, 16, 32, 64])
def test_factor_apply_matches_oracle(ctx, oracle, ordering, m):
    N = 24
    rp, ci, v = M.laplacian_neumann_2d(N)
    v = v.copy()
    E._add_to_diagonal(rp, ci, v, 1e-2)
    hf = E.HostFactorization((rp, ci, v), ordering, scale_rows=(m == 16))
    n = N * N
    X = rnd(n, m, 10)
    dF = E.Factor(ctx, hf)
    info = dF.info()
    assert info["lnz"] == hf.lnz and info["levels_L"] >= 1
    dX, dY = E.MultiVector.from_array(ctx, X), E.MultiVector(ctx, n, m)
    E.matmul_inverse_tallskinny_blocked(dY, dF, dX)
    ref, _ = oracle.factor_apply(hf.arrays(), X)
    got = dY.download()
    assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max()
    A = M.to_scipy((rp, ci, v))
    assert np.abs(A @ got - X).max() <= 1e-10 * np.abs(X).max()  # F^-1 really inverts A


def test_factor_apply_golden_and_errors(ctx, golden):
    F = {k[2:]: golden[k] for k in golden.files if k.startswith("f_") and k != "f_apply"}
    dF = E.Factor(ctx, F)
    dX, dY = E.MultiVector.from_array(ctx, golden["k_X"]), E.MultiVector(ctx, 36, 16)
    E.matmul_inverse_tallskinny_blocked(dY, dF, dX)
    np.testing.assert_allclose(dY.download(), golden["f_apply"], rtol=0, atol=1e-12)
    with pytest.raises(capi.DeError, match="Qout/Qin size mismatch"):  # kernels_cpp.hh:664-665
        E.matmul_inverse_tallskinny_blocked(E.MultiVector(ctx, 36, 8), dF, dX)
    with pytest.raises(capi.DeError, match="Factorization does not match"):  # kernels_cpp.hh:666-667
        E.matmul_inverse_tallskinny_blocked(E.MultiVector(ctx, 35, 16), dF, E.MultiVector(ctx, 35, 16))


def test_empty_inputs(ctx):
    dX, dY = E.MultiVector(ctx, 0, 8), E.MultiVector(ctx, 0, 8)
    assert np.array_equal(E.dot_products_diagonal_blocked(dX, dY), np.zeros(8))
    assert np.array_equal(E.dot_products_all_blocked(dX, dY), np.zeros((8, 8)))
    dA = E.Matrix(ctx, (np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0)))
    E.matmul_sparse_tallskinny_blocked(dY, dA, dX)
    assert dY.download().shape == (0, 8)


@pytest.mark.parametrize("n", [1, 59, 60, 61, 7777, 250000])
def test_two_operand_gram_wide_block(ctx, oracle, n):
    """G = X^T Y at m = 64 runs on the warp-specialised tensor-core kernel (csrc/kernels_gram2.cuh); against the reference's
    dot_products_all_blocked (kernels_cpp.hh:58-96) including blocks shorter than one tile and ragged tails"""
    rng = np.random.default_rng(n)
    X, Y = rng.standard_normal((n, 64)), rng.standard_normal((n, 64))
    dX, dY = E.MultiVector.from_array(ctx, X), E.MultiVector.from_array(ctx, Y)
    G = E.dot_products_all_blocked(dX, dY)
    ref = oracle.gram(X, Y)
    scale = np.sqrt(n) * 4.0
    assert np.abs(G - ref).max() <= 1e-13 * scale * max(1.0, np.abs(ref).max() / scale)
    assert np.abs(G - X.T @ Y).max() <= 1e-12 * n
    dX.close()
    dY.close()
