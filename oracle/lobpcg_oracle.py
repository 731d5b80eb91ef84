"""CPU restatement of the LOBPCG drivers (StandardLOBPCG / GeneralizedLOBPCG) in numpy -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: normallytangent/dune-eigensolver has no LOBPCG (its drivers are eigensolver.hh:28-112,
:116-198, :204-351), so there is no reference vector, golden file or reference run this could be pinned to. What it is
for: an implementation of the algorithm of dune_eigensolver_b200/csrc/lobpcg_core.hpp that shares no code with it
(scipy sparse products, LAPACK for the Rayleigh-Ritz problem, Householder QR / LAPACK Cholesky for the orthonormalisation), so that
tests can compare ITERATION COUNTS and eigenvalues -- a wrong search direction or coefficient block in the product's
orchestration would still converge (as steepest descent), only more slowly, and no spectrum check would notice.
Only tests/ may import this module; nothing under dune_eigensolver_b200/ does.

Algorithm (same as the product, stated once more): X B-orthonormal Ritz vectors; W = T (A X - B X diag(theta)) with T
the Jacobi-scaled Chebyshev polynomial preconditioner (or identity), B-orthogonalised against X and B-orthonormalised;
Rayleigh-Ritz on S = [X W P] with all Gram blocks formed explicitly; X <- S C, P <- [W P] C_{W,P}; A X, A P, B X, B P
recomputed by products; a Rayleigh-Ritz problem whose unit-diagonal-scaled S^T B S has a squared Cholesky pivot below
1e-10 is retried without P. Stop when ||A x_j - theta_j B x_j||_2 <= tol |theta_j| for all j < nev.
"""
import numpy as np
import scipy.linalg as sl
import scipy.sparse as sp


def _b_orthonormalize(W, B):
    """W <- W R^-1 with W^T B W = I, R upper triangular with positive diagonal (the unique factor the product's
    CholQR2 converges to): Householder QR without B, two Cholesky-QR passes in the B inner product with it"""
    if B is None:
        Q, R = np.linalg.qr(W)
        return Q * np.where(np.diag(R) < 0.0, -1.0, 1.0)[None, :]
    for _ in range(2):
        G = W.T @ (B @ W)
        R = sl.cholesky(0.5 * (G + G.T), lower=False)
        W = sl.solve_triangular(R, W.T, trans="T", lower=False).T
    return W


def _scaled_pivots_ok(GB, floor=1e-10):
    d = 1.0 / np.sqrt(np.diag(GB))
    Bs = GB * d[:, None] * d[None, :]
    try:
        L = sl.cholesky(0.5 * (Bs + Bs.T), lower=True)
    except sl.LinAlgError:
        return False
    return bool(np.all(np.diag(L) ** 2 > floor))


def _chebyshev(A, dinv, hi, degree, R):
    ratio = max(4.0, (degree + 1) ** 2 / 2.25)
    lo = hi / ratio
    theta, delta = 0.5 * (hi + lo), 0.5 * (hi - lo)
    sigma1 = theta / delta
    rho = 1.0 / sigma1
    Z, Zold = dinv[:, None] * R / theta, np.zeros_like(R)
    for _ in range(degree):
        rho_new = 1.0 / (2.0 * sigma1 - rho)
        Znew = Z + rho_new * rho * (Z - Zold) + (2.0 * rho_new / delta) * dinv[:, None] * (R - A @ Z)
        Zold, Z, rho = Z, Znew, rho_new
    return Z


def lobpcg(A, B, X0, nev, tol, maxiter, cheb_degree=0):
    """A, B: (rowptr, col, val) triples or scipy matrices (B may be None); X0: n x m start block.
    -> (theta[m], X[n, m], iterations, restarts, converged)"""
    def mat(Mx):
        if Mx is None or sp.issparse(Mx):
            return Mx
        rp, ci, v = Mx
        return sp.csr_matrix((v, ci, rp), shape=(len(rp) - 1, len(rp) - 1))

    A, B = mat(A), mat(B)
    n, m = X0.shape
    mul_b = (lambda V: B @ V) if B is not None else (lambda V: V)
    if cheb_degree > 0:
        diag = A.diagonal()
        dinv = 1.0 / diag
        hi = float((abs(A).sum(axis=1).A1 / diag).max())
    X = _b_orthonormalize(X0.copy(), B)
    AX = A @ X
    w, V = sl.eigh(X.T @ AX)
    X, AX, theta = X @ V, AX @ V, w
    BX = mul_b(X)
    P = AP = BP = None
    restarts = 0
    it = 0
    while True:
        R = AX - BX * theta
        rel = np.linalg.norm(R, axis=0)[:nev] / np.abs(theta[:nev])
        if rel.max() <= tol:
            return theta, X, it, restarts, True
        if it >= maxiter:
            return theta, X, it, restarts, False
        W = _chebyshev(A, dinv, hi, cheb_degree, R) if cheb_degree > 0 else R
        W = W - X @ (BX.T @ W)
        W = _b_orthonormalize(W, B)
        AW, BW = A @ W, mul_b(W)
        blocks = [(X, AX, BX), (W, AW, BW)] + ([(P, AP, BP)] if P is not None else [])
        while True:
            S = np.hstack([b[0] for b in blocks])
            AS = np.hstack([b[1] for b in blocks])
            BS = np.hstack([b[2] for b in blocks])
            GA, GB = S.T @ AS, S.T @ BS
            GA, GB = 0.5 * (GA + GA.T), 0.5 * (GB + GB.T)
            if _scaled_pivots_ok(GB):
                break
            if len(blocks) == 3:
                blocks = blocks[:2]
                restarts += 1
                continue
            raise np.linalg.LinAlgError("Rayleigh-Ritz problem on [X W] is singular")
        d = 1.0 / np.sqrt(np.diag(GB))
        w, C = sl.eigh(GA * d[:, None] * d[None, :], GB * d[:, None] * d[None, :])
        C = d[:, None] * C[:, :m]
        theta = w[:m]
        Pn = np.hstack([b[0] for b in blocks[1:]]) @ C[m:]
        X = X @ C[:m] + Pn
        P = Pn
        AX, AP = A @ X, A @ P
        BX, BP = mul_b(X), mul_b(P)
        it += 1
