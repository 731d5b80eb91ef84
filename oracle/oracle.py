"""TEST INFRASTRUCTURE ONLY -- CPU oracle loader. Never imported by the product package.

Two interchangeable checkers behind the same C interface (see oracle/ref_capi.cc, oracle/oracle_port.cc):

  kind "reference": oracle/_ref/libde_reference.so -- the reference's own headers compiled verbatim
                    (/root/reference/dune/eigensolver/{multivector,kernels_cpp,eigensolver}.hh)
  kind "port":      oracle/liborc_port.so          -- this repo's CPU restatement

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this module.
All multivector arguments are numpy (n, m) arrays; the conversion to the reference MultiVector<double,8>
layout (reference multivector.hh:130-133) happens here.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libde_reference.so")
PORT_LIB = os.path.join(_HERE, "liborc_port.so")

_dp = C.POINTER(C.c_double)
_lp = C.POINTER(C.c_long)


def to_panels(X):
    """(n, m) row-major -> reference block-column-major (panels of 8 columns, row-major inside a panel)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    n, m = X.shape
    assert m % 8 == 0
    # always a fresh buffer: for m == 8 the two layouts coincide and the reference clobbers some inputs
    return np.array(X.reshape(n, m // 8, 8).transpose(1, 0, 2), order="C", copy=True).reshape(-1)


def from_panels(p, n, m):
    return np.ascontiguousarray(np.asarray(p).reshape(m // 8, n, 8).transpose(1, 0, 2)).reshape(n, m)


def _d(a):
    return a.ctypes.data_as(_dp)


def _l(a):
    return a.ctypes.data_as(_lp)


def _csr(A):
    """scipy CSR (or (rowptr, col, val) tuple) -> int64 / float64 contiguous arrays."""
    if isinstance(A, tuple):
        rp, ci, v = A
    else:
        rp, ci, v = A.indptr, A.indices, A.data
    return (np.ascontiguousarray(rp, dtype=np.int64), np.ascontiguousarray(ci, dtype=np.int64),
            np.ascontiguousarray(v, dtype=np.float64))


class Oracle:
    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        L = self.lib
        L.orc_kind.restype = C.c_char_p
        L.orc_last_error.restype = C.c_char_p
        L.orc_flops_orthonormalize.restype = C.c_double
        L.orc_flops_orthonormalize.argtypes = [C.c_int, C.c_int]
        L.orc_bytes_orthonormalize_naive.restype = C.c_double
        L.orc_bytes_orthonormalize_naive.argtypes = [C.c_int, C.c_int]
        L.orc_bytes_orthonormalize_blocked.restype = C.c_double
        L.orc_bytes_orthonormalize_blocked.argtypes = [C.c_int, C.c_int, C.c_int]
        L.orc_set_factor_options.argtypes = [C.c_int, C.c_int]
        L.orc_start_block.argtypes = [C.c_long, C.c_long, C.c_uint, _dp]
        L.orc_spmm.argtypes = [C.c_long, _lp, _lp, _dp, C.c_long, _dp, _dp]
        L.orc_diag_dot.argtypes = [C.c_long, C.c_long, _dp, _dp, _dp]
        L.orc_gram.argtypes = [C.c_long, C.c_long, _dp, _dp, _dp]
        L.orc_orthonormalize.argtypes = [C.c_long, C.c_long, _dp]
        L.orc_orthonormalize_naive.argtypes = [C.c_long, C.c_long, _dp]
        L.orc_b_orthonormalize.argtypes = [C.c_long, _lp, _lp, _dp, C.c_long, _dp, _dp]
        L.orc_factor_apply.argtypes = [C.c_long, C.c_long, _lp, _lp, _dp, _lp, _lp, _dp, _lp, _lp, _dp, C.c_long,
                                       _dp, _dp]
        drv = [C.c_long, _lp, _lp, _dp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_uint, _dp, _dp, _lp]
        L.orc_standard_largest.argtypes = drv
        L.orc_standard_inverse.argtypes = drv
        L.orc_generalized_inverse.argtypes = [C.c_long, _lp, _lp, _dp, _lp, _lp, _dp, C.c_double, C.c_double,
                                              C.c_double, C.c_int, C.c_int, C.c_uint, _dp, _dp, _lp]
        self.kind = L.orc_kind().decode()

    def _chk(self, rc):
        if rc != 0:
            raise ValueError(self.lib.orc_last_error().decode())

    # ---- helpers -------------------------------------------------------------------------------
    def set_factor_options(self, ordering=1, scale_rows=False):
        self.lib.orc_set_factor_options(int(ordering), int(bool(scale_rows)))

    def start_block(self, n, m, seed=123):
        out = np.empty(n * m)
        self._chk(self.lib.orc_start_block(n, m, seed, _d(out)))
        return from_panels(out, n, m)

    # ---- kernels -------------------------------------------------------------------------------
    def spmm(self, A, X):
        rp, ci, v = _csr(A)
        n, m = X.shape
        xin, y = to_panels(X), np.empty(n * m)
        self._chk(self.lib.orc_spmm(n, _l(rp), _l(ci), _d(v), m, _d(xin), _d(y)))
        return from_panels(y, n, m)

    def diag_dot(self, X1, X2):
        n, m = X1.shape
        if X1.shape != X2.shape:
            # let the reference produce its own error
            pass
        a, b, dp = to_panels(X1), to_panels(X2), np.empty(m)
        self._chk(self.lib.orc_diag_dot(n, m, _d(a), _d(b), _d(dp)))
        return dp

    def gram(self, X1, X2):
        n, m = X1.shape
        a, b, g = to_panels(X1), to_panels(X2), np.empty(m * m)
        self._chk(self.lib.orc_gram(n, m, _d(a), _d(b), _d(g)))
        return g.reshape(m, m)

    def orthonormalize(self, X):
        n, m = X.shape
        a = to_panels(X)
        self._chk(self.lib.orc_orthonormalize(n, m, _d(a)))
        return from_panels(a, n, m)

    def orthonormalize_naive(self, X):
        n, m = X.shape
        a = np.ascontiguousarray(np.asarray(X, dtype=np.float64).T).reshape(-1)  # column-major
        self._chk(self.lib.orc_orthonormalize_naive(n, m, _d(a)))
        return np.ascontiguousarray(a.reshape(m, n).T)

    def b_orthonormalize(self, B, X):
        rp, ci, v = _csr(B)
        n, m = X.shape
        a = to_panels(X)
        nrm = C.c_double(0.0)
        self._chk(self.lib.orc_b_orthonormalize(n, _l(rp), _l(ci), _d(v), m, _d(a), C.byref(nrm)))
        return from_panels(a, n, m), nrm.value

    def factor_apply(self, F, X):
        """F: dict with Lp,Lj,Lx,Up,Ui,Ux,P,Q,Rs,do_recip (UMFPACK contract). Returns (A^-1 X, clobbered X)."""
        n, m = X.shape
        ia = {k: np.ascontiguousarray(F[k], dtype=np.int64) for k in ("Lp", "Lj", "Up", "Ui", "P", "Q")}
        da = {k: np.ascontiguousarray(F[k], dtype=np.float64) for k in ("Lx", "Ux", "Rs")}
        xin, xout = to_panels(X), np.empty(n * m)
        self._chk(self.lib.orc_factor_apply(n, m, _l(ia["Lp"]), _l(ia["Lj"]), _d(da["Lx"]), _l(ia["Up"]),
                                            _l(ia["Ui"]), _d(da["Ux"]), _l(ia["P"]), _l(ia["Q"]), _d(da["Rs"]),
                                            int(F["do_recip"]), _d(xin), _d(xout)))
        return from_panels(xout, n, m), from_panels(xin, n, m)

    # ---- drivers -------------------------------------------------------------------------------
    def _std(self, fn, A, shift, tol, maxiter, nev, seed):
        rp, ci, v = _csr(A)
        v = v.copy()
        n = len(rp) - 1
        ev, V, it = np.zeros(nev), np.zeros(nev * n), C.c_long(-1)
        self._chk(fn(n, _l(rp), _l(ci), _d(v), shift, tol, maxiter, nev, seed, _d(ev), _d(V), C.byref(it)))
        return ev, V.reshape(nev, n), it.value

    def standard_largest(self, A, shift, tol, maxiter, nev, seed=123):
        """-> (eval[nev], evec[nev, n], k) with k the reference's loop index at exit (eigensolver.hh:75-103)."""
        return self._std(self.lib.orc_standard_largest, A, shift, tol, maxiter, nev, seed)

    def standard_inverse(self, A, shift, tol, maxiter, nev, seed=123):
        return self._std(self.lib.orc_standard_inverse, A, shift, tol, maxiter, nev, seed)

    def generalized_inverse(self, A, B, shift, reg, tol, maxiter, nev, seed=123):
        """-> (eval[nev], evec[nev, n], iterations) (reference eigensolver.hh:204-351)."""
        rpa, cia, va = _csr(A)
        rpb, cib, vb = _csr(B)
        n = len(rpa) - 1
        ev, V, it = np.zeros(nev), np.zeros(nev * n), C.c_long(-1)
        self._chk(self.lib.orc_generalized_inverse(n, _l(rpa), _l(cia), _d(va), _l(rpb), _l(cib), _d(vb), shift, reg,
                                                   tol, maxiter, nev, seed, _d(ev), _d(V), C.byref(it)))
        return ev, V.reshape(nev, n), it.value

    # ---- the reference's analytic cost models ---------------------------------------------------
    def flops_orthonormalize(self, n, m):
        return self.lib.orc_flops_orthonormalize(n, m)

    def bytes_orthonormalize_naive(self, n, m):
        return self.lib.orc_bytes_orthonormalize_naive(n, m)

    def bytes_orthonormalize_blocked(self, n, m, b=8):
        return self.lib.orc_bytes_orthonormalize_blocked(n, m, b)


def load_reference():
    """The verbatim-compiled reference (None if it was never built)."""
    return Oracle(REF_LIB) if os.path.exists(REF_LIB) else None


def load_port():
    return Oracle(PORT_LIB) if os.path.exists(PORT_LIB) else None


def load_best():
    """Prefer the compiled reference, fall back to the port."""
    o = load_reference()
    return o if o is not None else load_port()
