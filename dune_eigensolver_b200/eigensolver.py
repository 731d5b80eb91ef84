"""Host-side mirror of the reference's interface for the hot path, over the C ABI.

Same names, argument meaning and error behaviour as the reference's free functions
(reference dune/eigensolver/eigensolver.hh, kernels_cpp.hh, multivector.hh); the C++ drop-in headers in
include/dune/eigensolver/ are the primary host side, this module is what the Python parity tests and
bench.py drive. Matrices are CSR triples `(rowptr, col, val)` (what the C++ header extracts from a
BCRSMatrix through its iterators); all compute runs on the GPU through libdune_eigensolver_b200.so.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import DeError, check, dptr, f64, i64, i64ptr, lptr


def padded_cols(nev, b=8):
    """m = smallest multiple of the block size >= nev (reference eigensolver.hh:43)."""
    return (nev // b + min(nev % b, 1)) * b


def to_panels(X):
    """(n, m) row-major -> reference MultiVector<double,8> storage order (multivector.hh:130-133)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    n, m = X.shape
    if m % 8 != 0:
        raise DeError(capi.DE_ERR_INVALID, "number of cols must be a multiple of block size")
    return np.ascontiguousarray(X.reshape(n, m // 8, 8).transpose(1, 0, 2)).reshape(-1)


def from_panels(p, n, m):
    return np.ascontiguousarray(np.asarray(p).reshape(m // 8, n, 8).transpose(1, 0, 2)).reshape(n, m)


def start_block(n, m, seed=123):
    """The reference's random start block (eigensolver.hh:50-55) in panel8 storage order."""
    out = np.empty(n * m)
    check(capi.lib().de_start_block(n, m, seed, dptr(out)))
    return out


# A/B switches: the LIBRARY reads no environment variable; this mirror maps DE_B200_<NAME> onto
# de_context_set_option for bench.py, the probes and the tests
ENV_OPTIONS = {"DE_B200_ONE_SWEEP": "one_sweep", "DE_B200_CHEB_EPILOGUE": "cheb_epilogue", "DE_B200_LINCOMB2": "lincomb2",
               "DE_B200_LOOP_GRAPH": "loop_graph", "DE_B200_FUSED_PUSH": "fused_push",
               "DE_B200_BRB_PLANE_POINTS": "brb_plane_points"}


def _apply_env_options(handle):
    import os

    for var, name in ENV_OPTIONS.items():
        val = os.environ.get(var, "")
        if val != "":
            check(capi.lib().de_context_set_option(handle, name.encode(), int(val)), handle)


class Context:
    """One GPU + stream + workspaces (+ NCCL communicator once init_comm was called)."""

    def __init__(self, device=0, stream=None):
        self._h = C.c_void_p()
        check(capi.lib().de_context_create(device, C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.device = device
        _apply_env_options(self._h)

    def set_option(self, name, value):
        """tuning / A-B switch of this context (include/dune_eigensolver_b200.h: de_context_set_option)"""
        check(capi.lib().de_context_set_option(self._h, name.encode(), int(value)), self._h)

    def close(self):
        if self._h:
            capi.lib().de_context_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(capi.lib().de_context_synchronize(self._h), self._h)

    def launch_count(self):
        c = C.c_int64(0)
        check(capi.lib().de_context_launch_count(self._h, C.byref(c)), self._h)
        return c.value

    def set_profiling(self, on=True, only=None):
        """per-kernel CUDA-event timers: all categories, or only those named in `only` (e.g. ["spmm"])"""
        names = ["spmm", "gram", "update", "small", "dot", "trsv", "misc", "spmm_boundary", "halo_push", "halo_wait"]
        code = int(bool(on))
        if on and only:
            code = sum(2 << names.index(c) for c in only)
        check(capi.lib().de_context_set_profiling(self._h, code), self._h)

    def profile(self, reset=False):
        """{category: (total_ms, launches)} of the per-kernel CUDA-event timers (synchronises the stream)."""
        names = ["spmm", "gram", "update", "small", "dot", "trsv", "misc", "spmm_boundary", "halo_push", "halo_wait"]
        out = {}
        for c, name in enumerate(names):
            ms, cnt = C.c_double(0.0), C.c_int64(0)
            last = c == len(names) - 1
            check(capi.lib().de_context_profile(self._h, c, C.byref(ms), C.byref(cnt), int(reset and last)), self._h)
            out[name] = (ms.value, cnt.value)
        return out

    def peer_window_create(self, halo_bytes):
        """-> 64-byte CUDA IPC handle of this rank's NVLink window (include/dune_eigensolver_b200.h)"""
        buf = (C.c_char * 64)()
        check(capi.lib().de_context_peer_window_create(self._h, int(halo_bytes), buf), self._h)
        return bytes(buf.raw)

    def peer_window_open(self, handles):
        """handles: the 64-byte handles of all ranks, concatenated in rank order"""
        buf = (C.c_char * len(handles)).from_buffer_copy(bytes(handles))
        check(capi.lib().de_context_peer_window_open(self._h, buf), self._h)

    def peer_ready(self):
        r = C.c_int(0)
        check(capi.lib().de_context_peer_ready(self._h, C.byref(r)), self._h)
        return bool(r.value)

    def init_comm(self, rank, nranks, unique_id):
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        check(capi.lib().de_context_init_comm(self._h, rank, nranks, buf), self._h)

    def rank(self):
        r, n = C.c_int(0), C.c_int(1)
        check(capi.lib().de_context_rank(self._h, C.byref(r), C.byref(n)), self._h)
        return r.value, n.value


def comm_unique_id():
    buf = (C.c_char * 128)()
    check(capi.lib().de_comm_unique_id(buf))
    return bytes(buf)


def _csr(A):
    if hasattr(A, "indptr"):
        A = (A.indptr, A.indices, A.data)
    rp, ci, v = A
    return i64(rp), i64(ci), f64(v)


class Matrix:
    """Device CSR matrix (what the drop-in header builds from a BCRSMatrix<FieldMatrix<double,1,1>>)."""

    def __init__(self, ctx, A=None, _handle=None):
        self.ctx = ctx
        self._h = C.c_void_p()
        if _handle is not None:
            self._h = _handle
        else:
            rp, ci, v = _csr(A)
            n = len(rp) - 1
            check(capi.lib().de_matrix_create_csr(ctx._h, n, len(ci), i64ptr(rp), i64ptr(ci), dptr(v),
                                                  C.byref(self._h)), ctx._h)
        n, nnz = C.c_int64(0), C.c_int64(0)
        check(capi.lib().de_matrix_rows(self._h, C.byref(n), C.byref(nnz)))
        self.n, self.nnz = n.value, nnz.value

    @classmethod
    def bcsr(cls, ctx, rowptr, col, blocks):
        """BCSR matrix with k x k blocks (BCRSMatrix<FieldMatrix<double,k,k>>): blocks has shape (nnzb, k, k). The result
        acts on vector blocks with nb*k rows (C ABI de_matrix_create_bcsr)."""
        rp, ci = i64(rowptr), i64(col)
        bl = f64(blocks)
        assert bl.ndim == 3 and bl.shape[1] == bl.shape[2] and bl.shape[0] == len(ci)
        h = C.c_void_p()
        check(capi.lib().de_matrix_create_bcsr(ctx._h, len(rp) - 1, len(ci), bl.shape[1], i64ptr(rp), i64ptr(ci), dptr(bl),
                                               C.byref(h)), ctx._h)
        return cls(ctx, _handle=h)

    @classmethod
    def distributed(cls, ctx, n_owned, n_halo, rowptr, col_local, val, peers, recv_counts, send_offsets, send_rows):
        rp, ci, v = i64(rowptr), i64(col_local), f64(val)
        peers = capi.i32(peers)
        rc, so, sr = i64(recv_counts), i64(send_offsets), i64(send_rows)
        h = C.c_void_p()
        check(capi.lib().de_matrix_create_distributed(ctx._h, n_owned, n_halo, len(ci), i64ptr(rp), i64ptr(ci), dptr(v),
                                                      len(peers), capi.i32ptr(peers), i64ptr(rc), i64ptr(so),
                                                      i64ptr(sr), C.byref(h)), ctx._h)
        return cls(ctx, _handle=h)

    @classmethod
    def rowblock(cls, ctx, rowptr, col_global, val, part, allgather=None):
        """This rank's part of a row-partitioned matrix from its rows with GLOBAL column indices (C ABI
        de_matrix_create_rowblock: halo planning and the exchange of the halo lists happen inside the library).
        allgather(send: bytes-like of length b) -> bytes of length nranks*b, in rank order."""
        rp, ci, v, part = i64(rowptr), i64(col_global), f64(val), i64(part)
        h = C.c_void_p()
        err = []

        def _ag(user, send, recv, nbytes):
            try:
                out = allgather(C.string_at(send, nbytes))
                C.memmove(recv, bytes(out), len(out))
                return 0
            except Exception as e:  # noqa: BLE001  (must not propagate through the C frame)
                err.append(e)
                return 1

        cb = capi.ALLGATHER_FN(_ag) if allgather is not None else C.cast(None, capi.ALLGATHER_FN)
        st = capi.lib().de_matrix_create_rowblock(ctx._h, len(rp) - 1, len(ci), i64ptr(rp), i64ptr(ci), dptr(v),
                                                  i64ptr(part), cb, None, C.byref(h))
        if err:
            raise err[0]
        check(st, ctx._h)
        return cls(ctx, _handle=h)

    def set_peer_deposit(self, deposit_rows, max_halo_rows_all_ranks):
        d = i64(deposit_rows)
        check(capi.lib().de_matrix_set_peer_deposit(self._h, i64ptr(d), int(max_halo_rows_all_ranks)), self.ctx._h)

    def set_spmm_format(self, fmt):
        """'auto' | 'csr' | 'brb' (include/dune_eigensolver_b200.h: de_matrix_set_spmm_format)"""
        code = {"auto": capi.DE_SPMM_AUTO, "csr": capi.DE_SPMM_CSR, "brb": capi.DE_SPMM_BRB}[fmt]
        check(capi.lib().de_matrix_set_spmm_format(self._h, code), self.ctx._h)

    def spmm_info(self):
        fmt, tiles, blocks, steps, umax = C.c_int(0), C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
        shape = (C.c_int * 3)()
        check(capi.lib().de_matrix_spmm_info(self._h, C.byref(fmt), C.byref(tiles), C.byref(blocks), C.byref(steps),
                                             C.byref(umax), shape), self.ctx._h)
        return {"format": {capi.DE_SPMM_CSR: "csr", capi.DE_SPMM_BRB: "brb"}[fmt.value], "tiles": tiles.value,
                "row_blocks": blocks.value, "steps": steps.value, "union_rows_max": umax.value,
                "tile_shape": tuple(shape)}

    def brb_selfcheck(self, A, ncols=None):
        """words in which the device-built BRB arrays differ from the host builder's (-1: no BRB form)"""
        rp, ci, v = _csr(A)
        n = len(rp) - 1
        bad = C.c_int64(0)
        check(capi.lib().de_matrix_brb_selfcheck(self._h, n, n if ncols is None else ncols, i64ptr(rp), i64ptr(ci), dptr(v),
                                                 C.byref(bad)), self.ctx._h)
        return bad.value

    def close(self):
        if self._h:
            capi.lib().de_matrix_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Multi:
    """Single-process multi-GPU front end (C ABI de_multi_*): one host thread and one context per GPU inside the
    library, NVLink peer windows of the same process. devices may repeat an ordinal (several ranks on one GPU)."""

    def __init__(self, devices, halo_bytes=0, timeout_s=None):
        self._h = C.c_void_p()
        d = capi.i32(list(devices))
        check(capi.lib().de_multi_create(capi.i32ptr(d), len(d), int(halo_bytes), C.byref(self._h)))
        self.ndev = len(d)
        if timeout_s is not None:
            self._check(capi.lib().de_multi_set_timeout(self._h, float(timeout_s)))
        for r in range(self.ndev):  # the DE_B200_* switches apply to every rank's context
            h = C.c_void_p()
            self._check(capi.lib().de_multi_context(self._h, r, C.byref(h)))
            _apply_env_options(h)

    def close(self):
        if self._h:
            capi.lib().de_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != capi.DE_OK:
            msg = capi.lib().de_multi_last_error(self._h)
            raise DeError(st, (msg or b"").decode() or "de_status %d" % st)

    def launch_count(self):
        c = C.c_int64(0)
        self._check(capi.lib().de_multi_launch_count(self._h, C.byref(c)))
        return c.value

    def _start(self, n, m, seed, start):
        return start_block(n, m, seed) if start is None else f64(start)

    def StandardLargest(self, A, shift, tol, maxiter, nev, verbose=0, seed=123, start=None, row_align=1):
        """reference StandardLargest (eigensolver.hh:28-112), rows split over the GPUs; A is shifted in place"""
        rp, ci, v = _csr(A)
        n, m = len(rp) - 1, padded_cols(nev)
        if shift != 0.0:
            _add_to_diagonal(rp, ci, v, shift)
        st = self._start(n, m, seed, start)
        ev, V, it = np.zeros(nev), np.zeros((nev, n)), C.c_int(0)
        self._check(capi.lib().de_multi_standard_largest(self._h, n, len(ci), i64ptr(rp), i64ptr(ci), dptr(v), int(row_align),
                                                         shift, tol, maxiter, nev, dptr(st), dptr(ev), dptr(V), verbose,
                                                         C.byref(it)))
        return Result(ev, V, it.value)

    def StandardLOBPCG(self, A, tol, maxiter, nev, verbose=0, seed=123, start=None, row_align=1):
        rp, ci, v = _csr(A)
        n, m = len(rp) - 1, padded_cols(nev)
        st = self._start(n, m, seed, start)
        ev, V, it = np.zeros(nev), np.zeros((nev, n)), C.c_int(0)
        self._check(capi.lib().de_multi_standard_lobpcg(self._h, n, len(ci), i64ptr(rp), i64ptr(ci), dptr(v), int(row_align),
                                                        tol, maxiter, nev, dptr(st), dptr(ev), dptr(V), verbose, C.byref(it)))
        return Result(ev, V, it.value)

    def GeneralizedLOBPCG(self, A, B, tol, maxiter, nev, verbose=0, seed=123, start=None, row_align=1):
        rp, ci, v = _csr(A)
        brp, bci, bv = _csr(B)
        n, m = len(rp) - 1, padded_cols(nev)
        st = self._start(n, m, seed, start)
        ev, V, it = np.zeros(nev), np.zeros((nev, n)), C.c_int(0)
        self._check(capi.lib().de_multi_generalized_lobpcg(self._h, n, len(ci), i64ptr(rp), i64ptr(ci), dptr(v), len(bci),
                                                           i64ptr(brp), i64ptr(bci), dptr(bv), int(row_align), tol, maxiter,
                                                           nev, dptr(st), dptr(ev), dptr(V), verbose, C.byref(it)))
        return Result(ev, V, it.value)


class MultiVector:
    """Device-resident n x m block; mirrors MultiVector<double,8> (multivector.hh:17-146): zero-initialised,
    m % 8 == 0 enforced with the reference's message."""

    blocksize = 8

    def __init__(self, ctx, n, m):
        self.ctx = ctx
        self._h = C.c_void_p()
        check(capi.lib().de_mv_create(ctx._h, n, m, C.byref(self._h)), ctx._h)
        self.n, self.m = n, m

    @classmethod
    def from_array(cls, ctx, X):
        X = f64(X)
        mv = cls(ctx, X.shape[0], X.shape[1])
        mv.upload(X)
        return mv

    def rows(self):
        return self.n

    def cols(self):
        return self.m

    def upload(self, X):
        """(n, m) array -> device, going through the reference storage order like the C++ boundary does."""
        p = to_panels(X)
        check(capi.lib().de_mv_upload_panel8(self._h, dptr(p)), self.ctx._h)

    def upload_panels(self, p):
        p = f64(p)
        check(capi.lib().de_mv_upload_panel8(self._h, dptr(p)), self.ctx._h)

    def upload_rowmajor(self, X):
        X = f64(X)
        check(capi.lib().de_mv_upload_rowmajor(self._h, dptr(X)), self.ctx._h)

    def download(self):
        p = np.empty(self.n * self.m)
        check(capi.lib().de_mv_download_panel8(self._h, dptr(p)), self.ctx._h)
        return from_panels(p, self.n, self.m)

    def download_rowmajor(self):
        X = np.empty((self.n, self.m))
        check(capi.lib().de_mv_download_rowmajor(self._h, dptr(X)), self.ctx._h)
        return X

    def copy_from(self, other):
        check(capi.lib().de_mv_copy(self._h, other._h), self.ctx._h)

    def device_ptr(self):
        p = C.c_void_p()
        check(capi.lib().de_mv_device_ptr(self._h, C.byref(p)))
        return p.value

    def close(self):
        if self._h:
            capi.lib().de_mv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostFactorization:
    """One-time host factorisation filling the UMFPACK field contract (reference umfpacktools.hh:26-44):
    stand-in for UMFPackFactorizedMatrix's constructor where UMFPACK is not installed."""

    def __init__(self, A, ordering=1, scale_rows=False, spd=False, nthreads=0, arrays=True):
        """spd: supernodal multifrontal Cholesky (C ABI de_host_factorize_spd; the matrix must be symmetric positive
        definite) instead of the scalar LU; arrays=False skips the expansion into the explicit L / U arrays of the
        UMFPACK contract (large factors: upload with Factor.from_host)."""
        rp, ci, v = _csr(A)
        self._h = C.c_void_p()
        self.supernodal = bool(spd)
        if spd:
            check(capi.lib().de_host_factorize_spd(len(rp) - 1, i64ptr(rp), i64ptr(ci), dptr(v), ordering, int(nthreads),
                                                   C.byref(self._h)))
            sn, n_, lnz_, st_, fl_ = C.c_int(0), C.c_int64(), C.c_int64(), C.c_int64(), C.c_double()
            sec = (C.c_double * 3)()
            check(capi.lib().de_host_factor_info(self._h, C.byref(sn), C.byref(n_), C.byref(lnz_), C.byref(st_),
                                                 C.byref(fl_), sec))
            self.info = {"n": n_.value, "lnz": lnz_.value, "stored": st_.value, "flops": fl_.value,
                         "seconds_ordering": sec[0], "seconds_symbolic": sec[1], "seconds_numeric": sec[2]}
            self.n, self.lnz, self.unz = n_.value, lnz_.value, lnz_.value
            if not arrays:
                return
        else:
            check(capi.lib().de_host_factorize(len(rp) - 1, i64ptr(rp), i64ptr(ci), dptr(v), ordering, int(scale_rows),
                                               C.byref(self._h)))
        n, lnz, unz, rec = C.c_int64(), C.c_int64(), C.c_int64(), C.c_long()
        ptrs = [C.POINTER(C.c_long)(), C.POINTER(C.c_long)(), C.POINTER(C.c_double)(), C.POINTER(C.c_long)(),
                C.POINTER(C.c_long)(), C.POINTER(C.c_double)(), C.POINTER(C.c_long)(), C.POINTER(C.c_long)(),
                C.POINTER(C.c_double)()]
        check(capi.lib().de_host_factor_arrays(self._h, C.byref(n), C.byref(lnz), C.byref(unz),
                                               *[C.byref(p) for p in ptrs], C.byref(rec)))
        self.n, self.lnz, self.unz, self.do_recip = n.value, lnz.value, unz.value, rec.value
        sizes = [self.n + 1, self.lnz, self.lnz, self.n + 1, self.unz, self.unz, self.n, self.n, self.n]
        names = ["Lp", "Lj", "Lx", "Up", "Ui", "Ux", "P", "Q", "Rs"]
        for name, p, s in zip(names, ptrs, sizes):
            setattr(self, name, np.ctypeslib.as_array(p, shape=(s,)).copy() if s > 0 else
                    np.zeros(0, dtype=np.float64 if name in ("Lx", "Ux", "Rs") else np.int64))

    def arrays(self):
        return {k: getattr(self, k) for k in ("Lp", "Lj", "Lx", "Up", "Ui", "Ux", "P", "Q", "Rs", "do_recip")}

    def close(self):
        if self._h:
            capi.lib().de_host_factor_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


AUTO_CHOLESKY_FROM_ROWS = 10000


def host_factorization(A, ordering=1, factorization="auto", nthreads=0):
    """the host provider the drivers use: see GeneralizedInverse"""
    rp, ci, v = _csr(A)
    n = len(rp) - 1
    if factorization == "auto":
        factorization = "lu"
        if n >= AUTO_CHOLESKY_FROM_ROWS:
            import scipy.sparse as sp

            S = sp.csr_matrix((v, ci, rp), shape=(n, n))
            D = abs(S - S.T)
            if D.nnz == 0 or D.max() <= 1e-13 * abs(S).max():
                try:
                    return HostFactorization((rp, ci, v), ordering, spd=True, nthreads=nthreads, arrays=False)
                except DeError as e:
                    if e.status != capi.DE_ERR_SINGULAR:
                        raise
    if factorization == "cholesky":
        return HostFactorization((rp, ci, v), ordering, spd=True, nthreads=nthreads, arrays=False)
    return HostFactorization((rp, ci, v), ordering)


class Factor:
    """Device copy of a factorisation in the UMFPACK field contract, with its level schedules."""

    def __init__(self, ctx, F):
        """F: HostFactorization or dict with Lp,Lj,Lx,Up,Ui,Ux,P,Q,Rs,do_recip."""
        self.ctx = ctx
        self._h = C.c_void_p()
        if isinstance(F, HostFactorization) and F.supernodal:
            # supernodal Cholesky factor: uploaded in its own form (dense panels; csrc/kernels_snode.cuh)
            check(capi.lib().de_factor_upload_host(ctx._h, F._h, C.byref(self._h)), ctx._h)
            self.n = F.n
            return
        if isinstance(F, HostFactorization):
            F = F.arrays()
        ia = {k: np.ascontiguousarray(F[k], dtype=np.int64) for k in ("Lp", "Lj", "Up", "Ui", "P", "Q")}
        da = {k: f64(F[k]) for k in ("Lx", "Ux", "Rs")}
        n = len(ia["Lp"]) - 1
        check(capi.lib().de_factor_upload(ctx._h, n, lptr(ia["Lp"]), lptr(ia["Lj"]), dptr(da["Lx"]), lptr(ia["Up"]),
                                          lptr(ia["Ui"]), dptr(da["Ux"]), lptr(ia["P"]), lptr(ia["Q"]), dptr(da["Rs"]),
                                          int(F["do_recip"]), C.byref(self._h)), ctx._h)
        self.n = n

    def info(self):
        n, lnz, unz, ll, lu = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int(), C.c_int()
        check(capi.lib().de_factor_info(self._h, C.byref(n), C.byref(lnz), C.byref(unz), C.byref(ll), C.byref(lu)))
        return {"n": n.value, "lnz": lnz.value, "unz": unz.value, "levels_L": ll.value, "levels_U": lu.value}

    def close(self):
        if self._h:
            capi.lib().de_factor_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- kernels: reference names (kernels_cpp.hh) --------------------------------------------------------------
def matmul_sparse_tallskinny_blocked(Qout, A, Qin):
    """Qout = A * Qin (kernels_cpp.hh:626-657)."""
    check(capi.lib().de_spmm(Qout._h, A._h, Qin._h), A.ctx._h)


def matmul_sparse_tallskinny_with_dots_and_gram(Qout, A, Qin):
    """Qout = A * Qin, diag(Qin^T Qout) and the Gram matrix Qout^T Qout from the same pass (de_spmm_gram)."""
    dp, G = np.empty(Qin.m), np.empty((Qin.m, Qin.m))
    check(capi.lib().de_spmm_gram(Qout._h, A._h, Qin._h, dptr(dp), dptr(G)), A.ctx._h)
    return dp, G


def matmul_sparse_tallskinny_with_dots(Qout, A, Qin):
    """Qout = A * Qin and the column-wise dot products diag(Qin^T Qout) from the same pass."""
    dp = np.empty(Qin.m)
    check(capi.lib().de_spmm_diag_dot(Qout._h, A._h, Qin._h, dptr(dp)), A.ctx._h)
    return dp


def dot_products_diagonal_blocked(Q1, Q2):
    """dp[j] = <Q1[:,j], Q2[:,j]> (kernels_cpp.hh:24-55); returns dp (the reference resizes its out-argument)."""
    dp = np.empty(Q1.m)
    check(capi.lib().de_diag_dot(dptr(dp), Q1._h, Q2._h), Q1.ctx._h)
    return dp


def dot_products_all_blocked(Q1, Q2):
    """Full Gram matrix Q1^T Q2 (kernels_cpp.hh:58-96)."""
    G = np.empty((Q1.m, Q1.m))
    check(capi.lib().de_gram(dptr(G), Q1._h, Q2._h), Q1.ctx._h)
    return G


def block_update(Q, R):
    """Q <- Q R (kernels_cpp.hh:293-305, :514-539)."""
    R = f64(R)
    if R.shape != (Q.m, Q.m):
        raise DeError(capi.DE_ERR_INVALID, "block_update: factor must be m x m")
    check(capi.lib().de_block_update(Q._h, dptr(R)), Q.ctx._h)


def block_project(Q, j0, k0, S):
    """Q[:, j0:j0+w] -= Q[:, k0:k0+w] S (kernels_cpp.hh:335-348)."""
    S = f64(S)
    check(capi.lib().de_block_project(Q._h, j0, k0, S.shape[0], dptr(S)), Q.ctx._h)


def orthonormalize_blocked(Q):
    """In-place thin QR with positive-diagonal triangular factor (kernels_cpp.hh:180-351)."""
    check(capi.lib().de_orthonormalize(Q._h), Q.ctx._h)


def B_orthonormalize_blocked(B, Q, BQ=None):
    """In-place B-orthonormalisation (kernels_cpp.hh:356-591); returns the `norm` diagnostic."""
    nrm = C.c_double(0.0)
    check(capi.lib().de_b_orthonormalize(B._h, Q._h, BQ._h if BQ is not None else None, C.byref(nrm)), Q.ctx._h)
    return nrm.value


def matmul_inverse_tallskinny_blocked(Qout, F, Qin):
    """Qout = A^-1 Qin through the factors; Qin is scratch (kernels_cpp.hh:660-755)."""
    check(capi.lib().de_factor_apply(Qout._h, F._h, Qin._h), F.ctx._h)


# ---- drivers: reference names (eigensolver.hh) ---------------------------------------------------------------
def _add_to_diagonal(rp, ci, v, s):
    n = len(rp) - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
    v[rows == ci] += s


class Result:
    def __init__(self, eval_, evec, iterations, relerror=None, time_factorization=None):
        self.eval, self.evec, self.iterations = eval_, evec, iterations
        self.relerror, self.time_factorization = relerror, time_factorization


def StandardLargest(ctx, A, shift, tol, maxiter, nev, verbose=0, seed=123, start=None):
    """reference StandardLargest (eigensolver.hh:28-112). `A = (rowptr, col, val)`; like the reference the
    caller's values are shifted in place when shift != 0."""
    rp, ci, v = A if isinstance(A, tuple) else (A.indptr, A.indices, A.data)
    n = len(rp) - 1
    m = padded_cols(nev)
    if start is None:
        start = start_block(n, m, seed)
    if shift != 0.0:
        _add_to_diagonal(np.asarray(rp), np.asarray(ci), v, shift)
    dA = Matrix(ctx, (rp, ci, v))
    try:
        ev, V, it = np.zeros(nev), np.zeros((nev, n)), C.c_int(0)
        check(capi.lib().de_standard_largest(ctx._h, dA._h, shift, tol, maxiter, nev, dptr(f64(start)), dptr(ev),
                                             dptr(V), verbose, C.byref(it)), ctx._h)
    finally:
        dA.close()
    return Result(ev, V, it.value)


def standard_largest_mv(ctx, dA, shift, tol, maxiter, Q, verbose=0):
    """Device-resident StandardLargest: Q (MultiVector) holds the start block on entry, the eigenvector block on
    return. Returns (eval[m], iterations)."""
    ev, it = np.zeros(Q.m), C.c_int(0)
    check(capi.lib().de_standard_largest_mv(ctx._h, dA._h, shift, tol, maxiter, Q._h, dptr(ev), verbose, C.byref(it)),
          ctx._h)
    return ev, it.value


def standard_inverse_mv(ctx, dA, dF, shift, tol, maxiter, Q, verbose=0):
    ev, it = np.zeros(Q.m), C.c_int(0)
    check(capi.lib().de_standard_inverse_mv(ctx._h, dA._h, dF._h, shift, tol, maxiter, Q._h, dptr(ev), verbose,
                                            C.byref(it)), ctx._h)
    return ev, it.value


def StandardInverse(ctx, A, shift, tol, maxiter, nev, verbose=0, seed=123, start=None, ordering=1, factorization="auto"):
    """reference StandardInverse (eigensolver.hh:116-198)."""
    rp, ci, v = A if isinstance(A, tuple) else (A.indptr, A.indices, A.data)
    n = len(rp) - 1
    m = padded_cols(nev)
    if start is None:
        start = start_block(n, m, seed)
    if shift != 0.0:
        _add_to_diagonal(np.asarray(rp), np.asarray(ci), v, shift)
    hF = host_factorization((rp, ci, v), ordering, factorization)
    dA = Matrix(ctx, (rp, ci, v))
    dF = Factor(ctx, hF)
    try:
        ev, V, it = np.zeros(nev), np.zeros((nev, n)), C.c_int(0)
        check(capi.lib().de_standard_inverse(ctx._h, dA._h, dF._h, shift, tol, maxiter, nev, dptr(f64(start)),
                                             dptr(ev), dptr(V), verbose, C.byref(it)), ctx._h)
    finally:
        dA.close()
        dF.close()
        hF.close()
    return Result(ev, V, it.value)


def GeneralizedInverse(ctx, inA, B, shift, reg, tol, maxiter, nev, verbose=0, seed=123, start=None, ordering=1,
                       factorization="auto", nthreads=0):
    """reference GeneralizedInverse (eigensolver.hh:204-351): A x = lambda B x by shift-invert subspace iteration.
    The input matrix is copied (eigensolver.hh:208); pattern(B) must be contained in pattern(A).
    factorization: "lu" (scalar sparse LU filling the UMFPACK contract), "cholesky" (supernodal multifrontal Cholesky
    for symmetric positive definite A + shift B: the provider for 3D problems) or "auto" (Cholesky for symmetric
    matrices with at least 10 000 rows, falling back to LU if a pivot is not positive -- what the C++ header does)."""
    import time

    rpa, cia, va = inA if isinstance(inA, tuple) else (inA.indptr, inA.indices, inA.data)
    rpb, cib, vb = B if isinstance(B, tuple) else (B.indptr, B.indices, B.data)
    rpa, cia, va = i64(rpa), i64(cia), f64(va).copy()
    rpb, cib, vb = i64(rpb), i64(cib), f64(vb)
    n = len(rpa) - 1
    m = padded_cols(nev)
    if start is None:
        start = start_block(n, m, seed)
    if shift != 0.0:  # A.axpy(shift, B) (eigensolver.hh:241-242)
        if len(cia) == len(cib) and np.array_equal(rpa, rpb) and np.array_equal(cia, cib):
            va += shift * vb
        else:
            rows_b = np.repeat(np.arange(n, dtype=np.int64), np.diff(rpb))
            keys_a = np.repeat(np.arange(n, dtype=np.int64), np.diff(rpa)) * n + cia
            pos = np.searchsorted(keys_a, rows_b * n + cib)
            if np.any(pos >= len(keys_a)) or np.any(keys_a[np.minimum(pos, len(keys_a) - 1)] != rows_b * n + cib):
                raise DeError(capi.DE_ERR_INVALID, "GeneralizedInverse: pattern of B not contained in pattern of A")
            np.add.at(va, pos, shift * vb)
    if reg != 0.0:
        _add_to_diagonal(rpa, cia, va, reg)
    t0 = time.perf_counter()
    hF = host_factorization((rpa, cia, va), ordering, factorization, nthreads)
    t_fact = time.perf_counter() - t0
    dA, dB = Matrix(ctx, (rpa, cia, va)), Matrix(ctx, (rpb, cib, vb))
    dF = Factor(ctx, hF)
    try:
        ev, V, it, rel = np.zeros(nev), np.zeros((nev, n)), C.c_int(0), C.c_double(0.0)
        check(capi.lib().de_generalized_inverse(ctx._h, dA._h, dB._h, dF._h, shift, tol, maxiter, nev,
                                                dptr(f64(start)), dptr(ev), dptr(V), verbose, C.byref(it),
                                                C.byref(rel)), ctx._h)
    finally:
        dA.close()
        dB.close()
        dF.close()
        hF.close()
    if verbose > 0:  # the reference's summary line (eigensolver.hh:344-350)
        print("GeneralizedInverse:  time_factorization=%g iterations=%d relerror=%g" % (t_fact, it.value, rel.value))
    res = Result(ev, V, it.value, rel.value, t_fact)
    res.factor_info = getattr(hF, "info", None)
    return res


# ---- LOBPCG drivers (new: the reference has none; SURVEY.md §8f rank 1, BASELINE.json configs[1]) ----------------
def StandardLOBPCG(ctx, A, tol, maxiter, nev, verbose=0, seed=123, start=None):
    """nev smallest eigenpairs of A x = lambda x without a factorisation. Parameter shape of the reference's drivers
    (eigensolver.hh:28-29): same start block, m = nev rounded up to 8, eval / evec as in StandardLargest.
    Convergence: ||A x - theta x||_2 <= tol |theta| for every wanted pair."""
    rp, ci, v = A if isinstance(A, tuple) else (A.indptr, A.indices, A.data)
    n = len(rp) - 1
    if start is None:
        start = start_block(n, padded_cols(nev), seed)
    dA = Matrix(ctx, (rp, ci, v))
    try:
        ev, V, it = np.zeros(nev), np.zeros((nev, n)), C.c_int(0)
        check(capi.lib().de_standard_lobpcg(ctx._h, dA._h, tol, maxiter, nev, dptr(f64(start)), dptr(ev), dptr(V),
                                            verbose, C.byref(it)), ctx._h)
    finally:
        dA.close()
    return Result(ev, V, it.value)


def GeneralizedLOBPCG(ctx, A, B, tol, maxiter, nev, verbose=0, seed=123, start=None):
    """nev smallest eigenpairs of A x = lambda B x (B symmetric positive definite) without a factorisation; the
    eigenvectors are B-orthonormal like GeneralizedInverse's (eigensolver.hh:204-351)."""
    rpa, cia, va = A if isinstance(A, tuple) else (A.indptr, A.indices, A.data)
    rpb, cib, vb = B if isinstance(B, tuple) else (B.indptr, B.indices, B.data)
    n = len(rpa) - 1
    if start is None:
        start = start_block(n, padded_cols(nev), seed)
    dA, dB = Matrix(ctx, (rpa, cia, va)), Matrix(ctx, (rpb, cib, vb))
    try:
        ev, V, it = np.zeros(nev), np.zeros((nev, n)), C.c_int(0)
        check(capi.lib().de_generalized_lobpcg(ctx._h, dA._h, dB._h, tol, maxiter, nev, dptr(f64(start)), dptr(ev),
                                               dptr(V), verbose, C.byref(it)), ctx._h)
    finally:
        dA.close()
        dB.close()
    return Result(ev, V, it.value)


def lobpcg_mv(ctx, dA, Q, tol, maxiter, nev=None, dB=None, dT=None, largest=False, verbose=0, cheb_degree=0):
    """Device-resident LOBPCG with all options (de_lobpcg_mv): Q holds the start block on entry and the m Ritz vectors
    on return; dB: mass matrix or None; dT: a Factor used as preconditioner or None; cheb_degree > 0: Chebyshev
    polynomial preconditioner with that many applications of A per iteration (the drivers' default is 8).
    Returns (eval[m], resnorm[m], iterations, restarts, converged)."""
    nev = Q.m if nev is None else nev
    ev, rn = np.zeros(Q.m), np.zeros(Q.m)
    it, rs, cv = C.c_int(0), C.c_int(0), C.c_int(0)
    check(capi.lib().de_lobpcg_mv(ctx._h, dA._h, dB._h if dB is not None else None, dT._h if dT is not None else None,
                                  1 if largest else 0, int(cheb_degree), tol, maxiter, nev, Q._h, dptr(ev), dptr(rn), verbose,
                                  C.byref(it), C.byref(rs), C.byref(cv)), ctx._h)
    return ev, rn, it.value, rs.value, bool(cv.value)


def block_lincomb(out, sources, coeffs, out2=None):
    """out = sum_s sources[s] @ coeffs[s]; out2 (optional) = the same sum without sources[0] -- one fused pass
    (de_block_lincomb). out may be sources[0]; out2 may be sources[1] or sources[2]."""
    ns = len(sources)
    Cm = f64(np.stack([f64(c) for c in coeffs]))
    arr = (C.c_void_p * ns)(*[s._h for s in sources])
    check(capi.lib().de_block_lincomb(out._h, out2._h if out2 is not None else None, ns, arr, dptr(Cm)), out.ctx._h)


def host_sym_eig(A):
    """(w ascending, V with eigenvector j in column j) of a dense symmetric matrix -- the host Rayleigh-Ritz solver."""
    A = f64(A)
    n = A.shape[0]
    w, V = np.zeros(n), np.zeros((n, n))
    check(capi.lib().de_host_sym_eig(n, dptr(A), dptr(w), dptr(V)))
    return w, V


def host_sym_gen_eig(GA, GB):
    """(w ascending, C with C^T GB C = I, smallest scaled Cholesky pivot) of GA c = w GB c."""
    GA, GB = f64(GA), f64(GB)
    n = GA.shape[0]
    w, Cm, piv = np.zeros(n), np.zeros((n, n)), C.c_double(0.0)
    check(capi.lib().de_host_sym_gen_eig(n, dptr(GA), dptr(GB), dptr(w), dptr(Cm), C.byref(piv)))
    return w, Cm, piv.value
