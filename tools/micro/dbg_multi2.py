"""rank-level primitives on the contexts of a de_multi (one Python thread per rank; ctypes releases the GIL)"""
import ctypes as C, os, sys, threading, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, "/root/repo")
import numpy as np
from dune_eigensolver_b200 import capi, eigensolver as E

ranks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
mode = sys.argv[2] if len(sys.argv) > 2 else "dot"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
mg = E.Multi([0] * ranks, timeout_s=3)
L = capi.lib()
errs = [None] * ranks

def work(r):
    h = C.c_void_p()
    capi.check(L.de_multi_context(mg._h, r, C.byref(h)))
    ctx = E.Context.__new__(E.Context)
    ctx._h, ctx.device = h, 0
    n, m = 1000 + 100 * r, 8
    X = E.MultiVector(ctx, n, m)
    Y = E.MultiVector(ctx, n, m)
    X.upload(np.full((n, m), 1.0))
    Y.upload(np.full((n, m), 2.0))
    want = 2.0 * sum(1000 + 100 * q for q in range(ranks))
    try:
        for it in range(iters):
            if mode == "dot":
                dp = E.dot_products_diagonal_blocked(X, Y)          # standalone all-reduce of m values
                assert np.allclose(dp, want), (it, dp[:2], want)
            elif mode == "gram":
                G = E.dot_products_all_blocked(X, Y)                # two-operand Gram: plain reduce + standalone all-reduce
                assert np.allclose(G, want), (it, G[0, :2], want)
            elif mode == "ortho":
                X.upload(np.random.default_rng(it).standard_normal((n, m)))
                E.orthonormalize_blocked(X)                         # fused tails
            elif mode == "mix":
                X.upload(np.random.default_rng(it).standard_normal((n, m)))
                E.orthonormalize_blocked(X)
                G = E.dot_products_all_blocked(X, X)
                dp = E.dot_products_diagonal_blocked(X, X)
    except Exception as e:
        errs[r] = "iteration %d: %s" % (it, str(e)[:300])
    X.close(); Y.close()
    ctx._h = C.c_void_p()  # borrowed

th = [threading.Thread(target=work, args=(r,)) for r in range(ranks)]
t0 = time.time()
[t.start() for t in th]; [t.join() for t in th]
print("ranks", ranks, "mode", mode, "iters", iters, "%.2f s" % (time.time() - t0), "errors:", [e for e in errs if e] or "none")
mg.close()
