import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", sys.argv[1] if len(sys.argv) > 1 else "32")
sys.path.insert(0, "/root/repo")
import numpy as np
from dune_eigensolver_b200 import eigensolver as E, matrices as M
shape = (7, 6, 9)
A = M.q1_stiffness(shape)
dense = np.linalg.eigvalsh(M.to_scipy(A).toarray())[:8]
ok = bad = 0
for trial in range(1):
    for ranks in (4,):
        mg = E.Multi([0] * ranks, timeout_s=3)
        t0 = time.time()
        try:
            r = mg.StandardLOBPCG(A, 1e-9, 2000, 8, row_align=42, verbose=0)
            assert np.abs(r.eval - dense).max() <= 1e-9 * np.abs(dense).max()
            r2 = mg.StandardLargest((A[0], A[1], A[2].copy()), 0.0, 1e-9, 3000, 8, row_align=42)
            ok += 1
        except Exception as e:
            bad += 1
            print("trial", trial, "ranks", ranks, "FAILED after %.1f s:" % (time.time() - t0), str(e)[:1500], flush=True)
        finally:
            mg.close()
print("connections", os.environ["CUDA_DEVICE_MAX_CONNECTIONS"], "ok", ok, "bad", bad)
