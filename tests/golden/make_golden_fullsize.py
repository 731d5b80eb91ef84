"""Generates tests/golden/reference_fullsize.npz: the REFERENCE ITSELF (oracle/_ref = its headers compiled verbatim) run
on BASELINE.json's configuration 2 at full size -- 3D Q1 27-point Laplace stiffness matrix, 100^3 nodes, 32 eigenpairs,
StandardLargest (reference eigensolver.hh:28-112), tol = 2e-3, maxiter = 4000, seed = 123, shift = 0 -- and on the 7-point
variant. Stored: iteration count, the 32 Rayleigh quotients, and per eigenvector the residual norm and a 64-entry
sample, enough to pin the GPU path at the size bench.py measures. Takes ~2 minutes of one CPU core per case.

    make -C oracle ref && python tests/golden/make_golden_fullsize.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from dune_eigensolver_b200 import matrices as M  # noqa: E402


def main():
    ref = O.load_reference()
    if ref is None or ref.kind != "reference":
        raise SystemExit("oracle/_ref/libde_reference.so missing: run `make -C oracle ref` where /root/reference exists")
    out = {}
    for tag, gen in (("q1", M.q1_stiffness), ("fd", M.laplacian_fd)):
        N, nev, tol = 100, 32, 2e-3
        A = gen((N, N, N))
        rp, ci, v = (np.ascontiguousarray(A[0], dtype=np.int64), np.ascontiguousarray(A[1], dtype=np.int64),
                     np.ascontiguousarray(A[2]))
        t0 = time.time()
        ev, V, k = ref.standard_largest((rp, ci, v.copy()), 0.0, tol, 4000, nev)
        print(tag, "iterations", k, "seconds", round(time.time() - t0, 1), "ev[:4]", ev[:4])
        S = M.to_scipy((rp, ci, v))
        V = np.asarray(V)  # nev x n
        res = np.array([np.linalg.norm(S @ V[j] - ev[j] * V[j]) for j in range(nev)])
        idx = np.linspace(0, N ** 3 - 1, 64).astype(np.int64)
        out[tag + "_N"], out[tag + "_nev"], out[tag + "_tol"] = N, nev, tol
        out[tag + "_iterations"] = k
        out[tag + "_eval"] = ev
        out[tag + "_residual"] = res
        out[tag + "_sample_idx"] = idx
        out[tag + "_sample"] = V[:, idx]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference_fullsize.npz"), **out)
    print("written")


if __name__ == "__main__":
    main()
