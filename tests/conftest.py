import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Tests that put several ranks of the single-process multi-GPU front end on ONE device (tests/test_multi_inprocess_gpu.py
# on a one-GPU box) need every rank's stream on a hardware work queue of its own: a kernel spinning on a peer's flag
# must never sit in front of that peer's kernels. The default of 8 queues is shared with torch's and the other test
# contexts' streams. Must be set before CUDA initialises; irrelevant with one rank per GPU (the production layout).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracles():
    """(reference-or-None, port): oracle/_ref is the reference compiled verbatim, the port is the restatement."""
    import subprocess

    from oracle import oracle as O

    if not os.path.exists(O.PORT_LIB) or (os.path.isdir("/root/reference") and not os.path.exists(O.REF_LIB)):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], check=True, capture_output=True)
    return O.load_reference(), O.load_port()


@pytest.fixture(scope="session")
def oracle(oracles):
    """the checker: the compiled reference when present, else the port (pinned to it by test_oracle.py)."""
    return oracles[0] if oracles[0] is not None else oracles[1]


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


@pytest.fixture(scope="session")
def ctx():
    """A GPU context. No skip and no fallback: a -m gpu run without the CUDA library or device must fail."""
    from dune_eigensolver_b200 import eigensolver as E

    c = E.Context(0)
    yield c
    c.close()
