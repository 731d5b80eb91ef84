"""Multi-GPU data path (NCCL halo exchange over NVLink + all-reduced reductions), one process per GPU via torchrun.
Needs >= 2 GPUs on the box; the host-side logic is covered on CPU by tests/test_distributed_cpu.py."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_gpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_partitioned_solve(world):
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    try:  # keep the workers' output where gpurun brings it back
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "mgpu_worker_%d.log" % world), "w") as f:
            f.write(out.stdout[-20000:] + "\n---- stderr ----\n" + out.stderr[-20000:])
    except OSError:
        pass
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "fails=0" in out.stdout
