"""GPU (-m gpu), needs >= 2 devices (skipped on a 1-GPU box): multi-GPU sessions over NCCL give factors
bit-identical to the single-GPU run (scripts/dist_check.py under torch.distributed.run)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("shape", ["small", "ml100k"])
def test_multi_gpu_equals_single_gpu(gpu, shape):
    n = gpu.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "scripts", "dist_check.py"), shape]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert out.returncode == 0 and lines, out.stdout[-3000:]
    rep = json.loads(lines[-1])
    assert rep["ok"] and rep["ccd_bitwise_equal_to_1gpu"] and rep["als_bitwise_equal_to_1gpu"], rep


def test_sessions_in_a_row_share_the_cached_peer_state(gpu):
    """Four multi-GPU CCD++ sessions on one NCCL unique id (cache hit, another shape, back again): every one bit-identical
    to the single-GPU run (scripts/dist_cache_check.py)."""
    n = gpu.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29518", os.path.join(ROOT, "scripts", "dist_cache_check.py")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert out.returncode == 0 and lines, out.stdout[-3000:]
    rep = json.loads(lines[-1])
    assert rep["ok"] and all(rep["sessions_bitwise_equal_to_1gpu"]), rep
