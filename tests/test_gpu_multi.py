"""Frame sharding over several GPUs gives bit-identical histograms: two NCCL ranks against one, when the box
has at least two GPUs (skipped otherwise; the 2-rank host logic is covered on CPU with gloo in tests/test_host_logic.py)."""
import json
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_two_ranks_equal_one(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs, %d visible" % (world, torch.cuda.device_count()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29581", os.path.join(ROOT, "tests", "tools", "dist_identity.py"), "--cells", "12", "--frames", "7"]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert res["n_gpus"] == world and res["hist_equals_single_rank"] and res["rows_equal_single_rank"] and res["frame0_parity"]


def test_single_rank_identity_script_runs():
    """The same script on one GPU (no process group): the check itself, and the frame-0 oracle parity, on every box."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "dist_identity.py"), "--cells", "8", "--frames", "3"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert res["n_gpus"] == 1 and res["hist_equals_single_rank"] and res["frame0_parity"]
