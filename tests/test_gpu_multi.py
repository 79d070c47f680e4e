"""Multi-GPU parity: rc_comm_all_reduce / rc_reduce_all_sharded / rc_reduce_axes_sharded against the
oracle on the unsharded array.  The work happens in tests/multi_worker.py, one process per GPU under torchrun.

  * world size 1 always runs on a GPU box (the communicator, the peer-window kernel and the fused
    "second pass + combine" path are all exercised with one rank);
  * world size 2 (and 4 when the box has them) runs when the box has that many GPUs -- both transports:
    the NVLink peer window (default) and NCCL only (RC_COMM_PEER=0).
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "multi_worker.py")


def _ngpu():
    import rstsr_b200 as rt
    try:
        return rt.DeviceCuda.device_count()
    except Exception:
        return 0


def _run(world, peer=True, port=29611):
    env = dict(os.environ)
    env["RC_COMM_PEER"] = "1" if peer else "0"
    env["RC_COMM_TIMEOUT_S"] = "60"
    env.pop("OMP_NUM_THREADS", None)
    if world == 1:
        cmd = [sys.executable, WORKER]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTI_OK" in r.stdout, f"worker failed\nstdout:\n{r.stdout[-3000:]}\nstderr:\n{r.stderr[-3000:]}"
    return r.stdout


@pytest.mark.gpu
def test_sharded_reductions_world1():
    out = _run(1)
    assert "world 1" in out


@pytest.mark.gpu
@pytest.mark.parametrize("peer", [True, False], ids=["peer-window", "nccl"])
def test_sharded_reductions_world2(peer):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    out = _run(2, peer=peer, port=29611 if peer else 29612)
    assert "world 2" in out and f"peer_window {peer}" in out


@pytest.mark.gpu
def test_sharded_reductions_world4():
    if _ngpu() < 4:
        pytest.skip("needs 4 GPUs")
    out = _run(4, port=29613)
    assert "world 4" in out
