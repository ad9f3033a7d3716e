"""GPU, needs >= 2 devices: launches tests/dist_worker.py under torchrun (N-rank result == 1-rank result)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exchange", ["peer", "nccl", "peer-sums+nccl"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_ranks_match_one_rank(precision, exchange):
    """exchange = peer: the one-shot all-reduce over NVLink peer memory (marf_peer_allreduce, symmetric memory);
    nccl: dist.all_reduce (MARF_NCCL_ALLREDUCE=1); peer-sums+nccl: the mix 8 ranks use."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(ROOT, "tests", "dist_worker.py"), "--precision", precision]
    env = dict(os.environ)
    env.pop("MARF_NCCL_ALLREDUCE", None)
    env.pop("MARF_PEER_ALLREDUCE_MAX_WORLD", None)
    if exchange == "nccl":
        env["MARF_NCCL_ALLREDUCE"] = "1"
    elif exchange == "peer-sums+nccl":                  # what 8 ranks use: loss sums over peer memory, gradients over NCCL
        env["MARF_PEER_ALLREDUCE_MAX_WORLD"] = "1"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0 and "all cases OK" in r.stdout
    assert f"exchange={exchange} " in r.stdout
    for other in ("peer", "nccl", "peer-sums+nccl"):
        assert other == exchange or f"exchange={other} " not in r.stdout
