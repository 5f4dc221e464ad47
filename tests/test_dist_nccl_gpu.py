"""NCCL data-parallel correctness: batch 8 on one GPU vs 2 x 4 -> equal losses and equal flat gradient
buffers after the (bucketed, overlapped or single) all-reduce, eager and CUDA-graph steps.  Needs >= 2 GPUs (skipped on a 1-GPU box); launched as the
driver launches bench.py: python -m torch.distributed.run --nproc-per-node 2 on 127.0.0.1."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("graph,overlap", [("0", "1"), ("1", "1"), ("1", "0")])
def test_two_rank_nccl_matches_single_gpu(graph, overlap):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    env = dict(os.environ, SPE_TEST_GRAPH=graph, SPE_AR_OVERLAP=overlap)     # overlap: bucketed all-reduce inside the step
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29731 + 2 * int(graph) + int(overlap)), os.path.join(ROOT, "tests", "_nccl_parity_worker.py")]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    assert r.stdout.count("-> OK") == 2, r.stdout[-3000:]
