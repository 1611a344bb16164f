"""Vocab-sharded multi-GPU path: every collective variant (NCCL, peer-memory triples, one-shot P2P all-reduce,
reduce-scatter fused into the K2a epilogue) against the oracle and a torch fp32 reference at the real head
(tests/multi_worker.py).  Needs >= 2 GPUs on the box (skipped on the 1-GPU boxes; the round-2 hardware logs of the
2- and 8-GPU runs are profiles/r2_multi_gpu_tests_n*.log)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_vocab_sharded_step_matches_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "multi_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(p.stdout[-4000:], p.stderr[-3000:])
    assert p.returncode == 0
