"""GPU (>= 2 devices; skipped on a one-GPU box): multi-rank parity on real NCCL through `bench.py --check`
(bench_extra.run_check): rank-local logits / loss against the oracle port on that rank's shard, all-reduced gradients
against the mean of the per-shard oracle gradients, graph-captured collective path against the eager one, replicas'
weights bit-identical after optimiser steps (SURVEY 8e "parity method")."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_parity_on_nccl():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "bench.py"), "--check", "--gpus", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stderr[-2000:]
    res = json.loads(lines[-1])
    print(json.dumps(res, indent=1))
    assert res["ok"] and r.returncode == 0
    assert res["fp32"]["rank_local_logits_rel"] < 1e-4 and res["fp32"]["identical_on_all_ranks"]
    assert res["graph_captured"]["replica_weights_bit_identical"]
