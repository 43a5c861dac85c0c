"""Level sharding across the GPUs of one node, on hardware (kernel K6): a 256-bit mul, an encrypted-amount
shift and a min with their wide PBS levels cut over 2 ranks, decrypted and checked on every rank - once through the
library's own peer-mapped exchange (fsc_peer_pool_*: the blind rotation's epilogue stores into both pools) and once
through the NCCL all-gather callback (fsc_set_level_exchange).  Needs 2 visible GPUs (gpurun --gpus 2); skipped otherwise."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_level_sharding_two_ranks_decrypts_256_bit_ops(exchange):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs on the node")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "multi_gpu_ops.py"), "--exchange", exchange, "--quick"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    lines = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 3 and all(ln["correct"] for ln in lines), lines
    assert all(ln["n_gpus"] == 2 and ln["exchange"] == exchange for ln in lines)
    assert lines[0]["sharded_levels"] > 0, "the 16 384-wide partial-product level of a 256-bit mul must have been sharded"
