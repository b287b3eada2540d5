"""N>1 on real GPUs: NCCL all-gather + merge kernel vs the unsharded oracle (needs >= 2 GPUs on the box)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_search_on_all_visible_gpus_equals_unsharded_oracle():
    import torch
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs at least 2 GPUs")
    g = 8 if g >= 8 else (4 if g >= 4 else 2)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={g}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count("True") == 10 and "False" not in p.stdout, p.stdout[-3000:]
