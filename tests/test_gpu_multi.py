"""GPU, >= 2 devices: N-GPU vs 1-GPU / oracle agreement through torchrun (skips on a 1-GPU box; bench.py's
N > 1 legs carry their own checks -- field_hash, subcube_vs_oracle, result_ok -- so that the driver's scaling
run verifies the multi-GPU paths even where this test cannot run).  Run twice: with the peers' memory
mapped (in-kernel combine / in-kernel halos) and with PH_NO_P2P=1 (the NCCL forms of the same entry points)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_multi_gpu_agreement(transport):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    env = dict(os.environ)
    if transport == "nccl":
        env["PH_NO_P2P"] = "1"
    port = "29631" if transport == "p2p" else "29632"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", port, os.path.join(ROOT, "tests", "mgpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert "MGPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
    want = "p2p=True" if transport == "p2p" else "p2p=False"
    assert want in out.stdout, out.stdout[-500:]
