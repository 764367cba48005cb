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


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_cpp_sharded_spec_one_process_per_gpu(transport, tmp_path):
    """tests/cpp/sharded_spec (Phase::ShardedNArray of include/ph_sharded.hpp vs the undivided array), one process
    per GPU; rank 0 hands the NCCL id to the others through a file."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    exe = os.path.join(ROOT, "tests", "cpp", "sharded_spec")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp")], check=True, capture_output=True)
    procs = []
    for r in range(n):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(n), PH_ID_FILE=str(tmp_path / "nccl_id"))
        if transport == "nccl":
            env["PH_NO_P2P"] = "1"
        procs.append(subprocess.Popen([exe], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "0 failed" in out, f"rank {r}:\n{out[-3000:]}"
    assert ("p2p=1" if transport == "p2p" else "p2p=0") in outs[0]


def test_cpp_sharded_spec_single_process():
    """The same spec with world = 1 (runs on the 1-GPU test box): every sharded operation degenerates to its local form."""
    exe = os.path.join(ROOT, "tests", "cpp", "sharded_spec")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp")], check=True, capture_output=True)
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "0 failed" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]


def test_agreement_checks_on_one_rank():
    """Every check of mgpu_check.py with a communicator of ONE rank (runs on the 1-GPU test box): slab runs with no
    neighbour, record-mode reductions, ShardedNArray operators / slicing / scatter / permute in their local forms --
    the host plans and kernels are the ones the N-GPU runs use, only the exchanges are empty."""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import ph_core_b200 as ph\nfrom ph_core_b200 import sharding as S\nimport mgpu_check\n"
            "ph.init(0); S.comm_init(None)\nr = mgpu_check.run_checks(1, 0)\nprint('ONE_RANK_OK', r['checks'])\n"
            % (ROOT, os.path.join(ROOT, "tests")))
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert "ONE_RANK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
