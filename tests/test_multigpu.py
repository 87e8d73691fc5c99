"""Point-sharded windows over several GPUs of one box: scripts/multigpu_check.py under torchrun, once
over NVLink peer memory (the product path) and once over the NCCL fallback.  Skipped on a single-GPU box."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _run(n, scale, env_extra, name="c4"):
    env = dict(os.environ, **env_extra)
    env.pop("OMP_NUM_THREADS", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_port()), str(ROOT / "scripts" / "multigpu_check.py"), str(scale), name]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and "MULTIGPU CHECK OK" in r.stdout, (r.stdout + r.stderr)[-4000:]
    return r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["peer", "nccl"])
def test_point_sharded_window_matches_single_gpu_and_oracle(path):
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    _run(2, 0.1, {} if path == "peer" else {"UBA_PEER": "0"})


@pytest.mark.gpu
def test_point_sharded_full_size_all_gpus():
    n = _n_gpus()
    if n < 4:
        pytest.skip("needs at least 4 GPUs")
    _run(n if n in (4, 8) else 4, 1.0, {})
