"""Multi-GPU path: slab partition + NCCL neighbour exchange (needs >= 2 GPUs; run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["full", "edge"])
@pytest.mark.parametrize("nproc", [2, 4])
def test_slab_exchange_matches_single_domain(nproc, mode):
    import torch

    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1", "--master-port",
           str(29500 + nproc + (10 if mode == "edge" else 0)), os.path.join(ROOT, "tests", "multirank_worker.py"), mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("MULTIRANK_OK") == 3, out.stdout[-3000:]
