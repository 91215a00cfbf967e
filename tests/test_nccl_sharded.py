"""2-rank NCCL run of the sample-sharded product path (SURVEY.md §8e, Appendix D "2/4/8-rank sharded vs single-rank"):
replicas bit-identical after every iteration, and within 1e-6 (one step) of the single-rank run on the concatenated
rows.  Needs two GPUs on the box: `gpurun --gpus 2 -- python -m pytest tests/test_nccl_sharded.py -m gpu`."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2])
def test_sharded_vbem_over_nccl_matches_single_rank(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "_nccl_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "NCCL_SHARDED_OK" in out.stdout, (out.stdout[-3000:], out.stderr[-3000:])
