"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): domains mapped 12/6 per GPU (BASELINE config 3
style), halo exchange over NCCL send/recv, every variant, ghost rows bit-identical to the owners' rows."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world,ndomains", [(2, 8), (2, 24), (4, 12), (8, 24)])
def test_nccl_exchange_bit_identical(world, ndomains, tmp_path):
    if ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_mg_worker.py"), str(ndomains)]
    env = dict(os.environ, OMP_NUM_THREADS="4", CFDP_MP_OUT=str(tmp_path))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for i in range(world):
        res = json.load(open(tmp_path / f"rank{i}.json"))
        assert res["errors"] == []
        assert res["remote"] > 0
