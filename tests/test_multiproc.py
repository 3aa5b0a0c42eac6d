"""The N>1 path on CPU: world_size-2 (and 4) gloo runs of the host side -- index handshake between
processes (comm_data.c:195-250), unified device numbering and the per-peer packed exchange plan."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,ndomains", [(2, 8), (2, 12), (4, 8)])
def test_gloo_handshake_and_exchange_plan(world, ndomains, tmp_path):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "_mp_worker.py"), str(ndomains)]
    env = dict(os.environ, OMP_NUM_THREADS="2", CUDA_VISIBLE_DEVICES="", CFDP_MP_OUT=str(tmp_path))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    results = [json.load(open(tmp_path / f"rank{i}.json")) for i in range(world)]
    assert len(results) == world
    for res in results:
        assert res["errors"] == []
        assert res["remote"] > 0 and len(res["peers"]) >= 1
        if ndomains // world > 1:
            assert res["local"] > 0        # neighbours hosted by the same process are copied on the device
