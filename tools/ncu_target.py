#!/usr/bin/env python
"""Minimal launch sequence for ncu: python tools/ncu_target.py MPOINTS ORDER KCFG [ITERS] [TILE_ORDER]
KCFG = version.chunk.persistent.order (see tools/kbench.py); runs ITERS comm_free iterations (gradient kernel only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfd_proxy_b200.mesh as M
from cfd_proxy_b200.driver import Session
from bench import lattice_for

mp, order, kcfg = float(sys.argv[1]), sys.argv[2], sys.argv[3]
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 30
torder = int(sys.argv[5]) if len(sys.argv) > 5 else 0
ver, chunk, pers, gorder = (int(x) for x in (kcfg.split(".") + ["0"])[:4])
os.environ["CFDP_ORDER"] = str(gorder)
with Session(8, device=0, tile_points=256, tile_order=torder) as S:
    S.load_spec(M.make_spec(lattice_for(mp), (2, 2, 2), order=order, brick=8, jitter=0.1, allow_big=True))
    S.setup()
    assert S.lib.cfdp_set_kernel(ver, chunk, pers) == ver
    ms = S.iterate("comm_free", iters) / iters
    print(f"{kcfg} {order}: {ms:.4f} ms per launch over {iters} launches")
