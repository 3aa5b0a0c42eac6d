"""Worker of tests/test_multiproc.py: one torchrun rank = one (virtual) GPU process hosting a block of
mesh domains.  CPU only: gloo carries the setup handshake, cfdp_plan() builds the exchange plan."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch.distributed as dist  # noqa: E402

import cfd_proxy_b200.mesh as M  # noqa: E402
from cfd_proxy_b200.driver import session_from_env  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    nd = int(sys.argv[1])
    grid = {8: (2, 2, 2), 12: (3, 2, 2), 4: (2, 2, 1)}[nd]
    spec = M.make_spec((18, 14, 10), grid, order="shuffle", hexfrac=0.3)
    doms = [M.gen_domain(spec, r) for r in range(nd)]
    recv, send = O.recvsend_index(doms)
    S = session_from_env(nd, backend="gloo")
    rank, world = S.proc_rank, S.nprocs
    S.load_spec(spec)
    S.setup(device=False)
    errors = []
    # 1. halo lists of the hosted domains == oracle (bit exact), whether the partner is local or remote
    for d in S.domains:
        si, ri = d.index_lists()
        a = d.rank
        if set(si) != set(send[a]) or set(ri) != set(recv[a]):
            errors.append(f"partner sets differ for domain {a}")
        for k in si:
            if not np.array_equal(si[k], send[a][k]):
                errors.append(f"sendindex[{a}][{k}] differs")
            if not np.array_equal(S.pack_list(d, k), send[a][k]):
                errors.append(f"device pack list [{a}][{k}] differs")
        for k in ri:
            if not np.array_equal(ri[k], recv[a][k]):
                errors.append(f"recvindex[{a}][{k}] differs")
            if not np.array_equal(S.unpack_list(d, k), recv[a][k]):
                errors.append(f"device unpack list [{a}][{k}] differs")
        cd = d.cd
        for i in range(cd.ncommdomains):  # offset tables of comm_data.c:343-352 and the exchanged remote offsets
            k = cd.commpartner[i]
            dk = doms[k]
            off = 0
            for kk in dk["commpartner"]:
                if kk == a:
                    break
                off += int(dk["recvcount"][kk]) * 21 * 8
            if cd.remote_recv_offset[k] != off:
                errors.append(f"remote_recv_offset[{a}][{k}] = {cd.remote_recv_offset[k]} != {off}")
            if cd.notification[k] != list(dk["commpartner"]).index(a):
                errors.append(f"notification[{a}][{k}] wrong")
    # 2. the per-peer packed buffers: what rank p sends to q must be, entry by entry, what q expects from p
    lib = S.lib
    plan = {}
    npeers = lib.cfdp_get_peer_plan(-1, None, None, None)
    soff = roff = 0
    for i in range(npeers):
        proc, ns, nr = C.c_int(), C.c_longlong(), C.c_longlong()
        lib.cfdp_get_peer_plan(i, C.byref(proc), C.byref(ns), C.byref(nr))
        sg, rg = [], []
        for j in range(ns.value):
            dom, pt = C.c_int(), C.c_int()
            assert lib.cfdp_get_exchange_entry(0, soff + j, C.byref(dom), C.byref(pt)) == 0
            sg.append(int(doms[dom.value]["global_id"][pt.value]))
        for j in range(nr.value):
            dom, pt = C.c_int(), C.c_int()
            assert lib.cfdp_get_exchange_entry(1, roff + j, C.byref(dom), C.byref(pt)) == 0
            if pt.value < doms[dom.value]["nown"]:
                errors.append("recv entry is not a ghost point")
            rg.append(int(doms[dom.value]["global_id"][pt.value]))
        soff += ns.value
        roff += nr.value
        plan[proc.value] = (sg, rg)
    allplans = [None] * world
    dist.all_gather_object(allplans, plan)
    for q, (sg, rg) in plan.items():
        their_s, their_r = allplans[q][rank]
        if sg != their_r:
            errors.append(f"rank {rank} -> {q}: send order does not match the peer's receive order")
        if rg != their_s:
            errors.append(f"rank {q} -> {rank}: receive order does not match the peer's send order")
    # 3. fused pack: every slot of the packed send buffer is filled exactly once, by the tile that owns the source point
    st0 = S.stats()
    filled = {}
    for t in range(st0.nboundary_tiles):
        ne = lib.cfdp_get_tile_exports(t, 0, None, None, None)
        src = (C.c_uint * max(ne, 1))(); dst = (C.c_uint * max(ne, 1))(); kind = (C.c_int * max(ne, 1))()
        lib.cfdp_get_tile_exports(t, ne, src, dst, kind)
        for i in range(ne):
            if kind[i] != 1:
                continue
            sd, sp, ed, ep = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            if lib.cfdp_get_row_owner(src[i], C.byref(sd), C.byref(sp)) != 0 or \
               lib.cfdp_get_exchange_entry(0, dst[i], C.byref(ed), C.byref(ep)) != 0:
                errors.append("bad export entry")
                continue
            if dst[i] in filled:
                errors.append(f"send slot {dst[i]} filled twice")
            filled[dst[i]] = 1
            if (sd.value, sp.value) != (ed.value, ep.value):
                errors.append(f"send slot {dst[i]} filled from the wrong point")
    if len(filled) != st0.send_rows_remote:
        errors.append(f"{len(filled)} send slots filled, {st0.send_rows_remote} expected")
    st = S.stats()
    total_remote = sum(len(v[0]) for v in plan.values())
    if total_remote != st.send_rows_remote:
        errors.append("send_rows_remote mismatch")
    out = dict(rank=rank, errors=errors, local=int(st.send_rows_local), remote=int(st.send_rows_remote), peers=sorted(plan))
    with open(os.path.join(os.environ["CFDP_MP_OUT"], f"rank{rank}.json"), "w") as f:
        json.dump(out, f)
    S.close()
    dist.barrier()
    dist.destroy_process_group()
    return 1 if errors else 0


if __name__ == "__main__":
    sys.exit(main())
