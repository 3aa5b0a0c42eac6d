"""Worker of tests/test_multigpu.py: one torchrun rank per GPU (NCCL).  Every rank checks its hosted
domains -- own rows AND ghost rows filled over NVLink -- bit for bit against the oracle, and the pseudo flux
(flux.c) computed from those exchanged rows."""
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
    grid = {8: (2, 2, 2), 12: (3, 2, 2), 24: (4, 3, 2)}[nd]
    spec = M.make_spec((36, 30, 24), grid, order="lex", hexfrac=0.3)
    doms = [M.gen_domain(spec, r) for r in range(nd)]
    recv, send = O.recvsend_index(doms)
    want = [O.gradients(d, M.var_for(d), is_send=O.is_send_mask(d, send[a]), order=1) for a, d in enumerate(doms)]
    want = O.exchange(want, recv, send)
    want_flux = [O.psd_flux(d, want[a], is_send=O.is_send_mask(d, send[a]), order=1) for a, d in enumerate(doms)]
    S = session_from_env(nd, backend="nccl")
    S.load_spec(spec)
    S.setup()
    errors = []
    transports = {}
    for variant in ("mpi_bulk_sync", "mpi_early_recv", "mpi_async", "gaspi_bulk_sync", "gaspi_async"):  # gaspi_* = CUDA-IPC put + notify
        for d in S.domains:
            d.grad[:] = np.nan
            d.psd_flux[:] = np.nan
        S.lib.cfdp_set_resident(1)
        S.set_flux(True)                 # every iteration: gradient + exchange, then the pseudo flux on the exchanged rows
        S.iterate(variant, 3)
        transports[variant] = int(S.stats().transport)
        S.download_grad()
        S.download_flux()
        for d in S.domains:
            bad = int((d.grad.view(np.uint64) != want[d.rank].view(np.uint64)).sum())
            if bad:
                errors.append(f"{variant}: domain {d.rank}: {bad} words differ")
            nown = doms[d.rank]["nown"]
            bad = int((d.psd_flux[:nown].view(np.uint64) != want_flux[d.rank][:nown].view(np.uint64)).sum())
            if bad:
                errors.append(f"{variant}: domain {d.rank}: {bad} pseudo-flux words differ")
    st = S.stats()
    want_t = dict(mpi_bulk_sync=2, mpi_early_recv=2, mpi_async=2, gaspi_bulk_sync=3 if st.ipc_ready else 2, gaspi_async=4 if st.direct_ready else (3 if st.ipc_ready else 2))
    if st.send_rows_remote and transports != want_t:
        errors.append(f"transports {transports} != {want_t}")
    out = dict(rank=S.proc_rank, errors=errors, transports=transports, local=int(st.send_rows_local), remote=int(st.send_rows_remote))
    with open(os.path.join(os.environ["CFDP_MP_OUT"], f"rank{S.proc_rank}.json"), "w") as f:
        json.dump(out, f)
    S.close()
    dist.barrier()
    dist.destroy_process_group()
    return 1 if errors else 0


if __name__ == "__main__":
    sys.exit(main())
