#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ by running the UNMODIFIED reference
(oracle/_ref/ref_harness = /root/reference/src compiled against the shims of oracle/shim, one rank per
domain over the shared-memory MPI shim) on small F6-schema stand-in meshes.  Run in the build container
(the GPU box has no /root/reference):   python tests/golden/make_golden.py

Each fixture <name>.npz holds, per domain d: the mesh spec (regenerated deterministically by
cfd_proxy_b200.mesh), grad_<variant>_t<threads>_d<d> (float64 bit patterns stored as uint64), sendindex /
recvindex as flat arrays, flux_<variant>_t1_d<d> = psd_flux after gradient + exchange + compute_psd_flux
(src/flux.c; one thread, the only thread count for which the reference's result is a function of the mesh), so that
tests can check
  * the oracle restatement (oracle/gg_oracle.c) against the real reference  (pins the oracle),
  * the CUDA path against the real reference's numbers on the GPU box.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cfd_proxy_b200.mesh as M  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = {
    # name: (lattice, domain grid, order, hexfrac, lvl)
    "f6like12_lvl4": ((16, 12, 10), (3, 2, 2), "lex", 0.4, 4),
    "tet8_shuffle": ((10, 8, 8), (2, 2, 2), "shuffle", 0.0, 1),
    "hex2": ((9, 7, 5), (2, 1, 1), "lex", 1.0, 1),
    "single": ((12, 10, 8), (1, 1, 1), "lex", 0.25, 1),
}


def main():
    assert O.have_ref(), "build oracle/_ref first: make ref"
    here = os.path.dirname(os.path.abspath(__file__))
    for name, (n, p, order, hexfrac, lvl) in CASES.items():
        spec = M.make_spec(n, p, order=order, brick=4, hexfrac=hexfrac)
        nd = p[0] * p[1] * p[2]
        tmp = tempfile.mkdtemp(prefix="golden_")
        prefix = os.path.join(tmp, "dualgrid")
        M.write_mesh(prefix, spec, lvl=lvl)
        out = dict(n=np.array(n), p=np.array(p), order=np.array(order), hexfrac=np.array(hexfrac), lvl=np.array(lvl))
        variants = ["comm_free"] + (["mpi_bulk_sync", "mpi_async"] if nd > 1 else [])
        for thr in (1, 3):
            for v in (variants if thr == 1 else variants[-1:]):
                res = O.run_ref(prefix, lvl, nd, v, 2, os.path.join(tmp, f"o_{v}_{thr}"), threads=thr)
                for d, r in enumerate(res):
                    out[f"grad_{v}_t{thr}_d{d}"] = r["grad"].view(np.uint64)
                    if thr == 1 and v == variants[-1]:
                        for k, idx in r["sendindex"].items():
                            out[f"sendindex_d{d}_k{k}"] = idx
                        for k, idx in r["recvindex"].items():
                            out[f"recvindex_d{d}_k{k}"] = idx
        res = O.run_ref(prefix, lvl, nd, variants[-1], 2, os.path.join(tmp, "o_flux"), threads=1, with_flux=True)
        for d, r in enumerate(res):
            assert np.array_equal(r["grad"].view(np.uint64), out[f"grad_{variants[-1]}_t1_d{d}"])
            out[f"flux_{variants[-1]}_t1_d{d}"] = r["psd_flux"].view(np.uint64)
        np.savez_compressed(os.path.join(here, name + ".npz"), **out)
        print(name, "->", os.path.getsize(os.path.join(here, name + ".npz")), "bytes")


if __name__ == "__main__":
    main()
