import os
import numpy as np
import cfd_proxy_b200.mesh as M

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = ["f6like12_lvl4", "tet8_shuffle", "hex2", "single"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n, p = tuple(int(x) for x in z["n"]), tuple(int(x) for x in z["p"])
    spec = M.make_spec(n, p, order=str(z["order"]), brick=4, hexfrac=float(z["hexfrac"]))
    nd = p[0] * p[1] * p[2]
    doms = [M.gen_domain(spec, r) for r in range(nd)]
    return z, spec, doms, int(z["lvl"])


def golden_grad(z, variant, threads, d):
    return z[f"grad_{variant}_t{threads}_d{d}"].view(np.float64)


def golden_flux(z, variant, d):
    return z[f"flux_{variant}_t1_d{d}"].view(np.float64)


def golden_index(z, d, nd):
    send = {k: z[f"sendindex_d{d}_k{k}"] for k in range(nd) if f"sendindex_d{d}_k{k}" in z}
    recv = {k: z[f"recvindex_d{d}_k{k}"] for k in range(nd) if f"recvindex_d{d}_k{k}" in z}
    return send, recv


def bits_differ(a, b):
    return int((np.ascontiguousarray(a).view(np.uint64) != np.ascontiguousarray(b).view(np.uint64)).sum())
