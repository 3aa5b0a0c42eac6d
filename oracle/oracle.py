"""oracle/oracle.py -- TEST INFRASTRUCTURE, not product code.

Python face of the oracle: ctypes wrapper over gg_oracle.c (the CPU restatement of the
gradient loop), numpy restatements of the halo list construction and exchange, and a runner
for the unmodified reference built into oracle/_ref (when present).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        ip, dp, up = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_ubyte)
        _lib.oracle_gradients.restype = C.c_long
        _lib.oracle_gradients.argtypes = [C.c_int, C.c_int, C.c_int, ip, dp, dp, dp, dp, up, C.c_int]
        _lib.oracle_psd_flux.restype = C.c_long
        _lib.oracle_psd_flux.argtypes = [C.c_int, C.c_int, C.c_int, ip, dp, dp, dp, up, C.c_int]
        _lib.oracle_error_scale.restype = None
        _lib.oracle_error_scale.argtypes = [C.c_int, C.c_int, ip, dp, dp, dp, dp]
        _lib.oracle_pack.argtypes = [dp, C.c_int, ip, C.c_int, dp]
        _lib.oracle_unpack.argtypes = [dp, C.c_int, ip, C.c_int, dp]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def gradients(dom, var, grad_in=None, is_send=None, order=1):
    """Green-Gauss gradients of one domain (gradients.c:25-147 semantics).  Returns grad[nall,7,3];
    rows no face writes (ghosts) keep grad_in (NaN by default)."""
    nall, nown = int(dom["nall"]), int(dom["nown"])
    fp = np.ascontiguousarray(dom["fpoint"], dtype=np.int32)
    fn = np.ascontiguousarray(dom["fnormal"], dtype=np.float64)
    pv = np.ascontiguousarray(dom["pvolume"], dtype=np.float64)
    var = np.ascontiguousarray(var, dtype=np.float64)
    grad = np.full((nall, 7, 3), np.nan) if grad_in is None else np.array(grad_in, dtype=np.float64).reshape(nall, 7, 3).copy()
    send = None
    if is_send is not None:
        send = np.ascontiguousarray(is_send, dtype=np.uint8)
    lib().oracle_gradients(len(fp), nown, nall, _p(fp, C.c_int), _p(fn, C.c_double), _p(pv, C.c_double),
                           _p(var, C.c_double), _p(grad, C.c_double),
                           _p(send, C.c_ubyte) if send is not None else None, order)
    return grad


def psd_flux(dom, grad, flux_in=None, is_send=None, order=1):
    """Pseudo flux of one domain from its (exchanged) grad[nall,7,3] (flux.c:111-201, single-thread semantics).
    Returns psd_flux[nall,3]; only own rows are defined, the others keep flux_in (NaN by default)."""
    nall, nown = int(dom["nall"]), int(dom["nown"])
    fp = np.ascontiguousarray(dom["fpoint"], dtype=np.int32)
    fn = np.ascontiguousarray(dom["fnormal"], dtype=np.float64)
    grad = np.ascontiguousarray(grad, dtype=np.float64).reshape(nall, 21)
    out = np.full((nall, 3), np.nan) if flux_in is None else np.array(flux_in, dtype=np.float64).reshape(nall, 3).copy()
    send = np.ascontiguousarray(is_send, dtype=np.uint8) if is_send is not None else None
    lib().oracle_psd_flux(len(fp), nown, nall, _p(fp, C.c_int), _p(fn, C.c_double), _p(grad, C.c_double), _p(out, C.c_double),
                          _p(send, C.c_ubyte) if send is not None else None, order)
    return out


def psd_flux_numpy(dom, grad):
    """Independent vectorised restatement of flux.c:111-201 for a single-thread run (np.add.at, file order):
    returns (psd_flux[nown,3], error scale S_p[nown]) -- cross-check of oracle_psd_flux and the tolerance of the
    fused-multiply-add mode: |a-b| <= 1e-12*|b| + 64*eps*S_p, S_p = sum over contributing faces of 4*|n_f|_1*max|d_f|."""
    nown = int(dom["nown"])
    fp, fn = dom["fpoint"], dom["fnormal"]
    keep = (fp[:, 0] < nown) | (fp[:, 1] < nown)
    fp, fn = fp[keep], fn[keep]
    g = np.asarray(grad, dtype=np.float64).reshape(-1, 21)[:, :9]
    d = 0.5 * (g[fp[:, 0]] + g[fp[:, 1]])                        # [F,9]: dvx_dx,dy,dz, dvy_..., dvz_...
    lam = -2.0 / 3.0
    sxx = lam * (d[:, 4] + d[:, 8] - 2.0 * d[:, 0]); syy = lam * (d[:, 0] + d[:, 8] - 2.0 * d[:, 4]); szz = lam * (d[:, 0] + d[:, 4] - 2.0 * d[:, 8])
    sxy = d[:, 1] + d[:, 3]; sxz = d[:, 2] + d[:, 6]; syz = d[:, 5] + d[:, 7]
    nx, ny, nz = fn[:, 0], fn[:, 1], fn[:, 2]
    fl = -np.stack([sxx * nx + sxy * ny + sxz * nz, sxy * nx + syy * ny + syz * nz, sxz * nx + syz * ny + szz * nz], axis=1)
    ftype = np.where(fp[:, 0] >= nown, 1, np.where(fp[:, 1] >= nown, 2, 3))
    out = np.zeros((nown, 3)); scale = np.zeros(nown)
    mag = 4.0 * np.abs(fn).sum(axis=1) * np.abs(d).max(axis=1)
    w0 = (ftype != 3) & (fp[:, 0] < nown)
    w1 = (ftype != 2) & (fp[:, 1] < nown)
    np.add.at(out, fp[w0, 0], fl[w0]); np.add.at(scale, fp[w0, 0], mag[w0])
    np.subtract.at(out, fp[w1, 1], fl[w1]); np.add.at(scale, fp[w1, 1], mag[w1])
    return out, scale


def error_scale(dom, var):
    nown = int(dom["nown"])
    fp = np.ascontiguousarray(dom["fpoint"], dtype=np.int32)
    fn = np.ascontiguousarray(dom["fnormal"], dtype=np.float64)
    pv = np.ascontiguousarray(dom["pvolume"], dtype=np.float64)
    var = np.ascontiguousarray(var, dtype=np.float64)
    out = np.zeros(nown)
    lib().oracle_error_scale(len(fp), nown, _p(fp, C.c_int), _p(fn, C.c_double), _p(pv, C.c_double),
                             _p(var, C.c_double), _p(out, C.c_double))
    return out


def gradients_numpy(dom, var):
    """Independent vectorised restatement (np.add.at, file order) -- cross-check of gg_oracle.c."""
    nall, nown = int(dom["nall"]), int(dom["nown"])
    fp, fn = dom["fpoint"], dom["fnormal"]
    keep = (fp[:, 0] < nown) | (fp[:, 1] < nown)
    fp, fn = fp[keep], fn[keep]
    val = 0.5 * (var[fp[:, 0]] + var[fp[:, 1]])                # [F,7]
    contrib = val[:, :, None] * fn[:, None, :]                  # [F,7,3]
    g = np.zeros((nall, 7, 3))
    w0, w1 = fp[:, 0] < nown, fp[:, 1] < nown
    np.add.at(g, fp[w0, 0], contrib[w0])
    np.subtract.at(g, fp[w1, 1], contrib[w1])
    g[:nown] *= (1.0 / dom["pvolume"][:nown])[:, None, None]
    g[nown:] = np.nan
    return g


def recvsend_index(doms):
    """Halo lists of all domains (comm_data.c:116-255 without the MPI handshake).

    recvindex[a][k] = nown_a + (positions j of a's addpoints owned by k, ascending)   (:163-174)
    sendindex[a][k] = what k asks of a = k's addpoint_idx over k's recvindex[a] order (:197-222)
    """
    nd = len(doms)
    recv = [dict() for _ in range(nd)]
    send = [dict() for _ in range(nd)]
    for a, d in enumerate(doms):
        for k in d["commpartner"]:
            k = int(k)
            if d["recvcount"][k] > 0:
                recv[a][k] = (int(d["nown"]) + np.nonzero(d["addpoint_owner"] == k)[0]).astype(np.int32)
    for a, d in enumerate(doms):
        for k in d["commpartner"]:
            k = int(k)
            if d["sendcount"][k] > 0:
                dk = doms[k]
                send[a][k] = dk["addpoint_idx"][recv[k][a] - int(dk["nown"])].astype(np.int32)
    return recv, send


def is_send_mask(dom, sendindex_a):
    m = np.zeros(int(dom["nown"]), dtype=np.uint8)
    for idx in sendindex_a.values():
        m[idx] = 1
    return m


def exchange(grads, recv, send):
    """threads.c:791-839: ghost rows <- owner rows, raw copies (all domains at once)."""
    out = [g.copy() for g in grads]
    for a in range(len(grads)):
        for k, ridx in recv[a].items():
            out[a][ridx] = grads[k][send[k][a]]
    return out


# ---------------------------------------------------------------------------------------
# the unmodified reference (oracle/_ref), when it has been built in this tree
# ---------------------------------------------------------------------------------------
def have_ref():
    return all(os.path.exists(os.path.join(REF_DIR, x)) for x in ("ref_harness", "mpirun_shim"))


def run_ref(prefix, lvl, ndomains, variant, niter, outprefix, threads=1, repeats=1, timeout=600, with_flux=False, timing_only=False):
    """Run the reference harness (one rank per domain) and return per-domain grad/index/time (and psd_flux);
    timing_only: no result dumps (large meshes), only the per-rank timing records."""
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), REF_WITH_FLUX="1" if with_flux else "0", REF_NO_DUMP="1" if timing_only else "0")
    cmd = [os.path.join(REF_DIR, "mpirun_shim"), "-np", str(ndomains), os.path.join(REF_DIR, "ref_harness"),
           "-lvl", str(lvl), prefix, variant, str(niter), outprefix, str(repeats)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"reference harness failed rc={r.returncode}\n{r.stdout}\n{r.stderr}")
    res = []
    if timing_only:
        return [dict(time=json.loads(open(f"{outprefix}_domain_{d}.time").read())) for d in range(ndomains)]
    for d in range(ndomains):
        g = np.fromfile(f"{outprefix}_domain_{d}.grad", dtype="<f8").reshape(-1, 7, 3)
        raw = np.fromfile(f"{outprefix}_domain_{d}.index", dtype="<i4")
        n, pos = int(raw[0]), 1
        sidx, ridx = {}, {}
        for _ in range(n):
            k, sc, rc = map(int, raw[pos:pos + 3])
            pos += 3
            sidx[k] = raw[pos:pos + sc].copy(); pos += sc
            ridx[k] = raw[pos:pos + rc].copy(); pos += rc
        t = json.loads(open(f"{outprefix}_domain_{d}.time").read())
        res.append(dict(grad=g, sendindex=sidx, recvindex=ridx, time=t))
        if with_flux:
            res[-1]["psd_flux"] = np.fromfile(f"{outprefix}_domain_{d}.flux", dtype="<f8").reshape(-1, 3)
    return res
