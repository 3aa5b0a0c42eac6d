"""Synthetic F6-schema mesh domains (wrapper over csrc/mesh_gen.c) and their NetCDF files."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import lib as L
from . import netcdf3

ORDER = {"lex": 0, "brick": 1, "shuffle": 2}
DEFAULT_SEED = 0x5DEECE66D


def make_spec(n, p, order="lex", brick=8, hexfrac=0.0, jitter=0.1, seed=DEFAULT_SEED, allow_big=False) -> L.MeshSpec:
    """n=(nx,ny,nz) lattice, p=(px,py,pz) domain grid, hexfrac = fraction of x that is hex-only."""
    s = L.MeshSpec()
    s.nx, s.ny, s.nz = map(int, n)
    s.px, s.py, s.pz = map(int, p)
    s.order = ORDER[order] if isinstance(order, str) else int(order)
    s.brick = int(brick)
    s.hexcut = int(round(hexfrac * s.nx))
    s.allow_big = int(allow_big)
    s.jitter = float(jitter)
    s.seed = int(seed)
    return s


def f6like_spec(ndomains=12, lvl=1, scale=1.0, **kw) -> L.MeshSpec:
    """F6-like stand-in: ~2 M points at level 1 (CFD-Proxy.pdf p.3), 8x coarser per level,
    hybrid (40 % of the x range hexahedral-dual, the rest tetrahedral-dual), 12 or 24 domains."""
    grids = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2), 12: (3, 2, 2), 24: (4, 3, 2), 48: (4, 4, 3)}
    base = np.array([160, 120, 108], dtype=float) * scale  # 2.07 M points
    n = np.maximum(4, np.round(base / (2 ** (lvl - 1)))).astype(int)
    kw.setdefault("hexfrac", 0.4)
    kw.setdefault("order", "lex")
    return make_spec(n, grids[ndomains], **kw)


def _np(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def gen_domain(spec: L.MeshSpec, rank: int) -> dict:
    """Generate one domain; returns numpy copies of all arrays."""
    lib = L.load()
    m = L.MeshDomain()
    rc = lib.cfdp_mesh_gen_domain(C.byref(spec), rank, C.byref(m))
    if rc != 0:
        raise RuntimeError(f"cfdp_mesh_gen_domain failed rc={rc}")
    try:
        d = dict(nfaces=m.nfaces, nown=m.nown, nall=m.nall, nadd=m.nadd, ndomains=m.ndomains,
                 ncommdomains=m.ncommdomains, rank=rank)
        d["fpoint"] = _np(m.fpoint, 2 * m.nfaces, np.int32).reshape(-1, 2)
        d["fnormal"] = _np(m.fnormal, 3 * m.nfaces, np.float64).reshape(-1, 3)
        d["pvolume"] = _np(m.pvolume, m.nall, np.float64)
        d["commpartner"] = _np(m.commpartner, m.ncommdomains, np.int32)
        d["sendcount"] = _np(m.sendcount, m.ndomains, np.int32)
        d["recvcount"] = _np(m.recvcount, m.ndomains, np.int32)
        d["addpoint_owner"] = _np(m.addpoint_owner, m.nadd, np.int32)
        d["addpoint_idx"] = _np(m.addpoint_idx, m.nadd, np.int32)
        d["global_id"] = _np(m.global_id, m.nall, np.int64)
    finally:
        lib.cfdp_mesh_free_domain(C.byref(m))
    return d


def var_for(dom: dict, seed=DEFAULT_SEED) -> np.ndarray:
    """var[p][eq] = u01(hash(seed, global id, eq)) + eq -- ghosts agree with their owners."""
    gid = dom["global_id"].astype(np.uint64)
    eq = np.arange(7, dtype=np.uint64)
    key = np.uint64(seed) ^ (gid[:, None] * np.uint64(8) + eq[None, :] + np.uint64(0x5DEECE66D))
    with np.errstate(over="ignore"):
        z = key + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) + eq[None, :].astype(np.float64)


def domain_path(prefix: str, rank: int, lvl: int) -> str:
    return f"{prefix}_domain_{rank}_lvl_{lvl}"  # hybrid.f6.c:57-62


def write_mesh(prefix: str, spec: L.MeshSpec, lvl: int = 1, version: int = 2, with_var: bool = True, seed=DEFAULT_SEED):
    """Write all domain files (+ raw little-endian `.var` side files for the oracle harness)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    nd = spec.px * spec.py * spec.pz
    doms = []
    for r in range(nd):
        d = gen_domain(spec, r)
        netcdf3.write_domain_file(domain_path(prefix, r, lvl), d, version=version)
        if with_var:
            var_for(d, seed).astype("<f8").tofile(domain_path(prefix, r, lvl) + ".var")
        doms.append(d)
    return doms
