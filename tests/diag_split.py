#!/usr/bin/env python
"""Diagnostic that located the split-roles bug (a stray 8-byte cp.async of the non-gathering lanes, profiles/README.md):
which rows come out wrong (tile, position in the tile, slot in the chunk), how many of their 21 columns, and whether a constant
var hides the damage (DIAG_CONST_VAR=1).  Kept as a template for localising parity damage; prints zeros on a healthy build."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # test infrastructure: uses the oracle as the checker
import numpy as np
import cfd_proxy_b200.mesh as M
from cfd_proxy_b200.driver import Session
from oracle import oracle as O

os.environ["CFDP_SPLIT_ROLES"] = "1"
chunk = int(os.environ.setdefault("CFDP_CHUNK", "3"))
spec = M.make_spec((24, 20, 16), (2, 2, 2), order="lex", hexfrac=0.25)
doms = [M.gen_domain(spec, r) for r in range(8)]
recv, send = O.recvsend_index(doms)
CONST = os.environ.get("DIAG_CONST_VAR") == "1"
def var_of(d):
    v = M.var_for(d)
    if CONST: v[:] = 1.25
    return v
want = [O.gradients(d, var_of(d), is_send=O.is_send_mask(d, send[a]), order=1) for a, d in enumerate(doms)]
for attempt in range(6):
    with Session(8, device=0, tile_points=256) as S:
        S.load_spec(spec)
        if CONST:
            for d in S.domains: d.var[:] = 1.25
        S.setup()
        for d in S.domains: d.grad[:] = np.nan
        S.iterate("mpi_bulk_sync", 1); S.download_grad()
        st = S.stats()
        bad_total = 0
        hist_pos, hist_slot, hist_kind, hist_ncol, hist_val = [collections.Counter() for _ in range(5)]
        gt_b = gt_i = 0
        # global tile order: boundary tiles of all domains, then interior tiles
        offs_b, offs_i = [], []
        nb_tot = sum(S.schedule(d)["nboundary"] for d in S.domains)
        cb, ci = 0, nb_tot
        for d in S.domains:
            sc = S.schedule(d); offs_b.append(cb); offs_i.append(ci); cb += sc["nboundary"]; ci += sc["ntiles"] - sc["nboundary"]
        for a, d in enumerate(S.domains):
            nown = doms[a]["nown"]
            wrong = np.nonzero((d.grad[:nown].view(np.uint64) != want[a][:nown].view(np.uint64)).any(axis=(1, 2)))[0]
            if len(wrong) == 0: continue
            bad_total += len(wrong)
            sc = S.schedule(d); rows = sc["row_of_point"][wrong]
            tiles = np.searchsorted(sc["tile_row0"], rows, side="right") - 1
            inv = np.full(sc["nrows"] + 1, -1, np.int64); inv[sc["row_of_point"][:nown]] = np.arange(nown)
            for p, r, t in zip(wrong, rows, tiles):
                pos = r - sc["tile_row0"][t]
                ncol = int((d.grad[p].view(np.uint64) != want[a][p].view(np.uint64)).sum())
                hist_ncol[ncol] += 1
                if np.isnan(d.grad[p]).any(): hist_val["nan"] += 1
                elif t > 0 and inv[sc["tile_row0"][t - 1] + pos] >= 0 and (d.grad[p].view(np.uint64) == want[a][inv[sc["tile_row0"][t - 1] + pos]].view(np.uint64)).all(): hist_val["prev tile row"] += 1
                else:
                    rel = np.abs(d.grad[p] - want[a][p]).max() / (np.abs(want[a][p]).max() + 1e-300)
                    hist_val["rel>1e-3" if rel > 1e-3 else "rel small"] += 1
                g = offs_b[a] + t if t < sc["nboundary"] else offs_i[a] + (t - sc["nboundary"])
                hist_pos[pos // 32] += 1; hist_slot[g % chunk] += 1; hist_kind["boundary" if t < sc["nboundary"] else "interior"] += 1
        print(f"attempt {attempt}: wrong points {bad_total}; by warp {dict(sorted(hist_pos.items()))}; by slot in chunk {dict(sorted(hist_slot.items()))}; {dict(hist_kind)}; tiles {st.ntiles} boundary {st.nboundary_tiles}; wrong columns per point {dict(sorted(hist_ncol.items()))}; values {dict(hist_val)}", flush=True)
