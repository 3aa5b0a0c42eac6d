"""CPU replay of the GPU face schedule: the tile blobs the kernels consume (gradient blob and pseudo-flux blob, raw
bytes through cfdp_get_tile_blob) are decoded and walked in numpy exactly the way gg_tile_pipe_kernel / psd_flux_pipe_kernel
walk them -- same entry order, IEEE multiply and add -- and the result must be BIT-IDENTICAL to the oracle (itself
bit-identical to the unmodified reference with one thread).  This pins the whole host side of the path (tiling, ELL
adjacency in the reference's face order, sign / ghost flags, slot and halo placement, device row numbering) without a GPU.
The replay lives in the tests: the product has no CPU path."""
import numpy as np
import pytest

import cfd_proxy_b200.mesh as M
from oracle import oracle as O
from helpers import bits_differ

PAD = 0xFFFFFFFF
CASES = [((20, 16, 12), (2, 2, 1), "shuffle", 0.4, 64, 0), ((16, 14, 12), (1, 1, 1), "lex", 0.0, 256, 0),
         ((16, 16, 16), (2, 1, 1), "brick", 0.0, 128, 1), ((9, 7, 5), (2, 1, 1), "lex", 1.0, 16, 0)]


def halo_base(npts):
    return (npts + 1) & ~1


def replay_setup(session_factory, n, p, order, hexfrac, tile, torder):
    spec = M.make_spec(n, p, order=order, brick=4, hexfrac=hexfrac)
    nd = p[0] * p[1] * p[2]
    doms = [M.gen_domain(spec, r) for r in range(nd)]
    recv, send = O.recvsend_index(doms) if nd > 1 else ([{}], [{}])
    S = session_factory(nd, tile_points=tile, tile_order=torder)
    S.load_spec(spec)
    S.setup(device=False)                                   # cfdp_plan only: no GPU involved
    return S, doms, recv, send


def device_rows(S, doms, per_point):
    """[nrows_total, ...] array in device row order from per-domain host arrays (what cfdp_var_to_device builds)."""
    total = int(S.stats().rows)
    out = np.zeros((total,) + per_point[0].shape[1:])
    base = 0
    for a, d in enumerate(S.domains):
        sc = S.schedule(d)
        out[base + sc["row_of_point"]] = per_point[a]
        base += sc["nrows"]
    return out


@pytest.mark.parametrize("n,p,order,hexfrac,tile,torder", CASES)
def test_gradient_blob_replay_bit_identical(session_factory, n, p, order, hexfrac, tile, torder):
    S, doms, recv, send = replay_setup(session_factory, n, p, order, hexfrac, tile, torder)
    hvar = device_rows(S, doms, [0.5 * M.var_for(d) for d in doms])        # the device keeps 0.5*var (exact scaling)
    pvol = device_rows(S, doms, [d["pvolume"][:, None] for d in doms])[:, 0]
    for a, d in enumerate(S.domains):
        dom = doms[a]
        want = O.gradients(dom, M.var_for(dom), is_send=O.is_send_mask(dom, send[a]), order=1)
        sc = S.schedule(d)
        for t in range(sc["ntiles"]):
            b = S.tile_blob(d, t)
            npts, nb = b["npts"], halo_base(b["npts"])
            assert b["row0"] % 16 == 0 and b["npad"] % 32 == 0 and b["npad"] >= npts
            # the slot the production kernel maps padding entries to: used by no entry, zero normal (TileDesc::zslot)
            used = (b["ell"][b["ell"] != PAD] >> 16) & 0x7FFF
            assert b["zslot"] < b["nfaces"] and b["zslot"] not in set(used.tolist()) and not b["normals"][b["zslot"]].any()
            for i in range(npts):
                acc = np.zeros((7, 3))
                hv = hvar[b["row0"] + i]
                for j in range(b["maxdeg"]):
                    e = int(b["ell"][j, i])
                    if e == PAD:
                        continue
                    loc, slot, is_p1 = e & 0x7FFF, (e >> 16) & 0x7FFF, e >> 31
                    row = b["row0"] + loc if loc < npts else int(b["halo_rows"][loc - nb])
                    assert row != PAD
                    nrm = -b["normals"][slot] if is_p1 else b["normals"][slot]
                    val = hv + hvar[row]                                 # == 0.5*(var[p0]+var[p1]), gradients.c:77
                    acc = acc + val[:, None] * nrm[None, :]              # one rounding per multiply, one per add
                acc = acc * (1.0 / pvol[b["row0"] + i])
                owner = S.row_owner(b["row0"] + i)
                assert owner is not None and owner[0] == d.rank
                assert bits_differ(acc, want[owner[1]]) == 0
            assert (b["ell"][:, npts:] == PAD).all()                       # padding lanes hold no entries


def test_replay_with_refined_placement(session_factory, monkeypatch):
    """CFDP_PLACE_REFINE (hill climbing on top of the greedy bank placement) permutes rows, slots and halo positions:
    the results must not change, the estimated wavefronts must not grow."""
    n, p, order, hexfrac, tile, torder = CASES[0]
    S, doms, recv, send = replay_setup(session_factory, n, p, order, hexfrac, tile, torder)
    base = S.stats().lds_wavefronts_est
    monkeypatch.setenv("CFDP_PLACE_REFINE", "3")
    test_gradient_blob_replay_bit_identical(session_factory, n, p, order, hexfrac, tile, torder)
    test_flux_blob_replay_bit_identical(session_factory, n, p, order, hexfrac, tile, torder)
    S2, _, _, _ = replay_setup(session_factory, n, p, order, hexfrac, tile, torder)
    assert S2.stats().lds_wavefronts_est <= base


@pytest.mark.parametrize("n,p,order,hexfrac,tile,torder", CASES)
def test_flux_blob_replay_bit_identical(session_factory, n, p, order, hexfrac, tile, torder):
    S, doms, recv, send = replay_setup(session_factory, n, p, order, hexfrac, tile, torder)
    nd = len(doms)
    grads = [O.gradients(dom, M.var_for(dom), is_send=O.is_send_mask(dom, send[a]), order=1) for a, dom in enumerate(doms)]
    grads = O.exchange(grads, recv, send) if nd > 1 else grads
    g9 = device_rows(S, doms, [np.nan_to_num(g).reshape(-1, 21)[:, :9] for g in grads])   # grad[p][IVX..IVZ][0..2]
    lam = -2.0 / 3.0
    total_entries = total_grad_entries = 0
    for a, d in enumerate(S.domains):
        dom = doms[a]
        want = O.psd_flux(dom, np.nan_to_num(grads[a]), is_send=O.is_send_mask(dom, send[a]), order=1)
        sc = S.schedule(d)
        for t in range(sc["ntiles"]):
            b, gb = S.tile_blob(d, t, flux=True), S.tile_blob(d, t)
            npts, nb = b["npts"], halo_base(b["npts"])
            assert b["row0"] == gb["row0"] and npts == gb["npts"] and b["npad"] == gb["npad"]
            assert b["maxdeg"] <= gb["maxdeg"] and b["nfaces"] <= gb["nfaces"] and b["nhalo"] <= gb["nhalo"]
            total_entries += int((b["ell"] != PAD).sum()); total_grad_entries += int((gb["ell"] != PAD).sum())
            for i in range(npts):
                acc = np.zeros(3)
                ga = g9[b["row0"] + i]
                for j in range(b["maxdeg"]):
                    e = int(b["ell"][j, i])
                    if e == PAD:
                        continue
                    loc, ghost, slot, is_p1 = e & 0x7FFF, (e >> 15) & 1, (e >> 16) & 0x7FFF, e >> 31
                    assert is_p1 or ghost                                 # only contributing entries are stored
                    row = b["row0"] + loc if loc < npts else int(b["halo_rows"][loc - nb])
                    owner = S.row_owner(row)
                    assert owner is not None and owner[0] == d.rank and (owner[1] >= dom["nown"]) == bool(ghost)
                    dd = 0.5 * (ga + g9[row])
                    nx, ny, nz = b["normals"][slot]
                    sxx = lam * (dd[4] + dd[8] - 2.0 * dd[0]); syy = lam * (dd[0] + dd[8] - 2.0 * dd[4]); szz = lam * (dd[0] + dd[4] - 2.0 * dd[8])
                    sxy = dd[1] + dd[3]; sxz = dd[2] + dd[6]; syz = dd[5] + dd[7]
                    f = -np.array([sxx * nx + sxy * ny + sxz * nz, sxy * nx + syy * ny + syz * nz, sxz * nx + syz * ny + szz * nz])
                    acc = acc - f if is_p1 else acc + f
                owner = S.row_owner(b["row0"] + i)
                assert bits_differ(acc, want[owner[1]]) == 0
    # about half of the gradient adjacency contributes to the pseudo flux
    assert 0.3 * total_grad_entries < total_entries < 0.75 * total_grad_entries
    assert S.stats().flux_blob_bytes > 0
