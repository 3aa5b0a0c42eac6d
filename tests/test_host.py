"""Host logic on CPU (no compute calls): C ABI surface, NetCDF loader, comm tables, GPU face schedule."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import cfd_proxy_b200.mesh as M
from cfd_proxy_b200 import lib as L
from cfd_proxy_b200 import netcdf3
from oracle import oracle as O
from helpers import GOLDEN, golden_index, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- C ABI ----------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cfdp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = re.sub(r"typedef[^;{]*\([^;]*;", "", hdr)          # function-pointer typedefs are not symbols
    declared = set(re.findall(r"\b([a-z_][a-z0-9_]*)\s*\([^;{]*\)\s*;", hdr))
    declared -= {"aligned", "__attribute__"}
    declared = {d for d in declared if not d.endswith("_fn")}
    lib = L.load()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(L.EXPORTED), declared ^ set(L.EXPORTED)


def test_struct_layout_matches_reference_headers():
    """sizeof/offsetof of the ctypes mirrors == the C structs of include/cfdp_b200.h (compiled here with gcc);
    field order and types follow solver_data.h:66-81 / comm_data.h:15-55."""
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "cfdp_b200.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu\n", sizeof(solver_data), offsetof(solver_data, fpoint), offsetof(solver_data, grad), offsetof(solver_data, fcolor), offsetof(solver_data, niter));
      printf("%zu %zu %zu %zu %zu %zu\n", sizeof(comm_data), offsetof(comm_data, addpoint_owner), offsetof(comm_data, recvindex), offsetof(comm_data, nreq), offsetof(comm_data, remote_recv_offset), offsetof(comm_data, recv_stage));
      printf("%zu %zu\n", sizeof(RangeList), sizeof(counter_t));
      return 0; }'''
    exe = "/tmp/cfdp_layout_test"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src, text=True, check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    got = list(map(int, out))
    SD, CD = L.SolverData, L.CommData
    want = [C.sizeof(SD), SD.fpoint.offset, SD.grad.offset, SD.fcolor.offset, SD.niter.offset,
            C.sizeof(CD), CD.addpoint_owner.offset, CD.recvindex.offset, CD.nreq.offset, CD.remote_recv_offset.offset,
            CD.recv_stage.offset, C.sizeof(L.RangeList), 64]
    assert got == want


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.load()


def test_compute_without_gpu_fails_loudly():
    """On a machine without a CUDA device the compute entry point must print an error and exit non-zero."""
    code = ("import cfd_proxy_b200.mesh as M\nfrom cfd_proxy_b200.driver import Session\n"
            "S=Session(1); S.load_spec(M.make_spec((6,5,4),(1,1,1))); S.setup(device=False); S.iterate('comm_free',1)\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


# ---- NetCDF loader (read_netcdf.c:20-61 replacement) -------------------------------------------------
@pytest.mark.parametrize("version", [1, 2])
def test_loader_reads_f6_schema(tmp_path, version):
    spec = M.make_spec((10, 8, 6), (2, 2, 1), hexfrac=0.3)
    prefix = str(tmp_path / "dualgrid")
    doms = M.write_mesh(prefix, spec, lvl=2, version=version)
    lib = L.load()
    for r, d in enumerate(doms):
        path = M.domain_path(prefix, r, 2)
        ncid = C.c_int(-1)
        assert lib.cfdp_nc_open(path.encode(), 0, C.byref(ncid)) == 0
        for name in ("nfaces", "nownpoints", "nallpoints", "ndomains", "naddpoints", "ncommdomains"):
            key = {"nownpoints": "nown", "nallpoints": "nall", "naddpoints": "nadd"}.get(name, name)
            assert lib.get_nc_val(ncid.value, name.encode()) == d[key]
        assert lib.get_nc_val(ncid.value, b"ncolors") == 1
        fp = np.zeros((d["nfaces"], 2), np.int32)
        lib.get_nc_int(ncid.value, b"fpoint", fp.ctypes.data_as(L.c_int_p))
        assert np.array_equal(fp, d["fpoint"])
        fn = np.zeros((d["nfaces"], 3))
        lib.get_nc_double(ncid.value, b"fnormal", fn.ctypes.data_as(L.c_dbl_p))
        assert np.array_equal(fn, d["fnormal"])
        idx = np.zeros(d["nadd"], np.int32)
        lib.get_nc_int(ncid.value, b"addpoint_idx", idx.ctypes.data_as(L.c_int_p))
        assert np.array_equal(idx, d["addpoint_idx"])
        assert lib.cfdp_nc_close(ncid.value) == 0
        # independent readers agree: numpy reader of this repo and scipy's
        dims, vars_ = netcdf3.read_cdf(path)
        assert dims["nfaces"] == d["nfaces"] and np.array_equal(vars_["pvolume"], d["pvolume"])
        from scipy.io import netcdf_file
        with netcdf_file(path, "r", mmap=False) as f:
            assert np.array_equal(f.variables["fpoint"][:], d["fpoint"])
            assert np.array_equal(f.variables["sendcount"][:], d["sendcount"])


def test_loader_errors(tmp_path):
    lib = L.load()
    ncid = C.c_int(-1)
    assert lib.cfdp_nc_open(str(tmp_path / "missing").encode(), 0, C.byref(ncid)) != 0
    bad = tmp_path / "bad"
    bad.write_bytes(b"HDF5 is not CDF" * 10)
    rc = lib.cfdp_nc_open(str(bad).encode(), 0, C.byref(ncid))
    assert rc != 0 and b"format" in lib.cfdp_nc_strerror(rc)
    # truncated data section
    spec = M.make_spec((6, 5, 4), (1, 1, 1))
    doms = M.write_mesh(str(tmp_path / "g"), spec, lvl=1)
    path = M.domain_path(str(tmp_path / "g"), 0, 1)
    data = open(path, "rb").read()
    open(path, "wb").write(data[: len(data) // 2])
    assert lib.cfdp_nc_open(path.encode(), 0, C.byref(ncid)) != 0
    # get_nc_* print "Error: ..." and exit(2) like the reference (error_handling.h:6-10)
    code = ("import ctypes as C\nfrom cfd_proxy_b200 import lib as L\nlib=L.load()\n"
            f"n=C.c_int(); assert lib.cfdp_nc_open({str(tmp_path / 'ok_domain_0_lvl_1')!r}.encode(),0,C.byref(n))==0\n"
            "lib.get_nc_val(n.value,b'no_such_dim')\n")
    M.write_mesh(str(tmp_path / "ok"), spec, lvl=1)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 2 and "Error:" in r.stdout


# ---- comm tables (comm_data.c:116-255, :309-443) ---------------------------------------------------
@pytest.mark.parametrize("name", GOLDEN[:3])
def test_comm_tables_match_reference_golden(session_factory, tmp_path, name):
    z, spec, doms, lvl = load_golden(name)
    nd = len(doms)
    prefix = str(tmp_path / "dualgrid")
    M.write_mesh(prefix, spec, lvl=lvl)
    S = session_factory(nd)
    S.load_files(prefix, lvl)
    S.setup(device=False)
    for a, d in enumerate(S.domains):
        gs, gr = golden_index(z, a, nd)
        si, ri = d.index_lists()
        assert set(si) == set(gs) and set(ri) == set(gr)
        for k in gs:
            assert np.array_equal(si[k], gs[k])
            assert np.array_equal(S.pack_list(d, k), gs[k])          # device pack order == cd->sendindex[k] order
        for k in gr:
            assert np.array_equal(ri[k], gr[k])
            assert np.array_equal(S.unpack_list(d, k), gr[k])
        cd = d.cd
        ssz = rsz = 0
        for i in range(cd.ncommdomains):                              # comm_data.c:343-352
            k = cd.commpartner[i]
            assert cd.local_send_offset[k] == ssz and cd.local_recv_offset[k] == rsz
            ssz += cd.sendcount[k] * 21 * 8
            rsz += cd.recvcount[k] * 21 * 8
        assert cd.nreq == 2 * cd.ncommdomains


def test_single_domain_has_no_comm(session_factory):
    S = session_factory(1)
    S.load_spec(M.make_spec((8, 6, 5), (1, 1, 1)))
    S.setup(device=False)
    cd = S.domains[0].cd
    assert cd.ndomains == 1 and cd.ncommdomains == 0 and not cd.sendindex    # comm_data.c:83-86
    st = S.stats()
    assert st.send_rows_local == 0 and st.send_rows_remote == 0 and st.nboundary_tiles == 0


# ---- GPU face schedule: the invariants eval.c:88-235 checks for the CPU schedule ---------------------
SCHED_CASES = [((20, 16, 12), (2, 2, 1), "shuffle", 0.4, 64, 0), ((24, 20, 16), (1, 1, 1), "lex", 0.0, 256, 0),
               ((16, 16, 16), (2, 1, 1), "brick", 0.0, 128, 1), ((9, 7, 5), (2, 1, 1), "lex", 1.0, 16, 0)]


@pytest.mark.parametrize("n,p,order,hexfrac,tile,torder", SCHED_CASES)
def test_schedule_invariants(session_factory, n, p, order, hexfrac, tile, torder):
    spec = M.make_spec(n, p, order=order, brick=4, hexfrac=hexfrac)
    nd = p[0] * p[1] * p[2]
    doms = [M.gen_domain(spec, r) for r in range(nd)]
    recv, send = O.recvsend_index(doms) if nd > 1 else ([{}], [{}])
    S = session_factory(nd, tile_points=tile, tile_order=torder)
    S.load_spec(spec)
    S.setup(device=False)
    for a, d in enumerate(S.domains):
        dom = doms[a]
        nown, nall = dom["nown"], dom["nall"]
        sc = S.schedule(d)
        rows = sc["row_of_point"]
        # every point (own + ghost) has exactly one device row (eval.c:96-123)
        assert len(set(rows.tolist())) == nall and rows.min() >= 0 and rows.max() < sc["nrows"]
        # own points: each in exactly one tile => zero-initialised and scaled exactly once, written by one block (eval.c:126-231)
        covered = np.zeros(nown, int)
        row0, npts = sc["tile_row0"], sc["tile_npts"]
        assert (row0[:-1] % 16 == 0).all() and (npts <= tile).all() and npts.sum() == nown
        inv = np.full(sc["nrows"], -1)
        inv[rows] = np.arange(nall)
        is_send = O.is_send_mask(dom, send[a]).astype(bool)
        keep = (dom["fpoint"][:, 0] < nown) | (dom["fpoint"][:, 1] < nown)
        seen_face_ends = 0
        for t in range(sc["ntiles"]):
            pts = inv[row0[t]:row0[t] + npts[t]]
            assert (pts >= 0).all() and (pts < nown).all()            # ghosts are never written (eval.c:190-199)
            covered[pts] += 1
            # boundary tiles first; a tile is boundary iff it holds a send point (early send)
            assert bool(sc["tile_is_boundary"][t]) == bool(is_send[pts].any())
            assert bool(sc["tile_is_boundary"][t]) == (t < sc["nboundary"])
            faces, halo = S.tile(d, t, int(sc["tile_nfaces"][t]), int(sc["tile_nhalo"][t]))
            inc = np.isin(dom["fpoint"][:, 0], pts) | np.isin(dom["fpoint"][:, 1], pts)
            assert sorted(faces.tolist()) == np.nonzero(inc)[0].tolist()   # exactly the incident faces, each once
            ends = dom["fpoint"][faces].ravel()
            assert set(halo.tolist()) == set(ends.tolist()) - set(pts.tolist())
            seen_face_ends += int(np.isin(dom["fpoint"][faces], pts).sum())
        assert (covered == 1).all()
        # every (face, own endpoint) pair is computed exactly once; ghost-ghost faces are dropped (rangelist.c:513-523)
        assert seen_face_ends == int((dom["fpoint"][keep] < nown).sum())
    st = S.stats()
    assert st.nfaces == sum(int(((d["fpoint"][:, 0] < d["nown"]) | (d["fpoint"][:, 1] < d["nown"])).sum()) for d in doms)
    assert st.alg_bytes == sum(int(((d["fpoint"][:, 0] < d["nown"]) | (d["fpoint"][:, 1] < d["nown"])).sum()) * 32 + d["nall"] * 56 + d["nown"] * 176 for d in doms)
    assert st.lds_wavefronts_est >= st.lds_wavefronts_min > 0


def test_mesh_generator_invariants():
    """what any stand-in mesh must satisfy for the reference to accept it."""
    spec = M.f6like_spec(12, lvl=4)
    doms = [M.gen_domain(spec, r) for r in range(12)]
    total = 0
    for a, d in enumerate(doms):
        nown = d["nown"]
        fp = d["fpoint"]
        assert ((fp[:, 0] < nown) | (fp[:, 1] < nown)).all()              # no ghost-ghost faces
        assert np.isin(np.arange(d["nall"]), fp).all()                     # every point has a face
        assert (d["pvolume"] > 0).all()
        for k in range(12):
            assert d["sendcount"][k] == doms[k]["recvcount"][a]
            assert (d["recvcount"][k] > 0) == (k in d["commpartner"]) or d["sendcount"][k] > 0
            assert d["recvcount"][k] == int((d["addpoint_owner"] == k).sum())
        o, i = d["addpoint_owner"], d["addpoint_idx"]
        for j in range(d["nadd"]):
            assert doms[o[j]]["global_id"][i[j]] == d["global_id"][nown + j]
        total += nown
    assert total == spec.nx * spec.ny * spec.nz
    # var is keyed by the global id: ghosts agree with owners; C and numpy generators agree
    lib = L.load()
    v = M.var_for(doms[0])
    assert v[3, 2] == lib.cfdp_mesh_var_value(M.DEFAULT_SEED, int(doms[0]["global_id"][3]), 2)


@pytest.mark.parametrize("n,p,order,hexfrac,tile", [((20, 16, 12), (2, 2, 1), "shuffle", 0.4, 64), ((24, 20, 16), (3, 2, 2), "lex", 0.25, 256)])
def test_fused_pack_export_lists_cover_every_halo_row_once(session_factory, n, p, order, hexfrac, tile):
    """The boundary tiles' export lists (fused pack) against the halo lists, the check the reference makes for its
    per-colour send lists (thread_comm.c:159-205): every (ghost point of a hosted partner) is written exactly once,
    by the tile that owns the source point, from the point cd->sendindex / cd->recvindex pair up."""
    spec = M.make_spec(n, p, order=order, brick=4, hexfrac=hexfrac)
    nd = p[0] * p[1] * p[2]
    doms = [M.gen_domain(spec, r) for r in range(nd)]
    recv, send = O.recvsend_index(doms)
    S = session_factory(nd, tile_points=tile)
    S.load_spec(spec)
    S.setup(device=False)
    lib, st = S.lib, S.stats()
    want = {}   # (dst domain, ghost point) -> (src domain, own point)
    for a in range(nd):
        for k, ridx in recv[a].items():
            for j, g in enumerate(ridx):
                want[(a, int(g))] = (k, int(send[k][a][j]))
    got = {}
    for t in range(st.nboundary_tiles):
        ne = lib.cfdp_get_tile_exports(t, 0, None, None, None)
        src = (C.c_uint * max(ne, 1))(); dst = (C.c_uint * max(ne, 1))(); kind = (C.c_int * max(ne, 1))()
        assert lib.cfdp_get_tile_exports(t, ne, src, dst, kind) == ne
        for i in range(ne):
            assert kind[i] == 0                       # one process hosts every domain: no send buffer
            sd, sp, dd, dp = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            assert lib.cfdp_get_row_owner(src[i], C.byref(sd), C.byref(sp)) == 0
            assert lib.cfdp_get_row_owner(dst[i], C.byref(dd), C.byref(dp)) == 0
            key = (dd.value, dp.value)
            assert key not in got                     # each ghost row written exactly once
            got[key] = (sd.value, sp.value)
    assert got == want
    assert st.send_rows_local == len(want)


def test_lean_host_mode_releases_mesh_arrays(session_factory, monkeypatch):
    """CFDP_LEAN_HOST=1 (256 M-point device-resident runs): no host mirrors of grad / psd_flux, fpoint / fnormal are
    released once the schedule is built; the plan itself is unchanged."""
    spec = M.make_spec((16, 14, 12), (2, 2, 1), order="lex", hexfrac=0.3)
    S = session_factory(4, tile_points=64)
    S.load_spec(spec)
    S.setup(device=False)
    ref = S.stats()
    ref_tiles = [S.schedule(d)["tile_row0"].copy() for d in S.domains]
    monkeypatch.setenv("CFDP_LEAN_HOST", "1")
    S = session_factory(4, tile_points=64)
    S.load_spec(spec)
    for d in S.domains:
        assert not d.sd.grad and not d.sd.psd_flux and d.sd.fpoint and d.sd.var
    S.setup(device=False)
    st = S.stats()
    for d in S.domains:
        assert not d.sd.fpoint and not d.sd.fnormal
    assert (st.nfaces, st.ntiles, st.rows, st.blob_bytes) == (ref.nfaces, ref.ntiles, ref.rows, ref.blob_bytes)
    for d, t0 in zip(S.domains, ref_tiles):
        assert np.array_equal(S.schedule(d)["tile_row0"], t0)
