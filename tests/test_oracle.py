"""Pins the oracle (oracle/gg_oracle.c + oracle/oracle.py) against the UNMODIFIED reference:
  * the committed golden fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py from
    oracle/_ref/ref_harness) -- runs everywhere, including the GPU box,
  * a live run of oracle/_ref when it has been built in this tree (the container with /root/reference).
The reference itself ships no golden vectors or numerical tests."""
import os

import numpy as np
import pytest

import cfd_proxy_b200.mesh as M
from oracle import oracle as O
from helpers import GOLDEN, bits_differ, golden_flux, golden_grad, golden_index, load_golden

EPS = np.finfo(np.float64).eps


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_bit_identical_to_reference_one_thread(name):
    z, spec, doms, lvl = load_golden(name)
    nd = len(doms)
    recv, send = O.recvsend_index(doms) if nd > 1 else ([{}], [{}])
    grads = [O.gradients(d, M.var_for(d), is_send=O.is_send_mask(d, send[a]), order=1) for a, d in enumerate(doms)]
    for a, d in enumerate(doms):
        ref = golden_grad(z, "comm_free", 1, a)
        assert bits_differ(ref[:d["nown"]], grads[a][:d["nown"]]) == 0
        assert np.isnan(ref[d["nown"]:]).all()            # comm_free never touches ghost rows (harness pre-fills NaN)
    if nd > 1:
        ex = O.exchange(grads, recv, send)
        for v in ("mpi_bulk_sync", "mpi_async"):
            for a in range(nd):
                assert bits_differ(golden_grad(z, v, 1, a), ex[a]) == 0      # own AND ghost rows
        for a in range(nd):
            gs, gr = golden_index(z, a, nd)
            assert set(gs) == set(send[a]) and set(gr) == set(recv[a])
            for k in gs:
                assert np.array_equal(gs[k], send[a][k])                       # comm_data.c:197-222
            for k in gr:
                assert np.array_equal(gr[k], recv[a][k])                       # comm_data.c:163-174


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_within_tolerance_of_threaded_reference(name):
    """With 3 OpenMP threads the reference sums in another order: SURVEY 8(c) tolerance."""
    z, spec, doms, lvl = load_golden(name)
    nd = len(doms)
    recv, send = O.recvsend_index(doms) if nd > 1 else ([{}], [{}])
    v = "mpi_async" if nd > 1 else "comm_free"
    for a, d in enumerate(doms):
        var = M.var_for(d)
        g = O.gradients(d, var, is_send=O.is_send_mask(d, send[a]), order=1)
        ref = golden_grad(z, v, 3, a)
        scale = O.error_scale(d, var)[:, None, None]
        err = np.abs(ref[:d["nown"]] - g[:d["nown"]])
        assert (err <= 1e-12 * np.abs(g[:d["nown"]]) + 64 * EPS * scale).all()


@pytest.mark.parametrize("name", GOLDEN)
def test_flux_oracle_bit_identical_to_reference_one_thread(name):
    """oracle_psd_flux (flux.c:111-201 restated) on the reference's own exchanged gradients == the reference's psd_flux."""
    z, spec, doms, lvl = load_golden(name)
    nd = len(doms)
    recv, send = O.recvsend_index(doms) if nd > 1 else ([{}], [{}])
    v = "mpi_async" if nd > 1 else "comm_free"
    for a, d in enumerate(doms):
        nown = d["nown"]
        fl = O.psd_flux(d, golden_grad(z, v, 1, a), is_send=O.is_send_mask(d, send[a]), order=1)
        assert bits_differ(golden_flux(z, v, a)[:nown], fl[:nown]) == 0
        assert np.isnan(fl[nown:]).all()                   # ghost rows are not defined and not written
        fl2, scale = O.psd_flux_numpy(d, golden_grad(z, v, 1, a))
        assert (np.abs(fl2 - fl[:nown]) <= 64 * EPS * scale[:, None]).all()   # independent restatement, other summation order


def test_numpy_restatement_agrees():
    _, spec, doms, _ = load_golden("single")
    d = doms[0]
    var = M.var_for(d)
    g1 = O.gradients(d, var, order=0)
    g2 = O.gradients_numpy(d, var)
    scale = O.error_scale(d, var)[:, None, None]
    assert (np.abs(g1[:d["nown"]] - g2[:d["nown"]]) <= 64 * EPS * scale).all()


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(1)
    data = rng.standard_normal((50, 21))
    idx = rng.permutation(50)[:17].astype(np.int32)
    import ctypes as C
    buf = np.zeros((17, 21))
    O.lib().oracle_pack(data.ctypes.data_as(C.POINTER(C.c_double)), 21, idx.ctypes.data_as(C.POINTER(C.c_int)), 17,
                        buf.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.array_equal(buf, data[idx])
    out = np.zeros_like(data)
    O.lib().oracle_unpack(out.ctypes.data_as(C.POINTER(C.c_double)), 21, idx.ctypes.data_as(C.POINTER(C.c_int)), 17,
                          buf.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.array_equal(out[idx], data[idx])


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("threads", [1, 4])
def test_live_reference_run(tmp_path, threads):
    spec = M.f6like_spec(12, lvl=4)
    prefix = str(tmp_path / "dualgrid")
    doms = M.write_mesh(prefix, spec, lvl=4)
    recv, send = O.recvsend_index(doms)
    ref = O.run_ref(prefix, 4, 12, "mpi_async", 2, str(tmp_path / "out"), threads=threads)
    grads = [O.gradients(d, M.var_for(d), is_send=O.is_send_mask(d, send[a]), order=1) for a, d in enumerate(doms)]
    grads = O.exchange(grads, recv, send)
    for a, d in enumerate(doms):
        if threads == 1:
            assert bits_differ(ref[a]["grad"], grads[a]) == 0
        else:
            scale = O.error_scale(d, M.var_for(d))[:, None, None]
            err = np.abs(ref[a]["grad"][:d["nown"]] - grads[a][:d["nown"]])
            assert (err <= 1e-12 * np.abs(grads[a][:d["nown"]]) + 64 * EPS * scale).all()
        for k in send[a]:
            assert np.array_equal(ref[a]["sendindex"][k], send[a][k])


@pytest.mark.skipif(not (O.have_ref() and os.path.exists(os.path.join(O.REF_DIR, "hybrid.f6.exe")) and os.path.exists(os.path.join(O.REF_DIR, "mesh_tool"))),
                    reason="oracle/_ref not built (needs /root/reference)")
def test_config1_reference_main_reports_its_timings():
    """BASELINE config 1 (tools/f6like_configs.py ref): the UNMODIFIED hybrid.f6.exe, 12 ranks over the shm-MPI shim, on the F6-like
    stand-in written by the standalone mesh_tool, prints its own report (solver.c:246-313) with non-zero MPI timings."""
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "f6like_configs.py"), "ref", "4"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "*** SUCCESS" in r.stdout
    t = dict(re.findall(r"^\s*(\w+):\s+([0-9.]+)\s*$", r.stdout, flags=re.M))
    assert int(float(t["nProc"])) == 12
    for k in ("comm_free", "exchange_dbl_mpi_bulk_sync", "exchange_dbl_mpi_early_recv", "exchange_dbl_mpi_async"):
        assert float(t[k]) > 0.0, k
    assert float(t["exchange_dbl_gaspi_async"]) == 0.0      # MPI-only build (USE_GASPI off)
