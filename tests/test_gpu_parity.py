"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle
(oracle/gg_oracle.c, itself bit-identical to the unmodified reference run with one thread).

Bar (BASELINE.json north_star / SURVEY 8(c)):
  * exact mode (default): gradients BIT-IDENTICAL to the oracle in the reference's single-thread
    summation order; ghost rows bit-identical to the owners' rows; halo lists bit-exact.
  * fma mode and any other summation order: |a-b| <= 1e-12*|b| + 64*eps*S_p  with
    S_p = sum_f |n_f|_1 max_eq|val_f| / vol_p   (fp64, differing only in rounding/summation order).
"""
import numpy as np
import pytest

import cfd_proxy_b200.mesh as M
from oracle import oracle as O

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def oracle_all(doms, exchange=True):
    recv, send = O.recvsend_index(doms) if len(doms) > 1 else ([{}], [{}])
    grads = [O.gradients(d, M.var_for(d), is_send=O.is_send_mask(d, send[a]), order=1) for a, d in enumerate(doms)]
    if exchange and len(doms) > 1:
        grads = O.exchange(grads, recv, send)
    return grads, recv, send


def bits_differ(a, b):
    return int((np.ascontiguousarray(a).view(np.uint64) != np.ascontiguousarray(b).view(np.uint64)).sum())


def within_tolerance(got, want, dom, var):
    scale = O.error_scale(dom, var)[:, None, None]
    nown = dom["nown"]
    err = np.abs(got[:nown] - want[:nown])
    return bool((err <= 1e-12 * np.abs(want[:nown]) + 64 * EPS * scale).all()), float(err.max())


CASES = [
    # (lattice, domain grid, point order, hexfrac, tile_points, tile_order)
    ((12, 10, 8), (1, 1, 1), "lex", 0.0, 256, 0),
    ((20, 16, 12), (1, 1, 1), "shuffle", 0.4, 64, 0),
    ((24, 20, 16), (2, 2, 2), "lex", 0.25, 256, 0),
    ((24, 20, 16), (3, 2, 2), "shuffle", 0.4, 128, 0),
    ((32, 24, 16), (2, 2, 1), "brick", 0.0, 256, 1),
    ((9, 7, 5), (2, 1, 1), "lex", 1.0, 16, 0),       # hex-only, tiny tiles, ragged sizes
]


@pytest.mark.parametrize("n,p,order,hexfrac,tile,torder", CASES)
@pytest.mark.parametrize("variant", ["comm_free", "mpi_bulk_sync", "mpi_early_recv", "mpi_async", "gaspi_async"])
@pytest.mark.parametrize("kernel", [2, 1])   # 2 = pipelined TMA kernel (production), 1 = one tile per CTA
def test_exact_mode_bit_identical(session_factory, monkeypatch, n, p, order, hexfrac, tile, torder, variant, kernel):
    monkeypatch.setenv("CFDP_KERNEL", str(kernel))
    monkeypatch.setenv("CFDP_CHUNK", "3")
    spec = M.make_spec(n, p, order=order, brick=4, hexfrac=hexfrac)
    nd = p[0] * p[1] * p[2]
    doms = [M.gen_domain(spec, r) for r in range(nd)]
    want, recv, send = oracle_all(doms, exchange=(variant != "comm_free"))
    S = session_factory(nd, device=0, tile_points=tile, tile_order=torder)
    S.load_spec(spec)
    S.setup()
    for d in S.domains:
        d.grad[:] = np.nan
    S.lib.cfdp_set_exact(1)
    S.iterate(variant, 2)
    S.download_grad()
    for a, d in enumerate(S.domains):
        nown = doms[a]["nown"]
        assert bits_differ(d.grad[:nown], want[a][:nown]) == 0
        if variant != "comm_free" and nd > 1:
            assert bits_differ(d.grad[nown:], want[a][nown:]) == 0      # ghost rows == owner rows
            for k in send[a]:
                assert np.array_equal(S.pack_list(d, k), send[a][k])
                assert np.array_equal(S.sendbuf(d, k), want[a][send[a][k]].reshape(-1, 21))  # threads.c:791-813


@pytest.mark.parametrize("n,p,order,hexfrac,tile,torder", CASES[:4])
def test_fma_mode_within_tolerance(session_factory, n, p, order, hexfrac, tile, torder):
    spec = M.make_spec(n, p, order=order, brick=4, hexfrac=hexfrac)
    nd = p[0] * p[1] * p[2]
    doms = [M.gen_domain(spec, r) for r in range(nd)]
    want, _, _ = oracle_all(doms)
    S = session_factory(nd, device=0, tile_points=tile, tile_order=torder)
    S.load_spec(spec)
    S.setup()
    S.lib.cfdp_set_exact(0)
    S.iterate("mpi_async", 1)
    S.download_grad()
    S.lib.cfdp_set_exact(1)
    for a, d in enumerate(S.domains):
        ok, mx = within_tolerance(d.grad, want[a], doms[a], M.var_for(doms[a]))
        assert ok, f"domain {a}: max abs err {mx}"
        nown = doms[a]["nown"]
        if nd > 1:
            # exchanged rows are raw copies whatever the arithmetic mode
            recv, send = O.recvsend_index(doms)
            for k, ridx in recv[a].items():
                assert bits_differ(d.grad[ridx], S.domains[k].grad[send[k][a]]) == 0


def test_dropin_call_sequence_from_files(session_factory, tmp_path):
    """The reference's main(): files -> read_* -> tables -> init_threads -> compute_gradients_gg_mpi_async,
    host buffers in and out (non-resident mode)."""
    import ctypes as C
    spec = M.f6like_spec(12, lvl=3)
    prefix = str(tmp_path / "dualgrid")
    doms = M.write_mesh(prefix, spec, lvl=3)
    want, _, _ = oracle_all(doms)
    S = session_factory(12, device=0)
    S.load_files(prefix, 3)
    S.setup()
    S.lib.cfdp_set_resident(0)
    for d in S.domains:
        d.grad[:] = np.nan
    for d in S.domains:   # every hosted "rank" makes the reference call; the first one drives the GPU
        S.lib.exchange_dbl_mpi_post_recv(C.byref(d.cd), 21)
        S.lib.compute_gradients_gg_mpi_async(C.byref(d.cd), C.byref(d.sd), 1)
    for a, d in enumerate(S.domains):
        assert bits_differ(d.grad, want[a]) == 0


def test_var_update_between_calls(session_factory):
    """sd->var is a host array the harness may rewrite between calls (SURVEY 8(b) ownership)."""
    import ctypes as C
    spec = M.make_spec((16, 12, 10), (1, 1, 1))
    dom = M.gen_domain(spec, 0)
    S = session_factory(1, device=0)
    S.load_spec(spec)
    S.setup()
    d = S.domains[0]
    rng = np.random.default_rng(7)
    for _ in range(2):
        v = rng.standard_normal((dom["nall"], 7))
        d.var[:] = v
        d.grad[:] = np.nan
        S.lib.compute_gradients_gg_comm_free(C.byref(d.cd), C.byref(d.sd), 0)
        want = O.gradients(dom, v, order=1)
        assert bits_differ(d.grad[:dom["nown"]], want[:dom["nown"]]) == 0


def test_var_refresh_on_the_device(session_factory):
    """cfdp_refresh_var / cfdp_set_var_refresh rebuild the per-tile copies of the halo var rows from the device var rows
    (a solver that changes var on the device): the result of an iteration that starts with the refresh is bit-identical,
    and the refresh is a real kernel launch with a device time."""
    spec = M.make_spec((24, 20, 16), (2, 2, 2), order="shuffle", hexfrac=0.25)
    doms = [M.gen_domain(spec, r) for r in range(8)]
    want, recv, send = oracle_all(doms, exchange=True)
    S = session_factory(8, device=0)
    S.load_spec(spec)
    S.setup()
    S.lib.cfdp_set_exact(1)
    l0 = S.stats().launches
    assert S.lib.cfdp_refresh_var(2) > 0.0
    assert S.stats().launches - l0 == 2
    S.lib.cfdp_set_var_refresh(1)
    for d in S.domains:
        d.grad[:] = np.nan
    S.iterate("mpi_async", 2)
    S.lib.cfdp_set_var_refresh(0)
    S.download_grad()
    for a, d in enumerate(S.domains):
        assert bits_differ(d.grad, want[a]) == 0


def test_large_mesh_properties(session_factory):
    """Full-size properties (no oracle run): linearity in var and exact reproducibility, and the two
    independent kernels agree bit for bit."""
    spec = M.make_spec((128, 128, 96), (2, 2, 2), order="lex", hexfrac=0.3)
    S = session_factory(8, device=0)
    S.load_spec(spec)
    S.setup()
    S.iterate("mpi_async", 1)
    S.download_grad()
    g1 = [d.grad.copy() for d in S.domains]
    S.iterate("mpi_bulk_sync", 1)
    S.download_grad()
    for a, d in enumerate(S.domains):
        assert bits_differ(d.grad, g1[a]) == 0            # deterministic, variant-independent
    for d in S.domains:
        d.var[:] *= 4.0                                    # power of two: exact scaling
    S.upload_var()
    S.iterate("mpi_async", 1)
    S.download_grad()
    for a, d in enumerate(S.domains):
        assert bits_differ(d.grad, 4.0 * g1[a]) == 0
    import os
    os.environ["CFDP_KERNEL"] = "1"
    try:
        S2 = session_factory(8, device=0)
        S2.load_spec(spec)
        S2.setup()
        S2.iterate("mpi_async", 1)
        S2.download_grad()
        for a, d in enumerate(S2.domains):
            assert bits_differ(d.grad, g1[a]) == 0
    finally:
        os.environ.pop("CFDP_KERNEL", None)


from helpers import GOLDEN, golden_grad, golden_index, load_golden  # noqa: E402


@pytest.mark.parametrize("name", GOLDEN)
def test_cuda_path_matches_reference_golden_vectors(session_factory, tmp_path, name):
    """The CUDA path against numbers the UNMODIFIED reference produced (tests/golden, one OpenMP thread):
    bit-identical own rows, ghost rows and halo lists; files read through the drop-in loader."""
    z, spec, doms, lvl = load_golden(name)
    nd = len(doms)
    prefix = str(tmp_path / "dualgrid")
    M.write_mesh(prefix, spec, lvl=lvl)
    S = session_factory(nd, device=0)
    S.load_files(prefix, lvl)
    S.setup()
    for d in S.domains:
        d.grad[:] = np.nan
    v = "mpi_async" if nd > 1 else "comm_free"
    S.iterate(v, 1)
    S.download_grad()
    for a, d in enumerate(S.domains):
        ref = golden_grad(z, v, 1, a)
        nown = doms[a]["nown"]
        assert bits_differ(d.grad[:nown], ref[:nown]) == 0
        if nd > 1:
            assert bits_differ(d.grad[nown:], ref[nown:]) == 0
            gs, gr = golden_index(z, a, nd)
            si, ri = d.index_lists()
            for k in gs:
                assert np.array_equal(si[k], gs[k])
            for k in gr:
                assert np.array_equal(ri[k], gr[k])
        # the 3-thread reference differs only by summation order
        ref3 = golden_grad(z, v, 3, a)
        ok, mx = within_tolerance(d.grad, ref3, doms[a], M.var_for(doms[a]))
        assert ok, mx


def test_reference_main_runs_on_the_library(tmp_path):
    """oracle/_ref/hybrid.f6.b200.exe = the reference's UNMODIFIED src/hybrid.f6.c linked against
    libcfdp_b200.so (built by `make -C oracle dropin` where /root/reference exists): the drop-in boundary
    carries the reference's own main() to '*** SUCCESS', and the gradients it leaves in sd->grad match the oracle."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(O.HERE), "oracle", "_ref", "hybrid.f6.b200.exe")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/hybrid.f6.b200.exe not built")
    spec = M.make_spec((24, 20, 16), (1, 1, 1), hexfrac=0.3)
    prefix = str(tmp_path / "dualgrid")
    doms = M.write_mesh(prefix, spec, lvl=1, with_var=False)
    r = subprocess.run([exe, "-lvl", "1", "dualgrid"], cwd=str(tmp_path), capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "*** SUCCESS" in r.stdout and "*** TIMINGS" in r.stdout and "comm_free:" in r.stdout
    # var == 1.0 (init_solver_data, solver_data.c:26-36): checksum of the oracle's gradients on the same mesh
    d = doms[0]
    g = O.gradients(d, np.ones((d["nall"], 7)), order=1)
    line = [l for l in r.stdout.splitlines() if "CHECKSUM (rank" in l][0]
    got = float(line.split(":")[1])
    want = 0.0
    for v in g[:d["nown"]].ravel():
        want += v
    assert got == want
    # ... and of the pseudo flux computed from them (solver.c:52, flux.c)
    fl = O.psd_flux(d, g, order=1)
    line = [l for l in r.stdout.splitlines() if "CHECKSUM psd_flux" in l][0]
    got = float(line.split(":")[1])
    want = 0.0
    for v in fl[:d["nown"]].ravel():
        want += v
    assert got == want
