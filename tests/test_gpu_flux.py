"""GPU parity of the pseudo flux (src/flux.c:111-201), the consumer of the exchanged gradients,
called through the C ABI (`compute_psd_flux`, `cfdp_set_flux` + `cfdp_iterate`).

Bar: exact mode BIT-IDENTICAL to the oracle restatement (itself bit-identical to the unmodified reference run with one
OpenMP thread, tests/test_oracle.py) and to the reference's golden vectors; fused-multiply-add mode within
|a-b| <= 1e-12*|b| + 64*eps*S_p, S_p = sum over the faces contributing to p of 4*|n_f|_1*max|d_f|.
Only own rows of psd_flux are compared: the reference never zeroes ghost rows (flux.c:128-134, :179-183).
"""
import numpy as np
import pytest

import cfd_proxy_b200.mesh as M
from oracle import oracle as O
from helpers import GOLDEN, bits_differ, golden_flux, golden_grad, load_golden

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps

CASES = [
    # (lattice, domain grid, point order, hexfrac, tile_points, tile_order)
    ((12, 10, 8), (1, 1, 1), "lex", 0.0, 256, 0),
    ((24, 20, 16), (2, 2, 2), "lex", 0.25, 256, 0),
    ((24, 20, 16), (3, 2, 2), "shuffle", 0.4, 128, 0),
    ((32, 24, 16), (2, 2, 1), "brick", 0.0, 256, 1),
    ((9, 7, 5), (2, 1, 1), "lex", 1.0, 16, 0),       # hex-only, tiny tiles, ragged sizes
]


def oracle_chain(doms, exchange=True):
    nd = len(doms)
    recv, send = O.recvsend_index(doms) if nd > 1 else ([{}], [{}])
    grads = [O.gradients(d, M.var_for(d), is_send=O.is_send_mask(d, send[a]), order=1) for a, d in enumerate(doms)]
    if exchange and nd > 1:
        grads = O.exchange(grads, recv, send)
    else:
        for a, d in enumerate(doms):
            grads[a][d["nown"]:] = 0.0                 # the device ghost rows before any exchange
    flux = [O.psd_flux(d, grads[a], is_send=O.is_send_mask(d, send[a]), order=1) for a, d in enumerate(doms)]
    return grads, flux


@pytest.mark.parametrize("n,p,order,hexfrac,tile,torder", CASES)
@pytest.mark.parametrize("variant", ["comm_free", "mpi_bulk_sync", "mpi_async", "gaspi_async"])
@pytest.mark.parametrize("kernel", [2, 1])   # 2 = chunked kernel with prefetched halo row numbers (production), 1 = one tile per CTA
def test_flux_after_gradient_and_exchange_bit_identical(session_factory, monkeypatch, n, p, order, hexfrac, tile, torder, variant, kernel):
    monkeypatch.setenv("CFDP_FLUX_KERNEL", str(kernel))
    monkeypatch.setenv("CFDP_CHUNK", "3")
    spec = M.make_spec(n, p, order=order, brick=4, hexfrac=hexfrac)
    nd = p[0] * p[1] * p[2]
    doms = [M.gen_domain(spec, r) for r in range(nd)]
    _, want = oracle_chain(doms, exchange=(variant != "comm_free"))
    S = session_factory(nd, device=0, tile_points=tile, tile_order=torder)
    S.load_spec(spec)
    S.setup()
    S.lib.cfdp_set_exact(1)
    S.set_flux(True)
    for d in S.domains:
        d.psd_flux[:] = 7.0
    S.iterate(variant, 2)                               # solver.c:45-55: gradient (+ exchange), pseudo flux, twice
    S.download_flux()
    for a, d in enumerate(S.domains):
        nown = doms[a]["nown"]
        assert np.isfinite(d.psd_flux[:nown]).all()
        assert bits_differ(d.psd_flux[:nown], want[a][:nown]) == 0
        assert (d.psd_flux[nown:] == 7.0).all()         # ghost rows of the host array are left alone


@pytest.mark.parametrize("n,p,order,hexfrac,tile,torder", CASES[1:4])
def test_compute_psd_flux_reads_host_grad(session_factory, n, p, order, hexfrac, tile, torder):
    """The reference-named entry point with host arrays: whatever stands in sd->grad (ghost rows included) goes in."""
    spec = M.make_spec(n, p, order=order, brick=4, hexfrac=hexfrac)
    nd = p[0] * p[1] * p[2]
    doms = [M.gen_domain(spec, r) for r in range(nd)]
    recv, send = O.recvsend_index(doms)
    S = session_factory(nd, device=0, tile_points=tile, tile_order=torder)
    S.load_spec(spec)
    S.setup()
    rng = np.random.default_rng(7)
    for d in S.domains:
        d.grad[:] = rng.standard_normal(d.grad.shape) * 10.0 ** rng.integers(-3, 4, size=(d.grad.shape[0], 1, 1))
        d.psd_flux[:] = -3.0
    S.psd_flux()
    for a, d in enumerate(S.domains):
        nown = doms[a]["nown"]
        want = O.psd_flux(doms[a], d.grad, is_send=O.is_send_mask(doms[a], send[a]), order=1)
        assert bits_differ(d.psd_flux[:nown], want[:nown]) == 0
        assert (d.psd_flux[nown:] == -3.0).all()
    # a second call after the host changed grad sees the new values
    for d in S.domains:
        d.grad[:] *= 0.5
    S.psd_flux()
    for a, d in enumerate(S.domains):
        nown = doms[a]["nown"]
        want = O.psd_flux(doms[a], d.grad, is_send=O.is_send_mask(doms[a], send[a]), order=1)
        assert bits_differ(d.psd_flux[:nown], want[:nown]) == 0


@pytest.mark.parametrize("name", GOLDEN)
def test_flux_matches_reference_golden_vectors(session_factory, tmp_path, name):
    """gradient + exchange + pseudo flux on the GPU against psd_flux of the UNMODIFIED reference (one thread)."""
    z, spec, doms, lvl = load_golden(name)
    nd = len(doms)
    prefix = str(tmp_path / "dualgrid")
    M.write_mesh(prefix, spec, lvl=lvl)
    S = session_factory(nd, device=0)
    S.load_files(prefix, lvl)
    S.setup()
    S.set_flux(True)
    v = "mpi_async" if nd > 1 else "comm_free"
    S.iterate(v, 2)
    S.download_flux()
    for a, d in enumerate(S.domains):
        nown = doms[a]["nown"]
        assert bits_differ(d.psd_flux[:nown], golden_flux(z, v, a)[:nown]) == 0
    # and the entry point fed with the reference's own gradients
    for a, d in enumerate(S.domains):
        d.grad[:] = golden_grad(z, v, 1, a)
    if nd == 1:
        S.domains[0].grad[doms[0]["nown"]:] = 0.0
    S.psd_flux()
    for a, d in enumerate(S.domains):
        nown = doms[a]["nown"]
        assert bits_differ(d.psd_flux[:nown], golden_flux(z, v, a)[:nown]) == 0


def test_flux_fma_mode_within_tolerance(session_factory):
    spec = M.make_spec((24, 20, 16), (2, 2, 2), order="lex", brick=4, hexfrac=0.25)
    doms = [M.gen_domain(spec, r) for r in range(8)]
    grads, want = oracle_chain(doms)
    S = session_factory(8, device=0)
    S.load_spec(spec)
    S.setup()
    S.lib.cfdp_set_exact(1)
    S.iterate("mpi_async", 1)                           # exact gradients on the device ...
    S.lib.cfdp_set_exact(0)
    S.flux_iterate(1)                                   # ... contracted pseudo flux
    S.download_flux()
    differs = 0
    for a, d in enumerate(S.domains):
        nown = doms[a]["nown"]
        _, scale = O.psd_flux_numpy(doms[a], grads[a])
        err = np.abs(d.psd_flux[:nown] - want[a][:nown])
        assert (err <= 1e-12 * np.abs(want[a][:nown]) + 64 * EPS * scale[:, None]).all(), float(err.max())
        differs += bits_differ(d.psd_flux[:nown], want[a][:nown])
    assert differs > 0                                  # the contracted build is really a different rounding


def test_flux_medium_mesh(session_factory):
    """262 k points in 8 domains, production tile size: every own row bit-identical, timing sane."""
    spec = M.make_spec((64, 64, 64), (2, 2, 2), order="lex", jitter=0.1)
    doms = [M.gen_domain(spec, r) for r in range(8)]
    _, want = oracle_chain(doms)
    S = session_factory(8, device=0)
    S.load_spec(spec)
    S.setup()
    S.set_flux(True)
    S.iterate("mpi_async", 3)
    S.download_flux()
    for a, d in enumerate(S.domains):
        nown = doms[a]["nown"]
        assert bits_differ(d.psd_flux[:nown], want[a][:nown]) == 0
    ms = S.flux_iterate(5) / 5
    st = S.stats()
    assert st.flux_alg_bytes > 0 and 0 < ms < 50
    assert abs(st.last_flux_ms - ms) < 1e-6
