"""GPU parity at BASELINE sizes, a soak test of the tile pipeline, and the inter-GPU data path driven on ONE GPU.

  * F6-like stand-in dualgrid.12 / dualgrid.24 at level 1 (2.07 M points: BASELINE configs[1], [2]), all domains on one
    GPU, and one 8 M-point domain (the size of a bench-mesh domain): every own and ghost row bit-identical to the oracle
    (oracle/gg_oracle.c, itself pinned to the unmodified reference), halo lists bit-exact.
  * Soak: 500 iterations per configuration with a bitwise comparison every 50 -- the guard of the mbarrier / TMA
    protocol of gg_tile_pipe_kernel while compute-sanitizer is closed on the pool (chunks of 1 / 3 / 8 tiles per CTA,
    interleaved persistent CTAs, 64- and 256-point tiles, shuffled numbering, exchange variants included).
  * Loopback (CFDP_LOOPBACK=1): halo rows between the domains of one GPU take the path rows between GPUs take -- fused
    pack into the send buffer + transfer + unpack (NCCL variants), put + notify through the receive window
    (one-sided bulk variants), direct stores into the "peer's" ghost rows with arrival counters and write credits
    (one-sided async variants) -- so that a one-GPU box verifies them against the oracle (exchange_data_mpi.c:189,
    thread_comm.c:159-205 are the reference's in-line checks of the same things).
"""
import ctypes as C

import numpy as np
import pytest

import cfd_proxy_b200.mesh as M
from oracle import oracle as O
from helpers import bits_differ

pytestmark = pytest.mark.gpu


def oracle_all(doms):
    recv, send = O.recvsend_index(doms) if len(doms) > 1 else ([{}], [{}])
    grads = [O.gradients(d, M.var_for(d), is_send=O.is_send_mask(d, send[a]), order=1) for a, d in enumerate(doms)]
    if len(doms) > 1:
        grads = O.exchange(grads, recv, send)
    return grads, recv, send


@pytest.mark.parametrize("ndomains", [12, 24])
def test_f6like_level1_bit_identical(session_factory, ndomains):
    spec = M.f6like_spec(ndomains, lvl=1)
    doms = [M.gen_domain(spec, r) for r in range(ndomains)]
    assert 2.0e6 < sum(d["nown"] for d in doms) < 2.2e6
    want, recv, send = oracle_all(doms)
    S = session_factory(ndomains, device=0)
    S.load_spec(spec)
    S.setup()
    for variant in ("mpi_async", "gaspi_bulk_sync"):
        for d in S.domains:
            d.grad[:] = np.nan
        S.iterate(variant, 2)
        S.download_grad()
        for a, d in enumerate(S.domains):
            assert bits_differ(d.grad, want[a]) == 0, f"{variant}: domain {a}"
    for a, d in enumerate(S.domains):                       # halo lists: bit-exact (comm_data.c:116-255)
        s_got, r_got = d.index_lists()
        assert set(s_got) == set(send[a]) and set(r_got) == set(recv[a])
        for k in send[a]:
            assert np.array_equal(s_got[k], send[a][k]) and np.array_equal(S.pack_list(d, k), send[a][k])
        for k in recv[a]:
            assert np.array_equal(r_got[k], recv[a][k]) and np.array_equal(S.unpack_list(d, k), recv[a][k])


def test_eight_million_point_domain_bit_identical(session_factory):
    spec = M.make_spec((200, 200, 200), (1, 1, 1), order="lex", jitter=0.1, allow_big=True)
    dom = M.gen_domain(spec, 0)
    assert dom["nown"] == 8_000_000
    want = O.gradients(dom, M.var_for(dom), order=1)
    S = session_factory(1, device=0)
    S.load_spec(spec)
    S.setup()
    S.domains[0].grad[:] = np.nan
    S.iterate("comm_free", 2)
    S.download_grad()
    assert bits_differ(S.domains[0].grad, want) == 0


SOAK = [
    # (lattice, grid, order, hexfrac, tile, chunk, persistent, variant, loopback)
    ((48, 40, 32), (2, 2, 2), "shuffle", 0.3, 64, 1, 0, "mpi_async", 0),
    ((48, 40, 32), (2, 2, 2), "lex", 0.0, 256, 3, 0, "mpi_async", 0),
    ((64, 48, 40), (2, 2, 2), "shuffle", 0.25, 256, 8, 0, "gaspi_async", 1),
    ((64, 48, 40), (2, 2, 1), "lex", 0.0, 256, 1, 296, "gaspi_bulk_sync", 1),
    ((40, 40, 40), (1, 1, 1), "brick", 0.5, 128, 5, 0, "comm_free", 0),
]


@pytest.mark.parametrize("n,p,order,hexfrac,tile,chunk,persistent,variant,loopback", SOAK)
def test_soak_500_iterations(session_factory, monkeypatch, n, p, order, hexfrac, tile, chunk, persistent, variant, loopback):
    if loopback:
        monkeypatch.setenv("CFDP_LOOPBACK", "1")
    spec = M.make_spec(n, p, order=order, brick=4, hexfrac=hexfrac)
    nd = p[0] * p[1] * p[2]
    doms = [M.gen_domain(spec, r) for r in range(nd)]
    want, _, _ = oracle_all(doms)
    if variant == "comm_free":
        want = [O.gradients(d, M.var_for(d), order=1) for d in doms]
    S = session_factory(nd, device=0, tile_points=tile)
    S.load_spec(spec)
    S.setup()
    assert S.lib.cfdp_set_kernel(2, chunk, persistent) == 2
    for rnd in range(10):
        for d in S.domains:
            d.grad[:] = np.nan
        if variant != "comm_free":
            S.upload_grad()              # ghost rows NaN on the device as well: every round must refill them
        S.iterate(variant, 50)
        S.download_grad()
        for a, d in enumerate(S.domains):
            nown = doms[a]["nown"]
            rows = slice(None) if variant != "comm_free" else slice(0, nown)
            assert bits_differ(d.grad[rows], want[a][rows]) == 0, f"round {rnd}, domain {a}"


@pytest.mark.parametrize("variant,transport", [("mpi_bulk_sync", 2), ("mpi_early_recv", 2), ("mpi_async", 2),
                                               ("gaspi_bulk_sync", 3), ("gaspi_async", 4)])
def test_loopback_drives_the_inter_gpu_path(session_factory, monkeypatch, variant, transport):
    monkeypatch.setenv("CFDP_LOOPBACK", "1")
    spec = M.make_spec((24, 20, 16), (3, 2, 2), order="shuffle", brick=4, hexfrac=0.4)
    doms = [M.gen_domain(spec, r) for r in range(12)]
    want, recv, send = oracle_all(doms)
    want_flux = [O.psd_flux(d, want[a], is_send=O.is_send_mask(d, send[a]), order=1) for a, d in enumerate(doms)]
    S = session_factory(12, device=0, tile_points=128)
    S.load_spec(spec)
    S.setup()
    st = S.stats()
    assert st.loopback == 1 and st.send_rows_local == 0 and st.send_rows_remote == sum(int(d["sendcount"].sum()) for d in doms)
    for d in S.domains:
        d.grad[:] = np.nan
        d.psd_flux[:] = np.nan
    S.upload_grad()
    S.set_flux(True)      # the pseudo flux reads the ghost rows of every iteration: the write credits of the direct stores are exercised
    S.iterate(variant, 4)
    S.set_flux(False)
    assert S.stats().transport == transport
    S.download_grad()
    S.download_flux()
    for a, d in enumerate(S.domains):
        nown = doms[a]["nown"]
        assert bits_differ(d.grad, want[a]) == 0, f"domain {a}"
        assert bits_differ(d.psd_flux[:nown], want_flux[a][:nown]) == 0
        for k in send[a]:
            assert np.array_equal(S.sendbuf(d, k), want[a][send[a][k]].reshape(-1, 21))      # threads.c:791-813


@pytest.mark.parametrize("loopback", [0, 1])
def test_one_sided_entry_points_with_several_domains(session_factory, monkeypatch, loopback):
    """compute_gradients_gg_mpifence_* / _mpipscw_* / _gaspi_* (gradients.h:15-25) with ndomains > 1, host buffers in and out."""
    if loopback:
        monkeypatch.setenv("CFDP_LOOPBACK", "1")
    spec = M.make_spec((20, 16, 12), (2, 2, 1), order="lex", hexfrac=0.5)
    doms = [M.gen_domain(spec, r) for r in range(4)]
    want, _, _ = oracle_all(doms)
    S = session_factory(4, device=0, tile_points=64)
    S.load_spec(spec)
    S.setup()
    S.lib.cfdp_set_resident(0)
    for name in ("mpifence_bulk_sync", "mpifence_async", "mpipscw_bulk_sync", "mpipscw_async", "gaspi_bulk_sync", "gaspi_async"):
        fn = getattr(S.lib, "compute_gradients_gg_" + name)
        for d in S.domains:
            d.grad[:] = np.nan
        for it in range(2):
            for d in S.domains:
                fn(C.byref(d.cd), C.byref(d.sd), int(it == 1))
        for a, d in enumerate(S.domains):
            assert bits_differ(d.grad, want[a]) == 0, f"{name}: domain {a}"


@pytest.mark.parametrize("env,tile,variant,loopback", [
    ({"CFDP_VARIANT": "0"}, 256, "mpi_async", 0),                 # no L2 prefetch of the next tile
    ({"CFDP_VARIANT": "2"}, 256, "mpi_async", 0),                 # all of the next tile asked into L2
    ({"CFDP_CTAS": "3"}, 128, "mpi_async", 0),                    # three CTAs of <= 160 threads per SM
    ({"CFDP_CTAS": "4"}, 128, "gaspi_async", 1),                  # four CTAs of 128 threads per SM, direct halo stores
    ({"CFDP_DIRECT_SPREAD": "60"}, 64, "gaspi_async", 1),         # boundary tiles dealt out over the walk (direct stores only)
    ({"CFDP_FLUX_VARIANT": "0"}, 256, "mpi_async", 0),            # pseudo flux without the L2 row prefetch
    ({"CFDP_FLUX_VARIANT": "3"}, 256, "mpi_async", 0),
])
def test_optional_kernel_shapes_bit_identical(session_factory, monkeypatch, env, tile, variant, loopback):
    """The switches that are off by default (measured slower, profiles/README.md) stay correct: same bits as the oracle."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    monkeypatch.setenv("CFDP_CHUNK", "3")
    if loopback:
        monkeypatch.setenv("CFDP_LOOPBACK", "1")
    spec = M.make_spec((28, 24, 20), (2, 2, 2), order="shuffle", brick=4, hexfrac=0.3)
    doms = [M.gen_domain(spec, r) for r in range(8)]
    want, recv, send = oracle_all(doms)
    want_flux = [O.psd_flux(d, want[a], is_send=O.is_send_mask(d, send[a]), order=1) for a, d in enumerate(doms)]
    S = session_factory(8, device=0, tile_points=tile)
    S.load_spec(spec)
    S.setup()
    for d in S.domains:
        d.grad[:] = np.nan
        d.psd_flux[:] = np.nan
    S.upload_grad()
    S.set_flux(True)
    S.iterate(variant, 3)
    S.set_flux(False)
    S.download_grad()
    S.download_flux()
    for a, d in enumerate(S.domains):
        nown = doms[a]["nown"]
        assert bits_differ(d.grad, want[a]) == 0, f"domain {a}"
        assert bits_differ(d.psd_flux[:nown], want_flux[a][:nown]) == 0
