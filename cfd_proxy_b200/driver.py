"""Host-side driver above the C ABI: the call sequence of the reference's main()
(src/hybrid.f6.c:27-101) for one process = one GPU hosting one or several mesh domains.

    init_communication -> nc_open -> read_solver_data -> init_solver_data ->
    read_communication_data -> compute_communication_tables -> init_threads -> iterate

Everything that computes goes through libcfdp_b200.so; this module only owns ctypes structs
and numpy views of the host arrays inside them.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from . import lib as L
from . import mesh as M


class HostedDomain:
    def __init__(self, rank: int):
        self.rank = rank
        self.cd = L.CommData()
        self.sd = L.SolverData()
        self.info = None  # generator dict when built from a spec

    # numpy views (no copies) of the host containers the C side allocated
    @property
    def var(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.sd.var, shape=(self.sd.nallpoints, 7))

    @property
    def grad(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.sd.grad, shape=(self.sd.nallpoints, 7, 3))

    @property
    def psd_flux(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.sd.psd_flux, shape=(self.sd.nallpoints, 3))

    @property
    def fpoint(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.sd.fpoint, shape=(self.sd.nfaces, 2))

    @property
    def fnormal(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.sd.fnormal, shape=(self.sd.nfaces, 3))

    @property
    def pvolume(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.sd.pvolume, shape=(self.sd.nallpoints,))

    def index_lists(self):
        """(sendindex, recvindex) dicts partner -> int32 array, read back from comm_data."""
        cd = self.cd
        send, recv = {}, {}
        if cd.ndomains <= 1:
            return send, recv
        for i in range(cd.ncommdomains):
            k = cd.commpartner[i]
            if cd.sendcount[k] > 0:
                send[k] = np.ctypeslib.as_array(cd.sendindex[k], shape=(cd.sendcount[k],)).copy()
            if cd.recvcount[k] > 0:
                recv[k] = np.ctypeslib.as_array(cd.recvindex[k], shape=(cd.recvcount[k],)).copy()
        return send, recv

    def as_dict(self):
        """The domain in the generator's dict form (for the oracle)."""
        cd, sd = self.cd, self.sd
        d = dict(nfaces=sd.nfaces, nown=sd.nownpoints, nall=sd.nallpoints, nadd=sd.nallpoints - sd.nownpoints,
                 ndomains=cd.ndomains, rank=self.rank, fpoint=self.fpoint.copy(), fnormal=self.fnormal.copy(),
                 pvolume=self.pvolume.copy())
        if cd.ndomains > 1:
            d["ncommdomains"] = cd.ncommdomains
            d["commpartner"] = np.ctypeslib.as_array(cd.commpartner, shape=(cd.ncommdomains,)).copy()
            d["sendcount"] = np.ctypeslib.as_array(cd.sendcount, shape=(cd.ndomains,)).copy()
            d["recvcount"] = np.ctypeslib.as_array(cd.recvcount, shape=(cd.ndomains,)).copy()
            d["addpoint_owner"] = np.ctypeslib.as_array(cd.addpoint_owner, shape=(cd.naddpoints,)).copy()
            d["addpoint_idx"] = np.ctypeslib.as_array(cd.addpoint_id, shape=(cd.naddpoints,)).copy()
        else:
            d["ncommdomains"] = 0
            d["commpartner"] = np.zeros(0, np.int32)
        return d


class Session:
    """One process = one GPU.  Hosts domains [first, first+count) of `ndomains`."""

    _active = None

    def __init__(self, ndomains: int, proc_rank: int = 0, nprocs: int = 1, device: int = -1,
                 tile_points: int | None = None, tile_order: int | None = None):
        if Session._active is not None:
            raise RuntimeError("one Session per process at a time (close() the previous one)")
        self.lib = L.load()
        if tile_points is not None:
            os.environ["CFDP_TILE_POINTS"] = str(tile_points)
        if tile_order is not None:
            os.environ["CFDP_TILE_ORDER"] = str(tile_order)
        rc = self.lib.cfdp_configure(proc_rank, nprocs, ndomains, device)
        if rc != 0:
            raise RuntimeError(f"cfdp_configure failed rc={rc}")
        self.ndomains, self.proc_rank, self.nprocs = ndomains, proc_rank, nprocs
        self.per_proc = ndomains // nprocs
        self.first = proc_rank * self.per_proc
        self.domains: list[HostedDomain] = []
        self._files = []
        self._setup_done = False
        self.timing = {}       # seconds spent in the setup phases (bench.py reports them)
        Session._active = self

    # ---- loading -------------------------------------------------------------------------
    def hosted_ranks(self):
        return range(self.first, self.first + self.per_proc)

    def load_files(self, prefix: str, lvl: int):
        """hybrid.f6.c:57-66: open `<prefix>_domain_<rank>_lvl_<lvl>` and read it (every hosted rank)."""
        lib = self.lib
        for r in self.hosted_ranks():
            d = HostedDomain(r)
            lib.cfdp_init_communication_domain(C.byref(d.cd), r)
            ncid = C.c_int(-1)
            path = M.domain_path(prefix, r, lvl)
            rc = lib.cfdp_nc_open(path.encode(), 0, C.byref(ncid))
            if rc != 0:
                raise RuntimeError(f"cfdp_nc_open({path}): {lib.cfdp_nc_strerror(rc).decode()}")
            lib.read_solver_data(ncid.value, C.byref(d.sd))
            lib.init_solver_data(C.byref(d.sd), 25)
            lib.read_communication_data(ncid.value, C.byref(d.cd))
            lib.cfdp_nc_close(ncid.value)
            vpath = path + ".var"
            if os.path.exists(vpath):
                d.var[:] = np.fromfile(vpath, dtype="<f8").reshape(-1, 7)
            self.domains.append(d)
        return self

    def load_spec(self, spec: L.MeshSpec, seed=M.DEFAULT_SEED, keep_info: bool = False):
        """Generate the hosted domains in memory (no files) and attach them."""
        lib = self.lib
        assert spec.px * spec.py * spec.pz == self.ndomains
        t0 = time.time()
        for r in self.hosted_ranks():
            d = HostedDomain(r)
            lib.cfdp_init_communication_domain(C.byref(d.cd), r)
            m = L.MeshDomain()
            rc = lib.cfdp_mesh_gen_domain(C.byref(spec), r, C.byref(m))
            if rc != 0:
                raise RuntimeError(f"cfdp_mesh_gen_domain rc={rc}")
            lib.cfdp_attach_mesh_take(C.byref(m), C.byref(d.cd), C.byref(d.sd))
            lib.cfdp_mesh_fill_var(C.byref(m), seed, d.sd.var)
            if keep_info:
                d.info = dict(global_id=np.ctypeslib.as_array(m.global_id, shape=(m.nall,)).copy())
            lib.cfdp_mesh_free_domain(C.byref(m))
            self.domains.append(d)
        self.timing["mesh_s"] = round(time.time() - t0, 2)
        return self

    # ---- setup ---------------------------------------------------------------------------
    def nccl_bootstrap(self, uid_bytes: bytes):
        buf = C.create_string_buffer(uid_bytes, 128)
        self.lib.cfdp_nccl_init(buf)

    def unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self.lib.cfdp_nccl_get_unique_id(buf)
        return buf.raw

    def setup(self, device: bool = True):
        lib = self.lib
        t0 = time.time()
        for d in self.domains:
            lib.compute_communication_tables(C.byref(d.cd))
        for d in self.domains:
            lib.init_threads(C.byref(d.cd), C.byref(d.sd), 1)
        t1 = time.time()
        lib.cfdp_plan()                 # host: face schedules, device row numbering, exchange row lists
        t2 = time.time()
        if device:
            lib.cfdp_commit()           # device: allocations, uploads, kernel configuration, IPC / NCCL setup
        self.timing.update(tables_s=round(t1 - t0, 2), plan_s=round(t2 - t1, 2), commit_s=round(time.time() - t2, 2))
        self._setup_done = True
        return self

    # ---- compute -------------------------------------------------------------------------
    def upload_var(self):
        for d in self.domains:
            self.lib.cfdp_var_to_device(C.byref(d.sd))
        self.lib.cfdp_device_synchronize()

    def download_grad(self):
        for d in self.domains:
            self.lib.cfdp_grad_to_host(C.byref(d.sd))

    def upload_grad(self):
        for d in self.domains:
            self.lib.cfdp_grad_to_device(C.byref(d.sd))
        self.lib.cfdp_device_synchronize()

    def download_flux(self):
        for d in self.domains:
            self.lib.cfdp_flux_to_host(C.byref(d.sd))

    def set_flux(self, on=True):
        """cfdp_iterate also runs the pseudo flux after every gradient + exchange (solver.c:45-55)."""
        self.lib.cfdp_set_flux(1 if on else 0)

    def flux_iterate(self, niter=1) -> float:
        """Pseudo-flux passes only, on the device grad as it stands; returns device milliseconds."""
        return self.lib.cfdp_flux_iterate(niter)

    def psd_flux(self):
        """The reference-named entry point (flux.h:12) for all hosted domains: host grad in, host psd_flux out."""
        self.lib.compute_psd_flux(C.byref(self.domains[0].sd))

    def iterate(self, variant="mpi_async", niter=1) -> float:
        """Device-resident iterations over all hosted domains; returns device milliseconds."""
        return self.lib.cfdp_iterate(L.VARIANTS[variant] if isinstance(variant, str) else variant, niter, 1)

    def step_e2e(self, variant="mpi_async") -> float:
        return self.lib.cfdp_step_e2e(L.VARIANTS[variant] if isinstance(variant, str) else variant)

    def stats(self) -> L.Stats:
        st = L.Stats()
        self.lib.cfdp_get_stats(C.byref(st))
        return st

    def schedule(self, d: HostedDomain):
        v = L.ScheduleView()
        rc = self.lib.cfdp_get_schedule(C.byref(d.sd), C.byref(v))
        if rc != 0:
            raise RuntimeError("schedule not built")
        nt = v.ntiles
        a = np.ctypeslib.as_array
        return dict(ntiles=nt, nboundary=v.nboundary_tiles, nrows=v.nrows,
                    row_of_point=a(v.row_of_point, shape=(d.sd.nallpoints,)).copy(),
                    tile_row0=a(v.tile_row0, shape=(nt + 1,)).copy(), tile_npts=a(v.tile_npts, shape=(nt,)).copy(),
                    tile_nfaces=a(v.tile_nfaces, shape=(nt,)).copy(), tile_nhalo=a(v.tile_nhalo, shape=(nt,)).copy(),
                    tile_is_boundary=a(v.tile_is_boundary, shape=(nt,)).copy())

    def tile(self, d: HostedDomain, t: int, nfaces: int, nhalo: int):
        f = np.zeros(max(nfaces, 1), np.int32)
        h = np.zeros(max(nhalo, 1), np.int32)
        n = self.lib.cfdp_get_tile(C.byref(d.sd), t, f.ctypes.data_as(L.c_int_p), h.ctypes.data_as(L.c_int_p))
        assert n == nfaces
        return f[:nfaces], h[:nhalo]

    def tile_blob(self, d: HostedDomain, t: int, flux: bool = False):
        """Raw blob of tile t of domain d (host copy, before cfdp_commit): dict with the descriptor fields, normals
        [nfaces,3], halo device rows [nhalo] and the ELL adjacency [maxdeg, npad] (uint32)."""
        desc = (C.c_uint * 8)()
        n = self.lib.cfdp_get_tile_blob(C.byref(d.sd), t, 1 if flux else 0, desc, None, 0)
        if n < 0:
            raise RuntimeError("tile blob not available (plan first, before commit)")
        buf = np.zeros(max(int(n), 1), np.uint8)
        self.lib.cfdp_get_tile_blob(C.byref(d.sd), t, 1 if flux else 0, desc, buf.ctypes.data_as(C.POINTER(C.c_ubyte)), int(n))
        row0, npts, nhalo, nfaces, maxdeg, npad, nbytes, halo_off = (int(x) for x in desc)
        nfaces, zslot = nfaces & 0xFFFF, nfaces >> 16
        adj_off = halo_off + (nhalo * 4 + 15) // 16 * 16
        return dict(row0=row0, npts=npts, nhalo=nhalo, nfaces=nfaces, zslot=zslot, maxdeg=maxdeg, npad=npad, bytes=nbytes,
                    normals=buf[:nfaces * 24].view(np.float64).reshape(nfaces, 3),
                    halo_rows=buf[halo_off:halo_off + nhalo * 4].view(np.uint32),
                    ell=buf[adj_off:adj_off + maxdeg * npad * 4].view(np.uint32).reshape(maxdeg, npad))

    def row_owner(self, row: int):
        dom, pt = C.c_int(), C.c_int()
        rc = self.lib.cfdp_get_row_owner(int(row), C.byref(dom), C.byref(pt))
        return (dom.value, pt.value) if rc == 0 else None

    def pack_list(self, d: HostedDomain, k: int) -> np.ndarray:
        out = np.zeros(max(d.cd.sendcount[k], 1), np.int32)
        n = self.lib.cfdp_get_pack_list(C.byref(d.cd), k, out.ctypes.data_as(L.c_int_p))
        return out[:n]

    def unpack_list(self, d: HostedDomain, k: int) -> np.ndarray:
        out = np.zeros(max(d.cd.recvcount[k], 1), np.int32)
        n = self.lib.cfdp_get_unpack_list(C.byref(d.cd), k, out.ctypes.data_as(L.c_int_p))
        return out[:n]

    def sendbuf(self, d: HostedDomain, k: int) -> np.ndarray:
        out = np.zeros((max(d.cd.sendcount[k], 1), 21))
        n = self.lib.cfdp_get_sendbuf(C.byref(d.cd), k, out.ctypes.data_as(L.c_dbl_p))
        return out[:n]

    def close(self):
        if Session._active is self:
            self.lib.cfdp_finalize()
            Session._active = None
        self.domains = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


# ------------------------------------------------------------------------------------------
# torch.distributed plumbing (one process per GPU, launched by torchrun)
# ------------------------------------------------------------------------------------------
_int_exchange_keepalive = None


def install_dist_int_exchange():
    """Route the setup-time index handshake (comm_data.c:195-250) through torch.distributed
    point-to-point ops.  Used with the gloo backend on machines without a GPU; with NCCL the
    library's own NCCL path is used instead."""
    global _int_exchange_keepalive
    import torch
    import torch.distributed as dist

    def cb(n, peer, sbuf, scount, rbuf, rcount):
        reqs, recvs = [], []
        for i in range(n):
            if scount[i] > 0:
                t = torch.from_numpy(np.ctypeslib.as_array(sbuf[i], shape=(scount[i],)).copy())
                reqs.append(dist.isend(t, dst=peer[i]))
            if rcount[i] > 0:
                t = torch.empty(rcount[i], dtype=torch.int32)
                reqs.append(dist.irecv(t, src=peer[i]))
                recvs.append((i, t))
        for r in reqs:
            r.wait()
        for i, t in recvs:
            np.ctypeslib.as_array(rbuf[i], shape=(rcount[i],))[:] = t.numpy()

    _int_exchange_keepalive = L.INT_EXCHANGE_FN(cb)
    L.load().cfdp_set_int_exchange(_int_exchange_keepalive)


def session_from_env(ndomains: int, backend: str | None = None, **kw) -> "Session":
    """Create the Session of this torchrun rank; bootstraps NCCL inside the library by shipping
    the unique id over torch.distributed (plumbing only)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        return Session(ndomains, 0, 1, device=kw.pop("device", 0), **kw)
    import torch
    import torch.distributed as dist
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if not dist.is_initialized():
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    S = Session(ndomains, rank, world, device=local if backend == "nccl" else -1, **kw)
    if backend == "nccl":
        obj = [S.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        S.nccl_bootstrap(obj[0])
    else:
        install_dist_int_exchange()
    return S
