"""ctypes binding of libcfdp_b200.so (the C ABI declared in include/cfdp_b200.h).

There is deliberately no fallback: if the shared library (and with it the CUDA kernels)
is missing, importing the compute entry points raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcfdp_b200.so")

NGRAD, NFLUX, DIM2 = 7, 3, 21

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)


class RangeList(C.Structure):
    pass


RangeList._fields_ = [
    ("succ", C.POINTER(RangeList)),
    ("start", C.c_int), ("stop", C.c_int), ("ftype", C.c_int),
    ("nall_points_of_color", C.c_int), ("all_points_of_color", c_int_p),
    ("nfirst_points_of_color", C.c_int), ("first_points_of_color", c_int_p),
    ("nlast_points_of_color", C.c_int), ("last_points_of_color", c_int_p),
    ("nsendcount", C.c_int), ("sendpartner", c_int_p), ("sendcount", c_int_p),
    ("sendindex", C.POINTER(c_int_p)), ("sendoffset", C.POINTER(c_int_p)),
    ("nrecvcount", C.c_int), ("recvpartner", c_int_p), ("recvcount", c_int_p),
    ("recvindex", C.POINTER(c_int_p)), ("recvoffset", C.POINTER(c_int_p)),
    ("tid", C.c_int),
]


class SolverData(C.Structure):  # solver_data, reference src/solver_data.h:66-81
    _fields_ = [
        ("nfaces", C.c_int), ("nallfaces", C.c_int), ("nownpoints", C.c_int), ("nallpoints", C.c_int),
        ("ncolors", C.c_int),
        ("fpoint", c_int_p), ("fnormal", c_dbl_p), ("pvolume", c_dbl_p),
        ("var", c_dbl_p), ("grad", c_dbl_p), ("psd_flux", c_dbl_p),
        ("fcolor", C.POINTER(RangeList)), ("niter", C.c_int),
    ]


class CommData(C.Structure):  # comm_data, reference src/comm_data.h:15-55
    _fields_ = [
        ("nProc", C.c_int), ("iProc", C.c_int), ("ndomains", C.c_int), ("ncommdomains", C.c_int),
        ("nownpoints", C.c_int), ("naddpoints", C.c_int),
        ("addpoint_owner", c_int_p), ("addpoint_id", c_int_p), ("commpartner", c_int_p),
        ("sendcount", c_int_p), ("recvcount", c_int_p),
        ("recvindex", C.POINTER(c_int_p)), ("sendindex", C.POINTER(c_int_p)),
        ("nreq", C.c_int), ("req", C.c_void_p), ("stat", C.c_void_p),
        ("recvbuf", C.POINTER(c_dbl_p)), ("sendbuf", C.POINTER(c_dbl_p)),
        ("remote_recv_offset", C.POINTER(C.c_ulong)), ("local_recv_offset", C.POINTER(C.c_ulong)),
        ("local_send_offset", C.POINTER(C.c_ulong)), ("notification", C.POINTER(C.c_ushort)),
        ("recv_flag", C.c_void_p), ("send_flag", C.c_void_p),
        ("recv_stage", C.c_int), ("send_stage", C.c_int), ("comm_stage", C.c_int),
    ]


class MeshSpec(C.Structure):
    _fields_ = [
        ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
        ("px", C.c_int), ("py", C.c_int), ("pz", C.c_int),
        ("order", C.c_int), ("brick", C.c_int), ("hexcut", C.c_int), ("allow_big", C.c_int),
        ("jitter", C.c_double), ("seed", C.c_ulonglong),
    ]


class MeshDomain(C.Structure):
    _fields_ = [
        ("nfaces", C.c_int), ("nown", C.c_int), ("nall", C.c_int), ("nadd", C.c_int),
        ("ndomains", C.c_int), ("ncommdomains", C.c_int),
        ("fpoint", c_int_p), ("fnormal", c_dbl_p), ("pvolume", c_dbl_p),
        ("commpartner", c_int_p), ("sendcount", c_int_p), ("recvcount", c_int_p),
        ("addpoint_owner", c_int_p), ("addpoint_idx", c_int_p),
        ("global_id", C.POINTER(C.c_longlong)),
    ]


class Stats(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in (
        "nfaces", "nown", "nall", "rows", "ntiles", "nboundary_tiles", "tile_faces", "halo_refs",
        "blob_bytes", "send_rows_local", "send_rows_remote", "alg_bytes", "h2d_bytes", "d2h_bytes", "launches",
        "lds_wavefronts_min", "lds_wavefronts_est")] + [
        ("last_kernel_ms", C.c_double),
        ("nprocs", C.c_int), ("proc_rank", C.c_int), ("ndomains_hosted", C.c_int),
        ("tile_points", C.c_int), ("smem_bytes", C.c_int), ("flux_smem_bytes", C.c_int),
        ("flux_alg_bytes", C.c_longlong), ("last_flux_ms", C.c_double), ("flux_blob_bytes", C.c_longlong),
        ("halo_pack_bytes", C.c_longlong), ("device_bytes", C.c_longlong),
        ("transport", C.c_int), ("ipc_ready", C.c_int), ("direct_ready", C.c_int), ("loopback", C.c_int),
    ]


TRANSPORTS = {0: "none", 1: "on-device copies", 2: "nccl send/recv", 3: "cuda-ipc put+notify", 4: "direct stores into peer memory (cuda-ipc mapping, nvlink)"}


class ScheduleView(C.Structure):
    _fields_ = [
        ("ntiles", C.c_int), ("nboundary_tiles", C.c_int), ("nrows", C.c_int),
        ("row_of_point", c_int_p), ("tile_row0", c_int_p), ("tile_npts", c_int_p),
        ("tile_nfaces", c_int_p), ("tile_nhalo", c_int_p), ("tile_is_boundary", c_int_p),
    ]


INT_EXCHANGE_FN = C.CFUNCTYPE(None, C.c_int, c_int_p, C.POINTER(c_int_p), c_int_p, C.POINTER(c_int_p), c_int_p)

VARIANTS = {
    "comm_free": 0, "mpi_bulk_sync": 1, "mpi_early_recv": 2, "mpi_async": 3,
    "gaspi_bulk_sync": 4, "gaspi_async": 5,
}

# every symbol include/cfdp_b200.h declares (tests check that the library exports all of them)
EXPORTED = [
    "init_communication", "read_communication_data", "compute_communication_tables",
    "free_communication_ressources", "read_solver_data", "init_solver_data", "init_threads",
    "compute_gradients_gg_comm_free", "compute_gradients_gg_mpi_bulk_sync",
    "compute_gradients_gg_mpi_early_recv", "compute_gradients_gg_mpi_async",
    "compute_gradients_gg_gaspi_bulk_sync", "compute_gradients_gg_gaspi_async",
    "compute_gradients_gg_mpifence_bulk_sync", "compute_gradients_gg_mpifence_async",
    "compute_gradients_gg_mpipscw_bulk_sync", "compute_gradients_gg_mpipscw_async",
    "exchange_dbl_mpi_post_recv", "compute_psd_flux", "cfdp_grad_to_device", "cfdp_flux_to_host", "cfdp_set_flux", "cfdp_flux_iterate", "cfdp_refresh_var", "cfdp_set_var_refresh",
    "get_nc_val", "get_nc_int", "get_nc_double",
    "cfdp_nc_open", "cfdp_nc_close", "cfdp_nc_strerror", "cfdp_nc_inq_dimid", "cfdp_nc_inq_dimlen",
    "cfdp_nc_inq_varid", "cfdp_nc_get_var_int", "cfdp_nc_get_var_double",
    "cfdp_configure", "cfdp_init_communication_domain", "cfdp_nccl_get_unique_id", "cfdp_nccl_init",
    "cfdp_commit", "cfdp_plan", "cfdp_set_int_exchange", "cfdp_get_peer_plan", "cfdp_get_exchange_entry", "cfdp_get_tile_exports", "cfdp_get_row_owner", "cfdp_var_to_device", "cfdp_grad_to_host", "cfdp_set_resident", "cfdp_set_exact", "cfdp_set_kernel", "cfdp_get_phase_profile",
    "cfdp_iterate", "cfdp_step_e2e", "cfdp_device_synchronize", "cfdp_finalize", "cfdp_get_stats",
    "cfdp_get_schedule", "cfdp_get_tile", "cfdp_get_tile_blob", "cfdp_get_pack_list", "cfdp_get_unpack_list", "cfdp_get_sendbuf",
    "cfdp_mesh_num_domains", "cfdp_mesh_count_faces_global", "cfdp_mesh_gen_domain",
    "cfdp_mesh_free_domain", "cfdp_mesh_fill_var", "cfdp_mesh_var_value", "cfdp_attach_mesh", "cfdp_attach_mesh_take",
]

_lib = None


def load() -> C.CDLL:
    """Load the shared library (fails loudly when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `make lib` (or __graft_entry__.build()). "
            "There is no CPU fallback for the gradient/halo path.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
    sd_p, cd_p = C.POINTER(SolverData), C.POINTER(CommData)
    P = C.POINTER

    def sig(name, res, *args):
        f = getattr(lib, name)
        f.restype, f.argtypes = res, list(args)

    sig("init_communication", None, C.c_int, P(C.c_char_p), cd_p)
    sig("read_communication_data", None, C.c_int, cd_p)
    sig("compute_communication_tables", None, cd_p)
    sig("free_communication_ressources", None, cd_p)
    sig("read_solver_data", None, C.c_int, sd_p)
    sig("init_solver_data", None, sd_p, C.c_int)
    sig("init_threads", None, cd_p, sd_p, C.c_int)
    for v in ("comm_free", "mpi_bulk_sync", "mpi_early_recv", "mpi_async", "gaspi_bulk_sync", "gaspi_async",
              "mpifence_bulk_sync", "mpifence_async", "mpipscw_bulk_sync", "mpipscw_async"):
        sig("compute_gradients_gg_" + v, None, cd_p, sd_p, C.c_int)
    sig("exchange_dbl_mpi_post_recv", None, cd_p, C.c_int)
    sig("compute_psd_flux", None, sd_p)
    sig("cfdp_grad_to_device", None, sd_p)
    sig("cfdp_flux_to_host", None, sd_p)
    sig("cfdp_set_flux", None, C.c_int)
    sig("cfdp_flux_iterate", C.c_double, C.c_int)
    sig("cfdp_refresh_var", C.c_double, C.c_int)
    sig("cfdp_set_var_refresh", None, C.c_int)
    sig("get_nc_val", C.c_int, C.c_int, C.c_char_p)
    sig("get_nc_int", None, C.c_int, C.c_char_p, c_int_p)
    sig("get_nc_double", None, C.c_int, C.c_char_p, c_dbl_p)
    sig("cfdp_nc_open", C.c_int, C.c_char_p, C.c_int, c_int_p)
    sig("cfdp_nc_close", C.c_int, C.c_int)
    sig("cfdp_nc_strerror", C.c_char_p, C.c_int)
    sig("cfdp_nc_inq_dimid", C.c_int, C.c_int, C.c_char_p, c_int_p)
    sig("cfdp_nc_inq_dimlen", C.c_int, C.c_int, C.c_int, P(C.c_size_t))
    sig("cfdp_nc_inq_varid", C.c_int, C.c_int, C.c_char_p, c_int_p)
    sig("cfdp_nc_get_var_int", C.c_int, C.c_int, C.c_int, c_int_p)
    sig("cfdp_nc_get_var_double", C.c_int, C.c_int, C.c_int, c_dbl_p)
    sig("cfdp_configure", C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
    sig("cfdp_init_communication_domain", None, cd_p, C.c_int)
    sig("cfdp_nccl_get_unique_id", C.c_int, C.c_void_p)
    sig("cfdp_nccl_init", C.c_int, C.c_void_p)
    sig("cfdp_commit", None)
    sig("cfdp_plan", None)
    sig("cfdp_set_int_exchange", None, INT_EXCHANGE_FN)
    sig("cfdp_get_peer_plan", C.c_int, C.c_int, c_int_p, P(C.c_longlong), P(C.c_longlong))
    sig("cfdp_get_exchange_entry", C.c_int, C.c_int, C.c_longlong, c_int_p, c_int_p)
    sig("cfdp_get_tile_exports", C.c_int, C.c_int, C.c_int, P(C.c_uint), P(C.c_uint), c_int_p)
    sig("cfdp_get_row_owner", C.c_int, C.c_longlong, c_int_p, c_int_p)
    sig("cfdp_var_to_device", None, sd_p)
    sig("cfdp_grad_to_host", None, sd_p)
    sig("cfdp_set_resident", None, C.c_int)
    sig("cfdp_set_exact", None, C.c_int)
    sig("cfdp_set_kernel", C.c_int, C.c_int, C.c_int, C.c_int)
    sig("cfdp_get_phase_profile", C.c_int, P(C.c_ulonglong), C.c_int)
    sig("cfdp_iterate", C.c_double, C.c_int, C.c_int, C.c_int)
    sig("cfdp_step_e2e", C.c_double, C.c_int)
    sig("cfdp_device_synchronize", None)
    sig("cfdp_finalize", None)
    sig("cfdp_get_stats", None, P(Stats))
    sig("cfdp_get_schedule", C.c_int, sd_p, P(ScheduleView))
    sig("cfdp_get_tile", C.c_int, sd_p, C.c_int, c_int_p, c_int_p)
    sig("cfdp_get_tile_blob", C.c_longlong, sd_p, C.c_int, C.c_int, P(C.c_uint), P(C.c_ubyte), C.c_longlong)
    sig("cfdp_get_pack_list", C.c_int, cd_p, C.c_int, c_int_p)
    sig("cfdp_get_unpack_list", C.c_int, cd_p, C.c_int, c_int_p)
    sig("cfdp_get_sendbuf", C.c_int, cd_p, C.c_int, c_dbl_p)
    sig("cfdp_mesh_num_domains", C.c_int, P(MeshSpec))
    sig("cfdp_mesh_count_faces_global", C.c_longlong, P(MeshSpec))
    sig("cfdp_mesh_gen_domain", C.c_int, P(MeshSpec), C.c_int, P(MeshDomain))
    sig("cfdp_mesh_free_domain", None, P(MeshDomain))
    sig("cfdp_mesh_fill_var", None, P(MeshDomain), C.c_ulonglong, c_dbl_p)
    sig("cfdp_mesh_var_value", C.c_double, C.c_ulonglong, C.c_longlong, C.c_int)
    sig("cfdp_attach_mesh", None, P(MeshDomain), cd_p, sd_p)
    sig("cfdp_attach_mesh_take", None, P(MeshDomain), cd_p, sd_p)
    _lib = lib
    return lib
