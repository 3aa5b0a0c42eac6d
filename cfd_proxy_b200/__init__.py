"""B200-native Green-Gauss gradient + halo exchange: the one hot path of CFD-Proxy.

Product code lives in ``csrc/`` (C host code + sm_100a CUDA kernels behind the C ABI of
``include/cfdp_b200.h``); this package is the thin Python plumbing above it: ctypes binding
(`lib`), synthetic F6-schema meshes and NetCDF-3 files (`mesh`, `netcdf3`) and the
multi-domain / multi-GPU driver used by tests and bench (`driver`).
"""
from . import lib  # noqa: F401

__all__ = ["lib", "mesh", "netcdf3", "driver"]
