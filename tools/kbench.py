#!/usr/bin/env python
"""Kernel-only sweep (comm_free iterations) over schedule / kernel configurations.
usage: python tools/kbench.py [--mpoints 16] cfg ...   cfg = kernel:tile:chunk:fma:order[:stages:split]  e.g. 2:256:16:0:lex:2:2"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfd_proxy_b200.mesh as M
from cfd_proxy_b200.driver import Session
from bench import lattice_for, measured_peak

def main():
    args = sys.argv[1:]
    mp = 16.0
    if args and args[0] == "--mpoints":
        mp = float(args[1]); args = args[2:]
    n = lattice_for(mp)
    peak, _ = measured_peak()
    f6 = None
    if args and args[0].startswith("--f6like"):
        f6 = int(args[0][8:]); args = args[1:]
    for cfg in args:
        parts = cfg.split(":") + ["2", "2"]
        k, tile, chunk, fma, order, stages, split = parts[:7]
        os.environ["CFDP_KERNEL"] = k
        os.environ["CFDP_CHUNK"] = chunk
        os.environ["CFDP_STAGES"] = stages
        os.environ["CFDP_SPLIT"] = split
        t0 = time.time()
        with Session(f6 or 8, device=0, tile_points=int(tile), tile_order=1 if order == "brickid" else 0) as S:
            if f6:   # BASELINE configs[1]/[2]: F6-like stand-in (hybrid hex/tet dual, ~2 M points at level 1), all domains on one GPU
                spec = M.f6like_spec(f6, lvl=1, order=order)
            else:
                spec = M.make_spec(n, (2, 2, 2), order="brick" if order.startswith("brick") else order, brick=8, jitter=0.1, allow_big=True)
            S.load_spec(spec); S.setup()
            S.lib.cfdp_set_exact(0 if fma == "1" else 1)
            S.iterate("comm_free", 3)
            ms = S.iterate("comm_free", 10) / 10
            ms_a = S.iterate("mpi_async", 10) / 10
            S.flux_iterate(3)
            ms_f = S.flux_iterate(10) / 10
            st = S.stats()
            prof = None
            if os.environ.get("CFDP_PHASE_PROF"):
                import ctypes as C
                buf = (C.c_ulonglong * 8)()
                S.lib.cfdp_get_phase_profile.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
                S.lib.cfdp_get_phase_profile(buf, 1)
                S.iterate("comm_free", 5)
                S.lib.cfdp_get_phase_profile(buf, 1)
                nt = max(buf[4], 1)
                prof = dict(wait=round(buf[0] / nt), walk=round(buf[1] / nt), rest=round(buf[2] / nt), stage_S2=round(buf[5] / nt), store_exports_earlyfetch=round(buf[6] / nt), wait_read=round(buf[7] / nt))
                print("  phase cycles per tile (thread 0):", prof, flush=True)
            print(json.dumps(dict(cfg=cfg, kernel_ms=round(ms, 4), gfaces=round(st.nfaces / ms / 1e6, 2), frac=round(st.alg_bytes / ms / 1e6 / peak, 4),
                                  async_ms=round(ms_a, 4), flux_ms=round(ms_f, 4), flux_frac=round(st.flux_alg_bytes / ms_f / 1e6 / peak, 4), flux_smem=st.flux_smem_bytes, flux_blob_B_per_face=round(st.flux_blob_bytes / st.nfaces, 2), smem=st.smem_bytes, tiles=st.ntiles, dup=round(st.tile_faces / st.nfaces, 3),
                                  blob_B_per_face=round(st.blob_bytes / st.nfaces, 2), setup_s=round(time.time() - t0, 1))), flush=True)

if __name__ == "__main__":
    main()
