#!/usr/bin/env python
"""Kernel-only sweep (comm_free iterations) over schedule / kernel configurations.
usage: python tools/kbench.py [--mpoints 16] [--f6likeN] [--iters K] mesh ...
  mesh = tile:order:fma[:stage budget[:CTAs per SM]]/kcfg,kcfg,...   kcfg = version.chunk.persistent[.variant]   e.g. 256:lex:0/2.8.0,3.8.0,2.1.296
The mesh and its schedule are built once per `mesh`; the kernel configurations are switched on the live session."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfd_proxy_b200.mesh as M
from cfd_proxy_b200.driver import Session
from bench import lattice_for, measured_peak

def main():
    args = sys.argv[1:]
    mp, f6, iters, rounds, warm_s = 16.0, None, 10, 3, 3.0
    while args and args[0].startswith("--"):
        if args[0] == "--mpoints": mp = float(args[1]); args = args[2:]
        elif args[0] == "--iters": iters = int(args[1]); args = args[2:]
        elif args[0] == "--rounds": rounds = int(args[1]); args = args[2:]
        elif args[0] == "--warm": warm_s = float(args[1]); args = args[2:]
        elif args[0].startswith("--f6like"): f6 = int(args[0][8:]); args = args[1:]
        else: raise SystemExit("unknown option " + args[0])
    n = lattice_for(mp)
    peak, _ = measured_peak()
    for mesh in args:
        head, kcfgs = mesh.split("/")
        head, _, envs = head.partition("@")          # 256:lex:0@CFDP_PLACE_REFINE=2,CFDP_X=1 : environment of this mesh's schedule build
        extra = dict(kv.split("=") for kv in envs.split(",") if kv)
        os.environ.update(extra)
        tile, order, fma, budget, ctas = (head.split(":") + ["0", "2"])[:5]
        os.environ["CFDP_CTAS"] = ctas
        if int(budget):
            os.environ["CFDP_STAGE_BUDGET"] = budget
        else:
            os.environ.pop("CFDP_STAGE_BUDGET", None)
        t0 = time.time()
        with Session(f6 or 8, device=0, tile_points=int(tile), tile_order=1 if order == "brickid" else 0) as S:
            if f6:   # BASELINE configs[1]/[2]: F6-like stand-in (hybrid hex/tet dual, ~2 M points at level 1), all domains on one GPU
                spec = M.f6like_spec(f6, lvl=1, order=order)
            else:
                spec = M.make_spec(n, (2, 2, 2), order="brick" if order.startswith("brick") else order, brick=8, jitter=0.1, allow_big=True)
            S.load_spec(spec); S.setup()
            S.lib.cfdp_set_exact(0 if fma == "1" else 1)
            setup_s = time.time() - t0
            # steady state first: under the 1 kW power cap the SM clock sags for the first seconds of load
            t_w = time.time()
            while time.time() - t_w < warm_s:
                S.iterate("comm_free", 20)
            res = {kc: dict(k=[], a=[], b=[]) for kc in kcfgs.split(",")}
            for rnd in range(rounds):     # configurations interleaved: drift hits all of them alike
                for kc in res:
                    ver, chunk, pers, order = (int(x) for x in (kc.split(".") + ["0"])[:4])
                    os.environ["CFDP_VARIANT"] = str(order)
                    assert S.lib.cfdp_set_kernel(ver, chunk, pers) == ver
                    S.iterate("comm_free", 2)
                    res[kc]["k"].append(S.iterate("comm_free", iters) / iters)
                    res[kc]["a"].append(S.iterate("mpi_async", iters) / iters)
                    res[kc]["b"].append(S.iterate("mpi_bulk_sync", iters) / iters)
            st = S.stats()
            for kc, r in res.items():
                ms, ms_a, ms_b = sorted(r["k"])[len(r["k"]) // 2], sorted(r["a"])[len(r["a"]) // 2], sorted(r["b"])[len(r["b"]) // 2]
                prof = None
                if os.environ.get("CFDP_PHASE_PROF"):
                    ver, chunk, pers, order = (int(x) for x in (kc.split(".") + ["0"])[:4])
                    os.environ["CFDP_VARIANT"] = str(order)
                    S.lib.cfdp_set_kernel(ver, chunk, pers)
                    buf = (C.c_ulonglong * 8)()
                    S.lib.cfdp_get_phase_profile(buf, 1)
                    S.iterate(os.environ.get("KBENCH_PROF_VARIANT", "comm_free"), 5)
                    S.lib.cfdp_get_phase_profile(buf, 1)
                    nt = max(buf[4], 1)
                    prof = dict(wait=round(buf[0] / nt), walk=round(buf[1] / nt), rest=round(buf[2] / nt), early=round(buf[5] / nt), stage_store=round(buf[6] / nt), wait_read=round(buf[7] / nt))
                print(json.dumps(dict(mesh=head, env=extra, kernel=kc, kernel_ms=round(ms, 4), spread=[round(min(r["k"]), 4), round(max(r["k"]), 4)], gfaces=round(st.nfaces / ms / 1e6, 2), frac=round(st.alg_bytes / ms / 1e6 / peak, 4),
                                      async_ms=round(ms_a, 4), bulk_ms=round(ms_b, 4), boundary_tiles=st.nboundary_tiles, smem=st.smem_bytes, tiles=st.ntiles, dup=round(st.tile_faces / st.nfaces, 3),
                                      blob_B_per_face=round(st.blob_bytes / st.nfaces, 2), setup_s=round(setup_s, 1), phase=prof)), flush=True)
            for k in extra: os.environ.pop(k, None)
            if os.environ.get("KBENCH_FLUX"):   # KBENCH_FLUX=0,1,2,3: pseudo-flux kernel variants (CFDP_FLUX_VARIANT), interleaved
                fv = os.environ["KBENCH_FLUX"].split(",")
                S.flux_iterate(3)
                fr = {v: [] for v in fv}
                for rnd in range(rounds):
                    for v in fv:
                        os.environ["CFDP_FLUX_VARIANT"] = v
                        S.flux_iterate(2)
                        fr[v].append(S.flux_iterate(iters) / iters)
                st = S.stats()
                for v in fv:
                    ms_f = sorted(fr[v])[len(fr[v]) // 2]
                    print(json.dumps(dict(mesh=head, flux_variant=v, flux_ms=round(ms_f, 4), spread=[round(min(fr[v]), 4), round(max(fr[v]), 4)],
                                          flux_frac=round(st.flux_alg_bytes / ms_f / 1e6 / peak, 4), flux_smem=st.flux_smem_bytes,
                                          flux_blob_B_per_face=round(st.flux_blob_bytes / st.nfaces, 2))), flush=True)

if __name__ == "__main__":
    main()
