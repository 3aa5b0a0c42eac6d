#!/usr/bin/env python
"""BASELINE.json configs 1-3 on the F6-like stand-in meshes (the shipped F6 files are not available offline).

  python tools/f6like_configs.py ref [LVL]        config 1: the UNMODIFIED reference (oracle/_ref/hybrid.f6.exe, MPI-only build)
                                                  with 12 ranks over the shm-MPI shim on dualgrid.12 level LVL (default 2), on
                                                  the host cores; prints the reference's own `*** TIMINGS` report (solver.c:246-313)
  python tools/f6like_configs.py gpu NDOMAINS     config 2 (NDOMAINS=12, one GPU) / config 3 (NDOMAINS=24 under torchrun with
                                                  2/4/8 ranks: 12/6/3 domains per GPU): level-1 stand-in (2.07 M points), every
                                                  exchange variant timed device-resident, faces/s over all ranks, one JSON line

The mesh files of `ref` come from oracle/_ref/mesh_tool (standalone writer): the reference processes never map libcfdp_b200.so.
"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref")


def run_ref(lvl):
    base = [160, 120, 108]
    n = [max(4, round(b / 2 ** (lvl - 1))) for b in base]
    hexcut = int(round(0.4 * n[0]))
    tmp = tempfile.mkdtemp(prefix="cfdp_f6like_")
    prefix = os.path.join(tmp, "dualgrid")
    r = subprocess.run([os.path.join(REF, "mesh_tool"), prefix, str(lvl), *map(str, n), "3", "2", "2", "0", "8", str(hexcut), "0.1", "0x5DEECE66D"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit("mesh_tool failed: " + r.stderr[-500:])
    print("# mesh:", r.stdout.strip().splitlines()[-1])
    ncores = os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS=str(max(1, ncores // 12)))
    cmd = [os.path.join(REF, "mpirun_shim"), "-np", "12", os.path.join(REF, "hybrid.f6.exe"), "-lvl", str(lvl), prefix]
    print("# %d host cores, OMP_NUM_THREADS=%s: %s" % (ncores, env["OMP_NUM_THREADS"], " ".join(cmd[:4] + ["hybrid.f6.exe", "-lvl", str(lvl), "dualgrid"])), flush=True)
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=3000)
    sys.stdout.write(r.stdout[-6000:])
    sys.stderr.write(r.stderr[-2000:])
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return r.returncode


def run_gpu(ndomains):
    import torch
    import cfd_proxy_b200.mesh as M
    from cfd_proxy_b200.driver import session_from_env
    from cfd_proxy_b200.lib import TRANSPORTS
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))
    S = session_from_env(ndomains)
    S.load_spec(M.f6like_spec(ndomains, lvl=1))
    S.setup()
    S.lib.cfdp_set_resident(1)
    st = S.stats()
    import torch.distributed as dist

    def allred(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    faces = allred(float(st.nfaces), dist.ReduceOp.SUM)
    points = allred(float(st.nown), dist.ReduceOp.SUM)
    variants = ["comm_free", "mpi_bulk_sync", "mpi_async", "gaspi_bulk_sync", "gaspi_async"]
    for v in variants:                      # steady clocks first
        S.iterate(v, 200)
    out, transports = {}, {}
    niter = 500
    for rnd in range(3):
        for v in variants:
            S.iterate(v, 5)
            S.lib.cfdp_device_synchronize()
            if world > 1:
                dist.barrier()
            ms = allred(S.iterate(v, niter) / niter, dist.ReduceOp.MAX)
            out.setdefault(v, []).append(ms)
            transports[v] = TRANSPORTS[int(S.stats().transport)]
    if rank == 0:
        med = {v: sorted(t)[1] for v, t in out.items()}
        print(json.dumps(dict(config="F6-like stand-in dualgrid.%d level 1 (hybrid hex/tet dual, %.2f M own points, %.2f M faces), %d domains per GPU on %d GPU(s)"
                                     % (ndomains, points / 1e6, faces / 1e6, ndomains // world, world),
                              n_gpus=world, iterations_per_variant=niter, us_per_iteration={v: round(1e3 * m, 2) for v, m in med.items()},
                              gfaces_per_s={v: round(faces / m / 1e6, 2) for v, m in med.items()}, transport=transports,
                              tiles=int(st.ntiles), boundary_tiles=int(st.nboundary_tiles))), flush=True)
    S.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    if len(sys.argv) >= 2 and sys.argv[1] == "ref":
        sys.exit(run_ref(int(sys.argv[2]) if len(sys.argv) > 2 else 2))
    if len(sys.argv) >= 3 and sys.argv[1] == "gpu":
        sys.exit(run_gpu(int(sys.argv[2])))
    raise SystemExit(__doc__)
