/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Drives the UNMODIFIED reference (objects compiled from /root/reference/src by
 * oracle/Makefile into oracle/_ref/) through its own public entry points, the way
 * src/hybrid.f6.c:27-101 and src/solver.c:35-120 do, but
 *   - overwrites sd.var with seeded data from "<meshfile>.var" (the shipped var == 1.0 makes
 *     interior gradients cancel), and pre-fills grad with NaN,
 *   - runs ONE variant for NITER iterations of gradient(+exchange) only; with REF_WITH_FLUX=1 in the
 *     environment every iteration also calls compute_psd_flux like solver.c:45-55 does, and psd_flux is dumped,
 *   - dumps grad, sendindex/recvindex and timing for the parity tests and the CPU baseline.
 *
 *   ref_harness -lvl L PREFIX VARIANT NITER OUTPREFIX [REPEATS]
 *   VARIANT: comm_free | mpi_bulk_sync | mpi_early_recv | mpi_async
 * Run as `mpirun_shim -np <ndomains> ref_harness ...` (one rank per mesh domain).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <omp.h>
#include <mpi.h>
#include <netcdf.h>
#include "comm_data.h"
#include "solver_data.h"
#include "rangelist.h"
#include "gradients.h"
#include "flux.h"
#include "exchange_data_mpi.h"
#include "error_handling.h"
#include "util.h"

typedef void (*grad_fn)(comm_data *, solver_data *, int);

static void dump_index(const char *path, comm_data *cd)
{
  /* binary int32: ncommdomains, then per partner: k, sendcount, recvcount, sendindex[], recvindex[] */
  FILE *f = fopen(path, "wb");
  ASSERT(f != NULL);
  int n = cd->ndomains > 1 ? cd->ncommdomains : 0;
  fwrite(&n, 4, 1, f);
  for (int i = 0; i < n; i++) {
    int k = cd->commpartner[i];
    fwrite(&k, 4, 1, f);
    fwrite(&cd->sendcount[k], 4, 1, f);
    fwrite(&cd->recvcount[k], 4, 1, f);
    fwrite(cd->sendindex[k], 4, (size_t)cd->sendcount[k], f);
    fwrite(cd->recvindex[k], 4, (size_t)cd->recvcount[k], f);
  }
  fclose(f);
}

int main(int argc, char *argv[])
{
  comm_data cd;
  solver_data sd;
  int ncid, retval;
  if (argc < 7 || strcmp(argv[1], "-lvl") != 0) {
    printf("Usage: %s -lvl L PREFIX VARIANT NITER OUTPREFIX [REPEATS]\n", argv[0]);
    exit(EXIT_FAILURE);
  }
  const char *variant = argv[4];
  const int niter = atoi(argv[5]);
  const char *outprefix = argv[6];
  const int repeats = argc > 7 ? atoi(argv[7]) : 1;
  char *env = getenv("OMP_NUM_THREADS");
  ASSERT(env != NULL);
  const int NTHREADS = atoi(env);
  const char *wf = getenv("REF_WITH_FLUX");
  const int with_flux = wf && atoi(wf);

  init_communication(argc, argv, &cd);
  char fname[512];
  snprintf(fname, sizeof fname, "%s_domain_%d_lvl_%d", argv[3], cd.iProc, atoi(argv[2]));
  ASSERT(f_exist(fname));
  if ((retval = nc_open(fname, NC_NOWRITE, &ncid))) ERR(retval);
  read_solver_data(ncid, &sd);
  init_solver_data(&sd, niter);
  read_communication_data(ncid, &cd);
  compute_communication_tables(&cd);

  char vname[600];
  snprintf(vname, sizeof vname, "%s.var", fname);
  FILE *vf = fopen(vname, "rb");
  if (vf) {
    size_t n = (size_t)sd.nallpoints * NGRAD;
    ASSERT(fread(&sd.var[0][0], sizeof(double), n, vf) == n);
    fclose(vf);
  }
  init_threads(&cd, &sd, NTHREADS);
  for (int p = 0; p < sd.nallpoints; p++)
    for (int e = 0; e < NGRAD; e++)
      for (int c = 0; c < 3; c++) sd.grad[p][e][c] = NAN;

  grad_fn fn = NULL;
  int post = 0;
  if (!strcmp(variant, "comm_free")) fn = compute_gradients_gg_comm_free;
  else if (!strcmp(variant, "mpi_bulk_sync")) fn = compute_gradients_gg_mpi_bulk_sync;
  else if (!strcmp(variant, "mpi_early_recv")) { fn = compute_gradients_gg_mpi_early_recv; post = 1; }
  else if (!strcmp(variant, "mpi_async")) { fn = compute_gradients_gg_mpi_async; post = 1; }
  ASSERT(fn != NULL);
  if (cd.ndomains == 1) { fn = compute_gradients_gg_comm_free; post = 0; }

  /* faces this rank computes: at least one own endpoint (rangelist.c:513-523) */
  long nf = 0;
  for (int f = 0; f < sd.nfaces; f++)
    if (sd.fpoint[f][0] < sd.nownpoints || sd.fpoint[f][1] < sd.nownpoints) nf++;

  double best = 1e300, sum = 0;
  for (int r = 0; r < repeats; r++) {
    double t = -now();
    MPI_Barrier(MPI_COMM_WORLD);
    if (post) exchange_dbl_mpi_post_recv(&cd, NGRAD * 3);
#pragma omp parallel default(none) shared(cd, sd, fn) firstprivate(with_flux)
    {
      for (int i = 0; i < sd.niter; ++i) {
        int final = (i == sd.niter - 1) ? 1 : 0;
        fn(&cd, &sd, final);
#pragma omp barrier
        if (with_flux) {
          compute_psd_flux(&sd);
#pragma omp barrier
        }
      }
    }
    MPI_Barrier(MPI_COMM_WORLD);
    t += now();
    if (t < best) best = t;
    sum += t;
  }

  char path[700];
  const char *nd = getenv("REF_NO_DUMP");            /* timing runs on large meshes: no result dumps */
  const int dump = !(nd && atoi(nd));
  if (dump) {
  snprintf(path, sizeof path, "%s_domain_%d.grad", outprefix, cd.iProc);
  FILE *gf = fopen(path, "wb");
  ASSERT(gf != NULL);
  fwrite(&sd.grad[0][0][0], sizeof(double), (size_t)sd.nallpoints * NGRAD * 3, gf);
  fclose(gf);
  }
  if (with_flux && dump) {
    snprintf(path, sizeof path, "%s_domain_%d.flux", outprefix, cd.iProc);
    FILE *ff = fopen(path, "wb");
    ASSERT(ff != NULL);
    fwrite(&sd.psd_flux[0][0], sizeof(double), (size_t)sd.nallpoints * NFLUX, ff);
    fclose(ff);
  }
  snprintf(path, sizeof path, "%s_domain_%d.index", outprefix, cd.iProc);
  if (dump) dump_index(path, &cd);
  snprintf(path, sizeof path, "%s_domain_%d.time", outprefix, cd.iProc);
  FILE *tf = fopen(path, "w");
  ASSERT(tf != NULL);
  fprintf(tf, "{\"variant\": \"%s\", \"rank\": %d, \"nranks\": %d, \"threads\": %d, \"niter\": %d, \"repeats\": %d, "
              "\"with_flux\": %d, \"faces\": %ld, \"best_s\": %.9g, \"mean_s\": %.9g}\n",
          variant, cd.iProc, cd.nProc, NTHREADS, niter, repeats, with_flux, nf, best, sum / repeats);
  fclose(tf);

  free_communication_ressources(&cd);
  if ((retval = nc_close(ncid))) ERR(retval);
  return 0;
}
