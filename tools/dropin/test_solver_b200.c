/*
 * test_solver_b200.c -- the benchmark loop the reference's main() calls (reference src/solver.c:35-314),
 * re-written for the B200 library: same protocol (N_MEDIAN = 25 repeats of NITER iterations per variant,
 * median reported), same "*** SETUP / *** TIMINGS" print format, every iteration = gradient + halo exchange
 * followed by the pseudo flux (solver.c:45-55).  Compiled against the REFERENCE headers: the structs
 * it passes to libcfdp_b200.so are the reference's own comm_data / solver_data.
 * Built by `make -C oracle dropin` together with the unmodified src/hybrid.f6.c.
 */
#include <stdio.h>
#include <stdlib.h>
#include <sys/time.h>
#include "comm_data.h"
#include "solver_data.h"
#include "gradients.h"
#include "flux.h"
#include "exchange_data_mpi.h"
#include "solver.h"

void cfdp_set_resident(int resident);
void cfdp_device_synchronize(void);
void cfdp_var_to_device(solver_data *sd);
void cfdp_grad_to_host(solver_data *sd);
void cfdp_flux_to_host(solver_data *sd);

#define N_MEDIAN 25
#define N_SOLVER 10

static double now_s(void)
{
  struct timeval tp;
  gettimeofday(&tp, NULL);
  return (double)tp.tv_sec + (double)tp.tv_usec * 1.e-6;
}
static int cmp_double(const void *a, const void *b)
{
  const double x = *(const double *)a, y = *(const double *)b;
  return x < y ? -1 : x > y;
}

typedef void (*grad_fn)(comm_data *, solver_data *, int);

void test_solver(comm_data *cd, solver_data *sd, int NTHREADS)
{
  static const grad_fn fn[N_SOLVER] = {
    compute_gradients_gg_comm_free, compute_gradients_gg_mpi_bulk_sync, compute_gradients_gg_mpi_early_recv,
    compute_gradients_gg_mpi_async, compute_gradients_gg_gaspi_bulk_sync, compute_gradients_gg_gaspi_async,
    compute_gradients_gg_mpifence_bulk_sync, compute_gradients_gg_mpifence_async,
    compute_gradients_gg_mpipscw_bulk_sync, compute_gradients_gg_mpipscw_async };
  static double median[N_SOLVER][N_MEDIAN];
  int k, v, i;
  cfdp_var_to_device(sd);     /* var/grad stay on the device during the timed loops */
  cfdp_set_resident(1);
  for (k = 0; k < N_MEDIAN; ++k) {
    for (v = 0; v < N_SOLVER; ++v) {
      if (v > 0 && cd->ndomains == 1) { median[v][k] = 0.0; continue; }   /* solver.c:61-64 */
      cfdp_device_synchronize();
      double t = -now_s();
      if (v == 2 || v == 3) exchange_dbl_mpi_post_recv(cd, NGRAD * 3);    /* solver.c:87,106 */
      for (i = 0; i < sd->niter; ++i) {
        fn[v](cd, sd, i == sd->niter - 1);
        compute_psd_flux(sd);                                             /* solver.c:52 */
      }
      cfdp_device_synchronize();
      t += now_s();
      median[v][k] = t;
    }
    if (cd->iProc == 0) { printf("."); fflush(stdout); }
  }
  cfdp_set_resident(0);
  cfdp_grad_to_host(sd);
  cfdp_flux_to_host(sd);
  if (cd->iProc == 0) {
    printf("\n*** COMPILE FLAGS\n -DCFDP_B200 (sm_100a CUDA kernels, NCCL halo exchange)");
    printf("\n\n*** SETUP\n");
    printf("                                 nProc: %d\n", cd->nProc);
    printf("                              NTHREADS: %d\n", NTHREADS);
    printf("                                 NITER: %d\n", sd->niter);
    printf("                              N_MEDIAN: %d\n", N_MEDIAN);
    for (v = 0; v < N_SOLVER; ++v) qsort(median[v], N_MEDIAN, sizeof(double), cmp_double);
    printf("\n*** TIMINGS\n");
    printf("                             comm_free: %10.6f\n", median[0][N_MEDIAN / 2]);
    printf("            exchange_dbl_mpi_bulk_sync: %10.6f\n", median[1][N_MEDIAN / 2]);
    printf("           exchange_dbl_mpi_early_recv: %10.6f\n", median[2][N_MEDIAN / 2]);
    printf("                exchange_dbl_mpi_async: %10.6f\n", median[3][N_MEDIAN / 2]);
    printf("          exchange_dbl_gaspi_bulk_sync: %10.6f\n", median[4][N_MEDIAN / 2]);
    printf("              exchange_dbl_gaspi_async: %10.6f\n", median[5][N_MEDIAN / 2]);
    printf("      exchange_dbl_mpi_fence_bulk_sync: %10.6f\n", median[6][N_MEDIAN / 2]);
    printf("          exchange_dbl_mpi_fence_async: %10.6f\n", median[7][N_MEDIAN / 2]);
    printf("       exchange_dbl_mpi_pscw_bulk_sync: %10.6f\n", median[8][N_MEDIAN / 2]);
    printf("           exchange_dbl_mpi_pscw_async: %10.6f\n", median[9][N_MEDIAN / 2]);
    /* one line the reference does not print: a checksum of grad so that runs can be compared */
    double s = 0.0;
    for (i = 0; i < sd->nownpoints; i++) for (v = 0; v < NGRAD; v++) for (k = 0; k < 3; k++) s += sd->grad[i][v][k];
    printf("\n*** CHECKSUM (rank 0 own rows): %.17g\n", s);
    s = 0.0;
    for (i = 0; i < sd->nownpoints; i++) for (k = 0; k < NFLUX; k++) s += sd->psd_flux[i][k];
    printf("*** CHECKSUM psd_flux (rank 0 own rows): %.17g\n", s);
  }
}
