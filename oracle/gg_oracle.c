/*
 * oracle/gg_oracle.c -- TEST INFRASTRUCTURE, not product code.  The checker, never the
 * thing measured or shipped: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.
 *
 * Plain-C restatement of the reference's Green-Gauss gradient path for ONE mesh domain:
 *   - face inclusion and write rules ..... src/rangelist.c:513-523, :567-608, :719-736
 *   - face order of a single-thread run .. src/rangelist.c:567-608 (ttype), src/util.c:113-136
 *     (merge sort by ttype, p1, p0), colours are consecutive chunks (rangelist.c:654-704)
 *   - arithmetic .......................... src/gradients.c:54-145 (zero at first touch,
 *     val = 0.5*(var[p0]+var[p1]), += / -= n*val, multiply by 1/pvolume at last touch)
 * Pinned against the unmodified reference (oracle/_ref/ref_harness, OMP_NUM_THREADS=1):
 * bit-identical grad on every test mesh (tests/test_oracle_vs_reference.py); the reference
 * itself ships no golden vectors.
 * oracle_psd_flux restates the step that consumes the exchanged gradients, src/flux.c:111-201
 * (row f3 of SURVEY 8f), again in the face order and with the colour face types of a
 * single-thread run; pinned the same way (ref_harness with REF_WITH_FLUX=1).
 * Compiled with -ffp-contract=off: the reference build has no FMA (x86-64 baseline).
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#define NGRAD 7

typedef struct { int tt, p1, p0, f; } fkey;

static int cmp_key(const void *a, const void *b)
{
  const fkey *x = (const fkey *)a, *y = (const fkey *)b;
  if (x->tt != y->tt) return x->tt < y->tt ? -1 : 1;
  if (x->p1 != y->p1) return x->p1 < y->p1 ? -1 : 1;
  if (x->p0 != y->p0) return x->p0 < y->p0 ? -1 : 1;
  return x->f < y->f ? -1 : (x->f > y->f); /* duplicate faces: not present in any test mesh */
}

/* order: 0 = faces in file order, 1 = the reference's single-thread order.
 * is_send[p] != 0 for own points listed in some cd->sendindex[k] (htype 2, rangelist.c:118-148);
 * may be NULL (no inner halo).  grad rows of points no face writes are left untouched,
 * like the reference.  Returns the number of faces computed (>= 1 own endpoint). */
long oracle_gradients(int nfaces, int nown, int nall, const int *fpoint, const double *fnormal,
                      const double *pvolume, const double *var, double *grad,
                      const unsigned char *is_send, int order)
{
  fkey *keys = (fkey *)malloc((size_t)(nfaces > 0 ? nfaces : 1) * sizeof(fkey));
  unsigned char *touched = (unsigned char *)calloc((size_t)(nall > 0 ? nall : 1), 1);
  long nf = 0;
  for (int f = 0; f < nfaces; f++) {
    const int p0 = fpoint[2 * f], p1 = fpoint[2 * f + 1];
    const int h0 = p0 >= nown ? 3 : (is_send && is_send[p0] ? 2 : 1);
    const int h1 = p1 >= nown ? 3 : (is_send && is_send[p1] ? 2 : 1);
    if (h0 == 3 && h1 == 3) continue;                       /* rangelist.c:518-519 */
    int tt = (h0 == 2 || h1 == 2) ? 0 : 3;                  /* rangelist.c:567-608, tid == pid everywhere */
    tt += (h0 == 3) ? 0 : (h1 == 3) ? 1 : 2;
    keys[nf].tt = order ? tt : 0; keys[nf].p1 = order ? p1 : 0; keys[nf].p0 = order ? p0 : 0; keys[nf].f = f;
    nf++;
  }
  if (order) qsort(keys, (size_t)nf, sizeof(fkey), cmp_key);
  for (long i = 0; i < nf; i++) {
    const int f = keys[i].f;
    const int p0 = fpoint[2 * f], p1 = fpoint[2 * f + 1];
    const double anx = fnormal[3 * f], any = fnormal[3 * f + 1], anz = fnormal[3 * f + 2];
    const int w0 = p0 < nown, w1 = p1 < nown;
    if (w0 && !touched[p0]) { touched[p0] = 1; memset(&grad[21 * (size_t)p0], 0, 21 * sizeof(double)); } /* gradients.c:54-63 */
    if (w1 && !touched[p1]) { touched[p1] = 1; memset(&grad[21 * (size_t)p1], 0, 21 * sizeof(double)); }
    for (int eq = 0; eq < NGRAD; eq++) {
      const double val = 0.5 * (var[NGRAD * (size_t)p0 + eq] + var[NGRAD * (size_t)p1 + eq]);
      const double vx = anx * val, vy = any * val, vz = anz * val;
      if (w0) { double *g = &grad[21 * (size_t)p0 + 3 * eq]; g[0] += vx; g[1] += vy; g[2] += vz; }
      if (w1) { double *g = &grad[21 * (size_t)p1 + 3 * eq]; g[0] -= vx; g[1] -= vy; g[2] -= vz; }
    }
  }
  for (int p = 0; p < nown; p++) {                           /* gradients.c:135-145 */
    if (!touched[p]) continue;
    const double tmp = 1 / pvolume[p];
    for (int c = 0; c < 21; c++) grad[21 * (size_t)p + c] *= tmp;
  }
  free(keys); free(touched);
  return nf;
}

/* The pseudo flux (flux.c:111-201) of one domain from grad[nall][7][3] (ghost rows already exchanged).
 * A colour of a single-thread run has face type 1 (p0 is a ghost), 2 (p1 is a ghost) or 3 (both own),
 * rangelist.c:719-736.  flux.c:178-189 tests the type differently from gradients.c: `ftype != 3` adds to p0,
 * `ftype != 2` subtracts from p1 -- so a face between two own points only updates p1, and a face whose p0 is
 * a ghost also "updates" that ghost row.  This is what the reference computes; it is restated, not repaired.
 * Only own rows of psd_flux are defined (ghost rows accumulate without ever being zeroed in the reference):
 * they are the only rows written here.  Own points no face touches keep their value. */
long oracle_psd_flux(int nfaces, int nown, int nall, const int *fpoint, const double *fnormal,
                     const double *grad, double *psd_flux, const unsigned char *is_send, int order)
{
  fkey *keys = (fkey *)malloc((size_t)(nfaces > 0 ? nfaces : 1) * sizeof(fkey));
  unsigned char *touched = (unsigned char *)calloc((size_t)(nall > 0 ? nall : 1), 1);
  long nf = 0;
  for (int f = 0; f < nfaces; f++) {
    const int p0 = fpoint[2 * f], p1 = fpoint[2 * f + 1];
    const int h0 = p0 >= nown ? 3 : (is_send && is_send[p0] ? 2 : 1);
    const int h1 = p1 >= nown ? 3 : (is_send && is_send[p1] ? 2 : 1);
    if (h0 == 3 && h1 == 3) continue;
    int tt = (h0 == 2 || h1 == 2) ? 0 : 3;
    tt += (h0 == 3) ? 0 : (h1 == 3) ? 1 : 2;
    keys[nf].tt = order ? tt : 0; keys[nf].p1 = order ? p1 : 0; keys[nf].p0 = order ? p0 : 0; keys[nf].f = f;
    nf++;
  }
  if (order) qsort(keys, (size_t)nf, sizeof(fkey), cmp_key);
  const double mue_eff = 1.0;
  for (long i = 0; i < nf; i++) {
    const int f = keys[i].f;
    const int p0 = fpoint[2 * f], p1 = fpoint[2 * f + 1];
    const double nx = fnormal[3 * f], ny = fnormal[3 * f + 1], nz = fnormal[3 * f + 2];
    const int ftype = p0 >= nown ? 1 : p1 >= nown ? 2 : 3;
    const double *g0 = &grad[21 * (size_t)p0], *g1 = &grad[21 * (size_t)p1];
    /* the points the colour zeroes at first touch are the ones gradients.c writes (flux.c:128-134) */
    if (p0 < nown && !touched[p0]) { touched[p0] = 1; psd_flux[3 * (size_t)p0] = psd_flux[3 * (size_t)p0 + 1] = psd_flux[3 * (size_t)p0 + 2] = 0.0; }
    if (p1 < nown && !touched[p1]) { touched[p1] = 1; psd_flux[3 * (size_t)p1] = psd_flux[3 * (size_t)p1 + 1] = psd_flux[3 * (size_t)p1 + 2] = 0.0; }
    const double dvx_dx = 0.5 * (g0[0] + g1[0]), dvx_dy = 0.5 * (g0[1] + g1[1]), dvx_dz = 0.5 * (g0[2] + g1[2]);
    const double dvy_dx = 0.5 * (g0[3] + g1[3]), dvy_dy = 0.5 * (g0[4] + g1[4]), dvy_dz = 0.5 * (g0[5] + g1[5]);
    const double dvz_dx = 0.5 * (g0[6] + g1[6]), dvz_dy = 0.5 * (g0[7] + g1[7]), dvz_dz = 0.5 * (g0[8] + g1[8]);
    const double lambda = -2.0 / 3.0 * mue_eff;
    const double sts_xx = lambda * (dvy_dy + dvz_dz - 2.0 * dvx_dx);
    const double sts_yy = lambda * (dvx_dx + dvz_dz - 2.0 * dvy_dy);
    const double sts_zz = lambda * (dvx_dx + dvy_dy - 2.0 * dvz_dz);
    const double sts_xy = mue_eff * (dvx_dy + dvy_dx);
    const double sts_xz = mue_eff * (dvx_dz + dvz_dx);
    const double sts_yz = mue_eff * (dvy_dz + dvz_dy);
    const double fx = -(sts_xx * nx + sts_xy * ny + sts_xz * nz);
    const double fy = -(sts_xy * nx + sts_yy * ny + sts_yz * nz);
    const double fz = -(sts_xz * nx + sts_yz * ny + sts_zz * nz);
    if (ftype != 3 && p0 < nown) { double *q = &psd_flux[3 * (size_t)p0]; q[0] += fx; q[1] += fy; q[2] += fz; }
    if (ftype != 2 && p1 < nown) { double *q = &psd_flux[3 * (size_t)p1]; q[0] -= fx; q[1] -= fy; q[2] -= fz; }
  }
  free(keys); free(touched);
  return nf;
}

/* per-point error scale S_p = (sum_f |n_f|_1 * max_eq|val_f|) / vol_p for the summation-order
 * tolerance of SURVEY 8(c): |a-b| <= 1e-12*|b| + 64*eps*S_p */
void oracle_error_scale(int nfaces, int nown, const int *fpoint, const double *fnormal,
                        const double *pvolume, const double *var, double *scale /* [nown] */)
{
  for (int p = 0; p < nown; p++) scale[p] = 0.0;
  for (int f = 0; f < nfaces; f++) {
    const int p0 = fpoint[2 * f], p1 = fpoint[2 * f + 1];
    double n1 = 0, vmax = 0;
    for (int c = 0; c < 3; c++) { double a = fnormal[3 * f + c]; n1 += a < 0 ? -a : a; }
    for (int eq = 0; eq < NGRAD; eq++) {
      double v = 0.5 * (var[NGRAD * (size_t)p0 + eq] + var[NGRAD * (size_t)p1 + eq]);
      v = v < 0 ? -v : v; if (v > vmax) vmax = v;
    }
    if (p0 < nown) scale[p0] += n1 * vmax;
    if (p1 < nown) scale[p1] += n1 * vmax;
  }
  for (int p = 0; p < nown; p++) { double v = pvolume[p]; scale[p] /= (v < 0 ? -v : v); }
}

/* threads.c:791-813 (pack) and :816-839 (unpack): raw row copies in list order */
void oracle_pack(const double *data, int dim2, const int *sendindex, int count, double *sbuf)
{
  for (int j = 0; j < count; j++) memcpy(&sbuf[(size_t)dim2 * j], &data[(size_t)dim2 * sendindex[j]], (size_t)dim2 * sizeof(double));
}
void oracle_unpack(double *data, int dim2, const int *recvindex, int count, const double *rbuf)
{
  for (int j = 0; j < count; j++) memcpy(&data[(size_t)dim2 * recvindex[j]], &rbuf[(size_t)dim2 * j], (size_t)dim2 * sizeof(double));
}
