/*
 * mesh_gen.c -- synthetic dual-mesh domains in the F6 per-domain schema.
 *
 * The shipped F6 meshes (f6/dualgrid.*.tgz) are not available offline, so every workload of
 * this repository runs on meshes generated here and laid out exactly like the files the
 * reference loader reads (reference: src/solver_data.c:98-144, src/comm_data.c:79-112):
 * own points [0,nown), ghost "addpoints" [nown,nall), faces (p0,p1,normal) with at least
 * one own endpoint, rank-indexed sendcount/recvcount, addpoint_owner/addpoint_idx.
 *
 * Geometry: an nx*ny*nz lattice whose edges follow the Kuhn (Freudenthal) triangulation,
 * i.e. up to 7 edge directions per point (3 axis, 3 face diagonals, 1 body diagonal): the
 * dual of a tetrahedral box mesh, 14 neighbours / ~7 faces per interior point.  A slab
 * x < hexcut keeps only the 3 axis directions (hexahedral dual, ~3 faces per point) which
 * gives an F6-like hybrid mesh with varying point degree.  Domains are the boxes of a
 * px*py*pz block decomposition; each domain is generated independently in closed form, so
 * a rank never materialises the global mesh.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include "cfdp_b200.h"

static const int DIRS[7][3] = { {1,0,0},{0,1,0},{0,0,1},{1,1,0},{1,0,1},{0,1,1},{1,1,1} };
static const double DIRW[7] = { 1.0, 1.0, 1.0, 0.5, 0.5, 0.5, 0.25 };

static inline uint64_t mix64(uint64_t z)
{
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline double u01(uint64_t key) { return (double)(mix64(key) >> 11) * (1.0 / 9007199254740992.0); }

/* seeded value of var[gid][eq]: keyed by the GLOBAL point id so that ghosts agree with owners */
double cfdp_mesh_var_value(unsigned long long seed, long long gid, int eq)
{
  return u01(seed ^ ((uint64_t)gid * 8u + (uint64_t)eq + 0x5DEECE66Dull)) + (double)eq;
}

typedef struct { int lo[3], hi[3]; } box_t;

static void split(int n, int parts, int i, int *lo, int *hi)
{
  *lo = (int)((long long)n * i / parts);
  *hi = (int)((long long)n * (i + 1) / parts);
}

static void domain_box(const cfdp_mesh_spec *s, int rank, box_t *b)
{
  int ix = rank % s->px, iy = (rank / s->px) % s->py, iz = rank / (s->px * s->py);
  split(s->nx, s->px, ix, &b->lo[0], &b->hi[0]);
  split(s->ny, s->py, iy, &b->lo[1], &b->hi[1]);
  split(s->nz, s->pz, iz, &b->lo[2], &b->hi[2]);
}

static int owner_of(const cfdp_mesh_spec *s, int x, int y, int z)
{
  /* inverse of split(): largest i with n*i/parts <= x */
  int ix = (int)(((long long)(x + 1) * s->px - 1) / s->nx);
  int iy = (int)(((long long)(y + 1) * s->py - 1) / s->ny);
  int iz = (int)(((long long)(z + 1) * s->pz - 1) / s->nz);
  return ix + s->px * (iy + s->py * iz);
}

static int edge_exists(const cfdp_mesh_spec *s, int x, int d)
{
  /* diagonal edges are dropped when their lower-x endpoint lies in the hex slab */
  if (d < 3) return 1;
  return x >= s->hexcut;
}

/* position of (x,y,z) in the own numbering of the domain with box b */
static long long own_rank_in_box(const cfdp_mesh_spec *s, const box_t *b, int x, int y, int z)
{
  long long sx = b->hi[0] - b->lo[0], sy = b->hi[1] - b->lo[1];
  long long lx = x - b->lo[0], ly = y - b->lo[1], lz = z - b->lo[2];
  if (s->order != CFDP_ORDER_BRICK || s->brick <= 1) return lx + sx * (ly + sy * lz);
  /* brick-blocked: bricks of brick^3 points (clipped at the box), lexicographic inside a brick */
  long long B = s->brick, sz = b->hi[2] - b->lo[2];
  long long bx = lx / B, by = ly / B, bz = lz / B;
  long long x0 = bx * B, y0 = by * B, z0 = bz * B;
  long long wx = (x0 + B <= sx ? B : sx - x0), wy = (y0 + B <= sy ? B : sy - y0), wz = (z0 + B <= sz ? B : sz - z0);
  /* points before this brick: full z-layers of bricks, full y-rows, then x */
  long long before = z0 * sx * sy + wz * (y0 * sx + wy * x0);
  (void)wz;
  return before + (lx - x0) + wx * ((ly - y0) + wy * (lz - z0));
}

/* seeded permutation for CFDP_ORDER_SHUFFLE (Fisher-Yates over the lexicographic rank) */
static int *shuffle_perm(const cfdp_mesh_spec *s, int rank, long long n)
{
  int *perm = (int *)malloc((size_t)n * sizeof(int));
  if (!perm) return NULL;
  for (long long i = 0; i < n; i++) perm[i] = (int)i;
  uint64_t st = s->seed ^ (0xA24BAED4963EE407ull * (uint64_t)(rank + 1));
  for (long long i = n - 1; i > 0; i--) {
    st = mix64(st);
    long long j = (long long)(st % (uint64_t)(i + 1));
    int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
  }
  return perm;
}

int cfdp_mesh_num_domains(const cfdp_mesh_spec *s) { return s->px * s->py * s->pz; }

long long cfdp_mesh_count_faces_global(const cfdp_mesh_spec *s)
{
  long long n = 0;
  for (int d = 0; d < 7; d++) {
    long long ex = s->nx - DIRS[d][0], ey = s->ny - DIRS[d][1], ez = s->nz - DIRS[d][2];
    if (ex <= 0 || ey <= 0 || ez <= 0) continue;
    if (d >= 3) {
      long long x0 = s->hexcut > 0 ? s->hexcut : 0;
      ex = ex - x0; if (ex < 0) ex = 0;
    }
    n += ex * ey * ez;
  }
  return n;
}

void cfdp_mesh_free_domain(cfdp_mesh_domain *m)
{
  if (!m) return;
  free(m->fpoint); free(m->fnormal); free(m->pvolume); free(m->commpartner);
  free(m->sendcount); free(m->recvcount); free(m->addpoint_owner); free(m->addpoint_idx);
  free(m->global_id);
  memset(m, 0, sizeof *m);
}

int cfdp_mesh_gen_domain(const cfdp_mesh_spec *s, int rank, cfdp_mesh_domain *m)
{
  memset(m, 0, sizeof *m);
  const int nd = cfdp_mesh_num_domains(s);
  if (rank < 0 || rank >= nd || s->nx < 2 || s->ny < 2 || s->nz < 2) return -1;
  box_t b; domain_box(s, rank, &b);
  const long long sx = b.hi[0] - b.lo[0], sy = b.hi[1] - b.lo[1], sz = b.hi[2] - b.lo[2];
  if (sx < 1 || sy < 1 || sz < 1) return -1;
  const long long nown = sx * sy * sz;
  /* extended box: one layer on every side, clipped to the lattice */
  int elo[3], ehi[3];
  const int n3[3] = { s->nx, s->ny, s->nz };
  for (int a = 0; a < 3; a++) { elo[a] = b.lo[a] > 0 ? b.lo[a] - 1 : 0; ehi[a] = b.hi[a] < n3[a] ? b.hi[a] + 1 : n3[a]; }
  const long long ex = ehi[0] - elo[0], ey = ehi[1] - elo[1], ez = ehi[2] - elo[2];
  const long long next = ex * ey * ez;
  if (nown * 21 >= 2147483647LL && !s->allow_big) return -2; /* reference int overflow limit, SURVEY 3.5 */

  int *lid = (int *)malloc((size_t)next * sizeof(int));
  if (!lid) return -3;
#define EIDX(x, y, z) (((long long)(x) - elo[0]) + ex * (((long long)(y) - elo[1]) + ey * ((long long)(z) - elo[2])))
#define INBOX(x, y, z) ((x) >= b.lo[0] && (x) < b.hi[0] && (y) >= b.lo[1] && (y) < b.hi[1] && (z) >= b.lo[2] && (z) < b.hi[2])

  int *perm = NULL;
  if (s->order == CFDP_ORDER_SHUFFLE) { perm = shuffle_perm(s, rank, nown); if (!perm) { free(lid); return -3; } }

  /* pass 1: mark own points and shell points adjacent to an own point through an existing edge */
#pragma omp parallel for schedule(static)
  for (long long zz = elo[2]; zz < ehi[2]; zz++)
    for (int y = elo[1]; y < ehi[1]; y++)
      for (int x = elo[0]; x < ehi[0]; x++) {
        int z = (int)zz;
        long long e = EIDX(x, y, z);
        if (INBOX(x, y, z)) {
          long long r = own_rank_in_box(s, &b, x, y, z);
          lid[e] = perm ? perm[r] : (int)r;
          continue;
        }
        int ghost = 0;
        for (int d = 0; d < 7 && !ghost; d++) {
          int qx = x + DIRS[d][0], qy = y + DIRS[d][1], qz = z + DIRS[d][2];
          if (qx < s->nx && qy < s->ny && qz < s->nz && edge_exists(s, x, d) && INBOX(qx, qy, qz)) ghost = 1;
          qx = x - DIRS[d][0]; qy = y - DIRS[d][1]; qz = z - DIRS[d][2];
          if (qx >= 0 && qy >= 0 && qz >= 0 && edge_exists(s, qx, d) && INBOX(qx, qy, qz)) ghost = 1;
        }
        lid[e] = ghost ? -2 : -1;
      }
  /* ghosts numbered in ascending global id (z,y,x) order */
  long long nadd = 0;
  for (long long e = 0; e < next; e++) if (lid[e] == -2) nadd++;
  m->nown = (int)nown; m->nadd = (int)nadd; m->nall = (int)(nown + nadd); m->ndomains = nd;
  m->pvolume = (double *)malloc((size_t)(nown + nadd) * sizeof(double));
  m->global_id = (long long *)malloc((size_t)(nown + nadd) * sizeof(long long));
  m->sendcount = (int *)calloc((size_t)nd, sizeof(int));
  m->recvcount = (int *)calloc((size_t)nd, sizeof(int));
  m->addpoint_owner = (int *)malloc((size_t)(nadd > 0 ? nadd : 1) * sizeof(int));
  m->addpoint_idx = (int *)malloc((size_t)(nadd > 0 ? nadd : 1) * sizeof(int));
  if (!m->pvolume || !m->global_id || !m->sendcount || !m->recvcount || !m->addpoint_owner || !m->addpoint_idx) { free(lid); free(perm); cfdp_mesh_free_domain(m); return -3; }

  int **pperm = NULL;
  if (s->order == CFDP_ORDER_SHUFFLE) pperm = (int **)calloc((size_t)nd, sizeof(int *));
  {
    long long j = 0;
    for (int z = elo[2]; z < ehi[2]; z++)
      for (int y = elo[1]; y < ehi[1]; y++)
        for (int x = elo[0]; x < ehi[0]; x++) {
          long long e = EIDX(x, y, z);
          long long gid = (long long)x + (long long)s->nx * ((long long)y + (long long)s->ny * z);
          if (lid[e] >= 0) { m->global_id[lid[e]] = gid; m->pvolume[lid[e]] = 0.5 + u01(s->seed ^ (uint64_t)gid * 3u); continue; }
          if (lid[e] != -2) continue;
          int ow = owner_of(s, x, y, z);
          box_t ob; domain_box(s, ow, &ob);
          long long r = own_rank_in_box(s, &ob, x, y, z);
          if (pperm) {
            if (!pperm[ow]) {
              long long on = (long long)(ob.hi[0] - ob.lo[0]) * (ob.hi[1] - ob.lo[1]) * (ob.hi[2] - ob.lo[2]);
              pperm[ow] = shuffle_perm(s, ow, on);
            }
            r = pperm[ow][r];
          }
          m->addpoint_owner[j] = ow;
          m->addpoint_idx[j] = (int)r;
          m->recvcount[ow]++;
          m->global_id[nown + j] = gid;
          m->pvolume[nown + j] = 0.5 + u01(s->seed ^ (uint64_t)gid * 3u);
          lid[e] = (int)(nown + j);
          j++;
        }
  }
  if (pperm) { for (int k = 0; k < nd; k++) free(pperm[k]); free(pperm); }
  free(perm);

  /* sendcount[k] = number of own points that are ghosts of k = recvcount_k[me]; by symmetry of
   * the edge stencil: own points adjacent (through an existing edge) to a point owned by k */
  {
    unsigned char *seen = (unsigned char *)calloc((size_t)nd, 1);
    for (int z = b.lo[2]; z < b.hi[2]; z++)
      for (int y = b.lo[1]; y < b.hi[1]; y++)
        for (int x = b.lo[0]; x < b.hi[0]; x++) {
          /* interior points cannot touch another domain */
          if (x > b.lo[0] && x < b.hi[0] - 1 && y > b.lo[1] && y < b.hi[1] - 1 && z > b.lo[2] && z < b.hi[2] - 1) { x = b.hi[0] - 2; continue; }
          int touched[14], nt = 0;
          for (int d = 0; d < 7; d++) {
            int qx = x + DIRS[d][0], qy = y + DIRS[d][1], qz = z + DIRS[d][2];
            if (qx < s->nx && qy < s->ny && qz < s->nz && edge_exists(s, x, d) && !INBOX(qx, qy, qz)) {
              int k = owner_of(s, qx, qy, qz); if (!seen[k]) { seen[k] = 1; touched[nt++] = k; }
            }
            qx = x - DIRS[d][0]; qy = y - DIRS[d][1]; qz = z - DIRS[d][2];
            if (qx >= 0 && qy >= 0 && qz >= 0 && edge_exists(s, qx, d) && !INBOX(qx, qy, qz)) {
              int k = owner_of(s, qx, qy, qz); if (!seen[k]) { seen[k] = 1; touched[nt++] = k; }
            }
          }
          for (int t = 0; t < nt; t++) { m->sendcount[touched[t]]++; seen[touched[t]] = 0; }
        }
    free(seen);
  }
  int ncomm = 0;
  for (int k = 0; k < nd; k++) if (m->sendcount[k] > 0 || m->recvcount[k] > 0) ncomm++;
  m->ncommdomains = ncomm;
  m->commpartner = (int *)malloc((size_t)(ncomm > 0 ? ncomm : 1) * sizeof(int));
  ncomm = 0;
  for (int k = 0; k < nd; k++) if (m->sendcount[k] > 0 || m->recvcount[k] > 0) m->commpartner[ncomm++] = k;

  /* pass 2: faces.  Scan lower endpoints p over the extended box (z-planes in parallel). */
  long long *zcount = (long long *)calloc((size_t)ez + 1, sizeof(long long));
#pragma omp parallel for schedule(static)
  for (long long zz = elo[2]; zz < ehi[2]; zz++) {
    int z = (int)zz; long long c = 0;
    for (int y = elo[1]; y < ehi[1]; y++)
      for (int x = elo[0]; x < ehi[0]; x++) {
        int l0 = lid[EIDX(x, y, z)]; if (l0 < 0) continue;
        for (int d = 0; d < 7; d++) {
          int qx = x + DIRS[d][0], qy = y + DIRS[d][1], qz = z + DIRS[d][2];
          if (qx >= ehi[0] || qy >= ehi[1] || qz >= ehi[2] || !edge_exists(s, x, d)) continue;
          int l1 = lid[EIDX(qx, qy, qz)]; if (l1 < 0) continue;
          if (l0 >= nown && l1 >= nown) continue;       /* ghost-ghost faces are not part of a domain */
          c++;
        }
      }
    zcount[zz - elo[2] + 1] = c;
  }
  for (long long i = 0; i < ez; i++) zcount[i + 1] += zcount[i];
  const long long nfaces = zcount[ez];
  m->nfaces = (int)nfaces;
  if (nfaces * 3 >= 2147483647LL && !s->allow_big) { free(lid); free(zcount); cfdp_mesh_free_domain(m); return -2; }
  m->fpoint = (int *)malloc((size_t)nfaces * 2 * sizeof(int));
  m->fnormal = (double *)malloc((size_t)nfaces * 3 * sizeof(double));
  if (!m->fpoint || !m->fnormal) { free(lid); free(zcount); cfdp_mesh_free_domain(m); return -3; }
#pragma omp parallel for schedule(static)
  for (long long zz = elo[2]; zz < ehi[2]; zz++) {
    int z = (int)zz; long long f = zcount[zz - elo[2]];
    for (int y = elo[1]; y < ehi[1]; y++)
      for (int x = elo[0]; x < ehi[0]; x++) {
        int l0 = lid[EIDX(x, y, z)]; if (l0 < 0) continue;
        long long gid = (long long)x + (long long)s->nx * ((long long)y + (long long)s->ny * z);
        for (int d = 0; d < 7; d++) {
          int qx = x + DIRS[d][0], qy = y + DIRS[d][1], qz = z + DIRS[d][2];
          if (qx >= ehi[0] || qy >= ehi[1] || qz >= ehi[2] || !edge_exists(s, x, d)) continue;
          int l1 = lid[EIDX(qx, qy, qz)]; if (l1 < 0) continue;
          if (l0 >= nown && l1 >= nown) continue;
          uint64_t key = s->seed ^ ((uint64_t)gid * 7u + (uint64_t)d) * 0x9E3779B97F4A7C15ull;
          double w = DIRW[d] * (1.0 + s->jitter * (2.0 * u01(key) - 1.0));
          m->fpoint[2 * f] = l0; m->fpoint[2 * f + 1] = l1;
          m->fnormal[3 * f + 0] = w * DIRS[d][0] + s->jitter * 0.25 * (2.0 * u01(key + 1) - 1.0);
          m->fnormal[3 * f + 1] = w * DIRS[d][1] + s->jitter * 0.25 * (2.0 * u01(key + 2) - 1.0);
          m->fnormal[3 * f + 2] = w * DIRS[d][2] + s->jitter * 0.25 * (2.0 * u01(key + 3) - 1.0);
          f++;
        }
      }
  }
  free(zcount);
  free(lid);
  return 0;
#undef EIDX
#undef INBOX
}

void cfdp_mesh_fill_var(const cfdp_mesh_domain *m, unsigned long long seed, double *var /* [nall][7] */)
{
#pragma omp parallel for schedule(static)
  for (long long p = 0; p < m->nall; p++)
    for (int eq = 0; eq < 7; eq++) var[7 * p + eq] = cfdp_mesh_var_value(seed, m->global_id[p], eq);
}
