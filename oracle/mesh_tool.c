/*
 * oracle/mesh_tool.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Standalone writer of synthetic F6-schema mesh files for the reference arm of bench.py and for the oracle rig:
 * links the mesh generator source (cfd_proxy_b200/csrc/mesh_gen.c) directly, so that the process that times the
 * reference's CPU path never maps the product library libcfdp_b200.so.  Writes, for every domain,
 *   <PREFIX>_domain_<rank>_lvl_<L>        NetCDF-3 64-bit-offset (CDF-2) file in the schema the reference loader reads
 *                                         (reference src/solver_data.c:98-144, src/comm_data.c:79-112, hybrid.f6.c:57-62)
 *   <PREFIX>_domain_<rank>_lvl_<L>.var    seeded var[nall][7], raw little-endian doubles (read by ref_harness.c)
 * and prints one line of JSON (points, faces with at least one own end point).
 *
 *   mesh_tool PREFIX LVL nx ny nz px py pz ORDER(0 lex,1 brick,2 shuffle) BRICK HEXCUT JITTER SEED
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <omp.h>
#include "cfdp_b200.h"

typedef struct { char *p; size_t n, cap; } buf_t;
static void put(buf_t *b, const void *src, size_t n)
{
  if (b->n + n > b->cap) { b->cap = (b->n + n) * 2 + 256; b->p = (char *)realloc(b->p, b->cap); if (!b->p) { perror("realloc"); exit(1); } }
  memcpy(b->p + b->n, src, n); b->n += n;
}
static void put_u32(buf_t *b, uint32_t v) { unsigned char c[4] = { (unsigned char)(v >> 24), (unsigned char)(v >> 16), (unsigned char)(v >> 8), (unsigned char)v }; put(b, c, 4); }
static void put_u64(buf_t *b, uint64_t v) { put_u32(b, (uint32_t)(v >> 32)); put_u32(b, (uint32_t)v); }
static void put_name(buf_t *b, const char *s)
{
  const size_t n = strlen(s); static const char z[4] = {0, 0, 0, 0};
  put_u32(b, (uint32_t)n); put(b, s, n); put(b, z, (4 - n % 4) % 4);
}

enum { NC_INT_T = 4, NC_DOUBLE_T = 6 };
typedef struct { const char *name; int ndims, dim[2], type; size_t count; const void *data; } var_t;

static void write_be(FILE *f, const void *data, size_t count, int width)
{
  enum { CH = 1 << 20 };
  unsigned char *tmp = (unsigned char *)malloc((size_t)CH * 8);
  const unsigned char *src = (const unsigned char *)data;
  for (size_t i = 0; i < count; i += CH) {
    const size_t n = count - i < CH ? count - i : CH;
#pragma omp parallel for schedule(static)
    for (size_t j = 0; j < n; j++)
      for (int k = 0; k < width; k++) tmp[j * width + k] = src[(i + j) * width + (width - 1 - k)];
    if (fwrite(tmp, (size_t)width, n, f) != n) { perror("fwrite"); exit(1); }
  }
  free(tmp);
}

static void write_domain(const char *path, const cfdp_mesh_domain *m)
{
  const char *dname[9] = { "ncolors", "nfaces", "nownpoints", "nallpoints", "ndomains", "two", "three", "naddpoints", "ncommdomains" };
  const size_t dlen[9] = { 1, (size_t)m->nfaces, (size_t)m->nown, (size_t)m->nall, (size_t)m->ndomains, 2, 3, (size_t)m->nadd, (size_t)m->ncommdomains };
  const int ndims = m->ndomains > 1 ? 9 : 7;
  int one_color = m->nall;
  int *ident = (int *)malloc((size_t)m->nall * sizeof(int));
  for (int i = 0; i < m->nall; i++) ident[i] = i;   /* simplest legal colouring; the reference reads and discards it (threads.c:748-749) */
  var_t v[10]; int nv = 0;
  v[nv++] = (var_t){ "fpoint", 2, {1, 5}, NC_INT_T, (size_t)m->nfaces * 2, m->fpoint };
  v[nv++] = (var_t){ "fnormal", 2, {1, 6}, NC_DOUBLE_T, (size_t)m->nfaces * 3, m->fnormal };
  v[nv++] = (var_t){ "pvolume", 1, {3, 0}, NC_DOUBLE_T, (size_t)m->nall, m->pvolume };
  v[nv++] = (var_t){ "fcolor_npoints", 1, {0, 0}, NC_INT_T, 1, &one_color };
  v[nv++] = (var_t){ "fcolor_points", 1, {3, 0}, NC_INT_T, (size_t)m->nall, ident };
  if (m->ndomains > 1) {
    v[nv++] = (var_t){ "commpartner", 1, {8, 0}, NC_INT_T, (size_t)m->ncommdomains, m->commpartner };
    v[nv++] = (var_t){ "sendcount", 1, {4, 0}, NC_INT_T, (size_t)m->ndomains, m->sendcount };
    v[nv++] = (var_t){ "recvcount", 1, {4, 0}, NC_INT_T, (size_t)m->ndomains, m->recvcount };
    v[nv++] = (var_t){ "addpoint_owner", 1, {7, 0}, NC_INT_T, (size_t)m->nadd, m->addpoint_owner };
    v[nv++] = (var_t){ "addpoint_idx", 1, {7, 0}, NC_INT_T, (size_t)m->nadd, m->addpoint_idx };
  }
  uint64_t begin[10];
  buf_t h = {0, 0, 0};
  for (int pass = 0; pass < 2; pass++) {   /* the header length does not depend on the offsets: lay it out twice */
    const size_t hlen = h.n;
    h.n = 0;
    put(&h, "CDF\002", 4); put_u32(&h, 0);
    put_u32(&h, 0x0A); put_u32(&h, (uint32_t)ndims);
    for (int d = 0; d < ndims; d++) { put_name(&h, dname[d]); put_u32(&h, (uint32_t)dlen[d]); }
    put_u32(&h, 0); put_u32(&h, 0);
    put_u32(&h, 0x0B); put_u32(&h, (uint32_t)nv);
    uint64_t pos = hlen;
    for (int i = 0; i < nv; i++) {
      const uint64_t nbytes = (uint64_t)v[i].count * (v[i].type == NC_INT_T ? 4 : 8);
      const uint64_t padded = (nbytes + 3) & ~(uint64_t)3;
      begin[i] = pos; pos += padded;
      put_name(&h, v[i].name); put_u32(&h, (uint32_t)v[i].ndims);
      for (int d = 0; d < v[i].ndims; d++) put_u32(&h, (uint32_t)v[i].dim[d]);
      put_u32(&h, 0); put_u32(&h, 0);
      put_u32(&h, (uint32_t)v[i].type); put_u32(&h, padded > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)padded);
      put_u64(&h, begin[i]);
    }
  }
  FILE *f = fopen(path, "wb");
  if (!f) { perror(path); exit(1); }
  if (fwrite(h.p, 1, h.n, f) != h.n) { perror("fwrite"); exit(1); }
  for (int i = 0; i < nv; i++) {
    write_be(f, v[i].data, v[i].count, v[i].type == NC_INT_T ? 4 : 8);
    const uint64_t nbytes = (uint64_t)v[i].count * (v[i].type == NC_INT_T ? 4 : 8);
    static const char z[4] = {0, 0, 0, 0};
    if (nbytes % 4) fwrite(z, 1, 4 - nbytes % 4, f);
  }
  fclose(f); free(h.p); free(ident);
}

int main(int argc, char **argv)
{
  if (argc < 14) { fprintf(stderr, "usage: %s PREFIX LVL nx ny nz px py pz ORDER BRICK HEXCUT JITTER SEED\n", argv[0]); return 2; }
  const char *prefix = argv[1]; const int lvl = atoi(argv[2]);
  cfdp_mesh_spec s; memset(&s, 0, sizeof s);
  s.nx = atoi(argv[3]); s.ny = atoi(argv[4]); s.nz = atoi(argv[5]);
  s.px = atoi(argv[6]); s.py = atoi(argv[7]); s.pz = atoi(argv[8]);
  s.order = atoi(argv[9]); s.brick = atoi(argv[10]); s.hexcut = atoi(argv[11]);
  s.jitter = atof(argv[12]); s.seed = strtoull(argv[13], NULL, 0); s.allow_big = 0;   /* the reference's int limits hold */
  const int nd = cfdp_mesh_num_domains(&s);
  long long points = 0, faces = 0;
  for (int r = 0; r < nd; r++) {
    cfdp_mesh_domain m;
    const int rc = cfdp_mesh_gen_domain(&s, r, &m);
    if (rc != 0) { fprintf(stderr, "cfdp_mesh_gen_domain rc=%d\n", rc); return 1; }
    char path[1024];
    snprintf(path, sizeof path, "%s_domain_%d_lvl_%d", prefix, r, lvl);
    write_domain(path, &m);
    double *var = (double *)malloc((size_t)m.nall * 7 * sizeof(double));
    cfdp_mesh_fill_var(&m, s.seed, var);
    strncat(path, ".var", sizeof path - strlen(path) - 1);
    FILE *vf = fopen(path, "wb");
    if (!vf || fwrite(var, sizeof(double), (size_t)m.nall * 7, vf) != (size_t)m.nall * 7) { perror(path); return 1; }
    fclose(vf); free(var);
    points += m.nown;
    for (int f = 0; f < m.nfaces; f++) if (m.fpoint[2 * f] < m.nown || m.fpoint[2 * f + 1] < m.nown) faces++;
    cfdp_mesh_free_domain(&m);
  }
  printf("{\"domains\": %d, \"points\": %lld, \"faces\": %lld}\n", nd, points, faces);
  return 0;
}
