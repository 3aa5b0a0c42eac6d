/*
 * oracle/shim/nc_cdf.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Minimal libnetcdf stand-in for the oracle rig: just enough of the nc_* API for the
 * unmodified reference loader (reference: src/read_netcdf.c:20-61) to read NetCDF-3
 * classic files (CDF-1 32-bit offsets, CDF-2 64-bit offsets; big-endian on disk).
 * Written independently of the product loader (cfd-proxy_b200/csrc/read_netcdf.c) so
 * that a bug in one cannot hide in the other.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "netcdf.h"

#define MAXF 64
#define MAXN 64

typedef struct { char name[128]; size_t len; } dim_t;
typedef struct { char name[128]; int ndims; int dimid[8]; int type; uint64_t begin; size_t nelem; } var_t;
typedef struct { FILE *fp; int used; int ndim, nvar; dim_t dim[MAXN]; var_t var[MAXN]; } file_t;

static file_t files[MAXF];

enum { E_OK = 0, E_OPEN = -31, E_FMT = -51, E_NOTFOUND = -49, E_TYPE = -45, E_IO = -60 };

const char *nc_strerror(int e)
{
  switch (e) {
  case E_OK: return "No error";
  case E_OPEN: return "oracle nc shim: cannot open file";
  case E_FMT: return "oracle nc shim: not a NetCDF-3 classic file";
  case E_NOTFOUND: return "oracle nc shim: dimension or variable not found";
  case E_TYPE: return "oracle nc shim: unexpected variable type";
  case E_IO: return "oracle nc shim: short read";
  }
  return "oracle nc shim: unknown error";
}

static int rd_u32(FILE *fp, uint32_t *v)
{
  unsigned char b[4];
  if (fread(b, 1, 4, fp) != 4) return E_IO;
  *v = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
  return 0;
}

static int rd_u64(FILE *fp, uint64_t *v)
{
  uint32_t hi, lo;
  if (rd_u32(fp, &hi) || rd_u32(fp, &lo)) return E_IO;
  *v = ((uint64_t)hi << 32) | lo;
  return 0;
}

static int rd_name(FILE *fp, char *out, size_t cap)
{
  uint32_t n;
  if (rd_u32(fp, &n)) return E_IO;
  size_t padded = (n + 3u) & ~3u;
  char tmp[512];
  if (padded > sizeof tmp || n >= cap) return E_FMT;
  if (fread(tmp, 1, padded, fp) != padded) return E_IO;
  memcpy(out, tmp, n);
  out[n] = 0;
  return 0;
}

static size_t type_size(int t)
{
  switch (t) { case 1: case 2: return 1; case 3: return 2; case 4: case 5: return 4; case 6: return 8; }
  return 0;
}

static int skip_attrs(FILE *fp)
{
  uint32_t tag, n;
  if (rd_u32(fp, &tag) || rd_u32(fp, &n)) return E_IO;
  if (tag == 0 && n == 0) return 0;
  if (tag != 0x0C) return E_FMT;
  for (uint32_t i = 0; i < n; i++) {
    char nm[256];
    uint32_t ty, ne;
    int r;
    if ((r = rd_name(fp, nm, sizeof nm))) return r;
    if (rd_u32(fp, &ty) || rd_u32(fp, &ne)) return E_IO;
    size_t bytes = ((size_t)ne * type_size((int)ty) + 3u) & ~(size_t)3u;
    if (fseeko(fp, (off_t)bytes, SEEK_CUR)) return E_IO;
  }
  return 0;
}

int nc_open(const char *path, int mode, int *ncidp)
{
  (void)mode;
  int id;
  for (id = 0; id < MAXF && files[id].used; id++) ;
  if (id == MAXF) return E_OPEN;
  file_t *f = &files[id];
  memset(f, 0, sizeof *f);
  f->fp = fopen(path, "rb");
  if (!f->fp) return E_OPEN;
  unsigned char magic[4];
  if (fread(magic, 1, 4, f->fp) != 4 || magic[0] != 'C' || magic[1] != 'D' || magic[2] != 'F' ||
      (magic[3] != 1 && magic[3] != 2)) { fclose(f->fp); return E_FMT; }
  int wide = magic[3] == 2;
  uint32_t numrecs, tag, n;
  int r;
  if (rd_u32(f->fp, &numrecs)) return E_IO;
  /* dim_list */
  if (rd_u32(f->fp, &tag) || rd_u32(f->fp, &n)) return E_IO;
  if (!(tag == 0 && n == 0)) {
    if (tag != 0x0A || n > MAXN) return E_FMT;
    for (uint32_t i = 0; i < n; i++) {
      uint32_t len;
      if ((r = rd_name(f->fp, f->dim[i].name, sizeof f->dim[i].name))) return r;
      if (rd_u32(f->fp, &len)) return E_IO;
      f->dim[i].len = len;
    }
    f->ndim = (int)n;
  }
  if ((r = skip_attrs(f->fp))) return r;
  /* var_list */
  if (rd_u32(f->fp, &tag) || rd_u32(f->fp, &n)) return E_IO;
  if (!(tag == 0 && n == 0)) {
    if (tag != 0x0B || n > MAXN) return E_FMT;
    for (uint32_t i = 0; i < n; i++) {
      var_t *v = &f->var[i];
      uint32_t nd, ty, vsize;
      if ((r = rd_name(f->fp, v->name, sizeof v->name))) return r;
      if (rd_u32(f->fp, &nd) || nd > 8) return E_FMT;
      v->ndims = (int)nd;
      v->nelem = 1;
      for (uint32_t d = 0; d < nd; d++) {
        uint32_t di;
        if (rd_u32(f->fp, &di) || (int)di >= f->ndim) return E_FMT;
        v->dimid[d] = (int)di;
        v->nelem *= f->dim[di].len;
      }
      if ((r = skip_attrs(f->fp))) return r;
      if (rd_u32(f->fp, &ty) || rd_u32(f->fp, &vsize)) return E_IO;
      v->type = (int)ty;
      if (wide) { if (rd_u64(f->fp, &v->begin)) return E_IO; }
      else { uint32_t b; if (rd_u32(f->fp, &b)) return E_IO; v->begin = b; }
    }
    f->nvar = (int)n;
  }
  f->used = 1;
  *ncidp = id;
  return 0;
}

int nc_close(int ncid)
{
  if (ncid < 0 || ncid >= MAXF || !files[ncid].used) return E_OPEN;
  fclose(files[ncid].fp);
  files[ncid].used = 0;
  return 0;
}

int nc_inq_dimid(int ncid, const char *name, int *dimidp)
{
  file_t *f = &files[ncid];
  for (int i = 0; i < f->ndim; i++)
    if (!strcmp(f->dim[i].name, name)) { *dimidp = i; return 0; }
  return E_NOTFOUND;
}

int nc_inq_dimlen(int ncid, int dimid, size_t *lenp)
{
  *lenp = files[ncid].dim[dimid].len;
  return 0;
}

int nc_inq_varid(int ncid, const char *name, int *varidp)
{
  file_t *f = &files[ncid];
  for (int i = 0; i < f->nvar; i++)
    if (!strcmp(f->var[i].name, name)) { *varidp = i; return 0; }
  return E_NOTFOUND;
}

static int read_raw(file_t *f, var_t *v, void *dst, size_t esz)
{
  if (fseeko(f->fp, (off_t)v->begin, SEEK_SET)) return E_IO;
  if (fread(dst, esz, v->nelem, f->fp) != v->nelem) return E_IO;
  return 0;
}

int nc_get_var_int(int ncid, int varid, int *ip)
{
  file_t *f = &files[ncid];
  var_t *v = &f->var[varid];
  if (v->type != 4) return E_TYPE;
  int r = read_raw(f, v, ip, 4);
  if (r) return r;
  uint32_t *u = (uint32_t *)ip;
  for (size_t i = 0; i < v->nelem; i++) u[i] = __builtin_bswap32(u[i]);
  return 0;
}

int nc_get_var_double(int ncid, int varid, double *dp)
{
  file_t *f = &files[ncid];
  var_t *v = &f->var[varid];
  if (v->type != 6) return E_TYPE;
  int r = read_raw(f, v, dp, 8);
  if (r) return r;
  uint64_t *u = (uint64_t *)dp;
  for (size_t i = 0; i < v->nelem; i++) u[i] = __builtin_bswap64(u[i]);
  return 0;
}
