/*
 * read_netcdf.c -- mesh loader of the drop-in boundary.
 *
 * Replaces the reference's three libnetcdf wrappers (reference: src/read_netcdf.c:20-61:
 * get_nc_val = nc_inq_dimid + nc_inq_dimlen, get_nc_int/get_nc_double = nc_inq_varid +
 * nc_get_var_*) and the eight libnetcdf calls they and the driver (src/hybrid.f6.c:65-66,
 * 89-90) make.  No libnetcdf exists in this image, so the NetCDF-3 "classic" container
 * (CDF-1 32-bit offsets / CDF-2 64-bit offsets, big-endian) is parsed
 * here: the file is mmap'ed, the header walked once at open time, and variables are
 * byte-swapped straight out of the mapping (OpenMP over the element range).
 * Errors follow the reference: get_nc_* print "Error: <text>" and exit(2)
 * (error_handling.h:6-10).
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include "cfdp_b200.h"

#define NC_MAX_OPEN 256
#define NC_ERRCODE 2

enum { NCE_OK = 0, NCE_BADID = -33, NCE_NFILE = -34, NCE_NOTNC = -51, NCE_BADDIM = -46, NCE_NOTVAR = -49,
       NCE_BADTYPE = -45, NCE_TRUNC = -60, NCE_OPEN = -31, NCE_NOMEM = -61 };

typedef struct { const char *name; uint32_t namelen; uint64_t len; } ncdim;
typedef struct { const char *name; uint32_t namelen; int type; uint64_t nelem; uint64_t begin; } ncvar;
typedef struct {
  int used;
  const unsigned char *map; size_t size;
  int ndims, nvars;
  ncdim *dims; ncvar *vars;
  char *path;
} ncfile;

static ncfile g_files[NC_MAX_OPEN];

const char *cfdp_nc_strerror(int err)
{
  switch (err) {
  case NCE_OK: return "No error";
  case NCE_BADID: return "NetCDF: Not a valid ID";
  case NCE_NFILE: return "NetCDF: Too many files open";
  case NCE_NOTNC: return "NetCDF: Unknown file format (not NetCDF-3 classic CDF-1/CDF-2)";
  case NCE_BADDIM: return "NetCDF: Invalid dimension ID or name";
  case NCE_NOTVAR: return "NetCDF: Variable not found";
  case NCE_BADTYPE: return "NetCDF: Not a valid data type or type mismatch";
  case NCE_TRUNC: return "NetCDF: file truncated";
  case NCE_OPEN: return "No such file or directory";
  case NCE_NOMEM: return "NetCDF: Memory allocation (malloc) failure";
  }
  return "NetCDF: Unknown Error";
}

typedef struct { const unsigned char *p, *end; int bad; } cursor;

static uint32_t get32(cursor *c)
{
  if (c->end - c->p < 4) { c->bad = 1; return 0; }
  uint32_t v; memcpy(&v, c->p, 4); c->p += 4;
  return __builtin_bswap32(v);
}
static uint64_t get64(cursor *c)
{
  uint64_t hi = get32(c), lo = get32(c);
  return (hi << 32) | lo;
}
static const char *getname(cursor *c, uint32_t *len)
{
  uint32_t n = get32(c);
  uint64_t padded = ((uint64_t)n + 3u) & ~(uint64_t)3u;
  if (c->bad || (uint64_t)(c->end - c->p) < padded) { c->bad = 1; return NULL; }
  const char *s = (const char *)c->p;
  c->p += padded; *len = n;
  return s;
}
static size_t nc_type_size(int t)
{
  switch (t) { case 1: case 2: return 1; case 3: return 2; case 4: case 5: return 4; case 6: return 8; }
  return 0;
}
static void skip_att_list(cursor *c)
{
  uint32_t tag = get32(c), n = get32(c);
  if (c->bad || (tag == 0 && n == 0)) return;
  if (tag != 0x0C) { c->bad = 1; return; }
  for (uint32_t i = 0; i < n && !c->bad; i++) {
    uint32_t nl; getname(c, &nl);
    uint32_t ty = get32(c), ne = get32(c);
    uint64_t bytes = ((uint64_t)ne * nc_type_size((int)ty) + 3u) & ~(uint64_t)3u;
    if (c->bad || nc_type_size((int)ty) == 0 || (uint64_t)(c->end - c->p) < bytes) { c->bad = 1; return; }
    c->p += bytes;
  }
}

int cfdp_nc_open(const char *path, int mode, int *ncidp)
{
  (void)mode;
  int id = -1;
  for (int i = 0; i < NC_MAX_OPEN; i++) if (!g_files[i].used) { id = i; break; }
  if (id < 0) return NCE_NFILE;
  int fd = open(path, O_RDONLY);
  if (fd < 0) return NCE_OPEN;
  struct stat st;
  if (fstat(fd, &st) || st.st_size < 32) { close(fd); return NCE_NOTNC; }
  void *map = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED) return NCE_NOMEM;
  ncfile *f = &g_files[id];
  memset(f, 0, sizeof *f);
  f->map = (const unsigned char *)map; f->size = (size_t)st.st_size;
  cursor c = { f->map, f->map + f->size, 0 };
  int rc = NCE_NOTNC;
  if (!(c.p[0] == 'C' && c.p[1] == 'D' && c.p[2] == 'F' && (c.p[3] == 1 || c.p[3] == 2))) goto fail;
  const int wide = c.p[3] == 2;
  c.p += 4;
  (void)get32(&c); /* numrecs: the F6 schema has no record dimension */
  uint32_t tag = get32(&c), n = get32(&c);
  if (c.bad) goto fail;
  if (!(tag == 0 && n == 0)) {
    if (tag != 0x0A) goto fail;
    f->dims = (ncdim *)calloc(n ? n : 1, sizeof(ncdim));
    if (!f->dims) { rc = NCE_NOMEM; goto fail; }
    for (uint32_t i = 0; i < n; i++) {
      f->dims[i].name = getname(&c, &f->dims[i].namelen);
      f->dims[i].len = get32(&c);
      if (c.bad) goto fail;
    }
    f->ndims = (int)n;
  }
  skip_att_list(&c);
  tag = get32(&c); n = get32(&c);
  if (c.bad) goto fail;
  if (!(tag == 0 && n == 0)) {
    if (tag != 0x0B) goto fail;
    f->vars = (ncvar *)calloc(n ? n : 1, sizeof(ncvar));
    if (!f->vars) { rc = NCE_NOMEM; goto fail; }
    for (uint32_t i = 0; i < n; i++) {
      ncvar *v = &f->vars[i];
      v->name = getname(&c, &v->namelen);
      uint32_t nd = get32(&c);
      if (c.bad || nd > 1024) goto fail;
      v->nelem = 1;
      for (uint32_t d = 0; d < nd; d++) {
        uint32_t di = get32(&c);
        if (c.bad || (int)di >= f->ndims) goto fail;
        if (f->dims[di].len != 0 && v->nelem > (uint64_t)-1 / 16 / f->dims[di].len) { rc = NCE_TRUNC; goto fail; } /* crafted header: product overflows */
        v->nelem *= f->dims[di].len;
      }
      skip_att_list(&c);
      v->type = (int)get32(&c);
      (void)get32(&c); /* vsize: redundant (and saturated for >4 GiB variables) */
      v->begin = wide ? get64(&c) : get32(&c);
      if (c.bad || nc_type_size(v->type) == 0) goto fail;
      /* overflow-checked: nelem * size cannot wrap (bounded above), begin is compared without adding to it */
      if (v->begin > f->size || v->nelem * nc_type_size(v->type) > f->size - v->begin) { rc = NCE_TRUNC; goto fail; }
    }
    f->nvars = (int)n;
  }
  f->path = strdup(path);
  f->used = 1;
  *ncidp = id;
  return NCE_OK;
fail:
  free(f->dims); free(f->vars);
  munmap(map, (size_t)st.st_size);
  memset(f, 0, sizeof *f);
  return rc;
}

static ncfile *file_of(int ncid) { return (ncid >= 0 && ncid < NC_MAX_OPEN && g_files[ncid].used) ? &g_files[ncid] : NULL; }

int cfdp_nc_close(int ncid)
{
  ncfile *f = file_of(ncid);
  if (!f) return NCE_BADID;
  munmap((void *)f->map, f->size);
  free(f->dims); free(f->vars); free(f->path);
  memset(f, 0, sizeof *f);
  return NCE_OK;
}

const char *cfdp_nc_path(int ncid) { ncfile *f = file_of(ncid); return f ? f->path : NULL; }

static int name_is(const char *s, uint32_t len, const char *name) { return strlen(name) == len && !memcmp(s, name, len); }

int cfdp_nc_inq_dimid(int ncid, const char *name, int *dimidp)
{
  ncfile *f = file_of(ncid);
  if (!f) return NCE_BADID;
  for (int i = 0; i < f->ndims; i++) if (name_is(f->dims[i].name, f->dims[i].namelen, name)) { *dimidp = i; return NCE_OK; }
  return NCE_BADDIM;
}
int cfdp_nc_inq_dimlen(int ncid, int dimid, size_t *lenp)
{
  ncfile *f = file_of(ncid);
  if (!f) return NCE_BADID;
  if (dimid < 0 || dimid >= f->ndims) return NCE_BADDIM;
  *lenp = (size_t)f->dims[dimid].len;
  return NCE_OK;
}
int cfdp_nc_inq_varid(int ncid, const char *name, int *varidp)
{
  ncfile *f = file_of(ncid);
  if (!f) return NCE_BADID;
  for (int i = 0; i < f->nvars; i++) if (name_is(f->vars[i].name, f->vars[i].namelen, name)) { *varidp = i; return NCE_OK; }
  return NCE_NOTVAR;
}
int cfdp_nc_get_var_int(int ncid, int varid, int *ip)
{
  ncfile *f = file_of(ncid);
  if (!f) return NCE_BADID;
  if (varid < 0 || varid >= f->nvars) return NCE_NOTVAR;
  const ncvar *v = &f->vars[varid];
  if (v->type != 4) return NCE_BADTYPE;
  const unsigned char *src = f->map + v->begin;
  const int64_t n = (int64_t)v->nelem;
#pragma omp parallel for schedule(static) if (n > (1 << 16))
  for (int64_t i = 0; i < n; i++) { uint32_t u; memcpy(&u, src + 4 * i, 4); u = __builtin_bswap32(u); memcpy(&ip[i], &u, 4); }
  return NCE_OK;
}
int cfdp_nc_get_var_double(int ncid, int varid, double *dp)
{
  ncfile *f = file_of(ncid);
  if (!f) return NCE_BADID;
  if (varid < 0 || varid >= f->nvars) return NCE_NOTVAR;
  const ncvar *v = &f->vars[varid];
  if (v->type != 6) return NCE_BADTYPE;
  const unsigned char *src = f->map + v->begin;
  const int64_t n = (int64_t)v->nelem;
#pragma omp parallel for schedule(static) if (n > (1 << 16))
  for (int64_t i = 0; i < n; i++) { uint64_t u; memcpy(&u, src + 8 * i, 8); u = __builtin_bswap64(u); memcpy(&dp[i], &u, 8); }
  return NCE_OK;
}

/* ---- the reference's wrappers (read_netcdf.c:20-61), same names and behaviour ---- */
#define NC_DIE(e) do { printf("Error: %s\n", cfdp_nc_strerror(e)); exit(NC_ERRCODE); } while (0)

void get_nc_int(int ncid, const char *name, int *array)
{
  int varid, rc;
  if ((rc = cfdp_nc_inq_varid(ncid, name, &varid))) NC_DIE(rc);
  if ((rc = cfdp_nc_get_var_int(ncid, varid, array))) NC_DIE(rc);
}
void get_nc_double(int ncid, const char *name, double *array)
{
  int varid, rc;
  if ((rc = cfdp_nc_inq_varid(ncid, name, &varid))) NC_DIE(rc);
  if ((rc = cfdp_nc_get_var_double(ncid, varid, array))) NC_DIE(rc);
}
int get_nc_val(int ncid, const char *name)
{
  int dimid, rc; size_t val;
  if ((rc = cfdp_nc_inq_dimid(ncid, name, &dimid))) NC_DIE(rc);
  if ((rc = cfdp_nc_inq_dimlen(ncid, dimid, &val))) NC_DIE(rc);
  return (int)val;
}
