/*
 * netcdf_compat.h -- maps the eight libnetcdf calls the reference makes (hybrid.f6.c:65-66,89-90;
 * read_netcdf.c:24-58) onto this library's NetCDF-3 reader, for builds without libnetcdf:
 *     gcc -include netcdf_compat.h ... hybrid.f6.c ... -lcfdp_b200
 */
#ifndef CFDP_NETCDF_COMPAT_H
#define CFDP_NETCDF_COMPAT_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
int cfdp_nc_open(const char *path, int mode, int *ncidp);
int cfdp_nc_close(int ncid);
const char *cfdp_nc_strerror(int err);
int cfdp_nc_inq_dimid(int ncid, const char *name, int *dimidp);
int cfdp_nc_inq_dimlen(int ncid, int dimid, size_t *lenp);
int cfdp_nc_inq_varid(int ncid, const char *name, int *varidp);
int cfdp_nc_get_var_int(int ncid, int varid, int *ip);
int cfdp_nc_get_var_double(int ncid, int varid, double *dp);
#ifdef __cplusplus
}
#endif
#define NC_NOWRITE 0
#define nc_open cfdp_nc_open
#define nc_close cfdp_nc_close
#define nc_strerror cfdp_nc_strerror
#define nc_inq_dimid cfdp_nc_inq_dimid
#define nc_inq_dimlen cfdp_nc_inq_dimlen
#define nc_inq_varid cfdp_nc_inq_varid
#define nc_get_var_int cfdp_nc_get_var_int
#define nc_get_var_double cfdp_nc_get_var_double
#endif
