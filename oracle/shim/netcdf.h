/*
 * oracle/shim/netcdf.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Prototype-only stand-in for <netcdf.h>: the 8 functions the reference's read_netcdf.c /
 * hybrid.f6.c call (reference: src/read_netcdf.c:20-61, src/hybrid.f6.c:65-66,89-90).
 * Implemented by nc_cdf.c, an independent NetCDF-3 classic (CDF-1/CDF-2) header parser.
 */
#ifndef CFDP_ORACLE_SHIM_NETCDF_H
#define CFDP_ORACLE_SHIM_NETCDF_H
#include <stddef.h>
#define NC_NOWRITE 0
#define NC_NOERR   0
int nc_open(const char *path, int mode, int *ncidp);
int nc_close(int ncid);
const char *nc_strerror(int err);
int nc_inq_dimid(int ncid, const char *name, int *dimidp);
int nc_inq_dimlen(int ncid, int dimid, size_t *lenp);
int nc_inq_varid(int ncid, const char *name, int *varidp);
int nc_get_var_int(int ncid, int varid, int *ip);
int nc_get_var_double(int ncid, int varid, double *dp);
#endif
