/* netcdf.h -- TEST INFRASTRUCTURE.  Declarations of the netCDF-C calls the REFERENCE's Grid.cpp and
 * Partitioner.cpp make (Grid.cpp:51-130, Partitioner.cpp:128-318), implemented over an in-memory
 * file table in oracle/ref_hostpath_shim.cpp so that those sources can be compiled where they lie
 * and run without libnetcdf.  Only int data, only what the reference calls.  Not part of the product. */
#ifndef DDC_REF_SHIM_NETCDF_H
#define DDC_REF_SHIM_NETCDF_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NC_NOERR 0
#define NC_NOWRITE 0
#define NC_CLOBBER 0
#define NC_NETCDF4 0x1000
#define NC_MPIIO 0x2000
#define NC_INT 4
#define NC_GLOBAL (-1)
#define NC_EBADID (-33)
#define NC_EBADDIM (-46)
#define NC_ENOTVAR (-49)
#define NC_EEDGE (-57)
#define NC_ENOGRP (-125)
#define NC_ENOENT 2

const char* nc_strerror(int err);
/* serial open / create: what THIS repository's host layer calls (one process drives the GPUs); the reference uses
   the parallel variants of netcdf_par.h */
int nc_open(const char* path, int mode, int* ncidp);
int nc_create(const char* path, int cmode, int* ncidp);
int nc_close(int ncid);
int nc_enddef(int ncid);
int nc_inq_ncid(int ncid, const char* name, int* grp_ncid);
int nc_inq_dimid(int ncid, const char* name, int* idp);
int nc_inq_dimlen(int ncid, int dimid, size_t* lenp);
int nc_inq_dimname(int ncid, int dimid, char* name);
int nc_inq_varid(int ncid, const char* name, int* varidp);
int nc_inq_vardimid(int ncid, int varid, int* dimidsp);
int nc_get_vara_int(int ncid, int varid, const size_t* startp, const size_t* countp, int* ip);
int nc_def_dim(int ncid, const char* name, size_t len, int* idp);
int nc_def_grp(int parent_ncid, const char* name, int* new_ncid);
int nc_def_var(int ncid, const char* name, int xtype, int ndims, const int* dimidsp, int* varidp);
int nc_put_att_int(int ncid, int varid, const char* name, int xtype, size_t len, const int* op);
int nc_put_vara_int(int ncid, int varid, const size_t* startp, const size_t* countp, const int* op);
int nc_put_var1_int(int ncid, int varid, const size_t* indexp, const int* op);

#ifdef __cplusplus
}
#endif
#endif
