/* netcdf_par.h -- TEST INFRASTRUCTURE, see netcdf.h in this directory. */
#ifndef DDC_REF_SHIM_NETCDF_PAR_H
#define DDC_REF_SHIM_NETCDF_PAR_H
#include <mpi.h>
#include <netcdf.h>

#ifdef __cplusplus
extern "C" {
#endif
#define NC_INDEPENDENT 0
#define NC_COLLECTIVE 1
int nc_open_par(const char* path, int mode, MPI_Comm comm, MPI_Info info, int* ncidp);
int nc_create_par(const char* path, int cmode, MPI_Comm comm, MPI_Info info, int* ncidp);
int nc_var_par_access(int ncid, int varid, int par_access);
#ifdef __cplusplus
}
#endif
#endif
