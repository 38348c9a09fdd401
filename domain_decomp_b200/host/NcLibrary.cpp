// NcLibrary.cpp -- see NcLibrary.hpp.  Serial netCDF-C calls (one host process drives the GPUs: no parallel I/O).
#ifdef HAVE_NETCDF
#include "NcLibrary.hpp"

#include <netcdf.h>

#include <stdexcept>

namespace ddc_host {
namespace {
void check(int rc, const char* what)
{
    if (rc != NC_NOERR) // the reference's NC_CHECK (Grid.hpp:243-248) throws the library's message
        throw std::runtime_error(std::string("ERROR: NetCDF: ") + nc_strerror(rc) + " (" + what + ")");
}
#define NCC(call) check((call), #call)

const char* const DIR_CHARS[4] = { "L", "R", "B", "T" };
const char* const DIR_NAMES[4] = { "left", "right", "bottom", "top" };
const char* const DIM_CHARS[2] = { "x", "y" };
} // namespace

NcGridMask nc_read_grid(const std::string& filename, const std::string& xdim, const std::string& ydim,
    const std::vector<int>& dim_order, const std::string& mask_name, bool ignore_mask)
{
    NcGridMask out;
    int ncid = -1, data_id = -1;
    NCC(nc_open(filename.c_str(), NC_NOWRITE, &ncid));
    try {
        // nextSIM restart files keep everything in group "data" (Grid.cpp:58-62); plain grids in the root
        if (nc_inq_ncid(ncid, "data", &data_id) != NC_NOERR)
            data_id = ncid;
        const std::string names[2] = { xdim, ydim };
        int dimid[2];
        size_t len[2];
        for (int d = 0; d < 2; d++) {
            if (nc_inq_dimid(data_id, names[d].c_str(), &dimid[d]) != NC_NOERR)
                NCC(nc_inq_dimid(ncid, names[d].c_str(), &dimid[d]));
            NCC(nc_inq_dimlen(data_id, dimid[d], &len[d]));
        }
        out.nx = (int)len[0];
        out.ny = (int)len[1];
        if (!ignore_mask) {
            int varid;
            NCC(nc_inq_varid(data_id, mask_name.c_str(), &varid));
            int vdims[2];
            NCC(nc_inq_vardimid(data_id, varid, vdims));
            for (int d = 0; d < 2; d++) { // the declared order must be what the caller said (Grid.cpp:104-114)
                char name[257];
                NCC(nc_inq_dimname(data_id, vdims[d], name));
                if (names[dim_order[d]] != name)
                    throw std::runtime_error("Dimension ordering provided does not match ordering in netCDF grid file");
            }
            const size_t start[2] = { 0, 0 }, count[2] = { len[dim_order[0]], len[dim_order[1]] };
            out.mask.resize(len[0] * len[1]);
            NCC(nc_get_vara_int(data_id, varid, start, count, out.mask.data()));
        }
    } catch (...) {
        nc_close(ncid);
        throw;
    }
    NCC(nc_close(ncid));
    return out;
}

void nc_write_mask(const std::string& filename, int nx, int ny, int num_parts, const int* pid)
{
    int ncid, dimid[2], varid;
    NCC(nc_create(filename.c_str(), NC_CLOBBER | NC_NETCDF4, &ncid));
    NCC(nc_put_att_int(ncid, NC_GLOBAL, "num_processes", NC_INT, 1, &num_parts));
    NCC(nc_def_dim(ncid, "y", (size_t)ny, &dimid[0])); // always (y, x) for nextSIM-DG
    NCC(nc_def_dim(ncid, "x", (size_t)nx, &dimid[1]));
    NCC(nc_def_var(ncid, "pid", NC_INT, 2, dimid, &varid));
    NCC(nc_enddef(ncid));
    const size_t start[2] = { 0, 0 }, count[2] = { (size_t)ny, (size_t)nx };
    NCC(nc_put_vara_int(ncid, varid, start, count, pid));
    NCC(nc_close(ncid));
}

void nc_write_metadata(const std::string& filename, int nx, int ny, const std::vector<std::vector<int>>& boxes,
    const std::vector<std::vector<int>>& counts, const std::vector<std::vector<int>>& ids,
    const std::vector<std::vector<int>>& halos, const std::vector<std::vector<int>>& starts)
{
    const size_t P = boxes[0].size();
    int ncid, dim_nx, dim_ny, dim_p, dim_e[8];
    NCC(nc_create(filename.c_str(), NC_CLOBBER | NC_NETCDF4, &ncid));
    // the order of the definitions is the reference's: it is the order `ncdump` prints
    NCC(nc_def_dim(ncid, "NX", (size_t)nx, &dim_nx));
    NCC(nc_def_dim(ncid, "NY", (size_t)ny, &dim_ny));
    NCC(nc_def_dim(ncid, "P", P, &dim_p));
    for (int l = 0; l < 8; l++) { // a length of 0 makes the dimension UNLIMITED, as in the reference's files
        const std::string name = std::string(DIR_CHARS[l & 3]) + (l >= 4 ? "_periodic" : "");
        NCC(nc_def_dim(ncid, name.c_str(), ids[l].size(), &dim_e[l]));
    }
    int g_box, g_con;
    NCC(nc_def_grp(ncid, "bounding_boxes", &g_box));
    NCC(nc_def_grp(ncid, "connectivity", &g_con));
    int v_top[2], v_cnt[2], v_num[8], v_ids[8], v_halo[8], v_start[8];
    for (int d = 0; d < 2; d++) {
        NCC(nc_def_var(g_box, (std::string("domain_") + DIM_CHARS[d]).c_str(), NC_INT, 1, &dim_p, &v_top[d]));
        NCC(nc_def_var(g_box, (std::string("domain_extent_") + DIM_CHARS[d]).c_str(), NC_INT, 1, &dim_p, &v_cnt[d]));
    }
    for (int l = 0; l < 8; l++) {
        const std::string dir = DIR_NAMES[l & 3], sfx = l >= 4 ? "_periodic" : "";
        NCC(nc_def_var(g_con, (dir + "_neighbours" + sfx).c_str(), NC_INT, 1, &dim_p, &v_num[l]));
        NCC(nc_def_var(g_con, (dir + "_neighbour_ids" + sfx).c_str(), NC_INT, 1, &dim_e[l], &v_ids[l]));
        NCC(nc_def_var(g_con, (dir + "_neighbour_halos" + sfx).c_str(), NC_INT, 1, &dim_e[l], &v_halo[l]));
        NCC(nc_def_var(g_con, (dir + "_neighbour_halo_starts" + sfx).c_str(), NC_INT, 1, &dim_e[l], &v_start[l]));
    }
    NCC(nc_enddef(ncid));
    const size_t zero = 0;
    for (int d = 0; d < 2; d++) {
        NCC(nc_put_vara_int(g_box, v_top[d], &zero, &P, boxes[d].data()));
        NCC(nc_put_vara_int(g_box, v_cnt[d], &zero, &P, boxes[2 + d].data()));
    }
    for (int l = 0; l < 8; l++) {
        std::vector<int> c = counts[l];
        c.resize(P, 0);
        NCC(nc_put_vara_int(g_con, v_num[l], &zero, &P, c.data()));
        const size_t n = ids[l].size();
        if (n) {
            NCC(nc_put_vara_int(g_con, v_ids[l], &zero, &n, ids[l].data()));
            NCC(nc_put_vara_int(g_con, v_halo[l], &zero, &n, halos[l].data()));
            NCC(nc_put_vara_int(g_con, v_start[l], &zero, &n, starts[l].data()));
        }
    }
    NCC(nc_close(ncid));
}
} // namespace ddc_host
#endif
