// host_netcdf_shim.cpp -- TEST INFRASTRUCTURE.  THIS repository's host layer (domain_decomp_b200/host, the same
// sources as the product) built with -DHAVE_NETCDF, i.e. with its netCDF-C backend (host/NcLibrary.cpp), on the
// in-memory netCDF of oracle/ref_shim/netcdf_mem.hpp and with the CPU oracle answering the ddc_* calls
// (ddc_oracle_stub.c).  host_nc_run() reads a grid through Grid::create, partitions, writes partition_mask.nc and
// partition_metadata.nc with the library calls and reports both files in the format of ref_host_run
// (oracle/ref_hostpath_shim.cpp), so that tests/test_host_layer_cpu.py can compare them with what the REFERENCE's
// own Partitioner.cpp writes for the same decomposition: dimensions (zero length = UNLIMITED included), groups,
// variables, attributes, values.  Not part of the product.
#include <functional>
#include <memory>

#include "netcdf_mem.hpp"

#include "Grid.hpp"
#include "Partitioner.hpp"

extern "C" {
int nc_open(const char* path, int, int* ncidp)
{
    std::lock_guard<std::mutex> lk(g_fs_mutex);
    const int f = find_file(path);
    if (f < 0)
        return NC_ENOENT;
    *ncidp = f << 8;
    return NC_NOERR;
}
int nc_create(const char* path, int, int* ncidp)
{
    std::lock_guard<std::mutex> lk(g_fs_mutex);
    int f = find_file(path);
    if (f >= 0)
        g_files[f] = MemFile(); // NC_CLOBBER
    else {
        g_files.push_back(MemFile());
        f = (int)g_files.size() - 1;
    }
    g_files[f].path = path;
    *ncidp = f << 8;
    return NC_NOERR;
}
}

namespace {
std::string g_report, g_error;
}

extern "C" {
// mask[ny][nx] is put into an in-memory grid.nc (dims xdim / ydim, variable maskname declared (ydim, xdim) or,
// file_order_xy, (xdim, ydim); data_group: inside group "data"), read back through Grid::create and decomposed
// into P parts.  Report: "grid nx ny", "gridmask v v ...", then the file / dim / att / var lines of both outputs.
__attribute__((visibility("default"))) const char* host_nc_run(int P, int nx, int ny, const int* mask, const char* xdim,
    const char* ydim, const char* maskname, int order_xy, int file_order_xy, int data_group, int ignore_mask, int px, int py)
{
    g_error.clear();
    g_report.clear();
    {
        std::lock_guard<std::mutex> lk(g_fs_mutex);
        g_files.clear();
        MemFile in;
        in.path = "grid.nc";
        in.dims = { { xdim, (size_t)nx }, { ydim, (size_t)ny } };
        if (data_group)
            in.groups.push_back("data");
        MemVar v;
        v.name = maskname;
        v.group = data_group ? 1 : 0;
        v.dimids = file_order_xy ? std::vector<int> { 0, 1 } : std::vector<int> { 1, 0 };
        v.data.assign(mask, mask + (size_t)nx * ny);
        v.written.assign(v.data.size(), 1);
        in.vars.push_back(v);
        g_files.push_back(in);
    }
    try {
        const std::vector<int> order = order_xy ? std::vector<int> { 0, 1 } : std::vector<int> { 1, 0 };
        std::unique_ptr<Grid> grid(Grid::create(ddc_shim_comm(0, 1), "grid.nc", xdim, ydim, order, maskname, ignore_mask != 0,
            px != 0, py != 0));
        std::ostringstream os;
        const std::vector<int> ext = grid->get_global_ext();
        os << "grid " << ext[0] << " " << ext[1] << "\ngridmask";
        for (size_t i = 0; i < (size_t)ext[0] * ext[1]; i++)
            os << " " << grid->get_global_land_mask()[i];
        os << "\n";
        std::unique_ptr<Partitioner> part(Partitioner::Factory::create(ddc_shim_comm(0, 1), 0, nullptr, PartitionerType::Cuda_RCB));
        part->set_num_parts(P);
        part->partition(*grid);
        part->save_mask("partition_mask.nc");
        part->save_metadata("partition_metadata.nc");
        {
            std::lock_guard<std::mutex> lk(g_fs_mutex);
            dump_file(os, "mask", "partition_mask.nc");
            dump_file(os, "metadata", "partition_metadata.nc");
        }
        g_report = os.str();
        return g_report.c_str();
    } catch (const std::exception& e) {
        g_error = e.what();
        return nullptr;
    }
}
__attribute__((visibility("default"))) const char* host_nc_error(void) { return g_error.c_str(); }
}
