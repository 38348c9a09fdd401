// NcLibrary.hpp -- the file formats of the reference through netCDF-C itself (builds with -DHAVE_NETCDF):
// any netCDF grid (classic or netCDF-4 / HDF5, group "data" of nextSIM restart files included) for the input, and
// for the outputs the netCDF-4 files the reference writes -- `partition_mask_<P>.nc` and, with its two groups
// `bounding_boxes` / `connectivity`, `partition_metadata_<P>.nc` (Partitioner.cpp:128-318), which is what
// nextSIM-DG opens at start-up.  Without HAVE_NETCDF the host library reads / writes netCDF classic binaries and
// CDL text on its own (NcClassic.cpp, CdlIO.cpp) and cannot produce the grouped file.
#pragma once
#ifdef HAVE_NETCDF
#include <string>
#include <vector>

#include "HostBuffer.hpp"

namespace ddc_host {
struct NcGridMask {
    int nx = 0, ny = 0;
    IntBuffer mask; // file order, converted like nc_get_vara_int; empty when ignore_mask
};
// Grid.cpp:51-130 of the reference: dimension lengths by name (group "data" first, then the root), the mask
// variable, its declared dimension order checked against dim_order ({1, 0}: (y, x))
NcGridMask nc_read_grid(const std::string& filename, const std::string& xdim, const std::string& ydim,
    const std::vector<int>& dim_order, const std::string& mask_name, bool ignore_mask);
// Partitioner.cpp:128-166
void nc_write_mask(const std::string& filename, int nx, int ny, int num_parts, const int* pid);
// Partitioner.cpp:168-318: boxes[4][P] = x0, y0, extent x, extent y; list l = periodic * 4 + edge
void nc_write_metadata(const std::string& filename, int nx, int ny, const std::vector<std::vector<int>>& boxes,
    const std::vector<std::vector<int>>& counts, const std::vector<std::vector<int>>& ids,
    const std::vector<std::vector<int>>& halos, const std::vector<std::vector<int>>& starts);
} // namespace ddc_host
#endif
