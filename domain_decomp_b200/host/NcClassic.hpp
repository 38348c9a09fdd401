// NcClassic.hpp -- reader / writer of the netCDF CLASSIC file formats (CDF-1, CDF-2 "64-bit offset",
// CDF-5 "64-bit data"), written from the published format grammar; no netCDF library is involved.
//
// Why: the reference reads its grid with netCDF-C (Grid.cpp:51-130) and writes partition_mask_<P>.nc
// with it (Partitioner.cpp:128-166).  This image has neither libnetcdf nor HDF5.  The classic formats
// are simple enough to implement directly and every netCDF reader (nextSIM-DG's included) opens
// them, so
//   * grids in classic format -- what `ncgen -b` makes of the reference's test/*.cdl inputs -- are
//     read as they are (a netCDF-4 / HDF5 file is recognised and refused with a hint);
//   * partition_mask_<P>.nc is written as a real netCDF file (same dimensions, variable, attribute
//     and values as the reference's netCDF-4 file: `ncdump` prints the same CDL).
// partition_metadata_<P>.nc uses groups, which only netCDF-4 / HDF5 has: it stays CDL text (CdlIO).
#pragma once
#include "CdlIO.hpp"

#include <cstdint>
#include <string>
#include <vector>

namespace ddc_host {

enum class FileKind { NetcdfClassic, Hdf5, Text };
// looks at the first bytes; throws std::runtime_error when the file cannot be opened
FileKind sniff_file_kind(const std::string& path);

// Header and data of a classic file as a CdlFile (one root group: classic files have no groups).
// only_var: when non-empty, only this variable's data is loaded (the others keep has_data = false).
// as_int: values are converted like nc_get_vara_int does (truncation toward zero) and stored in
// CdlVar::idata instead of CdlVar::data -- half the memory for a mask of 10^9 cells.
CdlFile read_netcdf_classic(const std::string& path, const std::string& only_var = "", bool as_int = false);

struct NcDim {
    std::string name;
    uint64_t len;
};
struct NcIntAttr {
    std::string name;
    int32_t value;
};
struct NcIntVar {
    std::string name;
    std::vector<int> dimids;
    const int32_t* data; // borrowed; product of the dimension lengths values
};
// Writes dims, global NC_INT attributes and NC_INT variables.  version: 1, 2 or 5; 0 = the smallest
// one that can hold the data (CDF-1 below 2 GiB, CDF-5 for a variable of 4 GiB or more, else CDF-2).
void write_netcdf_classic(const std::string& path, const std::vector<NcDim>& dims,
    const std::vector<NcIntAttr>& global_attrs, const std::vector<NcIntVar>& vars, int version = 0);

} // namespace ddc_host
