// CdlIO.hpp -- minimal CDL (netCDF text notation) reader and ncdump-style formatter.
// This image has no netCDF library, so CDL text is the in-tree file format: grids are read from
// what `ncdump grid.nc` prints and results are written as what `ncdump result.nc` would print.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "HostBuffer.hpp"

namespace ddc_host {

struct CdlVar {
    std::string type;
    std::vector<std::string> dims;
    std::vector<double> data;
    ddc_host::IntBuffer idata; // filled INSTEAD of `data` by readers asked for ints (large masks: 4 B/value, not 8)
    bool has_data = false;
};
struct CdlGroup {
    std::map<std::string, long> dims;
    std::map<std::string, CdlVar> vars;
    std::map<std::string, CdlGroup> groups;
};
struct CdlFile {
    std::string name;
    CdlGroup root;
};

// throws std::runtime_error("ERROR: ...") on malformed input or unreadable files
CdlFile read_cdl(const std::string& path);

// ncdump-style value list: "v, v, v ;" broken into lines of at most 80 columns.
// `first_prefix` is what already stands on the first line (e.g. "   domain_x = ").
std::string format_values(const std::string& first_prefix, const int* v, size_t n, const std::string& cont_indent);

// "partition_mask_3.nc" / "dir/partition_mask_3.cdl" -> "partition_mask_3"
std::string netcdf_name_of(const std::string& filename);
// replace a trailing ".nc" by ".cdl" (append ".cdl" when there is no such suffix)
std::string cdl_path_of(const std::string& filename);

} // namespace ddc_host
