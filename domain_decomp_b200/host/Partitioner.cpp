// Partitioner.cpp -- base class of the partitioner plug-ins: per-part queries and the two writers.
//
// The reference gathers one box per rank with MPI_Allgather and discovers neighbours on the host
// (Partitioner.cpp:329-435); here the concrete partitioner already delivers all boxes and the
// neighbour tables (built by the interval-intersection kernel), so this file only serves views of
// them and reproduces the wire format of save_mask / save_metadata (Partitioner.cpp:128-318).
#include "Partitioner.hpp"

#include "CdlIO.hpp"
#include "NcClassic.hpp"
#include "NcLibrary.hpp"
#include "CudaRcbPartitioner.hpp"

#include <fstream>
#include <iostream>
#include <stdexcept>

Partitioner::Partitioner(MPI_Comm comm)
    : _comm(comm)
{
    MPI_Comm_size(comm, &_total_num_procs);
    MPI_Comm_rank(comm, &_rank);
    _num_parts = _total_num_procs;
}

void Partitioner::set_num_parts(int nparts)
{
    if (nparts < 1)
        throw std::runtime_error("ERROR: the number of parts must be positive");
    _num_parts = nparts;
}

Partitioner* Partitioner::Factory::create(MPI_Comm comm, int argc, char** argv, PartitionerType type)
{
    // Zoltan_RCB is accepted so that existing callers keep working: the same decomposition is
    // produced by the CUDA implementation.
    if (type == PartitionerType::Zoltan_RCB || type == PartitionerType::Cuda_RCB)
        return CudaRcbPartitioner::create(comm, argc, argv);
    throw std::runtime_error("Invalid partitioner!");
}

// ---- per-part queries ---------------------------------------------------------------------------

void Partitioner::get_bounding_box(int part, int& global_0, int& global_1, int& local_ext_0, int& local_ext_1) const
{
    if (part < 0 || part >= (int)_boxes[0].size())
        throw std::runtime_error("ERROR: part index out of range (call partition() first)");
    global_0 = _boxes[0][part];
    global_1 = _boxes[1][part];
    local_ext_0 = _boxes[2][part];
    local_ext_1 = _boxes[3][part];
}

void Partitioner::get_bounding_box(int& global_0, int& global_1, int& local_ext_0, int& local_ext_1) const
{
    global_0 = _global_new[0];
    global_1 = _global_new[1];
    local_ext_0 = _local_ext_new[0];
    local_ext_1 = _local_ext_new[1];
}

namespace {
void append_lists(int part, int first_list, const std::vector<std::vector<int>>& counts,
    const std::vector<std::vector<int>>& offsets, const std::vector<std::vector<int>>& a,
    const std::vector<std::vector<int>>& b, const std::vector<std::vector<int>>& c,
    std::vector<std::vector<int>>& ids, std::vector<std::vector<int>>& halo_sizes,
    std::vector<std::vector<int>>& halo_starts)
{
    for (int e = 0; e < N_EDGE; e++) {
        const int l = first_list + e;
        if (part < 0 || part >= (int)counts[l].size())
            continue;
        const int o = offsets[l][part], n = counts[l][part];
        ids[e].insert(ids[e].end(), a[l].begin() + o, a[l].begin() + o + n);
        halo_sizes[e].insert(halo_sizes[e].end(), b[l].begin() + o, b[l].begin() + o + n);
        halo_starts[e].insert(halo_starts[e].end(), c[l].begin() + o, c[l].begin() + o + n);
    }
}
} // namespace

void Partitioner::get_neighbour_info(int part, std::vector<std::vector<int>>& ids,
    std::vector<std::vector<int>>& halo_sizes, std::vector<std::vector<int>>& halo_starts) const
{
    append_lists(part, 0, _nbr_counts, _nbr_offsets, _nbr_ids, _nbr_halos, _nbr_starts, ids, halo_sizes, halo_starts);
}

void Partitioner::get_neighbour_info_periodic(int part, std::vector<std::vector<int>>& ids,
    std::vector<std::vector<int>>& halo_sizes, std::vector<std::vector<int>>& halo_starts) const
{
    // the periodic tables only ever hold L/R entries when px and B/T entries when py
    append_lists(part, N_EDGE, _nbr_counts, _nbr_offsets, _nbr_ids, _nbr_halos, _nbr_starts, ids, halo_sizes, halo_starts);
}

void Partitioner::get_neighbour_info(std::vector<std::vector<int>>& ids,
    std::vector<std::vector<int>>& halo_sizes, std::vector<std::vector<int>>& halo_starts) const
{
    for (auto edge : edges) {
        for (const auto& kv : _neighbours[edge]) {
            ids[edge].push_back(kv.first);
            halo_sizes[edge].push_back(kv.second);
        }
        for (const auto& kv : _halo_starts[edge])
            halo_starts[edge].push_back(kv.second);
    }
}

void Partitioner::get_neighbour_info_periodic(std::vector<std::vector<int>>& ids,
    std::vector<std::vector<int>>& halo_sizes, std::vector<std::vector<int>>& halo_starts) const
{
    for (auto edge : edges) {
        const bool horizontal = edge == LEFT || edge == RIGHT;
        if ((horizontal && !_px) || (!horizontal && !_py))
            continue;
        for (const auto& kv : _neighbours_p[edge]) {
            ids[edge].push_back(kv.first);
            halo_sizes[edge].push_back(kv.second);
        }
        for (const auto& kv : _halo_starts_p[edge])
            halo_starts[edge].push_back(kv.second);
    }
}

void Partitioner::publish_rank_view()
{
    for (int e = 0; e < NNBRS; e++) {
        _neighbours[e].clear();
        _halo_starts[e].clear();
        _neighbours_p[e].clear();
        _halo_starts_p[e].clear();
    }
    const int P = (int)_boxes[0].size();
    if (_rank >= 0 && _rank < P) {
        _global_new = { _boxes[0][_rank], _boxes[1][_rank] };
        _local_ext_new = { _boxes[2][_rank], _boxes[3][_rank] };
        for (int l = 0; l < 2 * NNBRS; l++) {
            if (_rank >= (int)_nbr_counts[l].size())
                continue;
            const int o = _nbr_offsets[l][_rank], n = _nbr_counts[l][_rank];
            auto& sizes = l < NNBRS ? _neighbours[l] : _neighbours_p[l - NNBRS];
            auto& starts = l < NNBRS ? _halo_starts[l] : _halo_starts_p[l - NNBRS];
            for (int i = o; i < o + n; i++) {
                sizes[_nbr_ids[l][i]] = _nbr_halos[l][i];
                starts[_nbr_ids[l][i]] = _nbr_starts[l][i];
            }
        }
    } else {
        _global_new = { 0, 0 };
        _local_ext_new = { 0, 0 };
    }
    // pid slab of this rank's naive block, the payload of the reference's save_mask (save_mask here writes from the
    // global map; for a single-rank communicator the slab IS the global map and is not copied: 4 GiB at 32768^2)
    _proc_id.clear();
    if (_total_num_procs > 1 && !_pid_global.empty() && _local_ext[0] > 0 && _local_ext[1] > 0) {
        const int NX = _global_ext[0], NY = _global_ext[1];
        _proc_id.reserve((size_t)_local_ext[0] * _local_ext[1]);
        for (int j = 0; j < _local_ext[1]; j++)
            for (int i = 0; i < _local_ext[0]; i++) {
                const int gx = _global[0] + i, gy = _global[1] + j;
                _proc_id.push_back(gx < NX && gy < NY ? _pid_global[(size_t)gy * NX + gx] : -1);
            }
    }
}

// ---- writers ----------------------------------------------------------------------------------------

std::string Partitioner::mask_cdl(const std::string& netcdf_name) const
{
    const int NX = _global_ext[0], NY = _global_ext[1];
    if (_pid_global.size() != (size_t)NX * NY)
        throw std::runtime_error("ERROR: no partition ids (call partition() first)");
    std::string s = "netcdf " + netcdf_name + " {\ndimensions:\n";
    s += "\ty = " + std::to_string(NY) + " ;\n\tx = " + std::to_string(NX) + " ;\n";
    s += "variables:\n\tint pid(y, x) ;\n\n// global attributes:\n";
    s += "\t\t:num_processes = " + std::to_string(_num_parts) + " ;\ndata:\n\n pid =\n";
    for (int y = 0; y < NY; y++) {
        // one grid row per text row; the last value of the variable is followed by " ;"
        std::string row = "  ";
        size_t col = 2;
        for (int x = 0; x < NX; x++) {
            std::string tok = std::to_string(_pid_global[(size_t)y * NX + x]);
            const bool last = y == NY - 1 && x == NX - 1;
            tok += last ? " ;" : ",";
            if (col + tok.size() > 79 && col > 4) {
                row += "\n    ";
                col = 4;
            }
            row += tok;
            col += tok.size();
            if (x + 1 < NX) {
                row += " ";
                col++;
            }
        }
        s += row + "\n";
    }
    s += "}\n";
    return s;
}

std::string Partitioner::metadata_cdl(const std::string& netcdf_name) const
{
    const int P = (int)_boxes[0].size();
    if (P == 0)
        throw std::runtime_error("ERROR: no partition (call partition() first)");
    auto total = [&](int l) { return (long)_nbr_ids[l].size(); };
    auto dim_line = [](const std::string& name, long n) {
        return "\t" + name + " = " + (n > 0 ? std::to_string(n) + " ;" : std::string("UNLIMITED ; // (0 currently)")) + "\n";
    };
    std::string s = "netcdf " + netcdf_name + " {\ndimensions:\n";
    s += dim_line(global_extent_names[0], _global_ext[0]);
    s += dim_line(global_extent_names[1], _global_ext[1]);
    s += dim_line("P", P);
    for (int e = 0; e < N_EDGE; e++)
        s += dim_line(dir_chars[e], total(e));
    for (int e = 0; e < N_EDGE; e++)
        s += dim_line(dir_chars[e] + "_periodic", total(N_EDGE + e));

    auto data_stmt = [](const std::string& name, const std::vector<int>& v) {
        if (v.empty())
            return std::string(); // ncdump prints nothing for a variable without elements
        return "\n" + ddc_host::format_values("   " + name + " = ", v.data(), v.size(), "    ") + "\n";
    };

    s += "\ngroup: bounding_boxes {\n  variables:\n";
    for (int d = 0; d < NDIMS; d++) {
        s += "  \tint domain_" + dim_chars[d] + "(P) ;\n";
        s += "  \tint domain_extent_" + dim_chars[d] + "(P) ;\n";
    }
    s += "  data:\n";
    for (int d = 0; d < NDIMS; d++) {
        s += data_stmt("domain_" + dim_chars[d], _boxes[d]);
        s += data_stmt("domain_extent_" + dim_chars[d], _boxes[2 + d]);
    }
    s += "  } // group bounding_boxes\n";

    s += "\ngroup: connectivity {\n  variables:\n";
    for (int per = 0; per < 2; per++) {
        const std::string sfx = per ? "_periodic" : "";
        for (int e = 0; e < N_EDGE; e++) {
            const std::string dim = dir_chars[e] + sfx;
            s += "  \tint " + dir_names[e] + "_neighbours" + sfx + "(P) ;\n";
            s += "  \tint " + dir_names[e] + "_neighbour_ids" + sfx + "(" + dim + ") ;\n";
            s += "  \tint " + dir_names[e] + "_neighbour_halos" + sfx + "(" + dim + ") ;\n";
            s += "  \tint " + dir_names[e] + "_neighbour_halo_starts" + sfx + "(" + dim + ") ;\n";
        }
    }
    s += "  data:\n";
    for (int per = 0; per < 2; per++) {
        const std::string sfx = per ? "_periodic" : "";
        for (int e = 0; e < N_EDGE; e++) {
            const int l = per * N_EDGE + e;
            std::vector<int> counts = _nbr_counts[l];
            if (counts.empty())
                counts.assign(P, 0);
            s += data_stmt(dir_names[e] + "_neighbours" + sfx, counts);
            s += data_stmt(dir_names[e] + "_neighbour_ids" + sfx, _nbr_ids[l]);
            s += data_stmt(dir_names[e] + "_neighbour_halos" + sfx, _nbr_halos[l]);
            s += data_stmt(dir_names[e] + "_neighbour_halo_starts" + sfx, _nbr_starts[l]);
        }
    }
    s += "  } // group connectivity\n}\n";
    return s;
}

namespace {
void write_text(const std::string& path, const std::string& text)
{
    std::ofstream out(path, std::ios::binary);
    if (!out)
        throw std::runtime_error("ERROR: cannot write '" + path + "'");
    out << text;
}
} // namespace

void Partitioner::save_mask(const std::string& filename) const
{
    // every rank of the reference writes its slab collectively; here rank 0 holds the whole map
    if (_rank != 0)
        return;
    // A real netCDF file when a ".nc" name was asked for: classic format (the reference writes
    // netCDF-4; dimensions, variable, attribute and values are the same, so `ncdump` prints the same
    // CDL and every netCDF reader opens it) -- Partitioner.cpp:128-166.  The CDL text `ncdump` would
    // print is written beside it, except for maps of more than 2^24 cells next to a .nc file
    // (hundreds of MB of text nobody reads).
    const int NX = _global_ext[0], NY = _global_ext[1];
    const bool binary = filename.size() > 3 && filename.compare(filename.size() - 3, 3, ".nc") == 0;
#ifdef HAVE_NETCDF
    if (binary) { // the reference's own format: netCDF-4 through netCDF-C
        if (_pid_global.size() != (size_t)NX * NY)
            throw std::runtime_error("ERROR: no partition (call partition() first)");
        ddc_host::nc_write_mask(filename, NX, NY, _num_parts, _pid_global.data());
        return;
    }
#endif
    if (!binary || (size_t)NX * NY <= ((size_t)1 << 24))
        write_text(ddc_host::cdl_path_of(filename), mask_cdl(ddc_host::netcdf_name_of(filename)));
    if (binary) {
        ddc_host::write_netcdf_classic(filename, { { "y", (uint64_t)NY }, { "x", (uint64_t)NX } },
            { { "num_processes", _num_parts } }, { { "pid", { 0, 1 }, _pid_global.data() } });
    }
}

void Partitioner::save_metadata(const std::string& filename) const
{
    if (_rank != 0)
        return;
#ifdef HAVE_NETCDF
    if (filename.size() > 3 && filename.compare(filename.size() - 3, 3, ".nc") == 0) {
        // netCDF-4 with the groups bounding_boxes / connectivity, as nextSIM-DG reads it (Partitioner.cpp:168-318)
        if (_boxes[0].empty())
            throw std::runtime_error("ERROR: no partition (call partition() first)");
        ddc_host::nc_write_metadata(filename, _global_ext[0], _global_ext[1], _boxes, _nbr_counts, _nbr_ids, _nbr_halos,
            _nbr_starts);
        return;
    }
#endif
    // Groups need netCDF-4 / HDF5, which this build does not have: the CDL text `ncdump` prints for the reference's
    // file goes to <name>.cdl; `ncgen -k nc4 -b <name>.cdl` turns it into the file nextSIM-DG opens.
    const std::string cdl = ddc_host::cdl_path_of(filename);
    write_text(cdl, metadata_cdl(ddc_host::netcdf_name_of(filename)));
    if (cdl != filename)
        std::cerr << "NOTE: built without netCDF-C: wrote " << cdl << " (CDL text) instead of " << filename
                  << "; convert with `ncgen -k nc4 -b " << cdl << "` or rebuild with -DDDC_WITH_NETCDF=ON" << std::endl;
}
