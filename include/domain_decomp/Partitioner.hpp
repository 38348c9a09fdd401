// Partitioner.hpp -- the partitioner plug-in point, API-compatible with the reference's
// Partitioner.hpp:18-251: the PartitionerType enum, the abstract partition(Grid&), the factory,
// the per-rank getters and the two writers.
//
// Difference in kind: the reference holds ONE box per MPI rank and gathers the others with
// MPI_Allgather; here the concrete partitioner (CudaRcbPartitioner) computes ALL parts on the GPU,
// so the base class stores every part's box and the neighbour tables in CSR form and the per-rank
// getters are views of part `rank`.  The number of parts defaults to the communicator size (the
// reference's only mode) and can be set independently with set_num_parts().
#pragma once

#include <map>
#include <string>
#include <vector>

#include "DomainUtils.hpp"
#include "Grid.hpp"
#include "HostBuffer.hpp"
#include "domain_decomp_export.hpp"

enum class LIB_EXPORT PartitionerType {
    Zoltan_RCB, // kept so that existing callers compile; served by the CUDA implementation
    Cuda_RCB // rectilinear RCB on B200 (CudaRcbPartitioner)
};

class LIB_EXPORT Partitioner {
public:
    Partitioner(const Partitioner&) = delete;
    Partitioner& operator=(const Partitioner&) = delete;
    virtual ~Partitioner() {}

    // Decompose the grid's ocean cells into get_num_parts() rectangular boxes.
    virtual void partition(Grid& grid) = 0;

    // ---- reference API: the view of this rank (= part `rank`) --------------------------------
    void get_bounding_box(int& global_0, int& global_1, int& local_ext_0, int& local_ext_1) const;
    // ids / halo_sizes / halo_starts must hold N_EDGE vectors; results are appended in the order
    // left, right, bottom, top, ids ascending.
    void get_neighbour_info(std::vector<std::vector<int>>& ids, std::vector<std::vector<int>>& halo_sizes,
        std::vector<std::vector<int>>& halo_starts) const;
    void get_neighbour_info_periodic(std::vector<std::vector<int>>& ids,
        std::vector<std::vector<int>>& halo_sizes, std::vector<std::vector<int>>& halo_starts) const;
    // partition_mask / partition_metadata files (netCDF-4 layout of the reference; without a
    // netCDF library the same content is written as the CDL text `ncdump` would print, to
    // <filename with .nc replaced by .cdl>)
    void save_mask(const std::string& filename) const;
    void save_metadata(const std::string& filename) const;

    // ---- all parts (new) ----------------------------------------------------------------------
    void set_num_parts(int nparts); // before partition(); default: communicator size
    int get_num_parts() const { return _num_parts; }
    void get_bounding_box(int part, int& global_0, int& global_1, int& local_ext_0, int& local_ext_1) const;
    void get_neighbour_info(int part, std::vector<std::vector<int>>& ids,
        std::vector<std::vector<int>>& halo_sizes, std::vector<std::vector<int>>& halo_starts) const;
    void get_neighbour_info_periodic(int part, std::vector<std::vector<int>>& ids,
        std::vector<std::vector<int>>& halo_sizes, std::vector<std::vector<int>>& halo_starts) const;
    const ddc_host::IntBuffer& get_partition_ids() const { return _pid_global; } // [NY][NX], -1 on land
    // text of the two files exactly as `ncdump <file>` prints them (used by the golden tests)
    std::string mask_cdl(const std::string& netcdf_name) const;
    std::string metadata_cdl(const std::string& netcdf_name) const;

protected:
    Partitioner(MPI_Comm comm);
    // fills the per-rank members below from the all-parts tables (call at the end of partition())
    void publish_rank_view();

    MPI_Comm _comm;
    int _rank = -1;
    int _total_num_procs = -1;
    int _num_parts = -1;
    static const int NDIMS = 2;
    static const int NNBRS = 2 * NDIMS;
    bool _px = false;
    bool _py = false;

    std::vector<std::string> dim_chars = { "x", "y" };
    std::vector<std::string> dir_chars = { "L", "R", "B", "T" };
    std::vector<std::string> dir_names = { "left", "right", "bottom", "top" };
    std::vector<std::string> global_extent_names = { "NX", "NY" };

    // state mirrored from the Grid (reference: Partitioner.hpp:146-158)
    std::vector<int> _num_procs = std::vector<int>(NDIMS, -1);
    std::vector<int> _global_ext = std::vector<int>(NDIMS, 0);
    std::vector<int> _local_ext = std::vector<int>(NDIMS, 0);
    std::vector<int> _global = std::vector<int>(NDIMS, -1);
    // this rank's box after partitioning, its pid slab and neighbour maps (reference layout)
    std::vector<int> _local_ext_new = std::vector<int>(NDIMS, 0);
    std::vector<int> _global_new = std::vector<int>(NDIMS, -1);
    std::vector<int> _proc_id = {};
    std::vector<std::map<int, int>> _neighbours = std::vector<std::map<int, int>>(NNBRS);
    std::vector<std::map<int, int>> _halo_starts = std::vector<std::map<int, int>>(NNBRS);
    std::vector<std::map<int, int>> _neighbours_p = std::vector<std::map<int, int>>(NNBRS);
    std::vector<std::map<int, int>> _halo_starts_p = std::vector<std::map<int, int>>(NNBRS);

    // every part: boxes[4][P] = x0, y0, ext_x, ext_y; neighbour tables per list l = periodic*4 + edge
    std::vector<std::vector<int>> _boxes = std::vector<std::vector<int>>(4);
    std::vector<std::vector<int>> _nbr_counts = std::vector<std::vector<int>>(2 * NNBRS);
    std::vector<std::vector<int>> _nbr_offsets = std::vector<std::vector<int>>(2 * NNBRS);
    std::vector<std::vector<int>> _nbr_ids = std::vector<std::vector<int>>(2 * NNBRS);
    std::vector<std::vector<int>> _nbr_halos = std::vector<std::vector<int>>(2 * NNBRS);
    std::vector<std::vector<int>> _nbr_starts = std::vector<std::vector<int>>(2 * NNBRS);
    ddc_host::IntBuffer _pid_global = {}; // page-locked when large: the device writes it over PCIe

public:
    struct LIB_EXPORT Factory {
        // argc / argv are forwarded to the concrete partitioner (it understands `--parts N`
        // and `--device D`); throws std::runtime_error("Invalid partitioner!") for unknown types
        static Partitioner* create(MPI_Comm comm, int argc, char** argv, PartitionerType type);
    };
};
