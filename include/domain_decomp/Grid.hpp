// Grid.hpp -- the input side of the decomposer: extents, land-sea mask, naive block layout.
// API-compatible with the reference's Grid.hpp:30-241 (same factory functions, same getters, same
// ownership rules), re-implemented for a single host process that feeds GPUs: besides the
// per-rank block of the reference it keeps the GLOBAL mask, which is what the CUDA partitioner
// stages into HBM.
#pragma once

#ifdef DDC_HAVE_MPI
#include <mpi.h>
#else
#include "mpi_shim/mpi.h"
#endif
#include <string>
#include <vector>

#include "HostBuffer.hpp"
#include "domain_decomp_export.hpp"

class LIB_EXPORT Grid {
public:
    Grid(const Grid&) = delete;
    Grid& operator=(const Grid&) = delete;
    ~Grid() {}

    // Named constructors (heap only, delete before MPI_Finalize -- as the reference).
    //   filename    grid file: netCDF when the library was built with DDC_HAVE_NETCDF,
    //               otherwise CDL text (what `ncdump` prints)
    //   xdim_name / ydim_name / mask_name   names inside the file (defaults "x", "y", "mask")
    //   dim_order   {1, 0}: variables are dimensioned (y, x);  {0, 1}: (x, y)
    //   ignore_mask treat every cell as ocean
    //   px, py      periodic in x / y (only affects the periodic neighbour tables)
    static Grid* create(MPI_Comm comm, const std::string& filename, bool ignore_mask = false,
        bool px = false, bool py = false);
    static Grid* create(MPI_Comm comm, const std::string& filename, const std::string xdim_name,
        const std::string ydim_name, const std::vector<int> dim_order, const std::string mask_name,
        bool ignore_mask = false, bool px = false, bool py = false);
    // In-memory variant (not in the reference): mask[ny][nx], x fastest.
    static Grid* create_from_mask(MPI_Comm comm, const int* mask, int nx, int ny, bool px = false,
        bool py = false);

    // ---- the rank's block of the naive 2-D decomposition (what the reference exposes) ----
    int get_num_objects() const; // cells of the block
    int get_num_nonzero_objects() const; // ocean cells of the block
    std::vector<int> get_num_procs() const; // blocks per dimension
    std::vector<int> get_local_ext() const;
    std::vector<int> get_global() const; // first cell of the block
    std::vector<int> get_global_ext() const; // {NX, NY}
    const int* get_land_mask() const; // block mask, x fastest (borrowed)
    const int* get_sparse_to_dense() const; // ocean index -> block-local dense index (borrowed)
    const int* get_nonzero_object_ids() const; // ocean index -> global id y * NX + x (borrowed)
    void get_bounding_box(int& global_0, int& global_1, int& local_ext_0, int& local_ext_1) const;
    bool get_px() const;
    bool get_py() const;

    // ---- additions for the CUDA partitioner ----
    const int* get_global_land_mask() const; // [NY][NX], x fastest (borrowed)
    bool mask_ignored() const { return _ignore_mask; }
    MPI_Comm get_comm() const { return _comm; }

    static const int NDIMS = 2;

private:
    Grid(MPI_Comm comm, bool px, bool py);
    void load_file(const std::string& filename, const std::string& xdim, const std::string& ydim,
        const std::vector<int>& order, const std::string& mask_name, bool ignore_mask);
    void build_block();
    void build_block_arrays() const; // the block's mask slab and ocean id lists, on first use

    MPI_Comm _comm;
    int _rank = -1;
    int _total_num_procs = -1;
    std::vector<int> _num_procs = std::vector<int>(NDIMS, -1);
    std::vector<int> _global_ext = std::vector<int>(NDIMS, 0);
    std::vector<int> _local_ext = std::vector<int>(NDIMS, 0);
    std::vector<int> _global = std::vector<int>(NDIMS, -1);
    int _num_objects = 0;
    mutable int _num_nonzero_objects = 0;
    mutable bool _block_built = false;
    bool _px = false, _py = false, _ignore_mask = false;
    ddc_host::IntBuffer _global_mask; // the whole mask (page-locked when large: it is what crosses PCIe)
    mutable std::vector<int> _land_mask; // the rank's block (built on first use)
    mutable std::vector<int> _local_id, _global_id;
};
