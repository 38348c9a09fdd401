/*!
 * @file CudaRcbPartitioner.cpp
 * @brief Partitioner of the reference tree backed by the B200 library (see CudaRcbPartitioner.hpp).
 *
 * Mode of operation: "every rank asks its GPU".  The reference runs one MPI rank per part; the GPU
 * library computes ALL parts from the whole mask in about a millisecond, so every rank assembles the
 * global mask from the ranks' naive blocks (two collectives), runs the decomposition on its GPU and
 * keeps its own part.  What Zoltan used to deliver -- this rank's box, `changes`, the new owner of
 * every cell of this rank's block -- comes from the library; neighbour discovery and both writers
 * stay the reference's own code.  (One rank per GPU with a row-sharded mask and NCCL / peer-memory
 * exchange is the other mode the C ABI offers; it needs the mask read as row blocks, see
 * INTEGRATION.md.)
 */
#include "CudaRcbPartitioner.hpp"

#include "Utils.hpp"

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <ddc.h>

namespace {
void ddc_check(ddc_handle_s* h, int rc, const char* what)
{
    if (rc != DDC_OK)
        throw std::runtime_error(std::string("ERROR: ") + what + ": " + ddc_last_error(h));
}
} // namespace

CudaRcbPartitioner* CudaRcbPartitioner::create(MPI_Comm comm, int argc, char** argv)
{
    int device = 0;
    for (int i = 1; i + 1 < argc; i++)
        if (std::strcmp(argv[i], "--device") == 0)
            device = std::atoi(argv[i + 1]);
    return new CudaRcbPartitioner(comm, device);
}

CudaRcbPartitioner::CudaRcbPartitioner(MPI_Comm comm, int device)
    : Partitioner(comm)
{
    // replaces Zoltan_Initialize + new Zoltan(comm) (ZoltanPartitioner.cpp:69-86); one GPU, no NCCL
    ddc_handle_t h = nullptr;
    if (ddc_create(&h, device, 0, 1, nullptr) != DDC_OK)
        throw std::runtime_error(std::string("ERROR: ddc_create: ") + ddc_last_error(nullptr));
    _ddc = h;
}

CudaRcbPartitioner::~CudaRcbPartitioner() { ddc_destroy(_ddc); }

void CudaRcbPartitioner::partition(Grid& grid)
{
    // Load initial grid state
    _num_procs = grid.get_num_procs();
    _global_ext = grid.get_global_ext();
    grid.get_bounding_box(_global[0], _global[1], _local_ext[0], _local_ext[1]);
    _px = grid.get_px();
    _py = grid.get_py();
    const int NX = _global_ext[0], NY = _global_ext[1], P = _total_num_procs;
    const int n_own = grid.get_num_objects();
    const bool masked = n_own != grid.get_num_nonzero_objects();

    if (P == 1) { // the reference returns before neighbour discovery (ZoltanPartitioner.cpp:102-121)
        for (int d = 0; d < NDIMS; d++) {
            _global_new[d] = _global[d];
            _local_ext_new[d] = _local_ext[d];
        }
        _proc_id.assign(n_own, masked ? -1 : _rank);
        for (int i = 0; masked && i < n_own; i++)
            if (grid.get_land_mask()[i] > 0)
                _proc_id[i] = _rank;
        return;
    }

    // 1. every rank's naive block (Grid.cpp:150-166) and its slab of the mask -> the whole mask
    int mine[4] = { _global[0], _global[1], _local_ext[0], _local_ext[1] };
    std::vector<int> blocks(4 * (size_t)P);
    CHECK_MPI(MPI_Allgather(mine, 4, MPI_INT, blocks.data(), 4, MPI_INT, _comm));
    std::vector<int> counts(P), displs(P);
    size_t total = 0;
    for (int r = 0; r < P; r++) {
        counts[r] = blocks[4 * r + 2] * blocks[4 * r + 3];
        displs[r] = (int)total;
        total += (size_t)counts[r];
    }
    // a rank without land in its block may not have read a mask at all (--ignore-mask): all ocean
    std::vector<int> slab(n_own, 1);
    if (masked)
        slab.assign(grid.get_land_mask(), grid.get_land_mask() + n_own);
    std::vector<int> slabs(total);
    CHECK_MPI(MPI_Allgatherv(slab.data(), n_own, MPI_INT, slabs.data(), counts.data(), displs.data(), MPI_INT, _comm));
    std::vector<int> mask((size_t)NX * NY, 0);
    for (int r = 0; r < P; r++) {
        const int x0 = blocks[4 * r], y0 = blocks[4 * r + 1], ex = blocks[4 * r + 2], ey = blocks[4 * r + 3];
        for (int j = 0; j < ey; j++)
            std::memcpy(&mask[(size_t)(y0 + j) * NX + x0], &slabs[(size_t)displs[r] + (size_t)j * ex], sizeof(int) * ex);
    }

    // 2. the decomposition: replaces Set_Param x14, the four callbacks, LB_Partition and RCB_Box
    //    (ZoltanPartitioner.cpp:125-195; the ceil / clamp of the boxes and the `changes == 0`
    //    fallback to the naive blocks happen inside the library)
    ddc_check(_ddc, ddc_set_mask_host(_ddc, mask.data(), NX, NY, 0, NY), "ddc_set_mask_host");
    ddc_check(_ddc, ddc_partition(_ddc, P, _px, _py, DDC_WANT_PID), "ddc_partition");
    std::vector<int> bx(P), by(P), bex(P), bey(P);
    ddc_check(_ddc, ddc_get_boxes(_ddc, bx.data(), by.data(), bex.data(), bey.data()), "ddc_get_boxes");
    _global_new = { bx[_rank], by[_rank] };
    _local_ext_new = { bex[_rank], bey[_rank] };

    // 3. Find my neighbours: the reference's own code (Partitioner.cpp:329-435)
    discover_neighbours();

    // 4. the process ids of the grid points of my ORIGINAL block (payload of save_mask,
    //    ZoltanPartitioner.cpp:201-219): the library's owner map, -1 on land
    std::vector<int> pid((size_t)NX * NY);
    ddc_check(_ddc, ddc_get_pid_host(_ddc, pid.data()), "ddc_get_pid_host");
    _proc_id.resize(n_own);
    for (int i = 0; i < n_own; i++) {
        const int x = _global[0] + i % _local_ext[0], y = _global[1] + i / _local_ext[0];
        _proc_id[i] = pid[(size_t)y * NX + x];
    }
}
