/*!
 * @file CudaRcbPartitioner.cpp
 * @brief Partitioner of the reference tree backed by the B200 library (see CudaRcbPartitioner.hpp).
 *
 * Mode of operation: "rank 0 asks the GPU".  The reference runs one MPI rank per part; the GPU library
 * computes ALL parts from the whole mask in about a millisecond.  So the ranks' naive blocks of the mask are
 * gathered on rank 0 (MPI_Gatherv), rank 0 -- the only rank that ever touches the GPU -- runs the
 * decomposition once, the P boxes are broadcast (MPI_Bcast) and every rank receives the owners of the cells
 * of its own block (MPI_Scatterv).  What Zoltan used to deliver -- this rank's box, `changes`, the new owner of
 * every cell of this rank's block -- comes from the library; neighbour discovery and both writers stay the
 * reference's own code.  (One rank per GPU with a row-sharded mask and the peer-memory / NCCL exchange is the
 * other mode the C ABI offers; it needs the mask read as row blocks, see INTEGRATION.md.)
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
    // replaces Zoltan_Initialize + new Zoltan(comm) (ZoltanPartitioner.cpp:69-86); one GPU, no NCCL: only rank 0
    // holds a handle
    if (_rank == 0) {
        ddc_handle_t h = nullptr;
        if (ddc_create(&h, device, 0, 1, nullptr) != DDC_OK)
            throw std::runtime_error(std::string("ERROR: ddc_create: ") + ddc_last_error(nullptr));
        _ddc = h;
    }
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

    // 1. every rank's naive block (Grid.cpp:150-166) and its slab of the mask -> the whole mask, on rank 0
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
    std::vector<int> slabs(_rank == 0 ? total : 0);
    CHECK_MPI(MPI_Gatherv(slab.data(), n_own, MPI_INT, slabs.data(), counts.data(), displs.data(), MPI_INT, 0, _comm));

    // 2. the decomposition, once, on rank 0: replaces Set_Param x14, the four callbacks, LB_Partition and RCB_Box
    //    (ZoltanPartitioner.cpp:125-195; the ceil / clamp of the boxes and the `changes == 0` fallback to the
    //    naive blocks happen inside the library).  boxes[4 P] = x0[P] y0[P] ex[P] ey[P]; `status` travels in front
    //    so that a failure on rank 0 is an exception on every rank
    std::vector<int> boxes(1 + 4 * (size_t)P, 0), pid_blocks(_rank == 0 ? total : 0);
    std::string failure;
    if (_rank == 0) {
        try {
            std::vector<int> mask((size_t)NX * NY, 0);
            for (int r = 0; r < P; r++) {
                const int x0 = blocks[4 * r], y0 = blocks[4 * r + 1], ex = blocks[4 * r + 2], ey = blocks[4 * r + 3];
                for (int j = 0; j < ey; j++)
                    std::memcpy(&mask[(size_t)(y0 + j) * NX + x0], &slabs[(size_t)displs[r] + (size_t)j * ex], sizeof(int) * ex);
            }
            ddc_check(_ddc, ddc_set_mask_host(_ddc, mask.data(), NX, NY, 0, NY), "ddc_set_mask_host");
            ddc_check(_ddc, ddc_partition(_ddc, P, _px, _py, DDC_WANT_PID), "ddc_partition");
            int* b = boxes.data() + 1;
            ddc_check(_ddc, ddc_get_boxes(_ddc, b, b + P, b + 2 * P, b + 3 * P), "ddc_get_boxes");
            // the owner map (-1 on land), cut into the ranks' ORIGINAL blocks: what MPI_Scatterv hands out
            std::vector<int>& pid = mask; // (the mask is consumed: reuse its storage)
            ddc_check(_ddc, ddc_get_pid_host(_ddc, pid.data()), "ddc_get_pid_host");
            for (int r = 0; r < P; r++) {
                const int x0 = blocks[4 * r], y0 = blocks[4 * r + 1], ex = blocks[4 * r + 2], ey = blocks[4 * r + 3];
                for (int j = 0; j < ey; j++)
                    std::memcpy(&pid_blocks[(size_t)displs[r] + (size_t)j * ex], &pid[(size_t)(y0 + j) * NX + x0], sizeof(int) * ex);
            }
            boxes[0] = 1;
        } catch (const std::exception& e) {
            failure = e.what();
        }
    }
    CHECK_MPI(MPI_Bcast(boxes.data(), 1 + 4 * P, MPI_INT, 0, _comm));
    if (boxes[0] != 1)
        throw std::runtime_error(_rank == 0 ? failure : std::string("ERROR: the decomposition failed on rank 0"));
    _global_new = { boxes[1 + _rank], boxes[1 + P + _rank] };
    _local_ext_new = { boxes[1 + 2 * P + _rank], boxes[1 + 3 * P + _rank] };

    // 3. Find my neighbours: the reference's own code (Partitioner.cpp:329-435)
    discover_neighbours();

    // 4. the process ids of the grid points of my ORIGINAL block (payload of save_mask,
    //    ZoltanPartitioner.cpp:201-219)
    _proc_id.resize(n_own);
    CHECK_MPI(MPI_Scatterv(pid_blocks.data(), counts.data(), displs.data(), MPI_INT, _proc_id.data(), n_own, MPI_INT, 0, _comm));
}
