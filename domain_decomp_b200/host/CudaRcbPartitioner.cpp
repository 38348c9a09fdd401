// CudaRcbPartitioner.cpp -- ZoltanPartitioner's replacement: a thin host wrapper over the C ABI.
// It owns no arithmetic: boxes, pid and the neighbour tables all come from libddc_cuda.
#include "CudaRcbPartitioner.hpp"

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

namespace {
[[noreturn]] void raise(ddc_handle_t h, const char* what)
{
    throw std::runtime_error(std::string("ERROR: ") + what + ": " + ddc_last_error(h));
}
} // namespace

CudaRcbPartitioner::CudaRcbPartitioner(MPI_Comm comm, int argc, char** argv)
    : Partitioner(comm)
{
    int device = 0;
    for (int i = 1; i + 1 < argc; i++) {
        if (!argv || !argv[i])
            continue;
        if (std::strcmp(argv[i], "--parts") == 0)
            set_num_parts(std::atoi(argv[i + 1]));
        else if (std::strcmp(argv[i], "--device") == 0)
            device = std::atoi(argv[i + 1]);
    }
    // one host process drives one GPU; multi-GPU row sharding is reached through the C ABI with
    // one process per GPU
    if (ddc_create(&_h, device, 0, 1, nullptr) != DDC_OK)
        raise(nullptr, "cannot create the CUDA partitioner (there is no CPU fallback)");
}

CudaRcbPartitioner::~CudaRcbPartitioner() { ddc_destroy(_h); }

CudaRcbPartitioner* CudaRcbPartitioner::create(MPI_Comm comm, int argc, char** argv)
{
    return new CudaRcbPartitioner(comm, argc, argv);
}

void CudaRcbPartitioner::partition(Grid& grid)
{
    // grid state, as the reference copies it (ZoltanPartitioner.cpp:96-100)
    _num_procs = grid.get_num_procs();
    _global_ext = grid.get_global_ext();
    grid.get_bounding_box(_global[0], _global[1], _local_ext[0], _local_ext[1]);
    _px = grid.get_px();
    _py = grid.get_py();
    const int NX = _global_ext[0], NY = _global_ext[1], P = _num_parts;

    if (ddc_set_mask_host(_h, grid.get_global_land_mask(), NX, NY, 0, NY) != DDC_OK)
        raise(_h, "ddc_set_mask_host");
    int flags = DDC_WANT_PID | DDC_WANT_NEIGHBOURS | (_profile ? DDC_PROFILE : 0);
    if (ddc_partition(_h, P, _px, _py, flags) != DDC_OK)
        raise(_h, "ddc_partition");

    for (int i = 0; i < 4; i++)
        _boxes[i].assign(P, 0);
    if (ddc_get_boxes(_h, _boxes[0].data(), _boxes[1].data(), _boxes[2].data(), _boxes[3].data()) != DDC_OK)
        raise(_h, "ddc_get_boxes");
    _pid_global.assign((size_t)NX * NY, -1);
    if (ddc_get_pid_host(_h, _pid_global.data()) != DDC_OK)
        raise(_h, "ddc_get_pid_host");
    for (int per = 0; per < 2; per++)
        for (int e = 0; e < N_EDGE; e++) {
            const int l = per * N_EDGE + e;
            _nbr_counts[l].assign(P, 0);
            if (ddc_get_neighbour_counts(_h, e, per, _nbr_counts[l].data()) != DDC_OK)
                raise(_h, "ddc_get_neighbour_counts");
            int64_t total = 0;
            if (ddc_get_neighbour_total(_h, e, per, &total) != DDC_OK)
                raise(_h, "ddc_get_neighbour_total");
            _nbr_ids[l].assign((size_t)total, 0);
            _nbr_halos[l].assign((size_t)total, 0);
            _nbr_starts[l].assign((size_t)total, 0);
            if (ddc_get_neighbours(_h, e, per, _nbr_ids[l].data(), _nbr_halos[l].data(), _nbr_starts[l].data()) != DDC_OK)
                raise(_h, "ddc_get_neighbours");
            _nbr_offsets[l].assign(P, 0);
            int run = 0;
            for (int p = 0; p < P; p++) {
                _nbr_offsets[l][p] = run;
                run += _nbr_counts[l][p];
            }
        }
    if (ddc_get_stats(_h, &_stats) != DDC_OK)
        raise(_h, "ddc_get_stats");
    publish_rank_view();
}
