// CudaRcbPartitioner.cpp -- ZoltanPartitioner's replacement: a thin host wrapper over the C ABI.
// It owns no arithmetic: boxes, pid and the neighbour tables all come from libddc_cuda.
//
// One GPU: one handle.  `--gpus G`: the mask is row-sharded over G GPUs of this box (the analogue of
// `mpirun -n G`, main.cpp:78-94 of the reference) -- G handles in this process, connected through peer memory
// (ddc_peer_connect), each driven by its own host thread, so that the G shards cross PCIe on G links at once
// and the exchange steps inside ddc_partition find every rank enqueued.
#include "CudaRcbPartitioner.hpp"

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>

namespace {
[[noreturn]] void raise(ddc_handle_t h, const char* what)
{
    throw std::runtime_error(std::string("ERROR: ") + what + ": " + ddc_last_error(h));
}
} // namespace

CudaRcbPartitioner::CudaRcbPartitioner(MPI_Comm comm, int argc, char** argv)
    : Partitioner(comm)
{
    int device = 0, gpus = 1;
    for (int i = 1; i + 1 < argc; i++) {
        if (!argv || !argv[i])
            continue;
        if (std::strcmp(argv[i], "--parts") == 0)
            set_num_parts(std::atoi(argv[i + 1]));
        else if (std::strcmp(argv[i], "--device") == 0)
            device = std::atoi(argv[i + 1]);
        else if (std::strcmp(argv[i], "--gpus") == 0)
            gpus = std::atoi(argv[i + 1]);
    }
    if (gpus < 1 || gpus > 16)
        throw std::runtime_error("ERROR: --gpus must be between 1 and 16");
    _first_device = device;
    create_handles(gpus);
}

void CudaRcbPartitioner::create_handles(int gpus)
{
    destroy_handles();
    for (int g = 0; g < gpus; g++) {
        ddc_handle_t h = nullptr;
        if (ddc_create(&h, _first_device + g, g, gpus, nullptr) != DDC_OK) {
            const std::string why = ddc_last_error(nullptr);
            destroy_handles();
            throw std::runtime_error("ERROR: cannot create the CUDA partitioner (there is no CPU fallback): " + why);
        }
        _hs.push_back(h);
    }
    _connected_for[0] = _connected_for[1] = _connected_for[2] = 0;
}

void CudaRcbPartitioner::destroy_handles()
{
    for (ddc_handle_t h : _hs)
        ddc_peer_close(h);
    for (ddc_handle_t h : _hs)
        ddc_destroy(h);
    _hs.clear();
}

CudaRcbPartitioner::~CudaRcbPartitioner() { destroy_handles(); }

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
    const int G = (int)_hs.size();
    const int flags = DDC_WANT_PID | DDC_WANT_NEIGHBOURS | (_profile ? DDC_PROFILE : 0);
    const int* mask = grid.get_global_land_mask();
    _pid_global.resize((size_t)NX * NY); // not initialised: every cell is written by the device (land = -1)

    if (G == 1) {
        ddc_handle_t h = _hs[0];
        if (ddc_set_mask_host(h, mask, NX, NY, 0, NY) != DDC_OK)
            raise(h, "ddc_set_mask_host");
        if (ddc_partition(h, P, _px, _py, flags) != DDC_OK)
            raise(h, "ddc_partition");
        if (ddc_get_pid_host(h, _pid_global.data()) != DDC_OK)
            raise(h, "ddc_get_pid_host");
    } else {
        if (_connected_for[0] != NX || _connected_for[1] != NY || _connected_for[2] != P) {
            if (_connected_for[0]) // exchange buffers of another geometry: start over
                create_handles(G);
            if (ddc_peer_connect(_hs.data(), G, NX, NY, P) != DDC_OK)
                raise(nullptr, "ddc_peer_connect");
            _connected_for[0] = NX;
            _connected_for[1] = NY;
            _connected_for[2] = P;
        }
        std::vector<std::string> err(G);
        std::vector<std::thread> workers;
        for (int g = 0; g < G; g++)
            workers.emplace_back([&, g] {
                ddc_handle_t h = _hs[g];
                int yb = 0, yc = 0;
                ddc_shard_rows(NY, G, g, &yb, &yc);
                const char* what = nullptr;
                if (ddc_set_mask_host(h, mask + (size_t)yb * NX, NX, NY, yb, yc) != DDC_OK)
                    what = "ddc_set_mask_host";
                else if (ddc_partition(h, P, _px, _py, flags) != DDC_OK)
                    what = "ddc_partition";
                else if (ddc_get_pid_host(h, _pid_global.data() + (size_t)yb * NX) != DDC_OK)
                    what = "ddc_get_pid_host";
                if (what)
                    err[g] = std::string("ERROR: ") + what + " (GPU " + std::to_string(g) + "): " + ddc_last_error(h);
            });
        for (std::thread& t : workers)
            t.join();
        for (const std::string& e : err)
            if (!e.empty())
                throw std::runtime_error(e);
    }

    // the replicated results: every rank holds the same boxes and tables, rank 0 hands them out
    ddc_handle_t h = _hs[0];
    for (int i = 0; i < 4; i++)
        _boxes[i].assign(P, 0);
    if (ddc_get_boxes(h, _boxes[0].data(), _boxes[1].data(), _boxes[2].data(), _boxes[3].data()) != DDC_OK)
        raise(h, "ddc_get_boxes");
    for (int per = 0; per < 2; per++)
        for (int e = 0; e < N_EDGE; e++) {
            const int l = per * N_EDGE + e;
            _nbr_counts[l].assign(P, 0);
            if (ddc_get_neighbour_counts(h, e, per, _nbr_counts[l].data()) != DDC_OK)
                raise(h, "ddc_get_neighbour_counts");
            int64_t total = 0;
            if (ddc_get_neighbour_total(h, e, per, &total) != DDC_OK)
                raise(h, "ddc_get_neighbour_total");
            _nbr_ids[l].assign((size_t)total, 0);
            _nbr_halos[l].assign((size_t)total, 0);
            _nbr_starts[l].assign((size_t)total, 0);
            if (ddc_get_neighbours(h, e, per, _nbr_ids[l].data(), _nbr_halos[l].data(), _nbr_starts[l].data()) != DDC_OK)
                raise(h, "ddc_get_neighbours");
            _nbr_offsets[l].assign(P, 0);
            int run = 0;
            for (int p = 0; p < P; p++) {
                _nbr_offsets[l][p] = run;
                run += _nbr_counts[l][p];
            }
        }
    if (ddc_get_stats(h, &_stats) != DDC_OK)
        raise(h, "ddc_get_stats");
    publish_rank_view();
}
