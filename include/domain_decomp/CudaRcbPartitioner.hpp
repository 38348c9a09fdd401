// CudaRcbPartitioner.hpp -- the concrete partitioner that replaces ZoltanPartitioner
// (reference: ZoltanPartitioner.hpp:24-59).  It is a thin host wrapper over the C ABI in ddc.h.
#pragma once

#include "Grid.hpp"
#include "Partitioner.hpp"
#include "ddc.h"

class LIB_EXPORT CudaRcbPartitioner final : public Partitioner {
public:
    CudaRcbPartitioner(const CudaRcbPartitioner&) = delete;
    CudaRcbPartitioner& operator=(const CudaRcbPartitioner&) = delete;
    ~CudaRcbPartitioner();

    // argv may carry `--parts N` (number of boxes, default = communicator size), `--device D` (first CUDA
    // device, default 0) and `--gpus G` (row-shard the mask over the G GPUs D .. D + G - 1 of this box; the
    // reference's analogue is `mpirun -n G`, main.cpp:78-94).  Throws std::runtime_error when no GPU is
    // usable: there is no CPU fallback.
    static CudaRcbPartitioner* create(MPI_Comm comm, int argc, char** argv);

    void partition(Grid& grid) override;

    // statistics of the last partition() (device timings only when profiling was requested)
    const ddc_stats& stats() const { return _stats; }
    int num_gpus() const { return (int)_hs.size(); }
    void set_profile(bool on) { _profile = on; }

protected:
    CudaRcbPartitioner(MPI_Comm comm, int argc, char** argv);

private:
    void create_handles(int gpus);
    void destroy_handles();
    std::vector<ddc_handle_t> _hs; // one per GPU (rank g = device _first_device + g)
    int _first_device = 0;
    int _connected_for[3] = { 0, 0, 0 }; // NX, NY, P the exchange buffers were sized for
    ddc_stats _stats {};
    bool _profile = false;
};
