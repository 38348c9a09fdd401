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

    // argv may carry `--parts N` (number of boxes, default = communicator size) and
    // `--device D` (CUDA device, default 0).  Throws std::runtime_error when no GPU is usable:
    // there is no CPU fallback.
    static CudaRcbPartitioner* create(MPI_Comm comm, int argc, char** argv);

    void partition(Grid& grid) override;

    // statistics of the last partition() (device timings only when profiling was requested)
    const ddc_stats& stats() const { return _stats; }
    void set_profile(bool on) { _profile = on; }

protected:
    CudaRcbPartitioner(MPI_Comm comm, int argc, char** argv);

private:
    ddc_handle_t _h = nullptr;
    ddc_stats _stats {};
    bool _profile = false;
};
