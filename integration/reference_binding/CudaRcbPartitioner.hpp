/*!
 * @file CudaRcbPartitioner.hpp
 * @brief The file a maintainer ADDS to the reference tree (nextsimhub/domain_decomp) to switch the
 *        partitioner from Zoltan to the B200 library: a Partitioner subclass over the C ABI of
 *        include/ddc.h.  It is written against the reference's OWN headers (Partitioner.hpp,
 *        Grid.hpp) and replaces ZoltanPartitioner.hpp:24-59 / ZoltanPartitioner.cpp:69-224; nothing
 *        else in the reference changes (Factory::create gains one case, Partitioner.cpp:320-327).
 *
 * Tested: oracle/Makefile (`make ref`) compiles it together with the reference's Grid.cpp,
 * Partitioner.cpp and DomainUtils.cpp where they lie, and tests/test_zz_reference_binding.py runs
 * the reference's integration cases through it on a GPU.
 */
#pragma once

#include "Grid.hpp"
#include "Partitioner.hpp"

struct ddc_handle_s;

class CudaRcbPartitioner final : public Partitioner {
public:
    CudaRcbPartitioner(const CudaRcbPartitioner&) = delete;
    CudaRcbPartitioner& operator=(const CudaRcbPartitioner&) = delete;
    ~CudaRcbPartitioner();

    // same named-constructor idiom as ZoltanPartitioner::create (ZoltanPartitioner.hpp:43);
    // understands `--device D` in argv (default: rank modulo the number of visible devices is NOT
    // assumed -- every rank uses device 0 unless told otherwise)
    static CudaRcbPartitioner* create(MPI_Comm comm, int argc, char** argv);

    // replaces ZoltanPartitioner::partition (ZoltanPartitioner.cpp:93-224)
    void partition(Grid& grid) override;

protected:
    CudaRcbPartitioner(MPI_Comm comm, int device);

private:
    ddc_handle_s* _ddc = nullptr;
};
