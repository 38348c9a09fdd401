/*
 * mpi.h -- single-process stand-in for <mpi.h>, used when the host library is built without a real
 * MPI (DDC_HAVE_MPI undefined; this image has none).  It only provides what the public headers of
 * the reference API need: the MPI_Comm type in the signatures of Grid::create and
 * Partitioner::Factory::create, rank / size queries and init / finalize.
 *
 * A communicator is a (rank, size) pair.  MPI_COMM_WORLD is rank 0 of 1.  ddc_shim_comm(rank, size)
 * makes "rank r of n" so that code written against the reference API -- one box per rank, rank r
 * asks for its own box / neighbours / pid slab -- can be exercised rank by rank in one process
 * (the reference's tests do exactly that with MPI_TEST_CASE(name, N), test/MainMPI.cpp).
 * Nothing is ever communicated: the CUDA partitioner computes every part from the global mask.
 */
#ifndef DDC_MPI_SHIM_H
#define DDC_MPI_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef long long MPI_Comm; /* size in the upper, rank in the lower 32 bits: any int part count fits */
#define MPI_COMM_WORLD ((MPI_Comm)0x100000000LL) /* size 1, rank 0 */
#define MPI_SUCCESS 0
#define MPI_MAX_ERROR_STRING 64

static inline MPI_Comm ddc_shim_comm(int rank, int size)
{
    return (MPI_Comm)(((unsigned long long)(unsigned)size << 32) | (unsigned long long)(unsigned)rank);
}
static inline int MPI_Init(int* argc, char*** argv)
{
    (void)argc;
    (void)argv;
    return MPI_SUCCESS;
}
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm comm, int* rank)
{
    *rank = (int)(unsigned)((unsigned long long)comm & 0xffffffffULL);
    return MPI_SUCCESS;
}
static inline int MPI_Comm_size(MPI_Comm comm, int* size)
{
    *size = (int)(unsigned)((unsigned long long)comm >> 32);
    return MPI_SUCCESS;
}

#ifdef __cplusplus
}
#endif
#endif
