/* mpi.h -- TEST INFRASTRUCTURE.  A miniature MPI for compiling the REFERENCE's own Grid.cpp /
 * Partitioner.cpp where they lie (oracle/Makefile target `ref`): P "ranks" are P threads of one
 * process, a communicator is a (world, rank) pair, and the few collectives the reference's host
 * path uses (Grid.cpp:142-146, Partitioner.cpp:85-86,190-205,378-388; MPI_Gatherv / MPI_Bcast / MPI_Scatterv for integration/reference_binding) are implemented over a
 * generation barrier in oracle/ref_hostpath_shim.cpp.  Not part of the product. */
#ifndef DDC_REF_SHIM_MPI_H
#define DDC_REF_SHIM_MPI_H

#ifdef __cplusplus
extern "C" {
#endif

struct ref_shim_comm;
typedef struct ref_shim_comm* MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;

#define MPI_SUCCESS 0
#define MPI_MAX_ERROR_STRING 256
#define MPI_INT 1
#define MPI_SUM 1
#define MPI_INFO_NULL 0

int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Allgather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
    MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Allgatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, const int* recvcounts,
    const int* displs, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Gatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, const int* recvcounts,
    const int* displs, MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Scatterv(const void* sendbuf, const int* sendcounts, const int* displs, MPI_Datatype sendtype, void* recvbuf,
    int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Bcast(void* buffer, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Exscan(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Error_string(int errorcode, char* string, int* resultlen);
int MPI_Finalize(void);

#ifdef __cplusplus
}
#endif
#endif
