/* mpi.h — stand-in for the MPI calls streamParanumal's first-party code makes (reference: include/comm.hpp:30-575,
 * libs/core/comm.cpp:34-120; full symbol list: grep -rhoE '\bMPI_[A-Za-z_]+' include libs BS).  This image has no MPI.
 * With `-I include/shim` the reference's comm_t compiles unchanged against this header; the functions live in
 * streamparanumal_b200/shim/mpi_shim.cpp: the ranks of ONE node, started by `spnrun -np P <exe>` (the mpirun stand-in, which
 * sets SPN_RANK / SPN_SIZE / SPN_RENDEZVOUS), talk over a mesh of Unix-domain sockets; with no launcher the process is rank 0
 * of 1.  Host buffers only.  Point-to-point messages match by (communicator, source, tag) in arrival order; collectives are
 * built on them and reduce in rank order.
 */
#ifndef SPN_SHIM_MPI_H
#define SPN_SHIM_MPI_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_NULL ((MPI_Comm)0)
#define MPI_COMM_WORLD ((MPI_Comm)1)
#define MPI_REQUEST_NULL ((MPI_Request)0)
#define MPI_IN_PLACE ((void*)1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)

/* datatype handle = (kind << 16) | bytes per element; kind 0 = opaque bytes (MPI_Type_contiguous) */
#define MPI_CHAR ((MPI_Datatype)((1 << 16) | 1))
#define MPI_INT ((MPI_Datatype)((2 << 16) | 4))
#define MPI_LONG_LONG_INT ((MPI_Datatype)((3 << 16) | 8))
#define MPI_FLOAT ((MPI_Datatype)((4 << 16) | 4))
#define MPI_DOUBLE ((MPI_Datatype)((5 << 16) | 8))

#define MPI_MAX ((MPI_Op)1)
#define MPI_MIN ((MPI_Op)2)
#define MPI_SUM ((MPI_Op)3)
#define MPI_PROD ((MPI_Op)4)
#define MPI_LAND ((MPI_Op)5)
#define MPI_LOR ((MPI_Op)6)
#define MPI_LXOR ((MPI_Op)7)

int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int errorcode);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm* newcomm);
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm* newcomm);
int MPI_Comm_free(MPI_Comm* comm);
int MPI_Get_processor_name(char* name, int* resultlen);
int MPI_Barrier(MPI_Comm comm);

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype);
int MPI_Type_commit(MPI_Datatype* type);
int MPI_Type_free(MPI_Datatype* type);

int MPI_Send(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status* status);
int MPI_Isend(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request* request);
int MPI_Irecv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request* request);
int MPI_Wait(MPI_Request* request, MPI_Status* status);
int MPI_Waitall(int count, MPI_Request* requests, MPI_Status* statuses);

int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Iallreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm, MPI_Request* request);
int MPI_Scan(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Gather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype, int root,
               MPI_Comm comm);
int MPI_Gatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, const int* recvcounts, const int* displs,
                MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Scatter(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype, int root,
                MPI_Comm comm);
int MPI_Scatterv(const void* sendbuf, const int* sendcounts, const int* displs, MPI_Datatype sendtype, void* recvbuf, int recvcount,
                 MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Allgather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype,
                  MPI_Comm comm);
int MPI_Allgatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, const int* recvcounts, const int* displs,
                   MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Alltoall(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype,
                 MPI_Comm comm);
int MPI_Alltoallv(const void* sendbuf, const int* sendcounts, const int* sdispls, MPI_Datatype sendtype, void* recvbuf,
                  const int* recvcounts, const int* rdispls, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Ialltoallv(const void* sendbuf, const int* sendcounts, const int* sdispls, MPI_Datatype sendtype, void* recvbuf,
                   const int* recvcounts, const int* rdispls, MPI_Datatype recvtype, MPI_Comm comm, MPI_Request* request);

#ifdef __cplusplus
}
#endif
#endif
