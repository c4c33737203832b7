// Device-buffer communication over NCCL (include/spn_b200.h, "comm" section).
// Replaces comm_t::{Irecv,Isend,Waitall} on device pointers in the GPU-aware
// pairwise exchange (libs/ogs/ogsPairwise.cpp:135-152): one ncclGroup of
// send/recv per exchange, enqueued on a stream, no host wait.
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>
#include "spn_b200.h"

struct spn_comm {
  ncclComm_t nccl;
  int rank;
  int size;
};

static_assert(sizeof(ncclUniqueId) == SPN_UNIQUE_ID_BYTES, "ncclUniqueId size");

#define SPN_NCCL_TRY(expr)                           \
  do {                                               \
    ncclResult_t _r = (expr);                        \
    if (_r != ncclSuccess && _r != ncclInProgress) return SPN_ERR_NCCL; \
  } while (0)

extern "C" {

int spn_comm_unique_id(void* id128) {
  if (!id128) return SPN_ERR_INVALID_ARGUMENT;
  ncclUniqueId id;
  SPN_NCCL_TRY(ncclGetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return SPN_SUCCESS;
}

int spn_comm_init(spn_comm_t* comm, const void* id128, int rank, int size) {
  if (!comm || !id128 || rank < 0 || rank >= size) return SPN_ERR_INVALID_ARGUMENT;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  spn_comm* c = static_cast<spn_comm*>(calloc(1, sizeof(spn_comm)));
  if (!c) return SPN_ERR_INVALID_ARGUMENT;
  c->rank = rank;
  c->size = size;
  ncclResult_t r = ncclCommInitRank(&c->nccl, size, id, rank);
  if (r != ncclSuccess) { free(c); return SPN_ERR_NCCL; }
  *comm = c;
  return SPN_SUCCESS;
}

int spn_comm_destroy(spn_comm_t comm) {
  if (!comm) return SPN_SUCCESS;
  ncclCommDestroy(comm->nccl);
  free(comm);
  return SPN_SUCCESS;
}

int spn_comm_rank(spn_comm_t comm, int* rank) {
  if (!comm || !rank) return SPN_ERR_INVALID_ARGUMENT;
  *rank = comm->rank;
  return SPN_SUCCESS;
}
int spn_comm_size(spn_comm_t comm, int* size) {
  if (!comm || !size) return SPN_ERR_INVALID_ARGUMENT;
  *size = comm->size;
  return SPN_SUCCESS;
}

int spn_comm_exchange(spn_comm_t comm,
                      int nrecv, const int* recvRanks, const size_t* recvOffsets, const size_t* recvBytes, void* recvBuf,
                      int nsend, const int* sendRanks, const size_t* sendOffsets, const size_t* sendBytes, const void* sendBuf,
                      spn_stream_t stream) {
  if (!comm || nrecv < 0 || nsend < 0) return SPN_ERR_INVALID_ARGUMENT;
  if (nrecv + nsend == 0) return SPN_SUCCESS;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SPN_NCCL_TRY(ncclGroupStart());
  for (int r = 0; r < nrecv; ++r) {
    if (recvBytes[r] == 0) continue;
    SPN_NCCL_TRY(ncclRecv(static_cast<char*>(recvBuf) + recvOffsets[r], recvBytes[r], ncclChar, recvRanks[r], comm->nccl, st));
  }
  for (int s = 0; s < nsend; ++s) {
    if (sendBytes[s] == 0) continue;
    SPN_NCCL_TRY(ncclSend(static_cast<const char*>(sendBuf) + sendOffsets[s], sendBytes[s], ncclChar, sendRanks[s], comm->nccl, st));
  }
  SPN_NCCL_TRY(ncclGroupEnd());
  return SPN_SUCCESS;
}

int spn_comm_allreduce_sum_f64(spn_comm_t comm, double* buf, int count, spn_stream_t stream) {
  if (!comm || count < 0) return SPN_ERR_INVALID_ARGUMENT;
  if (count == 0) return SPN_SUCCESS;
  SPN_NCCL_TRY(ncclAllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, comm->nccl,
                             reinterpret_cast<cudaStream_t>(stream)));
  return SPN_SUCCESS;
}

int spn_comm_barrier(spn_comm_t comm, spn_stream_t stream) {
  if (!comm) return SPN_ERR_INVALID_ARGUMENT;
  static thread_local double* token = nullptr;
  if (!token) {
    if (cudaMalloc(&token, sizeof(double)) != cudaSuccess) return SPN_ERR_NCCL;
    cudaMemset(token, 0, sizeof(double));
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SPN_NCCL_TRY(ncclAllReduce(token, token, 1, ncclDouble, ncclSum, comm->nccl, st));
  if (cudaStreamSynchronize(st) != cudaSuccess) return SPN_ERR_NCCL;
  return SPN_SUCCESS;
}

}  // extern "C"
