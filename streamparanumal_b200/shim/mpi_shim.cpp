// Single-rank implementation of include/shim/mpi.h (`mpirun -np 1` semantics): the reference's comm_t
// (include/comm.hpp:95-575) compiles against the header and links against this file unchanged.
// Collectives over one rank are the identity (send buffer -> receive buffer unless MPI_IN_PLACE); a message a rank
// sends to itself is matched with the receive of the same (communicator, tag) in posting order.
#include <mpi.h>

#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <vector>

namespace {

inline size_t bytes_of(MPI_Datatype t) { return (size_t)(t & 0xffff) + (((size_t)t >> 24) << 16); }

struct Message {
  MPI_Comm comm;
  int tag;
  std::vector<char> data;
};
struct PendingRecv {
  MPI_Comm comm;
  int tag;
  void* buf;
  size_t bytes;
  int request;
};
std::deque<Message> g_unexpected;
std::deque<PendingRecv> g_posted;
std::vector<char> g_done(1, 1);  // request handle -> complete; handle 0 = MPI_REQUEST_NULL
int g_next_comm = 2;

int new_request(bool done) {
  g_done.push_back(done ? 1 : 0);
  return (int)g_done.size() - 1;
}

void fail(const char* what) {
  fprintf(stderr, "mpi shim (single rank): %s\n", what);
  abort();
}

void check_peer(int peer, const char* what) {
  if (peer != 0 && peer != MPI_ANY_SOURCE) fail(what);
}

void deliver(MPI_Comm comm, int tag, const void* buf, size_t bytes) {
  for (auto it = g_posted.begin(); it != g_posted.end(); ++it) {
    if (it->comm == comm && (it->tag == tag || it->tag == MPI_ANY_TAG)) {
      if (bytes > it->bytes) fail("message longer than the posted receive");
      if (bytes) memcpy(it->buf, buf, bytes);
      g_done[it->request] = 1;
      g_posted.erase(it);
      return;
    }
  }
  Message m{comm, tag, std::vector<char>((const char*)buf, (const char*)buf + bytes)};
  g_unexpected.push_back(std::move(m));
}

bool try_receive(MPI_Comm comm, int tag, void* buf, size_t bytes) {
  for (auto it = g_unexpected.begin(); it != g_unexpected.end(); ++it) {
    if (it->comm == comm && (it->tag == tag || tag == MPI_ANY_TAG)) {
      if (it->data.size() > bytes) fail("message longer than the receive buffer");
      if (!it->data.empty()) memcpy(buf, it->data.data(), it->data.size());
      g_unexpected.erase(it);
      return true;
    }
  }
  return false;
}

void copy_unless_in_place(const void* send, void* recv, size_t bytes) {
  if (send != MPI_IN_PLACE && send != recv && bytes) memcpy(recv, send, bytes);
}

}  // namespace

extern "C" {

int MPI_Init(int*, char***) { return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm, int errorcode) { _exit(errorcode ? errorcode : 1); }
int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = 0; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm, int* size) { *size = 1; return MPI_SUCCESS; }
int MPI_Comm_dup(MPI_Comm, MPI_Comm* newcomm) { *newcomm = g_next_comm++; return MPI_SUCCESS; }
int MPI_Comm_split(MPI_Comm, int, int, MPI_Comm* newcomm) { *newcomm = g_next_comm++; return MPI_SUCCESS; }
int MPI_Comm_free(MPI_Comm* comm) { *comm = MPI_COMM_NULL; return MPI_SUCCESS; }
int MPI_Get_processor_name(char* name, int* resultlen) {
  if (gethostname(name, MPI_MAX_PROCESSOR_NAME - 1) != 0) strcpy(name, "localhost");
  name[MPI_MAX_PROCESSOR_NAME - 1] = 0;
  *resultlen = (int)strlen(name);
  return MPI_SUCCESS;
}
int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype) {
  const size_t b = (size_t)count * bytes_of(oldtype);
  if (b >= ((size_t)1 << 23)) fail("MPI_Type_contiguous: type too large for this shim");
  *newtype = (MPI_Datatype)((b & 0xffff) | ((b >> 16) << 24));  // kind 0: opaque bytes
  return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype*) { return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype*) { return MPI_SUCCESS; }

int MPI_Send(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm) {
  check_peer(dest, "MPI_Send to a rank other than 0");
  deliver(comm, tag, buf, (size_t)count * bytes_of(type));
  return MPI_SUCCESS;
}
int MPI_Recv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status*) {
  check_peer(source, "MPI_Recv from a rank other than 0");
  if (!try_receive(comm, tag, buf, (size_t)count * bytes_of(type))) fail("MPI_Recv would block forever (no matching send)");
  return MPI_SUCCESS;
}
int MPI_Isend(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request* request) {
  check_peer(dest, "MPI_Isend to a rank other than 0");
  deliver(comm, tag, buf, (size_t)count * bytes_of(type));
  *request = new_request(true);
  return MPI_SUCCESS;
}
int MPI_Irecv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request* request) {
  check_peer(source, "MPI_Irecv from a rank other than 0");
  const size_t bytes = (size_t)count * bytes_of(type);
  if (try_receive(comm, tag, buf, bytes)) {
    *request = new_request(true);
  } else {
    *request = new_request(false);
    g_posted.push_back(PendingRecv{comm, tag, buf, bytes, *request});
  }
  return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request* request, MPI_Status*) {
  if (*request != MPI_REQUEST_NULL && !g_done[*request]) fail("MPI_Wait would block forever (no matching send)");
  *request = MPI_REQUEST_NULL;
  return MPI_SUCCESS;
}
int MPI_Waitall(int count, MPI_Request* requests, MPI_Status*) {
  for (int i = 0; i < count; ++i) MPI_Wait(requests + i, MPI_STATUS_IGNORE);
  return MPI_SUCCESS;
}

int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op, int, MPI_Comm) {
  copy_unless_in_place(sendbuf, recvbuf, (size_t)count * bytes_of(type));
  return MPI_SUCCESS;
}
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op, MPI_Comm) {
  copy_unless_in_place(sendbuf, recvbuf, (size_t)count * bytes_of(type));
  return MPI_SUCCESS;
}
int MPI_Iallreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op, MPI_Comm, MPI_Request* request) {
  copy_unless_in_place(sendbuf, recvbuf, (size_t)count * bytes_of(type));
  *request = new_request(true);
  return MPI_SUCCESS;
}
int MPI_Scan(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op, MPI_Comm) {
  copy_unless_in_place(sendbuf, recvbuf, (size_t)count * bytes_of(type));
  return MPI_SUCCESS;
}
int MPI_Gather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int, MPI_Datatype, int, MPI_Comm) {
  copy_unless_in_place(sendbuf, recvbuf, (size_t)sendcount * bytes_of(sendtype));
  return MPI_SUCCESS;
}
int MPI_Gatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, const int*, const int* displs,
                MPI_Datatype recvtype, int, MPI_Comm) {
  copy_unless_in_place(sendbuf, (char*)recvbuf + (size_t)displs[0] * bytes_of(recvtype), (size_t)sendcount * bytes_of(sendtype));
  return MPI_SUCCESS;
}
int MPI_Scatter(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int, MPI_Datatype, int, MPI_Comm) {
  if (recvbuf != MPI_IN_PLACE && recvbuf != sendbuf) memcpy(recvbuf, sendbuf, (size_t)sendcount * bytes_of(sendtype));
  return MPI_SUCCESS;
}
int MPI_Scatterv(const void* sendbuf, const int* sendcounts, const int* displs, MPI_Datatype sendtype, void* recvbuf, int,
                 MPI_Datatype, int, MPI_Comm) {
  if (recvbuf != MPI_IN_PLACE)
    memmove(recvbuf, (const char*)sendbuf + (size_t)displs[0] * bytes_of(sendtype), (size_t)sendcounts[0] * bytes_of(sendtype));
  return MPI_SUCCESS;
}
int MPI_Allgather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int, MPI_Datatype, MPI_Comm) {
  copy_unless_in_place(sendbuf, recvbuf, (size_t)sendcount * bytes_of(sendtype));
  return MPI_SUCCESS;
}
int MPI_Allgatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, const int*, const int* displs,
                   MPI_Datatype recvtype, MPI_Comm) {
  copy_unless_in_place(sendbuf, (char*)recvbuf + (size_t)displs[0] * bytes_of(recvtype), (size_t)sendcount * bytes_of(sendtype));
  return MPI_SUCCESS;
}
int MPI_Alltoall(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int, MPI_Datatype, MPI_Comm) {
  copy_unless_in_place(sendbuf, recvbuf, (size_t)sendcount * bytes_of(sendtype));
  return MPI_SUCCESS;
}
int MPI_Alltoallv(const void* sendbuf, const int* sendcounts, const int* sdispls, MPI_Datatype sendtype, void* recvbuf, const int*,
                  const int* rdispls, MPI_Datatype recvtype, MPI_Comm) {
  if (sendbuf != MPI_IN_PLACE)
    memmove((char*)recvbuf + (size_t)rdispls[0] * bytes_of(recvtype), (const char*)sendbuf + (size_t)sdispls[0] * bytes_of(sendtype),
            (size_t)sendcounts[0] * bytes_of(sendtype));
  return MPI_SUCCESS;
}
int MPI_Ialltoallv(const void* sendbuf, const int* sendcounts, const int* sdispls, MPI_Datatype sendtype, void* recvbuf,
                   const int* recvcounts, const int* rdispls, MPI_Datatype recvtype, MPI_Comm comm, MPI_Request* request) {
  MPI_Alltoallv(sendbuf, sendcounts, sdispls, sendtype, recvbuf, recvcounts, rdispls, recvtype, comm);
  *request = new_request(true);
  return MPI_SUCCESS;
}

}  // extern "C"
