// spn::comm_t — the communicator the host layer uses where the reference uses comm_t
// (include/comm.hpp:95-575, an MPI wrapper).  This image has no MPI, so:
//   * host-side collectives (the int/int64 Alltoallv/Allreduce/Allgather calls of ogs setup,
//     libs/ogs/ogsSetup.cpp:319-460,651-788) run over a full mesh of Unix-domain sockets between
//     the ranks of one node (ranks are separate processes: torchrun, or tools/spnrun);
//   * device buffers never touch these sockets: halo send/recv and scalar allreduce go through
//     the NCCL entry points of include/spn_b200.h (spn_comm_exchange, spn_comm_allreduce_sum_f64),
//     replacing comm_t::{Isend,Irecv,Waitall} on device pointers (libs/ogs/ogsPairwise.cpp:135-152).
// Method names follow the reference (rank, size, Barrier, Alltoall, Alltoallv, Allreduce,
// Allgather, Bcast, Scan).  Collectives are deterministic: reductions fold in rank order.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "spn_b200.h"

namespace spn {

struct exception : public std::runtime_error {
  explicit exception(const std::string& msg) : std::runtime_error(msg) {}
};

// LIBP_ABORT equivalent (include/utils.hpp:57-75): throw when cond holds
#define SPN_ABORT(msg, cond)                                                                       \
  do {                                                                                             \
    if (cond) throw ::spn::exception(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + (msg)); \
  } while (0)

// C-ABI status -> exception (the reference raises occa::exception / libp::exception)
void check_status(int status, const char* where);
#define SPN_CHECK(call) ::spn::check_status((call), #call)

namespace Comm {
enum Op { Sum, Max, Min, Prod };
}

class comm_t {
 public:
  comm_t() = default;
  // rendezvous: every rank binds <dir>/r<rank>.sock; rank i connects to every j<i
  comm_t(int rank, int size, const std::string& rendezvous_dir);
  // RANK/WORLD_SIZE/MASTER_PORT (torchrun) or SPN_RANK/SPN_SIZE/SPN_RENDEZVOUS; size 1 if unset
  static comm_t World();

  int rank() const { return impl_ ? rank_ : 0; }
  int size() const { return impl_ ? size_ : 1; }
  bool isInitialized() const { return (bool)impl_; }

  void Barrier();

  // byte-level primitives
  void sendrecv(int dst, const void* sbuf, size_t sbytes, int src, void* rbuf, size_t rbytes);
  void alltoallv_bytes(const void* sbuf, const size_t* scounts, const size_t* soffsets, void* rbuf,
                       const size_t* rcounts, const size_t* roffsets);

  template <typename T>
  void Alltoall(const T* send, T* recv, int cnt = 1) {
    std::vector<size_t> c((size_t)size(), sizeof(T) * cnt), o((size_t)size());
    for (int r = 0; r < size(); ++r) o[r] = sizeof(T) * cnt * (size_t)r;
    alltoallv_bytes(send, c.data(), o.data(), recv, c.data(), o.data());
  }
  // counts / offsets in elements, as MPI_Alltoallv (comm_t::Alltoallv, include/comm.hpp)
  template <typename T>
  void Alltoallv(const T* send, const int* scounts, const int* soffsets, T* recv, const int* rcounts,
                 const int* roffsets) {
    const int P = size();
    std::vector<size_t> sc(P), so(P), rc(P), ro(P);
    for (int r = 0; r < P; ++r) {
      sc[r] = sizeof(T) * (size_t)scounts[r];
      so[r] = sizeof(T) * (size_t)soffsets[r];
      rc[r] = sizeof(T) * (size_t)rcounts[r];
      ro[r] = sizeof(T) * (size_t)roffsets[r];
    }
    alltoallv_bytes(send, sc.data(), so.data(), recv, rc.data(), ro.data());
  }
  template <typename T>
  void Allgather(const T* send, T* recv, int cnt = 1) {
    const int P = size();
    std::vector<T> tmp((size_t)P * cnt);
    for (int r = 0; r < P; ++r)
      for (int i = 0; i < cnt; ++i) tmp[(size_t)r * cnt + i] = send[i];
    Alltoall(tmp.data(), recv, cnt);
  }
  template <typename T>
  void Allreduce(T* buf, int cnt, Comm::Op op = Comm::Sum) {
    const int P = size();
    if (P == 1) return;
    std::vector<T> all((size_t)P * cnt);
    Allgather(buf, all.data(), cnt);
    for (int i = 0; i < cnt; ++i) {
      T a = all[i];
      for (int r = 1; r < P; ++r) {
        const T b = all[(size_t)r * cnt + i];
        switch (op) {
          case Comm::Sum: a = a + b; break;
          case Comm::Prod: a = a * b; break;
          case Comm::Max: a = (b > a) ? b : a; break;
          case Comm::Min: a = (b < a) ? b : a; break;
        }
      }
      buf[i] = a;
    }
  }
  template <typename T>
  void Allreduce(T& v, Comm::Op op = Comm::Sum) { Allreduce(&v, 1, op); }
  // inclusive prefix over ranks (MPI_Scan)
  template <typename T>
  void Scan(const T& snd, T& rcv) {
    std::vector<T> all((size_t)size());
    Allgather(&snd, all.data(), 1);
    T a = all[0];
    for (int r = 1; r <= rank(); ++r) a = a + all[r];
    rcv = a;
  }
  template <typename T>
  void Bcast(T* buf, int cnt, int root) {
    const int P = size();
    if (P == 1) return;
    std::vector<size_t> sc(P, 0), so(P, 0), rc(P, 0), ro(P, 0);
    if (rank() == root)
      for (int r = 0; r < P; ++r) sc[r] = sizeof(T) * (size_t)cnt;
    rc[root] = sizeof(T) * (size_t)cnt;
    std::vector<T> tmp(buf, buf + cnt);
    alltoallv_bytes(tmp.data(), sc.data(), so.data(), buf, rc.data(), ro.data());
  }

  // ---- device side (NCCL), created on first use for the current CUDA device ----------------
  spn_comm_t device();          // NCCL communicator (unique id broadcast over the sockets)
  bool deviceReady() const;
  void AllreduceDevice(double* d_buf, int count, spn_stream_t stream);  // sum, in place

 private:
  struct Impl;
  std::shared_ptr<Impl> impl_;
  int rank_ = 0, size_ = 1;
};

}  // namespace spn
