// spn::comm_t implementation: Unix-domain-socket full mesh for host collectives, NCCL (through
// the C ABI) for device buffers.  See comm.hpp for what it replaces in the reference.
#include "comm.hpp"

#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include <chrono>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <functional>
#include <thread>

namespace spn {

void check_status(int status, const char* where) {
  if (status != 0) throw exception(std::string(where) + ": " + spn_error_string(status) + " (status " + std::to_string(status) + ")");
}

struct comm_t::Impl {
  int rank = 0, size = 1;
  std::vector<int> fd;  // fd[r] = stream socket to rank r (-1 for self)
  std::string sock_path;
  int listen_fd = -1;
  spn_comm_t nccl = nullptr;
  ~Impl() {
    if (nccl) spn_comm_destroy(nccl);
    for (int f : fd)
      if (f >= 0) close(f);
    if (listen_fd >= 0) close(listen_fd);
    if (!sock_path.empty()) unlink(sock_path.c_str());
  }
};

namespace {

constexpr int kTimeoutMs = 180000;

void set_nonblocking(int fd) {
  const int fl = fcntl(fd, F_GETFL, 0);
  fcntl(fd, F_SETFL, fl | O_NONBLOCK);
}

void write_all_blocking(int fd, const void* buf, size_t n) {
  const char* p = static_cast<const char*>(buf);
  while (n) {
    const ssize_t w = ::send(fd, p, n, MSG_NOSIGNAL);
    if (w < 0) {
      if (errno == EINTR) continue;
      if (errno == EAGAIN || errno == EWOULDBLOCK) {
        pollfd pf{fd, POLLOUT, 0};
        if (poll(&pf, 1, kTimeoutMs) <= 0) throw exception("spn::comm_t: send timed out");
        continue;
      }
      throw exception(std::string("spn::comm_t: send failed: ") + strerror(errno));
    }
    p += w;
    n -= (size_t)w;
  }
}

void read_all_blocking(int fd, void* buf, size_t n) {
  char* p = static_cast<char*>(buf);
  while (n) {
    const ssize_t r = ::recv(fd, p, n, 0);
    if (r == 0) throw exception("spn::comm_t: peer closed the connection");
    if (r < 0) {
      if (errno == EINTR) continue;
      if (errno == EAGAIN || errno == EWOULDBLOCK) {
        pollfd pf{fd, POLLIN, 0};
        if (poll(&pf, 1, kTimeoutMs) <= 0) throw exception("spn::comm_t: recv timed out");
        continue;
      }
      throw exception(std::string("spn::comm_t: recv failed: ") + strerror(errno));
    }
    p += r;
    n -= (size_t)r;
  }
}

sockaddr_un make_addr(const std::string& path) {
  sockaddr_un a;
  memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  SPN_ABORT("rendezvous path too long: " + path, path.size() >= sizeof(a.sun_path));
  strncpy(a.sun_path, path.c_str(), sizeof(a.sun_path) - 1);
  return a;
}

}  // namespace

namespace {
// 64-bit token of the job a rank belongs to: ranks of different jobs that meet in one rendezvous directory refuse each other
uint64_t job_token(const std::string& dir) {
  std::string key = dir;
  for (const char* name : {"TORCHELASTIC_RUN_ID", "MASTER_ADDR", "MASTER_PORT", "SPN_JOB_ID"})
    if (const char* v = getenv(name)) { key += '|'; key += v; }
  uint64_t h = 1469598103934665603ull;  // FNV-1a
  for (unsigned char c : key) { h ^= c; h *= 1099511628211ull; }
  return h;
}
}  // namespace

comm_t::comm_t(int rank, int size, const std::string& dir) : rank_(rank), size_(size) {
  SPN_ABORT("invalid rank/size", size < 1 || rank < 0 || rank >= size);
  impl_ = std::make_shared<Impl>();
  impl_->rank = rank;
  impl_->size = size;
  impl_->fd.assign((size_t)size, -1);
  if (size == 1) return;

  mkdir(dir.c_str(), 0700);
  impl_->sock_path = dir + "/r" + std::to_string(rank) + ".sock";
  unlink(impl_->sock_path.c_str());
  impl_->listen_fd = socket(AF_UNIX, SOCK_STREAM, 0);
  SPN_ABORT("socket() failed", impl_->listen_fd < 0);
  sockaddr_un me = make_addr(impl_->sock_path);
  SPN_ABORT(std::string("bind failed: ") + strerror(errno), bind(impl_->listen_fd, (sockaddr*)&me, sizeof(me)) != 0);
  SPN_ABORT("listen failed", listen(impl_->listen_fd, size) != 0);

  // connect to every lower rank (its socket may not exist yet, or be a stale file: retry)
  const auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < rank; ++r) {
    const sockaddr_un peer = make_addr(dir + "/r" + std::to_string(r) + ".sock");
    while (true) {
      const int fd = socket(AF_UNIX, SOCK_STREAM, 0);
      SPN_ABORT("socket() failed", fd < 0);
      if (connect(fd, (const sockaddr*)&peer, sizeof(peer)) == 0) {
        const uint64_t hello[2] = {(uint64_t)rank, job_token(dir)};
        write_all_blocking(fd, hello, sizeof(hello));
        impl_->fd[r] = fd;
        break;
      }
      close(fd);
      const auto dt = std::chrono::steady_clock::now() - t0;
      SPN_ABORT("rendezvous timed out waiting for rank " + std::to_string(r),
                std::chrono::duration_cast<std::chrono::milliseconds>(dt).count() > kTimeoutMs);
      std::this_thread::sleep_for(std::chrono::milliseconds(20));
    }
  }
  // accept every higher rank
  for (int n = rank + 1; n < size; ++n) {
    pollfd pf{impl_->listen_fd, POLLIN, 0};
    SPN_ABORT("rendezvous timed out in accept", poll(&pf, 1, kTimeoutMs) <= 0);
    const int fd = accept(impl_->listen_fd, nullptr, nullptr);
    SPN_ABORT("accept failed", fd < 0);
    uint64_t hello2[2] = {~0ull, 0};
    read_all_blocking(fd, hello2, sizeof(hello2));
    SPN_ABORT("rendezvous: a rank of another job connected to " + impl_->sock_path + " (set SPN_RENDEZVOUS to a private directory)",
              hello2[1] != job_token(dir));
    const int hello = (int)hello2[0];
    SPN_ABORT("bad rendezvous hello", hello <= rank || hello >= size || impl_->fd[hello] >= 0);
    impl_->fd[hello] = fd;
  }
  for (int r = 0; r < size; ++r)
    if (impl_->fd[r] >= 0) set_nonblocking(impl_->fd[r]);
  Barrier();
}

comm_t comm_t::World() {
  auto env = [](const char* a, const char* b) -> const char* {
    const char* v = getenv(a);
    return v ? v : getenv(b);
  };
  const char* r = env("SPN_RANK", "RANK");
  const char* s = env("SPN_SIZE", "WORLD_SIZE");
  const int rank = r ? atoi(r) : 0, size = s ? atoi(s) : 1;
  std::string dir;
  if (const char* d = getenv("SPN_RENDEZVOUS")) dir = d;
  else {
    // one directory per job: the launcher's run id / port when there is one, else the parent process (ranks forked by one shell)
    const char* port = getenv("MASTER_PORT");
    const char* run = getenv("TORCHELASTIC_RUN_ID");
    dir = std::string("/tmp/spn_rdv_") + std::to_string((long)getuid()) + "_" + (port ? port : "0") + "_" +
          (run ? std::to_string(std::hash<std::string>{}(run) % 1000003) : std::string("p") + std::to_string((long)getppid()));
  }
  return comm_t(rank, size, dir);
}

void comm_t::sendrecv(int dst, const void* sbuf, size_t sbytes, int src, void* rbuf, size_t rbytes) {
  if (dst == rank_ && src == rank_) {
    SPN_ABORT("self sendrecv size mismatch", sbytes != rbytes);
    if (sbytes) memcpy(rbuf, sbuf, sbytes);
    return;
  }
  const char* sp = static_cast<const char*>(sbuf);
  char* rp = static_cast<char*>(rbuf);
  const int sfd = sbytes ? impl_->fd[dst] : -1, rfd = rbytes ? impl_->fd[src] : -1;
  while (sbytes || rbytes) {
    pollfd pf[2];
    int n = 0, si = -1, ri = -1;
    if (sbytes) { pf[n] = {sfd, POLLOUT, 0}; si = n++; }
    if (rbytes) { pf[n] = {rfd, POLLIN, 0}; ri = n++; }
    const int rc = poll(pf, n, kTimeoutMs);
    if (rc < 0 && errno == EINTR) continue;
    SPN_ABORT("spn::comm_t: exchange timed out", rc <= 0);
    if (si >= 0 && (pf[si].revents & (POLLOUT | POLLERR | POLLHUP))) {
      const ssize_t w = ::send(sfd, sp, sbytes, MSG_NOSIGNAL);
      if (w > 0) { sp += w; sbytes -= (size_t)w; }
      else if (w < 0 && errno != EAGAIN && errno != EWOULDBLOCK && errno != EINTR)
        throw exception(std::string("spn::comm_t: send failed: ") + strerror(errno));
    }
    if (ri >= 0 && (pf[ri].revents & (POLLIN | POLLERR | POLLHUP))) {
      const ssize_t r = ::recv(rfd, rp, rbytes, 0);
      if (r > 0) { rp += r; rbytes -= (size_t)r; }
      else if (r == 0) throw exception("spn::comm_t: peer closed the connection");
      else if (errno != EAGAIN && errno != EWOULDBLOCK && errno != EINTR)
        throw exception(std::string("spn::comm_t: recv failed: ") + strerror(errno));
    }
  }
}

void comm_t::alltoallv_bytes(const void* sbuf, const size_t* sc, const size_t* so, void* rbuf, const size_t* rc,
                             const size_t* ro) {
  const char* s = static_cast<const char*>(sbuf);
  char* r = static_cast<char*>(rbuf);
  const int P = size(), me = rank();
  SPN_ABORT("alltoallv self size mismatch", sc[me] != rc[me]);
  if (sc[me] >= ((size_t)16 << 20)) {  // the self part of the setup's label routing is hundreds of MB: copy (and first-touch) it in parallel
    const size_t piece = (size_t)4 << 20, pieces = (sc[me] + piece - 1) / piece;
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < pieces; ++c) {
      const size_t lo = c * piece, len = std::min(piece, sc[me] - lo);
      memcpy(r + ro[me] + lo, s + so[me] + lo, len);
    }
  } else if (sc[me]) {
    memcpy(r + ro[me], s + so[me], sc[me]);
  }
  for (int step = 1; step < P; ++step) {
    const int dst = (me + step) % P, src = (me - step + P) % P;
    sendrecv(dst, s + so[dst], sc[dst], src, r + ro[src], rc[src]);
  }
}

void comm_t::Barrier() {
  int token = 1;
  Allreduce(token);
}

bool comm_t::deviceReady() const { return impl_ && impl_->nccl != nullptr; }

spn_comm_t comm_t::device() {
  SPN_ABORT("comm_t not initialised", !impl_);
  if (!impl_->nccl) {
    unsigned char id[SPN_UNIQUE_ID_BYTES];
    memset(id, 0, sizeof(id));
    if (rank() == 0) SPN_CHECK(spn_comm_unique_id(id));
    Bcast(id, SPN_UNIQUE_ID_BYTES, 0);
    SPN_CHECK(spn_comm_init(&impl_->nccl, id, rank(), size()));
  }
  return impl_->nccl;
}

void comm_t::AllreduceDevice(double* d_buf, int count, spn_stream_t stream) {
  if (size() == 1) return;  // the reference's BS3-5 never reduce across ranks; at one rank this is the identity
  SPN_CHECK(spn_comm_allreduce_sum_f64(device(), d_buf, count, stream));
}

}  // namespace spn
