// Implementation of include/shim/mpi.h: the MPI calls the reference's comm_t makes (include/comm.hpp:95-575), for the ranks
// of ONE node.  This image has no MPI; the reference's sources compile against the header and link against this file
// unchanged, and `spnrun -np P <exe>` (the mpirun stand-in) starts P copies with SPN_RANK / SPN_SIZE / SPN_RENDEZVOUS set.
//   P = 1 (no launcher): every collective is the single-rank identity, a message a rank sends to itself is matched with
//   the receive of the same (communicator, tag) in posting order.
//   P > 1: a full mesh of Unix-domain stream sockets.  Messages are framed (communicator, tag, bytes) and sent eagerly; a
//   blocked sender keeps draining its own incoming sockets, so two ranks that send to each other first never deadlock.
//   Receives match by (communicator, source, tag) in arrival order.  Collectives are built from these point-to-point
//   messages on reserved tags; reductions fold in rank order, so every result is deterministic.
// Host buffers only: device buffers of the native path never come through here (they move over NVLink / NCCL).
#include <mpi.h>

#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <execinfo.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <string>
#include <thread>
#include <vector>

namespace {

inline size_t bytes_of(MPI_Datatype t) { return (size_t)(t & 0xffff) + (((size_t)t >> 24) << 16); }
inline int kind_of(MPI_Datatype t) { return (t >> 16) & 0xff; }

[[noreturn]] void fail(const std::string& what) {
  fprintf(stderr, "mpi shim: %s\n", what.c_str());
  if (getenv("SPN_MPI_BACKTRACE")) {
    void* frames[48];
    backtrace_symbols_fd(frames, backtrace(frames, 48), 2);
  }
  abort();
}

// ---- world ------------------------------------------------------------------------------------------------------
int g_rank = 0, g_size = 1;
bool g_init = false;
std::vector<int> g_fd;  // socket to every other world rank
std::string g_sock_path;

struct Frame {  // message header on the wire
  int comm, tag;
  uint64_t bytes;
};
struct Message {
  int comm, tag, src;  // src: world rank
  std::vector<char> data;
};
struct Incoming {  // per-peer reassembly state
  Frame hdr;
  size_t have_hdr = 0;
  std::vector<char> data;
  size_t have = 0;
};
struct PendingRecv {
  int comm, tag, src;  // src: world rank or MPI_ANY_SOURCE
  void* buf;
  size_t bytes;
  int request;
};
std::vector<Incoming> g_in;
std::deque<Message> g_unexpected;
std::deque<PendingRecv> g_posted;
std::vector<char> g_done(1, 1);  // request handle -> complete; handle 0 = MPI_REQUEST_NULL

int new_request(bool done) {
  g_done.push_back(done ? 1 : 0);
  return (int)g_done.size() - 1;
}

// ---- communicators: ordered lists of world ranks -------------------------------------------------------------------
struct Comm {
  std::vector<int> ranks;  // world rank of every member
  int me = 0;              // my index
};
std::map<int, Comm> g_comms;
int g_next_comm = 2;

Comm& comm_of(MPI_Comm c) {
  auto it = g_comms.find(c);
  if (it == g_comms.end()) fail("use of an unknown or freed communicator");
  return it->second;
}

void deliver(Message&& m) {
  for (auto it = g_posted.begin(); it != g_posted.end(); ++it) {
    if (it->comm == m.comm && (it->tag == m.tag || it->tag == MPI_ANY_TAG) && (it->src == m.src || it->src == MPI_ANY_SOURCE)) {
      if (m.data.size() > it->bytes)
        fail("message longer than the posted receive (comm " + std::to_string(m.comm) + ", source " + std::to_string(m.src) + ", tag " +
             std::to_string(m.tag) + ": " + std::to_string(m.data.size()) + " bytes sent, " + std::to_string(it->bytes) + " posted)");
      if (!m.data.empty()) memcpy(it->buf, m.data.data(), m.data.size());
      g_done[it->request] = 1;
      g_posted.erase(it);
      return;
    }
  }
  g_unexpected.push_back(std::move(m));
}

bool try_receive(int comm, int src, int tag, void* buf, size_t bytes) {
  for (auto it = g_unexpected.begin(); it != g_unexpected.end(); ++it) {
    if (it->comm == comm && (it->tag == tag || tag == MPI_ANY_TAG) && (it->src == src || src == MPI_ANY_SOURCE)) {
      if (it->data.size() > bytes)
        fail("message longer than the receive buffer (comm " + std::to_string(comm) + ", source " + std::to_string(it->src) + ", tag " +
             std::to_string(it->tag) + ": " + std::to_string(it->data.size()) + " bytes sent, " + std::to_string(bytes) + " posted)");
      if (!it->data.empty()) memcpy(buf, it->data.data(), it->data.size());
      g_unexpected.erase(it);
      return true;
    }
  }
  return false;
}

// read whatever has arrived from peer r (non-blocking); completed messages are delivered
void drain(int r) {
  Incoming& in = g_in[r];
  while (g_fd[r] >= 0) {
    if (in.have_hdr < sizeof(Frame)) {
      const ssize_t n = ::recv(g_fd[r], reinterpret_cast<char*>(&in.hdr) + in.have_hdr, sizeof(Frame) - in.have_hdr, MSG_DONTWAIT);
      if (n == 0) {
        // the peer finished (its MPI_Finalize closed the socket): fine between messages, an error inside one
        if (in.have_hdr != 0) fail("rank " + std::to_string(r) + " closed its connection in the middle of a message");
        close(g_fd[r]);
        g_fd[r] = -1;
        return;
      }
      if (n < 0) {
        if (errno == EAGAIN || errno == EWOULDBLOCK) return;
        if (errno == EINTR) continue;
        fail(std::string("recv failed: ") + strerror(errno));
      }
      in.have_hdr += (size_t)n;
      if (in.have_hdr < sizeof(Frame)) continue;
      in.data.resize((size_t)in.hdr.bytes);
      in.have = 0;
    }
    if (in.have < in.data.size()) {
      const ssize_t n = ::recv(g_fd[r], in.data.data() + in.have, in.data.size() - in.have, MSG_DONTWAIT);
      if (n == 0) fail("rank " + std::to_string(r) + " closed its connection");
      if (n < 0) {
        if (errno == EAGAIN || errno == EWOULDBLOCK) return;
        if (errno == EINTR) continue;
        fail(std::string("recv failed: ") + strerror(errno));
      }
      in.have += (size_t)n;
      if (in.have < in.data.size()) continue;
    }
    Message m{in.hdr.comm, in.hdr.tag, r, std::move(in.data)};
    in = Incoming();
    deliver(std::move(m));
  }
}

// one step of the progress engine: wait (up to timeout_ms) for any socket to become readable and drain it
void progress(int timeout_ms) {
  if (g_size == 1) return;
  std::vector<pollfd> pf;
  std::vector<int> who;
  for (int r = 0; r < g_size; ++r)
    if (r != g_rank && g_fd[r] >= 0) { pf.push_back({g_fd[r], POLLIN, 0}); who.push_back(r); }
  if (pf.empty()) fail("waiting for a message, but every other rank has finished");
  const int rc = poll(pf.data(), pf.size(), timeout_ms);
  if (rc < 0 && errno != EINTR) fail("poll failed");
  if (rc == 0) fail("timed out waiting for a message");
  for (size_t i = 0; i < pf.size(); ++i)
    if (pf[i].revents & (POLLIN | POLLHUP | POLLERR)) drain(who[i]);
}

// eager send with progress: while the socket to `dst` is full, keep draining what others send to us
void send_bytes(int dst, const void* p, size_t n) {
  if (g_fd[dst] < 0) fail("send to rank " + std::to_string(dst) + ", which has already finished");
  const char* c = static_cast<const char*>(p);
  while (n) {
    const ssize_t w = ::send(g_fd[dst], c, n, MSG_NOSIGNAL | MSG_DONTWAIT);
    if (w > 0) { c += w; n -= (size_t)w; continue; }
    if (w < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) {
      std::vector<pollfd> pf;
      std::vector<int> who;
      pf.push_back({g_fd[dst], POLLOUT, 0});
      who.push_back(-1);
      for (int r = 0; r < g_size; ++r)
        if (r != g_rank && g_fd[r] >= 0) { pf.push_back({g_fd[r], POLLIN, 0}); who.push_back(r); }
      if (poll(pf.data(), pf.size(), 600000) <= 0) fail("send timed out");
      for (size_t i = 1; i < pf.size(); ++i)
        if (pf[i].revents & (POLLIN | POLLHUP | POLLERR)) drain(who[i]);
      continue;
    }
    if (w < 0 && errno == EINTR) continue;
    fail(std::string("send failed: ") + strerror(errno));
  }
}

bool g_trace = getenv("SPN_MPI_TRACE") != nullptr;
void send_message(int comm, int dst_world, int tag, const void* buf, size_t bytes) {
  if (g_trace) fprintf(stderr, "[%d] send -> %d comm %d tag %d bytes %zu (first int %d)\n", g_rank, dst_world, comm, tag, bytes, bytes >= 4 ? *(const int*)buf : 0);
  if (dst_world == g_rank) {
    Message m{comm, tag, g_rank, std::vector<char>((const char*)buf, (const char*)buf + bytes)};
    deliver(std::move(m));
    return;
  }
  Frame f{comm, tag, (uint64_t)bytes};
  send_bytes(dst_world, &f, sizeof(f));
  if (bytes) send_bytes(dst_world, buf, bytes);
}

void wait_request(int req) {
  if (req == MPI_REQUEST_NULL) return;
  while (!g_done[req]) {
    if (g_size == 1) fail("a receive would block forever (no matching send on one rank)");
    progress(600000);
  }
}

void recv_message(int comm, int src_world, int tag, void* buf, size_t bytes) {
  if (try_receive(comm, src_world, tag, buf, bytes)) return;
  const int req = new_request(false);
  g_posted.push_back(PendingRecv{comm, tag, src_world, buf, bytes, req});
  wait_request(req);
}

sockaddr_un make_addr(const std::string& path) {
  sockaddr_un a;
  memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  if (path.size() >= sizeof(a.sun_path)) fail("rendezvous path too long: " + path);
  strncpy(a.sun_path, path.c_str(), sizeof(a.sun_path) - 1);
  return a;
}

void connect_world() {
  auto env = [](const char* a, const char* b) -> const char* {
    const char* v = getenv(a);
    return v ? v : getenv(b);
  };
  const char* r = env("SPN_RANK", "RANK");
  const char* s = env("SPN_SIZE", "WORLD_SIZE");
  g_rank = r ? atoi(r) : 0;
  g_size = s ? atoi(s) : 1;
  if (g_size < 1 || g_rank < 0 || g_rank >= g_size) fail("bad SPN_RANK / SPN_SIZE");
  Comm world;
  for (int i = 0; i < g_size; ++i) world.ranks.push_back(i);
  world.me = g_rank;
  g_comms[MPI_COMM_WORLD] = world;
  g_fd.assign((size_t)g_size, -1);
  g_in.assign((size_t)g_size, Incoming());
  if (g_size == 1) return;
  std::string dir;
  if (const char* d = getenv("SPN_RENDEZVOUS")) dir = d;
  else {
    const char* port = getenv("MASTER_PORT");
    dir = std::string("/tmp/spn_mpi_") + std::to_string((long)getuid()) + "_" + (port ? port : "0") + "_p" + std::to_string((long)getppid());
  }
  mkdir(dir.c_str(), 0700);
  g_sock_path = dir + "/mpi" + std::to_string(g_rank) + ".sock";
  unlink(g_sock_path.c_str());
  const int lfd = socket(AF_UNIX, SOCK_STREAM, 0);
  if (lfd < 0) fail("socket() failed");
  sockaddr_un me = make_addr(g_sock_path);
  if (bind(lfd, (sockaddr*)&me, sizeof(me)) != 0) fail(std::string("bind failed: ") + strerror(errno));
  if (listen(lfd, g_size) != 0) fail("listen failed");
  const auto t0 = std::chrono::steady_clock::now();
  for (int peer = 0; peer < g_rank; ++peer) {  // connect to every lower rank
    const sockaddr_un addr = make_addr(dir + "/mpi" + std::to_string(peer) + ".sock");
    while (true) {
      const int fd = socket(AF_UNIX, SOCK_STREAM, 0);
      if (connect(fd, (const sockaddr*)&addr, sizeof(addr)) == 0) {
        const int32_t hello = g_rank;
        if (::send(fd, &hello, sizeof(hello), MSG_NOSIGNAL) != (ssize_t)sizeof(hello)) fail("hello failed");
        g_fd[peer] = fd;
        break;
      }
      close(fd);
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(180)) fail("rendezvous timed out");
      std::this_thread::sleep_for(std::chrono::milliseconds(10));
    }
  }
  for (int n = g_rank + 1; n < g_size; ++n) {  // accept every higher rank
    pollfd pf{lfd, POLLIN, 0};
    if (poll(&pf, 1, 180000) <= 0) fail("rendezvous timed out in accept");
    const int fd = accept(lfd, nullptr, nullptr);
    int32_t hello = -1;
    if (::recv(fd, &hello, sizeof(hello), MSG_WAITALL) != (ssize_t)sizeof(hello) || hello <= g_rank || hello >= g_size || g_fd[hello] >= 0)
      fail("bad rendezvous hello");
    g_fd[hello] = fd;
  }
  close(lfd);
}

// ---- reductions ------------------------------------------------------------------------------------------------------
template <typename T>
void reduce_typed(T* acc, const T* in, int count, MPI_Op op) {
  for (int i = 0; i < count; ++i) {
    switch (op) {
      case MPI_MAX: acc[i] = std::max(acc[i], in[i]); break;
      case MPI_MIN: acc[i] = std::min(acc[i], in[i]); break;
      case MPI_SUM: acc[i] = acc[i] + in[i]; break;
      case MPI_PROD: acc[i] = acc[i] * in[i]; break;
      case MPI_LAND: acc[i] = (T)((acc[i] != T(0)) && (in[i] != T(0))); break;
      case MPI_LOR: acc[i] = (T)((acc[i] != T(0)) || (in[i] != T(0))); break;
      case MPI_LXOR: acc[i] = (T)((acc[i] != T(0)) != (in[i] != T(0))); break;
      default: fail("unknown reduction op");
    }
  }
}
void reduce_into(void* acc, const void* in, int count, MPI_Datatype type, MPI_Op op) {
  switch (kind_of(type)) {
    case 1: reduce_typed((char*)acc, (const char*)in, count, op); break;
    case 2: reduce_typed((int*)acc, (const int*)in, count, op); break;
    case 3: reduce_typed((long long*)acc, (const long long*)in, count, op); break;
    case 4: reduce_typed((float*)acc, (const float*)in, count, op); break;
    case 5: reduce_typed((double*)acc, (const double*)in, count, op); break;
    default: fail("reduction on a derived datatype");
  }
}

constexpr int kTagColl = -1000;  // reserved tag space of the collectives (user tags are >= 0)

// every rank's `bytes` bytes gathered, in rank order, at every rank
std::vector<char> allgather_bytes(MPI_Comm c, const void* mine, size_t bytes) {
  Comm& cm = comm_of(c);
  const int P = (int)cm.ranks.size();
  std::vector<char> all((size_t)P * bytes);
  for (int i = 0; i < P; ++i)
    if (i != cm.me) send_message(c, cm.ranks[i], kTagColl - 1, mine, bytes);
  if (bytes) memcpy(all.data() + (size_t)cm.me * bytes, mine, bytes);
  for (int i = 0; i < P; ++i)
    if (i != cm.me) recv_message(c, cm.ranks[i], kTagColl - 1, all.data() + (size_t)i * bytes, bytes);
  return all;
}

}  // namespace

extern "C" {

int MPI_Init(int*, char***) {
  if (!g_init) { connect_world(); g_init = true; }
  return MPI_SUCCESS;
}
int MPI_Finalize(void) {
  if (g_size > 1) {
    MPI_Barrier(MPI_COMM_WORLD);
    for (int fd : g_fd)
      if (fd >= 0) close(fd);
    if (!g_sock_path.empty()) unlink(g_sock_path.c_str());
  }
  return MPI_SUCCESS;
}
int MPI_Abort(MPI_Comm, int errorcode) { _exit(errorcode ? errorcode : 1); }
int MPI_Comm_rank(MPI_Comm c, int* rank) {
  if (!g_init) MPI_Init(nullptr, nullptr);
  *rank = comm_of(c).me;
  return MPI_SUCCESS;
}
int MPI_Comm_size(MPI_Comm c, int* size) {
  if (!g_init) MPI_Init(nullptr, nullptr);
  *size = (int)comm_of(c).ranks.size();
  return MPI_SUCCESS;
}
int MPI_Comm_dup(MPI_Comm c, MPI_Comm* newcomm) {  // collective: every member takes the same next id
  if (!g_init) MPI_Init(nullptr, nullptr);
  Comm copy = comm_of(c);
  // ids must agree across the members even when ranks created different numbers of communicators before: take the max
  int id = g_next_comm;
  if (copy.ranks.size() > 1) {
    std::vector<char> all = allgather_bytes(c, &id, sizeof(int));
    for (size_t i = 0; i < copy.ranks.size(); ++i) id = std::max(id, reinterpret_cast<int*>(all.data())[i]);
  }
  g_next_comm = id + 1;
  g_comms[id] = copy;
  *newcomm = id;
  return MPI_SUCCESS;
}
int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* newcomm) {
  if (!g_init) MPI_Init(nullptr, nullptr);
  Comm& cm = comm_of(c);
  const int P = (int)cm.ranks.size();
  int mine[3] = {color, key, g_next_comm};
  std::vector<char> all = allgather_bytes(c, mine, sizeof(mine));
  const int* a = reinterpret_cast<const int*>(all.data());
  int id = 0;
  for (int i = 0; i < P; ++i) id = std::max(id, a[3 * i + 2]);
  std::vector<std::pair<std::pair<int, int>, int>> members;  // ((key, old index), world rank) of my colour
  for (int i = 0; i < P; ++i)
    if (a[3 * i] == color) members.push_back({{a[3 * i + 1], i}, cm.ranks[i]});
  std::sort(members.begin(), members.end());
  Comm nc;
  for (size_t i = 0; i < members.size(); ++i) {
    nc.ranks.push_back(members[i].second);
    if (members[i].second == g_rank) nc.me = (int)i;
  }
  // one id per colour would do; a single id for all colours is enough because messages also match by source
  g_next_comm = id + 1;
  g_comms[id] = nc;
  *newcomm = id;
  return MPI_SUCCESS;
}
int MPI_Comm_free(MPI_Comm* c) {
  if (*c != MPI_COMM_WORLD && *c != MPI_COMM_NULL) g_comms.erase(*c);
  *c = MPI_COMM_NULL;
  return MPI_SUCCESS;
}
int MPI_Get_processor_name(char* name, int* resultlen) {
  if (gethostname(name, MPI_MAX_PROCESSOR_NAME - 1) != 0) strcpy(name, "localhost");
  name[MPI_MAX_PROCESSOR_NAME - 1] = 0;
  *resultlen = (int)strlen(name);
  return MPI_SUCCESS;
}

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype) {
  const size_t b = (size_t)count * bytes_of(oldtype);
  if (b >= ((size_t)1 << 23)) fail("MPI_Type_contiguous: type too large for this shim");
  *newtype = (MPI_Datatype)((b & 0xffff) | ((b >> 16) << 24));  // kind 0: opaque bytes
  return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype*) { return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype*) { return MPI_SUCCESS; }

// ---- point to point ---------------------------------------------------------------------------------------------------
int MPI_Send(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm c) {
  send_message(c, comm_of(c).ranks.at(dest), tag, buf, (size_t)count * bytes_of(type));
  return MPI_SUCCESS;
}
int MPI_Recv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm c, MPI_Status*) {
  const int src = source == MPI_ANY_SOURCE ? MPI_ANY_SOURCE : comm_of(c).ranks.at(source);
  recv_message(c, src, tag, buf, (size_t)count * bytes_of(type));
  return MPI_SUCCESS;
}
int MPI_Isend(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm c, MPI_Request* request) {
  send_message(c, comm_of(c).ranks.at(dest), tag, buf, (size_t)count * bytes_of(type));  // eager: complete on return
  *request = new_request(true);
  return MPI_SUCCESS;
}
int MPI_Irecv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm c, MPI_Request* request) {
  const int src = source == MPI_ANY_SOURCE ? MPI_ANY_SOURCE : comm_of(c).ranks.at(source);
  const size_t bytes = (size_t)count * bytes_of(type);
  if (g_trace) fprintf(stderr, "[%d] irecv <- %d comm %d tag %d bytes %zu\n", g_rank, src, c, tag, bytes);
  if (try_receive(c, src, tag, buf, bytes)) {
    *request = new_request(true);
  } else {
    *request = new_request(false);
    g_posted.push_back(PendingRecv{c, tag, src, buf, bytes, *request});
  }
  return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request* request, MPI_Status*) {
  wait_request(*request);
  *request = MPI_REQUEST_NULL;
  return MPI_SUCCESS;
}
int MPI_Waitall(int count, MPI_Request* requests, MPI_Status*) {
  for (int i = 0; i < count; ++i) MPI_Wait(requests + i, MPI_STATUS_IGNORE);
  return MPI_SUCCESS;
}

// ---- collectives (over the point-to-point layer; rank-ordered, deterministic) ---------------------------------------------
int MPI_Barrier(MPI_Comm c) {
  if (!g_init) MPI_Init(nullptr, nullptr);
  char token = 0;
  if (comm_of(c).ranks.size() > 1) allgather_bytes(c, &token, 1);
  return MPI_SUCCESS;
}
int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm c) {
  Comm& cm = comm_of(c);
  const size_t bytes = (size_t)count * bytes_of(type);
  if (cm.me == root) {
    for (size_t i = 0; i < cm.ranks.size(); ++i)
      if ((int)i != root) send_message(c, cm.ranks[i], kTagColl - 2, buf, bytes);
  } else {
    recv_message(c, cm.ranks[root], kTagColl - 2, buf, bytes);
  }
  return MPI_SUCCESS;
}
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm c) {
  Comm& cm = comm_of(c);
  const size_t bytes = (size_t)count * bytes_of(type);
  const void* mine = (sendbuf == MPI_IN_PLACE) ? recvbuf : sendbuf;
  if (cm.ranks.size() == 1) {
    if (mine != recvbuf && bytes) memcpy(recvbuf, mine, bytes);
    return MPI_SUCCESS;
  }
  std::vector<char> all = allgather_bytes(c, mine, bytes);
  memcpy(recvbuf, all.data(), bytes);
  for (size_t i = 1; i < cm.ranks.size(); ++i) reduce_into(recvbuf, all.data() + i * bytes, count, type, op);
  return MPI_SUCCESS;
}
int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm c) {
  Comm& cm = comm_of(c);
  const size_t bytes = (size_t)count * bytes_of(type);
  const void* mine = (sendbuf == MPI_IN_PLACE) ? recvbuf : sendbuf;
  if (cm.me != root) {
    send_message(c, cm.ranks[root], kTagColl - 3, mine, bytes);
    return MPI_SUCCESS;
  }
  std::vector<char> acc((const char*)mine, (const char*)mine + bytes), tmp(bytes);
  // fold in rank order: ranks below the root first would change the order; keep plain rank order with the root in place
  std::vector<std::vector<char>> parts(cm.ranks.size());
  for (size_t i = 0; i < cm.ranks.size(); ++i) {
    if ((int)i == root) parts[i] = acc;
    else { parts[i].resize(bytes); recv_message(c, cm.ranks[i], kTagColl - 3, parts[i].data(), bytes); }
  }
  std::vector<char> out = parts[0];
  for (size_t i = 1; i < parts.size(); ++i) reduce_into(out.data(), parts[i].data(), count, type, op);
  if (bytes) memcpy(recvbuf, out.data(), bytes);
  return MPI_SUCCESS;
}
int MPI_Iallreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm c, MPI_Request* request) {
  MPI_Allreduce(sendbuf, recvbuf, count, type, op, c);
  *request = new_request(true);
  return MPI_SUCCESS;
}
int MPI_Scan(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm c) {
  Comm& cm = comm_of(c);
  const size_t bytes = (size_t)count * bytes_of(type);
  const void* mine = (sendbuf == MPI_IN_PLACE) ? recvbuf : sendbuf;
  std::vector<char> all = cm.ranks.size() > 1 ? allgather_bytes(c, mine, bytes) : std::vector<char>((const char*)mine, (const char*)mine + bytes);
  std::vector<char> out(all.begin(), all.begin() + bytes);
  for (int i = 1; i <= cm.me; ++i) reduce_into(out.data(), all.data() + (size_t)i * bytes, count, type, op);
  if (bytes) memcpy(recvbuf, out.data(), bytes);
  return MPI_SUCCESS;
}
int MPI_Gatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, const int* recvcounts, const int* displs,
                MPI_Datatype recvtype, int root, MPI_Comm c) {
  Comm& cm = comm_of(c);
  const size_t sb = (size_t)sendcount * bytes_of(sendtype), rb = bytes_of(recvtype);
  if (cm.me != root) {
    send_message(c, cm.ranks[root], kTagColl - 4, sendbuf, sb);
    return MPI_SUCCESS;
  }
  for (size_t i = 0; i < cm.ranks.size(); ++i) {
    char* dst = (char*)recvbuf + (size_t)displs[i] * rb;
    if ((int)i == root) { if (sendbuf != MPI_IN_PLACE && sb) memmove(dst, sendbuf, sb); }
    else recv_message(c, cm.ranks[i], kTagColl - 4, dst, (size_t)recvcounts[i] * rb);
  }
  return MPI_SUCCESS;
}
int MPI_Gather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype, int root,
               MPI_Comm c) {
  Comm& cm = comm_of(c);
  std::vector<int> counts(cm.ranks.size(), recvcount), displs(cm.ranks.size());
  for (size_t i = 0; i < displs.size(); ++i) displs[i] = (int)i * recvcount;
  return MPI_Gatherv(sendbuf, sendcount, sendtype, recvbuf, counts.data(), displs.data(), recvtype, root, c);
}
int MPI_Scatterv(const void* sendbuf, const int* sendcounts, const int* displs, MPI_Datatype sendtype, void* recvbuf, int recvcount,
                 MPI_Datatype recvtype, int root, MPI_Comm c) {
  Comm& cm = comm_of(c);
  const size_t sbt = bytes_of(sendtype);
  if (cm.me == root) {
    for (size_t i = 0; i < cm.ranks.size(); ++i) {
      const char* src = (const char*)sendbuf + (size_t)displs[i] * sbt;
      if ((int)i == root) { if (recvbuf != MPI_IN_PLACE) memmove(recvbuf, src, (size_t)sendcounts[i] * sbt); }
      else send_message(c, cm.ranks[i], kTagColl - 5, src, (size_t)sendcounts[i] * sbt);
    }
  } else {
    recv_message(c, cm.ranks[root], kTagColl - 5, recvbuf, (size_t)recvcount * bytes_of(recvtype));
  }
  return MPI_SUCCESS;
}
int MPI_Scatter(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype, int root,
                MPI_Comm c) {
  Comm& cm = comm_of(c);
  std::vector<int> counts(cm.ranks.size(), sendcount), displs(cm.ranks.size());
  for (size_t i = 0; i < displs.size(); ++i) displs[i] = (int)i * sendcount;
  return MPI_Scatterv(sendbuf, counts.data(), displs.data(), sendtype, recvbuf, recvcount, recvtype, root, c);
}
int MPI_Alltoallv(const void* sendbuf, const int* sendcounts, const int* sdispls, MPI_Datatype sendtype, void* recvbuf,
                  const int* recvcounts, const int* rdispls, MPI_Datatype recvtype, MPI_Comm c) {
  Comm& cm = comm_of(c);
  const int P = (int)cm.ranks.size();
  const size_t sb = bytes_of(sendtype), rb = bytes_of(recvtype);
  if (sendbuf == MPI_IN_PLACE) fail("MPI_Alltoallv with MPI_IN_PLACE is not used by the reference");
  for (int step = 0; step < P; ++step) {
    const int dst = (cm.me + step) % P;
    send_message(c, cm.ranks[dst], kTagColl - 6, (const char*)sendbuf + (size_t)sdispls[dst] * sb, (size_t)sendcounts[dst] * sb);
  }
  for (int step = 0; step < P; ++step) {
    const int src = (cm.me - step + P) % P;
    recv_message(c, cm.ranks[src], kTagColl - 6, (char*)recvbuf + (size_t)rdispls[src] * rb, (size_t)recvcounts[src] * rb);
  }
  return MPI_SUCCESS;
}
int MPI_Ialltoallv(const void* sendbuf, const int* sendcounts, const int* sdispls, MPI_Datatype sendtype, void* recvbuf,
                   const int* recvcounts, const int* rdispls, MPI_Datatype recvtype, MPI_Comm c, MPI_Request* request) {
  MPI_Alltoallv(sendbuf, sendcounts, sdispls, sendtype, recvbuf, recvcounts, rdispls, recvtype, c);
  *request = new_request(true);
  return MPI_SUCCESS;
}
int MPI_Alltoall(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype,
                 MPI_Comm c) {
  const int P = (int)comm_of(c).ranks.size();
  std::vector<int> sc(P, sendcount), sd(P), rc(P, recvcount), rd(P);
  for (int i = 0; i < P; ++i) { sd[i] = i * sendcount; rd[i] = i * recvcount; }
  return MPI_Alltoallv(sendbuf, sc.data(), sd.data(), sendtype, recvbuf, rc.data(), rd.data(), recvtype, c);
}
int MPI_Allgatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, const int* recvcounts, const int* displs,
                   MPI_Datatype recvtype, MPI_Comm c) {
  Comm& cm = comm_of(c);
  const int P = (int)cm.ranks.size();
  const size_t rb = bytes_of(recvtype);
  const void* mine = (sendbuf == MPI_IN_PLACE) ? (const char*)recvbuf + (size_t)displs[cm.me] * rb : sendbuf;
  const size_t mb = (sendbuf == MPI_IN_PLACE) ? (size_t)recvcounts[cm.me] * rb : (size_t)sendcount * bytes_of(sendtype);
  for (int i = 0; i < P; ++i)
    if (i != cm.me) send_message(c, cm.ranks[i], kTagColl - 7, mine, mb);
  if (sendbuf != MPI_IN_PLACE && mb) memmove((char*)recvbuf + (size_t)displs[cm.me] * rb, mine, mb);
  for (int i = 0; i < P; ++i)
    if (i != cm.me) recv_message(c, cm.ranks[i], kTagColl - 7, (char*)recvbuf + (size_t)displs[i] * rb, (size_t)recvcounts[i] * rb);
  return MPI_SUCCESS;
}
int MPI_Allgather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype,
                  MPI_Comm c) {
  const int P = (int)comm_of(c).ranks.size();
  std::vector<int> counts(P, recvcount), displs(P);
  for (int i = 0; i < P; ++i) displs[i] = i * recvcount;
  return MPI_Allgatherv(sendbuf, sendcount, sendtype, recvbuf, counts.data(), displs.data(), recvtype, c);
}

}  // extern "C"
