// Self-test of the MPI stand-in (include/shim/mpi.h, streamparanumal_b200/shim/mpi_shim.cpp) under `spnrun -np P`, P >= 1:
// the semantics the reference's comm_t and ogs setup rely on (include/comm.hpp:30-575, libs/ogs/ogsSetup.cpp, ogsPairwise.cpp,
// ogsCrystalRouter.cpp) — non-overtaking point-to-point matching by (communicator, source, tag) with messages of different
// sizes on one tag, zero-length messages, large messages in both directions at once, communicators that do not see each
// other's traffic, ragged collectives with empty parts, reductions in rank order for every type / op, MPI_IN_PLACE.
// Prints "mpi shim ok (P ranks)" on rank 0; any failed check aborts with its line number.
#include <mpi.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

static int g_rank = 0, g_size = 1;
#define CHECK(cond)                                                                         \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      fprintf(stderr, "rank %d: check failed at line %d: %s\n", g_rank, __LINE__, #cond);   \
      abort();                                                                              \
    }                                                                                       \
  } while (0)

static void point_to_point(MPI_Comm comm) {
  const int P = g_size, me = g_rank, right = (me + 1) % P, left = (me + P - 1) % P;
  // three messages of different sizes on ONE tag to the right neighbour, received in order (the crystal router's and the
  // pairwise setup's pattern: count, then payloads, all tagged with the sender's rank)
  {
    int count = 5 + me;
    std::vector<long long> ids(count);
    std::vector<int> ranks(count);
    for (int i = 0; i < count; ++i) { ids[i] = 1000LL * me + i; ranks[i] = me * 7 + i; }
    MPI_Request req[6];
    MPI_Isend(&count, 1, MPI_INT, right, me, comm, &req[0]);
    MPI_Isend(ids.data(), count, MPI_LONG_LONG_INT, right, me, comm, &req[1]);
    MPI_Isend(ranks.data(), count, MPI_INT, right, me, comm, &req[2]);
    int rcount = -1;
    MPI_Irecv(&rcount, 1, MPI_INT, left, left, comm, &req[3]);
    MPI_Wait(&req[3], MPI_STATUS_IGNORE);
    CHECK(rcount == 5 + left);
    std::vector<long long> rids(rcount);
    std::vector<int> rranks(rcount);
    MPI_Irecv(rids.data(), rcount, MPI_LONG_LONG_INT, left, left, comm, &req[4]);
    MPI_Irecv(rranks.data(), rcount, MPI_INT, left, left, comm, &req[5]);
    MPI_Waitall(6, req, MPI_STATUSES_IGNORE);
    for (int i = 0; i < rcount; ++i) CHECK(rids[i] == 1000LL * left + i && rranks[i] == left * 7 + i);
  }
  // zero-length message keeps its place in the order
  {
    int a = me + 1, b = -1, c = -1;
    MPI_Request req[4];
    MPI_Isend(&a, 0, MPI_INT, right, 9, comm, &req[0]);
    MPI_Isend(&a, 1, MPI_INT, right, 9, comm, &req[1]);
    MPI_Irecv(&b, 0, MPI_INT, left, 9, comm, &req[2]);
    MPI_Irecv(&c, 1, MPI_INT, left, 9, comm, &req[3]);
    MPI_Waitall(4, req, MPI_STATUSES_IGNORE);
    CHECK(b == -1 && c == left + 1);
  }
  // receives posted before the sends, different tags matched out of arrival order
  {
    int x = -1, y = -1;
    MPI_Request req[2];
    MPI_Irecv(&y, 1, MPI_INT, left, 22, comm, &req[0]);
    MPI_Irecv(&x, 1, MPI_INT, left, 21, comm, &req[1]);
    const int s1 = 100 + me, s2 = 200 + me;
    MPI_Send(&s1, 1, MPI_INT, right, 21, comm);
    MPI_Send(&s2, 1, MPI_INT, right, 22, comm);
    MPI_Waitall(2, req, MPI_STATUSES_IGNORE);
    CHECK(x == 100 + left && y == 200 + left);
  }
  // 24 MB in both directions at once (more than any socket buffer: a send must keep receiving while it is blocked)
  {
    const int n = 3 << 20;
    std::vector<double> out(n), in(n, -1.0);
    for (int i = 0; i < n; ++i) out[i] = me + i * 1e-6;
    MPI_Request req[2];
    MPI_Irecv(in.data(), n, MPI_DOUBLE, left, 5, comm, &req[0]);
    MPI_Isend(out.data(), n, MPI_DOUBLE, right, 5, comm, &req[1]);
    MPI_Waitall(2, req, MPI_STATUSES_IGNORE);
    for (int i = 0; i < n; i += 4097) CHECK(in[i] == left + i * 1e-6);
    CHECK(in[n - 1] == left + (n - 1) * 1e-6);
  }
  // a contiguous derived type (comm_t sends structs this way: include/comm.hpp:80-110)
  {
    struct Rec { long long id; int a, b; };
    MPI_Datatype t;
    MPI_Type_contiguous((int)sizeof(Rec), MPI_CHAR, &t);
    MPI_Type_commit(&t);
    Rec s[2] = {{me * 10LL, me, -me}, {me * 10LL + 1, me + 1, -me - 1}}, r[2];
    MPI_Request req[2];
    MPI_Irecv(r, 2, t, left, 3, comm, &req[0]);
    MPI_Isend(s, 2, t, right, 3, comm, &req[1]);
    MPI_Waitall(2, req, MPI_STATUSES_IGNORE);
    CHECK(r[0].id == left * 10LL && r[1].a == left + 1 && r[1].b == -left - 1);
    MPI_Type_free(&t);
  }
}

static void communicators() {
  const int P = g_size, me = g_rank;
  MPI_Comm dup, half;
  MPI_Comm_dup(MPI_COMM_WORLD, &dup);
  // the same (source, tag) on two communicators: each receive takes the message of ITS communicator, whatever the order
  const int right = (me + 1) % P, left = (me + P - 1) % P;
  int a = 1000 + me, b = 2000 + me, ra = -1, rb = -1;
  MPI_Request req[4];
  MPI_Isend(&a, 1, MPI_INT, right, 4, MPI_COMM_WORLD, &req[0]);
  MPI_Isend(&b, 1, MPI_INT, right, 4, dup, &req[1]);
  MPI_Irecv(&rb, 1, MPI_INT, left, 4, dup, &req[2]);
  MPI_Irecv(&ra, 1, MPI_INT, left, 4, MPI_COMM_WORLD, &req[3]);
  MPI_Waitall(4, req, MPI_STATUSES_IGNORE);
  CHECK(ra == 1000 + left && rb == 2000 + left);
  // split by parity, ranks ordered by descending world rank (key)
  MPI_Comm_split(MPI_COMM_WORLD, me & 1, -me, &half);
  int hr, hs;
  MPI_Comm_rank(half, &hr);
  MPI_Comm_size(half, &hs);
  const int members = (P + 1 - (me & 1)) / 2;
  CHECK(hs == members);
  std::vector<int> all(hs);
  MPI_Allgather(&me, 1, MPI_INT, all.data(), 1, MPI_INT, half);
  for (int i = 0; i < hs; ++i) CHECK((all[i] & 1) == (me & 1) && (i == 0 || all[i] < all[i - 1]));
  CHECK(all[hr] == me);
  int sum = me;
  MPI_Allreduce(MPI_IN_PLACE, &sum, 1, MPI_INT, MPI_SUM, half);
  CHECK(sum == std::accumulate(all.begin(), all.end(), 0));
  MPI_Comm_free(&half);
  MPI_Comm_free(&dup);
  CHECK(half == MPI_COMM_NULL && dup == MPI_COMM_NULL);
}

static void collectives(MPI_Comm comm) {
  const int P = g_size, me = g_rank;
  MPI_Barrier(comm);
  // reductions: every op, three types, in place and out of place
  {
    double d[2] = {me + 0.5, -(double)me}, dm[2];
    MPI_Allreduce(d, dm, 2, MPI_DOUBLE, MPI_MAX, comm);
    CHECK(dm[0] == P - 0.5 && dm[1] == 0.0);
    MPI_Allreduce(d, dm, 2, MPI_DOUBLE, MPI_MIN, comm);
    CHECK(dm[0] == 0.5 && dm[1] == -(double)(P - 1));
    long long s = me + 1, prod = (me % 3) + 1;
    MPI_Allreduce(MPI_IN_PLACE, &s, 1, MPI_LONG_LONG_INT, MPI_SUM, comm);
    CHECK(s == (long long)P * (P + 1) / 2);
    MPI_Allreduce(MPI_IN_PLACE, &prod, 1, MPI_LONG_LONG_INT, MPI_PROD, comm);
    long long expect = 1;
    for (int r = 0; r < P; ++r) expect *= (r % 3) + 1;
    CHECK(prod == expect);
    int flags[3] = {me == 0, 1, me & 1}, land[3], lor[3], lxor[3];
    MPI_Allreduce(flags, land, 3, MPI_INT, MPI_LAND, comm);
    MPI_Allreduce(flags, lor, 3, MPI_INT, MPI_LOR, comm);
    MPI_Allreduce(flags, lxor, 3, MPI_INT, MPI_LXOR, comm);
    CHECK(land[0] == (P == 1) && land[1] == 1 && lor[0] == 1 && lor[2] == (P > 1) && lxor[2] == ((P / 2) & 1));
    // floating-point sums fold in rank order: identical on every rank, equal to the sequential left fold
    float f = 1.0f / (float)(me + 3), fs = 0.0f, seq = 0.0f;
    MPI_Allreduce(&f, &fs, 1, MPI_FLOAT, MPI_SUM, comm);
    for (int r = 0; r < P; ++r) seq = (r == 0) ? 1.0f / 3.0f : seq + 1.0f / (float)(r + 3);
    CHECK(memcmp(&fs, &seq, sizeof(float)) == 0);
    int red = me + 1, out = -1;
    MPI_Reduce(&red, &out, 1, MPI_INT, MPI_SUM, P - 1, comm);
    CHECK(me != P - 1 || out == P * (P + 1) / 2);
    MPI_Request rq;
    int ia = me, ib = -1;
    MPI_Iallreduce(&ia, &ib, 1, MPI_INT, MPI_MAX, comm, &rq);
    MPI_Wait(&rq, MPI_STATUS_IGNORE);
    CHECK(ib == P - 1);
    long long sc = me + 1, pre = 0;
    MPI_Scan(&sc, &pre, 1, MPI_LONG_LONG_INT, MPI_SUM, comm);
    CHECK(pre == (long long)(me + 1) * (me + 2) / 2);
  }
  // broadcast, gather(v), scatter(v), allgather(v)
  {
    long long token[3] = {0, 0, 0};
    if (me == P - 1) { token[0] = 7; token[1] = -8; token[2] = 1LL << 40; }
    MPI_Bcast(token, 3, MPI_LONG_LONG_INT, P - 1, comm);
    CHECK(token[0] == 7 && token[1] == -8 && token[2] == (1LL << 40));
    std::vector<int> counts(P), displs(P + 1, 0);
    for (int r = 0; r < P; ++r) { counts[r] = (r * 3) % 4; displs[r + 1] = displs[r] + counts[r]; }  // some ranks contribute nothing
    std::vector<int> mine(counts[me]);
    for (int i = 0; i < counts[me]; ++i) mine[i] = me * 100 + i;
    std::vector<int> gathered(displs[P] + 1, -1);
    MPI_Gatherv(mine.data(), counts[me], MPI_INT, gathered.data(), counts.data(), displs.data(), MPI_INT, 0, comm);
    if (me == 0)
      for (int r = 0; r < P; ++r)
        for (int i = 0; i < counts[r]; ++i) CHECK(gathered[displs[r] + i] == r * 100 + i);
    std::vector<int> everyone(displs[P] + 1, -1);
    MPI_Allgatherv(mine.data(), counts[me], MPI_INT, everyone.data(), counts.data(), displs.data(), MPI_INT, comm);
    for (int r = 0; r < P; ++r)
      for (int i = 0; i < counts[r]; ++i) CHECK(everyone[displs[r] + i] == r * 100 + i);
    std::vector<int> back(counts[me] + 1, -1);
    MPI_Scatterv(everyone.data(), counts.data(), displs.data(), MPI_INT, back.data(), counts[me], MPI_INT, P / 2, comm);
    for (int i = 0; i < counts[me]; ++i) CHECK(back[i] == me * 100 + i);
    double pair[2] = {me * 1.0, me * 2.0};
    std::vector<double> g(2 * P, -1.0), piece(2, -1.0);
    MPI_Gather(pair, 2, MPI_DOUBLE, g.data(), 2, MPI_DOUBLE, P - 1, comm);
    if (me == P - 1)
      for (int r = 0; r < P; ++r) CHECK(g[2 * r] == r && g[2 * r + 1] == 2.0 * r);
    MPI_Bcast(g.data(), 2 * P, MPI_DOUBLE, P - 1, comm);
    MPI_Scatter(g.data(), 2, MPI_DOUBLE, piece.data(), 2, MPI_DOUBLE, 0, comm);
    CHECK(piece[0] == me && piece[1] == 2.0 * me);
  }
  // all-to-all: fixed and ragged (rank r sends (r + d) % 3 entries to rank d; zero-length parts included), blocking and not
  {
    std::vector<int> s(P), r(P, -1);
    for (int d = 0; d < P; ++d) s[d] = me * 1000 + d;
    MPI_Alltoall(s.data(), 1, MPI_INT, r.data(), 1, MPI_INT, comm);
    for (int q = 0; q < P; ++q) CHECK(r[q] == q * 1000 + me);
    std::vector<int> sc(P), sd(P + 1, 0), rc(P), rd(P + 1, 0);
    for (int d = 0; d < P; ++d) { sc[d] = (me + d) % 3; sd[d + 1] = sd[d] + sc[d]; }
    for (int q = 0; q < P; ++q) { rc[q] = (q + me) % 3; rd[q + 1] = rd[q] + rc[q]; }
    std::vector<long long> sv(sd[P] + 1), rv(rd[P] + 1, -1);
    for (int d = 0; d < P; ++d)
      for (int i = 0; i < sc[d]; ++i) sv[sd[d] + i] = ((long long)me << 32) | (unsigned)(d * 10 + i);
    for (int pass = 0; pass < 2; ++pass) {
      std::fill(rv.begin(), rv.end(), -1);
      if (pass == 0) {
        MPI_Alltoallv(sv.data(), sc.data(), sd.data(), MPI_LONG_LONG_INT, rv.data(), rc.data(), rd.data(), MPI_LONG_LONG_INT, comm);
      } else {
        MPI_Request rq;
        MPI_Ialltoallv(sv.data(), sc.data(), sd.data(), MPI_LONG_LONG_INT, rv.data(), rc.data(), rd.data(), MPI_LONG_LONG_INT, comm, &rq);
        MPI_Wait(&rq, MPI_STATUS_IGNORE);
      }
      for (int q = 0; q < P; ++q)
        for (int i = 0; i < rc[q]; ++i) CHECK(rv[rd[q] + i] == (((long long)q << 32) | (unsigned)(me * 10 + i)));
    }
  }
}

int main(int argc, char** argv) {
  MPI_Init(&argc, &argv);
  MPI_Comm_rank(MPI_COMM_WORLD, &g_rank);
  MPI_Comm_size(MPI_COMM_WORLD, &g_size);
  char name[MPI_MAX_PROCESSOR_NAME];
  int len = 0;
  MPI_Get_processor_name(name, &len);
  CHECK(len > 0 && len == (int)strlen(name));
  for (int round = 0; round < 3; ++round) {  // (repeated: nothing may be left over in the queues between rounds)
    point_to_point(MPI_COMM_WORLD);
    communicators();
    collectives(MPI_COMM_WORLD);
  }
  MPI_Barrier(MPI_COMM_WORLD);
  if (g_rank == 0) printf("mpi shim ok (%d ranks)\n", g_size);
  MPI_Finalize();
  return 0;
}
