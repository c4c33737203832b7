// ogs_t::Setup — builds the gather operators and the pairwise exchange plan from the node labels.
// Contract of the reference's ogsBase_t::Setup (libs/ogs/ogsSetup.cpp:87-290) and
// ogsPairwise_t (libs/ogs/ogsPairwise.cpp:176-387); the algorithm is our own:
//   1. every non-zero label is routed to a directory rank (|label| % size) in ONE alltoallv; the
//      directory groups copies by label, decides shared / local, picks the positive copy when
//      `unique`, and answers each copy with (shared flag, signed label) plus, per (label, rank) of
//      a shared label, the list of ranks holding it and whether each holds a positive copy;
//   2. rows are numbered locally (one parallel sort + linear passes);
//   3. the per-neighbour send/receive lists are ordered by LABEL on both sides, so sender and
//      receiver agree on the message layout without exchanging row indices (the reference ships
//      remote row ids through four more alltoallv rounds: ogsSetup.cpp:617-813).
#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <omp.h>
#include <parallel/algorithm>
#include <thread>

#include <cstdlib>
#include <string>

#include "ogs.hpp"

namespace spn {
namespace ogs {

size_t Sizeof(Type type) {
  switch (type) {
    case Float: return 4;
    case Double: return 8;
    case Int32: return 4;
    case Int64: return 8;
  }
  return 0;
}

namespace {

inline hlong habs(hlong x) { return x < 0 ? -x : x; }

// The reference's generator for the positive-copy pick: prim::random<int> after prim::seedRNG(rank)
// (libs/primitives/random.cpp:45-164, seeded at libs/ogs/ogsSetup.cpp:110-112): the 64-bit LCG
// X(n+1) = a X(n) + c, first state a * (~0 ^ seed) + c, value n = |int(X(n))|.
struct RefLcg {
  unsigned long long x;
  explicit RefLcg(unsigned long long seed) : x(6364136223846793005ULL * (~0ULL ^ seed) + 1ULL) {}
  int next() {
    const int v = static_cast<int>(x);
    x = 6364136223846793005ULL * x + 1ULL;
    return v < 0 ? (v == INT_MIN ? v : -v) : v;  // std::abs(int), INT_MIN left as the reference leaves it
  }
  // advance by k draws in O(log k): the k-fold composition of x -> a x + c by repeated squaring, so that a thread can
  // start at the draw of its first label group
  void skip(unsigned long long k) {
    unsigned long long curA = 6364136223846793005ULL, curC = 1ULL, accA = 1ULL, accC = 0ULL;
    for (; k; k >>= 1) {
      if (k & 1) { accC = accC * curA + curC; accA *= curA; }
      curC = (curA + 1ULL) * curC;
      curA *= curA;
    }
    x = accA * x + accC;
  }
};

// 0: members of a label group in (source rank, source order) — the stable order, the default;
// 1: members in the order the reference's own `prim::sort` leaves them in (std::sort on the values only,
//    libs/primitives/sort.cpp:176-190 built without GLIBCXX_PARALLEL; libs/ogs/ogsSetup.cpp:341-343): an exact replica
//    of the serial reference build, used to pin the maps against it (costs one more serial sort).
// The reference's own order is implementation-defined (an unstable sort; thread-count dependent with
// GLIBCXX_PARALLEL), so both are instances of its rule.  SPN_OGS_PICK=stdsort selects 1.
int g_pick_mode = -1;
int pick_mode() {
  if (g_pick_mode < 0) {
    const char* e = getenv("SPN_OGS_PICK");
    g_pick_mode = (e && std::string(e) == "stdsort") ? 1 : 0;
  }
  return g_pick_mode;
}

struct Key {
  hlong label;  // |label|
  dlong idx;    // position in the entry list
};

// Stable parallel LSD radix sort of keys[i] = {label, i} by label (11-bit digits, as many passes as the largest label
// needs).  Stable + the initial order i ascending means equal labels stay in ascending position: the (label, idx) order
// the grouping below relies on.  Replaces a comparison sort of 2^27 16-byte keys (3-5 s of the setup) by 3 streaming passes.
void sort_keys_by_label(hvec<Key>& keys) {
  const size_t n = keys.size();
  if (n < 2) return;
  hlong maxLabel = 0;
#pragma omp parallel for reduction(max : maxLabel)
  for (size_t i = 0; i < n; ++i) maxLabel = std::max(maxLabel, keys[i].label);
  constexpr int BITS = 11, BUCKETS = 1 << BITS;
  int passes = 1;
  while (passes * BITS < 63 && (maxLabel >> (passes * BITS)) != 0) ++passes;
  hvec<Key> tmp(n);
  Key* src = keys.data();
  Key* dst = tmp.data();
  const int T = omp_get_max_threads();
  std::vector<size_t> hist((size_t)T * BUCKETS);
  for (int pass = 0; pass < passes; ++pass) {
    const int shift = pass * BITS;
#pragma omp parallel num_threads(T)
    {
      const int t = omp_get_thread_num();
      const size_t lo = n * (size_t)t / T, hi = n * (size_t)(t + 1) / T;
      size_t* h = hist.data() + (size_t)t * BUCKETS;
      std::fill(h, h + BUCKETS, 0);
      for (size_t i = lo; i < hi; ++i) h[(src[i].label >> shift) & (BUCKETS - 1)]++;
#pragma omp barrier
#pragma omp single
      {
        size_t run = 0;  // bucket-major, thread-minor: thread t's share of bucket b starts after the lower threads' shares
        for (int b2 = 0; b2 < BUCKETS; ++b2)
          for (int t2 = 0; t2 < T; ++t2) {
            const size_t c = hist[(size_t)t2 * BUCKETS + b2];
            hist[(size_t)t2 * BUCKETS + b2] = run;
            run += c;
          }
      }
      for (size_t i = lo; i < hi; ++i) dst[h[(src[i].label >> shift) & (BUCKETS - 1)]++] = src[i];
    }
    std::swap(src, dst);
  }
  if (src != keys.data()) keys.swap(tmp);
}

// exclusive prefix sum of v[0..n) in place, returns the total (parallel two-pass)
template <typename T>
T exclusive_scan(T* v, size_t n) {
  const int P = omp_get_max_threads();
  std::vector<T> part((size_t)P + 1, 0);
#pragma omp parallel num_threads(P)
  {
    const int t = omp_get_thread_num();
    const size_t lo = n * (size_t)t / P, hi = n * (size_t)(t + 1) / P;
    T s = 0;
    for (size_t i = lo; i < hi; ++i) s += v[i];
    part[t + 1] = s;
#pragma omp barrier
#pragma omp single
    for (int q = 0; q < P; ++q) part[q + 1] += part[q];
    T run = part[t];
    for (size_t i = lo; i < hi; ++i) {
      const T x = v[i];
      v[i] = run;
      run += x;
    }
  }
  return part[P];
}

// Work split of [0, n) into `chunks` contiguous pieces, one per loop iteration of a static OpenMP loop: the order-
// preserving selections below count per chunk, prefix-sum the counts, and write each chunk at its offset.
inline dlong chunk_lo(dlong n, int c, int chunks) { return (dlong)((long long)n * c / chunks); }

// offsets[c] = number of i in the chunks before c with keep(i); offsets[chunks] = total
template <class Keep>
std::vector<dlong> count_selected(dlong n, int chunks, Keep keep) {
  std::vector<dlong> offsets((size_t)chunks + 1, 0);
#pragma omp parallel for schedule(static, 1)
  for (int c = 0; c < chunks; ++c) {
    dlong cnt = 0;
    for (dlong i = chunk_lo(n, c, chunks), e = chunk_lo(n, c + 1, chunks); i < e; ++i) cnt += keep(i) ? 1 : 0;
    offsets[c + 1] = cnt;
  }
  for (int c = 0; c < chunks; ++c) offsets[c + 1] += offsets[c];
  return offsets;
}

// emit(i, position) for every selected i, positions ascending with i
template <class Keep, class Emit>
void emit_selected(dlong n, int chunks, const std::vector<dlong>& offsets, Keep keep, Emit emit) {
#pragma omp parallel for schedule(static, 1)
  for (int c = 0; c < chunks; ++c) {
    dlong o = offsets[c];
    for (dlong i = chunk_lo(n, c, chunks), e = chunk_lo(n, c + 1, chunks); i < e; ++i)
      if (keep(i)) emit(i, o++);
  }
}

template <typename T>
void parallel_fill(T* v, size_t n, T value) {
#pragma omp parallel for
  for (size_t i = 0; i < n; ++i) v[i] = value;
}

// Number the label groups of one entry set and build both CSR variants.
// entries: signed labels `ids` at columns `cols` (cols strictly ascending).
// Rows = label groups in first-appearance order, groups with a positive copy first (libs/ogs/ogsSetup.cpp:492-559);
// columns ascend inside a row (stable SortByRow, :561-579).  Every pass is a parallel loop or a prefix sum.
// `sorted`: the entries already sorted by (|label|, entry) — on one rank the directory round has just produced exactly
// that order for the same entries, so the local operator does not sort again.
std::shared_ptr<ogsOperator_t> build_operator(Kind kind, dlong Ncols, const hvec<hlong>& ids, const hvec<dlong>& cols,
                                              hvec<Key>* sorted = nullptr) {
  auto op = std::make_shared<ogsOperator_t>();
  op->kind = (kind == Unsigned) ? Unsigned : Signed;
  op->Ncols = Ncols;
  const dlong n = (dlong)ids.size();
  op->nnzT = n;

  hvec<Key> keys;
  if (sorted && (dlong)sorted->size() == n) {
    keys.swap(*sorted);
  } else {
    keys.resize((size_t)n);
#pragma omp parallel for
    for (dlong i = 0; i < n; ++i) keys[i] = Key{habs(ids[i]), i};
    sort_keys_by_label(keys);  // groups contiguous, members in ascending entry (= column) order
  }

  // label groups = runs of equal labels in sorted order; groupStart[g] = sorted position of the first member
  const int chunks = std::max(1, omp_get_max_threads()) * 4;
  auto opens_group = [&](dlong i) { return i == 0 || keys[i].label != keys[i - 1].label; };
  const std::vector<dlong> groupOffsets = count_selected(n, chunks, opens_group);
  const dlong NT = groupOffsets[chunks];
  hvec<dlong> groupStart((size_t)NT + 1);
  emit_selected(n, chunks, groupOffsets, opens_group, [&](dlong i, dlong g) { groupStart[g] = i; });
  groupStart[NT] = n;
  hvec<char> groupPos((size_t)NT);  // has a positive copy on this rank
#pragma omp parallel for
  for (dlong g = 0; g < NT; ++g) {
    char pos = 0;
    for (dlong i = groupStart[g]; i < groupStart[g + 1]; ++i) pos |= (ids[keys[i].idx] > 0) ? 1 : 0;
    groupPos[g] = pos;
  }
  // first-appearance order, positive groups first: flags in ENTRY order at each group's first member, two prefix sums
  hvec<dlong> headP((size_t)n), headM((size_t)n);
  parallel_fill(headP.data(), (size_t)n, (dlong)0);
  parallel_fill(headM.data(), (size_t)n, (dlong)0);
#pragma omp parallel for
  for (dlong g = 0; g < NT; ++g) {
    const dlong first = keys[groupStart[g]].idx;  // smallest entry index of the group
    if (groupPos[g]) headP[first] = 1; else headM[first] = 1;
  }
  const dlong NP = n ? exclusive_scan(headP.data(), (size_t)n) : 0;
  if (n) exclusive_scan(headM.data(), (size_t)n);
  hvec<dlong> rowOfGroup((size_t)NT);
#pragma omp parallel for
  for (dlong g = 0; g < NT; ++g) {
    const dlong first = keys[groupStart[g]].idx;
    rowOfGroup[g] = groupPos[g] ? headP[first] : NP + headM[first];
  }
  headP.clear(); headP.shrink_to_fit();
  headM.clear(); headM.shrink_to_fit();
  op->NrowsT = NT;
  op->NrowsN = NP;

  // CSR "T": every copy.  Row lengths = group sizes; members of a group are already in ascending column order.
  op->rowLabel.resize((size_t)NT);
  op->rowStartsT.resize((size_t)NT + 1);
  op->rowStartsT[NT] = 0;
#pragma omp parallel for
  for (dlong g = 0; g < NT; ++g) {
    const dlong row = rowOfGroup[g];
    op->rowStartsT[row] = groupStart[g + 1] - groupStart[g];
    op->rowLabel[row] = keys[groupStart[g]].label;
  }
  if (NT) exclusive_scan(op->rowStartsT.data(), (size_t)NT + 1);
  op->colIdsT.resize((size_t)n);
#pragma omp parallel for
  for (dlong g = 0; g < NT; ++g) {
    dlong o = op->rowStartsT[rowOfGroup[g]];
    for (dlong i = groupStart[g]; i < groupStart[g + 1]; ++i) op->colIdsT[o++] = cols[keys[i].idx];
  }
  // CSR "N": positive copies only (libs/ogs/ogsOperator.cpp:620-643).  rowStartsN carries NrowsT+1
  // entries (rows >= NrowsN are empty) so kernels may walk the T row partition over it.
  if (op->kind == Signed) {
    op->rowStartsN.resize((size_t)NT + 1);
    op->rowStartsN[NT] = 0;
#pragma omp parallel for
    for (dlong g = 0; g < NT; ++g) {
      dlong c = 0;
      for (dlong i = groupStart[g]; i < groupStart[g + 1]; ++i) c += (ids[keys[i].idx] > 0) ? 1 : 0;
      op->rowStartsN[rowOfGroup[g]] = c;
    }
    const dlong nnzN = NT ? exclusive_scan(op->rowStartsN.data(), (size_t)NT + 1) : 0;
    op->nnzN = nnzN;
    op->colIdsN.resize((size_t)nnzN);
#pragma omp parallel for
    for (dlong g = 0; g < NT; ++g) {
      dlong o = op->rowStartsN[rowOfGroup[g]];
      for (dlong i = groupStart[g]; i < groupStart[g + 1]; ++i)
        if (ids[keys[i].idx] > 0) op->colIdsN[o++] = cols[keys[i].idx];
    }
  } else {
    op->nnzN = op->nnzT;
    op->NrowsN = op->NrowsT;
  }
  return op;
}

}  // namespace

void SetPositivePickOrder(int mode) { g_pick_mode = mode ? 1 : 0; }

void ogs_t::Free() {
  if (packedEvent) { spn_event_destroy(packedEvent); packedEvent = nullptr; }
  if (recvEvent) { spn_event_destroy(recvEvent); recvEvent = nullptr; }

  if (forkEvent) { spn_event_destroy(forkEvent); forkEvent = nullptr; }
  if (joinEvent) { spn_event_destroy(joinEvent); joinEvent = nullptr; }
  halo_forked = false;
  FreeGraphs();
  FreePeerExchange();
  for (HostStage& st : stage) {
    if (st.busy) spn_event_sync(st.d2hDone);
    if (st.h2dDone) spn_event_destroy(st.h2dDone);
    if (st.computeDone) spn_event_destroy(st.computeDone);
    if (st.d2hDone) spn_event_destroy(st.d2hDone);
    st = HostStage();
  }
  nextStage = 0;
  // every Setup makes its own streams: give them back (a platform outlives many ogs_t)
  for (spn_stream_t* st : {&commStream, &haloStream, &graphStream, &h2dStream, &d2hStream}) {
    platform.destroyStream(*st);
    *st = nullptr;
  }
  gatherLocal = nullptr;
  gatherHalo = nullptr;
  for (auto& d : data) d = ogsExchangeData_t();
  o_ownBuf.free();
  o_sendBuf.free();
  o_recvBuf.free();
  N = Ngather = NlocalT = NlocalP = NhaloT = NhaloP = 0;
  NgatherGlobal = 0;
  on_device = false;
  exchange_in_flight = false;
  halo_only = false;
}

// ---------------------------------------------------------------------------------------------
// peer-memory exchange: allocate the shared block, swap IPC handles and receive offsets (collective)
// ---------------------------------------------------------------------------------------------
void ogs_t::FreePeerExchange() {
  if (p2p.enabled) {
    // nobody may unmap or free a block a neighbour can still store into
    try {
      platform.finish();
      comm.Barrier();
    } catch (...) {
    }
  }
  for (void*& b : p2p.peerBlock)
    if (b) { spn_ipc_close_handle(b); b = nullptr; }
  p2p = PeerExchange();
  peer_exchange_in_flight = false;
}

void ogs_t::SetupPeerExchange() {
  const int size = comm.size(), rank = comm.rank();
  const char* mode_env = getenv("SPN_OGS_EXCHANGE");
  int want = !(mode_env && std::string(mode_env) == "nccl") && size <= 64 ? 1 : 0;
  for (int m = 0; m < 3; ++m)
    if ((int)data[m].sendRanks.size() > SPN_P2P_MAX_PEERS || (int)data[m].recvRanks.size() > SPN_P2P_MAX_PEERS) want = 0;
  comm.Allreduce(want, Comm::Min);
  if (!want) return;

  SPN_ABORT("peer exchange: more ranks than SPN_P2P_MAX_RANKS", size > SPN_P2P_MAX_RANKS);
  dlong maxRecv = 0;
  for (int m = 0; m < 3; ++m) maxRecv = std::max(maxRecv, data[m].Nrecv);
  // every rank lays its block out the same way: a sender computes the address of a destination's second half from the
  // common half size (ranks with different receive counts would otherwise disagree on where half 1 starts)
  comm.Allreduce(maxRecv, Comm::Max);
  auto align = [](size_t b) { return (b + 255) / 256 * 256; };
  p2p.flagBytes = align(2 * sizeof(unsigned long long) * (size_t)size);  // flags, then acks
  p2p.halfBytes = align(kPeerEntryBytes * (size_t)std::max<dlong>(maxRecv, 1));
  const size_t blockBytes = p2p.flagBytes + 2 * p2p.halfBytes;
  p2p.o_block = platform.malloc<char>(blockBytes);
  p2p.o_state = platform.malloc<char>(sizeof(spn_p2p_state_t));
  p2p.o_plan = platform.malloc<char>(3 * sizeof(spn_p2p_plan_t));
  SPN_CHECK(spn_memset(p2p.o_block.ptr(), 0, blockBytes, platform.getStream()));
  SPN_CHECK(spn_memset(p2p.o_state.ptr(), 0, sizeof(spn_p2p_state_t), platform.getStream()));
  platform.finish();

  // handles of every rank's block
  std::vector<char> mine(SPN_IPC_HANDLE_BYTES, 0), all((size_t)SPN_IPC_HANDLE_BYTES * size, 0);
  int ok = spn_ipc_get_handle(p2p.o_block.ptr(), mine.data()) == SPN_SUCCESS ? 1 : 0;
  comm.Allgather(mine.data(), all.data(), SPN_IPC_HANDLE_BYTES);
  comm.Allreduce(ok, Comm::Min);

  // neighbours in any mode
  std::vector<char> isNbr((size_t)size, 0);
  for (int m = 0; m < 3; ++m) {
    for (int r : data[m].sendRanks) isNbr[r] = 1;
    for (int r : data[m].recvRanks) isNbr[r] = 1;
  }
  p2p.peerBlock.assign((size_t)size, nullptr);
  if (ok) {
    for (int r = 0; r < size && ok; ++r) {
      if (!isNbr[r] || r == rank) continue;
      void* mapped = nullptr;
      if (spn_ipc_open_handle(all.data() + (size_t)SPN_IPC_HANDLE_BYTES * r, &mapped) != SPN_SUCCESS) ok = 0;
      else p2p.peerBlock[r] = mapped;
    }
  }
  comm.Allreduce(ok, Comm::Min);

  // address tables of the three modes: where my entries land in each destination (its recvOffsets entry for me),
  // which words to raise and which to watch
  std::vector<spn_p2p_plan_t> plans(3);
  memset(plans.data(), 0, sizeof(spn_p2p_plan_t) * 3);
  char* myBlock = p2p.o_block.ptr();
  const size_t ackBase = sizeof(unsigned long long) * (size_t)size;
  for (int m = 0; m < 3; ++m) {
    const ogsExchangeData_t& d = data[m];
    std::vector<dlong> tell((size_t)size, -1), got((size_t)size, -1);
    for (size_t i = 0; i < d.recvRanks.size(); ++i) tell[d.recvRanks[i]] = d.recvOffsets[i];
    comm.Alltoall(tell.data(), got.data());
    if (!ok) continue;  // (the collective above still has to run on every rank)
    spn_p2p_plan_t& pl = plans[m];
    pl.nsend = (int)d.sendRanks.size();
    pl.nrecv = (int)d.recvRanks.size();
    for (int i = 0; i < pl.nsend; ++i) {
      const int r = d.sendRanks[i];
      char* block = static_cast<char*>(p2p.peerBlock[r]);
      const dlong off = got[r];
      if (off < 0 || !block) { ok = 0; break; }  // the destination does not expect me: plans disagree
      pl.slot0[i] = d.sendOffsets[i];
      pl.dstRank[i] = r;
      pl.remoteOffset[i] = off;
      pl.dstHalf[0][i] = block + p2p.flagBytes;
      pl.dstHalf[1][i] = block + p2p.flagBytes + p2p.halfBytes;
      pl.flagRemote[i] = reinterpret_cast<unsigned long long*>(block) + rank;
      pl.ackLocal[i] = reinterpret_cast<const unsigned long long*>(myBlock + ackBase) + r;
    }
    pl.slot0[pl.nsend] = d.Nsend;
    for (int i = 0; i < pl.nrecv && ok; ++i) {
      const int r = d.recvRanks[i];
      char* block = static_cast<char*>(p2p.peerBlock[r]);
      if (!block) { ok = 0; break; }
      pl.flagLocal[i] = reinterpret_cast<const unsigned long long*>(myBlock) + r;
      pl.ackRemote[i] = reinterpret_cast<unsigned long long*>(block + ackBase) + rank;
    }
    pl.recvHalf[0] = myBlock + p2p.flagBytes;
    pl.recvHalf[1] = myBlock + p2p.flagBytes + p2p.halfBytes;
  }
  comm.Allreduce(ok, Comm::Min);
  if (!ok) {
    if (rank == 0) fprintf(stderr, "spn ogs: CUDA IPC peer access unavailable, halo exchange stays on NCCL send/recv\n");
    FreePeerExchange();
    return;
  }
  SPN_CHECK(spn_memcpy_h2d(p2p.o_plan.ptr(), plans.data(), sizeof(spn_p2p_plan_t) * 3, platform.getStream(), 0));
  p2p.enabled = true;
}

void ogs_t::SetupGlobalToLocalMapping(dlong* GlobalToLocal) const {
  SPN_ABORT("ogs handle is not set up.", NgatherGlobal == 0 || !gatherLocal || !gatherHalo);  // ogsSetup.cpp:835-836
#pragma omp parallel for
  for (dlong n = 0; n < N; ++n) GlobalToLocal[n] = -1;
  // the reference scatters the row numbers through both operators with NoTrans, i.e. over every copy (the T arrays)
  const ogsOperator_t* ops[2] = {gatherLocal.get(), gatherHalo.get()};
  const dlong base[2] = {0, NlocalT};
  for (int o = 0; o < 2; ++o) {
    const ogsOperator_t& op = *ops[o];
#pragma omp parallel for
    for (dlong r = 0; r < op.NrowsT; ++r)
      for (dlong g = op.rowStartsT[r]; g < op.rowStartsT[r + 1]; ++g) GlobalToLocal[op.colIdsT[g]] = base[o] + r;
  }
}

void ogs_t::AssertGatherDefined() const {
  SPN_ABORT("Gather / Scatter are not operations of a halo_t (Kind Halo): use Exchange / Combine", halo_only);
  SPN_ABORT("Gather operation not well-defined.", !gather_defined);  // libs/ogs/ogsSetup.cpp:827-830
}

void ogs_t::SetupHost(const dlong N_, hlong* ids, comm_t comm_, const Kind kind_, const bool unique_, const bool verbose) {
  Free();
  SetupPlan(N_, ids, comm_, kind_, unique_, verbose);
}

void ogs_t::Setup(const dlong N_, hlong* ids, comm_t comm_, const Kind kind_, const Method /*method: Pairwise over NCCL only*/,
                  const bool unique_, const bool verbose, platform_t& platform_) {
  Free();
  const timePoint_t start = Time();
  platform = platform_;
  SetupPlan(N_, ids, comm_, kind_, unique_, verbose);

  // device copies (reference: ogsOperator_t ctor uploads, libs/ogs/ogsOperator.cpp:618-665)
  if (!halo_only) gatherLocal->upload(platform, /*tile_plans=*/true);  // a halo_t never touches its local rows
  gatherHalo->upload(platform, /*tile_plans=*/false);
  for (int m = 0; m < 3; ++m) {
    ogsExchangeData_t& d = data[m];
    if (kind == Unsigned && m != Sym) {  // reference: data[NoTrans] = data[Trans] = data[Sym]
      d.o_sendIds = data[Sym].o_sendIds;
      d.o_cmbStarts = data[Sym].o_cmbStarts;
      d.o_cmbIds = data[Sym].o_cmbIds;
      d.o_rowSendStarts = data[Sym].o_rowSendStarts;
      d.o_rowSendDst = data[Sym].o_rowSendDst;
      d.o_slotDst = data[Sym].o_slotDst;
      continue;
    }
    if (d.Nsend) d.o_sendIds = platform.malloc<dlong>(d.sendIds);
    if (d.Nrows) d.o_cmbStarts = platform.malloc<dlong>(d.cmbStarts);
    if (d.Nrecv) d.o_cmbIds = platform.malloc<dlong>(d.cmbIds);
    if (NhaloT) d.o_rowSendStarts = platform.malloc<dlong>(d.rowSendStarts);
    if (d.Nsend) {
      d.o_rowSendDst = platform.malloc<dlong>(d.rowSendDst);
      d.o_slotDst = platform.malloc<dlong>(d.slotDst);
    }
  }
  if (comm.size() > 1) {
    commStream = platform.createStream();
    haloStream = platform.createStream(/*high_priority=*/true);
    SPN_CHECK(spn_event_create(&forkEvent));
    SPN_CHECK(spn_event_create(&joinEvent));
    SPN_CHECK(spn_event_create(&packedEvent));
    SPN_CHECK(spn_event_create(&recvEvent));
    // NCCL communicator: collective, every rank reaches this point.  SPN_OGS_NCCL=lazy leaves it to the first call that
    // sends over NCCL (StartExchange, also reached by every rank): ranks that share one GPU — the oversubscribed
    // peer-memory tests — cannot form an NCCL communicator and never need one.
    const char* ncclMode = getenv("SPN_OGS_NCCL");
    if (!(ncclMode && std::string(ncclMode) == "lazy")) comm.device();
    SetupPeerExchange();
  }
  AllocBuffer(sizeof(dfloat));
  on_device = true;

  platform.finish();
  comm.Barrier();
  if (comm.rank() == 0 && verbose)
    printf("ogs Setup Time: %g seconds.\n", ElapsedTime(start, Time()));
}

namespace {
// The setup sorts and scans run on OpenMP threads.  Launchers such as torchrun export OMP_NUM_THREADS=1 for every
// rank, which makes a 2^27-node setup three times slower than it has to be: use hardware threads / ranks of this
// node for the duration of the setup (SPN_SETUP_THREADS overrides), then restore the caller's setting.
struct SetupThreads {
  int saved;
  explicit SetupThreads(int ranks) : saved(omp_get_max_threads()) {
    const char* e = getenv("SPN_SETUP_THREADS");
    const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    const int n = e ? atoi(e) : std::max(1, hw / std::max(1, ranks));
    if (n > 0) omp_set_num_threads(n);
  }
  ~SetupThreads() { omp_set_num_threads(saved); }
};
}  // namespace

namespace {
struct PhaseTimer {  // SPN_SETUP_TIMING=1: seconds per phase of SetupPlan on rank 0
  bool on;
  timePoint_t t;
  PhaseTimer(int rank) : on(rank == 0 && getenv("SPN_SETUP_TIMING") != nullptr), t(Time()) {}
  void lap(const char* what) {
    if (!on) return;
    const timePoint_t n = Time();
    fprintf(stderr, "  ogs setup: %-38s %7.3f s\n", what, std::chrono::duration<double>(n - t).count());
    t = n;
  }
};
}  // namespace

void ogs_t::SetupPlan(const dlong N_, hlong* ids, comm_t comm_, const Kind kind_, const bool unique_, const bool verbose) {
  SetupThreads threads(comm_.size());
  PhaseTimer phase(comm_.rank());
  N = N_;
  comm = comm_;
  kind = kind_;
  unique = unique_;
  const int rank = comm.rank(), size = comm.size();

  SPN_ABORT("Invalid ogs setup requested", (kind == Unsigned && unique) || (kind == Halo && unique));
  // Kind Halo (halo_t::Setup, libs/ogs/ogsSetup.cpp:73-82): a Signed, non-unique operator of which only the rows
  // shared between ranks act — the reference builds no gatherLocal for it (ogsSetup.cpp:592-603)
  halo_only = (kind == Halo);

  const int chunks = std::max(1, omp_get_max_threads()) * 4;  // pieces of the order-preserving parallel passes

  // ---- non-zero labels, in column order ----------------------------------------------------
  auto nonzero = [&](dlong n) { return ids[n] != 0; };
  const std::vector<dlong> nzOffsets = count_selected(N, chunks, nonzero);
  const dlong Nids = nzOffsets[chunks];
  hvec<dlong> colIds((size_t)Nids);
  hvec<hlong> baseIds((size_t)Nids);
  emit_selected(N, chunks, nzOffsets, nonzero, [&](dlong n, dlong c) { colIds[c] = n; baseIds[c] = ids[n]; });

  phase.lap("non-zero labels");
  // ---- 1. directory round ------------------------------------------------------------------
  // label L lives on directory rank |L| % size; the send buffer is ordered by destination, entries of one destination
  // in column order (per-chunk histograms, destination-major / chunk-minor offsets)
  // On ONE rank the round is the identity (receive slot i = entry i): nothing is routed, copied or indexed through sendPos.
  const bool single = (size == 1);
  std::vector<int> sendCounts(size, 0), sendOffsets(size + 1, 0), recvCounts(size, 0), recvOffsets(size + 1, 0);
  hvec<dlong> sendPos;  // entry n sits at slot sendPos[n] of the send buffer (stable by dest)
  hvec<hlong> sendLabels, recvLabels;
  if (single) {
    sendCounts[0] = recvCounts[0] = Nids;
    sendOffsets[1] = recvOffsets[1] = Nids;
  } else {
    sendPos.resize((size_t)Nids);
    sendLabels.resize((size_t)Nids);
    std::vector<dlong> hist((size_t)chunks * size, 0);
#pragma omp parallel for schedule(static, 1)
    for (int c = 0; c < chunks; ++c) {
      dlong* h = hist.data() + (size_t)c * size;
      for (dlong n = chunk_lo(Nids, c, chunks), e = chunk_lo(Nids, c + 1, chunks); n < e; ++n) h[habs(baseIds[n]) % size]++;
    }
    dlong run = 0;
    for (int r = 0; r < size; ++r) {
      sendOffsets[r] = (int)run;
      for (int c = 0; c < chunks; ++c) {
        const dlong cnt = hist[(size_t)c * size + r];
        hist[(size_t)c * size + r] = run;
        run += cnt;
      }
      sendCounts[r] = (int)run - sendOffsets[r];
    }
    sendOffsets[size] = (int)run;
#pragma omp parallel for schedule(static, 1)
    for (int c = 0; c < chunks; ++c) {
      dlong* h = hist.data() + (size_t)c * size;
      for (dlong n = chunk_lo(Nids, c, chunks), e = chunk_lo(Nids, c + 1, chunks); n < e; ++n) {
        const dlong pos = h[habs(baseIds[n]) % size]++;
        sendPos[n] = pos;
        sendLabels[pos] = baseIds[n];
      }
    }
    comm.Alltoall(sendCounts.data(), recvCounts.data());
    for (int r = 0; r < size; ++r) recvOffsets[r + 1] = recvOffsets[r] + recvCounts[r];
    recvLabels.resize((size_t)recvOffsets[size]);
    comm.Alltoallv(sendLabels.data(), sendCounts.data(), sendOffsets.data(), recvLabels.data(), recvCounts.data(),
                   recvOffsets.data());
    sendLabels.clear();
    sendLabels.shrink_to_fit();
  }
  const dlong recvN = recvOffsets[size];
  const hlong* recvL = single ? baseIds.data() : recvLabels.data();  // label of every receive slot (read only)
  auto slot_of = [&](dlong n) { return single ? n : sendPos[n]; };

  phase.lap("route labels to directory ranks");
  // group the received copies by |label|; inside a group they stay in (source rank, source order)
  hvec<Key> keys((size_t)recvN);
#pragma omp parallel for
  for (dlong i = 0; i < recvN; ++i) keys[i] = Key{habs(recvL[i]), i};
  sort_keys_by_label(keys);
  phase.lap("sort received labels");
  auto rank_of = [&](dlong idx) {  // source rank of receive slot idx
    return (int)(std::upper_bound(recvOffsets.begin(), recvOffsets.end(), (int)idx) - recvOffsets.begin()) - 1;
  };

  // positive-copy pick of the reference (libs/ogs/ogsSetup.cpp:355-382): group n of this directory rank, in ascending
  // label order, takes member  rands[n] % (group size)  with rands from the rank-seeded LCG
  std::vector<std::pair<hlong, dlong>> refOrder;  // pick mode 1: (|label|, receive slot) as std::sort leaves them
  if (unique && pick_mode() == 1) {
    refOrder.resize((size_t)recvN);
    for (dlong i = 0; i < recvN; ++i) refOrder[i] = {habs(recvL[i]), i};
    std::sort(refOrder.begin(), refOrder.end(),
              [](const std::pair<hlong, dlong>& a, const std::pair<hlong, dlong>& b) { return a.first < b.first; });
  }

  // label groups = runs of equal |label| in the sorted keys
  auto opens_group = [&](dlong i) { return i == 0 || keys[i].label != keys[i - 1].label; };
  const std::vector<dlong> groupOffsets = count_selected(recvN, chunks, opens_group);
  const dlong Ngroups = groupOffsets[chunks];
  hvec<dlong> groupStart((size_t)Ngroups + 1);
  emit_selected(recvN, chunks, groupOffsets, opens_group, [&](dlong i, dlong g) { groupStart[g] = i; });
  groupStart[Ngroups] = recvN;

  // what travels back to the sources: one byte per copy, bit 0 = label shared between ranks, bit 1 = positive copy
  // (the reference returns the signed label and a flag, 12 bytes per copy: libs/ogs/ogsSetup.cpp:384-425)
  hvec<unsigned char> recvAnswer((size_t)recvN);
  hvec<char> groupShared((size_t)Ngroups);  // copies on more than one rank
  int is_unique = 1;
  // every group is independent once its draw is known: thread chunks of consecutive groups, the generator skipped
  // ahead to the chunk's first group (draw n belongs to group n, libs/ogs/ogsSetup.cpp:355-382)
#pragma omp parallel for schedule(static, 1) reduction(min : is_unique)
  for (int c = 0; c < chunks; ++c) {
    const dlong g0 = chunk_lo(Ngroups, c, chunks), g1 = chunk_lo(Ngroups, c + 1, chunks);
    RefLcg lcg((unsigned long long)rank);
    if (unique) lcg.skip((unsigned long long)g0);
    for (dlong g = g0; g < g1; ++g) {
      const dlong s0 = groupStart[g], e = groupStart[g + 1];
      // members are in ascending receive slot = ascending source rank: first and last differ iff the label is shared
      const unsigned char shared = (size > 1 && rank_of(keys[s0].idx) != rank_of(keys[e - 1].idx)) ? 1 : 0;
      groupShared[g] = (char)shared;
      if (unique) {
        // one positive copy per label (libs/ogs/ogsSetup.cpp:364-382)
        const int m = lcg.next() % (int)(e - s0);
        for (dlong i = s0; i < e; ++i) recvAnswer[keys[i].idx] = shared;
        const dlong picked = refOrder.empty() ? keys[s0 + m].idx : refOrder[(size_t)(s0 + m)].second;
        recvAnswer[picked] = shared | 2;
      } else {
        int positives = 0;
        for (dlong i = s0; i < e; ++i) {
          const int pos = (recvL[keys[i].idx] > 0) ? 1 : 0;
          positives += pos;
          recvAnswer[keys[i].idx] = (unsigned char)(shared | (pos << 1));
        }
        if (positives != 1) is_unique = 0;
      }
    }
  }

  // replies for the shared labels only (the surface of the partition), in ascending label order
  std::vector<std::vector<hlong>> reply(size);  // per source rank: [label, nranks, (rank*2+positive)...] per shared label
  hlong NsharedLabels = 0;
  {
    auto is_shared = [&](dlong g) { return groupShared[g] != 0; };
    const std::vector<dlong> sharedOffsets = count_selected(Ngroups, chunks, is_shared);
    NsharedLabels = sharedOffsets[chunks];
    hvec<dlong> sharedGroups((size_t)NsharedLabels);
    emit_selected(Ngroups, chunks, sharedOffsets, is_shared, [&](dlong g, dlong q) { sharedGroups[q] = g; });
    std::vector<hlong> sharers;
    for (dlong q = 0; q < (dlong)NsharedLabels; ++q) {
      const dlong g = sharedGroups[q];
      const dlong s0 = groupStart[g], e = groupStart[g + 1];
      const hlong label = keys[s0].label;
      sharers.clear();
      for (dlong i = s0; i < e;) {
        const int r = rank_of(keys[i].idx);
        int pos = 0;
        dlong j = i;
        for (; j < e && rank_of(keys[j].idx) == r; ++j) pos |= (recvAnswer[keys[j].idx] >> 1) & 1;
        sharers.push_back((hlong)r * 2 + pos);
        i = j;
      }
      for (hlong sp : sharers) {
        std::vector<hlong>& out = reply[(int)(sp >> 1)];
        out.push_back(label);
        out.push_back((hlong)sharers.size());
        out.insert(out.end(), sharers.begin(), sharers.end());
      }
    }
  }
  // one rank: receive slot i is entry i and nothing is shared, so these keys are the local operator's (see build_operator)
  hvec<Key> localSorted;
  if (size == 1) localSorted.swap(keys);
  keys.clear();
  keys.shrink_to_fit();
  groupStart.clear();
  groupStart.shrink_to_fit();

  phase.lap("groups: shared / pick");
  comm.Allreduce(is_unique, Comm::Min);
  gather_defined = (is_unique == 1);
  // libs/ogs/ogsSetup.cpp:405-408
  SPN_ABORT("Halo exchange not well-defined. Some baseId groups have no unique positive id.", kind == Halo && !gather_defined);
  if (verbose) {
    hlong total = NsharedLabels;
    comm.Allreduce(total);
    if (rank == 0) printf("ogs Setup: %lld unique labels shared.\n", total);
  }

  // answers back to the sources
  recvLabels.clear();
  recvLabels.shrink_to_fit();
  hvec<unsigned char> slotAnswer;
  if (single) {
    slotAnswer.swap(recvAnswer);
  } else {
    slotAnswer.resize((size_t)Nids);
    comm.Alltoallv(recvAnswer.data(), recvCounts.data(), recvOffsets.data(), slotAnswer.data(), sendCounts.data(), sendOffsets.data());
    recvAnswer.clear();
    recvAnswer.shrink_to_fit();
  }
  if (unique) {
#pragma omp parallel for
    for (dlong n = 0; n < Nids; ++n) {
      const hlong label = habs(baseIds[n]);
      baseIds[n] = (slotAnswer[slot_of(n)] & 2) ? label : -label;
      ids[colIds[n]] = baseIds[n];  // write the signs back (libs/ogs/ogsSetup.cpp:147-150)
    }
  }

  // sharer lists of the shared labels this rank holds
  std::vector<hlong> sharerStream;
  {
    std::vector<int> sc(size), so(size + 1, 0), rc(size), ro(size + 1, 0);
    std::vector<hlong> flat;
    for (int r = 0; r < size; ++r) {
      sc[r] = (int)reply[r].size();
      so[r + 1] = so[r] + sc[r];
      flat.insert(flat.end(), reply[r].begin(), reply[r].end());
    }
    comm.Alltoall(sc.data(), rc.data());
    for (int r = 0; r < size; ++r) ro[r + 1] = ro[r] + rc[r];
    sharerStream.resize((size_t)ro[size]);
    comm.Alltoallv(flat.data(), sc.data(), so.data(), sharerStream.data(), rc.data(), ro.data());
  }

  phase.lap("answers back to the sources");
  // ---- 2. split into local / halo entries and number the rows ---------------------------------
  hvec<hlong> localIds, haloIds;
  hvec<dlong> localCols, haloCols;
  if (single) {  // nothing is shared: every entry is local, in place
    localIds.swap(baseIds);
    localCols.swap(colIds);
  } else {
    auto is_halo = [&](dlong n) { return (slotAnswer[slot_of(n)] & 1) != 0; };
    auto is_local = [&](dlong n) { return (slotAnswer[slot_of(n)] & 1) == 0; };
    const std::vector<dlong> haloOffsets = count_selected(Nids, chunks, is_halo);
    std::vector<dlong> localOffsets((size_t)chunks + 1);
    for (int c = 0; c <= chunks; ++c) localOffsets[c] = chunk_lo(Nids, c, chunks) - haloOffsets[c];
    haloIds.resize((size_t)haloOffsets[chunks]);
    haloCols.resize((size_t)haloOffsets[chunks]);
    localIds.resize((size_t)localOffsets[chunks]);
    localCols.resize((size_t)localOffsets[chunks]);
    emit_selected(Nids, chunks, haloOffsets, is_halo, [&](dlong n, dlong q) { haloIds[q] = baseIds[n]; haloCols[q] = colIds[n]; });
    emit_selected(Nids, chunks, localOffsets, is_local, [&](dlong n, dlong q) { localIds[q] = baseIds[n]; localCols[q] = colIds[n]; });
  }
  sendPos.clear(); sendPos.shrink_to_fit();
  slotAnswer.clear(); slotAnswer.shrink_to_fit();
  baseIds.clear(); baseIds.shrink_to_fit();
  colIds.clear(); colIds.shrink_to_fit();
  phase.lap("split local / halo");
  gatherLocal = build_operator(kind, N, localIds, localCols, size == 1 ? &localSorted : nullptr);
  phase.lap("local operator");
  gatherHalo = build_operator(kind, N, haloIds, haloCols);
  NlocalT = gatherLocal->NrowsT;
  NlocalP = gatherLocal->NrowsN;
  NhaloT = gatherHalo->NrowsT;
  NhaloP = gatherHalo->NrowsN;
  Ngather = NlocalP + NhaloP;
  NgatherGlobal = Ngather;
  comm.Allreduce(NgatherGlobal);

  phase.lap("halo operator");
  // ---- 3. exchange plan -------------------------------------------------------------------------
  // halo row of a label: rowLabel sorted once
  std::vector<std::pair<hlong, dlong>> labelRow((size_t)NhaloT);
  for (dlong h = 0; h < NhaloT; ++h) labelRow[h] = {gatherHalo->rowLabel[h], h};
  std::sort(labelRow.begin(), labelRow.end());
  auto row_of_label = [&](hlong label) {
    auto it = std::lower_bound(labelRow.begin(), labelRow.end(), std::make_pair(label, (dlong)-1));
    SPN_ABORT("ogs setup: shared label without a halo row", it == labelRow.end() || it->first != label);
    return it->second;
  };

  struct Slot { int rank; hlong label; dlong row; };
  std::vector<Slot> sends[3], recvs[3];
  for (size_t p = 0; p < sharerStream.size();) {
    const hlong label = sharerStream[p];
    const int nr = (int)sharerStream[p + 1];
    const dlong h = row_of_label(label);
    const bool mePos = (h < NhaloP);
    for (int q = 0; q < nr; ++q) {
      const hlong sp = sharerStream[p + 2 + q];
      const int r = (int)(sp >> 1);
      const bool rPos = (sp & 1) != 0;
      if (r == rank) {
        SPN_ABORT("ogs setup: inconsistent positive flag", rPos != mePos);
        continue;
      }
      const Slot s{r, label, h};
      sends[Sym].push_back(s);
      recvs[Sym].push_back(s);
      if (mePos) sends[NoTrans].push_back(s);  // only positive copies send (ogsPairwise.cpp:315-345)
      if (rPos) recvs[NoTrans].push_back(s);
      if (rPos) sends[Trans].push_back(s);     // only towards the positive copy (ogsPairwise.cpp:347-377)
      if (mePos) recvs[Trans].push_back(s);
    }
    p += 2 + (size_t)nr;
  }
  const dlong rowsP[3] = {NhaloT, NhaloP, NhaloP};  // Sym, NoTrans, Trans: rows contributing their own partial
  const dlong rowsOut[3] = {NhaloT, NhaloT, NhaloP};
  auto slot_less = [](const Slot& a, const Slot& b) { return a.rank < b.rank || (a.rank == b.rank && a.label < b.label); };
  for (int m = 0; m < 3; ++m) {
    ogsExchangeData_t& d = data[m];
    const int src = (kind == Unsigned) ? (int)Sym : m;  // Unsigned: one plan for all modes
    std::vector<Slot>& S = sends[src];
    std::vector<Slot>& R = recvs[src];
    if (m == src) {
      std::sort(S.begin(), S.end(), slot_less);
      std::sort(R.begin(), R.end(), slot_less);
    }
    d.Nsend = (dlong)S.size();
    d.Nrecv = (dlong)R.size();
    d.NrowsP = (kind == Unsigned) ? NhaloT : rowsP[m];
    d.Nrows = (kind == Unsigned) ? NhaloT : rowsOut[m];
    d.sendIds.resize(S.size());
    d.recvRows.resize(R.size());
    for (size_t i = 0; i < S.size(); ++i) {
      d.sendIds[i] = S[i].row;
      if (i == 0 || S[i].rank != S[i - 1].rank) {
        d.sendRanks.push_back(S[i].rank);
        d.sendOffsets.push_back((dlong)i);
        d.sendCounts.push_back(0);
      }
      d.sendCounts.back()++;
    }
    for (size_t i = 0; i < R.size(); ++i) {
      d.recvRows[i] = R[i].row;
      if (i == 0 || R[i].rank != R[i - 1].rank) {
        d.recvRanks.push_back(R[i].rank);
        d.recvOffsets.push_back((dlong)i);
        d.recvCounts.push_back(0);
      }
      d.recvCounts.back()++;
    }
    // receive slots of every output row, ascending slot == ascending source rank
    d.cmbStarts.assign((size_t)d.Nrows + 1, 0);
    for (dlong i = 0; i < d.Nrecv; ++i) {
      SPN_ABORT("ogs setup: receive slot outside the output rows", d.recvRows[i] >= d.Nrows);
      d.cmbStarts[d.recvRows[i] + 1]++;
    }
    for (dlong r = 0; r < d.Nrows; ++r) d.cmbStarts[r + 1] += d.cmbStarts[r];
    d.cmbIds.resize((size_t)d.Nrecv);
    std::vector<dlong> fill(d.cmbStarts.begin(), d.cmbStarts.end() - 1);
    for (dlong i = 0; i < d.Nrecv; ++i) d.cmbIds[fill[d.recvRows[i]]++] = i;
    // send slots of every halo row, for the peer-memory pack kernel (one fold per row, fanned out)
    SPN_ABORT("ogs setup: more send slots per neighbour than the peer exchange encodes", d.Nsend >= (1 << 24));
    d.slotDst.resize((size_t)d.Nsend);
    for (size_t p = 0; p < d.sendRanks.size(); ++p)
      for (dlong q = 0; q < d.sendCounts[p]; ++q) d.slotDst[d.sendOffsets[p] + q] = (dlong)((p << 24) | (size_t)q);
    d.rowSendStarts.assign((size_t)NhaloT + 1, 0);
    for (dlong i = 0; i < d.Nsend; ++i) d.rowSendStarts[d.sendIds[i] + 1]++;
    for (dlong r = 0; r < NhaloT; ++r) d.rowSendStarts[r + 1] += d.rowSendStarts[r];
    d.rowSendDst.resize((size_t)d.Nsend);
    std::vector<dlong> fill2(d.rowSendStarts.begin(), d.rowSendStarts.end() - 1);
    for (dlong i = 0; i < d.Nsend; ++i) d.rowSendDst[fill2[d.sendIds[i]]++] = d.slotDst[i];
  }
  phase.lap("exchange plan");
}

}  // namespace ogs
}  // namespace spn
