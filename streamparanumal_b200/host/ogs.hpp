// spn::ogs — host layer of the gather/scatter library, mirroring the reference's public ogs API
// (include/ogs.hpp:178-341: enums, ogs_t::Setup, Gather / Scatter / GatherScatter and their
// Start/Finish halves on deviceMemory<T>) over the C ABI of include/spn_b200.h.
//
// Same contract as the reference:
//   * ids[n]==0: node n does not participate; ids[n]<0: "flagged" copy (Signed kind); with
//     unique=true Setup picks ONE positive copy per label and writes the signs back into ids
//     (libs/ogs/ogsSetup.cpp:120-150,355-382).
//   * rows (gathered numbering): labels whose copies all live on this rank are "local" rows, labels
//     shared with other ranks are "halo" rows; inside each set rows are numbered by first appearance,
//     rows with a positive copy first (ogsSetup.cpp:492-559); columns ascend inside a row (:561-579).
//   * a gathered vector is laid out [NlocalT local rows | NhaloP owned halo rows]
//     (libs/ogs/ogs.cpp:253,276,295,490), Ngather = NlocalP + NhaloP (ogsSetup.cpp:191).
//   * GatherScatter(Sym): every sharer sends its CSR-ordered partial to every other sharer; each
//     rank folds own partial first, then the remote partials in ascending source rank
//     (libs/ogs/ogsPairwise.cpp:214-232).  NoTrans: only positive copies send.  Trans: partials go
//     to the positive copy only (ogsPairwise.cpp:292-377).
// What is different (B200-native, by design):
//   * the exchange is NCCL send/recv on a dedicated stream ordered by events — no device.finish()
//     host syncs (the reference has 2-3 per call: ogs.cpp:74,104,117) and no host staging;
//   * halo gather+pack and combine+scatter are one launch each (csrc/ogs_halo.cu);
//   * the interior operator runs on the warp-tile plan (csrc/ogs_tile.cuh) when gather CSR == scatter CSR;
//   * only the Pairwise exchange exists (Method is accepted and ignored): on NVSwitch every peer is
//     one hop away, which is the case Pairwise is optimal for;
//   * the positive copy of a label is picked by the reference's rule (rank-seeded LCG of libs/primitives/random.cpp,
//     member  rands[group] % size  of the label group on the directory rank, libs/ogs/ogsSetup.cpp:355-382); the
//     member ORDER inside a group is implementation-defined in the reference (an unstable sort) — see
//     SetPositivePickOrder.
#pragma once
#include <memory>
#include <vector>

#include "platform.hpp"

namespace spn {
namespace ogs {

/* enums: numeric values equal the reference's (include/ogs.hpp:178-200) and spn_type / spn_op */
typedef enum { Float, Double, Int32, Int64 } Type;
typedef enum { Add, Mul, Max, Min } Op;
typedef enum { Sym, NoTrans, Trans } Transpose;
typedef enum { Auto, Pairwise, CrystalRouter, AllToAll } Method;
typedef enum { Unsigned, Signed, Halo } Kind;

constexpr Type Dfloat = Double;
constexpr Type Dlong = Int32;
constexpr Type Hlong = Int64;

template <typename T> struct ogsType;
template <> struct ogsType<float> { static constexpr Type get() { return Float; } };
template <> struct ogsType<double> { static constexpr Type get() { return Double; } };
template <> struct ogsType<int> { static constexpr Type get() { return Int32; } };
template <> struct ogsType<long long int> { static constexpr Type get() { return Int64; } };
size_t Sizeof(Type type);

// order of the copies of one label on its directory rank, which the positive pick indexes:
// 0 (default) stable (source rank, position); 1 the order the serial reference build's std::sort leaves them in
// (exact replica, one more serial sort at setup).  Also SPN_OGS_PICK=stdsort.
void SetPositivePickOrder(int mode);

// CSR gather operator with its "T" (all copies) and "N" (positive copies only) variants
// (include/ogs/ogsOperator.hpp:44-138).  Host arrays are always built; upload() makes the device copies.
class ogsOperator_t {
 public:
  Kind kind = Unsigned;
  dlong Ncols = 0, NrowsN = 0, NrowsT = 0, nnzN = 0, nnzT = 0;
  hvec<dlong> rowStartsT, colIdsT;  // NrowsT+1, nnzT
  hvec<dlong> rowStartsN, colIdsN;  // NrowsT+1 (rows >= NrowsN are empty), nnzN
  hvec<hlong> rowLabel;             // |label| of every row

  // device side
  deviceMemory<dlong> o_rowStartsT, o_colIdsT, o_rowStartsN, o_colIdsN, o_blockRowStartsT, o_blockRowStartsN;
  dlong NrowBlocksT = 0, NrowBlocksN = 0;  // launch partitions: T rows over rowStartsT, rows [0,NrowsN) over rowStartsN
  spn_ogs_plan_t planT = nullptr, planN = nullptr;  // warp-tile plans (interior operator only)

  ~ogsOperator_t();
  void upload(platform_t& platform, bool tile_plans);

  // device operations, semantics of libs/ogs/ogsOperator.cpp:176-204, 295-322, 442-481
  void Gather(void* o_gv, const void* o_v, int k, Type type, Op op, Transpose trans, spn_stream_t stream);
  void Scatter(void* o_v, const void* o_gv, int k, Type type, Transpose trans, spn_stream_t stream);
  void GatherScatter(void* o_v, int k, Type type, Op op, Transpose trans, spn_stream_t stream);
};

// one exchange mode of the pairwise plan (reference: ogsPairwise_t::data_t, include/ogs/ogsExchange.hpp:168-195)
struct ogsExchangeData_t {
  dlong Nsend = 0, Nrecv = 0;
  dlong NrowsP = 0;  // rows whose own partial takes part in the fold
  dlong Nrows = 0;   // rows produced by the fold
  std::vector<dlong> sendIds;                            // halo row of every send slot, (dest rank, label) order
  std::vector<dlong> recvRows;                           // halo row of every receive slot, (source rank, label) order
  std::vector<int> sendRanks, recvRanks;                 // neighbours, ascending
  std::vector<dlong> sendCounts, sendOffsets, recvCounts, recvOffsets;  // per neighbour, in slots
  std::vector<dlong> cmbStarts, cmbIds;                  // Nrows+1 / Nrecv: receive slots of each row, ascending source rank
  deviceMemory<dlong> o_sendIds, o_cmbStarts, o_cmbIds;
  // peer-memory exchange: send slots of every halo row (CSR over rows) and of every send slot, each entry
  // (index of the destination in sendRanks) << 24 | slot within that destination's message
  std::vector<dlong> rowSendStarts, rowSendDst, slotDst;
  deviceMemory<dlong> o_rowSendStarts, o_rowSendDst, o_slotDst;
};

class ogs_t {
 public:
  ogs_t() = default;
  ~ogs_t() { Free(); }
  ogs_t(const ogs_t&) = delete;
  ogs_t& operator=(const ogs_t&) = delete;

  platform_t platform;
  comm_t comm;

  dlong N = 0;
  dlong Ngather = 0;        // owned gathered rows: NlocalP + NhaloP
  hlong NgatherGlobal = 0;  // summed over ranks
  dlong NlocalT = 0, NlocalP = 0, NhaloT = 0, NhaloP = 0;
  Kind kind = Unsigned;
  bool unique = false;
  bool gather_defined = false;
  bool on_device = false;
  bool halo_only = false;  // Kind Halo: only the rows shared between ranks act (no interior operator)
  bool peer_exchange() const { return p2p.enabled; }  // true: halo data moves by P2P stores over NVLink, false: NCCL

  std::shared_ptr<ogsOperator_t> gatherLocal, gatherHalo;
  ogsExchangeData_t data[3];  // indexed by Transpose

  // Setup on a platform (device arrays, NCCL communicator when comm.size()>1).  ids is updated in
  // place when unique (reference: ogs_t::Setup, libs/ogs/ogs... ogsSetup.cpp:87-290).
  void Setup(const dlong N, hlong* ids, comm_t comm, const Kind kind, const Method method, const bool unique,
             const bool verbose, platform_t& platform);
  // host-only plan (no CUDA calls): what the CPU tests of the partition/exchange logic use
  void SetupHost(const dlong N, hlong* ids, comm_t comm, const Kind kind, const bool unique, const bool verbose);
  // GlobalToLocal[n] = position of node n's row in a gathered vector ([NlocalT local rows | halo rows]), -1 for a masked
  // node; N entries (reference: ogs_t::SetupGlobalToLocalMapping, libs/ogs/ogsSetup.cpp:833-849).  A host index map.
  void SetupGlobalToLocalMapping(dlong* GlobalToLocal) const;
  void Free();

  // ---- device operations: raw-pointer core (type-erased), used by the templates and the C API ----
  // Start + Finish.  On a multi-rank operator with the peer-memory exchange the whole call (fork, halo pack, interior
  // operator, halo combine, join: 3 launches + 4 event operations) is captured ONCE per (vector, k, type, op, trans,
  // stream) into a CUDA graph and replayed afterwards — possible because a call passes no per-call value from the host
  // (the epoch lives in device memory).  SPN_OGS_GRAPH=0 disables it.
  void GatherScatter(void* o_v, int k, Type type, Op op, Transpose trans);
  void GatherScatterStart(void* o_v, int k, Type type, Op op, Transpose trans);
  void GatherScatterFinish(void* o_v, int k, Type type, Op op, Transpose trans);
  void GatherStart(void* o_gv, const void* o_v, int k, Type type, Op op, Transpose trans);
  void GatherFinish(void* o_gv, const void* o_v, int k, Type type, Op op, Transpose trans);
  void ScatterStart(void* o_v, const void* o_gv, int k, Type type, Transpose trans);
  void ScatterFinish(void* o_v, const void* o_gv, int k, Type type, Transpose trans);

  // ---- host-buffer operation (the reference's host-memory overloads, include/ogs.hpp:251-267 on memory<T>, run
  //      on the CPU; here the data is staged through the GPU).  Start enqueues  H2D -> GatherScatter -> D2H  of
  //      h_v (pinned host memory, platform.hostMalloc) on dedicated copy streams and returns at once; Finish
  //      waits until h_v holds the result.  Two operations may be in flight (double-buffered device staging), so
  //      the upload of the next vector overlaps the download of the previous one (PCIe is full duplex). ----
  void HostGatherScatterStart(void* h_v, int k, Type type, Op op, Transpose trans);
  void HostGatherScatterFinish(void* h_v);
  template <typename T> void GatherScatterStart(pinnedMemory<T> h_v, const int k, const Op op, const Transpose trans) {
    HostGatherScatterStart(h_v.ptr(), k, ogsType<T>::get(), op, trans);
  }
  template <typename T> void GatherScatterFinish(pinnedMemory<T> h_v, const int, const Op, const Transpose) {
    HostGatherScatterFinish(h_v.ptr());
  }
  template <typename T> void GatherScatter(pinnedMemory<T> h_v, const int k, const Op op, const Transpose trans) {
    HostGatherScatterStart(h_v.ptr(), k, ogsType<T>::get(), op, trans);
    HostGatherScatterFinish(h_v.ptr());
  }

  // ---- the reference's signatures (include/ogs.hpp:251-341) ----
  template <typename T> void GatherScatter(deviceMemory<T> o_v, const int k, const Op op, const Transpose trans) {
    GatherScatter(o_v.ptr(), k, ogsType<T>::get(), op, trans);
  }
  template <typename T> void GatherScatterStart(deviceMemory<T> o_v, const int k, const Op op, const Transpose trans) {
    GatherScatterStart(o_v.ptr(), k, ogsType<T>::get(), op, trans);
  }
  template <typename T> void GatherScatterFinish(deviceMemory<T> o_v, const int k, const Op op, const Transpose trans) {
    GatherScatterFinish(o_v.ptr(), k, ogsType<T>::get(), op, trans);
  }
  template <typename T> void Gather(deviceMemory<T> o_gv, deviceMemory<T> o_v, const int k, const Op op, const Transpose trans) {
    GatherStart(o_gv, o_v, k, op, trans);
    GatherFinish(o_gv, o_v, k, op, trans);
  }
  template <typename T> void GatherStart(deviceMemory<T> o_gv, deviceMemory<T> o_v, const int k, const Op op, const Transpose trans) {
    GatherStart(o_gv.ptr(), o_v.ptr(), k, ogsType<T>::get(), op, trans);
  }
  template <typename T> void GatherFinish(deviceMemory<T> o_gv, deviceMemory<T> o_v, const int k, const Op op, const Transpose trans) {
    GatherFinish(o_gv.ptr(), o_v.ptr(), k, ogsType<T>::get(), op, trans);
  }
  template <typename T> void Scatter(deviceMemory<T> o_v, deviceMemory<T> o_gv, const int k, const Transpose trans) {
    ScatterStart(o_v, o_gv, k, trans);
    ScatterFinish(o_v, o_gv, k, trans);
  }
  template <typename T> void ScatterStart(deviceMemory<T> o_v, deviceMemory<T> o_gv, const int k, const Transpose trans) {
    ScatterStart(o_v.ptr(), o_gv.ptr(), k, ogsType<T>::get(), trans);
  }
  template <typename T> void ScatterFinish(deviceMemory<T> o_v, deviceMemory<T> o_gv, const int k, const Transpose trans) {
    ScatterFinish(o_v.ptr(), o_gv.ptr(), k, ogsType<T>::get(), trans);
  }

 private:
  void SetupPlan(const dlong N, hlong* ids, comm_t comm, const Kind kind, const bool unique, const bool verbose);
  void AssertGatherDefined() const;
  void AllocBuffer(size_t Nbytes);  // reference: ogsPairwise_t::AllocBuffer, libs/ogs/ogsPairwise.cpp:389-402
  void StartExchange(const ogsExchangeData_t& d, int k, Type type);   // pack done on the current stream -> NCCL on commStream
  void FinishExchange();                                              // current stream waits for the receive

  // halo rows and interior rows own disjoint columns, so the halo chain (pack -> exchange -> combine) runs on its own
  // high-priority stream next to the interior operator: ForkHalo makes it wait for everything queued on the current
  // stream so far, JoinHalo makes the current stream wait for it.
  spn_stream_t ForkHalo();
  void JoinHalo();
  spn_stream_t haloStream = nullptr;
  spn_event_t forkEvent = nullptr, joinEvent = nullptr;
  bool halo_forked = false;

  deviceMemory<char> o_ownBuf, o_sendBuf, o_recvBuf;
  spn_stream_t commStream = nullptr;
  spn_event_t packedEvent = nullptr, recvEvent = nullptr;
  bool exchange_in_flight = false;

  // peer-memory exchange (include/spn_b200.h "Peer-memory halo exchange"; csrc/ogs_halo.cu): every rank owns one
  // IPC-shared block  [flag word per source rank | ack word per destination rank | receive half 0 | receive half 1].
  // The pack kernel of a call stores its send slots into half (epoch & 1) of each destination and sets flag[me] = epoch
  // there; the combine kernel waits for flag[src] >= epoch, reads, and sets ack[me] = epoch in each source.  A pack kernel
  // does not store into half h of a destination before that destination acknowledged this rank's previous send into h —
  // the flow control that also covers the one-way neighbour relations of the Trans / NoTrans modes.  The epoch and the
  // per-destination send history live in device memory (spn_p2p_state_t), the address tables of each mode too
  // (spn_p2p_plan_t), so a call passes no per-call value from the host.  Used when an entry (K*sizeof(T)) fits
  // kPeerEntryBytes; otherwise, or when IPC is not available or SPN_OGS_EXCHANGE=nccl, the NCCL send/recv path runs.
  static constexpr size_t kPeerEntryBytes = 32;  // receive slots are sized for K*sizeof(T) up to 4 doubles
  struct PeerExchange {
    bool enabled = false;
    deviceMemory<char> o_block;
    size_t flagBytes = 0, halfBytes = 0;
    std::vector<void*> peerBlock;  // [size]: mapped block of every neighbour, nullptr otherwise
    deviceMemory<char> o_plan;     // spn_p2p_plan_t[3], indexed by Transpose
    deviceMemory<char> o_state;    // spn_p2p_state_t
  };
  PeerExchange p2p;
  void SetupPeerExchange();
  void FreePeerExchange();
  bool UsePeerExchange(size_t Nbytes) const { return p2p.enabled && Nbytes <= kPeerEntryBytes; }
  const spn_p2p_plan_t* PeerPlan(int mode) const { return reinterpret_cast<const spn_p2p_plan_t*>(p2p.o_plan.ptr()) + mode; }
  spn_p2p_state_t* PeerState() const { return reinterpret_cast<spn_p2p_state_t*>(const_cast<char*>(p2p.o_state.ptr())); }
  bool peer_exchange_in_flight = false;

  // captured calls (see GatherScatter): second call with the same key captures, later calls replay
  struct CallGraph {
    void* o_v = nullptr;
    spn_stream_t stream = nullptr;
    int k = 0, type = 0, op = 0, trans = 0, seen = 0;
    spn_graph_t graph = nullptr;
  };
  std::vector<CallGraph> graphs;
  spn_stream_t graphStream = nullptr;  // recording stream (owned by the platform)
  int use_graphs = -1;  // -1: not decided yet
  void FreeGraphs();

  // staging of the host-buffer operations
  struct HostStage {
    deviceMemory<char> o_buf;
    spn_event_t h2dDone = nullptr, computeDone = nullptr, d2hDone = nullptr;
    void* h_v = nullptr;
    bool busy = false;
  };
  HostStage stage[2];
  int nextStage = 0;
  spn_stream_t h2dStream = nullptr, d2hStream = nullptr;
};

// halo_t (include/ogs.hpp:347-397): element-halo exchange.  ids > 0 marks the one owned copy of a node, ids < 0 its
// halo copies on other ranks (or this one); Exchange sends the owner's value to every copy, Combine adds every copy
// into the owner.  Device-buffer versions only (the host-memory overloads of the reference run on the CPU).
class halo_t {
 public:
  dlong Nhalo = 0;  // NhaloT - NhaloP: rows whose owner is another rank
  ogs_t ogs;        // the Kind-Halo operator underneath

  void Setup(const dlong N, hlong* ids, comm_t comm, const Method method, const bool verbose, platform_t& platform);
  void SetupHost(const dlong N, hlong* ids, comm_t comm, const bool verbose);  // host-only plan, for CPU tests

  void ExchangeStart(void* o_v, int k, Type type);
  void ExchangeFinish(void* o_v, int k, Type type);
  void CombineStart(void* o_v, int k, Type type);
  void CombineFinish(void* o_v, int k, Type type);

  template <typename T> void Exchange(deviceMemory<T> o_v, const int k) { ExchangeStart(o_v, k); ExchangeFinish(o_v, k); }
  template <typename T> void ExchangeStart(deviceMemory<T> o_v, const int k) { ExchangeStart(o_v.ptr(), k, ogsType<T>::get()); }
  template <typename T> void ExchangeFinish(deviceMemory<T> o_v, const int k) { ExchangeFinish(o_v.ptr(), k, ogsType<T>::get()); }
  template <typename T> void Combine(deviceMemory<T> o_v, const int k) { CombineStart(o_v, k); CombineFinish(o_v, k); }
  template <typename T> void CombineStart(deviceMemory<T> o_v, const int k) { CombineStart(o_v.ptr(), k, ogsType<T>::get()); }
  template <typename T> void CombineFinish(deviceMemory<T> o_v, const int k) { CombineFinish(o_v.ptr(), k, ogsType<T>::get()); }
};

}  // namespace ogs

// Box mesh of degree-p hex elements, nx*ny*nz elements per rank on the Factor3(size) process grid
// (reference: libs/mesh/meshSetupBoxHex3D.cpp:31-131, libs/core/rankDecomp.cpp:55-91,166-213).
// ids[e*Np+n] is the label of node n=i+j*Nq+k*Nq^2 of local element e=ex+ey*nx+ez*nx*ny: the value the reference's
// mesh_t::ConnectNodes converges to (libs/mesh/meshConnectNodes.cpp:44-81) — the minimum over the node's coincident
// copies, on any rank, of 1 + e*Np + n + (nodes owned by the lower ranks) — computed in closed form for the box.
void Factor3(const int n, int& nx, int& ny, int& nz);
void RankDecomp3(int size_x, int size_y, int size_z, int& rank_x, int& rank_y, int& rank_z, const int rank);
void BoxHexGlobalIds(int rank, int size, int nx, int ny, int nz, int p, hlong* ids, hlong* NgatherGlobalExpected);

}  // namespace spn
