// Device operations of spn::ogs — orchestration of libs/ogs/ogs.cpp:40-125 (GatherScatter),
// :206-301 (Gather), :395-492 (Scatter) and the operator dispatch of libs/ogs/ogsOperator.cpp:
// 176-204, 295-322, 442-481, re-expressed on CUDA streams and events:
//
//   Start  (current stream)  halo gather + pack ............ spn_ogs_halo_gather_pack
//          (comm stream)     waits for the pack, then ONE ncclGroup of send/recv (spn_comm_exchange)
//   Finish (current stream)  interior operator ............. tile plan / CSR kernels, overlaps the exchange
//                            waits for the receive, then
//                            fold own + received, write copies  spn_ogs_halo_combine_scatter
//
// The reference needs device.finish() host syncs at ogs.cpp:74,104,117 and ogsPairwise.cpp:118-121;
// here the host never waits.
#include <algorithm>

#include <cstdlib>

#include "ogs.hpp"

namespace spn {
namespace ogs {

// timing probe (spn_host_debug_set(0, mask)): 1 skip the halo pack / send, 2 skip the interior operator,
// 4 skip the halo combine.  Results are wrong with any bit set, so the probe only exists in builds made with
// -DSPN_TIMING_PROBES (tools/halo_breakdown.py); the shipped library compiles the checks away.
#ifdef SPN_TIMING_PROBES
int g_debug_skip = 0;
#define SPN_PROBE_SKIP(bit) (g_debug_skip & (bit))
#else
#define SPN_PROBE_SKIP(bit) false
#endif

// ---------------------------------------------------------------------------------------------
// ogsOperator_t
// ---------------------------------------------------------------------------------------------
ogsOperator_t::~ogsOperator_t() {
  if (planT) spn_ogs_plan_destroy(planT);
  if (planN && planN != planT) spn_ogs_plan_destroy(planN);
}

void ogsOperator_t::upload(platform_t& platform, bool tile_plans) {
  if (NrowsT == 0) return;
  o_rowStartsT = platform.malloc<dlong>(rowStartsT);
  o_colIdsT = platform.malloc<dlong>(colIdsT);
  if (kind == Signed) {
    o_rowStartsN = platform.malloc<dlong>(rowStartsN);
    // a rank that holds no positive copy of any row (nnzN == 0) still passes colIdsN to the kernels, whose every row
    // is then empty: keep a one-entry array so that the handle is not null (the reference's kernels run on it as they are)
    o_colIdsN = nnzN ? platform.malloc<dlong>(colIdsN) : platform.malloc<dlong>(std::vector<dlong>(1, 0));
  } else {
    o_rowStartsN = o_rowStartsT;
    o_colIdsN = o_colIdsT;
  }
  // launch partitions of the CSR kernels: whole rows, <= spn_ogs_nodes_per_block() non-zeros per block
  // (the role of blockRows(), libs/ogs/ogsOperator.cpp:522-594; a longer single row gets its own block)
  auto partition = [&](const hvec<dlong>& rs, dlong nrows, dlong& nblocks) {
    const dlong limit = spn_ogs_nodes_per_block();
    std::vector<dlong> blocks(1, 0);
    dlong r = 0;
    while (r < nrows) {
      const dlong cap = rs[r] + limit;
      dlong e = (dlong)(std::upper_bound(rs.begin() + r, rs.begin() + nrows + 1, cap) - rs.begin()) - 1;
      e = std::min(std::max(e, r + 1), nrows);
      blocks.push_back(e);
      r = e;
    }
    nblocks = (dlong)blocks.size() - 1;
    return platform.malloc<dlong>(blocks);
  };
  o_blockRowStartsT = partition(rowStartsT, NrowsT, NrowBlocksT);
  if (kind == Signed) {
    if (NrowsN) o_blockRowStartsN = partition(rowStartsN, NrowsN, NrowBlocksN);
  } else {
    o_blockRowStartsN = o_blockRowStartsT;
    NrowBlocksN = NrowBlocksT;
  }
  if (tile_plans) {
    SPN_CHECK(spn_ogs_plan_create(&planT, NrowsT, Ncols, rowStartsT.data(), colIdsT.data(), o_rowStartsT.ptr(),
                                  o_colIdsT.ptr()));
    if (kind == Signed && nnzN != nnzT) {
      if (nnzN)
        SPN_CHECK(spn_ogs_plan_create(&planN, NrowsN, Ncols, rowStartsN.data(), colIdsN.data(), o_rowStartsN.ptr(),
                                      o_colIdsN.ptr()));
    } else {
      planN = planT;  // every copy is positive: the N operator is the T operator
    }
  }
}

void ogsOperator_t::Gather(void* o_gv, const void* o_v, int k, Type type, Op op, Transpose trans, spn_stream_t stream) {
  if (NrowsT == 0) return;
  const bool useN = (trans == NoTrans);  // libs/ogs/ogsOperator.cpp:185-203
  if (useN && nnzN == 0) return;
  spn_ogs_plan_t plan = useN ? planN : planT;
  if (plan) {
    SPN_CHECK(spn_ogs_plan_gather(plan, k, o_v, o_gv, type, op, stream));
    return;
  }
  if (useN)  // rows [0, NrowsN) only, as the reference's NrowBlocksN launch
    SPN_CHECK(spn_ogs_gather(NrowBlocksN, k, o_blockRowStartsN.ptr(), o_rowStartsN.ptr(), o_colIdsN.ptr(), o_v, o_gv, type, op, stream));
  else
    SPN_CHECK(spn_ogs_gather(NrowBlocksT, k, o_blockRowStartsT.ptr(), o_rowStartsT.ptr(), o_colIdsT.ptr(), o_v, o_gv, type, op, stream));
}

void ogsOperator_t::Scatter(void* o_v, const void* o_gv, int k, Type type, Transpose trans, spn_stream_t stream) {
  if (NrowsT == 0) return;
  const bool useN = (trans == Trans);  // libs/ogs/ogsOperator.cpp:303-321
  if (useN && nnzN == 0) return;
  spn_ogs_plan_t plan = useN ? planN : planT;
  if (plan) {
    SPN_CHECK(spn_ogs_plan_scatter(plan, k, o_gv, o_v, type, stream));
    return;
  }
  if (useN)
    SPN_CHECK(spn_ogs_scatter(NrowBlocksN, k, o_blockRowStartsN.ptr(), o_rowStartsN.ptr(), o_colIdsN.ptr(), o_gv, o_v, type, stream));
  else
    SPN_CHECK(spn_ogs_scatter(NrowBlocksT, k, o_blockRowStartsT.ptr(), o_rowStartsT.ptr(), o_colIdsT.ptr(), o_gv, o_v, type, stream));
}

void ogsOperator_t::GatherScatter(void* o_v, int k, Type type, Op op, Transpose trans, spn_stream_t stream) {
  if (NrowsT == 0) return;
  // gather / scatter CSR pairs of libs/ogs/ogsOperator.cpp:450-480 (Trans: T then N; Sym: T, T; NoTrans: N then T)
  const bool same = (trans == Sym) || kind != Signed || nnzN == nnzT;
  if (same) {
    if (planT) SPN_CHECK(spn_ogs_plan_gather_scatter(planT, k, o_v, type, op, stream));
    else
      SPN_CHECK(spn_ogs_gather_scatter(NrowBlocksT, k, o_blockRowStartsT.ptr(), o_rowStartsT.ptr(), o_colIdsT.ptr(),
                                       o_rowStartsT.ptr(), o_colIdsT.ptr(), o_v, type, op, stream));
    return;
  }
  const dlong* gRS = (trans == Trans) ? o_rowStartsT.ptr() : o_rowStartsN.ptr();
  const dlong* gCI = (trans == Trans) ? o_colIdsT.ptr() : o_colIdsN.ptr();
  const dlong* sRS = (trans == Trans) ? o_rowStartsN.ptr() : o_rowStartsT.ptr();
  const dlong* sCI = (trans == Trans) ? o_colIdsN.ptr() : o_colIdsT.ptr();
  SPN_CHECK(spn_ogs_gather_scatter(NrowBlocksT, k, o_blockRowStartsT.ptr(), gRS, gCI, sRS, sCI, o_v, type, op, stream));
}

// ---------------------------------------------------------------------------------------------
// ogs_t: buffers and the exchange
// ---------------------------------------------------------------------------------------------
void ogs_t::AllocBuffer(size_t Nbytes) {
  const ogsExchangeData_t& d = data[Sym];  // the Sym plan bounds every mode (ogsPairwise.cpp:389-402)
  if (o_ownBuf.size() < (size_t)NhaloT * Nbytes || o_sendBuf.size() < (size_t)d.Nsend * Nbytes ||
      o_recvBuf.size() < (size_t)d.Nrecv * Nbytes)
    FreeGraphs();  // captured calls hold the old buffers
  if (o_ownBuf.size() < (size_t)NhaloT * Nbytes) o_ownBuf = platform.malloc<char>((size_t)NhaloT * Nbytes);
  if (o_sendBuf.size() < (size_t)d.Nsend * Nbytes) o_sendBuf = platform.malloc<char>((size_t)d.Nsend * Nbytes);
  if (o_recvBuf.size() < (size_t)d.Nrecv * Nbytes) o_recvBuf = platform.malloc<char>((size_t)d.Nrecv * Nbytes);
}

spn_stream_t ogs_t::ForkHalo() {
  if (!haloStream) return platform.getStream();  // one rank: no separate halo chain
  if (!halo_forked) {
    SPN_CHECK(spn_event_record(forkEvent, platform.getStream()));
    SPN_CHECK(spn_stream_wait_event(haloStream, forkEvent));
    halo_forked = true;
  }
  return haloStream;
}

void ogs_t::JoinHalo() {
  if (!halo_forked) return;
  SPN_CHECK(spn_event_record(joinEvent, haloStream));
  SPN_CHECK(spn_stream_wait_event(platform.getStream(), joinEvent));
  halo_forked = false;
}

void ogs_t::StartExchange(const ogsExchangeData_t& d, int k, Type type) {
  if (comm.size() == 1) return;
  const size_t Nbytes = Sizeof(type) * (size_t)k;
  spn_stream_t cur = ForkHalo();  // the pack ran there
  SPN_CHECK(spn_event_record(packedEvent, cur));
  SPN_CHECK(spn_stream_wait_event(commStream, packedEvent));
  const int nr = (int)d.recvRanks.size(), ns = (int)d.sendRanks.size();
  std::vector<size_t> ro(nr), rb(nr), so(ns), sb(ns);
  for (int i = 0; i < nr; ++i) { ro[i] = (size_t)d.recvOffsets[i] * Nbytes; rb[i] = (size_t)d.recvCounts[i] * Nbytes; }
  for (int i = 0; i < ns; ++i) { so[i] = (size_t)d.sendOffsets[i] * Nbytes; sb[i] = (size_t)d.sendCounts[i] * Nbytes; }
  SPN_CHECK(spn_comm_exchange(comm.device(), nr, d.recvRanks.data(), ro.data(), rb.data(), o_recvBuf.ptr(), ns,
                              d.sendRanks.data(), so.data(), sb.data(), o_sendBuf.ptr(), commStream));
  SPN_CHECK(spn_event_record(recvEvent, commStream));
  exchange_in_flight = true;
}

void ogs_t::FinishExchange() {
  if (!exchange_in_flight) return;
  SPN_CHECK(spn_stream_wait_event(ForkHalo(), recvEvent));  // the combine runs there
  exchange_in_flight = false;
}

// ---------------------------------------------------------------------------------------------
// GatherScatter (libs/ogs/ogs.cpp:40-125)
// ---------------------------------------------------------------------------------------------
void ogs_t::FreeGraphs() {
  for (CallGraph& g : graphs)
    if (g.graph) spn_graph_destroy(g.graph);
  graphs.clear();
}

void ogs_t::GatherScatter(void* o_v, int k, Type type, Op op, Transpose trans) {
  SPN_ABORT("ogs_t used before Setup on a platform", !on_device);
  if (use_graphs < 0) {
    const char* e = getenv("SPN_OGS_GRAPH");
    use_graphs = (e && e[0] == '0') ? 0 : 1;
  }
  const size_t Nbytes = (size_t)k * Sizeof(type);
  // replay pays off where the call is launch-bound; from ~16 M entries per rank on the directly issued sequence is faster
  // (8 ranks: 16^3 elements per rank 54 us replayed vs 63 us direct, 32^3 109 vs 92 us; 2 ranks: 64^3/2 0.274 vs 0.263 ms)
  const bool graphable = use_graphs && comm.size() > 1 && UsePeerExchange(Nbytes) && !exchange_in_flight && !peer_exchange_in_flight &&
                         (size_t)N * (size_t)k <= ((size_t)1 << 23);
  if (!graphable) {
    GatherScatterStart(o_v, k, type, op, trans);
    GatherScatterFinish(o_v, k, type, op, trans);
    return;
  }
  spn_stream_t cur = platform.getStream();
  CallGraph* hit = nullptr;
  for (CallGraph& g : graphs)
    if (g.o_v == o_v && g.stream == cur && g.k == k && g.type == (int)type && g.op == (int)op && g.trans == (int)trans) { hit = &g; break; }
  if (hit && hit->graph) {
    SPN_CHECK(spn_graph_launch(hit->graph, cur));
    return;
  }
  if (!hit) {  // first sighting: run it directly (also sizes the buffers and warms every lazy attribute)
    if (graphs.size() >= 16) FreeGraphs();
    CallGraph g;
    g.o_v = o_v; g.stream = cur; g.k = k; g.type = (int)type; g.op = (int)op; g.trans = (int)trans; g.seen = 1;
    graphs.push_back(g);
    GatherScatterStart(o_v, k, type, op, trans);
    GatherScatterFinish(o_v, k, type, op, trans);
    return;
  }
  // second sighting: record the same sequence instead of running it, then replay.  The recording runs on a stream of
  // our own (the caller's current stream may be the legacy default stream, which cannot be captured); the graph itself
  // is launched into the caller's stream.
  AllocBuffer(Nbytes);  // (no allocation inside the capture)
  if (!graphStream) graphStream = platform.createStream();
  platform.setStream(graphStream);
  spn_graph_t graph = nullptr;
  try {
    SPN_CHECK(spn_graph_begin_capture(graphStream));
    try {
      GatherScatterStart(o_v, k, type, op, trans);
      GatherScatterFinish(o_v, k, type, op, trans);
    } catch (...) {
      spn_graph_end_capture(graphStream, &graph);
      if (graph) spn_graph_destroy(graph);
      throw;
    }
    SPN_CHECK(spn_graph_end_capture(graphStream, &graph));
  } catch (...) {
    platform.setStream(cur);
    throw;
  }
  platform.setStream(cur);
  hit->graph = graph;
  SPN_CHECK(spn_graph_launch(graph, cur));
}

void ogs_t::GatherScatterStart(void* o_v, int k, Type type, Op op, Transpose trans) {
  SPN_ABORT("ogs_t used before Setup on a platform", !on_device);
  if (NhaloT == 0 && comm.size() == 1) return;
  if (SPN_PROBE_SKIP(1)) return;
  AllocBuffer((size_t)k * Sizeof(type));
  const ogsExchangeData_t& d = data[trans];
  // halo partials: gather over the T copies unless NoTrans (positive copies only), straight into send order
  const bool useN = (trans == NoTrans);
  const dlong* hRS = useN ? gatherHalo->o_rowStartsN.ptr() : gatherHalo->o_rowStartsT.ptr();
  const dlong* hCI = useN ? gatherHalo->o_colIdsN.ptr() : gatherHalo->o_colIdsT.ptr();
  if (comm.size() > 1 && UsePeerExchange((size_t)k * Sizeof(type))) {
    // partials go straight into the neighbours' receive buffers (P2P stores), flags raised by the same launch
    SPN_CHECK(spn_ogs_halo_gather_pack_p2p(NhaloT, k, hRS, hCI, d.o_rowSendStarts.ptr(), d.o_rowSendDst.ptr(), o_v, o_ownBuf.ptr(),
                                           PeerPlan(trans), PeerState(), type, op, ForkHalo()));
    peer_exchange_in_flight = true;
    return;
  }
  SPN_CHECK(spn_ogs_halo_gather_pack(NhaloT, d.Nsend, k, hRS, hCI, d.o_sendIds.ptr(), o_v, o_ownBuf.ptr(), o_sendBuf.ptr(), type,
                                     op, ForkHalo()));
  StartExchange(d, k, type);
}

void ogs_t::GatherScatterFinish(void* o_v, int k, Type type, Op op, Transpose trans) {
  SPN_ABORT("ogs_t used before Setup on a platform", !on_device);
  spn_stream_t cur = platform.getStream();
  spn_stream_t hs = (NhaloT == 0 && comm.size() == 1) ? cur : ForkHalo();  // halo chain: its own stream, next to the interior operator
  if (!SPN_PROBE_SKIP(2) && !halo_only) gatherLocal->GatherScatter(o_v, k, type, op, trans, cur);  // interior rows
  if (NhaloT == 0 && comm.size() == 1) return;
  if (SPN_PROBE_SKIP(4)) { peer_exchange_in_flight = false; FinishExchange(); JoinHalo(); return; }
  const ogsExchangeData_t& d = data[trans];
  // scatter side: every copy, except Trans which writes positive copies only (ogsOperator.cpp:303-321)
  const bool sN = (trans == Trans);
  const dlong* sRS = sN ? gatherHalo->o_rowStartsN.ptr() : gatherHalo->o_rowStartsT.ptr();
  const dlong* sCI = sN ? gatherHalo->o_colIdsN.ptr() : gatherHalo->o_colIdsT.ptr();
  if (peer_exchange_in_flight) {
    SPN_CHECK(spn_ogs_halo_combine_scatter_p2p(d.Nrows, d.NrowsP, k, d.o_cmbStarts.ptr(), d.o_cmbIds.ptr(), o_ownBuf.ptr(),
                                               PeerPlan(trans), PeerState(), sRS, sCI, o_v, type, op, hs));
    peer_exchange_in_flight = false;
    JoinHalo();
    return;
  }
  FinishExchange();
  SPN_CHECK(spn_ogs_halo_combine_scatter(d.Nrows, d.NrowsP, k, d.o_cmbStarts.ptr(), d.o_cmbIds.ptr(), o_ownBuf.ptr(),
                                         o_recvBuf.ptr(), sRS, sCI, o_v, type, op, hs));
  JoinHalo();
}

// ---------------------------------------------------------------------------------------------
// GatherScatter on a pinned host vector: H2D | device operation | D2H on three streams, two stages
// ---------------------------------------------------------------------------------------------
void ogs_t::HostGatherScatterStart(void* h_v, int k, Type type, Op op, Transpose trans) {
  SPN_ABORT("ogs_t used before Setup on a platform", !on_device);
  SPN_ABORT("HostGatherScatterStart: null host vector", h_v == nullptr && N > 0);
  HostStage& s = stage[nextStage];
  SPN_ABORT("HostGatherScatterStart: two host-buffer operations are already in flight (call Finish first)", s.busy);
  nextStage ^= 1;
  const size_t bytes = (size_t)N * (size_t)k * Sizeof(type);
  if (!h2dStream) {
    h2dStream = platform.createStream();
    d2hStream = platform.createStream();
  }
  if (!s.h2dDone) {
    SPN_CHECK(spn_event_create(&s.h2dDone));
    SPN_CHECK(spn_event_create(&s.computeDone));
    SPN_CHECK(spn_event_create(&s.d2hDone));
  }
  if (s.o_buf.size() < bytes) s.o_buf = platform.malloc<char>(bytes);
  spn_stream_t cur = platform.getStream();
  if (bytes) SPN_CHECK(spn_memcpy_h2d(s.o_buf.ptr(), h_v, bytes, h2dStream, 1));
  SPN_CHECK(spn_event_record(s.h2dDone, h2dStream));
  SPN_CHECK(spn_stream_wait_event(cur, s.h2dDone));
  GatherScatterStart(s.o_buf.ptr(), k, type, op, trans);
  GatherScatterFinish(s.o_buf.ptr(), k, type, op, trans);
  SPN_CHECK(spn_event_record(s.computeDone, cur));
  SPN_CHECK(spn_stream_wait_event(d2hStream, s.computeDone));
  if (bytes) SPN_CHECK(spn_memcpy_d2h(h_v, s.o_buf.ptr(), bytes, d2hStream, 1));
  SPN_CHECK(spn_event_record(s.d2hDone, d2hStream));
  s.h_v = h_v;
  s.busy = true;
}

void ogs_t::HostGatherScatterFinish(void* h_v) {
  for (HostStage& s : stage) {
    if (s.busy && s.h_v == h_v) {
      SPN_CHECK(spn_event_sync(s.d2hDone));
      s.busy = false;
      s.h_v = nullptr;
      return;
    }
  }
  SPN_ABORT("HostGatherScatterFinish: no operation in flight on this host vector", true);
}

// ---------------------------------------------------------------------------------------------
// halo_t (libs/ogs/ogsHalo.cpp:39-400, device-buffer versions).  In the reference
//   Exchange = gatherHalo->Gather(NoTrans) -> exchange(NoTrans) -> gatherHalo->Scatter(NoTrans)
//   Combine  = gatherHalo->Gather(Trans)   -> exchange(Trans)   -> gatherHalo->Scatter(Trans)
// i.e. the halo half of GatherScatter(Add, NoTrans / Trans) with no interior operator — which is what an ogs_t
// set up with Kind Halo runs (halo_only).
// ---------------------------------------------------------------------------------------------
void halo_t::Setup(const dlong N, hlong* ids, comm_t comm, const Method method, const bool verbose, platform_t& platform) {
  ogs.Setup(N, ids, comm, Halo, method, /*unique=*/false, verbose, platform);
  Nhalo = ogs.NhaloT - ogs.NhaloP;  // halo_t::Setup, libs/ogs/ogsSetup.cpp:81
}
void halo_t::SetupHost(const dlong N, hlong* ids, comm_t comm, const bool verbose) {
  ogs.SetupHost(N, ids, comm, Halo, /*unique=*/false, verbose);
  Nhalo = ogs.NhaloT - ogs.NhaloP;
}
void halo_t::ExchangeStart(void* o_v, int k, Type type) { ogs.GatherScatterStart(o_v, k, type, Add, NoTrans); }
void halo_t::ExchangeFinish(void* o_v, int k, Type type) { ogs.GatherScatterFinish(o_v, k, type, Add, NoTrans); }
void halo_t::CombineStart(void* o_v, int k, Type type) { ogs.GatherScatterStart(o_v, k, type, Add, Trans); }
void halo_t::CombineFinish(void* o_v, int k, Type type) { ogs.GatherScatterFinish(o_v, k, type, Add, Trans); }

// ---------------------------------------------------------------------------------------------
// Gather (libs/ogs/ogs.cpp:206-301): o_gv = [NlocalT local rows | NhaloP owned halo rows]
// ---------------------------------------------------------------------------------------------
void ogs_t::GatherStart(void* o_gv, const void* o_v, int k, Type type, Op op, Transpose trans) {
  SPN_ABORT("ogs_t used before Setup on a platform", !on_device);
  AssertGatherDefined();
  if (NhaloT == 0) return;
  const size_t Nbytes = (size_t)k * Sizeof(type);
  if (trans == Trans) {  // partials travel to the positive copy
    AllocBuffer(Nbytes);
    const ogsExchangeData_t& d = data[Trans];
    if (comm.size() > 1 && UsePeerExchange(Nbytes)) {
      SPN_CHECK(spn_ogs_halo_gather_pack_p2p(NhaloT, k, gatherHalo->o_rowStartsT.ptr(), gatherHalo->o_colIdsT.ptr(),
                                             d.o_rowSendStarts.ptr(), d.o_rowSendDst.ptr(), o_v, o_ownBuf.ptr(), PeerPlan(Trans),
                                             PeerState(), type, op, ForkHalo()));
      peer_exchange_in_flight = true;
      return;
    }
    SPN_CHECK(spn_ogs_halo_gather_pack(NhaloT, d.Nsend, k, gatherHalo->o_rowStartsT.ptr(), gatherHalo->o_colIdsT.ptr(),
                                       d.o_sendIds.ptr(), o_v, o_ownBuf.ptr(), o_sendBuf.ptr(), type, op, ForkHalo()));
    StartExchange(d, k, type);
  } else {  // no communication: gather the halo rows in place behind the local rows
    gatherHalo->Gather(static_cast<char*>(o_gv) + Nbytes * (size_t)NlocalT, o_v, k, type, op, trans, ForkHalo());
  }
}

void ogs_t::GatherFinish(void* o_gv, const void* o_v, int k, Type type, Op op, Transpose trans) {
  SPN_ABORT("ogs_t used before Setup on a platform", !on_device);
  AssertGatherDefined();
  spn_stream_t cur = platform.getStream();
  spn_stream_t hs = (NhaloT > 0) ? ForkHalo() : cur;
  gatherLocal->Gather(o_gv, o_v, k, type, op, trans, cur);
  if (trans != Trans || NhaloT == 0) JoinHalo();
  if (trans == Trans && NhaloT > 0) {
    const ogsExchangeData_t& d = data[Trans];
    const size_t Nbytes = (size_t)k * Sizeof(type);
    if (peer_exchange_in_flight) {
      SPN_CHECK(spn_ogs_halo_combine_scatter_p2p(d.Nrows, d.NrowsP, k, d.o_cmbStarts.ptr(), d.o_cmbIds.ptr(), o_ownBuf.ptr(),
                                                 PeerPlan(Trans), PeerState(), nullptr, nullptr,
                                                 static_cast<char*>(o_gv) + Nbytes * (size_t)NlocalT, type, op, hs));
      peer_exchange_in_flight = false;
      JoinHalo();
      return;
    }
    FinishExchange();
    // fold straight into the tail of o_gv (the reference folds into o_recvBuf and copies: ogs.cpp:270-277)
    SPN_CHECK(spn_ogs_halo_combine_scatter(d.Nrows, d.NrowsP, k, d.o_cmbStarts.ptr(), d.o_cmbIds.ptr(), o_ownBuf.ptr(),
                                           o_recvBuf.ptr(), nullptr, nullptr,
                                           static_cast<char*>(o_gv) + Nbytes * (size_t)NlocalT, type, op, hs));
    JoinHalo();
  }
}

// ---------------------------------------------------------------------------------------------
// Scatter (libs/ogs/ogs.cpp:395-492)
// ---------------------------------------------------------------------------------------------
void ogs_t::ScatterStart(void* /*o_v*/, const void* o_gv, int k, Type type, Transpose trans) {
  SPN_ABORT("ogs_t used before Setup on a platform", !on_device);
  AssertGatherDefined();
  if (trans != NoTrans || NhaloT == 0) return;
  const size_t Nbytes = (size_t)k * Sizeof(type);
  AllocBuffer(Nbytes);
  const ogsExchangeData_t& d = data[NoTrans];
  // owners pack their rows of o_gv straight into send order (the reference copies, then extracts: ogs.cpp:423-431)
  if (comm.size() > 1 && UsePeerExchange(Nbytes)) {
    SPN_CHECK(spn_ogs_extract_p2p(d.Nsend, k, d.o_sendIds.ptr(), d.o_slotDst.ptr(),
                                  static_cast<const char*>(o_gv) + Nbytes * (size_t)NlocalT, PeerPlan(NoTrans), PeerState(), type,
                                  ForkHalo()));
    peer_exchange_in_flight = true;
    return;
  }
  SPN_CHECK(spn_ogs_extract(d.Nsend, k, d.o_sendIds.ptr(), static_cast<const char*>(o_gv) + Nbytes * (size_t)NlocalT,
                            o_sendBuf.ptr(), type, ForkHalo()));
  StartExchange(d, k, type);
}

void ogs_t::ScatterFinish(void* o_v, const void* o_gv, int k, Type type, Transpose trans) {
  SPN_ABORT("ogs_t used before Setup on a platform", !on_device);
  AssertGatherDefined();
  spn_stream_t cur = platform.getStream();
  spn_stream_t hs = (NhaloT > 0) ? ForkHalo() : cur;
  gatherLocal->Scatter(o_v, o_gv, k, type, trans, cur);
  if (NhaloT == 0) return;
  const size_t Nbytes = (size_t)k * Sizeof(type);
  const char* gvHalo = static_cast<const char*>(o_gv) + Nbytes * (size_t)NlocalT;
  if (trans == NoTrans) {
    const ogsExchangeData_t& d = data[NoTrans];
    if (peer_exchange_in_flight) {
      SPN_CHECK(spn_ogs_halo_combine_scatter_p2p(d.Nrows, d.NrowsP, k, d.o_cmbStarts.ptr(), d.o_cmbIds.ptr(), gvHalo, PeerPlan(NoTrans),
                                                 PeerState(), gatherHalo->o_rowStartsT.ptr(), gatherHalo->o_colIdsT.ptr(), o_v, type,
                                                 Add, hs));
      peer_exchange_in_flight = false;
      JoinHalo();
      return;
    }
    FinishExchange();
    // rows < NhaloP take their own value from o_gv, the others the owner's (identity (+) x, as postmpi.Gather does)
    SPN_CHECK(spn_ogs_halo_combine_scatter(d.Nrows, d.NrowsP, k, d.o_cmbStarts.ptr(), d.o_cmbIds.ptr(), gvHalo, o_recvBuf.ptr(),
                                           gatherHalo->o_rowStartsT.ptr(), gatherHalo->o_colIdsT.ptr(), o_v, type, Add, hs));
  } else {
    gatherHalo->Scatter(o_v, gvHalo, k, type, trans, hs);
  }
  JoinHalo();
}

}  // namespace ogs
}  // namespace spn
