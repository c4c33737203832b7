// C binding of the host layer (include/spn_host.h).
#include <string>
#include <cstdlib>
#include <unistd.h>

#include "ogs.hpp"
#include "spn_host.h"

using namespace spn;
#ifdef SPN_TIMING_PROBES
namespace spn { namespace ogs { extern int g_debug_skip; } }
#endif

struct spn_host_comm { comm_t comm; };
struct spn_host_platform { platform_t platform; };
struct spn_host_ogs { ogs::ogs_t ogs; };

namespace {
thread_local std::string g_last_error;

template <typename F>
int guarded(F&& f) {
  try {
    f();
    return SPN_SUCCESS;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return SPN_ERR_HOST;
  }
}
Comm::Op to_op(int op) { return op == 1 ? Comm::Max : op == 2 ? Comm::Min : Comm::Sum; }
}  // namespace

extern "C" {

const char* spn_host_last_error(void) { return g_last_error.c_str(); }

int spn_host_comm_create(spn_host_comm_t* comm, int rank, int size, const char* dir) {
  if (!comm) return SPN_ERR_INVALID_ARGUMENT;
  return guarded([&] {
    if (dir) *comm = new spn_host_comm{comm_t(rank, size, dir)};
    else {
      const char* port = getenv("MASTER_PORT");
      *comm = new spn_host_comm{comm_t(rank, size, std::string("/tmp/spn_rdv_") + std::to_string((long)getuid()) + "_" + (port ? port : "0"))};
    }
  });
}
int spn_host_comm_create_world(spn_host_comm_t* comm) {
  if (!comm) return SPN_ERR_INVALID_ARGUMENT;
  return guarded([&] { *comm = new spn_host_comm{comm_t::World()}; });
}
int spn_host_comm_destroy(spn_host_comm_t comm) { delete comm; return SPN_SUCCESS; }
int spn_host_comm_rank(spn_host_comm_t comm, int* rank) { if (!comm || !rank) return SPN_ERR_INVALID_ARGUMENT; *rank = comm->comm.rank(); return 0; }
int spn_host_comm_size(spn_host_comm_t comm, int* size) { if (!comm || !size) return SPN_ERR_INVALID_ARGUMENT; *size = comm->comm.size(); return 0; }
int spn_host_comm_barrier(spn_host_comm_t comm) { if (!comm) return SPN_ERR_INVALID_ARGUMENT; return guarded([&] { comm->comm.Barrier(); }); }
int spn_host_comm_allreduce_i64(spn_host_comm_t comm, long long* buf, int n, int op) {
  if (!comm || !buf) return SPN_ERR_INVALID_ARGUMENT;
  return guarded([&] { comm->comm.Allreduce(buf, n, to_op(op)); });
}
int spn_host_comm_allreduce_f64(spn_host_comm_t comm, double* buf, int n, int op) {
  if (!comm || !buf) return SPN_ERR_INVALID_ARGUMENT;
  return guarded([&] { comm->comm.Allreduce(buf, n, to_op(op)); });
}
int spn_host_comm_allreduce_device_f64(spn_host_comm_t comm, double* d_buf, int n, spn_stream_t stream) {
  if (!comm) return SPN_ERR_INVALID_ARGUMENT;
  return guarded([&] { comm->comm.AllreduceDevice(d_buf, n, stream); });
}

int spn_host_platform_create(spn_host_platform_t* platform, spn_host_comm_t comm, int device_id) {
  if (!platform || !comm) return SPN_ERR_INVALID_ARGUMENT;
  return guarded([&] { *platform = new spn_host_platform{platform_t(comm->comm, device_id)}; });
}
int spn_host_platform_destroy(spn_host_platform_t platform) { delete platform; return SPN_SUCCESS; }
int spn_host_platform_set_stream(spn_host_platform_t platform, spn_stream_t stream) {
  if (!platform) return SPN_ERR_INVALID_ARGUMENT;
  platform->platform.setStream(stream);
  return SPN_SUCCESS;
}
int spn_host_platform_finish(spn_host_platform_t platform) {
  if (!platform) return SPN_ERR_INVALID_ARGUMENT;
  return guarded([&] { platform->platform.finish(); });
}

int spn_host_rank_grid(int rank, int size, int* out) {
  if (!out || size < 1 || rank < 0 || rank >= size) return SPN_ERR_INVALID_ARGUMENT;
  return guarded([&] {
    Factor3(size, out[0], out[1], out[2]);
    RankDecomp3(out[0], out[1], out[2], out[3], out[4], out[5], rank);
  });
}
int spn_host_box_hex_ids(int rank, int size, int nx, int ny, int nz, int p, long long* ids, long long* Nglobal) {
  if (!ids || size < 1 || rank < 0 || rank >= size || nx < 1 || ny < 1 || nz < 1 || p < 1) return SPN_ERR_INVALID_ARGUMENT;
  return guarded([&] { BoxHexGlobalIds(rank, size, nx, ny, nz, p, ids, Nglobal); });
}

int spn_host_ogs_setup(spn_host_ogs_t* ogs, int N, long long* ids, spn_host_comm_t comm, int kind, int unique, int verbose,
                       spn_host_platform_t platform) {
  if (!ogs || !comm || N < 0 || (N > 0 && !ids)) return SPN_ERR_INVALID_ARGUMENT;
  *ogs = nullptr;
  spn_host_ogs* o = new spn_host_ogs();
  const int rc = guarded([&] {
    if (platform) o->ogs.Setup(N, ids, comm->comm, (ogs::Kind)kind, ogs::Auto, unique != 0, verbose != 0, platform->platform);
    else o->ogs.SetupHost(N, ids, comm->comm, (ogs::Kind)kind, unique != 0, verbose != 0);
  });
  if (rc) { delete o; return rc; }
  *ogs = o;
  return SPN_SUCCESS;
}
int spn_host_ogs_free(spn_host_ogs_t ogs) { delete ogs; return SPN_SUCCESS; }

int spn_host_ogs_info(spn_host_ogs_t h, long long* info) {
  if (!h || !info) return SPN_ERR_INVALID_ARGUMENT;
  const ogs::ogs_t& o = h->ogs;
  for (int i = 0; i < 16; ++i) info[i] = 0;
  info[0] = o.N; info[1] = o.Ngather; info[2] = o.NgatherGlobal; info[3] = o.NlocalT; info[4] = o.NlocalP;
  info[5] = o.NhaloT; info[6] = o.NhaloP; info[7] = o.gather_defined ? 1 : 0;
  if (o.gatherLocal) { info[8] = o.gatherLocal->nnzT; info[9] = o.gatherLocal->nnzN; }
  if (o.gatherHalo) { info[10] = o.gatherHalo->nnzT; info[11] = o.gatherHalo->nnzN; }
  if (o.gatherLocal && o.gatherLocal->planT) {
    int nt = 0;
    spn_ogs_plan_info(o.gatherLocal->planT, &nt, nullptr, nullptr);
    info[12] = nt;
  }
  info[13] = o.peer_exchange() ? 1 : 0;
  return SPN_SUCCESS;
}
int spn_host_ogs_exchange_info(spn_host_ogs_t h, int mode, long long* info) {
  if (!h || !info || mode < 0 || mode > 2) return SPN_ERR_INVALID_ARGUMENT;
  const ogs::ogsExchangeData_t& d = h->ogs.data[mode];
  info[0] = d.Nsend; info[1] = d.Nrecv; info[2] = d.NrowsP; info[3] = d.Nrows;
  info[4] = (long long)d.sendRanks.size(); info[5] = (long long)d.recvRanks.size(); info[6] = info[7] = 0;
  return SPN_SUCCESS;
}
int spn_host_ogs_array(spn_host_ogs_t h, int which, int mode, const int** ptr, long long* len) {
  if (!h || !ptr || !len || mode < 0 || mode > 2) return SPN_ERR_INVALID_ARGUMENT;
  const ogs::ogs_t& o = h->ogs;
  if (!o.gatherLocal || !o.gatherHalo) return SPN_ERR_INVALID_ARGUMENT;
  const ogs::ogsExchangeData_t& d = o.data[mode];
  const int* data = nullptr;
  size_t count = 0;
  auto set = [&](const auto& vec) { data = vec.data(); count = vec.size(); };
  // an Unsigned operator aliases N to T (libs/ogs/ogsOperator.cpp:644-650)
  auto nOrT = [&](const ogs::ogsOperator_t& op, bool starts) {
    if (op.kind == ogs::Signed) set(starts ? op.rowStartsN : op.colIdsN);
    else set(starts ? op.rowStartsT : op.colIdsT);
  };
  switch (which) {
    case 0: set(o.gatherLocal->rowStartsT); break;
    case 1: set(o.gatherLocal->colIdsT); break;
    case 2: nOrT(*o.gatherLocal, true); break;
    case 3: nOrT(*o.gatherLocal, false); break;
    case 4: set(o.gatherHalo->rowStartsT); break;
    case 5: set(o.gatherHalo->colIdsT); break;
    case 6: nOrT(*o.gatherHalo, true); break;
    case 7: nOrT(*o.gatherHalo, false); break;
    case 8: set(d.sendIds); break;
    case 9: set(d.sendRanks); break;
    case 10: set(d.sendCounts); break;
    case 11: set(d.sendOffsets); break;
    case 12: set(d.recvRanks); break;
    case 13: set(d.recvCounts); break;
    case 14: set(d.recvOffsets); break;
    case 15: set(d.cmbStarts); break;
    case 16: set(d.cmbIds); break;
    case 17: set(d.recvRows); break;
    case 18: set(d.rowSendStarts); break;
    case 19: set(d.rowSendDst); break;
    case 20: set(d.slotDst); break;
    default: return SPN_ERR_INVALID_ARGUMENT;
  }
  *ptr = data;
  *len = (long long)count;
  return SPN_SUCCESS;
}
int spn_host_ogs_halo_labels(spn_host_ogs_t h, const long long** ptr, long long* len) {
  if (!h || !ptr || !len || !h->ogs.gatherHalo) return SPN_ERR_INVALID_ARGUMENT;
  *ptr = h->ogs.gatherHalo->rowLabel.data();
  *len = (long long)h->ogs.gatherHalo->rowLabel.size();
  return SPN_SUCCESS;
}

int spn_host_ogs_global_to_local(spn_host_ogs_t h, int* out) {
  if (!h || !out) return SPN_ERR_INVALID_ARGUMENT;
  return guarded([&] { h->ogs.SetupGlobalToLocalMapping(out); });
}

int spn_host_ogs_gather_scatter(spn_host_ogs_t h, void* v, int k, int type, int op, int trans) {
  if (!h) return SPN_ERR_INVALID_ARGUMENT;
  return guarded([&] {
    h->ogs.GatherScatter(v, k, (ogs::Type)type, (ogs::Op)op, (ogs::Transpose)trans);
  });
}
int spn_host_ogs_gather_scatter_start(spn_host_ogs_t h, void* v, int k, int type, int op, int trans) {
  if (!h) return SPN_ERR_INVALID_ARGUMENT;
  return guarded([&] { h->ogs.GatherScatterStart(v, k, (ogs::Type)type, (ogs::Op)op, (ogs::Transpose)trans); });
}
int spn_host_ogs_gather_scatter_finish(spn_host_ogs_t h, void* v, int k, int type, int op, int trans) {
  if (!h) return SPN_ERR_INVALID_ARGUMENT;
  return guarded([&] { h->ogs.GatherScatterFinish(v, k, (ogs::Type)type, (ogs::Op)op, (ogs::Transpose)trans); });
}
int spn_host_ogs_gather_scatter_host_start(spn_host_ogs_t h, void* h_v, int k, int type, int op, int trans) {
  if (!h) return SPN_ERR_INVALID_ARGUMENT;
  return guarded([&] { h->ogs.HostGatherScatterStart(h_v, k, (ogs::Type)type, (ogs::Op)op, (ogs::Transpose)trans); });
}
int spn_host_ogs_gather_scatter_host_finish(spn_host_ogs_t h, void* h_v) {
  if (!h) return SPN_ERR_INVALID_ARGUMENT;
  return guarded([&] { h->ogs.HostGatherScatterFinish(h_v); });
}
int spn_host_debug_set(int key, int value) {
#ifdef SPN_TIMING_PROBES
  if (key == 0) { spn::ogs::g_debug_skip = value; return SPN_SUCCESS; }
#endif
  if (key == 1) { spn::ogs::SetPositivePickOrder(value); return SPN_SUCCESS; }
  return SPN_ERR_INVALID_ARGUMENT;
}
// halo_t operations on an ogs handle set up with kind 2 (Halo): Exchange = owner value to every copy, Combine = every
// copy added into the owner (libs/ogs/ogsHalo.cpp:39-400)
static int halo_op(spn_host_ogs_t h, void* v, int k, int type, int trans, int phase) {
  if (!h) return SPN_ERR_INVALID_ARGUMENT;
  return guarded([&] {
    SPN_ABORT("halo operation on an ogs_t that was not set up with Kind Halo", !h->ogs.halo_only);
    if (phase != 2) h->ogs.GatherScatterStart(v, k, (ogs::Type)type, ogs::Add, (ogs::Transpose)trans);
    if (phase != 1) h->ogs.GatherScatterFinish(v, k, (ogs::Type)type, ogs::Add, (ogs::Transpose)trans);
  });
}
int spn_host_halo_exchange(spn_host_ogs_t h, void* d_v, int k, int type) { return halo_op(h, d_v, k, type, ogs::NoTrans, 0); }
int spn_host_halo_exchange_start(spn_host_ogs_t h, void* d_v, int k, int type) { return halo_op(h, d_v, k, type, ogs::NoTrans, 1); }
int spn_host_halo_exchange_finish(spn_host_ogs_t h, void* d_v, int k, int type) { return halo_op(h, d_v, k, type, ogs::NoTrans, 2); }
int spn_host_halo_combine(spn_host_ogs_t h, void* d_v, int k, int type) { return halo_op(h, d_v, k, type, ogs::Trans, 0); }
int spn_host_halo_combine_start(spn_host_ogs_t h, void* d_v, int k, int type) { return halo_op(h, d_v, k, type, ogs::Trans, 1); }
int spn_host_halo_combine_finish(spn_host_ogs_t h, void* d_v, int k, int type) { return halo_op(h, d_v, k, type, ogs::Trans, 2); }
int spn_host_ogs_gather(spn_host_ogs_t h, void* gv, const void* v, int k, int type, int op, int trans) {
  if (!h) return SPN_ERR_INVALID_ARGUMENT;
  return guarded([&] {
    h->ogs.GatherStart(gv, v, k, (ogs::Type)type, (ogs::Op)op, (ogs::Transpose)trans);
    h->ogs.GatherFinish(gv, v, k, (ogs::Type)type, (ogs::Op)op, (ogs::Transpose)trans);
  });
}
int spn_host_ogs_scatter(spn_host_ogs_t h, void* v, const void* gv, int k, int type, int trans) {
  if (!h) return SPN_ERR_INVALID_ARGUMENT;
  return guarded([&] {
    h->ogs.ScatterStart(v, gv, k, (ogs::Type)type, (ogs::Transpose)trans);
    h->ogs.ScatterFinish(v, gv, k, (ogs::Type)type, (ogs::Transpose)trans);
  });
}

}  // extern "C"
