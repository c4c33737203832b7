// TEST INFRASTRUCTURE (oracle/): C binding of the reference's OWN ogs_t and mesh_t, compiled unchanged from
// /root/reference against include/shim/{occa.hpp,mpi.h} with the host backend (oracle/shim/occa_host.cpp).
// Gives the tests (i) the index maps the reference's setup produces (libs/ogs/ogsSetup.cpp:87-290,
// libs/ogs/ogsOperator.cpp:522-665) — the pin for the native host layer's maps — (ii) the reference's host-memory and
// "device" (compiled OKL kernels) Gather / Scatter / GatherScatter on the CPU, and (iii) the global node ids of the
// reference's box-hex mesh (libs/mesh/meshSetupBoxHex3D.cpp, meshConnectNodes.cpp:44-81).
// One rank (the MPI stand-in of include/shim/mpi.h).  Nothing under streamparanumal_b200/ links this.
#include <omp.h>

#include <cstring>
#include <string>

#include "mesh.hpp"
#include "ogs.hpp"
#include "ogs/ogsOperator.hpp"
#include "platform.hpp"
#include "timer.hpp"

using namespace libp;

namespace {

struct Env {
  comm_t comm;
  settings_t settings;
  platform_t platform;
  Env() {
    static int argc = 1;
    static char arg0[] = "ref";
    static char* argv_[] = {arg0, nullptr};
    static char** argv = argv_;
    static bool inited = false;
    if (!inited) { Comm::Init(argc, argv); inited = true; }
    comm = Comm::World().Dup();
    settings = settings_t(comm);
    platformAddSettings(settings);
    settings.changeSetting("THREAD MODEL", "Serial");
    platform = platform_t(settings);
  }
};
Env& env() {
  static Env e;
  return e;
}

// the operators are protected members of ogsBase_t: reach them through a derived class
struct OgsPeek : public ogs::ogs_t {
  ogs::ogsOperator_t* local() { return gatherLocal.get(); }
  ogs::ogsOperator_t* halo() { return gatherHalo.get(); }
};

struct Handle {
  OgsPeek ogs;
  memory<hlong> ids;
};

template <typename T, typename F>
void with_host(void* p, size_t n, F f) {
  memory<T> m(n);
  memcpy(m.ptr(), p, n * sizeof(T));
  f(m);
  memcpy(p, m.ptr(), n * sizeof(T));
}

}  // namespace

extern "C" {

// ids (N entries) is updated in place when unique != 0, as ogs_t::Setup does
void* ref_ogs_create(int N, long long* ids, int kind, int unique, int verbose) {
  Handle* h = new Handle();
  h->ids = memory<hlong>(N);
  memcpy(h->ids.ptr(), ids, sizeof(hlong) * (size_t)N);
  h->ogs.Setup(N, h->ids, env().comm, (ogs::Kind)kind, ogs::Pairwise, unique != 0, verbose != 0, env().platform);
  memcpy(ids, h->ids.ptr(), sizeof(hlong) * (size_t)N);
  return h;
}
void ref_ogs_destroy(void* handle) { delete static_cast<Handle*>(handle); }

// out: N, Ngather, NlocalT, NlocalP, NhaloT, NhaloP, local nnzT, local nnzN, local NrowBlocksT, local NrowBlocksN
void ref_ogs_dims(void* handle, long long* out) {
  Handle* h = static_cast<Handle*>(handle);
  ogs::ogsOperator_t* L = h->ogs.local();
  out[0] = h->ogs.N; out[1] = h->ogs.Ngather; out[2] = h->ogs.NlocalT; out[3] = h->ogs.NlocalP;
  out[4] = h->ogs.NhaloT; out[5] = h->ogs.NhaloP;
  out[6] = L ? L->nnzT : 0; out[7] = L ? L->nnzN : 0; out[8] = L ? L->NrowBlocksT : 0; out[9] = L ? L->NrowBlocksN : 0;
  out[10] = h->ogs.NgatherGlobal;
}

// name: {local,halo}_{rowStartsT,colIdsT,rowStartsN,colIdsN,blockRowStartsT,blockRowStartsN}; returns the host array
const int* ref_ogs_array(void* handle, const char* name, long long* len) {
  Handle* h = static_cast<Handle*>(handle);
  const std::string s(name);
  ogs::ogsOperator_t* op = s.rfind("local_", 0) == 0 ? h->ogs.local() : h->ogs.halo();
  *len = 0;
  if (!op) return nullptr;
  const std::string f = s.substr(s.find('_') + 1);
  memory<dlong> m;
  if (f == "rowStartsT") m = op->rowStartsT;
  else if (f == "colIdsT") m = op->colIdsT;
  else if (f == "rowStartsN") m = op->rowStartsN;
  else if (f == "colIdsN") m = op->colIdsN;
  else if (f == "blockRowStartsT") m = op->blockRowStartsT;
  else if (f == "blockRowStartsN") m = op->blockRowStartsN;
  else return nullptr;
  *len = (long long)m.length();
  return m.ptr();
}

// ogs_t::SetupGlobalToLocalMapping (libs/ogs/ogsSetup.cpp:833-849): out has N entries
void ref_ogs_global_to_local(void* handle, int* out) {
  Handle* h = static_cast<Handle*>(handle);
  memory<dlong> g2l(h->ogs.N > 0 ? h->ogs.N : 1);
  h->ogs.SetupGlobalToLocalMapping(g2l);
  for (dlong n = 0; n < h->ogs.N; ++n) out[n] = g2l[n];
}

// host-memory overloads (libs/ogs/ogs.cpp:127-204, 303-393, 494-578): type 0..3 = float, double, int, long long
#define BY_TYPE(type, ...)                              \
  switch (type) {                                       \
    case 0: { using T = float; __VA_ARGS__; } break;    \
    case 1: { using T = double; __VA_ARGS__; } break;   \
    case 2: { using T = int; __VA_ARGS__; } break;      \
    case 3: { using T = long long int; __VA_ARGS__; } break; \
    default: return -1;                                 \
  }

int ref_ogs_gather_scatter(void* handle, void* v, int k, int type, int op, int trans) {
  Handle* h = static_cast<Handle*>(handle);
  const size_t n = (size_t)h->ogs.N * k;
  BY_TYPE(type, with_host<T>(v, n, [&](memory<T>& m) { h->ogs.GatherScatter(m, k, (ogs::Op)op, (ogs::Transpose)trans); }));
  return 0;
}
int ref_ogs_gather(void* handle, void* gv, void* v, int k, int type, int op, int trans) {
  Handle* h = static_cast<Handle*>(handle);
  const size_t n = (size_t)h->ogs.N * k;
  const size_t ng = (size_t)(trans == (int)ogs::Trans ? h->ogs.Ngather : h->ogs.NlocalT + h->ogs.NhaloT) * k;
  (void)ng;
  const size_t ngv = (size_t)(h->ogs.NlocalT + h->ogs.NhaloT) * k;  // upper bound of what any mode writes
  BY_TYPE(type, {
    memory<T> mv(n), mg(ngv);
    memcpy(mv.ptr(), v, n * sizeof(T));
    memcpy(mg.ptr(), gv, ngv * sizeof(T));
    h->ogs.Gather(mg, mv, k, (ogs::Op)op, (ogs::Transpose)trans);
    memcpy(gv, mg.ptr(), ngv * sizeof(T));
  });
  return 0;
}
int ref_ogs_scatter(void* handle, void* v, void* gv, int k, int type, int trans) {
  Handle* h = static_cast<Handle*>(handle);
  const size_t n = (size_t)h->ogs.N * k;
  const size_t ngv = (size_t)(h->ogs.NlocalT + h->ogs.NhaloT) * k;
  BY_TYPE(type, {
    memory<T> mv(n), mg(ngv);
    memcpy(mv.ptr(), v, n * sizeof(T));
    memcpy(mg.ptr(), gv, ngv * sizeof(T));
    h->ogs.Scatter(mv, mg, k, (ogs::Transpose)trans);
    memcpy(v, mv.ptr(), n * sizeof(T));
  });
  return 0;
}
// the reference's DEVICE path on the host backend: ogs_t::GatherScatter(deviceMemory<T>) -> ogsOperator_t ->
// kernel_t -> the compiled OKL kernels
int ref_ogs_gather_scatter_device(void* handle, void* v, int k, int type, int op, int trans) {
  Handle* h = static_cast<Handle*>(handle);
  const size_t n = (size_t)h->ogs.N * k;
  BY_TYPE(type, {
    memory<T> m(n);
    memcpy(m.ptr(), v, n * sizeof(T));
    deviceMemory<T> o_v = env().platform.malloc<T>(m);
    h->ogs.GatherScatter(o_v, k, (ogs::Op)op, (ogs::Transpose)trans);
    o_v.copyTo(m);
    memcpy(v, m.ptr(), n * sizeof(T));
  });
  return 0;
}

// ---- halo_t (include/ogs.hpp:347-397, libs/ogs/ogsHalo.cpp): ids > 0 the owned copy, ids < 0 halo copies ----------------
struct HaloHandle {
  ogs::halo_t halo;
  memory<hlong> ids;
};
void* ref_halo_create(int N, const long long* ids, int verbose) {
  HaloHandle* h = new HaloHandle();
  h->ids = memory<hlong>(N);
  memcpy(h->ids.ptr(), ids, sizeof(hlong) * (size_t)N);
  h->halo.Setup(N, h->ids, env().comm, ogs::Pairwise, verbose != 0, env().platform);
  return h;
}
void ref_halo_destroy(void* handle) { delete static_cast<HaloHandle*>(handle); }
int ref_halo_nhalo(void* handle) { return static_cast<HaloHandle*>(handle)->halo.Nhalo; }
// host-memory Exchange (op 0) / Combine (op 1) of v (N*k entries) in place
int ref_halo_apply(void* handle, int which, void* v, int k, int type) {
  HaloHandle* h = static_cast<HaloHandle*>(handle);
  const size_t n = (size_t)h->halo.N * k;
  BY_TYPE(type, with_host<T>(v, n, [&](memory<T>& m) {
    if (which == 0) h->halo.Exchange(m, k);
    else h->halo.Combine(m, k);
  }));
  return 0;
}

// Times `steps` calls of the reference's own BS8 operation, mesh.ogs.GatherScatter(q, 1, ogs::Add, ogs::Sym)
// (BS/BS8/src/bs8Run.cpp:47-57), after `warmup` calls, on FP64 data resident in the reference's own containers:
//   path 0: the host-memory overload (memory<dfloat>; libs/ogs/ogs.cpp:127-204 -> the OpenMP-threaded host operator,
//           libs/ogs/ogsOperator.cpp:340-409)
//   path 1: the device-memory overload on the Serial backend (deviceMemory<dfloat> -> kernel_t -> the OKL gatherScatter
//           kernel), i.e. what `BS8 -m Serial` times
// Returns seconds per call; *threads = OpenMP threads the reference configured (platformDeviceConfig.cpp:103-150).
double ref_ogs_time_gather_scatter(void* handle, int path, int steps, int warmup, int* threads) {
  Handle* h = static_cast<Handle*>(handle);
  const size_t n = (size_t)h->ogs.N;
  memory<dfloat> q(n);
#pragma omp parallel for
  for (size_t i = 0; i < n; ++i) q[i] = (dfloat)((i * 2654435761ull) % 2001) * 1e-3 - 1.0;
  if (threads) {
#pragma omp parallel
    {
#pragma omp single
      *threads = omp_get_num_threads();
    }
  }
  deviceMemory<dfloat> o_q;
  if (path == 1) o_q = env().platform.malloc<dfloat>(q);
  auto call = [&]() {
    if (path == 1) h->ogs.GatherScatter(o_q, 1, ogs::Add, ogs::Sym);
    else h->ogs.GatherScatter(q, 1, ogs::Add, ogs::Sym);
  };
  for (int i = 0; i < warmup; ++i) call();
  const timePoint_t t0 = Time();
  for (int i = 0; i < steps; ++i) call();
  const timePoint_t t1 = Time();
  return ElapsedTime(t0, t1) / (steps > 0 ? steps : 1);
}

// global node ids of the reference's own box mesh (one rank) of degree-p elements of type `element` ("Tri", "Quad", "Tet",
// "Hex": libs/mesh/meshSetupBox{Tri2D,Quad2D,Tet3D,Hex3D}.cpp + meshConnectNodes.cpp): Nelements*Np entries.
// Returns the number of entries; writes them when ids_out is non-null and capacity suffices, else writes nothing
// (-1 when a non-null buffer is too small).  nz is ignored by the 2-D element types.
long long ref_mesh_box_ids(const char* element, int nx, int ny, int nz, int p, long long* ids_out, long long capacity) {
  settings_t settings(env().comm);
  platformAddSettings(settings);
  meshAddSettings(settings);
  settings.newSetting("-v", "--verbose", "VERBOSE", "FALSE", "Enable verbose output", {"TRUE", "FALSE"});  // as the BS drivers do
  settings.changeSetting("THREAD MODEL", "Serial");
  settings.changeSetting("ELEMENT TYPE", element);
  settings.changeSetting("BOX NX", std::to_string(nx));
  settings.changeSetting("BOX NY", std::to_string(ny));
  settings.changeSetting("BOX NZ", std::to_string(nz));
  settings.changeSetting("POLYNOMIAL DEGREE", std::to_string(p));
  platform_t platform(settings);
  mesh_t mesh(platform, settings, env().comm);
  const long long n = (long long)mesh.Nelements * mesh.Np;
  if (!ids_out) return n;
  if (n > capacity) return -1;
  memcpy(ids_out, mesh.globalIds.ptr(), sizeof(hlong) * (size_t)n);
  return n;
}
long long ref_mesh_box_hex_ids(int nx, int ny, int nz, int p, long long* ids_out, long long capacity) {
  return ref_mesh_box_ids("Hex", nx, ny, nz, p, ids_out, capacity);
}

}  // extern "C"
