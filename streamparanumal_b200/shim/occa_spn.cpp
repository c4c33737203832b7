// Product backend of include/shim/occa.hpp: the OCCA handle API of the reference's L1-L5 code on top of the C ABI
// of include/spn_b200.h (hand-written sm_100a kernels).  This is what replaces occa/src/occa/internal/modes/cuda
// for streamParanumal's thirteen kernels; mode "CUDA" is the only one it implements — there is no CPU fallback.
//   buildKernel(file, name, props)  ->  one precompiled entry point, picked by the kernel name and, for the four ogs
//   kernels, props["defines/T"] and props["defines/OGS_OP(a,b)"] (libs/ogs/ogsUtils.cpp:46-120).
// Kernel argument lists follow the reference call sites: BS/BS0..5/src/bs*Run.cpp, libs/ogs/ogsOperator.cpp:187-202,
// 305-320, 452-479, libs/ogs/ogsPairwise.cpp:115.
#include <occa.hpp>

#include <cstring>

#include "spn_b200.h"

namespace occa {
namespace backend {

namespace {
void check(int status, const char* where) {
  if (status != 0) throw occa::exception(std::string(where) + ": " + spn_error_string(status) + " (status " + std::to_string(status) + ")");
}
#define SPN_OK(call) check((call), #call)

int ogs_type_of(const json& props) {
  const std::string t = props.get("defines/T", std::string("double"));
  if (t == "float") return SPN_FLOAT;
  if (t == "double") return SPN_DOUBLE;
  if (t == "int") return SPN_INT32;
  if (t == "long long int") return SPN_INT64;
  throw occa::exception("buildKernel: no native ogs kernel for defines/T = '" + t + "'");
}
int ogs_op_of(const json& props) {
  const std::string o = props.get("defines/OGS_OP(a,b)", std::string("a+=b"));
  if (o == "a+=b") return SPN_ADD;
  if (o == "a*=b") return SPN_MUL;
  if (o == "if(b>a) a=b") return SPN_MAX;
  if (o == "if(b<a) a=b") return SPN_MIN;
  throw occa::exception("buildKernel: no native ogs kernel for OGS_OP = '" + o + "'");
}

// ctx of the ogs kernels: type | op << 8
inline int ctx_type(const void* ctx) { return (int)(reinterpret_cast<uintptr_t>(ctx) & 0xff); }
inline int ctx_op(const void* ctx) { return (int)((reinterpret_cast<uintptr_t>(ctx) >> 8) & 0xff); }
inline const void* make_ctx(int type, int op) { return reinterpret_cast<const void*>((uintptr_t)(type | (op << 8))); }

inline const double* cd(const kernelArg& a) { return static_cast<const double*>(a.asPointer()); }
inline double* md(const kernelArg& a) { return static_cast<double*>(a.asPointer()); }
inline const int* ci(const kernelArg& a) { return static_cast<const int*>(a.asPointer()); }

void k_bs0(const void*, const kernelArg* a, int, void* s) { SPN_OK(spn_bs0((int)a[0].asInt(), cd(a[1]), md(a[2]), s)); }
void k_bs1(const void*, const kernelArg* a, int, void* s) { SPN_OK(spn_copy_f64(a[0].asInt(), cd(a[1]), md(a[2]), s)); }
void k_bs2(const void*, const kernelArg* a, int, void* s) {
  SPN_OK(spn_axpby_f64(a[0].asInt(), a[1].asDouble(), cd(a[2]), a[3].asDouble(), md(a[4]), s));
}
void k_bs3_1(const void*, const kernelArg* a, int, void* s) { SPN_OK(spn_bs3_1((int)a[0].asInt(), (int)a[1].asInt(), cd(a[2]), md(a[3]), s)); }
void k_bs3_2(const void*, const kernelArg* a, int, void* s) { SPN_OK(spn_bs3_2((int)a[0].asInt(), cd(a[1]), md(a[2]), s)); }
void k_bs4_1(const void*, const kernelArg* a, int, void* s) {
  SPN_OK(spn_bs4_1((int)a[0].asInt(), (int)a[1].asInt(), cd(a[2]), cd(a[3]), md(a[4]), s));
}
void k_bs4_2(const void*, const kernelArg* a, int, void* s) { SPN_OK(spn_bs4_2((int)a[0].asInt(), cd(a[1]), md(a[2]), s)); }
void k_bs5_1(const void*, const kernelArg* a, int, void* s) {
  SPN_OK(spn_bs5_1((int)a[0].asInt(), (int)a[1].asInt(), cd(a[2]), a[3].asDouble(), md(a[4]), md(a[5]), s));
}
void k_bs5_2(const void*, const kernelArg* a, int, void* s) { SPN_OK(spn_bs5_2((int)a[0].asInt(), cd(a[1]), md(a[2]), s)); }

void k_gather(const void* ctx, const kernelArg* a, int, void* s) {
  SPN_OK(spn_ogs_gather((int)a[0].asInt(), (int)a[1].asInt(), ci(a[2]), ci(a[3]), ci(a[4]), a[5].asPointer(), a[6].asPointer(),
                        ctx_type(ctx), ctx_op(ctx), s));
}
void k_scatter(const void* ctx, const kernelArg* a, int, void* s) {
  SPN_OK(spn_ogs_scatter((int)a[0].asInt(), (int)a[1].asInt(), ci(a[2]), ci(a[3]), ci(a[4]), a[5].asPointer(), a[6].asPointer(),
                         ctx_type(ctx), s));
}
void k_gather_scatter(const void* ctx, const kernelArg* a, int, void* s) {
  SPN_OK(spn_ogs_gather_scatter((int)a[0].asInt(), (int)a[1].asInt(), ci(a[2]), ci(a[3]), ci(a[4]), ci(a[5]), ci(a[6]),
                                a[7].asPointer(), ctx_type(ctx), ctx_op(ctx), s));
}
void k_extract(const void* ctx, const kernelArg* a, int, void* s) {
  SPN_OK(spn_ogs_extract((int)a[0].asInt(), (int)a[1].asInt(), ci(a[2]), a[3].asPointer(), a[4].asPointer(), ctx_type(ctx), s));
}
}  // namespace

std::string deviceSetup(const std::string& mode, int device_id) {
  if (mode != "CUDA")
    throw occa::exception("occa shim: mode '" + mode + "' requested, but this backend runs sm_100a CUDA only (no CPU fallback)");
  SPN_OK(spn_set_device(device_id));
  return "CUDA";
}
int deviceCount(const std::string&) {
  int n = 0;
  spn_device_count(&n);
  return n;
}
void* deviceMalloc(size_t bytes) {
  void* p = nullptr;
  SPN_OK(spn_malloc(&p, bytes));
  return p;
}
void deviceFree(void* p) { spn_free(p); }
void* hostMalloc(size_t bytes) {
  void* p = nullptr;
  SPN_OK(spn_host_malloc(&p, bytes));
  return p;
}
void hostFree(void* p) { spn_host_free(p); }
void copy(void* dst, const void* src, size_t bytes, copyKind kind, void* stream, bool async) {
  switch (kind) {
    case h2d: SPN_OK(spn_memcpy_h2d(dst, src, bytes, stream, async)); break;
    case d2h: SPN_OK(spn_memcpy_d2h(dst, src, bytes, stream, async)); break;
    case d2d: SPN_OK(spn_memcpy_d2d(dst, src, bytes, stream, async)); break;
    case h2h:
      // pinned <-> pageable host copy: ordered after the stream like a blocking cuMemcpy would be
      if (!async) SPN_OK(spn_stream_sync(stream));
      memcpy(dst, src, bytes);
      break;
  }
}
void* streamCreate() {
  spn_stream_t s = nullptr;
  SPN_OK(spn_stream_create(&s));
  return s;
}
void streamSync(void* stream) { SPN_OK(spn_stream_sync(stream)); }

kernelEntry lookupKernel(const std::string& file, const std::string& name, const json& props) {
  kernelEntry e;
  auto set = [&](launch_fn f, int nargs, const void* ctx = nullptr) { e.launch = f; e.nargs = nargs; e.ctx = ctx; return e; };
  if (name == "bs0") return set(k_bs0, 3);
  if (name == "bs1") return set(k_bs1, 3);
  if (name == "bs2") return set(k_bs2, 5);
  if (name == "bs3_1") return set(k_bs3_1, 4);
  if (name == "bs3_2") return set(k_bs3_2, 3);
  if (name == "bs4_1") return set(k_bs4_1, 5);
  if (name == "bs4_2") return set(k_bs4_2, 3);
  if (name == "bs5_1") return set(k_bs5_1, 6);
  if (name == "bs5_2") return set(k_bs5_2, 3);
  if (name == "gather") return set(k_gather, 7, make_ctx(ogs_type_of(props), ogs_op_of(props)));
  if (name == "scatter") return set(k_scatter, 7, make_ctx(ogs_type_of(props), 0));
  if (name == "gatherScatter") return set(k_gather_scatter, 8, make_ctx(ogs_type_of(props), ogs_op_of(props)));
  if (name == "extract") return set(k_extract, 5, make_ctx(ogs_type_of(props), 0));
  throw occa::exception("buildKernel: '" + name + "' (" + file +
                        ") has no native sm_100a implementation; this backend has no JIT and no CPU fallback");
}

}  // namespace backend
}  // namespace occa
