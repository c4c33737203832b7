// TEST INFRASTRUCTURE (oracle/): host backend of include/shim/occa.hpp.  "Device" memory is host memory, streams are
// no-ops, and buildKernel resolves to the reference's OWN OKL kernels compiled for the host by oracle/Makefile
// (oracle/_ref/*.o, Serial semantics).  Linked with the reference's unmodified libs/core, libs/primitives, libs/ogs,
// libs/mesh and BS drivers it gives the reference's `-m Serial` executables and its ogs_t on the CPU
// (oracle/_ref/BS0..8, oracle/_ref/libref_ogs.so).  Nothing under streamparanumal_b200/ may link this file.
#include <occa.hpp>

#include <cstdlib>
#include <cstring>

extern "C" {
// BS/BS<n>/okl/bs<n>.okl (p_blockSize = 256, the Serial block size: BS/BS5/src/bs5Setup.cpp:37)
void bs0(const int N, const double* a, double* b);
void bs1(const int N, const double* a, double* b);
void bs2(const int N, const double alpha, const double* x, const double beta, double* y);
void bs3_1(const int Nblocks, const int N, const double* a, double* norm);
void bs3_2(const int Nblocks, const double* a, double* norm);
void bs4_1(const int Nblocks, const int N, const double* a, const double* b, double* dot);
void bs4_2(const int Nblocks, const double* a, double* dot);
void bs5_1(const int Nblocks, const int N, const double* Ap, const double alpha, double* r, double* rdotr);
void bs5_2(const int Nblocks, const double* a, double* rdotr);
// libs/ogs/okl/ogsKernels.okl, one instantiation per (T, OGS_OP)
#define OGS_DECL(TAG, OPTAG, T)                                                                                         \
  void ref_gatherScatter_##TAG##_##OPTAG(const int Nblocks, const int K, const int* blockStarts, const int* gatherStarts, \
                                         const int* gatherIds, const int* scatterStarts, const int* scatterIds, T* q);   \
  void ref_gather_##TAG##_##OPTAG(const int Nblocks, const int K, const int* blockStarts, const int* gatherStarts,       \
                                  const int* gatherIds, const T* q, T* gatherq);                                        \
  void ref_scatter_##TAG##_##OPTAG(const int Nblocks, const int K, const int* blockStarts, const int* scatterStarts,     \
                                   const int* scatterIds, const T* gatherq, T* q);                                      \
  void ref_extract_##TAG##_##OPTAG(const int N, const int K, const int* ids, const T* q, T* gatherq);
#define OGS_DECL4(TAG, T) OGS_DECL(TAG, add, T) OGS_DECL(TAG, mul, T) OGS_DECL(TAG, max, T) OGS_DECL(TAG, min, T)
OGS_DECL4(float, float)
OGS_DECL4(double, double)
OGS_DECL4(int32, int)
OGS_DECL4(int64, long long int)
}

namespace occa {
namespace backend {

namespace {
int type_of(const json& props) {
  const std::string t = props.get("defines/T", std::string("double"));
  if (t == "float") return 0;
  if (t == "double") return 1;
  if (t == "int") return 2;
  if (t == "long long int") return 3;
  throw occa::exception("host backend: no reference kernel for defines/T = '" + t + "'");
}
int op_of(const json& props) {
  const std::string o = props.get("defines/OGS_OP(a,b)", std::string("a+=b"));
  if (o == "a+=b") return 0;
  if (o == "a*=b") return 1;
  if (o == "if(b>a) a=b") return 2;
  if (o == "if(b<a) a=b") return 3;
  throw occa::exception("host backend: no reference kernel for OGS_OP = '" + o + "'");
}

typedef void (*gs_fn)(const int, const int, const int*, const int*, const int*, const int*, const int*, void*);
typedef void (*g_fn)(const int, const int, const int*, const int*, const int*, const void*, void*);
typedef void (*x_fn)(const int, const int, const int*, const void*, void*);
#define ROW(KIND, FN, TAG) {(FN)ref_##KIND##_##TAG##_add, (FN)ref_##KIND##_##TAG##_mul, (FN)ref_##KIND##_##TAG##_max, (FN)ref_##KIND##_##TAG##_min}
const gs_fn kGatherScatter[4][4] = {ROW(gatherScatter, gs_fn, float), ROW(gatherScatter, gs_fn, double), ROW(gatherScatter, gs_fn, int32),
                                    ROW(gatherScatter, gs_fn, int64)};
const g_fn kGather[4][4] = {ROW(gather, g_fn, float), ROW(gather, g_fn, double), ROW(gather, g_fn, int32), ROW(gather, g_fn, int64)};
const g_fn kScatter[4][4] = {ROW(scatter, g_fn, float), ROW(scatter, g_fn, double), ROW(scatter, g_fn, int32), ROW(scatter, g_fn, int64)};
const x_fn kExtract[4][4] = {ROW(extract, x_fn, float), ROW(extract, x_fn, double), ROW(extract, x_fn, int32), ROW(extract, x_fn, int64)};

inline int I(const kernelArg& a) { return (int)a.asInt(); }
inline double D(const kernelArg& a) { return a.asDouble(); }
inline double* P(const kernelArg& a) { return static_cast<double*>(a.asPointer()); }
inline const int* Q(const kernelArg& a) { return static_cast<const int*>(a.asPointer()); }

void k_bs0(const void*, const kernelArg* a, int, void*) { bs0(I(a[0]), P(a[1]), P(a[2])); }
void k_bs1(const void*, const kernelArg* a, int, void*) { bs1(I(a[0]), P(a[1]), P(a[2])); }
void k_bs2(const void*, const kernelArg* a, int, void*) { bs2(I(a[0]), D(a[1]), P(a[2]), D(a[3]), P(a[4])); }
void k_bs3_1(const void*, const kernelArg* a, int, void*) { bs3_1(I(a[0]), I(a[1]), P(a[2]), P(a[3])); }
void k_bs3_2(const void*, const kernelArg* a, int, void*) { bs3_2(I(a[0]), P(a[1]), P(a[2])); }
void k_bs4_1(const void*, const kernelArg* a, int, void*) { bs4_1(I(a[0]), I(a[1]), P(a[2]), P(a[3]), P(a[4])); }
void k_bs4_2(const void*, const kernelArg* a, int, void*) { bs4_2(I(a[0]), P(a[1]), P(a[2])); }
void k_bs5_1(const void*, const kernelArg* a, int, void*) { bs5_1(I(a[0]), I(a[1]), P(a[2]), D(a[3]), P(a[4]), P(a[5])); }
void k_bs5_2(const void*, const kernelArg* a, int, void*) { bs5_2(I(a[0]), P(a[1]), P(a[2])); }
void k_gs(const void* f, const kernelArg* a, int, void*) {
  ((gs_fn)f)(I(a[0]), I(a[1]), Q(a[2]), Q(a[3]), Q(a[4]), Q(a[5]), Q(a[6]), a[7].asPointer());
}
void k_g(const void* f, const kernelArg* a, int, void*) {
  ((g_fn)f)(I(a[0]), I(a[1]), Q(a[2]), Q(a[3]), Q(a[4]), a[5].asPointer(), a[6].asPointer());
}
void k_x(const void* f, const kernelArg* a, int, void*) { ((x_fn)f)(I(a[0]), I(a[1]), Q(a[2]), a[3].asPointer(), a[4].asPointer()); }
}  // namespace

std::string deviceSetup(const std::string& mode, int) {
  if (mode != "Serial" && mode != "OpenMP")
    throw occa::exception("host backend of the oracle: mode '" + mode + "' is not a CPU mode");
  return "Serial";  // the kernels are the Serial translation whichever CPU mode was asked for
}
int deviceCount(const std::string&) { return 1; }
void* deviceMalloc(size_t bytes) {
  void* p = nullptr;
  if (posix_memalign(&p, 64, bytes ? bytes : 64) != 0) throw occa::exception("host backend: out of memory");
  return p;
}
void deviceFree(void* p) { free(p); }
void* hostMalloc(size_t bytes) { return deviceMalloc(bytes); }
void hostFree(void* p) { free(p); }
void copy(void* dst, const void* src, size_t bytes, copyKind, void*, bool) { memmove(dst, src, bytes); }
void* streamCreate() { return nullptr; }
void streamSync(void*) {}

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
  if (name == "gatherScatter") return set(k_gs, 8, (const void*)kGatherScatter[type_of(props)][op_of(props)]);
  if (name == "gather") return set(k_g, 7, (const void*)kGather[type_of(props)][op_of(props)]);
  if (name == "scatter") return set(k_g, 7, (const void*)kScatter[type_of(props)][op_of(props)]);
  if (name == "extract") return set(k_x, 5, (const void*)kExtract[type_of(props)][op_of(props)]);
  throw occa::exception("host backend: no reference kernel named '" + name + "' (" + file + ")");
}

}  // namespace backend
}  // namespace occa
