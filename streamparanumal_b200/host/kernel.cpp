// buildKernel: (kernel name, defines) -> precompiled sm_100a entry point of include/spn_b200.h.
// Reference call sites: BS/BS{0..5}/src/bs*Setup.cpp (bs0, bs1, bs2, bs3_1/2, bs4_1/2, bs5_1/2) and
// libs/ogs/ogsUtils.cpp:105-120 (gatherScatter, gather, scatter, extract).
#include "kernel.hpp"

namespace spn {

namespace {
using A = const kernelArg_t*;

int ogs_type_of(const properties_t& props) {
  const std::string t = props.get("defines/T", "double");
  if (t == "float") return SPN_FLOAT;
  if (t == "double") return SPN_DOUBLE;
  if (t == "int") return SPN_INT32;
  if (t == "long long int") return SPN_INT64;
  throw exception("buildKernel: no native ogs kernel for defines/T = '" + t + "'");
}
int ogs_op_of(const properties_t& props) {
  const std::string o = props.get("defines/OGS_OP(a,b)", "a+=b");  // libs/ogs/ogsUtils.cpp:97-102
  if (o == "a+=b") return SPN_ADD;
  if (o == "a*=b") return SPN_MUL;
  if (o == "if(b>a) a=b") return SPN_MAX;
  if (o == "if(b<a) a=b") return SPN_MIN;
  throw exception("buildKernel: no native ogs kernel for OGS_OP = '" + o + "'");
}
inline const double* cd(const kernelArg_t& a) { return static_cast<const double*>(a.p); }
inline double* md(const kernelArg_t& a) { return static_cast<double*>(a.p); }
inline const int* ci(const kernelArg_t& a) { return static_cast<const int*>(a.p); }
inline int type_of(int ctx) { return ctx & 0xff; }
inline int op_of(int ctx) { return (ctx >> 8) & 0xff; }

int k_bs0(A a, spn_stream_t s, int) { return spn_bs0((int)a[0].i, cd(a[1]), md(a[2]), s); }
int k_bs1(A a, spn_stream_t s, int) { return spn_copy_f64(a[0].i, cd(a[1]), md(a[2]), s); }
int k_bs2(A a, spn_stream_t s, int) { return spn_axpby_f64(a[0].i, a[1].d, cd(a[2]), a[3].d, md(a[4]), s); }
int k_bs3_1(A a, spn_stream_t s, int) { return spn_bs3_1((int)a[0].i, (int)a[1].i, cd(a[2]), md(a[3]), s); }
int k_bs3_2(A a, spn_stream_t s, int) { return spn_bs3_2((int)a[0].i, cd(a[1]), md(a[2]), s); }
int k_bs4_1(A a, spn_stream_t s, int) { return spn_bs4_1((int)a[0].i, (int)a[1].i, cd(a[2]), cd(a[3]), md(a[4]), s); }
int k_bs4_2(A a, spn_stream_t s, int) { return spn_bs4_2((int)a[0].i, cd(a[1]), md(a[2]), s); }
int k_bs5_1(A a, spn_stream_t s, int) { return spn_bs5_1((int)a[0].i, (int)a[1].i, cd(a[2]), a[3].d, md(a[4]), md(a[5]), s); }
int k_bs5_2(A a, spn_stream_t s, int) { return spn_bs5_2((int)a[0].i, cd(a[1]), md(a[2]), s); }
int k_gather(A a, spn_stream_t s, int c) {
  return spn_ogs_gather((int)a[0].i, (int)a[1].i, ci(a[2]), ci(a[3]), ci(a[4]), a[5].p, a[6].p, type_of(c), op_of(c), s);
}
int k_scatter(A a, spn_stream_t s, int c) {
  return spn_ogs_scatter((int)a[0].i, (int)a[1].i, ci(a[2]), ci(a[3]), ci(a[4]), a[5].p, a[6].p, type_of(c), s);
}
int k_gather_scatter(A a, spn_stream_t s, int c) {
  return spn_ogs_gather_scatter((int)a[0].i, (int)a[1].i, ci(a[2]), ci(a[3]), ci(a[4]), ci(a[5]), ci(a[6]), a[7].p, type_of(c), op_of(c), s);
}
int k_extract(A a, spn_stream_t s, int c) { return spn_ogs_extract((int)a[0].i, (int)a[1].i, ci(a[2]), a[3].p, a[4].p, type_of(c), s); }
}  // namespace

kernel_t buildKernel(platform_t& platform, const std::string& fileName, const std::string& kernelName,
                     const properties_t& props) {
  (void)fileName;  // the OKL file is not read: every kernel it holds is already compiled for sm_100a
  const auto cur = platform.streamSlot();
  auto make = [&](int nargs, kernel_t::impl_t f, int ctx = 0) { return kernel_t(kernelName, nargs, f, ctx, cur); };

  if (kernelName == "bs0") return make(3, k_bs0);
  if (kernelName == "bs1") return make(3, k_bs1);
  if (kernelName == "bs2") return make(5, k_bs2);
  if (kernelName == "bs3_1") return make(4, k_bs3_1);
  if (kernelName == "bs4_1") return make(5, k_bs4_1);
  if (kernelName == "bs5_1") return make(6, k_bs5_1);
  if (kernelName == "bs3_2") return make(3, k_bs3_2);
  if (kernelName == "bs4_2") return make(3, k_bs4_2);
  if (kernelName == "bs5_2") return make(3, k_bs5_2);
  if (kernelName == "gather") return make(7, k_gather, ogs_type_of(props) | (ogs_op_of(props) << 8));
  if (kernelName == "scatter") return make(7, k_scatter, ogs_type_of(props));
  if (kernelName == "gatherScatter") return make(8, k_gather_scatter, ogs_type_of(props) | (ogs_op_of(props) << 8));
  if (kernelName == "extract") return make(5, k_extract, ogs_type_of(props));
  throw exception("buildKernel: '" + kernelName + "' (" + fileName +
                  ") has no native sm_100a implementation; this backend has no JIT and no CPU fallback");
}

kernel_t platform_t::buildKernel(const std::string& fileName, const std::string& kernelName, const properties_t& props) {
  return spn::buildKernel(*this, fileName, kernelName, props);
}

}  // namespace spn
