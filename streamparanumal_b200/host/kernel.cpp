// buildKernel: (kernel name, defines) -> precompiled sm_100a entry point of include/spn_b200.h.
// Reference call sites: BS/BS{0..5}/src/bs*Setup.cpp (bs0, bs1, bs2, bs3_1/2, bs4_1/2, bs5_1/2) and
// libs/ogs/ogsUtils.cpp:105-120 (gatherScatter, gather, scatter, extract).
#include "kernel.hpp"

namespace spn {

namespace {
using A = const std::vector<kernelArg_t>&;

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
const double* cd(const kernelArg_t& a) { return static_cast<const double*>(a.p); }
double* md(const kernelArg_t& a) { return static_cast<double*>(a.p); }
const int* ci(const kernelArg_t& a) { return static_cast<const int*>(a.p); }
}  // namespace

kernel_t buildKernel(platform_t& platform, const std::string& fileName, const std::string& kernelName,
                     const properties_t& props) {
  (void)fileName;  // the OKL file is not read: every kernel it holds is already compiled for sm_100a
  const auto cur = platform.streamSlot();
  auto make = [&](size_t nargs, kernel_t::impl_t f) { return kernel_t(kernelName, nargs, std::move(f), cur); };

  if (kernelName == "bs0") return make(3, [](A a, spn_stream_t s) { return spn_bs0((int)a[0].i, cd(a[1]), md(a[2]), s); });
  if (kernelName == "bs1") return make(3, [](A a, spn_stream_t s) { return spn_copy_f64(a[0].i, cd(a[1]), md(a[2]), s); });
  if (kernelName == "bs2")
    return make(5, [](A a, spn_stream_t s) { return spn_axpby_f64(a[0].i, a[1].d, cd(a[2]), a[3].d, md(a[4]), s); });
  if (kernelName == "bs3_1")
    return make(4, [](A a, spn_stream_t s) { return spn_bs3_1((int)a[0].i, (int)a[1].i, cd(a[2]), md(a[3]), s); });
  if (kernelName == "bs4_1")
    return make(5, [](A a, spn_stream_t s) { return spn_bs4_1((int)a[0].i, (int)a[1].i, cd(a[2]), cd(a[3]), md(a[4]), s); });
  if (kernelName == "bs5_1")
    return make(6, [](A a, spn_stream_t s) {
      return spn_bs5_1((int)a[0].i, (int)a[1].i, cd(a[2]), a[3].d, md(a[4]), md(a[5]), s);
    });
  if (kernelName == "bs3_2") return make(3, [](A a, spn_stream_t s) { return spn_bs3_2((int)a[0].i, cd(a[1]), md(a[2]), s); });
  if (kernelName == "bs4_2") return make(3, [](A a, spn_stream_t s) { return spn_bs4_2((int)a[0].i, cd(a[1]), md(a[2]), s); });
  if (kernelName == "bs5_2") return make(3, [](A a, spn_stream_t s) { return spn_bs5_2((int)a[0].i, cd(a[1]), md(a[2]), s); });

  if (kernelName == "gather") {
    const int type = ogs_type_of(props), op = ogs_op_of(props);
    return make(7, [type, op](A a, spn_stream_t s) {
      return spn_ogs_gather((int)a[0].i, (int)a[1].i, ci(a[2]), ci(a[3]), ci(a[4]), a[5].p, a[6].p, type, op, s);
    });
  }
  if (kernelName == "scatter") {
    const int type = ogs_type_of(props);
    return make(7, [type](A a, spn_stream_t s) {
      return spn_ogs_scatter((int)a[0].i, (int)a[1].i, ci(a[2]), ci(a[3]), ci(a[4]), a[5].p, a[6].p, type, s);
    });
  }
  if (kernelName == "gatherScatter") {
    const int type = ogs_type_of(props), op = ogs_op_of(props);
    return make(8, [type, op](A a, spn_stream_t s) {
      return spn_ogs_gather_scatter((int)a[0].i, (int)a[1].i, ci(a[2]), ci(a[3]), ci(a[4]), ci(a[5]), ci(a[6]), a[7].p,
                                    type, op, s);
    });
  }
  if (kernelName == "extract") {
    const int type = ogs_type_of(props);
    return make(5, [type](A a, spn_stream_t s) {
      return spn_ogs_extract((int)a[0].i, (int)a[1].i, ci(a[2]), a[3].p, a[4].p, type, s);
    });
  }
  throw exception("buildKernel: '" + kernelName + "' (" + fileName +
                  ") has no native sm_100a implementation; this backend has no JIT and no CPU fallback");
}

kernel_t platform_t::buildKernel(const std::string& fileName, const std::string& kernelName, const properties_t& props) {
  return spn::buildKernel(*this, fileName, kernelName, props);
}

}  // namespace spn
