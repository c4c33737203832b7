// spn::kernel_t / spn::properties_t — the reference's kernel handle and property bag
// (include/utils.hpp:44-47: kernel_t = occa::kernel, properties_t = occa::json) as the hot path
// uses them:
//   properties_t props;  props["defines/p_blockSize"] = 256;  props["defines/T"] = "double";
//   kernel_t k = platform.buildKernel("BS/BS5/okl/bs5.okl", "bs5_1", props);   (libs/core/platformBuildKernel.cpp:31-52)
//   k(Nblock, N, o_Ap, alpha, o_r, o_tmp);                                      (BS/BS5/src/bs5Run.cpp:62,87)
// Nothing is JIT-compiled here: buildKernel resolves (kernel name, defines/T, defines/OGS_OP) to
// one precompiled sm_100a entry point of include/spn_b200.h and fails loudly for anything else.
// A call enqueues on the platform's current stream and returns (occa::kernel::operator() on the
// CUDA backend: occa/src/occa/internal/modes/cuda/kernel.cpp:74-101).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "platform.hpp"

namespace spn {

// path-keyed property bag: the slice of occa::json the call sites touch (operator[] with "a/b"
// paths, assignment of int / double / string, copy)
class properties_t {
 public:
  class value_t {
   public:
    value_t& operator=(int v) { s_ = std::to_string(v); return *this; }
    value_t& operator=(long long v) { s_ = std::to_string(v); return *this; }
    value_t& operator=(double v) { s_ = std::to_string(v); return *this; }
    value_t& operator=(const char* v) { s_ = v; return *this; }
    value_t& operator=(const std::string& v) { s_ = v; return *this; }
    value_t& operator+=(const std::string& v) { s_ += v; return *this; }
    const std::string& string() const { return s_; }
    bool empty() const { return s_.empty(); }
   private:
    std::string s_;
  };
  properties_t() = default;
  properties_t(const std::string& key, bool v) { kv_[key] = v ? "true" : "false"; }  // properties_t("async", true)
  value_t& operator[](const std::string& path) { return kv_[path]; }
  std::string get(const std::string& path, const std::string& dflt = "") const {
    auto it = kv_.find(path);
    return it == kv_.end() || it->second.empty() ? dflt : it->second.string();
  }
  bool has(const std::string& path) const { return kv_.count(path) != 0; }
 private:
  std::map<std::string, value_t> kv_;
};

// one kernel argument: scalars by value, buffers as device pointers (what occa::kernelArg carries)
struct kernelArg_t {
  enum Kind { Int, Double, Pointer } kind;
  long long i = 0;
  double d = 0.0;
  void* p = nullptr;
  kernelArg_t() : kind(Int) {}
  kernelArg_t(int v) : kind(Int), i(v) {}
  kernelArg_t(long long v) : kind(Int), i(v) {}
  kernelArg_t(double v) : kind(Double), d(v) {}
  kernelArg_t(float v) : kind(Double), d(v) {}
  template <typename T> kernelArg_t(const deviceMemory<T>& m) : kind(Pointer), p((void*)m.ptr()) {}
  template <typename T> kernelArg_t(T* ptr) : kind(Pointer), p((void*)ptr) {}
};

class kernel_t {
 public:
  // plain function pointer + a small context word (ogs kernels: element type | reduction op << 8): a call costs no heap
  // allocation and no std::function indirection (BS0 measures exactly this path)
  using impl_t = int (*)(const kernelArg_t* args, spn_stream_t stream, int ctx);
  static constexpr int kMaxArgs = 8;
  kernel_t() = default;
  kernel_t(std::string name, int nargs, impl_t impl, int ctx, std::shared_ptr<spn_stream_t> cur)
      : name_(std::move(name)), nargs_(nargs), impl_(impl), ctx_(ctx), cur_(std::move(cur)) {}
  bool isInitialized() const { return impl_ != nullptr; }
  const std::string& name() const { return name_; }

  template <typename... Args>
  void operator()(Args&&... args) const {
    static_assert(sizeof...(Args) <= (size_t)kMaxArgs, "no kernel of this path takes more than 8 arguments");
    SPN_ABORT("kernel_t: call of an uninitialised kernel", !impl_);
    const kernelArg_t a[kMaxArgs] = {kernelArg_t(std::forward<Args>(args))...};
    SPN_ABORT("kernel " + name_ + ": expected " + std::to_string(nargs_) + " arguments, got " + std::to_string(sizeof...(Args)),
              (int)sizeof...(Args) != nargs_);
    check_status(impl_(a, cur_ ? *cur_ : nullptr, ctx_), name_.c_str());
  }

 private:
  std::string name_;
  int nargs_ = 0;
  impl_t impl_ = nullptr;
  int ctx_ = 0;
  std::shared_ptr<spn_stream_t> cur_;
};

// platform_t::buildKernel (include/platform.hpp:92-93).  Free function so that platform.hpp stays
// free of the kernel table; platform_t::buildKernel forwards here.
kernel_t buildKernel(platform_t& platform, const std::string& fileName, const std::string& kernelName,
                     const properties_t& props);

}  // namespace spn
