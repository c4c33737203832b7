/* occa.hpp — stand-in for the slice of the OCCA API that streamParanumal's first-party code touches
 * (reference: include/utils.hpp:44-47 re-exports occa::json / device / kernel / stream as properties_t / device_t /
 * kernel_t / stream_t; include/memory.hpp:283-690 subclasses occa::memory; include/platform.hpp:95-170;
 * libs/core/platformDeviceConfig.cpp:159; libs/core/platformBuildKernel.cpp:41-47; libs/ogs/ogs.cpp:64-119;
 * every kernel_t call site in BS/BSn/src/bsnRun.cpp, libs/ogs/ogsOperator.cpp and libs/ogs/ogsPairwise.cpp).
 *
 * With `-I include/shim` in front of the reference's own include path, the reference's sources
 * (libs/core, libs/primitives, libs/ogs, libs/mesh, BS/BSn) compile UNCHANGED against this header: buildKernel resolves
 * (kernel name, defines/T, defines/OGS_OP(a,b)) to a precompiled entry point instead of JIT-compiling OKL, and
 * occa::memory / stream / device map onto the backend hooks declared in occa::backend at the end of this file.
 * Two backends implement those hooks:
 *   streamparanumal_b200/shim/occa_spn.cpp   the product: sm_100a kernels through the C ABI of include/spn_b200.h
 *                                            (mode "CUDA" only — there is no CPU fallback);
 *   oracle/shim/occa_host.cpp                test infrastructure: host memory + the reference's OKL kernels
 *                                            compiled for the host (modes "Serial" / "OpenMP").
 * Semantics follow OCCA 1.6 (occa/include/occa/core/{device,memory,kernel,stream}.hpp):
 *   - handles are reference-counted and copyable; a default-constructed handle is "not initialised";
 *   - memory::operator+ / slice take offsets in ELEMENTS of the handle's dtype, size() is bytes, length() elements;
 *   - copyFrom / copyTo are blocking unless props["async"] is true, then they are enqueued on the device's current
 *     stream (occa/src/occa/internal/modes/cuda/memory.cpp:59-139);
 *   - kernel::operator() enqueues on the device's current stream (occa/.../modes/cuda/kernel.cpp:96-100);
 *   - device::finish() synchronises the current stream (occa/src/core/device.cpp:250-254).
 */
#ifndef SPN_SHIM_OCCA_HPP
#define SPN_SHIM_OCCA_HPP

#include <array>
#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace occa {

class exception : public std::runtime_error {
 public:
  explicit exception(const std::string& msg) : std::runtime_error(msg) {}
};

/* ---------------------------------------------------------------------------------------------------------
 * occa::json — path-addressed property tree: operator[]("a/b"), assignment of bool / integers / reals / strings,
 * string `+=`, asObject() / asArray(), json(key, value), copy.  Values are kept as text plus a kind tag.
 * ------------------------------------------------------------------------------------------------------- */
class json {
 public:
  enum kind_t { none_, object_, array_, boolean_, integer_, real_, string_ };

  json() = default;
  template <typename V>
  json(const std::string& key, const V& value) { (*this)[key] = value; }

  json& operator[](const std::string& path) {
    json* node = this;
    size_t start = 0;
    while (start <= path.size()) {
      const size_t slash = path.find('/', start);
      const std::string key = path.substr(start, slash == std::string::npos ? std::string::npos : slash - start);
      if (!key.empty()) {
        if (node->kind_ != object_) { node->kind_ = object_; node->text_.clear(); }
        node = &node->children_[key];
      }
      if (slash == std::string::npos) break;
      start = slash + 1;
    }
    return *node;
  }
  const json& operator[](const std::string& path) const {
    static const json missing;
    const json* node = this;
    size_t start = 0;
    while (start <= path.size()) {
      const size_t slash = path.find('/', start);
      const std::string key = path.substr(start, slash == std::string::npos ? std::string::npos : slash - start);
      if (!key.empty()) {
        auto it = node->children_.find(key);
        if (it == node->children_.end()) return missing;
        node = &it->second;
      }
      if (slash == std::string::npos) break;
      start = slash + 1;
    }
    return *node;
  }

  json& operator=(const bool v) { return set(boolean_, v ? "true" : "false"); }
  json& operator=(const char* v) { return set(string_, v); }
  json& operator=(const std::string& v) { return set(string_, v); }
  template <typename V, typename std::enable_if<std::is_arithmetic<V>::value && !std::is_same<V, bool>::value, int>::type = 0>
  json& operator=(const V v) {
    std::ostringstream ss;
    ss.precision(17);
    ss << v;
    return set(std::is_integral<V>::value ? integer_ : real_, ss.str());
  }
  json& operator+=(const std::string& v) {
    if (kind_ != string_) { kind_ = string_; text_.clear(); children_.clear(); }
    text_ += v;
    return *this;
  }
  json& operator+=(const char* v) { return *this += std::string(v); }

  json& asObject() { if (kind_ != object_) { kind_ = object_; text_.clear(); } return *this; }
  json& asArray() { if (kind_ != array_) { kind_ = array_; text_.clear(); } return *this; }

  bool isInitialized() const { return kind_ != none_; }
  bool has(const std::string& path) const { return (*this)[path].kind_ != none_; }
  kind_t kind() const { return kind_; }
  /* value as text ("" when absent) / as a truth value (absent = dflt) */
  const std::string& string() const { return text_; }
  std::string get(const std::string& path, const std::string& dflt = "") const {
    const json& n = (*this)[path];
    return n.kind_ == none_ || n.kind_ == object_ || n.kind_ == array_ ? dflt : n.text_;
  }
  bool get(const std::string& path, const bool dflt) const {
    const json& n = (*this)[path];
    if (n.kind_ == none_) return dflt;
    return n.text_ == "true" || n.text_ == "1";
  }
  const std::map<std::string, json>& children() const { return children_; }

 private:
  json& set(kind_t k, const std::string& t) { kind_ = k; text_ = t; children_.clear(); return *this; }
  kind_t kind_ = none_;
  std::string text_;
  std::map<std::string, json> children_;
};

/* ---------------------------------------------------------------------------------------------------------
 * occa::dtype — only the element size is needed (memory::setDtype(occa::dtype::get<T>()), include/memory.hpp:293)
 * ------------------------------------------------------------------------------------------------------- */
class dtype_t {
 public:
  dtype_t() = default;
  explicit dtype_t(size_t bytes) : bytes_(bytes) {}
  size_t bytes() const { return bytes_; }
 private:
  size_t bytes_ = 1;
};
namespace dtype {
template <typename T>
inline dtype_t get() { return dtype_t(sizeof(T)); }
}  // namespace dtype

class memory;
struct kernelArg;

/* ---------------------------------------------------------------------------------------------------------
 * backend hooks (one device per process, as in the reference: platformDeviceConfig.cpp picks device_id = local rank)
 * ------------------------------------------------------------------------------------------------------- */
namespace backend {
/* device.setup("{mode: 'CUDA', device_id: 0}"): returns the mode string the backend runs ("CUDA", "Serial", ...);
 * throws occa::exception for a mode it does not implement */
std::string deviceSetup(const std::string& mode, int device_id);
int deviceCount(const std::string& mode);
void* deviceMalloc(size_t bytes);  /* uninitialised, like cuMemAlloc (occa/.../cuda/buffer.cpp:65-67) */
void deviceFree(void* p);
void* hostMalloc(size_t bytes);    /* pinned host memory (props["host"] = true) */
void hostFree(void* p);
enum copyKind { h2d, d2h, d2d, h2h };
void copy(void* dst, const void* src, size_t bytes, copyKind kind, void* stream, bool async);
void* streamCreate();
void streamSync(void* stream);
/* resolves a kernel; the launch function receives the call's arguments in order */
typedef void (*launch_fn)(const void* ctx, const kernelArg* args, int nargs, void* stream);
struct kernelEntry { launch_fn launch = nullptr; const void* ctx = nullptr; int nargs = 0; };
kernelEntry lookupKernel(const std::string& file, const std::string& name, const json& props);
}  // namespace backend

/* ---------------------------------------------------------------------------------------------------------
 * occa::stream
 * ------------------------------------------------------------------------------------------------------- */
class stream {
 public:
  stream() = default;
  bool isInitialized() const { return (bool)impl_; }
  void* handle() const { return impl_ ? impl_->h : nullptr; }
  bool operator==(const stream& o) const { return impl_ == o.impl_; }
 private:
  friend class device;
  struct impl_t { void* h = nullptr; };
  std::shared_ptr<impl_t> impl_;
};

namespace detail {
/* process-wide device state shared by every handle */
struct deviceState {
  bool initialized = false;
  std::string mode;
  int device_id = 0;
  stream current;
};
inline deviceState& state() {
  static deviceState s;
  return s;
}
inline void* currentStream() { return state().current.handle(); }

struct buffer {
  void* ptr = nullptr;
  size_t bytes = 0;
  bool host = false;
  ~buffer() {
    if (ptr) {
      if (host) backend::hostFree(ptr);
      else backend::deviceFree(ptr);
    }
  }
};
}  // namespace detail

/* ---------------------------------------------------------------------------------------------------------
 * occa::memory
 * ------------------------------------------------------------------------------------------------------- */
class memory {
 public:
  memory() = default;

  bool isInitialized() const { return (bool)buf_; }
  void setDtype(const dtype_t& d) { dsize_ = d.bytes() ? d.bytes() : 1; }
  void* ptr() { return buf_ ? static_cast<char*>(buf_->ptr) + offset_ : nullptr; }
  const void* ptr() const { return buf_ ? static_cast<const char*>(buf_->ptr) + offset_ : nullptr; }
  size_t size() const { return buf_ ? bytes_ : 0; }      /* bytes */
  size_t length() const { return size() / dsize_; }      /* elements of the dtype */
  bool isHost() const { return buf_ && buf_->host; }

  /* offsets in elements of the dtype (occa::memory::slice / operator+) */
  memory slice(const std::ptrdiff_t offset, const std::ptrdiff_t count = -1) const {
    if (!buf_) throw exception("occa::memory::slice: memory not initialised");
    const size_t off = (size_t)offset * dsize_;
    const size_t cnt = count < 0 ? (off <= bytes_ ? bytes_ - off : 0) : (size_t)count * dsize_;
    if (offset < 0 || off + cnt > bytes_) throw exception("occa::memory::slice: out of bounds");
    memory m(*this);
    m.offset_ = offset_ + off;
    m.bytes_ = cnt;
    return m;
  }
  memory operator+(const std::ptrdiff_t offset) const { return slice(offset); }

  void free() { buf_.reset(); offset_ = bytes_ = 0; }
  void swap(memory& o) { std::swap(buf_, o.buf_); std::swap(offset_, o.offset_); std::swap(bytes_, o.bytes_); std::swap(dsize_, o.dsize_); }

  /* host pointer -> this */
  void copyFrom(const void* src, const std::ptrdiff_t bytes = -1, const std::ptrdiff_t offset = 0, const json& props = json()) {
    const size_t n = bytes < 0 ? size() - (size_t)offset : (size_t)bytes;
    if (!n) return;
    check(offset, n, "copyFrom");
    backend::copy(static_cast<char*>(ptr()) + offset, src, n, isHost() ? backend::h2h : backend::h2d, detail::currentStream(),
                  props.get("async", false));
  }
  void copyFrom(const void* src, const json& props) { copyFrom(src, -1, 0, props); }
  /* this -> host pointer */
  void copyTo(void* dest, const std::ptrdiff_t bytes = -1, const std::ptrdiff_t offset = 0, const json& props = json()) const {
    const size_t n = bytes < 0 ? size() - (size_t)offset : (size_t)bytes;
    if (!n) return;
    check(offset, n, "copyTo");
    backend::copy(dest, static_cast<const char*>(ptr()) + offset, n, isHost() ? backend::h2h : backend::d2h, detail::currentStream(),
                  props.get("async", false));
  }
  void copyTo(void* dest, const json& props) const { copyTo(dest, -1, 0, props); }
  /* memory -> this */
  void copyFrom(const memory src, const std::ptrdiff_t bytes = -1, const std::ptrdiff_t destOffset = 0,
                const std::ptrdiff_t srcOffset = 0, const json& props = json()) {
    const size_t n = bytes < 0 ? size() - (size_t)destOffset : (size_t)bytes;
    if (!n) return;
    check(destOffset, n, "copyFrom");
    src.check(srcOffset, n, "copyFrom (source)");
    backend::copy(static_cast<char*>(ptr()) + destOffset, static_cast<const char*>(src.ptr()) + srcOffset, n, kindOf(src, *this),
                  detail::currentStream(), props.get("async", false));
  }
  void copyFrom(const memory src, const json& props) { copyFrom(src, -1, 0, 0, props); }
  /* this -> memory */
  void copyTo(memory dest, const std::ptrdiff_t bytes = -1, const std::ptrdiff_t destOffset = 0, const std::ptrdiff_t srcOffset = 0,
              const json& props = json()) const {
    const size_t n = bytes < 0 ? size() - (size_t)srcOffset : (size_t)bytes;
    if (!n) return;
    check(srcOffset, n, "copyTo");
    dest.check(destOffset, n, "copyTo (destination)");
    backend::copy(static_cast<char*>(dest.ptr()) + destOffset, static_cast<const char*>(ptr()) + srcOffset, n, kindOf(*this, dest),
                  detail::currentStream(), props.get("async", false));
  }
  void copyTo(memory dest, const json& props) const { copyTo(dest, -1, 0, 0, props); }

 private:
  friend class device;
  static backend::copyKind kindOf(const memory& src, const memory& dst) {
    if (src.isHost()) return dst.isHost() ? backend::h2h : backend::h2d;
    return dst.isHost() ? backend::d2h : backend::d2d;
  }
  void check(std::ptrdiff_t offset, size_t n, const char* what) const {
    if (!buf_) throw exception(std::string("occa::memory::") + what + ": memory not initialised");
    if (offset < 0 || (size_t)offset + n > bytes_) throw exception(std::string("occa::memory::") + what + ": out of bounds");
  }
  std::shared_ptr<detail::buffer> buf_;
  size_t offset_ = 0, bytes_ = 0, dsize_ = 1;
};

/* ---------------------------------------------------------------------------------------------------------
 * occa::kernel — variadic call: scalars by value, buffers as occa::memory (deviceMemory<T> derives from it)
 * ------------------------------------------------------------------------------------------------------- */
struct kernelArg {
  enum kind_t { i32, i64, f32, f64, pointer } kind;
  union { int i; long long l; float f; double d; void* p; } v;
  kernelArg(const int x) : kind(i32) { v.l = 0; v.i = x; }
  kernelArg(const unsigned int x) : kind(i32) { v.l = 0; v.i = (int)x; }
  kernelArg(const long x) : kind(i64) { v.l = x; }
  kernelArg(const long long x) : kind(i64) { v.l = x; }
  kernelArg(const unsigned long x) : kind(i64) { v.l = (long long)x; }
  kernelArg(const float x) : kind(f32) { v.l = 0; v.f = x; }
  kernelArg(const double x) : kind(f64) { v.d = x; }
  kernelArg(const memory& m) : kind(pointer) { v.p = const_cast<void*>(m.ptr()); }
  kernelArg(const void* ptr) : kind(pointer) { v.p = const_cast<void*>(ptr); }
  /* conversions used by the backends */
  long long asInt() const { return kind == i32 ? v.i : kind == i64 ? v.l : kind == f32 ? (long long)v.f : (long long)v.d; }
  double asDouble() const { return kind == f64 ? v.d : kind == f32 ? (double)v.f : (double)asInt(); }
  void* asPointer() const { return kind == pointer ? v.p : nullptr; }
};

class kernel {
 public:
  kernel() = default;
  bool isInitialized() const { return entry_.launch != nullptr; }
  const std::string& name() const { return name_; }

  template <typename... Args>
  void operator()(Args&&... args) const {
    if (!entry_.launch) throw exception("occa::kernel: call of an uninitialised kernel");
    const kernelArg a[] = {kernelArg(std::forward<Args>(args))...};
    constexpr int n = (int)sizeof...(Args);
    if (entry_.nargs && n != entry_.nargs)
      throw exception("occa::kernel " + name_ + ": expected " + std::to_string(entry_.nargs) + " arguments, got " + std::to_string(n));
    entry_.launch(entry_.ctx, a, n, detail::currentStream());
  }
  void operator()() const {
    if (!entry_.launch) throw exception("occa::kernel: call of an uninitialised kernel");
    entry_.launch(entry_.ctx, nullptr, 0, detail::currentStream());
  }

 private:
  friend class device;
  std::string name_;
  backend::kernelEntry entry_;
};

/* ---------------------------------------------------------------------------------------------------------
 * occa::device
 * ------------------------------------------------------------------------------------------------------- */
class device {
 public:
  device() = default;

  /* "{mode: 'CUDA', device_id: 0}" (libs/core/platformDeviceConfig.cpp:57-100) */
  void setup(const std::string& info) {
    const std::string mode = field(info, "mode");
    const std::string id = field(info, "device_id");
    detail::deviceState& s = detail::state();
    s.device_id = id.empty() ? 0 : std::stoi(id);
    s.mode = backend::deviceSetup(mode, s.device_id);
    s.initialized = true;
    s.current = stream();
    s.current.impl_ = std::make_shared<stream::impl_t>();  /* the default stream: initialised, null handle */
  }
  bool isInitialized() const { return detail::state().initialized; }
  const std::string& mode() const { return detail::state().mode; }

  template <typename T = void>
  memory malloc(const size_t count, const json& props = json()) { return malloc<T>(count, nullptr, props); }
  template <typename T = void>
  memory malloc(const size_t count, const void* src, const json& props = json()) {
    constexpr size_t esize = sizeof(typename std::conditional<std::is_void<T>::value, char, T>::type);
    memory m;
    if (count == 0) return m;  /* a zero-sized allocation stays an uninitialised handle (libs/ogs/ogsPairwise.cpp:389-402 relies on it) */
    const bool host = props.get("host", false);
    m.buf_ = std::make_shared<detail::buffer>();
    m.buf_->bytes = count * esize;
    m.buf_->host = host;
    m.buf_->ptr = host ? backend::hostMalloc(m.buf_->bytes) : backend::deviceMalloc(m.buf_->bytes);
    m.bytes_ = m.buf_->bytes;
    m.dsize_ = esize;
    if (src) backend::copy(m.buf_->ptr, src, m.bytes_, host ? backend::h2h : backend::h2d, detail::currentStream(), false);
    return m;
  }

  kernel buildKernel(const std::string& file, const std::string& name, const json& props) {
    kernel k;
    k.name_ = name;
    k.entry_ = backend::lookupKernel(file, name, props);
    return k;
  }

  void finish() { backend::streamSync(detail::currentStream()); }
  stream getStream() { return detail::state().current; }
  void setStream(const stream& s) { detail::state().current = s; }
  stream createStream(const json& = json()) {
    stream s;
    s.impl_ = std::make_shared<stream::impl_t>();
    s.impl_->h = backend::streamCreate();
    return s;
  }

 private:
  /* value of `key` in a "{key: 'value', key2: 3}" setup string */
  static std::string field(const std::string& info, const std::string& key) {
    size_t p = info.find(key);
    if (p == std::string::npos) return "";
    p = info.find(':', p);
    if (p == std::string::npos) return "";
    ++p;
    while (p < info.size() && (info[p] == ' ' || info[p] == '\'' || info[p] == '"')) ++p;
    size_t e = p;
    while (e < info.size() && info[e] != '\'' && info[e] != '"' && info[e] != ',' && info[e] != '}' && info[e] != ' ') ++e;
    return info.substr(p, e - p);
  }
};

inline int getDeviceCount(const std::string& info) { return backend::deviceCount(info); }

namespace env {
inline void setOccaCacheDir(const std::string&) {}  /* nothing is JIT-compiled: there is no kernel cache */
}  // namespace env

}  // namespace occa

#endif
