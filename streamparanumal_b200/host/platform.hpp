// spn::platform_t / deviceMemory<T> / pinnedMemory<T> — the slice of the reference's platform and
// memory classes the hot path touches (include/platform.hpp:60-176, include/memory.hpp:284-454),
// backed by the C ABI of include/spn_b200.h instead of occa::device / occa::memory.
//   platform.malloc<T>(n[,src])   -> deviceMemory<T>     (include/platform.hpp:95-115)
//   platform.hostMalloc<T>(n)     -> pinnedMemory<T>     (include/platform.hpp:117-140)
//   platform.finish()             -> sync the current stream (occa device.finish())
//   o_a.ptr(), o_a + offset, o_a.length(), copyFrom / copyTo (counts and offsets in ELEMENTS, optional
//   async = enqueue on the platform's current stream), conversions deviceMemory<char> -> deviceMemory<T>.
// Null handles behave like the reference's: operator+ on a null handle returns a null handle,
// zero-count copies are no-ops (include/memory.hpp:319-324,338,372,400,434).
#pragma once
#include <chrono>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "comm.hpp"

namespace spn {

using dfloat = double;     // include/types.h:33-69
using dlong = int;
using hlong = long long int;

class platform_t;
class kernel_t;
class properties_t;

// std::vector whose resize / sized constructor leaves trivially-constructible elements uninitialised: the big index
// arrays of the ogs setup are written in full by parallel loops right after they are sized, and the serial value-
// initialisation of std::vector (one thread first-touching every page, 0.5 s per GB) was a third of the setup time.
template <typename T>
struct NoInitAlloc : std::allocator<T> {
  template <typename U>
  struct rebind { using other = NoInitAlloc<U>; };
  NoInitAlloc() = default;
  template <typename U>
  NoInitAlloc(const NoInitAlloc<U>&) {}
  template <typename U, typename... Args>
  void construct(U* p, Args&&... args) {
    if constexpr (sizeof...(Args) == 0) ::new ((void*)p) U;
    else ::new ((void*)p) U(std::forward<Args>(args)...);
  }
};
template <typename T>
using hvec = std::vector<T, NoInitAlloc<T>>;

namespace detail {
struct DeviceBlock {
  void* base = nullptr;
  size_t bytes = 0;
  bool pinned_host = false;
  ~DeviceBlock() {
    if (!base) return;
    if (pinned_host) spn_host_free(base); else spn_free(base);
  }
};
}  // namespace detail

template <typename T>
class deviceMemory {
 public:
  deviceMemory() = default;
  deviceMemory(std::shared_ptr<detail::DeviceBlock> blk, size_t offset_bytes, size_t length, std::shared_ptr<spn_stream_t> cur)
      : blk_(std::move(blk)), off_(offset_bytes), len_(length), cur_(std::move(cur)) {}
  // reinterpret, like deviceMemory<char> -> deviceMemory<T> (include/memory.hpp:296-303)
  template <typename U>
  deviceMemory(const deviceMemory<U>& o) : blk_(o.block()), off_(o.offsetBytes()), len_(o.size() / sizeof(T)), cur_(o.streamSlot()) {}

  bool isInitialized() const { return (bool)blk_; }
  T* ptr() const { return blk_ ? reinterpret_cast<T*>(static_cast<char*>(blk_->base) + off_) : nullptr; }
  size_t length() const { return len_; }
  size_t size() const { return len_ * sizeof(T); }  // bytes, as occa::memory::size()
  void free() { blk_.reset(); off_ = 0; len_ = 0; }

  deviceMemory<T> operator+(const std::ptrdiff_t offset) const {
    if (!blk_) return deviceMemory<T>();
    SPN_ABORT("deviceMemory offset out of range", offset < 0 || (size_t)offset > len_);
    return deviceMemory<T>(blk_, off_ + sizeof(T) * (size_t)offset, len_ - (size_t)offset, cur_);
  }
  deviceMemory<T>& operator+=(const std::ptrdiff_t offset) { *this = *this + offset; return *this; }

  void copyFrom(const T* src, std::ptrdiff_t count = -1, std::ptrdiff_t offset = 0, bool async = false) {
    const size_t n = count < 0 ? len_ : (size_t)count;
    if (n == 0) return;
    SPN_ABORT("copyFrom out of range", (size_t)offset + n > len_);
    SPN_CHECK(spn_memcpy_h2d(ptr() + offset, src, n * sizeof(T), stream(), async ? 1 : 0));
  }
  void copyFrom(const deviceMemory<T>& src, std::ptrdiff_t count = -1, std::ptrdiff_t offset = 0, bool async = false) {
    const size_t n = count < 0 ? len_ : (size_t)count;
    if (n == 0) return;
    SPN_ABORT("copyFrom out of range", (size_t)offset + n > len_ || n > src.length());
    SPN_CHECK(spn_memcpy_d2d(ptr() + offset, src.ptr(), n * sizeof(T), stream(), async ? 1 : 0));
  }
  void copyTo(T* dst, std::ptrdiff_t count = -1, std::ptrdiff_t offset = 0, bool async = false) const {
    const size_t n = count < 0 ? len_ : (size_t)count;
    if (n == 0) return;
    SPN_ABORT("copyTo out of range", (size_t)offset + n > len_);
    SPN_CHECK(spn_memcpy_d2h(dst, ptr() + offset, n * sizeof(T), stream(), async ? 1 : 0));
  }
  void copyTo(deviceMemory<T> dst, std::ptrdiff_t count = -1, std::ptrdiff_t offset = 0, bool async = false) const {
    const size_t n = count < 0 ? len_ : (size_t)count;
    if (n == 0) return;
    SPN_ABORT("copyTo out of range", (size_t)offset + n > len_ || n > dst.length());
    SPN_CHECK(spn_memcpy_d2d(dst.ptr(), ptr() + offset, n * sizeof(T), stream(), async ? 1 : 0));
  }

  const std::shared_ptr<detail::DeviceBlock>& block() const { return blk_; }
  size_t offsetBytes() const { return off_; }
  const std::shared_ptr<spn_stream_t>& streamSlot() const { return cur_; }

 private:
  spn_stream_t stream() const { return cur_ ? *cur_ : nullptr; }
  std::shared_ptr<detail::DeviceBlock> blk_;
  size_t off_ = 0, len_ = 0;
  std::shared_ptr<spn_stream_t> cur_;  // the owning platform's current-stream slot
};

// pinned host memory (platform.hostMalloc): a host pointer with the same handle semantics
template <typename T>
class pinnedMemory {
 public:
  pinnedMemory() = default;
  pinnedMemory(std::shared_ptr<detail::DeviceBlock> blk, size_t length) : blk_(std::move(blk)), len_(length) {}
  bool isInitialized() const { return (bool)blk_; }
  T* ptr() const { return blk_ ? static_cast<T*>(blk_->base) : nullptr; }
  T& operator[](size_t i) const { return ptr()[i]; }
  size_t length() const { return len_; }
  size_t size() const { return len_ * sizeof(T); }
 private:
  std::shared_ptr<detail::DeviceBlock> blk_;
  size_t len_ = 0;
};

using timePoint_t = std::chrono::time_point<std::chrono::high_resolution_clock>;

class platform_t {
 public:
  platform_t() = default;
  // device_id < 0: local rank (libs/core/platformDeviceConfig.cpp:84-96: device_id = localRank)
  explicit platform_t(comm_t comm, int device_id = -1);

  comm_t comm;
  int rank() const { return comm.rank(); }
  int size() const { return comm.size(); }
  int deviceId() const { return state_ ? state_->device : -1; }
  int smCount() const;

  template <typename T>
  deviceMemory<T> malloc(size_t count) {
    auto blk = std::make_shared<detail::DeviceBlock>();
    if (count) SPN_CHECK(spn_malloc(&blk->base, count * sizeof(T)));
    blk->bytes = count * sizeof(T);
    if (!count) return deviceMemory<T>();
    return deviceMemory<T>(blk, 0, count, state_->current);
  }
  template <typename T>
  deviceMemory<T> malloc(size_t count, const T* src) {
    deviceMemory<T> m = malloc<T>(count);
    m.copyFrom(src, (std::ptrdiff_t)count);
    return m;
  }
  template <typename T, typename A>
  deviceMemory<T> malloc(const std::vector<T, A>& src) { return malloc<T>(src.size(), src.data()); }
  template <typename T>
  pinnedMemory<T> hostMalloc(size_t count) {
    auto blk = std::make_shared<detail::DeviceBlock>();
    blk->pinned_host = true;
    if (count) SPN_CHECK(spn_host_malloc(&blk->base, count * sizeof(T)));
    blk->bytes = count * sizeof(T);
    if (!count) return pinnedMemory<T>();
    return pinnedMemory<T>(blk, count);
  }

  // stream handling: OCCA keeps one *current* stream per device (device.getStream/setStream),
  // kernels and async copies go to it (occa/src/occa/internal/modes/cuda/kernel.cpp:96-100)
  spn_stream_t getStream() const { return *state_->current; }
  void setStream(spn_stream_t s) { *state_->current = s; }
  spn_stream_t createStream(bool high_priority = false);
  void destroyStream(spn_stream_t s);  // a stream made by createStream (waits for its work); null is a no-op
  void finish() { SPN_CHECK(spn_stream_sync(*state_->current)); }  // device.finish()
  const std::shared_ptr<spn_stream_t>& streamSlot() const { return state_->current; }

  // platform.buildKernel(file, name, props) (include/platform.hpp:92-93): resolves the kernel name and
  // its defines to a precompiled sm_100a entry point (host/kernel.cpp); "CUDA" is the only mode
  kernel_t buildKernel(const std::string& fileName, const std::string& kernelName, const properties_t& props);
  std::string mode() const { return "CUDA"; }  // platform.device.mode()

 private:
  struct State {
    int device = 0;
    std::shared_ptr<spn_stream_t> current = std::make_shared<spn_stream_t>(nullptr);  // *current == NULL: legacy default stream
    std::vector<spn_stream_t> owned;
    ~State() { for (auto s : owned) spn_stream_destroy(s); }
  };
  std::shared_ptr<State> state_;
};

// libs/core/timer.cpp:32-65
timePoint_t Time();
timePoint_t PlatformTime(platform_t& platform);                 // device.finish(), then the host clock
timePoint_t GlobalPlatformTime(platform_t& platform);           // + barrier over the platform's comm
double ElapsedTime(const timePoint_t start, const timePoint_t end);  // seconds; whole microseconds like the reference (:64)

}  // namespace spn
