#include <algorithm>

#include "platform.hpp"

#include <cstdlib>

namespace spn {

platform_t::platform_t(comm_t comm_, int device_id) : comm(comm_) {
  state_ = std::make_shared<State>();
  int count = 0;
  SPN_CHECK(spn_device_count(&count));
  SPN_ABORT("no CUDA device: streamparanumal_b200 has no CPU fallback", count < 1);
  if (device_id < 0) {
    const char* lr = getenv("LOCAL_RANK");
    device_id = (lr ? atoi(lr) : comm.rank()) % count;
  }
  SPN_CHECK(spn_set_device(device_id));
  state_->device = device_id;
}

int platform_t::smCount() const {
  int sms = 0;
  SPN_CHECK(spn_device_sm_count(&sms));
  return sms;
}

spn_stream_t platform_t::createStream(bool high_priority) {
  spn_stream_t s = nullptr;
  if (high_priority) SPN_CHECK(spn_stream_create_priority(&s, 1));
  else SPN_CHECK(spn_stream_create(&s));
  state_->owned.push_back(s);
  return s;
}

void platform_t::destroyStream(spn_stream_t s) {
  if (!s || !state_) return;
  auto it = std::find(state_->owned.begin(), state_->owned.end(), s);
  if (it == state_->owned.end()) return;
  spn_stream_sync(s);
  spn_stream_destroy(s);
  state_->owned.erase(it);
}

timePoint_t Time() { return std::chrono::high_resolution_clock::now(); }

timePoint_t PlatformTime(platform_t& platform) {
  platform.finish();
  return Time();
}

timePoint_t GlobalPlatformTime(platform_t& platform) {
  platform.finish();
  platform.comm.Barrier();
  return Time();
}

double ElapsedTime(const timePoint_t start, const timePoint_t end) {
  return std::chrono::duration_cast<std::chrono::microseconds>(end - start).count() / 1.0e6;
}

}  // namespace spn
