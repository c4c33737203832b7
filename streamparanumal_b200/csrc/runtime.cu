// Device runtime entry points of include/spn_b200.h: the slice of the OCCA CUDA
// backend (occa/src/occa/internal/modes/cuda/{device,buffer,memory,stream}.cpp)
// that platform_t / deviceMemory<T> / ogs use on the hot path.
#include <mutex>
#include <stdio.h>
#include <string.h>
#include "common.cuh"

namespace spn {

static constexpr int kMaxDevices = 64;
static DeviceInfo g_info[kMaxDevices];
static bool g_info_ready[kMaxDevices];
static std::mutex g_info_mutex;

int device_info(DeviceInfo** out) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return SPN_ERR_NO_DEVICE; }
  if (dev < 0 || dev >= kMaxDevices) return SPN_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(g_info_mutex);
  if (!g_info_ready[dev]) {
    DeviceInfo& d = g_info[dev];
    d.device = dev;
    SPN_CUDA_TRY(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
    SPN_CUDA_TRY(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    SPN_CUDA_TRY(cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    // The fatbin holds sm_100a SASS only: anything else cannot run these kernels.
    if (d.cc_major != 10) return SPN_ERR_UNSUPPORTED;
    d.reduce_ws = nullptr;
    g_info_ready[dev] = true;
  }
  *out = &g_info[dev];
  return SPN_SUCCESS;
}

}  // namespace spn

using namespace spn;

extern "C" {

int spn_version(void) { return 100; }

const char* spn_error_string(int status) {
  switch (status) {
    case SPN_SUCCESS: return "success";
    case SPN_ERR_INVALID_ARGUMENT: return "spn: invalid argument";
    case SPN_ERR_UNSUPPORTED: return "spn: unsupported device (this library carries sm_100a code only)";
    case SPN_ERR_ROW_TOO_LONG: return "spn: row block exceeds spn_ogs_nodes_per_block()";
    case SPN_ERR_NCCL: return "spn: NCCL error";
    case SPN_ERR_NO_DEVICE: return "spn: no CUDA device available (there is no CPU fallback)";
    case SPN_ERR_OUT_OF_MEMORY: return "spn: out of host or device memory";
    default: break;
  }
  if (status > 0) return cudaGetErrorString((cudaError_t)status);
  return "spn: unknown error";
}

int spn_device_count(int* count) {
  if (!count) return SPN_ERR_INVALID_ARGUMENT;
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) { (void)cudaGetLastError(); *count = 0; return SPN_ERR_NO_DEVICE; }
  return SPN_SUCCESS;
}

int spn_set_device(int device_id) { SPN_CUDA_TRY(cudaSetDevice(device_id)); return SPN_SUCCESS; }
int spn_get_device(int* device_id) {
  if (!device_id) return SPN_ERR_INVALID_ARGUMENT;
  SPN_CUDA_TRY(cudaGetDevice(device_id));
  return SPN_SUCCESS;
}
int spn_device_sm_count(int* sms) {
  if (!sms) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  *sms = d->sms;
  return SPN_SUCCESS;
}
int spn_device_sync(void) { SPN_CUDA_TRY(cudaDeviceSynchronize()); return SPN_SUCCESS; }

int spn_stream_create(spn_stream_t* stream) {
  if (!stream) return SPN_ERR_INVALID_ARGUMENT;
  cudaStream_t s;
  SPN_CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = (spn_stream_t)s;
  return SPN_SUCCESS;
}
int spn_stream_destroy(spn_stream_t stream) {
  if (!stream) return SPN_SUCCESS;
  SPN_CUDA_TRY(cudaStreamDestroy(as_stream(stream)));
  return SPN_SUCCESS;
}
int spn_stream_sync(spn_stream_t stream) { SPN_CUDA_TRY(cudaStreamSynchronize(as_stream(stream))); return SPN_SUCCESS; }

int spn_malloc(void** dptr, size_t bytes) {
  if (!dptr) return SPN_ERR_INVALID_ARGUMENT;
  *dptr = nullptr;
  if (bytes == 0) return SPN_SUCCESS;
  SPN_CUDA_TRY(cudaMalloc(dptr, bytes));
  return SPN_SUCCESS;
}
int spn_free(void* dptr) {
  if (!dptr) return SPN_SUCCESS;
  SPN_CUDA_TRY(cudaFree(dptr));
  return SPN_SUCCESS;
}
int spn_host_malloc(void** hptr, size_t bytes) {
  if (!hptr) return SPN_ERR_INVALID_ARGUMENT;
  *hptr = nullptr;
  if (bytes == 0) return SPN_SUCCESS;
  SPN_CUDA_TRY(cudaMallocHost(hptr, bytes));
  return SPN_SUCCESS;
}
int spn_host_free(void* hptr) {
  if (!hptr) return SPN_SUCCESS;
  SPN_CUDA_TRY(cudaFreeHost(hptr));
  return SPN_SUCCESS;
}

int spn_memset(void* dptr, int value, size_t bytes, spn_stream_t stream) {
  if (bytes == 0) return SPN_SUCCESS;
  if (!dptr) return SPN_ERR_INVALID_ARGUMENT;
  SPN_CUDA_TRY(cudaMemsetAsync(dptr, value, bytes, as_stream(stream)));
  return SPN_SUCCESS;
}

static int do_copy(void* dst, const void* src, size_t bytes, spn_stream_t stream, int async,
                   cudaMemcpyKind kind) {
  if (bytes == 0) return SPN_SUCCESS;
  if (!dst || !src) return SPN_ERR_INVALID_ARGUMENT;
  SPN_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, kind, as_stream(stream)));
  if (!async) SPN_CUDA_TRY(cudaStreamSynchronize(as_stream(stream)));
  return SPN_SUCCESS;
}
int spn_memcpy_h2d(void* dst, const void* src, size_t bytes, spn_stream_t stream, int async) {
  return do_copy(dst, src, bytes, stream, async, cudaMemcpyHostToDevice);
}
int spn_memcpy_d2h(void* dst, const void* src, size_t bytes, spn_stream_t stream, int async) {
  return do_copy(dst, src, bytes, stream, async, cudaMemcpyDeviceToHost);
}
int spn_memcpy_d2d(void* dst, const void* src, size_t bytes, spn_stream_t stream, int async) {
  return do_copy(dst, src, bytes, stream, async, cudaMemcpyDeviceToDevice);
}

int spn_event_create(spn_event_t* ev) {
  if (!ev) return SPN_ERR_INVALID_ARGUMENT;
  cudaEvent_t e;
  SPN_CUDA_TRY(cudaEventCreate(&e));
  *ev = (spn_event_t)e;
  return SPN_SUCCESS;
}
int spn_event_destroy(spn_event_t ev) {
  if (!ev) return SPN_SUCCESS;
  SPN_CUDA_TRY(cudaEventDestroy((cudaEvent_t)ev));
  return SPN_SUCCESS;
}
int spn_event_record(spn_event_t ev, spn_stream_t stream) {
  SPN_CUDA_TRY(cudaEventRecord((cudaEvent_t)ev, as_stream(stream)));
  return SPN_SUCCESS;
}
int spn_event_sync(spn_event_t ev) { SPN_CUDA_TRY(cudaEventSynchronize((cudaEvent_t)ev)); return SPN_SUCCESS; }
int spn_stream_wait_event(spn_stream_t stream, spn_event_t ev) {
  SPN_CUDA_TRY(cudaStreamWaitEvent(as_stream(stream), (cudaEvent_t)ev, 0));
  return SPN_SUCCESS;
}
int spn_event_elapsed_ms(spn_event_t start, spn_event_t stop, float* ms) {
  if (!ms) return SPN_ERR_INVALID_ARGUMENT;
  SPN_CUDA_TRY(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return SPN_SUCCESS;
}

int spn_graph_begin_capture(spn_stream_t stream) {
  SPN_CUDA_TRY(cudaStreamBeginCapture(as_stream(stream), cudaStreamCaptureModeThreadLocal));
  return SPN_SUCCESS;
}
int spn_graph_end_capture(spn_stream_t stream, spn_graph_t* graph) {
  if (!graph) return SPN_ERR_INVALID_ARGUMENT;
  *graph = nullptr;
  cudaGraph_t g = nullptr;
  SPN_CUDA_TRY(cudaStreamEndCapture(as_stream(stream), &g));
  cudaGraphExec_t exec = nullptr;
  const cudaError_t e = cudaGraphInstantiate(&exec, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return (int)e; }
  *graph = reinterpret_cast<spn_graph_t>(exec);
  return SPN_SUCCESS;
}
int spn_graph_launch(spn_graph_t graph, spn_stream_t stream) {
  if (!graph) return SPN_ERR_INVALID_ARGUMENT;
  SPN_CUDA_TRY(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(graph), as_stream(stream)));
  return SPN_SUCCESS;
}
int spn_graph_destroy(spn_graph_t graph) {
  if (graph) SPN_CUDA_TRY(cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(graph)));
  return SPN_SUCCESS;
}

int spn_stream_create_priority(spn_stream_t* stream, int high) {
  if (!stream) return SPN_ERR_INVALID_ARGUMENT;
  int lo = 0, hi = 0;
  SPN_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // numerically lower = higher priority
  cudaStream_t s;
  SPN_CUDA_TRY(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high ? hi : lo));
  *stream = reinterpret_cast<spn_stream_t>(s);
  return SPN_SUCCESS;
}

// ---- CUDA IPC: share a spn_malloc block with the other ranks of the node (peer-memory halo exchange) ----
int spn_ipc_get_handle(void* dptr, void* handle64) {
  if (!dptr || !handle64) return SPN_ERR_INVALID_ARGUMENT;
  static_assert(sizeof(cudaIpcMemHandle_t) == SPN_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  cudaIpcMemHandle_t h;
  SPN_CUDA_TRY(cudaIpcGetMemHandle(&h, dptr));
  memcpy(handle64, &h, sizeof(h));
  return SPN_SUCCESS;
}
int spn_ipc_open_handle(const void* handle64, void** dptr) {
  if (!dptr || !handle64) return SPN_ERR_INVALID_ARGUMENT;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  SPN_CUDA_TRY(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
  return SPN_SUCCESS;
}
int spn_ipc_close_handle(void* dptr) {
  if (!dptr) return SPN_SUCCESS;
  SPN_CUDA_TRY(cudaIpcCloseMemHandle(dptr));
  return SPN_SUCCESS;
}

}  // extern "C"
