// C-ABI entry points for BS0–BS5 (include/spn_b200.h).  Launch geometry lives
// here; the kernels are in stream_kernels.cuh.
#include <algorithm>
#include <mutex>
#include "stream_kernels.cuh"

#ifndef SPN_EW_BLOCK
#define SPN_EW_BLOCK 256
#endif
#ifndef SPN_EW_UNROLL
#define SPN_EW_UNROLL 4
#endif
#ifndef SPN_RED_BLOCK
#define SPN_RED_BLOCK 512
#endif
#ifndef SPN_RED_UNROLL
#define SPN_RED_UNROLL 4
#endif
#ifndef SPN_RED_CTAS_PER_SM
#define SPN_RED_CTAS_PER_SM 4
#endif
// BS5 streams like axpby (read Ap, read+write r): the sweep in profiles/round1_notes.md has it
// fastest as 128-bit accesses, 8 in flight per thread, 256-thread CTAs and a grid that is NOT
// capped to the SM count (one CTA per 4096 entries, partials folded by the last CTA)
#ifndef SPN_CG_BLOCK
#define SPN_CG_BLOCK 256
#endif
#ifndef SPN_CG_UNROLL
#define SPN_CG_UNROLL 8
#endif
#ifndef SPN_CG_MAXVEC
#define SPN_CG_MAXVEC 2
#endif

using namespace spn;

namespace {

std::mutex g_ws_mutex;

int default_workspace(DeviceInfo* d, ReduceWs** ws) {
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  if (!d->reduce_ws) {
    void* p = nullptr;
    SPN_CUDA_TRY(cudaMalloc(&p, sizeof(ReduceWs)));
    SPN_CUDA_TRY(cudaMemset(p, 0, sizeof(ReduceWs)));
    d->reduce_ws = p;
  }
  *ws = static_cast<ReduceWs*>(d->reduce_ws);
  return SPN_SUCCESS;
}

// A caller-provided workspace must have been zero-filled once (spn_memset) before its
// first use; every launch leaves the arrival counter re-armed at zero.
int resolve_workspace(DeviceInfo* d, void* user, ReduceWs** ws) {
  if (!user) return default_workspace(d, ws);
  if (reinterpret_cast<uintptr_t>(user) & 7) return SPN_ERR_INVALID_ARGUMENT;
  *ws = static_cast<ReduceWs*>(user);
  return SPN_SUCCESS;
}

inline int ew_grid(int64_t n, int vec) {
  const int64_t nvec = n / vec;
  const int64_t chunk = (int64_t)SPN_EW_BLOCK * SPN_EW_UNROLL;
  int64_t g = (nvec + chunk - 1) / chunk;
  if (g < 1) g = 1;
  if (g > 0x7fffffff) g = 0x7fffffff;  // kernels grid-stride, so any cap is correct
  return (int)g;
}

inline int red_grid(const DeviceInfo* d, int64_t n, int vec) {
  const int64_t nvec = n / vec;
  const int64_t chunk = (int64_t)SPN_RED_BLOCK * SPN_RED_UNROLL;
  int64_t g = (nvec + chunk - 1) / chunk;
  const int64_t cap = std::min<int64_t>((int64_t)d->sms * SPN_RED_CTAS_PER_SM, kMaxPartials);
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

template <int MODE>
int launch_dot(int64_t n, const double* a, const double* b, double* out, int zero_tail, void* workspace,
               spn_stream_t stream) {
  if (n < 0 || !out || (n > 0 && (!a || (MODE == 1 && !b)))) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  ReduceWs* ws;
  rc = resolve_workspace(d, workspace, &ws);
  if (rc) return rc;
  const int vec = common_vec(a, MODE == 1 ? b : nullptr);
  const int grid = red_grid(d, n, vec);
  switch (vec) {
    case 4: dot_kernel<MODE, 4, SPN_RED_UNROLL, SPN_RED_BLOCK><<<grid, SPN_RED_BLOCK, 0, st>>>(n, a, b, ws, out, zero_tail); break;
    case 2: dot_kernel<MODE, 2, SPN_RED_UNROLL, SPN_RED_BLOCK><<<grid, SPN_RED_BLOCK, 0, st>>>(n, a, b, ws, out, zero_tail); break;
    default: dot_kernel<MODE, 1, SPN_RED_UNROLL, SPN_RED_BLOCK><<<grid, SPN_RED_BLOCK, 0, st>>>(n, a, b, ws, out, zero_tail); break;
  }
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

int launch_cg(int64_t n, const double* Ap, double alpha, double* r, double* out, int zero_tail,
              void* workspace, spn_stream_t stream) {
  if (n < 0 || !out || (n > 0 && (!Ap || !r))) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  ReduceWs* ws;
  rc = resolve_workspace(d, workspace, &ws);
  if (rc) return rc;
  const int vec = std::min(common_vec(Ap, r), SPN_CG_MAXVEC);
  int64_t g = (n / vec + (int64_t)SPN_CG_BLOCK * SPN_CG_UNROLL - 1) / ((int64_t)SPN_CG_BLOCK * SPN_CG_UNROLL);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(g, kMaxPartials));
  switch (vec) {
    case 4: cg_update_kernel<4, SPN_CG_UNROLL, SPN_CG_BLOCK><<<grid, SPN_CG_BLOCK, 0, st>>>(n, Ap, alpha, r, ws, out, zero_tail); break;
    case 2: cg_update_kernel<2, SPN_CG_UNROLL, SPN_CG_BLOCK><<<grid, SPN_CG_BLOCK, 0, st>>>(n, Ap, alpha, r, ws, out, zero_tail); break;
    default: cg_update_kernel<1, SPN_CG_UNROLL, SPN_CG_BLOCK><<<grid, SPN_CG_BLOCK, 0, st>>>(n, Ap, alpha, r, ws, out, zero_tail); break;
  }
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

int launch_final(int Nblocks, const double* tmp, double* out, spn_stream_t stream) {
  if (Nblocks < 0 || !out || (Nblocks > 0 && !tmp)) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  final_sum_kernel<256><<<1, 256, 0, as_stream(stream)>>>(Nblocks, tmp, out);
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

}  // namespace

extern "C" {

size_t spn_reduce_workspace_bytes(void) { return sizeof(ReduceWs); }

int spn_bs0(int N, const double* a, double* b, spn_stream_t stream) {
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  bs0_kernel<<<1, 1, 0, as_stream(stream)>>>(N, a, b);
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

int spn_copy_f64(int64_t n, const double* a, double* b, spn_stream_t stream) {
  if (n < 0) return SPN_ERR_INVALID_ARGUMENT;
  if (n == 0) return SPN_SUCCESS;
  if (!a || !b) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  const int vec = common_vec(a, b);
  const int grid = ew_grid(n, vec);
  switch (vec) {
    case 4: copy_kernel<4, SPN_EW_UNROLL, SPN_EW_BLOCK><<<grid, SPN_EW_BLOCK, 0, st>>>(n, a, b); break;
    case 2: copy_kernel<2, SPN_EW_UNROLL, SPN_EW_BLOCK><<<grid, SPN_EW_BLOCK, 0, st>>>(n, a, b); break;
    default: copy_kernel<1, SPN_EW_UNROLL, SPN_EW_BLOCK><<<grid, SPN_EW_BLOCK, 0, st>>>(n, a, b); break;
  }
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

int spn_axpby_f64(int64_t n, double alpha, const double* x, double beta, double* y, spn_stream_t stream) {
  if (n < 0) return SPN_ERR_INVALID_ARGUMENT;
  if (n == 0) return SPN_SUCCESS;
  if (!x || !y) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  const int vec = common_vec(x, y);
  const int grid = ew_grid(n, vec);
  switch (vec) {
    case 4: axpby_kernel<4, SPN_EW_UNROLL, SPN_EW_BLOCK><<<grid, SPN_EW_BLOCK, 0, st>>>(n, alpha, x, beta, y); break;
    case 2: axpby_kernel<2, SPN_EW_UNROLL, SPN_EW_BLOCK><<<grid, SPN_EW_BLOCK, 0, st>>>(n, alpha, x, beta, y); break;
    default: axpby_kernel<1, SPN_EW_UNROLL, SPN_EW_BLOCK><<<grid, SPN_EW_BLOCK, 0, st>>>(n, alpha, x, beta, y); break;
  }
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

int spn_norm2_f64(int64_t n, const double* a, double* out, void* workspace, spn_stream_t stream) {
  return launch_dot<0>(n, a, nullptr, out, 0, workspace, stream);
}
int spn_dot_f64(int64_t n, const double* a, const double* b, double* out, void* workspace, spn_stream_t stream) {
  return launch_dot<1>(n, a, b, out, 0, workspace, stream);
}
int spn_cg_update_f64(int64_t n, const double* Ap, double alpha, double* r, double* rdotr, void* workspace,
                      spn_stream_t stream) {
  return launch_cg(n, Ap, alpha, r, rdotr, 0, workspace, stream);
}

int spn_bs3_1(int Nblocks, int N, const double* a, double* tmp, spn_stream_t stream) {
  if (Nblocks < 1) return SPN_ERR_INVALID_ARGUMENT;
  return launch_dot<0>(N, a, nullptr, tmp, Nblocks, nullptr, stream);
}
int spn_bs3_2(int Nblocks, const double* tmp, double* out, spn_stream_t stream) {
  return launch_final(Nblocks, tmp, out, stream);
}
int spn_bs4_1(int Nblocks, int N, const double* a, const double* b, double* tmp, spn_stream_t stream) {
  if (Nblocks < 1) return SPN_ERR_INVALID_ARGUMENT;
  return launch_dot<1>(N, a, b, tmp, Nblocks, nullptr, stream);
}
int spn_bs4_2(int Nblocks, const double* tmp, double* out, spn_stream_t stream) {
  return launch_final(Nblocks, tmp, out, stream);
}
int spn_bs5_1(int Nblocks, int N, const double* Ap, double alpha, double* r, double* tmp, spn_stream_t stream) {
  if (Nblocks < 1) return SPN_ERR_INVALID_ARGUMENT;
  return launch_cg(N, Ap, alpha, r, tmp, Nblocks, nullptr, stream);
}
int spn_bs5_2(int Nblocks, const double* tmp, double* out, spn_stream_t stream) {
  return launch_final(Nblocks, tmp, out, stream);
}

}  // extern "C"
