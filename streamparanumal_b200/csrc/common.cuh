// Shared device/host helpers for the sm_100a kernels behind include/spn_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "spn_b200.h"

#define SPN_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) { (void)cudaGetLastError(); return (int)_e; } \
  } while (0)

#define SPN_LAUNCH_CHECK()                                   \
  do {                                                       \
    cudaError_t _e = cudaGetLastError();                     \
    if (_e != cudaSuccess) return (int)_e;                   \
  } while (0)

namespace spn {

struct DeviceInfo {
  int device;
  int sms;
  int cc_major;
  int cc_minor;
  void* reduce_ws;  // lazily allocated default reduction workspace
};

// Per-device cached properties (fails with SPN_ERR_NO_DEVICE / SPN_ERR_UNSUPPORTED).
int device_info(DeviceInfo** out);

static inline cudaStream_t as_stream(spn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------------------
// 256-bit / 128-bit streaming accesses.  sm_100 has LDG.E.256 / STG.E.256
// (PTX .v4.f64), which halves the LSU instruction count of a pure streaming
// kernel relative to 128-bit accesses.
// ---------------------------------------------------------------------------
template <int VEC> struct alignas(VEC * 8) dvec { double v[VEC]; };

// read-only for the whole kernel (.nc), no L1 allocation, evict-first in L2
template <int VEC> __device__ __forceinline__ dvec<VEC> ld_ro(const double* p);
template <> __device__ __forceinline__ dvec<4> ld_ro<4>(const double* p) {
  dvec<4> r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p));
  return r;
}
template <> __device__ __forceinline__ dvec<2> ld_ro<2>(const double* p) {
  dvec<2> r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
               : "=d"(r.v[0]), "=d"(r.v[1]) : "l"(p));
  return r;
}
template <> __device__ __forceinline__ dvec<1> ld_ro<1>(const double* p) {
  dvec<1> r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r.v[0]) : "l"(p));
  return r;
}

// read of a location this kernel also writes (no .nc)
template <int VEC> __device__ __forceinline__ dvec<VEC> ld_rw(const double* p);
template <> __device__ __forceinline__ dvec<4> ld_rw<4>(const double* p) {
  dvec<4> r;
  asm volatile("ld.global.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p) : "memory");
  return r;
}
template <> __device__ __forceinline__ dvec<2> ld_rw<2>(const double* p) {
  dvec<2> r;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
               : "=d"(r.v[0]), "=d"(r.v[1]) : "l"(p) : "memory");
  return r;
}
template <> __device__ __forceinline__ dvec<1> ld_rw<1>(const double* p) {
  dvec<1> r;
  asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(r.v[0]) : "l"(p) : "memory");
  return r;
}

template <int VEC> __device__ __forceinline__ void st_stream(double* p, const dvec<VEC>& x);
template <> __device__ __forceinline__ void st_stream<4>(double* p, const dvec<4>& x) {
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "d"(x.v[0]), "d"(x.v[1]), "d"(x.v[2]), "d"(x.v[3]) : "memory");
}
template <> __device__ __forceinline__ void st_stream<2>(double* p, const dvec<2>& x) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};"
               :: "l"(p), "d"(x.v[0]), "d"(x.v[1]) : "memory");
}
template <> __device__ __forceinline__ void st_stream<1>(double* p, const dvec<1>& x) {
  asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" :: "l"(p), "d"(x.v[0]) : "memory");
}

// widest vector (in doubles) usable for all of the given pointers
static inline int common_vec(const void* a, const void* b = nullptr, const void* c = nullptr) {
  uintptr_t m = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                reinterpret_cast<uintptr_t>(c);
  if ((m & 31) == 0) return 4;
  if ((m & 15) == 0) return 2;
  return 1;
}

}  // namespace spn
