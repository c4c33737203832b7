// BS0–BS5 device kernels (templates; instantiated by stream_abi.cu and by
// tools/tune_stream.cu).  Every kernel is HBM-bound: one pass over its arrays
// with 256-bit (LDG.E.256 / STG.E.256) accesses, UNROLL independent accesses in
// flight per thread, no L1 allocation.  Nothing here is a contraction, so no
// tensor cores.
//
// Arithmetic follows the reference *Serial* build (OCCA Serial mode compiles the
// OKL with "-O3", libs/core/platformProperties.cpp:64-68, i.e. no FMA
// contraction): products and sums that reach memory are rounded separately
// (__dmul_rn/__dadd_rn), so copy/axpby/r are bit-identical to the reference for
// any alpha/beta.  Reduction results differ only by summation order (1e-12 rel).
#pragma once
#include "common.cuh"

namespace spn {

// ---------------------------------------------------------------------------
// reduction workspace: [counter | pad to 128 B | partials[kMaxPartials]]
// ---------------------------------------------------------------------------
static constexpr int kMaxPartials = 65536;
struct ReduceWs {
  unsigned int counter;
  unsigned int pad[31];
  double partials[kMaxPartials];
};

template <int BLOCK>
__device__ __forceinline__ double block_reduce_sum(double x, double* red /* [32] shared */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = x;
  __syncthreads();
  if (warp == 0) {
    x = (lane < BLOCK / 32) ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
  }
  return x;  // valid on thread 0
}

// Second pass without a second launch: every CTA publishes its partial, the CTA
// that arrives last folds all partials in a fixed order (deterministic for a
// given grid) and re-arms the counter.  Replaces bs3_2/bs4_2/bs5_2
// (BS/BS3/okl/bs3.okl:66-102).
template <int BLOCK>
__device__ __forceinline__ void grid_finish(double blockSum, ReduceWs* ws, double* out, int zero_tail,
                                            double* red) {
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    ws->partials[blockIdx.x] = blockSum;
    __threadfence();
    const unsigned int t = atomicAdd(&ws->counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double s = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += BLOCK)
    s += *reinterpret_cast<volatile double*>(&ws->partials[i]);
  __syncthreads();  // red[] is reused
  s = block_reduce_sum<BLOCK>(s, red);
  if (threadIdx.x == 0) {
    out[0] = s;
    ws->counter = 0u;
  }
  for (int i = 1 + threadIdx.x; i < zero_tail; i += BLOCK) out[i] = 0.0;
}

// ---------------------------------------------------------------------------
// BS0: empty launch (BS/BS0/okl/bs0.okl:27-39)
// ---------------------------------------------------------------------------
__global__ void bs0_kernel(int N, const double* __restrict__ a, double* __restrict__ b) {
  if ((int)threadIdx.x == N && (int)blockIdx.x == N) b[0] = 0.0;  // never true for N>0; keeps the store alive
}

// ---------------------------------------------------------------------------
// BS1: b = a (BS/BS1/okl/bs1.okl:27-34)
// ---------------------------------------------------------------------------
template <int VEC, int UNROLL, int BLOCK>
__global__ void __launch_bounds__(BLOCK) copy_kernel(int64_t n, const double* __restrict__ a,
                                                      double* __restrict__ b) {
  const int64_t nvec = n / VEC;
  constexpr int64_t chunk = (int64_t)BLOCK * UNROLL;
  for (int64_t base = (int64_t)blockIdx.x * chunk; base < nvec; base += (int64_t)gridDim.x * chunk) {
    dvec<VEC> x[UNROLL];
    if (base + chunk <= nvec) {
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) x[j] = ld_ro<VEC>(a + (base + j * BLOCK + threadIdx.x) * VEC);
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) st_stream<VEC>(b + (base + j * BLOCK + threadIdx.x) * VEC, x[j]);
    } else {
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) {
        const int64_t i = base + j * BLOCK + threadIdx.x;
        if (i < nvec) x[j] = ld_ro<VEC>(a + i * VEC);
      }
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) {
        const int64_t i = base + j * BLOCK + threadIdx.x;
        if (i < nvec) st_stream<VEC>(b + i * VEC, x[j]);
      }
    }
  }
  if (VEC > 1 && blockIdx.x == 0) {
    const int64_t i = nvec * VEC + threadIdx.x;
    if (i < n) b[i] = a[i];
  }
}

// ---------------------------------------------------------------------------
// BS2: y = alpha*x + beta*y (BS/BS2/okl/bs2.okl:27-36), unfused
// ---------------------------------------------------------------------------
__device__ __forceinline__ double axpby1(double alpha, double x, double beta, double y) {
  return __dadd_rn(__dmul_rn(alpha, x), __dmul_rn(beta, y));
}

template <int VEC, int UNROLL, int BLOCK>
__global__ void __launch_bounds__(BLOCK) axpby_kernel(int64_t n, double alpha, const double* __restrict__ x,
                                                       double beta, double* y) {
  const int64_t nvec = n / VEC;
  constexpr int64_t chunk = (int64_t)BLOCK * UNROLL;
  for (int64_t base = (int64_t)blockIdx.x * chunk; base < nvec; base += (int64_t)gridDim.x * chunk) {
    dvec<VEC> xv[UNROLL], yv[UNROLL];
    const bool full = (base + chunk <= nvec);
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const int64_t i = base + j * BLOCK + threadIdx.x;
      if (full || i < nvec) {
        xv[j] = ld_ro<VEC>(x + i * VEC);
        yv[j] = ld_rw<VEC>(y + i * VEC);
      }
    }
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const int64_t i = base + j * BLOCK + threadIdx.x;
      if (full || i < nvec) {
#pragma unroll
        for (int l = 0; l < VEC; ++l) yv[j].v[l] = axpby1(alpha, xv[j].v[l], beta, yv[j].v[l]);
        st_stream<VEC>(y + i * VEC, yv[j]);
      }
    }
  }
  if (VEC > 1 && blockIdx.x == 0) {
    const int64_t i = nvec * VEC + threadIdx.x;
    if (i < n) y[i] = axpby1(alpha, x[i], beta, y[i]);
  }
}

// ---------------------------------------------------------------------------
// BS3: sum a^2   (BS/BS3/okl/bs3.okl:27-64)
// BS4: sum a*b   (BS/BS4/okl/bs4.okl:27-65)
// MODE 0 = norm2 (b unused), 1 = dot
// ---------------------------------------------------------------------------
template <int MODE, int VEC, int UNROLL, int BLOCK>
__global__ void __launch_bounds__(BLOCK) dot_kernel(int64_t n, const double* __restrict__ a,
                                                     const double* __restrict__ b, ReduceWs* ws,
                                                     double* out, int zero_tail) {
  __shared__ double red[32];
  const int64_t nvec = n / VEC;
  constexpr int64_t chunk = (int64_t)BLOCK * UNROLL;
  double acc[VEC];
#pragma unroll
  for (int l = 0; l < VEC; ++l) acc[l] = 0.0;

  for (int64_t base = (int64_t)blockIdx.x * chunk; base < nvec; base += (int64_t)gridDim.x * chunk) {
    dvec<VEC> av[UNROLL], bv[UNROLL];
    const bool full = (base + chunk <= nvec);
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const int64_t i = base + j * BLOCK + threadIdx.x;
      if (full || i < nvec) {
        av[j] = ld_ro<VEC>(a + i * VEC);
        if (MODE == 1) bv[j] = ld_ro<VEC>(b + i * VEC);
      } else {
#pragma unroll
        for (int l = 0; l < VEC; ++l) { av[j].v[l] = 0.0; if (MODE == 1) bv[j].v[l] = 0.0; }
      }
    }
#pragma unroll
    for (int j = 0; j < UNROLL; ++j)
#pragma unroll
      for (int l = 0; l < VEC; ++l)
        acc[l] = fma(av[j].v[l], (MODE == 1) ? bv[j].v[l] : av[j].v[l], acc[l]);
  }
  double s = acc[0];
#pragma unroll
  for (int l = 1; l < VEC; ++l) s += acc[l];
  if (VEC > 1 && blockIdx.x == 0) {
    const int64_t i = nvec * VEC + threadIdx.x;
    if (i < n) s = fma(a[i], (MODE == 1) ? b[i] : a[i], s);
  }
  s = block_reduce_sum<BLOCK>(s, red);
  grid_finish<BLOCK>(s, ws, out, zero_tail, red);
}

// ---------------------------------------------------------------------------
// BS5: r += alpha*Ap ; rdotr = sum r_new^2   (BS/BS5/okl/bs5.okl:27-67)
// ---------------------------------------------------------------------------
template <int VEC, int UNROLL, int BLOCK>
__global__ void __launch_bounds__(BLOCK) cg_update_kernel(int64_t n, const double* __restrict__ Ap,
                                                           double alpha, double* r, ReduceWs* ws,
                                                           double* out, int zero_tail) {
  __shared__ double red[32];
  const int64_t nvec = n / VEC;
  constexpr int64_t chunk = (int64_t)BLOCK * UNROLL;
  double acc[VEC];
#pragma unroll
  for (int l = 0; l < VEC; ++l) acc[l] = 0.0;

  for (int64_t base = (int64_t)blockIdx.x * chunk; base < nvec; base += (int64_t)gridDim.x * chunk) {
    dvec<VEC> pv[UNROLL], rv[UNROLL];
    const bool full = (base + chunk <= nvec);
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const int64_t i = base + j * BLOCK + threadIdx.x;
      if (full || i < nvec) {
        pv[j] = ld_ro<VEC>(Ap + i * VEC);
        rv[j] = ld_rw<VEC>(r + i * VEC);
      }
    }
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const int64_t i = base + j * BLOCK + threadIdx.x;
      if (full || i < nvec) {
#pragma unroll
        for (int l = 0; l < VEC; ++l) {
          const double rn = __dadd_rn(rv[j].v[l], __dmul_rn(alpha, pv[j].v[l]));
          rv[j].v[l] = rn;
          acc[l] = fma(rn, rn, acc[l]);
        }
        st_stream<VEC>(r + i * VEC, rv[j]);
      }
    }
  }
  double s = acc[0];
#pragma unroll
  for (int l = 1; l < VEC; ++l) s += acc[l];
  if (VEC > 1 && blockIdx.x == 0) {
    const int64_t i = nvec * VEC + threadIdx.x;
    if (i < n) {
      const double rn = __dadd_rn(r[i], __dmul_rn(alpha, Ap[i]));
      r[i] = rn;
      s = fma(rn, rn, s);
    }
  }
  s = block_reduce_sum<BLOCK>(s, red);
  grid_finish<BLOCK>(s, ws, out, zero_tail, red);
}

// Reference second pass (bs3_2/bs4_2/bs5_2): out[0] = sum tmp[0..Nblocks)
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) final_sum_kernel(int Nblocks, const double* __restrict__ tmp,
                                                           double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < Nblocks; i += BLOCK) s += tmp[i];
  s = block_reduce_sum<BLOCK>(s, red);
  if (threadIdx.x == 0) out[0] = s;
}

}  // namespace spn
