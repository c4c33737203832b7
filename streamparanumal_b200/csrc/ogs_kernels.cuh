// ogs device kernels: gather, scatter, fused gather+scatter, extract
// (reference: libs/ogs/okl/ogsKernels.okl:32-177).
//
// Layout (identical to the reference's device CSR, libs/ogs/ogsOperator.cpp:597-665):
//   rowStarts[Nrows+1], colIds[nnz]  int32, rows sorted, columns ascending per row
//   blockRowStarts[NrowBlocks+1]     row partition, <= kNodesPerBlock nnz per block
//   v[Ncols*K], gv[Nrows*K]          element vectors, entry (i,k) at k + i*K
//
// One CTA per row block.  Design for sm_100a / HBM3e:
//   * the contiguous index stream (colIds) is read with 128-bit loads, all of a
//     thread's index vectors first, then all of its dependent v[] loads, so each
//     thread keeps up to 12 independent 8-byte gathers in flight;
//   * rowStarts are staged once per block in shared memory (the reference re-reads
//     them 2-3x per row from global);
//   * row sums run out of shared memory in CSR order starting from the op identity,
//     which makes the result bit-identical to the reference Serial kernel;
//   * for the symmetric gather+scatter the column ids stay in registers between the
//     gather and the scatter phase: the CSR is read once, not twice;
//   * blocks up to 2048 nnz (reference: 512) give every thread 8 items.
// These kernels are HBM-bound (4 B index + 8..16 B payload per nnz): no tensor cores.
#pragma once
#include <cfloat>
#include <climits>
#include "common.cuh"

namespace spn {

#ifndef SPN_OGS_NNZB
#define SPN_OGS_NNZB 2048
#endif
static constexpr int kOgsBlock = 256;
static constexpr int kNodesPerBlock = SPN_OGS_NNZB;
// 128-bit index vectors per thread: the block's nnz range is widened down to a
// multiple of 4, so up to kNodesPerBlock/4+1 vectors are needed.
static constexpr int kIdxIter = kNodesPerBlock / (4 * kOgsBlock) + 1;

// ---- op identities and combiners (libs/ogs/ogsOperator.cpp:37-56, ogsUtils.cpp:56-103)
template <typename T> struct TypeMax;
template <> struct TypeMax<float> { __host__ __device__ static constexpr float value() { return FLT_MAX; } };
template <> struct TypeMax<double> { __host__ __device__ static constexpr double value() { return DBL_MAX; } };
template <> struct TypeMax<int> { __host__ __device__ static constexpr int value() { return INT_MAX; } };
template <> struct TypeMax<long long> { __host__ __device__ static constexpr long long value() { return LLONG_MAX; } };

template <typename T, int OP> struct OgsOp;
template <typename T> struct OgsOp<T, SPN_ADD> {
  __device__ static __forceinline__ T init() { return T(0); }
  __device__ static __forceinline__ void apply(T& a, const T b) { a += b; }
};
template <typename T> struct OgsOp<T, SPN_MUL> {
  __device__ static __forceinline__ T init() { return T(1); }
  __device__ static __forceinline__ void apply(T& a, const T b) { a *= b; }
};
template <typename T> struct OgsOp<T, SPN_MAX> {
  __device__ static __forceinline__ T init() { return -TypeMax<T>::value(); }
  __device__ static __forceinline__ void apply(T& a, const T b) { if (b > a) a = b; }
};
template <typename T> struct OgsOp<T, SPN_MIN> {
  __device__ static __forceinline__ T init() { return TypeMax<T>::value(); }
  __device__ static __forceinline__ void apply(T& a, const T b) { if (b < a) a = b; }
};

// ---- block geometry ------------------------------------------------------
struct BlockRange {
  int rb, re;        // rows [rb,re)
  int start, end;    // nnz  [start,end)
  int a0;            // start rounded down to a multiple of 4
};

__device__ __forceinline__ BlockRange block_range(const int* __restrict__ blockStarts,
                                                  const int* __restrict__ rowStarts) {
  BlockRange r;
  r.rb = __ldg(blockStarts + blockIdx.x);
  r.re = __ldg(blockStarts + blockIdx.x + 1);
  r.start = __ldg(rowStarts + r.rb);
  r.end = __ldg(rowStarts + r.re);
  r.a0 = r.start & ~3;
  return r;
}

// all of this thread's column ids for the block: vector j covers nnz
// [a0 + 4*(tid + j*BLOCK), +4)
__device__ __forceinline__ void load_ids(const int* __restrict__ colIds, const BlockRange& r,
                                         int4 (&c)[kIdxIter]) {
  // 128-bit loads need a 16-byte aligned index array (always true for whole allocations)
  const bool vec_ok = (reinterpret_cast<uintptr_t>(colIds) & 15) == 0;
#pragma unroll
  for (int j = 0; j < kIdxIter; ++j) {
    const int p = r.a0 + 4 * ((int)threadIdx.x + j * kOgsBlock);
    if (p + 4 <= r.end && vec_ok) {
      c[j] = __ldg(reinterpret_cast<const int4*>(colIds + p));
    } else if (p + 4 <= r.end) {
      c[j].x = __ldg(colIds + p + 0);
      c[j].y = __ldg(colIds + p + 1);
      c[j].z = __ldg(colIds + p + 2);
      c[j].w = __ldg(colIds + p + 3);
    } else {  // straddles the block end: never read past it
      c[j].x = (p + 0 < r.end) ? __ldg(colIds + p + 0) : 0;
      c[j].y = (p + 1 < r.end) ? __ldg(colIds + p + 1) : 0;
      c[j].z = (p + 2 < r.end) ? __ldg(colIds + p + 2) : 0;
      c[j].w = 0;
    }
  }
}

// stage v[k + col*K] for every nnz of the block into temp[nnz - start]
template <typename T, bool READONLY>
__device__ __forceinline__ void stage_values(const T* v, int K, int k, const BlockRange& r,
                                             const int4 (&c)[kIdxIter], T* temp) {
  T x[kIdxIter][4];
#pragma unroll
  for (int j = 0; j < kIdxIter; ++j) {
    const int p = r.a0 + 4 * ((int)threadIdx.x + j * kOgsBlock);
    const int ids[4] = {c[j].x, c[j].y, c[j].z, c[j].w};
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      if (p + l >= r.start && p + l < r.end) {
        const T* src = v + (k + (int64_t)ids[l] * K);
        x[j][l] = READONLY ? __ldg(src) : *src;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kIdxIter; ++j) {
    const int p = r.a0 + 4 * ((int)threadIdx.x + j * kOgsBlock);
#pragma unroll
    for (int l = 0; l < 4; ++l)
      if (p + l >= r.start && p + l < r.end) temp[p + l - r.start] = x[j][l];
  }
}

// write temp[nnz - start] to v[k + col*K] for every nnz of the block
template <typename T>
__device__ __forceinline__ void unstage_values(T* v, int K, int k, const BlockRange& r,
                                               const int4 (&c)[kIdxIter], const T* temp) {
#pragma unroll
  for (int j = 0; j < kIdxIter; ++j) {
    const int p = r.a0 + 4 * ((int)threadIdx.x + j * kOgsBlock);
    const int ids[4] = {c[j].x, c[j].y, c[j].z, c[j].w};
#pragma unroll
    for (int l = 0; l < 4; ++l)
      if (p + l >= r.start && p + l < r.end) v[k + (int64_t)ids[l] * K] = temp[p + l - r.start];
  }
}

__device__ __forceinline__ void stage_rows(const int* __restrict__ rowStarts, const BlockRange& r, int* srow) {
  const int nrows = r.re - r.rb;
  for (int i = threadIdx.x; i <= nrows; i += kOgsBlock) srow[i] = __ldg(rowStarts + r.rb + i) - r.start;
}

__device__ __forceinline__ bool block_fits(const BlockRange& r) {
  return (r.end - r.start) <= kNodesPerBlock && (r.re - r.rb) <= kNodesPerBlock;
}

// ---------------------------------------------------------------------------
// gather  (ogsKernels.okl:86-122)
// ---------------------------------------------------------------------------
template <typename T, int OP>
__global__ void __launch_bounds__(kOgsBlock)
gather_kernel(int K, const int* __restrict__ blockStarts, const int* __restrict__ rowStarts,
              const int* __restrict__ colIds, const T* __restrict__ v, T* __restrict__ gv) {
  __shared__ T temp[kNodesPerBlock];
  __shared__ int srow[kNodesPerBlock + 1];
  const int k = blockIdx.y;
  const BlockRange r = block_range(blockStarts, rowStarts);

  if (!block_fits(r)) {  // oversized block: correct but unstaged
    for (int row = r.rb + threadIdx.x; row < r.re; row += kOgsBlock) {
      T acc = OgsOp<T, OP>::init();
      for (int g = rowStarts[row]; g < rowStarts[row + 1]; ++g)
        OgsOp<T, OP>::apply(acc, v[k + (int64_t)colIds[g] * K]);
      gv[k + (int64_t)row * K] = acc;
    }
    return;
  }

  int4 c[kIdxIter];
  load_ids(colIds, r, c);
  stage_rows(rowStarts, r, srow);
  stage_values<T, true>(v, K, k, r, c, temp);
  __syncthreads();

  const int nrows = r.re - r.rb;
  for (int i = threadIdx.x; i < nrows; i += kOgsBlock) {
    const int s = srow[i], e = srow[i + 1];
    T acc = OgsOp<T, OP>::init();
    for (int g = s; g < e; ++g) OgsOp<T, OP>::apply(acc, temp[g]);
    gv[k + (int64_t)(r.rb + i) * K] = acc;
  }
}

// ---------------------------------------------------------------------------
// scatter  (ogsKernels.okl:127-164) — U is an unsigned type of the element size
// ---------------------------------------------------------------------------
template <typename U>
__global__ void __launch_bounds__(kOgsBlock)
scatter_kernel(int K, const int* __restrict__ blockStarts, const int* __restrict__ rowStarts,
               const int* __restrict__ colIds, const U* __restrict__ gv, U* __restrict__ v) {
  __shared__ U temp[kNodesPerBlock];
  __shared__ int srow[kNodesPerBlock + 1];
  const int k = blockIdx.y;
  const BlockRange r = block_range(blockStarts, rowStarts);

  if (!block_fits(r)) {
    for (int row = r.rb + threadIdx.x; row < r.re; row += kOgsBlock) {
      const U val = gv[k + (int64_t)row * K];
      for (int g = rowStarts[row]; g < rowStarts[row + 1]; ++g) v[k + (int64_t)colIds[g] * K] = val;
    }
    return;
  }

  int4 c[kIdxIter];
  load_ids(colIds, r, c);
  stage_rows(rowStarts, r, srow);
  __syncthreads();

  const int nrows = r.re - r.rb;
  for (int i = threadIdx.x; i < nrows; i += kOgsBlock) {
    const U val = __ldg(gv + (k + (int64_t)(r.rb + i) * K));
    const int s = srow[i], e = srow[i + 1];
    for (int g = s; g < e; ++g) temp[g] = val;
  }
  __syncthreads();
  unstage_values<U>(v, K, k, r, c, temp);
}

// ---------------------------------------------------------------------------
// gatherScatter, symmetric: gather CSR == scatter CSR  (ogsKernels.okl:32-81 with
// the arguments of libs/ogs/ogsOperator.cpp:460-469)
// ---------------------------------------------------------------------------
template <typename T, int OP>
__global__ void __launch_bounds__(kOgsBlock)
gather_scatter_sym_kernel(int K, const int* __restrict__ blockStarts, const int* __restrict__ rowStarts,
                          const int* __restrict__ colIds, T* v) {
  __shared__ T temp[kNodesPerBlock];
  __shared__ int srow[kNodesPerBlock + 1];
  const int k = blockIdx.y;
  const BlockRange r = block_range(blockStarts, rowStarts);

  if (!block_fits(r)) {
    for (int row = r.rb + threadIdx.x; row < r.re; row += kOgsBlock) {
      const int s = rowStarts[row], e = rowStarts[row + 1];
      T acc = OgsOp<T, OP>::init();
      for (int g = s; g < e; ++g) OgsOp<T, OP>::apply(acc, v[k + (int64_t)colIds[g] * K]);
      for (int g = s; g < e; ++g) v[k + (int64_t)colIds[g] * K] = acc;
    }
    return;
  }

  int4 c[kIdxIter];
  load_ids(colIds, r, c);
  stage_rows(rowStarts, r, srow);
  stage_values<T, false>(v, K, k, r, c, temp);
  __syncthreads();

  const int nrows = r.re - r.rb;
  for (int i = threadIdx.x; i < nrows; i += kOgsBlock) {
    const int s = srow[i], e = srow[i + 1];
    T acc = OgsOp<T, OP>::init();
    for (int g = s; g < e; ++g) OgsOp<T, OP>::apply(acc, temp[g]);
    for (int g = s; g < e; ++g) temp[g] = acc;
  }
  __syncthreads();
  unstage_values<T>(v, K, k, r, c, temp);
}

// ---------------------------------------------------------------------------
// gatherScatter, general: separate gather and scatter CSRs over the same rows
// (Trans / NoTrans on a Signed operator, libs/ogs/ogsOperator.cpp:450-459,470-479)
// ---------------------------------------------------------------------------
template <typename T, int OP>
__global__ void __launch_bounds__(kOgsBlock)
gather_scatter_kernel(int K, const int* __restrict__ blockStarts, const int* __restrict__ gRowStarts,
                      const int* __restrict__ gColIds, const int* __restrict__ sRowStarts,
                      const int* __restrict__ sColIds, T* v) {
  __shared__ T temp[kNodesPerBlock];
  __shared__ int srow[kNodesPerBlock + 1];
  const int k = blockIdx.y;
  const BlockRange r = block_range(blockStarts, gRowStarts);

  if (!block_fits(r)) {
    for (int row = r.rb + threadIdx.x; row < r.re; row += kOgsBlock) {
      T acc = OgsOp<T, OP>::init();
      for (int g = gRowStarts[row]; g < gRowStarts[row + 1]; ++g)
        OgsOp<T, OP>::apply(acc, v[k + (int64_t)gColIds[g] * K]);
      for (int g = sRowStarts[row]; g < sRowStarts[row + 1]; ++g) v[k + (int64_t)sColIds[g] * K] = acc;
    }
    return;
  }

  int4 c[kIdxIter];
  load_ids(gColIds, r, c);
  stage_rows(gRowStarts, r, srow);
  stage_values<T, false>(v, K, k, r, c, temp);
  __syncthreads();

  const int nrows = r.re - r.rb;
  for (int i = threadIdx.x; i < nrows; i += kOgsBlock) {
    const int row = r.rb + i;
    const int ss = __ldg(sRowStarts + row), se = __ldg(sRowStarts + row + 1);
    T acc = OgsOp<T, OP>::init();
    for (int g = srow[i]; g < srow[i + 1]; ++g) OgsOp<T, OP>::apply(acc, temp[g]);
    for (int g = ss; g < se; ++g) v[k + (int64_t)__ldg(sColIds + g) * K] = acc;
  }
}

// ---------------------------------------------------------------------------
// extract  (ogsKernels.okl:167-177)
// ---------------------------------------------------------------------------
template <typename U>
__global__ void __launch_bounds__(kOgsBlock)
extract_kernel(int64_t NK, int K, const int* __restrict__ ids, const U* __restrict__ q, U* __restrict__ out) {
  for (int64_t n = (int64_t)blockIdx.x * kOgsBlock + threadIdx.x; n < NK; n += (int64_t)gridDim.x * kOgsBlock) {
    const int64_t gid = n / K;
    const int k = (int)(n - gid * K);
    out[n] = __ldg(q + (k + (int64_t)__ldg(ids + gid) * K));
  }
}

}  // namespace spn
