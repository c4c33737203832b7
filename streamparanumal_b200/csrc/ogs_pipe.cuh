// Software-pipelined ogs kernels for sm_100a: gather, scatter and symmetric
// gather+scatter (reference: libs/ogs/okl/ogsKernels.okl:32-164).
//
// The reference kernel is one CTA per row block with three dependent global round
// trips (block bounds -> column ids -> values) and nothing else in flight, which
// leaves HBM3e idle most of the time.  Here a small persistent grid walks the row
// blocks round-robin and every CTA keeps three blocks in flight with LDGSTS
// (cp.async) copies that land directly in shared memory:
//
//      iteration i :  [A] 128-bit async copies of colIds / rowStarts of block i+2
//                     [B] 8-byte async gathers  v[colIds[..]] of block i+1 -> smem
//                     [C] CSR-ordered row folds of block i out of smem, coalesced
//                         stores (gv rows, or v[colIds[..]] for scatter / gather+scatter)
//
// so index traffic, indexed value traffic and the store stream of three different
// blocks overlap inside one CTA, and several CTAs per SM overlap their barriers.
// Block descriptors (row range, nnz range) are prefetched two blocks further ahead in
// registers.  Row sums fold left-to-right from the op identity exactly like the
// reference Serial kernel, so results are bit-identical.
// HBM-bound (4 B index + 8..16 B payload per nnz): no tensor cores.
#pragma once
#include "ogs_kernels.cuh"

namespace spn {

// ---- LDGSTS helpers ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// 16-byte async copy of four whole ints.  (The last, partial group of a block is copied word by word below: the
// src-size form of cp.async would do, but compute-sanitizer's initcheck counts it as a 16-byte read and flags the
// bytes past the end of the array.)
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
template <int BYTES>
__device__ __forceinline__ void cp_async_small(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- pipeline geometry -------------------------------------------------------
// kPipeIds: index/rowStart copies run this many blocks ahead of the fold
// kPipeVal: value gathers run this many blocks ahead of the fold
static constexpr int kPipeIds = 2;
static constexpr int kPipeVal = 1;
static constexpr int kIdSlots = kPipeIds + 2;   // see the hazard note in the kernel
static constexpr int kValSlots = kPipeVal + 1;
static constexpr int kSlotInts = kNodesPerBlock + 8;  // widened to 16-byte boundaries at both ends

template <typename T>
struct PipeSmem {
  int ids[kIdSlots][kSlotInts];
  int rows[kIdSlots][kSlotInts];
  T val[kValSlots][kNodesPerBlock];
};

struct BlockDesc {
  int rb, re;      // rows [rb,re)
  int start, end;  // nnz  [start,end)
  __device__ __forceinline__ int nnz() const { return end - start; }
  __device__ __forceinline__ int nrows() const { return re - rb; }
  __device__ __forceinline__ bool empty() const { return re <= rb; }
  __device__ __forceinline__ bool fits() const { return nnz() <= kNodesPerBlock && nrows() <= kNodesPerBlock; }
};

// [A] stage colIds[start,end) at ids[(start&3) + j] and rowStarts[rb..re] at rows[(rb&3) + i]
template <int THREADS>
__device__ __forceinline__ void issue_index_copies(const BlockDesc& d, const int* __restrict__ rowStarts,
                                                   const int* __restrict__ colIds, int* ids, int* rows,
                                                   bool ids_vec, bool rows_vec) {
  if (d.empty() || !d.fits()) return;
  {
    const int a0 = d.start & ~3;
    const int n = d.end - a0;
    if (ids_vec) {
      const int nvec = (n + 3) >> 2;
      for (int q = threadIdx.x; q < nvec; q += THREADS) {
        const int left = n - 4 * q;
        if (left >= 4) {
          cp_async_16(smem_u32(ids + 4 * q), colIds + a0 + 4 * q);
        } else {
          for (int e = 0; e < left; ++e) cp_async_small<4>(smem_u32(ids + 4 * q + e), colIds + a0 + 4 * q + e);
        }
      }
    } else {
      for (int j = (d.start - a0) + threadIdx.x; j < n; j += THREADS)
        cp_async_small<4>(smem_u32(ids + j), colIds + a0 + j);
    }
  }
  {
    const int r0 = d.rb & ~3;
    const int n = d.re + 1 - r0;
    if (rows_vec) {
      const int nvec = (n + 3) >> 2;
      for (int q = threadIdx.x; q < nvec; q += THREADS) {
        const int left = n - 4 * q;
        if (left >= 4) {
          cp_async_16(smem_u32(rows + 4 * q), rowStarts + r0 + 4 * q);
        } else {
          for (int e = 0; e < left; ++e) cp_async_small<4>(smem_u32(rows + 4 * q + e), rowStarts + r0 + 4 * q + e);
        }
      }
    } else {
      for (int j = (d.rb - r0) + threadIdx.x; j < n; j += THREADS)
        cp_async_small<4>(smem_u32(rows + j), rowStarts + r0 + j);
    }
  }
}

// [B] gather v[k + col*K] for every nnz of the block into val[j]
template <typename T, int THREADS>
__device__ __forceinline__ void issue_value_gathers(const BlockDesc& d, const int* ids, const T* v, int K, int k,
                                                    T* val) {
  if (d.empty() || !d.fits()) return;
  const int* c = ids + (d.start & 3);
  const int n = d.nnz();
#pragma unroll 4
  for (int j = threadIdx.x; j < n; j += THREADS)
    cp_async_small<sizeof(T)>(smem_u32(val + j), v + (k + (int64_t)c[j] * K));
}

// [B'] scatter: stage the contiguous gv rows of the block, gv[k + (rb+i)*K] -> val[i]
template <typename T, int THREADS>
__device__ __forceinline__ void issue_row_values(const BlockDesc& d, const T* gv, int K, int k, T* val) {
  if (d.empty() || !d.fits()) return;
  const int n = d.nrows();
#pragma unroll 4
  for (int i = threadIdx.x; i < n; i += THREADS)
    cp_async_small<sizeof(T)>(smem_u32(val + i), gv + (k + (int64_t)(d.rb + i) * K));
}

enum { OGS_GATHER = 0, OGS_SCATTER = 1, OGS_GATHER_SCATTER = 2 };

// MODE OGS_GATHER:          in = v,  out = gv
// MODE OGS_SCATTER:         in = gv, out = v     (OP unused)
// MODE OGS_GATHER_SCATTER:  in = out = v (symmetric: one CSR for both directions)
template <typename T, int OP, int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS)
ogs_pipe_kernel(int NB, int K, const int* __restrict__ blockStarts, const int* __restrict__ rowStarts,
                const int* __restrict__ colIds, const T* in, T* out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PipeSmem<T>& sm = *reinterpret_cast<PipeSmem<T>*>(smem_raw);
  // scatter needs a second value buffer (row values in, per-nnz values out); it lives
  // behind PipeSmem (see ogs_pipe_smem_bytes)
  T* expand = reinterpret_cast<T*>(smem_raw + sizeof(PipeSmem<T>));

  const int k = blockIdx.y;
  const bool ids_vec = (reinterpret_cast<uintptr_t>(colIds) & 15) == 0;
  const bool rows_vec = (reinterpret_cast<uintptr_t>(rowStarts) & 15) == 0;
  const int stride = gridDim.x;
  const int nIter = (NB > (int)blockIdx.x) ? (NB - 1 - (int)blockIdx.x) / stride + 1 : 0;

  auto block_of = [&](int it) { return (int)blockIdx.x + it * stride; };
  auto load_rows = [&](int it, int& rb, int& re) {
    if (it < nIter) {
      const int b = block_of(it);
      rb = __ldg(blockStarts + b);
      re = __ldg(blockStarts + b + 1);
    } else {
      rb = re = 0;
    }
  };
  auto finish_desc = [&](int rb, int re) {
    BlockDesc d;
    d.rb = rb;
    d.re = re;
    if (re > rb) {
      d.start = __ldg(rowStarts + rb);
      d.end = __ldg(rowStarts + re);
    } else {
      d.start = d.end = 0;
    }
    return d;
  };

  // descriptor ring: d0 = block `it`, d1 = it+1, d2 = it+2 ; dn = it+3 (loads in flight),
  // (q_rb,q_re) = row range of block it+4 (loads in flight)
  BlockDesc d0, d1, d2, dn;
  int q_rb, q_re;
  d0.rb = d0.re = d0.start = d0.end = 0;
  d1 = d0;
  d2 = d0;
  {
    int rb, re;
    load_rows(0, rb, re);
    dn = finish_desc(rb, re);   // becomes d2 at it = -2, i.e. block 0
    load_rows(1, q_rb, q_re);
  }

  for (int it = -kPipeIds; it < nIter; ++it) {
    // rotate descriptors: after this, d2 describes block it+2
    d0 = d1;
    d1 = d2;
    d2 = dn;
    dn = finish_desc(q_rb, q_re);         // block it+3, consumed next iteration
    load_rows(it + kPipeIds + 2, q_rb, q_re);  // block it+4

    // [A] index copies of block it+2.  Slot (it+2)%4 was last read by block it-2, which
    // every thread has finished: each thread passed a barrier of iteration it-1.
    {
      const int s = (it + kPipeIds) & (kIdSlots - 1);
      issue_index_copies<THREADS>(d2, rowStarts, colIds, sm.ids[s], sm.rows[s], ids_vec, rows_vec);
    }
    cp_async_commit();
    cp_async_wait<2 * (kPipeIds - kPipeVal)>();  // indices of block it+1 have landed
    __syncthreads();

    // [B] values of block it+1
    if (it + kPipeVal >= 0) {
      const int s = (it + kPipeVal) & (kIdSlots - 1);
      T* val = sm.val[(it + kPipeVal) & (kValSlots - 1)];
      if (MODE == OGS_SCATTER)
        issue_row_values<T, THREADS>(d1, in, K, k, val);
      else
        issue_value_gathers<T, THREADS>(d1, sm.ids[s], in, K, k, val);
    }
    cp_async_commit();
    cp_async_wait<2 * kPipeVal>();  // values of block it have landed
    __syncthreads();

    if (it < 0 || d0.empty()) continue;

    // [C] fold / expand block it
    const int s = it & (kIdSlots - 1);
    const int* c = sm.ids[s] + (d0.start & 3);
    const int* rs = sm.rows[s] + (d0.rb & 3);
    T* val = sm.val[it & (kValSlots - 1)];
    const int nrows = d0.nrows();
    const int start = d0.start;

    if (!d0.fits()) {  // oversized block: correct but unstaged
      for (int row = d0.rb + threadIdx.x; row < d0.re; row += THREADS) {
        const int gs = rowStarts[row], ge = rowStarts[row + 1];
        if (MODE == OGS_SCATTER) {
          const T x = in[k + (int64_t)row * K];
          for (int g = gs; g < ge; ++g) out[k + (int64_t)colIds[g] * K] = x;
        } else {
          T acc = OgsOp<T, OP>::init();
          for (int g = gs; g < ge; ++g) OgsOp<T, OP>::apply(acc, in[k + (int64_t)colIds[g] * K]);
          if (MODE == OGS_GATHER) {
            out[k + (int64_t)row * K] = acc;
          } else {
            for (int g = gs; g < ge; ++g) out[k + (int64_t)colIds[g] * K] = acc;
          }
        }
      }
      continue;
    }

    if (MODE == OGS_GATHER) {
      for (int i = threadIdx.x; i < nrows; i += THREADS) {
        const int gs = rs[i] - start, ge = rs[i + 1] - start;
        T acc = OgsOp<T, OP>::init();
        for (int g = gs; g < ge; ++g) OgsOp<T, OP>::apply(acc, val[g]);
        out[k + (int64_t)(d0.rb + i) * K] = acc;
      }
    } else {
      T* dst = (MODE == OGS_SCATTER) ? expand : val;
      for (int i = threadIdx.x; i < nrows; i += THREADS) {
        const int gs = rs[i] - start, ge = rs[i + 1] - start;
        T acc;
        if (MODE == OGS_SCATTER) {
          acc = val[i];
        } else {
          acc = OgsOp<T, OP>::init();
          for (int g = gs; g < ge; ++g) OgsOp<T, OP>::apply(acc, val[g]);
        }
        for (int g = gs; g < ge; ++g) dst[g] = acc;
      }
      __syncthreads();
      const int n = d0.nnz();
#pragma unroll 4
      for (int j = threadIdx.x; j < n; j += THREADS) out[k + (int64_t)c[j] * K] = dst[j];
      // `expand` is rewritten by the next iteration's fold only after two more barriers
    }
  }
}

template <typename T, int MODE>
constexpr size_t ogs_pipe_smem_bytes() {
  return sizeof(PipeSmem<T>) + (MODE == OGS_SCATTER ? sizeof(T) * kNodesPerBlock : 0);
}

}  // namespace spn
