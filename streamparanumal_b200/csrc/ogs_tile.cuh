// Warp-tile ogs kernels for sm_100a: gather, scatter and symmetric gather+scatter over a
// packed tile plan (reference kernels: libs/ogs/okl/ogsKernels.okl:32-164; the plan takes
// the place of the reference's device CSR + blockRowStarts launch partition,
// libs/ogs/ogsOperator.cpp:522-665).
//
// Plan (built once per operator, see ogs_plan.cu): the CSR rows are cut greedily into
// tiles of <=32 non-zeros made of WHOLE consecutive rows, and the column ids are re-laid
// out as one 128-byte line per tile:
//     colP[32*t + l] = colIds[start_t + l] | (nnz ends its row ? 1<<31 : 0),   0xFFFFFFFF = empty lane
//     row0[t]        = first row of tile t            (gather / scatter only)
// Row boundaries travel in the sign bit of the ids, so rowStarts is never read by the
// device: 4.0 B/nnz of index traffic instead of the reference's 4 + 4*Nrows/nnz.
// Rows that are empty or longer than 32 are "irregular": a thread-per-row kernel handles them.
//
// Kernel: one warp per tile, lane l owns one non-zero.  No shared memory, no barrier.
// The CSR-ordered left fold of every row (identity (+) x0 (+) x1 ...) runs as shuffle
// rounds: in round t the lane at position t of its row takes its left neighbour's prefix
// and folds its own value — exactly the reference's summation order, so results are
// bit-identical.  Each warp keeps U tiles in flight and prefetches the id lines of its
// next U tiles while the indexed loads of the current ones are outstanding; a persistent
// grid of full-occupancy CTAs keeps >50 KB per SM in flight.
// HBM-bound (4 B index + 8..16 B payload per nnz): no tensor cores.
#pragma once
#include "ogs_kernels.cuh"

namespace spn {

static constexpr unsigned kEmptyLane = 0xFFFFFFFFu;

// build-time descriptor (host -> pack kernel); not used by the hot kernels
struct OgsTileDesc {
  int start;      // first nnz
  int row0;       // first row
  unsigned last;  // bit l set <=> nnz start+l is the last of its row
};

enum { TILE_GATHER = 0, TILE_SCATTER = 1, TILE_GATHER_SCATTER = 2 };

__global__ void __launch_bounds__(256)
ogs_pack_tiles_kernel(int NT, const OgsTileDesc* __restrict__ desc, const int* __restrict__ colIds,
                      unsigned* __restrict__ colP, int* __restrict__ row0) {
  const int lane = threadIdx.x & 31;
  const int64_t t = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  if (t >= NT) return;
  const OgsTileDesc d = desc[t];
  const int count = 32 - __clz(d.last);
  unsigned w = kEmptyLane;
  if (lane < count) w = (unsigned)colIds[d.start + lane] | (((d.last >> lane) & 1u) << 31);
  colP[32 * t + lane] = w;
  if (lane == 0) row0[t] = d.row0;
}

template <typename T>
__device__ __forceinline__ T shfl_up1(T x) { return __shfl_up_sync(0xffffffffu, x, 1); }
template <typename T>
__device__ __forceinline__ T shfl_idx(T x, int src) { return __shfl_sync(0xffffffffu, x, src); }

// indexed load of v.  Deliberately a plain C++ load: an `asm volatile` load with a "memory"
// clobber pins every load behind the previous tile's stores and serialises the U tiles a warp
// keeps in flight (measured: 2539 -> 4877 GB/s on BS8, profiles/round1_notes.md).
template <typename T>
__device__ __forceinline__ T ld_once(const T* p) { return *p; }
__device__ __forceinline__ unsigned ld_ids(const unsigned* p) {
  unsigned r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// experiment knobs (OGS_PROBE in tools/quick_ogs.py): VAR bit1 = plain instead of streaming
// (.cs) indexed stores, bit2 = .nc indexed loads
template <int VAR, typename T>
__device__ __forceinline__ T ld_var(const T* p) {
  if (VAR & 4) return __ldg(p);
  return ld_once<T>(p);
}
template <int VAR, typename T>
__device__ __forceinline__ void st_var(T* p, T x) {
  if (VAR & 2) *p = x; else __stcs(p, x);
}

// MODE TILE_GATHER:          in = v,  out = gv
// MODE TILE_SCATTER:         in = gv, out = v      (OP unused)
// MODE TILE_GATHER_SCATTER:  in = out = v
// KONE: K == 1 fast path
template <typename T, int OP, int MODE, int U, bool KONE, int THREADS, int MINB, int VAR = 0>
__global__ void __launch_bounds__(THREADS, MINB)
ogs_tile_kernel(int NT, int K, const unsigned* __restrict__ colP, const int* __restrict__ row0, const T* in,
                T* out) {
  const int lane = threadIdx.x & 31;
  const int warp = (int)((blockIdx.x * (unsigned)THREADS + threadIdx.x) >> 5);
  const int nwarps = (int)((gridDim.x * (unsigned)THREADS) >> 5);
  const unsigned le_mask = 0xffffffffu >> (31 - lane);
  const int64_t step = (int64_t)nwarps * U;

  int64_t base = (int64_t)warp * U;
  if (base >= NT) return;

  unsigned w[U];
#pragma unroll
  for (int u = 0; u < U; ++u) w[u] = (base + u < NT) ? ld_ids(colP + 32 * (base + u) + lane) : kEmptyLane;

  for (; base < NT; base += step) {
    bool valid[U];
    int col[U];
    unsigned last[U];
    int r0[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      valid[u] = (w[u] != kEmptyLane);
      col[u] = (int)(w[u] & 0x7fffffffu);
      if (MODE != TILE_GATHER_SCATTER) r0[u] = (base + u < NT) ? __ldg(row0 + base + u) : 0;
    }

    // ---- indexed loads of the current tiles (K == 1)
    T x[U];
    if (KONE && MODE != TILE_SCATTER) {
#pragma unroll
      for (int u = 0; u < U; ++u) x[u] = valid[u] ? ld_var<VAR, T>(in + col[u]) : OgsOp<T, OP>::init();
    }

    // ---- prefetch the id lines of the next U tiles
    unsigned wn[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t tn = base + step + u;
      wn[u] = (tn < NT) ? ld_ids(colP + 32 * tn + lane) : kEmptyLane;
    }

    // ---- row structure from the sign bits
    int pos[U], rowidx[U], maxpos = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      last[u] = __ballot_sync(0xffffffffu, valid[u] && (w[u] >> 31));
      const unsigned heads = ((last[u] << 1) | 1u) & le_mask;
      pos[u] = valid[u] ? lane - (31 - __clz(heads)) : 0;
      rowidx[u] = __popc(heads) - 1;
      maxpos = max(maxpos, pos[u]);
    }

    if (KONE) {
      if (MODE == TILE_SCATTER) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (valid[u]) x[u] = __ldg(in + r0[u] + rowidx[u]);
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (valid[u]) st_var<VAR, T>(out + col[u], x[u]);
      } else {
        const int R = __reduce_max_sync(0xffffffffu, maxpos);
        T acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          acc[u] = OgsOp<T, OP>::init();
          OgsOp<T, OP>::apply(acc[u], x[u]);
        }
        for (int t = 1; t <= R; ++t) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            T y = shfl_up1(acc[u]);
            if (pos[u] == t) {
              OgsOp<T, OP>::apply(y, x[u]);
              acc[u] = y;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (MODE == TILE_GATHER) {
            if (valid[u] && (w[u] >> 31)) out[r0[u] + rowidx[u]] = acc[u];
          } else {
            const int lastlane = lane + __ffs(last[u] >> lane) - 1;
            const T total = shfl_idx(acc[u], lastlane & 31);
            if (valid[u]) st_var<VAR, T>(out + col[u], total);
          }
        }
      }
    } else {
      // general K: same tiles, one pass per k (ids are read once for all k)
      const int R = __reduce_max_sync(0xffffffffu, maxpos);
      for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t row = (int64_t)r0[u] + rowidx[u];
          if (MODE == TILE_SCATTER) {
            if (valid[u]) out[k + (int64_t)col[u] * K] = __ldg(in + (k + row * K));
            continue;
          }
          const T xv = valid[u] ? in[k + (int64_t)col[u] * K] : OgsOp<T, OP>::init();
          T a = OgsOp<T, OP>::init();
          OgsOp<T, OP>::apply(a, xv);
          for (int t = 1; t <= R; ++t) {
            T y = shfl_up1(a);
            if (pos[u] == t) {
              OgsOp<T, OP>::apply(y, xv);
              a = y;
            }
          }
          if (MODE == TILE_GATHER) {
            if (valid[u] && (w[u] >> 31)) out[k + row * K] = a;
          } else {
            const int lastlane = lane + __ffs(last[u] >> lane) - 1;
            const T total = shfl_idx(a, lastlane & 31);
            if (valid[u]) out[k + (int64_t)col[u] * K] = total;
          }
        }
      }
    }

#pragma unroll
    for (int u = 0; u < U; ++u) w[u] = wn[u];
  }
}

// Irregular rows (empty, or longer than a warp): one thread per row, sequential CSR-ordered
// fold straight from global memory.  Rare by construction.
template <typename T, int OP, int MODE>
__global__ void __launch_bounds__(128)
ogs_irregular_kernel(int nIrr, int K, const int* __restrict__ irrRows, const int* __restrict__ rowStarts,
                     const int* __restrict__ colIds, const T* in, T* out) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= nIrr) return;
  const int row = irrRows[i];
  const int s = rowStarts[row], e = rowStarts[row + 1];
  for (int k = 0; k < K; ++k) {
    if (MODE == TILE_SCATTER) {
      const T x = in[k + (int64_t)row * K];
      for (int g = s; g < e; ++g) out[k + (int64_t)colIds[g] * K] = x;
    } else {
      T acc = OgsOp<T, OP>::init();
      for (int g = s; g < e; ++g) OgsOp<T, OP>::apply(acc, in[k + (int64_t)colIds[g] * K]);
      if (MODE == TILE_GATHER) {
        out[k + (int64_t)row * K] = acc;
      } else {
        for (int g = s; g < e; ++g) out[k + (int64_t)colIds[g] * K] = acc;
      }
    }
  }
}

}  // namespace spn
