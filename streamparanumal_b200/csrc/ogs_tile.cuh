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
#include <type_traits>
#include "ogs_kernels.cuh"

namespace spn {

static constexpr unsigned kEmptyLane = 0xFFFFFFFFu;
// colP / row0 carry this many all-empty tiles after the last real one, so a warp may load its
// U consecutive tiles without per-tile bounds checks (U <= kTilePad)
static constexpr int kTilePad = 16;

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
  if (t >= NT + kTilePad) return;
  if (t >= NT) {  // padding tiles
    colP[32 * t + lane] = kEmptyLane;
    if (lane == 0) row0[t] = 0;
    return;
  }
  const OgsTileDesc d = desc[t];
  const int count = 32 - __clz(d.last);
  unsigned w = kEmptyLane;
  if (lane < count) w = (unsigned)colIds[d.start + lane] | (((d.last >> lane) & 1u) << 31);
  colP[32 * t + lane] = w;
  if (lane == 0) row0[t] = d.row0;
}

template <typename T>
__device__ __forceinline__ T shfl_up1(T x) { return __shfl_up_sync(0xffffffffu, x, 1); }

// one fold round: the lane at position t of its row takes its left neighbour's prefix y and
// folds its own value.  FP64 Add gets a predicated DADD (the generic form compiles to
// DADD + two selects).
template <typename T, int OP>
__device__ __forceinline__ void fold_step(T& acc, T y, const T x, int pos, int t) {
  if (pos == t) {
    OgsOp<T, OP>::apply(y, x);
    acc = y;
  }
}
template <>
__device__ __forceinline__ void fold_step<double, SPN_ADD>(double& acc, double y, const double x, int pos, int t) {
  asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %3, %4;\n\t@p add.rn.f64 %0, %1, %2;\n\t}"
      : "+d"(acc) : "d"(y), "d"(x), "r"(pos), "r"(t));
}
// fold round with a ROUND-INVARIANT predicate: every lane that is not the head of its row recomputes
// prefix(left neighbour) (+) own value in every round.  After round t the lanes at positions <= t hold their final
// prefix (induction on the position), later rounds recompute the same bits, and heads never change — so the result
// equals the position-by-position form above, without the per-round compare and the two 64-bit selects.
template <typename T, int OP>
__device__ __forceinline__ void fold_round(T& acc, T y, const T x, bool nonhead) {
  OgsOp<T, OP>::apply(y, x);
  if (nonhead) acc = y;
}
// FP64 Add without any select: heads shuffle from THEMSELVES (delta 0) and add +0.0, which leaves their prefix
// bit-for-bit unchanged — a prefix is the result of identity + x0 (+ ...), hence never -0.0, and p + (+0.0) == p for
// every other double (NaNs are already quiet).  Two SHFL + one DADD per round.
__device__ __forceinline__ void fold_round_add_f64(double& acc, double xeff, int delta) {
  acc = __shfl_up_sync(0xffffffffu, acc, delta) + xeff;
}

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

// column of a packed id word.  The AND is opaque PTX on purpose: when the compiler sees the mask it folds it into
// the 64-bit address arithmetic of EVERY indexed access as a six-instruction shift / mask / carry sequence (and
// re-reads the base pointer from the constant bank under a predicate) — 12 of the ~80 instructions per tile.
// With an opaque 32-bit column, `base + col` is one IMAD.WIDE.U32 and the access stays an LDG / STG.
// (Used by the gather+scatter kernel; the gather kernel at its 64-register budget spills with it and keeps the plain mask.)
template <bool OPAQUE>
__device__ __forceinline__ unsigned col_of(unsigned w) {
  if (!OPAQUE) return w & 0x7fffffffu;
  unsigned c;
  asm("and.b32 %0, %1, 0x7fffffff;" : "=r"(c) : "r"(w));
  return c;
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

  static_assert(U <= kTilePad, "tile padding too small");
  unsigned w[U];
  int r0[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    w[u] = ld_ids(colP + 32 * (base + u) + lane);
    r0[u] = (MODE != TILE_GATHER_SCATTER) ? __ldg(row0 + base + u) : 0;
  }

  for (; base < NT; base += step) {
    bool valid[U];
    int col[U];
    unsigned last[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      valid[u] = (w[u] != kEmptyLane);
      col[u] = (int)(w[u] & 0x7fffffffu);
    }

    // ---- indexed loads of the current tiles (K == 1)
    T x[U];
    if (KONE && MODE != TILE_SCATTER) {
#pragma unroll
      for (int u = 0; u < U; ++u) x[u] = valid[u] ? ld_var<VAR, T>(in + col[u]) : OgsOp<T, OP>::init();
    }

    // ---- prefetch the id lines of the next U tiles
    unsigned wn[U];
    int r0n[U];
    {
      // clamp instead of branching: the last iteration re-reads its own (cached) lines
      const int64_t tn = (base + step < NT) ? base + step : base;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        wn[u] = ld_ids(colP + 32 * (tn + u) + lane);
        r0n[u] = (MODE != TILE_GATHER_SCATTER) ? __ldg(row0 + tn + u) : 0;
      }
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
          if (valid[u]) out[col[u]] = x[u];  // plain: partial sectors must stay in L2 until their neighbours arrive
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
            fold_step<T, OP>(acc[u], shfl_up1(acc[u]), x[u], pos[u], t);
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
            fold_step<T, OP>(a, shfl_up1(a), xv, pos[u], t);
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
    for (int u = 0; u < U; ++u) {
      w[u] = wn[u];
      r0[u] = r0n[u];
    }
  }
}

// ---------------------------------------------------------------------------
// K == 1 gather / gather+scatter with a three-deep register pipeline per warp:
//   iteration i :  load ids of tiles i+2  |  indexed loads of tiles i+1  |  fold + store tiles i
// so a warp never waits for the memory it has just asked for: the values it folds were
// requested one whole iteration earlier.
// ---------------------------------------------------------------------------
template <typename T, int OP, int MODE, int U, int THREADS, int MINB, int VAR = 0>
__global__ void __launch_bounds__(THREADS, MINB)
ogs_tile_pf_kernel(int NTall, int chunkTiles, const unsigned* __restrict__ colP, const int* __restrict__ row0, const T* in, T* out) {
  static_assert(MODE == TILE_GATHER || MODE == TILE_GATHER_SCATTER, "scatter runs by columns");
  static_assert(U <= kTilePad, "tile padding too small");
  const int lane = threadIdx.x & 31;
  const unsigned le_mask = 0xffffffffu >> (31 - lane);
  // Work split.  chunkTiles == 0: persistent grid, warp w of the whole grid walks tiles w*U, w*U + step, ...
  // chunkTiles  > 0: CTA b owns the contiguous tiles [b*chunk, (b+1)*chunk) and retires after them, so that
  // concurrently resident CTAs sweep one contiguous window of the mesh and a higher-priority stream (the halo
  // kernels of a partitioned mesh) finds free SM slots every few microseconds.
  int64_t base, step;
  int NT;
  if (chunkTiles > 0) {
    const int64_t c0 = (int64_t)blockIdx.x * chunkTiles;
    base = c0 + (int64_t)(threadIdx.x >> 5) * U;
    step = (int64_t)(THREADS / 32) * U;
    NT = (int)min((int64_t)NTall, c0 + chunkTiles);
  } else {
    const int warp = (int)((blockIdx.x * (unsigned)THREADS + threadIdx.x) >> 5);
    const int nwarps = (int)((gridDim.x * (unsigned)THREADS) >> 5);
    base = (int64_t)warp * U;
    step = (int64_t)nwarps * U;
    NT = NTall;
  }
  if (base >= NT) return;

  unsigned wc[U], w1[U];
  int rc[U], r1[U];
  T xc[U];
  {
    const int64_t t1 = (base + step < NT) ? base + step : base;  // clamped: harmless re-read
#pragma unroll
    for (int u = 0; u < U; ++u) {
      wc[u] = ld_ids(colP + 32 * (base + u) + lane);
      w1[u] = ld_ids(colP + 32 * (t1 + u) + lane);
      rc[u] = (MODE == TILE_GATHER) ? __ldg(row0 + base + u) : 0;
      r1[u] = (MODE == TILE_GATHER) ? __ldg(row0 + t1 + u) : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      xc[u] = (wc[u] != kEmptyLane) ? ld_once<T>(in + col_of<MODE == TILE_GATHER_SCATTER>(wc[u])) : OgsOp<T, OP>::init();
  }

  while (true) {
    // ---- ids two iterations ahead, values one iteration ahead
    unsigned w2[U];
    int r2[U];
    T x1[U];
    {
      const int64_t t2 = (base + 2 * step < NT) ? base + 2 * step : base;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        w2[u] = ld_ids(colP + 32 * (t2 + u) + lane);
        r2[u] = (MODE == TILE_GATHER) ? __ldg(row0 + t2 + u) : 0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        x1[u] = (w1[u] != kEmptyLane) ? ld_once<T>(in + col_of<MODE == TILE_GATHER_SCATTER>(w1[u])) : OgsOp<T, OP>::init();
    }

    // ---- row structure of the current tiles from the sign bits
    unsigned last[U];
    int pos[U], maxpos = 0;
    T acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool valid = (wc[u] != kEmptyLane);
      last[u] = __ballot_sync(0xffffffffu, valid && (wc[u] >> 31));
      const unsigned heads = ((last[u] << 1) | 1u) & le_mask;
      pos[u] = valid ? lane - (31 - __clz(heads)) : 0;
      maxpos = max(maxpos, pos[u]);
      acc[u] = OgsOp<T, OP>::init();
      OgsOp<T, OP>::apply(acc[u], xc[u]);
    }
    const int R = __reduce_max_sync(0xffffffffu, maxpos);
    if constexpr (sizeof(T) == 8 && OP == SPN_ADD && !std::is_integral<T>::value) {
      double xeff[U];  // own value, or +0.0 on a row head
      int delta[U];    // 1, or 0 on a row head
#pragma unroll
      for (int u = 0; u < U; ++u) {
        xeff[u] = (pos[u] > 0) ? (double)xc[u] : 0.0;
        delta[u] = (pos[u] > 0) ? 1 : 0;
      }
      for (int t = 1; t <= R; ++t) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          double a = (double)acc[u];
          fold_round_add_f64(a, xeff[u], delta[u]);
          acc[u] = (T)a;
        }
      }
    } else {
      for (int t = 1; t <= R; ++t) {
#pragma unroll
        for (int u = 0; u < U; ++u) fold_round<T, OP>(acc[u], shfl_up1(acc[u]), xc[u], pos[u] > 0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool valid = (wc[u] != kEmptyLane);
      if (MODE == TILE_GATHER) {
        const unsigned heads = ((last[u] << 1) | 1u) & le_mask;
        if (valid && (wc[u] >> 31)) out[rc[u] + __popc(heads) - 1] = acc[u];
      } else {
        const int lastlane = lane + __ffs(last[u] >> lane) - 1;
        const T total = shfl_idx(acc[u], lastlane & 31);
        if (valid) __stcs(out + col_of<true>(wc[u]), total);
      }
    }

    base += step;
    if (base >= NT) break;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      wc[u] = w1[u]; w1[u] = w2[u];
      rc[u] = r1[u]; r1[u] = r2[u];
      xc[u] = x1[u];
    }
  }
}

// ---------------------------------------------------------------------------
// scatter by columns.  Every column belongs to at most one row (the ogs invariant,
// include/ogs/ogsOperator.hpp:41-43), so v[colIds[g]] = gv[row] for all g is the same
// as v[c] = gv[rowOf[c]] for every column c that has a row: a pure streaming kernel
// with fully coalesced 256-bit stores of v and mostly-contiguous reads of gv, instead
// of indexed 8-byte stores.  rowOf[c] = -1 marks a column outside the operator (left
// untouched, like the reference).  Index traffic is 4 B per column, what colIds costs.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ogs_build_rowof_kernel(int NT, const unsigned* __restrict__ colP, const int* __restrict__ row0,
                       int* __restrict__ rowOf) {
  const int lane = threadIdx.x & 31;
  const int64_t t = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  if (t >= NT) return;
  const unsigned w = colP[32 * t + lane];
  const bool valid = (w != kEmptyLane);
  const unsigned last = __ballot_sync(0xffffffffu, valid && (w >> 31));
  const unsigned heads = ((last << 1) | 1u) & (0xffffffffu >> (31 - lane));
  if (valid) rowOf[w & 0x7fffffffu] = row0[t] + __popc(heads) - 1;
}

__global__ void __launch_bounds__(128)
ogs_build_rowof_irregular_kernel(int nIrr, const int* __restrict__ irrRows, const int* __restrict__ rowStarts,
                                 const int* __restrict__ colIds, int* __restrict__ rowOf) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= nIrr) return;
  const int row = irrRows[i];
  for (int g = rowStarts[row]; g < rowStarts[row + 1]; ++g) rowOf[colIds[g]] = row;
}

// T is an unsigned type of the element size; VEC = columns per thread (4 needs v 32-byte
// aligned for 8-byte T / 16-byte aligned for 4-byte T, and rowOf 16-byte aligned)
template <typename T, int VEC, int UNROLL>
__global__ void __launch_bounds__(256)
ogs_scatter_cols_kernel(int64_t Ncols, const int* __restrict__ rowOf, const T* __restrict__ gv, T* __restrict__ v) {
  const int64_t ngroups = Ncols / VEC;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t g0 = (int64_t)blockIdx.x * 256 + threadIdx.x; g0 < ngroups; g0 += stride * UNROLL) {
    int r[UNROLL][VEC];
    T x[UNROLL][VEC];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const int64_t g = g0 + j * stride;
      if (g < ngroups) {
        if (VEC == 4) {
          const int4 q = __ldg(reinterpret_cast<const int4*>(rowOf) + g);
          r[j][0] = q.x; r[j][1] = q.y; r[j][2] = q.z; r[j][3] = q.w;
        } else {
          r[j][0] = __ldg(rowOf + g);
        }
      } else {
#pragma unroll
        for (int l = 0; l < VEC; ++l) r[j][l] = -1;
      }
    }
#pragma unroll
    for (int j = 0; j < UNROLL; ++j)
#pragma unroll
      for (int l = 0; l < VEC; ++l)
        if (r[j][l] >= 0) x[j][l] = __ldg(gv + r[j][l]);
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const int64_t g = g0 + j * stride;
      if (VEC == 4 && (r[j][0] | r[j][1] | r[j][2] | r[j][3]) >= 0) {
        if (sizeof(T) == 8) {
          asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(v + 4 * g), "l"((unsigned long long)x[j][0]),
                       "l"((unsigned long long)x[j][1]), "l"((unsigned long long)x[j][2]), "l"((unsigned long long)x[j][3]) : "memory");
        } else {
          asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(v + 4 * g), "r"((unsigned)x[j][0]), "r"((unsigned)x[j][1]),
                       "r"((unsigned)x[j][2]), "r"((unsigned)x[j][3]) : "memory");
        }
      } else {
#pragma unroll
        for (int l = 0; l < VEC; ++l)
          if (r[j][l] >= 0) v[VEC * g + l] = x[j][l];
      }
    }
  }
  if (VEC > 1 && blockIdx.x == 0 && threadIdx.x < VEC) {  // tail columns
    const int64_t c = ngroups * VEC + threadIdx.x;
    if (c < Ncols) {
      const int row = rowOf[c];
      if (row >= 0) v[c] = gv[row];
    }
  }
}

// general K: one thread per (column, k)
template <typename T>
__global__ void __launch_bounds__(256)
ogs_scatter_cols_k_kernel(int64_t Ncols, int K, const int* __restrict__ rowOf, const T* __restrict__ gv, T* __restrict__ v) {
  const int64_t total = Ncols * K;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t c = i / K;
    const int k = (int)(i - c * K);
    const int row = __ldg(rowOf + c);
    if (row >= 0) v[i] = __ldg(gv + (k + (int64_t)row * K));
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
