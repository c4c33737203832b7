// Class-tile gather+scatter for sm_100a (reference kernel: libs/ogs/okl/ogsKernels.okl:32-81,
// device dispatch libs/ogs/ogsOperator.cpp:442-481) — the K==1 in-place gatherScatter of BS8.
//
// The warp-tile kernel of ogs_tile.cuh keeps one non-zero per lane and folds rows ACROSS lanes
// with shuffle rounds; at 79 warp instructions per 32 non-zeros it is issue-bound on B200
// (profiles/r01_ncu_full_summary_k.txt).  Rows of an ogs operator own disjoint columns, so the
// operator may visit rows in any order as long as every row is folded in its CSR order.  This
// plan therefore regroups rows by LENGTH CLASS inside slices of consecutive rows:
//
//   slice  = kSliceRows consecutive rows (keeps the temporal locality of the first-appearance
//            row order: everything a slice touches is in flight at the same time and meets in L2)
//   class  = rows whose length rounds up to L in {1,2,4,8}
//   ctile  = kLines(8) lines of 32 ids (1 KB).  Lane l of a class-L tile owns 8/L whole rows:
//            row-group g (g < 8/L) occupies lines g*L .. g*L+L-1, slot s of the row in line g*L+s.
//            Line g*L+s therefore holds slot s of 32 CONSECUTIVE class-L rows of the slice, so the
//            indexed accesses of a line coalesce exactly as well as 32 consecutive rows do.
//            0xFFFFFFFF = empty slot (row shorter than L, or padding at the end of a class).
//
// A row lives in ONE lane: its fold is a chain of L dependent adds in registers in CSR order
// (identity (+) x0 (+) x1 ...: bit-identical to the reference), with no shuffle, no ballot and no
// row-boundary arithmetic; the row total is stored to every slot of the row.  Every tile carries
// the same 256 slots, so tiles are perfectly balanced.  Per tile: 8 id loads, 8 indexed loads,
// 8 adds, 8 indexed stores (~0.3 warp instructions per non-zero instead of 2.5).
// Rows longer than 8 stay in ogs_tile.cuh tiles ("long" tiles), rows longer than 32 or empty
// are irregular (thread per row) — both rare for the meshes the reference generates.
// HBM-bound: 4 B id + 8 B read + 8 B write per non-zero; rowStarts is never read.
#pragma once
#include "ogs_tile.cuh"

namespace spn {

static constexpr int kCtLines = 8;            // id lines per class tile
static constexpr int kCtSlots = 32 * kCtLines;  // ids per class tile
static constexpr int kCtSliceRowsDefault = 1 << 16;  // rows per slice
static constexpr int kCtPad = 2;              // all-empty tiles after the last one (prefetch may run past the end)

// fold the rows of one lane: lines g*L .. g*L+L-1 are one row
template <typename T, int OP, int L>
__device__ __forceinline__ void ctile_fold(T (&x)[kCtLines]) {
#pragma unroll
  for (int g = 0; g < kCtLines / L; ++g) {
    T acc = OgsOp<T, OP>::init();
#pragma unroll
    for (int s = 0; s < L; ++s) OgsOp<T, OP>::apply(acc, x[g * L + s]);  // empty slots hold the identity
#pragma unroll
    for (int s = 0; s < L; ++s) x[g * L + s] = acc;
  }
}

// in == out for the in-place form.  cls[t] = log2(L) of tile t.
template <typename T, int OP, int THREADS, int MINB, int VAR = 0>
__global__ void __launch_bounds__(THREADS, MINB)
ogs_ctile_gs_kernel(int NT, const unsigned* __restrict__ ids, const unsigned char* __restrict__ cls, const T* in, T* out) {
  const int lane = threadIdx.x & 31;
  const int warp = (int)((blockIdx.x * (unsigned)THREADS + threadIdx.x) >> 5);
  const int nwarps = (int)((gridDim.x * (unsigned)THREADS) >> 5);
  int t = warp;
  if (t >= NT) return;

  unsigned w[kCtLines];
  int c;
  {
    const unsigned* p = ids + (size_t)t * kCtSlots + lane;
#pragma unroll
    for (int j = 0; j < kCtLines; ++j) w[j] = ld_ids(p + 32 * j);
    c = cls[t];
  }
  while (true) {
    // ---- indexed loads of this tile (8 independent requests per lane)
    T x[kCtLines];
#pragma unroll
    for (int j = 0; j < kCtLines; ++j) x[j] = (w[j] != kEmptyLane) ? ld_once<T>(in + w[j]) : OgsOp<T, OP>::init();

    // ---- ids of the warp's next tile while those are outstanding (clamped: harmless re-read)
    const int tn = (t + nwarps < NT) ? t + nwarps : t;
    unsigned wn[kCtLines];
    {
      const unsigned* p = ids + (size_t)tn * kCtSlots + lane;
#pragma unroll
      for (int j = 0; j < kCtLines; ++j) wn[j] = ld_ids(p + 32 * j);
    }
    const int cn = cls[tn];

    // ---- per-lane folds, by class (warp-uniform branch)
    switch (c) {
      case 0: ctile_fold<T, OP, 1>(x); break;
      case 1: ctile_fold<T, OP, 2>(x); break;
      case 2: ctile_fold<T, OP, 4>(x); break;
      default: ctile_fold<T, OP, 8>(x); break;
    }
#pragma unroll
    for (int j = 0; j < kCtLines; ++j)
      if (w[j] != kEmptyLane) st_var<VAR, T>(out + w[j], x[j]);

    t += nwarps;
    if (t >= NT) break;
#pragma unroll
    for (int j = 0; j < kCtLines; ++j) w[j] = wn[j];
    c = cn;
  }
}

// Same operation with a three-deep register pipeline (ids two tiles ahead, values one tile ahead):
// a warp never waits for memory it has just asked for.
template <typename T, int OP, int THREADS, int MINB, int VAR = 0>
__global__ void __launch_bounds__(THREADS, MINB)
ogs_ctile_gs_pf_kernel(int NT, const unsigned* __restrict__ ids, const unsigned char* __restrict__ cls, const T* in, T* out) {
  const int lane = threadIdx.x & 31;
  const int warp = (int)((blockIdx.x * (unsigned)THREADS + threadIdx.x) >> 5);
  const int nwarps = (int)((gridDim.x * (unsigned)THREADS) >> 5);
  int t = warp;
  if (t >= NT) return;

  unsigned wc[kCtLines], w1[kCtLines];
  T xc[kCtLines];
  int cc, c1;
  {
    const int t1 = (t + nwarps < NT) ? t + nwarps : t;
    const unsigned* p = ids + (size_t)t * kCtSlots + lane;
    const unsigned* p1 = ids + (size_t)t1 * kCtSlots + lane;
#pragma unroll
    for (int j = 0; j < kCtLines; ++j) {
      wc[j] = ld_ids(p + 32 * j);
      w1[j] = ld_ids(p1 + 32 * j);
    }
    cc = cls[t];
    c1 = cls[t1];
#pragma unroll
    for (int j = 0; j < kCtLines; ++j) xc[j] = (wc[j] != kEmptyLane) ? ld_once<T>(in + wc[j]) : OgsOp<T, OP>::init();
  }
  while (true) {
    unsigned w2[kCtLines];
    T x1[kCtLines];
    const int t2 = (t + 2 * nwarps < NT) ? t + 2 * nwarps : t;
    {
      const unsigned* p2 = ids + (size_t)t2 * kCtSlots + lane;
#pragma unroll
      for (int j = 0; j < kCtLines; ++j) w2[j] = ld_ids(p2 + 32 * j);
    }
    const int c2 = cls[t2];
#pragma unroll
    for (int j = 0; j < kCtLines; ++j) x1[j] = (w1[j] != kEmptyLane) ? ld_once<T>(in + w1[j]) : OgsOp<T, OP>::init();

    switch (cc) {
      case 0: ctile_fold<T, OP, 1>(xc); break;
      case 1: ctile_fold<T, OP, 2>(xc); break;
      case 2: ctile_fold<T, OP, 4>(xc); break;
      default: ctile_fold<T, OP, 8>(xc); break;
    }
#pragma unroll
    for (int j = 0; j < kCtLines; ++j)
      if (wc[j] != kEmptyLane) st_var<VAR, T>(out + wc[j], xc[j]);

    t += nwarps;
    if (t >= NT) break;
#pragma unroll
    for (int j = 0; j < kCtLines; ++j) {
      wc[j] = w1[j];
      w1[j] = w2[j];
      xc[j] = x1[j];
    }
    cc = c1;
    c1 = c2;
  }
}

}  // namespace spn
