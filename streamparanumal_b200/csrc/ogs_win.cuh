// Window-staged symmetric gather+scatter for sm_100a (reference kernel: libs/ogs/okl/ogsKernels.okl:32-81,
// dispatched by ogsOperator_t::GatherScatter, libs/ogs/ogsOperator.cpp:442-481).
//
// The warp-tile kernel (ogs_tile.cuh) issues one indexed 8-byte load and one indexed 8-byte store per
// non-zero and keeps the ids it prefetches in registers: ~48 KB per SM in flight, 1.6x sector over-fetch
// between L1 and L2, 68 % of the DRAM peak (profiles/r01_ncu_full_summary_f3_ogs.txt).  Rows are numbered in
// first-appearance order (libs/ogs/ogsSetup.cpp:492-559), so the FIRST column of consecutive rows ascends
// and most columns of a run of rows fall into one short contiguous window of v.  This kernel therefore
// stages, per chunk of TC consecutive tiles,
//     window   v[c0, c0+W)                    one 1-D bulk copy (TMA) global -> shared, completes on an mbarrier
//     far      v[farCols[i]], i < nfar        8-byte cp.async global -> shared, no registers held
//     ids      16-bit shared-memory SLOT per non-zero (+ row-end flag in bit 15), 2 B / nnz instead of 4
// several chunks ahead (S stages) from ONE producer warp, while NW consumer warps fold every row out of
// shared memory in CSR order with the same shuffle rounds as the tile kernel (bit-identical to
// `init (+) x0 (+) x1 ...`) and store the row total straight to every column of the row in global memory
// (window columns: c0 + slot, contiguous runs; far columns: through the staged list).  Producer and consumers
// meet only through mbarriers (meta / full / empty per stage) — no __syncthreads in the loop.
// Every column belongs to at most one row (the ogs invariant, include/ogs/ogsOperator.hpp:41-43), hence to
// exactly one chunk: only that chunk reads it for use and writes it; a window that overlaps columns owned by
// another chunk merely carries them along unused.  So chunks never race, in place or out of place.
// HBM-bound (2-6 B index + 16 B payload per nnz): no tensor cores.
#pragma once
#include <type_traits>
#include "ogs_tile.cuh"

// timing probes (wrong results): 1 = no far copies, 2 = no stores, 4 = no window copies
#ifndef SPN_WIN_PROBE
#define SPN_WIN_PROBE 0
#endif

namespace spn {

struct OgsWinDesc {  // one per chunk, 16 bytes
  int c0;            // first column of the window (multiple of 16)
  int W;             // window length in columns (multiple of 16, <= WMAX)
  int nfar;          // columns of the chunk outside the window
  int farStart;      // offset of the chunk's list in farCols (multiple of 4)
};

static constexpr unsigned kWinEmpty = 0x7FFFu;  // slot of an empty lane (bit 15, the row-end flag, clear)

// ---------------------------------------------------------------------------------------------------
// plan build (device): chunk c = tiles [c*TC, (c+1)*TC) of the tile plan
// ---------------------------------------------------------------------------------------------------
// pass 1: window and far count of every chunk (one warp per chunk)
template <int TC, int WMAX>
__global__ void __launch_bounds__(256)
ogs_win_count_kernel(int NC, int NT, int Ncols, const unsigned* __restrict__ colP, OgsWinDesc* __restrict__ desc) {
  const int lane = threadIdx.x & 31;
  const int c = (int)(((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5);
  if (c >= NC) return;
  const int64_t t0 = (int64_t)c * TC;
  const int nt = min(TC, NT - (int)t0);
  const unsigned w0 = __shfl_sync(0xffffffffu, colP[32 * t0 + lane], 0);  // a tile is never empty
  const int c0 = (int)(w0 & 0x7fffffffu) & ~15;
  const int Wfull = min(WMAX, (Ncols - c0) & ~15);
  int nfar = 0, maxin = -1;
  for (int t = 0; t < nt; ++t) {
    const unsigned w = colP[32 * (t0 + t) + lane];
    const bool valid = (w != kEmptyLane);
    const unsigned off = (w & 0x7fffffffu) - (unsigned)c0;
    const bool inwin = valid && off < (unsigned)Wfull;
    nfar += __popc(__ballot_sync(0xffffffffu, valid && !inwin));
    if (inwin) maxin = max(maxin, (int)off);
  }
  maxin = __reduce_max_sync(0xffffffffu, maxin);
  if (lane == 0) {
    OgsWinDesc d;
    d.c0 = c0;
    d.W = min(Wfull, (maxin + 16) & ~15);
    d.nfar = nfar;
    d.farStart = 0;
    desc[c] = d;
  }
}

// pass 2: 16-bit slot ids and far column lists (one warp per chunk; farStart filled in)
template <int TC, int WMAX>
__global__ void __launch_bounds__(256)
ogs_win_fill_kernel(int NC, int NT, const unsigned* __restrict__ colP, const OgsWinDesc* __restrict__ desc,
                    unsigned short* __restrict__ ids16, int* __restrict__ farCols) {
  const int lane = threadIdx.x & 31;
  const int c = (int)(((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5);
  if (c >= NC) return;
  const OgsWinDesc d = desc[c];
  const int64_t t0 = (int64_t)c * TC;
  const int nt = min(TC, NT - (int)t0);
  const unsigned lt_mask = (1u << lane) - 1u;
  int run = 0;
  for (int t = 0; t < TC; ++t) {
    const unsigned w = (t < nt) ? colP[32 * (t0 + t) + lane] : kEmptyLane;
    const bool valid = (w != kEmptyLane);
    const unsigned col = w & 0x7fffffffu;
    const unsigned off = col - (unsigned)d.c0;
    const bool inwin = valid && off < (unsigned)d.W;
    const bool far = valid && !inwin;
    const unsigned fb = __ballot_sync(0xffffffffu, far);
    const int idx = run + __popc(fb & lt_mask);
    unsigned slot = kWinEmpty;
    if (inwin) {
      slot = off | ((w >> 31) << 15);
    } else if (far) {
      slot = (unsigned)(WMAX + idx) | ((w >> 31) << 15);
      farCols[d.farStart + idx] = (int)col;
    }
    ids16[32 * (t0 + t) + lane] = (unsigned short)slot;
    run += __popc(fb);
  }
  // padding of the far list up to a multiple of 4 (bulk copies move multiples of 16 bytes)
  const int padded = (d.nfar + 3) & ~3;
  if (d.nfar + lane < padded) farCols[d.farStart + d.nfar + lane] = 0;
}

// ---------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier, 1-D bulk copy (TMA), cp.async
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy; bytes: multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
template <int BYTES>
__device__ __forceinline__ void cp_async_small(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst)), "l"(src), "n"(BYTES) : "memory");
}
// arrive on `bar` once every cp.async this thread has issued so far has landed (counted in the barrier's init count)
__device__ __forceinline__ void cp_async_arrive(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// shared-memory footprint of one stage: window + far values, far columns, slot ids
template <typename T, int TC, int WMAX, int FMAX>
struct OgsWinStage {
  static constexpr int kVals = (WMAX + FMAX) * (int)sizeof(T);
  static constexpr int kFar = FMAX * 4;
  static constexpr int kIds = TC * 32 * 2;
  static constexpr int kBytes = kVals + kFar + kIds;  // every part is a multiple of 16 bytes
  static_assert(WMAX % 16 == 0 && FMAX % 16 == 0 && WMAX + FMAX < 0x7fff, "window geometry");
};
static constexpr int kWinHeader = 512;  // mbarriers + chunk descriptors
template <typename T, int TC, int WMAX, int FMAX, int S>
constexpr int ogs_win_smem_bytes() { return S * OgsWinStage<T, TC, WMAX, FMAX>::kBytes + kWinHeader; }

// ---------------------------------------------------------------------------------------------------
// the kernel: persistent grid, CTA b owns chunks b, b+G, b+2G, ... (concurrent CTAs sweep one contiguous
// region of the mesh, so windows that overlap their neighbours' meet in L2).
//   warp NW, lane 0 (producer):  chunk n -> stage n % S:  wait empty | bulk ids + far columns (meta barrier) |
//                                bulk window (full barrier)
//   warps 0..NW-1, iteration n:  wait meta of chunk n+L | cp.async this warp's share of its far values | arrive on
//                                that chunk's full barrier;  wait full of chunk n | fold U tiles | arrive on empty | store
// meta of chunk n+L can only complete after every warp released the stage's previous chunk, so the far copies
// never overwrite values somebody still reads.  The plan is only built when no chunk has more than FMAX far
// columns (otherwise the warp-tile kernel runs).
// ---------------------------------------------------------------------------------------------------
template <typename T, int OP, int TC, int WMAX, int FMAX, int S, int L, int NW, int MINB>
__global__ void __launch_bounds__((NW + 1) * 32, MINB)
ogs_win_gs_kernel(int NC, const OgsWinDesc* __restrict__ desc, const unsigned short* __restrict__ ids16,
                  const int* __restrict__ farCols, const T* in, T* out) {
  using St = OgsWinStage<T, TC, WMAX, FMAX>;
  constexpr int U = TC / NW;
  static_assert(TC % NW == 0 && U >= 1, "a chunk is a whole number of tiles per warp");
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "element size");
  static_assert(L >= 1 && L < S, "far copies run L chunks ahead of the fold, inside the S stages");
  static_assert(3 * S * 8 + S * 16 <= kWinHeader, "header");

  extern __shared__ __align__(128) unsigned char smem[];
  unsigned long long* barMeta = reinterpret_cast<unsigned long long*>(smem);
  unsigned long long* barFull = barMeta + S;
  unsigned long long* barEmpty = barFull + S;
  OgsWinDesc* sdesc = reinterpret_cast<OgsWinDesc*>(barEmpty + S);
  unsigned char* stage0 = smem + kWinHeader;
  auto vals_of = [&](int s) { return reinterpret_cast<T*>(stage0 + (size_t)s * St::kBytes); };
  auto far_of = [&](int s) { return reinterpret_cast<int*>(stage0 + (size_t)s * St::kBytes + St::kVals); };
  auto ids_of = [&](int s) { return reinterpret_cast<unsigned short*>(stage0 + (size_t)s * St::kBytes + St::kVals + St::kFar); };

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x;
  const int nIter = (NC - (int)blockIdx.x + G - 1) / G;  // grid <= NC

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(barMeta + s, 1);
      mbar_init(barFull + s, 1 + NW * 32);  // the window's expect_tx + one cp.async arrival per consumer lane
      mbar_init(barEmpty + s, NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == NW) {
    // ================= producer =================
    if (lane != 0) return;
    OgsWinDesc d = desc[blockIdx.x];
    int s = 0;
    unsigned ph = 1;  // parity of the PREVIOUS use of the stage
    for (int n = 0; n < nIter; ++n) {
      const int64_t c = (int64_t)blockIdx.x + (int64_t)n * G;
      OgsWinDesc dn = d;
      if (n + 1 < nIter) dn = desc[c + G];  // in flight while this chunk is issued
      if (n >= S) mbar_wait(barEmpty + s, ph);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      sdesc[s] = d;
      const unsigned farBytes = (unsigned)((d.nfar + 3) & ~3) * 4u;  // nfar <= FMAX by construction of the plan
      mbar_expect_tx(barMeta + s, (unsigned)St::kIds + farBytes);
      bulk_g2s(ids_of(s), ids16 + c * (TC * 32), St::kIds, barMeta + s);
      if (farBytes) bulk_g2s(far_of(s), farCols + d.farStart, farBytes, barMeta + s);
      const unsigned winBytes = (SPN_WIN_PROBE & 4) ? 0u : (unsigned)d.W * (unsigned)sizeof(T);
      mbar_expect_tx(barFull + s, winBytes);
      if (winBytes) bulk_g2s(vals_of(s), in + d.c0, winBytes, barFull + s);
      d = dn;
      if (++s == S) { s = 0; ph ^= 1u; }
    }
    return;
  }

  // ================= consumer warps =================
  // this warp's share of the far values of local chunk m (stage sf, parity pf)
  auto issue_far = [&](int sf, unsigned pf) {
    mbar_wait(barMeta + sf, pf);
    const int nfar = sdesc[sf].nfar;
    const int* fc = far_of(sf);
    T* fv = vals_of(sf) + WMAX;
#if !(SPN_WIN_PROBE & 1)
    for (int i = tid; i < nfar; i += NW * 32) cp_async_small<(int)sizeof(T)>(fv + i, in + fc[i]);
#endif
    cp_async_arrive(barFull + sf);
  };
  int sf = 0;
  unsigned pf = 0;
  for (int m = 0; m < L && m < nIter; ++m) {
    issue_far(sf, pf);
    if (++sf == S) { sf = 0; pf ^= 1u; }
  }

  const unsigned le_mask = 0xffffffffu >> (31 - lane);
  int s = 0;
  unsigned ph = 0;
  for (int n = 0; n < nIter; ++n) {
    if (n + L < nIter) {
      issue_far(sf, pf);
      if (++sf == S) { sf = 0; pf ^= 1u; }
    }
    mbar_wait(barFull + s, ph);  // (every lane has seen this chunk's meta barrier complete before arriving here)
    const int c0 = sdesc[s].c0;
    const T* vals = vals_of(s);
    const int* fc = far_of(s) - WMAX;  // indexed by slot
    const unsigned short* ids = ids_of(s) + (warp * U) * 32 + lane;

    unsigned slot[U], last[U];
    int pos[U], col[U], maxpos = 0;
    T acc[U], x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned w = ids[u * 32];
      last[u] = __ballot_sync(0xffffffffu, (w & 0x8000u) != 0u);
      slot[u] = w & 0x7fffu;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      x[u] = OgsOp<T, OP>::init();
      if (slot[u] != kWinEmpty) x[u] = vals[slot[u]];
      col[u] = c0 + (int)slot[u];
      if (slot[u] >= (unsigned)WMAX && slot[u] != kWinEmpty) col[u] = fc[slot[u]];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned heads = ((last[u] << 1) | 1u) & le_mask;
      pos[u] = (slot[u] != kWinEmpty) ? lane - (31 - __clz(heads)) : 0;
      maxpos = max(maxpos, pos[u]);
      acc[u] = OgsOp<T, OP>::init();
      OgsOp<T, OP>::apply(acc[u], x[u]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(barEmpty + s);  // everything this warp needs from the stage is in registers
    if (++s == S) { s = 0; ph ^= 1u; }
    const int R = __reduce_max_sync(0xffffffffu, maxpos);
    if constexpr (sizeof(T) == 8 && OP == SPN_ADD && !std::is_integral<T>::value) {
      double xeff[U];
      int delta[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        xeff[u] = (pos[u] > 0) ? (double)x[u] : 0.0;
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
        for (int u = 0; u < U; ++u) fold_round<T, OP>(acc[u], shfl_up1(acc[u]), x[u], pos[u] > 0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int lastlane = lane + __ffs(last[u] >> lane) - 1;
      const T total = shfl_idx(acc[u], lastlane & 31);
#if (SPN_WIN_PROBE & 2)
      if (slot[u] != kWinEmpty && total == T(12345)) out[col[u]] = total;
#else
      if (slot[u] != kWinEmpty) out[col[u]] = total;
#endif
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
}

}  // namespace spn
