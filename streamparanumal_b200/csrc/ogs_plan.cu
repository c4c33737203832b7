// Tile-plan entry points of include/spn_b200.h (spn_ogs_plan_*): the B200-native form of
// the reference's device operator (ogsOperator_t: libs/ogs/ogsOperator.cpp:597-665 uploads
// rowStarts/colIds/blockRowStarts; here the row partition is a warp-tile plan, see
// ogs_tile.cuh) and its three kernels (libs/ogs/ogsOperator.cpp:176-204,295-322,442-481).
#include <algorithm>
#include <atomic>
#include <new>
#include <thread>
#include <system_error>
#include <vector>
#include "ogs_tile.cuh"

// Launch geometry per operation, from the sweeps in profiles/r01_notes.md (B200, degree-7 hex; gather+scatter re-swept
// after the instruction diet of the fold: U=4 with four resident CTAs beats U=8 with two, 0.487 vs 0.517 ms):
// U = tiles a warp keeps in flight, MINB = resident 256-thread CTAs per SM the register
// allocator must allow.  -DSPN_TILE_U / -DSPN_TILE_MINB override all three for sweeps.
#ifndef SPN_TILE_THREADS
#define SPN_TILE_THREADS 256
#endif
#ifdef SPN_TILE_U
#define SPN_TILE_U_G SPN_TILE_U
#define SPN_TILE_U_S SPN_TILE_U
#define SPN_TILE_U_GS SPN_TILE_U
#else
#define SPN_TILE_U_G 4
#define SPN_TILE_U_S 8
#define SPN_TILE_U_GS 4
#endif
#ifdef SPN_TILE_MINB
#define SPN_TILE_MINB_G SPN_TILE_MINB
#define SPN_TILE_MINB_S SPN_TILE_MINB
#define SPN_TILE_MINB_GS SPN_TILE_MINB
#else
#define SPN_TILE_MINB_G 4
#define SPN_TILE_MINB_S 3
#define SPN_TILE_MINB_GS 4
#endif

// CTA size per operation (last sweep of the round, profiles/r01_notes.md): at equal registers per SM, gather is
// fastest with 512-thread CTAs (0.354 ms), gather+scatter with 128-thread CTAs (0.477 ms).
#ifdef SPN_TILE_THREADS_ALL
#define SPN_TILE_TH_G SPN_TILE_THREADS_ALL
#define SPN_TILE_TH_GS SPN_TILE_THREADS_ALL
#define SPN_TILE_CTAS_G SPN_TILE_MINB_G
#define SPN_TILE_CTAS_GS SPN_TILE_MINB_GS
#else
#define SPN_TILE_TH_G 512
#define SPN_TILE_TH_GS 128
#define SPN_TILE_CTAS_G (SPN_TILE_MINB_G * 256 / 512)
#define SPN_TILE_CTAS_GS (SPN_TILE_MINB_GS * 256 / 128)
#endif

template <int MODE> struct TileGeom;
template <> struct TileGeom<spn::TILE_GATHER> { static constexpr int U = SPN_TILE_U_G, MINB = SPN_TILE_MINB_G, PF_TH = SPN_TILE_TH_G, PF_MINB = SPN_TILE_CTAS_G; };
template <> struct TileGeom<spn::TILE_SCATTER> { static constexpr int U = SPN_TILE_U_S, MINB = SPN_TILE_MINB_S, PF_TH = SPN_TILE_THREADS, PF_MINB = SPN_TILE_MINB_S; };
template <> struct TileGeom<spn::TILE_GATHER_SCATTER> { static constexpr int U = SPN_TILE_U_GS, MINB = SPN_TILE_MINB_GS, PF_TH = SPN_TILE_TH_GS, PF_MINB = SPN_TILE_CTAS_GS; };

using namespace spn;

struct spn_ogs_plan {
  int Nrows;
  int NT;    // tiles
  int nIrr;  // irregular rows
  long long nnz;
  unsigned* d_colP;  // packed ids, one 128-byte line per tile (owned)
  int* d_row0;       // first row of each tile (owned)
  int* d_irr;        // irregular rows (owned)
  int* d_rowOf;      // row of every column, -1 = none (owned); drives scatter
  long long Ncols;
  const int* d_rowStarts;  // borrowed
  const int* d_colIds;     // borrowed
};

namespace {

// iterations a warp of the gather / gather+scatter tile kernel runs before its CTA retires (spn_debug_set key 4);
// 0 = persistent grid.  Short-lived CTAs let the halo kernels of a partitioned mesh (higher-priority stream) in.
// Measured at 64^3 (profiles/r01_notes.md): gather+scatter 0.483 ms with 8 iterations per CTA vs 0.547 ms persistent;
// gather 0.365 ms persistent vs 0.389 ms chunked — so gather stays persistent.
int g_chunk_iters = 8;
int g_chunk_iters_gather = 0;

// resident CTAs per SM of a kernel, cached PER DEVICE (the cache of one instantiation is an array indexed by the device
// ordinal; entries are written once with the value every thread would compute, through relaxed atomics)
struct OccupancyCache {
  std::atomic<int> per_device[64];
};
template <typename K>
int resident_ctas(K kern, int threads, int device, OccupancyCache& cache, int* out) {
  std::atomic<int>& slot = cache.per_device[device & 63];
  int v = slot.load(std::memory_order_relaxed);
  if (!v) {
    int per_sm = 0;
    SPN_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0));
    v = per_sm > 0 ? per_sm : 1;
    slot.store(v, std::memory_order_relaxed);
  }
  *out = v;
  return SPN_SUCCESS;
}

template <typename T, int OP, int MODE>
int launch_tiles(const spn_ogs_plan* p, int K, const void* in, void* out, cudaStream_t st) {
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  const unsigned* colP = p->d_colP;
  const int NT = p->NT;
  if (NT > 0) {
    constexpr int U = TileGeom<MODE>::U, TH = SPN_TILE_THREADS, MB = TileGeom<MODE>::MINB;
    const int warps_needed = (NT + U - 1) / U;
    const int ctas_needed = (warps_needed + TH / 32 - 1) / (TH / 32);
    if (K == 1 && MODE != TILE_SCATTER) {
      constexpr int M2 = (MODE == TILE_SCATTER) ? TILE_GATHER : MODE;  // (scatter never reaches this kernel)
      constexpr int PTH = TileGeom<MODE>::PF_TH, PMB = TileGeom<MODE>::PF_MINB;
      auto kern = ogs_tile_pf_kernel<T, OP, M2, U, PTH, PMB>;
      static OccupancyCache cache;
      int res = 0;
      rc = resident_ctas(kern, PTH, d->device, cache, &res);
      if (rc) return rc;
      const int pf_ctas_needed = (warps_needed + PTH / 32 - 1) / (PTH / 32);
      const int iters = (MODE == TILE_GATHER) ? g_chunk_iters_gather : g_chunk_iters;
      const int chunk = iters > 0 ? iters * (PTH / 32) * U : 0;  // every warp of the CTA runs `iters` iterations
      if (chunk > 0 && pf_ctas_needed > res * d->sms)
        kern<<<(NT + chunk - 1) / chunk, PTH, 0, st>>>(NT, chunk, colP, p->d_row0, static_cast<const T*>(in), static_cast<T*>(out));
      else
        kern<<<std::min(pf_ctas_needed, res * d->sms), PTH, 0, st>>>(NT, 0, colP, p->d_row0, static_cast<const T*>(in), static_cast<T*>(out));
    } else if (K == 1) {
      auto kern = ogs_tile_kernel<T, OP, MODE, U, true, TH, MB>;
      static OccupancyCache cache;
      int res = 0;
      rc = resident_ctas(kern, TH, d->device, cache, &res);
      if (rc) return rc;
      kern<<<std::min(ctas_needed, res * d->sms), TH, 0, st>>>(NT, K, colP, p->d_row0, static_cast<const T*>(in), static_cast<T*>(out));
    } else {
      auto kern = ogs_tile_kernel<T, OP, MODE, U, false, TH, MB>;
      static OccupancyCache cache;
      int res = 0;
      rc = resident_ctas(kern, TH, d->device, cache, &res);
      if (rc) return rc;
      kern<<<std::min(ctas_needed, res * d->sms), TH, 0, st>>>(NT, K, colP, p->d_row0, static_cast<const T*>(in), static_cast<T*>(out));
    }
    SPN_LAUNCH_CHECK();
  }
  if (p->nIrr > 0) {
    ogs_irregular_kernel<T, OP, MODE><<<(p->nIrr + 127) / 128, 128, 0, st>>>(
        p->nIrr, K, p->d_irr, p->d_rowStarts, p->d_colIds, static_cast<const T*>(in), static_cast<T*>(out));
    SPN_LAUNCH_CHECK();
  }
  return SPN_SUCCESS;
}

template <typename T, int MODE>
int dispatch_op(const spn_ogs_plan* p, int K, const void* in, void* out, int op, cudaStream_t st) {
  switch (op) {
    case SPN_ADD: return launch_tiles<T, SPN_ADD, MODE>(p, K, in, out, st);
    case SPN_MUL: return launch_tiles<T, SPN_MUL, MODE>(p, K, in, out, st);
    case SPN_MAX: return launch_tiles<T, SPN_MAX, MODE>(p, K, in, out, st);
    case SPN_MIN: return launch_tiles<T, SPN_MIN, MODE>(p, K, in, out, st);
    default: return SPN_ERR_INVALID_ARGUMENT;
  }
}

template <int MODE>
int dispatch_type(const spn_ogs_plan* p, int K, const void* in, void* out, int type, int op, cudaStream_t st) {
  switch (type) {
    case SPN_FLOAT: return dispatch_op<float, MODE>(p, K, in, out, op, st);
    case SPN_DOUBLE: return dispatch_op<double, MODE>(p, K, in, out, op, st);
    case SPN_INT32: return dispatch_op<int, MODE>(p, K, in, out, op, st);
    case SPN_INT64: return dispatch_op<long long, MODE>(p, K, in, out, op, st);
    default: return SPN_ERR_INVALID_ARGUMENT;
  }
}

}  // namespace

// rows [r0, r1) packed greedily into tiles of whole consecutive rows with at most 32 non-zeros; empty rows and rows
// longer than a tile go to the irregular list
constexpr int kPackChunkRows = 1 << 20;
struct PackOut {
  std::vector<OgsTileDesc> tiles;
  std::vector<int> irr;
  int bad = 0;  // SPN_ERR_* when the row starts are not monotone / memory ran out
};
static void pack_rows(const int* rowStarts, int r0, int r1, PackOut& out) {
  try {
    out.tiles.reserve((size_t)((long long)rowStarts[r1] - rowStarts[r0]) / 24 + 16);
    OgsTileDesc cur;
    int cur_n = 0;
    auto flush = [&]() {
      if (cur_n > 0) out.tiles.push_back(cur);
      cur_n = 0;
    };
    for (int r = r0; r < r1; ++r) {
      const int s = rowStarts[r], len = rowStarts[r + 1] - s;
      if (len < 0) { out.bad = SPN_ERR_INVALID_ARGUMENT; return; }
      if (len == 0 || len > 32) {
        flush();
        out.irr.push_back(r);
        continue;
      }
      if (cur_n + len > 32) flush();
      if (cur_n == 0) {
        cur.start = s;
        cur.row0 = r;
        cur.last = 0u;
      }
      cur_n += len;
      cur.last |= 1u << (cur_n - 1);
    }
    flush();
  } catch (const std::bad_alloc&) {
    out.bad = SPN_ERR_OUT_OF_MEMORY;
  }
}

extern "C" {

static int status_of_cuda(int rc) { return rc == (int)cudaErrorMemoryAllocation ? SPN_ERR_OUT_OF_MEMORY : rc; }

// device side of spn_ogs_plan_create; on failure the caller destroys whatever was allocated
static int build_device_plan(spn_ogs_plan* p, const std::vector<OgsTileDesc>& tiles, const std::vector<int>& irr,
                             OgsTileDesc** d_desc) {
  if (p->NT) {
    // re-lay the ids out as one 128-byte line per tile (device side; the descriptors are scratch)
    SPN_CUDA_TRY(cudaMalloc(d_desc, sizeof(OgsTileDesc) * tiles.size()));
    SPN_CUDA_TRY(cudaMemcpy(*d_desc, tiles.data(), sizeof(OgsTileDesc) * tiles.size(), cudaMemcpyHostToDevice));
    SPN_CUDA_TRY(cudaMalloc(&p->d_colP, sizeof(unsigned) * 32 * (tiles.size() + kTilePad)));
    SPN_CUDA_TRY(cudaMalloc(&p->d_row0, sizeof(int) * (tiles.size() + kTilePad)));
    const long long blocks = (((long long)p->NT + kTilePad) * 32 + 255) / 256;
    ogs_pack_tiles_kernel<<<(unsigned)blocks, 256>>>(p->NT, *d_desc, p->d_colIds, p->d_colP, p->d_row0);
    SPN_LAUNCH_CHECK();
    SPN_CUDA_TRY(cudaDeviceSynchronize());
  }
  if (p->nIrr) {
    SPN_CUDA_TRY(cudaMalloc(&p->d_irr, sizeof(int) * irr.size()));
    SPN_CUDA_TRY(cudaMemcpy(p->d_irr, irr.data(), sizeof(int) * irr.size(), cudaMemcpyHostToDevice));
  }
  if (p->Ncols > 0) {
    // inverse map for scatter: rowOf[c] = row holding column c, -1 if none
    SPN_CUDA_TRY(cudaMalloc(&p->d_rowOf, sizeof(int) * (size_t)p->Ncols));
    SPN_CUDA_TRY(cudaMemset(p->d_rowOf, 0xff, sizeof(int) * (size_t)p->Ncols));
    if (p->NT) {
      const long long blocks = ((long long)p->NT * 32 + 255) / 256;
      ogs_build_rowof_kernel<<<(unsigned)blocks, 256>>>(p->NT, p->d_colP, p->d_row0, p->d_rowOf);
      SPN_LAUNCH_CHECK();
    }
    if (p->nIrr) {
      ogs_build_rowof_irregular_kernel<<<(p->nIrr + 127) / 128, 128>>>(p->nIrr, p->d_irr, p->d_rowStarts, p->d_colIds, p->d_rowOf);
      SPN_LAUNCH_CHECK();
    }
    SPN_CUDA_TRY(cudaDeviceSynchronize());
  }
  return SPN_SUCCESS;
}

int spn_ogs_plan_create(spn_ogs_plan_t* plan, int Nrows, int Ncols, const int* h_rowStarts, const int* h_colIds,
                        const int* d_rowStarts, const int* d_colIds) {
  if (!plan || Nrows < 0 || Ncols < 0 || (Nrows > 0 && (!h_rowStarts || !d_rowStarts || !d_colIds)))
    return SPN_ERR_INVALID_ARGUMENT;
  *plan = nullptr;
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  spn_ogs_plan* p = new (std::nothrow) spn_ogs_plan();  // value-initialised: every pointer null
  if (!p) return SPN_ERR_OUT_OF_MEMORY;
  p->Nrows = Nrows;
  p->nnz = Nrows ? (long long)h_rowStarts[Nrows] - h_rowStarts[0] : 0;
  p->d_rowStarts = d_rowStarts;
  p->d_colIds = d_colIds;
  p->Ncols = Ncols;
  (void)h_colIds;  // reserved for layouts that need the host copy of the ids (none in use: profiles/r01_notes.md)

  // greedy packing of whole consecutive rows into <=32-nnz tiles, by chunks of kPackChunkRows rows on host threads
  // (a tile never spans a chunk boundary: at most one tile more per 2^20 rows than one serial sweep, and the tiling does
  // not depend on the thread count)
  std::vector<OgsTileDesc> tiles;
  std::vector<int> irr;
  try {
    const int nchunks = (Nrows + kPackChunkRows - 1) / kPackChunkRows;
    std::vector<PackOut> parts((size_t)nchunks);
    std::atomic<int> next{0};
    auto worker = [&]() {
      for (int c = next.fetch_add(1); c < nchunks; c = next.fetch_add(1))
        pack_rows(h_rowStarts, c * kPackChunkRows, std::min(Nrows, (c + 1) * kPackChunkRows), parts[c]);
    };
    const int nthreads = std::max(1, std::min({(int)std::thread::hardware_concurrency(), 16, nchunks}));
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(worker);
    worker();
    for (std::thread& t : pool) t.join();
    size_t nt = 0, ni = 0;
    for (const PackOut& o : parts) {
      if (o.bad) { delete p; return o.bad; }
      nt += o.tiles.size();
      ni += o.irr.size();
    }
    tiles.reserve(nt);
    irr.reserve(ni);
    for (const PackOut& o : parts) {
      tiles.insert(tiles.end(), o.tiles.begin(), o.tiles.end());
      irr.insert(irr.end(), o.irr.begin(), o.irr.end());
    }
  } catch (const std::bad_alloc&) {
    delete p;
    return SPN_ERR_OUT_OF_MEMORY;
  } catch (const std::system_error&) {  // no thread could be started
    delete p;
    return SPN_ERR_OUT_OF_MEMORY;
  }
  p->NT = (int)tiles.size();
  p->nIrr = (int)irr.size();
  OgsTileDesc* d_desc = nullptr;
  try {
    rc = build_device_plan(p, tiles, irr, &d_desc);
  } catch (const std::bad_alloc&) {
    rc = SPN_ERR_OUT_OF_MEMORY;
  }
  if (d_desc) cudaFree(d_desc);
  if (rc) {
    spn_ogs_plan_destroy(p);
    return status_of_cuda(rc);
  }
  *plan = p;
  return SPN_SUCCESS;
}

int spn_ogs_plan_destroy(spn_ogs_plan_t plan) {
  if (!plan) return SPN_SUCCESS;
  if (plan->d_colP) cudaFree(plan->d_colP);
  if (plan->d_row0) cudaFree(plan->d_row0);
  if (plan->d_rowOf) cudaFree(plan->d_rowOf);
  if (plan->d_irr) cudaFree(plan->d_irr);
  delete plan;
  return SPN_SUCCESS;
}

int spn_ogs_plan_info(spn_ogs_plan_t plan, int* Ntiles, int* Nirregular, size_t* plan_bytes) {
  if (!plan) return SPN_ERR_INVALID_ARGUMENT;
  if (Ntiles) *Ntiles = plan->NT;
  if (Nirregular) *Nirregular = plan->nIrr;
  if (plan_bytes)
    *plan_bytes = (sizeof(unsigned) * 32 + sizeof(int)) * (size_t)(plan->NT + kTilePad) + sizeof(int) * (size_t)plan->nIrr +
                  sizeof(int) * (size_t)plan->Ncols;
  return SPN_SUCCESS;
}

int spn_ogs_plan_gather(spn_ogs_plan_t plan, int K, const void* v, void* gv, int type, int op, spn_stream_t stream) {
  if (!plan || K < 1) return SPN_ERR_INVALID_ARGUMENT;
  if (plan->Nrows == 0) return SPN_SUCCESS;
  if (!v || !gv) return SPN_ERR_INVALID_ARGUMENT;
  return dispatch_type<TILE_GATHER>(plan, K, v, gv, type, op, as_stream(stream));
}

int spn_ogs_plan_scatter(spn_ogs_plan_t plan, int K, const void* gv, void* v, int type, spn_stream_t stream) {
  if (!plan || K < 1) return SPN_ERR_INVALID_ARGUMENT;
  if (plan->Nrows == 0) return SPN_SUCCESS;
  if (!v || !gv) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  const int64_t Ncols = plan->Ncols;
  if (Ncols == 0) return SPN_SUCCESS;
  const bool wide = (type == SPN_DOUBLE || type == SPN_INT64);
  if (type != SPN_FLOAT && type != SPN_INT32 && !wide) return SPN_ERR_INVALID_ARGUMENT;
  if (K == 1) {
    constexpr int UNR = 2;
    const uintptr_t vp = reinterpret_cast<uintptr_t>(v), rp = reinterpret_cast<uintptr_t>(plan->d_rowOf);
    const bool vec = ((vp & (wide ? 31 : 15)) == 0) && ((rp & 15) == 0);
    const int64_t groups = vec ? Ncols / 4 : Ncols;
    int64_t grid = (groups + 256 * UNR - 1) / (256 * UNR);
    grid = std::max<int64_t>(1, std::min<int64_t>(grid, 0x7fffffff));
    if (wide) {
      using T = unsigned long long;
      if (vec) ogs_scatter_cols_kernel<T, 4, UNR><<<(unsigned)grid, 256, 0, st>>>(Ncols, plan->d_rowOf, static_cast<const T*>(gv), static_cast<T*>(v));
      else ogs_scatter_cols_kernel<T, 1, UNR><<<(unsigned)grid, 256, 0, st>>>(Ncols, plan->d_rowOf, static_cast<const T*>(gv), static_cast<T*>(v));
    } else {
      using T = unsigned int;
      if (vec) ogs_scatter_cols_kernel<T, 4, UNR><<<(unsigned)grid, 256, 0, st>>>(Ncols, plan->d_rowOf, static_cast<const T*>(gv), static_cast<T*>(v));
      else ogs_scatter_cols_kernel<T, 1, UNR><<<(unsigned)grid, 256, 0, st>>>(Ncols, plan->d_rowOf, static_cast<const T*>(gv), static_cast<T*>(v));
    }
  } else {
    const int64_t total = Ncols * K;
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)d->sms * 32));
    if (wide) ogs_scatter_cols_k_kernel<unsigned long long><<<(unsigned)grid, 256, 0, st>>>(Ncols, K, plan->d_rowOf, static_cast<const unsigned long long*>(gv), static_cast<unsigned long long*>(v));
    else ogs_scatter_cols_k_kernel<unsigned int><<<(unsigned)grid, 256, 0, st>>>(Ncols, K, plan->d_rowOf, static_cast<const unsigned int*>(gv), static_cast<unsigned int*>(v));
  }
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

int spn_debug_set(int key, int value) {
  if (key == 4 && value >= 0) { g_chunk_iters = value; return SPN_SUCCESS; }
  if (key == 5 && value >= 0) { g_chunk_iters_gather = value; return SPN_SUCCESS; }
  return SPN_ERR_INVALID_ARGUMENT;
}

int spn_ogs_plan_gather_scatter_to(spn_ogs_plan_t plan, int K, const void* v, void* w, int type, int op, spn_stream_t stream) {
  if (!plan || K < 1) return SPN_ERR_INVALID_ARGUMENT;
  if (plan->Nrows == 0) return SPN_SUCCESS;
  if (!v || !w) return SPN_ERR_INVALID_ARGUMENT;
  return dispatch_type<TILE_GATHER_SCATTER>(plan, K, v, w, type, op, as_stream(stream));
}

int spn_ogs_plan_gather_scatter(spn_ogs_plan_t plan, int K, void* v, int type, int op, spn_stream_t stream) {
  if (!plan || K < 1) return SPN_ERR_INVALID_ARGUMENT;
  if (plan->Nrows == 0) return SPN_SUCCESS;
  if (!v) return SPN_ERR_INVALID_ARGUMENT;
  return dispatch_type<TILE_GATHER_SCATTER>(plan, K, v, v, type, op, as_stream(stream));
}

}  // extern "C"
