// Halo-side kernels of the pairwise exchange (include/spn_b200.h, "halo" section).
//
// The reference runs four kernels around its MPI exchange (libs/ogs/ogs.cpp:49-125,
// libs/ogs/ogsPairwise.cpp:106-156): halo gather -> extract (pack per neighbour) -> [MPI]
// -> postmpi gather (own partial + received partials) -> halo scatter.  The halo is ~1% of
// the nodes, so these kernels are latency-, not bandwidth-bound, and they sit on the critical
// path before the send and after the receive.  Here they are two launches:
//   spn_ogs_halo_gather_pack      rows -> ownBuf, and directly into the per-neighbour send order
//   spn_ogs_halo_combine_scatter  fold own + received in the reference's order, write every local copy
// One thread per (row, k); every fold runs in CSR order from the op identity, so results are
// bit-identical to the reference's four-kernel sequence.
#include "ogs_kernels.cuh"

using namespace spn;

namespace {

template <typename T, int OP>
__device__ __forceinline__ T fold_row(const int* __restrict__ rowStarts, const int* __restrict__ colIds,
                                      const T* __restrict__ v, int row, int k, int K) {
  T acc = OgsOp<T, OP>::init();
  const int s = __ldg(rowStarts + row), e = __ldg(rowStarts + row + 1);
  for (int g = s; g < e; ++g) OgsOp<T, OP>::apply(acc, v[k + (int64_t)__ldg(colIds + g) * K]);
  return acc;
}

// threads [0, Nrows*K): ownBuf[k + row*K] = fold of halo row `row`
// threads [Nrows*K, (Nrows+Nsend)*K): sendBuf[k + s*K] = fold of halo row sendIds[s]
// (a row sent to m neighbours is folded m+1 times: rows hold <= 8 local copies and the values
//  are L1/L2-resident, which is cheaper than a second dependent launch)
template <typename T, int OP>
__global__ void __launch_bounds__(256)
halo_gather_pack_kernel(int Nrows, int Nsend, int K, const int* __restrict__ rowStarts, const int* __restrict__ colIds,
                        const int* __restrict__ sendIds, const T* __restrict__ v, T* __restrict__ ownBuf,
                        T* __restrict__ sendBuf) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t nOwn = (int64_t)Nrows * K;
  if (i < nOwn) {
    const int row = (int)(i / K), k = (int)(i - (int64_t)row * K);
    ownBuf[i] = fold_row<T, OP>(rowStarts, colIds, v, row, k, K);
  } else {
    const int64_t j = i - nOwn;
    if (j >= (int64_t)Nsend * K) return;
    const int s = (int)(j / K), k = (int)(j - (int64_t)s * K);
    sendBuf[j] = fold_row<T, OP>(rowStarts, colIds, v, __ldg(sendIds + s), k, K);
  }
}

// acc = identity (+) own[row] (row < NrowsP) (+) recv[cmbIds[j]] for j in [cmbStarts[row], cmbStarts[row+1])
// (ascending source rank: libs/ogs/ogsPairwise.cpp:214-232), then
//   sRowStarts != NULL : out[k + sColIds[g]*K] = acc for every g of the row   (halo scatter, ogs.cpp:124)
//   sRowStarts == NULL : out[k + row*K] = acc                                 (gathered layout, ogs.cpp:276)
template <typename T, int OP>
__global__ void __launch_bounds__(256)
halo_combine_scatter_kernel(int Nrows, int NrowsP, int K, const int* __restrict__ cmbStarts,
                            const int* __restrict__ cmbIds, const T* __restrict__ ownBuf,
                            const T* __restrict__ recvBuf, const int* __restrict__ sRowStarts,
                            const int* __restrict__ sColIds, T* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= (int64_t)Nrows * K) return;
  const int row = (int)(i / K), k = (int)(i - (int64_t)row * K);
  T acc = OgsOp<T, OP>::init();
  if (row < NrowsP) OgsOp<T, OP>::apply(acc, ownBuf[i]);
  const int s = __ldg(cmbStarts + row), e = __ldg(cmbStarts + row + 1);
  for (int j = s; j < e; ++j) OgsOp<T, OP>::apply(acc, recvBuf[k + (int64_t)__ldg(cmbIds + j) * K]);
  if (sRowStarts) {
    const int gs = __ldg(sRowStarts + row), ge = __ldg(sRowStarts + row + 1);
    for (int g = gs; g < ge; ++g) out[k + (int64_t)__ldg(sColIds + g) * K] = acc;
  } else {
    out[i] = acc;
  }
}

template <typename T, int OP>
int launch_pack(int Nrows, int Nsend, int K, const int* rs, const int* ci, const int* sendIds, const void* v,
                void* own, void* send, cudaStream_t st) {
  const int64_t total = ((int64_t)Nrows + Nsend) * K;
  const int64_t grid = (total + 255) / 256;
  if (grid > 0x7fffffff) return SPN_ERR_INVALID_ARGUMENT;
  halo_gather_pack_kernel<T, OP><<<(unsigned)grid, 256, 0, st>>>(Nrows, Nsend, K, rs, ci, sendIds,
      static_cast<const T*>(v), static_cast<T*>(own), static_cast<T*>(send));
  return SPN_SUCCESS;
}

template <typename T, int OP>
int launch_combine(int Nrows, int NrowsP, int K, const int* cs, const int* cid, const void* own, const void* recv,
                   const int* srs, const int* sci, void* out, cudaStream_t st) {
  const int64_t total = (int64_t)Nrows * K;
  const int64_t grid = (total + 255) / 256;
  if (grid > 0x7fffffff) return SPN_ERR_INVALID_ARGUMENT;
  halo_combine_scatter_kernel<T, OP><<<(unsigned)grid, 256, 0, st>>>(Nrows, NrowsP, K, cs, cid,
      static_cast<const T*>(own), static_cast<const T*>(recv), srs, sci, static_cast<T*>(out));
  return SPN_SUCCESS;
}

#define SPN_OPS(FN, T, ...)                                                   \
  switch (op) {                                                               \
    case SPN_ADD: rc = FN<T, SPN_ADD>(__VA_ARGS__); break;                    \
    case SPN_MUL: rc = FN<T, SPN_MUL>(__VA_ARGS__); break;                    \
    case SPN_MAX: rc = FN<T, SPN_MAX>(__VA_ARGS__); break;                    \
    case SPN_MIN: rc = FN<T, SPN_MIN>(__VA_ARGS__); break;                    \
    default: return SPN_ERR_INVALID_ARGUMENT;                                 \
  }
#define SPN_DISPATCH_TYPE_OP(FN, ...)                                         \
  switch (type) {                                                             \
    case SPN_FLOAT: SPN_OPS(FN, float, __VA_ARGS__) break;                    \
    case SPN_DOUBLE: SPN_OPS(FN, double, __VA_ARGS__) break;                  \
    case SPN_INT32: SPN_OPS(FN, int, __VA_ARGS__) break;                      \
    case SPN_INT64: SPN_OPS(FN, long long, __VA_ARGS__) break;                \
    default: return SPN_ERR_INVALID_ARGUMENT;                                 \
  }

}  // namespace

extern "C" {

int spn_ogs_halo_gather_pack(int Nrows, int Nsend, int K, const int* rowStarts, const int* colIds, const int* sendIds,
                             const void* v, void* ownBuf, void* sendBuf, int type, int op, spn_stream_t stream) {
  if (Nrows < 0 || Nsend < 0 || K < 1) return SPN_ERR_INVALID_ARGUMENT;
  if (Nrows == 0) return SPN_SUCCESS;  // nothing shared (every P=1 call): no launch, like the reference's NrowBlocks==0 guards
  if (!rowStarts || !colIds || !v || !ownBuf || (Nsend > 0 && (!sendIds || !sendBuf))) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  SPN_DISPATCH_TYPE_OP(launch_pack, Nrows, Nsend, K, rowStarts, colIds, sendIds, v, ownBuf, sendBuf, st)
  if (rc) return rc;
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

int spn_ogs_halo_combine_scatter(int Nrows, int NrowsP, int K, const int* cmbStarts, const int* cmbIds,
                                 const void* ownBuf, const void* recvBuf, const int* sRowStarts, const int* sColIds,
                                 void* out, int type, int op, spn_stream_t stream) {
  if (Nrows < 0 || NrowsP < 0 || K < 1) return SPN_ERR_INVALID_ARGUMENT;
  if (Nrows == 0) return SPN_SUCCESS;
  if (!cmbStarts || !out || (NrowsP > 0 && !ownBuf) || (sRowStarts && !sColIds)) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  SPN_DISPATCH_TYPE_OP(launch_combine, Nrows, NrowsP, K, cmbStarts, cmbIds, ownBuf, recvBuf, sRowStarts, sColIds, out, st)
  if (rc) return rc;
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

}  // extern "C"
