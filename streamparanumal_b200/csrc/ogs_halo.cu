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

// ---- peer-memory variants (include/spn_b200.h, "Peer-memory halo exchange") ------------------------------
// after every thread of the launch has issued its peer stores: the last block raises the flags.
// stores -> fence.sys -> block barrier -> counter (gpu scope) ... last block: fence.sys -> release stores.
__device__ __forceinline__ void p2p_signal(const spn_p2p_send_t& t) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int old = atomicInc(t.counter, gridDim.x - 1);  // wraps to 0 after the last block: re-armed
    if (old == gridDim.x - 1) {
      __threadfence_system();
      for (int i = 0; i < t.npeers; ++i)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(t.flag[i]), "l"(t.epoch) : "memory");
    }
  }
}

__device__ __forceinline__ void p2p_wait(const spn_p2p_wait_t& w) {
  if (threadIdx.x < w.nsrc) {
    unsigned long long seen;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(w.flag[threadIdx.x]) : "memory");
    } while (seen < w.epoch);
  }
  __syncthreads();
}

template <typename T, int OP>
__global__ void __launch_bounds__(256)
halo_gather_pack_p2p_kernel(int Nrows, int Nsend, int K, const int* __restrict__ rowStarts, const int* __restrict__ colIds,
                            const int* __restrict__ sendIds, const T* __restrict__ v, T* __restrict__ ownBuf,
                            const __grid_constant__ spn_p2p_send_t send) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t nOwn = (int64_t)Nrows * K;
  if (i < nOwn) {
    const int row = (int)(i / K), k = (int)(i - (int64_t)row * K);
    ownBuf[i] = fold_row<T, OP>(rowStarts, colIds, v, row, k, K);
  } else {
    const int64_t j = i - nOwn;
    if (j < (int64_t)Nsend * K) {
      const int s = (int)(j / K), k = (int)(j - (int64_t)s * K);
      const T x = fold_row<T, OP>(rowStarts, colIds, v, __ldg(sendIds + s), k, K);
      // dst[] is in units of slots of K elements
      int lo = 0, hi = send.npeers - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (send.slot0[mid] <= s) lo = mid; else hi = mid - 1;
      }
      static_cast<T*>(send.dst[lo])[(int64_t)(s - send.slot0[lo]) * K + k] = x;
    }
  }
  p2p_signal(send);
}

template <typename U>
__global__ void __launch_bounds__(256)
extract_p2p_kernel(int64_t NK, int K, const int* __restrict__ ids, const U* __restrict__ q,
                   const __grid_constant__ spn_p2p_send_t send) {
  const int64_t n = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (n < NK) {
    const int64_t s = n / K;
    const int k = (int)(n - s * K);
    const U x = __ldg(q + (k + (int64_t)__ldg(ids + s) * K));
    int lo = 0, hi = send.npeers - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (send.slot0[mid] <= (int)s) lo = mid; else hi = mid - 1;
    }
    static_cast<U*>(send.dst[lo])[(s - send.slot0[lo]) * K + k] = x;
  }
  p2p_signal(send);
}

// acc = identity (+) own[row] (row < NrowsP) (+) recv[cmbIds[j]] for j in [cmbStarts[row], cmbStarts[row+1])
// (ascending source rank: libs/ogs/ogsPairwise.cpp:214-232), then
//   sRowStarts != NULL : out[k + sColIds[g]*K] = acc for every g of the row   (halo scatter, ogs.cpp:124)
//   sRowStarts == NULL : out[k + row*K] = acc                                 (gathered layout, ogs.cpp:276)
template <typename T, int OP, bool WAIT>
__global__ void __launch_bounds__(256)
halo_combine_scatter_kernel(int Nrows, int NrowsP, int K, const int* __restrict__ cmbStarts,
                            const int* __restrict__ cmbIds, const T* __restrict__ ownBuf,
                            const T* recvBuf, const int* __restrict__ sRowStarts,
                            const int* __restrict__ sColIds, T* __restrict__ out, const __grid_constant__ spn_p2p_wait_t wait) {
  if (WAIT) p2p_wait(wait);  // the receive buffer is written by peer GPUs: plain (coherent) loads below, no .nc
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
  spn_p2p_wait_t none;
  none.nsrc = 0;
  none.epoch = 0;
  halo_combine_scatter_kernel<T, OP, false><<<(unsigned)grid, 256, 0, st>>>(Nrows, NrowsP, K, cs, cid,
      static_cast<const T*>(own), static_cast<const T*>(recv), srs, sci, static_cast<T*>(out), none);
  return SPN_SUCCESS;
}

template <typename T, int OP>
int launch_combine_p2p(int Nrows, int NrowsP, int K, const int* cs, const int* cid, const void* own, const void* recv,
                       const spn_p2p_wait_t* wait, const int* srs, const int* sci, void* out, cudaStream_t st) {
  const int64_t total = (int64_t)Nrows * K;
  const int64_t grid = (total + 255) / 256;
  if (grid > 0x7fffffff) return SPN_ERR_INVALID_ARGUMENT;
  halo_combine_scatter_kernel<T, OP, true><<<(unsigned)grid, 256, 0, st>>>(Nrows, NrowsP, K, cs, cid,
      static_cast<const T*>(own), static_cast<const T*>(recv), srs, sci, static_cast<T*>(out), *wait);
  return SPN_SUCCESS;
}

template <typename T, int OP>
int launch_pack_p2p(int Nrows, int Nsend, int K, const int* rs, const int* ci, const int* sendIds, const void* v,
                    void* own, const spn_p2p_send_t* send, cudaStream_t st) {
  const int64_t total = ((int64_t)Nrows + Nsend) * K;
  const int64_t grid = (total + 255) / 256;
  if (grid > 0x7fffffff) return SPN_ERR_INVALID_ARGUMENT;
  halo_gather_pack_p2p_kernel<T, OP><<<(unsigned)grid, 256, 0, st>>>(Nrows, Nsend, K, rs, ci, sendIds,
      static_cast<const T*>(v), static_cast<T*>(own), *send);
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


static int check_send(const spn_p2p_send_t* send, int Nsend) {
  if (!send || send->npeers < 0 || send->npeers > SPN_P2P_MAX_PEERS || !send->counter) return SPN_ERR_INVALID_ARGUMENT;
  if (send->npeers == 0) return Nsend == 0 ? SPN_SUCCESS : SPN_ERR_INVALID_ARGUMENT;
  if (send->slot0[0] != 0 || send->slot0[send->npeers] != Nsend) return SPN_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < send->npeers; ++i)
    if (!send->dst[i] || !send->flag[i] || send->slot0[i + 1] < send->slot0[i]) return SPN_ERR_INVALID_ARGUMENT;
  return SPN_SUCCESS;
}

int spn_ogs_halo_gather_pack_p2p(int Nrows, int Nsend, int K, const int* rowStarts, const int* colIds, const int* sendIds,
                                 const void* v, void* ownBuf, const spn_p2p_send_t* send, int type, int op, spn_stream_t stream) {
  if (Nrows < 0 || Nsend < 0 || K < 1) return SPN_ERR_INVALID_ARGUMENT;
  int rc = check_send(send, Nsend);
  if (rc) return rc;
  if (Nrows == 0 && send->npeers == 0) return SPN_SUCCESS;
  if (Nrows > 0 && (!rowStarts || !colIds || !v || !ownBuf || (Nsend > 0 && !sendIds))) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  rc = device_info(&d);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  SPN_DISPATCH_TYPE_OP(launch_pack_p2p, Nrows, Nsend, K, rowStarts, colIds, sendIds, v, ownBuf, send, st)
  if (rc) return rc;
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

int spn_ogs_extract_p2p(int N, int K, const int* ids, const void* q, const spn_p2p_send_t* send, int type, spn_stream_t stream) {
  if (N < 0 || K < 1) return SPN_ERR_INVALID_ARGUMENT;
  int rc = check_send(send, N);
  if (rc) return rc;
  if (send->npeers == 0) return SPN_SUCCESS;
  if (N > 0 && (!ids || !q)) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  rc = device_info(&d);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  const int64_t NK = (int64_t)N * K;
  const int64_t grid = (NK + 255) / 256 > 0 ? (NK + 255) / 256 : 1;
  if (grid > 0x7fffffff) return SPN_ERR_INVALID_ARGUMENT;
  if (type == SPN_FLOAT || type == SPN_INT32)
    extract_p2p_kernel<unsigned int><<<(unsigned)grid, 256, 0, st>>>(NK, K, ids, static_cast<const unsigned int*>(q), *send);
  else if (type == SPN_DOUBLE || type == SPN_INT64)
    extract_p2p_kernel<unsigned long long><<<(unsigned)grid, 256, 0, st>>>(NK, K, ids, static_cast<const unsigned long long*>(q), *send);
  else
    return SPN_ERR_INVALID_ARGUMENT;
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

int spn_ogs_halo_combine_scatter_p2p(int Nrows, int NrowsP, int K, const int* cmbStarts, const int* cmbIds,
                                     const void* ownBuf, const void* recvBuf, const spn_p2p_wait_t* wait,
                                     const int* sRowStarts, const int* sColIds, void* out, int type, int op, spn_stream_t stream) {
  if (Nrows < 0 || NrowsP < 0 || K < 1 || !wait || wait->nsrc < 0 || wait->nsrc > SPN_P2P_MAX_PEERS) return SPN_ERR_INVALID_ARGUMENT;
  if (Nrows == 0) return SPN_SUCCESS;
  if (!cmbStarts || !out || (NrowsP > 0 && !ownBuf) || (sRowStarts && !sColIds)) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  SPN_DISPATCH_TYPE_OP(launch_combine_p2p, Nrows, NrowsP, K, cmbStarts, cmbIds, ownBuf, recvBuf, wait, sRowStarts, sColIds, out, st)
  if (rc) return rc;
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

}  // extern "C"
