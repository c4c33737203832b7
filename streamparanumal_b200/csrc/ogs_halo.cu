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
#include <algorithm>

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
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long x;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(x) : "l"(p) : "memory");
  return x;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long x) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(x) : "memory");
}

// Start of a sending launch: the epoch of this call (every block reads the same value: the last block advances it only
// after all blocks have passed their counter increment), and flow control — no store into half h of a destination
// before that destination has acknowledged this rank's previous send into the same half.
__device__ __forceinline__ unsigned long long p2p_begin_send(const spn_p2p_plan_t* __restrict__ plan, const spn_p2p_state_t* state) {
  const unsigned long long e = *reinterpret_cast<const volatile unsigned long long*>(&state->epoch) + 1ull;
  const int h = (int)(e & 1ull);
  if ((int)threadIdx.x < plan->nsend) {
    const unsigned long long need = *reinterpret_cast<const volatile unsigned long long*>(&state->lastSend[plan->dstRank[threadIdx.x]][h]);
    const unsigned long long* ack = plan->ackLocal[threadIdx.x];
    while (ld_acquire_sys(ack) < need) {}
  }
  __syncthreads();
  return e;
}

// End of a sending launch, after every thread has issued its peer stores: the last block raises the flags and advances
// the state.  stores -> fence.sys -> block barrier -> counter (gpu scope) ... last block: fence.sys -> release stores,
// one destination per thread (a single thread storing to 7 peers one after the other costs a fabric round trip each).
__device__ __forceinline__ void p2p_end_send(const spn_p2p_plan_t* __restrict__ plan, spn_p2p_state_t* state, unsigned long long e) {
  __shared__ int s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int old = atomicInc(&state->packCounter, gridDim.x - 1);  // wraps to 0 after the last block: re-armed
    s_last = (old == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  const int h = (int)(e & 1ull);
  if ((int)threadIdx.x < plan->nsend) {
    st_release_sys(plan->flagRemote[threadIdx.x], e);
    state->lastSend[plan->dstRank[threadIdx.x]][h] = e;
  }
  if (threadIdx.x == 0) state->epoch = e;
}

// one thread per (halo row, k): fold once, keep the partial in ownBuf, fan it out to the row's send slots
template <typename T, int OP>
__global__ void __launch_bounds__(256)
halo_gather_pack_p2p_kernel(int Nrows, int K, const int* __restrict__ rowStarts, const int* __restrict__ colIds,
                            const int* __restrict__ rowSendStarts, const int* __restrict__ rowSendDst, const T* __restrict__ v,
                            T* __restrict__ ownBuf, const spn_p2p_plan_t* __restrict__ plan, spn_p2p_state_t* state) {
  const unsigned long long e = p2p_begin_send(plan, state);
  const int h = (int)(e & 1ull);
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < (int64_t)Nrows * K) {
    const int row = (int)(i / K), k = (int)(i - (int64_t)row * K);
    const T x = fold_row<T, OP>(rowStarts, colIds, v, row, k, K);
    ownBuf[i] = x;
    const int s0 = __ldg(rowSendStarts + row), s1 = __ldg(rowSendStarts + row + 1);
    for (int j = s0; j < s1; ++j) {
      const unsigned w = (unsigned)__ldg(rowSendDst + j);
      const int peer = (int)(w >> 24);
      const int64_t slot = plan->remoteOffset[peer] + (int64_t)(w & 0xffffffu);
      static_cast<T*>(plan->dstHalf[h][peer])[slot * K + k] = x;
    }
  }
  p2p_end_send(plan, state, e);
}

template <typename U>
__global__ void __launch_bounds__(256)
extract_p2p_kernel(int64_t NK, int K, const int* __restrict__ ids, const int* __restrict__ slotDst, const U* __restrict__ q,
                   const spn_p2p_plan_t* __restrict__ plan, spn_p2p_state_t* state) {
  const unsigned long long e = p2p_begin_send(plan, state);
  const int h = (int)(e & 1ull);
  const int64_t n = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (n < NK) {
    const int64_t s = n / K;
    const int k = (int)(n - s * K);
    const U x = __ldg(q + (k + (int64_t)__ldg(ids + s) * K));
    const unsigned w = (unsigned)__ldg(slotDst + s);
    const int peer = (int)(w >> 24);
    const int64_t slot = plan->remoteOffset[peer] + (int64_t)(w & 0xffffffu);
    static_cast<U*>(plan->dstHalf[h][peer])[slot * K + k] = x;
  }
  p2p_end_send(plan, state, e);
}

// acc = identity (+) own[row] (row < NrowsP) (+) recv[cmbIds[j]] for j in [cmbStarts[row], cmbStarts[row+1])
// (ascending source rank: libs/ogs/ogsPairwise.cpp:214-232), then
//   sRowStarts != NULL : out[k + sColIds[g]*K] = acc for every g of the row   (halo scatter, ogs.cpp:124)
//   sRowStarts == NULL : out[k + row*K] = acc                                 (gathered layout, ogs.cpp:276)
template <typename T, int OP, bool P2P>
__global__ void __launch_bounds__(256)
halo_combine_scatter_kernel(int Nrows, int NrowsP, int K, const int* __restrict__ cmbStarts,
                            const int* __restrict__ cmbIds, const T* __restrict__ ownBuf,
                            const T* recvBuf, const int* __restrict__ sRowStarts,
                            const int* __restrict__ sColIds, T* __restrict__ out, const spn_p2p_plan_t* __restrict__ plan,
                            spn_p2p_state_t* state) {
  unsigned long long e = 0;
  if (P2P) {
    // the pack kernel of this call ran earlier on this stream: the epoch is the current one
    e = *reinterpret_cast<const volatile unsigned long long*>(&state->epoch);
    if ((int)threadIdx.x < plan->nrecv) {
      const unsigned long long* f = plan->flagLocal[threadIdx.x];
      while (ld_acquire_sys(f) < e) {}
    }
    __syncthreads();
    recvBuf = static_cast<const T*>(plan->recvHalf[e & 1ull]);  // written by peer GPUs: plain (coherent) loads below, no .nc
  }
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < (int64_t)Nrows * K) {
    const int row = (int)(i / K), k = (int)(i - (int64_t)row * K);
    T acc = OgsOp<T, OP>::init();
    if (row < NrowsP) OgsOp<T, OP>::apply(acc, ownBuf[i]);
    const int s = __ldg(cmbStarts + row), e2 = __ldg(cmbStarts + row + 1);
    for (int j = s; j < e2; ++j) OgsOp<T, OP>::apply(acc, recvBuf[k + (int64_t)__ldg(cmbIds + j) * K]);
    if (sRowStarts) {
      const int gs = __ldg(sRowStarts + row), ge = __ldg(sRowStarts + row + 1);
      for (int g = gs; g < ge; ++g) out[k + (int64_t)__ldg(sColIds + g) * K] = acc;
    } else {
      out[i] = acc;
    }
  }
  if (P2P) {
    // every block has read its part of the receive half: the last one acknowledges the sources, one per thread
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int old = atomicInc(&state->combineCounter, gridDim.x - 1);
      s_last = (old == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && (int)threadIdx.x < plan->nrecv) {
      __threadfence_system();
      st_release_sys(plan->ackRemote[threadIdx.x], e);
    }
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
  halo_combine_scatter_kernel<T, OP, false><<<(unsigned)grid, 256, 0, st>>>(Nrows, NrowsP, K, cs, cid,
      static_cast<const T*>(own), static_cast<const T*>(recv), srs, sci, static_cast<T*>(out), nullptr, nullptr);
  return SPN_SUCCESS;
}

template <typename T, int OP>
int launch_combine_p2p(int Nrows, int NrowsP, int K, const int* cs, const int* cid, const void* own,
                       const spn_p2p_plan_t* plan, spn_p2p_state_t* state, const int* srs, const int* sci, void* out, cudaStream_t st) {
  const int64_t total = (int64_t)Nrows * K;
  const int64_t grid = std::max<int64_t>(1, (total + 255) / 256);
  if (grid > 0x7fffffff) return SPN_ERR_INVALID_ARGUMENT;
  halo_combine_scatter_kernel<T, OP, true><<<(unsigned)grid, 256, 0, st>>>(Nrows, NrowsP, K, cs, cid,
      static_cast<const T*>(own), nullptr, srs, sci, static_cast<T*>(out), plan, state);
  return SPN_SUCCESS;
}

template <typename T, int OP>
int launch_pack_p2p(int Nrows, int K, const int* rs, const int* ci, const int* rss, const int* rsd, const void* v,
                    void* own, const spn_p2p_plan_t* plan, spn_p2p_state_t* state, cudaStream_t st) {
  const int64_t total = (int64_t)Nrows * K;
  const int64_t grid = std::max<int64_t>(1, (total + 255) / 256);
  if (grid > 0x7fffffff) return SPN_ERR_INVALID_ARGUMENT;
  halo_gather_pack_p2p_kernel<T, OP><<<(unsigned)grid, 256, 0, st>>>(Nrows, K, rs, ci, rss, rsd,
      static_cast<const T*>(v), static_cast<T*>(own), plan, state);
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


int spn_ogs_halo_gather_pack_p2p(int Nrows, int K, const int* rowStarts, const int* colIds, const int* rowSendStarts,
                                 const int* rowSendDst, const void* v, void* ownBuf, const spn_p2p_plan_t* d_plan,
                                 spn_p2p_state_t* d_state, int type, int op, spn_stream_t stream) {
  if (Nrows < 0 || K < 1 || !d_plan || !d_state) return SPN_ERR_INVALID_ARGUMENT;
  if (Nrows > 0 && (!rowStarts || !colIds || !rowSendStarts || !v || !ownBuf)) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  SPN_DISPATCH_TYPE_OP(launch_pack_p2p, Nrows, K, rowStarts, colIds, rowSendStarts, rowSendDst, v, ownBuf, d_plan, d_state, st)
  if (rc) return rc;
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

int spn_ogs_extract_p2p(int N, int K, const int* ids, const int* slotDst, const void* q, const spn_p2p_plan_t* d_plan,
                        spn_p2p_state_t* d_state, int type, spn_stream_t stream) {
  if (N < 0 || K < 1 || !d_plan || !d_state) return SPN_ERR_INVALID_ARGUMENT;
  if (N > 0 && (!ids || !slotDst || !q)) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  const int64_t NK = (int64_t)N * K;
  const int64_t grid = std::max<int64_t>(1, (NK + 255) / 256);
  if (grid > 0x7fffffff) return SPN_ERR_INVALID_ARGUMENT;
  if (type == SPN_FLOAT || type == SPN_INT32)
    extract_p2p_kernel<unsigned int><<<(unsigned)grid, 256, 0, st>>>(NK, K, ids, slotDst, static_cast<const unsigned int*>(q), d_plan, d_state);
  else if (type == SPN_DOUBLE || type == SPN_INT64)
    extract_p2p_kernel<unsigned long long><<<(unsigned)grid, 256, 0, st>>>(NK, K, ids, slotDst, static_cast<const unsigned long long*>(q), d_plan, d_state);
  else
    return SPN_ERR_INVALID_ARGUMENT;
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

int spn_ogs_halo_combine_scatter_p2p(int Nrows, int NrowsP, int K, const int* cmbStarts, const int* cmbIds,
                                     const void* ownBuf, const spn_p2p_plan_t* d_plan, spn_p2p_state_t* d_state,
                                     const int* sRowStarts, const int* sColIds, void* out, int type, int op, spn_stream_t stream) {
  if (Nrows < 0 || NrowsP < 0 || K < 1 || !d_plan || !d_state) return SPN_ERR_INVALID_ARGUMENT;
  if (Nrows > 0 && (!cmbStarts || !out || (NrowsP > 0 && !ownBuf) || (sRowStarts && !sColIds))) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  SPN_DISPATCH_TYPE_OP(launch_combine_p2p, Nrows, NrowsP, K, cmbStarts, cmbIds, ownBuf, d_plan, d_state, sRowStarts, sColIds, out, st)
  if (rc) return rc;
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

}  // extern "C"
