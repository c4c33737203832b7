// C-ABI entry points for the ogs kernels (include/spn_b200.h).  `type`/`op`
// select a precompiled instantiation where the reference JIT-compiles one OKL
// kernel per (T, OGS_OP) pair (libs/ogs/ogsUtils.cpp:46-123).
#include <atomic>
#include "ogs_pipe.cuh"

#ifndef SPN_OGS_THREADS
#define SPN_OGS_THREADS 512
#endif

using namespace spn;

namespace {

// Persistent launch of the pipelined kernels: grid.x = min(NrowBlocks, SMs x resident CTAs),
// grid.y = K.  The occupancy query and the opt-in to >48 KB dynamic shared memory run once
// per instantiation and device.
template <typename T, int OP, int MODE>
void launch_pipe(dim3 grid, cudaStream_t st, int K, const int* bs, const int* rs, const int* ci,
                 const T* in, T* out) {
  auto kern = ogs_pipe_kernel<T, OP, MODE, SPN_OGS_THREADS>;
  constexpr size_t smem = ogs_pipe_smem_bytes<T, MODE>();
  static std::atomic<int> resident[64];  // CTAs per device ordinal, 0 = not yet queried (every thread would store the same value)
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (resident[dev] == 0) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int per_sm = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SPN_OGS_THREADS, smem);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    resident[dev] = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 1);
  }
  const int NB = (int)grid.x;
  const int cap = resident[dev].load();
  const dim3 pgrid(NB < cap ? NB : cap, grid.y);
  kern<<<pgrid, SPN_OGS_THREADS, smem, st>>>(NB, K, bs, rs, ci, in, out);
}

template <typename T, int OP>
void launch_gather(dim3 grid, cudaStream_t st, int K, const int* bs, const int* rs, const int* ci,
                   const void* v, void* gv) {
  launch_pipe<T, OP, OGS_GATHER>(grid, st, K, bs, rs, ci, static_cast<const T*>(v), static_cast<T*>(gv));
}

template <typename T>
int dispatch_gather_op(int op, dim3 grid, cudaStream_t st, int K, const int* bs, const int* rs,
                       const int* ci, const void* v, void* gv) {
  switch (op) {
    case SPN_ADD: launch_gather<T, SPN_ADD>(grid, st, K, bs, rs, ci, v, gv); break;
    case SPN_MUL: launch_gather<T, SPN_MUL>(grid, st, K, bs, rs, ci, v, gv); break;
    case SPN_MAX: launch_gather<T, SPN_MAX>(grid, st, K, bs, rs, ci, v, gv); break;
    case SPN_MIN: launch_gather<T, SPN_MIN>(grid, st, K, bs, rs, ci, v, gv); break;
    default: return SPN_ERR_INVALID_ARGUMENT;
  }
  return SPN_SUCCESS;
}

template <typename T, int OP>
void launch_gs(dim3 grid, cudaStream_t st, int K, const int* bs, const int* grs, const int* gci,
               const int* srs, const int* sci, void* v) {
  if (grs == srs && gci == sci)
    launch_pipe<T, OP, OGS_GATHER_SCATTER>(grid, st, K, bs, grs, gci, static_cast<const T*>(v), static_cast<T*>(v));
  else
    gather_scatter_kernel<T, OP><<<grid, kOgsBlock, 0, st>>>(K, bs, grs, gci, srs, sci, static_cast<T*>(v));
}

template <typename T>
int dispatch_gs_op(int op, dim3 grid, cudaStream_t st, int K, const int* bs, const int* grs,
                   const int* gci, const int* srs, const int* sci, void* v) {
  switch (op) {
    case SPN_ADD: launch_gs<T, SPN_ADD>(grid, st, K, bs, grs, gci, srs, sci, v); break;
    case SPN_MUL: launch_gs<T, SPN_MUL>(grid, st, K, bs, grs, gci, srs, sci, v); break;
    case SPN_MAX: launch_gs<T, SPN_MAX>(grid, st, K, bs, grs, gci, srs, sci, v); break;
    case SPN_MIN: launch_gs<T, SPN_MIN>(grid, st, K, bs, grs, gci, srs, sci, v); break;
    default: return SPN_ERR_INVALID_ARGUMENT;
  }
  return SPN_SUCCESS;
}

inline int check_launch_dims(int NrowBlocks, int K) {
  if (NrowBlocks < 0 || K < 1 || K > 65535) return SPN_ERR_INVALID_ARGUMENT;
  return SPN_SUCCESS;
}

}  // namespace

extern "C" {

int spn_ogs_nodes_per_block(void) { return kNodesPerBlock; }

int spn_ogs_gather(int NrowBlocks, int K, const int* blockRowStarts, const int* rowStarts,
                   const int* colIds, const void* v, void* gv, int type, int op, spn_stream_t stream) {
  int rc = check_launch_dims(NrowBlocks, K);
  if (rc) return rc;
  if (NrowBlocks == 0) return SPN_SUCCESS;
  if (!blockRowStarts || !rowStarts || !colIds || !v || !gv) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  rc = device_info(&d);
  if (rc) return rc;
  const dim3 grid(NrowBlocks, K);
  cudaStream_t st = as_stream(stream);
  switch (type) {
    case SPN_FLOAT: rc = dispatch_gather_op<float>(op, grid, st, K, blockRowStarts, rowStarts, colIds, v, gv); break;
    case SPN_DOUBLE: rc = dispatch_gather_op<double>(op, grid, st, K, blockRowStarts, rowStarts, colIds, v, gv); break;
    case SPN_INT32: rc = dispatch_gather_op<int>(op, grid, st, K, blockRowStarts, rowStarts, colIds, v, gv); break;
    case SPN_INT64: rc = dispatch_gather_op<long long>(op, grid, st, K, blockRowStarts, rowStarts, colIds, v, gv); break;
    default: return SPN_ERR_INVALID_ARGUMENT;
  }
  if (rc) return rc;
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

int spn_ogs_scatter(int NrowBlocks, int K, const int* blockRowStarts, const int* rowStarts,
                    const int* colIds, const void* gv, void* v, int type, spn_stream_t stream) {
  int rc = check_launch_dims(NrowBlocks, K);
  if (rc) return rc;
  if (NrowBlocks == 0) return SPN_SUCCESS;
  if (!blockRowStarts || !rowStarts || !colIds || !v || !gv) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  rc = device_info(&d);
  if (rc) return rc;
  const dim3 grid(NrowBlocks, K);
  cudaStream_t st = as_stream(stream);
  switch (type) {
    case SPN_FLOAT:
    case SPN_INT32:
      launch_pipe<unsigned int, SPN_ADD, OGS_SCATTER>(grid, st, K, blockRowStarts, rowStarts, colIds,
          static_cast<const unsigned int*>(gv), static_cast<unsigned int*>(v));
      break;
    case SPN_DOUBLE:
    case SPN_INT64:
      launch_pipe<unsigned long long, SPN_ADD, OGS_SCATTER>(grid, st, K, blockRowStarts, rowStarts, colIds,
          static_cast<const unsigned long long*>(gv), static_cast<unsigned long long*>(v));
      break;
    default: return SPN_ERR_INVALID_ARGUMENT;
  }
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

int spn_ogs_gather_scatter(int NrowBlocks, int K, const int* blockRowStarts,
                           const int* gRowStarts, const int* gColIds,
                           const int* sRowStarts, const int* sColIds,
                           void* v, int type, int op, spn_stream_t stream) {
  int rc = check_launch_dims(NrowBlocks, K);
  if (rc) return rc;
  if (NrowBlocks == 0) return SPN_SUCCESS;
  if (!blockRowStarts || !gRowStarts || !gColIds || !sRowStarts || !sColIds || !v) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  rc = device_info(&d);
  if (rc) return rc;
  const dim3 grid(NrowBlocks, K);
  cudaStream_t st = as_stream(stream);
  switch (type) {
    case SPN_FLOAT: rc = dispatch_gs_op<float>(op, grid, st, K, blockRowStarts, gRowStarts, gColIds, sRowStarts, sColIds, v); break;
    case SPN_DOUBLE: rc = dispatch_gs_op<double>(op, grid, st, K, blockRowStarts, gRowStarts, gColIds, sRowStarts, sColIds, v); break;
    case SPN_INT32: rc = dispatch_gs_op<int>(op, grid, st, K, blockRowStarts, gRowStarts, gColIds, sRowStarts, sColIds, v); break;
    case SPN_INT64: rc = dispatch_gs_op<long long>(op, grid, st, K, blockRowStarts, gRowStarts, gColIds, sRowStarts, sColIds, v); break;
    default: return SPN_ERR_INVALID_ARGUMENT;
  }
  if (rc) return rc;
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

int spn_ogs_extract(int N, int K, const int* ids, const void* q, void* gatherq, int type, spn_stream_t stream) {
  if (N < 0 || K < 1) return SPN_ERR_INVALID_ARGUMENT;
  if (N == 0) return SPN_SUCCESS;
  if (!ids || !q || !gatherq) return SPN_ERR_INVALID_ARGUMENT;
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  const int64_t NK = (int64_t)N * K;
  int64_t g = (NK + kOgsBlock - 1) / kOgsBlock;
  const int64_t cap = (int64_t)d->sms * 32;
  if (g > cap) g = cap;
  cudaStream_t st = as_stream(stream);
  switch (type) {
    case SPN_FLOAT:
    case SPN_INT32:
      extract_kernel<unsigned int><<<(int)g, kOgsBlock, 0, st>>>(NK, K, ids, static_cast<const unsigned int*>(q),
                                                                 static_cast<unsigned int*>(gatherq));
      break;
    case SPN_DOUBLE:
    case SPN_INT64:
      extract_kernel<unsigned long long><<<(int)g, kOgsBlock, 0, st>>>(NK, K, ids, static_cast<const unsigned long long*>(q),
                                                                       static_cast<unsigned long long*>(gatherq));
      break;
    default: return SPN_ERR_INVALID_ARGUMENT;
  }
  SPN_LAUNCH_CHECK();
  return SPN_SUCCESS;
}

}  // extern "C"
