/*
 * ogs_oracle.c — TEST INFRASTRUCTURE ONLY (never linked or loaded by the product).
 *
 * CPU restatement of the ogs operator kernels, following the reference's host
 * twins (same CSR semantics and summation order as the OKL kernels):
 *   gather         libs/ogs/ogsOperator.cpp:61-114   (kernel: libs/ogs/okl/ogsKernels.okl:86-122)
 *   scatter        libs/ogs/ogsOperator.cpp:223-267  (kernel: ogsKernels.okl:127-164)
 *   gatherScatter  libs/ogs/ogsOperator.cpp:340-409  (kernel: ogsKernels.okl:32-81)
 *   extract        libs/ogs/ogsOperator.cpp:693-718  (kernel: ogsKernels.okl:167-177)
 *   op identities  libs/ogs/ogsOperator.cpp:37-56
 * A row value starts at the op identity and folds the row's columns in CSR order.
 * Pinned against the reference OKL kernels compiled for the host
 * (oracle/_ref/libref_okl.so) by tests/test_oracle_pin.py.
 *
 * `threads` > 1 runs rows in an OpenMP parallel-for exactly where the reference
 * host twins have one (ogsOperator.cpp:88,245,375); results do not depend on it.
 */
#include <float.h>
#include <limits.h>
#include <omp.h>

enum { OP_ADD = 0, OP_MUL = 1, OP_MAX = 2, OP_MIN = 3 }; /* ogs::Op, include/ogs.hpp:187 */

#define DEFINE_OGS(T, TAG, TMAX)                                                              \
  static inline T init_##TAG(int op) {                                                        \
    switch (op) {                                                                             \
      case OP_ADD: return (T)0;                                                               \
      case OP_MUL: return (T)1;                                                               \
      case OP_MAX: return -(TMAX);                                                            \
      default: return (TMAX);                                                                 \
    }                                                                                         \
  }                                                                                           \
  static inline T fold_##TAG(int op, T gv, T v) {                                             \
    switch (op) {                                                                             \
      case OP_ADD: return gv + v;                                                             \
      case OP_MUL: return gv * v;                                                             \
      case OP_MAX: return (v > gv) ? v : gv;                                                  \
      default: return (v < gv) ? v : gv;                                                      \
    }                                                                                         \
  }                                                                                           \
  void oracle_ogs_gather_##TAG(int Nrows, int K, const int* rowStarts, const int* colIds,     \
                               const T* v, T* gv, int op, int threads) {                      \
    _Pragma("omp parallel for num_threads(threads)")                                          \
    for (int n = 0; n < Nrows; ++n) {                                                         \
      for (int k = 0; k < K; ++k) {                                                           \
        T val = init_##TAG(op);                                                               \
        for (int g = rowStarts[n]; g < rowStarts[n + 1]; ++g)                                 \
          val = fold_##TAG(op, val, v[k + (long)colIds[g] * K]);                              \
        gv[k + (long)n * K] = val;                                                            \
      }                                                                                       \
    }                                                                                         \
  }                                                                                           \
  void oracle_ogs_scatter_##TAG(int Nrows, int K, const int* rowStarts, const int* colIds,    \
                                const T* gv, T* v, int threads) {                             \
    _Pragma("omp parallel for num_threads(threads)")                                          \
    for (int n = 0; n < Nrows; ++n)                                                           \
      for (int g = rowStarts[n]; g < rowStarts[n + 1]; ++g)                                   \
        for (int k = 0; k < K; ++k) v[k + (long)colIds[g] * K] = gv[k + (long)n * K];         \
  }                                                                                           \
  void oracle_ogs_gather_scatter_##TAG(int Nrows, int K, const int* gRowStarts,               \
                                       const int* gColIds, const int* sRowStarts,             \
                                       const int* sColIds, T* v, int op, int threads) {       \
    _Pragma("omp parallel for num_threads(threads)")                                          \
    for (int n = 0; n < Nrows; ++n) {                                                         \
      for (int k = 0; k < K; ++k) {                                                           \
        T val = init_##TAG(op);                                                               \
        for (int g = gRowStarts[n]; g < gRowStarts[n + 1]; ++g)                               \
          val = fold_##TAG(op, val, v[k + (long)gColIds[g] * K]);                             \
        for (int s = sRowStarts[n]; s < sRowStarts[n + 1]; ++s)                               \
          v[k + (long)sColIds[s] * K] = val;                                                  \
      }                                                                                       \
    }                                                                                         \
  }                                                                                           \
  void oracle_ogs_extract_##TAG(int N, int K, const int* ids, const T* q, T* out) {           \
    for (int n = 0; n < N; ++n)                                                               \
      for (int k = 0; k < K; ++k) out[k + (long)n * K] = q[k + (long)ids[n] * K];             \
  }

DEFINE_OGS(float, float, FLT_MAX)
DEFINE_OGS(double, double, DBL_MAX)
DEFINE_OGS(int, int32, INT_MAX)
DEFINE_OGS(long long, int64, LLONG_MAX)
