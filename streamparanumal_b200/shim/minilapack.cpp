// The four LAPACK routines the reference's libs/core wrappers call (libs/core/matrixEig.cpp:29-34,
// matrixInverse.cpp:29-35, matrixRightSolve.cpp:31-45), for the small dense matrices of the reference element
// (at most (N+1) x (N+1), N <= 15): the image has no system BLAS/LAPACK.  Column-major, Fortran calling convention.
//   dgeev_   real SYMMETRIC input only (the Jacobi matrix of the Gauss quadrature, libs/mesh/meshBasis1D.cpp:195-231):
//            cyclic Jacobi rotations; eigenvalues in WR (WI = 0), right eigenvectors in the columns of VR.
//            A non-symmetric matrix is refused (INFO = -4).
//   dgetrf_ / dgetri_ / dgesv_   LU with partial pivoting.
// The single-precision entry points convert to double and back.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" {

void dgeev_(char* JOBVL, char* JOBVR, int* N, double* A, int* LDA, double* WR, double* WI, double* VL, int* LDVL, double* VR, int* LDVR,
            double* WORK, int* LWORK, int* INFO) {
  (void)JOBVL; (void)VL; (void)LDVL; (void)WORK; (void)LWORK;
  const int n = *N, lda = *LDA;
  double scale = 0.0;
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) scale = std::fmax(scale, std::fabs(A[i + j * lda]));
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < j; ++i)
      if (std::fabs(A[i + j * lda] - A[j + i * lda]) > 1e-13 * scale) { *INFO = -4; return; }
  std::vector<double> a((size_t)n * n), v((size_t)n * n, 0.0);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) a[i + j * n] = 0.5 * (A[i + j * lda] + A[j + i * lda]);
  for (int i = 0; i < n; ++i) v[i + i * n] = 1.0;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) off += a[p + q * n] * a[p + q * n];
    if (off <= 1e-300 || std::sqrt(off) <= 1e-17 * scale) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = a[p + q * n];
        if (apq == 0.0) continue;
        const double theta = (a[q + q * n] - a[p + p * n]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {  // columns p, q
          const double akp = a[k + p * n], akq = a[k + q * n];
          a[k + p * n] = c * akp - s * akq;
          a[k + q * n] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {  // rows p, q
          const double apk = a[p + k * n], aqk = a[q + k * n];
          a[p + k * n] = c * apk - s * aqk;
          a[q + k * n] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = v[k + p * n], vkq = v[k + q * n];
          v[k + p * n] = c * vkp - s * vkq;
          v[k + q * n] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < n; ++i) {
    WR[i] = a[i + i * n];
    WI[i] = 0.0;
  }
  if (JOBVR && (*JOBVR == 'V' || *JOBVR == 'v') && VR) {
    const int ldvr = *LDVR;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) VR[i + j * ldvr] = v[i + j * n];
  }
  *INFO = 0;
}

void dgetrf_(int* M, int* N, double* A, int* LDA, int* IPIV, int* INFO) {
  const int m = *M, n = *N, lda = *LDA;
  *INFO = 0;
  for (int k = 0; k < (m < n ? m : n); ++k) {
    int piv = k;
    for (int i = k + 1; i < m; ++i)
      if (std::fabs(A[i + k * lda]) > std::fabs(A[piv + k * lda])) piv = i;
    IPIV[k] = piv + 1;
    if (A[piv + k * lda] == 0.0) { if (!*INFO) *INFO = k + 1; continue; }
    if (piv != k)
      for (int j = 0; j < n; ++j) { const double t = A[k + j * lda]; A[k + j * lda] = A[piv + j * lda]; A[piv + j * lda] = t; }
    for (int i = k + 1; i < m; ++i) {
      A[i + k * lda] /= A[k + k * lda];
      const double l = A[i + k * lda];
      for (int j = k + 1; j < n; ++j) A[i + j * lda] -= l * A[k + j * lda];
    }
  }
}

// solves with the LU factors in place: B (n x nrhs, ldb) <- A^{-1} B
static void lu_solve(int n, int nrhs, const double* A, int lda, const int* IPIV, double* B, int ldb) {
  for (int r = 0; r < nrhs; ++r) {
    double* b = B + (size_t)r * ldb;
    for (int k = 0; k < n; ++k) {
      const int p = IPIV[k] - 1;
      if (p != k) { const double t = b[k]; b[k] = b[p]; b[p] = t; }
    }
    for (int k = 0; k < n; ++k)
      for (int i = k + 1; i < n; ++i) b[i] -= A[i + k * lda] * b[k];
    for (int k = n - 1; k >= 0; --k) {
      b[k] /= A[k + k * lda];
      for (int i = 0; i < k; ++i) b[i] -= A[i + k * lda] * b[k];
    }
  }
}

void dgetri_(int* N, double* A, int* LDA, int* IPIV, double* WORK, int* LWORK, int* INFO) {
  (void)WORK; (void)LWORK;
  const int n = *N, lda = *LDA;
  std::vector<double> inv((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) inv[i + (size_t)i * n] = 1.0;
  for (int k = 0; k < n; ++k)
    if (A[k + k * lda] == 0.0) { *INFO = k + 1; return; }
  lu_solve(n, n, A, lda, IPIV, inv.data(), n);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) A[i + j * lda] = inv[i + (size_t)j * n];
  *INFO = 0;
}

void dgesv_(int* N, int* NRHS, double* A, int* LDA, int* IPIV, double* B, int* LDB, int* INFO) {
  dgetrf_(N, N, A, LDA, IPIV, INFO);
  if (*INFO) return;
  lu_solve(*N, *NRHS, A, *LDA, IPIV, B, *LDB);
}

// ---- single precision: through the double routines ----
void sgeev_(char* JOBVL, char* JOBVR, int* N, float* A, int* LDA, float* WR, float* WI, float* VL, int* LDVL, float* VR, int* LDVR,
            float* WORK, int* LWORK, int* INFO) {
  (void)VL; (void)WORK;
  const int n = *N;
  std::vector<double> a((size_t)n * n), wr(n), wi(n), vr((size_t)n * n);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) a[i + (size_t)j * n] = A[i + j * *LDA];
  int ld = n;
  dgeev_(JOBVL, JOBVR, N, a.data(), &ld, wr.data(), wi.data(), nullptr, LDVL, vr.data(), &ld, nullptr, LWORK, INFO);
  for (int i = 0; i < n; ++i) { WR[i] = (float)wr[i]; WI[i] = (float)wi[i]; }
  if (VR && JOBVR && (*JOBVR == 'V' || *JOBVR == 'v'))
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) VR[i + j * *LDVR] = (float)vr[i + (size_t)j * n];
}
void sgetrf_(int* M, int* N, float* A, int* LDA, int* IPIV, int* INFO) {
  const int m = *M, n = *N;
  std::vector<double> a((size_t)m * n);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) a[i + (size_t)j * m] = A[i + j * *LDA];
  int ld = m;
  dgetrf_(M, N, a.data(), &ld, IPIV, INFO);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) A[i + j * *LDA] = (float)a[i + (size_t)j * m];
}
void sgetri_(int* N, float* A, int* LDA, int* IPIV, float* WORK, int* LWORK, int* INFO) {
  (void)WORK;
  const int n = *N;
  std::vector<double> a((size_t)n * n);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) a[i + (size_t)j * n] = A[i + j * *LDA];
  int ld = n;
  dgetri_(N, a.data(), &ld, IPIV, nullptr, LWORK, INFO);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) A[i + j * *LDA] = (float)a[i + (size_t)j * n];
}
void sgesv_(int* N, int* NRHS, float* A, int* LDA, int* IPIV, float* B, int* LDB, int* INFO) {
  const int n = *N, r = *NRHS;
  std::vector<double> a((size_t)n * n), b((size_t)n * r);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) a[i + (size_t)j * n] = A[i + j * *LDA];
  for (int j = 0; j < r; ++j)
    for (int i = 0; i < n; ++i) b[i + (size_t)j * n] = B[i + j * *LDB];
  int ld = n;
  dgesv_(N, NRHS, a.data(), &ld, IPIV, b.data(), &ld, INFO);
  for (int j = 0; j < r; ++j)
    for (int i = 0; i < n; ++i) B[i + j * *LDB] = (float)b[i + (size_t)j * n];
}

}  // extern "C"
