/*
 * stream_oracle.c — TEST INFRASTRUCTURE ONLY (never linked or loaded by the product).
 *
 * CPU restatement of streamParanumal's BS1–BS5 kernels with the arithmetic and the
 * summation order of the reference's *Serial* path (OCCA Serial mode, "-O3", no FMA
 * contraction; this file is built with -ffp-contract=off).  Each function cites the
 * reference lines it follows.  Pinned against the reference's own OKL sources
 * compiled for the host (oracle/_ref/libref_okl.so) by tests/test_oracle_pin.py —
 * the reference ships no golden vectors for this path.
 *
 * The two-pass reductions keep the reference structure: Nblocks partial sums, thread
 * t of block b accumulates entries t + b*BS + j*BS*Nblocks, then a shared-memory
 * tree BS/2 .. 2 and a final s[0]+s[1] (BS/BS3/okl/bs3.okl:35-63).
 */
#include <stdlib.h>
#include <string.h>

/* tree over s[0..bs): the unrolled ladder of BS/BS3/okl/bs3.okl:46-63 (and the same
 * lines in bs4.okl / bs5.okl), valid for power-of-two bs >= 4 */
static double tree_sum(double* s, int bs) {
  for (int half = bs / 2; half >= 2; half /= 2)
    for (int t = 0; t < half; ++t) s[t] += s[t + half];
  return s[0] + s[1];
}

/* BS/BS1/okl/bs1.okl:27-34 */
void oracle_bs1(int N, const double* a, double* b) {
  for (int n = 0; n < N; ++n) b[n] = a[n];
}

/* BS/BS2/okl/bs2.okl:27-36 */
void oracle_bs2(int N, double alpha, const double* x, double beta, double* y) {
  for (int n = 0; n < N; ++n) y[n] = alpha * x[n] + beta * y[n];
}

/* BS/BS3/okl/bs3.okl:27-64 */
void oracle_bs3_1(int bs, int Nblocks, int N, const double* a, double* norm) {
  double* s = (double*)malloc(sizeof(double) * (size_t)bs);
  for (int b = 0; b < Nblocks; ++b) {
    for (int t = 0; t < bs; ++t) {
      double r = 0.0;
      for (long id = t + (long)b * bs; id < N; id += (long)bs * Nblocks) r += a[id] * a[id];
      s[t] = r;
    }
    norm[b] = tree_sum(s, bs);
  }
  free(s);
}

/* BS/BS4/okl/bs4.okl:27-65 */
void oracle_bs4_1(int bs, int Nblocks, int N, const double* a, const double* bb, double* dot) {
  double* s = (double*)malloc(sizeof(double) * (size_t)bs);
  for (int b = 0; b < Nblocks; ++b) {
    for (int t = 0; t < bs; ++t) {
      double r = 0.0;
      for (long id = t + (long)b * bs; id < N; id += (long)bs * Nblocks) r += a[id] * bb[id];
      s[t] = r;
    }
    dot[b] = tree_sum(s, bs);
  }
  free(s);
}

/* BS/BS5/okl/bs5.okl:27-67 */
void oracle_bs5_1(int bs, int Nblocks, int N, const double* Ap, double alpha, double* r, double* rdotr) {
  double* s = (double*)malloc(sizeof(double) * (size_t)bs);
  for (int b = 0; b < Nblocks; ++b) {
    for (int t = 0; t < bs; ++t) {
      double acc = 0.0;
      for (long id = t + (long)b * bs; id < N; id += (long)bs * Nblocks) {
        const double rr = r[id] + alpha * Ap[id];
        r[id] = rr;
        acc += rr * rr;
      }
      s[t] = acc;
    }
    rdotr[b] = tree_sum(s, bs);
  }
  free(s);
}

/* second pass, identical in bs3_2 / bs4_2 / bs5_2 (BS/BS3/okl/bs3.okl:66-102):
 * one block, thread t sums a[t], a[t+bs*Nblocks], ... (the stride is bs*Nblocks, so
 * with Nblocks<=bs each thread reads at most one entry) */
void oracle_final(int bs, int Nblocks, const double* a, double* out) {
  double* s = (double*)malloc(sizeof(double) * (size_t)bs);
  for (int t = 0; t < bs; ++t) {
    double r = 0.0;
    for (long id = t; id < Nblocks; id += (long)bs * Nblocks) r += a[id];
    s[t] = r;
  }
  out[0] = tree_sum(s, bs);
  free(s);
}

/* launch geometry of the drivers: Nblock = min(ceil(N/bs), bs)
 * (BS/BS3/src/bs3Run.cpp:80-81, bs4Run.cpp:80-81, bs5Run.cpp:84-85) */
static int driver_nblock(int N, int bs) {
  int nb = (N + bs - 1) / bs;
  return nb > bs ? bs : nb;
}

/* whole BS3/BS4/BS5 call as the driver issues it (kernel1 then kernel2) */
double oracle_norm2(int bs, int N, const double* a) {
  const int nb = driver_nblock(N, bs);
  double* tmp = (double*)calloc((size_t)bs, sizeof(double));
  double out = 0.0;
  oracle_bs3_1(bs, nb, N, a, tmp);
  oracle_final(bs, nb, tmp, &out);
  free(tmp);
  return out;
}
double oracle_dot(int bs, int N, const double* a, const double* b) {
  const int nb = driver_nblock(N, bs);
  double* tmp = (double*)calloc((size_t)bs, sizeof(double));
  double out = 0.0;
  oracle_bs4_1(bs, nb, N, a, b, tmp);
  oracle_final(bs, nb, tmp, &out);
  free(tmp);
  return out;
}
double oracle_cg_update(int bs, int N, const double* Ap, double alpha, double* r) {
  const int nb = driver_nblock(N, bs);
  double* tmp = (double*)calloc((size_t)bs, sizeof(double));
  double out = 0.0;
  oracle_bs5_1(bs, nb, N, Ap, alpha, r, tmp);
  oracle_final(bs, nb, tmp, &out);
  free(tmp);
  return out;
}
