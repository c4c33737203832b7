// Sweep of launch geometries for the BS1-BS5 kernels (same templates the library
// instantiates).  Build: make tune ; run on the B200: build/tune_stream [log2 n]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../streamparanumal_b200/csrc/stream_kernels.cuh"

namespace spn { int device_info(DeviceInfo**) { return 0; } }
using namespace spn;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

static int g_sms = 148;
static cudaEvent_t e0, e1;

template <class F>
float time_ms(F launch, int iters = 20) {
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGetLastError());
  return ms / iters;
}

static int grid_for(int64_t n, int vec, int block, int unroll, int ctas_per_sm) {
  int64_t chunk = (int64_t)block * unroll;
  int64_t g = (n / vec + chunk - 1) / chunk;
  if (ctas_per_sm > 0 && g > (int64_t)g_sms * ctas_per_sm) g = (int64_t)g_sms * ctas_per_sm;
  return (int)g;
}

template <int VEC, int UNROLL, int BLOCK>
void sweep(int64_t n, double* a, double* b, ReduceWs* ws, double* out) {
  const int cps[] = {0, 1, 2, 4, 8, 16, 32};
  for (int c : cps) {
    if (c > 0 && (int64_t)c * BLOCK > 2048) continue;
    int g = grid_for(n, VEC, BLOCK, UNROLL, c);
    float t1 = time_ms([&] { copy_kernel<VEC, UNROLL, BLOCK><<<g, BLOCK>>>(n, a, b); });
    float t2 = time_ms([&] { axpby_kernel<VEC, UNROLL, BLOCK><<<g, BLOCK>>>(n, 1.0, a, 1.0, b); });
    int gr = g > kMaxPartials ? kMaxPartials : g;
    float t3 = 0, t4 = 0, t5 = 0;
    if (c > 0 || g <= kMaxPartials) {
      t3 = time_ms([&] { dot_kernel<0, VEC, UNROLL, BLOCK><<<gr, BLOCK>>>(n, a, nullptr, ws, out, 0); });
      t4 = time_ms([&] { dot_kernel<1, VEC, UNROLL, BLOCK><<<gr, BLOCK>>>(n, a, b, ws, out, 0); });
      t5 = time_ms([&] { cg_update_kernel<VEC, UNROLL, BLOCK><<<gr, BLOCK>>>(n, a, 1.0, b, ws, out, 0); });
    }
    auto gbs = [&](double bytes, float ms) { return ms > 0 ? bytes / ms * 1e-6 : 0.0; };
    printf("vec%d unroll%d block%4d cta/sm%3d grid%8d | copy %7.1f axpby %7.1f norm %7.1f dot %7.1f cg %7.1f GB/s\n",
           VEC, UNROLL, BLOCK, c, g, gbs(16.0 * n, t1), gbs(24.0 * n, t2), gbs(8.0 * n, t3), gbs(16.0 * n, t4), gbs(24.0 * n, t5));
    fflush(stdout);
  }
}

int main(int argc, char** argv) {
  int lg = argc > 1 ? atoi(argv[1]) : 27;
  int64_t n = (int64_t)1 << lg;
  CK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0));
  double *a, *b, *out;
  ReduceWs* ws;
  CK(cudaMalloc(&a, n * 8));
  CK(cudaMalloc(&b, n * 8));
  CK(cudaMalloc(&out, 8));
  CK(cudaMalloc(&ws, sizeof(ReduceWs)));
  CK(cudaMemset(a, 0, n * 8));
  CK(cudaMemset(b, 0, n * 8));
  CK(cudaMemset(ws, 0, sizeof(ReduceWs)));
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  printf("n = 2^%d doubles per array, %d SMs\n", lg, g_sms);
  float tm = time_ms([&] { CK(cudaMemcpyAsync(b, a, n * 8, cudaMemcpyDeviceToDevice)); });
  printf("cudaMemcpy D2D: %.1f GB/s\n", 16.0 * n / tm * 1e-6);
  sweep<4, 1, 256>(n, a, b, ws, out);
  sweep<4, 2, 256>(n, a, b, ws, out);
  sweep<4, 4, 256>(n, a, b, ws, out);
  sweep<4, 8, 256>(n, a, b, ws, out);
  sweep<4, 2, 512>(n, a, b, ws, out);
  sweep<4, 4, 512>(n, a, b, ws, out);
  sweep<4, 2, 1024>(n, a, b, ws, out);
  sweep<4, 4, 128>(n, a, b, ws, out);
  sweep<2, 4, 256>(n, a, b, ws, out);
  sweep<2, 8, 256>(n, a, b, ws, out);
  sweep<2, 4, 512>(n, a, b, ws, out);
  sweep<1, 4, 256>(n, a, b, ws, out);
  return 0;
}
