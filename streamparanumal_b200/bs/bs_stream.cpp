// BS0 .. BS5 drivers: the reference executables' command line, warm-up / timing loops and output lines
// (BS/BS0/src/bs0Run.cpp:30-82, BS/BS1/src/bs1Run.cpp:29-91, bs2Run.cpp:29-94, bs3Run.cpp:29-99,
// bs4Run.cpp:29-99, bs5Run.cpp:29-103) on top of spn::platform_t / kernel_t, i.e. on the sm_100a kernels.
// One source, compiled once per benchmark with -DSPN_BS=<0..5>.
//
// What is the same: flags and defaults, N = B / (bytes per entry), 5 warm-up calls at Nmax, 20 timed calls per
// size between GlobalPlatformTime stamps, bytes counted as the reference counts them, the output format.
// What differs: with more than one rank the BS3-5 scalar is summed over the ranks by ncclAllReduce on the device; kernel1 of BS3-5 finishes the reduction itself (last-block-done), so kernel2 only folds the
// Nblock-entry tmp array the reference way; N is kept in 64 bits internally (the reference's `int N` overflows
// above 2^31 entries, which no BS flag reaches at 4 GiB).
#include <cstdio>

#include "kernel.hpp"
#include "settings.hpp"

#ifndef SPN_BS
#error "compile with -DSPN_BS=<0..5>"
#endif

using namespace spn;

namespace {

#if SPN_BS == 0
// ------------------------------------------------------------------------------------------- BS0
void run(platform_t& platform, settings_t& settings) {
  int Nlaunches = 10, NlaunchMin = 1, NlaunchMax = 1024, NlaunchStep = 1;
  settings.getSetting("NLAUNCHES", Nlaunches);
  settings.getSetting("NLAUNCHMIN", NlaunchMin);
  settings.getSetting("NLAUNCHMAX", NlaunchMax);
  settings.getSetting("NLAUNCHSTEP", NlaunchStep);
  if (!(Nlaunches | NlaunchMin | NlaunchMax)) Nlaunches = 10;  // default: one test of 10 launches
  if (NlaunchStep < 1) NlaunchStep = 1;

  properties_t kernelInfo;
  kernel_t kernel = platform.buildKernel("BS/BS0/okl/bs0.okl", "bs0", kernelInfo);
  const int N = 1;
  deviceMemory<dfloat> o_a = platform.malloc<dfloat>(N), o_b = platform.malloc<dfloat>(N);
  for (int w = 0; w < 100; ++w) kernel(N, o_a, o_b);

  const bool single = Nlaunches != 0;
  if (single) {
    NlaunchMin = NlaunchMax = Nlaunches;
    printf("BS0 = [");
  } else {
    printf("%%Launch Latency [Nlaunches, time per launch, elapsed]\n");
    printf("BS0 = [\n");
  }
  for (int n = NlaunchMin; n <= NlaunchMax; n += NlaunchStep) {
    timePoint_t start = PlatformTime(platform);
    for (int i = 0; i < n; ++i) kernel(N, o_a, o_b);
    timePoint_t end = PlatformTime(platform);
    const double elapsed = ElapsedTime(start, end);
    printf("%d %5.4e %5.4e", n, elapsed / n, elapsed);
    if (n < NlaunchMax) printf(";\n");
  }
  printf("]; %%Launch Latency [Nlaunches, time per launch, elapsed]\n");
}
void addSettings(settings_t& s) {
  s.newSetting("-n", "--Nlaunch", "NLAUNCHES", "0", "Number of small kernels to launch");
  s.newSetting("-nmin", "--NlaunchMin", "NLAUNCHMIN", "0", "Minimum number of small kernels to launch");
  s.newSetting("-nmax", "--NlaunchMax", "NLAUNCHMAX", "0", "Maximum number of small kernels to launch");
  s.newSetting("-s", "--step", "NLAUNCHSTEP", "1", "Step size of small kernels to launch");
}
#else
// ------------------------------------------------------------------------------------- BS1 .. BS5
struct Bench {
  const char* tag;     // "BS1"
  const char* title;   // text after '%' in the output
  int doublesMoved;    // bytes per entry / sizeof(dfloat), as the reference counts them
};
#if SPN_BS == 1
constexpr Bench kBench{"BS1", "Copy", 2};
#elif SPN_BS == 2
constexpr Bench kBench{"BS2", "AXPY", 3};
#elif SPN_BS == 3
constexpr Bench kBench{"BS3", "Norm", 1};
#elif SPN_BS == 4
constexpr Bench kBench{"BS4", "Dot", 2};
#elif SPN_BS == 5
constexpr Bench kBench{"BS5", "CG Update", 3};
#endif

void addSettings(settings_t& s) {
  s.newSetting("-b", "--bytes", "BYTES", "0", "Array size in bytes");
  s.newSetting("-bmin", "--bmin", "BMIN", "0", "Minimum array size in bytes");
  s.newSetting("-bmax", "--bmax", "BMAX", "0", "Maximum array size in bytes");
  s.newSetting("-s", "--bstep", "BSTEP", "1048576", "Array step size in bytes");
}

void run(platform_t& platform, settings_t& settings) {
  size_t B = 0, Bmin = 0, Bmax = 0, Bstep = 0;
  settings.getSetting("BYTES", B);
  settings.getSetting("BMIN", Bmin);
  settings.getSetting("BMAX", Bmax);
  settings.getSetting("BSTEP", Bstep);
  if (!(B | Bmin | Bmax)) B = 1073741824;  // default: one test moving 1 GiB
  if (B) Bmax = B;

  const size_t sc = kBench.doublesMoved * sizeof(dfloat);
  long long Nmin = (long long)(Bmin / sc), Nmax = (long long)(Bmax / sc);
  const long long Nstep = (Bstep / sc > 0) ? (long long)(Bstep / sc) : 1;
  SPN_ABORT("array of more than 2^31-1 entries: dlong is 32 bits (include/types.h)", Nmax > 0x7fffffffLL);

  // kernels and buffers (bsNSetup.cpp): blockSize is what the reference sizes o_tmp with
  properties_t kernelInfo;
  const int blockSize = (SPN_BS == 5) ? 1024 : 256;  // bs5Setup.cpp:37 (CUDA mode), bs3Setup.cpp / bs4Setup.cpp
  kernelInfo["defines/p_blockSize"] = blockSize;
  const std::string okl = std::string("BS/") + kBench.tag + "/okl/bs" + std::to_string(SPN_BS) + ".okl";
  const std::string base = "bs" + std::to_string(SPN_BS);
#if SPN_BS <= 2
  kernel_t kernel = platform.buildKernel(okl, base, kernelInfo);
#else
  kernel_t kernel1 = platform.buildKernel(okl, base + "_1", kernelInfo);
  kernel_t kernel2 = platform.buildKernel(okl, base + "_2", kernelInfo);
  deviceMemory<dfloat> o_tmp = platform.malloc<dfloat>(blockSize);
  deviceMemory<dfloat> o_out = platform.malloc<dfloat>(1);
#endif
  deviceMemory<dfloat> o_a = platform.malloc<dfloat>((size_t)Nmax);
#if SPN_BS != 3
  deviceMemory<dfloat> o_b = platform.malloc<dfloat>((size_t)Nmax);
#endif
  const dfloat alpha = 1.0, beta = 1.0;
  (void)alpha; (void)beta;

  auto call = [&](long long N) {
    const int n = (int)N;
#if SPN_BS == 1
    kernel(n, o_a, o_b);
#elif SPN_BS == 2
    kernel(n, alpha, o_a, beta, o_b);
#else
    int Nblock = (int)((N + blockSize - 1) / blockSize);
    Nblock = (Nblock > blockSize) ? blockSize : Nblock;  // limit to blockSize entries
    if (Nblock < 1) Nblock = 1;
#if SPN_BS == 3
    kernel1(Nblock, n, o_a, o_tmp);
#elif SPN_BS == 4
    kernel1(Nblock, n, o_a, o_b, o_tmp);
#else
    kernel1(Nblock, n, o_a, alpha, o_b, o_tmp);  // o_a = Ap, o_b = r
#endif
    kernel2(Nblock, o_tmp, o_out);
    // partitioned run: the scalar is the sum over ranks (north star: NCCL allreduce behind norm / dot / r.r; the
    // reference drivers leave one partial result per rank)
    if (platform.size() > 1) platform.comm.AllreduceDevice(o_out.ptr(), 1, platform.getStream());
#endif
  };

  for (int n = 0; n < 5; ++n) call(Nmax);  // warm up

  if (B) {
    Nmin = Nmax = (long long)(B / sc);
    printf("%s = [", kBench.tag);
  } else {
    printf("%%%s [DOFs, elapsed, DOFs/s, BW (GB/s)]\n", kBench.title);
    printf("%s = [\n", kBench.tag);
  }
  const int Ntests = 20;
  for (long long N = Nmin; N <= Nmax; N += Nstep) {
    timePoint_t start = GlobalPlatformTime(platform);
    for (int n = 0; n < Ntests; ++n) call(N);
    timePoint_t end = GlobalPlatformTime(platform);
    const double elapsedTime = ElapsedTime(start, end) / Ntests;
    const size_t bytes = (size_t)N * sc;
    printf("%lld %5.4e %5.4e %6.2f", N, elapsedTime, N / elapsedTime, (double)(bytes / 1.e9) / elapsedTime);
    if (N < Nmax) printf(";\n");
  }
  printf("]; %%%s [DOFs, elapsed, DOFs/s, BW (GB/s)]\n", kBench.title);
}
#endif

}  // namespace

int main(int argc, char** argv) {
  try {
    comm_t comm = comm_t::World();  // Comm::Init + Comm::World().Dup() of the reference mains
    settings_t settings(comm);
    platformAddSettings(settings);
    addSettings(settings);
    verboseAddSetting(settings);
    settings.parseSettings(argc, argv);
    if (settings.compareSetting("VERBOSE", "TRUE") && comm.rank() == 0) settings.report();
    platform_t platform = makePlatform(settings);
    run(platform, settings);
  } catch (const std::exception& e) {
    fprintf(stderr, "%s\n", e.what());
    return -1;  // LIBP_ERROR (include/utils.hpp:50-51)
  }
  return 0;  // LIBP_SUCCESS
}
