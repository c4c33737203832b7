// Level-0 end-to-end check, compiled by `make ref-drivers` into build/ref_drivers/ogs_check:
// the REFERENCE's own mesh_t + ogs_t (libs/mesh, libs/ogs, unchanged) on the sm_100a backend through include/shim.
// For the box-hex mesh of the command line it runs the reference's device path — ogs_t::GatherScatter / Gather / Scatter on
// deviceMemory<T>: ogsOperator_t -> kernel_t -> C ABI kernels, and at P > 1 its pairwise exchange over the MPI
// stand-in — and compares with the reference's own host-memory overloads (libs/ogs/ogs.cpp:127-204, 303-393, 494-578) on
// the same vectors: GatherScatter(Add) for every type (float, double, int, long long) x Transpose x K in {1, 3};
// Gather(Add / Max, Trans / NoTrans) and Scatter(NoTrans / Trans) for two types.
// Same code, same fold order: the results must agree bit for bit.
// (Only Add for GatherScatter: the reference's DEVICE operator launches gatherScatterKernel[type][Add] whatever op it is
// given — libs/ogs/ogsOperator.cpp:443-480 — so for Mul / Max / Min its device and host paths differ on the interior
// rows by construction, on any backend.  `OGS_CHECK_ALL_OPS=1` runs them anyway and shows exactly that: mismatches for
// Sym and Trans, none for NoTrans where every row gathers one copy.)  Rank 0 prints
//   ogs_check: P=<ranks> N=<degree> Nlocal=<nodes> Ngather=<global> gatherScatter=<mismatches> gather=<..> scatter=<..> cases=<n>
// and the exit code is non-zero on any mismatch.   usage: [spnrun -np P] ogs_check -nx 4 -ny 4 -nz 4 -p 3
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "mesh.hpp"
#include "ogs.hpp"
#include "platform.hpp"

using namespace libp;

namespace {

template <typename T>
long long mismatches(const memory<T>& a, const memory<T>& b, size_t n) {
  long long bad = 0;
  for (size_t i = 0; i < n; ++i) bad += (memcmp(&a[i], &b[i], sizeof(T)) != 0) ? 1 : 0;
  return bad;
}

// seeded values, different on every rank; small magnitudes so that products over a node's copies stay finite
template <typename T>
void fill(memory<T>& q, size_t n, int rank, int salt) {
  for (size_t i = 0; i < n; ++i) {
    const unsigned long long h = ((unsigned long long)(i + 1) * 2654435761ull + 97ull * (unsigned)rank + 7919ull * (unsigned)salt) % 1000003ull;
    if (std::is_floating_point<T>::value) q[i] = (T)((double)h / 500001.5 - 1.0);
    else q[i] = (T)((long long)(h % 7ull) - 3);
  }
}

struct Tally { long long gs = 0, g = 0, s = 0, cases = 0; };

template <typename T>
void check_gather_scatter(mesh_t& mesh, platform_t& platform, int rank, Tally& tally) {
  const size_t N = (size_t)mesh.Nelements * mesh.Np;
  const ogs::Op ops[4] = {ogs::Add, ogs::Mul, ogs::Max, ogs::Min};
  const ogs::Transpose modes[3] = {ogs::Sym, ogs::NoTrans, ogs::Trans};
  const int nops = getenv("OGS_CHECK_ALL_OPS") ? 4 : 1;
  for (int K = 1; K <= 3; K += 2) {
    memory<T> q(N * K), qHost(N * K), qDev(N * K);
    for (int o = 0; o < nops; ++o)
      for (int m = 0; m < 3; ++m) {
        fill(q, N * K, rank, 16 * o + 4 * m + K);
        qHost.copyFrom(q);
        mesh.ogs.GatherScatter(qHost, K, ops[o], modes[m]);
        deviceMemory<T> o_q = platform.malloc<T>(q);
        mesh.ogs.GatherScatter(o_q, K, ops[o], modes[m]);
        o_q.copyTo(qDev);
        const long long bad = mismatches(qHost, qDev, N * K);
        if (bad && getenv("OGS_CHECK_VERBOSE")) printf("rank %d: GatherScatter sizeof(T)=%d K=%d op=%d mode=%d: %lld mismatches\n", rank, (int)sizeof(T), K, o, m, bad);
        tally.gs += bad;
        tally.cases++;
      }
  }
}

template <typename T>
void check_gather_and_scatter(mesh_t& mesh, platform_t& platform, int rank, Tally& tally) {
  const size_t N = (size_t)mesh.Nelements * mesh.Np;
  const size_t Ng = (size_t)mesh.ogs.Ngather;
  for (int K = 1; K <= 3; K += 2) {
    memory<T> q(N * K);
    fill(q, N * K, rank, 1000 + K);
    const size_t ng = (Ng > 0 ? Ng : 1) * K;
    for (int variant = 0; variant < 3; ++variant) {
      const ogs::Op op = (variant == 2) ? ogs::Max : ogs::Add;
      const ogs::Transpose trans = (variant == 1) ? ogs::NoTrans : ogs::Trans;
      memory<T> gHost(ng), gDev(ng);
      mesh.ogs.Gather(gHost, q, K, op, trans);
      deviceMemory<T> o_v = platform.malloc<T>(q);
      deviceMemory<T> o_g = platform.malloc<T>(ng);
      mesh.ogs.Gather(o_g, o_v, K, op, trans);
      o_g.copyTo(gDev);
      tally.g += mismatches(gHost, gDev, Ng * K);
      tally.cases++;
      if (variant == 2) continue;
      // scatter the gathered vector back with the transpose mode that pairs with the gather
      const ogs::Transpose back = (trans == ogs::Trans) ? ogs::NoTrans : ogs::Trans;
      memory<T> sHost(N * K, (T)0), sDev(N * K, (T)0);
      mesh.ogs.Scatter(sHost, gHost, K, back);
      deviceMemory<T> o_s = platform.malloc<T>(sDev);
      mesh.ogs.Scatter(o_s, o_g, K, back);
      o_s.copyTo(sDev);
      tally.s += mismatches(sHost, sDev, N * K);
      tally.cases++;
    }
  }
}

}  // namespace

int main(int argc, char** argv) {
  Comm::Init(argc, argv);
  int rc = 0;
  {
    comm_t comm(Comm::World().Dup());
    settings_t settings(comm);
    platformAddSettings(settings);
    meshAddSettings(settings);
    settings.newSetting("-v", "--verbose", "VERBOSE", "FALSE", "Enable verbose output", {"TRUE", "FALSE"});
    settings.parseSettings(argc, argv);
    platform_t platform(settings);
    mesh_t mesh(platform, settings, comm);

    Tally tally;
    check_gather_scatter<float>(mesh, platform, comm.rank(), tally);
    check_gather_scatter<double>(mesh, platform, comm.rank(), tally);
    check_gather_scatter<int>(mesh, platform, comm.rank(), tally);
    check_gather_scatter<long long int>(mesh, platform, comm.rank(), tally);
    check_gather_and_scatter<double>(mesh, platform, comm.rank(), tally);
    check_gather_and_scatter<int>(mesh, platform, comm.rank(), tally);

    memory<long long> bad(3);
    bad[0] = tally.gs; bad[1] = tally.g; bad[2] = tally.s;
    comm.Allreduce(bad, Comm::Sum, 3);
    if (comm.rank() == 0)
      printf("ogs_check: P=%d N=%d Nlocal=%d Ngather=%lld gatherScatter=%lld gather=%lld scatter=%lld cases=%lld\n", comm.size(), mesh.N,
             (int)(mesh.Nelements * mesh.Np), (long long)mesh.ogs.NgatherGlobal, bad[0], bad[1], bad[2], tally.cases);
    rc = (bad[0] | bad[1] | bad[2]) ? 1 : 0;
  }
  Comm::Finalize();
  return rc;
}
