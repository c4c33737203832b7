// Level-0 end-to-end check, compiled by `make ref-drivers` into build/ref_drivers/ogs_check:
// the REFERENCE's own mesh_t + ogs_t (libs/mesh, libs/ogs, unchanged) on the sm_100a backend through include/shim.
// For the box-hex mesh of the command line it runs the reference's device path — ogs_t::GatherScatter / Gather / Scatter on
// deviceMemory<dfloat>: ogsOperator_t -> kernel_t -> C ABI kernels, and at P > 1 its pairwise exchange over the MPI
// stand-in — and compares with the reference's own host-memory overloads (libs/ogs/ogs.cpp:127-204, 303-393, 494-578) on
// the same vectors.  Same code, same fold order: the results must agree bit for bit.  Prints one line per rank 0:
//   ogs_check: P=<ranks> N=<degree> Nlocal=<nodes> Ngather=<global> gatherScatter=<mismatches> gather=<..> scatter=<..>
// and exits non-zero on any mismatch.   usage: [spnrun -np P] ogs_check -nx 4 -ny 4 -nz 4 -p 3
#include <cstdio>
#include <cstring>

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

    const dlong N = mesh.Nelements * mesh.Np;
    const dlong Ngather = mesh.ogs.Ngather;
    memory<dfloat> q(N), qHost(N), qDev(N);
    for (dlong n = 0; n < N; ++n)  // seeded, different on every rank, uniform(-1,1)-like with full mantissas
      q[n] = (dfloat)(((unsigned long long)(n + 1) * 2654435761ull + 97ull * comm.rank()) % 1000003ull) / 500001.5 - 1.0;

    // GatherScatter(Add, Sym)
    qHost.copyFrom(q);
    mesh.ogs.GatherScatter(qHost, 1, ogs::Add, ogs::Sym);
    deviceMemory<dfloat> o_q = platform.malloc<dfloat>(q);
    mesh.ogs.GatherScatter(o_q, 1, ogs::Add, ogs::Sym);
    o_q.copyTo(qDev);
    long long badGS = mismatches(qHost, qDev, (size_t)N);

    // Gather(Add, Trans) and Scatter(NoTrans)
    memory<dfloat> gHost(Ngather > 0 ? Ngather : 1), gDev(Ngather > 0 ? Ngather : 1);
    mesh.ogs.Gather(gHost, q, 1, ogs::Add, ogs::Trans);
    deviceMemory<dfloat> o_v = platform.malloc<dfloat>(q);
    deviceMemory<dfloat> o_g = platform.malloc<dfloat>(Ngather > 0 ? Ngather : 1);
    mesh.ogs.Gather(o_g, o_v, 1, ogs::Add, ogs::Trans);
    o_g.copyTo(gDev);
    long long badG = mismatches(gHost, gDev, (size_t)Ngather);

    memory<dfloat> sHost(N, 0.0), sDev(N, 0.0);
    mesh.ogs.Scatter(sHost, gHost, 1, ogs::NoTrans);
    deviceMemory<dfloat> o_s = platform.malloc<dfloat>(sDev);
    mesh.ogs.Scatter(o_s, o_g, 1, ogs::NoTrans);
    o_s.copyTo(sDev);
    long long badS = mismatches(sHost, sDev, (size_t)N);

    long long bad[3] = {badGS, badG, badS};
    memory<long long> badm(3);
    for (int i = 0; i < 3; ++i) badm[i] = bad[i];
    comm.Allreduce(badm, Comm::Sum, 3);
    if (comm.rank() == 0)
      printf("ogs_check: P=%d N=%d Nlocal=%d Ngather=%lld gatherScatter=%lld gather=%lld scatter=%lld\n", comm.size(), mesh.N, (int)N,
             (long long)mesh.ogs.NgatherGlobal, badm[0], badm[1], badm[2]);
    rc = (badm[0] | badm[1] | badm[2]) ? 1 : 0;
  }
  Comm::Finalize();
  return rc;
}
