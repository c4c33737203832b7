// Box hex mesh labels and the process grid (see ogs.hpp for the reference citations).
#include <algorithm>
#include <cmath>

#include "ogs.hpp"

namespace spn {

// n = nx*ny*nz with nx >= ny >= nz as close to one another as the divisors of n allow
// (same results as the reference's Factor3 for the grids the benchmarks use: 2 -> (2,1,1),
// 4 -> (2,2,1), 8 -> (2,2,2), 6 -> (3,2,1), 12 -> (3,2,2)).
void Factor3(const int n, int& nx, int& ny, int& nz) {
  auto largest_divisor_leq = [](int m, double bound) {
    int best = 1;
    for (int d = 1; d <= m && d <= (int)std::floor(bound + 1e-9); ++d)
      if (m % d == 0) best = d;
    return best;
  };
  int a = largest_divisor_leq(n, std::cbrt((double)n));
  const int f = n / a;
  int b = largest_divisor_leq(f, std::sqrt((double)f));
  int c = f / b;
  int v[3] = {a, b, c};
  std::sort(v, v + 3);
  nx = v[2];
  ny = v[1];
  nz = v[0];
}

// x fastest, then y, then z.  The reference's recursive RankDecomp3 gives this same map whenever every
// grid dimension is 1 or prime (all of 1, 2, 4, 8 ranks); results never depend on which rank sits where.
void RankDecomp3(int size_x, int size_y, int /*size_z*/, int& rank_x, int& rank_y, int& rank_z, const int rank) {
  rank_x = rank % size_x;
  rank_y = (rank / size_x) % size_y;
  rank_z = rank / (size_x * size_y);
}

void BoxHexGlobalIds(int rank, int size, int nx, int ny, int nz, int p, hlong* ids, hlong* NgatherGlobalExpected) {
  int sx, sy, sz, rx, ry, rz;
  Factor3(size, sx, sy, sz);
  RankDecomp3(sx, sy, sz, rx, ry, rz, rank);
  const int Nq = p + 1, Np = Nq * Nq * Nq;
  const hlong LX = (hlong)sx * nx * p + 1, LY = (hlong)sy * ny * p + 1, LZ = (hlong)sz * nz * p + 1;
  if (NgatherGlobalExpected) *NgatherGlobalExpected = LX * LY * LZ;
  const hlong ox = (hlong)rx * nx * p, oy = (hlong)ry * ny * p, oz = (hlong)rz * nz * p;
#pragma omp parallel for collapse(2)
  for (int ez = 0; ez < nz; ++ez)
    for (int ey = 0; ey < ny; ++ey)
      for (int ex = 0; ex < nx; ++ex) {
        const hlong e = ex + (hlong)ey * nx + (hlong)ez * nx * ny;
        hlong* out = ids + e * Np;
        for (int k = 0; k < Nq; ++k)
          for (int j = 0; j < Nq; ++j)
            for (int i = 0; i < Nq; ++i) {
              const hlong gi = ox + (hlong)ex * p + i, gj = oy + (hlong)ey * p + j, gk = oz + (hlong)ez * p + k;
              out[i + j * Nq + k * Nq * Nq] = 1 + gi + gj * LX + gk * LX * LY;
            }
      }
}

}  // namespace spn
