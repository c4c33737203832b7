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

// Labels the reference's mesh_t::ConnectNodes converges to (libs/mesh/meshConnectNodes.cpp:44-81): every node starts
// as 1 + e*Np + n + (nodes of the lower ranks) and repeatedly takes the smaller label of the node across each face, so
// a node ends with the MINIMUM of that expression over all of its coincident copies, on this rank and on others.
// For the box that minimum has a closed form: along each axis a lattice coordinate on an element boundary belongs
// to two elements, so a node has at most 8 copies, enumerated directly.
void BoxHexGlobalIds(int rank, int size, int nx, int ny, int nz, int p, hlong* ids, hlong* NgatherGlobalExpected) {
  int sx, sy, sz, rx, ry, rz;
  Factor3(size, sx, sy, sz);
  RankDecomp3(sx, sy, sz, rx, ry, rz, rank);
  const int Nq = p + 1, Np = Nq * Nq * Nq;
  const hlong NX = (hlong)sx * nx, NY = (hlong)sy * ny, NZ = (hlong)sz * nz;  // global elements per axis
  if (NgatherGlobalExpected) *NgatherGlobalExpected = (NX * p + 1) * (NY * p + 1) * (NZ * p + 1);
  const hlong nodesPerRank = (hlong)nx * ny * nz * Np;  // comm.Scan(localNnodes): every rank owns the same count
  // label of node (i,j,k) of the element at GLOBAL element coordinates (EX,EY,EZ)
  auto label = [&](hlong EX, hlong EY, hlong EZ, int i, int j, int k) -> hlong {
    const int qx = (int)(EX / nx), qy = (int)(EY / ny), qz = (int)(EZ / nz);
    const hlong r = qx + (hlong)qy * sx + (hlong)qz * sx * sy;  // inverse of RankDecomp3 above
    const hlong e = (EX - (hlong)qx * nx) + (EY - (hlong)qy * ny) * nx + (EZ - (hlong)qz * nz) * nx * ny;
    return 1 + r * nodesPerRank + e * Np + (i + j * Nq + k * Nq * Nq);
  };
#pragma omp parallel for collapse(2)
  for (int ez = 0; ez < nz; ++ez)
    for (int ey = 0; ey < ny; ++ey)
      for (int ex = 0; ex < nx; ++ex) {
        const hlong e = ex + (hlong)ey * nx + (hlong)ez * nx * ny;
        const hlong EX = (hlong)rx * nx + ex, EY = (hlong)ry * ny + ey, EZ = (hlong)rz * nz + ez;
        hlong* out = ids + e * Np;
        for (int k = 0; k < Nq; ++k)
          for (int j = 0; j < Nq; ++j)
            for (int i = 0; i < Nq; ++i) {
              // the other copy along an axis: node 0 of this element == node p of the previous one, and vice versa
              const int ax = (i == 0 && EX > 0) ? -1 : (i == p && EX + 1 < NX) ? 1 : 0;
              const int ay = (j == 0 && EY > 0) ? -1 : (j == p && EY + 1 < NY) ? 1 : 0;
              const int az = (k == 0 && EZ > 0) ? -1 : (k == p && EZ + 1 < NZ) ? 1 : 0;
              hlong best = label(EX, EY, EZ, i, j, k);
              for (int cz = 0; cz <= (az != 0); ++cz)
                for (int cy = 0; cy <= (ay != 0); ++cy)
                  for (int cx = 0; cx <= (ax != 0); ++cx) {
                    if (!(cx | cy | cz)) continue;
                    const int ii = cx ? (ax < 0 ? p : 0) : i, jj = cy ? (ay < 0 ? p : 0) : j, kk = cz ? (az < 0 ? p : 0) : k;
                    const hlong v = label(EX + cx * ax, EY + cy * ay, EZ + cz * az, ii, jj, kk);
                    if (v < best) best = v;
                  }
              out[i + j * Nq + k * Nq * Nq] = best;
            }
      }
}

}  // namespace spn
