// TEST INFRASTRUCTURE.  Host replay of k_pcl_normals (noether_b200/csrc/kernels_cluster.cu): the kernel's own per-point
// functions — cluster_ops.h (FLANN's float distance test) and pcl_normal_ops.h (shifted moments -> pcl::eigen33 -> flip)
// — driven by a loop over the points with an all-pairs neighbour scan in a scrambled order, so that the CPU tests compare the
// device arithmetic with the oracle (oracle/oracle_normals_pcl.cpp) on a machine without a GPU.
#include "../../noether_b200/csrc/cluster_ops.h"
#include "../../noether_b200/csrc/pcl_normal_ops.h"

#include <cmath>
#include <cstdint>
#include <vector>

extern "C" int emu_pcl_normals(const float* xyz, uint64_t n, double radius, double vx, double vy, double vz, unsigned seed, float* normals,
                               float* curvature, uint32_t* count)
{
  using namespace nb2;
  const float r2 = static_cast<float>(radius * radius);
  const float vp[3] = { static_cast<float>(vx), static_cast<float>(vy), static_cast<float>(vz) };
  for (uint64_t p = 0; p < n; ++p)
  {
    const float* P = xyz + 3 * p;
    double S[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    unsigned k = 0;
    // neighbours visited in a scrambled order (the device walks hash chains in whatever order the atomics left them)
    const uint64_t step = n > 1 ? (2 * ((seed + p) % (n / 2 + 1)) + 1) : 1;
    uint64_t q = (seed * 7919ull + p) % n;
    uint64_t g = step;
    {
      // a stride coprime with n visits every point once
      auto gcd = [](uint64_t a, uint64_t b) { while (b) { const uint64_t t = a % b; a = b; b = t; } return a; };
      while (gcd(g, n) != 1)
        ++g;
    }
    for (uint64_t i = 0; i < n; ++i, q = (q + g) % n)
    {
      const float* Q = xyz + 3 * q;
      if (!(radius > 0.0) || !clu::isNeighbour(P[0], P[1], P[2], Q[0], Q[1], Q[2], r2))
        continue;
      const double x = static_cast<double>(Q[0] - P[0]), y = static_cast<double>(Q[1] - P[1]), z = static_cast<double>(Q[2] - P[2]);
      S[0] += x * x, S[1] += x * y, S[2] += x * z, S[3] += y * y, S[4] += y * z, S[5] += z * z, S[6] += x, S[7] += y, S[8] += z;
      ++k;
    }
    float nn[3] = { std::nanf(""), std::nanf(""), std::nanf("") }, c = std::nanf("");
    if (k >= 3)
      pclne::normalFromSums(S, k, P, vp, nn, c);
    normals[3 * p] = nn[0], normals[3 * p + 1] = nn[1], normals[3 * p + 2] = nn[2];
    curvature[p] = c;
    if (count)
      count[p] = k;
  }
  return 0;
}
