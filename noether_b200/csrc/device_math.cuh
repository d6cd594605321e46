// Exact (never FMA-contracted) device arithmetic restating the reference's CPU evaluation order.
//
// The reference evaluates plane scalars and cut-point interpolation on an x86-64 CPU without FMA contraction
// (vtkPlane::EvaluateFunction, vtkTriangle::Contour — third-party VTK 9.1).  Bit-exact
// connectivity needs the same IEEE-754 double operations in the same order, so everything here goes through the
// round-to-nearest intrinsics, which nvcc never fuses.  (The library is additionally compiled with -fmad=false.)
//
// The same functions compile for the host (plain IEEE operations; the host code is built with -ffp-contract=off), which
// is how the CPU test suite replays the contouring and linking of the kernels without a GPU (tests/cpp/link_emulation.cpp).
#pragma once
#include "nb2_internal.h"

#if defined(__CUDACC__)
#define NB2_DEV __host__ __device__ __forceinline__
#else
#define NB2_DEV inline
#endif

namespace nb2
{
#if defined(__CUDA_ARCH__)
NB2_DEV double dmul(double a, double b) { return __dmul_rn(a, b); }
NB2_DEV double dadd(double a, double b) { return __dadd_rn(a, b); }
NB2_DEV double dsub(double a, double b) { return __dsub_rn(a, b); }
NB2_DEV double ddiv(double a, double b) { return __ddiv_rn(a, b); }
NB2_DEV float d2f(double a) { return __double2float_rn(a); }
#else
NB2_DEV double dmul(double a, double b) { return a * b; }
NB2_DEV double dadd(double a, double b) { return a + b; }
NB2_DEV double dsub(double a, double b) { return a - b; }
NB2_DEV double ddiv(double a, double b) { return a / b; }
NB2_DEV float d2f(double a) { return static_cast<float>(a); }
#endif

/** Origin of lattice plane k: start_loc + (k * line_spacing_) * cut_normal  (plane_slicer_raster_planner.cpp:598) */
struct PlaneOrigin
{
  double x, y, z;
};
NB2_DEV PlaneOrigin planeOrigin(const LatticeDev& L, long long k)
{
  const double is = dmul(static_cast<double>(k), L.spacing);
  PlaneOrigin o;
  o.x = dadd(L.start[0], dmul(is, L.n[0]));
  o.y = dadd(L.start[1], dmul(is, L.n[1]));
  o.z = dadd(L.start[2], dmul(is, L.n[2]));
  return o;
}

/** vtkPlane::EvaluateFunction: n0*(x0-o0) + n1*(x1-o1) + n2*(x2-o2), double, left to right */
NB2_DEV double planeScalar(const LatticeDev& L, const PlaneOrigin& o, float x, float y, float z)
{
  const double a = dmul(L.n[0], dsub(static_cast<double>(x), o.x));
  const double b = dmul(L.n[1], dsub(static_cast<double>(y), o.y));
  const double c = dmul(L.n[2], dsub(static_cast<double>(z), o.z));
  return dadd(dadd(a, b), c);
}

/** (a.x*d0 + a.y*d1) + a.z*d2 with float->double operands: orientation / ordering key of chains */
NB2_DEV double dotLR(double ax, double ay, double az, const double d[3])
{
  return dadd(dadd(dmul(ax, d[0]), dmul(ay, d[1])), dmul(az, d[2]));
}

// vtkTriangle::Contour tables (VTK 9.1 Common/DataModel/vtkTriangle.cxx):
//   edges     = {0,1} {1,2} {2,0}                      -> edge e joins local vertices e and (e+1)%3
//   lineCases = {-} {0,2} {1,0} {1,2} {2,1} {0,1} {2,0} {-}   (pairs of edge numbers, indexed by the case)
// packed two bits per case so the lookup is register arithmetic instead of a local-memory table.
NB2_DEV int lineCaseEdge(int index, int which)
{
  const unsigned packed = which == 0 ? 0x2250u : 0x0588u;
  return (packed >> (2 * index)) & 3;
}

// register-array selects (constant indices only, so the arrays never spill to local memory)
template <typename T>
NB2_DEV T pick3(const T a[3], int i)
{
  return i == 0 ? a[0] : (i == 1 ? a[1] : a[2]);
}

/** One endpoint of the contour line of a triangle: which mesh edge, where, with what interpolated normal */
struct CutPoint
{
  float x, y, z;     // float32 position as stored in vtkCutter's output points
  float nx, ny, nz;  // float32 interpolated normal (vtkDataSetAttributes::InterpolateEdge)
  uint32_t lo, hi;   // mesh vertices of the generating edge: lower scalar first
};

/** Contour case index of a triangle for plane scalars s[0..2]: bit i set when s[i] >= 0 (vtkTriangle::Contour) */
NB2_DEV int contourCase(const double s[3])
{
  return (s[0] >= 0.0 ? 1 : 0) | (s[1] >= 0.0 ? 2 : 0) | (s[2] >= 0.0 ? 4 : 0);
}

/** Where endpoint `which` (0/1) of the line of case `index` lies: on the triangle edge from local vertex e1 (lower scalar) to
 *  e2, at parameter t = (0 - s_lo) / (s_hi - s_lo) (vtkTriangle::Contour; t = 0 when the edge lies in the plane) */
struct CutParam
{
  int e1, e2;
  double t;
};
NB2_DEV CutParam cutParam(int index, int which, const double s[3])
{
  const int edge = lineCaseEdge(index, which);
  const int v0 = edge, v1 = edge == 2 ? 0 : edge + 1;
  double delta = dsub(pick3(s, v1), pick3(s, v0));
  CutParam c;
  if (delta > 0.0)
  {
    c.e1 = v0;
    c.e2 = v1;
  }
  else
  {
    c.e1 = v1;
    c.e2 = v0;
    delta = -delta;
  }
  c.t = (delta == 0.0) ? 0.0 : ddiv(dsub(0.0, pick3(s, c.e1)), delta);
  return c;
}
/** x_lo + t (x_hi - x_lo) in double, stored as float (the cut point's coordinate) */
NB2_DEV float lerpPosition(float a, float b, double t)
{
  return d2f(dadd(static_cast<double>(a), dmul(t, dsub(static_cast<double>(b), static_cast<double>(a)))));
}
/** a_lo (1 - t) + a_hi t in double, stored as float (vtkDataSetAttributes::InterpolateEdge) */
NB2_DEV float lerpAttribute(float a, float b, double t, double omt)
{
  return d2f(dadd(dmul(static_cast<double>(a), omt), dmul(static_cast<double>(b), t)));
}

/** Interpolate endpoint `which` (0/1) of the line of case `index`; normals only when nv != nullptr (nv = the vertex
 *  normals of the triangle's three vertices, already in registers).
 *  Restates vtkTriangle::Contour: interpolate from the lower- to the higher-scalar end, t = (0 - s_lo) / (s_hi - s_lo),
 *  x = x_lo + t (x_hi - x_lo) in double, stored as float; attributes a_lo (1 - t) + a_hi t in double, stored as float. */
NB2_DEV CutPoint interpolateCut(int index, int which, const uint32_t ids[3], const float4 p[3],
                                                   const double s[3], const float4* nv)
{
  const int edge = lineCaseEdge(index, which);
  const int v0 = edge, v1 = edge == 2 ? 0 : edge + 1;
  double delta = dsub(pick3(s, v1), pick3(s, v0));
  int e1, e2;
  if (delta > 0.0)
  {
    e1 = v0;
    e2 = v1;
  }
  else
  {
    e1 = v1;
    e2 = v0;
    delta = -delta;
  }
  const double t = (delta == 0.0) ? 0.0 : ddiv(dsub(0.0, pick3(s, e1)), delta);
  CutPoint c;
  const float4 a = pick3(p, e1), b = pick3(p, e2);
  c.x = d2f(dadd(static_cast<double>(a.x), dmul(t, dsub(static_cast<double>(b.x), static_cast<double>(a.x)))));
  c.y = d2f(dadd(static_cast<double>(a.y), dmul(t, dsub(static_cast<double>(b.y), static_cast<double>(a.y)))));
  c.z = d2f(dadd(static_cast<double>(a.z), dmul(t, dsub(static_cast<double>(b.z), static_cast<double>(a.z)))));
  c.lo = pick3(ids, e1);
  c.hi = pick3(ids, e2);
  if (nv != nullptr)
  {
    const float4 n1 = pick3(nv, e1), n2 = pick3(nv, e2);
    const double omt = dsub(1.0, t);
    c.nx = d2f(dadd(dmul(static_cast<double>(n1.x), omt), dmul(static_cast<double>(n2.x), t)));
    c.ny = d2f(dadd(dmul(static_cast<double>(n1.y), omt), dmul(static_cast<double>(n2.y), t)));
    c.nz = d2f(dadd(dmul(static_cast<double>(n1.z), omt), dmul(static_cast<double>(n2.z), t)));
  }
  else
  {
    c.nx = c.ny = c.nz = 0.f;
  }
  return c;
}

/** Both endpoints of the contour line of case `index` (1 .. 6), positions only — what k_emit_hits needs of every hit.
 *  The same arithmetic as interpolateCut(index, 0 / 1, ...), arranged around the triangle's LONE vertex (the one on its
 *  own side of the plane): both cut edges start there, so the three vertices are rotated once (lone vertex first) and
 *  everything after is fixed positions instead of a select chain per operand.  In vtkTriangle's tables, after that
 *  rotation, a case with one vertex inside reads {edge 0, edge 2} and a case with two inside {edge 2, edge 0}; an edge is
 *  interpolated from its vertex outside (scalar < 0) to its vertex inside, t = (0 - s_out) / (s_in - s_out)
 *  (-(a - b) and b - a are the same double, so the sign flip of cutParam needs no restating). */
NB2_DEV void cutBothPositions(int index, const float4 p[3], const double s[3], float c0[3], float c1[3])
{
  const bool single = index == 1 || index == 2 || index == 4;  // one vertex inside: it is the lone one
  const int hot = single ? index : (7 ^ index);                 // one-hot position of the lone vertex
  const int l = hot >> 1;                                       // 1, 2, 4 -> 0, 1, 2
  const int l1 = l == 2 ? 0 : l + 1, l2 = l == 0 ? 2 : l - 1;
  const float4 q0 = pick3(p, l), q1 = pick3(p, l1), q2 = pick3(p, l2);
  const double s0 = pick3(s, l), s1 = pick3(s, l1), s2 = pick3(s, l2);
  // line end 0 / 1: single -> edges (q0,q1) / (q2,q0), from q1 / q2 towards q0;  double -> edges (q2,q0) / (q0,q1), from q0
  const float4 a0 = single ? q1 : q0, b0 = single ? q0 : q2;
  const float4 a1 = single ? q2 : q0, b1 = single ? q0 : q1;
  const double lo0 = single ? s1 : s0, hi0 = single ? s0 : s2;
  const double lo1 = single ? s2 : s0, hi1 = single ? s0 : s1;
  const double d0 = dsub(hi0, lo0), d1 = dsub(hi1, lo1);
  const double t0 = (d0 == 0.0) ? 0.0 : ddiv(dsub(0.0, lo0), d0);
  const double t1 = (d1 == 0.0) ? 0.0 : ddiv(dsub(0.0, lo1), d1);
  c0[0] = lerpPosition(a0.x, b0.x, t0);
  c0[1] = lerpPosition(a0.y, b0.y, t0);
  c0[2] = lerpPosition(a0.z, b0.z, t0);
  c1[0] = lerpPosition(a1.x, b1.x, t1);
  c1[1] = lerpPosition(a1.y, b1.y, t1);
  c1[2] = lerpPosition(a1.z, b1.z, t1);
}

}  // namespace nb2