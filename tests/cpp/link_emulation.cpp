// TEST INFRASTRUCTURE.  Host replay of the contouring (k_emit_hits), the fast linking path (k_link_planes) and the
// waypoint pass (k_waypoints) of
// noether_b200/csrc/kernels_slice.cu: the kernels' own per-element functions — device_math.cuh (plane scalars, vtkTriangle
// case table, cut-point and normal interpolation) and link2_ops.h (every phase of the linker) — driven by loops instead of
// CTAs, "threads" visited in a scrambled order and the plane's cut lines stored in a scrambled order (the order the
// kernel's atomics leave them in is arbitrary), so that the CUDA path's logic is compared with the oracle bit for bit on
// a machine without a GPU.  Planes the fast path declines (it returns before writing anything) are reported with the
// reason; on the device those go to the general path.
//
// Build (tests/test_link_emulation.py): g++ -std=c++17 -O2 -fPIC -shared -ffp-contract=off -fno-fast-math
//        -I/usr/local/cuda/include tests/cpp/link_emulation.cpp -o tests/cpp/build/liblink_emulation.so
#include "../../noether_b200/csrc/device_math.cuh"
#include "../../noether_b200/csrc/link2_ops.h"

#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

using namespace nb2;
using namespace nb2::link2;

namespace
{
struct HostOps
{
  static void orr(unsigned* a, unsigned v) { *a |= v; }
  static unsigned cas(unsigned* a, unsigned cmp, unsigned v)
  {
    const unsigned seen = *a;
    if (seen == cmp)
      *a = v;
    return seen;
  }
  static unsigned add(unsigned* a, unsigned v)
  {
    const unsigned old = *a;
    *a = old + v;
    return old;
  }
  static unsigned claim(unsigned* counter, bool wanted)
  {
    if (!wanted)
      return 0u;
    return (*counter)++;
  }
  static unsigned vload(const unsigned* a) { return *a; }
  static Rec4 load(const Rec4* p) { return *p; }
  static unsigned long long ld64(const unsigned long long* p) { return *p; }
  static void st64(unsigned long long* p, unsigned long long v) { *p = v; }
};

struct Result
{
  std::vector<int64_t> path_plane;
  std::vector<uint32_t> path_segments;  // segments per non-empty plane
  std::vector<uint32_t> seg_len;
  std::vector<float> pos, nrm;
  std::vector<uint32_t> elo, ehi;
  std::vector<int64_t> declined_plane;
  std::vector<uint32_t> declined_reason;
  uint64_t lines = 0;
  std::vector<uint32_t> declined_misc;  // the first eight misc words of every declined plane (diagnostics)
  unsigned max_rounds = 0;              // most claim rounds any plane needed
  uint64_t stragglers = 0;              // endpoints the plain rounds left to the compare-and-swap loop
} g;

/** visit 0 .. n-1 in a scrambled (but complete) order */
template <typename F>
void scrambled(unsigned n, unsigned seed, F f)
{
  if (n == 0)
    return;
  unsigned stride = 1;
  for (unsigned cand : { 7919u, 104729u, 1299709u, 15485863u })
    if (std::gcd(cand, n) == 1)
      stride = cand % n ? cand % n : 1;
  if (std::gcd(stride, n) != 1)
    stride = 1;
  unsigned i = seed % n;
  for (unsigned k = 0; k < n; ++k)
  {
    f(i);
    i = (i + stride) % n;
  }
}
}  // namespace

extern "C" {

/** Slice lattice planes [plane_begin, plane_end) of the mesh; nth "threads" per plane.  Returns the number of planes the
 *  fast path declined (their waypoints are missing from the result). */
int emu_slice(const float* xyz, const float* nrm, uint64_t V, const uint32_t* tri, uint64_t F, const double n[3],
              const double start[3], double spacing, const double dir[3], int64_t num_planes, int64_t plane_begin,
              int64_t plane_end, unsigned nth, unsigned seed, int want_edges)
{
  (void)V;
  g = Result{};
  LatticeDev L{};
  for (int i = 0; i < 3; ++i)
    L.n[i] = n[i], L.start[i] = start[i], L.dir[i] = dir[i];
  L.spacing = spacing;
  L.num_planes = num_planes;
  L.plane_begin = plane_begin;
  L.plane_end = plane_end;
  std::vector<Rec4> rec, nrec;
  std::vector<uint32_t> erec;
  std::vector<uint8_t> area;
  for (int64_t k = plane_begin; k < plane_end; ++k)
  {
    // ---- k_emit_hits for this plane: every triangle the plane cuts, contoured exactly as on the device -----------------
    const PlaneOrigin org = planeOrigin(L, k);
    std::vector<Rec4> r, nr;
    std::vector<uint32_t> er;
    for (uint64_t t = 0; t < F; ++t)
    {
      const uint32_t ids[3] = { tri[3 * t], tri[3 * t + 1], tri[3 * t + 2] };
      float4 p[3], nv[3];
      double s[3];
      for (int j = 0; j < 3; ++j)
      {
        p[j] = make_float4(xyz[3 * ids[j]], xyz[3 * ids[j] + 1], xyz[3 * ids[j] + 2], 0.f);
        nv[j] = make_float4(nrm[3 * ids[j]], nrm[3 * ids[j] + 1], nrm[3 * ids[j] + 2], 0.f);
        s[j] = planeScalar(L, org, p[j].x, p[j].y, p[j].z);
      }
      const int index = contourCase(s);
      if (index == 0 || index == 7)
        continue;
      for (int w = 0; w < 2; ++w)
      {
        const CutPoint c = interpolateCut(index, w, ids, p, s, nv);
        Rec4 a{ c.x, c.y, c.z, 0.f };
        const unsigned rank = static_cast<unsigned>(2 * t + w);
        std::memcpy(&a.w, &rank, 4);
        r.push_back(a);
        nr.push_back(Rec4{ c.nx, c.ny, c.nz, 0.f });
        er.push_back(c.lo);
        er.push_back(c.hi);
      }
    }
    const unsigned nl = static_cast<unsigned>(r.size() / 2);
    if (nl == 0)
      continue;
    g.lines += nl;
    // the bucket order is arbitrary on the device: scramble the lines
    rec.resize(2 * nl);
    nrec.resize(2 * nl);
    erec.resize(4 * nl);
    {
      unsigned at = 0;
      scrambled(nl, seed + static_cast<unsigned>(k) * 31u, [&](unsigned j) {
        for (int w = 0; w < 2; ++w)
        {
          rec[2 * at + w] = r[2 * j + w];
          nrec[2 * at + w] = nr[2 * j + w];
          erec[4 * at + 2 * w] = er[4 * j + 2 * w];
          erec[4 * at + 2 * w + 1] = er[4 * j + 2 * w + 1];
        }
        ++at;
      });
    }
    // ---- the fast path of k_link_planes ------------------------------------------------------------------------------
    if (nl > kMaxLines || !fitsThreadSet(nl, nth))
    {
      g.declined_plane.push_back(k);
      g.declined_reason.push_back(F_CAPACITY);
      continue;
    }
    area.assign(workBytes(nl) + 64, 0xcd);
    uint8_t* base = area.data() + (16 - reinterpret_cast<uintptr_t>(area.data()) % 16) % 16;
    const Work W = carve(base, nl);
    auto declined = [&]() {
      if (!W.misc[M_FAIL])
        return false;
      g.declined_plane.push_back(k);
      g.declined_reason.push_back(W.misc[M_FAIL]);
      g.declined_misc.insert(g.declined_misc.end(), W.misc, W.misc + 8);
      return true;
    };
    unsigned phase = 0;
    auto run = [&](auto fn) { scrambled(nth, seed + 17u * ++phase, [&](unsigned tid) { fn(tid); }); };
    std::vector<unsigned> todo(nth, 0u);  // (a register of each thread on the device)
    run([&](unsigned tid) { phaseInit<HostOps>(W, tid, nth); });
    run([&](unsigned tid) { todo[tid] = phaseHash<HostOps>(W, rec.data(), tid, nth); });
    // (at least the round whose claim rode along with the hashing — the device always settles that one; also: few plain
    //  rounds, the rest left to the compare-and-swap loop)
    const unsigned plain_rounds = (seed & 1u) ? kPlainRounds : 1u + (seed % 6u);
    for (unsigned round = 0; round < plain_rounds; ++round)
    {
      if (round)  // (the first claim is part of phaseHash)
        run([&](unsigned tid) { phaseClaim<HostOps>(W, todo[tid], tid, nth); });
      bool any = false;
      run([&](unsigned tid) {
        todo[tid] = phaseSettle<HostOps>(W, todo[tid], tid, nth);
        any = any || todo[tid] != 0u;
      });
      g.max_rounds = std::max(g.max_rounds, round + 1);
      if (!any)
        break;
    }
    for (unsigned tid = 0; tid < nth; ++tid)
      g.stragglers += static_cast<unsigned>(__builtin_popcount(todo[tid]));
    run([&](unsigned tid) { phaseStragglers<HostOps>(W, todo[tid], tid, nth); });
    run([&](unsigned tid) { phasePair<HostOps>(W, tid, nth); });
    if (W.misc[M_ANYDEAD])
      run([&](unsigned tid) { phaseDead<HostOps>(W, tid, nth); });
    if (declined())
      continue;
    run([&](unsigned tid) { phaseLink<HostOps>(W, tid, nth); });
    if (declined())
      continue;
    run([&](unsigned tid) { phaseWalk<HostOps>(W, tid, nth); });
    for (unsigned round = 0, rounds = jumpRounds(W.misc[M_NS]); round < rounds; ++round)
    {
      bool pending = false;
      run([&](unsigned tid) { pending = phaseJump<HostOps>(W, tid, nth) || pending; });
      if (!pending)
        break;
    }
    run([&](unsigned tid) { phaseChains<HostOps>(W, tid, nth); });
    if (declined())
      continue;
    run([&](unsigned tid) { phaseHeads<HostOps>(W, rec.data(), L.dir, tid, nth); });
    if (declined())
      continue;
    run([&](unsigned tid) { phaseOrder<HostOps>(W, tid, nth); });
    scanChainsSerial(W);
    const unsigned Wk = W.misc[M_WK], Sk = W.misc[M_NH];
    std::vector<unsigned> tags(Wk, 0xffffffffu), lens(Sk, 0u);
    Out o{ tags.data(), lens.data() };
    run([&](unsigned tid) { phaseOutput<HostOps>(W, rec.data(), o, tid, nth); });
    if (declined())
      continue;
    // ---- k_waypoints: the first inserter's triangle contoured again, normals interpolated along the cut edge ----------
    for (unsigned i = 0; i < Wk; ++i)
    {
      const unsigned t = tags[i] >> 1;
      if (t >= F)
        return -2;  // a waypoint nobody named
      const uint32_t ids[3] = { tri[3 * t], tri[3 * t + 1], tri[3 * t + 2] };
      float4 p[3], nv[3];
      double s[3];
      for (int j = 0; j < 3; ++j)
      {
        p[j] = make_float4(xyz[3 * ids[j]], xyz[3 * ids[j] + 1], xyz[3 * ids[j] + 2], 0.f);
        nv[j] = make_float4(nrm[3 * ids[j]], nrm[3 * ids[j] + 1], nrm[3 * ids[j] + 2], 0.f);
        s[j] = planeScalar(L, org, p[j].x, p[j].y, p[j].z);
      }
      const CutPoint c = interpolateCut(contourCase(s), static_cast<int>(tags[i] & 1u), ids, p, s, nv);
      g.pos.insert(g.pos.end(), { c.x, c.y, c.z });
      g.nrm.insert(g.nrm.end(), { c.nx, c.ny, c.nz });
      if (want_edges)
      {
        g.elo.push_back(c.lo);
        g.ehi.push_back(c.hi);
      }
    }
    g.seg_len.insert(g.seg_len.end(), lens.begin(), lens.end());
    g.path_plane.push_back(k);
    g.path_segments.push_back(Sk);
  }
  return static_cast<int>(g.declined_plane.size());
}

uint64_t emu_num_paths() { return g.path_plane.size(); }
uint64_t emu_num_segments() { return g.seg_len.size(); }
uint64_t emu_num_waypoints() { return g.pos.size() / 3; }
uint64_t emu_num_lines() { return g.lines; }
uint64_t emu_num_declined() { return g.declined_plane.size(); }

void emu_export(int64_t* path_plane, uint32_t* path_segments, uint32_t* seg_len, float* pos, float* nrm, uint32_t* elo,
                uint32_t* ehi, int64_t* declined_plane, uint32_t* declined_reason)
{
  auto copy = [](auto* dst, const auto& v) {
    if (dst && !v.empty())
      std::memcpy(dst, v.data(), v.size() * sizeof(v[0]));
  };
  copy(path_plane, g.path_plane);
  copy(path_segments, g.path_segments);
  copy(seg_len, g.seg_len);
  copy(pos, g.pos);
  copy(nrm, g.nrm);
  copy(elo, g.elo);
  copy(ehi, g.ehi);
  copy(declined_plane, g.declined_plane);
  copy(declined_reason, g.declined_reason);
}

uint64_t emu_work_bytes(unsigned n) { return workBytes(n); }

/** cutBothPositions (what k_emit_hits contours with) against interpolateCut (the statement of vtkTriangle::Contour the
 *  waypoint pass and this replay use), bit for bit, on `n` random triangles per contour case incl. vertices exactly in the
 *  plane; returns the number of mismatches */
uint64_t emu_check_cut_both(unsigned seed, unsigned n)
{
  uint64_t bad = 0;
  uint64_t x = 0x9E3779B97F4A7C15ull * (seed + 1);
  auto rnd = [&]() {
    x ^= x << 13, x ^= x >> 7, x ^= x << 17;
    return static_cast<double>(x >> 11) / 9007199254740992.0;  // [0, 1)
  };
  const uint32_t ids[3] = { 10, 20, 30 };
  for (unsigned it = 0; it < n; ++it)
    for (int index = 1; index <= 6; ++index)
    {
      float4 p[3];
      double s[3];
      for (int v = 0; v < 3; ++v)
      {
        p[v] = make_float4(static_cast<float>(rnd() * 20 - 10), static_cast<float>(rnd() * 20 - 10), static_cast<float>(rnd() * 20 - 10), 0.f);
        const bool inside = (index >> v) & 1;
        double mag = rnd() < 0.15 ? 0.0 : rnd() * (rnd() < 0.3 ? 1e-9 : 3.0);
        if (!inside && mag == 0.0)
          mag = 1e-300;  // outside means strictly negative
        s[v] = inside ? mag : -mag;
      }
      if (contourCase(s) != index)
        continue;
      float c0[3], c1[3];
      cutBothPositions(index, p, s, c0, c1);
      const CutPoint r0 = interpolateCut(index, 0, ids, p, s, nullptr), r1 = interpolateCut(index, 1, ids, p, s, nullptr);
      const float ref[6] = { r0.x, r0.y, r0.z, r1.x, r1.y, r1.z }, got[6] = { c0[0], c0[1], c0[2], c1[0], c1[1], c1[2] };
      if (std::memcmp(ref, got, sizeof(ref)) != 0)
        ++bad;
    }
  return bad;
}
unsigned emu_max_rounds() { return g.max_rounds; }
uint64_t emu_stragglers() { return g.stragglers; }
void emu_declined_misc(uint32_t* out) { std::memcpy(out, g.declined_misc.data(), 4 * g.declined_misc.size()); }
}
