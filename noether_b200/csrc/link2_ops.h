// Per-plane linking of cut lines into tool-path segments, second generation: the phases of the fast
// path of k_link_planes, written once as __host__ __device__ functions over (tid, nth).  The kernel (kernels_slice.cu) runs
// them with one CTA per plane and a block barrier between phases; tests/cpp/link_emulation.cpp runs the very same
// functions with a loop over tid per phase, so the logic is compared with the oracle bit for bit on a machine without a GPU.
//
// What is computed (the reference: vtkCutter's point merging, vtkStripper, processPolyLines —
// plane_slicer_raster_planner.cpp:183-378,600-618): the n cut lines of one lattice plane, each given by two endpoint
// records {x, y, z, rank} (rank = 2 * triangle + line end = the order in which vtkCutter would have inserted the point),
// become maximal chains; a chain runs along cut_direction; chains are ordered by (first point . cut_direction, rank of the
// first line end); every waypoint carries the position and the interpolated normal of its FIRST INSERTER (the merged
// endpoint with the smallest rank: vtkMergePoints keeps the first point, vtkCutter interpolates attributes only then).
//
// How (differences from the first-generation fast path, profiles/r01_summary.md):
//   * cut points are merged through a 4-byte-per-slot hash {16-bit fingerprint, endpoint} WITHOUT staging the positions in
//     shared memory: a fingerprint match is confirmed against the holder's record in global memory (L1 / L2 resident: the
//     CTA has just streamed the plane's records).  A point has two line ends on a manifold cut, so the second arrival IS
//     the pairing: no pairing cells, no per-point atomicMin (the first inserter of a waypoint is the smaller rank of the
//     two merged ends, decided at output time).  One shared-memory CAS per cut point instead of about five atomics.
//   * a line gets its direction from geometry ((p1 - p0) . cut_direction), so only ONE dart per line is ranked instead
//     of two; a chain whose lines do not all agree (an overhang with respect to cut_direction) is left to the general path.
//   * list ranking (Helman & JaJa) ranks FROM THE HEAD: splitters every 8th line and chain heads, sublist walks, pointer
//     jumping on {previous splitter, root, distance} packed in one 64-bit word — position in chain and chain identity come
//     out together, no per-dart rank / last arrays.
//   * the work area is 19 bytes per line plus 4-byte slots (about 78 KB for a 4000-line plane, against 213 KB), so two
//     planes are resident per SM and their phases overlap; the output phase streams the plane's position and normal
//     records once and writes the FINAL waypoints (no tag array, no waypoint kernel).
// Anything else — a cut point with more than two line ends or a zero-length line (plane through a mesh vertex), closed
// contours, non-monotone chains, ties — makes the fast path return false BEFORE it writes any output, and the plane is
// linked by the general path (linkPlane), which remains the specification.
#pragma once
#include "host_math.h"

#include <cstddef>
#include <cstdint>

namespace nb2
{
namespace link2
{
/** 16-byte record: endpoint {x, y, z, rank bits} or normal {nx, ny, nz, 0} (same layout as float4) */
struct Rec4
{
  float x, y, z, w;
};

constexpr unsigned kNone = 0xffffu;
constexpr unsigned kEmpty = 0xffffffffu;
constexpr unsigned kSplitMask = 7u;
constexpr unsigned kRetryCap = 32u;
constexpr unsigned kMaxLines = 32766u;  // endpoint ids are 16 bit, 0xffff = none

// misc words
enum : unsigned
{
  M_FAIL = 0,     // why the fast path gives up (bits below); nothing has been written by then
  M_NS = 1,       // splitters
  M_NH = 2,       // chain heads
  M_RETRY = 3,    // endpoints whose fingerprint match was a different point
  M_WK = 4,       // waypoints of the plane
  M_RETRY_LIST = 8,
  M_WORDS = 8 + kRetryCap
};
enum : unsigned
{
  F_DEAD = 1u,        // a zero-length line (both ends on one cut point): its ends only matter for the first-inserter rank
  F_CROWDED = 2u,     // a cut point with more than two line ends
  F_DIRECTION = 4u,   // consecutive lines of a chain disagree on the direction of travel
  F_LOOP = 8u,        // a closed contour
  F_CAPACITY = 16u,   // more splitters / heads / retries than the work area holds
  F_TIE = 32u         // a line or a chain exactly perpendicular to cut_direction
};

struct Head
{
  double key;     // first point . cut_direction
  unsigned tag;   // rank of the chain's first line end
  unsigned info;  // [15:0] root splitter, [31:16] lines in the chain
};

struct Work
{
  unsigned* slots;           // [cap]   hash phase: {fingerprint:16, endpoint:16}.  Afterwards the region holds:
  uint16_t* owner;           // [n]       splitter whose sublist holds the line
  uint16_t* hop;             // [n]       hops from that splitter
  Head* heads;               // [headCap]
  unsigned* sortedLen;       // [headCap] points of the chains in output order, then their exclusive prefix
  uint16_t* twin;            // [2 n]   the other line end at the same cut point
  uint8_t* epFlag;           // [2 n]   1: the endpoint holds its point's hash slot
  uint8_t* lineDir;          // [n]     which end the line starts at when travelled along cut_direction
  unsigned long long* ent;   // [nsCap] splitter: {previous splitter:16, root:16, lines from the root:32}
  unsigned* tail;            // [nsCap] splitter whose sublist ends the chain: {hops:16, last line end:16}
  unsigned* chain;           // [nsCap] per root: {lines:16, last line end:16}, later the chain's first waypoint (plane local)
  uint16_t* splLine;         // [nsCap]
  unsigned* misc;            // [M_WORDS]
  unsigned n, cap, nsCap, headCap;
};

NB2_HD inline unsigned slotCapacity(unsigned n)
{
  unsigned cap = 64;
  while (cap < 2u * n)
    cap <<= 1;
  return cap;
}
NB2_HD inline unsigned splitterCapacity(unsigned n) { return n / 4u + 64u; }
NB2_HD inline size_t align16(size_t x) { return (x + 15) & ~static_cast<size_t>(15); }

/** bytes of the work area of an n-line plane */
NB2_HD inline size_t workBytes(unsigned n)
{
  const size_t cap = slotCapacity(n), ns = splitterCapacity(n);
  return align16(4 * cap) + align16(4 * static_cast<size_t>(n)) + align16(2 * static_cast<size_t>(n)) + align16(n) + align16(8 * ns) +
         align16(4 * ns) + align16(4 * ns) + align16(2 * ns) + align16(4 * M_WORDS);
}

NB2_HD inline Work carve(uint8_t* base, unsigned n)
{
  Work W;
  W.n = n;
  W.cap = slotCapacity(n);
  W.nsCap = splitterCapacity(n);
  uint8_t* p = base;
  W.slots = reinterpret_cast<unsigned*>(p);
  W.owner = reinterpret_cast<uint16_t*>(p);
  W.hop = W.owner + n;
  const size_t lines_bytes = align16(4 * static_cast<size_t>(n));
  const size_t region = align16(4 * static_cast<size_t>(W.cap));
  // heads (16 bytes) and sortedLen (4 bytes) share what the owner / hop arrays leave of the slot region
  W.headCap = static_cast<unsigned>((region - lines_bytes) / 20u);
  if (W.headCap > W.nsCap)
    W.headCap = W.nsCap;
  W.heads = reinterpret_cast<Head*>(p + lines_bytes);
  W.sortedLen = reinterpret_cast<unsigned*>(p + lines_bytes + align16(16 * static_cast<size_t>(W.headCap)));
  p += region;
  W.twin = reinterpret_cast<uint16_t*>(p);
  p += align16(4 * static_cast<size_t>(n));
  W.epFlag = p;
  p += align16(2 * static_cast<size_t>(n));
  W.lineDir = p;
  p += align16(n);
  W.ent = reinterpret_cast<unsigned long long*>(p);
  p += align16(8 * static_cast<size_t>(W.nsCap));
  W.tail = reinterpret_cast<unsigned*>(p);
  p += align16(4 * static_cast<size_t>(W.nsCap));
  W.chain = reinterpret_cast<unsigned*>(p);
  p += align16(4 * static_cast<size_t>(W.nsCap));
  W.splLine = reinterpret_cast<uint16_t*>(p);
  p += align16(2 * static_cast<size_t>(W.nsCap));
  W.misc = reinterpret_cast<unsigned*>(p);
  return W;
}

/** hash of a float3 (-0 folded onto +0: vtkMergePoints compares with operator==); upper half = fingerprint */
NB2_HD inline unsigned hashPoint(float x, float y, float z)
{
  union
  {
    float f;
    unsigned u;
  } a, b, c;
  a.f = x == 0.f ? 0.f : x;
  b.f = y == 0.f ? 0.f : y;
  c.f = z == 0.f ? 0.f : z;
  unsigned h = a.u * 0x9E3779B1u + b.u * 0x85EBCA77u + c.u * 0xC2B2AE3Du;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  return h ^ (h >> 13);
}
NB2_HD inline unsigned bitsOf(float f)
{
  union
  {
    float f;
    unsigned u;
  } v;
  v.f = f;
  return v.u;
}
NB2_HD inline bool samePoint(const Rec4& a, const Rec4& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

/** ((b - a) . d) with float -> double operands, left to right: the orientation / ordering arithmetic of the general path
 *  (device_math.cuh dotLR; the library is built without FMA contraction on both sides) */
NB2_HD inline double dotDiff(const Rec4& a, const Rec4& b, const double d[3])
{
  const double x = static_cast<double>(b.x) - static_cast<double>(a.x), y = static_cast<double>(b.y) - static_cast<double>(a.y),
               z = static_cast<double>(b.z) - static_cast<double>(a.z);
  return (x * d[0] + y * d[1]) + z * d[2];
}
NB2_HD inline double dotPoint(const Rec4& a, const double d[3])
{
  return (static_cast<double>(a.x) * d[0] + static_cast<double>(a.y) * d[1]) + static_cast<double>(a.z) * d[2];
}

constexpr unsigned long long kEntHi = 0xffffffff00000000ull;
NB2_HD inline unsigned long long packEnt(unsigned prev, unsigned root, unsigned dist)
{
  return static_cast<unsigned long long>(prev | (root << 16)) | (static_cast<unsigned long long>(dist) << 32);
}
NB2_HD inline unsigned entPrev(unsigned long long e) { return static_cast<unsigned>(e) & 0xffffu; }
NB2_HD inline unsigned entRoot(unsigned long long e) { return (static_cast<unsigned>(e) >> 16) & 0xffffu; }
NB2_HD inline unsigned entDist(unsigned long long e) { return static_cast<unsigned>(e >> 32); }

// ---------------------------------------------------------------------------------------------------------------------
// phases.  X supplies the memory operations: cas / add / orr on 32-bit words, vload (a load that sees other threads'
// stores of the same phase), load (a 16-byte record from global memory), ld64 / st64.
// ---------------------------------------------------------------------------------------------------------------------
template <class X>
NB2_HD inline void phaseInit(const Work& W, unsigned tid, unsigned nth)
{
  for (unsigned s = tid; s < W.cap; s += nth)
    W.slots[s] = kEmpty;
  unsigned* tw = reinterpret_cast<unsigned*>(W.twin);  // two 16-bit entries per word
  for (unsigned j = tid; j < W.n; j += nth)
    tw[j] = 0xffffffffu;
  uint16_t* fl = reinterpret_cast<uint16_t*>(W.epFlag);
  for (unsigned j = tid; j < W.n; j += nth)
    fl[j] = 0;
  if (tid < M_WORDS)
    W.misc[tid] = 0;
}

/** one endpoint into the hash: claims an empty slot (it then HOLDS the point) or meets a slot with its fingerprint (the
 *  holder becomes its candidate twin; phaseVerify confirms the coordinates) */
template <class X>
NB2_HD inline void insertEndpoint(const Work& W, unsigned e, const Rec4& r)
{
  const unsigned mask = W.cap - 1u;
  const unsigned h = hashPoint(r.x, r.y, r.z);
  const unsigned fp = h >> 16;
  unsigned s = h & mask;
  for (;;)
  {
    unsigned v = X::vload(W.slots + s);
    if (v == kEmpty)
    {
      v = X::cas(W.slots + s, kEmpty, (fp << 16) | e);
      if (v == kEmpty)
      {
        W.epFlag[e] = 1;
        return;
      }
    }
    if ((v >> 16) == fp)
    {
      W.twin[e] = static_cast<uint16_t>(v & 0xffffu);
      return;
    }
    s = (s + 1u) & mask;
  }
}

template <class X>
NB2_HD inline void phaseInsert(const Work& W, const Rec4* rec, unsigned tid, unsigned nth)
{
  for (unsigned j = tid; j < W.n; j += nth)
  {
    const Rec4 r0 = X::load(rec + 2 * j), r1 = X::load(rec + 2 * j + 1);
    if (samePoint(r0, r1))
    {
      X::orr(W.misc + M_FAIL, F_DEAD);
      continue;
    }
    insertEndpoint<X>(W, 2 * j, r0);
    insertEndpoint<X>(W, 2 * j + 1, r1);
  }
}

/** candidates confirmed against their holder's coordinates; the line's direction of travel */
template <class X>
NB2_HD inline void phaseVerify(const Work& W, const Rec4* rec, const double dir[3], unsigned tid, unsigned nth)
{
  for (unsigned j = tid; j < W.n; j += nth)
  {
    const Rec4 r0 = X::load(rec + 2 * j), r1 = X::load(rec + 2 * j + 1);
    // both candidates' holder records are requested before either is looked at
    const bool c0 = !W.epFlag[2 * j], c1 = !W.epFlag[2 * j + 1];
    const unsigned g0 = W.twin[2 * j], g1 = W.twin[2 * j + 1];
    Rec4 h0 = r0, h1 = r1;
    if (c0)
      h0 = X::load(rec + g0);
    if (c1)
      h1 = X::load(rec + g1);
    const double d = dotDiff(r0, r1, dir);
    if (d == 0.0)
      X::orr(W.misc + M_FAIL, F_TIE);
    W.lineDir[j] = d > 0.0 ? 0 : 1;
    for (unsigned w = 0; w < 2; ++w)
    {
      if (!(w ? c1 : c0))
        continue;
      const unsigned e = 2 * j + w, g = w ? g1 : g0;
      if (samePoint(w ? r1 : r0, w ? h1 : h0))
        W.twin[g] = static_cast<uint16_t>(e);  // (a second confirmed candidate overwrites: phaseLink notices)
      else
      {
        W.twin[e] = static_cast<uint16_t>(kNone);
        const unsigned i = X::add(W.misc + M_RETRY, 1u);
        if (i < kRetryCap)
          W.misc[M_RETRY_LIST + i] = e;
        else
          X::orr(W.misc + M_FAIL, F_CAPACITY);
      }
    }
  }
}

/** rare: endpoints that met another point's fingerprint go on probing, coordinates checked at every match */
template <class X>
NB2_HD inline void phaseRetry(const Work& W, const Rec4* rec, unsigned tid, unsigned nth)
{
  const unsigned count = W.misc[M_RETRY] < kRetryCap ? W.misc[M_RETRY] : kRetryCap;
  const unsigned mask = W.cap - 1u;
  for (unsigned i = tid; i < count; i += nth)
  {
    const unsigned e = W.misc[M_RETRY_LIST + i];
    const Rec4 r = X::load(rec + e);
    const unsigned h = hashPoint(r.x, r.y, r.z);
    const unsigned fp = h >> 16;
    unsigned s = h & mask;
    for (;;)
    {
      unsigned v = X::vload(W.slots + s);
      if (v == kEmpty)
      {
        v = X::cas(W.slots + s, kEmpty, (fp << 16) | e);
        if (v == kEmpty)
        {
          W.epFlag[e] = 1;
          break;
        }
      }
      if ((v >> 16) == fp)
      {
        const unsigned g = v & 0xffffu;
        if (samePoint(r, X::load(rec + g)))
        {
          W.twin[e] = static_cast<uint16_t>(g);
          W.twin[g] = static_cast<uint16_t>(e);
          break;
        }
      }
      s = (s + 1u) & mask;
    }
  }
}

/** twins must be mutual (else the point has a third line end), the next line must be entered at its start; splitters */
template <class X>
NB2_HD inline void phaseLink(const Work& W, unsigned tid, unsigned nth)
{
  for (unsigned j = tid; j < W.n; j += nth)
  {
    const unsigned s = 2 * j + W.lineDir[j], t = s ^ 1u;
    const unsigned ps = W.twin[s], pt = W.twin[t];
    if ((ps != kNone && W.twin[ps] != s) || (pt != kNone && W.twin[pt] != t))
      X::orr(W.misc + M_FAIL, F_CROWDED);
    if (pt != kNone && (pt & 1u) != W.lineDir[pt >> 1])
      X::orr(W.misc + M_FAIL, F_DIRECTION);
    if (ps != kNone && (ps & 1u) == W.lineDir[ps >> 1])
      X::orr(W.misc + M_FAIL, F_DIRECTION);
    if (ps == kNone || (j & kSplitMask) == 0u)
    {
      const unsigned c = X::add(W.misc + M_NS, 1u);
      if (c < W.nsCap)
      {
        W.splLine[c] = static_cast<uint16_t>(j);
        W.owner[j] = static_cast<uint16_t>(c);
        W.hop[j] = 0;
        X::st64(W.ent + c, packEnt(kNone, c, 0u));
        W.tail[c] = kEmpty;
        W.chain[c] = kEmpty;
      }
      else
        X::orr(W.misc + M_FAIL, F_CAPACITY);
    }
    else
      W.owner[j] = static_cast<uint16_t>(kNone);
  }
}

/** every splitter walks its sublist up to the next splitter (which learns its predecessor) or to the end of the chain */
template <class X>
NB2_HD inline void phaseWalk(const Work& W, unsigned tid, unsigned nth)
{
  const unsigned Ns = W.misc[M_NS];
  for (unsigned c = tid; c < Ns; c += nth)
  {
    unsigned j = W.splLine[c], h = 0;
    for (unsigned guard = 0; guard <= W.n; ++guard)
    {
      const unsigned t = 2 * j + (1u - W.lineDir[j]);
      const unsigned nx = W.twin[t];
      if (nx == kNone)
      {
        W.tail[c] = (h << 16) | t;
        break;
      }
      const unsigned nj = nx >> 1;
      if ((nj & kSplitMask) == 0u)  // a splitter (a chain head has no predecessor, so it is never met here)
      {
        const unsigned cn = W.owner[nj];
        X::st64(W.ent + cn, packEnt(c, cn, h + 1u));
        break;
      }
      W.owner[nj] = static_cast<uint16_t>(c);
      W.hop[nj] = static_cast<uint16_t>(h + 1u);
      j = nj;
      ++h;
    }
  }
}

/** one round of pointer jumping towards the chain heads, in place (an entry is one 64-bit word, so whichever version of
 *  its predecessor a thread reads is a consistent triple); returns whether this thread saw unfinished work */
template <class X>
NB2_HD inline bool phaseJump(const Work& W, unsigned tid, unsigned nth)
{
  const unsigned Ns = W.misc[M_NS];
  bool pending = false;
  for (unsigned c = tid; c < Ns; c += nth)
  {
    const unsigned long long me = X::ld64(W.ent + c);
    const unsigned pv = entPrev(me);
    if (pv != kNone)
    {
      const unsigned long long o = X::ld64(W.ent + pv);
      X::st64(W.ent + c, (o & 0xffffffffull) | ((me & kEntHi) + (o & kEntHi)));
      pending = pending || entPrev(o) != kNone;
    }
  }
  return pending;
}
NB2_HD inline unsigned jumpRounds(unsigned Ns)
{
  unsigned rounds = 1;
  while ((1u << rounds) < Ns + 1u)
    ++rounds;
  return rounds + 1u;
}

/** closed contours (a splitter that never reached a head, a line no splitter reached); chain lengths and last ends */
template <class X>
NB2_HD inline void phaseChains(const Work& W, unsigned tid, unsigned nth)
{
  const unsigned Ns = W.misc[M_NS];
  for (unsigned c = tid; c < Ns; c += nth)
  {
    const unsigned long long e = X::ld64(W.ent + c);
    if (entPrev(e) != kNone)
    {
      X::orr(W.misc + M_FAIL, F_LOOP);
      continue;
    }
    const unsigned tl = W.tail[c];
    if (tl != kEmpty)
    {
      const unsigned lines = entDist(e) + (tl >> 16) + 1u;
      W.chain[entRoot(e)] = (lines << 16) | (tl & 0xffffu);
    }
  }
  for (unsigned j = tid; j < W.n; j += nth)
    if (W.owner[j] == kNone)
      X::orr(W.misc + M_FAIL, F_LOOP);
}

/** chain heads: the chain must run along cut_direction from its head (first -> last point), as its lines do */
template <class X>
NB2_HD inline void phaseHeads(const Work& W, const Rec4* rec, const double dir[3], unsigned tid, unsigned nth)
{
  const unsigned Ns = W.misc[M_NS];
  for (unsigned c = tid; c < Ns; c += nth)
  {
    if (entRoot(X::ld64(W.ent + c)) != c)
      continue;
    const unsigned j = W.splLine[c];
    const unsigned s = 2 * j + W.lineDir[j];
    const unsigned info = W.chain[c];
    if (info == kEmpty)  // (cannot happen on an open chain; a loop has been flagged already)
    {
      X::orr(W.misc + M_FAIL, F_LOOP);
      continue;
    }
    const Rec4 a = X::load(rec + s), b = X::load(rec + (info & 0xffffu));
    if (!(dotDiff(a, b, dir) > 0.0))
      X::orr(W.misc + M_FAIL, F_TIE);
    const unsigned h = X::add(W.misc + M_NH, 1u);
    if (h < W.headCap)
    {
      W.heads[h].key = dotPoint(a, dir);
      W.heads[h].tag = bitsOf(a.w);
      W.heads[h].info = c | (info & 0xffff0000u);
    }
    else
      X::orr(W.misc + M_FAIL, F_CAPACITY);
  }
}

/** order of the chains: (first point . cut_direction, rank of the first line end) */
template <class X>
NB2_HD inline void phaseOrder(const Work& W, unsigned tid, unsigned nth)
{
  const unsigned S = W.misc[M_NH];
  for (unsigned h = tid; h < S; h += nth)
  {
    unsigned ord = 0;
    const double key = W.heads[h].key;
    const unsigned tg = W.heads[h].tag;
    for (unsigned g = 0; g < S; ++g)
    {
      const double kg = W.heads[g].key;
      ord += (kg < key) || (kg == key && W.heads[g].tag < tg);
    }
    W.sortedLen[ord] = (W.heads[h].info >> 16) + 1u;  // lines + 1 points
    W.chain[W.heads[h].info & 0xffffu] = ord;          // root -> place of its chain
  }
}

/** exclusive prefix of the chain sizes (serial form; the kernel does the same with one warp) */
NB2_HD inline void scanChainsSerial(const Work& W)
{
  const unsigned S = W.misc[M_NH];
  unsigned acc = 0;
  for (unsigned i = 0; i < S; ++i)
  {
    const unsigned v = W.sortedLen[i];
    W.sortedLen[i] = acc;
    acc += v;
  }
  W.misc[M_WK] = acc;
}

/** where the plane's results go: final arrays (the plane's first waypoint / segment already added to the pointers) */
struct Out
{
  float* pos;        // [3 W]
  float* nrm;        // [3 W]
  uint32_t* edgeLo;  // [W] or null
  uint32_t* edgeHi;
  unsigned* segLen;  // [S]
};

template <class X>
NB2_HD inline void writePoint(const Out& o, unsigned i, const Rec4& p, const Rec4& nv, const uint32_t* edge)
{
  o.pos[3 * i] = p.x;
  o.pos[3 * i + 1] = p.y;
  o.pos[3 * i + 2] = p.z;
  o.nrm[3 * i] = nv.x;
  o.nrm[3 * i + 1] = nv.y;
  o.nrm[3 * i + 2] = nv.z;
  if (o.edgeLo)
  {
    o.edgeLo[i] = edge[0];
    o.edgeHi[i] = edge[1];
  }
}

/** Every line streams its own records once and writes the waypoints it is the first inserter of: its start point is
 *  waypoint p of its chain, its end point waypoint p + 1; of two merged line ends the smaller rank writes.
 *  erec: generating mesh edge {lo, hi} of every endpoint (only when edges are wanted). */
template <class X>
NB2_HD inline void phaseOutput(const Work& W, const Rec4* rec, const Rec4* nrec, const uint32_t* erec, const Out& o, unsigned tid,
                               unsigned nth)
{
  const unsigned S = W.misc[M_NH];
  for (unsigned h = tid; h < S; h += nth)
    o.segLen[W.chain[W.heads[h].info & 0xffffu]] = (W.heads[h].info >> 16) + 1u;
  for (unsigned j = tid; j < W.n; j += nth)
  {
    const unsigned long long e = X::ld64(W.ent + W.owner[j]);
    const unsigned first = W.sortedLen[W.chain[entRoot(e)]] + entDist(e) + W.hop[j];
    const unsigned s = 2 * j + W.lineDir[j], t = s ^ 1u;
    const unsigned ps = W.twin[s], pt = W.twin[t];
    // the ranks of the two partners are requested together with the line's own records
    const Rec4 r0 = X::load(rec + 2 * j), r1 = X::load(rec + 2 * j + 1);
    unsigned rank_ps = kEmpty, rank_pt = kEmpty;
    if (ps != kNone)
      rank_ps = X::loadRank(rec + ps);
    if (pt != kNone)
      rank_pt = X::loadRank(rec + pt);
    const Rec4& rs = (s & 1u) ? r1 : r0;
    const Rec4& rt = (t & 1u) ? r1 : r0;
    const bool mine_s = bitsOf(rs.w) < rank_ps, mine_t = bitsOf(rt.w) < rank_pt;  // (ranks are below 2^31: kEmpty loses)
    if (mine_s)
      writePoint<X>(o, first, rs, X::load(nrec + s), erec ? erec + 2 * static_cast<size_t>(s) : nullptr);
    if (mine_t)
      writePoint<X>(o, first + 1u, rt, X::load(nrec + t), erec ? erec + 2 * static_cast<size_t>(t) : nullptr);
  }
}

}  // namespace link2
}  // namespace nb2
