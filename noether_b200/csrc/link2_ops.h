// Per-plane linking of cut lines into tool-path segments, second generation: the phases of the fast
// path of k_link_planes, written once as __host__ __device__ functions over (tid, nth).  The kernel (kernels_slice.cu) runs
// them with one CTA per plane and a block barrier between phases; tests/cpp/link_emulation.cpp runs the very same
// functions with a loop over tid per phase, so the logic is compared with the oracle bit for bit on a machine without a GPU.
//
// What is computed (the reference: vtkCutter's point merging, vtkStripper, processPolyLines —
// plane_slicer_raster_planner.cpp:183-378,600-618): the n cut lines of one lattice plane, each given by two endpoint
// records {x, y, z, rank} (rank = 2 * triangle + line end = the order in which vtkCutter would have inserted the point),
// become maximal chains; a chain runs along cut_direction; chains are ordered by (first point . cut_direction, rank of the
// first line end); every waypoint is named by the rank of its FIRST INSERTER (the merged endpoint with the smallest rank:
// vtkMergePoints keeps the first point, vtkCutter interpolates attributes only then), from which k_waypoints produces
// the position and the interpolated normal.
//
// How (differences from the first-generation fast path; measurements in profiles/r02_summary.md):
//   * NO SHARED-MEMORY ATOMICS on the common path (the first generation issues three per endpoint; whether they are what
//     it waits for was not confirmed: profiles/r02_summary.md).  Cut points are merged through a hash whose slots are claimed with
//     plain stores in rounds: every unsettled endpoint writes {fingerprint, endpoint} into its slot if the slot looks
//     empty; after a barrier it reads the slot back — its own value: it holds the point; the same 32-bit hash: the holder
//     is its twin; something else: next slot, next round.  Whichever store wins, everybody sees the same winner.
//   * the hash values of the endpoints are staged in shared memory (4 bytes per endpoint), not the coordinates (16): the
//     plane's records are streamed once for the hashing and once, with each line's two partners, at output time, where
//     the coordinates of merged endpoints are compared for real (a 32-bit collision sends the plane to the general path).
//   * a point has two line ends on a manifold cut, so the second arrival IS the pairing: no pairing cells, no per-point
//     atomicMin (the first inserter of a waypoint is the smaller rank of the two merged ends, decided at output time).
//   * vtkTriangle's case table orients every cut line consistently with the triangle's winding (the inside of the plane
//     on a fixed side), so on a consistently wound mesh line end 1 always meets line end 0 of the next line: only ONE dart
//     per line (end 0 -> end 1) is ranked instead of two, and whether a chain is output in that direction or reversed is
//     decided once per chain, at its head, with the general path's rule.  (Deciding from coordinates does not work: where
//     a plane grazes a row of mesh vertices, rounding makes sliver lines point backwards.)  A plane on which some line end
//     1 meets another end 1 (inconsistent winding) is left to the general path.
//   * zero-length lines (both ends on one cut point: the plane passes through, or within rounding of, a mesh vertex) take
//     no part in the pairing, as in vtkTriangle::Contour, but their ranks count for the first inserter of their point:
//     they look their point up after the pairing and leave an override the output phase honours.
//   * list ranking (Helman & JaJa) ranks FROM THE HEAD: splitters every 8th line and chain heads, sublist walks, pointer
//     jumping on {previous splitter, root, distance} packed in one 64-bit word — position in chain and chain identity come
//     out together, no per-dart rank / last arrays.
//   * the work area is about 26 bytes per line (104 KB for a 4000-line plane, against 213 KB), so two planes are resident
//     per SM and each hides the other's barriers and global-memory phases.
// Anything else — a cut point with more than two proper line ends, closed contours, inconsistent winding, a hash collision,
// more chains / zero-length lines than the work area holds — makes the fast path return false, and the plane is linked by
// the general path (linkPlane), which remains the specification.  (Whatever the fast path wrote by then lies in the
// plane's own output windows and is overwritten.)
#pragma once
#include "host_math.h"

// every function here must end up inlined into the kernel: the work area's pointers then stay in registers and the
// compiler sees that they point into shared memory (LDS / STS / ATOMS instead of generic accesses through a struct in
// local memory)
#if defined(__CUDACC__)
#define NB2_HDI __host__ __device__ __forceinline__
#else
#define NB2_HDI inline
#endif

#include <cstddef>
#include <cstdint>

namespace nb2
{
namespace link2
{
/** 16-byte record: endpoint {x, y, z, rank bits} (same layout as float4) */
struct Rec4
{
  float x, y, z, w;
};

constexpr unsigned kNone = 0xffffu;
constexpr unsigned kEmpty = 0xffffffffu;
constexpr unsigned kSplitMask = 7u;
constexpr unsigned kDeadCap = 96u;      // zero-length lines whose point is used by a chain, per plane
constexpr unsigned kMaxLines = 32700u;  // endpoint ids and slot numbers are 16 bit, 0xffff = none

// misc words
enum : unsigned
{
  M_FAIL = 0,     // why the fast path gives up (bits below)
  M_NS = 1,       // splitters
  M_NH = 2,       // chain heads
  M_WK = 4,       // waypoints of the plane
  M_DEAD = 5,     // overrides left by zero-length lines
  M_ANYDEAD = 6,  // the plane has zero-length lines
  M_DEAD_LIST = 8,  // pairs {holder of the point, line end 0 of the zero-length line}
  M_WORDS = M_DEAD_LIST + 2 * kDeadCap
};
enum : unsigned
{
  F_COLLISION = 1u,   // two different cut points with the same 32-bit hash
  F_CROWDED = 2u,     // a cut point with more than two proper line ends
  F_DIRECTION = 4u,   // a line end 1 meets a line end 1 (or 0 meets 0): the triangles are not consistently wound
  F_LOOP = 8u,        // a closed contour
  F_CAPACITY = 16u    // more splitters / heads / zero-length lines than the work area holds
};
// endpoint states
enum : unsigned
{
  E_UNSETTLED = 0u,
  E_HOLDER = 1u,     // holds its point's hash slot
  E_CANDIDATE = 2u,  // met a holder with its hash: twin = that holder
  E_DEAD = 4u        // end of a zero-length line
};

struct Head
{
  double key;     // first point . cut_direction
  unsigned tag;   // rank of the chain's first line end
  unsigned info;  // [15:0] root splitter, [30:16] lines in the chain, [31] output reversed (from line end 1 of its last line)
};

struct Work
{
  unsigned* slots;           // [cap]   hash phase: {fingerprint:16, endpoint:16}.  Afterwards the region holds:
  Head* heads;               // [headCap]
  unsigned* sortedLen;       // [headCap] points of the chains in output order, then their exclusive prefix
  unsigned* hsh;             // [2 n]   hash phase: 32-bit hash of every endpoint's coordinates.  Afterwards the region holds:
  uint16_t* owner;           // [n]       splitter whose sublist holds the line
  uint16_t* hop;             // [n]       hops from that splitter
  unsigned long long* ent;   // [nsCap]   splitter: {previous splitter:16, root:16, lines from the root:32}
  uint16_t* twin;            // [2 n]   hash phase: the slot an unsettled endpoint probes next; then the other line end at the point
  uint8_t* state;            // [2 n]   E_*
  uint8_t* lineDead;         // [n]     1: zero-length line (takes no part in the chains)
  unsigned* tail;            // [nsCap] splitter whose sublist ends the chain: {hops:16, last line end:16}
  unsigned* chain;           // [nsCap] per root: {lines:15, last line end:16}, later {reversed:1, lines:15, place of the chain:16}
  uint16_t* splLine;         // [nsCap]
  unsigned* misc;            // [M_WORDS]
  unsigned n, cap, nsCap, headCap;
};

/** hash slots of an n-line plane: two per cut point (a plane has about n + chains points), not rounded to a power of
 *  two — the home slot is a multiply-shift of the hash, so the work area grows smoothly with n */
NB2_HDI unsigned slotCapacity(unsigned n) { return 2u * n + 64u; }
NB2_HDI unsigned homeSlot(unsigned h, unsigned cap)
{
  return static_cast<unsigned>((static_cast<unsigned long long>(h * 0x85EBCA6Bu) * cap) >> 32);
}
NB2_HDI unsigned splitterCapacity(unsigned n) { return n / 4u + 64u; }
NB2_HDI size_t align16(size_t x) { return (x + 15) & ~static_cast<size_t>(15); }
NB2_HDI size_t regionB(unsigned n)
{
  const size_t hashing = 8 * static_cast<size_t>(n), ranking = align16(4 * static_cast<size_t>(n)) + 8 * static_cast<size_t>(splitterCapacity(n));
  return align16(hashing > ranking ? hashing : ranking);
}

/** bytes of the work area of an n-line plane */
NB2_HDI size_t workBytes(unsigned n)
{
  const size_t ns = splitterCapacity(n);
  return align16(4 * static_cast<size_t>(slotCapacity(n))) + regionB(n) + align16(4 * static_cast<size_t>(n)) +
         align16(2 * static_cast<size_t>(n)) + align16(n) + align16(4 * ns) + align16(4 * ns) + align16(2 * ns) + align16(4 * M_WORDS);
}

NB2_HDI Work carve(uint8_t* base, unsigned n)
{
  Work W;
  W.n = n;
  W.cap = slotCapacity(n);
  W.nsCap = splitterCapacity(n);
  uint8_t* p = base;
  W.slots = reinterpret_cast<unsigned*>(p);
  const size_t regionA = align16(4 * static_cast<size_t>(W.cap));
  W.headCap = static_cast<unsigned>(regionA / 20u);  // heads (16 bytes) and sortedLen (4 bytes) share the slot region
  if (W.headCap > W.nsCap)
    W.headCap = W.nsCap;
  W.heads = reinterpret_cast<Head*>(p);
  W.sortedLen = reinterpret_cast<unsigned*>(p + align16(16 * static_cast<size_t>(W.headCap)));
  p += regionA;
  W.hsh = reinterpret_cast<unsigned*>(p);
  W.owner = reinterpret_cast<uint16_t*>(p);
  W.hop = W.owner + n;
  W.ent = reinterpret_cast<unsigned long long*>(p + align16(4 * static_cast<size_t>(n)));
  p += regionB(n);
  W.twin = reinterpret_cast<uint16_t*>(p);
  p += align16(4 * static_cast<size_t>(n));
  W.state = p;
  p += align16(2 * static_cast<size_t>(n));
  W.lineDead = p;
  p += align16(n);
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
NB2_HDI unsigned hashPoint(float x, float y, float z)
{
  union
  {
    float f;
    unsigned u;
  } a, b, c;
  a.f = x == 0.f ? 0.f : x;
  b.f = y == 0.f ? 0.f : y;
  c.f = z == 0.f ? 0.f : z;
  // (the coordinates enter rotated against one another: with plain sums the sign bits of two coordinates cancel and the
  //  points (x, y, z) and (-x, y, -z) of a point-symmetric part would always collide)
  const unsigned br = (b.u << 11) | (b.u >> 21), cr = (c.u << 22) | (c.u >> 10);
  unsigned h = a.u * 0x9E3779B1u + br * 0x85EBCA77u + cr * 0xC2B2AE3Du;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  return h ^ (h >> 13);
}
NB2_HDI unsigned bitsOf(float f)
{
  union
  {
    float f;
    unsigned u;
  } v;
  v.f = f;
  return v.u;
}
NB2_HDI bool samePoint(const Rec4& a, const Rec4& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

/** ((b - a) . d) with float -> double operands, left to right: the orientation / ordering arithmetic of the general path
 *  (device_math.cuh dotLR; the library is built without FMA contraction on both sides) */
NB2_HDI double dotDiff(const Rec4& a, const Rec4& b, const double d[3])
{
  const double x = static_cast<double>(b.x) - static_cast<double>(a.x), y = static_cast<double>(b.y) - static_cast<double>(a.y),
               z = static_cast<double>(b.z) - static_cast<double>(a.z);
  return (x * d[0] + y * d[1]) + z * d[2];
}
NB2_HDI double dotPoint(const Rec4& a, const double d[3])
{
  return (static_cast<double>(a.x) * d[0] + static_cast<double>(a.y) * d[1]) + static_cast<double>(a.z) * d[2];
}

constexpr unsigned long long kEntHi = 0xffffffff00000000ull;
NB2_HDI unsigned long long packEnt(unsigned prev, unsigned root, unsigned dist)
{
  return static_cast<unsigned long long>(prev | (root << 16)) | (static_cast<unsigned long long>(dist) << 32);
}
NB2_HDI unsigned entPrev(unsigned long long e) { return static_cast<unsigned>(e) & 0xffffu; }
NB2_HDI unsigned entRoot(unsigned long long e) { return (static_cast<unsigned>(e) >> 16) & 0xffffu; }
NB2_HDI unsigned entDist(unsigned long long e) { return static_cast<unsigned>(e >> 32); }

// ---------------------------------------------------------------------------------------------------------------------
// phases.  X supplies what differs between the device and the host replay: load (a 16-byte record from global memory),
// orr (set bits in a word other threads set bits in too), add / cas (fetch-and-add, compare-and-swap: rare events), claim (a place in a
// compacted list: one shared-memory atomic per WARP on the device, to be called by all threads of the block together),
// vload (a load that must not be hoisted), ld64 / st64 (a 64-bit word that another thread may be rewriting).  Thread `tid` of `nth` owns
// lines tid, tid + nth, ...; everything between two phases is a block barrier.
// ---------------------------------------------------------------------------------------------------------------------
template <class X>
NB2_HDI void phaseInit(const Work& W, unsigned tid, unsigned nth)
{
  for (unsigned s = tid; s < W.cap; s += nth)
    W.slots[s] = kEmpty;
  for (unsigned i = tid; i < M_WORDS; i += nth)
    W.misc[i] = 0;
}

/** The plane's records streamed once: hash of every endpoint, zero-length lines, and the first claim of every endpoint
 *  (needs the slots emptied: a barrier after phaseInit).  Returns the thread's set of unsettled endpoints (bit 2 i + w =
 *  line end w of the thread's i-th line). */
template <class X>
NB2_HDI unsigned phaseHash(const Work& W, const Rec4* rec, unsigned tid, unsigned nth)
{
  unsigned todo = 0u;
  unsigned i = 0;
  for (unsigned j = tid; j < W.n; j += nth, ++i)
  {
    const Rec4 r0 = X::load(rec + 2 * j), r1 = X::load(rec + 2 * j + 1);
    const bool dead = samePoint(r0, r1);  // vtkTriangle::Contour drops the line (pts[0] == pts[1]); see phaseDead
    const unsigned h0 = hashPoint(r0.x, r0.y, r0.z), h1 = dead ? h0 : hashPoint(r1.x, r1.y, r1.z);
    W.lineDead[j] = dead ? 1 : 0;
    W.hsh[2 * j] = h0;
    W.hsh[2 * j + 1] = h1;
    const unsigned s0 = homeSlot(h0, W.cap), s1 = homeSlot(h1, W.cap);
    W.twin[2 * j] = static_cast<uint16_t>(s0);
    W.twin[2 * j + 1] = static_cast<uint16_t>(s1);
    W.state[2 * j] = W.state[2 * j + 1] = static_cast<uint8_t>(dead ? E_DEAD : E_UNSETTLED);
    if (dead)
      W.misc[M_ANYDEAD] = 1u;
    else
    {
      todo |= 3u << (2u * i);
      // the first claim (phaseClaim) rides along: hash and slot are in registers
      if (X::vload(W.slots + s0) == kEmpty)
        W.slots[s0] = (h0 & 0xffff0000u) | (2 * j);
      if (X::vload(W.slots + s1) == kEmpty)
        W.slots[s1] = (h1 & 0xffff0000u) | (2 * j + 1);
    }
  }
  return todo;
}
/** lines a thread can own with a 32-bit set of endpoints */
NB2_HDI bool fitsThreadSet(unsigned n, unsigned nth) { return (n + nth - 1u) / nth <= 16u; }

/** index of the lowest set bit */
NB2_HDI unsigned lowestBit(unsigned x)
{
#if defined(__CUDA_ARCH__)
  return static_cast<unsigned>(__ffs(static_cast<int>(x))) - 1u;
#else
  return static_cast<unsigned>(__builtin_ctz(x));
#endif
}
constexpr unsigned kPlainRounds = 4u;  // claim rounds with plain stores; the few endpoints still unsettled then use CAS

/** round, first half: every unsettled endpoint writes itself into its slot if the slot looks empty (plain store) */
template <class X>
NB2_HDI void phaseClaim(const Work& W, unsigned todo, unsigned tid, unsigned nth)
{
  while (todo)
  {
    const unsigned b = lowestBit(todo);
    todo &= todo - 1u;
    const unsigned e = 2u * (tid + (b >> 1) * nth) + (b & 1u);
    const unsigned s = W.twin[e];
    if (X::vload(W.slots + s) == kEmpty)
      W.slots[s] = (W.hsh[e] & 0xffff0000u) | e;
  }
}
/** round, second half.  Its own value in its slot: it holds the point.  Otherwise it walks on through the occupied slots
 *  (their contents are final: slots are only ever claimed while empty) — a slot with its hash: that holder is its twin;
 *  the first empty slot: that is where it claims next round.  So a round is lost only to another endpoint that claimed
 *  the same empty slot in the same round, and the number of rounds is the longest series of lost races (a handful), not
 *  the longest probe sequence.  The table keeps the linear-probing invariant (a holder's slots from its home slot on are
 *  all occupied), which the look-ups of phaseDead rely on.  Returns the endpoints still unsettled. */
template <class X>
NB2_HDI unsigned phaseSettle(const Work& W, unsigned todo, unsigned tid, unsigned nth)
{
  unsigned left = 0u;
  while (todo)
  {
    const unsigned b = lowestBit(todo);
    todo &= todo - 1u;
    const unsigned e = 2u * (tid + (b >> 1) * nth) + (b & 1u);
    unsigned s = W.twin[e];
    const unsigned h = W.hsh[e];
    unsigned v = W.slots[s];
    if ((v & 0xffffu) == e)
    {
      W.state[e] = static_cast<uint8_t>(E_HOLDER);
      W.twin[e] = static_cast<uint16_t>(kNone);
      continue;
    }
    for (;;)
    {
      const unsigned g = v & 0xffffu;
      if (((v ^ h) >> 16) == 0u && W.hsh[g] == h)
      {
        W.state[e] = static_cast<uint8_t>(E_CANDIDATE);
        W.twin[e] = static_cast<uint16_t>(g);
        break;
      }
      s = s + 1u == W.cap ? 0u : s + 1u;
      v = W.slots[s];
      if (v == kEmpty)
      {
        W.twin[e] = static_cast<uint16_t>(s);
        left |= 1u << b;
        break;
      }
    }
  }
  return left;
}
/** What the plain rounds left over (a percent or so of the endpoints, in the longest probe sequences) finishes with
 *  compare-and-swap: settled slots never change, so the two protocols agree. */
template <class X>
NB2_HDI void phaseStragglers(const Work& W, unsigned todo, unsigned tid, unsigned nth)
{
  while (todo)
  {
    const unsigned b = lowestBit(todo);
    todo &= todo - 1u;
    const unsigned e = 2u * (tid + (b >> 1) * nth) + (b & 1u);
    const unsigned h = W.hsh[e];
    unsigned s = W.twin[e];
    for (;;)
    {
      unsigned v = X::vload(W.slots + s);
      if (v == kEmpty)
      {
        v = X::cas(W.slots + s, kEmpty, (h & 0xffff0000u) | e);
        if (v == kEmpty)
        {
          W.state[e] = static_cast<uint8_t>(E_HOLDER);
          W.twin[e] = static_cast<uint16_t>(kNone);
          break;
        }
      }
      const unsigned g = v & 0xffffu;
      if (((v ^ h) >> 16) == 0u && X::vload(W.hsh + g) == h)
      {
        W.state[e] = static_cast<uint8_t>(E_CANDIDATE);
        W.twin[e] = static_cast<uint16_t>(g);
        break;
      }
      s = s + 1u == W.cap ? 0u : s + 1u;
    }
  }
}

/** every candidate tells its holder (a second candidate of the same holder overwrites the first: phaseLink notices) */
template <class X>
NB2_HDI void phasePair(const Work& W, unsigned tid, unsigned nth)
{
  const unsigned E = 2u * W.n;
  for (unsigned e = tid; e < E; e += nth)
    if (W.state[e] == E_CANDIDATE)
      W.twin[W.twin[e]] = static_cast<uint16_t>(e);
}

/** Zero-length lines look their point up (no insertion): when a chain uses the point, the line leaves {holder, its line
 *  end 0} on the override list — vtkCutter inserted the point for this triangle too, so if its rank is the smallest the
 *  waypoint is this triangle's (phaseOutput).  Runs while the hash is still intact. */
template <class X>
NB2_HDI void phaseDead(const Work& W, unsigned tid, unsigned nth)
{
  for (unsigned j = tid; j < W.n; j += nth)
  {
    if (!W.lineDead[j])
      continue;
    const unsigned h = W.hsh[2 * j];
    unsigned s = homeSlot(h, W.cap);
    for (;;)
    {
      const unsigned v = W.slots[s];
      if (v == kEmpty)
        break;  // no line ends here: vtkStripper never sees the point
      const unsigned g = v & 0xffffu;
      if (((v ^ h) >> 16) == 0u && W.hsh[g] == h)
      {
        const unsigned i = X::add(W.misc + M_DEAD, 1u);  // (not X::claim: the threads that get here are not whole warps)
        if (i < kDeadCap)
        {
          W.misc[M_DEAD_LIST + 2 * i] = g;
          W.misc[M_DEAD_LIST + 2 * i + 1] = 2 * j;
        }
        else
          X::orr(W.misc + M_FAIL, F_CAPACITY);
        break;
      }
      s = s + 1u == W.cap ? 0u : s + 1u;
    }
  }
}

/** twins must be mutual (else the point has a third line end) and join a line end 1 to a line end 0; splitters.
 *  (The hash region becomes owner / hop / ent here: the hash is dead.)  Every thread makes the same number of trips
 *  through the loop: X::claim is a warp-wide operation on the device. */
template <class X>
NB2_HDI void phaseLink(const Work& W, unsigned tid, unsigned nth)
{
  for (unsigned j0 = 0; j0 < W.n; j0 += nth)
  {
    const unsigned j = j0 + tid;
    bool splitter = false;
    if (j < W.n && !W.lineDead[j])
    {
      const unsigned s = 2 * j, t = s + 1u;
      const unsigned ps = W.twin[s], pt = W.twin[t];
      if ((ps != kNone && W.twin[ps] != s) || (pt != kNone && W.twin[pt] != t))
        X::orr(W.misc + M_FAIL, F_CROWDED);
      if ((pt != kNone && (pt & 1u) != 0u) || (ps != kNone && (ps & 1u) == 0u))
        X::orr(W.misc + M_FAIL, F_DIRECTION);
      splitter = ps == kNone || (j & kSplitMask) == 0u;
    }
    const unsigned c = X::claim(W.misc + M_NS, splitter);
    if (j < W.n)
    {
      if (splitter && c < W.nsCap)
      {
        W.splLine[c] = static_cast<uint16_t>(j);
        W.owner[j] = static_cast<uint16_t>(c);
        W.hop[j] = 0;
        X::st64(W.ent + c, packEnt(kNone, c, 0u));
        W.tail[c] = kEmpty;
        W.chain[c] = kEmpty;
      }
      else
      {
        W.owner[j] = static_cast<uint16_t>(kNone);
        if (splitter)
          X::orr(W.misc + M_FAIL, F_CAPACITY);
      }
    }
  }
}

/** every splitter walks its sublist (line end 0 -> line end 1 -> next line) up to the next splitter, which learns its
 *  predecessor, or to the end of the chain */
template <class X>
NB2_HDI void phaseWalk(const Work& W, unsigned tid, unsigned nth)
{
  const unsigned Ns = W.misc[M_NS];
  for (unsigned c = tid; c < Ns; c += nth)
  {
    unsigned j = W.splLine[c], h = 0;
    for (unsigned guard = 0; guard <= W.n; ++guard)
    {
      const unsigned t = 2 * j + 1u;
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
NB2_HDI bool phaseJump(const Work& W, unsigned tid, unsigned nth)
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
NB2_HDI unsigned jumpRounds(unsigned Ns)
{
  unsigned rounds = 1;
  while ((1u << rounds) < Ns + 1u)
    ++rounds;
  return rounds + 1u;
}

/** closed contours (a splitter that never reached a head, a line no splitter reached); chain lengths and last ends */
template <class X>
NB2_HDI void phaseChains(const Work& W, unsigned tid, unsigned nth)
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
    if (W.owner[j] == kNone && !W.lineDead[j])
      X::orr(W.misc + M_FAIL, F_LOOP);
}

/** Chain heads.  A chain has been ranked from its open line end 0 (a) to its open line end 1 (b); it is OUTPUT along
 *  cut_direction: from a when (b - a) . cut_direction > 0, from b when < 0, and on a tie from the end with the smaller rank
 *  — the general path's rule for choosing between the two dart lists of a chain (linkPlane, "C: chain heads").
 *  Every thread makes the same number of trips through the loop (X::claim). */
template <class X>
NB2_HDI void phaseHeads(const Work& W, const Rec4* rec, const double dir[3], unsigned tid, unsigned nth)
{
  const unsigned Ns = W.misc[M_NS];
  for (unsigned c0 = 0; c0 < Ns; c0 += nth)
  {
    const unsigned c = c0 + tid;
    const bool root = c < Ns && entRoot(X::ld64(W.ent + c)) == c;
    unsigned info = kEmpty;
    if (root)
    {
      info = W.chain[c];
      if (info == kEmpty)  // (cannot happen on an open chain; a loop has been flagged already)
        X::orr(W.misc + M_FAIL, F_LOOP);
    }
    const bool head = root && info != kEmpty;
    const unsigned h = X::claim(W.misc + M_NH, head);
    if (!head)
      continue;
    if (h >= W.headCap)
    {
      X::orr(W.misc + M_FAIL, F_CAPACITY);
      continue;
    }
    const Rec4 a = X::load(rec + 2 * W.splLine[c]), b = X::load(rec + (info & 0xffffu));
    const double d = dotDiff(a, b, dir);
    const unsigned ta = bitsOf(a.w), tb = bitsOf(b.w);
    const bool forward = d > 0.0 || (d == 0.0 && ta < tb);
    W.heads[h].key = dotPoint(forward ? a : b, dir);
    W.heads[h].tag = forward ? ta : tb;
    W.heads[h].info = c | (info & 0x7fff0000u) | (forward ? 0u : 0x80000000u);
  }
}

/** order of the chains: (first point . cut_direction, rank of the first line end) */
template <class X>
NB2_HDI void phaseOrder(const Work& W, unsigned tid, unsigned nth)
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
    const unsigned info = W.heads[h].info;
    W.sortedLen[ord] = ((info >> 16) & 0x7fffu) + 1u;                             // lines + 1 points
    W.chain[info & 0xffffu] = ord | (info & 0x80000000u) | (info & 0x7fff0000u);  // root -> {place, lines, reversed}
  }
}

/** exclusive prefix of the chain sizes (serial form; the kernel does the same with one warp) */
NB2_HDI void scanChainsSerial(const Work& W)
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

/** Where the plane's results go: the rank of the first inserter of every waypoint and the lengths of the segments, in
 *  the plane's own windows of the output (a plane with n lines has at most 2 n waypoints and n segments) */
struct Out
{
  unsigned* wpTag;   // [W]
  unsigned* segLen;  // [S]
};

/** the waypoint of the cut point where line end e (record pe) meets `partner` (record pp; kNone: an open end): named by
 *  the smaller of the two ranks — unless a zero-length line at the same point was inserted even earlier.  The line end
 *  with the smaller rank writes; the two records must be the same point (else two points share a 32-bit hash). */
template <class X>
NB2_HDI void emitWaypoint(const Work& W, const Rec4* rec, const Out& o, unsigned i, unsigned e, unsigned partner, const Rec4& pe,
                          const Rec4& pp)
{
  unsigned best = bitsOf(pe.w);
  if (partner != kNone)
  {
    if (!samePoint(pe, pp))
      X::orr(W.misc + M_FAIL, F_COLLISION);
    if (!(best < bitsOf(pp.w)))
      return;
  }
  const unsigned nd = W.misc[M_DEAD] < kDeadCap ? W.misc[M_DEAD] : kDeadCap;
  if (nd)
  {
    const unsigned holder = W.state[e] == E_HOLDER ? e : partner;
    for (unsigned k = 0; k < nd; ++k)
      if (W.misc[M_DEAD_LIST + 2 * k] == holder)
      {
        const Rec4 pd = X::load(rec + W.misc[M_DEAD_LIST + 2 * k + 1]);
        if (!samePoint(pe, pd))
          X::orr(W.misc + M_FAIL, F_COLLISION);
        if (bitsOf(pd.w) < best)
          best = bitsOf(pd.w);
      }
  }
  o.wpTag[i] = best;
}

/** Every line streams its own records and those of its two partners and names the waypoints it is the first inserter
 *  of.  Line p of a chain of L lines (counted from the chain's open line end 0) starts at waypoint p and ends at waypoint
 *  p + 1 — L - p and L - p - 1 when the chain is output reversed. */
template <class X>
NB2_HDI void phaseOutput(const Work& W, const Rec4* rec, const Out& o, unsigned tid, unsigned nth)
{
  const unsigned S = W.misc[M_NH];
  for (unsigned h = tid; h < S; h += nth)
    o.segLen[W.chain[W.heads[h].info & 0xffffu] & 0xffffu] = ((W.heads[h].info >> 16) & 0x7fffu) + 1u;
  for (unsigned j = tid; j < W.n; j += nth)
  {
    if (W.lineDead[j])
      continue;
    const unsigned ps = W.twin[2 * j], pt = W.twin[2 * j + 1];
    // the four records are requested together
    const Rec4 r0 = X::load(rec + 2 * j), r1 = X::load(rec + 2 * j + 1);
    const Rec4 q0 = ps != kNone ? X::load(rec + ps) : r0, q1 = pt != kNone ? X::load(rec + pt) : r1;
    const unsigned long long e = X::ld64(W.ent + W.owner[j]);
    const unsigned ci = W.chain[entRoot(e)];
    const unsigned p = entDist(e) + W.hop[j], lines = (ci >> 16) & 0x7fffu;
    const unsigned first = W.sortedLen[ci & 0xffffu];
    const bool reversed = (ci & 0x80000000u) != 0u;
    emitWaypoint<X>(W, rec, o, first + (reversed ? lines - p : p), 2 * j, ps, r0, q0);
    emitWaypoint<X>(W, rec, o, first + (reversed ? lines - p - 1u : p + 1u), 2 * j + 1, pt, r1, q1);
  }
}

}  // namespace link2
}  // namespace nb2
