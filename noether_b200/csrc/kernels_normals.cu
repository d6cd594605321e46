// NormalsFromMeshFacesMeshModifier on the device (sm_100a).
//
// The reference (normals_from_mesh_faces_modifier.cpp:21-29, getFaceNormal utils.cpp:135-162) walks the faces in order
// and does a float `+=` of the unit face normal onto each of the face's vertices, then normalises per vertex.  The sum
// at a vertex is therefore the float sum of its incident face normals in ascending face index.  To reproduce that
// bit for bit without serialising, the faces incident to each vertex are gathered through a CSR vertex->face list
// and summed in ascending face order by the one thread that owns the vertex: no float atomics, order-deterministic.
//
//   N1  k_face_normals   : unit face normal per triangle (float, reference operation order) + vertex valence count
//                          (+ range check of the vertex indices: offending faces are skipped and reported)
//   N2  k_scan_u32       : single-pass decoupled look-back exclusive scan of the valences -> CSR row offsets
//   N3  k_csr_fill       : scatter face ids into the rows (integer atomics on the row cursors only)
//   N4  k_vertex_normals : per vertex: sort its (short) row, sum face normals in that order, normalise
#include "nb2_internal.h"

namespace nb2
{
namespace
{
constexpr unsigned kFull = 0xffffffffu;

__global__ void __launch_bounds__(256)
    k_face_normals(const uint32_t* __restrict__ tri, unsigned long long F, unsigned long long V,
                   const float4* __restrict__ pos, float4* __restrict__ faceNrm, unsigned* __restrict__ valence,
                   uint16_t* __restrict__ rank, unsigned* __restrict__ rankOverflow, unsigned* __restrict__ maxIndex)
{
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long f = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; f < F; f += stride)
  {
    const uint32_t a = __ldg(tri + 3 * f), b = __ldg(tri + 3 * f + 1), c = __ldg(tri + 3 * f + 2);
    // range check of the vertex indices (the reference would read out of bounds; here the call fails instead): the
    // face is left out of every later pass and the host reports the largest offending index
    if (max(a, max(b, c)) >= V)
    {
      atomicMax(maxIndex, max(a, max(b, c)));
      continue;
    }
    const float4 p0 = __ldg(pos + a), p1 = __ldg(pos + b), pn = __ldg(pos + c);
    // edge_01.cross(edge_0n).normalized(), float, no contraction
    const float ux = __fsub_rn(p1.x, p0.x), uy = __fsub_rn(p1.y, p0.y), uz = __fsub_rn(p1.z, p0.z);
    const float vx = __fsub_rn(pn.x, p0.x), vy = __fsub_rn(pn.y, p0.y), vz = __fsub_rn(pn.z, p0.z);
    float nx = __fsub_rn(__fmul_rn(uy, vz), __fmul_rn(uz, vy));
    float ny = __fsub_rn(__fmul_rn(uz, vx), __fmul_rn(ux, vz));
    float nz = __fsub_rn(__fmul_rn(ux, vy), __fmul_rn(uy, vx));
    // Eigen 3.4 normalized(): |n|^2 = x^2 + (y^2 + z^2); divide by sqrt when > 0, else leave (zero stays zero)
    const float z = __fadd_rn(__fmul_rn(nx, nx), __fadd_rn(__fmul_rn(ny, ny), __fmul_rn(nz, nz)));
    if (z > 0.f)
    {
      const float s = __fsqrt_rn(z);
      nx = __fdiv_rn(nx, s);
      ny = __fdiv_rn(ny, s);
      nz = __fdiv_rn(nz, s);
    }
    faceNrm[f] = make_float4(nx, ny, nz, 0.f);
    // the count doubles as the corner's slot inside its vertex's CSR row: the fill pass then needs no second round of
    // atomics (the order inside a row is arbitrary; k_vertex_normals sorts each row by face id)
    const unsigned ra = atomicAdd(valence + a, 1u), rb = atomicAdd(valence + b, 1u), rc = atomicAdd(valence + c, 1u);
    if (max(ra, max(rb, rc)) >= 0xffffu)
      *rankOverflow = 1u;  // a vertex with 65535 or more incident faces: the fill pass falls back to its own cursors
    rank[3 * f] = static_cast<uint16_t>(ra);
    rank[3 * f + 1] = static_cast<uint16_t>(rb);
    rank[3 * f + 2] = static_cast<uint16_t>(rc);
  }
}

// ---- single-pass exclusive scan (decoupled look-back, Merrill & Garland) -------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;
constexpr unsigned long long kScanAggregate = 1ull << 62, kScanPrefix = 2ull << 62, kScanMask = (1ull << 62) - 1;

__global__ void __launch_bounds__(kScanThreads)
    k_scan_u32(const unsigned* __restrict__ in, unsigned* __restrict__ out, unsigned long long n,
               unsigned long long* __restrict__ state, unsigned* __restrict__ ticket)
{
  __shared__ unsigned s_warp[kScanThreads / 32];
  __shared__ unsigned s_tile;
  __shared__ unsigned long long s_prefix;
  if (threadIdx.x == 0)
    s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const unsigned tile = s_tile;
  const unsigned long long base = static_cast<unsigned long long>(tile) * kScanTile + threadIdx.x * kScanItems;

  unsigned v[kScanItems];
  unsigned local = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
  {
    v[i] = (base + i < n) ? in[base + i] : 0u;
    local += v[i];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned x = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const unsigned y = __shfl_up_sync(kFull, x, o);
    if (lane >= o)
      x += y;
  }
  if (lane == 31)
    s_warp[warp] = x;
  __syncthreads();
  unsigned warp_excl = 0, tile_total = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w)
  {
    const unsigned t = s_warp[w];
    if (w < warp)
      warp_excl += t;
    tile_total += t;
  }
  if (warp == 0)
  {
    // look-back by one warp, 32 predecessor tiles per step: lane l reads tile j - l; the nearest tile that already
    // knows its inclusive prefix ends the walk
    volatile unsigned long long* st_v = state;
    if (lane == 0 && tile > 0)
      st_v[tile] = kScanAggregate | tile_total;
    unsigned long long excl = 0;
    for (long long j = static_cast<long long>(tile) - 1; j >= 0; j -= 32)
    {
      const long long idx = j - lane;
      unsigned long long st = kScanPrefix;  // before tile 0: an empty prefix
      if (idx >= 0)
        while (((st = st_v[idx]) >> 62) == 0ull)
        {
        }
      const unsigned closed = __ballot_sync(kFull, (st >> 62) == 2ull);
      const int stop = closed ? __ffs(closed) - 1 : 31;
      unsigned long long part = lane <= stop ? (st & kScanMask) : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        part += __shfl_xor_sync(kFull, part, o);
      excl += part;
      if (closed)
        break;
    }
    if (lane == 0)
    {
      st_v[tile] = kScanPrefix | (excl + tile_total);
      s_prefix = excl;
    }
  }
  __syncthreads();
  unsigned run = static_cast<unsigned>(s_prefix) + warp_excl + (x - local);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
  {
    if (base + i < n)
      out[base + i] = run;
    run += v[i];
  }
}

__global__ void __launch_bounds__(256)
    k_csr_fill(const uint32_t* __restrict__ tri, unsigned long long F, unsigned long long V,
               const unsigned* __restrict__ rowOff, unsigned* __restrict__ cursor, const uint16_t* __restrict__ rank,
               const unsigned* __restrict__ rankOverflow, uint32_t* __restrict__ adj)
{
  const bool use_rank = *rankOverflow == 0u;
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long f = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; f < F; f += stride)
  {
    const uint32_t id[3] = { __ldg(tri + 3 * f), __ldg(tri + 3 * f + 1), __ldg(tri + 3 * f + 2) };
    if (max(id[0], max(id[1], id[2])) >= V)
      continue;  // k_face_normals reported it
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
      const uint32_t v = id[k];
      const unsigned slot = __ldg(rowOff + v) + (use_rank ? static_cast<unsigned>(rank[3 * f + k]) : atomicAdd(cursor + v, 1u));
      adj[slot] = static_cast<uint32_t>(f);
    }
  }
}

constexpr int kSortCap = 16;

__global__ void __launch_bounds__(256)
    k_vertex_normals(unsigned long long V, const unsigned* __restrict__ rowOff, const uint32_t* __restrict__ adj,
                     const float4* __restrict__ faceNrm, float4* __restrict__ nrm)
{
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long v = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V; v += stride)
  {
    const unsigned b = rowOff[v], e = rowOff[v + 1];
    const unsigned d = e - b;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    if (d <= kSortCap)
    {
      uint32_t ids[kSortCap];
      for (unsigned i = 0; i < d; ++i)
      {
        // insertion sort while loading
        const uint32_t f = adj[b + i];
        unsigned j = i;
        while (j > 0 && ids[j - 1] > f)
        {
          ids[j] = ids[j - 1];
          --j;
        }
        ids[j] = f;
      }
      for (unsigned i = 0; i < d; ++i)
      {
        const float4 n = __ldg(faceNrm + ids[i]);
        sx = __fadd_rn(sx, n.x);
        sy = __fadd_rn(sy, n.y);
        sz = __fadd_rn(sz, n.z);
      }
    }
    else
    {
      // high-valence vertex: repeated selection of the next larger face id (a face lists a vertex at most 3 times;
      // equal ids are consumed together in the count)
      long long prev = -1;
      unsigned done = 0;
      while (done < d)
      {
        uint32_t best = 0xffffffffu;
        unsigned mult = 0;
        for (unsigned i = 0; i < d; ++i)
        {
          const uint32_t f = adj[b + i];
          if (static_cast<long long>(f) > prev)
          {
            if (f < best)
            {
              best = f;
              mult = 1;
            }
            else if (f == best)
              ++mult;
          }
        }
        const float4 n = __ldg(faceNrm + best);
        for (unsigned m = 0; m < mult; ++m)
        {
          sx = __fadd_rn(sx, n.x);
          sy = __fadd_rn(sy, n.y);
          sz = __fadd_rn(sz, n.z);
        }
        done += mult;
        prev = best;
      }
    }
    const float z = __fadd_rn(__fmul_rn(sx, sx), __fadd_rn(__fmul_rn(sy, sy), __fmul_rn(sz, sz)));
    if (z > 0.f)
    {
      const float s = __fsqrt_rn(z);
      sx = __fdiv_rn(sx, s);
      sy = __fdiv_rn(sy, s);
      sz = __fdiv_rn(sz, s);
    }
    nrm[v] = make_float4(sx, sy, sz, 0.f);
  }
}

__global__ void __launch_bounds__(256) k_max_index(const uint32_t* __restrict__ tri, unsigned long long n3, unsigned* out)
{
  unsigned m = 0;
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n3; i += stride)
    m = max(m, __ldg(tri + i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    m = max(m, __shfl_down_sync(kFull, m, o));
  if ((threadIdx.x & 31) == 0 && m)
    atomicMax(out, m);
}

inline int gridFor(unsigned long long n, int threads, int cap)
{
  unsigned long long b = (n + threads - 1) / threads;
  if (b < 1)
    b = 1;
  if (b > static_cast<unsigned long long>(cap))
    b = cap;
  return static_cast<int>(b);
}
}  // namespace

void launchValidateTriangles(nb2_context* c)
{
  NB2_CUDA(cudaMemsetAsync(c->counters.as<unsigned>() + 12, 0, sizeof(unsigned), c->stream));
  if (c->F == 0)
    return;
  k_max_index<<<gridFor(3 * c->F, 256, c->sm_count * 8), 256, 0, c->stream>>>(c->tri.as<uint32_t>(), 3 * c->F,
                                                                              c->counters.as<unsigned>() + 12);
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

/** exclusive prefix sum of n 32-bit counters (out may not alias in); out[n - 1] of an input with a trailing 0 = total */
void launchScanU32(nb2_context* c, const unsigned* in, unsigned* out, unsigned long long n)
{
  const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
  c->scanState.ensure(sizeof(unsigned long long) * (tiles + 1));
  NB2_CUDA(cudaMemsetAsync(c->scanState.ptr, 0, sizeof(unsigned long long) * (tiles + 1), c->stream));
  NB2_CUDA(cudaMemsetAsync(c->counters.as<unsigned>() + 13, 0, sizeof(unsigned), c->stream));
  k_scan_u32<<<static_cast<unsigned>(tiles), kScanThreads, 0, c->stream>>>(in, out, n, c->scanState.as<unsigned long long>(),
                                                                           c->counters.as<unsigned>() + 13);
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

void launchVertexNormals(nb2_context* c)
{
  const uint64_t V = c->V, F = c->F;
  c->faceNrm.ensure(sizeof(float4) * std::max<uint64_t>(F, 1));
  c->valence.ensure(sizeof(unsigned) * (V + 1));
  c->csrOff.ensure(sizeof(unsigned) * (V + 1));
  c->csrCursor.ensure(sizeof(unsigned) * (V + 1));
  c->csrAdj.ensure(sizeof(uint32_t) * 3 * std::max<uint64_t>(F, 1));
  c->csrRank.ensure(sizeof(uint16_t) * 3 * std::max<uint64_t>(F, 1));
  c->nrm.ensure(sizeof(float4) * V);
  c->nrm_base = c->nrm.as<uint8_t>();  // the computed normals replace whatever the cloud carried
  c->nrm_stride = 16;
  const uint64_t tiles = (V + 1 + kScanTile - 1) / kScanTile;
  c->scanState.ensure(sizeof(unsigned long long) * (tiles + 1));

  NB2_CUDA(cudaMemsetAsync(c->valence.ptr, 0, sizeof(unsigned) * (V + 1), c->stream));
  NB2_CUDA(cudaMemsetAsync(c->csrCursor.ptr, 0, sizeof(unsigned) * (V + 1), c->stream));
  NB2_CUDA(cudaMemsetAsync(c->scanState.ptr, 0, sizeof(unsigned long long) * (tiles + 1), c->stream));
  // counters[12] largest out-of-range vertex index, [13] scan ticket, [15] rank overflow flag
  NB2_CUDA(cudaMemsetAsync(c->counters.as<unsigned>() + 12, 0, 4 * sizeof(unsigned), c->stream));

  if (F)
  {
    k_face_normals<<<gridFor(F, 256, c->sm_count * 16), 256, 0, c->stream>>>(
        c->tri.as<uint32_t>(), F, V, c->pos.as<float4>(), c->faceNrm.as<float4>(), c->valence.as<unsigned>(),
        c->csrRank.as<uint16_t>(), c->counters.as<unsigned>() + 15, c->counters.as<unsigned>() + 12);
    NB2_CUDA(cudaGetLastError());
    ++c->launches;
  }
  c->timer.mark("face normals + valence", c->stream);
  k_scan_u32<<<static_cast<unsigned>(tiles), kScanThreads, 0, c->stream>>>(c->valence.as<unsigned>(), c->csrOff.as<unsigned>(), V + 1,
                                                                           c->scanState.as<unsigned long long>(),
                                                                           c->counters.as<unsigned>() + 13);
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
  c->timer.mark("scan", c->stream);
  if (F)
  {
    k_csr_fill<<<gridFor(F, 256, c->sm_count * 16), 256, 0, c->stream>>>(c->tri.as<uint32_t>(), F, V, c->csrOff.as<unsigned>(),
                                                                         c->csrCursor.as<unsigned>(), c->csrRank.as<uint16_t>(),
                                                                         c->counters.as<unsigned>() + 15, c->csrAdj.as<uint32_t>());
    NB2_CUDA(cudaGetLastError());
    ++c->launches;
  }
  c->timer.mark("csr fill", c->stream);
  k_vertex_normals<<<gridFor(V, 256, c->sm_count * 16), 256, 0, c->stream>>>(V, c->csrOff.as<unsigned>(), c->csrAdj.as<uint32_t>(),
                                                                             c->faceNrm.as<float4>(), c->nrm.as<float4>());
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
  c->timer.mark("vertex gather", c->stream);
}

}  // namespace nb2
