// Mesh files decoded on the device.
//
// The reference loads its meshes with pcl::io::loadPolygonFile (noether_gui/src/widgets/tpp_widget.cpp:426;
// noether_tpp/test/tool_path_planner_utest.cpp:24): vtkPLYReader -> pcl::io::vtk2mesh -> PCLPointCloud2 + one heap
// vector per face, on one core, and only then does the planner see the data.  Here the file image goes to the GPU as it
// is (one PCIe copy) and two streaming kernels turn the records into what the slicer reads: a 32-byte aligned cloud
// record {x y z 0 nx ny nz 0} per vertex and uint32 x 3 per face.  Records in a PLY body sit at arbitrary byte
// offsets (the header has no padding, face records are 13 bytes), so every field is assembled from aligned 32-bit words
// with a funnel shift; consecutive threads read consecutive records, so the loads of a warp cover one contiguous span.
#include "nb2_internal.h"

namespace nb2
{
namespace
{
constexpr unsigned kFull = 0xffffffffu;

/** Little-endian 32-bit word at an arbitrary byte offset of a word-aligned image (the image is padded by 8 bytes) */
__device__ __forceinline__ unsigned loadU32(const unsigned* __restrict__ img, unsigned long long byte_off)
{
  const unsigned long long w = byte_off >> 2;
  const unsigned sh = static_cast<unsigned>(byte_off & 3u) * 8u;
  const unsigned lo = __ldg(img + w);
  if (sh == 0)
    return lo;
  const unsigned hi = __ldg(img + w + 1);
  return __funnelshift_r(lo, hi, sh);
}

/** Unsigned little-endian field of 1, 2 or 4 bytes */
__device__ __forceinline__ unsigned loadField(const unsigned* __restrict__ img, unsigned long long byte_off, unsigned bytes)
{
  const unsigned v = loadU32(img, byte_off);
  return bytes == 4 ? v : (bytes == 2 ? (v & 0xffffu) : (v & 0xffu));
}

/** float32 property, or float64 narrowed as the reader's `(float)value` would */
__device__ __forceinline__ float loadReal(const unsigned* __restrict__ img, unsigned long long byte_off, bool f64)
{
  const unsigned lo = loadU32(img, byte_off);
  if (!f64)
    return __uint_as_float(lo);
  const unsigned hi = loadU32(img, byte_off + 4);
  return __double2float_rn(__hiloint2double(static_cast<int>(hi), static_cast<int>(lo)));
}

struct PlyVertexFields
{
  unsigned off[6];
  unsigned char f64[6];
};

/** One thread per vertex: unaligned x y z [nx ny nz] -> aligned float4 record(s) */
__global__ void __launch_bounds__(256)
    k_ply_vertices(const unsigned* __restrict__ img, unsigned long long begin, unsigned step, PlyVertexFields fld, int normals,
                   unsigned long long V, float4* __restrict__ rec)
{
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long v = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V; v += stride)
  {
    const unsigned long long at = begin + v * step;
    const float x = loadReal(img, at + fld.off[0], fld.f64[0]);
    const float y = loadReal(img, at + fld.off[1], fld.f64[1]);
    const float z = loadReal(img, at + fld.off[2], fld.f64[2]);
    if (!normals)
      rec[v] = make_float4(x, y, z, 0.f);
    else
    {
      const float nx = loadReal(img, at + fld.off[3], fld.f64[3]);
      const float ny = loadReal(img, at + fld.off[4], fld.f64[4]);
      const float nz = loadReal(img, at + fld.off[5], fld.f64[5]);
      rec[2 * v] = make_float4(x, y, z, 0.f);
      rec[2 * v + 1] = make_float4(nx, ny, nz, 0.f);
    }
  }
}

/** One thread per face: {[scalars] count i0 i1 i2 [scalars]} -> uint32 x 3; faces whose count is not 3 are counted (the
 *  fixed record size assumed for every later face is then wrong, so the upload is refused); the largest vertex index is
 *  folded for the range check (same slot as k_max_index). */
__global__ void __launch_bounds__(256)
    k_ply_faces(const unsigned* __restrict__ img, unsigned long long begin, unsigned step, unsigned pre, unsigned count_bytes,
                unsigned index_bytes, unsigned long long F, uint32_t* __restrict__ tri, unsigned* __restrict__ not_triangles,
                unsigned* __restrict__ max_index)
{
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  unsigned m = 0, bad = 0;
  for (unsigned long long f = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; f < F; f += stride)
  {
    const unsigned long long at = begin + f * step + pre;
    const unsigned n = loadField(img, at, count_bytes);
    const unsigned a = loadField(img, at + count_bytes, index_bytes);
    const unsigned b = loadField(img, at + count_bytes + index_bytes, index_bytes);
    const unsigned c = loadField(img, at + count_bytes + 2 * index_bytes, index_bytes);
    bad += n != 3u;
    tri[3 * f] = a;
    tri[3 * f + 1] = b;
    tri[3 * f + 2] = c;
    m = max(m, max(a, max(b, c)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    m = max(m, __shfl_down_sync(kFull, m, o));
    bad += __shfl_down_sync(kFull, bad, o);
  }
  if ((threadIdx.x & 31) == 0)
  {
    if (m)
      atomicMax(max_index, m);
    if (bad)
      atomicAdd(not_triangles, bad);
  }
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

void launchPlyDecode(nb2_context* c, const PlyLayout& L)
{
  const unsigned* img = c->fileImg.as<unsigned>();
  const size_t rec = L.has_normals ? 32 : 16;
  PlyVertexFields fld;
  for (int k = 0; k < 6; ++k)
  {
    fld.off[k] = L.off_field[k];
    fld.f64[k] = L.field_f64[k];
  }
  c->blob.ensure(rec * L.n_points);
  c->tri.ensure(sizeof(uint32_t) * 3 * std::max<uint64_t>(L.n_faces, 1));
  k_ply_vertices<<<gridFor(L.n_points, 256, c->sm_count * 16), 256, 0, c->stream>>>(
      img, L.vertex_begin, L.point_step, fld, L.has_normals ? 1 : 0, L.n_points, c->blob.as<float4>());
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
  NB2_CUDA(cudaMemsetAsync(c->counters.as<unsigned>() + 12, 0, sizeof(unsigned), c->stream));
  NB2_CUDA(cudaMemsetAsync(c->counters.as<unsigned>() + 16, 0, sizeof(unsigned), c->stream));
  if (L.n_faces)
  {
    k_ply_faces<<<gridFor(L.n_faces, 256, c->sm_count * 16), 256, 0, c->stream>>>(
        img, L.face_begin, L.face_step, L.face_pre, L.face_count_bytes, L.face_index_bytes, L.n_faces, c->tri.as<uint32_t>(),
        c->counters.as<unsigned>() + 16, c->counters.as<unsigned>() + 12);
    NB2_CUDA(cudaGetLastError());
    ++c->launches;
  }
}

}  // namespace nb2
