// noether_b200 plugin library: the B200-native raster-slicing path exported through the reference's own
// boost_plugin_loader factories, sections and YAML keys.
//
// This is the only translation unit of the product that sees noether / PCL / Eigen / yaml-cpp / boost_plugin_loader
// types.  Everything below the C ABI (include/noether_b200.h) is hand-written sm_100a CUDA in libnoether_b200.so.
//
//   alias                 section  replaces (reference, noether_tpp)
//   PlaneSlicer           tpp      Plugin_PlaneSlicerRasterPlanner        src/plugins.cpp:252-278
//   PlaneSlicerLegacy     tpp      Plugin_PlaneSlicerLegacyRasterPlanner  src/plugins.cpp:151-179
//   PrincipalAxis         dg       PrincipalAxisDirectionGenerator        src/plugins.cpp:89
//   PCARotated            dg       Plugin_PCARotatedDirectionGenerator    src/plugins.cpp:91-107
//   Centroid              og       CentroidOriginGenerator                src/plugins.cpp:111
//   AABBCenter            og       AABBCenterOriginGenerator              src/plugins.cpp:110
//   NormalsFromMeshFaces  mesh     NormalsFromMeshFacesMeshModifier       src/plugins.cpp:62
//   FaceMidpointSubdivision mesh   FaceMidpointSubdivisionMeshModifier    src/plugins.cpp:58
//   FaceSubdivisionByArea mesh     FaceSubdivisionByAreaMeshModifier      src/plugins.cpp:59
//   EuclideanClustering   mesh     EuclideanClusteringMeshModifier        src/plugins.cpp:57
//
// Built against the real headers inside a noether workspace (plugin/CMakeLists.txt).  In the development image those
// packages do not exist; there the file is compiled against plugin/compat/ (minimal stand-ins declaring the same
// interfaces) so that the shim and the C ABI are exercised by tests/cpp/plugin_shim_test.cpp.
#include <noether_b200.h>

#include <noether_tpp/core/mesh_modifier.h>
#include <noether_tpp/core/tool_path_planner.h>
#include <noether_tpp/core/types.h>
#include <noether_tpp/plugin_interface.h>
#include <noether_tpp/serialization.h>
#include <noether_tpp/tool_path_planners/raster/raster_planner.h>

#include <pcl/PCLPointCloud2.h>
#include <pcl/PCLPointField.h>
#include <pcl/PolygonMesh.h>
#include <pcl/common/io.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace noether_b200
{
// ---------------------------------------------------------------------------------------------------------------------
// device context shared by the components of this library
// ---------------------------------------------------------------------------------------------------------------------
/** One nb2_context per process and device.  The reference's component methods are all `const` and may be called from
 *  any thread (plugin_interface.h:39-40, tool_path_planner.h:37): calls are serialised here so that "upload, then
 *  compute on the resident mesh" stays atomic per call. */
class Device
{
public:
  static Device& instance()
  {
    static Device dev;
    return dev;
  }
  std::mutex& mutex() { return mutex_; }
  nb2_context* ctx()
  {
    if (!ctx_)
    {
      int ordinal = 0;
      if (const char* env = std::getenv("NOETHER_B200_DEVICE"))
        ordinal = std::atoi(env);
      check(nb2_context_create(ordinal, &ctx_));
    }
    return ctx_;
  }
  /** status -> exception, the way the reference reports failures (std::runtime_error, caught and re-thrown nested by
   *  ToolPathPlannerPipeline::plan, tool_path_planner_pipeline.cpp:64-101) */
  static void check(int status)
  {
    if (status != NB2_OK)
      throw std::runtime_error(nb2_last_error());
  }

  /** What the context holds, by content.  A noether pipeline hands the same mesh to several components in a row (the
   *  normals modifier's output to the planner; the planner's mesh to its generators; a cross-hatch planner's mesh to
   *  two plane slicers): a component whose mesh is already resident skips the upload. */
  struct Resident
  {
    bool valid = false;
    std::uint64_t n_points = 0, n_triangles = 0;
    std::uint64_t h_xyz = 0, h_tri = 0, h_normals = 0;
    bool has_normals = false;
  };
  Resident resident;
  std::vector<std::uint32_t> triangles;  // flattened polygons of the last mesh seen (kept: no page faults next time)

private:
  Device() = default;
  ~Device()
  {
    if (ctx_)
      nb2_context_destroy(ctx_);
  }
  std::mutex mutex_;
  nb2_context* ctx_ = nullptr;
};

// ---------------------------------------------------------------------------------------------------------------------
// pcl::PolygonMesh -> nb2_mesh_view (replaces the by-name accessors of utils.cpp:49-117 and the per-plan copies of
// plane_slicer_raster_planner.cpp:536-537,589-590)
// ---------------------------------------------------------------------------------------------------------------------
namespace
{
const pcl::PCLPointField* findField(const pcl::PCLPointCloud2& cloud, const char* name)
{
  for (const auto& f : cloud.fields)
    if (f.name == name)
      return &f;
  return nullptr;
}

const pcl::PCLPointField& findFieldOrThrow(const pcl::PCLPointCloud2& cloud, const char* name)
{
  const pcl::PCLPointField* f = findField(cloud, name);
  if (!f)
    throw std::runtime_error(std::string("Failed to find field '") + name + "'");  // utils.cpp:59-60
  return *f;
}

bool hasNormals(const pcl::PCLPointCloud2& cloud)  // utils.cpp:65-72
{
  return findField(cloud, "normal_x") && findField(cloud, "normal_y") && findField(cloud, "normal_z");
}

/** NOETHER_B200_TRACE=1: wall-clock laps of the host-side steps on stderr */
class Trace
{
public:
  explicit Trace(const char* what) : what_(what), on_(std::getenv("NOETHER_B200_TRACE") != nullptr), t_(now()) {}
  void lap(const char* step)
  {
    if (!on_)
      return;
    const double t = now();
    std::fprintf(stderr, "[noether_b200] %s: %s %.1f ms\n", what_, step, (t - t_) * 1e3);
    t_ = t;
  }

private:
  static double now()
  {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }
  const char* what_;
  bool on_;
  double t_;
};

// 64-bit content hash: fixed-size chunks hashed independently (so the work spreads over the cores whatever their
// number) and folded in chunk order.  Not cryptographic; it only has to tell apart the meshes of one pipeline run.
constexpr std::size_t kHashChunk = 65536;
static_assert(kHashChunk == NB2_UPLOAD_GRAIN, "the streamed upload hands out ranges that begin on a fingerprint chunk");
inline std::uint64_t mix(std::uint64_t h, std::uint64_t w)
{
  h = (h ^ w) * 0x9E3779B97F4A7C15ull;
  return h ^ (h >> 29);
}
inline std::uint64_t fold(const std::vector<std::uint64_t>& chunk_hashes, std::uint64_t seed)
{
  std::uint64_t h = seed;
  for (std::uint64_t c : chunk_hashes)
    h = mix(h, c);
  return h;
}

/** fn(chunk index, begin, end) over [0, n) in chunks of kHashChunk, spread over the hardware threads */
template <typename Fn>
void forChunks(std::size_t n, Fn fn)
{
  const std::size_t chunks = (n + kHashChunk - 1) / kHashChunk;
  const std::size_t workers = std::min<std::size_t>(std::max(1u, std::thread::hardware_concurrency()), chunks);
  std::atomic<std::size_t> next{ 0 };
  auto run = [&]() {
    for (std::size_t c = next.fetch_add(1); c < chunks; c = next.fetch_add(1))
      fn(c, c * kHashChunk, std::min(n, (c + 1) * kHashChunk));
  };
  std::vector<std::thread> pool;
  for (std::size_t w = 1; w < workers; ++w)
    pool.emplace_back(run);
  run();
  for (auto& t : pool)
    t.join();
}

/** hash of n records of three 32-bit words found at base + i * stride */
std::uint64_t hashTriples(const std::uint8_t* base, std::size_t stride, std::size_t n, std::uint64_t seed)
{
  std::vector<std::uint64_t> part((n + kHashChunk - 1) / kHashChunk);
  forChunks(n, [&](std::size_t c, std::size_t b, std::size_t e) {
    std::uint64_t h = seed + c;
    for (std::size_t i = b; i < e; ++i)
    {
      std::uint32_t w[3];
      std::memcpy(w, base + i * stride, 12);
      h = mix(h, (static_cast<std::uint64_t>(w[0]) << 32) | w[1]);
      h = mix(h, w[2]);
    }
    part[c] = h;
  });
  return fold(part, seed);
}

/** hashTriples of the positions and (off_nrm >= 0) of the normals of an interleaved cloud, in one pass over it */
void hashCloud(const std::uint8_t* data, std::size_t stride, std::size_t n, std::size_t off_xyz, long off_nrm,
               std::uint64_t& h_xyz, std::uint64_t& h_nrm)
{
  constexpr std::uint64_t kSeedXyz = 0x78797a21ull, kSeedNrm = 0x6e726d21ull;
  std::vector<std::uint64_t> px((n + kHashChunk - 1) / kHashChunk), pn(off_nrm >= 0 ? px.size() : 0);
  forChunks(n, [&](std::size_t c, std::size_t b, std::size_t e) {
    std::uint64_t hx = kSeedXyz + c, hn = kSeedNrm + c;
    for (std::size_t i = b; i < e; ++i)
    {
      std::uint32_t w[3];
      std::memcpy(w, data + i * stride + off_xyz, 12);
      hx = mix(hx, (static_cast<std::uint64_t>(w[0]) << 32) | w[1]);
      hx = mix(hx, w[2]);
      if (off_nrm >= 0)
      {
        std::memcpy(w, data + i * stride + off_nrm, 12);
        hn = mix(hn, (static_cast<std::uint64_t>(w[0]) << 32) | w[1]);
        hn = mix(hn, w[2]);
      }
    }
    px[c] = hx;
    if (off_nrm >= 0)
      pn[c] = hn;
  });
  h_xyz = fold(px, kSeedXyz);
  h_nrm = off_nrm >= 0 ? fold(pn, kSeedNrm) : 0;
}

/** The mesh as the C ABI wants it (triangles flattened into the device object's scratch) and its content hashes */
struct MeshView
{
  nb2_mesh_view view{};
  Device::Resident content;
};

MeshView makeView(Device& dev, const pcl::PolygonMesh& mesh)
{
  MeshView m;
  const pcl::PCLPointCloud2& cloud = mesh.cloud;
  const auto& x = findFieldOrThrow(cloud, "x");
  const auto& y = findFieldOrThrow(cloud, "y");
  const auto& z = findFieldOrThrow(cloud, "z");
  if ((y.offset - x.offset != 4) || (z.offset - y.offset != 4) || x.datatype != pcl::PCLPointField::FLOAT32)
    throw std::runtime_error("XYZ fields are not contiguous floats");  // utils.cpp:82-83
  m.view.cloud_data = cloud.data.data();
  m.view.n_points = static_cast<std::uint64_t>(cloud.width) * cloud.height;  // utils.cpp:169
  m.view.point_step = cloud.point_step;
  m.view.offset_xyz = x.offset;
  m.view.offset_normal = -1;
  if (hasNormals(cloud))
  {
    const auto& nx = findFieldOrThrow(cloud, "normal_x");
    const auto& ny = findFieldOrThrow(cloud, "normal_y");
    const auto& nz = findFieldOrThrow(cloud, "normal_z");
    if ((ny.offset - nx.offset != 4) || (nz.offset - ny.offset != 4) || nx.datatype != pcl::PCLPointField::FLOAT32)
      throw std::runtime_error("Normal fields are not contiguous floats");  // utils.cpp:104-105
    m.view.offset_normal = static_cast<std::int32_t>(nx.offset);
  }
  if (cloud.data.size() < m.view.n_points * cloud.point_step)
    throw std::runtime_error("point cloud data is shorter than width * height * point_step");
  // pcl::Vertices is one heap block per face: flattening 10^7 of them is pointer chasing, so it is spread over the
  // cores; the same pass hashes the indices
  const std::size_t F = mesh.polygons.size();
  dev.triangles.resize(3 * F);
  std::uint32_t* tri = dev.triangles.data();
  std::atomic<int> bad{ 0 };  // 1: fewer than 3 vertices, 2: not a triangle
  std::vector<std::uint64_t> part((F + kHashChunk - 1) / kHashChunk);
  forChunks(F, [&](std::size_t c, std::size_t b, std::size_t e) {
    std::uint64_t h = 0x7472697321ull + c;
    for (std::size_t f = b; f < e; ++f)
    {
      const auto& v = mesh.polygons[f].vertices;
      if (v.size() != 3)
      {
        bad.store(v.size() < 3 ? 1 : 2);
        return;
      }
      const std::uint32_t i0 = static_cast<std::uint32_t>(v[0]), i1 = static_cast<std::uint32_t>(v[1]),
                          i2 = static_cast<std::uint32_t>(v[2]);
      tri[3 * f] = i0, tri[3 * f + 1] = i1, tri[3 * f + 2] = i2;
      h = mix(h, (static_cast<std::uint64_t>(i0) << 32) | i1);
      h = mix(h, i2);
    }
    part[c] = h;
  });
  if (bad.load() == 1)
    throw std::runtime_error("Polygon has fewer than 3 vertices");  // getFaceNormal, utils.cpp:137-142
  if (bad.load() == 2)
    throw std::runtime_error("noether_b200 slices triangle meshes only; triangulate the mesh first");
  m.view.triangles = tri;
  m.view.n_triangles = F;

  m.content.valid = true;
  m.content.n_points = m.view.n_points;
  m.content.n_triangles = F;
  m.content.h_tri = fold(part, 0x7472697321ull);
  m.content.has_normals = m.view.offset_normal >= 0;
  hashCloud(cloud.data.data(), cloud.point_step, m.view.n_points, x.offset, m.view.offset_normal, m.content.h_xyz,
            m.content.h_normals);
  return m;
}

/** The mesh handed to the device piece by piece (nb2_mesh_upload_stream): the library's staging threads call these on
 *  disjoint ranges; each call compacts / flattens its range straight into pinned memory and fingerprints it on the way, so
 *  the host data is read once in all. */
struct StreamJob
{
  const pcl::PolygonMesh* mesh = nullptr;
  std::size_t in_step = 0, off_xyz = 0;
  long off_nrm = -1;
  std::vector<std::uint64_t> hx, hn, ht;  // per kHashChunk
  std::atomic<int> bad{ 0 };              // 1: fewer than 3 vertices, 2: not a triangle
  Device* dev = nullptr;
  bool need_normals = false;
  Device::Resident content;
  bool hit = false;

  std::atomic<long long> ns_cloud{ 0 }, ns_tri{ 0 };  // NOETHER_B200_TRACE: time the staging threads spent in the callbacks
  bool timed = false;
  struct Lap
  {
    std::atomic<long long>* acc;
    std::chrono::steady_clock::time_point t0;
    Lap(std::atomic<long long>* a) : acc(a), t0(a ? std::chrono::steady_clock::now() : std::chrono::steady_clock::time_point()) {}
    ~Lap()
    {
      if (acc)
        acc->fetch_add(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count());
    }
  };

  static int cloud(void* user, std::uint64_t first, std::uint64_t count, void* dst)
  {
    StreamJob& j = *static_cast<StreamJob*>(user);
    Lap lap(j.timed ? &j.ns_cloud : nullptr);
    constexpr std::uint64_t kSeedXyz = 0x78797a21ull, kSeedNrm = 0x6e726d21ull;
    const std::uint8_t* p = j.mesh->cloud.data.data() + first * j.in_step;
    std::uint8_t* o = static_cast<std::uint8_t*>(dst);
    const bool nrm = j.off_nrm >= 0;
    for (std::uint64_t b = first; b < first + count; b += kHashChunk)
    {
      const std::uint64_t c = b / kHashChunk, e = std::min(first + count, b + kHashChunk);
      std::uint64_t hx = kSeedXyz + c, hn = kSeedNrm + c;
      for (std::uint64_t i = b; i < e; ++i, p += j.in_step)
      {
        std::uint32_t w[3];
        std::memcpy(w, p + j.off_xyz, 12);
        std::memcpy(o, w, 12);
        o += 12;
        hx = mix(hx, (static_cast<std::uint64_t>(w[0]) << 32) | w[1]);
        hx = mix(hx, w[2]);
        if (nrm)
        {
          std::memcpy(w, p + j.off_nrm, 12);
          std::memcpy(o, w, 12);
          o += 12;
          hn = mix(hn, (static_cast<std::uint64_t>(w[0]) << 32) | w[1]);
          hn = mix(hn, w[2]);
        }
      }
      j.hx[c] = hx;
      if (nrm)
        j.hn[c] = hn;
    }
    return 0;
  }
  static int triangles(void* user, std::uint64_t first, std::uint64_t count, std::uint32_t* dst)
  {
    StreamJob& j = *static_cast<StreamJob*>(user);
    Lap lap(j.timed ? &j.ns_tri : nullptr);
    const auto& polygons = j.mesh->polygons;
    for (std::uint64_t b = first; b < first + count; b += kHashChunk)
    {
      const std::uint64_t c = b / kHashChunk, e = std::min(first + count, b + kHashChunk);
      std::uint64_t h = 0x7472697321ull + c;
      for (std::uint64_t f = b; f < e; ++f, dst += 3)
      {
        const auto& v = polygons[f].vertices;
        if (v.size() != 3)
        {
          j.bad.store(v.size() < 3 ? 1 : 2);
          return 1;
        }
        const std::uint32_t i0 = static_cast<std::uint32_t>(v[0]), i1 = static_cast<std::uint32_t>(v[1]),
                            i2 = static_cast<std::uint32_t>(v[2]);
        dst[0] = i0, dst[1] = i1, dst[2] = i2;
        h = mix(h, (static_cast<std::uint64_t>(i0) << 32) | i1);
        h = mix(h, i2);
      }
      j.ht[c] = h;
    }
    return 0;
  }
  /** everything is staged: is it what the device already holds? */
  static int accept(void* user)
  {
    StreamJob& j = *static_cast<StreamJob*>(user);
    j.content.h_xyz = fold(j.hx, 0x78797a21ull);
    j.content.h_normals = j.off_nrm >= 0 ? fold(j.hn, 0x6e726d21ull) : 0;
    j.content.h_tri = fold(j.ht, 0x7472697321ull);
    const Device::Resident& r = j.dev->resident;
    const Device::Resident& m = j.content;
    const bool same_geometry = r.valid && r.n_points == m.n_points && r.n_triangles == m.n_triangles && r.h_xyz == m.h_xyz &&
                               r.h_tri == m.h_tri;
    const bool same_normals = r.has_normals == m.has_normals && (!r.has_normals || r.h_normals == m.h_normals);
    j.hit = same_geometry && (same_normals || !j.need_normals);
    return j.hit ? 0 : 1;
  }
};

/** Make `mesh` the context's resident mesh.  The mesh is streamed to the device while it is fingerprinted (one pass over
 *  the host data); when the fingerprint says the device already holds these vertices, triangles and, where the caller
 *  needs them, normals, the staged copy is dropped and the resident mesh stays, with everything derived from it (PCA,
 *  extents, computed normals).  Call with the device mutex held. */
void ensureResident(Device& dev, const pcl::PolygonMesh& mesh, bool need_normals, bool whole_records = false)
{
  Trace trace("ensureResident");
  if (whole_records)
  {
    // the caller works on every byte of the cloud records (face subdivision copies them), and the fingerprints cover
    // positions, normals and indices only — the same geometry in another record layout, or with other colours /
    // curvature, would pass them.  Such callers always upload, records as they are.
    const MeshView m = makeView(dev, mesh);
    trace.lap("flatten + fingerprint");
    dev.resident.valid = false;
    Device::check(nb2_mesh_upload(dev.ctx(), &m.view));
    dev.resident = m.content;
    trace.lap("upload");
    return;
  }
  const pcl::PCLPointCloud2& cloud = mesh.cloud;
  const auto& x = findFieldOrThrow(cloud, "x");
  const auto& y = findFieldOrThrow(cloud, "y");
  const auto& z = findFieldOrThrow(cloud, "z");
  if ((y.offset - x.offset != 4) || (z.offset - y.offset != 4) || x.datatype != pcl::PCLPointField::FLOAT32)
    throw std::runtime_error("XYZ fields are not contiguous floats");  // utils.cpp:82-83
  StreamJob job;
  job.mesh = &mesh;
  job.dev = &dev;
  job.need_normals = need_normals;
  job.in_step = cloud.point_step;
  job.off_xyz = x.offset;
  if (hasNormals(cloud))
  {
    const auto& nx = findFieldOrThrow(cloud, "normal_x");
    const auto& ny = findFieldOrThrow(cloud, "normal_y");
    const auto& nz = findFieldOrThrow(cloud, "normal_z");
    if ((ny.offset - nx.offset != 4) || (nz.offset - ny.offset != 4) || nx.datatype != pcl::PCLPointField::FLOAT32)
      throw std::runtime_error("Normal fields are not contiguous floats");  // utils.cpp:104-105
    job.off_nrm = static_cast<long>(nx.offset);
  }
  const std::uint64_t V = static_cast<std::uint64_t>(cloud.width) * cloud.height, F = mesh.polygons.size();  // utils.cpp:169
  if (cloud.data.size() < V * cloud.point_step)
    throw std::runtime_error("point cloud data is shorter than width * height * point_step");
  job.hx.assign((V + kHashChunk - 1) / kHashChunk, 0);
  job.hn.assign(job.off_nrm >= 0 ? job.hx.size() : 0, 0);
  job.ht.assign((F + kHashChunk - 1) / kHashChunk, 0);
  job.content.valid = true;
  job.content.n_points = V;
  job.content.n_triangles = F;
  job.content.has_normals = job.off_nrm >= 0;
  nb2_mesh_stream st{};
  st.n_points = V;
  st.point_step = job.off_nrm >= 0 ? 24u : 12u;  // what travels: {x y z [nx ny nz]}
  st.offset_xyz = 0;
  st.offset_normal = job.off_nrm >= 0 ? 12 : -1;
  st.n_triangles = F;
  st.cloud = &StreamJob::cloud;
  st.triangles = &StreamJob::triangles;
  st.accept = &StreamJob::accept;
  st.user = &job;
  job.timed = std::getenv("NOETHER_B200_TRACE") != nullptr;
  int committed = 0;
  const Device::Resident before = dev.resident;
  dev.resident.valid = false;
  const int rc = nb2_mesh_upload_stream(dev.ctx(), &st, &committed);
  if (job.bad.load() == 1)
    throw std::runtime_error("Polygon has fewer than 3 vertices");  // getFaceNormal, utils.cpp:137-142
  if (job.bad.load() == 2)
    throw std::runtime_error("noether_b200 slices triangle meshes only; triangulate the mesh first");
  Device::check(rc);
  dev.resident = committed ? job.content : before;
  trace.lap(committed ? "streamed to the device" : "streamed, already resident");
  if (job.timed)
    std::fprintf(stderr, "[noether_b200] ensureResident: staging threads spent %.1f ms on the cloud, %.1f ms on the triangles (summed over threads)\n",
                 job.ns_cloud.load() * 1e-6, job.ns_tri.load() * 1e-6);
}

/** SoA result -> noether::ToolPaths (core/types.h:31-50): nested vectors of Eigen::Isometry3d */
noether::ToolPaths toToolPaths(const nb2_result* r)
{
  nb2_result_view v{};
  Device::check(nb2_result_layout(r, &v));  // (positions / normals may still be arriving: download = 2)
  // Eigen::Isometry3d is exactly its 4x4 double matrix, column-major, last row 0 0 0 1 — the layout the library writes —
  // so every segment's storage is handed over and filled in place (in parallel, no intermediate copy)
  static_assert(sizeof(Eigen::Isometry3d) == 16 * sizeof(double), "Isometry3d is not a bare 4x4 double matrix");
  noether::ToolPaths out(v.n_paths);
  std::vector<double*> segment_out(v.n_segments);
  for (std::uint64_t p = 0; p < v.n_paths; ++p)
  {
    noether::ToolPath& path = out[p];
    path.resize(v.path_offsets[p + 1] - v.path_offsets[p]);
    for (std::uint32_t s = v.path_offsets[p]; s < v.path_offsets[p + 1]; ++s)
    {
      noether::ToolPathSegment& seg = path[s - v.path_offsets[p]];
      seg.resize(v.segment_offsets[s + 1] - v.segment_offsets[s]);
      segment_out[s] = seg.empty() ? nullptr : seg.front().matrix().data();
    }
  }
  Device::check(nb2_result_poses_segments(r, segment_out.data()));
  return out;
}
}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// generators
// ---------------------------------------------------------------------------------------------------------------------
/** PrincipalAxisDirectionGenerator (principal_axis_direction_generator.cpp:14-26): the PCA runs on the GPU as part of
 *  the ingest pass. */
class PrincipalAxisDirectionGenerator : public noether::DirectionGenerator
{
public:
  explicit PrincipalAxisDirectionGenerator(double rotation_offset = 0.0) : rotation_offset_(rotation_offset) {}
  double rotationOffset() const { return rotation_offset_; }

  Eigen::Vector3d generate(const pcl::PolygonMesh& mesh) const override final
  {
    Device& dev = Device::instance();
    std::lock_guard<std::mutex> lock(dev.mutex());
    ensureResident(dev, mesh, false);
    double d[3];
    Device::check(nb2_principal_axis_direction(dev.ctx(), rotation_offset_, d));
    return Eigen::Vector3d(d[0], d[1], d[2]);
  }

private:
  double rotation_offset_;
};

/** PCARotatedDirectionGenerator (pca_rotated_direction_generator.cpp:13-25): the direction of a nested generator rotated
 *  about the smallest principal axis; the PCA runs on the GPU. */
class PCARotatedDirectionGenerator : public noether::DirectionGenerator
{
public:
  PCARotatedDirectionGenerator(noether::DirectionGenerator::ConstPtr dir_gen, double rotation_angle)
    : dir_gen_(std::move(dir_gen)), rotation_angle_(rotation_angle)
  {
  }

  Eigen::Vector3d generate(const pcl::PolygonMesh& mesh) const override final
  {
    const Eigen::Vector3d nominal = dir_gen_->generate(mesh);  // may use the device itself: call it before locking
    const double v[3] = { nominal.x(), nominal.y(), nominal.z() };
    Device& dev = Device::instance();
    std::lock_guard<std::mutex> lock(dev.mutex());
    ensureResident(dev, mesh, false);
    double d[3];
    Device::check(nb2_pca_rotated_direction(dev.ctx(), rotation_angle_, v, d));
    return Eigen::Vector3d(d[0], d[1], d[2]);
  }

private:
  noether::DirectionGenerator::ConstPtr dir_gen_;
  double rotation_angle_;
};

/** CentroidOriginGenerator (centroid_origin_generator.cpp:8-17) */
class CentroidOriginGenerator : public noether::OriginGenerator
{
public:
  Eigen::Vector3d generate(const pcl::PolygonMesh& mesh) const override final
  {
    Device& dev = Device::instance();
    std::lock_guard<std::mutex> lock(dev.mutex());
    ensureResident(dev, mesh, false);
    double o[3];
    Device::check(nb2_centroid_origin(dev.ctx(), o));
    return Eigen::Vector3d(o[0], o[1], o[2]);
  }
};

/** AABBCenterOriginGenerator (aabb_center_origin_generator.cpp:9-17): (max - min) / 2 of the float AABB, which the ingest
 *  pass leaves on the device */
class AABBCenterOriginGenerator : public noether::OriginGenerator
{
public:
  Eigen::Vector3d generate(const pcl::PolygonMesh& mesh) const override final
  {
    Device& dev = Device::instance();
    std::lock_guard<std::mutex> lock(dev.mutex());
    ensureResident(dev, mesh, false);
    double o[3];
    Device::check(nb2_aabb_center_origin(dev.ctx(), o));
    return Eigen::Vector3d(o[0], o[1], o[2]);
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// planners
// ---------------------------------------------------------------------------------------------------------------------
/**
 * PlaneSlicerRasterPlanner (plane_slicer_raster_planner.cpp:518-631).  Deriving from noether::RasterPlanner keeps the
 * reference's own `plan` (raster_planner.cpp:16-35, final): it calls this planImpl and then applies the reference's
 * RasterOrganizationModifier and DirectionOfTravelOrientationModifier, so post-processing is the reference's code.
 *
 * When the configured generators are this library's PrincipalAxis / Centroid, their values come out of the same mesh
 * upload as the plan (one H2D copy, one PCA); any other generator (FixedDirection, PCARotated, Offset, ... from the
 * reference's plugin library) is evaluated through its own `generate` and passed as an explicit vector.
 */
class PlaneSlicerRasterPlanner : public noether::RasterPlanner
{
public:
  PlaneSlicerRasterPlanner(noether::DirectionGenerator::ConstPtr dir_gen, noether::OriginGenerator::ConstPtr origin_gen)
    : noether::RasterPlanner(std::move(dir_gen), std::move(origin_gen))
  {
  }
  void generateRastersBidirectionally(const bool bidirectional) { bidirectional_ = bidirectional; }

protected:
  virtual void fillLegacy(nb2_plan_params&) const {}

  noether::ToolPaths planImpl(const pcl::PolygonMesh& mesh) const override
  {
    if (!hasNormals(mesh.cloud))  // same message as plane_slicer_raster_planner.cpp:520-528
      throw std::runtime_error("The input mesh does not have vertex normals, which are required for the plane slice "
                               "raster tool path planner. Use a MeshModifier to generate vertex normals for the mesh, "
                               "or provide a different mesh with vertex normals.");
    nb2_plan_params prm;
    nb2_plan_params_default(&prm);
    prm.line_spacing = line_spacing_;
    prm.bidirectional = bidirectional_ ? 1 : 0;
    prm.raster_post_ops = 0;  // RasterPlanner::plan applies the reference's own modifiers after planImpl
    prm.emit_edges = 0;
    prm.download = 2;         // the waypoints arrive while the ToolPaths containers are sized and filled
    fillLegacy(prm);

    // foreign generators run first: they may use the device themselves (through this library or not)
    if (const auto* pa = dynamic_cast<const PrincipalAxisDirectionGenerator*>(dir_gen_.get()))
    {
      prm.direction_mode = NB2_DIRECTION_PRINCIPAL_AXIS;
      prm.rotation_offset = pa->rotationOffset();
    }
    else
    {
      const Eigen::Vector3d d = dir_gen_->generate(mesh);
      prm.direction_mode = NB2_DIRECTION_EXPLICIT;
      prm.direction[0] = d.x(), prm.direction[1] = d.y(), prm.direction[2] = d.z();
    }
    if (dynamic_cast<const CentroidOriginGenerator*>(origin_gen_.get()))
      prm.origin_mode = NB2_ORIGIN_CENTROID;
    else
    {
      const Eigen::Vector3d o = origin_gen_->generate(mesh);
      prm.origin_mode = NB2_ORIGIN_EXPLICIT;
      prm.origin[0] = o.x(), prm.origin[1] = o.y(), prm.origin[2] = o.z();
    }

    Device& dev = Device::instance();
    std::lock_guard<std::mutex> lock(dev.mutex());
    ensureResident(dev, mesh, true);
    Trace trace("planImpl");
    nb2_result* r = nullptr;
    Device::check(nb2_plan(dev.ctx(), &prm, &r));
    trace.lap("nb2_plan");
    std::unique_ptr<nb2_result, void (*)(nb2_result*)> guard(r, nb2_result_free);
    noether::ToolPaths out = toToolPaths(r);
    trace.lap("ToolPaths");
    return out;
  }

  bool bidirectional_ = true;
};

/** PlaneSlicerLegacyRasterPlanner (plane_slicer_legacy_raster_planner.cpp:22-33): the plane slicer followed by
 *  JoinCloseSegments, MinimumSegmentLength and UniformSpacing(degree 1). */
class PlaneSlicerLegacyRasterPlanner : public PlaneSlicerRasterPlanner
{
public:
  PlaneSlicerLegacyRasterPlanner(noether::DirectionGenerator::ConstPtr dir_gen, noether::OriginGenerator::ConstPtr origin_gen,
                                 const double point_spacing, const double min_segment_size, const double min_hole_size)
    : PlaneSlicerRasterPlanner(std::move(dir_gen), std::move(origin_gen))
    , point_spacing_(point_spacing)
    , min_segment_size_(min_segment_size)
    , min_hole_size_(min_hole_size)
  {
  }

protected:
  void fillLegacy(nb2_plan_params& prm) const override
  {
    prm.legacy = 1;
    prm.point_spacing = point_spacing_;
    prm.min_segment_size = min_segment_size_;
    prm.min_hole_size = min_hole_size_;
  }

  double point_spacing_, min_segment_size_, min_hole_size_;
};

// ---------------------------------------------------------------------------------------------------------------------
// mesh modifier
// ---------------------------------------------------------------------------------------------------------------------
/** The output of the normal-producing mesh modifiers: the input's header and polygons and concatenateFields(mesh.cloud,
 *  normals) — what pcl::toPCLPointCloud2(pcl::PointCloud<pcl::Normal>) yields (32-byte points: normal_x/y/z at 0/4/8,
 *  curvature at 16, zero padding) in front of the input's fields (normals_from_mesh_faces_modifier.cpp:34-45,
 *  normal_estimation_pcl.cpp:36-47).  curvature may be NULL (zero, as NormalsFromMeshFaces leaves it). */
std::vector<pcl::PolygonMesh> withNormalRecords(const pcl::PolygonMesh& mesh, const float* normals, const float* curvature)
{
  const std::size_t n = static_cast<std::size_t>(mesh.cloud.width) * mesh.cloud.height;
  pcl::PCLPointCloud2 normals_pc2;
  normals_pc2.header = mesh.cloud.header;
  normals_pc2.width = mesh.cloud.width;
  normals_pc2.height = mesh.cloud.height;
  normals_pc2.is_bigendian = false;
  normals_pc2.is_dense = true;
  normals_pc2.point_step = 32;
  normals_pc2.row_step = 32 * mesh.cloud.width;
  const char* names[4] = { "normal_x", "normal_y", "normal_z", "curvature" };
  const std::uint32_t offsets[4] = { 0, 4, 8, 16 };
  for (int i = 0; i < 4; ++i)
  {
    pcl::PCLPointField f;
    f.name = names[i];
    f.offset = offsets[i];
    f.datatype = pcl::PCLPointField::FLOAT32;
    f.count = 1;
    normals_pc2.fields.push_back(f);
  }
  normals_pc2.data.assign(32 * n, 0);
  for (std::size_t i = 0; i < n; ++i)
  {
    std::memcpy(normals_pc2.data.data() + 32 * i, normals + 3 * i, 12);
    if (curvature)
      std::memcpy(normals_pc2.data.data() + 32 * i + 16, curvature + i, 4);
  }
  // the output mesh: the input's header and polygons (one deep copy of the polygon list: the interface returns meshes
  // by value) and the concatenated cloud.  Moved, not copied, into the result (`return { output }` would copy the
  // 10^7 polygon vectors a second time).
  std::vector<pcl::PolygonMesh> result(1);
  pcl::PolygonMesh& output = result.front();
  output.header = mesh.header;
  output.polygons = mesh.polygons;
  if (!pcl::concatenateFields(mesh.cloud, normals_pc2, output.cloud))
    throw std::runtime_error("Failed to concatenate normals into mesh vertex cloud");  // :42-43
  return result;
}

/** NormalsFromMeshFacesMeshModifier (normals_from_mesh_faces_modifier.cpp:15-46): vertex normals on the GPU, then the
 *  reference's own field concatenation (pcl::concatenateFields) so the output cloud has the layout the reference
 *  produces: a 32-byte pcl::Normal record (normal_x, normal_y, normal_z, curvature) followed by the input's fields. */
class NormalsFromMeshFacesMeshModifier : public noether::MeshModifier
{
public:
  std::vector<pcl::PolygonMesh> modify(const pcl::PolygonMesh& mesh) const override final
  {
    const std::size_t n = static_cast<std::size_t>(mesh.cloud.width) * mesh.cloud.height;
    std::vector<float> normals(3 * n);
    {
      Device& dev = Device::instance();
      std::lock_guard<std::mutex> lock(dev.mutex());
      ensureResident(dev, mesh, false);
      dev.resident.valid = false;
      Device::check(nb2_normals_from_mesh_faces(dev.ctx(), normals.data()));
      // the context now holds this geometry with the computed normals: exactly the mesh returned below, which the
      // pipeline hands to the planner next
      dev.resident.has_normals = true;
      dev.resident.h_normals = hashTriples(reinterpret_cast<const std::uint8_t*>(normals.data()), 12, n, 0x6e726d21ull);
      dev.resident.valid = true;
    }
    return withNormalRecords(mesh, normals.data(), nullptr);
  }
};

/** NormalEstimationPCLMeshModifier (normal_estimation_pcl.cpp:15-50): pcl::NormalEstimation with a radius search and a view
 *  point, one k-d tree search per point in the reference, on the GPU here (nb2_normal_estimation_pcl); the output cloud is the
 *  reference's concatenation, curvature included. */
class NormalEstimationPCLMeshModifier : public noether::MeshModifier
{
public:
  NormalEstimationPCLMeshModifier(double radius, double vx, double vy, double vz) : radius_(radius), vx_(vx), vy_(vy), vz_(vz) {}
  std::vector<pcl::PolygonMesh> modify(const pcl::PolygonMesh& mesh) const override final
  {
    const std::size_t n = static_cast<std::size_t>(mesh.cloud.width) * mesh.cloud.height;
    std::vector<float> normals(3 * n), curvature(n);
    {
      Device& dev = Device::instance();
      std::lock_guard<std::mutex> lock(dev.mutex());
      ensureResident(dev, mesh, false);
      dev.resident.valid = false;
      Device::check(nb2_normal_estimation_pcl(dev.ctx(), radius_, vx_, vy_, vz_, normals.data(), curvature.data()));
      if (radius_ > 0.0)
      {
        // the context now holds this geometry with the estimated normals: the mesh returned below
        dev.resident.has_normals = true;
        dev.resident.h_normals = hashTriples(reinterpret_cast<const std::uint8_t*>(normals.data()), 12, n, 0x6e726d21ull);
        dev.resident.valid = true;
      }
    }
    return withNormalRecords(mesh, normals.data(), curvature.data());
  }

private:
  double radius_, vx_, vy_, vz_;
};

/** Host side of EuclideanClusteringMeshModifier::modify, from the labels the device computed (label = smallest point index
 *  of the point's cluster): size filter, PCL's size sort, one sub-mesh per cluster.  No device involved. */
std::vector<pcl::PolygonMesh> cutClusterMeshes(const pcl::PolygonMesh& mesh, const std::uint32_t* labels, int min_cluster_size,
                                               int max_cluster_size)
{
  const std::size_t V = static_cast<std::size_t>(mesh.cloud.width) * mesh.cloud.height;
  // clusters in order of discovery (ascending seed) with their sizes; the size filter of extractEuclideanClusters
  std::vector<std::uint32_t> size_of(V, 0);
  for (std::size_t v = 0; v < V; ++v)
    ++size_of[labels[v]];
  const std::uint32_t min_pts = static_cast<std::uint32_t>(min_cluster_size);  // setMinClusterSize takes pcl::uindex_t
  const std::uint32_t max_pts = max_cluster_size <= 0 ? static_cast<std::uint32_t>(V) : static_cast<std::uint32_t>(max_cluster_size);  // :53
  std::vector<std::uint32_t> seeds, sizes;
  for (std::size_t v = 0; v < V; ++v)
    if (size_of[v] && size_of[v] >= min_pts && size_of[v] <= max_pts)
    {
      seeds.push_back(static_cast<std::uint32_t>(v));
      sizes.push_back(size_of[v]);
    }
  std::vector<std::uint32_t> order(seeds.size());
  Device::check(nb2_cluster_order(sizes.data(), sizes.size(), order.data()));
  // faces by cluster, in face order (a face belongs to a cluster when its first three vertices do)
  std::vector<std::int64_t> slot_of_seed(V, -1);
  for (std::size_t k = 0; k < order.size(); ++k)
    slot_of_seed[seeds[order[k]]] = static_cast<std::int64_t>(k);
  std::vector<std::vector<std::uint32_t>> faces(order.size());
  for (std::size_t f = 0; f < mesh.polygons.size(); ++f)
  {
    const auto& v = mesh.polygons[f].vertices;
    const std::uint32_t l = labels[v[0]];
    if (labels[v[1]] == l && labels[v[2]] == l && slot_of_seed[l] >= 0)
      faces[static_cast<std::size_t>(slot_of_seed[l])].push_back(static_cast<std::uint32_t>(f));
  }
  std::vector<pcl::PolygonMesh> result(order.size());
  std::vector<std::int64_t> new_index(V, -1);
  for (std::size_t k = 0; k < order.size(); ++k)
  {
    pcl::PolygonMesh& out = result[k];
    out.header = mesh.header;             // subset_extractor.cpp:52-53
    out.cloud.header = mesh.cloud.header;
    if (faces[k].empty())
      continue;  // simplify() returns at once: the output stays a default-constructed mesh
    std::vector<std::uint32_t> used;
    for (std::uint32_t f : faces[k])
      for (auto v : mesh.polygons[f].vertices)
        if (new_index[v] < 0)
        {
          new_index[v] = static_cast<std::int64_t>(used.size());
          used.push_back(static_cast<std::uint32_t>(v));
        }
    out.polygons.reserve(faces[k].size());
    if (used.size() == V)
    {
      out.cloud = mesh.cloud;  // every point is used: the cloud as it is
      for (std::uint32_t f : faces[k])
        out.polygons.push_back(mesh.polygons[f]);
    }
    else
    {
      out.cloud.fields = mesh.cloud.fields;
      out.cloud.point_step = mesh.cloud.point_step;
      out.cloud.is_bigendian = mesh.cloud.is_bigendian;
      out.cloud.height = 1;
      out.cloud.width = static_cast<std::uint32_t>(used.size());
      out.cloud.row_step = out.cloud.point_step * out.cloud.width;
      out.cloud.is_dense = false;
      out.cloud.data.resize(static_cast<std::size_t>(out.cloud.width) * out.cloud.point_step);
      for (std::size_t i = 0; i < used.size(); ++i)
        std::memcpy(out.cloud.data.data() + i * out.cloud.point_step,
                    mesh.cloud.data.data() + static_cast<std::size_t>(used[i]) * mesh.cloud.point_step, mesh.cloud.point_step);
      for (std::uint32_t f : faces[k])
      {
        pcl::Vertices p;
        p.vertices.resize(mesh.polygons[f].vertices.size());
        for (std::size_t i = 0; i < p.vertices.size(); ++i)
          p.vertices[i] = static_cast<decltype(p.vertices)::value_type>(new_index[mesh.polygons[f].vertices[i]]);
        out.polygons.push_back(std::move(p));
      }
    }
    for (std::uint32_t v : used)
      new_index[v] = -1;
  }
  return result;
}

/** EuclideanClusteringMeshModifier (euclidean_clustering_modifier.cpp:40-64): the clustering itself — one k-d tree radius
 *  search per point in the reference — runs on the GPU (nb2_euclidean_cluster_labels: connected components under FLANN's
 *  radius test, label = smallest point index = PCL's seed).  The rest is index bookkeeping on the labels, done here where
 *  the meshes are materialised: the size filter of pcl::extractEuclideanClusters, PCL's size sort (nb2_cluster_order),
 *  extractSubMeshFromInlierVertices (subset_extractor.cpp:22-56) and SimplificationRemoveUnusedVertices. */
class EuclideanClusteringMeshModifier : public noether::MeshModifier
{
public:
  EuclideanClusteringMeshModifier(double tolerance, int min_cluster_size, int max_cluster_size)
    : tolerance_(tolerance), min_cluster_size_(min_cluster_size), max_cluster_size_(max_cluster_size)
  {
  }

  std::vector<pcl::PolygonMesh> modify(const pcl::PolygonMesh& mesh) const override final
  {
    const std::size_t V = static_cast<std::size_t>(mesh.cloud.width) * mesh.cloud.height;
    std::vector<std::uint32_t> labels(V);
    {
      Device& dev = Device::instance();
      std::lock_guard<std::mutex> lock(dev.mutex());
      ensureResident(dev, mesh, false);
      Device::check(nb2_euclidean_cluster_labels(dev.ctx(), tolerance_, labels.data(), labels.size()));
    }
    return cutClusterMeshes(mesh, labels.data(), min_cluster_size_, max_cluster_size_);
  }

private:
  double tolerance_;
  int min_cluster_size_, max_cluster_size_;
};

/** The face subdivision mesh modifiers (face_subdivision_modifier.cpp:202-339) on the GPU: the mesh is subdivided where it
 *  is resident and read back once — a cloud with the input's fields (new records are copies of their first source with
 *  x y z / normals / rgba averaged) and 3-vertex polygons in the reference's depth-first order. */
class FaceSubdivisionOnDevice : public noether::MeshModifier
{
public:
  std::vector<pcl::PolygonMesh> modify(const pcl::PolygonMesh& mesh) const override final
  {
    std::int32_t offset_rgba = -1;  // has_color = the cloud has a field called "rgba" (:223,267)
    for (const pcl::PCLPointField& f : mesh.cloud.fields)
      if (f.name == "rgba")
        offset_rgba = static_cast<std::int32_t>(f.offset);
    std::vector<pcl::PolygonMesh> result(1);
    pcl::PolygonMesh& output = result.front();
    output.header = mesh.header;
    output.cloud = mesh.cloud;  // fields, point_step, flags; the data is replaced below
    std::vector<std::uint32_t> tri;
    nb2_mesh_file_info info{};
    {
      Device& dev = Device::instance();
      std::lock_guard<std::mutex> lock(dev.mutex());
      // the records themselves are subdivided: the cloud on the device must be this very cloud, byte for byte (after
      // NormalsFromMeshFaces the context holds the INPUT's records, not the concatenated cloud that modifier returned)
      ensureResident(dev, mesh, false, /*whole_records=*/true);
      dev.resident.valid = false;  // the context holds the subdivided mesh from here on
      Device::check(subdivide(dev.ctx(), offset_rgba, &info));
      std::uint32_t step = 0;
      Device::check(nb2_mesh_record_layout(dev.ctx(), &step, nullptr, nullptr));
      if (step != mesh.cloud.point_step)
        throw std::runtime_error("noether_b200: the resident cloud has " + std::to_string(step) +
                                 "-byte records but the mesh being subdivided has " + std::to_string(mesh.cloud.point_step));
      output.cloud.data.resize(info.n_points * mesh.cloud.point_step);
      tri.resize(3 * info.n_triangles);
      Device::check(nb2_mesh_download_cloud(dev.ctx(), output.cloud.data.data(), output.cloud.data.size()));
      if (info.n_triangles)
        Device::check(nb2_mesh_download_triangles(dev.ctx(), tri.data(), info.n_triangles));
    }
    // the flattened cloud of the reference (:218-220,241 / :262-264,305)
    output.cloud.height = 1;
    output.cloud.width = static_cast<std::uint32_t>(info.n_points);
    output.cloud.row_step = output.cloud.width * output.cloud.point_step;
    output.polygons.resize(info.n_triangles);
    forChunks(info.n_triangles, [&](std::size_t, std::size_t b, std::size_t e) {
      for (std::size_t f = b; f < e; ++f)
        output.polygons[f].vertices = { tri[3 * f], tri[3 * f + 1], tri[3 * f + 2] };
    });
    return result;
  }

protected:
  virtual int subdivide(nb2_context* ctx, std::int32_t offset_rgba, nb2_mesh_file_info* info) const = 0;
};

class FaceMidpointSubdivisionMeshModifier : public FaceSubdivisionOnDevice
{
public:
  explicit FaceMidpointSubdivisionMeshModifier(unsigned n_iterations = 1) : n_iterations_(n_iterations) {}

protected:
  int subdivide(nb2_context* ctx, std::int32_t offset_rgba, nb2_mesh_file_info* info) const override
  {
    return nb2_face_midpoint_subdivision(ctx, n_iterations_, offset_rgba, info);
  }
  unsigned n_iterations_;
};

class FaceSubdivisionByAreaMeshModifier : public FaceSubdivisionOnDevice
{
public:
  explicit FaceSubdivisionByAreaMeshModifier(float max_area) : max_area_(max_area) {}

protected:
  int subdivide(nb2_context* ctx, std::int32_t offset_rgba, nb2_mesh_file_info* info) const override
  {
    return nb2_face_subdivision_by_area(ctx, max_area_, offset_rgba, 0, info);
  }
  float max_area_;
};

// ---------------------------------------------------------------------------------------------------------------------
// tool path modifiers
// ---------------------------------------------------------------------------------------------------------------------
/** A noether::ToolPathModifier that runs on the GPU (nb2_result_modify): the poses go to the device once, every modifier
 *  of the chain is one streaming kernel over them, and they come back once.  One object stands for one modifier of the
 *  reference, or — built by the CompoundToolPathModifier plugin below — for a whole chain of them. */
class DeviceToolPathModifier : public noether::ToolPathModifier
{
public:
  explicit DeviceToolPathModifier(std::vector<nb2_modifier> chain) : chain_(std::move(chain)) {}
  const std::vector<nb2_modifier>& chain() const { return chain_; }

  noether::ToolPaths modify(noether::ToolPaths tool_paths) const override final
  {
    static_assert(sizeof(Eigen::Isometry3d) == 16 * sizeof(double), "Isometry3d is not a bare 4x4 double matrix");
    std::vector<std::uint32_t> path_off(1, 0u), seg_off(1, 0u);
    std::vector<const double*> seg_data;
    std::uint64_t w = 0;
    for (const noether::ToolPath& path : tool_paths)
    {
      for (const noether::ToolPathSegment& seg : path)
      {
        seg_data.push_back(seg.empty() ? nullptr : seg.front().matrix().data());
        w += seg.size();
        if (w > 0xfffffff0ull)
          throw std::runtime_error("tool paths exceed 2^32 waypoints");
        seg_off.push_back(static_cast<std::uint32_t>(w));
      }
      path_off.push_back(static_cast<std::uint32_t>(seg_data.size()));
    }
    Device& dev = Device::instance();
    std::lock_guard<std::mutex> lock(dev.mutex());
    Trace trace("ToolPathModifier::modify");
    nb2_result* in = nullptr;
    Device::check(nb2_result_from_pose_segments(dev.ctx(), tool_paths.size(), path_off.data(), seg_data.size(), seg_off.data(),
                                                seg_data.data(), &in));
    std::unique_ptr<nb2_result, void (*)(nb2_result*)> in_guard(in, nb2_result_free);
    nb2_result* r = nullptr;
    Device::check(nb2_result_modify(dev.ctx(), in, chain_.data(), static_cast<int>(chain_.size()), &r));
    trace.lap("nb2_result_modify");
    std::unique_ptr<nb2_result, void (*)(nb2_result*)> guard(r, nb2_result_free);
    noether::ToolPaths out = toToolPaths(r);
    trace.lap("ToolPaths");
    return out;
  }

private:
  std::vector<nb2_modifier> chain_;
};

namespace
{
nb2_modifier record(int kind)
{
  nb2_modifier m{};
  m.kind = kind;
  return m;
}
/** Eigen::Vector3d as serialization.h:59-83 reads it */
void readVector3(const YAML::Node& node, double out[3])
{
  out[0] = YAML::getMember<double>(node, "x");
  out[1] = YAML::getMember<double>(node, "y");
  out[2] = YAML::getMember<double>(node, "z");
}
/** Eigen::Isometry3d as serialization.h:41-55 reads it: Translation3d(x, y, z) * Quaterniond(qw, qx, qy, qz), column-major
 *  (the quaternion as given; Eigen 3.4 QuaternionBase::toRotationMatrix) */
void readIsometry(const YAML::Node& node, double T[16])
{
  const double x = YAML::getMember<double>(node, "qx"), y = YAML::getMember<double>(node, "qy"),
               z = YAML::getMember<double>(node, "qz"), w = YAML::getMember<double>(node, "qw");
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  T[0] = 1.0 - (tyy + tzz), T[4] = txy - twz, T[8] = txz + twy;
  T[1] = txy + twz, T[5] = 1.0 - (txx + tzz), T[9] = tyz - twx;
  T[2] = txz - twy, T[6] = tyz + twx, T[10] = 1.0 - (txx + tyy);
  T[3] = T[7] = T[11] = 0.0;
  readVector3(node, T + 12);
  T[15] = 1.0;
}
}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// plugins: same aliases, sections and YAML keys as the reference (src/plugins.cpp)
// ---------------------------------------------------------------------------------------------------------------------
struct Plugin_PrincipalAxisDirectionGenerator : public noether::Plugin<noether::DirectionGenerator>
{
  std::unique_ptr<noether::DirectionGenerator> create(const YAML::Node& config,
                                                      std::shared_ptr<const noether::Factory> /*factory*/) const override final
  {
    // principal_axis_direction_generator.cpp:41-45
    return std::make_unique<PrincipalAxisDirectionGenerator>(YAML::getMember<double>(config, "rotation_offset"));
  }
};

struct Plugin_PCARotatedDirectionGenerator : public noether::Plugin<noether::DirectionGenerator>
{
  std::unique_ptr<noether::DirectionGenerator> create(const YAML::Node& config,
                                                      std::shared_ptr<const noether::Factory> factory) const override final
  {
    // plugins.cpp:91-107: a nested generator (created through the factory) and `rotation_offset`
    auto dir_gen = factory->createDirectionGenerator(YAML::getMember<YAML::Node>(config, "direction_generator"));
    return std::make_unique<PCARotatedDirectionGenerator>(std::move(dir_gen), YAML::getMember<double>(config, "rotation_offset"));
  }
};

struct Plugin_CentroidOriginGenerator : public noether::Plugin<noether::OriginGenerator>
{
  std::unique_ptr<noether::OriginGenerator> create(const YAML::Node& /*config*/,
                                                   std::shared_ptr<const noether::Factory> /*factory*/) const override final
  {
    return std::make_unique<CentroidOriginGenerator>();
  }
};

struct Plugin_AABBCenterOriginGenerator : public noether::Plugin<noether::OriginGenerator>
{
  std::unique_ptr<noether::OriginGenerator> create(const YAML::Node& /*config*/,
                                                   std::shared_ptr<const noether::Factory> /*factory*/) const override final
  {
    return std::make_unique<AABBCenterOriginGenerator>();
  }
};

struct Plugin_NormalsFromMeshFacesMeshModifier : public noether::Plugin<noether::MeshModifier>
{
  std::unique_ptr<noether::MeshModifier> create(const YAML::Node& /*config*/,
                                                std::shared_ptr<const noether::Factory> /*factory*/) const override final
  {
    return std::make_unique<NormalsFromMeshFacesMeshModifier>();
  }
};

/** NormalEstimationPCL: YAML keys radius, vx, vy, vz (normal_estimation_pcl.cpp:67-75) */
struct Plugin_NormalEstimationPCLMeshModifier : public noether::Plugin<noether::MeshModifier>
{
  std::unique_ptr<noether::MeshModifier> create(const YAML::Node& config,
                                                std::shared_ptr<const noether::Factory> /*factory*/) const override final
  {
    return std::make_unique<NormalEstimationPCLMeshModifier>(YAML::getMember<double>(config, "radius"), YAML::getMember<double>(config, "vx"),
                                                             YAML::getMember<double>(config, "vy"), YAML::getMember<double>(config, "vz"));
  }
};

/** EuclideanClustering: YAML keys tolerance, min_cluster_size, max_cluster_size (euclidean_clustering_modifier.cpp:79-85) */
struct Plugin_EuclideanClusteringMeshModifier : public noether::Plugin<noether::MeshModifier>
{
  std::unique_ptr<noether::MeshModifier> create(const YAML::Node& config,
                                                std::shared_ptr<const noether::Factory> /*factory*/) const override final
  {
    return std::make_unique<EuclideanClusteringMeshModifier>(YAML::getMember<double>(config, "tolerance"),
                                                             YAML::getMember<int>(config, "min_cluster_size"),
                                                             YAML::getMember<int>(config, "max_cluster_size"));
  }
};

/** FaceMidpointSubdivision: YAML key n_iterations, read as a float and cast (face_subdivision_modifier.cpp:354-359) */
struct Plugin_FaceMidpointSubdivisionMeshModifier : public noether::Plugin<noether::MeshModifier>
{
  std::unique_ptr<noether::MeshModifier> create(const YAML::Node& config,
                                                std::shared_ptr<const noether::Factory> /*factory*/) const override final
  {
    return std::make_unique<FaceMidpointSubdivisionMeshModifier>(static_cast<unsigned>(YAML::getMember<float>(config, "n_iterations")));
  }
};

/** FaceSubdivisionByArea: YAML key max_area (face_subdivision_modifier.cpp:368-373) */
struct Plugin_FaceSubdivisionByAreaMeshModifier : public noether::Plugin<noether::MeshModifier>
{
  std::unique_ptr<noether::MeshModifier> create(const YAML::Node& config,
                                                std::shared_ptr<const noether::Factory> /*factory*/) const override final
  {
    return std::make_unique<FaceSubdivisionByAreaMeshModifier>(YAML::getMember<float>(config, "max_area"));
  }
};

struct Plugin_PlaneSlicerRasterPlanner : public noether::Plugin<noether::ToolPathPlanner>
{
  std::unique_ptr<noether::ToolPathPlanner> create(const YAML::Node& config,
                                                   std::shared_ptr<const noether::Factory> factory) const override final
  {
    // nested generators are created through the factory, so the reference's FixedDirection / FixedOrigin / PCARotated /
    // Offset generators compose with this planner (plugins.cpp:257-266)
    auto dir_gen = factory->createDirectionGenerator(YAML::getMember<YAML::Node>(config, "direction_generator"));
    auto origin_gen = factory->createOriginGenerator(YAML::getMember<YAML::Node>(config, "origin_generator"));
    auto tpp = std::make_unique<PlaneSlicerRasterPlanner>(std::move(dir_gen), std::move(origin_gen));
    tpp->setLineSpacing(YAML::getMember<double>(config, "line_spacing"));
    tpp->generateRastersBidirectionally(YAML::getMember<bool>(config, "bidirectional"));
    return tpp;
  }
};

struct Plugin_PlaneSlicerLegacyRasterPlanner : public noether::Plugin<noether::ToolPathPlanner>
{
  std::unique_ptr<noether::ToolPathPlanner> create(const YAML::Node& config,
                                                   std::shared_ptr<const noether::Factory> factory) const override final
  {
    auto dir_gen = factory->createDirectionGenerator(YAML::getMember<YAML::Node>(config, "direction_generator"));
    auto origin_gen = factory->createOriginGenerator(YAML::getMember<YAML::Node>(config, "origin_generator"));
    auto tpp = std::make_unique<PlaneSlicerLegacyRasterPlanner>(std::move(dir_gen), std::move(origin_gen),
                                                                YAML::getMember<double>(config, "point_spacing"),
                                                                YAML::getMember<double>(config, "min_segment_size"),
                                                                YAML::getMember<double>(config, "min_hole_size"));
    tpp->setLineSpacing(YAML::getMember<double>(config, "line_spacing"));
    tpp->generateRastersBidirectionally(YAML::getMember<bool>(config, "bidirectional"));
    return tpp;
  }
};

/** The tool path modifiers that run on the device, under the reference's aliases and YAML keys (plugins.cpp:182-199 and
 *  the YAML::convert<>::decode of each modifier) */
template <int KIND>
struct Plugin_DeviceToolPathModifier : public noether::Plugin<noether::ToolPathModifier>
{
  std::unique_ptr<noether::ToolPathModifier> create(const YAML::Node& config,
                                                    std::shared_ptr<const noether::Factory> /*factory*/) const override final
  {
    nb2_modifier m = record(KIND);
    switch (KIND)
    {
      case NB2_MOD_LINEAR_APPROACH:   // linear_approach_modifier.cpp:68-73
      case NB2_MOD_LINEAR_DEPARTURE:  // linear_departure_modifier.cpp (same keys)
        readVector3(YAML::getMember<YAML::Node>(config, "offset"), m.v);
        m.n_points = YAML::getMember<int>(config, "n_points");
        break;
      case NB2_MOD_OFFSET:  // offset_modifier.cpp:37-41
        readIsometry(YAML::getMember<YAML::Node>(config, "offset"), m.v);
        break;
      case NB2_MOD_FIXED_ORIENTATION:  // fixed_orientation_modifier.cpp:50-54: the member is assigned, not normalised
        readVector3(YAML::getMember<YAML::Node>(config, "ref_x_dir"), m.v);
        m.v[3] = 1.0;
        break;
      case NB2_MOD_TOOL_DRAG_ORIENTATION:         // tool_drag_orientation_modifier.cpp:51-57
      case NB2_MOD_BIASED_TOOL_DRAG_ORIENTATION:  // biased_tool_drag_orientation_modifier.cpp (same keys)
        m.v[0] = YAML::getMember<double>(config, "angle_offset");
        m.v[1] = YAML::getMember<double>(config, "tool_radius");
        break;
      case NB2_MOD_JOIN_CLOSE_SEGMENTS:  // join_close_segments_modifier.cpp:55-59
        m.v[0] = YAML::getMember<double>(config, "distance");
        break;
      case NB2_MOD_MINIMUM_SEGMENT_LENGTH:  // minimum_segment_length_modifier.cpp:53-58
        m.v[0] = YAML::getMember<double>(config, "minimum_length");
        break;
      case NB2_MOD_CIRCULAR_LEAD_IN:   // circular_lead_in_modifier.cpp:60-66
      case NB2_MOD_CIRCULAR_LEAD_OUT:  // circular_lead_out_modifier.cpp (same keys)
        m.v[0] = YAML::getMember<double>(config, "arc_angle");
        m.v[1] = YAML::getMember<double>(config, "arc_radius");
        m.n_points = YAML::getMember<int>(config, "n_points");
        break;
      case NB2_MOD_MOVING_AVERAGE_ORIENTATION_SMOOTHING:  // moving_average_orientation_smoothing_modifier.cpp:99-105
        m.n_points = static_cast<int>(static_cast<std::size_t>(YAML::getMember<double>(config, "window_size")));
        break;
      default:  // SnakeOrganization, UniformOrientation, Concatenate, DirectionOfTravelOrientation: no keys
        break;
    }
    return std::make_unique<DeviceToolPathModifier>(std::vector<nb2_modifier>{ m });
  }
};
using Plugin_SnakeOrganization = Plugin_DeviceToolPathModifier<NB2_MOD_SNAKE_ORGANIZATION>;
using Plugin_LinearApproach = Plugin_DeviceToolPathModifier<NB2_MOD_LINEAR_APPROACH>;
using Plugin_LinearDeparture = Plugin_DeviceToolPathModifier<NB2_MOD_LINEAR_DEPARTURE>;
using Plugin_Offset = Plugin_DeviceToolPathModifier<NB2_MOD_OFFSET>;
using Plugin_FixedOrientation = Plugin_DeviceToolPathModifier<NB2_MOD_FIXED_ORIENTATION>;
using Plugin_UniformOrientation = Plugin_DeviceToolPathModifier<NB2_MOD_UNIFORM_ORIENTATION>;
using Plugin_Concatenate = Plugin_DeviceToolPathModifier<NB2_MOD_CONCATENATE>;
using Plugin_ToolDragOrientation = Plugin_DeviceToolPathModifier<NB2_MOD_TOOL_DRAG_ORIENTATION>;
using Plugin_BiasedToolDragOrientation = Plugin_DeviceToolPathModifier<NB2_MOD_BIASED_TOOL_DRAG_ORIENTATION>;
using Plugin_DirectionOfTravelOrientation = Plugin_DeviceToolPathModifier<NB2_MOD_DIRECTION_OF_TRAVEL_ORIENTATION>;
using Plugin_RasterOrganization = Plugin_DeviceToolPathModifier<NB2_MOD_RASTER_ORGANIZATION>;
using Plugin_JoinCloseSegments = Plugin_DeviceToolPathModifier<NB2_MOD_JOIN_CLOSE_SEGMENTS>;
using Plugin_MinimumSegmentLength = Plugin_DeviceToolPathModifier<NB2_MOD_MINIMUM_SEGMENT_LENGTH>;
using Plugin_CircularLeadIn = Plugin_DeviceToolPathModifier<NB2_MOD_CIRCULAR_LEAD_IN>;
using Plugin_CircularLeadOut = Plugin_DeviceToolPathModifier<NB2_MOD_CIRCULAR_LEAD_OUT>;
using Plugin_MovingAverageOrientationSmoothing = Plugin_DeviceToolPathModifier<NB2_MOD_MOVING_AVERAGE_ORIENTATION_SMOOTHING>;

/** CompoundModifier (compound_modifier.cpp) whose runs of consecutive device modifiers are fused into one device chain;
 *  any other modifier the factory returns (the reference's own) is applied in its place in the sequence. */
class CompoundToolPathModifier : public noether::ToolPathModifier
{
public:
  explicit CompoundToolPathModifier(std::vector<noether::ToolPathModifier::ConstPtr> modifiers)
  {
    for (auto& m : modifiers)
    {
      const auto* dev = dynamic_cast<const DeviceToolPathModifier*>(m.get());
      const auto* last = stages_.empty() ? nullptr : dynamic_cast<const DeviceToolPathModifier*>(stages_.back().get());
      if (dev && last)
      {
        std::vector<nb2_modifier> chain = last->chain();
        chain.insert(chain.end(), dev->chain().begin(), dev->chain().end());
        stages_.back() = std::make_unique<DeviceToolPathModifier>(std::move(chain));
      }
      else
        stages_.push_back(std::move(m));
    }
  }
  noether::ToolPaths modify(noether::ToolPaths tool_paths) const override final
  {
    for (const auto& stage : stages_)
      tool_paths = stage->modify(std::move(tool_paths));
    return tool_paths;
  }
  std::size_t stages() const { return stages_.size(); }

private:
  std::vector<noether::ToolPathModifier::ConstPtr> stages_;
};

struct Plugin_CompoundToolPathModifier : public noether::Plugin<noether::ToolPathModifier>
{
  std::unique_ptr<noether::ToolPathModifier> create(const YAML::Node& config,
                                                    std::shared_ptr<const noether::Factory> factory) const override final
  {
    // plugins.cpp:202-218
    std::vector<noether::ToolPathModifier::ConstPtr> modifiers;
    const YAML::Node modifiers_config = YAML::getMember<YAML::Node>(config, "modifiers");
    for (const YAML::Node& entry : modifiers_config)
      modifiers.push_back(factory->createToolPathModifier(entry));
    return std::make_unique<CompoundToolPathModifier>(std::move(modifiers));
  }
};

}  // namespace noether_b200

EXPORT_TOOL_PATH_MODIFIER_PLUGIN(noether_b200::Plugin_SnakeOrganization, SnakeOrganization)
EXPORT_TOOL_PATH_MODIFIER_PLUGIN(noether_b200::Plugin_LinearApproach, LinearApproach)
EXPORT_TOOL_PATH_MODIFIER_PLUGIN(noether_b200::Plugin_LinearDeparture, LinearDeparture)
EXPORT_TOOL_PATH_MODIFIER_PLUGIN(noether_b200::Plugin_Offset, Offset)
EXPORT_TOOL_PATH_MODIFIER_PLUGIN(noether_b200::Plugin_FixedOrientation, FixedOrientation)
EXPORT_TOOL_PATH_MODIFIER_PLUGIN(noether_b200::Plugin_UniformOrientation, UniformOrientation)
EXPORT_TOOL_PATH_MODIFIER_PLUGIN(noether_b200::Plugin_Concatenate, Concatenate)
EXPORT_TOOL_PATH_MODIFIER_PLUGIN(noether_b200::Plugin_ToolDragOrientation, ToolDragOrientation)
EXPORT_TOOL_PATH_MODIFIER_PLUGIN(noether_b200::Plugin_BiasedToolDragOrientation, BiasedToolDragOrientation)
EXPORT_TOOL_PATH_MODIFIER_PLUGIN(noether_b200::Plugin_DirectionOfTravelOrientation, DirectionOfTravelOrientation)
EXPORT_TOOL_PATH_MODIFIER_PLUGIN(noether_b200::Plugin_RasterOrganization, RasterOrganization)
EXPORT_TOOL_PATH_MODIFIER_PLUGIN(noether_b200::Plugin_JoinCloseSegments, JoinCloseSegments)
EXPORT_TOOL_PATH_MODIFIER_PLUGIN(noether_b200::Plugin_MinimumSegmentLength, MinimumSegmentLength)
EXPORT_TOOL_PATH_MODIFIER_PLUGIN(noether_b200::Plugin_CircularLeadIn, CircularLeadIn)
EXPORT_TOOL_PATH_MODIFIER_PLUGIN(noether_b200::Plugin_CircularLeadOut, CircularLeadOut)
EXPORT_TOOL_PATH_MODIFIER_PLUGIN(noether_b200::Plugin_MovingAverageOrientationSmoothing, MovingAverageOrientationSmoothing)
EXPORT_TOOL_PATH_MODIFIER_PLUGIN(noether_b200::Plugin_CompoundToolPathModifier, CompoundToolPathModifier)
EXPORT_MESH_MODIFIER_PLUGIN(noether_b200::Plugin_NormalsFromMeshFacesMeshModifier, NormalsFromMeshFaces)
EXPORT_MESH_MODIFIER_PLUGIN(noether_b200::Plugin_EuclideanClusteringMeshModifier, EuclideanClustering)
EXPORT_MESH_MODIFIER_PLUGIN(noether_b200::Plugin_NormalEstimationPCLMeshModifier, NormalEstimationPCL)
/** test hook (tests/cpp/plugin_shim_test.cpp --cpu-only): the host side of EuclideanClustering without a device */
extern "C" void noether_b200_cut_cluster_meshes(const pcl::PolygonMesh* mesh, const std::uint32_t* labels, int min_cluster_size,
                                                int max_cluster_size, std::vector<pcl::PolygonMesh>* out)
{
  *out = noether_b200::cutClusterMeshes(*mesh, labels, min_cluster_size, max_cluster_size);
}
EXPORT_MESH_MODIFIER_PLUGIN(noether_b200::Plugin_FaceMidpointSubdivisionMeshModifier, FaceMidpointSubdivision)
EXPORT_MESH_MODIFIER_PLUGIN(noether_b200::Plugin_FaceSubdivisionByAreaMeshModifier, FaceSubdivisionByArea)
EXPORT_DIRECTION_GENERATOR_PLUGIN(noether_b200::Plugin_PrincipalAxisDirectionGenerator, PrincipalAxis)
EXPORT_DIRECTION_GENERATOR_PLUGIN(noether_b200::Plugin_PCARotatedDirectionGenerator, PCARotated)
EXPORT_ORIGIN_GENERATOR_PLUGIN(noether_b200::Plugin_CentroidOriginGenerator, Centroid)
EXPORT_ORIGIN_GENERATOR_PLUGIN(noether_b200::Plugin_AABBCenterOriginGenerator, AABBCenter)
EXPORT_TOOL_PATH_PLANNER_PLUGIN(noether_b200::Plugin_PlaneSlicerRasterPlanner, PlaneSlicer)
EXPORT_TOOL_PATH_PLANNER_PLUGIN(noether_b200::Plugin_PlaneSlicerLegacyRasterPlanner, PlaneSlicerLegacy)
