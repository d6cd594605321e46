// The reference's own planner / mesh-modifier tests, re-expressed through the plugin boundary of noether_b200:
//
//   RasterPlannerTestFixture.FlatSquareMesh      noether_tpp/test/tool_path_planner_utest.cpp:74-120
//   RasterPlannerTestFixture.SemiPlanarMeshFile  noether_tpp/test/tool_path_planner_utest.cpp:122-177
//   MeshModifier field retention                 noether_tpp/test/mesh_modifier_utest.cpp:104-113,154-183
//
// The components are created the way the reference tests create them: by alias, from a YAML node, through a
// noether::Factory that resolves plugins out of a shared library (tool_path_planner_utest.cpp:51-59).  Here that
// library is plugin/b200_plugins.cpp compiled against plugin/compat (the development image has no PCL / Eigen /
// yaml-cpp / boost_plugin_loader); `FixedDirection` and `FixedOrigin`, which the reference tests take from the
// reference's plugin library, are supplied by this executable.  Needs a B200: run by tests/test_plugin_shim.py.
//
//   usage: plugin_shim_test <libnoether_b200_plugins.so> <wavy_mesh_with_hole.ply>
#include <noether_tpp/plugin_interface.h>
#include <noether_tpp/serialization.h>

#include <dlfcn.h>

#include <chrono>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>

namespace
{
int g_failures = 0;
#define CHECK(cond)                                                                                                    \
  do                                                                                                                   \
  {                                                                                                                    \
    if (!(cond))                                                                                                       \
    {                                                                                                                  \
      std::printf("  FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);                                                  \
      ++g_failures;                                                                                                    \
    }                                                                                                                  \
  } while (0)

Eigen::Vector3d readVector(const YAML::Node& n)
{
  return Eigen::Vector3d(YAML::getMember<double>(n, "x"), YAML::getMember<double>(n, "y"), YAML::getMember<double>(n, "z"));
}
YAML::Node vectorNode(double x, double y, double z)
{
  YAML::Node n;
  n["x"] = x;
  n["y"] = y;
  n["z"] = z;
  return n;
}

// ---- the two reference generators the reference tests configure (fixed_direction_generator.cpp, fixed_origin_generator.cpp)
struct FixedDirectionGenerator : noether::DirectionGenerator
{
  Eigen::Vector3d dir;
  Eigen::Vector3d generate(const pcl::PolygonMesh&) const override { return dir; }
};
struct FixedOriginGenerator : noether::OriginGenerator
{
  Eigen::Vector3d origin;
  Eigen::Vector3d generate(const pcl::PolygonMesh&) const override { return origin; }
};
struct Plugin_FixedDirection : noether::Plugin<noether::DirectionGenerator>
{
  std::unique_ptr<noether::DirectionGenerator> create(const YAML::Node& config,
                                                      std::shared_ptr<const noether::Factory>) const override
  {
    auto g = std::make_unique<FixedDirectionGenerator>();
    g->dir = readVector(YAML::getMember<YAML::Node>(config, "direction"));
    return g;
  }
};
struct Plugin_FixedOrigin : noether::Plugin<noether::OriginGenerator>
{
  std::unique_ptr<noether::OriginGenerator> create(const YAML::Node& config,
                                                   std::shared_ptr<const noether::Factory>) const override
  {
    auto g = std::make_unique<FixedOriginGenerator>();
    g->origin = readVector(YAML::getMember<YAML::Node>(config, "origin"));
    return g;
  }
};
}  // namespace
EXPORT_DIRECTION_GENERATOR_PLUGIN(Plugin_FixedDirection, FixedDirection)
EXPORT_ORIGIN_GENERATOR_PLUGIN(Plugin_FixedOrigin, FixedOrigin)

namespace
{
// ---- fixtures ---------------------------------------------------------------------------------------------------------
void addField(pcl::PCLPointCloud2& c, const char* name, unsigned offset, std::uint8_t type = pcl::PCLPointField::FLOAT32)
{
  pcl::PCLPointField f;
  f.name = name;
  f.offset = offset;
  f.datatype = type;
  f.count = 1;
  c.fields.push_back(f);
}

/** createPlaneMesh (utils.cpp:249-291): 4 pcl::PointNormal vertices (48-byte records), polygons {0,1,3} {1,2,3};
 *  R (row-major 3x3) and t place it in space as the reference test's random transform does */
pcl::PolygonMesh planeMesh(float lx, float ly, const double R[9], const double t[3])
{
  pcl::PolygonMesh mesh;
  pcl::PCLPointCloud2& c = mesh.cloud;
  c.width = 4;
  c.height = 1;
  c.point_step = 48;
  c.row_step = 4 * 48;
  c.is_dense = 1;
  addField(c, "x", 0), addField(c, "y", 4), addField(c, "z", 8);
  addField(c, "normal_x", 16), addField(c, "normal_y", 20), addField(c, "normal_z", 24), addField(c, "curvature", 32);
  const float pts[4][3] = { { -lx / 2, ly / 2, 0 }, { -lx / 2, -ly / 2, 0 }, { lx / 2, -ly / 2, 0 }, { lx / 2, ly / 2, 0 } };
  c.data.assign(4 * 48, 0);
  for (int i = 0; i < 4; ++i)
  {
    float rec[12] = { 0 };
    for (int r = 0; r < 3; ++r)
    {
      rec[r] = static_cast<float>(R[3 * r] * pts[i][0] + R[3 * r + 1] * pts[i][1] + R[3 * r + 2] * pts[i][2] + t[r]);
      rec[4 + r] = static_cast<float>(R[3 * r + 2]);  // R * (0, 0, 1)
    }
    std::memcpy(c.data.data() + 48 * i, rec, 48);
  }
  pcl::Vertices a, b;
  a.vertices = { 0, 1, 3 };
  b.vertices = { 1, 2, 3 };
  mesh.polygons = { a, b };
  return mesh;
}

/** wavy_mesh_with_hole.ply (binary little endian: x y z nx ny nz float, r g b a uchar; faces uchar n + int[n]) as
 *  pcl::io::loadPolygonFile lays it out: PointXYZRGBNormal, 48-byte records */
pcl::PolygonMesh loadPly(const std::string& path)
{
  std::ifstream f(path, std::ios::binary);
  if (!f)
    throw std::runtime_error("cannot open " + path);
  std::string line;
  std::size_t nv = 0, nf = 0;
  while (std::getline(f, line))
  {
    std::istringstream ss(line);
    std::string a, b;
    ss >> a >> b;
    if (a == "element" && b == "vertex")
      ss >> nv;
    if (a == "element" && b == "face")
      ss >> nf;
    if (a == "end_header")
      break;
  }
  pcl::PolygonMesh mesh;
  pcl::PCLPointCloud2& c = mesh.cloud;
  c.width = static_cast<unsigned>(nv);
  c.height = 1;
  c.point_step = 48;
  c.row_step = c.point_step * c.width;
  addField(c, "x", 0), addField(c, "y", 4), addField(c, "z", 8);
  addField(c, "normal_x", 16), addField(c, "normal_y", 20), addField(c, "normal_z", 24);
  addField(c, "rgb", 32), addField(c, "curvature", 36);
  c.data.assign(nv * 48, 0);
  for (std::size_t i = 0; i < nv; ++i)
  {
    char rec[28];
    f.read(rec, 28);
    std::memcpy(c.data.data() + 48 * i, rec, 12);            // x y z
    std::memcpy(c.data.data() + 48 * i + 16, rec + 12, 12);  // normals
    std::memcpy(c.data.data() + 48 * i + 32, rec + 24, 4);   // rgba
  }
  for (std::size_t i = 0; i < nf; ++i)
  {
    unsigned char n;
    f.read(reinterpret_cast<char*>(&n), 1);
    pcl::Vertices v;
    for (int k = 0; k < n; ++k)
    {
      std::int32_t id;
      f.read(reinterpret_cast<char*>(&id), 4);
      v.vertices.push_back(static_cast<pcl::index_t>(id));
    }
    mesh.polygons.push_back(v);
  }
  if (!f)
    throw std::runtime_error("truncated PLY " + path);
  return mesh;
}

/** the reference tests' parameter YAML (tool_path_planner_utest.cpp:259-275), including the stale keys the plugin must
 *  ignore */
YAML::Node planeSlicerConfig(const char* alias, double line_spacing, const Eigen::Vector3d& dir, const Eigen::Vector3d& origin)
{
  YAML::Node config;
  config["name"] = alias;
  config["line_spacing"] = line_spacing;
  config["point_spacing"] = 0.25;
  config["min_segment_size"] = 0.5;
  config["min_hole_size"] = 0.0;
  config["search_radius"] = 0.1;
  config["bidirectional"] = true;
  YAML::Node dg;
  dg["name"] = "FixedDirection";
  dg["direction"] = vectorNode(dir.x(), dir.y(), dir.z());
  config["direction_generator"] = dg;
  YAML::Node og;
  og["name"] = "FixedOrigin";
  og["origin"] = vectorNode(origin.x(), origin.y(), origin.z());
  config["origin_generator"] = og;
  return config;
}

Eigen::Vector3d sub(const Eigen::Vector3d& a, const Eigen::Vector3d& b) { return { a.x() - b.x(), a.y() - b.y(), a.z() - b.z() }; }
double dot(const Eigen::Vector3d& a, const Eigen::Vector3d& b) { return a.x() * b.x() + a.y() * b.y() + a.z() * b.z(); }
double norm(const Eigen::Vector3d& a) { return std::sqrt(dot(a, a)); }

// ---- tests --------------------------------------------------------------------------------------------------------------
void testFlatSquareMesh(const noether::Factory& factory, const char* alias)
{
  std::printf("[FlatSquareMesh / %s]\n", alias);
  // a rigid transform with no special axis (the reference draws one from mt19937(0))
  const double a = 0.7, b = -1.1, c = 2.3;
  const double ca = std::cos(a), sa = std::sin(a), cb = std::cos(b), sb = std::sin(b), cc = std::cos(c), sc = std::sin(c);
  const double R[9] = { cc * cb, cc * sb * sa - sc * ca, cc * sb * ca + sc * sa, sc * cb, sc * sb * sa + cc * ca,
                        sc * sb * ca - cc * sa,  -sb, cb * sa, cb * ca };
  const double t[3] = { 3.5, -1.25, 4.0 };
  const pcl::PolygonMesh mesh = planeMesh(1.f, 1.f, R, t);
  const Eigen::Vector3d direction(R[0], R[3], R[6]);  // R * UnitX
  const unsigned n_lines = 10;
  YAML::Node config = planeSlicerConfig(alias, 0.95 * 1.0 / n_lines, direction, Eigen::Vector3d(t[0], t[1], t[2]));
  std::shared_ptr<const noether::ToolPathPlanner> planner = factory.createToolPathPlanner(config);
  noether::ToolPaths tool_paths = planner->plan(mesh);
  CHECK(tool_paths.size() == n_lines + 1);
  for (const noether::ToolPath& path : tool_paths)
  {
    CHECK(path.size() == 1);
    if (path.empty() || path.front().size() < 2)
    {
      CHECK(false);
      continue;
    }
    const Eigen::Vector3d first = path.front().front().translation(), last = path.back().back().translation();
    const Eigen::Vector3d d = sub(last, first);
    CHECK(std::abs(dot(d, direction)) / norm(d) > std::cos(1.0 * M_PI / 180.0));
    for (const auto& wp : path.front())
    {
      // every pose is a rigid transform whose z axis is the mesh normal R * UnitZ
      const auto& m = wp.matrix();
      CHECK(m(3, 0) == 0.0 && m(3, 1) == 0.0 && m(3, 2) == 0.0 && m(3, 3) == 1.0);
      CHECK(std::abs(m(0, 2) - R[2]) < 1e-6 && std::abs(m(1, 2) - R[5]) < 1e-6 && std::abs(m(2, 2) - R[8]) < 1e-6);
    }
  }
}

void testSemiPlanarMeshFile(const noether::Factory& factory, const pcl::PolygonMesh& mesh, const char* alias, bool legacy)
{
  std::printf("[SemiPlanarMeshFile / %s]\n", alias);
  const Eigen::Vector3d direction(1, 0, 0);
  const unsigned n_lines = 10;
  YAML::Node config = planeSlicerConfig(alias, 10.5 / n_lines, direction, Eigen::Vector3d(0, 0, 0));
  std::shared_ptr<const noether::ToolPathPlanner> planner = factory.createToolPathPlanner(config);
  noether::ToolPaths tool_paths = planner->plan(mesh);
  CHECK(tool_paths.size() == n_lines);
  // planImpl returns the rasters in lattice order; the reference's RasterOrganizationModifier (applied by the real
  // RasterPlanner::plan) may reverse that order, so the hole rows are 4,5,6 counted from one end or the other
  unsigned two_from_front = 0, two_from_back = 0;
  for (std::size_t i = 0; i < tool_paths.size(); ++i)
  {
    const noether::ToolPath& path = tool_paths[i];
    CHECK(path.size() >= 1);
    CHECK(!path.empty() && path.front().size() >= 2);
    if (i > 3 && i < 7)
      two_from_front += path.size() == 2;
    if (tool_paths.size() - 1 - i > 3 && tool_paths.size() - 1 - i < 7)
      two_from_back += path.size() == 2;
    if (path.empty() || path.front().empty())
      continue;
    // direction check of the reference (10 degrees); segments of a raster may be stored in either order before the
    // organisation modifier runs, so take the extreme waypoints along the direction
    double lo = 1e300, hi = -1e300;
    Eigen::Vector3d pl, ph;
    for (const auto& seg : path)
      for (const auto& wp : seg)
      {
        const double k = dot(wp.translation(), direction);
        if (k < lo)
          lo = k, pl = wp.translation();
        if (k > hi)
          hi = k, ph = wp.translation();
      }
    const Eigen::Vector3d d = sub(ph, pl);
    CHECK(std::abs(dot(d, direction)) / norm(d) > std::cos(10.0 * M_PI / 180.0));
  }
  CHECK(two_from_front == 3 || two_from_back == 3);
  (void)legacy;
}

void testNormalsFieldRetention(const noether::Factory& factory, const pcl::PolygonMesh& mesh)
{
  std::printf("[MeshModifier field retention / NormalsFromMeshFaces]\n");
  YAML::Node config;
  config["name"] = "NormalsFromMeshFaces";
  std::shared_ptr<const noether::MeshModifier> mod = factory.createMeshModifier(config);
  const std::vector<pcl::PolygonMesh> out = mod->modify(mesh);
  CHECK(out.size() == 1);
  if (out.empty())
    return;
  const pcl::PolygonMesh& m = out.front();
  CHECK(m.polygons.size() == mesh.polygons.size());
  CHECK(m.cloud.width * m.cloud.height == mesh.cloud.width * mesh.cloud.height);
  CHECK(m.cloud.data.size() == static_cast<std::size_t>(m.cloud.point_step) * m.cloud.width * m.cloud.height);
  // every input field is still present, with its datatype and count (mesh_modifier_utest.cpp:104-113)
  for (const auto& f : mesh.cloud.fields)
  {
    bool found = false;
    for (const auto& g : m.cloud.fields)
      found = found || (g.name == f.name && g.datatype == f.datatype && g.count == f.count);
    CHECK(found);
  }
  // the new normals are unit vectors that agree with the file's own vertex normals on this smooth surface
  unsigned off_new = 0, off_old = 0, off_xyz_new = 0, off_xyz_old = 0;
  for (const auto& g : m.cloud.fields)
  {
    if (g.name == "normal_x")
      off_new = g.offset;
    if (g.name == "x")
      off_xyz_new = g.offset;
  }
  for (const auto& g : mesh.cloud.fields)
  {
    if (g.name == "normal_x")
      off_old = g.offset;
    if (g.name == "x")
      off_xyz_old = g.offset;
  }
  double worst = 1.0;
  bool xyz_same = true;
  const std::size_t n = static_cast<std::size_t>(mesh.cloud.width) * mesh.cloud.height;
  for (std::size_t i = 0; i < n; ++i)
  {
    float a[3], b[3];
    std::memcpy(a, m.cloud.data.data() + i * m.cloud.point_step + off_new, 12);
    std::memcpy(b, mesh.cloud.data.data() + i * mesh.cloud.point_step + off_old, 12);
    const double na = std::sqrt(double(a[0]) * a[0] + double(a[1]) * a[1] + double(a[2]) * a[2]);
    CHECK(std::abs(na - 1.0) < 1e-5);
    worst = std::min(worst, (double(a[0]) * b[0] + double(a[1]) * b[1] + double(a[2]) * b[2]) / na);
    xyz_same = xyz_same && std::memcmp(m.cloud.data.data() + i * m.cloud.point_step + off_xyz_new,
                                       mesh.cloud.data.data() + i * mesh.cloud.point_step + off_xyz_old, 12) == 0;
  }
  CHECK(xyz_same);
  CHECK(worst > 0.9);
  // and the planner accepts the modifier's output cloud (normals first, 80-byte records)
  YAML::Node pc = planeSlicerConfig("PlaneSlicer", 1.05, Eigen::Vector3d(1, 0, 0), Eigen::Vector3d(0, 0, 0));
  CHECK(factory.createToolPathPlanner(pc)->plan(m).size() == 10);
}

/** NormalEstimationPCL by alias with the reference test's YAML (mesh_modifier_utest.cpp:26-30: radius 0.5, view point 0 0 10):
 *  every input field retained, a curvature field added, every normal a unit vector towards the view point or NaN (a radius
 *  search with fewer than three points), close to the file's own vertex normals on this smooth surface; the planner accepts
 *  the result where the normals are finite */
void testNormalEstimationPCL(const noether::Factory& factory, const pcl::PolygonMesh& mesh)
{
  std::printf("[MeshModifier NormalEstimationPCL]\n");
  YAML::Node config;
  config["name"] = "NormalEstimationPCL";
  config["radius"] = 0.5;
  config["vx"] = 0.0;
  config["vy"] = 0.0;
  config["vz"] = 10.0;
  std::shared_ptr<const noether::MeshModifier> mod = factory.createMeshModifier(config);
  const std::vector<pcl::PolygonMesh> out = mod->modify(mesh);
  CHECK(out.size() == 1);
  if (out.empty())
    return;
  const pcl::PolygonMesh& m = out.front();
  CHECK(m.polygons.size() == mesh.polygons.size());
  CHECK(m.cloud.width * m.cloud.height == mesh.cloud.width * mesh.cloud.height);
  for (const auto& f : mesh.cloud.fields)
  {
    bool found = false;
    for (const auto& g : m.cloud.fields)
      found = found || (g.name == f.name && g.datatype == f.datatype && g.count == f.count);
    CHECK(found);
  }
  unsigned off_new = ~0u, off_old = 0, off_curv = ~0u, off_xyz = 0;
  for (const auto& g : m.cloud.fields)
  {
    if (g.name == "normal_x")
      off_new = g.offset;
    if (g.name == "curvature")
      off_curv = g.offset;
    if (g.name == "x")
      off_xyz = g.offset;
  }
  for (const auto& g : mesh.cloud.fields)
    if (g.name == "normal_x")
      off_old = g.offset;
  CHECK(off_new == 0 && off_curv == 16);
  const std::size_t n = static_cast<std::size_t>(mesh.cloud.width) * mesh.cloud.height;
  std::size_t finite = 0, close = 0;
  double worst = 1.0;
  bool towards = true, curv_ok = true;
  for (std::size_t i = 0; i < n; ++i)
  {
    float a[3], b[3], p[3], c;
    std::memcpy(a, m.cloud.data.data() + i * m.cloud.point_step + off_new, 12);
    std::memcpy(&c, m.cloud.data.data() + i * m.cloud.point_step + off_curv, 4);
    std::memcpy(p, m.cloud.data.data() + i * m.cloud.point_step + off_xyz, 12);
    std::memcpy(b, mesh.cloud.data.data() + i * mesh.cloud.point_step + off_old, 12);
    if (std::isnan(a[0]))  // fewer than three neighbours (curvature NaN too), or neighbours on a line (eigenvector 0 / 0)
    {
      curv_ok = curv_ok && (std::isnan(c) || c >= 0.f);
      continue;
    }
    ++finite;
    const double na = std::sqrt(double(a[0]) * a[0] + double(a[1]) * a[1] + double(a[2]) * a[2]);
    CHECK(std::abs(na - 1.0) < 1e-4);
    towards = towards && (double(0 - p[0]) * a[0] + double(0 - p[1]) * a[1] + double(10 - p[2]) * a[2]) >= 0.0;
    curv_ok = curv_ok && (std::isnan(c) || (c >= 0.f && c <= 1.0f));
    const double nb = std::sqrt(double(b[0]) * b[0] + double(b[1]) * b[1] + double(b[2]) * b[2]);
    const double cs = (double(a[0]) * b[0] + double(a[1]) * b[1] + double(a[2]) * b[2]) / (na * nb);
    worst = std::min(worst, cs);
    close += cs > 0.9;
  }
  // (this mesh is coarse against the radius: a third of the searches return fewer than three points, and some of the rest
  //  three or four points nearly on a line, whose "plane" is arbitrary — in the reference as here)
  std::printf("  %zu of %zu normals finite, %zu within 26 degrees of the file's normals, smallest cosine %.3f\n", finite, n, close, worst);
  CHECK(finite > n / 2);
  CHECK(towards);
  CHECK(curv_ok);
  CHECK(close > finite * 8 / 10);
}

void testGeneratorsAndErrors(const noether::Factory& factory, const pcl::PolygonMesh& mesh)
{
  std::printf("[PrincipalAxis / Centroid / error behaviour]\n");
  YAML::Node dg;
  dg["name"] = "PrincipalAxis";
  dg["rotation_offset"] = 0.0;
  const Eigen::Vector3d d = factory.createDirectionGenerator(dg)->generate(mesh);
  CHECK(std::abs(norm(d) - 1.0) < 1e-6);
  CHECK(std::abs(d.z()) < 0.2);  // the second principal axis of this 10.5 x 10.5 x 4 surface lies near the xy plane
  {
    // PCARotated (plugins.cpp:91-107): a nested generator's direction rotated about the smallest principal axis
    YAML::Node rot;
    rot["name"] = "PCARotated";
    rot["direction_generator"] = dg;
    rot["rotation_offset"] = 0.0;
    const Eigen::Vector3d r0 = factory.createDirectionGenerator(rot)->generate(mesh);
    CHECK(std::abs(r0.x() - d.x()) < 1e-12 && std::abs(r0.y() - d.y()) < 1e-12 && std::abs(r0.z() - d.z()) < 1e-12);
    rot["rotation_offset"] = M_PI / 2;
    const Eigen::Vector3d r90 = factory.createDirectionGenerator(rot)->generate(mesh);
    CHECK(std::abs(dot(r90, d)) < 1e-9 && std::abs(norm(r90) - 1.0) < 1e-9);
  }
  YAML::Node og;
  og["name"] = "Centroid";
  const Eigen::Vector3d o = factory.createOriginGenerator(og)->generate(mesh);
  CHECK(std::abs(o.x() - 5.25) < 0.5 && std::abs(o.y() - 5.25) < 0.5);
  // PlaneSlicer with this library's own generators: one upload, PCA direction and centroid taken on the device
  YAML::Node config;
  config["name"] = "PlaneSlicer";
  config["line_spacing"] = 1.05;
  config["bidirectional"] = true;
  config["direction_generator"] = dg;
  config["origin_generator"] = og;
  const noether::ToolPaths tp = factory.createToolPathPlanner(config)->plan(mesh);
  CHECK(tp.size() >= 9 && tp.size() <= 16);

  // missing YAML key: the reference's message (serialization.h:13-14)
  YAML::Node bad = planeSlicerConfig("PlaneSlicer", 1.0, Eigen::Vector3d(1, 0, 0), Eigen::Vector3d(0, 0, 0));
  YAML::Node bad2;
  bad2["name"] = "PlaneSlicer";
  bad2["direction_generator"] = bad["direction_generator"];
  bad2["origin_generator"] = bad["origin_generator"];
  bad2["bidirectional"] = true;
  try
  {
    factory.createToolPathPlanner(bad2);
    CHECK(false);
  }
  catch (const std::exception& e)
  {
    CHECK(std::string(e.what()) == "Required member 'line_spacing' not found in YAML node");
  }
  // mesh without normals: the reference's message (plane_slicer_raster_planner.cpp:520-528)
  pcl::PolygonMesh stripped = mesh;
  for (auto& f : stripped.cloud.fields)
    if (f.name.rfind("normal_", 0) == 0)
      f.name = "was_" + f.name;
  try
  {
    factory.createToolPathPlanner(bad)->plan(stripped);
    CHECK(false);
  }
  catch (const std::exception& e)
  {
    CHECK(std::string(e.what()).find("does not have vertex normals") != std::string::npos);
  }
  // a quad face is refused rather than mis-sliced
  pcl::PolygonMesh quad = mesh;
  quad.polygons[0].vertices.push_back(quad.polygons[1].vertices[2]);
  try
  {
    factory.createToolPathPlanner(bad)->plan(quad);
    CHECK(false);
  }
  catch (const std::exception& e)
  {
    CHECK(std::string(e.what()).find("triangle") != std::string::npos);
  }
}
}  // namespace

/** wavy(nx, ny) of SURVEY.md §8(d) as a pcl::PolygonMesh (PointXYZ cloud, one pcl::Vertices per face) */
pcl::PolygonMesh wavyMesh(unsigned nx, unsigned ny)
{
  const double lx = 1.0, ly = 0.8, amp = 0.02, lam = 0.25;
  pcl::PolygonMesh mesh;
  pcl::PCLPointCloud2& c = mesh.cloud;
  c.width = (nx + 1) * (ny + 1);
  c.height = 1;
  c.point_step = 16;
  c.row_step = c.point_step * c.width;
  addField(c, "x", 0), addField(c, "y", 4), addField(c, "z", 8);
  c.data.assign(static_cast<std::size_t>(c.width) * 16, 0);
  for (unsigned r = 0; r <= ny; ++r)
    for (unsigned q = 0; q <= nx; ++q)
    {
      const double x = q * (lx / nx), y = r * (ly / ny);
      const float p[3] = { static_cast<float>(x), static_cast<float>(y),
                           static_cast<float>(amp * std::sin(2 * M_PI * x / lam) * std::sin(2 * M_PI * y / lam)) };
      std::memcpy(c.data.data() + 16 * (static_cast<std::size_t>(q) + static_cast<std::size_t>(nx + 1) * r), p, 12);
    }
  mesh.polygons.resize(2 * static_cast<std::size_t>(nx) * ny);
  std::size_t f = 0;
  for (int upper = 0; upper < 2; ++upper)
    for (unsigned r = 0; r < ny; ++r)
      for (unsigned q = 0; q < nx; ++q)
      {
        const pcl::index_t i = q + (nx + 1) * r;
        mesh.polygons[f++].vertices = upper ? std::vector<pcl::index_t>{ i, i + nx + 2, i + nx + 1 }
                                            : std::vector<pcl::index_t>{ i, i + 1, i + nx + 2 };
      }
  return mesh;
}

/** What a noether application sees: NormalsFromMeshFaces, then PlaneSlicer(PrincipalAxis, Centroid), both by alias,
 *  pcl::PolygonMesh in, nested noether::ToolPaths out.  Wall clock, second of two runs. */
bool samePoses(const noether::ToolPaths& a, const noether::ToolPaths& b)
{
  if (a.size() != b.size())
    return false;
  for (std::size_t p = 0; p < a.size(); ++p)
  {
    if (a[p].size() != b[p].size())
      return false;
    for (std::size_t s = 0; s < a[p].size(); ++s)
    {
      if (a[p][s].size() != b[p][s].size())
        return false;
      for (std::size_t w = 0; w < a[p][s].size(); ++w)
        if (std::memcmp(a[p][s][w].matrix().data(), b[p][s][w].matrix().data(), 16 * sizeof(double)) != 0)
          return false;
    }
  }
  return true;
}

std::size_t fieldOffset(const pcl::PCLPointCloud2& cloud, const char* name)
{
  for (const auto& f : cloud.fields)
    if (f.name == name)
      return f.offset;
  return 0;
}

/** The plugin library keeps the last mesh on the device and recognises it by content: a pipeline's components (and a
 *  cross-hatch pair of planners) share one upload.  What must never happen is a stale hit. */
void testResidentMeshReuse(const noether::Factory& factory, const pcl::PolygonMesh& raw)
{
  std::printf("[resident mesh reuse]\n");
  YAML::Node mod;
  mod["name"] = "NormalsFromMeshFaces";
  const pcl::PolygonMesh m1 = factory.createMeshModifier(mod)->modify(raw).front();
  YAML::Node dg, og, pc;
  dg["name"] = "PrincipalAxis";
  dg["rotation_offset"] = 0.0;
  og["name"] = "Centroid";
  pc["name"] = "PlaneSlicer";
  pc["line_spacing"] = 0.7;
  pc["bidirectional"] = true;
  pc["direction_generator"] = dg;
  pc["origin_generator"] = og;
  auto planner = factory.createToolPathPlanner(pc);
  // cross-hatch partner: the same slicer, its direction turned by 90 degrees about the smallest principal axis
  YAML::Node rot, pc2;
  rot["name"] = "PCARotated";
  rot["rotation_offset"] = M_PI / 2;
  rot["direction_generator"] = dg;
  pc2["name"] = "PlaneSlicer";
  pc2["line_spacing"] = 0.7;
  pc2["bidirectional"] = true;
  pc2["direction_generator"] = rot;
  pc2["origin_generator"] = og;
  auto crossed = factory.createToolPathPlanner(pc2);

  const noether::ToolPaths a = planner->plan(m1);     // hit: the modifier left exactly this mesh on the device
  const noether::ToolPaths ax = crossed->plan(m1);    // hit again (generator + planner)
  CHECK(!a.empty() && !ax.empty() && !samePoses(a, ax));

  // same sizes, one coordinate changed: must be seen as another mesh
  pcl::PolygonMesh m2 = m1;
  float z;
  std::uint8_t* pz = m2.cloud.data.data() + 100 * m2.cloud.point_step + fieldOffset(m2.cloud, "z");
  std::memcpy(&z, pz, 4);
  z += 0.25f;
  std::memcpy(pz, &z, 4);
  const noether::ToolPaths b = planner->plan(m2);
  CHECK(!samePoses(a, b));
  // same geometry, other normals: the planner must not reuse the resident normals
  pcl::PolygonMesh m3 = m1;
  const std::size_t on = fieldOffset(m3.cloud, "normal_x");
  for (std::size_t i = 0; i < static_cast<std::size_t>(m3.cloud.width) * m3.cloud.height; ++i)
    for (int k = 0; k < 3; ++k)
    {
      float v;
      std::uint8_t* q = m3.cloud.data.data() + i * m3.cloud.point_step + on + 4 * k;
      std::memcpy(&v, q, 4);
      v = -v;
      std::memcpy(q, &v, 4);
    }
  const noether::ToolPaths c = planner->plan(m2);  // (hit on m2)
  CHECK(samePoses(b, c));
  const noether::ToolPaths d = planner->plan(m3);
  CHECK(!samePoses(a, d));
  // one triangle re-indexed
  pcl::PolygonMesh m4 = m1;
  std::swap(m4.polygons[7].vertices[0], m4.polygons[7].vertices[1]);
  (void)planner->plan(m4);
  // and back to the first mesh after all of that: identical to the first answers
  CHECK(samePoses(a, planner->plan(m1)));
  CHECK(samePoses(ax, crossed->plan(m1)));
}

void timePipeline(const noether::Factory& factory, unsigned nx, unsigned ny, double spacing)
{
  const pcl::PolygonMesh raw = wavyMesh(nx, ny);
  YAML::Node mod;
  mod["name"] = "NormalsFromMeshFaces";
  YAML::Node dg, og, pc;
  dg["name"] = "PrincipalAxis";
  dg["rotation_offset"] = 0.0;
  og["name"] = "Centroid";
  pc["name"] = "PlaneSlicer";
  pc["line_spacing"] = spacing;
  pc["bidirectional"] = true;
  pc["direction_generator"] = dg;
  pc["origin_generator"] = og;
  auto modifier = factory.createMeshModifier(mod);
  auto planner = factory.createToolPathPlanner(pc);
  for (int run = 0; run < 2; ++run)
  {
    const auto t0 = std::chrono::steady_clock::now();
    const std::vector<pcl::PolygonMesh> meshes = modifier->modify(raw);
    const auto t1 = std::chrono::steady_clock::now();
    const noether::ToolPaths tp = planner->plan(meshes.front());
    const auto t2 = std::chrono::steady_clock::now();
    std::size_t wp = 0;
    for (const auto& path : tp)
      for (const auto& seg : path)
        wp += seg.size();
    if (run == 1)
      std::printf("[pipeline %zu triangles, spacing %g] modify %.1f ms, plan %.1f ms -> %zu tool paths, %zu waypoints "
                  "(nested Isometry3d vectors on the host)\n",
                  raw.polygons.size(), spacing, std::chrono::duration<double, std::milli>(t1 - t0).count(),
                  std::chrono::duration<double, std::milli>(t2 - t1).count(), tp.size(), wp);
    CHECK(tp.size() > 10 && wp > 1000);
    if (run == 1)
    {
      // the cross-hatch partner on the same mesh: direction generator and planner find the mesh resident
      YAML::Node rot, pc2;
      rot["name"] = "PCARotated";
      rot["rotation_offset"] = M_PI / 2;
      rot["direction_generator"] = dg;
      pc2["name"] = "PlaneSlicer";
      pc2["line_spacing"] = spacing;
      pc2["bidirectional"] = true;
      pc2["direction_generator"] = rot;
      pc2["origin_generator"] = og;
      auto crossed = factory.createToolPathPlanner(pc2);
      const auto t3 = std::chrono::steady_clock::now();
      const noether::ToolPaths tx = crossed->plan(meshes.front());
      const auto t4 = std::chrono::steady_clock::now();
      std::printf("  cross-hatch partner (PCARotated 90 deg) on the same mesh: plan %.1f ms -> %zu tool paths\n",
                  std::chrono::duration<double, std::milli>(t4 - t3).count(), tx.size());
      CHECK(tx.size() > 10);
    }
  }
}

/** bench.py's end-to-end leg (`e2e`): K plans through the reference-facing call — Factory::createToolPathPlanner("PlaneSlicer")
 *  ->plan(pcl::PolygonMesh) -> nested noether::ToolPaths (core/types.h:31-50) — on host meshes the device has not seen:
 *  two copies of the workload that differ in one coordinate alternate, so every plan flattens, fingerprints and uploads its
 *  mesh, plans, downloads and fills the Isometry3d vectors.  Wall clock per plan; one JSON line on stdout. */
int benchPlanner(const noether::Factory& factory, unsigned nx, unsigned ny, double spacing, int steps, int warmup)
{
  const pcl::PolygonMesh raw = wavyMesh(nx, ny);
  YAML::Node mod;
  mod["name"] = "NormalsFromMeshFaces";
  YAML::Node dg, og, pc;
  dg["name"] = "PrincipalAxis";
  dg["rotation_offset"] = 0.0;
  og["name"] = "Centroid";
  pc["name"] = "PlaneSlicer";
  pc["line_spacing"] = spacing;
  pc["bidirectional"] = true;
  pc["direction_generator"] = dg;
  pc["origin_generator"] = og;
  auto modifier = factory.createMeshModifier(mod);
  auto planner = factory.createToolPathPlanner(pc);
  std::vector<pcl::PolygonMesh> mesh(2);
  mesh[0] = modifier->modify(raw).front();
  mesh[1] = mesh[0];
  {
    // the partner: the first vertex moved by one ulp along z (same counts, another content fingerprint)
    float z;
    std::uint8_t* p = mesh[1].cloud.data.data() + 8;
    std::memcpy(&z, p, 4);
    z = std::nextafter(z, 1.0f);
    std::memcpy(p, &z, 4);
  }
  std::vector<double> ms;
  std::size_t wp = 0, paths = 0, segs = 0;
  for (int i = 0; i < warmup + steps; ++i)
  {
    const auto t0 = std::chrono::steady_clock::now();
    {
      const noether::ToolPaths tp = planner->plan(mesh[i & 1]);
      const auto t1 = std::chrono::steady_clock::now();
      if (i >= warmup)
        ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
      paths = tp.size();
      wp = segs = 0;
      for (const auto& path : tp)
      {
        segs += path.size();
        for (const auto& seg : path)
          wp += seg.size();
      }
    }  // (the tool paths are released outside the timed interval, as a caller that keeps them would)
  }
  double sum = 0, best = 1e300;
  for (double v : ms)
    sum += v, best = std::min(best, v);
  const std::size_t V = static_cast<std::size_t>(mesh[0].cloud.width) * mesh[0].cloud.height, F = mesh[0].polygons.size();
  std::printf("{\"plugin_e2e\": true, \"triangles\": %zu, \"vertices\": %zu, \"point_step\": %u, \"steps\": %d, \"warmup\": %d, "
              "\"ms_per_plan\": %.4f, \"ms_best\": %.4f, \"tool_paths\": %zu, \"segments\": %zu, \"waypoints\": %zu, "
              "\"host_in_bytes\": %zu, \"host_out_bytes\": %zu}\n",
              F, V, mesh[0].cloud.point_step, steps, warmup, sum / std::max<std::size_t>(1, ms.size()), best, paths, segs, wp,
              V * mesh[0].cloud.point_step + 12 * F, wp * sizeof(Eigen::Isometry3d));
  return wp > 0 ? 0 : 1;
}

// ---- ToolPathModifierTestFixture / OneTimeToolPathModifierTestFixture (tool_path_modifier_utest.cpp:14-96,120-214,261-304)
//      for the modifiers this library runs on the device, created by alias from the reference's test YAML
noether::ToolPaths gridToolPaths(unsigned p, unsigned s, unsigned w, bool identity)
{
  noether::ToolPaths paths(p);
  for (unsigned pi = 0; pi < p; ++pi)
  {
    paths[pi].resize(s);
    for (unsigned si = 0; si < s; ++si)
    {
      paths[pi][si].resize(w);
      for (unsigned wi = 0; wi < w; ++wi)
      {
        double* T = paths[pi][si][wi].matrix().data();
        for (int k = 0; k < 16; ++k)
          T[k] = (k % 5 == 0) ? 1.0 : 0.0;
        if (!identity)
        {
          T[12] = static_cast<double>(si * w + wi);  // createRasterGridToolPath :138-170
          T[13] = static_cast<double>(pi);
        }
      }
    }
  }
  return paths;
}

bool sameToolPaths(const noether::ToolPaths& a, const noether::ToolPaths& b, double tol)
{
  if (a.size() != b.size())
    return false;
  for (std::size_t i = 0; i < a.size(); ++i)
  {
    if (a[i].size() != b[i].size())
      return false;
    for (std::size_t j = 0; j < a[i].size(); ++j)
    {
      if (a[i][j].size() != b[i][j].size())
        return false;
      for (std::size_t k = 0; k < a[i][j].size(); ++k)
        for (int e = 0; e < 16; ++e)
          if (!(std::fabs(a[i][j][k].matrix().data()[e] - b[i][j][k].matrix().data()[e]) <= tol))
            return false;
    }
  }
  return true;
}

void testToolPathModifiers(const noether::Factory& factory)
{
  std::printf("[ToolPathModifier plugins on the device]\n");
  YAML::Node offset = vectorNode(0.1, 0.2, 0.3);
  YAML::Node iso = vectorNode(0.0, 0.0, 0.0);
  iso["qw"] = 1.0, iso["qx"] = 0.0, iso["qy"] = 0.0, iso["qz"] = 0.0;
  std::vector<std::pair<YAML::Node, bool>> configs;  // (config, one-time modifier)
  auto entry = [&](const char* name, bool one_time) -> YAML::Node& {
    YAML::Node n;
    n["name"] = name;
    configs.emplace_back(n, one_time);
    return configs.back().first;
  };
  entry("SnakeOrganization", false);
  {
    YAML::Node& n = entry("LinearApproach", false);
    n["offset"] = offset, n["n_points"] = 5;
  }
  {
    YAML::Node& n = entry("LinearDeparture", false);
    n["offset"] = offset, n["n_points"] = 5;
  }
  entry("Offset", false)["offset"] = iso;
  entry("FixedOrientation", true)["ref_x_dir"] = vectorNode(1.0, 0.0, 0.0);
  entry("UniformOrientation", true);
  entry("Concatenate", true);
  entry("DirectionOfTravelOrientation", true);
  entry("MovingAverageOrientationSmoothing", false)["window_size"] = 3;
  {
    YAML::Node& n = entry("ToolDragOrientation", false);
    n["angle_offset"] = 0.1745, n["tool_radius"] = 0.025;
  }
  {
    YAML::Node& n = entry("BiasedToolDragOrientation", false);
    n["angle_offset"] = 0.1745, n["tool_radius"] = 0.025;
  }
  for (const char* lead : { "CircularLeadIn", "CircularLeadOut" })
  {
    YAML::Node& n = entry(lead, false);
    n["arc_angle"] = 1.571, n["arc_radius"] = 0.1, n["n_points"] = 5;
  }
  {
    YAML::Node& n = entry("CompoundToolPathModifier", false);
    YAML::Node list;
    list.push_back(configs[1].first);
    list.push_back(configs[2].first);
    n["modifiers"] = list;
  }
  const std::vector<noether::ToolPaths> inputs = { gridToolPaths(4, 2, 10, true), gridToolPaths(4, 2, 10, false) };
  for (const auto& c : configs)
  {
    const std::string name = c.first["name"].as<std::string>();
    auto modifier = factory.createToolPathModifier(c.first);
    CHECK(modifier != nullptr);
    for (const noether::ToolPaths& input : inputs)
    {
      const noether::ToolPaths once = modifier->modify(input);
      if (c.second)  // OneTimeToolPathModifierTestFixture: applying it again changes nothing
        CHECK(sameToolPaths(once, modifier->modify(once), 1e-12));
      if (name == "Concatenate")
        CHECK(once.size() == 1 && once[0].size() == 8);
      else
        CHECK(once.size() == 4 && once[0].size() == 2);
      const std::size_t want = name == "CompoundToolPathModifier" ? 20 : (name.rfind("Linear", 0) == 0 || name.rfind("Circular", 0) == 0 ? 15 : 10);
      CHECK(once[0][0].size() == want);
    }
    std::printf("  %s ok\n", name.c_str());
  }
  // closed forms on the raster grid: snake reverses the odd tool paths; approach points lead up to the first waypoint
  {
    const noether::ToolPaths snake = factory.createToolPathModifier(configs[0].first)->modify(inputs[1]);
    CHECK(snake[0][0][0].translation().x() == 0.0 && snake[1][0][0].translation().x() == 19.0 &&
          snake[1][1][9].translation().x() == 0.0 && snake[2][1][9].translation().x() == 19.0);
    const noether::ToolPaths lead = factory.createToolPathModifier(configs[1].first)->modify(inputs[1]);
    CHECK(std::fabs(lead[0][0][0].translation().z() - 0.3) < 1e-15 && std::fabs(lead[0][0][4].translation().z() - 0.06) < 1e-15 &&
          lead[0][0][5].translation().z() == 0.0 && lead[0][1][5].translation().x() == 10.0);
  }
  // error behaviour: a std::exception where the reference throws / reads out of range
  {
    noether::ToolPaths single(1);
    single[0].resize(1);
    single[0][0] = gridToolPaths(1, 1, 1, true)[0][0];
    bool threw = false;
    try
    {
      factory.createToolPathModifier(configs[9].first)->modify(single);
    }
    catch (const std::exception& e)
    {
      threw = std::string(e.what()).find("Insufficient number of points") != std::string::npos;
    }
    CHECK(threw);
  }
}

/** FaceMidpointSubdivision / FaceSubdivisionByArea through the plugin boundary, with the YAML of the reference's own test
 *  (mesh_modifier_utest.cpp:20-23: n_iterations 2, max_area 0.01): fields retained (:104-113), 4^n faces, one new vertex
 *  per edge and iteration, every kept triangle within the area limit, surface area unchanged, and a plan on the output. */
void testFaceSubdivision(const noether::Factory& factory, const pcl::PolygonMesh& mesh)
{
  std::printf("[MeshModifier FaceMidpointSubdivision / FaceSubdivisionByArea]\n");
  unsigned off_xyz = 0;
  for (const auto& g : mesh.cloud.fields)
    if (g.name == "x")
      off_xyz = g.offset;
  auto point = [&](const pcl::PolygonMesh& m, std::uint32_t i) {
    float p[3];
    std::memcpy(p, m.cloud.data.data() + static_cast<std::size_t>(i) * m.cloud.point_step + off_xyz, 12);
    return Eigen::Vector3d(p[0], p[1], p[2]);
  };
  auto faceArea = [&](const pcl::PolygonMesh& m, const pcl::Vertices& f) {
    const Eigen::Vector3d a = point(m, f.vertices[0]), b = point(m, f.vertices[1]), c = point(m, f.vertices[2]);
    const Eigen::Vector3d u(b.x() - a.x(), b.y() - a.y(), b.z() - a.z()), v(c.x() - a.x(), c.y() - a.y(), c.z() - a.z());
    const Eigen::Vector3d n(u.y() * v.z() - u.z() * v.y(), u.z() * v.x() - u.x() * v.z(), u.x() * v.y() - u.y() * v.x());
    return 0.5 * norm(n);
  };
  auto totalArea = [&](const pcl::PolygonMesh& m) {
    double s = 0;
    for (const auto& f : m.polygons)
      s += faceArea(m, f);
    return s;
  };
  auto checkCommon = [&](const pcl::PolygonMesh& m) {
    CHECK(m.cloud.height == 1 && m.cloud.point_step == mesh.cloud.point_step);
    CHECK(m.cloud.data.size() == static_cast<std::size_t>(m.cloud.point_step) * m.cloud.width);
    CHECK(m.cloud.row_step == m.cloud.width * m.cloud.point_step);
    CHECK(m.cloud.fields.size() == mesh.cloud.fields.size());
    for (const auto& f : mesh.cloud.fields)
    {
      bool found = false;
      for (const auto& g : m.cloud.fields)
        found = found || (g.name == f.name && g.datatype == f.datatype && g.count == f.count && g.offset == f.offset);
      CHECK(found);
    }
    // the input's records come first, untouched
    CHECK(std::memcmp(m.cloud.data.data(), mesh.cloud.data.data(), mesh.cloud.data.size()) == 0);
    std::uint32_t top = 0;
    for (const auto& f : m.polygons)
    {
      CHECK(f.vertices.size() == 3);
      for (auto v : f.vertices)
        top = std::max<std::uint32_t>(top, v);
    }
    CHECK(top + 1 == m.cloud.width);  // every new vertex is referenced
    CHECK(std::abs(totalArea(m) - totalArea(mesh)) < 1e-3 * totalArea(mesh));
  };
  std::size_t edges = 0;
  {
    std::vector<std::pair<std::uint32_t, std::uint32_t>> e;
    for (const auto& f : mesh.polygons)
      for (int k = 0; k < 3; ++k)
        e.emplace_back(std::min<std::uint32_t>(f.vertices[k], f.vertices[(k + 1) % 3]), std::max<std::uint32_t>(f.vertices[k], f.vertices[(k + 1) % 3]));
    std::sort(e.begin(), e.end());
    edges = std::unique(e.begin(), e.end()) - e.begin();
  }
  {
    YAML::Node config;
    config["name"] = "FaceMidpointSubdivision";
    config["n_iterations"] = 2;
    const std::vector<pcl::PolygonMesh> out = factory.createMeshModifier(config)->modify(mesh);
    CHECK(out.size() == 1);
    const pcl::PolygonMesh& m = out.front();
    checkCommon(m);
    CHECK(m.polygons.size() == 16 * mesh.polygons.size());
    // iteration 1 adds one vertex per edge; the refined mesh has 2 E + 3 F edges
    CHECK(m.cloud.width == mesh.cloud.width + edges + (2 * edges + 3 * mesh.polygons.size()));
    YAML::Node pc = planeSlicerConfig("PlaneSlicer", 1.05, Eigen::Vector3d(1, 0, 0), Eigen::Vector3d(0, 0, 0));
    CHECK(factory.createToolPathPlanner(pc)->plan(m).size() == 10);
  }
  {
    YAML::Node config;
    config["name"] = "FaceSubdivisionByArea";
    config["max_area"] = 0.01;
    const std::vector<pcl::PolygonMesh> out = factory.createMeshModifier(config)->modify(mesh);
    CHECK(out.size() == 1);
    const pcl::PolygonMesh& m = out.front();
    checkCommon(m);
    CHECK(m.polygons.size() > mesh.polygons.size());
    CHECK((m.polygons.size() - mesh.polygons.size()) == 2 * (m.cloud.width - mesh.cloud.width));  // a split adds 1 vertex, 2 faces
    bool within = true;
    for (const auto& f : m.polygons)
      within = within && faceArea(m, f) <= 0.01 * (1 + 1e-5);
    CHECK(within);
  }
  {
    // The pipeline [NormalsFromMeshFaces -> FaceMidpointSubdivision]: the first modifier leaves the INPUT's records on the
    // device and returns a concatenated cloud with another record layout (32-byte pcl::Normal + the input's fields); the
    // subdivision must work on that returned cloud, not on what happens to be resident (same geometry, other records).
    YAML::Node nc;
    nc["name"] = "NormalsFromMeshFaces";
    const std::vector<pcl::PolygonMesh> with_normals = factory.createMeshModifier(nc)->modify(mesh);
    CHECK(with_normals.size() == 1);
    const pcl::PolygonMesh& in = with_normals.front();
    CHECK(in.cloud.point_step == mesh.cloud.point_step + 32);
    YAML::Node config;
    config["name"] = "FaceMidpointSubdivision";
    config["n_iterations"] = 1;
    const std::vector<pcl::PolygonMesh> out = factory.createMeshModifier(config)->modify(in);
    CHECK(out.size() == 1);
    const pcl::PolygonMesh& m = out.front();
    CHECK(m.cloud.point_step == in.cloud.point_step && m.cloud.height == 1);
    CHECK(m.cloud.data.size() == static_cast<std::size_t>(m.cloud.point_step) * m.cloud.width);
    CHECK(m.cloud.width == in.cloud.width + edges);
    CHECK(m.polygons.size() == 4 * in.polygons.size());
    CHECK(m.cloud.data.size() >= in.cloud.data.size() &&
          std::memcmp(m.cloud.data.data(), in.cloud.data.data(), in.cloud.data.size()) == 0);
    // and the other way round: the same geometry in the ORIGINAL layout right after (the device now holds the subdivided
    // 80-byte records)
    const std::vector<pcl::PolygonMesh> again = factory.createMeshModifier(config)->modify(mesh);
    CHECK(again.size() == 1 && again.front().cloud.point_step == mesh.cloud.point_step);
    CHECK(again.front().cloud.data.size() == static_cast<std::size_t>(mesh.cloud.point_step) * (mesh.cloud.width + edges));
    CHECK(std::memcmp(again.front().cloud.data.data(), mesh.cloud.data.data(), mesh.cloud.data.size()) == 0);
  }
}

/** the surface twice, the copy shifted by 100 in z and without its last 200 faces */
pcl::PolygonMesh twoSurfaces(const pcl::PolygonMesh& mesh)
{
  unsigned off_z = 0;
  for (const auto& g : mesh.cloud.fields)
    if (g.name == "z")
      off_z = g.offset;
  pcl::PolygonMesh two = mesh;
  const std::uint32_t V = mesh.cloud.width * mesh.cloud.height;
  two.cloud.data.insert(two.cloud.data.end(), mesh.cloud.data.begin(), mesh.cloud.data.end());
  two.cloud.width = 2 * V;
  two.cloud.height = 1;
  two.cloud.row_step = two.cloud.width * two.cloud.point_step;
  for (std::uint32_t v = V; v < 2 * V; ++v)
  {
    float z;
    std::memcpy(&z, two.cloud.data.data() + static_cast<std::size_t>(v) * two.cloud.point_step + off_z, 4);
    z += 100.f;
    std::memcpy(two.cloud.data.data() + static_cast<std::size_t>(v) * two.cloud.point_step + off_z, &z, 4);
  }
  for (std::size_t f = 0; f + 200 < mesh.polygons.size(); ++f)
  {
    pcl::Vertices p = mesh.polygons[f];
    for (auto& v : p.vertices)
      v += V;
    two.polygons.push_back(p);
  }
  return two;
}

void checkTwoSurfaces(const std::vector<pcl::PolygonMesh>& out, const pcl::PolygonMesh& mesh, std::uint32_t V)
{
  CHECK(out.size() == 2);
  if (out.size() == 2)
  {
    // equal cluster sizes: whichever order the size sort leaves, one mesh has all faces and the other 200 fewer
    const std::size_t a = out[0].polygons.size(), b = out[1].polygons.size();
    CHECK(std::max(a, b) == mesh.polygons.size() && std::min(a, b) == mesh.polygons.size() - 200);
    for (const auto& m : out)
    {
      CHECK(m.cloud.height == 1 && m.cloud.width <= V && m.cloud.width > 0);
      CHECK(m.cloud.data.size() == static_cast<std::size_t>(m.cloud.width) * m.cloud.point_step);
      CHECK(m.cloud.fields.size() == mesh.cloud.fields.size());
      std::uint32_t top = 0;
      for (const auto& f : m.polygons)
        for (auto v : f.vertices)
          top = std::max<std::uint32_t>(top, v);
      CHECK(top + 1 == m.cloud.width);
    }
  }
}

/** The host side of EuclideanClustering (size filter, PCL's size sort, sub-mesh cutting) through the plugin library's test
 *  hook, with the labels the device would return for two surfaces far apart: no GPU needed. */
void testClusterMeshesHost(void* lib, const pcl::PolygonMesh& mesh)
{
  std::printf("[EuclideanClustering, host side]\n");
  using Hook = void (*)(const pcl::PolygonMesh*, const std::uint32_t*, int, int, std::vector<pcl::PolygonMesh>*);
  auto hook = reinterpret_cast<Hook>(dlsym(lib, "noether_b200_cut_cluster_meshes"));
  CHECK(hook != nullptr);
  if (!hook)
    return;
  const std::uint32_t V = mesh.cloud.width * mesh.cloud.height;
  const pcl::PolygonMesh two = twoSurfaces(mesh);
  std::vector<std::uint32_t> labels(2 * V, 0);
  std::fill(labels.begin() + V, labels.end(), V);
  std::vector<pcl::PolygonMesh> out;
  hook(&two, labels.data(), 1, -1, &out);
  checkTwoSurfaces(out, mesh, V);
  // one cluster that uses every point: the input itself
  std::vector<std::uint32_t> zeros(V, 0);
  hook(&mesh, zeros.data(), 1, -1, &out);
  CHECK(out.size() == 1 && out[0].polygons.size() == mesh.polygons.size() && out[0].cloud.data == mesh.cloud.data &&
        out[0].cloud.width == mesh.cloud.width && out[0].cloud.height == mesh.cloud.height);
  // every point its own cluster: V default-constructed meshes; with min_cluster_size 2 nothing is left
  std::vector<std::uint32_t> iota(V);
  for (std::uint32_t v = 0; v < V; ++v)
    iota[v] = v;
  hook(&mesh, iota.data(), 1, -1, &out);
  CHECK(out.size() == V);
  bool all_empty = true;
  for (const auto& m : out)
    all_empty = all_empty && m.polygons.empty() && m.cloud.data.empty() && m.cloud.fields.empty();
  CHECK(all_empty);
  hook(&mesh, iota.data(), 2, -1, &out);
  CHECK(out.empty());
  // the size filter and the size sort: the copy's cluster made one point larger by moving a point of the original over
  labels[0] = V;   // point 0 now belongs to the second cluster (size V + 1 against V - 1): it comes first
  hook(&two, labels.data(), 1, -1, &out);
  CHECK(out.size() == 2 && out[0].polygons.size() == mesh.polygons.size() - 200);
  hook(&two, labels.data(), static_cast<int>(V), -1, &out);
  CHECK(out.size() == 1);
  hook(&two, labels.data(), 1, static_cast<int>(V), &out);
  CHECK(out.size() == 1 && out[0].polygons.size() < mesh.polygons.size());
}

/** EuclideanClustering through the plugin boundary with the YAML of the reference's own test (mesh_modifier_utest.cpp:15-19:
 *  tolerance 1.0, min 1, max -1): non-empty meshes with polygons and the input's fields (:130-183) — on this connected
 *  surface one cluster that uses every point, i.e. the input itself; and two copies of the surface far apart come back as
 *  two meshes, largest first, each renumbered from zero. */
void testEuclideanClustering(const noether::Factory& factory, const pcl::PolygonMesh& mesh)
{
  std::printf("[MeshModifier EuclideanClustering]\n");
  YAML::Node config;
  config["name"] = "EuclideanClustering";
  config["tolerance"] = 1.0;
  config["min_cluster_size"] = 1;
  config["max_cluster_size"] = -1;
  auto modifier = factory.createMeshModifier(config);
  {
    const std::vector<pcl::PolygonMesh> out = modifier->modify(mesh);
    CHECK(out.size() == 1);
    if (!out.empty())
    {
      CHECK(out[0].polygons.size() == mesh.polygons.size());
      CHECK(out[0].cloud.width * out[0].cloud.height == mesh.cloud.width * mesh.cloud.height);
      CHECK(out[0].cloud.fields.size() == mesh.cloud.fields.size());
      CHECK(out[0].cloud.data == mesh.cloud.data);
    }
  }
  const pcl::PolygonMesh two = twoSurfaces(mesh);
  const std::uint32_t V = mesh.cloud.width * mesh.cloud.height;
  const std::vector<pcl::PolygonMesh> out = modifier->modify(two);
  checkTwoSurfaces(out, mesh, V);
}

int main(int argc, char** argv)
{
  if (argc < 3)
  {
    std::printf("usage: %s <plugin library> <wavy_mesh_with_hole.ply>\n", argv[0]);
    return 2;
  }
  void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib)
  {
    std::printf("dlopen failed: %s\n", dlerror());
    return 2;
  }
  auto loader = std::make_shared<boost_plugin_loader::PluginLoader>();
  loader->libraries = { lib, dlopen(nullptr, RTLD_NOW) };  // our plugin library first, then this executable's generators
  const noether::Factory factory(loader);
  try
  {
    for (const char* alias : { "PlaneSlicer", "PlaneSlicerLegacy", "PrincipalAxis", "PCARotated", "Centroid", "NormalsFromMeshFaces",
                                "FaceMidpointSubdivision", "FaceSubdivisionByArea", "EuclideanClustering", "NormalEstimationPCL" })
      CHECK(dlsym(lib, alias) != nullptr);
    if (argc > 8 && std::string(argv[3]) == "--bench")  // --bench nx ny spacing steps warmup
      return benchPlanner(factory, static_cast<unsigned>(std::atoi(argv[4])), static_cast<unsigned>(std::atoi(argv[5])),
                          std::atof(argv[6]), std::atoi(argv[7]), std::atoi(argv[8]));
    const pcl::PolygonMesh ply = loadPly(argv[2]);
    CHECK(ply.cloud.width == 757 && ply.polygons.size() == 1347);
    if (argc > 3 && std::string(argv[3]) == "--cpu-only")
    {
      testClusterMeshesHost(lib, ply);
      std::printf(g_failures ? "%d check(s) FAILED\n" : "all checks passed\n", g_failures);
      return g_failures ? 1 : 0;
    }
    testFlatSquareMesh(factory, "PlaneSlicer");
    testSemiPlanarMeshFile(factory, ply, "PlaneSlicer", false);
    testSemiPlanarMeshFile(factory, ply, "PlaneSlicerLegacy", true);
    testNormalsFieldRetention(factory, ply);
    testNormalEstimationPCL(factory, ply);
    testFaceSubdivision(factory, ply);
    testGeneratorsAndErrors(factory, ply);
    testResidentMeshReuse(factory, ply);
    testToolPathModifiers(factory);
    if (argc > 3 && std::string(argv[3]) == "--clustering")
      testEuclideanClustering(factory, ply);
    if (argc > 3 && std::string(argv[3]) == "--time")
    {
      timePipeline(factory, 800, 625, 0.005);    // cfg2: 1 M triangles
      timePipeline(factory, 2500, 2000, 0.001);  // cfg3: 10 M triangles
    }
  }
  catch (const std::exception& e)
  {
    std::printf("  FAILED with exception: %s\n", e.what());
    ++g_failures;
  }
  std::printf(g_failures ? "%d check(s) FAILED\n" : "all checks passed\n", g_failures);
  return g_failures ? 1 : 0;
}
