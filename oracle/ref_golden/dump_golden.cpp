// Golden-vector dumper for the hot path — TEST INFRASTRUCTURE, built and run ONLY inside a real noether workspace
// (the reference's own docker image: oracle/ref_golden/generate.sh).  It cannot be compiled in the development image of
// this repository (no VTK / PCL / Eigen / yaml-cpp / boost_plugin_loader there), which is why the oracle's parity is
// "unpinned" until the files this program writes exist under tests/golden/ref/.
//
// It runs the UNMODIFIED reference through its public plugin API (noether::Factory, the calls its own tests make:
// noether_tpp/test/tool_path_planner_utest.cpp:51-58,74-177, mesh_modifier_utest.cpp:154-183) on a handful of small
// cases and writes everything a parity test needs, inputs included, so that nothing has to be regenerated elsewhere:
//
//   <case>.bin :=  "NB2GOLD1"  { name_len:u32 name dtype:u32 ndim:u32 dims:u64[ndim] data }*
//                  dtype 0 = float64, 1 = float32, 2 = uint32
//     xyz            float32 [V,3]   vertex positions of the input mesh (pcl::PointXYZ view of mesh.cloud)
//     normals        float32 [V,3]   vertex normals of the mesh handed to the planner (input's, or NormalsFromMeshFaces')
//     triangles      uint32  [F,3]
//     direction      float64 [3]     DirectionGenerator::generate(mesh)        (raster_planner.h:43)
//     origin         float64 [3]     OriginGenerator::generate(mesh)           (raster_planner.h:56)
//     path_sizes     uint32  [P]     segments per tool path                    (ToolPaths, core/types.h:31-50)
//     segment_sizes  uint32  [S]     waypoints per segment
//     poses          float64 [W,16]  Eigen::Isometry3d::matrix(), column-major, in ToolPaths order
//     params         float64 [8]     line_spacing, bidirectional, legacy, point_spacing, min_segment_size, min_hole_size, 0, 0
//
// Usage: dump_golden <noether_tpp/test/meshes/wavy_mesh_with_hole.ply> <output directory>
#include <noether_tpp/plugin_interface.h>
#include <noether_tpp/serialization.h>
#include <noether_tpp/utils.h>

#include <boost_plugin_loader/plugin_loader.h>
#include <pcl/conversions.h>
#include <pcl/io/vtk_lib_io.h>
#include <pcl/point_types.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <random>
#include <string>
#include <vector>

using namespace noether;

namespace
{
struct Writer
{
  std::ofstream f;
  explicit Writer(const std::string& path) : f(path, std::ios::binary) { f.write("NB2GOLD1", 8); }
  template <typename T>
  void array(const std::string& name, std::uint32_t dtype, const std::vector<std::uint64_t>& dims, const T* data)
  {
    const std::uint32_t n = static_cast<std::uint32_t>(name.size()), nd = static_cast<std::uint32_t>(dims.size());
    f.write(reinterpret_cast<const char*>(&n), 4);
    f.write(name.data(), n);
    f.write(reinterpret_cast<const char*>(&dtype), 4);
    f.write(reinterpret_cast<const char*>(&nd), 4);
    std::uint64_t count = 1;
    for (std::uint64_t d : dims)
    {
      f.write(reinterpret_cast<const char*>(&d), 8);
      count *= d;
    }
    f.write(reinterpret_cast<const char*>(data), static_cast<std::streamsize>(count * sizeof(T)));
  }
};

void dumpMesh(Writer& w, const pcl::PolygonMesh& mesh)
{
  const std::size_t V = static_cast<std::size_t>(mesh.cloud.width) * mesh.cloud.height;
  std::vector<float> xyz(3 * V), nrm;
  for (std::size_t i = 0; i < V; ++i)
  {
    const Eigen::Vector3f p = getPoint(mesh.cloud, static_cast<std::uint32_t>(i));  // utils.cpp:74-95
    xyz[3 * i] = p.x(), xyz[3 * i + 1] = p.y(), xyz[3 * i + 2] = p.z();
  }
  w.array("xyz", 1, { V, 3 }, xyz.data());
  if (hasNormals(mesh.cloud))  // utils.cpp:65-72
  {
    nrm.resize(3 * V);
    for (std::size_t i = 0; i < V; ++i)
    {
      const Eigen::Vector3f n = getNormal(mesh.cloud, static_cast<std::uint32_t>(i));  // utils.cpp:97-117
      nrm[3 * i] = n.x(), nrm[3 * i + 1] = n.y(), nrm[3 * i + 2] = n.z();
    }
    w.array("normals", 1, { V, 3 }, nrm.data());
  }
  std::vector<std::uint32_t> tri;
  for (const pcl::Vertices& p : mesh.polygons)
    for (auto v : p.vertices)
      tri.push_back(static_cast<std::uint32_t>(v));
  w.array("triangles", 2, { mesh.polygons.size(), 3 }, tri.data());
}

void dumpPaths(Writer& w, const ToolPaths& tps)
{
  std::vector<std::uint32_t> path_sizes, segment_sizes;
  std::vector<double> poses;
  for (const ToolPath& path : tps)
  {
    path_sizes.push_back(static_cast<std::uint32_t>(path.size()));
    for (const ToolPathSegment& seg : path)
    {
      segment_sizes.push_back(static_cast<std::uint32_t>(seg.size()));
      for (const Eigen::Isometry3d& T : seg)
        poses.insert(poses.end(), T.matrix().data(), T.matrix().data() + 16);
    }
  }
  w.array("path_sizes", 2, { path_sizes.size() }, path_sizes.data());
  w.array("segment_sizes", 2, { segment_sizes.size() }, segment_sizes.data());
  w.array("poses", 0, { poses.size() / 16, 16 }, poses.data());
}

YAML::Node vec(const Eigen::Vector3d& v)
{
  YAML::Node n;
  n["x"] = v.x(), n["y"] = v.y(), n["z"] = v.z();
  return n;
}

struct Case
{
  std::string name;
  pcl::PolygonMesh mesh;
  YAML::Node dir_gen, origin_gen;
  double line_spacing = 0.1;
  bool bidirectional = true;
  bool legacy = false;
  double point_spacing = 0.0, min_segment_size = 0.0, min_hole_size = 0.0;
  bool compute_normals = false;  // run NormalsFromMeshFaces on the mesh first
};

void run(const Factory& factory, Case c, const std::string& out_dir)
{
  Writer w(out_dir + "/" + c.name + ".bin");
  if (c.compute_normals)
  {
    YAML::Node mc;
    mc["name"] = "NormalsFromMeshFaces";
    c.mesh = factory.createMeshModifier(mc)->modify(c.mesh).at(0);  // normals_from_mesh_faces_modifier.cpp:15-46
  }
  dumpMesh(w, c.mesh);
  const Eigen::Vector3d d = factory.createDirectionGenerator(c.dir_gen)->generate(c.mesh);
  const Eigen::Vector3d o = factory.createOriginGenerator(c.origin_gen)->generate(c.mesh);
  w.array("direction", 0, { 3 }, d.data());
  w.array("origin", 0, { 3 }, o.data());
  YAML::Node pc;
  pc["name"] = c.legacy ? "PlaneSlicerLegacy" : "PlaneSlicer";
  pc["direction_generator"] = c.dir_gen;
  pc["origin_generator"] = c.origin_gen;
  pc["line_spacing"] = c.line_spacing;
  pc["bidirectional"] = c.bidirectional;
  pc["point_spacing"] = c.point_spacing;
  pc["min_segment_size"] = c.min_segment_size;
  pc["min_hole_size"] = c.min_hole_size;
  const ToolPaths tps = factory.createToolPathPlanner(pc)->plan(c.mesh);  // raster_planner.cpp:16-35
  dumpPaths(w, tps);
  const double params[8] = { c.line_spacing, c.bidirectional ? 1.0 : 0.0, c.legacy ? 1.0 : 0.0, c.point_spacing,
                             c.min_segment_size, c.min_hole_size, 0.0, 0.0 };
  w.array("params", 0, { 8 }, params);
  std::printf("%s: %zu tool paths\n", c.name.c_str(), tps.size());
}

/** SURVEY.md §8(d) wavy(nx, ny): (nx+1)(ny+1) vertices row-major, z = A sin(2 pi x / l) sin(2 pi y / l) in double -> float;
 *  all lower triangles (i, i+1, i+nx+2) first, then all upper (i, i+nx+2, i+nx+1) */
pcl::PolygonMesh wavy(int nx, int ny, double lx = 1.0, double ly = 0.8, double amp = 0.02, double lambda = 0.25)
{
  pcl::PointCloud<pcl::PointXYZ> cloud;
  for (int r = 0; r <= ny; ++r)
    for (int c = 0; c <= nx; ++c)
    {
      const double x = c * lx / nx, y = r * ly / ny;
      cloud.emplace_back(static_cast<float>(x), static_cast<float>(y),
                         static_cast<float>(amp * std::sin(2 * M_PI * x / lambda) * std::sin(2 * M_PI * y / lambda)));
    }
  pcl::PolygonMesh mesh;
  pcl::toPCLPointCloud2(cloud, mesh.cloud);
  for (int pass = 0; pass < 2; ++pass)
    for (int r = 0; r < ny; ++r)
      for (int c = 0; c < nx; ++c)
      {
        const std::uint32_t i = static_cast<std::uint32_t>(c + (nx + 1) * r);
        pcl::Vertices t;
        if (pass == 0)
          t.vertices = { i, i + 1, i + static_cast<std::uint32_t>(nx) + 2 };
        else
          t.vertices = { i, i + static_cast<std::uint32_t>(nx) + 2, i + static_cast<std::uint32_t>(nx) + 1 };
        mesh.polygons.push_back(t);
      }
  return mesh;
}

/** tool_path_planner_utest.cpp:30-45 */
Eigen::Isometry3d createRandomTransform(const double translation_limit, const double rotation_limit)
{
  std::mt19937 rand_gen(0);
  std::uniform_real_distribution<double> pos_dist(-std::abs(translation_limit), std::abs(translation_limit));
  Eigen::Vector3d pos = Eigen::Vector3d::NullaryExpr([&rand_gen, &pos_dist]() { return pos_dist(rand_gen); });
  std::uniform_real_distribution<double> rot_dist(-std::abs(rotation_limit), std::abs(rotation_limit));
  Eigen::Vector3d ea = Eigen::Vector3d::NullaryExpr([&rand_gen, &rot_dist]() { return rot_dist(rand_gen); });
  return Eigen::Translation3d(pos) * Eigen::AngleAxisd(ea(0), Eigen::Vector3d::UnitX()) *
         Eigen::AngleAxisd(ea(1), Eigen::Vector3d::UnitY()) * Eigen::AngleAxisd(ea(2), Eigen::Vector3d::UnitZ());
}
}  // namespace

int main(int argc, char** argv)
{
  if (argc < 3)
  {
    std::printf("usage: %s <wavy_mesh_with_hole.ply> <output directory>\n", argv[0]);
    return 2;
  }
  const std::string out_dir = argv[2];
  const Factory factory;  // the stock plugin library, found the default way (plugin_interface.cpp:52-59)
  pcl::PolygonMesh ply;
  if (pcl::io::loadPolygonFile(argv[1], ply) <= 0)
  {
    std::printf("cannot load %s\n", argv[1]);
    return 1;
  }
  YAML::Node fixed_x, fixed_o, centroid, principal;
  fixed_x["name"] = "FixedDirection";
  fixed_x["direction"] = vec(Eigen::Vector3d::UnitX());
  fixed_o["name"] = "FixedOrigin";
  fixed_o["origin"] = vec(Eigen::Vector3d::Zero());
  centroid["name"] = "Centroid";
  principal["name"] = "PrincipalAxis";
  principal["rotation_offset"] = 0.0;

  {  // SemiPlanarMeshFile's geometry (tool_path_planner_utest.cpp:122-177): 10 tool paths, 4..6 split by the hole
    Case c{ "fixture_fixed_1p05", ply, fixed_x, fixed_o };
    c.line_spacing = 1.05;
    run(factory, c, out_dir);
  }
  {  // the YAML of the reference's raster planner tests (:259-275): spacing 0.1, Centroid origin
    Case c{ "fixture_default_yaml", ply, fixed_x, centroid };
    c.line_spacing = 0.1;
    run(factory, c, out_dir);
  }
  {  // PCA direction + centroid origin on the fixture (a2-a5)
    Case c{ "fixture_principal_axis", ply, principal, centroid };
    c.line_spacing = 0.25;
    run(factory, c, out_dir);
  }
  {  // unidirectional lattice from the centroid
    Case c{ "fixture_unidirectional", ply, fixed_x, centroid };
    c.line_spacing = 0.37;
    c.bidirectional = false;
    run(factory, c, out_dir);
  }
  {  // FlatSquareMesh (:74-119): 2 triangles under the test's random transform
    const Eigen::Isometry3d T = createRandomTransform(5.0, M_PI);
    YAML::Node dir, org;
    dir["name"] = "FixedDirection";
    dir["direction"] = vec(T.rotation() * Eigen::Vector3d::UnitX());
    org["name"] = "FixedOrigin";
    org["origin"] = vec(Eigen::Vector3d(T.translation()));
    Case c{ "flat_square", createPlaneMesh(1.0f, 1.0f, T), dir, org };
    c.line_spacing = 0.095;
    run(factory, c, out_dir);
  }
  {  // the synthetic bench surface at smoke size, normals from the reference's own NormalsFromMeshFaces (a10)
    Case c{ "wavy_96_80", wavy(96, 80), principal, centroid };
    c.line_spacing = 0.0107;
    c.compute_normals = true;
    run(factory, c, out_dir);
  }
  {  // PlaneSlicerLegacy (a11) on the fixture
    Case c{ "fixture_legacy", ply, fixed_x, fixed_o };
    c.line_spacing = 1.05;
    c.legacy = true;
    c.point_spacing = 0.25;
    c.min_segment_size = 0.5;
    c.min_hole_size = 0.0;
    run(factory, c, out_dir);
  }
  return 0;
}
