// TEST INFRASTRUCTURE ONLY — never linked or loaded by the product.
//
// What pcl::io::loadPolygonFile(path, mesh) yields for a PLY file, restated for the checker.  The reference calls it at
// noether_gui/src/widgets/tpp_widget.cpp:426 and noether_tpp/test/tool_path_planner_utest.cpp:24 (and
// test/mesh_modifier_utest.cpp); the implementation is third party and absent here: PCL 1.12 io/src/vtk_lib_io.cpp
// (loadPolygonFilePLY -> vtkPLYReader -> vtk2mesh) over VTK 9.1 IO/PLY/vtkPLYReader.cxx + vtkPLY.cxx.  Published
// behaviour restated:
//   * header: "ply", "format {ascii|binary_little_endian|binary_big_endian} 1.0", comment / obj_info lines skipped,
//     "element <name> <count>", "property <type> <name>", "property list <count type> <index type> <name>";
//   * vtkPLYReader asks the vertex element for x, y, z as float (ply::get_element converts whatever type is stored),
//     nx, ny, nz as float -> point normals when all three exist, and the face element for the list vertex_indices
//     (or vertex_index) as int; other properties are read past;
//   * vtk2mesh copies the points in file order (float), the normals into normal_x / normal_y / normal_z, and every
//     polygon's ids in file order into pcl::Vertices.
// PARITY STATUS: unpinned (no PLY golden data beyond the planner expectations on wavy_mesh_with_hole.ply, which the
// slicing tests already cover through this reader's output).
#include "oracle.h"
#include "oracle_internal.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

struct orc_mesh
{
  std::vector<float> xyz, nrm;
  std::vector<uint32_t> face_size, index;
  bool has_normals = false;
};

namespace
{
int plyError(const std::string& msg)
{
  orc::setError(msg);
  return ORC_ERR_INVALID;
}
enum class Fmt
{
  Ascii,
  LE,
  BE
};
struct Prop
{
  std::string name;
  int type = 0;        // 0 i8 1 u8 2 i16 3 u16 4 i32 5 u32 6 f32 7 f64
  int count_type = -1; // >= 0: list
};
struct Elem
{
  std::string name;
  uint64_t count = 0;
  std::vector<Prop> props;
};
int typeOf(const std::string& t)
{
  static const char* names[][2] = { { "char", "int8" },   { "uchar", "uint8" }, { "short", "int16" },   { "ushort", "uint16" },
                                    { "int", "int32" },   { "uint", "uint32" }, { "float", "float32" }, { "double", "float64" } };
  for (int i = 0; i < 8; ++i)
    if (t == names[i][0] || t == names[i][1])
      return i;
  return -1;
}
const int kSize[8] = { 1, 1, 2, 2, 4, 4, 4, 8 };

struct Cursor
{
  const uint8_t* p;
  const uint8_t* end;
  Fmt fmt;
  bool ok = true;
  /** next scalar of the stored type, as double (what ply::get_stored_item / get_ascii_item hand to the converter) */
  double next(int type)
  {
    if (fmt == Fmt::Ascii)
    {
      while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t'))
        ++p;
      if (p >= end)
      {
        ok = false;
        return 0;
      }
      const uint8_t* q = p;
      while (q < end && !(*q == ' ' || *q == '\n' || *q == '\r' || *q == '\t'))
        ++q;
      const std::string tok(reinterpret_cast<const char*>(p), q - p);
      p = q;
      return std::strtod(tok.c_str(), nullptr);
    }
    const int n = kSize[type];
    if (end - p < n)
    {
      ok = false;
      return 0;
    }
    uint8_t b[8];
    for (int i = 0; i < n; ++i)
      b[i] = fmt == Fmt::LE ? p[i] : p[n - 1 - i];
    p += n;
    switch (type)
    {
      case 0: { int8_t v; std::memcpy(&v, b, 1); return v; }
      case 1: return b[0];
      case 2: { int16_t v; std::memcpy(&v, b, 2); return v; }
      case 3: { uint16_t v; std::memcpy(&v, b, 2); return v; }
      case 4: { int32_t v; std::memcpy(&v, b, 4); return v; }
      case 5: { uint32_t v; std::memcpy(&v, b, 4); return v; }
      case 6: { float v; std::memcpy(&v, b, 4); return v; }
      default: { double v; std::memcpy(&v, b, 8); return v; }
    }
  }
};
}  // namespace

extern "C" {

int orc_read_ply(const uint8_t* data, uint64_t n_bytes, orc_mesh** out)
{
  if (!data || !out)
    return plyError("orc_read_ply: null argument");
  *out = nullptr;
  if (n_bytes < 4 || std::memcmp(data, "ply", 3) != 0)
    return plyError("orc_read_ply: not a PLY file");
  std::vector<Elem> elems;
  Fmt fmt = Fmt::Ascii;
  bool have_fmt = false;
  uint64_t pos = 0, body = 0;
  while (pos < n_bytes)
  {
    uint64_t e = pos;
    while (e < n_bytes && data[e] != '\n')
      ++e;
    if (e >= n_bytes)
      break;
    std::string line(reinterpret_cast<const char*>(data + pos), e - pos);
    pos = e + 1;
    std::istringstream is(line);
    std::vector<std::string> tok;
    for (std::string t; is >> t;)
      tok.push_back(t);
    if (tok.empty() || tok[0] == "ply" || tok[0] == "comment" || tok[0] == "obj_info")
      continue;
    if (tok[0] == "end_header")
    {
      body = pos;
      break;
    }
    if (tok[0] == "format" && tok.size() >= 2)
    {
      have_fmt = true;
      if (tok[1] == "ascii")
        fmt = Fmt::Ascii;
      else if (tok[1] == "binary_little_endian")
        fmt = Fmt::LE;
      else if (tok[1] == "binary_big_endian")
        fmt = Fmt::BE;
      else
        return plyError("orc_read_ply: unknown format");
    }
    else if (tok[0] == "element" && tok.size() >= 3)
    {
      Elem el;
      el.name = tok[1];
      el.count = std::strtoull(tok[2].c_str(), nullptr, 10);
      elems.push_back(el);
    }
    else if (tok[0] == "property" && !elems.empty())
    {
      Prop p;
      if (tok.size() >= 5 && tok[1] == "list")
      {
        p.count_type = typeOf(tok[2]);
        p.type = typeOf(tok[3]);
        p.name = tok[4];
        if (p.count_type < 0 || p.type < 0)
          return plyError("orc_read_ply: unknown list type");
      }
      else if (tok.size() >= 3)
      {
        p.type = typeOf(tok[1]);
        p.name = tok[2];
        if (p.type < 0)
          return plyError("orc_read_ply: unknown property type");
      }
      else
        return plyError("orc_read_ply: malformed property");
      elems.back().props.push_back(p);
    }
    else
      return plyError("orc_read_ply: unexpected header line");
  }
  if (!body || !have_fmt)
    return plyError("orc_read_ply: incomplete header");

  auto mesh = new orc_mesh();
  Cursor cur{ data + body, data + n_bytes, fmt };
  for (const Elem& el : elems)
  {
    const bool is_vertex = el.name == "vertex", is_face = el.name == "face";
    int ix = -1, iy = -1, iz = -1, inx = -1, iny = -1, inz = -1, ilist = -1;
    for (size_t k = 0; k < el.props.size(); ++k)
    {
      const Prop& p = el.props[k];
      if (is_vertex && p.count_type < 0)
      {
        if (p.name == "x") ix = static_cast<int>(k);
        if (p.name == "y") iy = static_cast<int>(k);
        if (p.name == "z") iz = static_cast<int>(k);
        if (p.name == "nx") inx = static_cast<int>(k);
        if (p.name == "ny") iny = static_cast<int>(k);
        if (p.name == "nz") inz = static_cast<int>(k);
      }
      if (is_face && p.count_type >= 0 && (p.name == "vertex_indices" || p.name == "vertex_index"))
        ilist = static_cast<int>(k);
    }
    const bool normals = is_vertex && inx >= 0 && iny >= 0 && inz >= 0;
    if (is_vertex)
    {
      if (ix < 0 || iy < 0 || iz < 0)
      {
        delete mesh;
        return plyError("orc_read_ply: vertex element without x, y, z");
      }
      mesh->has_normals = normals;
      mesh->xyz.resize(3 * el.count);
      if (normals)
        mesh->nrm.resize(3 * el.count);
    }
    std::vector<double> scalar(el.props.size());
    for (uint64_t r = 0; r < el.count; ++r)
    {
      for (size_t k = 0; k < el.props.size(); ++k)
      {
        const Prop& p = el.props[k];
        if (p.count_type < 0)
        {
          scalar[k] = cur.next(p.type);
          continue;
        }
        const long long cnt = static_cast<long long>(cur.next(p.count_type));
        if (static_cast<int>(k) == ilist)
          mesh->face_size.push_back(static_cast<uint32_t>(cnt));
        for (long long j = 0; j < cnt && cur.ok; ++j)
        {
          const double v = cur.next(p.type);
          if (static_cast<int>(k) == ilist)
            mesh->index.push_back(static_cast<uint32_t>(static_cast<int32_t>(static_cast<long long>(v))));  // stored -> int
        }
      }
      if (!cur.ok)
      {
        delete mesh;
        return plyError("orc_read_ply: file ends inside element '" + el.name + "'");
      }
      if (is_vertex)
      {
        mesh->xyz[3 * r] = static_cast<float>(scalar[ix]);
        mesh->xyz[3 * r + 1] = static_cast<float>(scalar[iy]);
        mesh->xyz[3 * r + 2] = static_cast<float>(scalar[iz]);
        if (normals)
        {
          mesh->nrm[3 * r] = static_cast<float>(scalar[inx]);
          mesh->nrm[3 * r + 1] = static_cast<float>(scalar[iny]);
          mesh->nrm[3 * r + 2] = static_cast<float>(scalar[inz]);
        }
      }
    }
  }
  *out = mesh;
  return ORC_OK;
}

/* pcl::io::loadPolygonFile for a binary STL (PCL 1.12 vtk_lib_io.cpp loadPolygonFileSTL -> vtkSTLReader -> vtk2mesh; VTK 9.1
 * IO/Geometry/vtkSTLReader.cxx), restated:
 *   * ReadBinarySTL: 80-byte header, 4-byte little-endian facet count (read but not relied upon), then 50-byte facets
 *     {float n[3], v1[3], v2[3], v3[3]; uint16 attribute} for as long as fread(&facet, 48, 1, fp) succeeds; the three
 *     corners of every facet are appended to the point list, one triangle per facet;
 *   * Merging (on by default): every corner in facet order goes through vtkMergePoints::InsertUniquePoint — float
 *     comparison x == x', y == y', z == z' inside the corner's bucket, so -0 equals +0, the first inserted point stays,
 *     ids count first occurrences; a facet is kept when its three ids differ;
 *   * vtk2mesh: points in id order as float xyz (no normals: the reader drops the facet normals), polygons in order.
 * Text STL ("solid" with a size that does not match the binary layout) is outside this restatement, as it is outside
 * the product.  PARITY STATUS: unpinned (the reference ships no STL fixture). */
int orc_read_stl(const uint8_t* data, uint64_t n_bytes, orc_mesh** out)
{
  if (!data || !out)
    return plyError("orc_read_stl: NULL argument");
  if (n_bytes < 84)
    return plyError("orc_read_stl: shorter than a binary STL header");
  uint32_t declared;
  std::memcpy(&declared, data + 80, 4);
  if (n_bytes != 84 + 50ull * declared && std::memcmp(data, "solid", 5) == 0)
    return plyError("orc_read_stl: ASCII STL");
  auto* mesh = new orc_mesh();
  struct Bucketed
  {
    std::vector<uint32_t> ids;
  };
  // any bucketing that sends equal coordinates (with -0 == +0) to one bucket reproduces vtkMergePoints' result
  std::vector<std::vector<uint32_t>> buckets(1u << 16);
  auto canon = [](float f) {
    uint32_t b;
    std::memcpy(&b, &f, 4);
    return b == 0x80000000u ? 0u : b;
  };
  for (uint64_t at = 84; at + 48 <= n_bytes; at += 50)
  {
    uint32_t node[3];
    for (int k = 0; k < 3; ++k)
    {
      float x[3];
      std::memcpy(x, data + at + 12 + 12 * k, 12);
      for (int d = 0; d < 3; ++d)
        if (!std::isfinite(x[d]))
        {
          delete mesh;
          return plyError("orc_read_stl: non-finite coordinate (undefined bucket index in vtkMergePoints)");
        }
      const uint32_t h = (canon(x[0]) * 0x9e3779b1u) ^ (canon(x[1]) * 0x85ebca77u) ^ (canon(x[2]) * 0xc2b2ae3du);
      std::vector<uint32_t>& bucket = buckets[(h ^ (h >> 16)) & 0xffffu];
      uint32_t found = UINT32_MAX;
      for (uint32_t id : bucket)
        if (mesh->xyz[3 * id] == x[0] && mesh->xyz[3 * id + 1] == x[1] && mesh->xyz[3 * id + 2] == x[2])
        {
          found = id;
          break;
        }
      if (found == UINT32_MAX)
      {
        found = static_cast<uint32_t>(mesh->xyz.size() / 3);
        mesh->xyz.insert(mesh->xyz.end(), x, x + 3);
        bucket.push_back(found);
      }
      node[k] = found;
    }
    if (node[0] != node[1] && node[0] != node[2] && node[1] != node[2])
    {
      mesh->face_size.push_back(3);
      mesh->index.insert(mesh->index.end(), node, node + 3);
    }
  }
  *out = mesh;
  return ORC_OK;
}

uint64_t orc_mesh_num_vertices(const orc_mesh* m) { return m ? m->xyz.size() / 3 : 0; }
uint64_t orc_mesh_num_faces(const orc_mesh* m) { return m ? m->face_size.size() : 0; }
uint64_t orc_mesh_num_indices(const orc_mesh* m) { return m ? m->index.size() : 0; }
int orc_mesh_has_normals(const orc_mesh* m) { return m && m->has_normals; }
void orc_mesh_export(const orc_mesh* m, float* xyz, float* nrm, uint32_t* face_size, uint32_t* index)
{
  if (!m)
    return;
  if (xyz)
    std::memcpy(xyz, m->xyz.data(), m->xyz.size() * sizeof(float));
  if (nrm && m->has_normals)
    std::memcpy(nrm, m->nrm.data(), m->nrm.size() * sizeof(float));
  if (face_size)
    std::memcpy(face_size, m->face_size.data(), m->face_size.size() * sizeof(uint32_t));
  if (index)
    std::memcpy(index, m->index.data(), m->index.size() * sizeof(uint32_t));
}
void orc_mesh_free(orc_mesh* m) { delete m; }

}  // extern "C"
