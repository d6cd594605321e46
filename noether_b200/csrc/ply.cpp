// Header of a binary little-endian PLY file -> the layout the device decoders need (pure host code, no CUDA).
//
// The reference reads meshes with pcl::io::loadPolygonFile (noether_gui/src/widgets/tpp_widget.cpp:426; the tests:
// test/tool_path_planner_utest.cpp:24), i.e. vtkPLYReader -> vtk2mesh -> PCLPointCloud2 + one pcl::Vertices per face,
// on one core.  nb2_mesh_upload_ply copies the file image to the GPU once and decodes it there (api.cu, kernels_mesh.cu).
#include "nb2_internal.h"

#include <cstring>
#include <string>
#include <vector>

namespace nb2
{
namespace
{
int scalarSize(const std::string& t)
{
  if (t == "char" || t == "uchar" || t == "int8" || t == "uint8")
    return 1;
  if (t == "short" || t == "ushort" || t == "int16" || t == "uint16")
    return 2;
  if (t == "int" || t == "uint" || t == "int32" || t == "uint32" || t == "float" || t == "float32")
    return 4;
  if (t == "double" || t == "float64")
    return 8;
  return 0;
}
bool isFloat32(const std::string& t) { return t == "float" || t == "float32"; }

std::vector<std::string> split(const std::string& line)
{
  std::vector<std::string> out;
  size_t i = 0;
  while (i < line.size())
  {
    while (i < line.size() && (line[i] == ' ' || line[i] == '\t' || line[i] == '\r'))
      ++i;
    size_t j = i;
    while (j < line.size() && line[j] != ' ' && line[j] != '\t' && line[j] != '\r')
      ++j;
    if (j > i)
      out.emplace_back(line.substr(i, j - i));
    i = j;
  }
  return out;
}

struct Property
{
  std::string name, type;  // scalar type, or the index type of a list
  bool list = false;
  int count_bytes = 0;  // lists only
  int bytes = 0;        // scalar size (lists: size of one index)
};
struct Element
{
  std::string name;
  uint64_t count = 0;
  std::vector<Property> props;
};
}  // namespace

PlyLayout parsePlyHeader(const uint8_t* data, uint64_t n_bytes)
{
  auto bad = [](const std::string& what) -> PlyLayout { fail(NB2_ERR_INVALID_ARGUMENT, "PLY: " + what); };
  if (!data || n_bytes < 4 || std::memcmp(data, "ply", 3) != 0 || (data[3] != '\n' && data[3] != '\r'))
    return bad("not a PLY file (magic 'ply' missing)");
  // the header is ASCII lines up to and including "end_header\n"
  std::vector<Element> elems;
  bool format_seen = false;
  uint64_t pos = 0, header_end = 0;
  const uint64_t scan_limit = n_bytes < (1u << 20) ? n_bytes : (1u << 20);
  while (pos < scan_limit)
  {
    uint64_t e = pos;
    while (e < scan_limit && data[e] != '\n')
      ++e;
    if (e >= scan_limit)
      break;
    const std::string line(reinterpret_cast<const char*>(data + pos), e - pos);
    pos = e + 1;
    const std::vector<std::string> tok = split(line);
    if (tok.empty() || tok[0] == "ply" || tok[0] == "comment" || tok[0] == "obj_info")
      continue;
    if (tok[0] == "end_header")
    {
      header_end = pos;
      break;
    }
    if (tok[0] == "format")
    {
      if (tok.size() < 3 || tok[1] != "binary_little_endian")
        return bad("only binary_little_endian files are decoded on the device (found '" + (tok.size() > 1 ? tok[1] : std::string()) +
                   "'); convert ASCII / big-endian files first");
      format_seen = true;
    }
    else if (tok[0] == "element")
    {
      if (tok.size() < 3)
        return bad("malformed element line");
      Element el;
      el.name = tok[1];
      el.count = std::strtoull(tok[2].c_str(), nullptr, 10);
      elems.push_back(el);
    }
    else if (tok[0] == "property")
    {
      if (elems.empty())
        return bad("property before any element");
      Property p;
      if (tok.size() >= 5 && tok[1] == "list")
      {
        p.list = true;
        p.count_bytes = scalarSize(tok[2]);
        p.type = tok[3];
        p.bytes = scalarSize(tok[3]);
        p.name = tok[4];
        if (!p.count_bytes || !p.bytes || p.count_bytes == 8 || p.bytes == 8)
          return bad("unsupported list property types");
      }
      else if (tok.size() >= 3)
      {
        p.type = tok[1];
        p.bytes = scalarSize(tok[1]);
        p.name = tok[2];
        if (!p.bytes)
          return bad("unknown property type '" + tok[1] + "'");
      }
      else
        return bad("malformed property line");
      elems.back().props.push_back(p);
    }
    else
      return bad("unexpected header line '" + tok[0] + "'");
  }
  if (!header_end)
    return bad("end_header not found");
  if (!format_seen)
    return bad("format line missing");

  PlyLayout L{};
  L.header_bytes = header_end;
  bool have_vertex = false, have_face = false;
  uint64_t at = header_end;
  for (const Element& el : elems)
  {
    if (have_vertex && have_face)
      break;  // whatever follows (edges, materials, ...) is not needed
    if (el.name == "vertex")
    {
      uint32_t step = 0;
      // vtkPLYReader asks for x, y, z and nx, ny, nz by name and converts the stored type to float; float32 and float64
      // properties are decoded on the device, integer-typed coordinates are refused
      static const char* wanted[6] = { "x", "y", "z", "nx", "ny", "nz" };
      bool found[6] = { false, false, false, false, false, false };
      for (const Property& p : el.props)
      {
        if (p.list)
          return bad("list property in the vertex element");
        for (int k = 0; k < 6; ++k)
          if (p.name == wanted[k] && !found[k])
          {
            if (!isFloat32(p.type) && p.bytes != 8)
              return bad("vertex property '" + p.name + "' is neither float nor double");
            found[k] = true;
            L.off_field[k] = step;
            L.field_f64[k] = p.bytes == 8;
          }
        step += p.bytes;
      }
      if (!found[0] || !found[1] || !found[2])
        return bad("the vertex element needs the properties x, y, z");
      L.n_points = el.count;
      L.point_step = step;
      L.has_normals = found[3] && found[4] && found[5];
      L.vertex_begin = at;
      at += el.count * step;
      have_vertex = true;
    }
    else if (el.name == "face")
    {
      int lists = 0;
      uint32_t pre = 0, post = 0;
      for (const Property& p : el.props)
      {
        if (p.list)
        {
          if (p.name != "vertex_indices" && p.name != "vertex_index")
            return bad("unsupported list property '" + p.name + "' in the face element");
          ++lists;
          L.face_count_bytes = p.count_bytes;
          L.face_index_bytes = p.bytes;
        }
        else
          (lists ? post : pre) += p.bytes;
      }
      if (lists != 1)
        return bad("the face element needs exactly one vertex_indices list");
      L.n_faces = el.count;
      L.face_pre = pre;
      // every face must be a triangle (checked on the device): the records then have one size
      L.face_step = pre + L.face_count_bytes + 3 * L.face_index_bytes + post;
      L.face_begin = at;
      at += el.count * L.face_step;
      have_face = true;
    }
    else
    {
      uint64_t step = 0;
      for (const Property& p : el.props)
      {
        if (p.list)
          return bad("element '" + el.name + "' with list properties ahead of the mesh data");
        step += p.bytes;
      }
      at += el.count * step;
    }
  }
  if (!have_vertex)
    return bad("no vertex element");
  if (!have_face)
  {
    L.n_faces = 0;
    L.face_begin = at;
    L.face_step = 0;
  }
  if (L.n_points >= (1ull << 31) || L.n_faces >= (1ull << 31))
    return bad("2^31 or more vertices / faces");
  if (L.vertex_begin + L.n_points * L.point_step > n_bytes || L.face_begin + L.n_faces * L.face_step > n_bytes)
    return bad("file is shorter than its header announces (or holds faces that are not triangles)");
  if (L.n_faces)
  {
    // cheap early refusal (the device checks every face): the first face already tells a quad mesh from a triangle mesh
    const uint8_t* f0 = data + L.face_begin + L.face_pre;
    uint32_t n0 = 0;
    std::memcpy(&n0, f0, L.face_count_bytes);  // little-endian host
    if (n0 != 3)
      fail(NB2_ERR_NOT_TRIANGLES, "PLY: the file holds faces that are not triangles (first face has " + std::to_string(n0) +
                                      " vertices; the plane slicer path takes triangle meshes)");
  }
  return L;
}

}  // namespace nb2
