// TEST INFRASTRUCTURE ONLY — see oracle.h.
//
// Restatement of the plane loop of noether::PlaneSlicerRasterPlanner::planImpl
// (noether_tpp/src/tool_path_planners/raster/plane_slicer_raster_planner.cpp:596-628) and of the VTK 9.1 filters it
// drives.  VTK sources are not under /root/reference; their behaviour is restated from the published algorithm
// (SURVEY.md §8c V1-V6):
//   vtkPlane::EvaluateFunction              -> planeScalar
//   vtkCutter::DataSetCutter (SortByCell)   -> cutPlane            (call site :604-608)
//   vtkTriangle::Contour + vtkMergePoints   -> cutPlane
//   vtkStripper (lines, JoinContiguousSegmentsOff) -> strip        (call site :610-615)
//   unifySegmentDirection (:183-204), topologicalSortSegments (:87-172, boost::topological_sort semantics),
//   joinContiguousSegments (:242-288,302-358), convertToPoses/createTransform (:380-448)
#include "oracle_internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <stack>
#include <unordered_map>

namespace orc
{
namespace
{
struct Mesh
{
  const float* xyz;
  const float* nrm;
  uint64_t V;
  const uint32_t* tri;
  uint64_t F;
};

/** Output of one vtkCutter::Update(): merged float points with interpolated normals, and 2-point lines */
struct CutOutput
{
  std::vector<V3f> pts;
  std::vector<V3f> nrms;
  std::vector<uint32_t> lo, hi;
  std::vector<std::array<int64_t, 2>> lines;
  std::vector<uint64_t> line_cell;  // input cell that produced each line
};

struct FloatKey
{
  uint32_t b[3];
  bool operator==(const FloatKey& o) const { return b[0] == o.b[0] && b[1] == o.b[1] && b[2] == o.b[2]; }
};
struct FloatKeyHash
{
  size_t operator()(const FloatKey& k) const
  {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 3; ++i)
    {
      h ^= k.b[i];
      h *= 1099511628211ull;
    }
    return static_cast<size_t>(h);
  }
};
/** vtkMergePoints compares float coordinates with operator== : -0.0f equals +0.0f */
inline FloatKey makeKey(const V3f& p)
{
  FloatKey k;
  for (int i = 0; i < 3; ++i)
  {
    float f = p[i];
    if (f == 0.f)
      f = 0.f;  // collapses -0 onto +0
    std::memcpy(&k.b[i], &f, 4);
  }
  return k;
}

/** Plane i of the lattice: current_loc = start_loc + i * line_spacing_ * cut_normal (:598) */
inline V3d planeOrigin(const orc_lattice& L, uint64_t i)
{
  const double is = static_cast<double>(i) * L.line_spacing;
  return { L.start_loc[0] + is * L.cut_normal[0], L.start_loc[1] + is * L.cut_normal[1],
           L.start_loc[2] + is * L.cut_normal[2] };
}

/** vtkPlane::EvaluateFunction: n . (x - o), double, left-to-right */
inline double planeScalar(const double n[3], const V3d& o, const float* x)
{
  return n[0] * (static_cast<double>(x[0]) - o.x) + n[1] * (static_cast<double>(x[1]) - o.y) +
         n[2] * (static_cast<double>(x[2]) - o.z);
}

// vtkTriangle::Contour tables
const int kEdges[3][2] = { { 0, 1 }, { 1, 2 }, { 2, 0 } };
const int kLineCases[8][2] = { { -1, -1 }, { 0, 2 }, { 1, 0 }, { 1, 2 }, { 2, 1 }, { 0, 1 }, { 2, 0 }, { -1, -1 } };

/** Run the cutter for one plane over `cells` (ascending cell ids).  `scalars` is either the per-vertex scalar array
 *  (ALL_CELLS) or null (scalars are evaluated per cell; identical values). */
void cutPlane(const Mesh& m,
              const orc_lattice& L,
              const V3d& origin,
              const uint32_t* cells,
              uint64_t n_cells,
              bool all_cells,
              const double* scalars,
              CutOutput& out)
{
  std::unordered_map<FloatKey, int64_t, FloatKeyHash> locator;
  const uint64_t count = all_cells ? m.F : n_cells;
  for (uint64_t ci = 0; ci < count; ++ci)
  {
    const uint64_t cell = all_cells ? ci : cells[ci];
    const uint32_t ids[3] = { m.tri[3 * cell], m.tri[3 * cell + 1], m.tri[3 * cell + 2] };
    double s[3];
    for (int i = 0; i < 3; ++i)
      s[i] = scalars ? scalars[ids[i]] : planeScalar(L.cut_normal, origin, m.xyz + 3 * ids[i]);

    int index = 0;
    for (int i = 0; i < 3; ++i)
      if (s[i] >= 0.0)
        index |= (1 << i);
    const int* edge = kLineCases[index];
    if (edge[0] < 0)
      continue;

    int64_t pts[2];
    for (int i = 0; i < 2; ++i)
    {
      const int* vert = kEdges[edge[i]];
      double deltaScalar = s[vert[1]] - s[vert[0]];
      int e1, e2;
      if (deltaScalar > 0)
      {
        e1 = vert[0];
        e2 = vert[1];
      }
      else
      {
        e1 = vert[1];
        e2 = vert[0];
        deltaScalar = -deltaScalar;
      }
      const double t = (deltaScalar == 0.0) ? 0.0 : (0.0 - s[e1]) / deltaScalar;
      const float* x1f = m.xyz + 3 * ids[e1];
      const float* x2f = m.xyz + 3 * ids[e2];
      V3f xf;
      for (int j = 0; j < 3; ++j)
      {
        const double x1 = x1f[j], x2 = x2f[j];
        xf[j] = static_cast<float>(x1 + t * (x2 - x1));
      }
      // vtkMergePoints::InsertUniquePoint
      const FloatKey key = makeKey(xf);
      auto it = locator.find(key);
      if (it != locator.end())
      {
        pts[i] = it->second;
      }
      else
      {
        pts[i] = static_cast<int64_t>(out.pts.size());
        locator.emplace(key, pts[i]);
        out.pts.push_back(xf);
        // vtkDataSetAttributes::InterpolateEdge -> InterpolateTuple: a1 * (1 - t) + a2 * t in double, cast to float
        V3f nf = { 0.f, 0.f, 0.f };
        if (m.nrm)
        {
          const float* a1 = m.nrm + 3 * ids[e1];
          const float* a2 = m.nrm + 3 * ids[e2];
          const double oneMinusT = 1.0 - t;
          for (int j = 0; j < 3; ++j)
            nf[j] = static_cast<float>(static_cast<double>(a1[j]) * oneMinusT + static_cast<double>(a2[j]) * t);
        }
        out.nrms.push_back(nf);
        out.lo.push_back(ids[e1]);
        out.hi.push_back(ids[e2]);
      }
    }
    if (pts[0] != pts[1])
    {
      out.lines.push_back({ pts[0], pts[1] });
      out.line_cell.push_back(cell);
    }
  }
}

using IdList = std::vector<int64_t>;

/** vtkStripper on 2-point lines (V6): seed = lowest unvisited line, extend in one direction only */
std::vector<IdList> strip(const CutOutput& cut)
{
  const size_t nl = cut.lines.size();
  // point -> incident lines, ascending line id (vtkPolyData::BuildLinks)
  std::vector<std::vector<int64_t>> links(cut.pts.size());
  for (size_t l = 0; l < nl; ++l)
  {
    links[cut.lines[l][0]].push_back(static_cast<int64_t>(l));
    links[cut.lines[l][1]].push_back(static_cast<int64_t>(l));
  }
  std::vector<char> visited(nl, 0);
  std::vector<IdList> polylines;
  for (size_t cellId = 0; cellId < nl; ++cellId)
  {
    if (visited[cellId])
      continue;
    visited[cellId] = 1;
    const auto& linePts = cut.lines[cellId];
    IdList pts(2);
    int64_t neighbor = -1;
    bool foundOne = false;
    for (int i = 0; !foundOne && i < 2; ++i)
    {
      pts[0] = linePts[i];
      pts[1] = linePts[(i + 1) % 2];
      for (int64_t nb : links[pts[1]])
      {
        if (nb != static_cast<int64_t>(cellId) && !visited[nb])
        {
          neighbor = nb;
          foundOne = true;
          break;
        }
      }
    }
    if (!foundOne)
    {
      polylines.push_back({ linePts[0], linePts[1] });
      continue;
    }
    while (neighbor >= 0)
    {
      visited[neighbor] = 1;
      const auto& lp = cut.lines[neighbor];
      const int64_t next_pt = (lp[0] != pts.back()) ? lp[0] : lp[1];
      pts.push_back(next_pt);
      neighbor = -1;
      for (int64_t id : links[next_pt])
      {
        if (!visited[id])
        {
          neighbor = id;
          break;
        }
      }
    }
    polylines.push_back(pts);
  }
  return polylines;
}

inline V3d pointD(const CutOutput& cut, int64_t id) { return toDouble(cut.pts[id]); }

/** unifySegmentDirection (:183-204) */
void unifySegmentDirection(const CutOutput& cut, std::vector<IdList>& segs, const V3d& ref)
{
  for (auto& seg : segs)
  {
    const V3d dir = pointD(cut, seg[1]) - pointD(cut, seg[0]);
    if (eigenDot(dir, ref) < 0.0)
      std::reverse(seg.begin(), seg.end());
  }
}

/** vtkIdList::IntersectWith then the reference's checks (:31-44).  Returns 0 none, 1 found, -1 error */
int findConnectingVertex(const IdList& s1, const IdList& s2, int64_t& out)
{
  int n = 0;
  for (int64_t id : s1)
  {
    if (std::find(s2.begin(), s2.end(), id) != s2.end())
    {
      if (n == 0)
        out = id;
      ++n;
    }
  }
  if (n > 1)
  {
    setError("Too many connecting vertices found between adjacent segments");
    return -1;
  }
  return n;
}

/** topologicalSortSegments (:87-172) with boost::directed_graph / boost::topological_sort semantics: depth-first
 *  search over vertices in insertion order, out-edges in insertion order, vertices emitted on finish, back edge ->
 *  not_a_dag; the result is then reversed and reduced to first occurrences of segment indices. */
int topologicalSortSegments(std::vector<IdList>& segs)
{
  const size_t n_segments = segs.size();
  if (n_segments <= 1)
    return ORC_OK;

  struct GV
  {
    size_t seg;
    int64_t id;
  };
  std::vector<GV> gv;
  std::vector<std::vector<size_t>> adj;
  std::vector<size_t> seg_first(n_segments);
  for (size_t i = 0; i < n_segments; ++i)
  {
    seg_first[i] = gv.size();
    for (size_t j = 0; j < segs[i].size(); ++j)
    {
      gv.push_back({ i, segs[i][j] });
      adj.emplace_back();
      if (j > 0)
        adj[gv.size() - 2].push_back(gv.size() - 1);
    }
  }
  auto findVertexDescriptor = [&](size_t seg, int64_t id) -> size_t {
    // linear scan over all graph vertices in insertion order (:60-70); the first match is inside segment `seg`
    for (size_t k = seg_first[seg]; k < seg_first[seg] + segs[seg].size(); ++k)
      if (gv[k].id == id)
        return k;
    return static_cast<size_t>(-1);
  };
  for (size_t i = 0; i + 1 < n_segments; ++i)
  {
    for (size_t j = i + 1; j < n_segments; ++j)
    {
      int64_t shared = -1;
      const int r = findConnectingVertex(segs[i], segs[j], shared);
      if (r < 0)
        return ORC_ERR_TOO_MANY_CONNECTING;
      if (r == 1)
        adj[findVertexDescriptor(i, shared)].push_back(findVertexDescriptor(j, shared));
    }
  }

  // depth_first_search + topo_sort_visitor
  const size_t N = gv.size();
  std::vector<char> color(N, 0);  // 0 white, 1 gray, 2 black
  std::vector<size_t> order;      // finish order
  order.reserve(N);
  std::vector<std::pair<size_t, size_t>> st;
  for (size_t root = 0; root < N; ++root)
  {
    if (color[root] != 0)
      continue;
    color[root] = 1;
    st.push_back({ root, 0 });
    while (!st.empty())
    {
      auto& top = st.back();
      const size_t u = top.first;
      if (top.second < adj[u].size())
      {
        const size_t v = adj[u][top.second++];
        if (color[v] == 0)
        {
          color[v] = 1;
          st.push_back({ v, 0 });
        }
        else if (color[v] == 1)
        {
          setError("The graph must be a DAG.");
          return ORC_ERR_NOT_A_DAG;
        }
      }
      else
      {
        color[u] = 2;
        order.push_back(u);
        st.pop_back();
      }
    }
  }
  std::reverse(order.begin(), order.end());

  std::vector<size_t> sorted_idx;
  for (size_t v : order)
    if (std::find(sorted_idx.begin(), sorted_idx.end(), gv[v].seg) == sorted_idx.end())
      sorted_idx.push_back(gv[v].seg);
  if (sorted_idx.size() != n_segments)
  {
    setError("Topological sort failed");
    return ORC_ERR_TOPO_SORT;
  }
  std::vector<IdList> sorted;
  sorted.reserve(n_segments);
  for (size_t idx : sorted_idx)
    sorted.push_back(segs[idx]);
  segs.swap(sorted);
  return ORC_OK;
}

IdList concatenateUnique(const IdList& l1, const IdList& l2)
{
  IdList out = l1;
  for (int64_t id : l2)
    if (std::find(out.begin(), out.end(), id) == out.end())
      out.push_back(id);
  return out;
}
IdList reversed(const IdList& l) { return IdList(l.rbegin(), l.rend()); }

/** joinContiguousSegments(s1, s2, connecting_vertex) (:242-288) */
int joinPair(const IdList& s1, const IdList& s2, int64_t cv, IdList& out)
{
  const size_t idx_s1 = std::find(s1.begin(), s1.end(), cv) - s1.begin();
  const size_t idx_s2 = std::find(s2.begin(), s2.end(), cv) - s2.begin();
  if (idx_s1 == 0)
  {
    if (idx_s2 == s2.size() - 1)
      out = concatenateUnique(s2, s1);
    else if (idx_s2 == 0)
      out = concatenateUnique(reversed(s2), s1);
    else
    {
      setError("Connecting vertex exists at the beginning of one segment but not at the end or beginning of the "
               "adjacent segment");
      return ORC_ERR_JOIN;
    }
    return ORC_OK;
  }
  else if (idx_s1 == s1.size() - 1)
  {
    if (idx_s2 == 0)
      out = concatenateUnique(s1, s2);
    else if (idx_s2 == s2.size() - 1)
      out = concatenateUnique(s1, reversed(s2));
    else
    {
      setError("Connecting vertex exists at the end of one segment but not at the beginning or end of the adjacent "
               "segment");
      return ORC_ERR_JOIN;
    }
    return ORC_OK;
  }
  setError("Connecting vertex does not exist at the start or end of a segment");
  return ORC_ERR_JOIN;
}

/** joinContiguousSegments(polydata) (:302-358) */
int joinContiguousSegments(std::vector<IdList>& segs)
{
  const size_t n_segments = segs.size();
  if (n_segments <= 1)
    return ORC_OK;
  std::vector<IdList> output;
  std::stack<IdList> stack;
  for (size_t i = n_segments; i-- > 0;)
    stack.push(segs[i]);
  while (!stack.empty())
  {
    IdList s1 = stack.top();
    stack.pop();
    if (stack.empty())
    {
      output.push_back(s1);
      continue;
    }
    IdList s2 = stack.top();
    stack.pop();
    int64_t cv = -1;
    const int r = findConnectingVertex(s1, s2, cv);
    if (r < 0)
      return ORC_ERR_TOO_MANY_CONNECTING;
    if (r == 0)
    {
      output.push_back(s1);
      stack.push(s2);
    }
    else
    {
      IdList joined;
      const int rc = joinPair(s1, s2, cv, joined);
      if (rc != ORC_OK)
        return rc;
      stack.push(joined);
    }
  }
  segs.swap(output);
  return ORC_OK;
}

/** Chains of the cut graph in a canonical orientation and order — the reference's documented contract ("joins
 *  contiguous line segments", plane_slicer_raster_planner.cpp:290-301) made independent of traversal order; see
 *  DESIGN.md "intent mode".
 *
 *  Every 2-point line has two ends; end tag = 2 * cell id + end index (the order in which vtkCutter produced them).
 *   - pairing: at each merged cut point the incident line ends are sorted by tag and paired consecutively
 *     (1st with 2nd, 3rd with 4th, ...); an unpaired end is an open end.  At an ordinary point (two incident lines)
 *     this is simply "the two lines continue into each other".
 *   - open chain: walk from an open end to the other open end.  Orientation: (p_last - p_first) . cut_direction > 0
 *     (double, left-to-right sum); on a zero dot the open end with the lower tag comes first.
 *   - closed loop (no open end): starts at end 0 of its lowest-tag line and first walks that line (what vtkStripper
 *     does from its seed); reversed, keeping the start point, when that first step . cut_direction < 0.  The start
 *     point is repeated at the end, as vtkStripper emits a closed polyline.
 *   - order within the plane: ascending p_first . cut_direction, ties by the tag of the chain's first line end. */
int maximalChains(const CutOutput& cut, const V3d& ref, std::vector<IdList>& segs)
{
  const size_t np = cut.pts.size();
  const size_t nl = cut.lines.size();
  const size_t ne = 2 * nl;
  auto tagOf = [&](size_t e) { return 2 * cut.line_cell[e >> 1] + (e & 1); };
  auto pointOf = [&](size_t e) { return cut.lines[e >> 1][e & 1]; };
  // incident ends per point, ascending tag (lines are stored in cell order, so ascending end index)
  std::vector<std::vector<size_t>> ends(np);
  for (size_t e = 0; e < ne; ++e)
    ends[pointOf(e)].push_back(e);
  const size_t kNone = static_cast<size_t>(-1);
  std::vector<size_t> twin(ne, kNone);
  for (const auto& v : ends)
    for (size_t i = 0; i + 1 < v.size(); i += 2)
    {
      twin[v[i]] = v[i + 1];
      twin[v[i + 1]] = v[i];
    }

  struct Keyed
  {
    double key;
    uint64_t tag;
    IdList ids;
  };
  std::vector<Keyed> chains;
  std::vector<char> used(nl, 0);
  auto dotLR = [&](const V3d& a) { return (a.x * ref.x + a.y * ref.y) + a.z * ref.z; };

  // open chains
  for (size_t e0 = 0; e0 < ne; ++e0)
  {
    if (twin[e0] != kNone || used[e0 >> 1])
      continue;
    IdList ids;
    size_t e = e0, last_end = e0;
    while (true)
    {
      used[e >> 1] = 1;
      ids.push_back(pointOf(e));
      last_end = e ^ 1;
      const size_t nx = twin[e ^ 1];
      if (nx == kNone)
        break;
      e = nx;
    }
    ids.push_back(pointOf(last_end));
    const double dd = dotLR(pointD(cut, ids.back()) - pointD(cut, ids.front()));
    uint64_t first_tag = tagOf(e0);
    if (dd < 0.0 || (dd == 0.0 && tagOf(last_end) < tagOf(e0)))
    {
      std::reverse(ids.begin(), ids.end());
      first_tag = tagOf(last_end);
    }
    chains.push_back({ dotLR(pointD(cut, ids.front())), first_tag, std::move(ids) });
  }
  // closed loops
  for (size_t l = 0; l < nl; ++l)
  {
    if (used[l])
      continue;
    IdList ids;
    size_t e = 2 * l;
    while (!used[e >> 1])
    {
      used[e >> 1] = 1;
      ids.push_back(pointOf(e));
      e = twin[e ^ 1];
    }
    ids.push_back(ids.front());
    const double dd = dotLR(pointD(cut, ids[1]) - pointD(cut, ids[0]));
    if (dd < 0.0)
      std::reverse(ids.begin(), ids.end());
    chains.push_back({ dotLR(pointD(cut, ids.front())), tagOf(2 * l), std::move(ids) });
  }
  std::sort(chains.begin(), chains.end(), [](const Keyed& a, const Keyed& b) {
    if (a.key != b.key)
      return a.key < b.key;
    return a.tag < b.tag;
  });
  segs.clear();
  for (auto& c : chains)
    segs.push_back(std::move(c.ids));
  return ORC_OK;
}

/** createTransform (:380-394) */
void createTransform(const V3d& position, const V3d& normal, const V3d& path_dir, double T[16])
{
  const V3d y = eigenCross(normal, path_dir);
  const V3d x = eigenCross(y, normal);
  const V3d xn = eigenNormalized(x), yn = eigenNormalized(y), zn = eigenNormalized(normal);
  const double M[16] = { xn.x, xn.y, xn.z, 0.0, yn.x, yn.y, yn.z, 0.0, zn.x, zn.y, zn.z, 0.0, position.x, position.y,
                         position.z, 1.0 };
  std::copy(M, M + 16, T);
}

/** convertToPoses (:396-448) */
Path convertToPoses(const CutOutput& cut, const std::vector<IdList>& segs, int64_t plane)
{
  Path path;
  path.plane = plane;
  for (const IdList& ids : segs)
  {
    Segment seg;
    const size_t n = ids.size();
    for (size_t i = 0; i < n; ++i)
    {
      const int64_t id = ids[i];
      const V3d point = pointD(cut, id);
      const V3d normal = toDouble(cut.nrms[id]);
      V3d dir;
      if (i < n - 1)
        dir = pointD(cut, ids[i + 1]) - point;
      else
        dir = point - pointD(cut, ids[i - 1]);
      Waypoint w;
      createTransform(point, normal, dir, w.T);
      for (int c = 0; c < 3; ++c)
      {
        w.pos[c] = cut.pts[id][c];
        w.nrm[c] = cut.nrms[id][c];
      }
      w.edge_lo = cut.lo[id];
      w.edge_hi = cut.hi[id];
      seg.push_back(w);
    }
    path.segments.push_back(std::move(seg));
  }
  return path;
}

}  // namespace

int slice(const float* xyz,
          const float* nrm,
          uint64_t V,
          const uint32_t* tri,
          uint64_t F,
          const orc_lattice& L,
          int slice_mode,
          int scan_mode,
          uint64_t plane_begin,
          uint64_t plane_end,
          orc_paths& out)
{
  const Mesh m{ xyz, nrm, V, tri, F };
  for (uint64_t f = 0; f < 3 * F; ++f)
  {
    if (tri[f] >= V)
    {
      setError("triangle references a vertex out of range");
      return ORC_ERR_INVALID;
    }
  }
  plane_end = std::min<uint64_t>(plane_end, L.num_planes);
  const V3d ref = { L.cut_direction[0], L.cut_direction[1], L.cut_direction[2] };
  out.paths.clear();

  // Optional conservative binning of cells by lattice index (ORC_SCAN_BINNED).  A cell missing from a plane's list
  // would have produced vtkTriangle case 0 or 7 (no output, no locator insertion), so results are identical.
  std::vector<uint64_t> bin_off;
  std::vector<uint32_t> bin_cells;
  if (scan_mode == ORC_SCAN_BINNED && plane_end > plane_begin)
  {
    const uint64_t np = plane_end - plane_begin;
    auto interval = [&](uint64_t f, int64_t& k0, int64_t& k1) {
      double dmin = 1e300, dmax = -1e300;
      for (int i = 0; i < 3; ++i)
      {
        const float* x = xyz + 3 * tri[3 * f + i];
        const double d = (L.cut_normal[0] * (x[0] - L.start_loc[0]) + L.cut_normal[1] * (x[1] - L.start_loc[1]) +
                          L.cut_normal[2] * (x[2] - L.start_loc[2])) /
                         L.line_spacing;
        dmin = std::min(dmin, d);
        dmax = std::max(dmax, d);
      }
      k0 = static_cast<int64_t>(std::floor(dmin)) - 1;
      k1 = static_cast<int64_t>(std::ceil(dmax)) + 1;
      k0 = std::max<int64_t>(k0, static_cast<int64_t>(plane_begin));
      k1 = std::min<int64_t>(k1, static_cast<int64_t>(plane_end) - 1);
    };
    bin_off.assign(np + 1, 0);
    for (uint64_t f = 0; f < F; ++f)
    {
      int64_t k0, k1;
      interval(f, k0, k1);
      for (int64_t k = k0; k <= k1; ++k)
        ++bin_off[k - plane_begin + 1];
    }
    for (uint64_t k = 0; k < np; ++k)
      bin_off[k + 1] += bin_off[k];
    bin_cells.resize(bin_off[np]);
    std::vector<uint64_t> cursor(bin_off.begin(), bin_off.end() - 1);
    for (uint64_t f = 0; f < F; ++f)
    {
      int64_t k0, k1;
      interval(f, k0, k1);
      for (int64_t k = k0; k <= k1; ++k)
        bin_cells[cursor[k - plane_begin]++] = static_cast<uint32_t>(f);
    }
  }

  std::vector<double> scalars;
  for (uint64_t i = plane_begin; i < plane_end; ++i)
  {
    const V3d origin = planeOrigin(L, i);
    CutOutput cut;
    if (scan_mode == ORC_SCAN_ALL_CELLS)
    {
      // vtkCutter evaluates the implicit function at every input point first
      scalars.resize(V);
      for (uint64_t v = 0; v < V; ++v)
        scalars[v] = planeScalar(L.cut_normal, origin, xyz + 3 * v);
      cutPlane(m, L, origin, nullptr, 0, true, scalars.data(), cut);
    }
    else
    {
      const uint64_t b = bin_off[i - plane_begin], e = bin_off[i - plane_begin + 1];
      cutPlane(m, L, origin, bin_cells.data() + b, e - b, false, nullptr, cut);
    }
    if (cut.lines.empty())
      continue;

    std::vector<IdList> segs;
    if (slice_mode == ORC_SLICE_FAITHFUL)
    {
      segs = strip(cut);
      unifySegmentDirection(cut, segs, ref);
      int rc = topologicalSortSegments(segs);
      if (rc != ORC_OK)
        return rc;
      rc = joinContiguousSegments(segs);
      if (rc != ORC_OK)
        return rc;
    }
    else
    {
      const int rc = maximalChains(cut, ref, segs);
      if (rc != ORC_OK)
        return rc;
    }
    Path path = convertToPoses(cut, segs, static_cast<int64_t>(i));
    if (!path.segments.empty())
      out.paths.push_back(std::move(path));
  }
  return ORC_OK;
}

}  // namespace orc
