"""TEST INFRASTRUCTURE ONLY — ctypes binding of the CPU oracle (oracle/liboracle.so).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this package.  The product package ``noether_b200`` never does.  See ``oracle/oracle.h`` for what each entry point
restates (reference file:line).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

MOMENTS_F32_SEQUENTIAL = 0
MOMENTS_F64 = 1
SLICE_FAITHFUL = 0
SLICE_INTENT = 1
SCAN_ALL_CELLS = 0
SCAN_BINNED = 1

OK = 0
ERR_INVALID = 1
ERR_TOO_MANY_CONNECTING = 2
ERR_NOT_A_DAG = 3
ERR_TOPO_SORT = 4
ERR_JOIN = 5
ERR_NON_MANIFOLD = 6
ERR_CLOSED_LOOP = 7
ERR_TOO_FEW_POINTS = 8
ERR_SEGMENT_TOO_SHORT = 9


class OracleError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile (gcc only, no reference sources involved)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h")) or f == "Makefile"]
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


class PcaResult(C.Structure):
    _fields_ = [("mean", C.c_float * 3), ("evecs", C.c_float * 9), ("evals", C.c_float * 3),
                ("scales", C.c_float * 3), ("cov", C.c_float * 9)]


class Lattice(C.Structure):
    _fields_ = [("cut_direction", C.c_double * 3), ("cut_normal", C.c_double * 3), ("cut_origin", C.c_double * 3),
                ("start_loc", C.c_double * 3), ("line_spacing", C.c_double), ("d_min", C.c_double),
                ("d_max", C.c_double), ("num_planes", C.c_uint64), ("no_cuts", C.c_int32), ("pad_", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    f32p = C.POINTER(C.c_float)
    f64p = C.POINTER(C.c_double)
    u32p = C.POINTER(C.c_uint32)
    u64p = C.POINTER(C.c_uint64)
    i64p = C.POINTER(C.c_int64)
    L.orc_last_error.restype = C.c_char_p
    L.orc_pca.argtypes = [f32p, C.c_uint64, C.c_int, C.POINTER(PcaResult)]
    L.orc_centroid.argtypes = [f32p, C.c_uint64, C.c_int, f64p]
    L.orc_aabb_center_origin.argtypes = [f32p, C.c_uint64, f64p]
    L.orc_principal_axis_direction.argtypes = [C.POINTER(PcaResult), C.c_double, f64p]
    L.orc_pca_rotated_direction.argtypes = [C.POINTER(PcaResult), C.c_double, f64p, f64p]
    L.orc_eigen_solver_3f.argtypes = [f32p, f32p, f32p]
    L.orc_make_lattice.argtypes = [C.POINTER(PcaResult), f64p, f64p, C.c_double, C.c_int, C.POINTER(Lattice)]
    L.orc_normals_from_mesh_faces.argtypes = [f32p, C.c_uint64, u32p, C.c_uint64, f32p]
    L.orc_slice.argtypes = [f32p, f32p, C.c_uint64, u32p, C.c_uint64, C.POINTER(Lattice), C.c_int, C.c_int,
                            C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p)]
    L.orc_paths_create.restype = C.c_void_p
    L.orc_paths_free.argtypes = [C.c_void_p]
    L.orc_paths_clone.argtypes = [C.c_void_p]
    L.orc_paths_clone.restype = C.c_void_p
    L.orc_paths_add_path.argtypes = [C.c_void_p]
    L.orc_paths_add_segment.argtypes = [C.c_void_p]
    L.orc_paths_add_waypoint.argtypes = [C.c_void_p, f64p]
    for n in ("orc_paths_num_paths", "orc_paths_num_segments", "orc_paths_num_waypoints"):
        getattr(L, n).argtypes = [C.c_void_p]
        getattr(L, n).restype = C.c_uint64
    L.orc_paths_export.argtypes = [C.c_void_p, u64p, u64p, i64p, f32p, f32p, u32p, u32p, f64p]
    L.orc_paths_export.restype = None
    for n in ("orc_raster_organization", "orc_direction_of_travel", "orc_snake_organization",
              "orc_uniform_orientation", "orc_concatenate"):
        getattr(L, n).argtypes = [C.c_void_p]
    for n in ("orc_linear_approach", "orc_linear_departure"):
        getattr(L, n).argtypes = [C.c_void_p, f64p, C.c_int]
    L.orc_offset.argtypes = [C.c_void_p, f64p]
    L.orc_moving_average_orientation_smoothing.argtypes = [C.c_void_p, C.c_int]
    for n in ("orc_circular_lead_in", "orc_circular_lead_out"):
        getattr(L, n).argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int]
    L.orc_fixed_orientation.argtypes = [C.c_void_p, f64p]
    for n in ("orc_tool_drag_orientation", "orc_biased_tool_drag_orientation"):
        getattr(L, n).argtypes = [C.c_void_p, C.c_double, C.c_double]
    for n in ("orc_join_close_segments", "orc_minimum_segment_length", "orc_uniform_spacing"):
        getattr(L, n).argtypes = [C.c_void_p, C.c_double]
    L.orc_read_ply.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    L.orc_read_stl.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    for n in ("orc_mesh_num_vertices", "orc_mesh_num_faces", "orc_mesh_num_indices"):
        getattr(L, n).argtypes = [C.c_void_p]
        getattr(L, n).restype = C.c_uint64
    L.orc_mesh_has_normals.argtypes = [C.c_void_p]
    L.orc_mesh_export.argtypes = [C.c_void_p, f32p, f32p, u32p, u32p]
    L.orc_mesh_export.restype = None
    L.orc_mesh_free.argtypes = [C.c_void_p]
    L.orc_mesh_free.restype = None
    for n in ("orc_face_midpoint_subdivision", "orc_face_subdivision_by_area"):
        getattr(L, n).argtypes = ([C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, u32p, C.c_uint64]
                                  + ([C.c_uint32] if "midpoint" in n else [C.c_float, C.c_uint64]) + [C.POINTER(C.c_void_p)])
    for n in ("orc_submesh_num_points", "orc_submesh_num_triangles"):
        getattr(L, n).argtypes = [C.c_void_p]
        getattr(L, n).restype = C.c_uint64
    L.orc_submesh_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_submesh_export.restype = None
    L.orc_submesh_free.argtypes = [C.c_void_p]
    L.orc_submesh_free.restype = None
    L.orc_euclidean_clustering.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, u32p, C.c_uint64, C.c_double, C.c_int,
                                           C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.orc_clusters_count.argtypes = [C.c_void_p]
    L.orc_clusters_count.restype = C.c_uint64
    L.orc_clusters_labels.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_clusters_labels.restype = None
    L.orc_cluster_info.argtypes = [C.c_void_p, C.c_uint64, u64p]
    L.orc_cluster_info.restype = None
    L.orc_cluster_export.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_cluster_export.restype = None
    L.orc_clusters_free.argtypes = [C.c_void_p]
    L.orc_clusters_free.restype = None
    L.orc_fixture_random_transform.argtypes = [C.c_double, C.c_double, f64p]
    L.orc_fixture_random_transform.restype = None
    L.orc_fixture_plane_mesh.argtypes = [C.c_float, C.c_float, f64p, f32p, f32p, u32p]
    L.orc_fixture_plane_mesh.restype = None
    _lib = L
    return L


def _check(rc: int):
    if rc != OK:
        raise OracleError(rc, lib().orc_last_error().decode())


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint32))


@dataclass
class Pca:
    mean: np.ndarray      # (3,) float32
    evecs: np.ndarray     # (3,3) float32, evecs[:, i] = i-th axis (descending eigenvalue)
    evals: np.ndarray     # (3,) float32 descending
    scales: np.ndarray    # (3,) float32
    cov: np.ndarray       # (3,3) float32
    raw: PcaResult


def pca(xyz, mode=MOMENTS_F32_SEQUENTIAL) -> Pca:
    a, p = _f32(xyz)
    r = PcaResult()
    _check(lib().orc_pca(p, a.shape[0], mode, C.byref(r)))
    return Pca(mean=np.array(r.mean, dtype=np.float32),
               evecs=np.array(r.evecs, dtype=np.float32).reshape(3, 3).T.copy(),
               evals=np.array(r.evals, dtype=np.float32), scales=np.array(r.scales, dtype=np.float32),
               cov=np.array(r.cov, dtype=np.float32).reshape(3, 3).T.copy(), raw=r)


def pca_from_parts(mean, evecs, scales) -> Pca:
    """Build a Pca from externally computed mean / eigenvectors (columns) / scales (e.g. the GPU's)."""
    r = PcaResult()
    ev = np.asarray(evecs, dtype=np.float32)
    for i in range(3):
        r.mean[i] = float(mean[i])
        r.scales[i] = float(scales[i])
        for row in range(3):
            r.evecs[i * 3 + row] = float(ev[row, i])
    return Pca(mean=np.asarray(mean, np.float32), evecs=ev, evals=np.zeros(3, np.float32),
               scales=np.asarray(scales, np.float32), cov=np.zeros((3, 3), np.float32), raw=r)


def centroid(xyz, mode=MOMENTS_F32_SEQUENTIAL) -> np.ndarray:
    a, p = _f32(xyz)
    out = np.zeros(3)
    _check(lib().orc_centroid(p, a.shape[0], mode, out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


def aabb_center_origin(xyz) -> np.ndarray:
    """AABBCenterOriginGenerator::generate (aabb_center_origin_generator.cpp:9-17)"""
    a, p = _f32(xyz)
    out = np.zeros(3)
    _check(lib().orc_aabb_center_origin(p, a.shape[0], out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


def principal_axis_direction(p: Pca, rotation_offset: float = 0.0) -> np.ndarray:
    out = np.zeros(3)
    _check(lib().orc_principal_axis_direction(C.byref(p.raw), rotation_offset,
                                              out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


def pca_rotated_direction(p: Pca, rotation_angle: float, nominal) -> np.ndarray:
    """PCARotatedDirectionGenerator::generate given the nested generator's direction."""
    _, v = _f64(nominal)
    out = np.zeros(3)
    _check(lib().orc_pca_rotated_direction(C.byref(p.raw), rotation_angle, v, out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


def eigen_solver_3f(cov) -> tuple[np.ndarray, np.ndarray]:
    c = np.ascontiguousarray(np.asarray(cov, dtype=np.float32).T)  # column-major
    evals = np.zeros(3, np.float32)
    evecs = np.zeros(9, np.float32)
    fp = C.POINTER(C.c_float)
    lib().orc_eigen_solver_3f(c.ctypes.data_as(fp), evals.ctypes.data_as(fp), evecs.ctypes.data_as(fp))
    return evals, evecs.reshape(3, 3).T.copy()


def make_lattice(p: Pca, cut_direction, cut_origin, line_spacing: float, bidirectional: bool = True) -> Lattice:
    _, d = _f64(cut_direction)
    _, o = _f64(cut_origin)
    L = Lattice()
    _check(lib().orc_make_lattice(C.byref(p.raw), d, o, line_spacing, int(bidirectional), C.byref(L)))
    return L


def normal_estimation_pcl(xyz, radius, viewpoint=(0.0, 0.0, 0.0)):
    """NormalEstimationPCLMeshModifier (normal_estimation_pcl.cpp:15-50): (normals (n,3) float32 with NaN rows where a radius
    search returns fewer than 3 points, curvature (n,) float32, neighbour counts (n,) uint32)"""
    a, ap = _f32(xyz)
    n = a.shape[0]
    nrm = np.zeros((n, 3), np.float32)
    cur = np.zeros(n, np.float32)
    cnt = np.zeros(n, np.uint32)
    L = lib()
    L.orc_normal_estimation_pcl.argtypes = [C.POINTER(C.c_float), C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_double,
                                            C.c_void_p, C.c_void_p, C.c_void_p]
    _check(L.orc_normal_estimation_pcl(ap, n, float(radius), float(viewpoint[0]), float(viewpoint[1]), float(viewpoint[2]),
                                       nrm.ctypes.data, cur.ctypes.data, cnt.ctypes.data))
    return nrm, cur, cnt


def normals_from_mesh_faces(xyz, tri) -> np.ndarray:
    a, ap = _f32(xyz)
    t, tp = _u32(tri)
    out = np.zeros((a.shape[0], 3), np.float32)
    _check(lib().orc_normals_from_mesh_faces(ap, a.shape[0], tp, t.shape[0], out.ctypes.data_as(C.POINTER(C.c_float))))
    return out


@dataclass
class FlatPaths:
    """noether::ToolPaths flattened: paths -> segments -> waypoints."""
    path_off: np.ndarray   # (P+1,) uint64 into segments
    seg_off: np.ndarray    # (S+1,) uint64 into waypoints
    path_plane: np.ndarray  # (P,) int64 lattice index
    pos: np.ndarray        # (W,3) float32
    nrm: np.ndarray        # (W,3) float32 (interpolated, not normalised)
    edge_lo: np.ndarray    # (W,) uint32
    edge_hi: np.ndarray    # (W,) uint32
    pose: np.ndarray       # (W,4,4) float64 (row, col)

    @property
    def num_paths(self):
        return len(self.path_off) - 1

    def segments_per_path(self):
        return np.diff(self.path_off).astype(np.int64)


class Paths:
    """Owning handle on an orc_paths (noether::ToolPaths + provenance)."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)

    def __del__(self):
        try:
            if self._h:
                lib().orc_paths_free(self._h)
        except Exception:
            pass

    @classmethod
    def from_nested(cls, nested):
        """nested: list of paths, each a list of segments, each an array (n,4,4) of poses."""
        L = lib()
        h = cls(L.orc_paths_create())
        for path in nested:
            L.orc_paths_add_path(h._h)
            for seg in path:
                L.orc_paths_add_segment(h._h)
                for T in seg:
                    m = np.ascontiguousarray(np.asarray(T, dtype=np.float64).T)  # column-major
                    L.orc_paths_add_waypoint(h._h, m.ctypes.data_as(C.POINTER(C.c_double)))
        return h

    def clone(self):
        return Paths(lib().orc_paths_clone(self._h))

    def flat(self) -> FlatPaths:
        L = lib()
        P = L.orc_paths_num_paths(self._h)
        S = L.orc_paths_num_segments(self._h)
        W = L.orc_paths_num_waypoints(self._h)
        path_off = np.zeros(P + 1, np.uint64)
        seg_off = np.zeros(S + 1, np.uint64)
        plane = np.zeros(P, np.int64)
        pos = np.zeros((W, 3), np.float32)
        nrm = np.zeros((W, 3), np.float32)
        lo = np.zeros(W, np.uint32)
        hi = np.zeros(W, np.uint32)
        pose = np.zeros((W, 16), np.float64)
        L.orc_paths_export(self._h, path_off.ctypes.data_as(C.POINTER(C.c_uint64)),
                           seg_off.ctypes.data_as(C.POINTER(C.c_uint64)),
                           plane.ctypes.data_as(C.POINTER(C.c_int64)),
                           pos.ctypes.data_as(C.POINTER(C.c_float)), nrm.ctypes.data_as(C.POINTER(C.c_float)),
                           lo.ctypes.data_as(C.POINTER(C.c_uint32)), hi.ctypes.data_as(C.POINTER(C.c_uint32)),
                           pose.ctypes.data_as(C.POINTER(C.c_double)))
        pose = pose.reshape(W, 4, 4).transpose(0, 2, 1).copy()
        return FlatPaths(path_off, seg_off, plane, pos, nrm, lo, hi, pose)

    def nested_poses(self):
        f = self.flat()
        out = []
        for p in range(f.num_paths):
            path = []
            for s in range(int(f.path_off[p]), int(f.path_off[p + 1])):
                path.append(f.pose[int(f.seg_off[s]):int(f.seg_off[s + 1])])
            out.append(path)
        return out

    # RasterPlanner::plan post-ops and the legacy modifiers, in place
    def raster_organization(self):
        _check(lib().orc_raster_organization(self._h))
        return self

    def direction_of_travel(self):
        _check(lib().orc_direction_of_travel(self._h))
        return self

    def join_close_segments(self, distance):
        _check(lib().orc_join_close_segments(self._h, distance))
        return self

    def minimum_segment_length(self, min_length):
        _check(lib().orc_minimum_segment_length(self._h, min_length))
        return self

    def uniform_spacing(self, point_spacing):
        _check(lib().orc_uniform_spacing(self._h, point_spacing))
        return self

    # tool-path modifiers chained after a raster plan, in place
    def snake_organization(self):
        _check(lib().orc_snake_organization(self._h))
        return self

    def linear_approach(self, offset, n_points):
        v = np.ascontiguousarray(offset, np.float64)
        _check(lib().orc_linear_approach(self._h, v.ctypes.data_as(C.POINTER(C.c_double)), int(n_points)))
        return self

    def linear_departure(self, offset, n_points):
        v = np.ascontiguousarray(offset, np.float64)
        _check(lib().orc_linear_departure(self._h, v.ctypes.data_as(C.POINTER(C.c_double)), int(n_points)))
        return self

    def offset(self, T):
        m = np.ascontiguousarray(np.asarray(T, np.float64).T)  # column-major
        _check(lib().orc_offset(self._h, m.ctypes.data_as(C.POINTER(C.c_double))))
        return self

    def fixed_orientation(self, ref_x):
        v = np.ascontiguousarray(ref_x, np.float64)
        _check(lib().orc_fixed_orientation(self._h, v.ctypes.data_as(C.POINTER(C.c_double))))
        return self

    def uniform_orientation(self):
        _check(lib().orc_uniform_orientation(self._h))
        return self

    def concatenate(self):
        _check(lib().orc_concatenate(self._h))
        return self

    def tool_drag_orientation(self, angle_offset, tool_radius):
        _check(lib().orc_tool_drag_orientation(self._h, angle_offset, tool_radius))
        return self

    def circular_lead_in(self, arc_angle, arc_radius, n_points):
        _check(lib().orc_circular_lead_in(self._h, arc_angle, arc_radius, int(n_points)))
        return self

    def circular_lead_out(self, arc_angle, arc_radius, n_points):
        _check(lib().orc_circular_lead_out(self._h, arc_angle, arc_radius, int(n_points)))
        return self

    def moving_average_orientation_smoothing(self, window_size):
        _check(lib().orc_moving_average_orientation_smoothing(self._h, int(window_size)))
        return self

    def biased_tool_drag_orientation(self, angle_offset, tool_radius):
        _check(lib().orc_biased_tool_drag_orientation(self._h, angle_offset, tool_radius))
        return self


def slice_mesh(xyz, normals, tri, lattice: Lattice, slice_mode=SLICE_FAITHFUL, scan_mode=SCAN_ALL_CELLS,
               plane_begin=0, plane_end=None) -> Paths:
    """PlaneSlicerRasterPlanner::planImpl's plane loop (plane_slicer_raster_planner.cpp:596-628)."""
    a, ap = _f32(xyz)
    n, npn = _f32(normals)
    t, tp = _u32(tri)
    if plane_end is None:
        plane_end = lattice.num_planes
    out = C.c_void_p()
    _check(lib().orc_slice(ap, npn, a.shape[0], tp, t.shape[0], C.byref(lattice), slice_mode, scan_mode,
                           plane_begin, plane_end, C.byref(out)))
    return Paths(out.value)


def plan_plane_slicer(xyz, normals, tri, line_spacing, bidirectional=True, direction=None, origin=None,
                      rotation_offset=0.0, moments_mode=MOMENTS_F32_SEQUENTIAL, slice_mode=SLICE_FAITHFUL,
                      scan_mode=SCAN_ALL_CELLS, post_ops=True, frame: Pca | None = None):
    """RasterPlanner::plan for PlaneSlicerRasterPlanner (raster_planner.cpp:16-35).

    direction=None -> PrincipalAxisDirectionGenerator(rotation_offset); origin=None -> CentroidOriginGenerator.
    frame: use this PCA frame (mean, axes, scales — 15 floats) instead of accumulating moments here; the centroid
    generator then returns the frame's mean.  Lets a test hand the oracle exactly the frame another implementation
    computed, so that everything downstream of those 15 scalars can be compared bit for bit.
    Returns (Paths, Lattice, Pca).
    """
    p = pca(xyz, moments_mode) if frame is None else frame
    d = principal_axis_direction(p, rotation_offset) if direction is None else np.asarray(direction, np.float64)
    if origin is not None:
        o = np.asarray(origin, np.float64)
    elif frame is not None:
        o = np.asarray(frame.mean, np.float32).astype(np.float64)
    else:
        o = centroid(xyz, moments_mode)
    lat = make_lattice(p, d, o, line_spacing, bidirectional)
    paths = slice_mesh(xyz, normals, tri, lat, slice_mode, scan_mode)
    if post_ops and paths.flat().num_paths > 0:
        paths.raster_organization().direction_of_travel()
    return paths, lat, p


def plan_plane_slicer_legacy(xyz, normals, tri, line_spacing, point_spacing, min_segment_size, min_hole_size,
                             bidirectional=True, direction=None, origin=None, **kw):
    """PlaneSlicerLegacyRasterPlanner (plane_slicer_legacy_raster_planner.cpp:22-33) through RasterPlanner::plan."""
    paths, lat, p = plan_plane_slicer(xyz, normals, tri, line_spacing, bidirectional, direction, origin,
                                      post_ops=False, **kw)
    paths.join_close_segments(min_hole_size).minimum_segment_length(min_segment_size).uniform_spacing(point_spacing)
    if paths.flat().num_paths > 0:
        paths.raster_organization().direction_of_travel()
    return paths, lat, p


def fixture_random_transform(translation_limit, rotation_limit) -> np.ndarray:
    out = np.zeros(16)
    lib().orc_fixture_random_transform(translation_limit, rotation_limit, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out.reshape(4, 4).T.copy()


def fixture_plane_mesh(lx, ly, T=None):
    T = np.eye(4) if T is None else np.asarray(T, np.float64)
    m = np.ascontiguousarray(T.T)
    xyz = np.zeros((4, 3), np.float32)
    nrm = np.zeros((4, 3), np.float32)
    tri = np.zeros((2, 3), np.uint32)
    lib().orc_fixture_plane_mesh(lx, ly, m.ctypes.data_as(C.POINTER(C.c_double)),
                                 xyz.ctypes.data_as(C.POINTER(C.c_float)), nrm.ctypes.data_as(C.POINTER(C.c_float)),
                                 tri.ctypes.data_as(C.POINTER(C.c_uint32)))
    return xyz, nrm, tri


def read_stl(data: bytes):
    """pcl::io::loadPolygonFile for a binary STL image (vtkSTLReader with Merging + vtk2mesh restated): returns
    (xyz float32 (V,3), triangles uint32 (F,3))."""
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    h = C.c_void_p()
    _check(lib().orc_read_stl(buf.ctypes.data, buf.size, C.byref(h)))
    try:
        nv, ni = lib().orc_mesh_num_vertices(h), lib().orc_mesh_num_indices(h)
        xyz = np.zeros((nv, 3), np.float32)
        idx = np.zeros(ni, np.uint32)
        lib().orc_mesh_export(h, xyz.ctypes.data_as(C.POINTER(C.c_float)), None, None, idx.ctypes.data_as(C.POINTER(C.c_uint32)))
    finally:
        lib().orc_mesh_free(h)
    return xyz, idx.reshape(-1, 3)


def read_ply(data: bytes):
    """pcl::io::loadPolygonFile for a PLY image (vtkPLYReader + vtk2mesh restated, oracle_ply.cpp): returns
    (xyz float32 (V,3), normals float32 (V,3) or None, face_sizes uint32 (F,), indices uint32 (sum of sizes,))."""
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    h = C.c_void_p()
    _check(lib().orc_read_ply(buf.ctypes.data, buf.size, C.byref(h)))
    try:
        nv, nf, ni = lib().orc_mesh_num_vertices(h), lib().orc_mesh_num_faces(h), lib().orc_mesh_num_indices(h)
        xyz = np.zeros((nv, 3), np.float32)
        nrm = np.zeros((nv, 3), np.float32) if lib().orc_mesh_has_normals(h) else None
        fs = np.zeros(nf, np.uint32)
        idx = np.zeros(ni, np.uint32)
        lib().orc_mesh_export(h, xyz.ctypes.data_as(C.POINTER(C.c_float)),
                              nrm.ctypes.data_as(C.POINTER(C.c_float)) if nrm is not None else None,
                              fs.ctypes.data_as(C.POINTER(C.c_uint32)), idx.ctypes.data_as(C.POINTER(C.c_uint32)))
    finally:
        lib().orc_mesh_free(h)
    return xyz, nrm, fs, idx


def _subdivide(call, cloud, point_step, off_xyz, off_normal, off_rgba, tri, *args):
    blob = np.ascontiguousarray(np.frombuffer(bytes(cloud), dtype=np.uint8))
    n = blob.size // point_step
    t, tp = _u32(np.asarray(tri).reshape(-1, 3))
    h = C.c_void_p()
    _check(call(blob.ctypes.data, n, point_step, off_xyz, off_normal, off_rgba, tp, t.shape[0], *args, C.byref(h)))
    try:
        nv, nf = lib().orc_submesh_num_points(h), lib().orc_submesh_num_triangles(h)
        out = np.zeros(nv * point_step, np.uint8)
        idx = np.zeros((nf, 3), np.uint32)
        lib().orc_submesh_export(h, out.ctypes.data, idx.ctypes.data)
    finally:
        lib().orc_submesh_free(h)
    return out, idx


def face_midpoint_subdivision(cloud, point_step, off_xyz, off_normal, off_rgba, tri, n_iterations):
    """FaceMidpointSubdivisionMeshModifier::modify (face_subdivision_modifier.cpp:202-248) on a raw PCLPointCloud2 blob:
    returns (new blob uint8, triangles uint32 (F', 3))."""
    return _subdivide(lib().orc_face_midpoint_subdivision, cloud, point_step, off_xyz, off_normal, off_rgba, tri,
                      int(n_iterations))


def face_subdivision_by_area(cloud, point_step, off_xyz, off_normal, off_rgba, tri, max_area, max_points=0):
    """FaceSubdivisionByAreaMeshModifier::modify (face_subdivision_modifier.cpp:250-339)."""
    return _subdivide(lib().orc_face_subdivision_by_area, cloud, point_step, off_xyz, off_normal, off_rgba, tri,
                      C.c_float(max_area), int(max_points))


@dataclass
class ClusterMesh:
    indices: np.ndarray      # the cluster (ascending point indices)
    cloud: np.ndarray        # uint8 records of the sub-mesh (empty when `empty`)
    tri: np.ndarray          # (F', 3) uint32
    whole_input: bool        # every point is used: the reference returns the input cloud unchanged
    empty: bool              # no face survives: the reference returns a default-constructed mesh


def euclidean_clustering(cloud, point_step, off_xyz, tri, tolerance, min_cluster_size=1, max_cluster_size=-1, brute=False):
    """EuclideanClusteringMeshModifier::modify (euclidean_clustering_modifier.cpp:40-64): returns (list of ClusterMesh in
    the reference's output order, labels uint32 (V,) = smallest index of every point's connected component)."""
    blob = np.ascontiguousarray(np.frombuffer(bytes(cloud), dtype=np.uint8))
    n = blob.size // point_step
    t, tp = _u32(np.asarray(tri).reshape(-1, 3))
    h = C.c_void_p()
    _check(lib().orc_euclidean_clustering(blob.ctypes.data, n, point_step, off_xyz, tp, t.shape[0], float(tolerance),
                                          int(min_cluster_size), int(max_cluster_size), int(bool(brute)), C.byref(h)))
    try:
        labels = np.zeros(n, np.uint32)
        lib().orc_clusters_labels(h, labels.ctypes.data)
        out = []
        for k in range(lib().orc_clusters_count(h)):
            info = (C.c_uint64 * 5)()
            lib().orc_cluster_info(h, k, info)
            idx = np.zeros(info[0], np.uint32)
            pts = np.zeros(info[1] * point_step, np.uint8)
            tr = np.zeros((info[2], 3), np.uint32)
            lib().orc_cluster_export(h, k, idx.ctypes.data, pts.ctypes.data, tr.ctypes.data)
            out.append(ClusterMesh(idx, pts, tr, bool(info[3]), bool(info[4])))
    finally:
        lib().orc_clusters_free(h)
    return out, labels
