"""Host-side mirror of the reference's component interface for the raster-slicing path, over the C ABI.

Same names, argument meaning and error behaviour as noether_tpp 0.16.0:

  PlaneSlicerRasterPlanner / PlaneSlicerLegacyRasterPlanner   tool_path_planners/raster/plane_slicer_raster_planner.h,
                                                             plane_slicer_legacy_raster_planner.h; RasterPlanner::plan
  PrincipalAxisDirectionGenerator, FixedDirectionGenerator    raster/direction_generators/*.h (DirectionGenerator::generate)
  CentroidOriginGenerator, FixedOriginGenerator               raster/origin_generators/*.h (OriginGenerator::generate)
  NormalsFromMeshFacesMeshModifier                            mesh_modifiers/normals_from_mesh_faces_modifier.h
  Factory.create_*(config)                                    plugin_interface.h:106-129, plugins.cpp (aliases + YAML keys)

The production drop-in is the C++ boost_plugin_loader shim in plugin/b200_plugins.cpp; this module is the same thing
for Python callers, tests and the benchmark.  Everything computes on the GPU through libnoether_b200.so; nothing here
falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi
from .meshes import MeshBlob


class Context:
    """One nb2_context (one GPU)."""

    def __init__(self, device: int = 0):
        self._lib = _capi.load()
        h = C.c_void_p()
        _capi.check(self._lib.nb2_context_create(device, C.byref(h)))
        self._h = h
        self._mesh_key = None
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.nb2_context_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- mesh -------------------------------------------------------------------------------------------------
    def upload(self, mesh: MeshBlob, force: bool = False, sharded: bool = False, geometry_only: bool = False):
        """nb2_mesh_upload; sharded=True (collective, after comm_init): every rank copies 1 / world of the bytes and the
        shares are all-gathered over NVLink (nb2_mesh_upload_sharded).  geometry_only=True (NB2_MESH_GEOMETRY_ONLY): only
        positions and normals will be used on the device, so the host cores compact the cloud to 12- / 24-byte records on
        its way through the pinned staging ring (callers that work on whole records — the face subdivisions — leave it off)."""
        # (a resident mesh serves every caller that needs positions and normals, compacted or not; callers that need whole
        #  records force their upload)
        key = (id(mesh.data), id(mesh.tri), mesh.n_vertices, mesh.point_step, mesh.off_xyz, mesh.off_normal)
        if not force and key == self._mesh_key:
            return
        self._mesh_key = None
        data = np.ascontiguousarray(mesh.data)
        tri = np.ascontiguousarray(mesh.tri, dtype=np.uint32)
        mv = _capi.MeshView(data.ctypes.data, mesh.n_vertices, mesh.point_step, mesh.off_xyz, mesh.off_normal,
                            _capi.NB2_MESH_GEOMETRY_ONLY if geometry_only else 0,
                            tri.ctypes.data if tri.size else None, tri.shape[0])
        fn = self._lib.nb2_mesh_upload_sharded if sharded else self._lib.nb2_mesh_upload
        _capi.check(fn(self._h, C.byref(mv)))
        self._mesh_key = key
        self._keep = (data, tri)

    def upload_raw(self, data_ptr: int, n_vertices: int, point_step: int, off_xyz: int, off_normal: int,
                   tri_ptr: int, n_triangles: int):
        """Upload from raw host pointers (e.g. pinned buffers owned by the caller)."""
        self._mesh_key = None
        mv = _capi.MeshView(data_ptr, n_vertices, point_step, off_xyz, off_normal, 0, tri_ptr, n_triangles)
        _capi.check(self._lib.nb2_mesh_upload(self._h, C.byref(mv)))

    def upload_ply(self, source) -> _capi.MeshFileInfo:
        """pcl::io::loadPolygonFile + upload for a binary little-endian PLY (noether_gui/src/widgets/tpp_widget.cpp:426):
        `source` is a path or the bytes of the file; the records are decoded on the device (nb2_mesh_upload_ply)."""
        self._mesh_key = None
        info = _capi.MeshFileInfo()
        if isinstance(source, (bytes, bytearray, memoryview, np.ndarray)):
            img = np.frombuffer(source, dtype=np.uint8) if not isinstance(source, np.ndarray) else np.ascontiguousarray(source).view(np.uint8)
            _capi.check(self._lib.nb2_mesh_upload_ply(self._h, img.ctypes.data, img.size, C.byref(info)))
        else:
            _capi.check(self._lib.nb2_mesh_upload_ply_file(self._h, os.fspath(source).encode(), C.byref(info)))
        return info

    def upload_stl(self, source) -> _capi.MeshFileInfo:
        """pcl::io::loadPolygonFile + upload for a binary STL: `source` is a path or the bytes of the file; the facet
        corners are welded on the device as vtkSTLReader's point merging does (nb2_mesh_upload_stl)."""
        self._mesh_key = None
        info = _capi.MeshFileInfo()
        if isinstance(source, (bytes, bytearray, memoryview, np.ndarray)):
            img = np.frombuffer(source, dtype=np.uint8) if not isinstance(source, np.ndarray) else np.ascontiguousarray(source).view(np.uint8)
            _capi.check(self._lib.nb2_mesh_upload_stl(self._h, img.ctypes.data, img.size, C.byref(info)))
        else:
            _capi.check(self._lib.nb2_mesh_upload_file(self._h, os.fspath(source).encode(), C.byref(info)))
        return info

    def upload_file(self, path) -> _capi.MeshFileInfo:
        """pcl::io::loadPolygonFile(path, mesh): .ply or .stl by extension (nb2_mesh_upload_file)"""
        self._mesh_key = None
        info = _capi.MeshFileInfo()
        _capi.check(self._lib.nb2_mesh_upload_file(self._h, os.fspath(path).encode(), C.byref(info)))
        return info

    def download_triangles(self, n_triangles: int) -> np.ndarray:
        tri = np.zeros((n_triangles, 3), np.uint32)
        _capi.check(self._lib.nb2_mesh_download_triangles(self._h, tri.ctypes.data, int(n_triangles)))
        return tri

    def download_cloud(self, n_vertices: int) -> np.ndarray:
        """the resident cloud as PCLPointCloud2::data (n_vertices records of point_step bytes)"""
        step, ox, on = self.record_layout()
        data = np.zeros(n_vertices * step, np.uint8)
        _capi.check(self._lib.nb2_mesh_download_cloud(self._h, data.ctypes.data, int(data.size)))
        return data

    def record_layout(self) -> tuple[int, int, int]:
        step, ox, on = C.c_uint32(), C.c_uint32(), C.c_int32()
        _capi.check(self._lib.nb2_mesh_record_layout(self._h, C.byref(step), C.byref(ox), C.byref(on)))
        return step.value, ox.value, on.value

    def face_midpoint_subdivision(self, n_iterations: int, off_rgba: int = -1) -> _capi.MeshFileInfo:
        """nb2_face_midpoint_subdivision on the resident mesh; the subdivided mesh becomes the resident one"""
        self._mesh_key = None
        info = _capi.MeshFileInfo()
        _capi.check(self._lib.nb2_face_midpoint_subdivision(self._h, int(n_iterations), int(off_rgba), C.byref(info)))
        return info

    def face_subdivision_by_area(self, max_area: float, off_rgba: int = -1, max_points: int = 0) -> _capi.MeshFileInfo:
        """nb2_face_subdivision_by_area on the resident mesh; the subdivided mesh becomes the resident one"""
        self._mesh_key = None
        info = _capi.MeshFileInfo()
        _capi.check(self._lib.nb2_face_subdivision_by_area(self._h, float(max_area), int(off_rgba), int(max_points), C.byref(info)))
        return info

    def euclidean_cluster_labels(self, tolerance: float, n_vertices: int) -> np.ndarray:
        """nb2_euclidean_cluster_labels: smallest point index of every point's connected component"""
        labels = np.zeros(n_vertices, np.uint32)
        _capi.check(self._lib.nb2_euclidean_cluster_labels(self._h, float(tolerance), labels.ctypes.data, int(n_vertices)))
        return labels

    def set_normals(self, normals: np.ndarray):
        n = np.ascontiguousarray(normals, dtype=np.float32)
        _capi.check(self._lib.nb2_mesh_set_normals(self._h, n.ctypes.data))

    def download(self, n_vertices: int):
        xyz = np.zeros((n_vertices, 3), np.float32)
        nrm = np.zeros((n_vertices, 3), np.float32)
        _capi.check(self._lib.nb2_mesh_download(self._h, xyz.ctypes.data, nrm.ctypes.data))
        return xyz, nrm

    # -- frame ------------------------------------------------------------------------------------------------
    def pca(self) -> _capi.Pca:
        out = _capi.Pca()
        _capi.check(self._lib.nb2_mesh_pca(self._h, C.byref(out)))
        return out

    def principal_axis_direction(self, rotation_offset: float) -> np.ndarray:
        out = (C.c_double * 3)()
        _capi.check(self._lib.nb2_principal_axis_direction(self._h, rotation_offset, out))
        return np.array(out)

    def pca_rotated_direction(self, rotation_angle: float, nominal) -> np.ndarray:
        v = (C.c_double * 3)(*[float(x) for x in nominal])
        out = (C.c_double * 3)()
        _capi.check(self._lib.nb2_pca_rotated_direction(self._h, rotation_angle, v, out))
        return np.array(out)

    def centroid_origin(self) -> np.ndarray:
        out = (C.c_double * 3)()
        _capi.check(self._lib.nb2_centroid_origin(self._h, out))
        return np.array(out)

    def normals_from_mesh_faces(self, n_vertices: int, download: bool = True):
        out = np.zeros((n_vertices, 3), np.float32) if download else None
        _capi.check(self._lib.nb2_normals_from_mesh_faces(self._h, out.ctypes.data if download else None))
        return out

    def normal_estimation_pcl(self, n_vertices: int, radius: float, viewpoint=(0.0, 0.0, 0.0)):
        """nb2_normal_estimation_pcl on the resident cloud: (normals (n,3) float32 — NaN rows where the radius search finds
        fewer than 3 points — and curvature (n,) float32); the estimated normals become the resident ones"""
        nrm = np.zeros((n_vertices, 3), np.float32)
        cur = np.zeros(n_vertices, np.float32)
        _capi.check(self._lib.nb2_normal_estimation_pcl(self._h, C.c_double(radius), C.c_double(viewpoint[0]), C.c_double(viewpoint[1]),
                                                        C.c_double(viewpoint[2]), nrm.ctypes.data_as(C.c_void_p),
                                                        cur.ctypes.data_as(C.c_void_p)))
        return nrm, cur

    # -- planning ---------------------------------------------------------------------------------------------
    def plan(self, params: _capi.PlanParams) -> "ToolPaths":
        r = C.c_void_p()
        _capi.check(self._lib.nb2_plan(self._h, C.byref(params), C.byref(r)))
        return ToolPaths(self, r)

    def plan_mesh_raw(self, data_ptr: int, n_vertices: int, point_step: int, off_xyz: int, off_normal: int,
                      tri_ptr: int, n_triangles: int, params: _capi.PlanParams) -> "ToolPaths":
        """nb2_plan_mesh: upload (or adopt, for device pointers) and plan in one call — planImpl(mesh)."""
        self._mesh_key = None
        mv = _capi.MeshView(data_ptr, n_vertices, point_step, off_xyz, off_normal, 0, tri_ptr, n_triangles)
        r = C.c_void_p()
        _capi.check(self._lib.nb2_plan_mesh(self._h, C.byref(mv), C.byref(params), C.byref(r)))
        return ToolPaths(self, r, layout_only=params.download == 2)

    def slice(self, lattice: _capi.Lattice, plane_begin: int = 0, plane_end: int | None = None) -> "ToolPaths":
        r = C.c_void_p()
        pe = _capi.NB2_ALL_PLANES if plane_end is None else plane_end
        _capi.check(self._lib.nb2_slice(self._h, C.byref(lattice), plane_begin, pe, C.byref(r)))
        return ToolPaths(self, r)

    def last_lattice(self) -> _capi.Lattice:
        out = _capi.Lattice()
        _capi.check(self._lib.nb2_last_lattice(self._h, C.byref(out)))
        return out

    def timings(self) -> dict:
        names = (C.c_char_p * 32)()
        ms = (C.c_float * 32)()
        n = self._lib.nb2_last_timings(self._h, names, ms, 32)
        return {names[i].decode(): float(ms[i]) for i in range(n)}

    def kernel_launches(self) -> int:
        return int(self._lib.nb2_kernel_launches(self._h))

    def link_profile(self) -> list[int]:
        """cycles per phase of the linker's fast path in the last plan (NB2_LINK_PROFILE=1 in the environment)"""
        out = (C.c_uint64 * 16)()
        _capi.check(self._lib.nb2_link_profile(self._h, out))
        return [int(v) for v in out]

    def general_path_planes(self) -> int:
        """planes of the last plan whose cut lines the linker's general path joined (diagnostics)"""
        return int(self._lib.nb2_last_plan_general_path_planes(self._h))

    def timer_start(self):
        _capi.check(self._lib.nb2_timer_start(self._h))

    def timer_stop(self) -> float:
        """Elapsed device time (ms) since timer_start, from CUDA events on the context's stream."""
        ms = C.c_float()
        _capi.check(self._lib.nb2_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def last_plan_span_ms(self) -> float:
        """Device time (ms) of the last plan on this context: CUDA events at the entry of the planning call and behind its
        last slicing kernel (nb2_last_plan_span)."""
        ms = C.c_float()
        _capi.check(self._lib.nb2_last_plan_span(self._h, C.byref(ms)))
        return float(ms.value)

    # -- multi-GPU --------------------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * _capi.NB2_UNIQUE_ID_BYTES)()
        _capi.check(_capi.load().nb2_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, world_size: int):
        buf = (C.c_uint8 * _capi.NB2_UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        _capi.check(self._lib.nb2_comm_init(self._h, buf, rank, world_size))

    def gather(self, local: "ToolPaths", root: int = 0) -> "ToolPaths | None":
        out = C.c_void_p()
        _capi.check(self._lib.nb2_result_gather(self._h, local._r, root, C.byref(out)))
        return ToolPaths(self, out) if out.value else None


def stl_inspect(data) -> _capi.MeshFileInfo:
    """facets of a binary STL image and the point count before welding (host only, nb2_stl_inspect)"""
    img = np.frombuffer(data, dtype=np.uint8)
    info = _capi.MeshFileInfo()
    _capi.check(_capi.load().nb2_stl_inspect(img.ctypes.data, img.size, C.byref(info)))
    return info


def ply_inspect(data) -> _capi.MeshFileInfo:
    """nb2_ply_inspect: vertex / triangle counts and normals flag of a binary little-endian PLY image (host only)."""
    img = np.frombuffer(bytes(data), dtype=np.uint8)
    info = _capi.MeshFileInfo()
    _capi.check(_capi.load().nb2_ply_inspect(img.ctypes.data, img.size, C.byref(info)))
    return info


def make_lattice(pca: _capi.Pca, direction, origin, line_spacing: float, bidirectional: bool = True) -> _capi.Lattice:
    d = (C.c_double * 3)(*[float(v) for v in direction])
    o = (C.c_double * 3)(*[float(v) for v in origin])
    out = _capi.Lattice()
    _capi.check(_capi.load().nb2_make_lattice(C.byref(pca), d, o, line_spacing, int(bidirectional), C.byref(out)))
    return out


class ToolPaths:
    """noether::ToolPaths as returned by the C ABI: SoA arrays (views into the result's pinned block)."""

    def __init__(self, ctx: Context, handle, layout_only: bool = False):
        """layout_only: do not wait for the waypoint arrays of a download = 2 result (nb2_result_layout): counts and offsets
        are valid at once, `positions` / `normals` after wait(); poses() consumes them as they land."""
        self._ctx = ctx
        self._lib = ctx._lib
        self._r = handle
        v = _capi.ResultView()
        _capi.check((self._lib.nb2_result_layout if layout_only else self._lib.nb2_result_get)(self._r, C.byref(v)))
        self.n_paths, self.n_segments, self.n_waypoints = int(v.n_paths), int(v.n_segments), int(v.n_waypoints)
        P, S, W = self.n_paths, self.n_segments, self.n_waypoints

        def arr(ptr, n, dt):
            if n == 0 or not ptr:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(ptr, shape=(n,)).view(dt)

        self.path_offsets = arr(v.path_offsets, P + 1, np.uint32)
        self.segment_offsets = arr(v.segment_offsets, S + 1, np.uint32)
        self.path_plane = arr(v.path_plane, P, np.int64)
        self.positions = arr(v.positions, 3 * W, np.float32).reshape(-1, 3)
        self.normals = arr(v.normals, 3 * W, np.float32).reshape(-1, 3)
        self.edge_lo = arr(v.edge_lo, W, np.uint32) if v.edge_lo else None
        self.edge_hi = arr(v.edge_hi, W, np.uint32) if v.edge_hi else None
        self.post_ops_applied = bool(v.raster_post_ops_applied)
        self.legacy = bool(v.legacy)

    @classmethod
    def from_nested(cls, nested, context: "Context | None" = None) -> "ToolPaths":
        """noether::ToolPaths given as list[tool path] of list[segment] of (n, 4, 4) poses (nb2_result_from_poses)."""
        ctx = context or default_context()
        path_off = [0]
        seg_off = [0]
        chunks = []
        for path in nested:
            for seg in path:
                a = np.asarray(seg, np.float64).reshape(-1, 4, 4)
                chunks.append(a.transpose(0, 2, 1).reshape(-1, 16))  # column-major storage
                seg_off.append(seg_off[-1] + len(a))
            path_off.append(len(seg_off) - 1)
        poses = np.ascontiguousarray(np.concatenate(chunks) if chunks else np.zeros((0, 16)), np.float64)
        po = np.asarray(path_off, np.uint32)
        so = np.asarray(seg_off, np.uint32)
        h = C.c_void_p()
        _capi.check(ctx._lib.nb2_result_from_poses(ctx._h, len(po) - 1, po.ctypes.data, len(so) - 1, so.ctypes.data,
                                                   poses.ctypes.data if len(poses) else None, C.byref(h)))
        return cls(ctx, h)

    def wait(self):
        """block until every waypoint of a download = 2 result is in host memory (nb2_result_get)"""
        v = _capi.ResultView()
        _capi.check(self._lib.nb2_result_get(self._r, C.byref(v)))

    def free(self):
        if self._r:
            self._lib.nb2_result_free(self._r)
            self._r = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def __len__(self):
        return self.n_paths

    def segments_per_path(self) -> np.ndarray:
        return np.diff(self.path_offsets.astype(np.int64))

    def poses(self) -> np.ndarray:
        """(W, 4, 4) float64, [row, col] — Eigen::Isometry3d matrices"""
        out = np.zeros((self.n_waypoints, 16), np.float64)
        _capi.check(self._lib.nb2_result_poses(self._r, out.ctypes.data))
        return out.reshape(-1, 4, 4).transpose(0, 2, 1).copy()

    def nested(self):
        """list[tool path] of list[segment] of (n, 4, 4) pose arrays — the shape of noether::ToolPaths"""
        poses = self.poses()
        out = []
        for p in range(self.n_paths):
            path = []
            for s in range(int(self.path_offsets[p]), int(self.path_offsets[p + 1])):
                path.append(poses[int(self.segment_offsets[s]):int(self.segment_offsets[s + 1])])
            out.append(path)
        return out


# ---------------------------------------------------------------------------------------------------------------------
# components (reference interface)
# ---------------------------------------------------------------------------------------------------------------------
_default_contexts: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_contexts:
        _default_contexts[device] = Context(device)
    return _default_contexts[device]


class DirectionGenerator:
    def generate(self, mesh: MeshBlob) -> np.ndarray:  # noqa: D401
        raise NotImplementedError


class OriginGenerator:
    def generate(self, mesh: MeshBlob) -> np.ndarray:
        raise NotImplementedError


class FixedDirectionGenerator(DirectionGenerator):
    """direction_generators/fixed_direction_generator.cpp: returns the configured direction (host only)"""

    def __init__(self, direction):
        self.direction = np.asarray(direction, np.float64)

    def generate(self, mesh):
        return self.direction


class FixedOriginGenerator(OriginGenerator):
    def __init__(self, origin=(0.0, 0.0, 0.0)):
        self.origin = np.asarray(origin, np.float64)

    def generate(self, mesh):
        return self.origin


class PrincipalAxisDirectionGenerator(DirectionGenerator):
    """principal_axis_direction_generator.cpp:14-26 — AngleAxis(rotation_offset, e2) * e1 from the device PCA"""

    def __init__(self, rotation_offset: float = 0.0, context: Context | None = None):
        self.rotation_offset = float(rotation_offset)
        self._ctx = context

    def generate(self, mesh):
        ctx = self._ctx or default_context()
        ctx.upload(mesh)
        return ctx.principal_axis_direction(self.rotation_offset)


class PCARotatedDirectionGenerator(DirectionGenerator):
    """pca_rotated_direction_generator.cpp:13-25 — a nested generator's direction rotated about the smallest principal
    axis (device PCA)"""

    def __init__(self, dir_gen: DirectionGenerator, rotation_angle: float, context: Context | None = None):
        self.dir_gen = dir_gen
        self.rotation_angle = float(rotation_angle)
        self._ctx = context

    def generate(self, mesh):
        nominal = self.dir_gen.generate(mesh)
        ctx = self._ctx or default_context()
        ctx.upload(mesh)
        return ctx.pca_rotated_direction(self.rotation_angle, nominal)


class CentroidOriginGenerator(OriginGenerator):
    """centroid_origin_generator.cpp:8-17 — mean of the vertices from the device moment reduction"""

    def __init__(self, context: Context | None = None):
        self._ctx = context

    def generate(self, mesh):
        ctx = self._ctx or default_context()
        ctx.upload(mesh)
        return ctx.centroid_origin()


class AABBCenterOriginGenerator(OriginGenerator):
    """aabb_center_origin_generator.cpp:9-17 — (max - min) / 2 of the float AABB (the reference's own arithmetic: half
    the range of the box), from the AABB the ingest pass leaves on the device"""

    def __init__(self, context: Context | None = None):
        self._ctx = context

    def generate(self, mesh):
        ctx = self._ctx or default_context()
        ctx.upload(mesh)
        out = (C.c_double * 3)()
        _capi.check(ctx._lib.nb2_aabb_center_origin(ctx._h, out))
        return np.array(out[:], np.float64)


class PlaneSlicerRasterPlanner:
    """RasterPlanner::plan for the plane slicer (raster_planner.cpp:16-35, plane_slicer_raster_planner.cpp:518-631)"""

    def __init__(self, dir_gen: DirectionGenerator, origin_gen: OriginGenerator, context: Context | None = None):
        self.dir_gen = dir_gen
        self.origin_gen = origin_gen
        self.line_spacing = None  # the reference leaves it unset until setLineSpacing (raster_planner.h:85)
        self.bidirectional = True
        self._ctx = context
        self.emit_edges = True
        self.raster_post_ops = True

    def set_line_spacing(self, line_spacing: float):
        self.line_spacing = float(line_spacing)

    def generate_rasters_bidirectionally(self, bidirectional: bool):
        self.bidirectional = bool(bidirectional)

    def _params(self, mesh) -> _capi.PlanParams:
        if self.line_spacing is None:
            raise ValueError("line_spacing has not been set (RasterPlanner::setLineSpacing)")
        p = _capi.PlanParams()
        _capi.load().nb2_plan_params_default(C.byref(p))
        p.line_spacing = self.line_spacing
        p.bidirectional = int(self.bidirectional)
        p.emit_edges = int(self.emit_edges)
        p.raster_post_ops = int(self.raster_post_ops)
        if isinstance(self.dir_gen, PrincipalAxisDirectionGenerator):
            p.direction_mode = _capi.NB2_DIRECTION_PRINCIPAL_AXIS
            p.rotation_offset = self.dir_gen.rotation_offset
        else:
            p.direction_mode = _capi.NB2_DIRECTION_EXPLICIT
            d = self.dir_gen.generate(mesh)
            p.direction[:] = [float(v) for v in d]
        if isinstance(self.origin_gen, CentroidOriginGenerator):
            p.origin_mode = _capi.NB2_ORIGIN_CENTROID
        else:
            p.origin_mode = _capi.NB2_ORIGIN_EXPLICIT
            o = self.origin_gen.generate(mesh)
            p.origin[:] = [float(v) for v in o]
        return p

    def plan(self, mesh: MeshBlob, shard_rank: int = 0, shard_count: int = 1, sharded_upload: bool = False) -> ToolPaths:
        ctx = self._ctx or default_context()
        ctx.upload(mesh, sharded=sharded_upload, geometry_only=not sharded_upload)
        p = self._params(mesh)
        p.shard_rank, p.shard_count = shard_rank, shard_count
        return ctx.plan(p)


    def plan_resident(self, shard_rank: int = 0, shard_count: int = 1) -> ToolPaths:
        """plan() on the mesh the context already holds (e.g. after Context.upload_ply): no upload."""
        ctx = self._ctx or default_context()
        p = self._params(None)
        p.shard_rank, p.shard_count = shard_rank, shard_count
        return ctx.plan(p)


class MultiToolPathPlanner:
    """MultiToolPathPlanner::plan (multi_tool_path_planner.cpp:10-20; alias `Multi`, plugins.cpp:134-149): the tool paths of
    every planner, one after the other.  A cross-hatch (noether_gui cross_hatch_plane_slicer_raster_planner.cpp) is two
    plane slicers, the second with a PCARotated direction: both find the mesh resident — one upload, one set of moments
    and extents."""

    def __init__(self, planners, context: Context | None = None):
        self.planners = list(planners)
        self._ctx = context

    def plan_each(self, mesh: MeshBlob) -> list[ToolPaths]:
        return [p.plan(mesh) for p in self.planners]

    def plan(self, mesh: MeshBlob) -> ToolPaths:
        nested = [path for tp in self.plan_each(mesh) for path in tp.nested()]
        return ToolPaths.from_nested(nested, self._ctx or (self.planners[0]._ctx if self.planners else None))


class PlaneSlicerLegacyRasterPlanner(PlaneSlicerRasterPlanner):
    """plane_slicer_legacy_raster_planner.cpp:10-33"""

    def __init__(self, dir_gen, origin_gen, point_spacing: float, min_segment_size: float, min_hole_size: float,
                 context: Context | None = None):
        super().__init__(dir_gen, origin_gen, context)
        self.point_spacing = float(point_spacing)
        self.min_segment_size = float(min_segment_size)
        self.min_hole_size = float(min_hole_size)

    def _params(self, mesh):
        p = super()._params(mesh)
        p.legacy = 1
        p.point_spacing = self.point_spacing
        p.min_segment_size = self.min_segment_size
        p.min_hole_size = self.min_hole_size
        return p


class NormalsFromMeshFacesMeshModifier:
    """normals_from_mesh_faces_modifier.cpp:15-46.  Returns [mesh] whose cloud is concatenateFields(mesh.cloud,
    normals): a 32-byte pcl::Normal record (normal_x0 normal_y4 normal_z8 curvature16) followed by the input's
    fields (the input's own normal fields, if any, are overridden by name)."""

    def __init__(self, context: Context | None = None):
        self._ctx = context

    def modify(self, mesh: MeshBlob) -> list[MeshBlob]:
        ctx = self._ctx or default_context()
        ctx.upload(mesh)
        normals = ctx.normals_from_mesh_faces(mesh.n_vertices)
        n = mesh.n_vertices
        step = 32 + mesh.point_step
        data = np.zeros(n * step, np.uint8)
        rec = data.reshape(n, step)
        rec[:, 0:12] = normals.view(np.uint8).reshape(n, 12)
        rec[:, 32:] = np.asarray(mesh.data).reshape(n, mesh.point_step)
        out = MeshBlob(data, n, step, 32 + mesh.off_xyz, 0, mesh.tri)
        return [out]


class NormalEstimationPCLMeshModifier:
    """normal_estimation_pcl.cpp:15-50 (alias NormalEstimationPCL, YAML keys radius, vx, vy, vz).  Returns [mesh] whose cloud is
    concatenateFields(mesh.cloud, normals): the 32-byte pcl::Normal record (normal_x0 normal_y4 normal_z8 curvature16)
    followed by the input's fields."""

    def __init__(self, radius: float, vx: float = 0.0, vy: float = 0.0, vz: float = 0.0, context: Context | None = None):
        self.radius, self.vx, self.vy, self.vz = float(radius), float(vx), float(vy), float(vz)
        self._ctx = context

    def modify(self, mesh: MeshBlob) -> list[MeshBlob]:
        ctx = self._ctx or default_context()
        ctx.upload(mesh)
        normals, curvature = ctx.normal_estimation_pcl(mesh.n_vertices, self.radius, (self.vx, self.vy, self.vz))
        ctx._mesh_key = None  # the resident normals are the estimated ones now, not the uploaded cloud's
        n = mesh.n_vertices
        step = 32 + mesh.point_step
        data = np.zeros(n * step, np.uint8)
        rec = data.reshape(n, step)
        rec[:, 0:12] = normals.view(np.uint8).reshape(n, 12)
        rec[:, 16:20] = curvature.view(np.uint8).reshape(n, 4)
        rec[:, 32:] = np.asarray(mesh.data).reshape(n, mesh.point_step)
        return [MeshBlob(data, n, step, 32 + mesh.off_xyz, 0, mesh.tri)]


def cluster_order(sizes) -> np.ndarray:
    """nb2_cluster_order (host only): PCL's reverse std::sort by size of clusters given in order of discovery"""
    sizes = np.ascontiguousarray(sizes, np.uint32)
    order = np.zeros(len(sizes), np.uint32)
    _capi.check(_capi.load().nb2_cluster_order(sizes.ctypes.data, len(sizes), order.ctypes.data))
    return order


def cluster_submeshes(labels, mesh: MeshBlob, min_cluster_size: int, max_cluster_size: int) -> list:
    """Host side of EuclideanClusteringMeshModifier::modify after the clustering itself: the size filter of
    pcl::extractEuclideanClusters, PCL's size sort (cluster_order), extractSubMeshFromInlierVertices
    (subset_extractor.cpp:22-56) and pcl::surface::SimplificationRemoveUnusedVertices.  Returns, in the reference's
    output order, MeshBlob objects (sub-mesh with the used points in order of first use; the input itself when every point
    is used) or None where the reference returns a default-constructed mesh (no face has all three vertices in the
    cluster)."""
    labels = np.asarray(labels, np.uint32)
    V = mesh.n_vertices
    seeds, sizes = np.unique(labels, return_counts=True)          # ascending seed = order of discovery
    min_pts = int(np.uint32(np.int64(min_cluster_size) & 0xffffffff))   # setMinClusterSize takes an unsigned
    max_pts = V if max_cluster_size <= 0 else int(max_cluster_size)     # euclidean_clustering_modifier.cpp:53
    keep = (sizes >= min_pts) & (sizes <= max_pts)
    seeds, sizes = seeds[keep], sizes[keep]
    tri = np.ascontiguousarray(mesh.tri, np.uint32).reshape(-1, 3)
    tri_label = labels[tri] if len(tri) else np.zeros((0, 3), np.uint32)
    face_seed = np.where((tri_label[:, 0] == tri_label[:, 1]) & (tri_label[:, 0] == tri_label[:, 2]), tri_label[:, 0].astype(np.int64), -1)
    rec = np.asarray(mesh.data, np.uint8).reshape(V, mesh.point_step)
    by_seed = np.argsort(face_seed, kind="stable")                # faces grouped by cluster, face order kept inside a group
    sorted_seed = face_seed[by_seed]
    out = []
    for k in cluster_order(sizes):
        lo, hi = np.searchsorted(sorted_seed, [int(seeds[k]), int(seeds[k]) + 1])
        kept = tri[by_seed[lo:hi]]
        if len(kept) == 0:
            out.append(None)
            continue
        flat = kept.ravel()
        uniq, first = np.unique(flat, return_index=True)
        used = uniq[np.argsort(first, kind="stable")]              # order of first use by the kept faces
        if len(used) == V:
            out.append(MeshBlob(np.asarray(mesh.data, np.uint8).copy(), V, mesh.point_step, mesh.off_xyz, mesh.off_normal, kept.copy()))
            continue
        new_index = np.zeros(V, np.uint32)
        new_index[used] = np.arange(len(used), dtype=np.uint32)
        out.append(MeshBlob(np.ascontiguousarray(rec[used]).reshape(-1), len(used), mesh.point_step, mesh.off_xyz, mesh.off_normal,
                            new_index[kept]))
    return out


class EuclideanClusteringMeshModifier:
    """euclidean_clustering_modifier.cpp:40-64 (plugin alias EuclideanClustering; YAML keys tolerance, min_cluster_size,
    max_cluster_size): the clustering runs on the GPU (connected components under FLANN's radius test), the sub-meshes
    are cut on the host from the labels.  An entry of the result is None where the reference returns an empty mesh."""

    def __init__(self, tolerance: float, min_cluster_size: int = 1, max_cluster_size: int = -1, context: Context | None = None):
        self._ctx = context
        self.tolerance, self.min_cluster_size, self.max_cluster_size = tolerance, min_cluster_size, max_cluster_size

    def modify(self, mesh: MeshBlob) -> list:
        ctx = self._ctx or default_context()
        ctx.upload(mesh)
        labels = ctx.euclidean_cluster_labels(self.tolerance, mesh.n_vertices)
        return cluster_submeshes(labels, mesh, self.min_cluster_size, self.max_cluster_size)


class _FaceSubdivisionBase:
    """Common part of the face subdivision mesh modifiers (face_subdivision_modifier.cpp): upload, subdivide on the
    device, read the new cloud (same fields as the input, utils.cpp copyPointToEnd) and triangle list back.
    `off_rgba` is the byte offset of the cloud's "rgba" field (-1: the cloud has none)."""

    def __init__(self, context: Context | None = None, off_rgba: int = -1):
        self._ctx = context
        self._off_rgba = off_rgba

    def _run(self, ctx: Context) -> _capi.MeshFileInfo:
        raise NotImplementedError

    def modify(self, mesh: MeshBlob) -> list[MeshBlob]:
        ctx = self._ctx or default_context()
        ctx.upload(mesh, force=True)
        info = self._run(ctx)
        data = ctx.download_cloud(info.n_points)
        tri = ctx.download_triangles(info.n_triangles)
        return [MeshBlob(data, int(info.n_points), mesh.point_step, mesh.off_xyz, mesh.off_normal, tri)]


class FaceMidpointSubdivisionMeshModifier(_FaceSubdivisionBase):
    """face_subdivision_modifier.cpp:202-248 (plugin alias FaceMidpointSubdivision, YAML key n_iterations)."""

    def __init__(self, n_iterations: int = 1, context: Context | None = None, off_rgba: int = -1):
        super().__init__(context, off_rgba)
        self.n_iterations = n_iterations

    def _run(self, ctx):
        return ctx.face_midpoint_subdivision(self.n_iterations, self._off_rgba)


class FaceSubdivisionByAreaMeshModifier(_FaceSubdivisionBase):
    """face_subdivision_modifier.cpp:250-339 (plugin alias FaceSubdivisionByArea, YAML key max_area)."""

    def __init__(self, max_area: float, context: Context | None = None, off_rgba: int = -1, max_points: int = 0):
        super().__init__(context, off_rgba)
        self.max_area = max_area
        self.max_points = max_points

    def _run(self, ctx):
        return ctx.face_subdivision_by_area(self.max_area, self._off_rgba, self.max_points)


# ---------------------------------------------------------------------------------------------------------------------
# tool path modifiers on the device (noether_tpp/src/tool_path_modifiers/; nb2_result_modify)
# ---------------------------------------------------------------------------------------------------------------------
class ToolPathModifier:
    """noether::ToolPathModifier (core/tool_path_modifier.h): modify(ToolPaths) -> ToolPaths.  `tool_paths` may be a
    ToolPaths of this package (a plan stays compact until the chain expands it on the device) or nested pose lists."""

    def __init__(self, context: Context | None = None):
        self._ctx = context

    def _records(self) -> list:
        raise NotImplementedError

    def modify(self, tool_paths) -> ToolPaths:
        ctx = self._ctx or (tool_paths._ctx if isinstance(tool_paths, ToolPaths) else default_context())
        if not isinstance(tool_paths, ToolPaths):
            tool_paths = ToolPaths.from_nested(tool_paths, ctx)
        recs = self._records()
        arr = (_capi.Modifier * max(len(recs), 1))(*recs)
        h = C.c_void_p()
        _capi.check(ctx._lib.nb2_result_modify(ctx._h, tool_paths._r, arr, len(recs), C.byref(h)))
        return ToolPaths(ctx, h)


def _record(kind: int, v=(), n_points: int = 0) -> _capi.Modifier:
    m = _capi.Modifier()
    m.kind = kind
    m.n_points = int(n_points)
    for i, x in enumerate(v):
        m.v[i] = float(x)
    return m


class SnakeOrganizationModifier(ToolPathModifier):
    """snake_organization_modifier.cpp:5-22"""

    def _records(self):
        return [_record(_capi.NB2_MOD_SNAKE_ORGANIZATION)]


class RasterOrganizationModifier(ToolPathModifier):
    """raster_organization_modifier.cpp:7-46"""

    def _records(self):
        return [_record(_capi.NB2_MOD_RASTER_ORGANIZATION)]


class JoinCloseSegmentsToolPathModifier(ToolPathModifier):
    """join_close_segments_modifier.cpp:6-39"""

    def __init__(self, distance: float, context: Context | None = None):
        super().__init__(context)
        self.distance = float(distance)

    def _records(self):
        return [_record(_capi.NB2_MOD_JOIN_CLOSE_SEGMENTS, [self.distance])]


class MinimumSegmentLengthToolPathModifier(ToolPathModifier):
    """minimum_segment_length_modifier.cpp:7-37"""

    def __init__(self, minimum_length: float, context: Context | None = None):
        super().__init__(context)
        self.minimum_length = float(minimum_length)

    def _records(self):
        return [_record(_capi.NB2_MOD_MINIMUM_SEGMENT_LENGTH, [self.minimum_length])]


class LinearApproachModifier(ToolPathModifier):
    """linear_approach_modifier.cpp:23-55"""
    _kind = _capi.NB2_MOD_LINEAR_APPROACH

    def __init__(self, offset, n_points: int, context: Context | None = None):
        super().__init__(context)
        self.offset, self.n_points = np.asarray(offset, np.float64), int(n_points)

    def _records(self):
        return [_record(self._kind, self.offset, self.n_points)]


class LinearDepartureModifier(LinearApproachModifier):
    """linear_departure_modifier.cpp:5-27"""
    _kind = _capi.NB2_MOD_LINEAR_DEPARTURE


class OffsetModifier(ToolPathModifier):
    """offset_modifier.cpp:6-22; offset = 4 x 4 isometry"""

    def __init__(self, offset, context: Context | None = None):
        super().__init__(context)
        self.offset = np.asarray(offset, np.float64).reshape(4, 4)

    def _records(self):
        return [_record(_capi.NB2_MOD_OFFSET, self.offset.T.reshape(-1))]


class FixedOrientationModifier(ToolPathModifier):
    """fixed_orientation_modifier.cpp:5-33 (the constructor normalises the direction, the YAML decoder does not)"""

    def __init__(self, reference_x_direction, context: Context | None = None, _as_is: bool = False):
        super().__init__(context)
        self.ref, self._as_is = np.asarray(reference_x_direction, np.float64), _as_is

    def _records(self):
        return [_record(_capi.NB2_MOD_FIXED_ORIENTATION, [*self.ref, 1.0 if self._as_is else 0.0])]


class UniformOrientationModifier(ToolPathModifier):
    """uniform_orientation_modifier.cpp:5-30"""

    def _records(self):
        return [_record(_capi.NB2_MOD_UNIFORM_ORIENTATION)]


class ConcatenateModifier(ToolPathModifier):
    """concatenate_modifier.cpp:5-15"""

    def _records(self):
        return [_record(_capi.NB2_MOD_CONCATENATE)]


class DirectionOfTravelOrientationModifier(ToolPathModifier):
    """direction_of_travel_orientation_modifier.cpp:6-48"""

    def _records(self):
        return [_record(_capi.NB2_MOD_DIRECTION_OF_TRAVEL_ORIENTATION)]


class CircularLeadInModifier(ToolPathModifier):
    """circular_lead_in_modifier.cpp:7-44"""
    _kind = _capi.NB2_MOD_CIRCULAR_LEAD_IN

    def __init__(self, arc_angle: float, arc_radius: float, n_points: int, context: Context | None = None):
        super().__init__(context)
        self.arc_angle, self.arc_radius, self.n_points = float(arc_angle), float(arc_radius), int(n_points)

    def _records(self):
        return [_record(self._kind, [self.arc_angle, self.arc_radius], self.n_points)]


class CircularLeadOutModifier(CircularLeadInModifier):
    """circular_lead_out_modifier.cpp:7-44"""
    _kind = _capi.NB2_MOD_CIRCULAR_LEAD_OUT


class MovingAverageOrientationSmoothingModifier(ToolPathModifier):
    """moving_average_orientation_smoothing_modifier.cpp:6-84 (in-place recurrence along each segment)"""

    def __init__(self, window_size: int, context: Context | None = None):
        super().__init__(context)
        self.window_size = int(window_size)

    def _records(self):
        return [_record(_capi.NB2_MOD_MOVING_AVERAGE_ORIENTATION_SMOOTHING, (), self.window_size)]


class ToolDragOrientationToolPathModifier(ToolPathModifier):
    """tool_drag_orientation_modifier.cpp:6-35"""
    _kind = _capi.NB2_MOD_TOOL_DRAG_ORIENTATION

    def __init__(self, angle_offset: float, tool_radius: float, context: Context | None = None):
        super().__init__(context)
        self.angle_offset, self.tool_radius = float(angle_offset), float(tool_radius)

    def _records(self):
        return [_record(self._kind, [self.angle_offset, self.tool_radius])]


class BiasedToolDragOrientationToolPathModifier(ToolDragOrientationToolPathModifier):
    """biased_tool_drag_orientation_modifier.cpp:6-35"""
    _kind = _capi.NB2_MOD_BIASED_TOOL_DRAG_ORIENTATION


class CompoundModifier(ToolPathModifier):
    """compound_modifier.cpp: the modifiers in order; here one device-resident chain, one upload and one download"""

    def __init__(self, modifiers, context: Context | None = None):
        super().__init__(context)
        self.modifiers = list(modifiers)

    def _records(self):
        return [r for m in self.modifiers for r in m._records()]


def _isometry_from_yaml(node: dict) -> np.ndarray:
    """Translation3d(x, y, z) * Quaterniond(qw, qx, qy, qz) (serialization.h:41-55; Eigen 3.4
    QuaternionBase::toRotationMatrix, the quaternion is taken as given)"""
    x, y, z, w = (float(_member(node, k)) for k in ("qx", "qy", "qz", "qw"))
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    T = np.eye(4)
    T[0, 0], T[0, 1], T[0, 2] = 1.0 - (tyy + tzz), txy - twz, txz + twy
    T[1, 0], T[1, 1], T[1, 2] = txy + twz, 1.0 - (txx + tzz), tyz - twx
    T[2, 0], T[2, 1], T[2, 2] = txz - twy, tyz + twx, 1.0 - (txx + tyy)
    T[:3, 3] = [float(_member(node, k)) for k in ("x", "y", "z")]
    return T


# ---------------------------------------------------------------------------------------------------------------------
# Factory (aliases + YAML keys of plugins.cpp)
# ---------------------------------------------------------------------------------------------------------------------
def _member(config: dict, key: str):
    if key not in config:
        # serialization.h:10-16
        raise RuntimeError(f"Required member '{key}' not found in YAML node")
    return config[key]


def _vec3(node) -> np.ndarray:
    if isinstance(node, dict):
        return np.array([_member(node, "x"), _member(node, "y"), _member(node, "z")], np.float64)
    return np.asarray(node, np.float64)


class Factory:
    """noether::Factory for the aliases on this path (plugin_interface.h:106-129)."""

    def __init__(self, context: Context | None = None):
        self._ctx = context

    def create_direction_generator(self, config: dict) -> DirectionGenerator:
        name = _member(config, "name")
        if name == "PrincipalAxis":
            return PrincipalAxisDirectionGenerator(_member(config, "rotation_offset"), self._ctx)
        if name == "FixedDirection":
            return FixedDirectionGenerator(_vec3(_member(config, "direction")))
        if name == "PCARotated":  # plugins.cpp:91-107
            return PCARotatedDirectionGenerator(self.create_direction_generator(_member(config, "direction_generator")),
                                                _member(config, "rotation_offset"), self._ctx)
        raise RuntimeError(f"Failed to find plugin '{name}' in section 'dg'")

    def create_origin_generator(self, config: dict) -> OriginGenerator:
        name = _member(config, "name")
        if name == "Centroid":
            return CentroidOriginGenerator(self._ctx)
        if name == "AABBCenter":  # plugins.cpp:110
            return AABBCenterOriginGenerator(self._ctx)
        if name == "FixedOrigin":
            return FixedOriginGenerator(_vec3(_member(config, "origin")))
        raise RuntimeError(f"Failed to find plugin '{name}' in section 'og'")

    def create_tool_path_planner(self, config: dict) -> PlaneSlicerRasterPlanner:
        name = _member(config, "name")
        if name == "Multi":  # plugins.cpp:134-149
            return MultiToolPathPlanner([self.create_tool_path_planner(e) for e in _member(config, "planners")], self._ctx)
        if name not in ("PlaneSlicer", "PlaneSlicerLegacy"):
            raise RuntimeError(f"Failed to find plugin '{name}' in section 'tpp'")
        dir_gen = self.create_direction_generator(_member(config, "direction_generator"))
        origin_gen = self.create_origin_generator(_member(config, "origin_generator"))
        if name == "PlaneSlicer":
            tpp = PlaneSlicerRasterPlanner(dir_gen, origin_gen, self._ctx)
        else:
            tpp = PlaneSlicerLegacyRasterPlanner(dir_gen, origin_gen, float(_member(config, "point_spacing")),
                                                 float(_member(config, "min_segment_size")),
                                                 float(_member(config, "min_hole_size")), self._ctx)
        tpp.set_line_spacing(float(_member(config, "line_spacing")))
        tpp.generate_rasters_bidirectionally(bool(_member(config, "bidirectional")))
        return tpp

    def create_tool_path_modifier(self, config: dict) -> ToolPathModifier:
        """aliases and YAML keys of plugins.cpp:182-219 for the modifiers that run on the device"""
        name = _member(config, "name")
        c = self._ctx
        if name == "SnakeOrganization":
            return SnakeOrganizationModifier(c)
        if name == "RasterOrganization":
            return RasterOrganizationModifier(c)
        if name == "JoinCloseSegments":
            return JoinCloseSegmentsToolPathModifier(float(_member(config, "distance")), c)
        if name == "MinimumSegmentLength":
            return MinimumSegmentLengthToolPathModifier(float(_member(config, "minimum_length")), c)
        if name in ("LinearApproach", "LinearDeparture"):
            cls = LinearApproachModifier if name == "LinearApproach" else LinearDepartureModifier
            return cls(_vec3(_member(config, "offset")), int(_member(config, "n_points")), c)
        if name == "Offset":
            return OffsetModifier(_isometry_from_yaml(_member(config, "offset")), c)
        if name == "FixedOrientation":
            return FixedOrientationModifier(_vec3(_member(config, "ref_x_dir")), c, _as_is=True)
        if name == "UniformOrientation":
            return UniformOrientationModifier(c)
        if name == "Concatenate":
            return ConcatenateModifier(c)
        if name == "DirectionOfTravelOrientation":
            return DirectionOfTravelOrientationModifier(c)
        if name in ("CircularLeadIn", "CircularLeadOut"):
            cls = CircularLeadInModifier if name == "CircularLeadIn" else CircularLeadOutModifier
            return cls(float(_member(config, "arc_angle")), float(_member(config, "arc_radius")), int(_member(config, "n_points")), c)
        if name == "MovingAverageOrientationSmoothing":  # window_size is read as a double and cast (:103)
            return MovingAverageOrientationSmoothingModifier(int(float(_member(config, "window_size"))), c)
        if name in ("ToolDragOrientation", "BiasedToolDragOrientation"):
            cls = ToolDragOrientationToolPathModifier if name == "ToolDragOrientation" else BiasedToolDragOrientationToolPathModifier
            return cls(float(_member(config, "angle_offset")), float(_member(config, "tool_radius")), c)
        if name == "CompoundToolPathModifier":
            return CompoundModifier([self.create_tool_path_modifier(e) for e in _member(config, "modifiers")], c)
        raise RuntimeError(f"Failed to find plugin '{name}' in section 'mod'")

    def create_mesh_modifier(self, config: dict) -> NormalsFromMeshFacesMeshModifier:
        name = _member(config, "name")
        if name == "NormalsFromMeshFaces":
            return NormalsFromMeshFacesMeshModifier(self._ctx)
        if name == "NormalEstimationPCL":      # normal_estimation_pcl.cpp:67-75
            return NormalEstimationPCLMeshModifier(float(_member(config, "radius")), float(_member(config, "vx")),
                                                   float(_member(config, "vy")), float(_member(config, "vz")), self._ctx)
        if name == "FaceMidpointSubdivision":  # n_iterations is read as a float and cast (face_subdivision_modifier.cpp:357)
            return FaceMidpointSubdivisionMeshModifier(int(float(_member(config, "n_iterations"))), self._ctx)
        if name == "EuclideanClustering":      # euclidean_clustering_modifier.cpp:79-85
            return EuclideanClusteringMeshModifier(float(_member(config, "tolerance")), int(_member(config, "min_cluster_size")),
                                                   int(_member(config, "max_cluster_size")), self._ctx)
        if name == "FaceSubdivisionByArea":    # :371
            return FaceSubdivisionByAreaMeshModifier(float(_member(config, "max_area")), self._ctx)
        raise RuntimeError(f"Failed to find plugin '{name}' in section 'mesh'")

