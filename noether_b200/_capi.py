"""ctypes binding of libnoether_b200.so — a 1:1 transcription of include/noether_b200.h.

Loading fails loudly when the library is missing or cannot be loaded: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NOETHER_B200_LIB") or os.path.join(HERE, "libnoether_b200.so")  # (override: kernel-variant experiments)

NB2_OK = 0
NB2_ERR_INVALID_ARGUMENT = 1
NB2_ERR_CUDA = 2
NB2_ERR_NO_NORMALS = 3
NB2_ERR_NON_MANIFOLD = 4
NB2_ERR_CLOSED_LOOP = 5
NB2_ERR_TOO_FEW_POINTS = 6
NB2_ERR_NOT_TRIANGLES = 7
NB2_ERR_NON_FINITE = 8
NB2_ERR_SEGMENT_TOO_SHORT = 9
NB2_ERR_NCCL = 10
NB2_ERR_NO_MESH = 11
NB2_ERR_INTERNAL = 12

NB2_DIRECTION_EXPLICIT = 0
NB2_DIRECTION_PRINCIPAL_AXIS = 1
NB2_ORIGIN_EXPLICIT = 0
NB2_ORIGIN_CENTROID = 1
NB2_ALL_PLANES = 2**64 - 1
NB2_MESH_GEOMETRY_ONLY = 1
NB2_UNIQUE_ID_BYTES = 128

# every symbol include/noether_b200.h declares (tests check the library exports all of them)
DECLARED_SYMBOLS = [
    "nb2_abi_version", "nb2_last_error", "nb2_context_create", "nb2_context_destroy", "nb2_context_device",
    "nb2_mesh_upload", "nb2_mesh_upload_stream", "nb2_mesh_set_normals", "nb2_mesh_download", "nb2_mesh_pca", "nb2_principal_axis_direction",
    "nb2_centroid_origin", "nb2_pca_rotated_direction", "nb2_normals_from_mesh_faces", "nb2_normal_estimation_pcl", "nb2_make_lattice", "nb2_eigen_solver_3f", "nb2_plan_params_default", "nb2_slice",
    "nb2_plan", "nb2_plan_mesh", "nb2_last_lattice", "nb2_result_get", "nb2_result_layout", "nb2_result_poses", "nb2_result_poses_segments", "nb2_expand_poses", "nb2_result_free", "nb2_last_timings",
    "nb2_kernel_launches", "nb2_last_plan_general_path_planes", "nb2_link_profile", "nb2_last_plan_span", "nb2_timer_start", "nb2_timer_stop", "nb2_comm_unique_id", "nb2_comm_init",
    "nb2_result_gather", "nb2_mesh_upload_sharded", "nb2_comm_destroy",
    "nb2_ply_inspect", "nb2_mesh_upload_ply", "nb2_mesh_upload_ply_file", "nb2_mesh_download_triangles",
    "nb2_result_from_poses", "nb2_result_from_pose_segments", "nb2_result_modify",
    "nb2_stl_inspect", "nb2_mesh_upload_stl", "nb2_mesh_upload_file",
    "nb2_face_midpoint_subdivision", "nb2_face_subdivision_by_area", "nb2_mesh_download_cloud", "nb2_mesh_record_layout",
    "nb2_euclidean_cluster_labels", "nb2_cluster_order", "nb2_aabb_center_origin",
]

NB2_MOD_SNAKE_ORGANIZATION = 1
NB2_MOD_LINEAR_APPROACH = 2
NB2_MOD_LINEAR_DEPARTURE = 3
NB2_MOD_OFFSET = 4
NB2_MOD_FIXED_ORIENTATION = 5
NB2_MOD_UNIFORM_ORIENTATION = 6
NB2_MOD_CONCATENATE = 7
NB2_MOD_TOOL_DRAG_ORIENTATION = 8
NB2_MOD_BIASED_TOOL_DRAG_ORIENTATION = 9
NB2_MOD_DIRECTION_OF_TRAVEL_ORIENTATION = 10
NB2_MOD_MOVING_AVERAGE_ORIENTATION_SMOOTHING = 11
NB2_MOD_CIRCULAR_LEAD_IN = 12
NB2_MOD_CIRCULAR_LEAD_OUT = 13
NB2_MOD_RASTER_ORGANIZATION = 14
NB2_MOD_JOIN_CLOSE_SEGMENTS = 15
NB2_MOD_MINIMUM_SEGMENT_LENGTH = 16


class Modifier(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_points", C.c_int32), ("v", C.c_double * 16)]


class MeshView(C.Structure):
    _fields_ = [("cloud_data", C.c_void_p), ("n_points", C.c_uint64), ("point_step", C.c_uint32),
                ("offset_xyz", C.c_uint32), ("offset_normal", C.c_int32), ("flags", C.c_uint32),
                ("triangles", C.c_void_p), ("n_triangles", C.c_uint64)]


class Pca(C.Structure):
    _fields_ = [("mean", C.c_float * 3), ("evecs", C.c_float * 9), ("evals", C.c_float * 3),
                ("scales", C.c_float * 3), ("cov", C.c_float * 9), ("aabb_min", C.c_float * 3),
                ("aabb_max", C.c_float * 3), ("pad_", C.c_float), ("mean64", C.c_double * 3),
                ("cov64", C.c_double * 9)]


class Lattice(C.Structure):
    _fields_ = [("cut_direction", C.c_double * 3), ("cut_normal", C.c_double * 3), ("cut_origin", C.c_double * 3),
                ("start_loc", C.c_double * 3), ("line_spacing", C.c_double), ("d_min", C.c_double),
                ("d_max", C.c_double), ("num_planes", C.c_uint64), ("no_cuts", C.c_int32), ("pad_", C.c_int32)]


class PlanParams(C.Structure):
    _fields_ = [("line_spacing", C.c_double), ("bidirectional", C.c_int32), ("direction_mode", C.c_int32),
                ("direction", C.c_double * 3), ("rotation_offset", C.c_double), ("origin_mode", C.c_int32),
                ("raster_post_ops", C.c_int32), ("origin", C.c_double * 3), ("legacy", C.c_int32),
                ("emit_edges", C.c_int32), ("point_spacing", C.c_double), ("min_segment_size", C.c_double),
                ("min_hole_size", C.c_double), ("shard_rank", C.c_int32), ("shard_count", C.c_int32),
                ("download", C.c_int32), ("reserved_", C.c_int32)]


class ResultView(C.Structure):
    _fields_ = [("n_paths", C.c_uint64), ("n_segments", C.c_uint64), ("n_waypoints", C.c_uint64),
                ("path_offsets", C.POINTER(C.c_uint32)), ("segment_offsets", C.POINTER(C.c_uint32)),
                ("path_plane", C.POINTER(C.c_int64)), ("positions", C.POINTER(C.c_float)),
                ("normals", C.POINTER(C.c_float)), ("edge_lo", C.POINTER(C.c_uint32)),
                ("edge_hi", C.POINTER(C.c_uint32)), ("poses", C.POINTER(C.c_double)),
                ("raster_post_ops_applied", C.c_int32), ("legacy", C.c_int32)]


class MeshFileInfo(C.Structure):
    _fields_ = [("n_points", C.c_uint64), ("n_triangles", C.c_uint64), ("has_normals", C.c_int32), ("reserved_", C.c_int32)]


class Nb2Error(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(msg)
        self.status = status


_lib = None


def load():
    """dlopen the in-tree library; raises if it is absent (build it with `python -m noether_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m noether_b200.build` "
                           "(noether_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.nb2_abi_version.restype = C.c_int
    L.nb2_last_error.restype = C.c_char_p
    L.nb2_context_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.nb2_context_destroy.argtypes = [vp]
    L.nb2_context_destroy.restype = None
    L.nb2_context_device.argtypes = [vp]
    L.nb2_mesh_upload.argtypes = [vp, C.POINTER(MeshView)]
    L.nb2_mesh_upload_sharded.argtypes = [vp, C.POINTER(MeshView)]
    L.nb2_mesh_set_normals.argtypes = [vp, C.c_void_p]
    L.nb2_mesh_download.argtypes = [vp, C.c_void_p, C.c_void_p]
    L.nb2_ply_inspect.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(MeshFileInfo)]
    L.nb2_mesh_upload_ply.argtypes = [vp, C.c_void_p, C.c_uint64, C.POINTER(MeshFileInfo)]
    L.nb2_mesh_upload_ply_file.argtypes = [vp, C.c_char_p, C.POINTER(MeshFileInfo)]
    L.nb2_mesh_download_triangles.argtypes = [vp, C.c_void_p, C.c_uint64]
    L.nb2_stl_inspect.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(MeshFileInfo)]
    L.nb2_mesh_upload_stl.argtypes = [vp, C.c_void_p, C.c_uint64, C.POINTER(MeshFileInfo)]
    L.nb2_mesh_upload_file.argtypes = [vp, C.c_char_p, C.POINTER(MeshFileInfo)]
    L.nb2_face_midpoint_subdivision.argtypes = [vp, C.c_uint32, C.c_int32, C.POINTER(MeshFileInfo)]
    L.nb2_face_subdivision_by_area.argtypes = [vp, C.c_float, C.c_int32, C.c_uint64, C.POINTER(MeshFileInfo)]
    L.nb2_mesh_download_cloud.argtypes = [vp, C.c_void_p, C.c_uint64]
    L.nb2_euclidean_cluster_labels.argtypes = [vp, C.c_double, C.c_void_p, C.c_uint64]
    L.nb2_cluster_order.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    L.nb2_aabb_center_origin.argtypes = [vp, C.POINTER(C.c_double)]
    L.nb2_mesh_record_layout.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]
    L.nb2_mesh_pca.argtypes = [vp, C.POINTER(Pca)]
    L.nb2_principal_axis_direction.argtypes = [vp, C.c_double, C.POINTER(C.c_double)]
    L.nb2_centroid_origin.argtypes = [vp, C.POINTER(C.c_double)]
    L.nb2_pca_rotated_direction.argtypes = [vp, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.nb2_normals_from_mesh_faces.argtypes = [vp, C.c_void_p]
    L.nb2_eigen_solver_3f.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.nb2_make_lattice.argtypes = [C.POINTER(Pca), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double,
                                   C.c_int, C.POINTER(Lattice)]
    L.nb2_plan_params_default.argtypes = [C.POINTER(PlanParams)]
    L.nb2_plan_params_default.restype = None
    L.nb2_slice.argtypes = [vp, C.POINTER(Lattice), C.c_uint64, C.c_uint64, C.POINTER(vp)]
    L.nb2_plan.argtypes = [vp, C.POINTER(PlanParams), C.POINTER(vp)]
    L.nb2_plan_mesh.argtypes = [vp, C.POINTER(MeshView), C.POINTER(PlanParams), C.POINTER(vp)]
    L.nb2_last_lattice.argtypes = [vp, C.POINTER(Lattice)]
    L.nb2_result_get.argtypes = [vp, C.POINTER(ResultView)]
    L.nb2_result_poses.argtypes = [vp, C.c_void_p]
    L.nb2_result_poses_segments.argtypes = [vp, C.c_void_p]
    L.nb2_result_free.argtypes = [vp]
    L.nb2_result_free.restype = None
    L.nb2_last_timings.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
    L.nb2_kernel_launches.argtypes = [vp]
    L.nb2_kernel_launches.restype = C.c_uint64
    L.nb2_last_plan_general_path_planes.argtypes = [vp]
    L.nb2_last_plan_general_path_planes.restype = C.c_uint64
    L.nb2_link_profile.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.nb2_timer_start.argtypes = [vp]
    L.nb2_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.nb2_comm_unique_id.argtypes = [C.c_void_p]
    L.nb2_comm_init.argtypes = [vp, C.c_void_p, C.c_int, C.c_int]
    L.nb2_result_gather.argtypes = [vp, vp, C.c_int, C.POINTER(vp)]
    L.nb2_comm_destroy.argtypes = [vp]
    L.nb2_result_from_poses.argtypes = [vp, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(vp)]
    L.nb2_result_from_pose_segments.argtypes = [vp, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(vp)]
    L.nb2_result_modify.argtypes = [vp, vp, C.POINTER(Modifier), C.c_int, C.POINTER(vp)]
    _lib = L
    return L


def check(status: int):
    if status != NB2_OK:
        raise Nb2Error(status, load().nb2_last_error().decode())
