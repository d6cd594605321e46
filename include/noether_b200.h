/*
 * noether_b200 — C ABI of the B200-native raster-slicing path.
 *
 * This is the drop-in boundary: a plain C interface (pointers + sizes, no C++/PCL/Eigen/torch types) that the
 * boost_plugin_loader plugin shim (plugin/b200_plugins.cpp) binds so that it can export the reference's aliases
 * `PlaneSlicer`, `PlaneSlicerLegacy`, `PrincipalAxis`, `Centroid` and `NormalsFromMeshFaces`
 * (reference: noether_tpp/src/plugins.cpp:62,89,111,151-179,252-278).  Each entry point names the reference code it
 * replaces (paths relative to the reference checkout, noether_tpp 0.16.0).
 *
 * All computation runs in hand-written sm_100a CUDA kernels inside libnoether_b200.so.  There is no CPU fallback: a
 * missing / failing device makes every entry point return NB2_ERR_CUDA.
 *
 * Threading: a context serialises its own calls internally (mutex); distinct contexts are independent.
 * Errors: every function returns an nb2_status; nb2_last_error() gives the message the shim re-throws as
 * std::runtime_error (reference error behaviour: noether_tpp/src/core/tool_path_planner_pipeline.cpp:64-101).
 */
#ifndef NOETHER_B200_H
#define NOETHER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB2_ABI_VERSION 3

typedef enum nb2_status
{
  NB2_OK = 0,
  NB2_ERR_INVALID_ARGUMENT = 1,
  NB2_ERR_CUDA = 2,            /* no usable device, launch or memory failure */
  NB2_ERR_NO_NORMALS = 3,      /* plane_slicer_raster_planner.cpp:520-528 */
  NB2_ERR_NON_MANIFOLD = 4,    /* a cut point with more than two incident cut lines */
  NB2_ERR_CLOSED_LOOP = 5,     /* cut lines forming a cycle (the reference throws "Topological sort failed"/not_a_dag) */
  NB2_ERR_TOO_FEW_POINTS = 6,  /* pcl::PCA needs >= 3 points; centroid needs >= 1 (centroid_origin_generator.cpp:13) */
  NB2_ERR_NOT_TRIANGLES = 7,
  NB2_ERR_NON_FINITE = 8,      /* non-finite vertex coordinate */
  NB2_ERR_SEGMENT_TOO_SHORT = 9, /* uniform_spacing_modifier.cpp:20-21 */
  NB2_ERR_NCCL = 10,
  NB2_ERR_NO_MESH = 11,
  NB2_ERR_INTERNAL = 12
} nb2_status;

typedef struct nb2_context nb2_context; /* one per GPU: stream, workspace arenas, the resident mesh */
typedef struct nb2_result nb2_result;   /* one planned set of tool paths (host SoA + device copies) */

int nb2_abi_version(void);
/* Message of the last failing call on this thread (never NULL). */
const char* nb2_last_error(void);

/* Create / destroy the per-GPU context (device = CUDA ordinal). */
int nb2_context_create(int device, nb2_context** out);
void nb2_context_destroy(nb2_context* ctx);
int nb2_context_device(const nb2_context* ctx);

/* ------------------------------------------------------------------------------------------------------------------
 * Mesh ingest.  Replaces pcl::fromPCLPointCloud2 + pcl::io::mesh2vtk and the by-name field accessors
 * (plane_slicer_raster_planner.cpp:536-537,589-590; utils.cpp:65-117).  The shim resolves field names to byte
 * offsets; `cloud_data` is pcl::PCLPointCloud2::data, `triangles` is mesh.polygons flattened to uint32[3 * n].
 * Host pointers may be pageable or pinned; pointers into this GPU's memory are adopted in place (no copy, no
 * synchronisation: the caller keeps them alive until the next upload, and input errors surface at the next call that
 * synchronises).  The mesh stays resident in HBM (SoA float4 positions, uint32 x 3 faces; normals are gathered from
 * the cloud itself) until the next upload.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct nb2_mesh_view
{
  const void* cloud_data;   /* n_points * point_step bytes */
  uint64_t n_points;        /* cloud.width * cloud.height */
  uint32_t point_step;
  uint32_t offset_xyz;      /* offset of field "x"; "y","z" must follow contiguously (utils.cpp:82-83) */
  int32_t offset_normal;    /* offset of "normal_x" (normal_y, normal_z contiguous, utils.cpp:104-105); -1 = none */
  uint32_t flags;           /* NB2_MESH_*; 0 = the records travel and stay resident as they are */
  const uint32_t* triangles; /* 3 * n_triangles vertex indices */
  uint64_t n_triangles;
} nb2_mesh_view;
/* flags: only positions and normals of this cloud will be used on the device (planning, normals, clustering — not the
 * face subdivisions, which copy whole records): a pageable host cloud is compacted to {x y z [nx ny nz]} records on its way
 * through the pinned staging ring, so a 48-byte PointNormal record costs 24 bytes of PCIe and HBM traffic (a pinned cloud
 * is copied as it is: the copy engine needs no host core for it).
 * nb2_mesh_record_layout reports the resident layout. */
#define NB2_MESH_GEOMETRY_ONLY 1u

/* Host memory that is not pinned, and clouds to be compacted, travel through a ring of pinned slots filled by the host
 * cores (one cudaMemcpyAsync per slot behind them); pinned arrays go out in one piece. */
int nb2_mesh_upload(nb2_context* ctx, const nb2_mesh_view* mesh);

/* The same upload with the mesh PRODUCED by the caller, piece by piece, straight into the pinned slots — for callers
 * whose mesh is not two flat arrays (pcl::PolygonMesh::polygons is one std::vector per face: plugin/b200_plugins.cpp
 * flattens, compacts and fingerprints in these callbacks, one pass over the host data in all).
 *   cloud(user, first, count, dst)      writes records [first, first + count) as count * point_step bytes
 *   triangles(user, first, count, dst)  writes triangles [first, first + count) as 3 * count indices
 * are called from several library threads at once on disjoint ranges; `first` is a multiple of NB2_UPLOAD_GRAIN; a
 * non-zero return aborts the upload (NB2_ERR_INVALID_ARGUMENT).  When everything is staged, accept(user) decides:
 * non-zero, the staged mesh becomes the resident one (*committed = 1); zero, it is dropped and the resident mesh, with
 * everything derived from it on the device (PCA, extents, computed normals), stays (*committed = 0) — the caller found,
 * from what its callbacks saw, that the content is what the device already holds.  The staged mesh lands in a second
 * pair of device buffers, so nothing resident is touched before accept().  accept may be NULL (always commit). */
#define NB2_UPLOAD_GRAIN 65536u
typedef struct nb2_mesh_stream
{
  uint64_t n_points;
  uint32_t point_step;     /* bytes per record as written by `cloud` */
  uint32_t offset_xyz;
  int32_t offset_normal;   /* -1 = none */
  uint32_t reserved_;
  uint64_t n_triangles;
  int (*cloud)(void* user, uint64_t first, uint64_t count, void* dst);
  int (*triangles)(void* user, uint64_t first, uint64_t count, uint32_t* dst);
  int (*accept)(void* user);
  void* user;
} nb2_mesh_stream;
int nb2_mesh_upload_stream(nb2_context* ctx, const nb2_mesh_stream* mesh, int* committed /* may be NULL */);
/* Replace the resident normals with packed float[3 * n_points] from the host (e.g. another modifier's output). */
int nb2_mesh_set_normals(nb2_context* ctx, const float* normals);
/* Copy the resident SoA back (packed float[3 * n]); either pointer may be NULL.  Test / debugging aid. */
int nb2_mesh_download(nb2_context* ctx, float* xyz, float* normals);

/* ------------------------------------------------------------------------------------------------------------------
 * Mesh ingest from a file image.  Replaces pcl::io::loadPolygonFile(path, mesh) for binary little-endian PLY files
 * (noether_gui/src/widgets/tpp_widget.cpp:426; noether_tpp/test/tool_path_planner_utest.cpp:24 — vtkPLYReader ->
 * pcl::io::vtk2mesh -> PCLPointCloud2 + one std::vector per face on one host core) followed by nb2_mesh_upload: the
 * header is parsed on the host, the records go to the GPU in one copy and are decoded there (vertex properties x y z
 * [nx ny nz | normal_x normal_y normal_z] as float32, any other scalar properties skipped; face list
 * vertex_indices / vertex_index with 1-, 2- or 4-byte counts and indices).  Every face must be a triangle
 * (NB2_ERR_NOT_TRIANGLES otherwise, like the rest of this path); ASCII and big-endian files are refused with a message.
 * The mesh is then resident exactly as after nb2_mesh_upload (nb2_plan, nb2_normals_from_mesh_faces, ... follow).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct nb2_mesh_file_info
{
  uint64_t n_points;
  uint64_t n_triangles;
  int32_t has_normals; /* the file carries vertex normals (pcl::io::vtk2mesh would add normal_x/y/z fields) */
  int32_t reserved_;
} nb2_mesh_file_info;

/* Header only, on the host (no device needed): sizes a caller can allocate for, and the refusals above. */
int nb2_ply_inspect(const void* file_image, uint64_t n_bytes, nb2_mesh_file_info* info);
int nb2_mesh_upload_ply(nb2_context* ctx, const void* file_image, uint64_t n_bytes, nb2_mesh_file_info* info /* may be NULL */);
/* The same from a path: the file is memory-mapped and copied to the GPU straight out of the page cache. */
int nb2_mesh_upload_ply_file(nb2_context* ctx, const char* path, nb2_mesh_file_info* info /* may be NULL */);
/* Binary STL (pcl::io::loadPolygonFile -> vtkSTLReader with Merging on -> pcl::io::vtk2mesh): 80-byte header, a 4-byte
 * facet count that is not trusted, 50-byte facets read until the file ends.  The facet corners are welded on the device
 * exactly as vtkMergePoints does it in the reader — coincident corners (float equality, -0 == +0) become one point,
 * numbered by first occurrence in the file; facets with two equal ids are dropped — and the welded mesh (PointXYZ cloud,
 * no normals: follow with nb2_normals_from_mesh_faces) is resident as after nb2_mesh_upload.  ASCII STL is refused.
 * nb2_stl_inspect reports the sizes before welding (3 points per facet): upper bounds a caller can allocate for. */
int nb2_stl_inspect(const void* file_image, uint64_t n_bytes, nb2_mesh_file_info* info);
int nb2_mesh_upload_stl(nb2_context* ctx, const void* file_image, uint64_t n_bytes, nb2_mesh_file_info* info /* may be NULL */);
/* pcl::io::loadPolygonFile(path, mesh): dispatch on the extension (.ply / .stl); the file is memory-mapped. */
int nb2_mesh_upload_file(nb2_context* ctx, const char* path, nb2_mesh_file_info* info /* may be NULL */);
/* Copy the resident triangle list back (uint32[3 * n_triangles]); `capacity_triangles` = triangles the buffer holds
 * (refused with NB2_ERR_INVALID_ARGUMENT when the resident mesh has more). */
int nb2_mesh_download_triangles(nb2_context* ctx, uint32_t* triangles, uint64_t capacity_triangles);

/* ------------------------------------------------------------------------------------------------------------------
 * PCA frame.  Replaces the pcl::PCA block of planImpl (plane_slicer_raster_planner.cpp:534-553), and feeds
 * PrincipalAxisDirectionGenerator::generate (principal_axis_direction_generator.cpp:14-26) and
 * CentroidOriginGenerator::generate (centroid_origin_generator.cpp:8-17).  One fused pass accumulates centroid,
 * covariance and AABB in fp64 (warp-shuffle + block reduction, fixed order => run-to-run deterministic); the
 * symmetric 3x3 eigen-solve restates Eigen 3.4's SelfAdjointEigenSolver<Matrix3f>; a second pass takes the float
 * extents of the cloud in the PCA frame (pcl::PCA::project + getMinMax3D).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct nb2_pca
{
  float mean[3];   /* pca.getMean() */
  float evecs[9];  /* pca.getEigenVectors(), column-major, column 0 = largest eigenvalue */
  float evals[3];  /* descending */
  float scales[3]; /* max - min of the projected cloud */
  float cov[9];    /* covariance handed to the eigen solver, column-major */
  float aabb_min[3];
  float aabb_max[3];
  float pad_;
  double mean64[3]; /* fp64 moments before the float cast */
  double cov64[9];
} nb2_pca;

int nb2_mesh_pca(nb2_context* ctx, nb2_pca* out);
int nb2_principal_axis_direction(nb2_context* ctx, double rotation_offset, double out_direction[3]);
int nb2_centroid_origin(nb2_context* ctx, double out_origin[3]);
/* AABBCenterOriginGenerator::generate (aabb_center_origin_generator.cpp:9-17): pcl::getMinMax3D of the vertices (the float
 * AABB of the ingest pass), then (max - min) / 2 — half the range of the box, as the reference computes it. */
int nb2_aabb_center_origin(nb2_context* ctx, double out_origin[3]);
/* PCARotatedDirectionGenerator::generate (pca_rotated_direction_generator.cpp:13-25): the direction of a nested
 * generator rotated by rotation_angle about the smallest principal axis of the resident mesh. */
int nb2_pca_rotated_direction(nb2_context* ctx, double rotation_angle, const double nominal_direction[3],
                              double out_direction[3]);

/* ------------------------------------------------------------------------------------------------------------------
 * NormalsFromMeshFacesMeshModifier::modify (normals_from_mesh_faces_modifier.cpp:15-46; getFaceNormal
 * utils.cpp:135-162).  Unit face normals, accumulated per vertex through a CSR vertex->face list in ascending face
 * order (no float atomics; bit-identical to the reference's sequential float sums), normalised.  The result
 * replaces the resident normals and, when `normals_out` is not NULL, is copied to the host as packed float[3 * n].
 * The shim builds the concatenated output cloud (pcl::concatenateFields) from it.
 * ------------------------------------------------------------------------------------------------------------------ */
int nb2_normals_from_mesh_faces(nb2_context* ctx, float* normals_out);

/* ------------------------------------------------------------------------------------------------------------------
 * Face subdivision mesh modifiers applied to the resident mesh:
 *   nb2_face_midpoint_subdivision  = FaceMidpointSubdivisionMeshModifier::modify (face_subdivision_modifier.cpp:202-248;
 *                                    YAML key n_iterations, plugins.cpp:58): every triangle becomes 4^n triangles, one new
 *                                    vertex per edge (shared between the triangles that meet there);
 *   nb2_face_subdivision_by_area   = FaceSubdivisionByAreaMeshModifier::modify (:250-339; YAML key max_area,
 *                                    plugins.cpp:59): a triangle whose area exceeds max_area gets a centroid vertex and
 *                                    three children, recursively.
 * A new vertex is a byte copy of the record of its first source whose x y z, normal_x/y/z (when the cloud has them) and
 * rgba (offset_rgba >= 0: offset of the field "rgba", utils.cpp:119-133; -1 = none) are the mean of the sources
 * (copyPointToEnd / updatePointData, :58-114).  Vertex numbers, record bytes and face order are exactly those of the
 * reference's sequential depth-first walk.  The subdivided mesh replaces the resident one (same record layout; moments
 * and PCA axes recomputed), so nb2_plan / nb2_normals_from_mesh_faces can follow without a host round trip; the shim
 * reads it back with nb2_mesh_download_cloud + nb2_mesh_download_triangles (sizes in `info`).
 * The modifiers work on the cloud as uploaded (the reference's modifiers see nothing else either): normals computed on
 * the device by nb2_normals_from_mesh_faces live beside the cloud and are not subdivided — upload the concatenated cloud
 * the shim builds from them, or run nb2_normals_from_mesh_faces again on the subdivided mesh.  The subdivision and the
 * hand-over of its result are two critical sections of the context: do not upload to the same context from another
 * thread in between (the plugin shim serialises its calls per device anyway).
 * Refusals: n_iterations = 0 (never terminates in the reference) or beyond the 2^31 triangle limit of this path;
 * a non-finite triangle area (NB2_ERR_NON_FINITE, "Triangle area is not finite", :332-333); a by-area subdivision that
 * would exceed max_points points (0 = the 2^31 limit; the reference has no limit and would exhaust memory).
 * ------------------------------------------------------------------------------------------------------------------ */
int nb2_face_midpoint_subdivision(nb2_context* ctx, uint32_t n_iterations, int32_t offset_rgba, nb2_mesh_file_info* info /* may be NULL */);
int nb2_face_subdivision_by_area(nb2_context* ctx, float max_area, int32_t offset_rgba, uint64_t max_points,
                                 nb2_mesh_file_info* info /* may be NULL */);
/* Copy the resident cloud back: n_points records of point_step bytes (pcl::PCLPointCloud2::data); `capacity_bytes` =
 * size of the buffer (refused with NB2_ERR_INVALID_ARGUMENT when the resident cloud is larger: the record layout on the
 * device is the uploaded one, nb2_mesh_record_layout, not necessarily the caller's). */
int nb2_mesh_download_cloud(nb2_context* ctx, void* cloud_data, uint64_t capacity_bytes);
/* Record layout of the resident cloud; any pointer may be NULL. */
int nb2_mesh_record_layout(const nb2_context* ctx, uint32_t* point_step, uint32_t* offset_xyz, int32_t* offset_normal);

/* ------------------------------------------------------------------------------------------------------------------
 * EuclideanClusteringMeshModifier::modify (euclidean_clustering_modifier.cpp:40-64), device part: the
 * connected components pcl::EuclideanClusterExtraction grows with one k-d tree radius search per point — two points are
 * linked when FLANN's float squared distance ((dx^2) + dy^2) + dz^2 is strictly below float(tolerance^2).
 * labels[v] (uint32[n_points], host memory, `capacity_labels` entries) = the smallest point index of v's component = the seed PCL starts that
 * cluster from, so ascending labels are PCL's clusters in order of discovery.  tolerance <= 0 finds no neighbours
 * (labels[v] = v).  The size filter (min_cluster_size / max_cluster_size), PCL's reverse std::sort by size and
 * extractSubMeshFromInlierVertices (subset_extractor.cpp:22-56) are index bookkeeping on the labels and are done by the
 * caller (plugin shim), where the meshes are materialised.
 * Refusals: a tolerance so large against the point spacing that the 27-cell neighbourhoods hold more than 4e11 candidate
 * pairs (the reference's radius searches would each return a large
 * part of the cloud as well).
 * ------------------------------------------------------------------------------------------------------------------ */
int nb2_euclidean_cluster_labels(nb2_context* ctx, double tolerance, uint32_t* labels, uint64_t capacity_labels);

/* NormalEstimationPCLMeshModifier::modify (noether_tpp/src/mesh_modifiers/normal_estimation_pcl.cpp:15-50):
 * pcl::NormalEstimation<PointXYZ, Normal> with setRadiusSearch(radius) and setViewPoint(vx, vy, vz) on the resident cloud.
 * One k-d tree radius search per point in the reference; here a uniform grid of cells of the radius and one thread per
 * point: FLANN's float distance test decides the neighbours (the same sets), their shifted second moments are summed in
 * double, pcl::eigen33's closed-form solver gives the eigenvector of the smallest eigenvalue, flipped towards the viewpoint.
 * normals: float[3 * n_points], NaN where the search returns fewer than 3 points; curvature: float[n_points] (may be
 * NULL) = |lambda_0 / trace|.  The estimated normals become the resident normals (nb2_plan may follow).  Agreement with
 * the reference is a tolerance on the angle (sums in double instead of float, CUDA's atan2f / cosf / sinf), not bit
 * equality.  A radius that is large against the point spacing is refused like nb2_euclidean_cluster_labels' tolerance. */
int nb2_normal_estimation_pcl(nb2_context* ctx, double radius, double vx, double vy, double vz, float* normals, float* curvature);
/* Host only (no context): the order pcl::EuclideanClusterExtraction::extract leaves its clusters in —
 * std::sort(clusters.rbegin(), clusters.rend(), comparePointClusters), largest first, equal sizes as the library's
 * introsort leaves them — for clusters given in order of discovery (ascending label) with these sizes, after the size
 * filter.  order[k] = index into cluster_sizes of the k-th mesh the reference returns. */
int nb2_cluster_order(const uint32_t* cluster_sizes, uint64_t n_clusters, uint32_t* order);

/* ------------------------------------------------------------------------------------------------------------------
 * Plane lattice (plane_slicer_raster_planner.cpp:556-586 and getDistancesToMinMaxCuts :460-501).  Pure host
 * arithmetic in double, restated with the reference's quirks (OBB centred on the mean, `else if` min/max update,
 * floored start offset, ceil'ed plane count).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct nb2_lattice
{
  double cut_direction[3];
  double cut_normal[3];
  double cut_origin[3];
  double start_loc[3];
  double line_spacing;
  double d_min, d_max;
  uint64_t num_planes;
  int32_t no_cuts; /* !bidirectional && d_max < 0 -> planImpl returns {} (:569-570) */
  int32_t pad_;
} nb2_lattice;

/* Host only (no context): the symmetric 3x3 float eigen-solver the PCA frame uses, on the host and in the ingest kernel's
 * epilogue alike — Eigen 3.4 SelfAdjointEigenSolver<Matrix3f>::compute as pcl::PCA::initCompute calls it
 * (plane_slicer_raster_planner.cpp:540-541): lower triangle scaled into [-1, 1], closed-form 3x3 tridiagonalisation,
 * implicit QR steps with Wilkinson shift and Eigen's deflation test, ascending selection sort.  cov and evecs are
 * column-major; evals ascending, evecs[3 * i .. 3 * i + 2] = eigenvector of evals[i], signs as that sequence of rotations
 * leaves them (pcl::PCA then reverses the order).  Exposed so that the restatement can be tested in isolation. */
int nb2_eigen_solver_3f(const float cov[9], float evals_ascending[3], float evecs_ascending[9]);

int nb2_make_lattice(const nb2_pca* pca,
                     const double cut_direction[3],
                     const double cut_origin[3],
                     double line_spacing,
                     int bidirectional,
                     nb2_lattice* out);

/* ------------------------------------------------------------------------------------------------------------------
 * Planning.
 * nb2_slice  = the plane loop of PlaneSlicerRasterPlanner::planImpl (:596-628: vtkCutter + vtkStripper +
 *              processPolyLines + convertToPoses) for planes [plane_begin, plane_end) of a given lattice.
 * nb2_plan   = RasterPlanner::plan (raster_planner.cpp:16-35) for PlaneSlicerRasterPlanner /
 *              PlaneSlicerLegacyRasterPlanner: PCA, generators, lattice, slice, optional legacy modifiers, then
 *              RasterOrganizationModifier + DirectionOfTravelOrientationModifier.
 * ------------------------------------------------------------------------------------------------------------------ */
#define NB2_DIRECTION_EXPLICIT 0       /* value produced by any noether::DirectionGenerator */
#define NB2_DIRECTION_PRINCIPAL_AXIS 1 /* PrincipalAxisDirectionGenerator(rotation_offset), computed here */
#define NB2_ORIGIN_EXPLICIT 0          /* value produced by any noether::OriginGenerator */
#define NB2_ORIGIN_CENTROID 1          /* CentroidOriginGenerator, computed here */
#define NB2_ALL_PLANES UINT64_MAX

typedef struct nb2_plan_params
{
  double line_spacing;    /* RasterPlanner::setLineSpacing */
  int32_t bidirectional;  /* PlaneSlicerRasterPlanner::generateRastersBidirectionally */
  int32_t direction_mode;
  double direction[3];
  double rotation_offset;
  int32_t origin_mode;
  int32_t raster_post_ops; /* 1 = apply RasterOrganization + DirectionOfTravel (what RasterPlanner::plan returns) */
  double origin[3];
  /* PlaneSlicerLegacyRasterPlanner (plane_slicer_legacy_raster_planner.cpp:22-33); legacy = 0 ignores the rest */
  int32_t legacy;
  int32_t emit_edges;      /* 1 = also return the generating mesh edge of every waypoint (connectivity provenance) */
  double point_spacing;
  double min_segment_size;
  double min_hole_size;
  /* plane-slab sharding: this context plans lattice planes [shard_rank * P / shard_count, (shard_rank+1) * P / shard_count) */
  int32_t shard_rank;
  int32_t shard_count; /* 0 or 1 = unsharded */
  /* 1 (default) = copy the tool paths to the host.  0 = leave them in HBM: the result carries the counts only and the
   * host-side modifiers (legacy, raster post-ops) are skipped — used to time the device path on resident data.
   * 2 = copy them BEHIND the call: nb2_plan returns once the segment structure is on the host; positions and normals
   * follow in chunks while the caller sizes its containers (nb2_result_layout) and nb2_result_poses_segments expands
   * the chunks that have landed (PCIe and the host cores overlap).  nb2_result_get and every other consumer of the result
   * wait for the rest first.  Ignored (= 1) for legacy plans, raster_post_ops, emit_edges and sharded plans. */
  int32_t download;
  int32_t reserved_;
} nb2_plan_params;

void nb2_plan_params_default(nb2_plan_params* p);

int nb2_slice(nb2_context* ctx, const nb2_lattice* lattice, uint64_t plane_begin, uint64_t plane_end, nb2_result** out);
int nb2_plan(nb2_context* ctx, const nb2_plan_params* params, nb2_result** out);
/* nb2_mesh_upload + nb2_plan in one call: exactly PlaneSlicerRasterPlanner::planImpl(mesh)
 * (plane_slicer_raster_planner.cpp:518-631), what the plugin's planImpl binds. */
int nb2_plan_mesh(nb2_context* ctx, const nb2_mesh_view* mesh, const nb2_plan_params* params, nb2_result** out);
/* Lattice used by the last nb2_plan on this context. */
int nb2_last_lattice(const nb2_context* ctx, nb2_lattice* out);

/* ------------------------------------------------------------------------------------------------------------------
 * Results.  noether::ToolPaths (core/types.h:31-50) as SoA on the host: tool paths -> segments -> waypoints.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct nb2_result_view
{
  uint64_t n_paths;
  uint64_t n_segments;
  uint64_t n_waypoints;
  const uint32_t* path_offsets;    /* n_paths + 1, indexes segments */
  const uint32_t* segment_offsets; /* n_segments + 1, indexes waypoints */
  const int64_t* path_plane;       /* n_paths, lattice index of each tool path */
  const float* positions;          /* 3 * n_waypoints, float32 cut points (what vtkCutter stores) */
  const float* normals;            /* 3 * n_waypoints, float32 interpolated vertex normals (not normalised) */
  const uint32_t* edge_lo;         /* n_waypoints: generating mesh edge, lower-scalar vertex */
  const uint32_t* edge_hi;         /* n_waypoints: generating mesh edge, higher-scalar vertex */
  const double* poses;             /* NULL unless the result carries explicit poses (legacy resampling): 16 * n */
  int32_t raster_post_ops_applied;
  int32_t legacy;
} nb2_result_view;

int nb2_result_get(const nb2_result* r, nb2_result_view* out);
/* The same view without waiting for a download = 2 result's positions / normals: counts, path_offsets, segment_offsets and
 * path_plane are valid, the waypoint arrays may still be arriving (read them after nb2_result_get, or let
 * nb2_result_poses_segments consume them). */
int nb2_result_layout(const nb2_result* r, nb2_result_view* out);
/* Expand to Eigen::Isometry3d storage: 16 doubles per waypoint, column-major, last row 0 0 0 1
 * (createTransform :380-394, then DirectionOfTravelOrientationModifier when post-ops were applied). */
int nb2_result_poses(const nb2_result* r, double* out16);
/* The same expansion, written segment by segment into storage the caller owns: segment_out[s] receives the
 * 16 * (waypoints of segment s) doubles of segment s (s < n_segments) — e.g. the data() of each
 * noether::ToolPathSegment (std::vector<Eigen::Isometry3d>), so the plugin fills noether::ToolPaths without an
 * intermediate copy.  The segments are filled in parallel on the host cores. */
int nb2_result_poses_segments(const nb2_result* r, double* const* segment_out);
/* The same expansion on SoA arrays the caller holds (no context, no device): positions / normals float[3 * W], the
 * segments' offsets uint32[n_segments + 1] into the waypoints; raster_post_ops_applied as in nb2_result_view; out16
 * receives 16 * W doubles.  On x86-64 with AVX2 four waypoints are expanded per step, bit-identical to the scalar code
 * (NB2_NO_AVX2 in the environment selects the scalar code: tests compare the two). */
int nb2_expand_poses(const float* positions, const float* normals, const uint32_t* segment_offsets, uint64_t n_segments,
                     int raster_post_ops_applied, double* out16);
void nb2_result_free(nb2_result* r);

/* ------------------------------------------------------------------------------------------------------------------
 * Tool-path modifiers on the device (the noether::ToolPathModifier implementations most often chained after a raster
 * plan, noether_tpp/src/tool_path_modifiers/).  The poses of a result are expanded (or uploaded) once, stay in HBM as
 * Eigen::Isometry3d storage (16 doubles per waypoint) while the chain runs — one streaming kernel per modifier, each
 * reading one pose buffer and writing the other — and come back to the host once, with the float positions / z axes.
 * Arithmetic: Eigen 3.4's Transform operations in the reference's operand order, no FMA contraction.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef enum nb2_modifier_kind
{
  NB2_MOD_SNAKE_ORGANIZATION = 1,          /* snake_organization_modifier.cpp:5-22 */
  NB2_MOD_LINEAR_APPROACH = 2,             /* linear_approach_modifier.cpp:33-55: v[0..2] = offset, n_points */
  NB2_MOD_LINEAR_DEPARTURE = 3,            /* linear_departure_modifier.cpp:5-27: v[0..2] = offset, n_points */
  NB2_MOD_OFFSET = 4,                      /* offset_modifier.cpp:8-22: v[0..15] = offset, column-major 4 x 4 */
  NB2_MOD_FIXED_ORIENTATION = 5,           /* fixed_orientation_modifier.cpp:5-33: v[0..2] = reference x direction, normalised
                                              here as the constructor does (:7); v[3] != 0: taken as is, as the YAML decoder
                                              stores it (:52) */
  NB2_MOD_UNIFORM_ORIENTATION = 6,         /* uniform_orientation_modifier.cpp:5-30 */
  NB2_MOD_CONCATENATE = 7,                 /* concatenate_modifier.cpp:5-15 */
  NB2_MOD_TOOL_DRAG_ORIENTATION = 8,       /* tool_drag_orientation_modifier.cpp:11-35: v[0] = angle_offset, v[1] = tool_radius */
  NB2_MOD_BIASED_TOOL_DRAG_ORIENTATION = 9,/* biased_tool_drag_orientation_modifier.cpp:12-35: v[0] = angle_offset, v[1] = tool_radius */
  NB2_MOD_DIRECTION_OF_TRAVEL_ORIENTATION = 10, /* direction_of_travel_orientation_modifier.cpp:6-48 */
  NB2_MOD_RASTER_ORGANIZATION = 14,        /* raster_organization_modifier.cpp:7-46 (decided on the host from the first /
                                              last translation of every segment, read back from the device) */
  NB2_MOD_JOIN_CLOSE_SEGMENTS = 15,        /* join_close_segments_modifier.cpp:11-39: v[0] = distance (offsets only: the merged
                                              segments are already adjacent in memory) */
  NB2_MOD_MINIMUM_SEGMENT_LENGTH = 16,     /* minimum_segment_length_modifier.cpp:12-37, computeLength utils.cpp:213-234:
                                              v[0] = minimum_length */
  NB2_MOD_CIRCULAR_LEAD_IN = 12,           /* circular_lead_in_modifier.cpp:12-44: v[0] = arc_angle, v[1] = arc_radius, n_points */
  NB2_MOD_CIRCULAR_LEAD_OUT = 13,          /* circular_lead_out_modifier.cpp:12-44: v[0] = arc_angle, v[1] = arc_radius, n_points */
  NB2_MOD_MOVING_AVERAGE_ORIENTATION_SMOOTHING = 11 /* moving_average_orientation_smoothing_modifier.cpp:6-84:
                                              n_points = window_size.  The reference smooths in place, so a window holds
                                              the already smoothed poses before its centre: a recurrence along each
                                              segment, run by one thread per segment */
} nb2_modifier_kind;

typedef struct nb2_modifier
{
  int32_t kind;     /* nb2_modifier_kind */
  int32_t n_points; /* LinearApproach / LinearDeparture */
  double v[16];
} nb2_modifier;

/* noether::ToolPaths handed in by a caller (what ToolPathModifier::modify receives): path_offsets[n_paths + 1] index
 * segments, segment_offsets[n_segments + 1] index waypoints, poses16 = 16 doubles per waypoint (column-major
 * Eigen::Isometry3d).  The result owns a copy; its positions / normals are the float translation / z axis. */
int nb2_result_from_poses(nb2_context* ctx, uint64_t n_paths, const uint32_t* path_offsets, uint64_t n_segments,
                          const uint32_t* segment_offsets, const double* poses16, nb2_result** out);
/* The same from segment-wise storage: segment_poses[s] points at the 16 * (waypoints of segment s) doubles of segment s —
 * the data() of each noether::ToolPathSegment — so a plugin hands noether::ToolPaths over without flattening them first. */
int nb2_result_from_pose_segments(nb2_context* ctx, uint64_t n_paths, const uint32_t* path_offsets, uint64_t n_segments,
                                  const uint32_t* segment_offsets, const double* const* segment_poses, nb2_result** out);
/* CompoundModifier::modify (compound_modifier.cpp) over the kinds above: applies chain[0], chain[1], ... to the tool
 * paths of `in` (left untouched) and returns a result with explicit poses.  A result without explicit poses (a plan) is
 * expanded on the device exactly as nb2_result_poses would. */
int nb2_result_modify(nb2_context* ctx, const nb2_result* in, const nb2_modifier* chain, int n_chain, nb2_result** out);

/* Device timings (CUDA events on the context's stream, milliseconds) of the last nb2_mesh_upload followed by those of
 * the last nb2_plan / nb2_slice / nb2_normals_from_mesh_faces on the context; `names` points at static strings.
 * Returns the number of stages written (<= capacity). */
int nb2_last_timings(const nb2_context* ctx, const char** names, float* ms, int capacity);
/* Number of kernel launches issued by this context since creation. */
uint64_t nb2_kernel_launches(const nb2_context* ctx);
/* Diagnostics: lattice planes of the last nb2_plan / nb2_slice whose cut lines the linker's general path joined because no
 * fast path took them (a cut point with more than two line ends, a closed contour, inconsistently wound triangles, a plane
 * beyond the fast paths' work area).  Same result either way; the general path is slower. */
uint64_t nb2_last_plan_general_path_planes(const nb2_context* ctx);
/* Diagnostics: with NB2_LINK_PROFILE and NB2_LINK_GEN=2 set in the environment the linking kernel sums, per phase of its
 * second-generation fast path, the clock cycles thread 0 of every CTA spent (9 phases: init, hashing, claim rounds, pairing
 * + zero-length lines, pairing check + splitters, walks, pointer jumping, chain heads + order, output); cycles[0..15] of the
 * last plan, zeros otherwise. */
int nb2_link_profile(nb2_context* ctx, uint64_t cycles[16]);
/* Device time of the last nb2_plan / nb2_plan_mesh on this context: from a CUDA event recorded on the context's stream at
 * the entry of the call (for nb2_plan_mesh: before the upload) to one recorded behind the plan's last slicing kernel —
 * the copies of the result to the host and the host's own work after the call are not in it. */
int nb2_last_plan_span(nb2_context* ctx, float* ms);
/* Bracket any sequence of calls on this context with CUDA events on the context's stream.  nb2_timer_stop waits for the
 * stop event and returns the elapsed device time in milliseconds. */
int nb2_timer_start(nb2_context* ctx);
int nb2_timer_stop(nb2_context* ctx, float* ms);

/* ------------------------------------------------------------------------------------------------------------------
 * Multi-GPU (one process per GPU).  Plane slabs are independent (the reference's loop :596-628 carries no state), so
 * the only exchange is the gather of per-slab polylines to rank 0, over NCCL (NVLink 5 / NVSwitch).
 * ------------------------------------------------------------------------------------------------------------------ */
#define NB2_UNIQUE_ID_BYTES 128
int nb2_comm_unique_id(uint8_t out[NB2_UNIQUE_ID_BYTES]);
int nb2_comm_init(nb2_context* ctx, const uint8_t unique_id[NB2_UNIQUE_ID_BYTES], int rank, int world_size);
/* Collective: every rank passes the result of its sharded nb2_plan (shard_count > 1; such a result keeps its tool paths
 * in the GPU's output buffers, valid until the next plan on that context).  The slabs travel GPU to GPU (grouped
 * ncclSend / ncclRecv of the SoA arrays); on `root` *gathered receives the concatenation in rank (= plane) order, on
 * other ranks NULL.  Legacy modifiers and raster post-ops, when requested in the params, are applied to the gathered
 * result. */
int nb2_result_gather(nb2_context* ctx, const nb2_result* local, int root, nb2_result** gathered);
/* Collective nb2_mesh_upload for a sharded job whose ranks all hold the same host mesh: every rank copies 1 / world of the
 * cloud and of the triangle list over its own PCIe link, and an in-place ncclAllGather over NVLink gives each GPU the
 * whole mesh.  Without a communicator (or for device-resident input) it is nb2_mesh_upload. */
int nb2_mesh_upload_sharded(nb2_context* ctx, const nb2_mesh_view* mesh);
int nb2_comm_destroy(nb2_context* ctx);

#ifdef __cplusplus
}
#endif
#endif /* NOETHER_B200_H */
