/* TEST INFRASTRUCTURE ONLY.
 *
 * C API of the CPU oracle: a restatement of the raster-slicing hot path of ros-industrial/noether (noether_tpp 0.16.0)
 * and of the third-party calls it makes (VTK 9.1 vtkCutter/vtkTriangle::Contour/vtkMergePoints/vtkStripper,
 * PCL pcl::PCA/computeCentroid/getMinMax3D, Eigen 3.4 fixed-size kernels, Boost.Graph topological_sort).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (libnoether_b200.so) never links or calls it.
 *
 * PARITY STATUS: the reference cannot be compiled in this image (VTK, PCL, Eigen, Boost, yaml-cpp and
 * boost_plugin_loader are absent, no network), so this oracle is pinned against the reference's own test expectations
 * (tool_path_planner_utest.cpp:104-119,151-176; tool_path_modifier_utest.cpp:361-399) — which pin counts and topology
 * only.  Coordinates, normals, PCA axes and centroid are "parity unpinned" at the third-party boundary.
 */
#ifndef NOETHER_ORACLE_H
#define NOETHER_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* moment accumulation mode */
#define ORC_MOMENTS_F32_SEQUENTIAL 0 /* what PCL does: float accumulators, vertex order */
#define ORC_MOMENTS_F64 1            /* double accumulators, result cast to float (GPU-comparable) */

/* slicing mode */
#define ORC_SLICE_FAITHFUL 0 /* vtkStripper + unify/topological sort/stack join, as the reference */
#define ORC_SLICE_INTENT 1   /* one segment per maximal open chain (reference's documented contract) */

/* which cells each plane visits */
#define ORC_SCAN_ALL_CELLS 0 /* every vertex + every cell per plane, like vtkCutter (Theta(P(V+F))) */
#define ORC_SCAN_BINNED 1    /* only cells whose conservative plane interval contains the plane; same results */

/* status codes */
#define ORC_OK 0
#define ORC_ERR_INVALID 1
#define ORC_ERR_TOO_MANY_CONNECTING 2 /* plane_slicer_raster_planner.cpp:37-38 */
#define ORC_ERR_NOT_A_DAG 3           /* boost::not_a_dag out of topological_sort (:148) */
#define ORC_ERR_TOPO_SORT 4           /* :158-159 */
#define ORC_ERR_JOIN 5                /* :264,282,287 */
#define ORC_ERR_NON_MANIFOLD 6        /* intent mode: cut point with more than two incident lines */
#define ORC_ERR_CLOSED_LOOP 7         /* intent mode: cut lines forming a cycle */
#define ORC_ERR_TOO_FEW_POINTS 8
#define ORC_ERR_SEGMENT_TOO_SHORT 9 /* uniform_spacing_modifier.cpp:20-21 */
#define ORC_ERR_NON_FINITE 10       /* face_subdivision_modifier.cpp:332-333 */

const char* orc_last_error(void);

/* ---- PCA / generators (plane_slicer_raster_planner.cpp:534-553, principal_axis_direction_generator.cpp:14-26,
 *      centroid_origin_generator.cpp:8-17) ---- */
typedef struct orc_pca_result
{
  float mean[3];
  float evecs[9]; /* column-major 3x3; column 0 = largest eigenvalue, column 2 = smallest (pcl::PCA order) */
  float evals[3]; /* descending */
  float scales[3]; /* max - min of the cloud projected on evecs (pcl::PCA::project + getMinMax3D) */
  float cov[9];   /* column-major covariance handed to the eigen solver */
} orc_pca_result;

int orc_pca(const float* xyz, uint64_t n_vertices, int moments_mode, orc_pca_result* out);
int orc_centroid(const float* xyz, uint64_t n_vertices, int moments_mode, double out[3]);
/* aabb_center_origin_generator.cpp:9-17: (max - min) / 2 of the float AABB (half the range, as the reference has it) */
int orc_aabb_center_origin(const float* xyz, uint64_t n_vertices, double out[3]);
int orc_principal_axis_direction(const orc_pca_result* pca, double rotation_offset, double out[3]);
/* PCARotatedDirectionGenerator::generate (pca_rotated_direction_generator.cpp:13-25): the nominal direction of a nested
 * generator rotated about the smallest principal axis */
int orc_pca_rotated_direction(const orc_pca_result* pca, double rotation_angle, const double nominal[3], double out[3]);
/* Eigen 3.4 SelfAdjointEigenSolver<Matrix3f>::compute restated; cov column-major, evecs column-major ascending */
int orc_eigen_solver_3f(const float cov[9], float evals_ascending[3], float evecs_ascending[9]);

/* ---- plane lattice (plane_slicer_raster_planner.cpp:550-586, 460-501) ---- */
typedef struct orc_lattice
{
  double cut_direction[3];
  double cut_normal[3];
  double cut_origin[3];
  double start_loc[3];
  double line_spacing;
  double d_min, d_max;
  uint64_t num_planes;
  int32_t no_cuts; /* set when !bidirectional && d_max < 0 (:569-570) */
  int32_t pad_;
} orc_lattice;

int orc_make_lattice(const orc_pca_result* pca,
                     const double cut_direction[3],
                     const double cut_origin[3],
                     double line_spacing,
                     int bidirectional,
                     orc_lattice* out);

/* ---- vertex normals (normals_from_mesh_faces_modifier.cpp:15-46, utils.cpp:135-162) ---- */
int orc_normals_from_mesh_faces(const float* xyz,
                                uint64_t n_vertices,
                                const uint32_t* tri,
                                uint64_t n_tri,
                                float* normals_out /* 3 per vertex */);

/* ---- tool paths ---- */
typedef struct orc_paths orc_paths; /* opaque: noether::ToolPaths + per-waypoint provenance */

/* planImpl's plane loop (:596-628) on planes [plane_begin, plane_end) of the lattice */
int orc_slice(const float* xyz,
              const float* normals,
              uint64_t n_vertices,
              const uint32_t* tri,
              uint64_t n_tri,
              const orc_lattice* lattice,
              int slice_mode,
              int scan_mode,
              uint64_t plane_begin,
              uint64_t plane_end,
              orc_paths** out);

orc_paths* orc_paths_create(void);
void orc_paths_free(orc_paths* p);
orc_paths* orc_paths_clone(const orc_paths* p);
/* append a tool path / a segment to the last tool path / a waypoint (4x4 column-major) to the last segment */
void orc_paths_add_path(orc_paths* p);
void orc_paths_add_segment(orc_paths* p);
void orc_paths_add_waypoint(orc_paths* p, const double pose16[16]);

uint64_t orc_paths_num_paths(const orc_paths* p);
uint64_t orc_paths_num_segments(const orc_paths* p);
uint64_t orc_paths_num_waypoints(const orc_paths* p);
/* flatten: any pointer may be NULL.  path_off has num_paths+1 entries indexing segments, seg_off has num_segments+1
 * entries indexing waypoints.  edge_lo/edge_hi = mesh edge that generated the cut point (lower-scalar vertex first),
 * plane = lattice index of the tool path. */
void orc_paths_export(const orc_paths* p,
                      uint64_t* path_off,
                      uint64_t* seg_off,
                      int64_t* path_plane,
                      float* pos,
                      float* nrm,
                      uint32_t* edge_lo,
                      uint32_t* edge_hi,
                      double* pose16);

/* RasterPlanner::plan post-ops (raster_planner.cpp:16-35) */
int orc_raster_organization(orc_paths* p);        /* raster_organization_modifier.cpp:7-46 */
int orc_direction_of_travel(orc_paths* p);        /* direction_of_travel_orientation_modifier.cpp:6-48 */
/* legacy post-ops (plane_slicer_legacy_raster_planner.cpp:22-33) */
int orc_join_close_segments(orc_paths* p, double distance);     /* join_close_segments_modifier.cpp:11-39 */
int orc_minimum_segment_length(orc_paths* p, double min_length); /* minimum_segment_length_modifier.cpp:12-37 */
int orc_uniform_spacing(orc_paths* p, double point_spacing);     /* uniform_spacing_modifier.cpp:18-70, degree 1 */

/* tool-path modifiers commonly chained after a raster plan */
int orc_snake_organization(orc_paths* p);                                        /* snake_organization_modifier.cpp:5-22 */
int orc_linear_approach(orc_paths* p, const double offset[3], int n_points);     /* linear_approach_modifier.cpp:33-55 */
int orc_linear_departure(orc_paths* p, const double offset[3], int n_points);    /* linear_departure_modifier.cpp:5-27 */
int orc_offset(orc_paths* p, const double offset16[16]);                         /* offset_modifier.cpp:8-22 (column-major 4x4) */
int orc_fixed_orientation(orc_paths* p, const double ref_x[3]);                  /* fixed_orientation_modifier.cpp:5-33 */
int orc_uniform_orientation(orc_paths* p);                                       /* uniform_orientation_modifier.cpp:5-30 */
int orc_concatenate(orc_paths* p);                                               /* concatenate_modifier.cpp:5-15 */
int orc_tool_drag_orientation(orc_paths* p, double angle_offset, double tool_radius);        /* tool_drag_orientation_modifier.cpp:11-35 */
int orc_biased_tool_drag_orientation(orc_paths* p, double angle_offset, double tool_radius); /* biased_tool_drag_orientation_modifier.cpp:12-35 */
int orc_circular_lead_in(orc_paths* p, double arc_angle, double arc_radius, int n_points);  /* circular_lead_in_modifier.cpp:12-44 */
int orc_circular_lead_out(orc_paths* p, double arc_angle, double arc_radius, int n_points); /* circular_lead_out_modifier.cpp:12-44 */
int orc_moving_average_orientation_smoothing(orc_paths* p, int window_size); /* moving_average_orientation_smoothing_modifier.cpp:6-84 */

/* ---- mesh files: what pcl::io::loadPolygonFile yields for a PLY (noether_gui/src/widgets/tpp_widget.cpp:426;
 *      vtkPLYReader + pcl::io::vtk2mesh restated, oracle_ply.cpp) ---- */
typedef struct orc_mesh orc_mesh;
int orc_read_ply(const uint8_t* data, uint64_t n_bytes, orc_mesh** out);
/* binary STL: vtkSTLReader (Merging on) + vtk2mesh restated (oracle_ply.cpp) */
int orc_read_stl(const uint8_t* data, uint64_t n_bytes, orc_mesh** out);
uint64_t orc_mesh_num_vertices(const orc_mesh* m);
uint64_t orc_mesh_num_faces(const orc_mesh* m);
uint64_t orc_mesh_num_indices(const orc_mesh* m); /* sum of the face sizes */
int orc_mesh_has_normals(const orc_mesh* m);
/* any pointer may be NULL: xyz / nrm 3 floats per vertex, face_size one entry per face, index the concatenated ids */
void orc_mesh_export(const orc_mesh* m, float* xyz, float* nrm, uint32_t* face_size, uint32_t* index);
void orc_mesh_free(orc_mesh* m);

/* ---- face subdivision mesh modifiers (face_subdivision_modifier.cpp, oracle_subdivide.cpp): the cloud is the raw
 *      PCLPointCloud2::data (n_points records of point_step bytes; offsets of x, normal_x (-1 = none), rgba (-1 = none)),
 *      the faces are triangles.  The result is a handle: new cloud (same record layout, more records) + triangles ---- */
int orc_face_midpoint_subdivision(const uint8_t* cloud, uint64_t n_points, uint32_t point_step, uint32_t off_xyz, int32_t off_normal,
                                  int32_t off_rgba, const uint32_t* triangles, uint64_t n_triangles, uint32_t n_iterations,
                                  void** out); /* :202-248 */
/* max_points: stop with ORC_ERR_INVALID once the cloud would grow beyond it (0 = no limit; the reference has none) */
int orc_face_subdivision_by_area(const uint8_t* cloud, uint64_t n_points, uint32_t point_step, uint32_t off_xyz, int32_t off_normal,
                                 int32_t off_rgba, const uint32_t* triangles, uint64_t n_triangles, float max_area,
                                 uint64_t max_points, void** out); /* :250-339 */
uint64_t orc_submesh_num_points(const void* h);
uint64_t orc_submesh_num_triangles(const void* h);
void orc_submesh_export(const void* h, uint8_t* cloud, uint32_t* triangles); /* either may be NULL */
void orc_submesh_free(void* h);

/* ---- EuclideanClusteringMeshModifier::modify (euclidean_clustering_modifier.cpp:40-64; pcl::EuclideanClusterExtraction,
 *      extractSubMeshFromInlierVertices, SimplificationRemoveUnusedVertices restated in oracle_cluster.cpp).  The result is a
 *      handle: the clusters in the order the reference returns their meshes.  force_brute: all-pairs neighbour search
 *      whatever the size (the grid search must agree with it) ---- */
int orc_euclidean_clustering(const uint8_t* cloud, uint64_t n_points, uint32_t point_step, uint32_t off_xyz, const uint32_t* tri,
                             uint64_t n_triangles, double tolerance, int min_cluster_size, int max_cluster_size, int force_brute,
                             void** out);
uint64_t orc_clusters_count(const void* h);
void orc_clusters_labels(const void* h, uint32_t* labels); /* smallest point index of every point's connected component */
void orc_cluster_info(const void* h, uint64_t k, uint64_t out[5]); /* cluster size, sub-mesh points, sub-mesh triangles, whole input, empty */
void orc_cluster_export(const void* h, uint64_t k, uint32_t* indices, uint8_t* cloud, uint32_t* tri); /* any may be NULL */
void orc_clusters_free(void* h);

/* ---- f4': NormalEstimationPCLMeshModifier::modify (normal_estimation_pcl.cpp:15-50; pcl::NormalEstimation with a radius search,
 *      computeMeanAndCovarianceMatrix, pcl::eigen33, flipNormalTowardsViewpoint restated in oracle_normals_pcl.cpp).
 *      normals float[3 * n] (NaN where a point has fewer than 3 neighbours), curvature float[n] or NULL, neighbour_count
 *      uint32[n] or NULL (the size of every radius search, the query point included) ---- */
int orc_normal_estimation_pcl(const float* xyz, uint64_t n_points, double radius, double vx, double vy, double vz, float* normals,
                              float* curvature, uint32_t* neighbour_count);

/* ---- restated test fixtures ---- */
/* createRandomTransform(translation_limit, rotation_limit): tool_path_planner_utest.cpp:32-46 (std::mt19937(0)) */
void orc_fixture_random_transform(double translation_limit, double rotation_limit, double out16[16]);
/* createPlaneMesh(lx, ly, origin): utils.cpp:249-291 -> 4 vertices (xyz + normal), 2 triangles */
void orc_fixture_plane_mesh(float lx, float ly, const double origin16[16], float xyz[12], float nrm[12], uint32_t tri[6]);

#ifdef __cplusplus
}
#endif
#endif
