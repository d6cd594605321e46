// Internal declarations shared by the host side and the CUDA kernels of libnoether_b200.so.
#pragma once
#include "../../include/noether_b200.h"

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

namespace nb2
{
// ---------------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------------
struct Error
{
  int status;
  std::string msg;
};
void setLastError(const std::string& msg);
[[noreturn]] void fail(int status, const std::string& msg);

#define NB2_CUDA(expr)                                                                                                 \
  do                                                                                                                   \
  {                                                                                                                    \
    cudaError_t err__ = (expr);                                                                                        \
    if (err__ != cudaSuccess)                                                                                          \
      ::nb2::fail(NB2_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(err__) + " at " + __FILE__ + ":" +    \
                                    std::to_string(__LINE__) + " (" #expr ")");                                        \
  } while (0)

// every extern "C" entry point: exceptions -> status + thread-local message
#define NB2_API_BEGIN                                                                                                  \
  try                                                                                                                  \
  {
#define NB2_API_END                                                                                                    \
  }                                                                                                                    \
  catch (const nb2::Error& e)                                                                                          \
  {                                                                                                                    \
    nb2::setLastError(e.msg);                                                                                          \
    return e.status;                                                                                                   \
  }                                                                                                                    \
  catch (const std::exception& e)                                                                                      \
  {                                                                                                                    \
    nb2::setLastError(std::string("internal error: ") + e.what());                                                     \
    return NB2_ERR_INTERNAL;                                                                                           \
  }                                                                                                                    \
  return NB2_OK;

// ---------------------------------------------------------------------------------------------------------------
// growable device buffer (arena slot).  Capacity only grows; contents are not preserved across growth.
// ---------------------------------------------------------------------------------------------------------------
struct DevBuf
{
  void* ptr = nullptr;
  size_t cap = 0;
  // a caller-owned device pointer may temporarily stand in for the owned allocation (zero-copy ingest)
  void* own_ptr = nullptr;
  size_t own_cap = 0;
  bool borrowed = false;
  void ensure(size_t bytes);
  void adopt(void* device_ptr);
  void release();
  template <typename T>
  T* as() const
  {
    return static_cast<T*>(ptr);
  }
};

struct PinnedBuf
{
  void* ptr = nullptr;
  size_t cap = 0;
  void ensure(size_t bytes);
  void release();
  template <typename T>
  T* as() const
  {
    return static_cast<T*>(ptr);
  }
};

// ---------------------------------------------------------------------------------------------------------------
// device-side PODs
// ---------------------------------------------------------------------------------------------------------------
/** fp64 partial moments of one thread block: sums of (x - pivot), of its outer product, and float AABB */
struct MomentPartial
{
  double s[3];
  double q[6];  // xx xy xz yy yz zz
  float mn[3];
  float mx[3];
  unsigned long long non_finite;
};

/** Lattice constants handed to kernels by value (plane_slicer_raster_planner.cpp:584-602) */
struct LatticeDev
{
  double n[3];      // cut_normal
  double start[3];  // start_loc
  double spacing;   // line_spacing_
  double dir[3];    // cut_direction (orientation / ordering of chains)
  long long num_planes;
  long long plane_begin;  // slab handled by this launch
  long long plane_end;
};

/** Inputs of the device-side lattice construction (k_frame_lattice) */
struct FrameParams
{
  double line_spacing;
  double direction[3];
  double origin[3];
  double sin_rotation, cos_rotation;  // of PrincipalAxis' rotation_offset, evaluated by the host libm
  int bidirectional;
  int direction_mode, origin_mode;
  int shard_rank, shard_count;
  unsigned plane_cap;  // planes the per-plane arrays are sized for
};

/** Frame state kept in device memory between the kernels of a plan; read back once, with the hit count */
struct FrameDev
{
  nb2_pca pca;
  nb2_lattice lat;
  LatticeDev L;
  unsigned ext[6];  // ordered bits of the min / max of the cloud in the PCA frame (k_extents)
  int have_scales;
  int lattice_status;  // frame::kLattice*
  int plane_cap_exceeded;
  int pad_;
  unsigned long long planes_wanted;
  unsigned long long non_finite;
};

/** Where a binary little-endian PLY file keeps what the slicer needs (ply.cpp parses the header on the host, the
 *  records are decoded on the device by kernels_ply.cu) */
struct PlyLayout
{
  uint64_t header_bytes;
  uint64_t n_points, n_faces;
  uint64_t vertex_begin, face_begin;  // byte offsets into the file image
  uint32_t point_step;                // bytes per vertex record
  uint32_t off_field[6];              // of the properties x y z nx ny nz inside a vertex record
  uint8_t field_f64[6];               // stored as float64 (converted to float as vtkPLYReader does), else float32
  bool has_normals;                   // nx, ny and nz are all present
  uint32_t face_step;                 // bytes per face record when every face is a triangle
  uint32_t face_pre;                  // bytes of scalar properties ahead of the vertex_indices list
  uint32_t face_count_bytes, face_index_bytes;
};
PlyLayout parsePlyHeader(const uint8_t* data, uint64_t n_bytes);

constexpr int kMomentBlocks = 1184;  // fixed (not SM-count derived) so that sums are identical on any device
constexpr int kMomentThreads = 256;

/** Per-stage CUDA-event timing */
struct StageTimer
{
  std::vector<const char*> names;
  std::vector<cudaEvent_t> events;  // events[i] .. events[i+1] brackets stage i
  size_t used = 0;
  void begin(cudaStream_t s);
  void mark(const char* name, cudaStream_t s);
  int collect(const char** out_names, float* ms, int capacity) const;
  void destroy();
};

}  // namespace nb2

// ---------------------------------------------------------------------------------------------------------------
// the context
// ---------------------------------------------------------------------------------------------------------------
struct nb2_context
{
  int device = 0;
  int sm_count = 148;
  size_t smem_optin = 0;
  size_t smem_per_sm = 0;
  size_t l2_window_max = 0;   // largest access-policy window
  size_t l2_persist_max = 0;  // bytes of L2 that may be set aside for persisting lines (0: not available / switched off)
  cudaStream_t stream = nullptr;
  mutable std::mutex mu;
  uint64_t launches = 0;

  // resident mesh (SoA)
  uint64_t V = 0, F = 0;
  bool has_normals = false;
  bool tri_checked = false;  // vertex indices of the resident triangles are known to be < V
  nb2::DevBuf blob;      // staging copy of the interleaved cloud
  std::vector<nb2::DevBuf> areaPool;  // FaceSubdivisionByArea: the per-level buffers, kept between calls
  nb2::DevBuf blobSpare, triSpare;  // nb2_mesh_upload_stream stages into these; they are swapped in when the mesh is accepted
  nb2::PinnedBuf stageRing;         // pinned slots the host cores fill on the way to the device (api.cu, stageToDevice)
  std::vector<cudaEvent_t> stageEvents;  // one per slot: its last copy has left
  nb2::DevBuf pos;       // float4[V]  (x, y, z, unused)
  nb2::DevBuf nrm;       // float4[V], only once normals were computed here or set by the caller
  // Where the slicer reads vertex normals: the uploaded cloud itself (normal_x of vertex v at nrm_base + v * nrm_stride;
  // the ingest pass does not copy them) or the SoA array above (stride 16).
  const uint8_t* nrm_base = nullptr;
  uint32_t nrm_stride = 0;
  nb2::DevBuf tri;       // uint32[3F]
  nb2::DevBuf fileImg;   // image of a mesh file being decoded on the device (nb2_mesh_upload_ply)
  // moments / frame
  nb2::DevBuf partials;  // MomentPartial[kMomentBlocks] + final
  nb2::DevBuf counters;  // misc device counters (zeroed per plan)
  unsigned plan_serial = 0;  // stamp of the report the link kernel writes into pinned host memory
  nb2::DevBuf frame;     // FrameDev
  bool extents_done = false;    // k_extents has run for the resident mesh (OBB extents are in FrameDev::ext)
  bool checks_deferred = false; // mesh adopted from device memory: non-finite / index checks happen at the next sync
  bool pca_valid = false;       // host copy below is current
  nb2_pca pca{};
  unsigned plane_cap = 1u << 16;
  // capacities the post-count stage was last sized for (0 = not yet known: the first plan waits for its hit count)
  uint64_t hit_cap = 0;      // hits (triangle x plane)
  uint32_t scratch_hits = 0; // hits of the largest plane, when it exceeds what fits shared memory
  size_t ingest_smem = 0;  // dynamic shared memory k_ingest_moments_tma was last prepared for
  int ingest_grid = 0;     // grid of the last ingest pass (its block partials: partials[0 .. grid), total at [grid])
  int link_cfg_threads = 0, link_cfg_ctas = 1;  // launch configuration k_link_planes was last prepared for
  size_t link_cfg_smem = 0;
  bool link_cfg_wide = false;
  uint32_t last_general_planes = 0;  // planes of the last plan that the linker's second-generation fast path declined
  uint32_t grid_planes = 0, grid_nmax = 0;  // planes / largest plane of the last plan: launch geometry of k_waypoints
  // slicing workspace
  nb2::DevBuf vertK;       // int32[V]
  nb2::DevBuf planeCnt;    // uint32[P + 1]
  nb2::DevBuf planeOff;    // uint32[P + 1]
  nb2::DevBuf planeCursor; // uint32[P]
  nb2::DevBuf hitRec;      // float4[2 H] endpoint records {x, y, z, rank} of every (triangle, plane) hit, grouped by plane
  nb2::DevBuf wpTag;       // uint32[2 H] first-inserter rank of every waypoint, in per-plane windows
  nb2::DevBuf segLen;      // uint32[H] segment lengths, in per-plane windows
  nb2::DevBuf planePrefix; // uint2[P + 1] waypoints / segments before each plane
  nb2::DevBuf quadList;    // uint32[F / 4] quads (groups of four triangles) cut by this slab, sharded plans only
  bool list_quads = false; // k_count_hits lists them, k_emit_hits walks the list (set per plan)
  nb2::DevBuf planeMeta;   // uint32[2 * P]: waypoints, segments of each plane
  nb2::DevBuf outPos;      // float[3 * Wcap]
  nb2::DevBuf outNrm;      // float[3 * Wcap]
  nb2::DevBuf outEdge;     // uint32[2 * Wcap]
  nb2::DevBuf outSegLen;   // uint32[Scap]
  nb2::DevBuf linkProfile;   // uint64[16] cycles per phase of the linker's fast path (NB2_LINK_PROFILE diagnostics)
  nb2::DevBuf linkScratch; // global scratch for planes too large for shared memory
  // legacy modifiers on the device (kernels_legacy.cu)
  nb2::DevBuf legCum;     // double[W] running chord length of every waypoint inside its (merged) segment
  nb2::DevBuf legSeg;     // uint4[S] surviving merged segments {begin, end, samples, first sample} relative to their plane
  nb2::DevBuf legPlane;   // uint4[2 (P + 1)] per plane {has path, segments, samples} and its exclusive prefix
  nb2::DevBuf legPose;    // double[16 W'] resampled poses
  nb2::DevBuf legPos, legNrm, legSegLen;
  // tool-path modifiers on the device (modifiers.cu)
  nb2::DevBuf modPose[2];  // double[16 W] pose buffers the modifiers of a chain ping-pong between
  nb2::DevBuf modSegOff;   // uint32[S + 1] segment offsets of the current structure
  nb2::DevBuf modSegMap;   // mod::SegMap[S] source of every output segment (restructuring modifiers)
  nb2::DevBuf modSegPath;  // uint32[S] tool path of every segment
  nb2::DevBuf modPathOff;  // uint32[P + 1]
  nb2::DevBuf modPathSign; // int32[P] tool-drag sign of every tool path
  nb2::DevBuf modPos, modNrm;  // float[3 W] compact input / float views of the output
  nb2::DevBuf modRot;          // double[18 n_points] step rotations of a circular lead, for either sign of a tool path
  nb2::DevBuf modEnds;         // double[6 S + 3] first / last translation of every segment, direction of tool path 0
  nb2::DevBuf modQuat;         // double4[W] quaternions of the poses (MovingAverageOrientationSmoothing)
  // STL welding workspace (kernels_stl.cu)
  nb2::DevBuf stlSlots;      // uint32[slots] first corner of the point hashed to each slot
  nb2::DevBuf stlCorner;     // uint32[3 T] slot, later point id, of every facet corner
  nb2::DevBuf stlFlag;       // uint32[3 T + 1] first-occurrence flags, later per-facet keep flags
  nb2::DevBuf stlPrefix;     // uint32[3 T + 1]
  nb2::DevBuf stlKeepPrefix; // uint32[T + 1]
  // face subdivision (kernels_subdivide.cu): the subdivided cloud / triangles ping-pong between two slots, because the
  // resident mesh (blob / tri adopt them in place) is the input of the next subdivision
  nb2::DevBuf subBlob[2], subTri[2];
  nb2::DevBuf subLevels;    // uint32[3 * triangles of levels 2 .. n] of the midpoint scheme
  nb2::DevBuf subKeys;      // uint64[2 * slots]: 16-byte slots { edge key, smallest request sequence number } of the level
  nb2::DevBuf subReq;       // uint32[3 * triangles of the level] slot of every request
  nb2::DevBuf subNodeCnt, subNodeBase;  // uint32[nodes + 1] vertices created by every tree node (pre-order), their prefix
  nb2::DevBuf subNodeMask;  // uint8[nodes] which of the node's three requests create a vertex
  nb2::DevBuf subFresh;     // centroid records of the by-area sweep in creation order
  nb2::DevBuf subRank;      // uint32[centroids] final number of each
  // Euclidean clustering (kernels_cluster.cu): chain head per cell slot, chain link / union-find parent / label per point
  nb2::DevBuf cluHead, cluNext, cluParent, cluLabel;
  nb2::DevBuf cluCellSize;  // uint32[slots] points per cell slot, then 2 x uint64 {chain steps of the link pass, longest walk}
  // record layout of the resident cloud (set by every ingest)
  uint32_t rec_step = 0, rec_off_xyz = 0;
  int32_t rec_off_nrm = -1;
  // normals workspace
  nb2::DevBuf faceNrm;   // float4[F]
  nb2::DevBuf valence;   // uint32[V + 1]
  nb2::DevBuf csrOff;    // uint32[V + 1]
  nb2::DevBuf csrCursor; // uint32[V]
  nb2::DevBuf csrAdj;    // uint32[3F]
  nb2::DevBuf csrRank;   // uint16[3F] slot of every triangle corner inside its vertex's CSR row
  nb2::DevBuf scanState; // uint64[num tiles] look-back status for the generic scan
  // host staging
  nb2::PinnedBuf hostSmall;  // small read-backs
  struct HostBlock
  {
    void* ptr;
    size_t cap;
  };
  std::vector<HostBlock> hostPool;  // pinned blocks recycled between results

  nb2_lattice last_lattice{};
  bool have_lattice = false;
  nb2::StageTimer timer;         // last plan / slice / normals call
  nb2::StageTimer timer_upload;  // last mesh upload
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;  // nb2_timer_start / nb2_timer_stop
  std::vector<cudaEvent_t> chunkEvents;  // lazy result download: one event per chunk of waypoints (reused plan after plan)
  cudaEvent_t ev_span_begin = nullptr, ev_span_end = nullptr;  // nb2_last_plan_span: entry of the call .. end of its last kernel
  bool span_open = false;   // nb2_plan_mesh recorded the begin event: nb2_plan must not move it
  bool span_valid = false;
  double host_t0 = 0.0;          // steady clock at the entry of the planning call ...
  float host_enqueue_ms = 0.f;   // ... and how long the host took to enqueue the plan's launches (reported as "host:enqueue")

  // NCCL (dlopen'ed lazily)
  void* nccl_comm = nullptr;
  int comm_rank = 0, comm_world = 1;
};

// ---------------------------------------------------------------------------------------------------------------
// the result: one pinned host block carved into the SoA arrays (device -> host copies land in their final place)
// ---------------------------------------------------------------------------------------------------------------
struct nb2_result
{
  nb2_context* owner = nullptr;  // the block returns to the owner's pool on nb2_result_free
  void* block = nullptr;
  size_t block_cap = 0;
  uint64_t n_paths = 0, n_segments = 0, n_waypoints = 0;
  uint32_t* path_offsets = nullptr;     // n_paths + 1
  uint32_t* segment_offsets = nullptr;  // n_segments + 1
  int64_t* path_plane = nullptr;        // n_paths
  float* positions = nullptr;           // 3 W
  float* normals = nullptr;             // 3 W
  uint32_t* edge_lo = nullptr;          // W (or null)
  uint32_t* edge_hi = nullptr;
  double* poses = nullptr;              // 16 W, only for legacy (resampled) results
  bool post_ops_applied = false;
  bool legacy = false;
  // A slab of a sharded plan stays in the owner's device buffers (outPos / outNrm / outEdge / outSegLen / planeMeta)
  // until nb2_result_gather ships it to the root over NCCL; valid until the next plan on that context.
  bool device_resident = false;
  unsigned plan_serial = 0;    // the owner's plan_serial when the slab was produced (a later plan overwrites the buffers)
  uint64_t dev_w_cap = 0;      // outEdge holds edge_lo at [0, W) and edge_hi at [dev_w_cap, dev_w_cap + W)
  uint64_t slab_begin = 0;     // first lattice plane of the slab
  uint64_t slab_planes = 0;    // planes in the slab (entries of planeMeta)
  bool dev_edges = false;
  nb2_plan_params params{};
  nb2_lattice lattice{};
  // download = 2: positions / normals arrive in `lazy_chunks` pieces of `chunk_w` waypoints behind the call that made the
  // result; chunk k is complete when the owner's chunkEvents[k] has happened (nb2::waitWaypoints)
  int lazy_chunks = 0;
  uint64_t chunk_w = 0;
  std::atomic<int> chunks_landed{ 0 };
};

// ---------------------------------------------------------------------------------------------------------------
// kernel launchers (implemented in the .cu files)
// ---------------------------------------------------------------------------------------------------------------
namespace nb2
{
void launchIngestMoments(nb2_context* c, uint32_t point_step, uint32_t off_xyz, int32_t off_nrm);
void launchFrameLattice(nb2_context* c, const FrameParams& prm);
void launchFrameSet(nb2_context* c, const nb2_lattice& lat, uint64_t plane_begin, uint64_t plane_end, unsigned plane_cap);
// with lattice_prm the last block of k_extents also builds the plan's lattice and resets the slicing counters
void launchExtents(nb2_context* c, const FrameParams* lattice_prm = nullptr);
// the three kernels before the plan's only mid-way synchronisation read the lattice from device memory (FrameDev::L)
void launchClassify(nb2_context* c);
void launchCountHits(nb2_context* c);  // + scan of the per-plane counters: writes {H, n_max} to counters[4..5]
void launchEmitHits(nb2_context* c, uint64_t hit_cap);
/** Where the last CTA of k_link_planes leaves the plan's report (pinned host memory, device-visible addresses) */
struct LinkReport
{
  unsigned* frame = nullptr;  // FrameDev
  unsigned* sizes = nullptr;  // [16]: counters[4..13]; [15] = stamp once everything above is written
  unsigned* meta = nullptr;   // [2 * planes] leading entries of planeMeta
  unsigned planes = 0;
  unsigned stamp = 0;
};
void launchLinkPlanes(nb2_context* c, uint64_t hit_cap, uint32_t scratch_hits, const LinkReport& report);
void launchWaypoints(nb2_context* c, uint64_t w_cap, int emit_edges);
void launchSetNormals(nb2_context* c, const float* dev_packed);
void launchPackOut(nb2_context* c, float* xyz_packed, float* nrm_packed);  // normals from nrm_base / nrm_stride
void launchVertexNormals(nb2_context* c);
void launchValidateTriangles(nb2_context* c);
// PLY records in c->fileImg -> c->blob (32-byte records {x y z 0 nx ny nz 0}, 16-byte ones without normals) and c->tri;
// counters[16] counts faces that are not triangles, counters[12] receives the largest vertex index
void launchPlyDecode(nb2_context* c, const PlyLayout& L);
// binary STL image in c->fileImg -> welded mesh in c->blob (16-byte records {x y z 0}) and c->tri (kernels_stl.cu)
void launchStlDecode(nb2_context* c, uint64_t n_facets, uint64_t* n_points_out, uint64_t* n_kept_out);
// face subdivision of the resident mesh into subBlob[slot] / subTri[slot] (kernels_subdivide.cu); both synchronise
namespace sub
{
struct Layout;
}
void launchMidpointSubdivision(nb2_context* c, const sub::Layout& L, int n_iterations, int slot, uint64_t* n_points, uint64_t* n_triangles);
void launchAreaSubdivision(nb2_context* c, const sub::Layout& L, float max_area, uint64_t max_points, int slot, uint64_t* n_points,
                           uint64_t* n_triangles);
// connected components of the resident cloud under FLANN's radius test (kernels_cluster.cu); synchronises
namespace clu
{
struct GridParams;
}
void launchClusterLabels(nb2_context* c, const clu::GridParams& g, double max_tests, uint32_t* labels_host);
void launchNormalEstimationPcl(nb2_context* c, const clu::GridParams& g, double max_tests, const float viewpoint[3], float* normals_host,
                               float* curvature_host);
// single-pass decoupled look-back exclusive scan (kernels_normals.cu)
void launchScanU32(nb2_context* c, const unsigned* in, unsigned* out, unsigned long long n);

// host-side frame arithmetic (frame.cpp)
const char* latticeErrorMessage(int code);
void eigenSolver3f(const float A[9], float evals[3], float evecs[9]);
void principalAxisDirection(const nb2_pca& pca, double rotation_offset, double out[3]);
void pcaRotatedDirection(const nb2_pca& pca, double rotation_angle, const double nominal[3], double out[3]);
int makeLattice(const nb2_pca& pca, const double dir[3], const double origin[3], double spacing, int bidirectional,
                nb2_lattice& L);

// PCIe-parallel upload for a sharded job (comm.cpp): rank r copies bytes [r * chunk, (r + 1) * chunk) of `src` to the
// same offsets of `dst`, then an in-place ncclAllGather over NVLink completes every rank's copy.  `dst` must hold
// shardedCapacity(bytes) bytes (the chunk is padded to a multiple of 256).
size_t shardedCapacity(const nb2_context* c, size_t bytes);
void shardedHostToDevice(nb2_context* c, void* dst, const void* src, size_t bytes);

// result storage (api.cu)
nb2_result* allocResult(nb2_context* c, uint64_t n_paths, uint64_t n_segments, uint64_t n_waypoints, bool edges, bool poses);
void freeResult(nb2_result* r);

// legacy modifiers on the device (kernels_legacy.cu): returns {paths, segments, samples} and the per-plane counts
void launchLegacy(nb2_context* c, uint64_t nP, uint64_t W, uint64_t S, double point_spacing, double min_segment_size,
                  double min_hole_size, uint64_t totals[3], std::vector<uint32_t>& host_planeL);

// host-side tool path post-ops (toolpath.cpp)
/** block until waypoints [0, w_end) of a lazily downloaded result are in host memory (no-op for ordinary results) */
void waitWaypoints(const nb2_result& r, uint64_t w_end);
void rasterOrganization(nb2_result*& r);
void expandPoses(const nb2_result& r, double* out16);
void expandPosesSegments(const nb2_result& r, double* const* segment_out);
void legacyModifiers(nb2_result*& r, double point_spacing, double min_segment_size, double min_hole_size);
}  // namespace nb2
