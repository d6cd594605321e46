// C ABI of libnoether_b200.so (include/noether_b200.h): context, mesh ingest, PCA frame, planning, results.
#include "host_math.h"
#include "nb2_internal.h"
#include "subdivide_ops.h"
#include "cluster_ops.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <atomic>
#include <chrono>
#include <functional>
#include <memory>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace nb2
{
namespace
{
thread_local std::string t_last_error;
}
void setLastError(const std::string& msg) { t_last_error = msg; }
void fail(int status, const std::string& msg) { throw Error{ status, msg }; }

void DevBuf::adopt(void* device_ptr)
{
  if (!borrowed)
  {
    own_ptr = ptr;
    own_cap = cap;
  }
  ptr = device_ptr;
  cap = 0;
  borrowed = true;
}
void DevBuf::ensure(size_t bytes)
{
  if (borrowed)
  {
    ptr = own_ptr;
    cap = own_cap;
    borrowed = false;
  }
  if (bytes <= cap)
    return;
  if (ptr)
    NB2_CUDA(cudaFree(ptr));
  ptr = nullptr;
  cap = 0;
  const size_t want = bytes + bytes / 8 + 256;
  NB2_CUDA(cudaMalloc(&ptr, want));
  cap = want;
}
void DevBuf::release()
{
  if (borrowed)
  {
    ptr = own_ptr;
    cap = own_cap;
    borrowed = false;
  }
  if (ptr)
    cudaFree(ptr);
  ptr = nullptr;
  cap = 0;
}
void PinnedBuf::ensure(size_t bytes)
{
  if (bytes <= cap)
    return;
  if (ptr)
    NB2_CUDA(cudaFreeHost(ptr));
  ptr = nullptr;
  cap = 0;
  NB2_CUDA(cudaMallocHost(&ptr, bytes));
  cap = bytes;
}
void PinnedBuf::release()
{
  if (ptr)
    cudaFreeHost(ptr);
  ptr = nullptr;
  cap = 0;
}

// NB2_STAGE_TIMERS=0: no events between the kernels of a call (nb2_last_timings then reports nothing)
static bool stageTimersOn()
{
  static const bool on = [] {
    const char* e = std::getenv("NB2_STAGE_TIMERS");
    return !e || std::atoi(e) != 0;
  }();
  return on;
}
void StageTimer::begin(cudaStream_t s)
{
  used = 0;
  names.clear();
  if (!stageTimersOn())
    return;
  if (events.empty())
  {
    events.resize(24);
    for (auto& e : events)
      NB2_CUDA(cudaEventCreate(&e));
  }
  NB2_CUDA(cudaEventRecord(events[0], s));
}
void StageTimer::mark(const char* name, cudaStream_t s)
{
  if (!stageTimersOn() || used + 2 > events.size())
    return;
  ++used;
  names.push_back(name);
  NB2_CUDA(cudaEventRecord(events[used], s));
}
int StageTimer::collect(const char** out_names, float* ms, int capacity) const
{
  int n = 0;
  for (size_t i = 0; i < used && n < capacity; ++i, ++n)
  {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, events[i], events[i + 1]) != cudaSuccess)
      t = -1.f;
    out_names[n] = names[i];
    ms[n] = t;
  }
  return n;
}
void StageTimer::destroy()
{
  for (auto& e : events)
    cudaEventDestroy(e);
  events.clear();
}

// ---------------------------------------------------------------------------------------------------------------
// result storage
// ---------------------------------------------------------------------------------------------------------------
static size_t align64(size_t x) { return (x + 63) & ~size_t(63); }

nb2_result* allocResult(nb2_context* c, uint64_t P, uint64_t S, uint64_t W, bool edges, bool poses)
{
  size_t off = 0;
  const size_t o_path = off;
  off = align64(off + 4 * (P + 1));
  const size_t o_seg = off;
  off = align64(off + 4 * (S + 1));
  const size_t o_plane = off;
  off = align64(off + 8 * (P + 1));
  const size_t o_pos = off;
  off = align64(off + 12 * (W + 1));
  const size_t o_nrm = off;
  off = align64(off + 12 * (W + 1));
  size_t o_lo = 0, o_hi = 0, o_pose = 0;
  if (edges)
  {
    o_lo = off;
    off = align64(off + 4 * (W + 1));
    o_hi = off;
    off = align64(off + 4 * (W + 1));
  }
  if (poses)
  {
    o_pose = off;
    off = align64(off + 128 * (W + 1));
  }
  const size_t need = off;

  auto* r = new nb2_result();
  r->owner = c;
  {
    // recycle the smallest pooled pinned block that fits
    int best = -1;
    for (size_t i = 0; i < c->hostPool.size(); ++i)
      if (c->hostPool[i].cap >= need && (best < 0 || c->hostPool[i].cap < c->hostPool[best].cap))
        best = static_cast<int>(i);
    if (best >= 0)
    {
      r->block = c->hostPool[best].ptr;
      r->block_cap = c->hostPool[best].cap;
      c->hostPool.erase(c->hostPool.begin() + best);
    }
    else
    {
      const size_t cap = need + need / 8 + 4096;
      cudaError_t err = cudaMallocHost(&r->block, cap);
      if (err != cudaSuccess)
      {
        delete r;
        fail(NB2_ERR_CUDA, std::string("cudaMallocHost failed: ") + cudaGetErrorString(err));
      }
      r->block_cap = cap;
    }
  }
  auto* base = static_cast<uint8_t*>(r->block);
  r->n_paths = P;
  r->n_segments = S;
  r->n_waypoints = W;
  r->path_offsets = reinterpret_cast<uint32_t*>(base + o_path);
  r->segment_offsets = reinterpret_cast<uint32_t*>(base + o_seg);
  r->path_plane = reinterpret_cast<int64_t*>(base + o_plane);
  r->positions = reinterpret_cast<float*>(base + o_pos);
  r->normals = reinterpret_cast<float*>(base + o_nrm);
  r->edge_lo = edges ? reinterpret_cast<uint32_t*>(base + o_lo) : nullptr;
  r->edge_hi = edges ? reinterpret_cast<uint32_t*>(base + o_hi) : nullptr;
  r->poses = poses ? reinterpret_cast<double*>(base + o_pose) : nullptr;
  r->path_offsets[0] = 0;
  r->segment_offsets[0] = 0;
  return r;
}

// contexts that are still alive: a result may outlive its context (garbage-collected bindings), in which case its
// pinned block is released directly instead of going back to the context's pool
std::mutex g_live_mu;
std::vector<nb2_context*> g_live;
void registerContext(nb2_context* c)
{
  std::lock_guard<std::mutex> l(g_live_mu);
  g_live.push_back(c);
}
void unregisterContext(nb2_context* c)
{
  std::lock_guard<std::mutex> l(g_live_mu);
  g_live.erase(std::remove(g_live.begin(), g_live.end(), c), g_live.end());
}
bool contextAlive(nb2_context* c)
{
  std::lock_guard<std::mutex> l(g_live_mu);
  return std::find(g_live.begin(), g_live.end(), c) != g_live.end();
}

void waitWaypoints(const nb2_result& r, uint64_t w_end)
{
  if (!r.lazy_chunks || !r.chunk_w)
    return;
  auto& landed = const_cast<nb2_result&>(r).chunks_landed;
  const int want = static_cast<int>(std::min<uint64_t>(r.lazy_chunks, (w_end + r.chunk_w - 1) / r.chunk_w));
  if (landed.load(std::memory_order_acquire) >= want)
    return;
  if (r.owner && contextAlive(r.owner) && want > 0)
  {
    // (the events are the owner's and are recorded again by its next lazy plan: the later point in the stream they then
    //  stand for implies this one)
    cudaSetDevice(r.owner->device);
    if (cudaEventSynchronize(r.owner->chunkEvents[want - 1]) != cudaSuccess)
      cudaGetLastError();
  }
  int seen = landed.load();
  while (seen < want && !landed.compare_exchange_weak(seen, want, std::memory_order_release))
  {
  }
}

void freeResult(nb2_result* r)
{
  if (!r)
    return;
  if (r->lazy_chunks)
    waitWaypoints(*r, r->n_waypoints);  // copies still in flight write into the block
  if (r->block && !r->owner)
    cudaFreeHost(r->block);
  if (r->block && r->owner)
  {
    auto& pool = r->owner->hostPool;
    if (pool.size() < 8)
      pool.push_back({ r->block, r->block_cap });
    else
      cudaFreeHost(r->block);
  }
  delete r;
}

// ---------------------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------------------
namespace
{
struct DeviceGuard
{
  int prev = -1;
  explicit DeviceGuard(int dev)
  {
    cudaGetDevice(&prev);
    if (prev != dev)
      NB2_CUDA(cudaSetDevice(dev));
    else
      prev = -1;
  }
  ~DeviceGuard()
  {
    if (prev >= 0)
      cudaSetDevice(prev);
  }
};

void requireMesh(const nb2_context* c)
{
  if (c->V == 0)
    fail(NB2_ERR_NO_MESH, "no mesh has been uploaded to this context");
}

const char* kNoNormalsMessage =
    "The input mesh does not have vertex normals, which are required for the plane slice raster tool path planner. Use a "
    "MeshModifier to generate vertex normals for the mesh, or provide a different mesh with vertex normals.";

/** k_extents once per resident mesh (the eigen axes were finalised on the device by the ingest pass); true when the
 *  kernel ran.  With `lattice_prm` its last block also runs the lattice step of the plan. */
bool ensureExtents(nb2_context* c, const FrameParams* lattice_prm = nullptr)
{
  if (c->extents_done || c->V < 3)
    return false;
  launchExtents(c, lattice_prm);
  c->timer.mark(lattice_prm ? "extents+lattice" : "extents", c->stream);
  c->extents_done = true;
  return true;
}

/** Checks that were postponed because the mesh was adopted from device memory without a synchronisation */
void deferredMeshChecks(nb2_context* c, const FrameDev& f)
{
  if (!c->checks_deferred)
    return;
  c->checks_deferred = false;
  if (f.non_finite)
  {
    c->V = 0;
    fail(NB2_ERR_NON_FINITE, std::to_string(f.non_finite) + " vertices have non-finite coordinates");
  }
}

float fromOrderedBits(unsigned u)
{
  const unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  float f;
  std::memcpy(&f, &v, 4);
  return f;
}

/** Host copy of the PCA frame (mean, axes, OBB extents): one read-back of the device frame state */
void ensureHostPca(nb2_context* c)
{
  requireMesh(c);
  if (c->pca_valid)
    return;
  ensureExtents(c);
  c->hostSmall.ensure(4096 + sizeof(FrameDev));
  auto* hf = c->hostSmall.as<FrameDev>();
  NB2_CUDA(cudaMemcpyAsync(hf, c->frame.ptr, sizeof(FrameDev), cudaMemcpyDeviceToHost, c->stream));
  NB2_CUDA(cudaStreamSynchronize(c->stream));
  deferredMeshChecks(c, *hf);
  c->pca = hf->pca;
  if (!hf->have_scales && c->V >= 3)
    for (int i = 0; i < 3; ++i)
      c->pca.scales[i] = fromOrderedBits(hf->ext[3 + i]) - fromOrderedBits(hf->ext[i]);
  c->pca_valid = true;
}

void requirePcaPoints(const nb2_context* c)
{
  if (c->V < 3)
    fail(NB2_ERR_TOO_FEW_POINTS, "[pcl::PCA::initCompute] number of points < 3");
}

nb2_result* emptyResult(nb2_context* c, const nb2_lattice& L)
{
  nb2_result* r = allocResult(c, 0, 0, 0, true, false);
  r->lattice = L;
  return r;
}

/**
 * planImpl's plane loop on the device.  The lattice comes either from the device-side construction (`fp`: PCA axes,
 * generators and lattice arithmetic run as single-thread kernels) or from the caller (`explicit_lat`, nb2_slice).
 *
 * Every kernel reads the sizes it depends on (plane count, hit count, largest plane) from device memory, so the whole
 * plan is enqueued without waiting for any of them: the buffers behind the hit count are sized from the capacities of
 * the previous plan on this context, and a kernel that finds a capacity too small writes nothing and raises a flag, on
 * which the host grows the buffers and repeats that stage.  Only the first plan of a context waits for its hit count.
 */
/** nb2_last_plan_span: the begin event at the entry of the planning call, the end event behind the plan's last kernel */
static void spanBegin(nb2_context* c)
{
  if (!c->ev_span_begin)
  {
    NB2_CUDA(cudaEventCreate(&c->ev_span_begin));
    NB2_CUDA(cudaEventCreate(&c->ev_span_end));
  }
  c->span_valid = false;
  c->host_t0 = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  NB2_CUDA(cudaEventRecord(c->ev_span_begin, c->stream));
}
static void spanEnd(nb2_context* c)
{
  if (c->ev_span_end)
  {
    NB2_CUDA(cudaEventRecord(c->ev_span_end, c->stream));
    c->span_valid = true;
    c->host_enqueue_ms = static_cast<float>(
        1e3 * (std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() - c->host_t0));
  }
}

nb2_result* sliceImpl(nb2_context* c, const FrameParams* fp, const nb2_lattice* explicit_lat, uint64_t plane_begin,
                      uint64_t plane_end, bool emit_edges, bool download = true, bool lazy = false)
{
  requireMesh(c);
  if (!c->has_normals)
    fail(NB2_ERR_NO_NORMALS, kNoNormalsMessage);
  // a slab of a sharded plan is cut by a fraction of the triangles: the counting pass lists those for the contouring pass
  c->list_quads = fp && fp->shard_count > 1;
  if (const char* e = std::getenv("NB2_LIST_QUADS"))
    c->list_quads = std::atoi(e) != 0;

  FrameDev* hf = nullptr;
  unsigned* h = nullptr;
  // counters: [4] H, [5] n_max, [6] waypoints, [7] segments, [8] link ticket, [10] status, [11] failing plane,
  //           [9] quads on the cut-quad list, [12] largest out-of-range vertex index, [13] capacity overflow bits
  // (start-of-plan state: written by the lattice step on the device, resetPlanCounters)
  auto resetCounters = [&](int first, int last = 16) {
    unsigned init[16] = { 0 };
    init[11] = 0xffffffffu;
    NB2_CUDA(cudaMemcpyAsync(c->counters.as<unsigned>() + first, init + first, (last - first) * sizeof(unsigned),
                             cudaMemcpyHostToDevice, c->stream));
  };
  // The per-plane {waypoints, segments} counts ride along with the sizes when the plane count of the previous plan is
  // known (it sizes the copy; a plan with more planes fetches the rest afterwards): one synchronisation at the end
  // of a plan instead of two.
  const uint64_t meta_hint = std::min<uint64_t>(c->grid_planes, c->plane_cap);
  c->hostSmall.ensure(8192 + sizeof(FrameDev) + sizeof(unsigned) * 2 * (meta_hint + 1));
  hf = c->hostSmall.as<FrameDev>();
  h = reinterpret_cast<unsigned*>(c->hostSmall.as<uint8_t>() + ((sizeof(FrameDev) + 63) & ~size_t(63)));
  bool meta_fetched = false;
  // After a complete stage two the sizes need no copy at all: the last CTA of k_link_planes has written them (frame,
  // counters, leading per-plane counts) into this pinned staging area, stamped with the plan's serial number.
  LinkReport report;
  report.frame = reinterpret_cast<unsigned*>(hf);
  report.sizes = h;
  report.meta = h + 16;
  report.planes = static_cast<unsigned>(meta_hint);
  auto readSizes = [&](bool with_meta = false) {
    if (with_meta)
    {
      NB2_CUDA(cudaStreamSynchronize(c->stream));
      if (h[15] == report.stamp)
      {
        meta_fetched = meta_hint > 0;
        return;
      }
    }
    NB2_CUDA(cudaMemcpyAsync(hf, c->frame.ptr, sizeof(FrameDev), cudaMemcpyDeviceToHost, c->stream));
    NB2_CUDA(cudaMemcpyAsync(h, c->counters.as<unsigned>() + 4, 11 * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    meta_fetched = with_meta && meta_hint > 0;
    if (meta_fetched)
      NB2_CUDA(cudaMemcpyAsync(h + 16, c->planeMeta.ptr, sizeof(unsigned) * 2 * meta_hint, cudaMemcpyDeviceToHost, c->stream));
    NB2_CUDA(cudaStreamSynchronize(c->stream));
  };
  auto setCapacities = [&](uint64_t H, uint32_t n_max) {
    c->hit_cap = H + H / 4 + 1024;
    c->scratch_hits = n_max + n_max / 4;  // (any plane may need the first-generation paths' global work area)
    c->grid_nmax = n_max;
  };
  auto stageTwo = [&]() {
    const uint64_t cap = c->hit_cap, w_cap = 2 * cap;
    c->hitRec.ensure(sizeof(float4) * 2 * cap);
    c->outPos.ensure(sizeof(float) * 3 * w_cap);
    c->outNrm.ensure(sizeof(float) * 3 * w_cap);
    if (emit_edges)
      c->outEdge.ensure(sizeof(uint32_t) * 2 * w_cap);
    c->outSegLen.ensure(sizeof(unsigned) * cap);
    launchEmitHits(c, cap);
    c->timer.mark("emit", c->stream);
    report.stamp = ++c->plan_serial ? c->plan_serial : ++c->plan_serial;  // never 0
    h[15] = 0;
    launchLinkPlanes(c, cap, c->scratch_hits, report);
    c->timer.mark("link", c->stream);
    launchWaypoints(c, w_cap, emit_edges ? 1 : 0);
    c->timer.mark("waypoints", c->stream);
    spanEnd(c);
  };

  bool sizes_known = false;
  for (int attempt = 0;; ++attempt)
  {
    if (fp)
    {
      FrameParams p = *fp;
      p.plane_cap = c->plane_cap;
      // first plan on this mesh: the lattice step rides on the extents reduction
      if (!ensureExtents(c, &p))
      {
        launchFrameLattice(c, p);
        c->timer.mark("lattice", c->stream);
      }
    }
    else
    {
      launchFrameSet(c, *explicit_lat, plane_begin, plane_end, c->plane_cap);
      c->timer.mark("lattice", c->stream);
    }
    launchClassify(c);
    c->timer.mark("classify", c->stream);
    launchCountHits(c);
    c->timer.mark("count+scan", c->stream);

    sizes_known = false;
    if (c->hit_cap == 0)
    {
      // first plan on this context: nothing to size the buffers from, wait for the hit count
      readSizes();
      sizes_known = true;
      if (!hf->plane_cap_exceeded && hf->lattice_status == 0)
      {
        setCapacities(h[0], h[1]);
        c->grid_planes = static_cast<uint32_t>(hf->L.plane_end - hf->L.plane_begin);
      }
    }
    if (!sizes_known || (c->hit_cap && !hf->plane_cap_exceeded && hf->lattice_status == 0 && h[0] > 0))
    {
      stageTwo();
      readSizes(true);
      sizes_known = true;
    }
    deferredMeshChecks(c, *hf);
    if (hf->lattice_status != 0)
      fail(NB2_ERR_INVALID_ARGUMENT, latticeErrorMessage(hf->lattice_status));
    if (!hf->plane_cap_exceeded)
      break;
    if (attempt > 0 || hf->planes_wanted >= (1ull << 31))
      fail(NB2_ERR_INTERNAL, "could not size the per-plane arrays");
    unsigned cap = c->plane_cap;
    while (cap < hf->planes_wanted)
      cap *= 2;
    c->plane_cap = cap;  // grow and run the plan again
  }
  if (hf->have_scales)
  {
    c->pca = hf->pca;
    c->pca_valid = true;
  }
  const nb2_lattice lat = hf->lat;
  const LatticeDev L = hf->L;
  c->last_lattice = lat;
  c->have_lattice = true;
  const uint64_t nP = static_cast<uint64_t>(L.plane_end - L.plane_begin);
  plane_begin = static_cast<uint64_t>(L.plane_begin);

  const uint64_t H = h[0];
  const uint32_t n_max = h[1];
  c->last_general_planes = h[10];  // counters[14]: planes the linker's second-generation fast path declined
  c->grid_planes = static_cast<uint32_t>(nP);  // launch geometry hints for the next plan
  c->grid_nmax = n_max;
  if (c->F && h[8] >= c->V)  // counters[12]: largest out-of-range vertex index met by k_count_hits
    fail(NB2_ERR_INVALID_ARGUMENT, "triangle references vertex " + std::to_string(h[8]) + " but the cloud has " +
                                       std::to_string(c->V) + " points");
  c->tri_checked = true;
  if (nP == 0 || H == 0)
  {
    nb2_result* r = emptyResult(c, lat);
    if (!download)
    {
      r->device_resident = true;
      r->plan_serial = c->plan_serial;
      r->dev_edges = emit_edges;
      r->slab_begin = plane_begin;
      r->slab_planes = 0;
    }
    return r;
  }
  if (H >= (1ull << 30))
    fail(NB2_ERR_INVALID_ARGUMENT, "more than 2^30 triangle/plane intersections in one slab; shard the plan");
  if (n_max >= (1u << 19))
    fail(NB2_ERR_INVALID_ARGUMENT, "a single cutting plane intersects 2^19 or more triangles; not supported");
  if (h[9])  // counters[13]: a capacity did not suffice; the kernels left the buffers alone
  {
    setCapacities(H, n_max);
    NB2_CUDA(cudaMemsetAsync(c->planeCursor.ptr, 0, sizeof(unsigned) * (static_cast<size_t>(c->plane_cap) + 1), c->stream));
    resetCounters(6, 9);  // (not [9]: the list of cut quads k_count_hits left stays valid)
    resetCounters(10);
    stageTwo();
    readSizes(true);
    if (h[9])
      fail(NB2_ERR_INTERNAL, "buffer capacities could not be settled");
  }
  const uint64_t w_cap = 2 * c->hit_cap;

  // per-plane meta and status (h[2..3] = counters[6..7] = waypoint / segment totals, h[6..7] = counters[10..11] = link
  // status).  The meta came with the sizes unless this plan has more planes than the previous one.
  unsigned* hmeta = h + 16;
  if (!meta_fetched || nP > meta_hint)
  {
    c->hostSmall.ensure(8192 + sizeof(FrameDev) + sizeof(unsigned) * 2 * (nP + 1));  // (may move the staging area)
    h = reinterpret_cast<unsigned*>(c->hostSmall.as<uint8_t>() + ((sizeof(FrameDev) + 63) & ~size_t(63)));
    hmeta = h + 16;
    NB2_CUDA(cudaMemcpyAsync(h, c->counters.as<unsigned>() + 4, 11 * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    NB2_CUDA(cudaMemcpyAsync(hmeta, c->planeMeta.ptr, sizeof(unsigned) * 2 * nP, cudaMemcpyDeviceToHost, c->stream));
    NB2_CUDA(cudaStreamSynchronize(c->stream));
  }
  const unsigned* hstat = h + 6;
  if (hstat[0])
  {
    const std::string where = " (lattice plane " + std::to_string(plane_begin + hstat[1]) + ")";
    if (hstat[0] & 1u)
      fail(NB2_ERR_NON_MANIFOLD, "cut point with more than two incident cut lines: the mesh is not a manifold surface there" + where);
    if (hstat[0] & 2u)
      fail(NB2_ERR_CLOSED_LOOP, "cut lines form a closed loop; the reference aborts such plans (Topological sort failed / "
                                "not a DAG)" + where);
    fail(NB2_ERR_INTERNAL, "linking failed" + where);
  }
  const uint64_t Wt = h[2];
  const uint64_t St = h[3];
  uint64_t Pn = 0;
  for (uint64_t k = 0; k < nP; ++k)
    Pn += hmeta[2 * k + 1] > 0;

  if (!download)
  {
    // tool paths stay in HBM: report the counts only
    nb2_result* r = allocResult(c, 0, 0, 0, false, false);
    r->lattice = lat;
    r->n_paths = Pn;
    r->n_segments = St;
    r->n_waypoints = Wt;
    r->path_offsets = r->segment_offsets = nullptr;
    r->path_plane = nullptr;
    r->positions = r->normals = nullptr;
    r->device_resident = true;
    r->plan_serial = c->plan_serial;
    r->dev_w_cap = w_cap;
    r->dev_edges = emit_edges;
    r->slab_begin = plane_begin;
    r->slab_planes = nP;
    return r;
  }
  nb2_result* r = allocResult(c, Pn, St, Wt, emit_edges, false);
  r->lattice = lat;
  try
  {
    // lazy: the segment lengths first and alone behind the synchronisation below; positions and normals follow in chunks,
    // each with an event, while the caller already sizes its containers and expands the chunks that have landed
    const bool in_chunks = lazy && !emit_edges && Wt >= (1u << 18);
    if (Wt && in_chunks)
    {
      NB2_CUDA(cudaMemcpyAsync(r->segment_offsets + 1, c->outSegLen.ptr, sizeof(unsigned) * St, cudaMemcpyDeviceToHost,
                               c->stream));
      c->timer.mark("d2h", c->stream);
      NB2_CUDA(cudaStreamSynchronize(c->stream));
      constexpr int kChunks = 16;
      if (c->chunkEvents.empty())
      {
        c->chunkEvents.resize(kChunks);
        for (auto& e : c->chunkEvents)
          NB2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      }
      r->chunk_w = (Wt + kChunks - 1) / kChunks;
      r->lazy_chunks = static_cast<int>((Wt + r->chunk_w - 1) / r->chunk_w);
      for (int k = 0; k < r->lazy_chunks; ++k)
      {
        const uint64_t w0 = k * r->chunk_w, wn = std::min<uint64_t>(r->chunk_w, Wt - w0);
        NB2_CUDA(cudaMemcpyAsync(r->positions + 3 * w0, c->outPos.as<float>() + 3 * w0, sizeof(float) * 3 * wn, cudaMemcpyDeviceToHost, c->stream));
        NB2_CUDA(cudaMemcpyAsync(r->normals + 3 * w0, c->outNrm.as<float>() + 3 * w0, sizeof(float) * 3 * wn, cudaMemcpyDeviceToHost, c->stream));
        NB2_CUDA(cudaEventRecord(c->chunkEvents[k], c->stream));
      }
    }
    else if (Wt)
    {
      NB2_CUDA(cudaMemcpyAsync(r->positions, c->outPos.ptr, sizeof(float) * 3 * Wt, cudaMemcpyDeviceToHost, c->stream));
      NB2_CUDA(cudaMemcpyAsync(r->normals, c->outNrm.ptr, sizeof(float) * 3 * Wt, cudaMemcpyDeviceToHost, c->stream));
      if (emit_edges)
      {
        NB2_CUDA(cudaMemcpyAsync(r->edge_lo, c->outEdge.ptr, sizeof(uint32_t) * Wt, cudaMemcpyDeviceToHost, c->stream));
        NB2_CUDA(cudaMemcpyAsync(r->edge_hi, c->outEdge.as<uint32_t>() + w_cap, sizeof(uint32_t) * Wt,
                                 cudaMemcpyDeviceToHost, c->stream));
      }
      NB2_CUDA(cudaMemcpyAsync(r->segment_offsets + 1, c->outSegLen.ptr, sizeof(unsigned) * St, cudaMemcpyDeviceToHost,
                               c->stream));
    }
    if (!in_chunks)
    {
      c->timer.mark("d2h", c->stream);
      NB2_CUDA(cudaStreamSynchronize(c->stream));
    }
  }
  catch (...)
  {
    freeResult(r);
    throw;
  }
  // segment lengths -> offsets; planes with segments -> tool paths
  for (uint64_t s = 0; s < St; ++s)
    r->segment_offsets[s + 1] += r->segment_offsets[s];
  uint64_t p = 0, s_acc = 0;
  for (uint64_t k = 0; k < nP; ++k)
  {
    const unsigned sk = hmeta[2 * k + 1];
    if (!sk)
      continue;
    r->path_offsets[p] = static_cast<uint32_t>(s_acc);
    r->path_plane[p] = static_cast<int64_t>(plane_begin + k);
    s_acc += sk;
    ++p;
  }
  r->path_offsets[p] = static_cast<uint32_t>(s_acc);
  return r;
}

/** JoinCloseSegments + MinimumSegmentLength + UniformSpacing on the slab `raw` left on the device (kernels_legacy.cu);
 *  consumes `raw`, returns the resampled result (explicit poses) on the host. */
nb2_result* legacyOnDevice(nb2_context* c, nb2_result* raw, const nb2_plan_params& prm)
{
  const uint64_t nP = raw->slab_planes, plane_begin = raw->slab_begin;
  const nb2_lattice lat = raw->lattice;
  uint64_t totals[3] = { 0, 0, 0 };
  std::vector<uint32_t> planeL;
  if (nP && raw->n_waypoints)
    launchLegacy(c, nP, raw->n_waypoints, raw->n_segments, prm.point_spacing, prm.min_segment_size, prm.min_hole_size,
                 totals, planeL);
  nb2_result* out = allocResult(c, totals[0], totals[1], totals[2], true, true);
  out->legacy = true;
  out->post_ops_applied = false;
  out->params = prm;
  out->lattice = lat;
  try
  {
    if (totals[2])
    {
      NB2_CUDA(cudaMemcpyAsync(out->poses, c->legPose.ptr, sizeof(double) * 16 * totals[2], cudaMemcpyDeviceToHost, c->stream));
      NB2_CUDA(cudaMemcpyAsync(out->positions, c->legPos.ptr, sizeof(float) * 3 * totals[2], cudaMemcpyDeviceToHost, c->stream));
      NB2_CUDA(cudaMemcpyAsync(out->normals, c->legNrm.ptr, sizeof(float) * 3 * totals[2], cudaMemcpyDeviceToHost, c->stream));
      NB2_CUDA(cudaMemcpyAsync(out->segment_offsets + 1, c->legSegLen.ptr, sizeof(unsigned) * totals[1], cudaMemcpyDeviceToHost,
                               c->stream));
    }
    c->timer.mark("d2h", c->stream);
    NB2_CUDA(cudaStreamSynchronize(c->stream));
  }
  catch (...)
  {
    freeResult(out);
    throw;
  }
  for (uint64_t s = 0; s < totals[1]; ++s)
    out->segment_offsets[s + 1] += out->segment_offsets[s];
  std::memset(out->edge_lo, 0xff, sizeof(uint32_t) * totals[2]);  // resampled waypoints lie on no single mesh edge
  std::memset(out->edge_hi, 0xff, sizeof(uint32_t) * totals[2]);
  uint64_t p = 0, s_acc = 0;
  for (uint64_t k = 0; k < nP && !planeL.empty(); ++k)
  {
    if (!planeL[4 * k])
      continue;
    out->path_offsets[p] = static_cast<uint32_t>(s_acc);
    out->path_plane[p] = static_cast<int64_t>(plane_begin + k);
    s_acc += planeL[4 * k + 1];
    ++p;
  }
  out->path_offsets[p] = static_cast<uint32_t>(s_acc);
  freeResult(raw);
  return out;
}

}  // namespace

void directionOfTravelInPlace(nb2_result& r);

}  // namespace nb2

using namespace nb2;

extern "C" {

int nb2_abi_version(void) { return NB2_ABI_VERSION; }
const char* nb2_last_error(void) { return t_last_error.c_str(); }

int nb2_context_create(int device, nb2_context** out)
{
  NB2_API_BEGIN
  if (!out)
    fail(NB2_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count == 0)
    fail(NB2_ERR_CUDA, std::string("no CUDA device available (this library has no CPU fallback): ") +
                           (err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0"));
  if (device < 0 || device >= count)
    fail(NB2_ERR_INVALID_ARGUMENT, "device ordinal out of range");
  DeviceGuard g(device);
  cudaDeviceProp prop{};
  NB2_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    fail(NB2_ERR_CUDA, std::string("device '") + prop.name + "' is not a Blackwell (sm_100a) GPU; this library ships sm_100a code only");
  auto c = std::make_unique<nb2_context>();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  c->smem_per_sm = prop.sharedMemPerMultiprocessor;
  NB2_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  // part of L2 may be set aside for lines that kernels on this stream mark as persisting: the SoA positions, which the
  // ingest pass writes and the extents / classification / contouring / waypoint passes of the same plan read again
  c->l2_persist_max = 0;
  // (measured on cfg3, profiles/r02_summary.md: extents and classification gain 5 us between them, but the contouring and
  //  linking passes lose 40 us to the smaller remainder of L2, so the window is opt-in: NB2_L2_PERSIST=1)
  if (prop.persistingL2CacheMaxSize > 0 && std::getenv("NB2_L2_PERSIST") && std::atoi(std::getenv("NB2_L2_PERSIST")) != 0)
  {
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, prop.persistingL2CacheMaxSize) == cudaSuccess)
    {
      c->l2_persist_max = prop.persistingL2CacheMaxSize;
      c->l2_window_max = prop.accessPolicyMaxWindowSize;
    }
    else
      cudaGetLastError();
  }
  c->counters.ensure(64 * sizeof(unsigned));
  NB2_CUDA(cudaMemsetAsync(c->counters.ptr, 0, 64 * sizeof(unsigned), c->stream));
  NB2_CUDA(cudaStreamSynchronize(c->stream));
  registerContext(c.get());
  *out = c.release();
  NB2_API_END
}

void nb2_context_destroy(nb2_context* c)
{
  if (!c)
    return;
  unregisterContext(c);
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  if (c->stream)
    cudaStreamSynchronize(c->stream);
  nb2_comm_destroy(c);
  for (DevBuf* b : { &c->blob, &c->blobSpare, &c->triSpare, &c->fileImg, &c->pos, &c->nrm, &c->tri, &c->partials, &c->frame, &c->counters, &c->vertK, &c->planeCnt,
                     &c->planeOff, &c->planeCursor, &c->hitRec, &c->linkProfile, &c->wpTag, &c->segLen, &c->planePrefix, &c->quadList, &c->legCum, &c->legSeg, &c->legPlane, &c->legPose, &c->legPos,
                     &c->legNrm, &c->legSegLen, &c->planeMeta, &c->outPos, &c->outNrm,
                     &c->outEdge, &c->outSegLen, &c->linkScratch, &c->faceNrm, &c->valence, &c->csrOff, &c->csrCursor,
                     &c->csrAdj, &c->csrRank, &c->scanState, &c->modPose[0], &c->modPose[1], &c->modSegOff, &c->modSegMap,
                     &c->modSegPath, &c->modPathOff, &c->modPathSign, &c->modPos, &c->modNrm, &c->modQuat, &c->modRot, &c->modEnds, &c->stlSlots, &c->stlCorner,
                     &c->stlFlag, &c->stlPrefix, &c->stlKeepPrefix, &c->subBlob[0], &c->subBlob[1], &c->subTri[0], &c->subTri[1], &c->subLevels,
                     &c->subKeys, &c->cluHead, &c->cluNext, &c->cluParent, &c->cluLabel, &c->cluCellSize, &c->subReq, &c->subNodeCnt, &c->subNodeBase, &c->subNodeMask, &c->subFresh, &c->subRank })
    b->release();
  for (auto& b : c->areaPool)
    b.release();
  c->areaPool.clear();
  c->hostSmall.release();
  c->stageRing.release();
  for (auto& e : c->stageEvents)
    cudaEventDestroy(e);
  c->stageEvents.clear();
  for (auto& hb : c->hostPool)
    cudaFreeHost(hb.ptr);
  c->timer.destroy();
  c->timer_upload.destroy();
  for (auto& e : c->chunkEvents)
    cudaEventDestroy(e);
  c->chunkEvents.clear();
  if (c->ev_span_begin)
  {
    cudaEventDestroy(c->ev_span_begin);
    cudaEventDestroy(c->ev_span_end);
  }
  if (c->stream)
    cudaStreamDestroy(c->stream);
  if (prev >= 0)
    cudaSetDevice(prev);
  delete c;
}

int nb2_context_device(const nb2_context* c) { return c ? c->device : -1; }

// ---------------------------------------------------------------------------------------------------------------
// Host -> device staging.  A pageable source (pcl::PCLPointCloud2::data is a std::vector) reaches the GPU at a fraction
// of the PCIe rate through cudaMemcpy's own bounce buffer, and what the plan needs of a PointNormal cloud is 24 of its 48
// bytes per vertex.  So host input is produced in pieces by the host cores — copied, compacted to {x y z [nx ny nz]}
// records, or generated by the caller (nb2_mesh_upload_stream: the plugin flattens one pcl::Vertices per face and
// fingerprints the content in the same pass) — straight into a ring of pinned slots, and every slot goes out with its own
// cudaMemcpyAsync while the cores fill the next ones.  One cudaMemcpyAsync per slot, all on the context's stream.
// ---------------------------------------------------------------------------------------------------------------
namespace
{
struct SourceError
{
  int code;
};
constexpr uint64_t kGrain = NB2_UPLOAD_GRAIN;
constexpr size_t kSlotBytes = size_t(3) << 19;  // (1.5 MB x 24 slots measured best on the 16-core bench host: the ring stays in its 60 MB L3)
constexpr unsigned kSlots = 24;
// experiments: NB2_STAGING_SLOT_KB / NB2_STAGING_SLOTS shrink the ring (a ring that fits the host's last-level cache is
// written and read by the copy engine without touching host DRAM)
size_t slotBytes()
{
  static const size_t n = [] {
    const char* e = std::getenv("NB2_STAGING_SLOT_KB");
    return e ? std::max<size_t>(size_t(64) << 10, static_cast<size_t>(std::atoll(e)) << 10) : kSlotBytes;
  }();
  return n;
}
unsigned slotCount()
{
  static const unsigned n = [] {
    const char* e = std::getenv("NB2_STAGING_SLOTS");
    return e ? static_cast<unsigned>(std::min(64, std::max(2, std::atoi(e)))) : kSlots;
  }();
  return n;
}

struct StageSource
{
  uint32_t step = 0;  // bytes per staged cloud record
  std::function<int(uint64_t first, uint64_t count, void* dst)> cloud;
  std::function<int(uint64_t first, uint64_t count, uint32_t* dst)> tri;
};

unsigned stagingWorkers()
{
  static const unsigned n = [] {
    unsigned hw = std::thread::hardware_concurrency();
    if (const char* e = std::getenv("NB2_STAGING_THREADS"))
      hw = static_cast<unsigned>(std::max(1, std::atoi(e)));
    return std::min(std::max(hw, 1u), 60u);
  }();
  return n;
}

/** V records of S.step bytes to `d_cloud`, F triangles to `d_tri`, through the pinned ring; the copies are enqueued on
 *  c->stream (not waited for).  Throws when a source reports an error. */
void stageToDevice(nb2_context* c, const StageSource& S, uint64_t V, uint64_t F, void* d_cloud, void* d_tri)
{
  const unsigned kSlots = slotCount();
  const uint64_t pts_per = std::max<uint64_t>(kGrain, (slotBytes() / S.step) / kGrain * kGrain);
  const uint64_t tri_per = std::max<uint64_t>(kGrain, (slotBytes() / 12) / kGrain * kGrain);
  const size_t slot_bytes = std::max<size_t>(pts_per * S.step, tri_per * 12);
  c->stageRing.ensure(slot_bytes * kSlots);
  if (c->stageEvents.size() < kSlots)
  {
    for (auto& e : c->stageEvents)
      cudaEventDestroy(e);
    c->stageEvents.assign(kSlots, nullptr);
    for (auto& e : c->stageEvents)
      NB2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  const uint64_t n_cloud = (V + pts_per - 1) / pts_per, n_tri = (F + tri_per - 1) / tri_per, n_items = n_cloud + n_tri;
  std::vector<std::atomic<int>> enqueued(n_items);
  for (auto& a : enqueued)
    a.store(0, std::memory_order_relaxed);
  std::atomic<uint64_t> next{ 0 };
  std::atomic<int> error{ 0 };
  std::atomic<int> cuda_error{ 0 };
  uint8_t* const ring = c->stageRing.as<uint8_t>();
  auto work = [&]() {
    if (cudaSetDevice(c->device) != cudaSuccess)
    {
      cuda_error.store(1);
      return;
    }
    for (;;)
    {
      const uint64_t i = next.fetch_add(1);
      if (i >= n_items)
        return;
      const unsigned slot = static_cast<unsigned>(i % kSlots);
      bool bad = error.load() || cuda_error.load();
      if (!bad && i >= kSlots)
      {
        // the slot's previous tenant must have left for the device
        while (!enqueued[i - kSlots].load(std::memory_order_acquire))
          std::this_thread::yield();
        if (enqueued[i - kSlots].load() == 1 && cudaEventSynchronize(c->stageEvents[slot]) != cudaSuccess)
          cuda_error.store(1), bad = true;
      }
      int state = 2;  // 2 = nothing enqueued for this item
      if (!bad)
      {
        uint8_t* dst = ring + slot * slot_bytes;
        int rc;
        void* d;
        size_t bytes;
        if (i < n_cloud)
        {
          const uint64_t first = i * pts_per, count = std::min(pts_per, V - first);
          rc = S.cloud(first, count, dst);
          d = static_cast<uint8_t*>(d_cloud) + first * S.step;
          bytes = count * S.step;
        }
        else
        {
          const uint64_t first = (i - n_cloud) * tri_per, count = std::min(tri_per, F - first);
          rc = S.tri(first, count, reinterpret_cast<uint32_t*>(dst));
          d = static_cast<uint8_t*>(d_tri) + first * 12;
          bytes = count * 12;
        }
        if (rc != 0)
        {
          int expected = 0;
          error.compare_exchange_strong(expected, rc);
        }
        else if (cudaMemcpyAsync(d, dst, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
                 cudaEventRecord(c->stageEvents[slot], c->stream) != cudaSuccess)
          cuda_error.store(1);
        else
          state = 1;
      }
      enqueued[i].store(state, std::memory_order_release);
    }
  };
  const unsigned workers = static_cast<unsigned>(std::min<uint64_t>(std::min(stagingWorkers(), kSlots > 4 ? kSlots - 2 : 2u), n_items));
  std::vector<std::thread> pool;
  for (unsigned w = 1; w < workers; ++w)
    pool.emplace_back(work);
  work();
  for (auto& t : pool)
    t.join();
  if (cuda_error.load())
  {
    cudaStreamSynchronize(c->stream);
    fail(NB2_ERR_CUDA, std::string("staged upload: ") + cudaGetErrorString(cudaGetLastError()));
  }
  if (error.load())
  {
    cudaStreamSynchronize(c->stream);
    throw SourceError{ error.load() };
  }
}

bool pinnedHost(const void* p)
{
  cudaPointerAttributes a{};
  if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess)
  {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}
}  // namespace

static void readUploadVerdicts(nb2_context* c);

static void uploadImpl(nb2_context* c, const nb2_mesh_view* m, bool sharded)
{
  if (!c || !m)
    fail(NB2_ERR_INVALID_ARGUMENT, "context or mesh is NULL");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  if (m->n_points == 0)
    fail(NB2_ERR_TOO_FEW_POINTS, "mesh has no vertices");
  if (!m->cloud_data || (m->n_triangles && !m->triangles))
    fail(NB2_ERR_INVALID_ARGUMENT, "mesh data pointer is NULL");
  if (m->point_step % 4 || m->offset_xyz % 4 || m->offset_xyz + 12 > m->point_step)
    fail(NB2_ERR_INVALID_ARGUMENT, "XYZ fields are not contiguous floats");  // utils.cpp:82-83
  if (m->offset_normal >= 0 && (m->offset_normal % 4 || static_cast<uint32_t>(m->offset_normal) + 12 > m->point_step))
    fail(NB2_ERR_INVALID_ARGUMENT, "Normal fields are not contiguous floats");  // utils.cpp:104-105
  if (m->n_points >= (1ull << 31) || m->n_triangles >= (1ull << 31))
    fail(NB2_ERR_INVALID_ARGUMENT, "meshes with 2^31 or more vertices / triangles are not supported");

  c->V = 0;
  c->pca_valid = false;
  c->extents_done = false;
  const uint64_t V = m->n_points, F = m->n_triangles;
  c->timer_upload.begin(c->stream);
  // Inputs that already live in this GPU's memory are used in place (no copy; the caller keeps them alive until the
  // next upload).  Anything else is copied host -> device.
  auto onThisDevice = [&](const void* p) {
    cudaPointerAttributes a{};
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
      cudaGetLastError();
      return false;
    }
    return a.type == cudaMemoryTypeDevice && a.device == c->device;
  };
  c->pos.ensure(sizeof(float4) * V);
  const bool cloud_on_device = onThisDevice(m->cloud_data), tri_on_device = F && onThisDevice(m->triangles);
  // layout of the records as they will lie in HBM
  uint32_t step = m->point_step, off_xyz = m->offset_xyz;
  int32_t off_nrm = m->offset_normal;
  // (a pinned cloud goes out as it is: the copy engine moves it at the PCIe rate without the host cores, which a batch of
  //  meshes on several contexts needs for itself — cfg5: 96 ms against 155 ms with compaction; pageable memory has to pass
  //  through the cores anyway, and they compact it on the way)
  const bool compact = (m->flags & NB2_MESH_GEOMETRY_ONLY) && !cloud_on_device && !sharded &&
                       m->point_step > (m->offset_normal >= 0 ? 24u : 12u) && !pinnedHost(m->cloud_data);
  if (compact)
  {
    step = m->offset_normal >= 0 ? 24u : 12u;
    off_xyz = 0;
    off_nrm = m->offset_normal >= 0 ? 12 : -1;
  }
  // Host sources worth the staging threads: pageable memory (any size beyond a few slots) and clouds to be compacted;
  // pinned memory otherwise goes out in one piece per array, straight from where it lies
  const size_t cloud_bytes = V * m->point_step, tri_bytes = sizeof(uint32_t) * 3 * F;
  const bool stage_cloud = !cloud_on_device && !sharded && (compact || (cloud_bytes >= 4 * kSlotBytes && !pinnedHost(m->cloud_data)));
  const bool stage_tri = F && !tri_on_device && !sharded && tri_bytes >= 4 * kSlotBytes && !pinnedHost(m->triangles);
  if (cloud_on_device)
    c->blob.adopt(const_cast<void*>(m->cloud_data));
  else if (sharded)
  {
    // every rank copies 1 / world of the bytes over its own PCIe link; NVLink does the rest (comm.cpp)
    c->blob.ensure(shardedCapacity(c, cloud_bytes));
    shardedHostToDevice(c, c->blob.ptr, m->cloud_data, cloud_bytes);
  }
  else
  {
    c->blob.ensure(V * step);
    if (!stage_cloud)
      NB2_CUDA(cudaMemcpyAsync(c->blob.ptr, m->cloud_data, cloud_bytes, cudaMemcpyHostToDevice, c->stream));
  }
  bool tri_adopted = false;
  if (tri_on_device)
  {
    c->tri.adopt(const_cast<uint32_t*>(m->triangles));
    tri_adopted = true;
  }
  else if (sharded && F)
  {
    c->tri.ensure(shardedCapacity(c, tri_bytes));
    shardedHostToDevice(c, c->tri.ptr, m->triangles, tri_bytes);
  }
  else
  {
    c->tri.ensure(sizeof(uint32_t) * 3 * std::max<uint64_t>(F, 1));
    if (F && !stage_tri)
      NB2_CUDA(cudaMemcpyAsync(c->tri.ptr, m->triangles, tri_bytes, cudaMemcpyHostToDevice, c->stream));
  }
  if (stage_cloud || stage_tri)
  {
    StageSource S;
    S.step = step;
    const uint8_t* src = static_cast<const uint8_t*>(m->cloud_data);
    const uint32_t in_step = m->point_step, in_xyz = m->offset_xyz;
    const int32_t in_nrm = m->offset_normal;
    if (compact)
      S.cloud = [=](uint64_t first, uint64_t count, void* dst) {
        uint8_t* o = static_cast<uint8_t*>(dst);
        const uint8_t* p = src + first * in_step;
        if (in_nrm >= 0)
          for (uint64_t i = 0; i < count; ++i, p += in_step, o += 24)
          {
            std::memcpy(o, p + in_xyz, 12);
            std::memcpy(o + 12, p + in_nrm, 12);
          }
        else
          for (uint64_t i = 0; i < count; ++i, p += in_step, o += 12)
            std::memcpy(o, p + in_xyz, 12);
        return 0;
      };
    else
      S.cloud = [=](uint64_t first, uint64_t count, void* dst) {
        std::memcpy(dst, src + first * in_step, count * in_step);
        return 0;
      };
    const uint32_t* tsrc = m->triangles;
    S.tri = [=](uint64_t first, uint64_t count, uint32_t* dst) {
      std::memcpy(dst, tsrc + 3 * first, 12 * count);
      return 0;
    };
    stageToDevice(c, S, stage_cloud ? V : 0, stage_tri ? F : 0, c->blob.ptr, c->tri.ptr);
  }
  c->timer_upload.mark("h2d", c->stream);
  c->V = V;
  c->F = F;
  c->has_normals = off_nrm >= 0;
  // normals are read in place, out of the cloud (SoA copies exist only for normals computed or set afterwards)
  c->nrm_base = c->has_normals ? c->blob.as<uint8_t>() + off_nrm : nullptr;
  c->nrm_stride = step;
  launchIngestMoments(c, step, off_xyz, off_nrm);
  c->timer_upload.mark("ingest+moments", c->stream);
  // Triangles copied from the host are range-checked now (noise next to the PCIe copy).  Triangles adopted in place
  // from device memory are checked by the first kernel that streams them anyway (k_count_hits / the normals pass).
  c->tri_checked = false;
  if (!tri_adopted)
  {
    NB2_CUDA(cudaMemsetAsync(c->counters.as<unsigned>() + 12, 0, sizeof(unsigned), c->stream));
    launchValidateTriangles(c);
    c->timer_upload.mark("validate", c->stream);
  }
  // (mean, covariance and eigen axes were finalised on the device by the last block of the ingest pass)
  c->extents_done = false;
  c->pca_valid = false;
  std::memset(&c->pca, 0, sizeof(c->pca));
  if (tri_adopted || cloud_on_device)
  {
    // device-resident input: nothing is read back here; non-finite coordinates and out-of-range indices are reported
    // by the next call that synchronises (nb2_plan, nb2_mesh_pca, nb2_normals_from_mesh_faces, ...)
    c->checks_deferred = true;
    c->tri_checked = false;
    return;
  }
  readUploadVerdicts(c);
}

/** end of a host upload: wait for the ingest pass and the index check, report what they found */
static void readUploadVerdicts(nb2_context* c)
{
  const uint64_t V = c->V, F = c->F;
  c->checks_deferred = false;
  c->hostSmall.ensure(4096 + sizeof(FrameDev));
  auto* h = c->hostSmall.as<uint8_t>();
  const int grid = c->ingest_grid;  // (the total lies behind the block partials of the last ingest pass)
  NB2_CUDA(cudaMemcpyAsync(h, c->partials.as<MomentPartial>() + grid, sizeof(MomentPartial), cudaMemcpyDeviceToHost, c->stream));
  NB2_CUDA(cudaMemcpyAsync(h + 512, c->counters.as<unsigned>() + 12, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  NB2_CUDA(cudaStreamSynchronize(c->stream));
  MomentPartial total;
  std::memcpy(&total, h, sizeof(total));
  unsigned max_index;
  std::memcpy(&max_index, h + 512, 4);
  if (total.non_finite)
  {
    c->V = 0;
    fail(NB2_ERR_NON_FINITE, std::to_string(total.non_finite) + " vertices have non-finite coordinates");
  }
  if (F && max_index >= V)
  {
    c->V = 0;
    fail(NB2_ERR_INVALID_ARGUMENT, "triangle references vertex " + std::to_string(max_index) + " but the cloud has " +
                                       std::to_string(V) + " points");
  }
  c->tri_checked = true;
}

/** nb2_mesh_upload_stream: the caller's callbacks fill the pinned slots; the staged mesh lands in the context's spare
 *  pair of buffers and replaces the resident one only if `accept` says so */
static void uploadStreamImpl(nb2_context* c, const nb2_mesh_stream* m, int* committed)
{
  if (!c || !m)
    fail(NB2_ERR_INVALID_ARGUMENT, "context or mesh is NULL");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  if (committed)
    *committed = 0;
  if (m->n_points == 0)
    fail(NB2_ERR_TOO_FEW_POINTS, "mesh has no vertices");
  if (!m->cloud || (m->n_triangles && !m->triangles))
    fail(NB2_ERR_INVALID_ARGUMENT, "mesh source callback is NULL");
  if (m->point_step == 0 || m->point_step % 4 || m->offset_xyz % 4 || m->offset_xyz + 12 > m->point_step)
    fail(NB2_ERR_INVALID_ARGUMENT, "XYZ fields are not contiguous floats");
  if (m->offset_normal >= 0 && (m->offset_normal % 4 || static_cast<uint32_t>(m->offset_normal) + 12 > m->point_step))
    fail(NB2_ERR_INVALID_ARGUMENT, "Normal fields are not contiguous floats");
  if (m->n_points >= (1ull << 31) || m->n_triangles >= (1ull << 31))
    fail(NB2_ERR_INVALID_ARGUMENT, "meshes with 2^31 or more vertices / triangles are not supported");
  const uint64_t V = m->n_points, F = m->n_triangles;
  c->timer_upload.begin(c->stream);
  c->blobSpare.ensure(V * m->point_step);
  c->triSpare.ensure(sizeof(uint32_t) * 3 * std::max<uint64_t>(F, 1));
  StageSource S;
  S.step = m->point_step;
  void* const user = m->user;
  const auto cloud_fn = m->cloud;
  const auto tri_fn = m->triangles;
  S.cloud = [=](uint64_t first, uint64_t count, void* dst) { return cloud_fn(user, first, count, dst); };
  S.tri = [=](uint64_t first, uint64_t count, uint32_t* dst) { return tri_fn(user, first, count, dst); };
  try
  {
    stageToDevice(c, S, V, F, c->blobSpare.ptr, c->triSpare.ptr);
  }
  catch (const SourceError& e)
  {
    fail(NB2_ERR_INVALID_ARGUMENT, "mesh source callback reported error " + std::to_string(e.code));
  }
  if (m->accept && !m->accept(user))
  {
    // same content as the resident mesh: the resident mesh and everything derived from it stay
    c->timer_upload.mark("h2d", c->stream);
    return;
  }
  std::swap(c->blob, c->blobSpare);
  std::swap(c->tri, c->triSpare);
  c->V = 0;
  c->pos.ensure(sizeof(float4) * V);
  c->timer_upload.mark("h2d", c->stream);
  c->V = V;
  c->F = F;
  c->has_normals = m->offset_normal >= 0;
  c->nrm_base = c->has_normals ? c->blob.as<uint8_t>() + m->offset_normal : nullptr;
  c->nrm_stride = m->point_step;
  launchIngestMoments(c, m->point_step, m->offset_xyz, m->offset_normal);
  c->timer_upload.mark("ingest+moments", c->stream);
  c->tri_checked = false;
  NB2_CUDA(cudaMemsetAsync(c->counters.as<unsigned>() + 12, 0, sizeof(unsigned), c->stream));
  launchValidateTriangles(c);
  c->timer_upload.mark("validate", c->stream);
  c->extents_done = false;
  c->pca_valid = false;
  std::memset(&c->pca, 0, sizeof(c->pca));
  readUploadVerdicts(c);
  if (committed)
    *committed = 1;
}

/** nb2_mesh_upload_ply: header on the host, records on the device (ply.cpp, kernels_ply.cu) */
static void uploadPlyImpl(nb2_context* c, const void* file_image, uint64_t n_bytes, nb2_mesh_file_info* info)
{
  if (!c || !file_image)
    fail(NB2_ERR_INVALID_ARGUMENT, "context or file image is NULL");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  const PlyLayout L = parsePlyHeader(static_cast<const uint8_t*>(file_image), n_bytes);
  if (L.n_points == 0)
    fail(NB2_ERR_TOO_FEW_POINTS, "mesh has no vertices");
  c->V = 0;
  c->pca_valid = false;
  c->extents_done = false;
  c->timer_upload.begin(c->stream);
  // one PCIe copy of the records (the header stays behind; byte offsets keep referring to the file)
  const uint64_t first = L.vertex_begin & ~uint64_t(255);
  const uint64_t last = std::max(L.vertex_begin + L.n_points * L.point_step, L.face_begin + L.n_faces * L.face_step);
  c->fileImg.ensure(((last - first + 15) & ~uint64_t(7)) + 8);
  NB2_CUDA(cudaMemcpyAsync(c->fileImg.ptr, static_cast<const uint8_t*>(file_image) + first, last - first,
                           cudaMemcpyHostToDevice, c->stream));
  c->timer_upload.mark("h2d", c->stream);
  PlyLayout D = L;
  D.vertex_begin -= first;
  D.face_begin -= first;
  c->pos.ensure(sizeof(float4) * L.n_points);
  launchPlyDecode(c, D);
  c->timer_upload.mark("decode", c->stream);
  const uint32_t step = L.has_normals ? 32u : 16u;
  c->V = L.n_points;
  c->F = L.n_faces;
  c->has_normals = L.has_normals;
  c->nrm_base = c->has_normals ? c->blob.as<uint8_t>() + 16 : nullptr;
  c->nrm_stride = step;
  launchIngestMoments(c, step, 0, c->has_normals ? 16 : -1);
  c->timer_upload.mark("ingest+moments", c->stream);
  std::memset(&c->pca, 0, sizeof(c->pca));
  c->checks_deferred = false;
  c->tri_checked = false;
  c->hostSmall.ensure(4096 + sizeof(FrameDev));
  auto* h = c->hostSmall.as<uint8_t>();
  const int grid = c->ingest_grid;  // (the total lies behind the block partials of the last ingest pass)
  NB2_CUDA(cudaMemcpyAsync(h, c->partials.as<MomentPartial>() + grid, sizeof(MomentPartial), cudaMemcpyDeviceToHost, c->stream));
  NB2_CUDA(cudaMemcpyAsync(h + 512, c->counters.as<unsigned>() + 12, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  NB2_CUDA(cudaMemcpyAsync(h + 516, c->counters.as<unsigned>() + 16, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  NB2_CUDA(cudaStreamSynchronize(c->stream));
  MomentPartial total;
  std::memcpy(&total, h, sizeof(total));
  unsigned max_index, not_triangles;
  std::memcpy(&max_index, h + 512, 4);
  std::memcpy(&not_triangles, h + 516, 4);
  const uint64_t V = c->V;
  if (not_triangles)
  {
    c->V = 0;
    fail(NB2_ERR_NOT_TRIANGLES, "PLY: the file holds faces that are not triangles (the plane slicer path takes triangle meshes)");
  }
  if (total.non_finite)
  {
    c->V = 0;
    fail(NB2_ERR_NON_FINITE, std::to_string(total.non_finite) + " vertices have non-finite coordinates");
  }
  if (c->F && max_index >= V)
  {
    c->V = 0;
    fail(NB2_ERR_INVALID_ARGUMENT, "triangle references vertex " + std::to_string(max_index) + " but the cloud has " +
                                       std::to_string(V) + " points");
  }
  c->tri_checked = true;
  if (info)
  {
    info->n_points = c->V;
    info->n_triangles = c->F;
    info->has_normals = c->has_normals ? 1 : 0;
    info->reserved_ = 0;
  }
}

/** Facets a binary STL image holds, as vtkSTLReader::ReadBinarySTL counts them: the 4-byte count of the header is not
 *  trusted, facets are read while 48 more bytes are there (the trailing 2-byte attribute may be cut off).  Text files
 *  ("solid ..." whose size does not match the binary layout) are refused. */
static uint64_t stlFacetCount(const uint8_t* data, uint64_t n_bytes)
{
  if (n_bytes < 84)
    fail(NB2_ERR_INVALID_ARGUMENT, "STL: the file is shorter than a binary STL header (84 bytes)");
  uint32_t declared;
  std::memcpy(&declared, data + 80, 4);
  const bool exact = n_bytes == 84 + 50ull * declared;
  if (!exact && std::memcmp(data, "solid", 5) == 0)
    fail(NB2_ERR_INVALID_ARGUMENT, "STL: ASCII files are not decoded on the device (binary STL only)");
  const uint64_t body = n_bytes - 84;
  const uint64_t facets = body / 50 + (body % 50 >= 48 ? 1 : 0);
  if (facets > 1400000000ull)
    fail(NB2_ERR_INVALID_ARGUMENT, "STL: more than 1.4e9 facets");
  return facets;
}

/** nb2_mesh_upload_stl: one copy of the file image, welded on the device (kernels_stl.cu) */
static void uploadStlImpl(nb2_context* c, const void* file_image, uint64_t n_bytes, nb2_mesh_file_info* info)
{
  if (!c || !file_image)
    fail(NB2_ERR_INVALID_ARGUMENT, "context or file image is NULL");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  const uint64_t facets = stlFacetCount(static_cast<const uint8_t*>(file_image), n_bytes);
  if (facets == 0)
    fail(NB2_ERR_TOO_FEW_POINTS, "mesh has no vertices");
  c->V = 0;
  c->pca_valid = false;
  c->extents_done = false;
  c->timer_upload.begin(c->stream);
  const uint64_t used = std::min<uint64_t>(n_bytes, 84 + 50 * facets);
  c->fileImg.ensure(((used + 15) & ~uint64_t(7)) + 8);
  NB2_CUDA(cudaMemcpyAsync(c->fileImg.ptr, file_image, used, cudaMemcpyHostToDevice, c->stream));
  c->timer_upload.mark("h2d", c->stream);
  uint64_t V = 0, F = 0;
  launchStlDecode(c, facets, &V, &F);
  c->timer_upload.mark("weld", c->stream);
  c->pos.ensure(sizeof(float4) * std::max<uint64_t>(V, 1));
  c->V = V;
  c->F = F;
  c->has_normals = false;
  c->nrm_base = nullptr;
  c->nrm_stride = 16;
  launchIngestMoments(c, 16, 0, -1);
  c->timer_upload.mark("ingest+moments", c->stream);
  std::memset(&c->pca, 0, sizeof(c->pca));
  c->checks_deferred = false;
  c->hostSmall.ensure(4096 + sizeof(FrameDev));
  auto* h = c->hostSmall.as<uint8_t>();
  const int grid = c->ingest_grid;  // (the total lies behind the block partials of the last ingest pass)
  NB2_CUDA(cudaMemcpyAsync(h, c->partials.as<MomentPartial>() + grid, sizeof(MomentPartial), cudaMemcpyDeviceToHost, c->stream));
  NB2_CUDA(cudaStreamSynchronize(c->stream));
  MomentPartial total;
  std::memcpy(&total, h, sizeof(total));
  if (total.non_finite)
  {
    c->V = 0;
    fail(NB2_ERR_NON_FINITE, std::to_string(total.non_finite) + " vertices have non-finite coordinates");
  }
  c->tri_checked = true;  // the ids were assigned here: all < V
  if (info)
  {
    info->n_points = c->V;
    info->n_triangles = c->F;
    info->has_normals = 0;
    info->reserved_ = 0;
  }
}

int nb2_ply_inspect(const void* file_image, uint64_t n_bytes, nb2_mesh_file_info* info)
{
  NB2_API_BEGIN
  if (!file_image || !info)
    fail(NB2_ERR_INVALID_ARGUMENT, "file image or info is NULL");
  const PlyLayout L = parsePlyHeader(static_cast<const uint8_t*>(file_image), n_bytes);
  info->n_points = L.n_points;
  info->n_triangles = L.n_faces;
  info->has_normals = L.has_normals ? 1 : 0;
  info->reserved_ = 0;
  NB2_API_END
}

int nb2_mesh_upload_ply(nb2_context* c, const void* file_image, uint64_t n_bytes, nb2_mesh_file_info* info)
{
  NB2_API_BEGIN
  uploadPlyImpl(c, file_image, n_bytes, info);
  NB2_API_END
}

int nb2_mesh_upload_ply_file(nb2_context* c, const char* path, nb2_mesh_file_info* info)
{
  NB2_API_BEGIN
  if (!path)
    fail(NB2_ERR_INVALID_ARGUMENT, "path is NULL");
  // the file is mapped, not read: the PCIe copy pulls the pages straight out of the page cache
  const int fd = ::open(path, O_RDONLY);
  if (fd < 0)
    fail(NB2_ERR_INVALID_ARGUMENT, std::string("cannot open '") + path + "'");
  struct stat st{};
  if (::fstat(fd, &st) != 0 || st.st_size <= 0)
  {
    ::close(fd);
    fail(NB2_ERR_INVALID_ARGUMENT, std::string("cannot stat '") + path + "' (or the file is empty)");
  }
  void* map = ::mmap(nullptr, static_cast<size_t>(st.st_size), PROT_READ, MAP_PRIVATE, fd, 0);
  ::close(fd);
  if (map == MAP_FAILED)
    fail(NB2_ERR_INVALID_ARGUMENT, std::string("cannot map '") + path + "'");
  struct Unmap
  {
    void* p;
    size_t n;
    ~Unmap() { ::munmap(p, n); }
  } unmap{ map, static_cast<size_t>(st.st_size) };
  uploadPlyImpl(c, map, static_cast<uint64_t>(st.st_size), info);
  NB2_API_END
}

int nb2_stl_inspect(const void* file_image, uint64_t n_bytes, nb2_mesh_file_info* info)
{
  NB2_API_BEGIN
  if (!file_image || !info)
    fail(NB2_ERR_INVALID_ARGUMENT, "file image or info is NULL");
  const uint64_t facets = stlFacetCount(static_cast<const uint8_t*>(file_image), n_bytes);
  info->n_points = 3 * facets;  // before welding: upper bounds
  info->n_triangles = facets;
  info->has_normals = 0;
  info->reserved_ = 0;
  NB2_API_END
}

int nb2_mesh_upload_stl(nb2_context* c, const void* file_image, uint64_t n_bytes, nb2_mesh_file_info* info)
{
  NB2_API_BEGIN
  uploadStlImpl(c, file_image, n_bytes, info);
  NB2_API_END
}

int nb2_mesh_upload_file(nb2_context* c, const char* path, nb2_mesh_file_info* info)
{
  NB2_API_BEGIN
  if (!path)
    fail(NB2_ERR_INVALID_ARGUMENT, "path is NULL");
  // pcl::io::loadPolygonFile dispatches on the extension (vtk_lib_io.cpp)
  const std::string p(path);
  const size_t dot = p.find_last_of('.');
  std::string ext = dot == std::string::npos ? std::string() : p.substr(dot);
  for (char& ch : ext)
    ch = static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
  const bool ply = ext == ".ply", stl = ext == ".stl";
  if (!ply && !stl)
    fail(NB2_ERR_INVALID_ARGUMENT, "unsupported mesh file extension '" + ext + "' (binary .ply and .stl are decoded on the device)");
  const int fd = ::open(path, O_RDONLY);
  if (fd < 0)
    fail(NB2_ERR_INVALID_ARGUMENT, std::string("cannot open '") + path + "'");
  struct stat st{};
  if (::fstat(fd, &st) != 0 || st.st_size <= 0)
  {
    ::close(fd);
    fail(NB2_ERR_INVALID_ARGUMENT, std::string("cannot stat '") + path + "' (or the file is empty)");
  }
  void* map = ::mmap(nullptr, static_cast<size_t>(st.st_size), PROT_READ, MAP_PRIVATE, fd, 0);
  ::close(fd);
  if (map == MAP_FAILED)
    fail(NB2_ERR_INVALID_ARGUMENT, std::string("cannot map '") + path + "'");
  struct Unmap
  {
    void* p;
    size_t n;
    ~Unmap() { ::munmap(p, n); }
  } unmap{ map, static_cast<size_t>(st.st_size) };
  if (ply)
    uploadPlyImpl(c, map, static_cast<uint64_t>(st.st_size), info);
  else
    uploadStlImpl(c, map, static_cast<uint64_t>(st.st_size), info);
  NB2_API_END
}

int nb2_mesh_download_triangles(nb2_context* c, uint32_t* triangles, uint64_t capacity_triangles)
{
  NB2_API_BEGIN
  if (!c || !triangles)
    fail(NB2_ERR_INVALID_ARGUMENT, "context or triangles is NULL");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  requireMesh(c);
  if (capacity_triangles < c->F)
    fail(NB2_ERR_INVALID_ARGUMENT, "nb2_mesh_download_triangles: the buffer holds " + std::to_string(capacity_triangles) +
                                       " triangles but the resident mesh has " + std::to_string(c->F));
  if (c->F)
    NB2_CUDA(cudaMemcpyAsync(triangles, c->tri.ptr, sizeof(uint32_t) * 3 * c->F, cudaMemcpyDeviceToHost, c->stream));
  NB2_CUDA(cudaStreamSynchronize(c->stream));
  NB2_API_END
}

int nb2_mesh_download_cloud(nb2_context* c, void* cloud_data, uint64_t capacity_bytes)
{
  NB2_API_BEGIN
  if (!c || !cloud_data)
    fail(NB2_ERR_INVALID_ARGUMENT, "context or cloud_data is NULL");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  requireMesh(c);
  if (capacity_bytes < c->V * c->rec_step)
    fail(NB2_ERR_INVALID_ARGUMENT, "nb2_mesh_download_cloud: the buffer holds " + std::to_string(capacity_bytes) +
                                       " bytes but the resident cloud has " + std::to_string(c->V) + " records of " +
                                       std::to_string(c->rec_step) + " bytes");
  NB2_CUDA(cudaMemcpyAsync(cloud_data, c->blob.ptr, c->V * c->rec_step, cudaMemcpyDeviceToHost, c->stream));
  NB2_CUDA(cudaStreamSynchronize(c->stream));
  NB2_API_END
}

int nb2_mesh_record_layout(const nb2_context* c, uint32_t* point_step, uint32_t* offset_xyz, int32_t* offset_normal)
{
  NB2_API_BEGIN
  if (!c)
    fail(NB2_ERR_INVALID_ARGUMENT, "context is NULL");
  requireMesh(c);
  if (point_step)
    *point_step = c->rec_step;
  if (offset_xyz)
    *offset_xyz = c->rec_off_xyz;
  if (offset_normal)
    *offset_normal = c->rec_off_nrm;
  NB2_API_END
}

/** the subdivided mesh (in subBlob[slot] / subTri[slot]) becomes the resident one: same ingest as an upload of
 *  device-resident arrays (adopted in place, moments and PCA axes recomputed) */
static void adoptSubdivided(nb2_context* c, int slot, uint64_t V, uint64_t F, const sub::Layout& L, nb2_mesh_file_info* info)
{
  nb2_mesh_view v{};
  v.cloud_data = c->subBlob[slot].ptr;
  v.n_points = V;
  v.point_step = L.step;
  v.offset_xyz = L.off_xyz;
  v.offset_normal = L.off_nrm;
  v.triangles = c->subTri[slot].as<uint32_t>();
  v.n_triangles = F;
  uploadImpl(c, &v, false);
  if (info)
  {
    info->n_points = V;
    info->n_triangles = F;
    info->has_normals = L.off_nrm >= 0;
    info->reserved_ = 0;
  }
}

/** which of the two output slots the resident mesh does NOT live in; the layout of the resident cloud */
static int subdivisionSetup(nb2_context* c, int32_t offset_rgba, sub::Layout* L)
{
  requireMesh(c);
  if (c->rec_step == 0 || c->rec_step % 4)
    fail(NB2_ERR_INVALID_ARGUMENT, "the resident cloud has no record layout");
  if (offset_rgba >= 0 && static_cast<uint32_t>(offset_rgba) + 4 > c->rec_step)
    fail(NB2_ERR_INVALID_ARGUMENT, "rgba field lies outside the point record");
  *L = sub::Layout{ c->rec_step, c->rec_off_xyz, c->rec_off_nrm, offset_rgba };
  if (c->checks_deferred || !c->tri_checked)
  {
    // mesh adopted from device memory (e.g. the output of a previous subdivision): make sure its indices are in range
    // before they are used as addresses
    NB2_CUDA(cudaMemsetAsync(c->counters.as<unsigned>() + 12, 0, sizeof(unsigned), c->stream));
    launchValidateTriangles(c);
    c->hostSmall.ensure(4096);
    unsigned* h = c->hostSmall.as<unsigned>();
    NB2_CUDA(cudaMemcpyAsync(h, c->counters.as<unsigned>() + 12, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    NB2_CUDA(cudaStreamSynchronize(c->stream));
    if (c->F && h[0] >= c->V)
      fail(NB2_ERR_INVALID_ARGUMENT, "triangle references vertex " + std::to_string(h[0]) + " but the cloud has " +
                                         std::to_string(c->V) + " points");
    c->tri_checked = true;
  }
  return (c->blob.ptr == c->subBlob[0].ptr || c->tri.ptr == c->subTri[0].ptr) ? 1 : 0;
}

int nb2_face_midpoint_subdivision(nb2_context* c, uint32_t n_iterations, int32_t offset_rgba, nb2_mesh_file_info* info)
{
  NB2_API_BEGIN
  if (!c)
    fail(NB2_ERR_INVALID_ARGUMENT, "context is NULL");
  if (n_iterations < 1 || n_iterations > static_cast<uint32_t>(sub::kMaxIterations))
    fail(NB2_ERR_INVALID_ARGUMENT, "FaceMidpointSubdivision: n_iterations must be between 1 and " +
                                       std::to_string(sub::kMaxIterations) + " (0 never terminates in the reference)");
  sub::Layout L{};
  int slot = 0;
  uint64_t V = 0, F = 0;
  {
    std::lock_guard<std::mutex> lock(c->mu);
    DeviceGuard g(c->device);
    slot = subdivisionSetup(c, offset_rgba, &L);
    launchMidpointSubdivision(c, L, static_cast<int>(n_iterations), slot, &V, &F);
  }
  adoptSubdivided(c, slot, V, F, L, info);
  NB2_API_END
}

int nb2_face_subdivision_by_area(nb2_context* c, float max_area, int32_t offset_rgba, uint64_t max_points, nb2_mesh_file_info* info)
{
  NB2_API_BEGIN
  if (!c)
    fail(NB2_ERR_INVALID_ARGUMENT, "context is NULL");
  if (!(max_area == max_area))
    fail(NB2_ERR_INVALID_ARGUMENT, "FaceSubdivisionByArea: max_area is NaN");
  sub::Layout L{};
  int slot = 0;
  uint64_t V = 0, F = 0;
  {
    std::lock_guard<std::mutex> lock(c->mu);
    DeviceGuard g(c->device);
    slot = subdivisionSetup(c, offset_rgba, &L);
    launchAreaSubdivision(c, L, max_area, max_points, slot, &V, &F);
  }
  adoptSubdivided(c, slot, V, F, L, info);
  NB2_API_END
}

int nb2_cluster_order(const uint32_t* cluster_sizes, uint64_t n_clusters, uint32_t* order)
{
  NB2_API_BEGIN
  if (n_clusters && (!cluster_sizes || !order))
    fail(NB2_ERR_INVALID_ARGUMENT, "cluster_sizes or order is NULL");
  // pcl::EuclideanClusterExtraction::extract: std::sort(clusters.rbegin(), clusters.rend(), comparePointClusters) with
  // comparePointClusters(a, b) = a.indices.size() < b.indices.size().  The order it leaves clusters of equal size in is
  // whatever the library's introsort does; the very same call on (size, id) pairs takes the same decisions.
  struct Entry
  {
    uint32_t size, id;
  };
  std::vector<Entry> v(n_clusters);
  for (uint64_t i = 0; i < n_clusters; ++i)
    v[i] = { cluster_sizes[i], static_cast<uint32_t>(i) };
  std::sort(v.rbegin(), v.rend(), [](const Entry& a, const Entry& b) { return a.size < b.size; });
  for (uint64_t i = 0; i < n_clusters; ++i)
    order[i] = v[i].id;
  NB2_API_END
}

int nb2_euclidean_cluster_labels(nb2_context* c, double tolerance, uint32_t* labels, uint64_t capacity_labels)
{
  NB2_API_BEGIN
  if (!c || !labels)
    fail(NB2_ERR_INVALID_ARGUMENT, "context or labels is NULL");
  if (!(tolerance == tolerance))
    fail(NB2_ERR_INVALID_ARGUMENT, "EuclideanClustering: tolerance is NaN");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  requireMesh(c);
  const uint64_t V = c->V;
  if (capacity_labels < V)
    fail(NB2_ERR_INVALID_ARGUMENT, "nb2_euclidean_cluster_labels: the buffer holds " + std::to_string(capacity_labels) +
                                       " labels but the resident cloud has " + std::to_string(V) + " points");
  if (!(tolerance > 0.0))
  {
    // a radius search with radius <= 0 finds nothing, not even the query point: every point is its own cluster
    for (uint64_t v = 0; v < V; ++v)
      labels[v] = static_cast<uint32_t>(v);
    return NB2_OK;
  }
  ensureHostPca(c);  // the AABB of the cloud (and the deferred checks of a device-adopted mesh)
  clu::GridParams grid{};
  grid.h = tolerance * (1.0 + 1.0 / 1024.0);
  grid.r2 = static_cast<float>(tolerance * tolerance);
  for (int i = 0; i < 3; ++i)
  {
    grid.mn[i] = static_cast<double>(c->pca.aabb_min[i]) - grid.h;  // one spare cell below the cloud
    const double cells = (static_cast<double>(c->pca.aabb_max[i]) - grid.mn[i]) / grid.h;
    if (!(cells < clu::kMaxCells))
      fail(NB2_ERR_INVALID_ARGUMENT, "EuclideanClustering: tolerance " + std::to_string(tolerance) +
                                         " is less than 2^-40 of the extent of the cloud");
  }
  launchClusterLabels(c, grid, 4e11, labels);
  NB2_API_END
}

int nb2_normal_estimation_pcl(nb2_context* c, double radius, double vx, double vy, double vz, float* normals, float* curvature)
{
  NB2_API_BEGIN
  if (!c)
    fail(NB2_ERR_INVALID_ARGUMENT, "context is NULL");
  if (!(radius == radius))
    fail(NB2_ERR_INVALID_ARGUMENT, "NormalEstimationPCL: radius is NaN");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  requireMesh(c);
  const uint64_t V = c->V;
  ensureHostPca(c);  // the AABB of the cloud (and the deferred checks of a device-adopted mesh)
  clu::GridParams grid{};
  // a radius search with radius <= 0 finds nothing, not even the query point: every normal is NaN.  The grid is then built
  // with a nominal cell (nothing passes the distance test: r2 = 0 and the test is strict)
  const double r = radius > 0.0 ? radius : 0.0;
  grid.r2 = static_cast<float>(r * r);
  double extent = 0.0;
  for (int i = 0; i < 3; ++i)
    extent = std::max(extent, static_cast<double>(c->pca.aabb_max[i]) - static_cast<double>(c->pca.aabb_min[i]));
  grid.h = r > 0.0 ? r * (1.0 + 1.0 / 1024.0) : std::max(extent, 1.0);
  for (int i = 0; i < 3; ++i)
  {
    grid.mn[i] = static_cast<double>(c->pca.aabb_min[i]) - grid.h;  // one spare cell below the cloud
    const double cells = (static_cast<double>(c->pca.aabb_max[i]) - grid.mn[i]) / grid.h;
    if (!(cells < clu::kMaxCells))
      fail(NB2_ERR_INVALID_ARGUMENT, "NormalEstimationPCL: radius " + std::to_string(radius) + " is less than 2^-40 of the extent of the cloud");
  }
  const float vp[3] = { static_cast<float>(vx), static_cast<float>(vy), static_cast<float>(vz) };  // setViewPoint takes floats
  if (!(r > 0.0))
  {
    // (no search at all: V NaN normals; the resident mesh keeps the normals it had)
    const float nan = std::nanf("");
    if (normals)
      std::fill(normals, normals + 3 * V, nan);
    if (curvature)
      std::fill(curvature, curvature + V, nan);
    return NB2_OK;
  }
  launchNormalEstimationPcl(c, grid, 4e11, vp, normals, curvature);
  NB2_API_END
}

int nb2_mesh_upload(nb2_context* c, const nb2_mesh_view* m)
{
  NB2_API_BEGIN
  uploadImpl(c, m, false);
  NB2_API_END
}

int nb2_mesh_upload_stream(nb2_context* c, const nb2_mesh_stream* m, int* committed)
{
  NB2_API_BEGIN
  uploadStreamImpl(c, m, committed);
  NB2_API_END
}

int nb2_mesh_upload_sharded(nb2_context* c, const nb2_mesh_view* m)
{
  NB2_API_BEGIN
  if (!c || c->comm_world <= 1 || !c->nccl_comm)
    uploadImpl(c, m, false);  // no communicator: an ordinary upload
  else
    uploadImpl(c, m, true);
  NB2_API_END
}

int nb2_mesh_set_normals(nb2_context* c, const float* normals)
{
  NB2_API_BEGIN
  if (!c || !normals)
    fail(NB2_ERR_INVALID_ARGUMENT, "context or normals is NULL");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  requireMesh(c);
  c->faceNrm.ensure(sizeof(float) * 3 * c->V);  // staging
  NB2_CUDA(cudaMemcpyAsync(c->faceNrm.ptr, normals, sizeof(float) * 3 * c->V, cudaMemcpyHostToDevice, c->stream));
  launchSetNormals(c, c->faceNrm.as<float>());
  NB2_CUDA(cudaStreamSynchronize(c->stream));
  c->has_normals = true;
  NB2_API_END
}

int nb2_mesh_download(nb2_context* c, float* xyz, float* normals)
{
  NB2_API_BEGIN
  if (!c)
    fail(NB2_ERR_INVALID_ARGUMENT, "context is NULL");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  requireMesh(c);
  c->outPos.ensure(sizeof(float) * 3 * c->V);
  c->outNrm.ensure(sizeof(float) * 3 * c->V);
  launchPackOut(c, xyz ? c->outPos.as<float>() : nullptr, normals ? c->outNrm.as<float>() : nullptr);
  if (xyz)
    NB2_CUDA(cudaMemcpyAsync(xyz, c->outPos.ptr, sizeof(float) * 3 * c->V, cudaMemcpyDeviceToHost, c->stream));
  if (normals)
    NB2_CUDA(cudaMemcpyAsync(normals, c->outNrm.ptr, sizeof(float) * 3 * c->V, cudaMemcpyDeviceToHost, c->stream));
  NB2_CUDA(cudaStreamSynchronize(c->stream));
  NB2_API_END
}

int nb2_mesh_pca(nb2_context* c, nb2_pca* out)
{
  NB2_API_BEGIN
  if (!c || !out)
    fail(NB2_ERR_INVALID_ARGUMENT, "context or out is NULL");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  requireMesh(c);
  requirePcaPoints(c);
  ensureHostPca(c);
  *out = c->pca;
  NB2_API_END
}

int nb2_principal_axis_direction(nb2_context* c, double rotation_offset, double out[3])
{
  NB2_API_BEGIN
  if (!c || !out)
    fail(NB2_ERR_INVALID_ARGUMENT, "context or out is NULL");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  requireMesh(c);
  requirePcaPoints(c);
  ensureHostPca(c);
  principalAxisDirection(c->pca, rotation_offset, out);
  NB2_API_END
}

int nb2_pca_rotated_direction(nb2_context* c, double rotation_angle, const double nominal[3], double out[3])
{
  NB2_API_BEGIN
  if (!c || !nominal || !out)
    fail(NB2_ERR_INVALID_ARGUMENT, "NULL argument");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  requireMesh(c);
  requirePcaPoints(c);
  ensureHostPca(c);
  pcaRotatedDirection(c->pca, rotation_angle, nominal, out);
  NB2_API_END
}

int nb2_aabb_center_origin(nb2_context* c, double out[3])
{
  NB2_API_BEGIN
  if (!c || !out)
    fail(NB2_ERR_INVALID_ARGUMENT, "context or out is NULL");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  ensureHostPca(c);
  // AABBCenterOriginGenerator::generate (aabb_center_origin_generator.cpp:9-17): pcl::getMinMax3D, then
  // (max - min) in float, cast to double, halved — half the RANGE of the box, not its centre; restated as it is
  for (int i = 0; i < 3; ++i)
    out[i] = static_cast<double>(c->pca.aabb_max[i] - c->pca.aabb_min[i]) / 2.0;
  NB2_API_END
}

int nb2_centroid_origin(nb2_context* c, double out[3])
{
  NB2_API_BEGIN
  if (!c || !out)
    fail(NB2_ERR_INVALID_ARGUMENT, "context or out is NULL");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  ensureHostPca(c);
  for (int i = 0; i < 3; ++i)
    out[i] = static_cast<double>(c->pca.mean[i]);
  NB2_API_END
}

int nb2_normals_from_mesh_faces(nb2_context* c, float* normals_out)
{
  NB2_API_BEGIN
  if (!c)
    fail(NB2_ERR_INVALID_ARGUMENT, "context is NULL");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  requireMesh(c);
  c->timer.begin(c->stream);
  // (vertex indices not checked yet, as for triangles adopted from device memory, are checked by the kernels
  // themselves: offending faces are skipped and the largest index is read back with the final synchronisation)
  launchVertexNormals(c);
  c->hostSmall.ensure(4096);
  unsigned* h_max = c->hostSmall.as<unsigned>();
  NB2_CUDA(cudaMemcpyAsync(h_max, c->counters.as<unsigned>() + 12, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  if (normals_out)
  {
    c->outPos.ensure(sizeof(float) * 3 * c->V);
    launchPackOut(c, nullptr, c->outPos.as<float>());
    c->timer.mark("pack", c->stream);
    NB2_CUDA(cudaMemcpyAsync(normals_out, c->outPos.ptr, sizeof(float) * 3 * c->V, cudaMemcpyDeviceToHost, c->stream));
    c->timer.mark("d2h", c->stream);
  }
  NB2_CUDA(cudaStreamSynchronize(c->stream));
  if (c->F && h_max[0] >= c->V)
  {
    const uint64_t V = c->V;
    c->V = 0;  // the mesh is unusable
    fail(NB2_ERR_INVALID_ARGUMENT, "triangle references vertex " + std::to_string(h_max[0]) + " but the cloud has " +
                                       std::to_string(V) + " points");
  }
  c->tri_checked = true;
  c->has_normals = true;
  NB2_API_END
}

int nb2_eigen_solver_3f(const float cov[9], float evals_ascending[3], float evecs_ascending[9])
{
  NB2_API_BEGIN
  if (!cov || !evals_ascending || !evecs_ascending)
    fail(NB2_ERR_INVALID_ARGUMENT, "NULL argument");
  eigenSolver3f(cov, evals_ascending, evecs_ascending);
  NB2_API_END
}

int nb2_make_lattice(const nb2_pca* pca, const double dir[3], const double origin[3], double spacing, int bidirectional,
                     nb2_lattice* out)
{
  NB2_API_BEGIN
  if (!pca || !dir || !origin || !out)
    fail(NB2_ERR_INVALID_ARGUMENT, "NULL argument");
  const int rc = makeLattice(*pca, dir, origin, spacing, bidirectional, *out);
  if (rc != NB2_OK)
    return rc;
  NB2_API_END
}

void nb2_plan_params_default(nb2_plan_params* p)
{
  if (!p)
    return;
  std::memset(p, 0, sizeof(*p));
  p->line_spacing = 0.1;  // GUI default (raster_planner_widget.ui:89-100)
  p->bidirectional = 1;
  p->direction_mode = NB2_DIRECTION_PRINCIPAL_AXIS;
  p->origin_mode = NB2_ORIGIN_CENTROID;
  p->raster_post_ops = 1;
  p->emit_edges = 1;
  p->point_spacing = 0.025;
  p->min_segment_size = 0.1;
  p->min_hole_size = 0.1;
  p->shard_rank = 0;
  p->shard_count = 1;
  p->download = 1;
}

int nb2_slice(nb2_context* c, const nb2_lattice* lat, uint64_t plane_begin, uint64_t plane_end, nb2_result** out)
{
  NB2_API_BEGIN
  if (!c || !lat || !out)
    fail(NB2_ERR_INVALID_ARGUMENT, "NULL argument");
  *out = nullptr;
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  if (!(lat->line_spacing > 0.0))
    fail(NB2_ERR_INVALID_ARGUMENT, "line_spacing must be a positive finite number");
  c->timer.begin(c->stream);
  *out = sliceImpl(c, nullptr, lat, plane_begin, plane_end, true);
  NB2_API_END
}

int nb2_plan(nb2_context* c, const nb2_plan_params* prm, nb2_result** out)
{
  NB2_API_BEGIN
  if (!c || !prm || !out)
    fail(NB2_ERR_INVALID_ARGUMENT, "NULL argument");
  *out = nullptr;
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  requireMesh(c);
  if (!c->has_normals)
    fail(NB2_ERR_NO_NORMALS, kNoNormalsMessage);
  if (!c->span_open)
    spanBegin(c);
  c->span_open = false;
  requirePcaPoints(c);
  if (!(prm->line_spacing > 0.0) || !std::isfinite(prm->line_spacing))
    fail(NB2_ERR_INVALID_ARGUMENT, "line_spacing must be a positive finite number");
  if (prm->shard_count > 1 && (prm->shard_rank < 0 || prm->shard_rank >= prm->shard_count))
    fail(NB2_ERR_INVALID_ARGUMENT, "shard_rank out of range");
  c->timer.begin(c->stream);
  FrameParams fp{};
  fp.line_spacing = prm->line_spacing;
  fp.bidirectional = prm->bidirectional;
  fp.direction_mode = prm->direction_mode;
  fp.origin_mode = prm->origin_mode;
  for (int i = 0; i < 3; ++i)
  {
    fp.direction[i] = prm->direction[i];
    fp.origin[i] = prm->origin[i];
  }
  fp.sin_rotation = std::sin(prm->rotation_offset);
  fp.cos_rotation = std::cos(prm->rotation_offset);
  fp.shard_rank = prm->shard_rank;
  fp.shard_count = prm->shard_count > 1 ? prm->shard_count : 1;
  // a slab of a sharded plan stays in HBM: nb2_result_gather sends it to the root GPU-to-GPU
  const bool download = prm->download != 0 && prm->shard_count <= 1;
  // the legacy planner's modifiers run on the device, on the slicer's result while it is still in HBM: only the
  // resampled poses are copied to the host (a sharded plan applies them on the root after the gather, toolpath.cpp)
  const bool device_legacy = prm->legacy != 0 && download;
  if (device_legacy && !(prm->point_spacing > 0.0))
    fail(NB2_ERR_INVALID_ARGUMENT, "point_spacing must be positive");
  // download = 2: the tool paths' positions and normals land in host memory behind this call (nb2_result_poses_segments
  // expands them as they arrive).  Plans whose result the host must post-process here are downloaded in one piece.
  const bool lazy = prm->download == 2 && download && !prm->legacy && !prm->raster_post_ops;
  nb2_result* r = sliceImpl(c, &fp, nullptr, 0, NB2_ALL_PLANES, prm->emit_edges != 0 || (prm->legacy != 0 && !device_legacy),
                            device_legacy ? false : download, lazy);
  r->params = *prm;
  if (device_legacy)
  {
    try
    {
      r = legacyOnDevice(c, r, *prm);
    }
    catch (...)
    {
      freeResult(r);
      throw;
    }
  }
  try
  {
    const bool sharded = prm->shard_count > 1;
    // sharded plans defer the modifiers to nb2_result_gather: they need the whole set of tool paths
    if (!sharded && download)
    {
      if (prm->legacy && !device_legacy && r->n_paths)
        legacyModifiers(r, prm->point_spacing, prm->min_segment_size, prm->min_hole_size);
      if (prm->raster_post_ops && r->n_paths)
      {
        rasterOrganization(r);
        r->post_ops_applied = true;
        directionOfTravelInPlace(*r);
      }
    }
  }
  catch (...)
  {
    freeResult(r);
    throw;
  }
  *out = r;
  NB2_API_END
}

int nb2_plan_mesh(nb2_context* c, const nb2_mesh_view* mesh, const nb2_plan_params* prm, nb2_result** out)
{
  if (out)
    *out = nullptr;
  if (c)
  {
    // (the span of nb2_last_plan_span begins here: the upload's copies and ingest pass belong to the plan)
    std::lock_guard<std::mutex> lock(c->mu);
    DeviceGuard g(c->device);
    c->span_open = false;
    if (cudaSuccess == cudaGetLastError())
    {
      try
      {
        spanBegin(c);
        c->span_open = true;
      }
      catch (const Error&)
      {
      }
    }
  }
  const int rc = nb2_mesh_upload(c, mesh);
  if (rc != NB2_OK && c)
    c->span_open = false;
  return rc != NB2_OK ? rc : nb2_plan(c, prm, out);
}

int nb2_last_plan_span(nb2_context* c, float* ms)
{
  NB2_API_BEGIN
  if (!c || !ms)
    fail(NB2_ERR_INVALID_ARGUMENT, "NULL argument");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  if (!c->span_valid)
    fail(NB2_ERR_INVALID_ARGUMENT, "no plan has run on this context");
  NB2_CUDA(cudaEventSynchronize(c->ev_span_end));
  NB2_CUDA(cudaEventElapsedTime(ms, c->ev_span_begin, c->ev_span_end));
  NB2_API_END
}

int nb2_last_lattice(const nb2_context* c, nb2_lattice* out)
{
  NB2_API_BEGIN
  if (!c || !out)
    fail(NB2_ERR_INVALID_ARGUMENT, "NULL argument");
  if (!c->have_lattice)
    fail(NB2_ERR_INVALID_ARGUMENT, "no plan has run on this context");
  *out = c->last_lattice;
  NB2_API_END
}

static int resultView(const nb2_result* r, nb2_result_view* out, bool layout_only)
{
  NB2_API_BEGIN
  if (!r || !out)
    fail(NB2_ERR_INVALID_ARGUMENT, "NULL argument");
  if (!layout_only)
    waitWaypoints(*r, r->n_waypoints);  // the caller may read positions / normals as soon as this returns
  out->n_paths = r->n_paths;
  out->n_segments = r->n_segments;
  out->n_waypoints = r->n_waypoints;
  out->path_offsets = r->path_offsets;
  out->segment_offsets = r->segment_offsets;
  out->path_plane = r->path_plane;
  out->positions = r->positions;
  out->normals = r->normals;
  out->edge_lo = r->edge_lo;
  out->edge_hi = r->edge_hi;
  out->poses = r->poses;
  out->raster_post_ops_applied = r->post_ops_applied ? 1 : 0;
  out->legacy = r->legacy ? 1 : 0;
  NB2_API_END
}

int nb2_result_get(const nb2_result* r, nb2_result_view* out) { return resultView(r, out, false); }
int nb2_result_layout(const nb2_result* r, nb2_result_view* out) { return resultView(r, out, true); }

int nb2_result_poses(const nb2_result* r, double* out16)
{
  NB2_API_BEGIN
  if (!r || (!out16 && r->n_waypoints))
    fail(NB2_ERR_INVALID_ARGUMENT, "NULL argument");
  if (r->n_waypoints && !r->positions)
    fail(NB2_ERR_INVALID_ARGUMENT, "the tool paths of this result were left on the device (download = 0 or a slab of a "
                                   "sharded plan): gather or re-plan with download = 1");
  expandPoses(*r, out16);
  NB2_API_END
}

int nb2_result_poses_segments(const nb2_result* r, double* const* segment_out)
{
  NB2_API_BEGIN
  if (!r || (!segment_out && r->n_segments))
    fail(NB2_ERR_INVALID_ARGUMENT, "NULL argument");
  if (r->n_waypoints && !r->positions)
    fail(NB2_ERR_INVALID_ARGUMENT, "the tool paths of this result were left on the device (download = 0 or a slab of a "
                                   "sharded plan): gather or re-plan with download = 1");
  expandPosesSegments(*r, segment_out);
  NB2_API_END
}

int nb2_expand_poses(const float* positions, const float* normals, const uint32_t* segment_offsets, uint64_t n_segments,
                     int raster_post_ops_applied, double* out16)
{
  NB2_API_BEGIN
  if (!segment_offsets || (n_segments && segment_offsets[n_segments] && (!positions || !normals || !out16)))
    fail(NB2_ERR_INVALID_ARGUMENT, "NULL argument");
  nb2_result r;
  r.n_segments = n_segments;
  r.n_waypoints = n_segments ? segment_offsets[n_segments] : 0;
  r.segment_offsets = const_cast<uint32_t*>(segment_offsets);
  r.positions = const_cast<float*>(positions);
  r.normals = const_cast<float*>(normals);
  r.post_ops_applied = raster_post_ops_applied != 0;
  expandPoses(r, out16);
  NB2_API_END
}

void nb2_result_free(nb2_result* r)
{
  if (!r)
    return;
  if (r->owner && contextAlive(r->owner))
  {
    std::lock_guard<std::mutex> lock(r->owner->mu);
    freeResult(r);
  }
  else
  {
    r->owner = nullptr;  // context already destroyed: release the pinned block directly
    freeResult(r);
  }
}

int nb2_last_timings(const nb2_context* c, const char** names, float* ms, int capacity)
{
  if (!c || !names || !ms)
    return 0;
  // stages of the last upload, then the stages of the last plan / slice / normals call
  int n = c->timer_upload.collect(names, ms, capacity);
  n += c->timer.collect(names + n, ms + n, capacity - n);
  // not a device stage: the host time from the entry of the last planning call to its last kernel launch (when the host
  // falls behind the device — short kernels, a busy host — the difference shows up as gaps between the kernels)
  if (n < capacity && c->span_valid)
  {
    names[n] = "host:enqueue";
    ms[n] = c->host_enqueue_ms;
    ++n;
  }
  return n;
}

uint64_t nb2_kernel_launches(const nb2_context* c) { return c ? c->launches : 0; }
uint64_t nb2_last_plan_general_path_planes(const nb2_context* c) { return c ? c->last_general_planes : 0; }

int nb2_link_profile(nb2_context* c, uint64_t cycles[16])
{
  NB2_API_BEGIN
  if (!c || !cycles)
    fail(NB2_ERR_INVALID_ARGUMENT, "NULL argument");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  for (int i = 0; i < 16; ++i)
    cycles[i] = 0;
  if (c->linkProfile.ptr)
    NB2_CUDA(cudaMemcpy(cycles, c->linkProfile.ptr, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  NB2_API_END
}

int nb2_timer_start(nb2_context* c)
{
  NB2_API_BEGIN
  if (!c)
    fail(NB2_ERR_INVALID_ARGUMENT, "context is NULL");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  if (!c->ev_start)
  {
    NB2_CUDA(cudaEventCreate(&c->ev_start));
    NB2_CUDA(cudaEventCreate(&c->ev_stop));
  }
  NB2_CUDA(cudaEventRecord(c->ev_start, c->stream));
  NB2_API_END
}

int nb2_timer_stop(nb2_context* c, float* ms)
{
  NB2_API_BEGIN
  if (!c || !ms || !c->ev_start)
    fail(NB2_ERR_INVALID_ARGUMENT, "timer not started");
  std::lock_guard<std::mutex> lock(c->mu);
  DeviceGuard g(c->device);
  NB2_CUDA(cudaEventRecord(c->ev_stop, c->stream));
  NB2_CUDA(cudaEventSynchronize(c->ev_stop));
  NB2_CUDA(cudaEventElapsedTime(ms, c->ev_start, c->ev_stop));
  NB2_API_END
}

}  // extern "C"
