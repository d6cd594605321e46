// Multi-GPU plumbing: one process per GPU, plane slabs per rank, gather of per-slab polylines to the root over NCCL.
//
// The reference has no distributed backend at all; its plane loop
// (plane_slicer_raster_planner.cpp:596-628) carries no state between iterations, so slabs of the lattice are
// independent and the only exchange is this gather.  NCCL is resolved with dlopen at nb2_comm_init time so that
// single-GPU users do not need libnccl at all.
#include "nb2_internal.h"

#include <dlfcn.h>

#include <cstring>

namespace nb2
{
namespace
{
// minimal NCCL surface (ABI-stable since NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct
{
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;  // ncclSuccess == 0
enum
{
  ncclInt8 = 0,
  ncclUint8 = 1,
  ncclInt32 = 2,
  ncclUint32 = 3,
  ncclInt64 = 4,
  ncclUint64 = 5,
  ncclFloat32 = 7
};

struct NcclApi
{
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& api()
{
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : { "libnccl.so.2", "libnccl.so" })
    {
      a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle)
        break;
    }
    if (!a.handle)
      return;
    auto sym = [&](const char* n) { return dlsym(a.handle, n); };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
    a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
    a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
  });
  if (!a.handle || !a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.Send || !a.Recv || !a.GroupStart || !a.GroupEnd)
    fail(NB2_ERR_NCCL, "libnccl.so.2 could not be loaded (dlopen) or lacks a required symbol");
  return a;
}

void check(ncclResult_t r, const char* what)
{
  if (r != 0)
    fail(NB2_ERR_NCCL, std::string("NCCL failure in ") + what + ": " + (api().GetErrorString ? api().GetErrorString(r) : "?"));
}

struct SlabHeader
{
  unsigned long long paths, segments, waypoints, edges;
};
}  // namespace

void directionOfTravelInPlace(nb2_result& r);
}  // namespace nb2

using namespace nb2;

extern "C" {

int nb2_comm_unique_id(uint8_t out[NB2_UNIQUE_ID_BYTES])
{
  try
  {
    static_assert(sizeof(ncclUniqueId) == NB2_UNIQUE_ID_BYTES, "unique id size");
    ncclUniqueId id;
    check(api().GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(out, &id, sizeof(id));
  }
  catch (const Error& e)
  {
    setLastError(e.msg);
    return e.status;
  }
  return NB2_OK;
}

int nb2_comm_init(nb2_context* c, const uint8_t unique_id[NB2_UNIQUE_ID_BYTES], int rank, int world)
{
  try
  {
    if (!c || !unique_id || world < 1 || rank < 0 || rank >= world)
      fail(NB2_ERR_INVALID_ARGUMENT, "bad communicator arguments");
    std::lock_guard<std::mutex> lock(c->mu);
    NB2_CUDA(cudaSetDevice(c->device));
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof(id));
    ncclComm_t comm = nullptr;
    check(api().CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
    c->nccl_comm = comm;
    c->comm_rank = rank;
    c->comm_world = world;
  }
  catch (const Error& e)
  {
    setLastError(e.msg);
    return e.status;
  }
  return NB2_OK;
}

int nb2_comm_destroy(nb2_context* c)
{
  if (c && c->nccl_comm)
  {
    try
    {
      api().CommDestroy(static_cast<ncclComm_t>(c->nccl_comm));
    }
    catch (...)
    {
    }
    c->nccl_comm = nullptr;
    c->comm_world = 1;
    c->comm_rank = 0;
  }
  return NB2_OK;
}

int nb2_result_gather(nb2_context* c, const nb2_result* local, int root, nb2_result** gathered)
{
  try
  {
    if (!c || !local || !gathered)
      fail(NB2_ERR_INVALID_ARGUMENT, "NULL argument");
    *gathered = nullptr;
    if (!c->nccl_comm)
      fail(NB2_ERR_NCCL, "nb2_comm_init has not been called on this context");
    std::lock_guard<std::mutex> lock(c->mu);
    NB2_CUDA(cudaSetDevice(c->device));
    auto comm = static_cast<ncclComm_t>(c->nccl_comm);
    const int world = c->comm_world, rank = c->comm_rank;
    if (root < 0 || root >= world)
      fail(NB2_ERR_INVALID_ARGUMENT, "root out of range");
    NcclApi& N = api();
    cudaStream_t st = c->stream;

    // 1. everybody learns every slab's sizes
    c->hostSmall.ensure(4096 + sizeof(SlabHeader) * (world + 1));
    auto* hh = reinterpret_cast<SlabHeader*>(c->hostSmall.as<uint8_t>() + 4096);
    hh[world] = { local->n_paths, local->n_segments, local->n_waypoints, local->edge_lo ? 1ull : 0ull };
    c->scanState.ensure(sizeof(SlabHeader) * (world + 1));
    auto* dh = c->scanState.as<SlabHeader>();
    NB2_CUDA(cudaMemcpyAsync(dh + world, hh + world, sizeof(SlabHeader), cudaMemcpyHostToDevice, st));
    check(N.AllGather(dh + world, dh, sizeof(SlabHeader), ncclUint8, comm, st), "ncclAllGather(sizes)");
    NB2_CUDA(cudaMemcpyAsync(hh, dh, sizeof(SlabHeader) * world, cudaMemcpyDeviceToHost, st));
    NB2_CUDA(cudaStreamSynchronize(st));

    uint64_t P = 0, S = 0, W = 0;
    bool edges = true;
    for (int r = 0; r < world; ++r)
    {
      P += hh[r].paths;
      S += hh[r].segments;
      W += hh[r].waypoints;
      edges = edges && hh[r].edges;
    }

    // 2. stage this slab on the device: positions | normals | edge_lo | edge_hi | segment lengths | path plane | path segs
    //    (float / uint32 words; path planes as two words each)
    auto words = [&](const SlabHeader& h) {
      return 6 * h.waypoints + (edges ? 2 * h.waypoints : 0) + h.segments + 3 * h.paths;
    };
    const uint64_t my_words = words(hh[rank]);
    std::vector<uint32_t> pack(my_words ? my_words : 1);
    {
      uint32_t* p = pack.data();
      std::memcpy(p, local->positions, 12 * local->n_waypoints);
      p += 3 * local->n_waypoints;
      std::memcpy(p, local->normals, 12 * local->n_waypoints);
      p += 3 * local->n_waypoints;
      if (edges)
      {
        std::memcpy(p, local->edge_lo, 4 * local->n_waypoints);
        p += local->n_waypoints;
        std::memcpy(p, local->edge_hi, 4 * local->n_waypoints);
        p += local->n_waypoints;
      }
      for (uint64_t s = 0; s < local->n_segments; ++s)
        *p++ = local->segment_offsets[s + 1] - local->segment_offsets[s];
      for (uint64_t q = 0; q < local->n_paths; ++q)
      {
        const uint64_t plane = static_cast<uint64_t>(local->path_plane[q]);
        *p++ = static_cast<uint32_t>(plane & 0xffffffffull);
        *p++ = static_cast<uint32_t>(plane >> 32);
        *p++ = local->path_offsets[q + 1] - local->path_offsets[q];
      }
    }
    uint64_t total_words = 0;
    std::vector<uint64_t> word_off(world + 1, 0);
    for (int r = 0; r < world; ++r)
    {
      word_off[r] = total_words;
      total_words += words(hh[r]);
    }
    word_off[world] = total_words;

    // the per-rank send buffer lives in linkScratch (free between plans); the root's receive area follows it
    const size_t send_bytes = sizeof(uint32_t) * (my_words + 1);
    const size_t recv_bytes = rank == root ? sizeof(uint32_t) * (total_words + 1) : 0;
    c->linkScratch.ensure(((send_bytes + 255) & ~size_t(255)) + recv_bytes);
    auto* d_send = c->linkScratch.as<uint32_t>();
    auto* d_recv = reinterpret_cast<uint32_t*>(c->linkScratch.as<uint8_t>() + ((send_bytes + 255) & ~size_t(255)));
    if (my_words)
      NB2_CUDA(cudaMemcpyAsync(d_send, pack.data(), sizeof(uint32_t) * my_words, cudaMemcpyHostToDevice, st));

    // 3. grouped point-to-point gather (NCCL has no gatherv)
    check(N.GroupStart(), "ncclGroupStart");
    if (rank == root)
    {
      for (int r = 0; r < world; ++r)
      {
        const uint64_t n = word_off[r + 1] - word_off[r];
        if (r == root || n == 0)
          continue;
        check(N.Recv(d_recv + word_off[r], n, ncclUint32, r, comm, st), "ncclRecv");
      }
    }
    else if (my_words)
    {
      check(N.Send(d_send, my_words, ncclUint32, root, comm, st), "ncclSend");
    }
    check(N.GroupEnd(), "ncclGroupEnd");
    if (rank != root)
    {
      NB2_CUDA(cudaStreamSynchronize(st));
      return NB2_OK;
    }
    if (my_words)
      NB2_CUDA(cudaMemcpyAsync(d_recv + word_off[root], d_send, sizeof(uint32_t) * my_words, cudaMemcpyDeviceToDevice, st));

    // 4. root: unpack into one result, slabs in rank order = plane order
    std::vector<uint32_t> all(total_words ? total_words : 1);
    if (total_words)
      NB2_CUDA(cudaMemcpyAsync(all.data(), d_recv, sizeof(uint32_t) * total_words, cudaMemcpyDeviceToHost, st));
    NB2_CUDA(cudaStreamSynchronize(st));

    nb2_result* out = allocResult(c, P, S, W, edges, false);
    out->params = local->params;
    out->lattice = local->lattice;
    uint64_t p_acc = 0, s_acc = 0, w_acc = 0;
    for (int r = 0; r < world; ++r)
    {
      const SlabHeader& h = hh[r];
      const uint32_t* p = all.data() + word_off[r];
      std::memcpy(out->positions + 3 * w_acc, p, 12 * h.waypoints);
      p += 3 * h.waypoints;
      std::memcpy(out->normals + 3 * w_acc, p, 12 * h.waypoints);
      p += 3 * h.waypoints;
      if (edges)
      {
        std::memcpy(out->edge_lo + w_acc, p, 4 * h.waypoints);
        p += h.waypoints;
        std::memcpy(out->edge_hi + w_acc, p, 4 * h.waypoints);
        p += h.waypoints;
      }
      uint64_t w_run = w_acc;
      for (uint64_t s = 0; s < h.segments; ++s)
      {
        out->segment_offsets[s_acc + s] = static_cast<uint32_t>(w_run);
        w_run += *p++;
      }
      uint64_t s_run = s_acc;
      for (uint64_t q = 0; q < h.paths; ++q)
      {
        const uint64_t plane = static_cast<uint64_t>(p[0]) | (static_cast<uint64_t>(p[1]) << 32);
        out->path_plane[p_acc + q] = static_cast<int64_t>(plane);
        out->path_offsets[p_acc + q] = static_cast<uint32_t>(s_run);
        s_run += p[2];
        p += 3;
      }
      p_acc += h.paths;
      s_acc += h.segments;
      w_acc += h.waypoints;
    }
    out->path_offsets[P] = static_cast<uint32_t>(S);
    out->segment_offsets[S] = static_cast<uint32_t>(W);

    // 5. the modifiers that need the whole set of tool paths
    try
    {
      const nb2_plan_params& prm = out->params;
      if (prm.legacy && out->n_paths)
        legacyModifiers(out, prm.point_spacing, prm.min_segment_size, prm.min_hole_size);
      if (prm.raster_post_ops && out->n_paths)
      {
        rasterOrganization(out);
        out->post_ops_applied = true;
        directionOfTravelInPlace(*out);
      }
    }
    catch (...)
    {
      freeResult(out);
      throw;
    }
    *gathered = out;
  }
  catch (const Error& e)
  {
    setLastError(e.msg);
    return e.status;
  }
  return NB2_OK;
}

}  // extern "C"
