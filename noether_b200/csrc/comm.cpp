// Multi-GPU plumbing: one process per GPU, plane slabs per rank, gather of per-slab polylines to the root over NCCL.
//
// The reference has no distributed backend at all; its plane loop
// (plane_slicer_raster_planner.cpp:596-628) carries no state between iterations, so slabs of the lattice are
// independent and the only exchange is this gather.  NCCL is resolved with dlopen at nb2_comm_init time so that
// single-GPU users do not need libnccl at all.
#include "nb2_internal.h"

#include <dlfcn.h>

#include <algorithm>
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
  unsigned long long paths, segments, waypoints, edges, planes, plane_begin;
};
}  // namespace

void directionOfTravelInPlace(nb2_result& r);

static size_t shardChunk(const nb2_context* c, size_t bytes)
{
  const size_t world = static_cast<size_t>(c->comm_world > 1 ? c->comm_world : 1);
  const size_t per = (bytes + world - 1) / world;
  return (per + 255) & ~size_t(255);
}

size_t shardedCapacity(const nb2_context* c, size_t bytes)
{
  return shardChunk(c, bytes) * static_cast<size_t>(c->comm_world > 1 ? c->comm_world : 1);
}

void shardedHostToDevice(nb2_context* c, void* dst, const void* src, size_t bytes)
{
  if (c->comm_world <= 1 || !c->nccl_comm)
  {
    NB2_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return;
  }
  const size_t chunk = shardChunk(c, bytes);
  const size_t begin = chunk * static_cast<size_t>(c->comm_rank);
  const size_t mine = begin < bytes ? std::min(chunk, bytes - begin) : 0;
  auto* d = static_cast<uint8_t*>(dst);
  if (mine)
    NB2_CUDA(cudaMemcpyAsync(d + begin, static_cast<const uint8_t*>(src) + begin, mine, cudaMemcpyHostToDevice, c->stream));
  // in place: each rank's contribution already sits at its own offset of the receive buffer
  check(api().AllGather(d + begin, d, chunk, ncclUint8, static_cast<ncclComm_t>(c->nccl_comm), c->stream),
        "ncclAllGather(mesh shares)");
}
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
  catch (const std::exception& e)  // std::bad_alloc of the host vectors: nothing may unwind through the C ABI
  {
    setLastError(std::string("internal error: ") + e.what());
    return NB2_ERR_INTERNAL;
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
  catch (const std::exception& e)  // std::bad_alloc of the host vectors: nothing may unwind through the C ABI
  {
    setLastError(std::string("internal error: ") + e.what());
    return NB2_ERR_INTERNAL;
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
    if (!local->device_resident || local->owner != c)
      fail(NB2_ERR_INVALID_ARGUMENT, "nb2_result_gather takes the slab result of a sharded nb2_plan (shard_count > 1) on "
                                     "this context");
    std::lock_guard<std::mutex> lock(c->mu);
    // the slab lives in the context's output buffers until the next plan overwrites (or reallocates) them
    if (local->plan_serial != c->plan_serial)
      fail(NB2_ERR_INVALID_ARGUMENT, "nb2_result_gather: another plan ran on this context after the slab result was "
                                     "produced; its device buffers no longer hold that slab");
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
    hh[world] = { local->n_paths, local->n_segments, local->n_waypoints, local->dev_edges ? 1ull : 0ull,
                  local->slab_planes, local->slab_begin };
    c->scanState.ensure(sizeof(SlabHeader) * (world + 1));
    auto* dh = c->scanState.as<SlabHeader>();
    NB2_CUDA(cudaMemcpyAsync(dh + world, hh + world, sizeof(SlabHeader), cudaMemcpyHostToDevice, st));
    check(N.AllGather(dh + world, dh, sizeof(SlabHeader), ncclUint8, comm, st), "ncclAllGather(sizes)");
    NB2_CUDA(cudaMemcpyAsync(hh, dh, sizeof(SlabHeader) * world, cudaMemcpyDeviceToHost, st));
    NB2_CUDA(cudaStreamSynchronize(st));
    const std::vector<SlabHeader> hv(hh, hh + world);  // the pinned staging area is reused below
    hh = nullptr;

    uint64_t P = 0, S = 0, W = 0, K = 0;
    bool edges = true;
    for (int r = 0; r < world; ++r)
    {
      P += hv[r].paths;
      S += hv[r].segments;
      W += hv[r].waypoints;
      K += hv[r].planes;
      edges = edges && hv[r].edges;
    }

    // 2. the slab arrays go GPU to GPU, straight out of the buffers the plan wrote (NVLink; NCCL has no gatherv, so a
    //    group of point-to-point transfers): positions, normals, [edge_lo, edge_hi], segment lengths, per-plane counts
    const uint64_t Wl = local->n_waypoints, Sl = local->n_segments, Kl = local->slab_planes;
    const float* d_pos = c->outPos.as<float>();
    const float* d_nrm = c->outNrm.as<float>();
    const uint32_t* d_elo = c->outEdge.as<uint32_t>();
    const uint32_t* d_ehi = c->outEdge.as<uint32_t>() + local->dev_w_cap;
    const unsigned* d_seg = c->outSegLen.as<unsigned>();
    const unsigned* d_meta = c->planeMeta.as<unsigned>();
    if (rank != root)
    {
      check(N.GroupStart(), "ncclGroupStart");
      if (Wl)
      {
        check(N.Send(d_pos, 3 * Wl, ncclFloat32, root, comm, st), "ncclSend(positions)");
        check(N.Send(d_nrm, 3 * Wl, ncclFloat32, root, comm, st), "ncclSend(normals)");
        if (edges)
        {
          check(N.Send(d_elo, Wl, ncclUint32, root, comm, st), "ncclSend(edge_lo)");
          check(N.Send(d_ehi, Wl, ncclUint32, root, comm, st), "ncclSend(edge_hi)");
        }
      }
      if (Sl)
        check(N.Send(d_seg, Sl, ncclUint32, root, comm, st), "ncclSend(segments)");
      if (Kl)
        check(N.Send(d_meta, 2 * Kl, ncclUint32, root, comm, st), "ncclSend(planes)");
      check(N.GroupEnd(), "ncclGroupEnd");
      NB2_CUDA(cudaStreamSynchronize(st));
      return NB2_OK;
    }

    // root: receive area (the hit records are dead between plans): pos | nrm | elo | ehi | seg | meta, slabs in rank order
    const size_t words = 6 * W + (edges ? 2 * W : 0) + S + 2 * K + 64;
    c->hitRec.ensure(sizeof(uint32_t) * words);
    uint32_t* base = c->hitRec.as<uint32_t>();
    float* r_pos = reinterpret_cast<float*>(base);
    float* r_nrm = r_pos + 3 * W;
    uint32_t* r_elo = reinterpret_cast<uint32_t*>(r_nrm + 3 * W);
    uint32_t* r_ehi = r_elo + (edges ? W : 0);
    uint32_t* r_seg = r_ehi + (edges ? W : 0);
    uint32_t* r_meta = r_seg + S;
    std::vector<uint64_t> w_off(world + 1, 0), s_off(world + 1, 0), k_off(world + 1, 0);
    for (int r = 0; r < world; ++r)
    {
      w_off[r + 1] = w_off[r] + hv[r].waypoints;
      s_off[r + 1] = s_off[r] + hv[r].segments;
      k_off[r + 1] = k_off[r] + hv[r].planes;
    }
    check(N.GroupStart(), "ncclGroupStart");
    for (int r = 0; r < world; ++r)
    {
      if (r == root)
        continue;
      const uint64_t Wr = hv[r].waypoints, Sr = hv[r].segments, Kr = hv[r].planes;
      if (Wr)
      {
        check(N.Recv(r_pos + 3 * w_off[r], 3 * Wr, ncclFloat32, r, comm, st), "ncclRecv(positions)");
        check(N.Recv(r_nrm + 3 * w_off[r], 3 * Wr, ncclFloat32, r, comm, st), "ncclRecv(normals)");
        if (edges)
        {
          check(N.Recv(r_elo + w_off[r], Wr, ncclUint32, r, comm, st), "ncclRecv(edge_lo)");
          check(N.Recv(r_ehi + w_off[r], Wr, ncclUint32, r, comm, st), "ncclRecv(edge_hi)");
        }
      }
      if (Sr)
        check(N.Recv(r_seg + s_off[r], Sr, ncclUint32, r, comm, st), "ncclRecv(segments)");
      if (Kr)
        check(N.Recv(r_meta + 2 * k_off[r], 2 * Kr, ncclUint32, r, comm, st), "ncclRecv(planes)");
    }
    check(N.GroupEnd(), "ncclGroupEnd");
    // the root's own slab joins the others on the device
    if (Wl)
    {
      NB2_CUDA(cudaMemcpyAsync(r_pos + 3 * w_off[root], d_pos, 12 * Wl, cudaMemcpyDeviceToDevice, st));
      NB2_CUDA(cudaMemcpyAsync(r_nrm + 3 * w_off[root], d_nrm, 12 * Wl, cudaMemcpyDeviceToDevice, st));
      if (edges)
      {
        NB2_CUDA(cudaMemcpyAsync(r_elo + w_off[root], d_elo, 4 * Wl, cudaMemcpyDeviceToDevice, st));
        NB2_CUDA(cudaMemcpyAsync(r_ehi + w_off[root], d_ehi, 4 * Wl, cudaMemcpyDeviceToDevice, st));
      }
    }
    if (Sl)
      NB2_CUDA(cudaMemcpyAsync(r_seg + s_off[root], d_seg, 4 * Sl, cudaMemcpyDeviceToDevice, st));
    if (Kl)
      NB2_CUDA(cudaMemcpyAsync(r_meta + 2 * k_off[root], d_meta, 8 * Kl, cudaMemcpyDeviceToDevice, st));

    // 3. one device -> pinned-host copy per array, into the final result
    nb2_result* out = allocResult(c, P, S, W, edges, false);
    out->params = local->params;
    out->lattice = local->lattice;
    c->hostSmall.ensure(4096 + 8 * K + 64);
    uint32_t* h_meta = reinterpret_cast<uint32_t*>(c->hostSmall.as<uint8_t>() + 4096);
    try
    {
      if (W)
      {
        NB2_CUDA(cudaMemcpyAsync(out->positions, r_pos, 12 * W, cudaMemcpyDeviceToHost, st));
        NB2_CUDA(cudaMemcpyAsync(out->normals, r_nrm, 12 * W, cudaMemcpyDeviceToHost, st));
        if (edges)
        {
          NB2_CUDA(cudaMemcpyAsync(out->edge_lo, r_elo, 4 * W, cudaMemcpyDeviceToHost, st));
          NB2_CUDA(cudaMemcpyAsync(out->edge_hi, r_ehi, 4 * W, cudaMemcpyDeviceToHost, st));
        }
      }
      if (S)
        NB2_CUDA(cudaMemcpyAsync(out->segment_offsets + 1, r_seg, 4 * S, cudaMemcpyDeviceToHost, st));
      if (K)
        NB2_CUDA(cudaMemcpyAsync(h_meta, r_meta, 8 * K, cudaMemcpyDeviceToHost, st));
      NB2_CUDA(cudaStreamSynchronize(st));
    }
    catch (...)
    {
      freeResult(out);
      throw;
    }
    // segment lengths -> offsets; planes with segments -> tool paths (slabs in rank order = plane order)
    out->segment_offsets[0] = 0;
    for (uint64_t s2 = 0; s2 < S; ++s2)
      out->segment_offsets[s2 + 1] += out->segment_offsets[s2];
    uint64_t p_acc = 0, s_acc = 0;
    for (int r = 0; r < world; ++r)
      for (uint64_t k = 0; k < hv[r].planes; ++k)
      {
        const unsigned sk = h_meta[2 * (k_off[r] + k) + 1];
        if (!sk)
          continue;
        out->path_offsets[p_acc] = static_cast<uint32_t>(s_acc);
        out->path_plane[p_acc] = static_cast<int64_t>(hv[r].plane_begin + k);
        s_acc += sk;
        ++p_acc;
      }
    out->path_offsets[p_acc] = static_cast<uint32_t>(s_acc);
    if (p_acc != P || s_acc != S)
    {
      freeResult(out);
      fail(NB2_ERR_INTERNAL, "gathered slab metadata is inconsistent");
    }

    // 4. the modifiers that need the whole set of tool paths
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
  catch (const std::exception& e)  // std::bad_alloc of the host vectors: nothing may unwind through the C ABI
  {
    setLastError(std::string("internal error: ") + e.what());
    return NB2_ERR_INTERNAL;
  }
  return NB2_OK;
}

}  // extern "C"
