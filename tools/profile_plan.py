#!/usr/bin/env python
"""One device-resident plan of a bench workload, bracketed by cudaProfilerStart/Stop, for ncu:

    python tools/profile_plan.py cfg3 > gpurun_out/plain.log 2>&1 &&
    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/prof \
        python tools/profile_plan.py cfg3

The mesh (interleaved cloud + triangles) is resident in HBM, as in bench.py's `value` leg; two warm plans run before
the profiled one.  Prints the CUDA-event stage times of the profiled plan (meaningless under ncu).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench
    import noether_b200 as nb
    from noether_b200 import _capi, meshes

    name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    emit_edges = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    xyz, tri, spacing, desc = bench.build_workload(name)
    V, F = xyz.shape[0], tri.shape[0]
    dev = torch.device("cuda", 0)
    ctx = nb.Context(0)
    ctx.upload(meshes.pack_blob(xyz, None, tri, "PointXYZ"), force=True)
    nrm = ctx.normals_from_mesh_faces(V)
    # resident layout: what nb2_mesh_upload(NB2_MESH_GEOMETRY_ONLY) leaves in HBM ({xyz, normal} records of 24 bytes), or the
    # 48-byte PointNormal records of the host cloud (NB2_PROFILE_LAYOUT=PointNormal)
    blob = meshes.pack_blob(xyz, nrm, tri, os.environ.get("NB2_PROFILE_LAYOUT", "Packed24"))
    d_cloud = torch.from_numpy(blob.data).to(dev)
    d_tri = torch.from_numpy(tri.reshape(-1).view(np.int32)).to(dev)
    torch.cuda.synchronize()

    prm = _capi.PlanParams()
    _capi.load().nb2_plan_params_default(prm)
    prm.line_spacing = spacing
    prm.emit_edges = emit_edges
    prm.download = 0
    # NB2_PROFILE_SHARDS=N: the slab rank N // 2 of an N-way sharded plan plans on this GPU (what a rank of an N-GPU job runs)
    shards = int(os.environ.get("NB2_PROFILE_SHARDS", "1"))
    if shards > 1:
        prm.shard_count, prm.shard_rank = shards, shards // 2

    def step():
        r = ctx.plan_mesh_raw(d_cloud.data_ptr(), V, blob.point_step, blob.off_xyz, blob.off_normal, d_tri.data_ptr(), F, prm)
        n = (r.n_paths, r.n_segments, r.n_waypoints)
        r.free()
        return n

    flush = torch.empty(256 * 2**20, dtype=torch.uint8, device=dev)
    for _ in range(2):
        step()
    flush.zero_()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ctx.timer_start()
    n = step()
    total = ctx.timer_stop()
    st = ctx.timings()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(name, desc)
    print("paths/segments/waypoints", n)
    print("stages ms", {k: round(v, 4) for k, v in st.items()})
    print("step ms (events) %.4f, sum of stages %.4f, library span %.4f" % (total, sum(v for k, v in st.items() if not k.startswith("host:")), ctx.last_plan_span_ms()))
    print("planes linked by the general path:", ctx.general_path_planes())
    if os.environ.get("NB2_LINK_PROFILE"):
        cyc = ctx.link_profile()
        names = ["init", "hash", "claim rounds", "pair+dead", "link", "walk", "jump", "heads+order", "output"]
        tot = max(1, sum(cyc))
        print("linker phases (cycles of thread 0 summed over CTAs and planes):",
              {n: f"{100 * c / tot:.1f}%" for n, c in zip(names, cyc)}, "total cycles", tot)
    ctx.close()


if __name__ == "__main__":
    main()
