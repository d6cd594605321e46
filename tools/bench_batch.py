#!/usr/bin/env python
"""BASELINE.json configs[4] (cfg5): a batch of 64 part meshes (1.5 - 2.5 M triangles each, random poses), whole meshes
sharded over the GPUs of one box by LPT on the triangle counts (noether_b200.dist.lpt_assign); 2 mm raster spacing,
PlaneSlicer(PrincipalAxis, Centroid) on vertex normals from NormalsFromMeshFaces.  No exchange between ranks.

    python tools/bench_batch.py [--meshes 64] [--threads 2]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_batch.py

Per rank the meshes sit in pinned host memory; each is uploaded, planned and its tool paths downloaded (e2e).  With
--threads 2 a rank drives two contexts from two host threads, so the PCIe copies of one mesh overlap the kernels of the
other.  Prints one JSON line (rank 0): batch wall time = max over ranks, meshes / s, triangles / s.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(argv=None):
    """returns the result dict on rank 0 (None elsewhere); bench.py --workload cfg5 wraps it into the bench line"""
    ap = argparse.ArgumentParser()
    ap.add_argument("--meshes", type=int, default=64)
    ap.add_argument("--threads", type=int, default=2)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--quiet", action="store_true")
    args = ap.parse_args(argv)

    import torch

    import noether_b200 as nb
    from noether_b200 import meshes
    from noether_b200.dist import RankInfo, init_process_group, lpt_assign, max_over_ranks

    info = RankInfo.from_env()
    torch.cuda.set_device(info.local_rank)
    dev = torch.device("cuda", info.local_rank)
    init_process_group(info, "nccl")
    sizes = meshes.cfg5_triangle_counts(args.meshes)
    share = lpt_assign(sizes, info.world)[info.rank]

    # this rank's meshes, with normals, in pinned host memory
    prep = nb.Context(info.local_rank)
    batch = []
    for k in share:
        xyz, tri, _ = meshes.cfg5_mesh(k)
        prep.upload(meshes.pack_blob(xyz, None, tri, "PointXYZ"), force=True)
        nrm = prep.normals_from_mesh_faces(xyz.shape[0])
        blob = meshes.pack_blob(xyz, nrm, tri)
        h_cloud = torch.empty(blob.data.size, dtype=torch.uint8, pin_memory=True)
        h_cloud.numpy()[:] = blob.data
        h_tri = torch.empty(tri.size, dtype=torch.int32, pin_memory=True)
        h_tri.numpy()[:] = tri.reshape(-1).view(np.int32)
        batch.append((k, meshes.MeshBlob(h_cloud.numpy(), xyz.shape[0], blob.point_step, blob.off_xyz, blob.off_normal,
                                         h_tri.numpy().view(np.uint32).reshape(-1, 3)), (h_cloud, h_tri)))
    prep.close()

    ctxs = [nb.Context(info.local_rank) for _ in range(max(1, args.threads))]
    planners = []
    for c in ctxs:
        p = nb.PlaneSlicerRasterPlanner(nb.PrincipalAxisDirectionGenerator(0.0, c), nb.CentroidOriginGenerator(c), c)
        p.set_line_spacing(0.002)
        p.emit_edges = False
        planners.append(p)
    counts = {}

    def worker(t):
        for j in range(t, len(batch), len(ctxs)):
            k, mesh, _ = batch[j]
            ctxs[t]._mesh_key = None
            r = planners[t].plan(mesh)
            counts[k] = (r.n_paths, r.n_waypoints)
            r.free()

    def run_once():
        th = [threading.Thread(target=worker, args=(t,)) for t in range(len(ctxs))]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    run_once()  # warm-up: buffers, pinned result blocks
    times = []
    for _ in range(args.repeat):
        if info.world > 1:
            torch.distributed.barrier()
        times.append(max_over_ranks(run_once(), info.world, dev))
    res = None
    launches = sum(c.kernel_launches() for c in ctxs)
    if info.rank == 0:
        best = min(times)
        tris = sum(sizes)
        res = {"workload": f"cfg5: {args.meshes} meshes, {tris} triangles, 2 mm spacing, whole meshes by LPT",
               "n_gpus": info.world, "host_threads_per_gpu": len(ctxs), "batch_s": best, "batch_s_mean": float(np.mean(times)),
               "meshes_per_s": args.meshes / best, "triangles_per_s": tris / best, "triangles": int(tris),
               "rank0_kernel_launches": int(launches), "repeat": args.repeat,
               "h2d_bytes": int(sum(b[1].data.size + b[1].tri.size * 4 for b in batch)),
               "rank0_share": share, "rank0_paths_waypoints": [counts[k] for k in share][:3]}
        if not args.quiet:
            print(json.dumps(res))
    for c in ctxs:
        c.close()
    return res


if __name__ == "__main__":
    main()
