#!/usr/bin/env python
"""BASELINE.json configs[3] (cfg4): NormalsFromMeshFaces + PCA frame on the 50 M-triangle mesh with permuted vertex /
face order, device-resident input; prints CUDA-event stage times and the achieved fraction of the HBM roofline
(B_normals = 12 F + 24 V, B_pca = 12 V).

    python tools/profile_normals.py [cfg4|cfg4-grid-order|cfg3|cfg2]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SIZES = {"cfg4": (6250, 4000, True), "cfg4-grid-order": (6250, 4000, False), "cfg3": (2500, 2000, False),
         "cfg2": (800, 625, False)}


def measure(name="cfg4", iters=4):
    """stage times of NormalsFromMeshFaces and of the PCA frame on the device-resident mesh: a dict (bench.py --workload cfg4
    wraps it into the bench line)"""
    import torch

    import noether_b200 as nb
    from noether_b200 import meshes

    nx, ny, permuted = SIZES[name]
    xyz, tri = meshes.wavy(nx, ny)
    if permuted:
        xyz, tri = meshes.permute(xyz, tri, seed=42)
    V, F = xyz.shape[0], tri.shape[0]
    blob = meshes.pack_blob(xyz, None, tri, "PointXYZ")
    dev = torch.device("cuda", 0)
    d_cloud = torch.from_numpy(blob.data).to(dev)
    d_tri = torch.from_numpy(tri.reshape(-1).view(np.int32)).to(dev)
    torch.cuda.synchronize()
    ctx = nb.Context(0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    out = []
    launches0 = ctx.kernel_launches()
    for it in range(max(2, iters)):
        ctx.upload_raw(d_cloud.data_ptr(), V, blob.point_step, blob.off_xyz, blob.off_normal, d_tri.data_ptr(), F)
        ctx.timer_start()
        ctx.normals_from_mesh_faces(V, download=False)
        t_n = ctx.timer_stop()
        st = ctx.timings()
        ctx.timer_start()
        ctx.upload_raw(d_cloud.data_ptr(), V, blob.point_step, blob.off_xyz, blob.off_normal, d_tri.data_ptr(), F)
        ctx.pca()
        t_p = ctx.timer_stop()
        out.append((t_n, t_p, st))
    t_n = min(o[0] for o in out[1:])
    t_p = min(o[1] for o in out[1:])
    b_n, b_p = 12 * F + 24 * V, 12 * V
    res = {"workload": name, "V": V, "F": F, "permuted": permuted,
           "normals_ms": t_n, "normals_GBps": b_n / t_n / 1e6, "normals_frac_of_peak": b_n / t_n / 1e6 / peak,
           "pca_frame_ms": t_p, "pca_GBps": b_p / t_p / 1e6, "pca_frac_of_peak": b_p / t_p / 1e6 / peak,
           "normals_ms_mean": float(np.mean([o[0] for o in out[1:]])), "pca_frame_ms_mean": float(np.mean([o[1] for o in out[1:]])),
           "launches_per_iteration": (ctx.kernel_launches() - launches0) // max(2, iters), "peak_GBps": peak,
           "stages_ms": {k: round(v, 4) for k, v in out[-1][2].items()}}
    ctx.close()
    return res


def main():
    print(json.dumps(measure(sys.argv[1] if len(sys.argv) > 1 else "cfg4")))


if __name__ == "__main__":
    main()
