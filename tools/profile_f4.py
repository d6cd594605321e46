#!/usr/bin/env python
"""NormalEstimationPCL and EuclideanClustering on the cfg3 cloud (5.0 M points), for ncu:

    python tools/profile_f4.py > gpurun_out/plain_f4.log 2>&1 &&
    ncu --set full --clock-control none -k regex:"k_pcl_normals|k_clu_link|k_clu_insert" -o gpurun_out/prof_f4 python tools/profile_f4.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    import noether_b200 as nb
    from noether_b200 import meshes

    xyz, tri, spacing, desc = bench.build_workload("cfg3")
    ctx = nb.Context(0)
    ctx.upload(meshes.pack_blob(xyz, None, tri, "PointXYZ"), force=True)
    h = 1.0 / 2500  # point spacing of the surface
    for r in (3.0 * h,):
        nrm, cur = ctx.normal_estimation_pcl(xyz.shape[0], r, (0.0, 0.0, 10.0))
        print("NormalEstimationPCL radius %.5f" % r, {k: round(v, 3) for k, v in ctx.timings().items()}, "NaN rows", int(np.isnan(nrm).any(1).sum()))
    labels = np.zeros(xyz.shape[0], np.uint32)
    from noether_b200 import _capi
    import ctypes as C
    _capi.check(ctx._lib.nb2_euclidean_cluster_labels(ctx._h, C.c_double(1.5 * h), labels.ctypes.data_as(C.c_void_p), C.c_uint64(labels.size)))
    print("EuclideanClustering tolerance %.5f: %d clusters" % (1.5 * h, np.unique(labels).size))
    ctx.close()


if __name__ == "__main__":
    main()
