"""CPU baseline, optimised variant (SURVEY.md §8 d: "interval-binned, ... on all host cores"): the oracle's plane loop with
its cells binned by lattice interval (ORC_SCAN_BINNED: each plane visits only the triangles whose plane interval
contains it; identical results to the every-cell scan of vtkCutter) and the lattice cut into contiguous plane slabs, one
forked worker per host core.  Test / measurement infrastructure: run by bench.py's cpu_baseline leg as its own process
(it forks, so it must not share a process with CUDA), never by the product.

    python tools/cpu_optimised.py cfg3 [workers]      -> one JSON line
"""
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import oracle as orc  # noqa: E402
from noether_b200 import meshes  # noqa: E402

WORKLOADS = {"cfg2": (800, 625, 0.005), "cfg3": (2500, 2000, 0.001)}
_state = {}


def _slab(be):
    t = time.perf_counter()
    r = orc.slice_mesh(_state["xyz"], _state["nrm"], _state["tri"], _state["lat"], orc.SLICE_FAITHFUL, orc.SCAN_BINNED, be[0], be[1])
    f = r.flat()
    return int(f.num_paths), int(f.seg_off[-1]) if len(f.seg_off) else 0, time.perf_counter() - t


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    workers = int(sys.argv[2]) if len(sys.argv) > 2 else min(cores, 64)
    nx, ny, spacing = WORKLOADS[name]
    orc.build()
    xyz, tri = meshes.wavy(nx, ny)
    nrm = orc.normals_from_mesh_faces(xyz, tri)
    t0 = time.perf_counter()
    p = orc.pca(xyz, orc.MOMENTS_F32_SEQUENTIAL)
    d = orc.principal_axis_direction(p, 0.0)
    o = orc.centroid(xyz, orc.MOMENTS_F32_SEQUENTIAL)
    lat = orc.make_lattice(p, d, o, spacing, True)
    t_pro = time.perf_counter() - t0
    P = int(lat.num_planes)
    workers = max(1, min(workers, P))
    _state.update(xyz=xyz, nrm=nrm, tri=tri, lat=lat)
    n_slabs = min(P, 2 * workers)   # two slabs per worker, handed out as workers finish (each slab bins the cells again)
    bounds = np.linspace(0, P, n_slabs + 1).round().astype(int)
    slabs = [(int(bounds[i]), int(bounds[i + 1])) for i in range(n_slabs)]
    with mp.get_context("fork").Pool(workers) as pool:   # the mesh is inherited copy-on-write, nothing is pickled in
        pool.map(_slab, [(0, 1)] * workers)              # workers up before the clock starts
        t1 = time.perf_counter()
        res = pool.map(_slab, slabs, chunksize=1)
        t_slices = time.perf_counter() - t1
    t_plan = t_pro + t_slices
    print(json.dumps({"workload": name, "triangles": int(len(tri)), "planes": P, "workers": workers,
                      "host_cores_available": os.cpu_count(), "t_prologue_s": round(t_pro, 4), "t_slices_s": round(t_slices, 4),
                      "t_plan_s": round(t_plan, 4), "tri_per_s": len(tri) / t_plan, "tool_paths": sum(r[0] for r in res),
                      "waypoints": sum(r[1] for r in res), "slab_seconds": [round(r[2], 3) for r in res]}))


if __name__ == "__main__":
    main()
