#!/usr/bin/env python
"""How often does the reference's ACTUAL linking differ from its documented INTENT?  (CPU only; oracle = test infrastructure)

    python tools/faithful_vs_intent.py [cfg1 cfg1-default cfg2 cfg3 hole ...]  > profiles/r02_faithful_vs_intent.json

The reference joins the pieces vtkStripper returns with a topological sort and a stack join that only joins pieces that end
up stack-adjacent (plane_slicer_raster_planner.cpp:302-358); when the sorted order interleaves pieces of two chains it
returns more segments than there are connected chains.  The oracle restates both behaviours:
SLICE_FAITHFUL (stripper + sort + join as written) and SLICE_INTENT (one segment per maximal open chain — the documented
contract, plane_slicer_raster_planner.cpp:290-301, and what the GPU path produces).  This tool slices each configuration in
both modes and reports, per configuration, on how many lattice planes the two differ, and how.

The waypoint SET of a plane is the same in both modes (FAITHFUL only cuts chains into more pieces); that is asserted here.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle as orc  # noqa: E402
from noether_b200 import meshes  # noqa: E402


def configs():
    ply = os.path.join(ROOT, "tests", "golden", "wavy_mesh_with_hole.ply")

    def fixture(spacing):
        xyz, nrm, sizes, idx = orc.read_ply(open(ply, "rb").read())
        assert (sizes == 3).all()
        tri = idx.reshape(-1, 3)
        return xyz, nrm, tri, dict(line_spacing=spacing, direction=[1, 0, 0], origin=[0, 0, 0])

    def wavy(nx, ny, spacing, hole=False):
        xyz, tri = meshes.wavy(nx, ny)
        if hole:
            xyz, tri = meshes.cut_hole(xyz, tri)
        nrm = orc.normals_from_mesh_faces(xyz, tri)
        return xyz, nrm, tri, dict(line_spacing=spacing)

    return {
        "cfg1": ("reference test mesh, the reference test's settings (line_spacing 1.05, FixedDirection x, FixedOrigin 0)",
                 lambda: fixture(1.05)),
        "cfg1-default": ("reference test mesh, default YAML spacing 0.1 (SURVEY A.1: 3 planes graze the hole rim)",
                         lambda: fixture(0.1)),
        "cfg1-0.05": ("reference test mesh, spacing 0.05", lambda: fixture(0.05)),
        "cfg2": ("wavy(800,625) 1 M triangles, 5 mm, PrincipalAxis(0) + Centroid", lambda: wavy(800, 625, 0.005)),
        "cfg2-hole": ("wavy(800,625) with the disc hole, 5 mm", lambda: wavy(800, 625, 0.005, hole=True)),
        "cfg3": ("wavy(2500,2000) 10 M triangles, 1 mm, PrincipalAxis(0) + Centroid", lambda: wavy(2500, 2000, 0.001)),
        "cfg3-hole": ("wavy(2500,2000) with the disc hole, 1 mm", lambda: wavy(2500, 2000, 0.001, hole=True)),
    }


def per_plane(flat):
    """{lattice plane: (segments, sorted waypoint rows)}"""
    out = {}
    for p in range(flat.num_paths):
        s0, s1 = int(flat.path_off[p]), int(flat.path_off[p + 1])
        w0, w1 = int(flat.seg_off[s0]), int(flat.seg_off[s1])
        out[int(flat.path_plane[p])] = (s1 - s0, w1 - w0, s0, s1)
    return out


def compare(name, desc, make):
    t0 = time.perf_counter()
    xyz, nrm, tri, kw = make()
    res = {}
    for mode, tag in ((orc.SLICE_INTENT, "intent"), (orc.SLICE_FAITHFUL, "faithful")):
        paths, lat, _ = orc.plan_plane_slicer(xyz, nrm, tri, kw["line_spacing"], True, kw.get("direction"), kw.get("origin"),
                                              slice_mode=mode, scan_mode=orc.SCAN_BINNED, post_ops=False)
        res[tag] = paths.flat()
    fi, ff = res["intent"], res["faithful"]
    pi, pf = per_plane(fi), per_plane(ff)
    assert set(pi) == set(pf), "the two modes disagree on which planes are empty"
    differing = []
    for k in sorted(pi):
        si, wi, a0, a1 = pi[k]
        sf, wf, b0, b1 = pf[k]
        # FAITHFUL repeats the vertex two pieces share when it fails to join them: more waypoints, same point set
        ui = np.unique(fi.pos[int(fi.seg_off[a0]):int(fi.seg_off[a1])].view(np.uint32), axis=0)
        uf = np.unique(ff.pos[int(ff.seg_off[b0]):int(ff.seg_off[b1])].view(np.uint32), axis=0)
        assert np.array_equal(ui, uf), f"plane {k}: the two modes cut different points"
        # the segments of the plane as a set of point sequences, each in its lexicographically smaller direction: the raster
        # post-ops (RasterOrganization) re-orient and re-order the segments of a plane anyway
        def canon(f, s0, s1):
            segs = []
            for s in range(s0, s1):
                p = f.pos[int(f.seg_off[s]):int(f.seg_off[s + 1])].view(np.uint32)
                a, b = p.tobytes(), p[::-1].tobytes()
                segs.append(min(a, b))
            return sorted(segs)
        if canon(fi, a0, a1) != canon(ff, b0, b1):
            differing.append({"plane": k, "segments_intent": si, "segments_faithful": sf, "waypoints_intent": wi,
                              "waypoints_faithful": wf})
    return {
        "config": name, "what": desc, "triangles": int(tri.shape[0]), "lattice_planes": int(lat.num_planes),
        "non_empty_planes": len(pi), "planes_where_faithful_differs": len(differing),
        "fraction": len(differing) / max(1, len(pi)), "segments_intent": int(len(fi.seg_off) - 1),
        "segments_faithful": int(len(ff.seg_off) - 1), "differing_planes": differing[:32],
        "seconds": round(time.perf_counter() - t0, 1),
    }


def main():
    orc.build()
    table = configs()
    wanted = sys.argv[1:] or ["cfg1", "cfg1-default", "cfg1-0.05", "cfg2", "cfg2-hole", "cfg3", "cfg3-hole"]
    out = [compare(n, *table[n]) for n in wanted]
    print(json.dumps({"note": "oracle SLICE_FAITHFUL (vtkStripper + topological sort + stack join as the reference writes "
                              "them, restated from VTK 9.1 / Boost.Graph behaviour: parity unpinned) against SLICE_INTENT "
                              "(maximal chains; the GPU path's target), lattice-binned scan, before the raster post-ops",
                      "results": out}, indent=1))


if __name__ == "__main__":
    main()
