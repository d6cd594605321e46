"""Mesh files decoded on the device: stage times of nb2_mesh_upload_ply / nb2_mesh_upload_stl for a bench
workload written as a binary PLY (vertices with normals + 13-byte faces) and as a binary STL (50-byte facets, corners
welded on the device), next to the oracle's sequential reader (vtkPLYReader / vtkSTLReader + vtk2mesh restated) on one host
core for a bounded sample.

    python tools/time_ingest.py [cfg2|cfg3] > profiles/r01_ingest_<cfg>.json
"""
import json
import os
import struct
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import noether_b200 as nb  # noqa: E402
from noether_b200 import meshes  # noqa: E402


def ply_image(xyz, nrm, tri):
    head = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
            "property float nx\nproperty float ny\nproperty float nz\nelement face %d\n"
            "property list uchar int vertex_indices\nend_header\n" % (len(xyz), len(tri))).encode()
    v = np.concatenate([xyz, nrm], axis=1).astype("<f4")
    f = np.zeros(len(tri), np.dtype([("n", "u1"), ("i", "<i4", 3)]))
    f["n"] = 3
    f["i"] = tri
    return head + v.tobytes() + f.tobytes()


def stl_image(xyz, tri):
    p = xyz[tri]
    rec = np.zeros((len(tri), 50), np.uint8)
    rec[:, 12:48] = np.ascontiguousarray(p, np.float32).view(np.uint8).reshape(-1, 36)
    return b"noether_b200".ljust(80, b"\0") + struct.pack("<I", len(tri)) + rec.tobytes()


def main():
    import oracle
    oracle.build()
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    nx, ny = {"cfg2": (800, 625), "cfg3": (2500, 2000)}[name]
    xyz, tri = meshes.wavy(nx, ny)
    ctx = nb.Context(0)
    ctx.upload(meshes.pack_blob(xyz, None, tri, "PointXYZ"), force=True)
    nrm = ctx.normals_from_mesh_faces(len(xyz))
    out = {"workload": name, "vertices": len(xyz), "triangles": len(tri)}
    for kind, img in (("ply", ply_image(xyz, nrm, tri)), ("stl", stl_image(xyz, tri))):
        raw = np.frombuffer(img, np.uint8)
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            info = ctx.upload_ply(raw) if kind == "ply" else ctx.upload_stl(raw)
            wall = 1e3 * (time.perf_counter() - t0)
            st = {k: round(v, 4) for k, v in ctx.timings().items() if k in ("h2d", "decode", "weld", "ingest+moments")}
            if best is None or wall < best[0]:
                best = (wall, st)
        assert (info.n_points, info.n_triangles) == (len(xyz), len(tri))
        # the sequential reader on one host core, on the leading part of the same file (bounded: ~1 M faces)
        if kind == "stl":
            n_s = min(len(tri), 1_000_000)
            sample = img[:84 + 50 * n_s]
            t0 = time.perf_counter()
            oracle.read_stl(sample)
            cpu = time.perf_counter() - t0
        else:
            n_s = len(tri)
            t0 = time.perf_counter()
            oracle.read_ply(img)
            cpu = time.perf_counter() - t0
        out[kind] = {"file_MB": round(len(img) / 1e6, 1), "wall_ms_pageable_host_image": round(best[0], 2), "stage_ms": best[1],
                     "device_ms_without_h2d": round(sum(v for k, v in best[1].items() if k != "h2d"), 4),
                     "oracle_reader_1_core": {"faces": n_s, "seconds": round(cpu, 3),
                                              "extrapolated_full_file_s": round(cpu * len(tri) / n_s, 3)}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
