"""Face subdivision on the device: device time of nb2_face_midpoint_subdivision (n = 1) and
nb2_face_subdivision_by_area on a bench workload (PointNormal cloud, resident in HBM; CUDA events on the library's stream
around the whole call, which includes its read-backs of the new sizes and the re-ingest of the subdivided mesh), next to
the oracle's sequential walk on one host core for a bounded sample, and the result checked against the oracle on that
sample.

    python tools/time_subdivide.py [cfg2|cfg3] > profiles/r01_subdivide_<cfg>.json
    ncu ... python tools/time_subdivide.py cfg3 --profile     (launch list / --set full capture)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import noether_b200 as nb  # noqa: E402
from noether_b200 import meshes  # noqa: E402


def peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6541.8, "fallback"


def main():
    import oracle
    oracle.build()
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    profile = "--profile" in sys.argv   # under ncu: one pass of each modifier, no oracle leg
    nx, ny = {"cfg2": (800, 625), "cfg3": (2500, 2000)}[name]
    xyz, tri = meshes.wavy(nx, ny)
    ctx = nb.Context(0)
    ctx.upload(meshes.pack_blob(xyz, None, tri, "PointXYZ"), force=True)
    nrm = ctx.normals_from_mesh_faces(len(xyz))
    blob = meshes.pack_blob(xyz, nrm, tri, "PointNormal")
    step = blob.point_step
    V, F = len(xyz), len(tri)
    peak, peak_src = peak_gbs()
    out = {"workload": name, "vertices": V, "triangles": F, "point_step": step, "peak_GBs": peak, "peak_source": peak_src}
    cell_area = 0.5 * (1.0 / nx) * (0.8 / ny)

    def timed(fn, reps=1 if profile else 3):
        best = None
        for _ in range(reps):
            ctx.upload(blob, force=True)
            l0 = ctx.kernel_launches()
            ctx.timer_start()
            info = fn()
            ms = ctx.timer_stop()
            launches = ctx.kernel_launches() - l0
            if best is None or ms < best[0]:
                best = (ms, info, launches)
        return best

    ms, info, launches = timed(lambda: ctx.face_midpoint_subdivision(1))
    # compulsory traffic: triangles in, 4x triangles out, the cloud copied and the new records written
    alg = 12 * F + 48 * F + step * V + step * info.n_points
    out["midpoint_n1"] = {"device_ms": round(ms, 4), "launches": launches, "points_out": info.n_points, "triangles_out": info.n_triangles,
                          "algorithmic_MB": round(alg / 1e6, 1), "achieved_GBs": round(alg / ms / 1e6, 1),
                          "frac_of_peak": round(alg / ms / 1e6 / peak, 4), "triangles_in_per_s": round(F / ms * 1e3)}
    limit = cell_area * 1.02 / 3   # flat-ish cells split once, the steeper (larger) ones twice
    ms2, info2, launches2 = timed(lambda: ctx.face_subdivision_by_area(limit))
    alg2 = 12 * F + 12 * info2.n_triangles + step * V + step * info2.n_points
    out["by_area"] = {"max_area": limit, "device_ms": round(ms2, 4), "launches": launches2, "points_out": info2.n_points,
                      "triangles_out": info2.n_triangles, "algorithmic_MB": round(alg2 / 1e6, 1),
                      "achieved_GBs": round(alg2 / ms2 / 1e6, 1), "frac_of_peak": round(alg2 / ms2 / 1e6 / peak, 4),
                      "triangles_in_per_s": round(F / ms2 * 1e3)}
    if profile:
        print(json.dumps(out))
        return
    # the oracle (sequential depth-first walk, std::map per edge) on the leading rows of the same mesh: one host core
    rows = max(1, min(ny, 1_000_000 // (2 * nx)))
    xs, ts = meshes.wavy(nx, rows, ly=0.8 * rows / ny)
    sub = meshes.pack_blob(xs, np.zeros_like(xs), ts, "PointNormal")
    t0 = time.perf_counter()
    ref_c, ref_t = oracle.face_midpoint_subdivision(sub.data, step, 0, 16, -1, ts, 1)
    cpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref2_c, ref2_t = oracle.face_subdivision_by_area(sub.data, step, 0, 16, -1, ts, limit)
    cpu2 = time.perf_counter() - t0
    ctx.upload(sub, force=True)
    i1 = ctx.face_midpoint_subdivision(1)
    same1 = bool(np.array_equal(ctx.download_cloud(i1.n_points), ref_c) and np.array_equal(ctx.download_triangles(i1.n_triangles), ref_t))
    ctx.upload(sub, force=True)
    i2 = ctx.face_subdivision_by_area(limit)
    same2 = bool(np.array_equal(ctx.download_cloud(i2.n_points), ref2_c) and np.array_equal(ctx.download_triangles(i2.n_triangles), ref2_t))
    out["oracle_1_core"] = {"sample_triangles": len(ts), "midpoint_n1_s": round(cpu, 3), "midpoint_triangles_per_s": round(len(ts) / cpu),
                            "by_area_s": round(cpu2, 3), "by_area_triangles_per_s": round(len(ts) / cpu2),
                            "gpu_equals_oracle_on_sample": {"midpoint": same1, "by_area": same2}}
    print(json.dumps(out))
    assert same1 and same2


if __name__ == "__main__":
    main()
