"""Device timing of a tool-path modifier chain on a plan's result: per-stage CUDA-event times of
nb2_result_modify and the HBM roofline fraction of each modifier kernel (128 W in + 128 W' out bytes).

    python tools/time_modifiers.py [cfg2|cfg3] > profiles/r01_modifiers_<cfg>.json
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import noether_b200 as nb  # noqa: E402
from noether_b200 import meshes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
nx, ny, sp = {"cfg2": (800, 625, 0.005), "cfg3": (2500, 2000, 0.001)}[name]
peak = 6541.8
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
xyz, tri = meshes.wavy(nx, ny)
ctx = nb.Context(0)
ctx.upload(meshes.pack_blob(xyz, None, tri, "PointXYZ"), force=True)
nrm = ctx.normals_from_mesh_faces(xyz.shape[0])
pl = nb.PlaneSlicerRasterPlanner(nb.PrincipalAxisDirectionGenerator(0.0, ctx), nb.CentroidOriginGenerator(ctx), ctx)
pl.set_line_spacing(sp)
pl.emit_edges = False
tp = pl.plan(meshes.pack_blob(xyz, nrm, tri))
T = np.eye(4)
T[:3, 3] = [0.0, 0.0, 0.01]
chain = nb.CompoundModifier([nb.SnakeOrganizationModifier(), nb.OffsetModifier(T), nb.ToolDragOrientationToolPathModifier(0.1745, 0.025),
                             nb.UniformOrientationModifier(), nb.LinearApproachModifier([0.0, 0.0, 0.05], 5),
                             nb.LinearDepartureModifier([0.0, 0.0, 0.05], 5)], ctx)
best = None
for it in range(5):
    t0 = time.perf_counter()
    out = chain.modify(tp)
    wall = 1e3 * (time.perf_counter() - t0)
    stages = ctx.timings()
    if best is None or sum(stages.values()) < sum(best[1].values()):
        best = (wall, dict(stages), out.n_waypoints, out.n_segments)
    out.free()
wall, stages, w_out, s_out = best
W = tp.n_waypoints
kernels = {}
w_now = W
for k, ms in stages.items():
    if k in ("h2d", "ingest+moments", "validate", "h2d compact", "h2d poses", "d2h", "float views"):  # (the first three: last upload)
        continue
    if k == "expand poses":
        b = 24 * W + 128 * W
    else:
        w_next = w_now + (5 * s_out if k in ("linear approach", "linear departure") else 0)
        b = 128 * w_now + 128 * w_next
        w_now = w_next
    kernels[k] = {"ms": round(ms, 4), "algorithmic_bytes": b, "GB/s": round(b / ms / 1e6, 1), "frac_of_peak": round(b / ms / 1e6 / peak, 3)}
# the one sequential modifier: MovingAverageOrientationSmoothing (a recurrence along each segment, one thread per segment)
smooth = nb.CompoundModifier([nb.MovingAverageOrientationSmoothingModifier(3)], ctx)
sm_ms = None
for it in range(2):
    out = smooth.modify(tp)
    ms = ctx.timings().get("moving average orientation smoothing")
    sm_ms = ms if sm_ms is None else min(sm_ms, ms)
    out.free()
kernels["moving average orientation smoothing (window 3)"] = {
    "ms": round(sm_ms, 3), "note": "%d segments = %d threads walking %d poses each in order (4 x 4 Jacobi SVD per pose)"
    % (tp.n_segments, tp.n_segments, W // max(tp.n_segments, 1))}
print(json.dumps({"workload": name, "waypoints_in": W, "waypoints_out": w_out, "segments": s_out, "wall_ms": round(wall, 2),
                  "stage_ms": {k: round(v, 4) for k, v in stages.items()}, "kernels": kernels, "hbm_peak_GB/s": peak}))
