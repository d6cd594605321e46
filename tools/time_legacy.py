import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import noether_b200 as nb
from noether_b200 import meshes
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
nx, ny, sp = {'cfg2': (800, 625, 0.005), 'cfg3': (2500, 2000, 0.001)}[name]
xyz, tri = meshes.wavy(nx, ny)
ctx = nb.Context(0)
ctx.upload(meshes.pack_blob(xyz, None, tri, "PointXYZ"), force=True)
nrm = ctx.normals_from_mesh_faces(xyz.shape[0])
blob = meshes.pack_blob(xyz, nrm, tri)
for legacy in (False, True):
    if legacy:
        pl = nb.PlaneSlicerLegacyRasterPlanner(nb.PrincipalAxisDirectionGenerator(0.0, ctx), nb.CentroidOriginGenerator(ctx), sp * 2.5, sp * 10, 0.0, ctx)
    else:
        pl = nb.PlaneSlicerRasterPlanner(nb.PrincipalAxisDirectionGenerator(0.0, ctx), nb.CentroidOriginGenerator(ctx), ctx)
    pl.set_line_spacing(sp)
    pl.emit_edges = False
    for it in range(3):
        ctx._mesh_key = None
        t0 = time.perf_counter()
        r = pl.plan(blob)
        t1 = time.perf_counter()
        p = r.poses()
        t2 = time.perf_counter()
        print(name, 'legacy' if legacy else 'plain', 'plan %.1f ms' % (1e3 * (t1 - t0)), 'poses %.1f ms' % (1e3 * (t2 - t1)), r.n_paths, r.n_segments, r.n_waypoints)
        r.free()
print({k: round(v, 4) for k, v in ctx.timings().items()})
