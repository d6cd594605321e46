"""Opt-in randomised parity campaign (GPU against the oracle, bit for bit) beyond the fixed seeds of
test_gpu_parity.py::test_fuzz_plans_match_oracle:

    NB2_FUZZ=small:1000:600,big:5000:60,topo:0:300 python -m pytest tests/test_fuzz_campaign.py -m gpu -q -s

Each item is kind:first_seed:count; kinds: small / big (the sheet generator of test_gpu_parity.py) topo (closed tubes and
tori, unwelded sheets, degenerate / duplicated triangles, fins) and legacy (PlaneSlicerLegacy with random point spacing,
minimum segment size and hole size on the sheet generator's meshes; poses bit for bit) and normals (NormalsFromMeshFaces on
the topology generator's meshes), mods (random chains of the device tool-path modifiers on random ragged tool paths, poses
bit for bit) and stl (random triangle soups with coincident, signed-zero and collapsing corners through the STL welder).
Skipped unless NB2_FUZZ is set (the default suites stay short)."""
import os

import numpy as np
import pytest

from test_gpu_parity import _fuzz_case
from util import assert_same_paths, oracle_pca_from_capi

pytestmark = pytest.mark.gpu


def _fuzz_case_topology(seed, orc):
    """Meshes the sheet generator never produces: closed tubes and tori (closed contours), partly unwelded sheets
    (cut points that merge on coordinates only), degenerate and duplicated triangles, fins (three triangles on an edge)."""
    from noether_b200 import meshes
    rng = np.random.default_rng(77000 + seed)
    kind = int(rng.integers(4))
    if kind == 0:      # tube: closed around, open at both ends
        n, m = int(rng.integers(8, 48)), int(rng.integers(6, 60))
        th = np.linspace(0, 2 * np.pi, n, endpoint=False)
        zs = np.linspace(0, float(rng.uniform(1, 10)), m)
        a, b, w = float(rng.uniform(0.3, 1.5)), float(rng.uniform(0.3, 1.5)), float(rng.uniform(0, 0.2))
        xyz = np.array([[a * np.cos(t) * (1 + w * np.sin(3 * z)), b * np.sin(t), z] for z in zs for t in th], np.float32)
        tri = np.array([t for j in range(m - 1) for i in range(n)
                        for t in ([j * n + i, j * n + (i + 1) % n, (j + 1) * n + i],
                                  [j * n + (i + 1) % n, (j + 1) * n + (i + 1) % n, (j + 1) * n + i])], np.uint32)
    elif kind == 1:    # torus: closed both ways
        n, m = int(rng.integers(8, 40)), int(rng.integers(8, 60))
        R, r = float(rng.uniform(1.0, 3.0)), float(rng.uniform(0.2, 0.8))
        xyz = np.array([[(R + r * np.cos(t)) * np.cos(p), (R + r * np.cos(t)) * np.sin(p), r * np.sin(t)]
                        for p in np.linspace(0, 2 * np.pi, m, endpoint=False)
                        for t in np.linspace(0, 2 * np.pi, n, endpoint=False)], np.float32)
        tri = np.array([t for j in range(m) for i in range(n)
                        for t in ([j * n + i, j * n + (i + 1) % n, ((j + 1) % m) * n + i],
                                  [j * n + (i + 1) % n, ((j + 1) % m) * n + (i + 1) % n, ((j + 1) % m) * n + i])], np.uint32)
    else:              # sheet, then damaged below
        nx, ny = int(rng.integers(6, 50)), int(rng.integers(6, 50))
        xyz, tri = meshes.wavy(nx, ny, lx=float(rng.uniform(0.5, 2.0)), ly=float(rng.uniform(0.3, 1.4)),
                               amp=float(rng.uniform(0.0, 0.08)), lam=float(rng.uniform(0.2, 1.5)))
    if kind >= 2 or rng.random() < 0.3:
        # unweld: a random subset of triangles gets private copies of its vertices
        sel = np.flatnonzero(rng.random(tri.shape[0]) < rng.uniform(0.05, 0.6))
        extra = xyz[tri[sel].reshape(-1)]
        tri = tri.copy()
        tri[sel] = (xyz.shape[0] + np.arange(3 * sel.size, dtype=np.uint32)).reshape(-1, 3)
        xyz = np.concatenate([xyz, extra])
    if kind == 3:
        t = [tri]
        k = max(1, tri.shape[0] // 20)
        pick = tri[rng.integers(0, tri.shape[0], k)]
        t.append(np.stack([pick[:, 0], pick[:, 0], pick[:, 1]], 1))   # a vertex twice
        t.append(pick[:, [1, 2, 0]])                                  # the same triangle again, rotated
        if rng.random() < 0.5:                                        # fins: a third triangle on an existing edge
            apex = xyz[pick[:, 0]] + np.float32(0.05) * rng.normal(size=(k, 3)).astype(np.float32)
            fin = np.stack([pick[:, 0], pick[:, 1], xyz.shape[0] + np.arange(k, dtype=np.uint32)], 1)
            xyz = np.concatenate([xyz, apex.astype(np.float32)])
            t.append(fin.astype(np.uint32))
        tri = np.concatenate(t).astype(np.uint32)
        tri = tri[rng.permutation(tri.shape[0])]
    if rng.random() < 0.6:
        T = np.eye(4)
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        w, x, y, z = q
        T[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
        T[:3, 3] = rng.uniform(-3, 3, size=3)
        xyz = meshes.rigid_transform(xyz, T)
    xyz = np.ascontiguousarray(xyz, np.float32)
    nrm = orc.normals_from_mesh_faces(xyz, tri)
    ext = xyz.max(0) - xyz.min(0)
    kw = dict(bidirectional=bool(rng.random() < 0.8), spacing=float(rng.uniform(0.02, 0.2) * np.linalg.norm(ext)))
    d = rng.normal(size=3)
    kw["direction"] = (d / np.linalg.norm(d)).tolist()
    if rng.random() < 0.5:
        kw["origin"] = (xyz.mean(0) + rng.uniform(-0.2, 0.2, size=3)).tolist()
    layout = ["PointNormal", "NormalsFirst", "Packed24"][int(rng.integers(3))]
    return meshes.pack_blob(xyz, nrm, tri, layout), xyz, nrm, tri, kw


def _campaign():
    spec = os.environ.get("NB2_FUZZ", "")
    for item in filter(None, spec.split(",")):
        kind, first, count = item.split(":")
        yield kind, int(first), int(count)


@pytest.mark.skipif(not os.environ.get("NB2_FUZZ"), reason="set NB2_FUZZ=kind:first_seed:count[,...] to run the campaign")
def test_fuzz_campaign(ctx, orc):
    import noether_b200 as nb
    from noether_b200 import _capi
    ran = refused = 0
    failures = []
    errors = {}
    for kind, first, count in _campaign():
        if kind in ("mods", "stl"):
            continue  # test_fuzz_modifiers_and_stl below
        big = kind
        for seed in range(first, first + count):
            blob, xyz, nrm, tri, kw = (_fuzz_case_topology(seed, orc) if kind in ("topo", "normals")
                                       else _fuzz_case(seed, orc, kind == "big"))
            if kind == "legacy" and "direction" not in kw and "rotation_offset" in kw:
                pass  # (PrincipalAxis with a rotation offset is fine for the legacy planner too)
            dg = (nb.FixedDirectionGenerator(kw["direction"]) if "direction" in kw
                  else nb.PrincipalAxisDirectionGenerator(kw.get("rotation_offset", 0.0), ctx))
            og = nb.FixedOriginGenerator(kw["origin"]) if "origin" in kw else nb.CentroidOriginGenerator(ctx)
            if kind == "normals":
                # NormalsFromMeshFaces on the topology generator's meshes (degenerate faces, isolated and duplicated
                # vertices, high valence at the fins): bit for bit against the face-ordered float sums
                from noether_b200 import meshes
                from util import bits
                ran += 1
                ctx.upload(meshes.pack_blob(xyz, None, tri, "PointXYZ"), force=True)
                got = ctx.normals_from_mesh_faces(xyz.shape[0])
                if not np.array_equal(bits(got), bits(nrm)):
                    failures.append(f"seed {seed} normals: {int((bits(got) != bits(nrm)).any(1).sum())} vertices differ")
                continue
            legacy = None
            if kind == "legacy":
                # PlaneSlicerLegacy: the three legacy modifiers run on the device on the slicer's result
                lrng = np.random.default_rng(31000 + seed)
                step = kw["spacing"] * float(lrng.uniform(0.15, 1.5))
                legacy = (step, float(lrng.choice([0.0, 1.0, 3.0, 8.0])) * step, float(lrng.choice([0.0, 0.5, 2.0, 6.0])) * step)
                planner = nb.PlaneSlicerLegacyRasterPlanner(dg, og, legacy[0], legacy[1], legacy[2], ctx)
            else:
                planner = nb.PlaneSlicerRasterPlanner(dg, og, ctx)
            planner.set_line_spacing(kw["spacing"])
            planner.generate_rasters_bidirectionally(kw["bidirectional"])
            ran += 1
            try:
                gpu = planner.plan(blob)
            except _capi.Nb2Error as e:
                allowed = (_capi.NB2_ERR_INVALID_ARGUMENT, _capi.NB2_ERR_NON_MANIFOLD) + (
                    (_capi.NB2_ERR_CLOSED_LOOP,) if kind == "topo" else ())
                if legacy and e.status == _capi.NB2_ERR_SEGMENT_TOO_SHORT:
                    # UniformSpacing refuses two-point segments (uniform_spacing_modifier.cpp:20-21): so must the oracle
                    import oracle
                    try:
                        orc.plan_plane_slicer_legacy(xyz, nrm, tri, kw["spacing"], legacy[0], legacy[1], legacy[2],
                                                     kw["bidirectional"], kw.get("direction"), kw.get("origin"),
                                                     rotation_offset=kw.get("rotation_offset", 0.0),
                                                     slice_mode=orc.SLICE_INTENT, frame=oracle_pca_from_capi(orc, ctx.pca()))
                        failures.append(f"seed {seed} legacy: the GPU refused a segment the oracle resamples")
                    except oracle.OracleError:
                        pass
                elif e.status not in allowed:
                    failures.append(f"seed {seed} big={big}: unexpected error {e}")
                errors[e.status] = errors.get(e.status, 0) + 1
                refused += 1
                continue
            if legacy:
                import oracle
                try:
                    paths, lat, _ = orc.plan_plane_slicer_legacy(
                        xyz, nrm, tri, kw["spacing"], legacy[0], legacy[1], legacy[2], kw["bidirectional"],
                        kw.get("direction"), kw.get("origin"), rotation_offset=kw.get("rotation_offset", 0.0),
                        slice_mode=orc.SLICE_INTENT, frame=oracle_pca_from_capi(orc, ctx.pca()))
                    ref = paths.flat()
                    assert gpu.legacy and gpu.n_paths == ref.num_paths, "paths"
                    np.testing.assert_array_equal(gpu.path_offsets.astype(np.uint64), ref.path_off)
                    np.testing.assert_array_equal(gpu.segment_offsets.astype(np.uint64), ref.seg_off)
                    np.testing.assert_array_equal(gpu.poses(), ref.pose)
                except oracle.OracleError as e:
                    failures.append(f"seed {seed} legacy {legacy}: oracle refuses ({e}) what the GPU planned")
                except AssertionError as e:
                    failures.append(f"seed {seed} legacy {legacy}: {str(e)[:300]}")
                finally:
                    gpu.free()
                continue
            try:
                paths, lat, _ = orc.plan_plane_slicer(xyz, nrm, tri, kw["spacing"], kw["bidirectional"],
                                                      direction=kw.get("direction"), origin=kw.get("origin"),
                                                      rotation_offset=kw.get("rotation_offset", 0.0),
                                                      slice_mode=orc.SLICE_INTENT, scan_mode=orc.SCAN_BINNED,
                                                      frame=oracle_pca_from_capi(orc, ctx.pca()))
                assert bytes(ctx.last_lattice()) == bytes(lat), "lattice"
                assert_same_paths(gpu, paths.flat(), f"seed {seed} {kw}")
            except AssertionError as e:
                failures.append(f"seed {seed} big={big}: {str(e)[:300]}")
            finally:
                gpu.free()
    print(f"fuzz campaign: {ran} cases, {refused} refused {errors}, {len(failures)} failures")
    assert not failures, "\n".join(failures[:20])
    if ran and all(kind in ("small", "big") for kind, _, _ in _campaign()):
        assert refused <= max(3, ran // 20)  # (topo: fins are refused; legacy: two-point segments, as the oracle does)


@pytest.mark.skipif(not os.environ.get("NB2_FUZZ"), reason="set NB2_FUZZ=kind:first_seed:count[,...] to run the campaign")
def test_fuzz_modifiers_and_stl(ctx, orc):
    import test_modifiers as tm
    import test_stl as ts
    from noether_b200 import _capi, meshes
    ran = refused = 0
    failures = []
    names = sorted(tm.MODIFIERS)
    for kind, first, count in _campaign():
        for seed in range(first, first + count):
            if kind == "mods":
                nested, chain = tm.random_case(seed)   # (a one-waypoint segment now and then: several modifiers refuse it)
                ran += 1
                try:
                    ref = tm.oracle_apply(nested, chain)
                    ref_ok = True
                except Exception:
                    ref_ok = False
                try:
                    po, so, poses, _ = tm.gpu_apply(ctx, nested, chain)
                    gpu_ok = True
                except _capi.Nb2Error:
                    gpu_ok = False
                if gpu_ok != ref_ok:
                    failures.append(f"mods seed {seed} {chain}: gpu {'ok' if gpu_ok else 'refuses'}, oracle {'ok' if ref_ok else 'refuses'}")
                elif not gpu_ok:
                    refused += 1
                else:
                    try:
                        tm.assert_matches_oracle((po, so, poses), ref, f"mods seed {seed} {chain}")
                    except AssertionError as e:
                        failures.append(str(e)[:300])
            elif kind == "stl":
                rng = np.random.default_rng(91000 + seed)
                nx, ny = int(rng.integers(2, 40)), int(rng.integers(2, 40))
                xyz, tri = meshes.wavy(nx, ny, amp=float(rng.uniform(0, 0.1)))
                xyz = xyz.copy()
                if rng.random() < 0.5:   # snap to a coarse lattice: many coincident vertices, collapsing facets, zeros
                    q = np.float32(rng.choice([0.05, 0.125, 0.25]))
                    xyz = (np.round(xyz / q) * q).astype(np.float32)
                if rng.random() < 0.5:   # signed zeros
                    xyz[rng.random(xyz.shape) < 0.2] *= np.float32(-0.0) if rng.random() < 0.5 else np.float32(1.0)
                    z = xyz == 0
                    xyz[z & (rng.random(xyz.shape) < 0.5)] = np.float32(-0.0)
                if rng.random() < 0.5:
                    xyz, tri = meshes.permute(xyz, tri, seed=int(rng.integers(1 << 30)))
                raw = ts.write_stl(xyz, tri, count=int(rng.integers(0, 2 * len(tri) + 2)), tail=b"\x07" * int(rng.integers(0, 48)))
                ran += 1
                try:
                    ts._check_against_oracle(ctx, orc, raw)
                except AssertionError as e:
                    failures.append(f"stl seed {seed}: {str(e)[:300]}")
    print(f"fuzz campaign (modifiers / stl): {ran} cases, {refused} refused by both, {len(failures)} failures")
    assert not failures, "\n".join(failures[:20])


@pytest.mark.skipif(not os.environ.get("NB2_FUZZ"), reason="set NB2_FUZZ=kind:first_seed:count[,...] to run the campaign")
def test_fuzz_mesh_modifiers(ctx, orc):
    """kinds `sub` (face subdivision: random triangle soups over few vertices, every cloud layout, both modifiers) and
    `clu` (Euclidean clustering: random clouds, snapped lattices, tolerances around the point spacing)"""
    import noether_b200 as nb
    import test_subdivide as tsub
    from noether_b200 import meshes
    ran = 0
    failures = []
    names = sorted(tsub.LAYOUTS)
    for kind, first, count in _campaign():
        for seed in range(first, first + count):
            rng = np.random.default_rng(77000 + seed)
            if kind == "sub":
                V, F = int(rng.integers(3, 200)), int(rng.integers(1, 600))
                xyz = (rng.normal(size=(V, 3)) * rng.choice([1e-3, 1.0, 50.0])).astype(np.float32)
                tri = rng.integers(0, V, size=(F, 3)).astype(np.uint32)
                if rng.random() < 0.3:   # a real surface instead of a soup
                    xyz, tri = meshes.wavy(int(rng.integers(2, 40)), int(rng.integers(2, 40)))
                layout = names[int(rng.integers(len(names)))]
                step, ox, on, oc = tsub.LAYOUTS[layout]
                cloud = tsub.make_cloud(xyz, layout, seed=seed)
                blob = tsub.gpu_blob(cloud, layout, tri)
                ran += 1
                try:
                    n = int(rng.integers(1, 4))
                    out = nb.FaceMidpointSubdivisionMeshModifier(n, ctx, off_rgba=oc).modify(blob)[0]
                    tsub.assert_same_mesh((out.data, out.tri), orc.face_midpoint_subdivision(cloud, step, ox, on, oc, tri, n),
                                          f"sub seed {seed} midpoint n={n} {layout}")
                    max_area = float(rng.choice(tsub.area_limits(xyz, tri)))
                    out = nb.FaceSubdivisionByAreaMeshModifier(max_area, ctx, off_rgba=oc).modify(blob)[0]
                    tsub.assert_same_mesh((out.data, out.tri), orc.face_subdivision_by_area(cloud, step, ox, on, oc, tri, max_area),
                                          f"sub seed {seed} by area {max_area} {layout}")
                except AssertionError as e:
                    failures.append(str(e)[:300])
            elif kind == "clu":
                V = int(rng.integers(1, 3000))
                scale = float(rng.choice([1e-3, 1.0, 250.0]))
                xyz = (rng.normal(size=(V, 3)) * scale).astype(np.float32)
                if rng.random() < 0.4:
                    xyz = (np.round(xyz / (0.2 * scale)) * (0.2 * scale)).astype(np.float32)
                tri = rng.integers(0, V, size=(int(rng.integers(0, 2 * V)), 3)).astype(np.uint32)
                blob = meshes.pack_blob(xyz, None, tri, "PointXYZ")
                tol = float(rng.choice([0.0, 0.05, 0.2, 0.2000001, 0.4, 1.0, 3.0])) * scale
                mn, mx = int(rng.choice([1, 1, 2, 5])), int(rng.choice([-1, -1, 50]))
                ran += 1
                try:
                    ref, labels = orc.euclidean_clustering(blob.data, 16, 0, tri, tol, mn, mx)
                    ctx.upload(blob, force=True)
                    assert np.array_equal(ctx.euclidean_cluster_labels(tol, V), labels), f"clu seed {seed}: labels differ"
                    got = nb.EuclideanClusteringMeshModifier(tol, mn, mx, ctx).modify(blob)
                    assert len(got) == len(ref), f"clu seed {seed}: {len(got)} meshes, oracle {len(ref)}"
                    for k, (g, r) in enumerate(zip(got, ref)):
                        assert (g is None) == r.empty, f"clu seed {seed}: mesh {k} emptiness"
                        if g is not None:
                            assert np.array_equal(g.tri, r.tri) and np.array_equal(np.asarray(g.data), r.cloud), f"clu seed {seed}: mesh {k}"
                except AssertionError as e:
                    failures.append(str(e)[:300])
    print(f"fuzz campaign (mesh modifiers): {ran} cases, {len(failures)} failures")
    assert not failures, "\n".join(failures[:20])
