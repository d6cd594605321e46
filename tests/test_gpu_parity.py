"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Every test needs a B200: -m gpu."""
import ctypes as C

import numpy as np
import pytest

from util import assert_same_paths, bits, capi_lattice_from_oracle, oracle_pca_from_capi

pytestmark = pytest.mark.gpu


def _wavy_mesh(nx, ny, layout="PointNormal", hole=False, **kw):
    import oracle
    from noether_b200 import meshes
    xyz, tri = meshes.wavy(nx, ny, **kw)
    if hole:
        xyz, tri = meshes.cut_hole(xyz, tri)
    nrm = oracle.normals_from_mesh_faces(xyz, tri)
    return meshes.pack_blob(xyz, nrm, tri, layout), xyz, nrm, tri


# ---------------------------------------------------------------------------------------------------------------------
# K0: ingest
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("layout", ["PointNormal", "NormalsFirst", "Packed24"])
def test_ingest_roundtrip(ctx, orc, layout):
    blob, xyz, nrm, tri = _wavy_mesh(37, 23, layout)
    ctx.upload(blob, force=True)
    gx, gn = ctx.download(blob.n_vertices)
    np.testing.assert_array_equal(bits(gx), bits(xyz))
    np.testing.assert_array_equal(bits(gn), bits(nrm))


def test_ingest_rejects_bad_input(ctx):
    from noether_b200 import _capi, meshes
    xyz, tri = meshes.wavy(4, 4)
    nrm = np.tile(np.array([0, 0, 1], np.float32), (xyz.shape[0], 1))
    bad = xyz.copy()
    bad[3, 1] = np.nan
    with pytest.raises(_capi.Nb2Error) as e:
        ctx.upload(meshes.pack_blob(bad, nrm, tri), force=True)
    assert e.value.status == _capi.NB2_ERR_NON_FINITE
    t2 = tri.copy()
    t2[5, 2] = xyz.shape[0] + 7
    with pytest.raises(_capi.Nb2Error) as e:
        ctx.upload(meshes.pack_blob(xyz, nrm, t2), force=True)
    assert e.value.status == _capi.NB2_ERR_INVALID_ARGUMENT


def test_device_resident_triangles_are_checked_by_the_normals_pass(ctx, orc):
    """Triangles adopted in place from device memory are not validated at upload; the normals kernels skip the
    offending faces and the call reports the largest out-of-range index."""
    import torch
    from noether_b200 import _capi, meshes
    xyz, tri = meshes.wavy(40, 30)
    blob = meshes.pack_blob(xyz, None, tri, "PointXYZ")
    dev = torch.device("cuda", 0)
    d_cloud = torch.from_numpy(blob.data).to(dev)
    bad = tri.copy()
    bad[17, 1] = xyz.shape[0] + 5
    bad[400, 0] = xyz.shape[0] + 2
    d_bad = torch.from_numpy(bad.reshape(-1).view(np.int32)).to(dev)
    d_ok = torch.from_numpy(tri.reshape(-1).view(np.int32)).to(dev)
    torch.cuda.synchronize()
    V, F = xyz.shape[0], tri.shape[0]
    ctx.upload_raw(d_cloud.data_ptr(), V, blob.point_step, blob.off_xyz, blob.off_normal, d_bad.data_ptr(), F)
    with pytest.raises(_capi.Nb2Error) as e:
        ctx.normals_from_mesh_faces(V)
    assert e.value.status == _capi.NB2_ERR_INVALID_ARGUMENT and str(V + 5) in str(e.value)
    # the same buffers with valid indices: bit-exact normals
    ctx.upload_raw(d_cloud.data_ptr(), V, blob.point_step, blob.off_xyz, blob.off_normal, d_ok.data_ptr(), F)
    nrm = ctx.normals_from_mesh_faces(V)
    assert np.array_equal(bits(nrm), bits(orc.normals_from_mesh_faces(xyz, tri)))


# ---------------------------------------------------------------------------------------------------------------------
# K1: PCA frame, generators
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(40, 25), (200, 160), (313, 97)])
def test_pca_matches_fp64_oracle(ctx, orc, shape):
    blob, xyz, nrm, tri = _wavy_mesh(*shape)
    # move it away from the origin and rotate so no axis is special
    T = orc.fixture_random_transform(3.0, np.pi)
    from noether_b200 import meshes
    xyz = meshes.rigid_transform(xyz, T)
    blob = meshes.pack_blob(xyz, nrm, tri)
    ctx.upload(blob, force=True)
    g = ctx.pca()
    o = orc.pca(xyz, orc.MOMENTS_F64)
    # fp64 tree sums differ from the oracle's sequential long-double sums by ~1e-16 relative: after the float cast the
    # frame is identical
    np.testing.assert_array_equal(bits(np.array(g.mean)), bits(o.mean))
    np.testing.assert_array_equal(bits(np.array(g.cov)), bits(o.cov.T.ravel()))
    np.testing.assert_array_equal(bits(np.array(g.evecs)), bits(o.evecs.T.ravel()))
    np.testing.assert_array_equal(bits(np.array(g.scales)), bits(o.scales))
    # and it agrees with the reference-faithful float32 sequential accumulation within float noise
    f = orc.pca(xyz, orc.MOMENTS_F32_SEQUENTIAL)
    diag = np.linalg.norm(xyz.max(0) - xyz.min(0))
    assert np.abs(np.array(g.mean) - f.mean).max() < 1e-5 * diag
    for i in range(3):
        c = abs(float(np.dot(np.array(g.evecs).reshape(3, 3)[i], f.evecs[:, i])))
        assert c > 1 - 1e-6
    # generators
    np.testing.assert_array_equal(ctx.centroid_origin(), orc.centroid(xyz, orc.MOMENTS_F64))
    for rot in (0.0, 0.3, -1.2):
        np.testing.assert_array_equal(ctx.principal_axis_direction(rot), orc.principal_axis_direction(o, rot))


def test_lattice_matches_oracle(ctx, orc, ply_mesh):
    from noether_b200 import make_lattice
    ctx.upload(ply_mesh, force=True)
    g = ctx.pca()
    o = oracle_pca_from_capi(orc, g)
    for spacing, bidir, d, org in [(0.1, True, [1, 0, 0], [0, 0, 0]), (1.05, True, [1, 0, 0], [0, 0, 0]),
                                   (0.37, False, [0.3, 1, 0.1], [5, 5, 0]), (0.05, False, [1, 0, 0], [0, 20, 0])]:
        a = make_lattice(g, d, org, spacing, bidir)
        b = orc.make_lattice(o, d, org, spacing, bidir)
        assert bytes(a) == bytes(b)


# ---------------------------------------------------------------------------------------------------------------------
# K2: vertex normals
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("permuted", [False, True])
def test_normals_bit_exact(ctx, orc, permuted):
    from noether_b200 import meshes
    xyz, tri = meshes.wavy(120, 90)
    if permuted:
        xyz, tri = meshes.permute(xyz, tri, seed=7)
    blob = meshes.pack_blob(xyz, None, tri, "PointXYZ")
    ctx.upload(blob, force=True)
    g = ctx.normals_from_mesh_faces(blob.n_vertices)
    o = orc.normals_from_mesh_faces(xyz, tri)
    np.testing.assert_array_equal(bits(g), bits(o))


def test_normals_degenerate_and_isolated(ctx, orc):
    from noether_b200 import meshes
    xyz, tri = meshes.wavy(6, 5)
    xyz = np.concatenate([xyz, [[9, 9, 9]]]).astype(np.float32)       # isolated vertex -> zero normal
    tri = np.concatenate([tri, [[0, 0, 1]], [[2, 3, 3]]]).astype(np.uint32)  # degenerate faces add zero
    # a high-valence fan (> 16 faces at vertex 0)
    fan = np.array([[0, 8 + i, 9 + i] for i in range(20)], np.uint32)
    tri = np.concatenate([tri, fan]).astype(np.uint32)
    blob = meshes.pack_blob(xyz, None, tri, "PointXYZ")
    ctx.upload(blob, force=True)
    g = ctx.normals_from_mesh_faces(blob.n_vertices)
    o = orc.normals_from_mesh_faces(xyz, tri)
    np.testing.assert_array_equal(bits(g), bits(o))
    assert np.all(g[-1] == 0)


def test_normals_modifier_field_contract(ctx, ply_mesh):
    # mesh_modifier_utest.cpp:154-183: every input field is still present; normals come first (concatenateFields)
    import noether_b200 as nb
    out = nb.NormalsFromMeshFacesMeshModifier(ctx).modify(ply_mesh)
    assert len(out) == 1
    m = out[0]
    assert m.point_step == 32 + ply_mesh.point_step and m.off_normal == 0 and m.off_xyz == 32 + ply_mesh.off_xyz
    np.testing.assert_array_equal(m.xyz(), ply_mesh.xyz())
    assert np.allclose(np.linalg.norm(m.normals(), axis=1), 1.0, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------------------
# K3-K6: slicing with an injected lattice -> bit-exact against the oracle
# ---------------------------------------------------------------------------------------------------------------------
def _slice_both(ctx, orc, blob, xyz, nrm, tri, spacing, direction, origin, bidirectional=True, mode=None):
    ctx.upload(blob, force=True)
    p = orc.pca(xyz, orc.MOMENTS_F64)
    lat = orc.make_lattice(p, direction, origin, spacing, bidirectional)
    ref = orc.slice_mesh(xyz, nrm, tri, lat, orc.SLICE_INTENT if mode is None else mode, orc.SCAN_BINNED).flat()
    gpu = ctx.slice(capi_lattice_from_oracle(lat))
    return gpu, ref, lat


def test_slice_reference_fixture(ctx, orc, ply_mesh):
    xyz, nrm, tri = ply_mesh.xyz(), ply_mesh.normals(), ply_mesh.tri
    for spacing in (1.05, 0.1, 0.05):
        gpu, ref, _ = _slice_both(ctx, orc, ply_mesh, xyz, nrm, tri, spacing, [1, 0, 0], [0, 0, 0])
        assert_same_paths(gpu, ref, f"ply spacing {spacing}")
    # the reference's own expectation (tool_path_planner_utest.cpp:151-167)
    gpu, ref, _ = _slice_both(ctx, orc, ply_mesh, xyz, nrm, tri, 1.05, [1, 0, 0], [0, 0, 0])
    assert gpu.n_paths == 10
    assert list(gpu.segments_per_path()[4:7]) == [2, 2, 2]


@pytest.mark.parametrize("nx,ny,spacing,hole", [(40, 32, 0.02, False), (64, 50, 0.013, True), (200, 160, 0.005, True),
                                                (96, 77, 0.0371, False)])
def test_slice_wavy(ctx, orc, nx, ny, spacing, hole):
    blob, xyz, nrm, tri = _wavy_mesh(nx, ny, hole=hole)
    p = orc.pca(xyz, orc.MOMENTS_F64)
    d = orc.principal_axis_direction(p, 0.0)
    o = orc.centroid(xyz, orc.MOMENTS_F64)
    gpu, ref, _ = _slice_both(ctx, orc, blob, xyz, nrm, tri, spacing, d, o)
    assert ref.num_paths > 5
    assert_same_paths(gpu, ref, "wavy")


def test_slice_vertex_hits_and_in_plane_edges(ctx, orc):
    # flat grid, planes exactly through vertex rows (snap-to-vertex merges, zero-length lines, in-plane edges)
    from noether_b200 import meshes
    xyz, tri = meshes.wavy(16, 16, lx=1.0, ly=1.0, amp=0.0)
    nrm = np.tile(np.array([0, 0, 1], np.float32), (xyz.shape[0], 1))
    blob = meshes.pack_blob(xyz, nrm, tri)
    for spacing in (0.0625, 0.125, 0.03125):  # multiples / fractions of the 1/16 cell
        gpu, ref, lat = _slice_both(ctx, orc, blob, xyz, nrm, tri, spacing, [1, 0, 0], [0, 0.5, 0])
        assert ref.num_paths > 3
        assert_same_paths(gpu, ref, f"grid-aligned spacing {spacing}")
    # diagonal direction: planes through the diagonal vertices of every cell
    gpu, ref, lat = _slice_both(ctx, orc, blob, xyz, nrm, tri, 0.0625 / np.sqrt(2), [1, 1, 0], [0.5, 0.5, 0])
    assert_same_paths(gpu, ref, "diagonal")


def test_slice_permutation_invariance(ctx, orc):
    # shuffled vertex / face order: same polylines after mapping vertex ids back
    from noether_b200 import meshes
    from util import canonical_segments
    blob, xyz, nrm, tri = _wavy_mesh(48, 40, hole=True)
    gpu0, ref0, lat = _slice_both(ctx, orc, blob, xyz, nrm, tri, 0.021, [0.1, 1, 0], [0.5, 0.4, 0])
    xyz2, tri2, nrm2 = meshes.permute(xyz, tri, seed=3, normals=nrm)
    blob2 = meshes.pack_blob(xyz2, nrm2, tri2)
    ctx.upload(blob2, force=True)
    gpu1 = ctx.slice(capi_lattice_from_oracle(lat))
    ref1 = orc.slice_mesh(xyz2, nrm2, tri2, lat, orc.SLICE_INTENT, orc.SCAN_BINNED).flat()
    assert_same_paths(gpu1, ref1, "permuted")
    # positions are the same set per segment regardless of the permutation
    assert gpu1.n_waypoints == gpu0.n_waypoints and gpu1.n_segments == gpu0.n_segments
    np.testing.assert_array_equal(np.sort(bits(gpu1.positions).view(np.uint32).ravel()),
                                  np.sort(bits(gpu0.positions).ravel()))


def test_slice_plane_slabs_concatenate(ctx, orc):
    blob, xyz, nrm, tri = _wavy_mesh(80, 64, hole=True)
    gpu, ref, lat = _slice_both(ctx, orc, blob, xyz, nrm, tri, 0.011, [0, 1, 0], [0.5, 0.4, 0])
    P = int(lat.num_planes)
    cl = capi_lattice_from_oracle(lat)
    pos, lo, planes = [], [], []
    for r in range(3):
        part = ctx.slice(cl, P * r // 3, P * (r + 1) // 3)
        pos.append(part.positions.copy())
        lo.append(part.edge_lo.copy())
        planes.append(part.path_plane.copy())
    np.testing.assert_array_equal(np.concatenate(pos), gpu.positions)
    np.testing.assert_array_equal(np.concatenate(lo), gpu.edge_lo)
    np.testing.assert_array_equal(np.concatenate(planes), gpu.path_plane)


@pytest.mark.parametrize("ny,longest", [(2600, 4200), (9000, 17000)])
def test_large_plane_uses_global_scratch(ctx, orc, ny, longest):
    # planes with more hits than fit the shared-memory work area (about 4600): 5200 hits take the linker's fast path with
    # its work area in global memory (16-bit indices), 18000 hits the general path with 32-bit indices
    blob, xyz, nrm, tri = _wavy_mesh(8, ny, lx=0.05, ly=2.0)
    gpu, ref, lat = _slice_both(ctx, orc, blob, xyz, nrm, tri, 0.0125, [0, 1, 0], [0.0281, 1.0, 0])
    assert np.diff(ref.seg_off.astype(np.int64)).max() > longest
    assert_same_paths(gpu, ref, "large plane")


# ---------------------------------------------------------------------------------------------------------------------
# whole plan: RasterPlanner::plan
# ---------------------------------------------------------------------------------------------------------------------
def test_plan_end_to_end_matches_oracle(ctx, orc):
    import noether_b200 as nb
    blob, xyz, nrm, tri = _wavy_mesh(100, 80, hole=True)
    planner = nb.PlaneSlicerRasterPlanner(nb.PrincipalAxisDirectionGenerator(0.0, ctx), nb.CentroidOriginGenerator(ctx), ctx)
    planner.set_line_spacing(0.0123)
    planner.generate_rasters_bidirectionally(True)
    gpu = planner.plan(blob)
    # 1. the frame (15 floats) agrees with the oracle's own accumulation within float noise ...
    g = ctx.pca()
    o64 = orc.pca(xyz, orc.MOMENTS_F64)
    diag = float(np.linalg.norm(xyz.max(0) - xyz.min(0)))
    assert np.abs(np.array(g.mean) - o64.mean).max() <= 1e-7 * diag
    assert np.abs(np.array(g.scales) - o64.scales).max() <= 1e-6 * diag
    for i in range(3):
        assert abs(float(np.dot(np.array(g.evecs).reshape(3, 3)[i], o64.evecs[:, i]))) > 1 - 1e-9
    # 2. ... and given those 15 floats everything downstream is bit-exact: lattice, connectivity, positions, normals,
    #    raster organisation, poses.  (The mean of this mesh's z is ~1e-18: a sum that cancels to nothing has no
    #    meaningful last bits, and on this mesh a lattice plane lies within 1e-20 of a vertex column, so the frame is
    #    handed over rather than re-accumulated.)
    paths, lat, p = orc.plan_plane_slicer(xyz, nrm, tri, 0.0123, True, slice_mode=orc.SLICE_INTENT,
                                          scan_mode=orc.SCAN_BINNED, frame=oracle_pca_from_capi(orc, g))
    assert bytes(ctx.last_lattice()) == bytes(lat)
    assert gpu.post_ops_applied
    assert_same_paths(gpu, paths.flat(), "plan")


def test_plan_reference_tests_through_factory(ctx, orc, ply_mesh):
    """The reference's two planner tests, written the way tool_path_planner_utest.cpp:74-177 writes them."""
    import noether_b200 as nb
    from noether_b200 import meshes
    factory = nb.Factory(ctx)
    base = {"name": "PlaneSlicer", "direction_generator": {"name": "FixedDirection", "direction": {"x": 1.0, "y": 0.0, "z": 0.0}},
            "origin_generator": {"name": "Centroid"}, "line_spacing": 0.1, "point_spacing": 0.05, "min_segment_size": 0.01,
            "min_hole_size": 0.05, "bidirectional": True, "search_radius": 0.01}
    n_lines = 10

    # FlatSquareMesh
    dim = 1.0
    T = orc.fixture_random_transform(dim * 5.0, np.pi)
    direction = T[:3, :3] @ np.array([1.0, 0, 0])
    xyz, nrm, tri = orc.fixture_plane_mesh(dim, dim, T)
    mesh = meshes.pack_blob(xyz, nrm, tri)
    cfg = dict(base, line_spacing=0.95 * dim / n_lines)
    cfg["direction_generator"] = {"name": "FixedDirection", "direction": list(direction)}
    cfg["origin_generator"] = {"name": "FixedOrigin", "origin": list(T[:3, 3])}
    tool_paths = factory.create_tool_path_planner(cfg).plan(mesh).nested()
    assert len(tool_paths) == n_lines + 1
    for path in tool_paths:
        assert len(path) == 1
        d = path[-1][-1][:3, 3] - path[0][0][:3, 3]
        d /= np.linalg.norm(d)
        assert abs(d @ direction) > np.cos(np.radians(1.0))

    # SemiPlanarMeshFile
    dim = 10.5
    cfg = dict(base, line_spacing=dim / n_lines)
    cfg["origin_generator"] = {"name": "FixedOrigin", "origin": [0.0, 0.0, 0.0]}
    tool_paths = factory.create_tool_path_planner(cfg).plan(ply_mesh).nested()
    assert len(tool_paths) == n_lines
    for i, path in enumerate(tool_paths):
        assert len(path) >= 1 and len(path[0]) >= 2
        if 3 < i < 7:
            assert len(path) == 2 and len(path[-1]) >= 2
        d = path[-1][-1][:3, 3] - path[0][0][:3, 3]
        d /= np.linalg.norm(d)
        assert abs(d @ np.array([1.0, 0, 0])) > np.cos(np.radians(10.0))


def test_plan_default_yaml_on_reference_mesh(ctx, orc, ply_mesh):
    # cfg1: the reference's default YAML (line_spacing 0.1, FixedDirection x, Centroid, bidirectional)
    import noether_b200 as nb
    cfg = {"name": "PlaneSlicer", "direction_generator": {"name": "FixedDirection", "direction": [1.0, 0.0, 0.0]},
           "origin_generator": {"name": "Centroid"}, "line_spacing": 0.1, "bidirectional": True}
    gpu = nb.Factory(ctx).create_tool_path_planner(cfg).plan(ply_mesh)
    xyz, nrm, tri = ply_mesh.xyz(), ply_mesh.normals(), ply_mesh.tri
    paths, lat, _ = orc.plan_plane_slicer(xyz, nrm, tri, 0.1, True, direction=[1, 0, 0], moments_mode=orc.MOMENTS_F64,
                                          slice_mode=orc.SLICE_INTENT)
    assert_same_paths(gpu, paths.flat(), "cfg1")
    assert gpu.n_paths > 100


def test_plan_unidirectional_and_empty(ctx, orc):
    import noether_b200 as nb
    blob, xyz, nrm, tri = _wavy_mesh(30, 24)
    # origin beyond the mesh along the cut normal: unidirectional plans nothing (:569-570)
    for origin, expect_empty in (([5.0, 0.4, 0], None), ([-5.0, 0.4, 0], None)):
        pl = nb.PlaneSlicerRasterPlanner(nb.FixedDirectionGenerator([0, 1, 0]), nb.FixedOriginGenerator(origin), ctx)
        pl.set_line_spacing(0.05)
        pl.generate_rasters_bidirectionally(False)
        gpu = pl.plan(blob)
        paths, lat, _ = orc.plan_plane_slicer(xyz, nrm, tri, 0.05, False, direction=[0, 1, 0], origin=origin,
                                              slice_mode=orc.SLICE_INTENT, frame=oracle_pca_from_capi(orc, ctx.pca()))
        assert_same_paths(gpu, paths.flat(), f"unidirectional {origin}")
    # exactly one of the two origins yields no cuts at all
    # (which one depends on the sign of the PCA normal, which the oracle shares)


def test_plan_errors(ctx, orc):
    import noether_b200 as nb
    from noether_b200 import _capi, meshes
    xyz, tri = meshes.wavy(10, 8)
    no_normals = meshes.pack_blob(xyz, None, tri, "PointXYZ")
    pl = nb.PlaneSlicerRasterPlanner(nb.FixedDirectionGenerator([1, 0, 0]), nb.FixedOriginGenerator(), ctx)
    pl.set_line_spacing(0.05)
    with pytest.raises(_capi.Nb2Error) as e:
        pl.plan(no_normals)
    assert e.value.status == _capi.NB2_ERR_NO_NORMALS and "does not have vertex normals" in str(e.value)
    # closed loops: a closed tube sliced across its axis
    n, m = 24, 40
    th = np.linspace(0, 2 * np.pi, n, endpoint=False)
    zs = np.linspace(0, 10, m)
    v = np.array([[np.cos(t), np.sin(t) * 0.5, z] for z in zs for t in th], np.float32)
    f = []
    for j in range(m - 1):
        for i in range(n):
            a, b = j * n + i, j * n + (i + 1) % n
            f += [[a, b, a + n], [b, b + n, a + n]]
    f = np.array(f, np.uint32)
    nr = orc.normals_from_mesh_faces(v, f)
    tube = meshes.pack_blob(v, nr, f)
    pl = nb.PlaneSlicerRasterPlanner(nb.FixedDirectionGenerator([1, 0, 0]), nb.FixedOriginGenerator([0, 0, 5.0]), ctx)
    pl.set_line_spacing(0.7)
    ctx.upload(tube, force=True)
    p = orc.pca(v, orc.MOMENTS_F64)
    lat = orc.make_lattice(p, [1, 0, 0], [0, 0, 5.0], 0.7, True)
    # closed contours come out as closed segments (first point repeated), exactly as the reference's vtkStripper
    # emits a loop it can walk in one piece: intent and faithful modes agree, and so does the GPU
    ref = orc.slice_mesh(v, nr, f, lat, orc.SLICE_INTENT).flat()
    faithful = orc.slice_mesh(v, nr, f, lat, orc.SLICE_FAITHFUL).flat()
    np.testing.assert_array_equal(ref.pose, faithful.pose)
    gpu = ctx.slice(capi_lattice_from_oracle(lat))
    assert_same_paths(gpu, ref, "closed loops")
    b, e = int(gpu.segment_offsets[0]), int(gpu.segment_offsets[1])
    assert e - b == n * 2 + 1 or e - b > n
    np.testing.assert_array_equal(gpu.positions[b], gpu.positions[e - 1])
    with pytest.raises(_capi.Nb2Error):
        pl.set_line_spacing(0.0)
        pl.plan(tube)


def test_legacy_planner(ctx, orc, ply_mesh):
    import noether_b200 as nb
    xyz, nrm, tri = ply_mesh.xyz(), ply_mesh.normals(), ply_mesh.tri
    cfg = {"name": "PlaneSlicerLegacy", "direction_generator": {"name": "FixedDirection", "direction": [1.0, 0.0, 0.0]},
           "origin_generator": {"name": "FixedOrigin", "origin": [0.0, 0.0, 0.0]}, "line_spacing": 1.05,
           "point_spacing": 0.25, "min_segment_size": 0.5, "min_hole_size": 0.0, "bidirectional": True}
    gpu = nb.Factory(ctx).create_tool_path_planner(cfg).plan(ply_mesh)
    paths, _, _ = orc.plan_plane_slicer_legacy(xyz, nrm, tri, 1.05, 0.25, 0.5, 0.0, True, [1, 0, 0], [0, 0, 0],
                                               moments_mode=orc.MOMENTS_F64, slice_mode=orc.SLICE_INTENT)
    ref = paths.flat()
    assert gpu.legacy and gpu.n_paths == ref.num_paths
    np.testing.assert_array_equal(gpu.segment_offsets.astype(np.uint64), ref.seg_off)
    gp = gpu.poses()
    diag = np.linalg.norm(xyz.max(0) - xyz.min(0))
    assert np.abs(gp[:, :3, 3] - ref.pose[:, :3, 3]).max() <= 1e-6 * diag
    assert np.abs(gp[:, :3, :3] - ref.pose[:, :3, :3]).max() <= 1e-5
    # uniform spacing inside every segment
    for s in range(gpu.n_segments):
        b, e = int(gpu.segment_offsets[s]), int(gpu.segment_offsets[s + 1])
        # samples are 0.25 apart along the polyline, so their chords are at most 0.25 (and close to it on this mesh)
        step = np.linalg.norm(np.diff(gp[b:e, :3, 3], axis=0), axis=1)
        assert np.all(step[1:-1] <= 0.25 + 1e-9) and np.all(step[1:-1] > 0.2)


# ---------------------------------------------------------------------------------------------------------------------
# benchmark-sized properties (1 M triangles) — structure checks that do not need the oracle's full run
# ---------------------------------------------------------------------------------------------------------------------
def test_cfg2_properties(ctx, orc):
    import noether_b200 as nb
    from noether_b200 import meshes
    xyz, tri = meshes.wavy(800, 625)
    blob0 = meshes.pack_blob(xyz, None, tri, "PointXYZ")
    ctx.upload(blob0, force=True)
    nrm = ctx.normals_from_mesh_faces(blob0.n_vertices)
    blob = meshes.pack_blob(xyz, nrm, tri)
    pl = nb.PlaneSlicerRasterPlanner(nb.PrincipalAxisDirectionGenerator(0.0, ctx), nb.CentroidOriginGenerator(ctx), ctx)
    pl.set_line_spacing(0.005)
    gpu = pl.plan(blob)
    lat = ctx.last_lattice()
    n = np.array(lat.cut_normal)
    assert 190 <= gpu.n_paths <= 210
    assert np.all(gpu.segments_per_path() == 1)           # no holes: one chain per plane
    assert np.all(np.diff(gpu.path_plane) == 1)           # consecutive lattice planes
    # every waypoint lies on its plane (float32 storage error only) and on its generating mesh edge
    seg_of_wp = np.repeat(np.arange(gpu.n_segments), np.diff(gpu.segment_offsets.astype(np.int64)))
    path_of_seg = np.repeat(np.arange(gpu.n_paths), gpu.segments_per_path())
    plane = gpu.path_plane[path_of_seg[seg_of_wp]]
    o = np.array(lat.start_loc)[None, :] + (plane[:, None] * lat.line_spacing) * n[None, :]
    dist = ((gpu.positions.astype(np.float64) - o) * n[None, :]).sum(1)
    assert np.abs(dist).max() < 2e-7
    a, b = xyz[gpu.edge_lo].astype(np.float64), xyz[gpu.edge_hi].astype(np.float64)
    ab = b - a
    t = ((gpu.positions - a) * ab).sum(1) / np.maximum((ab * ab).sum(1), 1e-30)
    assert t.min() > -1e-6 and t.max() < 1 + 1e-6
    assert np.abs(a + t[:, None] * ab - gpu.positions).max() < 2e-7
    # consecutive waypoints of a segment share a triangle: their generating edges share a vertex
    e0 = np.stack([gpu.edge_lo, gpu.edge_hi], 1)
    same_seg = seg_of_wp[1:] == seg_of_wp[:-1]
    share = (e0[1:, :, None] == e0[:-1, None, :]).any(axis=(1, 2))
    # (a waypoint that landed on a mesh vertex reports the edge of its first inserter, which may belong to a
    # neighbouring triangle of the fan)
    on_vertex = (np.minimum(np.abs(gpu.positions - xyz[gpu.edge_hi]).max(1),
                            np.abs(gpu.positions - xyz[gpu.edge_lo]).max(1)) < 1e-9)
    assert np.all((share | on_vertex[1:] | on_vertex[:-1])[same_seg])
    # and the binned oracle agrees bit for bit on the whole 1 M-triangle plan
    paths, lat_o, _ = orc.plan_plane_slicer(xyz, nrm, tri, 0.005, True, slice_mode=orc.SLICE_INTENT,
                                            scan_mode=orc.SCAN_BINNED, frame=oracle_pca_from_capi(orc, ctx.pca()))
    assert bytes(lat_o) == bytes(lat)
    assert_same_paths(gpu, paths.flat(), "cfg2", poses=False)


def test_cfg3_full_size_plan_matches_oracle(ctx, orc):
    """BASELINE.json configs[2] — the configuration the headline metric is quoted on — at FULL size: wavy(2500, 2000),
    10,000,000 triangles, 1 mm spacing, PrincipalAxis(0) + Centroid.  The oracle's lattice-binned plan (same output as its
    every-cell scan, about 11 s on one core) against the GPU plan: lattice bytes, segment offsets, generating edges,
    positions and normals bit for bit over all 4.0 M waypoints."""
    import noether_b200 as nb
    from noether_b200 import meshes
    xyz, tri = meshes.wavy(2500, 2000)
    assert tri.shape[0] == 10_000_000 and xyz.shape[0] == 5_004_501
    blob0 = meshes.pack_blob(xyz, None, tri, "PointXYZ")
    ctx.upload(blob0, force=True)
    nrm = ctx.normals_from_mesh_faces(blob0.n_vertices)
    assert np.array_equal(nrm.view(np.uint32), orc.normals_from_mesh_faces(xyz, tri).view(np.uint32))
    blob = meshes.pack_blob(xyz, nrm, tri)
    pl = nb.PlaneSlicerRasterPlanner(nb.PrincipalAxisDirectionGenerator(0.0, ctx), nb.CentroidOriginGenerator(ctx), ctx)
    pl.set_line_spacing(0.001)
    gpu = pl.plan(blob)
    lat = ctx.last_lattice()
    assert 995 <= gpu.n_paths <= 1005 and gpu.n_waypoints > 3_900_000
    assert np.all(gpu.segments_per_path() == 1) and np.all(np.diff(gpu.path_plane) == 1)
    paths, lat_o, _ = orc.plan_plane_slicer(xyz, nrm, tri, 0.001, True, slice_mode=orc.SLICE_INTENT,
                                            scan_mode=orc.SCAN_BINNED, frame=oracle_pca_from_capi(orc, ctx.pca()))
    assert bytes(lat_o) == bytes(lat)
    assert_same_paths(gpu, paths.flat(), "cfg3", poses=False)
    # the same plan with the bench's settings (no generating edges, tool paths left on the device): same counts
    n = (gpu.n_paths, gpu.n_segments, gpu.n_waypoints)
    pl.emit_edges = False
    again = pl.plan(blob)
    assert (again.n_paths, again.n_segments, again.n_waypoints) == n
    assert np.array_equal(again.positions.view(np.uint32), gpu.positions.view(np.uint32))
    assert np.array_equal(again.normals.view(np.uint32), gpu.normals.view(np.uint32))


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[3] (cfg4) and configs[4] (cfg5) at full size
# ---------------------------------------------------------------------------------------------------------------------
def test_cfg4_normals_and_pca_on_50m_triangle_scan_mesh(ctx, orc):
    """NormalsFromMeshFaces + PrincipalAxis / Centroid on the 50 M-triangle mesh with scanner-like (permuted) vertex and
    face order: normals bit-exact against the oracle's face-ordered float sums, PCA frame against its fp64 moments."""
    import torch
    from noether_b200 import meshes
    if torch.cuda.mem_get_info()[0] < 24 * 2**30:
        pytest.skip("needs 24 GB of free device memory")
    xyz, tri = meshes.wavy(6250, 4000)
    assert tri.shape[0] == 50_000_000 and xyz.shape[0] == 25_010_251
    xyz, tri = meshes.permute(xyz, tri, seed=42)
    blob = meshes.pack_blob(xyz, None, tri, "PointXYZ")
    ctx.upload(blob, force=True)
    nrm = ctx.normals_from_mesh_faces(blob.n_vertices)
    ref = orc.normals_from_mesh_faces(xyz, tri)
    assert np.array_equal(bits(nrm), bits(ref))
    # unit length, and close to the analytic normal of z = A sin(2 pi x / L) sin(2 pi y / L)
    assert np.abs(np.linalg.norm(nrm.astype(np.float64), axis=1) - 1.0).max() < 1e-6
    A, lam = 0.02, 0.25
    sub = np.random.default_rng(0).integers(0, xyz.shape[0], 200_000)
    x, y = xyz[sub, 0].astype(np.float64), xyz[sub, 1].astype(np.float64)
    interior = (x > 1e-3) & (x < 1 - 1e-3) & (y > 1e-3) & (y < 0.8 - 1e-3)
    w = 2 * np.pi / lam
    g = np.stack([-A * w * np.cos(w * x) * np.sin(w * y), -A * w * np.sin(w * x) * np.cos(w * y), np.ones_like(x)], 1)
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    cosang = (g * nrm[sub].astype(np.float64)).sum(1)[interior]
    assert np.arccos(np.clip(cosang, -1, 1)).max() < 2e-3
    # PCA frame: centroid, axes and OBB extents against the fp64 oracle
    g_pca = ctx.pca()
    o = orc.pca(xyz, orc.MOMENTS_F64)
    diag = float(np.linalg.norm(xyz.max(0) - xyz.min(0)))
    assert np.abs(np.array(g_pca.mean) - o.mean).max() <= 1e-7 * diag
    assert np.abs(np.array(g_pca.scales) - o.scales).max() <= 1e-6 * diag
    ev = np.array(g_pca.evecs, np.float32).reshape(3, 3).T
    for i in range(3):
        assert abs(float(ev[:, i] @ o.evecs[:, i])) > 1 - 1e-9
    d = ctx.principal_axis_direction(0.0)
    assert abs(abs(d[1]) - 1.0) < 1e-4          # Lx = 1.0 > Ly = 0.8: the second principal axis is y
    assert np.abs(ctx.centroid_origin() - np.array(o.mean, np.float64)).max() <= 1e-7 * diag


def test_cfg5_share_of_the_mesh_batch(ctx, orc):
    """Whole-mesh sharding: the meshes rank 0 of 8 receives under the LPT assignment, planned back to back on one context
    at 2 mm spacing; structure checks on all of them, bit-exact oracle comparison on the first."""
    import noether_b200 as nb
    from noether_b200 import meshes
    from noether_b200.dist import lpt_assign
    share = lpt_assign(meshes.cfg5_triangle_counts(64), 8)[0]
    assert len(share) == 8
    pl = nb.PlaneSlicerRasterPlanner(nb.PrincipalAxisDirectionGenerator(0.0, ctx), nb.CentroidOriginGenerator(ctx), ctx)
    pl.set_line_spacing(0.002)
    for n, k in enumerate(share[:4]):
        xyz, tri, T = meshes.cfg5_mesh(k)
        assert 1_450_000 <= tri.shape[0] <= 2_550_000
        ctx.upload(meshes.pack_blob(xyz, None, tri, "PointXYZ"), force=True)
        nrm = ctx.normals_from_mesh_faces(xyz.shape[0])
        gpu = pl.plan(meshes.pack_blob(xyz, nrm, tri))
        lat = ctx.last_lattice()
        # rasters step along the long side (1 m) of the part whatever its pose: about 500 planes, one chain each.  (The
        # first / last lattice plane may graze the boundary column, which is no longer exactly planar once the posed
        # vertices are rounded to float32: it is then cut into fragments, as the reference would cut it.)
        assert 480 <= gpu.n_paths <= 520 and np.all(gpu.segments_per_path()[1:-1] == 1)
        assert np.all(np.diff(gpu.path_plane) == 1)
        nvec = np.array(lat.cut_normal)
        assert abs(abs(float(nvec @ T[:3, 0])) - 1.0) < 1e-3      # cut normal = the part's own x axis
        seg_of_wp = np.repeat(np.arange(gpu.n_segments), np.diff(gpu.segment_offsets.astype(np.int64)))
        path_of_seg = np.repeat(np.arange(gpu.n_paths), gpu.segments_per_path())
        plane = gpu.path_plane[path_of_seg[seg_of_wp]]
        o = np.array(lat.start_loc)[None, :] + (plane[:, None] * lat.line_spacing) * nvec[None, :]
        dist = ((gpu.positions.astype(np.float64) - o) * nvec[None, :]).sum(1)
        assert np.abs(dist).max() < 1e-6                          # float32 storage of coordinates up to ~2 m
        if n == 0:
            paths, lat_o, _ = orc.plan_plane_slicer(xyz, nrm, tri, 0.002, True, slice_mode=orc.SLICE_INTENT,
                                                    scan_mode=orc.SCAN_BINNED, frame=oracle_pca_from_capi(orc, ctx.pca()))
            assert bytes(lat_o) == bytes(lat)
            assert_same_paths(gpu, paths.flat(), f"cfg5 mesh {k}", poses=False)
        gpu.free()


# ---------------------------------------------------------------------------------------------------------------------
# randomised parity: many small meshes, every knob drawn at random, GPU against the oracle bit for bit
# ---------------------------------------------------------------------------------------------------------------------
def _fuzz_case(seed, orc, big=False):
    from noether_b200 import meshes
    rng = np.random.default_rng(1000 + seed)
    # small meshes: planes of a few dozen hits (the linker's general path); big ones: hundreds per plane (its fast path)
    nx, ny = (int(rng.integers(120, 380)), int(rng.integers(120, 380))) if big else (int(rng.integers(6, 70)), int(rng.integers(6, 70)))
    lx, ly = float(rng.uniform(0.3, 2.0)), float(rng.uniform(0.3, 2.0))
    if abs(lx - ly) < 0.1:
        lx += 0.25  # keep the in-plane principal axes distinguishable
    xyz, tri = meshes.wavy(nx, ny, lx=lx, ly=ly, amp=float(rng.uniform(0.0, 0.08)), lam=float(rng.uniform(0.2, 1.5)))
    if rng.random() < 0.5:
        xyz, tri = meshes.cut_hole(xyz, tri, centre_frac=(rng.uniform(0.3, 0.7), rng.uniform(0.3, 0.7)),
                                   radius_frac=float(rng.uniform(0.08, 0.3)))
    aligned = rng.random() < 0.25  # lattice planes through whole rows of vertices: vertex hits, in-plane edges
    if not aligned and rng.random() < 0.7:
        T = np.eye(4)
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        w, x, y, z = q
        T[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
        T[:3, 3] = rng.uniform(-3, 3, size=3)
        xyz = meshes.rigid_transform(xyz, T)
    nrm = orc.normals_from_mesh_faces(xyz, tri)
    if rng.random() < 0.5:
        xyz, tri, nrm = meshes.permute(xyz, tri, seed=int(rng.integers(1 << 30)), normals=nrm)
    layout = ["PointNormal", "NormalsFirst", "Packed24"][int(rng.integers(3))]
    kw = dict(bidirectional=bool(rng.random() < 0.8))
    if aligned:
        # z == 0 would make the PCA normal ill-defined; keep the waves, cut along y through vertex columns
        kw.update(direction=[0.0, 1.0, 0.0], origin=[float(xyz[:, 0].min()), 0.0, 0.0],
                  spacing=float(np.float32(lx / nx)) * int(rng.integers(1, 4)))
    else:
        kw["spacing"] = float(rng.uniform(0.6, 6.0) * min(lx / nx, ly / ny) * rng.choice([1.0, 3.0]))
        mode = rng.random()
        if mode < 0.4:
            kw["rotation_offset"] = float(rng.uniform(-np.pi, np.pi)) if rng.random() < 0.5 else 0.0
        else:
            d = rng.normal(size=3)
            kw["direction"] = (d / np.linalg.norm(d)).tolist()
        if rng.random() < 0.5:
            kw["origin"] = (xyz.mean(0) + rng.uniform(-0.2, 0.2, size=3)).tolist()
    return meshes.pack_blob(xyz, nrm, tri, layout), xyz, nrm, tri, kw


@pytest.mark.parametrize("block", range(6))
def test_fuzz_plans_match_oracle(ctx, orc, block):
    import noether_b200 as nb
    from noether_b200 import _capi
    refused = 0
    big = block >= 4
    for seed in (range(200 + 12 * block, 212 + 12 * block) if big else range(25 * block, 25 * block + 25)):
        blob, xyz, nrm, tri, kw = _fuzz_case(seed, orc, big)
        dg = (nb.FixedDirectionGenerator(kw["direction"]) if "direction" in kw
              else nb.PrincipalAxisDirectionGenerator(kw.get("rotation_offset", 0.0), ctx))
        og = nb.FixedOriginGenerator(kw["origin"]) if "origin" in kw else nb.CentroidOriginGenerator(ctx)
        planner = nb.PlaneSlicerRasterPlanner(dg, og, ctx)
        planner.set_line_spacing(kw["spacing"])
        planner.generate_rasters_bidirectionally(kw["bidirectional"])
        try:
            gpu = planner.plan(blob)
        except _capi.Nb2Error as e:
            # a direction (nearly) parallel to the mesh normal has no cutting planes; a cut point shared by more line
            # ends than the linker re-pairs is reported, never mis-linked
            assert e.status in (_capi.NB2_ERR_INVALID_ARGUMENT, _capi.NB2_ERR_NON_MANIFOLD), f"seed {seed}: {e}"
            refused += 1
            continue
        paths, lat, _ = orc.plan_plane_slicer(xyz, nrm, tri, kw["spacing"], kw["bidirectional"],
                                              direction=kw.get("direction"), origin=kw.get("origin"),
                                              rotation_offset=kw.get("rotation_offset", 0.0), slice_mode=orc.SLICE_INTENT,
                                              scan_mode=orc.SCAN_BINNED, frame=oracle_pca_from_capi(orc, ctx.pca()))
        assert bytes(ctx.last_lattice()) == bytes(lat), f"seed {seed}: lattice"
        assert_same_paths(gpu, paths.flat(), f"seed {seed} {kw}")
        gpu.free()
    assert refused <= 3


def test_pca_rotated_direction_generator(ctx, orc, ply_mesh):
    """PCARotatedDirectionGenerator (pca_rotated_direction_generator.cpp:13-25) through the factory, bit-exact against the
    oracle given the same frame; composes with the planner like any foreign generator."""
    import noether_b200 as nb
    xyz, nrm, tri = ply_mesh.xyz(), ply_mesh.normals(), ply_mesh.tri
    f = nb.Factory(ctx)
    for nested, angle in (({"name": "FixedDirection", "direction": [1.0, 0.25, 0.0]}, 0.7),
                          ({"name": "PrincipalAxis", "rotation_offset": 0.3}, -1.9)):
        gen = f.create_direction_generator({"name": "PCARotated", "direction_generator": nested, "rotation_offset": angle})
        got = gen.generate(ply_mesh)
        frame = oracle_pca_from_capi(orc, ctx.pca())
        nominal = (np.array(nested["direction"], np.float64) if nested["name"] == "FixedDirection"
                   else orc.principal_axis_direction(frame, nested["rotation_offset"]))
        want = orc.pca_rotated_direction(frame, angle, nominal)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
        assert abs(np.linalg.norm(got) - np.linalg.norm(nominal)) < 1e-12
    cfg = {"name": "PlaneSlicer", "line_spacing": 1.05, "bidirectional": True,
           "direction_generator": {"name": "PCARotated", "rotation_offset": 0.5,
                                   "direction_generator": {"name": "PrincipalAxis", "rotation_offset": 0.0}},
           "origin_generator": {"name": "Centroid"}}
    gpu = f.create_tool_path_planner(cfg).plan(ply_mesh)
    frame = oracle_pca_from_capi(orc, ctx.pca())
    d = orc.pca_rotated_direction(frame, 0.5, orc.principal_axis_direction(frame, 0.0))
    paths, lat, _ = orc.plan_plane_slicer(xyz, nrm, tri, 1.05, True, direction=d, slice_mode=orc.SLICE_INTENT, frame=frame)
    assert bytes(ctx.last_lattice()) == bytes(lat)
    assert_same_paths(gpu, paths.flat(), "PCARotated plan")


def test_cross_hatch_shares_one_upload_and_one_frame(ctx, orc):
    """a cross-hatch is a `Multi` planner of two plane slicers, the second with a PCARotated(90 deg)
    direction (noether_gui cross_hatch_plane_slicer_raster_planner.cpp:55; multi_tool_path_planner.cpp:10-20).  Both
    plans run on the mesh the context already holds: the second neither uploads nor recomputes moments / extents, and
    the concatenation equals the oracle's two plans bit for bit."""
    import noether_b200 as nb
    blob, xyz, nrm, tri = _wavy_mesh(200, 160, hole=True)
    base = {"name": "PlaneSlicer", "line_spacing": 0.013, "bidirectional": True, "origin_generator": {"name": "Centroid"}}
    pa = {"name": "PrincipalAxis", "rotation_offset": 0.0}
    cfg = {"name": "Multi", "planners": [dict(base, direction_generator=pa),
                                         dict(base, direction_generator={"name": "PCARotated", "rotation_offset": np.pi / 2,
                                                                         "direction_generator": pa})]}
    multi = nb.Factory(ctx).create_tool_path_planner(cfg)
    ctx._mesh_key = None
    first, second = multi.plan_each(blob)
    stages_second = ctx.timings()
    frame = oracle_pca_from_capi(orc, ctx.pca())
    d0 = orc.principal_axis_direction(frame, 0.0)
    d1 = orc.pca_rotated_direction(frame, np.pi / 2, d0)
    refs = []
    for gpu, d in ((first, d0), (second, d1)):
        paths, _, _ = orc.plan_plane_slicer(xyz, nrm, tri, 0.013, True, direction=d, slice_mode=orc.SLICE_INTENT, frame=frame)
        assert_same_paths(gpu, paths.flat(), "cross-hatch plan")
        refs.append(paths.flat())
    assert first.n_paths > 40 and second.n_paths > 40
    # the two plans cut along directions a quarter turn apart
    assert abs(np.dot(d0, d1)) < 1e-6
    # the whole of it as one ToolPaths, in planner order
    both = multi.plan(blob)
    assert both.n_paths == first.n_paths + second.n_paths
    np.testing.assert_array_equal(both.poses(), np.concatenate([refs[0].pose, refs[1].pose]))
    # the second plan found extents and lattice inputs on the device: its stage list has no extents pass
    assert "classify" in stages_second and "extents+lattice" not in stages_second, stages_second


@pytest.mark.parametrize("min_hole,min_seg", [(0.0, 0.03), (0.0, 0.12), (0.35, 0.03), (0.35, 0.3)])
def test_legacy_modifiers_on_device_match_oracle(ctx, orc, min_hole, min_seg):
    """JoinCloseSegments / MinimumSegmentLength / UniformSpacing run on the GPU (kernels_legacy.cu) on the slicer's
    result while it is still in HBM.  Against the oracle's sequential double arithmetic: same structure, poses bit for
    bit.  The mesh has a hole (two segments per raster there): min_hole_size 0.35 bridges it where it is narrow,
    min_segment_size drops the short rasters near the rim."""
    import noether_b200 as nb
    blob, xyz, nrm, tri = _wavy_mesh(160, 128, hole=True)
    pl = nb.PlaneSlicerLegacyRasterPlanner(nb.FixedDirectionGenerator([0.0, 1.0, 0.0]), nb.FixedOriginGenerator([0.0, 0.0, 0.0]),
                                           0.013, min_seg, min_hole, ctx)
    pl.set_line_spacing(0.0171)
    if min_hole == 0.0 and min_seg == 0.03:
        # without a length filter a two-point sliver at the rim of the hole reaches UniformSpacing, which refuses it
        # (uniform_spacing_modifier.cpp:20-21): same error from the device path and from the oracle
        from noether_b200 import _capi
        import oracle
        p0 = nb.PlaneSlicerLegacyRasterPlanner(nb.FixedDirectionGenerator([0.0, 1.0, 0.0]),
                                               nb.FixedOriginGenerator([0.0, 0.0, 0.0]), 0.013, 0.0, 0.0, ctx)
        p0.set_line_spacing(0.0171)
        try:
            p0.plan(blob).free()
            gpu_refuses = False
        except _capi.Nb2Error as e:
            gpu_refuses = e.status == _capi.NB2_ERR_SEGMENT_TOO_SHORT
        try:
            orc.plan_plane_slicer_legacy(xyz, nrm, tri, 0.0171, 0.013, 0.0, 0.0, True, [0, 1, 0], [0, 0, 0],
                                         slice_mode=orc.SLICE_INTENT, frame=oracle_pca_from_capi(orc, ctx.pca()))
            orc_refuses = False
        except oracle.OracleError:
            orc_refuses = True
        assert gpu_refuses == orc_refuses
    gpu = pl.plan(blob)
    paths, lat, _ = orc.plan_plane_slicer_legacy(xyz, nrm, tri, 0.0171, 0.013, min_seg, min_hole, True, [0, 1, 0], [0, 0, 0],
                                                 slice_mode=orc.SLICE_INTENT, frame=oracle_pca_from_capi(orc, ctx.pca()))
    ref = paths.flat()
    assert gpu.legacy and gpu.n_paths == ref.num_paths > 40
    np.testing.assert_array_equal(gpu.path_offsets.astype(np.uint64), ref.path_off)
    np.testing.assert_array_equal(gpu.segment_offsets.astype(np.uint64), ref.seg_off)
    np.testing.assert_array_equal(gpu.path_plane, ref.path_plane)
    if min_hole == 0.0:
        assert (gpu.segments_per_path() == 2).sum() > 10      # the hole splits rasters ...
    else:
        assert (gpu.segments_per_path() == 2).sum() < (np.diff(ref.path_off.astype(np.int64)) >= 1).sum()
    np.testing.assert_array_equal(gpu.poses(), ref.pose)      # bit for bit
    np.testing.assert_array_equal(bits(gpu.positions), bits(ref.pos))
    np.testing.assert_array_equal(bits(gpu.normals), bits(ref.nrm))


@pytest.mark.gpu
def test_aabb_center_origin_matches_oracle(ctx, orc, ply_mesh):
    """AABBCenterOriginGenerator (aabb_center_origin_generator.cpp:9-17) from the AABB of the ingest pass, bit for bit, and as
    the origin generator of a plan"""
    import noether_b200 as nb
    from noether_b200 import meshes
    for blob in (ply_mesh, meshes.pack_blob(*(lambda x, t: (x, None, t))(*meshes.wavy(64, 48)), "PointXYZ")):
        got = nb.AABBCenterOriginGenerator(ctx).generate(blob)
        assert np.array_equal(got, orc.aabb_center_origin(blob.xyz()))
    gen = nb.Factory(ctx).create_origin_generator({"name": "AABBCenter"})
    assert np.array_equal(gen.generate(ply_mesh), orc.aabb_center_origin(ply_mesh.xyz()))


@pytest.mark.gpu
def test_lazy_download_equals_ordinary_download():
    """download = 2 (positions / normals land behind nb2_plan, nb2_result_poses expands them as they arrive) gives the very
    poses, positions and normals of the ordinary download — on a plan large enough to travel in chunks, from pageable host
    arrays through the staging ring with the cloud compacted (NB2_MESH_GEOMETRY_ONLY)"""
    import ctypes as C

    import noether_b200 as nb
    from noether_b200 import _capi, meshes
    ctx = nb.Context(0)
    xyz, tri = meshes.wavy(640, 500)
    blob0 = meshes.pack_blob(xyz, None, tri, "PointXYZ")
    ctx.upload(blob0, force=True)
    nrm = ctx.normals_from_mesh_faces(xyz.shape[0])
    blob = meshes.pack_blob(xyz, nrm, tri, "PointNormal")
    t32 = np.ascontiguousarray(tri.reshape(-1), np.uint32)
    prm = _capi.PlanParams()
    _capi.load().nb2_plan_params_default(prm)
    prm.line_spacing = 0.002
    prm.raster_post_ops = 0
    prm.emit_edges = 0
    out = {}
    for mode, flags in ((1, 0), (2, _capi.NB2_MESH_GEOMETRY_ONLY)):
        prm.download = mode
        mv = _capi.MeshView(blob.data.ctypes.data, xyz.shape[0], blob.point_step, blob.off_xyz, blob.off_normal, flags,
                            t32.ctypes.data, tri.shape[0])
        r = C.c_void_p()
        _capi.check(ctx._lib.nb2_plan_mesh(ctx._h, C.byref(mv), C.byref(prm), C.byref(r)))
        from noether_b200.planner import ToolPaths
        tp = ToolPaths(ctx, r, layout_only=mode == 2)
        assert tp.n_waypoints >= 1 << 18, "the plan is too small to be downloaded in chunks"
        poses = tp.poses()       # (mode 2: consumes the chunks as they land)
        tp.wait()
        out[mode] = (poses, tp.positions.copy(), tp.normals.copy(), tp.segment_offsets.copy(), tp.path_plane.copy())
        tp.free()
    for a, b in zip(out[1], out[2]):
        assert a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))
    ctx.close()


def test_second_generation_linker_matches_oracle():
    """NB2_LINK_GEN=2 (the linker's second-generation fast path: plain-store claim rounds, one dart per line, two CTAs per SM;
    not the default since the first generation measured faster) must give the same bits: the slicing tests of this file once
    more, in a process of their own with the variable set (the library reads it once per process)"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if os.environ.get("NB2_LINK_GEN") == "2":
        pytest.skip("already inside the second-generation run")
    env = dict(os.environ, NB2_LINK_GEN="2")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q",
                        "-k", "test_slice_ or test_plan_end_to_end or test_cfg2_properties or test_large_plane or test_fuzz_plans"],
                       env=env, capture_output=True, text=True, cwd=root)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_bulk_copy_ingest_variant():
    """NB2_INGEST_TMA=1 (k_ingest_moments_tma: cp.async.bulk tiles into shared memory, an mbarrier per stage; opt-in because it
    measured level with the plain kernel) gives the same positions, the same frame within the tolerance of the fp64 tree sums
    and the same plans: the ingest, PCA and planning tests of this file once more in a process with the variable set"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if os.environ.get("NB2_INGEST_TMA") == "1":
        pytest.skip("already inside the bulk-copy run")
    env = dict(os.environ, NB2_INGEST_TMA="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q",
                        "-k", "test_ingest_ or test_pca_ or test_lattice_ or test_plan_end_to_end or test_cfg2_properties or test_cfg3_full or test_lazy_download"],
                       env=env, capture_output=True, text=True, cwd=root, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
