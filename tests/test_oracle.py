"""The CPU oracle pinned against the reference's own test expectations and closed-form known answers (no GPU).

The reference's tests pin counts and topology only; everything numeric is pinned here against
closed-form cases instead, and the third-party restatements (Eigen solver, stripper + sort + join) against numpy and
against the intent-mode chains.
"""
import os

import numpy as np
import pytest

from util import canonical_segments


# ---------------------------------------------------------------------------------------------------------------------
# the reference's own tests (tool_path_planner_utest.cpp, tool_path_modifier_utest.cpp)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("slice_mode", ["faithful", "intent"])
def test_ref_flat_square_mesh(orc, slice_mode):
    # tool_path_planner_utest.cpp:74-120
    mode = orc.SLICE_FAITHFUL if slice_mode == "faithful" else orc.SLICE_INTENT
    dim, n_lines = 1.0, 10
    T = orc.fixture_random_transform(dim * 5.0, np.pi)
    assert np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-12) and np.abs(T[:3, 3]).max() <= 5.0
    direction = T[:3, :3] @ np.array([1.0, 0, 0])
    xyz, nrm, tri = orc.fixture_plane_mesh(dim, dim, T)
    paths, lat, _ = orc.plan_plane_slicer(xyz, nrm, tri, 0.95 * dim / n_lines, True, direction=direction,
                                          origin=T[:3, 3], slice_mode=mode)
    nested = paths.nested_poses()
    assert len(nested) == n_lines + 1
    for path in nested:
        assert len(path) == 1
        d = path[-1][-1][:3, 3] - path[0][0][:3, 3]
        d /= np.linalg.norm(d)
        assert abs(d @ direction) > np.cos(np.radians(1.0))


@pytest.mark.parametrize("slice_mode", ["faithful", "intent"])
def test_ref_semi_planar_mesh_file(orc, ply_mesh, slice_mode):
    # tool_path_planner_utest.cpp:122-177
    mode = orc.SLICE_FAITHFUL if slice_mode == "faithful" else orc.SLICE_INTENT
    dim, n_lines = 10.5, 10
    xyz, nrm, tri = ply_mesh.xyz(), ply_mesh.normals(), ply_mesh.tri
    assert xyz.shape == (757, 3) and tri.shape == (1347, 3)
    paths, lat, _ = orc.plan_plane_slicer(xyz, nrm, tri, dim / n_lines, True, direction=[1, 0, 0], origin=[0, 0, 0],
                                          slice_mode=mode)
    nested = paths.nested_poses()
    assert len(nested) == n_lines
    for i, path in enumerate(nested):
        assert len(path) >= 1 and len(path[0]) >= 2
        if 3 < i < 7:
            assert len(path) == 2 and len(path[-1]) >= 2
        d = path[-1][-1][:3, 3] - path[0][0][:3, 3]
        d /= np.linalg.norm(d)
        assert abs(d @ np.array([1.0, 0, 0])) > np.cos(np.radians(10.0))
    # lines per plane and components, measured independently during the survey
    raw = orc.slice_mesh(xyz, nrm, tri, lat, mode).flat()
    assert list(raw.path_plane) == list(range(1, 11))
    assert list(raw.segments_per_path()) == [1, 1, 1, 1, 2, 2, 2, 2, 1, 1]
    per_plane_points = [int(raw.seg_off[raw.path_off[p + 1]] - raw.seg_off[raw.path_off[p]]) for p in range(10)]
    assert per_plane_points == [58, 60, 63, 61, 51, 40, 41, 48, 57, 57]
    assert np.allclose(np.array(lat.cut_normal), [0, 0.99354, 0.11351], atol=1e-5)


def _raster_grid(p, s, w):
    # createRasterGridToolPath, tool_path_modifier_utest.cpp:138-168
    out = []
    for pi in range(p):
        path = []
        for si in range(s):
            seg = []
            for wi in range(w):
                T = np.eye(4)
                T[:3, 3] = [si * w + wi, pi, 0.0]
                seg.append(T)
            path.append(np.array(seg))
        out.append(path)
    return out


def test_ref_organization_modifier(orc):
    # tool_path_modifier_utest.cpp:361-379: raster-organise(grid) == grid and raster-organise(shuffle(grid)) == grid
    grid = _raster_grid(4, 2, 10)
    a = orc.Paths.from_nested(grid).raster_organization().nested_poses()
    for p in range(4):
        for s in range(2):
            np.testing.assert_array_equal(a[p][s], grid[p][s])
    # shuffle(tool_paths, false) (:214-238): segment order inside every tool path, and the order of tool paths 1..n
    # (tool path 0 stays first: the reference direction is taken from it); waypoints are left alone
    for seed in range(8):
        rng = np.random.default_rng(seed)
        shuffled = [[path[i] for i in rng.permutation(2)] for path in grid]
        shuffled = [shuffled[0]] + [shuffled[1 + i] for i in rng.permutation(3)]
        b = orc.Paths.from_nested(shuffled).raster_organization().nested_poses()
        assert len(b) == 4
        for p in range(4):
            assert len(b[p]) == 2
            for s in range(2):
                np.testing.assert_array_equal(b[p][s], grid[p][s])
    # reversed segments are turned to follow tool path 0's direction
    flipped = [[seg[::-1] for seg in path] for path in grid]
    flipped[0] = grid[0]
    c = orc.Paths.from_nested(flipped).raster_organization().nested_poses()
    for p in range(4):
        for s in range(2):
            np.testing.assert_array_equal(c[p][s], grid[p][s])


# ---------------------------------------------------------------------------------------------------------------------
# closed-form known answers
# ---------------------------------------------------------------------------------------------------------------------
def test_kat_flat_square_closed_form(orc):
    # A.4 #1: unit square, identity pose, direction x, origin 0, spacing 0.095 -> 11 rasters at y = k * 0.095,
    # each 3 points (boundary edge, the 1-3 diagonal of createPlaneMesh, boundary edge), z axis (0, 0, 1)
    xyz, nrm, tri = orc.fixture_plane_mesh(1.0, 1.0)
    paths, lat, _ = orc.plan_plane_slicer(xyz, nrm, tri, 0.095, True, direction=[1, 0, 0], origin=[0, 0, 0])
    f = paths.flat()
    assert f.num_paths == 11 and list(f.segments_per_path()) == [1] * 11
    ys = sorted(float(f.pos[int(f.seg_off[int(f.path_off[p])])][1]) for p in range(11))
    assert np.allclose(ys, [k * 0.095 for k in range(-5, 6)], atol=1e-7)
    for s in range(11):
        b, e = int(f.seg_off[s]), int(f.seg_off[s + 1])
        assert e - b == 3
        y = f.pos[b][1]
        xs = sorted(f.pos[b:e, 0])
        # the diagonal joins vertices 1 (-.5,-.5) and 3 (.5,.5): it is crossed at x = y
        assert np.allclose(xs, [-0.5, y, 0.5], atol=1e-7)
        assert np.allclose(f.pose[b:e, :3, 2], [0, 0, 1])
        assert np.allclose(f.pose[b:e, 3, :], [0, 0, 0, 1])
        # DirectionOfTravel: x axis along the travel direction
        step = f.pos[b + 1] - f.pos[b]
        assert np.allclose(f.pose[b, :3, 0], step / np.linalg.norm(step), atol=1e-7)


def test_kat_vertex_hits_and_in_plane_edges(orc):
    # A.4 #2/#3: planes exactly through vertex rows: snapped points merge, zero-length lines drop, an edge lying in the
    # plane is emitted once
    from noether_b200 import meshes
    xyz, tri = meshes.wavy(8, 8, lx=1.0, ly=1.0, amp=0.0)
    nrm = np.tile(np.array([0, 0, 1], np.float32), (xyz.shape[0], 1))
    p = orc.pca(xyz)
    lat = orc.make_lattice(p, [1, 0, 0], [0, 0.5, 0], 0.125, True)
    for mode in (orc.SLICE_FAITHFUL, orc.SLICE_INTENT):
        f = orc.slice_mesh(xyz, nrm, tri, lat, mode).flat()
        assert f.num_paths >= 7
        for s in range(len(f.seg_off) - 1):
            b, e = int(f.seg_off[s]), int(f.seg_off[s + 1])
            ys = f.pos[b:e, 1]
            assert np.all(ys == ys[0])                      # on the row
            assert (ys[0] * 8) == round(ys[0] * 8)          # which is a vertex row
            assert e - b == 9                               # the 9 row vertices, no duplicates
            assert np.allclose(np.sort(f.pos[b:e, 0]), np.arange(9) / 8.0)


def test_kat_normals(orc):
    from noether_b200 import meshes
    xyz, tri = meshes.wavy(10, 7, amp=0.0)
    n = orc.normals_from_mesh_faces(xyz, tri)
    assert np.all(n == np.array([0, 0, 1], np.float32))
    # sphere cap: vertex normals within O(h^2) of radial
    u, v = np.meshgrid(np.linspace(-0.3, 0.3, 41), np.linspace(-0.3, 0.3, 41))
    z = np.sqrt(1 - u ** 2 - v ** 2)
    pts = np.stack([u.ravel(), v.ravel(), z.ravel()], 1).astype(np.float32)
    _, t = meshes.wavy(40, 40)
    n = orc.normals_from_mesh_faces(pts, t)
    inner = np.ones(41 * 41, bool).reshape(41, 41)
    inner[[0, -1], :] = False
    inner[:, [0, -1]] = False
    ang = np.arccos(np.clip((n * pts).sum(1) / np.linalg.norm(pts, axis=1), -1, 1))
    assert ang[inner.ravel()].max() < 2e-3


def test_kat_pca_axes_and_centroid(orc):
    rng = np.random.default_rng(1)
    pts = (rng.random((20000, 3)) - 0.5) * np.array([4.0, 2.0, 0.5])
    R = orc.fixture_random_transform(0.0, np.pi)[:3, :3]
    t = np.array([3.0, -2.0, 7.0])
    X = (pts @ R.T + t).astype(np.float32)
    for mode in (orc.MOMENTS_F32_SEQUENTIAL, orc.MOMENTS_F64):
        p = orc.pca(X, mode)
        for i in range(3):
            assert abs(p.evecs[:, i] @ R[:, i]) > 0.999
        assert np.allclose(p.mean, t, atol=0.05)
        assert np.allclose(p.scales, [4.0, 2.0, 0.5], rtol=0.02)
        assert np.allclose(orc.centroid(X, mode), X.astype(np.float64).mean(0), atol=1e-4)
        d = orc.principal_axis_direction(p, 0.0)
        assert abs(d @ R[:, 1]) > 0.999
        d90 = orc.principal_axis_direction(p, np.pi / 2)
        assert abs(d90 @ R[:, 0]) > 0.999


def test_eigen_solver_against_numpy(orc):
    rng = np.random.default_rng(2)
    for _ in range(200):
        A = rng.standard_normal((3, 3)) * rng.choice([1e-3, 1.0, 1e3])
        C = (A @ A.T).astype(np.float32)
        w, V = orc.eigen_solver_3f(C)
        w_np, V_np = np.linalg.eigh(C.astype(np.float64))
        assert np.all(np.diff(w) >= 0)
        assert np.allclose(w, w_np, rtol=2e-5, atol=2e-6 * abs(w_np).max())
        assert np.allclose(V.T @ V, np.eye(3), atol=2e-6)
        assert np.allclose(C @ V, V * w, atol=3e-5 * abs(w_np).max())
    # diagonal / degenerate inputs
    w, V = orc.eigen_solver_3f(np.diag([3.0, 1.0, 2.0]))
    assert np.allclose(w, [1, 2, 3])
    w, V = orc.eigen_solver_3f(np.zeros((3, 3)))
    assert np.all(w == 0) and np.allclose(V, np.eye(3))


def _spd_batch(rng, n):
    """random symmetric positive semi-definite 3x3 float32 matrices R diag(w) R^T over twelve decades of scale: well
    separated spectra, near-degenerate pairs (relative gaps 1e-7 .. 1e-2), exactly repeated eigenvalues, rank-deficient
    ones and plain A A^T"""
    out = []
    for i in range(n):
        kind = i % 5
        scale = 10.0 ** rng.uniform(-6, 6)
        if kind == 4:
            A = rng.standard_normal((3, 3))
            M = A @ A.T * scale
        else:
            q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
            w = np.sort(rng.uniform(0.05, 1.0, 3))
            if kind == 1:
                w[1] = w[0] * (1.0 + 10.0 ** rng.uniform(-7, -2))      # near-degenerate pair at the bottom
            elif kind == 2:
                w[2] = w[1] * (1.0 + 10.0 ** rng.uniform(-7, -2))      # ... at the top (the raster direction pair)
            elif kind == 3:
                w[rng.integers(0, 3)] = 0.0 if rng.random() < 0.5 else w[0]  # rank deficient / exactly repeated
            M = (q * (w * scale)) @ q.T
        M = M.astype(np.float32)
        out.append(np.tril(M) + np.tril(M, -1).T)  # exactly symmetric (the solver reads the lower triangle)
    return out


def test_eigen_solver_10k_matrices_against_numpy_and_the_product(orc):
    """The restated Eigen 3.4 SelfAdjointEigenSolver<Matrix3f> (oracle/oracle_pca.cpp) and the product's own restatement
    (noether_b200/csrc/frame_math.h through nb2_eigen_solver_3f) on 10^4 symmetric matrices, against numpy.linalg.eigh of
    the same float32 matrix in float64.

    Conventions checked: eigenvalues ASCENDING (pcl::PCA reverses them afterwards); eigenvectors are the COLUMNS,
    orthonormal; the sign of a column is not defined by the mathematics — it is whatever the solver's sequence of Givens
    rotations yields, so columns are compared up to sign, and within a (near-)degenerate cluster only the spanned subspace
    is compared.  The two restatements must agree with each other bit for bit, signs included."""
    import ctypes as C
    from noether_b200 import _capi
    from noether_b200.build import build_library
    build_library()
    L = _capi.load()
    fp = C.POINTER(C.c_float)
    rng = np.random.default_rng(20240607)
    eps = float(np.finfo(np.float32).eps)
    for M in _spd_batch(rng, 10_000):
        w, V = orc.eigen_solver_3f(M)
        cm = np.ascontiguousarray(M.T)
        w2 = np.zeros(3, np.float32)
        v2 = np.zeros(9, np.float32)
        assert L.nb2_eigen_solver_3f(cm.ctypes.data_as(fp), w2.ctypes.data_as(fp), v2.ctypes.data_as(fp)) == 0
        assert np.array_equal(w.view(np.uint32), w2.view(np.uint32))
        assert np.array_equal(V.view(np.uint32), v2.reshape(3, 3).T.copy().view(np.uint32))
        w_np, V_np = np.linalg.eigh(M.astype(np.float64))
        top = max(abs(w_np).max(), 1e-300)
        assert np.all(np.diff(w) >= 0)
        assert np.abs(w - w_np).max() <= 16 * eps * top
        assert np.abs(V.T.astype(np.float64) @ V - np.eye(3)).max() <= 16 * eps
        assert np.abs(M.astype(np.float64) @ V - V * w.astype(np.float64)).max() <= 32 * eps * top
        # eigenvectors: group numpy's eigenvalues into clusters closer than the float32 resolution of the problem and
        # compare the projector onto each cluster's eigenspace
        tol_gap = 4096 * eps * top
        start = 0
        for end in range(1, 4):
            if end == 3 or w_np[end] - w_np[end - 1] > tol_gap:
                P_np = V_np[:, start:end] @ V_np[:, start:end].T
                Vc = V[:, start:end].astype(np.float64)
                gap = min([abs(w_np[i] - w_np[j]) for i in range(start, end) for j in range(3) if not start <= j < end] or [top])
                assert np.abs(Vc @ Vc.T - P_np).max() <= 64 * eps * top / gap
                start = end


def test_kat_lattice_quirks(orc):
    # unit cube cloud centred at (0.5, 0.5, 0.05): OBB centred on the mean; planes from a floored start
    from noether_b200 import meshes
    xyz, tri = meshes.wavy(10, 10, lx=1.0, ly=0.7, amp=0.0)
    p = orc.pca(xyz)
    lat = orc.make_lattice(p, [0, 1, 0], [0.5, 0.35, 0], 0.3, True)
    n = np.array(lat.cut_normal)
    assert abs(abs(n[0]) - 1) < 1e-6
    assert lat.num_planes == int(np.ceil(abs(lat.d_max - lat.d_min) / 0.3))
    k = np.floor(abs(lat.d_min) / 0.3)
    assert np.allclose(np.array(lat.start_loc), np.array([0.5, 0.35, 0]) + n * np.copysign(k * 0.3, lat.d_min))
    # unidirectional: origin in front of the whole mesh -> no cuts
    far = orc.make_lattice(p, [0, 1, 0], np.array([0.5, 0.35, 0]) + 5 * n, 0.3, False)
    near = orc.make_lattice(p, [0, 1, 0], np.array([0.5, 0.35, 0]) - 5 * n, 0.3, False)
    assert (far.no_cuts, near.no_cuts) in ((1, 0), (0, 1))
    with pytest.raises(orc.OracleError):
        orc.make_lattice(p, [0, 1, 0], [0, 0, 0], 0.0, True)


def test_kat_legacy_uniform_spacing(orc):
    # A.4 #9: straight 1 m polyline, 0.3 m spacing -> samples at 0, 0.2, 0.5, 0.8, 1.0
    seg = []
    for x in np.linspace(0, 1, 11):
        T = np.eye(4)
        T[0, 3] = x
        seg.append(T)
    out = orc.Paths.from_nested([[np.array(seg)]]).uniform_spacing(0.3).nested_poses()
    assert np.allclose(out[0][0][:, 0, 3], [0, 0.2, 0.5, 0.8, 1.0], atol=1e-12)
    assert np.allclose(out[0][0][:, :3, 0], [1, 0, 0]) and np.allclose(out[0][0][:, :3, 2], [0, 0, 1])
    with pytest.raises(orc.OracleError) as e:
        orc.Paths.from_nested([[np.array(seg[:2])]]).uniform_spacing(0.3)
    assert e.value.code == orc.ERR_SEGMENT_TOO_SHORT
    # join / prune
    a = np.array(seg[:5])
    b = np.array(seg[6:])
    joined = orc.Paths.from_nested([[a, b]]).join_close_segments(0.25).nested_poses()
    assert len(joined[0]) == 1 and len(joined[0][0]) == 10
    kept = orc.Paths.from_nested([[a, b]]).join_close_segments(0.15).minimum_segment_length(0.45).nested_poses()
    assert len(kept) == 0 or all(len(s) for s in kept[0])


# ---------------------------------------------------------------------------------------------------------------------
# consistency of the restatements with each other
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("spacing", [0.1, 0.05, 0.37])
def test_binned_scan_equals_full_scan(orc, ply_mesh, spacing):
    xyz, nrm, tri = ply_mesh.xyz(), ply_mesh.normals(), ply_mesh.tri
    p = orc.pca(xyz)
    lat = orc.make_lattice(p, [1, 0, 0], orc.centroid(xyz), spacing, True)
    for mode in (orc.SLICE_FAITHFUL, orc.SLICE_INTENT):
        a = orc.slice_mesh(xyz, nrm, tri, lat, mode, orc.SCAN_ALL_CELLS).flat()
        b = orc.slice_mesh(xyz, nrm, tri, lat, mode, orc.SCAN_BINNED).flat()
        np.testing.assert_array_equal(a.pose, b.pose)
        np.testing.assert_array_equal(a.seg_off, b.seg_off)
        np.testing.assert_array_equal(a.edge_lo, b.edge_lo)


def test_faithful_is_a_refinement_of_intent(orc, ply_mesh):
    """the reference's stack join only merges container-adjacent pieces, so it may return a maximal
    chain in several pieces.  Every faithful segment must be a contiguous sub-chain of an intent segment, and on the
    reference's own test setting the two coincide."""
    xyz, nrm, tri = ply_mesh.xyz(), ply_mesh.normals(), ply_mesh.tri
    p = orc.pca(xyz)
    under_joined = 0
    for spacing, origin in ((1.05, [0, 0, 0]), (0.1, orc.centroid(xyz)), (0.05, orc.centroid(xyz))):
        lat = orc.make_lattice(p, [1, 0, 0], origin, spacing, True)
        fa = orc.slice_mesh(xyz, nrm, tri, lat, orc.SLICE_FAITHFUL).flat()
        it = orc.slice_mesh(xyz, nrm, tri, lat, orc.SLICE_INTENT).flat()
        assert list(fa.path_plane) == list(it.path_plane)
        ca, ci = canonical_segments(fa), canonical_segments(it)
        for pa, pi in zip(ca, ci):
            if pa == pi:
                continue
            under_joined += 1
            assert len(pa) > len(pi)
            for seg in pa:
                ok = False
                for whole in pi:
                    for cand in (whole, whole[::-1]):
                        for i in range(len(cand) - len(seg) + 1):
                            if cand[i:i + len(seg)] == seg:
                                ok = True
                    if ok:
                        break
                assert ok, "faithful piece is not a sub-chain of an intent chain"
        if spacing == 1.05:
            assert ca == ci
    assert under_joined >= 0


def test_intent_is_order_invariant(orc):
    # vertex / face permutation: same chains (as ordered float positions) after canonicalisation
    from noether_b200 import meshes
    xyz, tri = meshes.wavy(30, 24)
    xyz, tri = meshes.cut_hole(xyz, tri)
    nrm = orc.normals_from_mesh_faces(xyz, tri)
    p = orc.pca(xyz, orc.MOMENTS_F64)
    lat = orc.make_lattice(p, [0.2, 1, 0], [0.5, 0.4, 0], 0.031, True)
    a = orc.slice_mesh(xyz, nrm, tri, lat, orc.SLICE_INTENT).flat()
    x2, t2, n2 = meshes.permute(xyz, tri, seed=11, normals=nrm)
    b = orc.slice_mesh(x2, n2, t2, lat, orc.SLICE_INTENT).flat()
    np.testing.assert_array_equal(a.seg_off, b.seg_off)
    np.testing.assert_array_equal(a.pos.view(np.uint32), b.pos.view(np.uint32))
    np.testing.assert_array_equal(a.nrm.view(np.uint32), b.nrm.view(np.uint32))


def test_closed_loops_match_the_stripper(orc):
    # a tube cut across its axis: vtkStripper walks each loop in one piece; intent mode emits the same closed segment
    n, m = 16, 12
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
    p = orc.pca(v)
    lat = orc.make_lattice(p, [1, 0, 0], [0, 0, 5.0], 0.7, True)
    a = orc.slice_mesh(v, nr, f, lat, orc.SLICE_INTENT).flat()
    b = orc.slice_mesh(v, nr, f, lat, orc.SLICE_FAITHFUL).flat()
    assert a.num_paths > 5 and np.all(a.segments_per_path() == 1)
    np.testing.assert_array_equal(a.pose, b.pose)
    s0, s1 = int(a.seg_off[0]), int(a.seg_off[1])
    np.testing.assert_array_equal(a.pos[s0], a.pos[s1 - 1])


def test_kat_pca_rotated_direction(orc):
    """PCARotatedDirectionGenerator: rotation about the smallest principal axis (pca_rotated_direction_generator.cpp:24)."""
    from noether_b200 import meshes
    xyz, _ = meshes.wavy(40, 25)
    p = orc.pca(xyz, orc.MOMENTS_F64)
    e2 = p.evecs[:, 2].astype(np.float64)
    e2 /= np.linalg.norm(e2)
    v = np.array([0.3, -1.2, 0.4])
    for ang in (0.0, 0.5, -2.0, np.pi):
        r = orc.pca_rotated_direction(p, ang, v)
        want = v * np.cos(ang) + np.cross(e2, v) * np.sin(ang) + e2 * (e2 @ v) * (1 - np.cos(ang))  # Rodrigues
        assert np.abs(r - want).max() < 1e-14
    assert np.array_equal(orc.pca_rotated_direction(p, 0.0, v), v * 1.0) or np.abs(orc.pca_rotated_direction(p, 0.0, v) - v).max() < 1e-16
    # with the nominal direction = normalised e1 it is PrincipalAxis with the same offset
    e1 = p.evecs[:, 1].astype(np.float64)
    e1 /= np.linalg.norm(e1)
    assert np.abs(orc.pca_rotated_direction(p, 0.9, e1) - orc.principal_axis_direction(p, 0.9)).max() < 1e-15


def test_cpu_optimised_baseline_slabs_cover_the_plan(orc):
    """tools/cpu_optimised.py (bench.py's second CPU figure): plane slabs over forked workers produce the plan of one
    pass over the whole lattice — same tool paths, same waypoints"""
    import json
    import subprocess
    import sys
    from noether_b200 import meshes
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "cpu_optimised.py"), "cfg2", "3"], capture_output=True,
                       text=True, timeout=300, check=True)
    o = json.loads(r.stdout.strip().splitlines()[-1])
    xyz, tri = meshes.wavy(800, 625)
    nrm = orc.normals_from_mesh_faces(xyz, tri)
    whole, lat, _ = orc.plan_plane_slicer(xyz, nrm, tri, 0.005, True, slice_mode=orc.SLICE_FAITHFUL, scan_mode=orc.SCAN_BINNED,
                                          post_ops=False)
    f = whole.flat()
    assert o["workers"] == 3 and o["planes"] == int(lat.num_planes)
    assert (o["tool_paths"], o["waypoints"]) == (f.num_paths, int(f.seg_off[-1]))
    assert o["tri_per_s"] > 0


# ---------------------------------------------------------------------------------------------------------------------
# properties of the slicer oracle on random surfaces (SURVEY.md §8 c: the oracle is the de-facto specification, so it is
# checked against what a cut must satisfy independently of how it is computed)
# ---------------------------------------------------------------------------------------------------------------------
def _random_surface(rng):
    from noether_b200 import meshes
    nx, ny = int(rng.integers(3, 22)), int(rng.integers(3, 22))
    xyz, tri = meshes.wavy(nx, ny, amp=float(rng.uniform(0, 0.06)), lam=float(rng.uniform(0.15, 0.6)))
    if rng.random() < 0.5 and min(nx, ny) >= 6:
        xyz, tri = meshes.cut_hole(xyz, tri, radius_frac=float(rng.uniform(0.08, 0.3)))
    if rng.random() < 0.5:
        a = rng.uniform(-np.pi, np.pi, 3)
        cx, sx, cy, sy, cz, sz = np.cos(a[0]), np.sin(a[0]), np.cos(a[1]), np.sin(a[1]), np.cos(a[2]), np.sin(a[2])
        R = (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
             @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]))
        xyz = (xyz.astype(np.float64) @ R.T + rng.uniform(-1, 1, 3)).astype(np.float32)
    if rng.random() < 0.5:
        xyz, tri = meshes.permute(xyz, tri, seed=int(rng.integers(1 << 30)))
    return xyz, tri


@pytest.mark.parametrize("seed", range(25))
def test_property_cut_points_lie_on_their_plane_and_edge_and_chains_are_the_components(orc, seed):
    rng = np.random.default_rng(31000 + seed)
    xyz, tri = _random_surface(rng)
    nrm = orc.normals_from_mesh_faces(xyz, tri)
    spacing = float(rng.uniform(0.03, 0.2))
    paths, lat, _ = orc.plan_plane_slicer(xyz, nrm, tri, spacing, True, slice_mode=orc.SLICE_INTENT, scan_mode=orc.SCAN_BINNED,
                                          post_ops=False)
    f = paths.flat()
    if f.num_paths == 0:
        return
    n = np.array(lat.cut_normal)
    start = np.array(lat.start_loc)
    diag = float(np.linalg.norm(xyz.max(0) - xyz.min(0)))
    X = xyz.astype(np.float64)
    seg_of_wp = np.repeat(np.arange(len(f.seg_off) - 1), np.diff(f.seg_off).astype(np.int64))
    path_of_seg = np.repeat(np.arange(f.num_paths), np.diff(f.path_off).astype(np.int64))
    plane = f.path_plane[path_of_seg[seg_of_wp]]
    pos = f.pos.astype(np.float64)
    # (1) on the plane: n . (x - (start + k s n)) = 0 up to the float32 rounding of the stored point
    origin = start[None, :] + (plane[:, None] * lat.line_spacing) * n[None, :]
    assert np.abs(((pos - origin) * n).sum(1)).max() <= 4e-7 * max(diag, np.abs(pos).max())
    # (2) on its generating mesh edge, between the endpoints, and that edge is an edge of a triangle
    lo, hi = X[f.edge_lo], X[f.edge_hi]
    e = hi - lo
    L = np.linalg.norm(e, axis=1)
    t = ((pos - lo) * e).sum(1) / np.maximum(L * L, 1e-300)
    assert (t >= -1e-6).all() and (t <= 1 + 1e-6).all()
    off = np.linalg.norm(pos - (lo + t[:, None] * e), axis=1)
    assert off.max() <= 4e-7 * max(diag, np.abs(pos).max())
    edges = set()
    for a, b, c in tri.tolist():
        edges.update({(min(a, b), max(a, b)), (min(b, c), max(b, c)), (min(a, c), max(a, c))})
    assert all((min(a, b), max(a, b)) in edges or a == b for a, b in zip(f.edge_lo.tolist(), f.edge_hi.tolist()))
    # (3) per plane: segments = connected components of the cut graph built independently from the per-plane scalars
    #     (triangles with vertices on both sides of the plane, joined through the cut points they share)
    for p in range(f.num_paths):
        k = int(f.path_plane[p])
        o = start + k * lat.line_spacing * n
        s = (X - o) @ n
        inside = s >= 0
        cut = [i for i, (a, b, c) in enumerate(tri.tolist()) if not (inside[a] == inside[b] == inside[c])]
        parent = list(range(len(cut)))

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x
        by_point = {}
        for i, ti in enumerate(cut):
            a, b, c = tri[ti].tolist()
            for u, v in ((a, b), (b, c), (c, a)):
                if inside[u] != inside[v]:
                    lo_v, hi_v = (u, v) if s[u] < s[v] else (v, u)
                    tt = (0.0 - s[lo_v]) / (s[hi_v] - s[lo_v])
                    key = (X[lo_v] + tt * (X[hi_v] - X[lo_v])).astype(np.float32).tobytes()   # vtkMergePoints: float equality
                    if key in by_point:
                        ra, rb = find(by_point[key]), find(i)
                        parent[ra] = rb
                    else:
                        by_point[key] = i
        components = len({find(i) for i in range(len(cut))})
        n_seg = int(f.path_off[p + 1] - f.path_off[p])
        assert n_seg == components, f"plane {k}: {n_seg} segments, {components} components of the cut graph"
    # (4) maximal open chains: inside a segment consecutive waypoints differ; no two segments of a plane share an end point
    for p in range(f.num_paths):
        ends = []
        for sgm in range(int(f.path_off[p]), int(f.path_off[p + 1])):
            a, b = int(f.seg_off[sgm]), int(f.seg_off[sgm + 1])
            assert b - a >= 2
            assert (np.abs(np.diff(f.pos[a:b], axis=0)).sum(1) > 0).all()
            if f.pos[a].tobytes() != f.pos[b - 1].tobytes():      # an open chain
                ends += [f.pos[a].tobytes(), f.pos[b - 1].tobytes()]
        assert len(ends) == len(set(ends))


def test_kat_aabb_center_origin_is_half_the_range(orc):
    """aabb_center_origin_generator.cpp:15-16 returns (max - min) / 2, not the centre of the box: restated as it is"""
    xyz = np.array([[1, 2, 3], [4, 0, 5], [-2, 1, 9]], np.float32)
    assert orc.aabb_center_origin(xyz).tolist() == [3.0, 1.0, 3.0]
    rng = np.random.default_rng(5)
    x = (rng.normal(size=(1000, 3)) * 37).astype(np.float32)
    assert np.array_equal(orc.aabb_center_origin(x), (x.max(0) - x.min(0)).astype(np.float64) / 2.0)
