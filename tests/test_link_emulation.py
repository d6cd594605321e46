"""The CUDA path's contouring and fast linking logic on the CPU: tests/cpp/link_emulation.cpp drives the kernels' own
per-element functions (noether_b200/csrc/device_math.cuh, link2_ops.h) with loops instead of CTAs — "threads" and cut
lines in scrambled orders — and the result must equal the oracle's slice (INTENT mode: one segment per maximal chain) bit
for bit: segment structure, generating mesh edges, float32 positions, float32 interpolated normals."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

F_COLLISION, F_CROWDED, F_DIRECTION, F_LOOP, F_CAPACITY = 1, 2, 4, 8, 16


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(ROOT, "tests", "cpp", "link_emulation.cpp")
    out = os.path.join(ROOT, "tests", "cpp", "build")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "liblink_emulation.so")
    deps = [src] + [os.path.join(ROOT, "noether_b200", "csrc", f) for f in ("device_math.cuh", "link2_ops.h", "nb2_internal.h", "host_math.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                        "-I/usr/local/cuda/include", src, "-o", lib], check=True)
    L = C.CDLL(lib)
    dp = C.POINTER(C.c_double)
    L.emu_slice.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, dp, dp, C.c_double, dp, C.c_int64,
                            C.c_int64, C.c_int64, C.c_uint, C.c_uint, C.c_int]
    for f in ("emu_num_paths", "emu_num_segments", "emu_num_waypoints", "emu_num_lines", "emu_num_declined"):
        getattr(L, f).restype = C.c_uint64
    L.emu_export.argtypes = [C.c_void_p] * 9
    L.emu_work_bytes.restype = C.c_uint64
    L.emu_work_bytes.argtypes = [C.c_uint]
    return L


@pytest.fixture(scope="module")
def orc():
    import oracle
    oracle.build()
    return oracle


def emu_slice(L, xyz, nrm, tri, lat, nth=96, seed=1, edges=True, planes=None):
    xyz = np.ascontiguousarray(xyz, np.float32)
    nrm = np.ascontiguousarray(nrm, np.float32)
    tri = np.ascontiguousarray(tri, np.uint32)
    d3 = lambda a: (C.c_double * 3)(*[float(x) for x in a])
    b, e = planes if planes else (0, int(lat.num_planes))
    L.emu_slice(xyz.ctypes.data, nrm.ctypes.data, len(xyz), tri.ctypes.data, len(tri), d3(lat.cut_normal), d3(lat.start_loc),
                float(lat.line_spacing), d3(lat.cut_direction), int(lat.num_planes), b, e, nth, seed, int(edges))
    P, S, W, D = L.emu_num_paths(), L.emu_num_segments(), L.emu_num_waypoints(), L.emu_num_declined()
    out = dict(path_plane=np.zeros(P, np.int64), path_segments=np.zeros(P, np.uint32), seg_len=np.zeros(S, np.uint32),
               pos=np.zeros((W, 3), np.float32), nrm=np.zeros((W, 3), np.float32), elo=np.zeros(W, np.uint32),
               ehi=np.zeros(W, np.uint32), declined_plane=np.zeros(D, np.int64), declined_reason=np.zeros(D, np.uint32))
    L.emu_export(*[out[k].ctypes.data if out[k].size else None for k in
                   ("path_plane", "path_segments", "seg_len", "pos", "nrm", "elo", "ehi", "declined_plane", "declined_reason")])
    out["lines"] = L.emu_num_lines()
    return out


def assert_equals_oracle(orc, got, xyz, nrm, tri, lat, what, planes=None, max_declined=0, reasons=0):
    """every plane the fast path accepted against the oracle's plane, bit for bit; at most `max_declined` planes declined, and
    only for `reasons`"""
    b, e = planes if planes else (0, int(lat.num_planes))
    f = orc.slice_mesh(xyz, nrm, tri, lat, orc.SLICE_INTENT, orc.SCAN_BINNED, b, e).flat()
    declined = set(int(k) for k in got["declined_plane"])
    assert len(declined) <= max_declined, f"{what}: fast path declined planes {got['declined_plane']} ({got['declined_reason']})"
    assert all(int(r) & ~reasons == 0 for r in got["declined_reason"]), f"{what}: reasons {got['declined_reason']}"
    keep_path = np.array([int(k) not in declined for k in f.path_plane], bool)
    assert np.array_equal(got["path_plane"], f.path_plane[keep_path]), what
    seg_per_path = np.diff(f.path_off).astype(np.int64)
    assert np.array_equal(got["path_segments"], seg_per_path[keep_path].astype(np.uint32)), what
    keep_seg = np.repeat(keep_path, seg_per_path)
    seg_len = np.diff(f.seg_off).astype(np.int64)
    assert np.array_equal(got["seg_len"], seg_len[keep_seg].astype(np.uint32)), what
    keep_wp = np.repeat(keep_seg, seg_len)
    assert np.array_equal(got["elo"], f.edge_lo[keep_wp]) and np.array_equal(got["ehi"], f.edge_hi[keep_wp]), what
    assert np.array_equal(got["pos"].view(np.uint32), f.pos[keep_wp].view(np.uint32)), what
    assert np.array_equal(got["nrm"].view(np.uint32), f.nrm[keep_wp].view(np.uint32)), what
    return len(declined)


def _frame(orc, xyz, spacing, direction=None, origin=None):
    p = orc.pca(xyz, orc.MOMENTS_F64)
    d = orc.principal_axis_direction(p, 0.0) if direction is None else np.asarray(direction, np.float64)
    o = orc.centroid(xyz, orc.MOMENTS_F64) if origin is None else np.asarray(origin, np.float64)
    return orc.make_lattice(p, d, o, spacing, True)


@pytest.mark.parametrize("nx,ny,spacing,hole,nth", [(40, 32, 0.02, False, 32), (64, 50, 0.013, True, 96), (120, 90, 0.0071, True, 512),
                                                     (30, 200, 0.05, False, 32)])
def test_wavy_surfaces_bit_exact(emu, orc, nx, ny, spacing, hole, nth):
    from noether_b200 import meshes
    xyz, tri = meshes.wavy(nx, ny)
    if hole:
        xyz, tri = meshes.cut_hole(xyz, tri)
    nrm = orc.normals_from_mesh_faces(xyz, tri)
    lat = _frame(orc, xyz, spacing)
    for seed in (1, 2):
        got = emu_slice(emu, xyz, nrm, tri, lat, nth=nth, seed=seed)
        # (a lattice plane that CONTAINS a column of mesh vertices and the edges between them — x = 0.5 on these grids — has
        # cut points with more than two line ends: the general path's business)
        assert_equals_oracle(orc, got, xyz, nrm, tri, lat, f"wavy {nx}x{ny} seed {seed}", max_declined=1, reasons=F_CROWDED | F_DIRECTION)
        assert got["lines"] > 0


def test_reference_fixture_and_permuted_rigid_copies(emu, orc, golden_dir):
    from noether_b200 import meshes
    xyz, nrm, sizes, idx = orc.read_ply(open(os.path.join(golden_dir, "wavy_mesh_with_hole.ply"), "rb").read())
    tri = idx.reshape(-1, 3)
    for spacing in (1.05, 0.1, 0.05):
        lat = _frame(orc, xyz, spacing, [1, 0, 0], [0, 0, 0])
        got = emu_slice(emu, xyz, nrm, tri, lat, nth=64, seed=3)
        assert_equals_oracle(orc, got, xyz, nrm, tri, lat, f"fixture {spacing}")
    # a permuted copy under a rigid transform: same structure, slab by slab as well
    rng = np.random.default_rng(5)
    x2, t2, n2 = meshes.permute(xyz, tri, seed=9, normals=nrm)
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    T = np.eye(4)
    T[:3, :3] = q * np.sign(np.linalg.det(q))
    T[:3, 3] = [3.0, -2.0, 0.5]
    x3 = meshes.rigid_transform(x2, T)
    n3 = (n2.astype(np.float64) @ T[:3, :3].T).astype(np.float32)
    lat = _frame(orc, x3, 0.21)
    got = emu_slice(emu, x3, n3, t2, lat, nth=128, seed=11)
    assert_equals_oracle(orc, got, x3, n3, t2, lat, "permuted rigid copy")
    half = int(lat.num_planes) // 2
    got = emu_slice(emu, x3, n3, t2, lat, nth=128, seed=12, planes=(half, int(lat.num_planes)))
    assert_equals_oracle(orc, got, x3, n3, t2, lat, "upper slab", planes=(half, int(lat.num_planes)))


def test_random_surfaces_with_holes(emu, orc):
    from noether_b200 import meshes
    for seed in range(12):
        rng = np.random.default_rng(300 + seed)
        nx, ny = int(rng.integers(12, 70)), int(rng.integers(12, 70))
        xyz, tri = meshes.wavy(nx, ny, amp=float(rng.uniform(0, 0.05)), lam=float(rng.uniform(0.15, 0.6)))
        if rng.random() < 0.6:
            xyz, tri = meshes.cut_hole(xyz, tri, centre_frac=tuple(rng.uniform(0.3, 0.7, 2)), radius_frac=float(rng.uniform(0.05, 0.3)))
        nrm = orc.normals_from_mesh_faces(xyz, tri)
        xyz, tri, nrm = meshes.permute(xyz, tri, seed=seed, normals=nrm)
        lat = _frame(orc, xyz, float(rng.uniform(0.011, 0.09)))
        got = emu_slice(emu, xyz, nrm, tri, lat, nth=int(rng.choice([32, 100, 256, 1024])), seed=seed)
        assert_equals_oracle(orc, got, xyz, nrm, tri, lat, f"random surface {seed}")


def test_fast_path_declines_what_it_does_not_handle(emu, orc):
    """lattice planes through whole rows of mesh vertices (zero-length lines, cut points with many line ends) and a closed
    tube (closed contours; chains that fold back): the fast path must say so — with the reason — before writing anything;
    the planes it does accept still equal the oracle"""
    from noether_b200 import meshes
    # planes within rounding of the vertex columns (x = k * 0.05 is not a float): sliver and zero-length lines, all linked
    xyz, tri = meshes.wavy(20, 16, lx=1.0, ly=0.8, amp=0.0)
    nrm = orc.normals_from_mesh_faces(xyz, tri)
    lat = _frame(orc, xyz, 0.05, [0, 1, 0], [0.0, 0.0, 0.0])
    got = emu_slice(emu, xyz, nrm, tri, lat, nth=64, seed=1)
    assert_equals_oracle(orc, got, xyz, nrm, tri, lat, "grazed vertex columns", max_declined=2, reasons=F_CROWDED | F_DIRECTION)
    # planes that CONTAIN the vertex columns (x = k / 16 exactly, axis-aligned normal) and the mesh edges between them: the
    # in-plane edges come out once (from the triangle on the negative side), every fan triangle adds a zero-length line
    xyz, tri = meshes.wavy(16, 12, lx=1.0, ly=0.75, amp=0.03)
    nrm = orc.normals_from_mesh_faces(xyz, tri)
    lat = _frame(orc, xyz, 0.03125, [0, 1, 0], [0.0, 0.0, 0.0])
    for i, v in enumerate((1.0, 0.0, 0.0)):     # (the PCA's third axis is only nearly z)
        lat.cut_normal[i] = v
        lat.start_loc[i] = 0.0
    lat.num_planes = 33
    got = emu_slice(emu, xyz, nrm, tri, lat, nth=64, seed=1)
    assert_equals_oracle(orc, got, xyz, nrm, tri, lat, "vertex columns", max_declined=0)
    # a fin: three triangles around one edge -> a cut point with three line ends
    xyz = np.array([[0, 0, 0], [0, 1, 0], [1, 0.5, 0], [-1, 0.5, 0.5], [-1, 0.5, -0.5]], np.float32)
    tri = np.array([[0, 1, 2], [1, 0, 3], [0, 1, 4]], np.uint32)
    nrm = orc.normals_from_mesh_faces(xyz, tri)
    lat = _frame(orc, xyz, 0.3, [1, 0, 0], [0.0, 0.05, 0.0])
    for i in range(3):                          # planes y = 0.05 + 0.3 k, chains along x
        lat.cut_normal[i] = (0.0, 1.0, 0.0)[i]
        lat.start_loc[i] = (0.0, 0.05, 0.0)[i]
        lat.cut_direction[i] = (1.0, 0.0, 0.0)[i]
    lat.num_planes = 3
    got = emu_slice(emu, xyz, nrm, tri, lat, nth=8, seed=1)
    assert got["lines"] == 9
    assert len(got["declined_plane"]) > 0 and all(r & (F_CROWDED | F_DIRECTION) for r in got["declined_reason"])
    # a tube around the y axis sliced by planes y = const: every contour closes
    m, k = 48, 20
    a = np.linspace(0, 2 * np.pi, m, endpoint=False)
    ring = np.stack([np.cos(a), np.zeros(m), np.sin(a)], 1)
    xyz = np.concatenate([ring + [0, 0.1 * i, 0] for i in range(k)]).astype(np.float32)
    tri = []
    for i in range(k - 1):
        for j in range(m):
            p, q = i * m + j, i * m + (j + 1) % m
            tri += [(p, q, q + m), (p, q + m, p + m)]
    tri = np.asarray(tri, np.uint32)
    nrm = orc.normals_from_mesh_faces(xyz, tri)
    lat = _frame(orc, xyz, 0.033, [1, 0, 0], [0.0, 0.017, 0.0])
    for i in range(3):                          # planes y = 0.017 + 0.033 k
        lat.cut_normal[i] = (0.0, 1.0, 0.0)[i]
        lat.start_loc[i] = (0.0, 0.017, 0.0)[i]
        lat.cut_direction[i] = (1.0, 0.0, 0.0)[i]
    lat.num_planes = 50
    got = emu_slice(emu, xyz, nrm, tri, lat, nth=64, seed=1)
    assert len(got["declined_plane"]) == 50 and all(r & F_LOOP for r in got["declined_reason"]), got["declined_reason"]


def test_work_area_fits_two_planes_per_sm(emu):
    """cfg3's planes (about 4000 cut lines) must fit two to an SM (227 KB of shared memory per SM, 1 KB reserved per CTA)"""
    assert emu.emu_work_bytes(4200) <= 113 * 1024
    assert emu.emu_work_bytes(1300) <= 37 * 1024    # cfg2's planes: six to an SM


def test_contouring_of_both_line_ends_equals_the_one_at_a_time_statement(emu):
    """k_emit_hits contours both ends of a cut line around the triangle's lone vertex (cutBothPositions); the waypoint pass
    and the oracle-checked replay use interpolateCut, the line-by-line statement of vtkTriangle::Contour: same bits"""
    emu.emu_check_cut_both.restype = C.c_uint64
    emu.emu_check_cut_both.argtypes = [C.c_uint, C.c_uint]
    for seed in range(4):
        assert emu.emu_check_cut_both(seed, 50000) == 0
