"""Face subdivision mesh modifiers: FaceMidpointSubdivision / FaceSubdivisionByArea
(face_subdivision_modifier.cpp).  CPU: the oracle against closed forms and invariants, and the host emulation of the
device launch sequence (tests/cpp/subdivide_emulation.cpp: the same per-element functions as the kernels, visited in a
scrambled order) against the oracle, bit for bit.  GPU: the CUDA path through the C ABI against the oracle, bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from noether_b200 import _capi, meshes

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


# ---------------------------------------------------------------------------------------------------------------------
# clouds: the PCL layouts the path meets, with every byte outside x y z / normal filled so that "byte copy of the first
# source" is visible, and an rgba field for the colour mean
# ---------------------------------------------------------------------------------------------------------------------
LAYOUTS = {
    # name: (point_step, off_xyz, off_normal, off_rgba)
    "PointXYZ": (16, 0, -1, -1),
    "PointNormal": (48, 0, 16, -1),
    "PointXYZRGBNormal": (48, 0, 16, 32),   # rgba at 32, curvature at 36
    "NormalsFirst": (48, 32, 0, -1),        # output of NormalsFromMeshFaces
    "PointXYZRGBA": (32, 0, -1, 16),
    "Packed24": (24, 0, 12, -1),            # step not a multiple of 16: the word-by-word record copy
    "OddRgba": (24, 4, -1, 17),             # rgba not on a word boundary (spills into the padding word)
}


def make_cloud(xyz, layout, seed=0):
    step, ox, on, oc = LAYOUTS[layout]
    rng = np.random.default_rng(seed)
    n = len(xyz)
    rec = rng.integers(0, 256, size=(n, step), dtype=np.uint8)
    rec[:, ox:ox + 12] = np.ascontiguousarray(xyz, np.float32).view(np.uint8).reshape(n, 12)
    if on >= 0:
        nrm = rng.normal(size=(n, 3)).astype(np.float32)
        nrm[rng.random(n) < 0.1] = 0
        rec[:, on:on + 12] = nrm.view(np.uint8).reshape(n, 12)
    return np.ascontiguousarray(rec).reshape(-1)


def case_meshes():
    out = {}
    out["one_triangle"] = (np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.uint32))
    xyz, tri = meshes.wavy(7, 5)
    out["wavy"] = (xyz, tri)
    xyz, tri = meshes.cut_hole(*meshes.wavy(12, 9))
    out["wavy_hole"] = (xyz, tri)
    xyz, tri = meshes.permute(*meshes.wavy(9, 6), seed=3)
    out["wavy_permuted"] = (xyz, tri)
    # degenerate and repeated corners, a duplicated face, both windings of one edge, signed zeros
    xyz = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [-0.0, -0.0, -0.0], [2, 0, 1]], np.float32)
    tri = np.array([[0, 1, 2], [2, 1, 3], [0, 0, 1], [1, 1, 1], [0, 1, 2], [2, 1, 0], [4, 4, 0], [3, 5, 1], [0, 4, 5]], np.uint32)
    out["degenerate"] = (xyz, tri)
    rng = np.random.default_rng(11)
    out["soup"] = (rng.normal(size=(40, 3)).astype(np.float32), rng.integers(0, 40, size=(90, 3)).astype(np.uint32))
    out["no_faces"] = (np.array([[0, 0, 0], [1, 0, 0]], np.float32), np.zeros((0, 3), np.uint32))
    return out


CASES = case_meshes()


def assert_same_mesh(got, ref, what):
    gc, gt = got
    rc, rt = ref
    assert gt.shape == rt.shape, f"{what}: {gt.shape[0]} triangles, oracle {rt.shape[0]}"
    assert gc.size == rc.size, f"{what}: cloud of {gc.size} bytes, oracle {rc.size}"
    assert np.array_equal(gt, rt), f"{what}: triangles differ (first at {np.argwhere(gt != rt)[:3].tolist()})"
    assert np.array_equal(gc, rc), f"{what}: cloud bytes differ (first at byte {int(np.flatnonzero(gc != rc)[0])})"


# ---------------------------------------------------------------------------------------------------------------------
# the oracle against closed forms
# ---------------------------------------------------------------------------------------------------------------------
def test_oracle_midpoint_counts_and_geometry(orc):
    xyz, tri = CASES["wavy"]
    step, ox, on, oc = LAYOUTS["PointNormal"]
    cloud = make_cloud(xyz, "PointNormal")
    for n in (1, 2, 3):
        blob, t = orc.face_midpoint_subdivision(cloud, step, ox, on, oc, tri, n)
        V = blob.size // step
        assert t.shape[0] == 4 ** n * len(tri)
        # Euler: a disc keeps V - E + F = 1; every iteration adds one vertex per edge
        E, F, Vn = 3 * 7 * 5 + 7 + 5, len(tri), len(xyz)
        for _ in range(n):
            Vn, E, F = Vn + E, 2 * E + 3 * F, 4 * F
        assert V == Vn
        assert np.array_equal(blob.reshape(V, step)[:len(xyz)], cloud.reshape(-1, step))   # the input records are kept
        assert len(np.unique(t)) == V                                                      # every vertex is used
        p = np.ndarray((V, 3), np.float32, blob, ox, (step, 4)).astype(np.float64)
        area = lambda P, T: 0.5 * np.linalg.norm(np.cross(P[T[:, 1]] - P[T[:, 0]], P[T[:, 2]] - P[T[:, 0]]), axis=1).sum()
        assert abs(area(p, t) - area(xyz.astype(np.float64), tri)) < 1e-6                  # midpoints lie on the edges


def test_oracle_midpoint_first_source_bytes_and_means(orc):
    xyz, tri = CASES["one_triangle"]
    step, ox, on, oc = LAYOUTS["PointXYZRGBNormal"]
    cloud = make_cloud(xyz, "PointXYZRGBNormal", seed=5)
    blob, t = orc.face_midpoint_subdivision(cloud, step, ox, on, oc, tri, 1)
    rec = blob.reshape(-1, step)
    src = cloud.reshape(-1, step)
    assert t.tolist() == [[0, 3, 5], [3, 1, 4], [5, 4, 2], [3, 4, 5]]
    for new, (a, b) in zip((3, 4, 5), ((0, 1), (1, 2), (2, 0))):
        assert np.array_equal(rec[new, 12:16], src[a, 12:16]) and np.array_equal(rec[new, 36:], src[a, 36:])   # copy of the start
        assert np.array_equal(rec[new, 32:36], ((src[a, 32:36].astype(int) + src[b, 32:36].astype(int)) // 2).astype(np.uint8))
        for off in (ox, on):
            fa, fb = src[a, off:off + 12].view(np.float32), src[b, off:off + 12].view(np.float32)
            assert np.array_equal(rec[new, off:off + 12].view(np.float32), (np.float32(0) + fa + fb) / np.float32(2))


def test_oracle_by_area_order_and_limit(orc):
    xyz, tri = CASES["one_triangle"]
    step, ox, on, oc = LAYOUTS["PointXYZ"]
    cloud = make_cloud(xyz, "PointXYZ")
    blob, t = orc.face_subdivision_by_area(cloud, step, ox, on, oc, tri, 0.1)
    # area 0.5 -> 3 x 0.1667 -> 9 x 0.0556: the stack pops { t0 m t2 } first, then { m t1 t2 }, then { t0 t1 m }
    assert t.tolist() == [[0, 4, 2], [4, 3, 2], [0, 3, 4], [3, 5, 2], [5, 1, 2], [3, 1, 5], [0, 6, 3], [6, 1, 3], [0, 1, 6]]
    p = blob.view(np.float32).reshape(-1, 4)[:, :3]
    assert np.array_equal(p[3], (np.float32(0) + xyz[0] + xyz[1] + xyz[2]) / np.float32(3))
    # nothing to do: faces come out in reverse order (the stack), the cloud is untouched
    xyz2, tri2 = CASES["wavy"]
    c2 = make_cloud(xyz2, "PointXYZ")
    b2, t2 = orc.face_subdivision_by_area(c2, step, ox, on, oc, tri2, 10.0)
    assert np.array_equal(b2, c2) and np.array_equal(t2, tri2[::-1])
    with pytest.raises(orc.OracleError):
        orc.face_subdivision_by_area(cloud, step, ox, on, oc, tri, -1.0, max_points=1000)   # never terminates in the reference
    bad = make_cloud(np.array([[0, 0, 0], [3e38, 0, 0], [0, 3e38, 0]], np.float32), "PointXYZ")
    with pytest.raises(orc.OracleError, match="not finite"):
        orc.face_subdivision_by_area(bad, step, ox, on, oc, tri, 1.0)


# ---------------------------------------------------------------------------------------------------------------------
# host emulation of the device launch sequence == oracle, bit for bit
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emu():
    out = os.path.join(HERE, "cpp", "build")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "libsubdivide_emulation.so")
    src = os.path.join(HERE, "cpp", "subdivide_emulation.cpp")
    deps = [src] + [os.path.join(ROOT, "noether_b200", "csrc", f) for f in ("subdivide_ops.h", "host_math.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", src, "-o", lib],
                       check=True)
    L = C.CDLL(lib)
    L.emu_sub_last_error.restype = C.c_char_p
    common = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, C.c_void_p, C.c_uint64]
    L.emu_midpoint.argtypes = common + [C.c_int]
    L.emu_by_area.argtypes = common + [C.c_float, C.c_uint64]
    L.emu_sub_num_points.argtypes = [C.c_uint32]
    L.emu_sub_num_points.restype = C.c_uint64
    L.emu_sub_num_triangles.restype = C.c_uint64
    L.emu_sub_export.argtypes = [C.c_void_p, C.c_void_p]
    L.emu_sub_export.restype = None
    return L


def emu_run(L, fn, cloud, layout, tri, *args):
    step, ox, on, oc = LAYOUTS[layout]
    tri = np.ascontiguousarray(tri, np.uint32)
    cloud = np.ascontiguousarray(cloud)
    rc = fn(cloud.ctypes.data, cloud.size // step, step, ox, on, oc, tri.ctypes.data, len(tri), *args)
    if rc:
        raise RuntimeError(L.emu_sub_last_error().decode())
    out = np.zeros(L.emu_sub_num_points(step) * step, np.uint8)
    t = np.zeros((L.emu_sub_num_triangles(), 3), np.uint32)
    L.emu_sub_export(out.ctypes.data, t.ctypes.data)
    return out, t


def area_limits(xyz, tri):
    if len(tri) == 0:
        return [1.0]
    p = xyz.astype(np.float64)
    a = 0.5 * np.linalg.norm(np.cross(p[tri[:, 1]] - p[tri[:, 0]], p[tri[:, 2]] - p[tri[:, 0]]), axis=1)
    top = float(a.max()) if a.max() > 0 else 1.0
    return [top * 2, top / 2.5, top / 30, float(np.median(a)) / 7 + top / 500]


@pytest.mark.parametrize("layout", sorted(LAYOUTS))
@pytest.mark.parametrize("case", sorted(CASES))
def test_emulated_midpoint_matches_oracle(emu, orc, case, layout):
    xyz, tri = CASES[case]
    step, ox, on, oc = LAYOUTS[layout]
    cloud = make_cloud(xyz, layout, seed=len(case))
    for n in (1, 2, 3):
        ref = orc.face_midpoint_subdivision(cloud, step, ox, on, oc, tri, n)
        assert_same_mesh(emu_run(emu, emu.emu_midpoint, cloud, layout, tri, n), ref, f"{case} {layout} n={n}")


@pytest.mark.parametrize("layout", sorted(LAYOUTS))
@pytest.mark.parametrize("case", sorted(CASES))
def test_emulated_by_area_matches_oracle(emu, orc, case, layout):
    xyz, tri = CASES[case]
    step, ox, on, oc = LAYOUTS[layout]
    cloud = make_cloud(xyz, layout, seed=len(case))
    for max_area in area_limits(xyz, tri):
        ref = orc.face_subdivision_by_area(cloud, step, ox, on, oc, tri, max_area)
        got = emu_run(emu, emu.emu_by_area, cloud, layout, tri, C.c_float(max_area), 0)
        assert_same_mesh(got, ref, f"{case} {layout} max_area={max_area}")


def test_emulated_random_meshes_match_oracle(emu, orc):
    """ragged random cases: triangle soups over few vertices (many shared and repeated edges), uneven areas"""
    names = sorted(LAYOUTS)
    for seed in range(300):   # the first 60 are the cases the GPU test draws
        rng = np.random.default_rng(7000 + seed)
        V, F = int(rng.integers(3, 30)), int(rng.integers(1, 60))
        xyz = (rng.normal(size=(V, 3)) * rng.choice([1e-3, 1.0, 50.0])).astype(np.float32)
        tri = rng.integers(0, V, size=(F, 3)).astype(np.uint32)
        layout = names[int(rng.integers(len(names)))]
        step, ox, on, oc = LAYOUTS[layout]
        cloud = make_cloud(xyz, layout, seed=seed)
        n = int(rng.integers(1, 4))
        assert_same_mesh(emu_run(emu, emu.emu_midpoint, cloud, layout, tri, n),
                         orc.face_midpoint_subdivision(cloud, step, ox, on, oc, tri, n), f"seed {seed} midpoint n={n}")
        max_area = float(rng.choice(area_limits(xyz, tri)))
        assert_same_mesh(emu_run(emu, emu.emu_by_area, cloud, layout, tri, C.c_float(max_area), 0),
                         orc.face_subdivision_by_area(cloud, step, ox, on, oc, tri, max_area), f"seed {seed} by area {max_area}")


# ---------------------------------------------------------------------------------------------------------------------
# the CUDA path through the C ABI == oracle, bit for bit
# ---------------------------------------------------------------------------------------------------------------------
def gpu_blob(cloud, layout, tri):
    step, ox, on, oc = LAYOUTS[layout]
    return meshes.MeshBlob(np.ascontiguousarray(cloud), cloud.size // step, step, ox, on, np.ascontiguousarray(tri, np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("layout", sorted(LAYOUTS))
def test_gpu_midpoint_matches_oracle(ctx, orc, layout):
    import noether_b200 as nb
    step, ox, on, oc = LAYOUTS[layout]
    for case, (xyz, tri) in sorted(CASES.items()):
        cloud = make_cloud(xyz, layout, seed=len(case))
        for n in (1, 2, 3):
            out = nb.FaceMidpointSubdivisionMeshModifier(n, ctx, off_rgba=oc).modify(gpu_blob(cloud, layout, tri))[0]
            assert_same_mesh((out.data, out.tri), orc.face_midpoint_subdivision(cloud, step, ox, on, oc, tri, n),
                             f"{case} {layout} n={n}")


@pytest.mark.gpu
@pytest.mark.parametrize("layout", sorted(LAYOUTS))
def test_gpu_by_area_matches_oracle(ctx, orc, layout):
    import noether_b200 as nb
    step, ox, on, oc = LAYOUTS[layout]
    for case, (xyz, tri) in sorted(CASES.items()):
        cloud = make_cloud(xyz, layout, seed=len(case))
        for max_area in area_limits(xyz, tri):
            out = nb.FaceSubdivisionByAreaMeshModifier(max_area, ctx, off_rgba=oc).modify(gpu_blob(cloud, layout, tri))[0]
            assert_same_mesh((out.data, out.tri), orc.face_subdivision_by_area(cloud, step, ox, on, oc, tri, max_area),
                             f"{case} {layout} max_area={max_area}")


@pytest.mark.gpu
def test_gpu_random_meshes_match_oracle(ctx, orc):
    import noether_b200 as nb
    names = sorted(LAYOUTS)
    for seed in range(60):
        rng = np.random.default_rng(7000 + seed)
        V, F = int(rng.integers(3, 30)), int(rng.integers(1, 60))
        xyz = (rng.normal(size=(V, 3)) * rng.choice([1e-3, 1.0, 50.0])).astype(np.float32)
        tri = rng.integers(0, V, size=(F, 3)).astype(np.uint32)
        layout = names[int(rng.integers(len(names)))]
        step, ox, on, oc = LAYOUTS[layout]
        cloud = make_cloud(xyz, layout, seed=seed)
        n = int(rng.integers(1, 4))
        out = nb.FaceMidpointSubdivisionMeshModifier(n, ctx, off_rgba=oc).modify(gpu_blob(cloud, layout, tri))[0]
        assert_same_mesh((out.data, out.tri), orc.face_midpoint_subdivision(cloud, step, ox, on, oc, tri, n), f"seed {seed} midpoint")
        max_area = float(rng.choice(area_limits(xyz, tri)))
        out = nb.FaceSubdivisionByAreaMeshModifier(max_area, ctx, off_rgba=oc).modify(gpu_blob(cloud, layout, tri))[0]
        assert_same_mesh((out.data, out.tri), orc.face_subdivision_by_area(cloud, step, ox, on, oc, tri, max_area), f"seed {seed} by area")


@pytest.mark.gpu
def test_gpu_subdivision_large_and_chained(ctx, orc):
    """1.2 M triangles -> 4.8 M (two launches per level over many CTAs), then a second subdivision of the resident result
    and a plan on it without a host round trip; refusals"""
    import noether_b200 as nb
    xyz, tri = meshes.wavy(1000, 600)
    layout = "PointNormal"
    step, ox, on, oc = LAYOUTS[layout]
    cloud = make_cloud(xyz, layout, seed=1)
    blob = gpu_blob(cloud, layout, tri)
    ctx.upload(blob, force=True)
    info = ctx.face_midpoint_subdivision(1)
    ref_c, ref_t = orc.face_midpoint_subdivision(cloud, step, ox, on, oc, tri, 1)
    assert (info.n_points, info.n_triangles) == (ref_c.size // step, len(ref_t))
    assert_same_mesh((ctx.download_cloud(info.n_points), ctx.download_triangles(info.n_triangles)), (ref_c, ref_t), "large midpoint")
    # chained on the resident result: by area, against the oracle applied to the oracle's own output
    a = 0.5 * (1.0 / 1000) * (0.8 / 600) / 4
    info2 = ctx.face_subdivision_by_area(a * 0.9)
    ref2_c, ref2_t = orc.face_subdivision_by_area(ref_c, step, ox, on, oc, ref_t, a * 0.9)
    assert (info2.n_points, info2.n_triangles) == (ref2_c.size // step, len(ref2_t))
    assert_same_mesh((ctx.download_cloud(info2.n_points), ctx.download_triangles(info2.n_triangles)), (ref2_c, ref2_t), "chained by area")
    # the resident subdivided mesh is a mesh like any other: its moments are those of the new cloud
    g = ctx.pca()
    p = np.ndarray((info2.n_points, 3), np.float32, ref2_c, ox, (step, 4)).astype(np.float64)
    assert np.abs(np.array(g.mean) - p.mean(0)).max() < 1e-5
    # refusals
    ctx.upload(gpu_blob(make_cloud(CASES["wavy"][0], "PointXYZ"), "PointXYZ", CASES["wavy"][1]), force=True)
    with pytest.raises(_capi.Nb2Error) as e:
        ctx.face_midpoint_subdivision(0)
    assert e.value.status == _capi.NB2_ERR_INVALID_ARGUMENT
    with pytest.raises(_capi.Nb2Error) as e:
        ctx.face_subdivision_by_area(-1.0, max_points=100000)
    assert e.value.status == _capi.NB2_ERR_INVALID_ARGUMENT
    bad = make_cloud(np.array([[0, 0, 0], [3e38, 0, 0], [0, 3e38, 0]], np.float32), "PointXYZ")
    ctx.upload(gpu_blob(bad, "PointXYZ", np.array([[0, 1, 2]], np.uint32)), force=True)
    with pytest.raises(_capi.Nb2Error, match="not finite") as e:
        ctx.face_subdivision_by_area(1.0)
    assert e.value.status == _capi.NB2_ERR_NON_FINITE


# ---------------------------------------------------------------------------------------------------------------------
# the oracle against an independent restatement (plain Python: recursion + dict, a list used as a stack; numpy float32
# scalars for the arithmetic) written from the reference's text, not from the oracle's source
# ---------------------------------------------------------------------------------------------------------------------
class _PyCloud:
    def __init__(self, cloud, layout):
        self.step, self.ox, self.on, self.oc = LAYOUTS[layout]
        self.rec = [bytes(r) for r in np.asarray(cloud, np.uint8).reshape(-1, self.step)]

    def f3(self, i, off):
        return np.frombuffer(self.rec[i], np.float32, 3, off)

    def append_mean(self, sources):   # copyPointToEnd(first source) + updatePointData
        rec = bytearray(self.rec[sources[0]])
        n = np.float32(len(sources))
        for off in (self.ox, self.on):
            if off < 0:
                continue
            mean = np.zeros(3, np.float32)
            for s in sources:
                mean = mean + self.f3(s, off)
            rec[off:off + 12] = (mean / n).astype(np.float32).tobytes()
        if self.oc >= 0:
            for ch in range(4):
                rec[self.oc + ch] = sum(self.rec[s][self.oc + ch] for s in sources) // len(sources)
        self.rec.append(bytes(rec))
        return len(self.rec) - 1

    def blob(self):
        return np.frombuffer(b"".join(self.rec), np.uint8)


def py_midpoint(cloud, layout, tri, n):
    c, edge, faces = _PyCloud(cloud, layout), {}, []

    def mid(a, b):
        key = (min(a, b), max(a, b))
        if key not in edge:
            edge[key] = c.append_mean([a, b])
        return edge[key]

    def rec(t, depth):
        m0, m1, m2 = mid(t[0], t[1]), mid(t[1], t[2]), mid(t[2], t[0])
        sub = [(t[0], m0, m2), (m0, t[1], m1), (m2, m1, t[2]), (m0, m1, m2)]
        if depth == 1:
            faces.extend(sub)
        else:
            for s in sub:
                rec(s, depth - 1)

    for t in np.asarray(tri).tolist():
        rec(tuple(t), n)
    return c.blob(), np.array(faces, np.uint32).reshape(-1, 3)


def py_by_area(cloud, layout, tri, max_area):
    c, faces = _PyCloud(cloud, layout), []
    stack = [tuple(t) for t in np.asarray(tri).tolist()]
    max_area = np.float32(max_area)
    while stack:
        f = stack.pop()
        v0, v1, v2 = (c.f3(i, c.ox) for i in f)
        a, b = v1 - v0, v2 - v0
        cr = np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], np.float32)
        sq = cr * cr
        area = np.float32(0.5) * np.sqrt(sq[0] + (sq[1] + sq[2]))
        if not area > max_area:
            faces.append(f)
            continue
        m = c.append_mean(list(f))
        stack += [(f[0], f[1], m), (m, f[1], f[2]), (f[0], m, f[2])]
    return c.blob(), np.array(faces, np.uint32).reshape(-1, 3)


@pytest.mark.parametrize("layout", ["PointXYZ", "PointXYZRGBNormal", "OddRgba"])
def test_oracle_matches_independent_python_restatement(orc, layout):
    step, ox, on, oc = LAYOUTS[layout]
    for case in ("one_triangle", "wavy", "degenerate", "soup"):
        xyz, tri = CASES[case]
        cloud = make_cloud(xyz, layout, seed=3)
        for n in (1, 2):
            assert_same_mesh(orc.face_midpoint_subdivision(cloud, step, ox, on, oc, tri, n), py_midpoint(cloud, layout, tri, n),
                             f"{case} {layout} midpoint n={n}")
        for max_area in area_limits(xyz, tri)[:3]:
            with np.errstate(all="ignore"):
                ref = py_by_area(cloud, layout, tri, max_area)
            assert_same_mesh(orc.face_subdivision_by_area(cloud, step, ox, on, oc, tri, max_area), ref,
                             f"{case} {layout} by area {max_area}")
