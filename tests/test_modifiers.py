"""Tool-path modifiers on the device against the oracle restatement of
noether_tpp/src/tool_path_modifiers/*.cpp.

CPU (`not gpu`): the oracle against closed-form expectations and the reference's own fixtures
(test/tool_path_modifier_utest.cpp:120-214: arbitrary / raster-grid / square tool paths, the one-time-modifier idempotence
test :285-304), and the host emulation of the device launch sequence (tests/cpp/modifier_emulation.cpp: same program
builder, same __host__ __device__ arithmetic, kernels replaced by loops) bit for bit against the oracle.
GPU: nb2_result_modify through the C ABI, bit for bit against the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle as orc
from noether_b200 import _capi

from util import oracle_pca_from_capi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module", autouse=True)
def _oracle_built():
    orc.build()
    orc.lib()


# ---------------------------------------------------------------------------------------------------------------------
# fixtures of the reference's modifier tests
# ---------------------------------------------------------------------------------------------------------------------
def _rot(axis, a):
    c, s = np.cos(a), np.sin(a)
    R = np.eye(4)
    i, j = {"x": (1, 2), "y": (2, 0), "z": (0, 1)}[axis]
    R[i, i], R[i, j], R[j, i], R[j, j] = c, -s, s, c
    return R


def _trans(v):
    T = np.eye(4)
    T[:3, 3] = v
    return T


def arbitrary_tool_paths(p, s, w, rng=None):
    """createArbitraryToolPath (tool_path_modifier_utest.cpp:120-136); rng None = identity waypoints"""
    def wp():
        if rng is None:
            return np.eye(4)
        t = rng.uniform(-1, 1, 3)
        r = rng.uniform(-1, 1, 3) * np.pi
        return _rot("x", r[0]) @ _rot("y", r[1]) @ _rot("z", r[2]) @ _trans(t)
    return [[np.stack([wp() for _ in range(w)]) for _ in range(s)] for _ in range(p)]


def raster_grid_tool_paths(p, s, w):
    """createRasterGridToolPath (:138-170)"""
    out = []
    for pi in range(p):
        path = []
        for si in range(s):
            seg = np.tile(np.eye(4), (w, 1, 1))
            seg[:, 0, 3] = si * w + np.arange(w)
            seg[:, 1, 3] = pi
            path.append(seg)
        out.append(path)
    return out


def square_tool_paths(p, w):
    """createSquareToolPaths (:172-214)"""
    scales = np.linspace(1.0, 2.0, p)[::-1]
    out = []
    for pi in range(p):
        off = np.array([scales[pi], 0.0, 0.0])
        path = []
        start = np.eye(4)
        for side in range(4):
            if side:
                start = path[-1][-1] @ _rot("z", -np.pi / 2)
            path.append(np.stack([start @ _trans(off * k) for k in range(w)]))
        out.append(path)
    return out


def ragged_tool_paths(rng, n_paths=7):
    """random poses, ragged segment lengths (>= 2), 1-4 segments per path"""
    out = []
    for _ in range(n_paths):
        path = []
        for _ in range(int(rng.integers(1, 5))):
            n = int(rng.integers(2, 40))
            start = arbitrary_tool_paths(1, 1, 1, rng)[0][0][0]
            step = rng.uniform(-0.1, 0.1, 3)
            seg = []
            for k in range(n):
                T = start @ _trans(step * k) @ _rot("z", rng.uniform(-0.2, 0.2))
                seg.append(T)
            path.append(np.stack(seg))
        out.append(path)
    return out


def all_inputs():
    rng = np.random.default_rng(0)
    return {
        "shuffled grid": shuffled(raster_grid_tool_paths(5, 3, 8), 5),
        "identity": arbitrary_tool_paths(4, 2, 10),
        "random": arbitrary_tool_paths(4, 2, 10, rng),
        "grid": raster_grid_tool_paths(4, 2, 10),
        "square": square_tool_paths(4, 10),
        "ragged": ragged_tool_paths(rng),
        "long": [[np.stack([_rot("y", 0.3) @ _trans([0.01 * k, 0.002 * (k % 7), 0.1 * p]) for k in range(700)])] for p in range(5)],
    }


# name -> (oracle call, C-ABI record); the parameters of the reference's test YAML (:14-96) plus non-trivial ones
def _rec(kind, v=(), n_points=0):
    m = _capi.Modifier()
    m.kind, m.n_points = kind, n_points
    for i, x in enumerate(v):
        m.v[i] = float(x)
    return m


_OFFSET = _rot("x", 0.3) @ _rot("z", -1.1) @ _trans([0.01, -0.02, 0.03])
MODIFIERS = {
    "SnakeOrganization": (lambda p: p.snake_organization(), lambda: _rec(_capi.NB2_MOD_SNAKE_ORGANIZATION)),
    "LinearApproach": (lambda p: p.linear_approach([0.1, 0.2, 0.3], 5),
                       lambda: _rec(_capi.NB2_MOD_LINEAR_APPROACH, [0.1, 0.2, 0.3], 5)),
    "LinearDeparture": (lambda p: p.linear_departure([0.1, 0.2, 0.3], 5),
                        lambda: _rec(_capi.NB2_MOD_LINEAR_DEPARTURE, [0.1, 0.2, 0.3], 5)),
    "LinearApproach1": (lambda p: p.linear_approach([0.0, 0.0, -0.05], 1),
                        lambda: _rec(_capi.NB2_MOD_LINEAR_APPROACH, [0.0, 0.0, -0.05], 1)),
    "OffsetIdentity": (lambda p: p.offset(np.eye(4)), lambda: _rec(_capi.NB2_MOD_OFFSET, np.eye(4).T.reshape(-1))),
    "Offset": (lambda p: p.offset(_OFFSET), lambda: _rec(_capi.NB2_MOD_OFFSET, _OFFSET.T.reshape(-1))),
    "FixedOrientation": (lambda p: p.fixed_orientation([1.0, 0.0, 0.0]),
                         lambda: _rec(_capi.NB2_MOD_FIXED_ORIENTATION, [1.0, 0.0, 0.0])),
    "FixedOrientationOblique": (lambda p: p.fixed_orientation([0.3, -2.0, 0.7]),
                                lambda: _rec(_capi.NB2_MOD_FIXED_ORIENTATION, [0.3, -2.0, 0.7])),
    "UniformOrientation": (lambda p: p.uniform_orientation(), lambda: _rec(_capi.NB2_MOD_UNIFORM_ORIENTATION)),
    "Concatenate": (lambda p: p.concatenate(), lambda: _rec(_capi.NB2_MOD_CONCATENATE)),
    "ToolDragOrientation": (lambda p: p.tool_drag_orientation(0.1745, 0.025),
                            lambda: _rec(_capi.NB2_MOD_TOOL_DRAG_ORIENTATION, [0.1745, 0.025])),
    "BiasedToolDragOrientation": (lambda p: p.biased_tool_drag_orientation(0.1745, 0.025),
                                  lambda: _rec(_capi.NB2_MOD_BIASED_TOOL_DRAG_ORIENTATION, [0.1745, 0.025])),
    "DirectionOfTravelOrientation": (lambda p: p.direction_of_travel(),
                                     lambda: _rec(_capi.NB2_MOD_DIRECTION_OF_TRAVEL_ORIENTATION)),
    "RasterOrganization": (lambda p: p.raster_organization(), lambda: _rec(_capi.NB2_MOD_RASTER_ORGANIZATION)),
    "JoinCloseSegments": (lambda p: p.join_close_segments(1.0),   # the reference's test YAML (:37-38)
                          lambda: _rec(_capi.NB2_MOD_JOIN_CLOSE_SEGMENTS, [1.0])),
    "JoinCloseSegments3": (lambda p: p.join_close_segments(3.0), lambda: _rec(_capi.NB2_MOD_JOIN_CLOSE_SEGMENTS, [3.0])),
    "MinimumSegmentLength": (lambda p: p.minimum_segment_length(1.0),   # the reference's test YAML (:49-50)
                             lambda: _rec(_capi.NB2_MOD_MINIMUM_SEGMENT_LENGTH, [1.0])),
    "MinimumSegmentLength7": (lambda p: p.minimum_segment_length(7.5),
                              lambda: _rec(_capi.NB2_MOD_MINIMUM_SEGMENT_LENGTH, [7.5])),
    "CircularLeadIn": (lambda p: p.circular_lead_in(1.571, 0.1, 5),   # the reference's test YAML (:18-25)
                       lambda: _rec(_capi.NB2_MOD_CIRCULAR_LEAD_IN, [1.571, 0.1], 5)),
    "CircularLeadOut": (lambda p: p.circular_lead_out(1.571, 0.1, 5),
                        lambda: _rec(_capi.NB2_MOD_CIRCULAR_LEAD_OUT, [1.571, 0.1], 5)),
    "CircularLeadIn2": (lambda p: p.circular_lead_in(0.6, 0.03, 2),
                        lambda: _rec(_capi.NB2_MOD_CIRCULAR_LEAD_IN, [0.6, 0.03], 2)),
    "MovingAverageOrientationSmoothing": (lambda p: p.moving_average_orientation_smoothing(3),   # the reference's test YAML
                                          lambda: _rec(_capi.NB2_MOD_MOVING_AVERAGE_ORIENTATION_SMOOTHING, (), 3)),
    "MovingAverageOrientationSmoothing8": (lambda p: p.moving_average_orientation_smoothing(8),
                                           lambda: _rec(_capi.NB2_MOD_MOVING_AVERAGE_ORIENTATION_SMOOTHING, (), 8)),
}
CHAINS = [
    ["LinearApproach", "LinearDeparture"],  # the reference's CompoundToolPathModifier test entry (:86-89)
    ["SnakeOrganization", "Offset", "LinearApproach", "LinearDeparture", "Concatenate"],
    ["DirectionOfTravelOrientation", "UniformOrientation", "ToolDragOrientation", "SnakeOrganization", "LinearApproach1"],
    ["Offset", "BiasedToolDragOrientation", "FixedOrientationOblique", "Concatenate", "SnakeOrganization"],
    ["DirectionOfTravelOrientation", "MovingAverageOrientationSmoothing8", "LinearApproach", "MovingAverageOrientationSmoothing"],
    ["SnakeOrganization", "CircularLeadIn", "CircularLeadOut", "Offset", "CircularLeadIn2", "Concatenate"],
    ["SnakeOrganization", "RasterOrganization", "LinearApproach", "Offset", "RasterOrganization", "SnakeOrganization"],
    ["JoinCloseSegments", "MinimumSegmentLength", "LinearDeparture", "JoinCloseSegments3", "MinimumSegmentLength7", "SnakeOrganization"],
]


def shuffled(nested, seed=0, waypoints=False):
    """shuffle() of the reference's modifier tests (tool_path_modifier_utest.cpp:216-241): segment order inside every tool
    path, the order of tool paths 1..n, and (here, as a snake pattern does) the direction of some segments of the tool
    paths after the first — tool path 0 defines the reference direction"""
    rng = np.random.default_rng(seed)
    out = []
    for k, path in enumerate(nested):
        segs = [seg[::-1].copy() if k and rng.random() < 0.5 else seg.copy() for seg in path]
        rng.shuffle(segs)
        out.append(segs)
    rest = out[1:]
    rng.shuffle(rest)
    return [out[0]] + rest


def test_oracle_raster_organization_restores_the_grid():
    """OrganizationModifiersTest (tool_path_modifier_utest.cpp:361-399): the shuffled raster grid comes back exactly"""
    grid = raster_grid_tool_paths(4, 3, 10)
    f = oracle_apply(shuffled(grid, 3), ["RasterOrganization"])
    np.testing.assert_array_equal(f.pose, np.concatenate([seg for path in grid for seg in path]))


def oracle_apply(nested, names):
    p = orc.Paths.from_nested(nested)
    for n in names:
        MODIFIERS[n][0](p)
    return p.flat()


def records(names):
    recs = [MODIFIERS[n][1]() for n in names]
    return (_capi.Modifier * len(recs))(*recs)


def flatten(nested):
    path_off, seg_off, chunks = [0], [0], []
    for path in nested:
        for seg in path:
            chunks.append(np.asarray(seg, np.float64).transpose(0, 2, 1).reshape(-1, 16))
            seg_off.append(seg_off[-1] + len(seg))
        path_off.append(len(seg_off) - 1)
    poses = np.ascontiguousarray(np.concatenate(chunks)) if chunks else np.zeros((0, 16))
    return np.asarray(path_off, np.uint32), np.asarray(seg_off, np.uint32), poses


# ---------------------------------------------------------------------------------------------------------------------
# oracle: closed-form checks and the reference's own tests
# ---------------------------------------------------------------------------------------------------------------------
def test_oracle_snake_reverses_odd_paths():
    nested = raster_grid_tool_paths(4, 2, 10)
    f = oracle_apply(nested, ["SnakeOrganization"])
    x = f.pose[:, 0, 3].reshape(4, 20)
    np.testing.assert_array_equal(x[0], np.arange(20.0))
    np.testing.assert_array_equal(x[1], np.arange(20.0)[::-1])
    np.testing.assert_array_equal(x[2], np.arange(20.0))
    np.testing.assert_array_equal(x[3], np.arange(20.0)[::-1])


def test_oracle_linear_approach_and_departure_points():
    nested = raster_grid_tool_paths(2, 2, 4)
    off = np.array([0.1, 0.2, 0.3])
    f = oracle_apply(nested, ["LinearApproach"])
    assert len(f.pose) == 2 * 2 * (4 + 5)
    seg = f.pose[:9]
    # farthest point first, the original first waypoint sixth (linear_approach_modifier.cpp:50)
    for j in range(5):
        np.testing.assert_allclose(seg[j][:3, 3], nested[0][0][0][:3, 3] + off * (5 - j) / 5, rtol=0, atol=1e-15)
        np.testing.assert_array_equal(seg[j][:3, :3], np.eye(3))
    np.testing.assert_array_equal(seg[5], nested[0][0][0])
    g = oracle_apply(nested, ["LinearDeparture"])
    seg = g.pose[:9]
    for j in range(5):
        np.testing.assert_allclose(seg[4 + j][:3, 3], nested[0][0][3][:3, 3] + off * (j + 1) / 5, rtol=0, atol=1e-15)


def test_oracle_offset_matches_matrix_product():
    nested = all_inputs()["random"]
    f = oracle_apply(nested, ["Offset"])
    want = np.concatenate([seg for path in nested for seg in path]) @ _OFFSET
    np.testing.assert_allclose(f.pose, want, rtol=0, atol=1e-14)
    np.testing.assert_array_equal(f.pose[:, 3, :], np.tile([0.0, 0.0, 0.0, 1.0], (len(want), 1)))


def test_oracle_tool_drag_is_translate_then_rotate():
    nested = raster_grid_tool_paths(2, 1, 5)
    a, r = 0.1745, 0.025
    f = oracle_apply(nested, ["ToolDragOrientation"])
    # grid paths run along +x with identity orientation: z x dir = y, so sign = +1 and R = Ry(-a)
    want = np.concatenate([seg for path in nested for seg in path]) @ _trans([0, 0, r * np.sin(a)]) @ _rot("y", -a)
    np.testing.assert_allclose(f.pose, want, rtol=0, atol=1e-15)
    g = oracle_apply(nested, ["BiasedToolDragOrientation"])
    want = np.concatenate([seg for path in nested for seg in path]) @ _rot("y", -a) @ _trans([r, 0, 0])
    np.testing.assert_allclose(g.pose, want, rtol=0, atol=1e-15)


def test_oracle_uniform_orientation_flips_half_a_turn():
    nested = raster_grid_tool_paths(2, 1, 4)
    nested[1][0][2] = nested[1][0][2] @ _rot("z", np.pi)  # x axis now points against the travel direction
    f = oracle_apply(nested, ["UniformOrientation"])
    assert (f.pose[:, 0, 0] > 0.99).all()


@pytest.mark.parametrize("name", ["FixedOrientation", "Concatenate", "UniformOrientation", "DirectionOfTravelOrientation"])
def test_oracle_one_time_modifiers_are_idempotent(name):
    """OneTimeToolPathModifierTestFixture (tool_path_modifier_utest.cpp:285-304) for the one-time modifiers built here"""
    for key, nested in all_inputs().items():
        if name in ("DirectionOfTravelOrientation", "UniformOrientation") and key == "identity":
            continue  # coincident waypoints: no direction of travel (the reference's isApprox would compare NaNs)
        p = orc.Paths.from_nested(nested)
        MODIFIERS[name][0](p)
        once = p.flat()
        MODIFIERS[name][0](p)
        twice = p.flat()
        np.testing.assert_array_equal(once.seg_off, twice.seg_off)
        np.testing.assert_allclose(once.pose, twice.pose, rtol=0, atol=1e-12, err_msg=f"{name} on {key}")


def test_oracle_circular_leads_lie_on_the_arc():
    """lead points sit on a circle of arc_radius about anchor + z * r, in the anchor's x-z plane, keep the anchor's
    orientation, and end next to the anchor (circular_lead_in_modifier.cpp:25-40)"""
    nested = raster_grid_tool_paths(2, 1, 6)
    r, a, n = 0.1, 1.571, 5
    f = oracle_apply(nested, ["CircularLeadIn"])
    seg = f.pose[:6 + n]
    centre = nested[0][0][0][:3, 3] + np.array([0.0, 0.0, r])
    for j in range(n):
        np.testing.assert_allclose(np.linalg.norm(seg[j][:3, 3] - centre), r, rtol=0, atol=1e-15)
        assert abs(seg[j][1, 3] - nested[0][0][0][1, 3]) < 1e-15
        np.testing.assert_array_equal(seg[j][:3, :3], np.eye(3))
    # grid paths run along +x with identity orientation: sign = +1; step i sits at angle delta * (i + 1) about +y
    d = a / (n - 1)
    for j in range(n):
        i = n - 1 - j
        want = centre + np.array([-r * np.sin(d * (i + 1)), 0.0, -r * np.cos(d * (i + 1))])
        np.testing.assert_allclose(seg[j][:3, 3], want, rtol=0, atol=1e-15)
    np.testing.assert_array_equal(seg[n], nested[0][0][0])
    g = oracle_apply(nested, ["CircularLeadOut"])
    seg = g.pose[:6 + n]
    centre = nested[0][0][5][:3, 3] + np.array([0.0, 0.0, r])
    for j in range(n):
        want = centre + np.array([r * np.sin(d * (j + 1)), 0.0, -r * np.cos(d * (j + 1))])
        np.testing.assert_allclose(seg[6 + j][:3, 3], want, rtol=0, atol=1e-15)


def test_oracle_moving_average_matches_an_eigen_decomposition():
    """Markley's mean = dominant eigenvector of sum q q^T; the oracle's restated JacobiSVD against numpy.linalg.eigh,
    with the reference's in-place recurrence (a window holds the already smoothed poses before its centre)"""
    def quat(R):
        t = np.trace(R)
        if t > 0:
            t = np.sqrt(t + 1)
            return np.array([(R[2, 1] - R[1, 2]) * 0.5 / t, (R[0, 2] - R[2, 0]) * 0.5 / t, (R[1, 0] - R[0, 1]) * 0.5 / t, 0.5 * t])
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        t = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1)
        q = np.zeros(4)
        q[i], q[3], q[j], q[k] = 0.5 * t, (R[k, j] - R[j, k]) * 0.5 / t, (R[j, i] + R[i, j]) * 0.5 / t, (R[k, i] + R[i, k]) * 0.5 / t
        return q

    def rot(q):
        x, y, z, w = q
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])

    nested = ragged_tool_paths(np.random.default_rng(4), 4)
    for window in (3, 6):
        f = orc.Paths.from_nested(nested).moving_average_orientation_smoothing(window).flat()
        at = 0
        for path in nested:
            for seg in path:
                seg = seg.copy()
                n = len(seg)
                for i in range(n):
                    q = quat(seg[i][:3, :3])
                    if 0 < i < n - 1:
                        lo, hi = max(i - window // 2, 0), min(i + window // 2, n - 1)
                        M = sum(np.outer(quat(seg[k][:3, :3]), quat(seg[k][:3, :3])) for k in range(lo, hi + 1))
                        q = np.linalg.eigh(M)[1][:, -1]
                    seg[i][:3, :3] = rot(q)
                np.testing.assert_allclose(f.pose[at:at + n], seg, rtol=0, atol=1e-13)
                at += n


def test_oracle_refuses_what_the_reference_cannot_do():
    single = [[np.eye(4)[None]]]
    with pytest.raises(Exception):
        oracle_apply(single, ["UniformOrientation"])
    with pytest.raises(Exception):
        oracle_apply(single, ["ToolDragOrientation"])


# ---------------------------------------------------------------------------------------------------------------------
# host emulation of the device launch sequence == oracle, bit for bit
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emu():
    out = os.path.join(HERE, "cpp", "build")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "libmodifier_emulation.so")
    src = [os.path.join(HERE, "cpp", "modifier_emulation.cpp"), os.path.join(ROOT, "noether_b200", "csrc", "modifier_plan.cpp")]
    deps = src + [os.path.join(ROOT, "noether_b200", "csrc", f) for f in ("modifier_ops.h", "modifier_plan.h", "host_math.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                        "-I/usr/local/cuda/include", *src, "-o", lib], check=True)
    L = C.CDLL(lib)
    L.emu_last_error.restype = C.c_char_p
    L.emu_modify.argtypes = [C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                             C.POINTER(_capi.Modifier), C.c_int]
    for n in ("emu_num_paths", "emu_num_segments", "emu_num_waypoints"):
        getattr(L, n).restype = C.c_uint64
    L.emu_export.argtypes = [C.c_void_p] * 3
    L.emu_export.restype = None
    return L


def emu_apply(L, nested, names):
    po, so, poses = flatten(nested)
    st = L.emu_modify(len(po) - 1, po.ctypes.data, len(so) - 1, so.ctypes.data, poses.ctypes.data, None, None, 0,
                      records(names), len(names))
    if st != 0:
        raise _capi.Nb2Error(st, L.emu_last_error().decode())
    P, S, W = L.emu_num_paths(), L.emu_num_segments(), L.emu_num_waypoints()
    opo, oso, op = np.zeros(P + 1, np.uint32), np.zeros(S + 1, np.uint32), np.zeros((W, 16))
    L.emu_export(opo.ctypes.data, oso.ctypes.data, op.ctypes.data)
    return opo, oso, op.reshape(W, 4, 4).transpose(0, 2, 1)


def assert_matches_oracle(got, ref, what):
    po, so, poses = got
    np.testing.assert_array_equal(po.astype(np.uint64), ref.path_off, err_msg=f"{what}: segments per path")
    np.testing.assert_array_equal(so.astype(np.uint64), ref.seg_off, err_msg=f"{what}: waypoints per segment")
    np.testing.assert_array_equal(np.ascontiguousarray(poses).view(np.uint64), np.ascontiguousarray(ref.pose).view(np.uint64),
                                  err_msg=f"{what}: poses (bits)")


@pytest.mark.parametrize("name", sorted(MODIFIERS))
def test_emulated_device_sequence_matches_oracle(emu, name):
    for key, nested in all_inputs().items():
        if name == "DirectionOfTravelOrientation" and key == "identity":
            pass  # zero steps: normalized() leaves zero vectors alone on both sides
        assert_matches_oracle(emu_apply(emu, nested, [name]), oracle_apply(nested, [name]), f"{name} on {key}")


@pytest.mark.parametrize("chain", CHAINS, ids=lambda c: "+".join(c))
def test_emulated_chains_match_oracle(emu, chain):
    for key, nested in all_inputs().items():
        assert_matches_oracle(emu_apply(emu, nested, chain), oracle_apply(nested, chain), f"{chain} on {key}")


def random_case(seed):
    """the `mods` kind of the fuzz campaign: ragged random tool paths, sometimes with a one-waypoint segment, and a random
    chain of 1 - 6 modifiers"""
    names = sorted(MODIFIERS)
    rng = np.random.default_rng(52000 + seed)
    nested = ragged_tool_paths(rng, n_paths=int(rng.integers(1, 12)))
    if rng.random() < 0.2:
        nested[int(rng.integers(len(nested)))].append(nested[0][0][:1].copy())
    chain = [names[int(i)] for i in rng.integers(0, len(names), size=int(rng.integers(1, 7)))]
    return nested, chain


def test_emulated_random_chains_match_oracle(emu):
    """the seeds the GPU campaign draws, replayed through the host emulation: same results and the same refusals"""
    refused = 0
    for seed in range(3000, 3250):
        nested, chain = random_case(seed)
        try:
            ref = oracle_apply(nested, chain)
        except Exception:
            ref = None
        try:
            got = emu_apply(emu, nested, chain)
        except _capi.Nb2Error:
            got = None
        assert (got is None) == (ref is None), f"seed {seed} {chain}: emulation {'refuses' if got is None else 'ok'}, oracle {'refuses' if ref is None else 'ok'}"
        if got is None:
            refused += 1
        else:
            assert_matches_oracle(got, ref, f"seed {seed} {chain}")
    assert 0 < refused < 60


def test_emulated_refusals(emu):
    single = [[np.eye(4)[None]]]
    for name, status in (("UniformOrientation", _capi.NB2_ERR_INVALID_ARGUMENT),
                         ("ToolDragOrientation", _capi.NB2_ERR_TOO_FEW_POINTS),
                         ("DirectionOfTravelOrientation", _capi.NB2_ERR_TOO_FEW_POINTS)):
        with pytest.raises(_capi.Nb2Error) as e:
            emu_apply(emu, single, [name])
        assert e.value.status == status
    with pytest.raises(_capi.Nb2Error):
        emu_apply(emu, [[np.zeros((0, 4, 4)), np.eye(4)[None]]], ["LinearApproach"])


def test_emulated_expansion_of_a_compact_result_matches_the_pose_expansion(emu):
    """k_mod_expand's arithmetic == createTransform + DirectionOfTravel of the oracle's slicer output"""
    from noether_b200 import meshes
    xyz, tri = meshes.wavy(40, 32)
    nrm = orc.normals_from_mesh_faces(xyz, tri)
    for post_ops in (False, True):
        paths, _, _ = orc.plan_plane_slicer(xyz, nrm, tri, 0.05, slice_mode=orc.SLICE_INTENT, post_ops=post_ops)
        f = paths.flat()
        po, so = f.path_off.astype(np.uint32), f.seg_off.astype(np.uint32)
        pos, nn = np.ascontiguousarray(f.pos), np.ascontiguousarray(f.nrm)
        st = emu.emu_modify(len(po) - 1, po.ctypes.data, len(so) - 1, so.ctypes.data, None, pos.ctypes.data, nn.ctypes.data,
                            int(post_ops), records([]), 0)
        assert st == 0
        W = emu.emu_num_waypoints()
        opo, oso, op = np.zeros(len(po), np.uint32), np.zeros(len(so), np.uint32), np.zeros((W, 16))
        emu.emu_export(opo.ctypes.data, oso.ctypes.data, op.ctypes.data)
        np.testing.assert_array_equal(op.reshape(W, 4, 4).transpose(0, 2, 1), f.pose)


# ---------------------------------------------------------------------------------------------------------------------
# the device path through the C ABI == oracle, bit for bit
# ---------------------------------------------------------------------------------------------------------------------
def gpu_apply(ctx, nested, names):
    from noether_b200 import ToolPaths
    tp = ToolPaths.from_nested(nested, ctx)
    h = C.c_void_p()
    _capi.check(ctx._lib.nb2_result_modify(ctx._h, tp._r, records(names), len(names), C.byref(h)))
    out = ToolPaths(ctx, h)
    return out.path_offsets.copy(), out.segment_offsets.copy(), out.poses(), out


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MODIFIERS))
def test_gpu_modifier_matches_oracle(ctx, name):
    for key, nested in all_inputs().items():
        po, so, poses, out = gpu_apply(ctx, nested, [name])
        ref = oracle_apply(nested, [name])
        assert_matches_oracle((po, so, poses), ref, f"{name} on {key}")
        np.testing.assert_array_equal(out.positions, ref.pose[:, :3, 3].astype(np.float32))
        np.testing.assert_array_equal(out.normals, ref.pose[:, :3, 2].astype(np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("chain", CHAINS, ids=lambda c: "+".join(c))
def test_gpu_chains_match_oracle(ctx, chain):
    for key, nested in all_inputs().items():
        po, so, poses, _ = gpu_apply(ctx, nested, chain)
        assert_matches_oracle((po, so, poses), oracle_apply(nested, chain), f"{chain} on {key}")


@pytest.mark.gpu
def test_gpu_refusals(ctx):
    single = [[np.eye(4)[None]]]
    for name, status in (("UniformOrientation", _capi.NB2_ERR_INVALID_ARGUMENT),
                         ("ToolDragOrientation", _capi.NB2_ERR_TOO_FEW_POINTS)):
        with pytest.raises(_capi.Nb2Error) as e:
            gpu_apply(ctx, single, [name])
        assert e.value.status == status


@pytest.mark.gpu
@pytest.mark.parametrize("post_ops", [False, True])
def test_gpu_chain_on_a_plan_matches_oracle(ctx, post_ops):
    """PlaneSlicer plan (compact result) -> chain on the device == oracle plan -> oracle modifiers, bit for bit;
    with an empty chain the device expansion equals nb2_result_poses."""
    from noether_b200 import (CompoundModifier, FixedDirectionGenerator, FixedOriginGenerator, LinearApproachModifier,
                              LinearDepartureModifier, OffsetModifier, PlaneSlicerRasterPlanner, SnakeOrganizationModifier,
                              ToolDragOrientationToolPathModifier, meshes)
    xyz, tri = meshes.cut_hole(*meshes.wavy(160, 128))
    nrm = orc.normals_from_mesh_faces(xyz, tri)
    blob = meshes.pack_blob(xyz, nrm, tri)
    planner = PlaneSlicerRasterPlanner(FixedDirectionGenerator([0.0, 1.0, 0.0]), FixedOriginGenerator([0.0, 0.0, 0.0]), ctx)
    planner.set_line_spacing(0.02)
    planner.raster_post_ops = post_ops
    tp = planner.plan(blob)
    assert tp.post_ops_applied == post_ops
    paths, _, _ = orc.plan_plane_slicer(xyz, nrm, tri, 0.02, direction=[0.0, 1.0, 0.0], origin=[0.0, 0.0, 0.0],
                                        slice_mode=orc.SLICE_INTENT, post_ops=post_ops,
                                        frame=oracle_pca_from_capi(orc, ctx.pca()))
    empty = CompoundModifier([], ctx).modify(tp)
    np.testing.assert_array_equal(empty.poses(), tp.poses())
    chain = CompoundModifier([SnakeOrganizationModifier(), OffsetModifier(_OFFSET), ToolDragOrientationToolPathModifier(0.1745, 0.025),
                              LinearApproachModifier([0.0, 0.0, 0.05], 3), LinearDepartureModifier([0.0, 0.0, 0.05], 3)], ctx)
    got = chain.modify(tp)
    paths.snake_organization().offset(_OFFSET).tool_drag_orientation(0.1745, 0.025)
    paths.linear_approach([0.0, 0.0, 0.05], 3).linear_departure([0.0, 0.0, 0.05], 3)
    ref = paths.flat()
    assert_matches_oracle((got.path_offsets, got.segment_offsets, got.poses()), ref, "plan + chain")
    np.testing.assert_array_equal(got.path_plane, ref.path_plane)
