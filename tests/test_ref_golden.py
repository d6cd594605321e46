"""Golden vectors written by the UNMODIFIED reference (oracle/ref_golden/dump_golden.cpp run inside the reference's own
docker image by oracle/ref_golden/generate.sh) against the oracle and against the CUDA path.

The development image of this repository cannot build the reference (no VTK / PCL / Eigen / Boost / yaml-cpp, no
network), so tests/golden/ref/*.bin do not exist yet and every test here SKIPS with that reason: until they do, the
oracle's parity stays "unpinned" (DESIGN.md §5).  The moment the files are generated and committed these tests pin

  * DirectionGenerator / OriginGenerator outputs (PCA axes, centroid)            within 1e-6 x bounding-box diagonal
  * the number of tool paths and, in the oracle's FAITHFUL mode, the segmentation the reference really returns
  * every waypoint position (1e-6 x diagonal) and orientation (1e-5 rad)
  * for the CUDA path (-m gpu): the same, against the reference's segments re-joined into maximal chains
"""
import glob
import os
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ref", "*.bin")))
REASON = ("tests/golden/ref/*.bin not generated yet: run oracle/ref_golden/generate.sh where docker and the network are "
          "available (the reference cannot be built in this image) — parity unpinned until then")
needs_golden = pytest.mark.skipif(not GOLDEN, reason=REASON)
_DTYPES = {0: np.float64, 1: np.float32, 2: np.uint32}


def load_golden(path):
    """{name: array} from the NB2GOLD1 container dump_golden.cpp writes"""
    out = {}
    with open(path, "rb") as f:
        assert f.read(8) == b"NB2GOLD1", path
        while True:
            head = f.read(4)
            if not head:
                break
            (n,) = struct.unpack("<I", head)
            name = f.read(n).decode()
            dtype, ndim = struct.unpack("<II", f.read(8))
            dims = struct.unpack(f"<{ndim}Q", f.read(8 * ndim))
            dt = np.dtype(_DTYPES[dtype])
            count = int(np.prod(dims)) if ndim else 1
            out[name] = np.frombuffer(f.read(count * dt.itemsize), dt).reshape(dims).copy()
    return out


def test_container_round_trip(tmp_path):
    """the loader reads what the dumper's Writer writes (format restated here byte for byte)"""
    p = tmp_path / "x.bin"
    xyz = np.arange(12, dtype=np.float32).reshape(4, 3)
    sizes = np.array([2, 1], np.uint32)
    with open(p, "wb") as f:
        f.write(b"NB2GOLD1")
        for name, code, a in (("xyz", 1, xyz), ("path_sizes", 2, sizes), ("direction", 0, np.array([1.0, 0, 0]))):
            f.write(struct.pack("<I", len(name)) + name.encode() + struct.pack("<II", code, a.ndim))
            f.write(struct.pack(f"<{a.ndim}Q", *a.shape) + a.tobytes())
    g = load_golden(str(p))
    assert np.array_equal(g["xyz"], xyz) and np.array_equal(g["path_sizes"], sizes) and g["direction"].dtype == np.float64


def _angles(Ra, Rb):
    """rotation angle between corresponding 3x3 rotations"""
    tr = np.einsum("nij,nij->n", Ra, Rb)
    return np.arccos(np.clip((tr - 1.0) / 2.0, -1.0, 1.0))


def _poses(g):
    return g["poses"].reshape(-1, 4, 4).transpose(0, 2, 1)  # column-major storage -> (row, col)


def _oracle_plan(orc, g, slice_mode):
    prm = g["params"]
    kw = dict(direction=g["direction"], origin=g["origin"], slice_mode=slice_mode, scan_mode=orc.SCAN_ALL_CELLS,
              moments_mode=orc.MOMENTS_F32_SEQUENTIAL)
    if prm[2]:
        return orc.plan_plane_slicer_legacy(g["xyz"], g["normals"], g["triangles"], prm[0], prm[3], prm[4], prm[5], bool(prm[1]), **kw)[0]
    return orc.plan_plane_slicer(g["xyz"], g["normals"], g["triangles"], prm[0], bool(prm[1]), **kw)[0]


@needs_golden
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_generators_match_the_reference(path):
    import oracle as orc
    orc.build()
    g = load_golden(path)
    diag = float(np.linalg.norm(g["xyz"].max(0) - g["xyz"].min(0)))
    name = os.path.basename(path)
    if "principal_axis" in name or "wavy" in name:
        p = orc.pca(g["xyz"], orc.MOMENTS_F32_SEQUENTIAL)
        d = orc.principal_axis_direction(p, 0.0)
        assert abs(float(d @ g["direction"])) > 1 - 1e-6, "PrincipalAxis direction (up to the solver's sign)"
        assert float(d @ g["direction"]) > 0, "sign of the principal axis differs from the reference's Eigen build"
    if "fixed_1p05" not in name and "flat_square" not in name and "legacy" not in name:
        assert np.abs(orc.centroid(g["xyz"], orc.MOMENTS_F32_SEQUENTIAL) - g["origin"]).max() <= 1e-6 * diag


@needs_golden
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_faithful_plan_matches_the_reference(path):
    """structure exactly, coordinates within the north-star tolerances (and reported when bit-exact)"""
    import oracle as orc
    orc.build()
    g = load_golden(path)
    f = _oracle_plan(orc, g, orc.SLICE_FAITHFUL).flat()
    assert np.array_equal(np.diff(f.path_off).astype(np.uint32), g["path_sizes"])
    assert np.array_equal(np.diff(f.seg_off).astype(np.uint32), g["segment_sizes"])
    ref = _poses(g)
    diag = float(np.linalg.norm(g["xyz"].max(0) - g["xyz"].min(0)))
    assert np.abs(f.pose[:, :3, 3] - ref[:, :3, 3]).max(initial=0.0) <= 1e-6 * diag
    assert _angles(f.pose[:, :3, :3], ref[:, :3, :3]).max(initial=0.0) <= 1e-5
    print(os.path.basename(path), "bit-exact poses:", bool(np.array_equal(f.pose, ref)))


def _rejoin(path_sizes, segment_sizes, poses):
    """the reference's segments of each tool path re-joined into maximal chains: consecutive pieces that share an end
    point (bit-equal translation) are one chain (what SLICE_INTENT and the CUDA path produce); returns per path a sorted
    list of chains as position arrays, each in its lexicographically smaller direction"""
    out, s, w = [], 0, 0
    for n_seg in path_sizes:
        pieces = []
        for _ in range(int(n_seg)):
            n = int(segment_sizes[s])
            pieces.append(poses[w:w + n, :3, 3].copy())
            s, w = s + 1, w + n
        changed = True
        while changed:
            changed = False
            for i in range(len(pieces)):
                for j in range(len(pieces)):
                    if i == j:
                        continue
                    a, b = pieces[i], pieces[j]
                    for aa in (a, a[::-1]):
                        for bb in (b, b[::-1]):
                            if not changed and np.array_equal(aa[-1], bb[0]):
                                pieces[i] = np.concatenate([aa, bb[1:]])
                                pieces.pop(j)
                                changed = True
                    if changed:
                        break
                if changed:
                    break
        canon = [min(p.astype(np.float32).tobytes(), p[::-1].astype(np.float32).tobytes()) for p in pieces]
        out.append(sorted(canon))
    return out


@needs_golden
@pytest.mark.gpu
@pytest.mark.parametrize("path", [p for p in GOLDEN if "legacy" not in p] or ["none"],
                         ids=[os.path.basename(p) for p in GOLDEN if "legacy" not in p] or ["none"])
def test_cuda_plan_matches_the_reference(path):
    """the CUDA path (explicit direction / origin as the reference's generators returned them) against the reference's
    tool paths re-joined into maximal chains: same chains, point for point (float32 positions bit for bit)"""
    import noether_b200 as nb
    from noether_b200 import meshes
    g = load_golden(path)
    ctx = nb.Context(0)
    try:
        blob = meshes.pack_blob(g["xyz"], g["normals"], g["triangles"])
        cfg = {"name": "PlaneSlicer", "line_spacing": float(g["params"][0]), "bidirectional": bool(g["params"][1]),
               "direction_generator": {"name": "FixedDirection", "direction": [float(x) for x in g["direction"]]},
               "origin_generator": {"name": "FixedOrigin", "origin": [float(x) for x in g["origin"]]}}
        gpu = nb.Factory(ctx).create_tool_path_planner(cfg).plan(blob)
        ref = _rejoin(g["path_sizes"], g["segment_sizes"], _poses(g))
        assert gpu.n_paths == len(ref)
        gp = gpu.poses()
        got = _rejoin(np.diff(gpu.path_offsets), np.diff(gpu.segment_offsets), gp)
        assert got == ref
    finally:
        ctx.close()
