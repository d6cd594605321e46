"""Mesh ingest from a binary STL image: the oracle's reader (vtkSTLReader with point merging + vtk2mesh
restated) against hand-made files and closed forms, the host-only header handling of the C ABI, and — on the GPU — the
welding kernels bit for bit against the oracle."""
import os
import struct

import numpy as np
import pytest

import noether_b200 as nb
from noether_b200 import _capi, meshes
from noether_b200.planner import stl_inspect


def write_stl(xyz, tri, header=b"noether_b200 test", count=None, tail=b""):
    """triangle soup: every facet stores its three corners (the facet normal is what a CAD exporter would write)"""
    p = xyz[tri]  # (F, 3, 3)
    n = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]).astype(np.float32)
    rec = np.zeros((len(tri), 50), np.uint8)
    rec[:, 0:12] = n.view(np.uint8).reshape(-1, 12)
    rec[:, 12:48] = np.ascontiguousarray(p, np.float32).view(np.uint8).reshape(-1, 36)
    rec[:, 48:50] = 0
    return header.ljust(80, b"\0")[:80] + struct.pack("<I", len(tri) if count is None else count) + rec.tobytes() + tail


def _soup(nx=9, ny=7):
    return meshes.wavy(nx, ny)


def test_oracle_welds_in_first_occurrence_order(orc):
    xyz, tri = _soup()
    oxyz, otri = orc.read_stl(write_stl(xyz, tri))
    # every vertex of the grid is used, so welding recovers the vertex set; ids count first occurrences in facet order
    assert len(oxyz) == len(xyz) and len(otri) == len(tri)
    order = []
    seen = set()
    for v in tri.reshape(-1):
        if int(v) not in seen:
            seen.add(int(v))
            order.append(int(v))
    np.testing.assert_array_equal(oxyz.view(np.uint32), xyz[order].view(np.uint32))
    remap = np.empty(len(xyz), np.uint32)
    remap[order] = np.arange(len(order), dtype=np.uint32)
    np.testing.assert_array_equal(otri, remap[tri])


def test_oracle_signed_zero_and_degenerate_facets(orc):
    xyz = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [-0.0, 0.0, -0.0], [1.0, 1.0, 0.0]], np.float32)
    tri = np.array([[0, 1, 2], [3, 1, 4], [0, 3, 1], [2, 4, 1]], np.uint32)
    oxyz, otri = orc.read_stl(write_stl(xyz, tri))
    # vertex 3 (-0, 0, -0) merges into vertex 0 and keeps the first representation; facet 2 collapses and is dropped
    assert len(oxyz) == 4 and not np.signbit(oxyz[0]).any()
    np.testing.assert_array_equal(otri, [[0, 1, 2], [0, 1, 3], [2, 3, 1]])


def test_oracle_reads_until_the_file_ends(orc):
    xyz, tri = _soup(4, 3)
    good = write_stl(xyz, tri)
    a = orc.read_stl(good)
    for raw in (write_stl(xyz, tri, count=7), write_stl(xyz, tri, count=10**6), good[:-2], good + b"\0" * 17):
        b = orc.read_stl(raw)  # bogus counts are ignored, a cut-off attribute word is tolerated, trailing junk < 48 B too
        np.testing.assert_array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
        np.testing.assert_array_equal(a[1], b[1])
    assert len(orc.read_stl(good[:-3])[1]) == len(tri) - 1  # 47 bytes of a facet: not read


def test_inspect_and_refusals(orc):
    xyz, tri = _soup(4, 3)
    info = stl_inspect(write_stl(xyz, tri))
    assert (info.n_triangles, info.n_points, info.has_normals) == (len(tri), 3 * len(tri), 0)
    assert stl_inspect(write_stl(xyz, tri, count=3)).n_triangles == len(tri)
    assert stl_inspect(write_stl(xyz, tri)[:-3]).n_triangles == len(tri) - 1
    for bad, what in [(b"solid part\nfacet normal 0 0 1\n outer loop\n" + b" " * 80, "ASCII"), (b"\0" * 50, "shorter")]:
        with pytest.raises(_capi.Nb2Error) as e:
            stl_inspect(bad)
        assert e.value.status == _capi.NB2_ERR_INVALID_ARGUMENT and what in str(e.value)
        with pytest.raises(Exception):
            orc.read_stl(bad)
    # a binary file whose header happens to start with "solid" (some exporters do that) is still binary
    assert stl_inspect(write_stl(xyz, tri, header=b"solid exported by a CAD tool")).n_triangles == len(tri)


# ---------------------------------------------------------------------------------------------------------------------
# device
# ---------------------------------------------------------------------------------------------------------------------
def _check_against_oracle(ctx, orc, raw):
    oxyz, otri = orc.read_stl(raw)
    info = ctx.upload_stl(raw)
    assert (info.n_points, info.n_triangles, info.has_normals) == (len(oxyz), len(otri), 0)
    gxyz, _ = ctx.download(len(oxyz))
    np.testing.assert_array_equal(gxyz.view(np.uint32), oxyz.view(np.uint32))
    np.testing.assert_array_equal(ctx.download_triangles(len(otri)), otri)
    return oxyz, otri


@pytest.mark.gpu
def test_device_weld_matches_oracle(ctx, orc):
    xyz, tri = _soup()
    _check_against_oracle(ctx, orc, write_stl(xyz, tri))
    # permuted facet order and vertex order: first-occurrence numbering changes, still identical to the sequential reader
    pxyz, ptri = meshes.permute(xyz, tri, seed=3)
    _check_against_oracle(ctx, orc, write_stl(pxyz, ptri))
    # bogus count, cut-off attribute word, trailing junk
    _check_against_oracle(ctx, orc, write_stl(xyz, tri, count=5))
    _check_against_oracle(ctx, orc, write_stl(xyz, tri)[:-2])
    _check_against_oracle(ctx, orc, write_stl(xyz, tri, tail=b"\1" * 31))


@pytest.mark.gpu
def test_device_weld_signed_zero_duplicates_and_degenerates(ctx, orc):
    xyz = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [-0.0, 0.0, -0.0], [1.0, 1.0, 0.0]], np.float32)
    tri = np.array([[0, 1, 2], [3, 1, 4], [0, 3, 1], [2, 4, 1], [0, 1, 2], [4, 4, 4]], np.uint32)
    oxyz, otri = _check_against_oracle(ctx, orc, write_stl(xyz, tri))
    assert len(oxyz) == 4 and len(otri) == 4
    # many coincident corners on few points: heavy contention on a handful of slots
    rng = np.random.default_rng(11)
    few = rng.normal(size=(13, 3)).astype(np.float32)
    t = rng.integers(0, 13, size=(5000, 3)).astype(np.uint32)
    _check_against_oracle(ctx, orc, write_stl(few, t))


@pytest.mark.gpu
def test_plan_from_stl_equals_oracle(ctx, orc, tmp_path):
    """loadPolygonFile(.stl) -> NormalsFromMeshFaces -> PlaneSlicer, file decoded and welded on the device, against the
    oracle on the oracle-read mesh: bit for bit."""
    from util import assert_same_paths, oracle_pca_from_capi
    xyz, tri = meshes.cut_hole(*meshes.wavy(120, 96))
    raw = write_stl(xyz, tri)
    path = tmp_path / "part.STL"
    path.write_bytes(raw)
    oxyz, otri = orc.read_stl(raw)
    info = ctx.upload_file(path)
    assert (info.n_points, info.n_triangles) == (len(oxyz), len(otri))
    got_n = ctx.normals_from_mesh_faces(len(oxyz))
    onrm = orc.normals_from_mesh_faces(oxyz, otri)
    np.testing.assert_array_equal(got_n.view(np.uint32), onrm.view(np.uint32))
    cfg = {"name": "PlaneSlicer", "line_spacing": 0.03, "bidirectional": True,
           "direction_generator": {"name": "PrincipalAxis", "rotation_offset": 0.0},
           "origin_generator": {"name": "Centroid"}}
    dev = nb.Factory(ctx).create_tool_path_planner(cfg).plan_resident()
    ref, _, _ = orc.plan_plane_slicer(oxyz, onrm, otri, 0.03, True, slice_mode=orc.SLICE_INTENT,
                                      frame=oracle_pca_from_capi(orc, ctx.pca()))
    assert dev.n_paths > 20
    assert_same_paths(dev, ref.flat(), "plan from an STL file")


@pytest.mark.gpu
def test_device_refusals_and_large_file(ctx, orc):
    xyz, tri = _soup(4, 3)
    nan = xyz.copy()
    nan[2, 0] = np.inf
    with pytest.raises(_capi.Nb2Error) as e:
        ctx.upload_stl(write_stl(nan, tri))
    assert e.value.status == _capi.NB2_ERR_NON_FINITE
    with pytest.raises(_capi.Nb2Error):
        ctx.upload_stl(b"solid x\n" + b" " * 200)
    # 1 M facets: welded mesh identical to the sequential reader's
    big_xyz, big_tri = meshes.wavy(1000, 500)
    _check_against_oracle(ctx, orc, write_stl(big_xyz, big_tri))
