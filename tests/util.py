"""Shared helpers for the parity tests."""
from __future__ import annotations

import ctypes as C

import numpy as np


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_same_paths(gpu, ref, what="", poses=True, exact=True):
    """GPU ToolPaths (noether_b200.ToolPaths) against oracle FlatPaths: structure, connectivity, coordinates.

    Connectivity (path/segment structure + the ordered mesh-edge sequence of every segment) must be bit-exact.
    Positions / normals are compared bit-exactly too when `exact`, else within the north-star tolerances
    (1e-6 x bbox diagonal, 1e-5 rad).
    """
    assert gpu.n_paths == ref.num_paths, f"{what}: tool path count {gpu.n_paths} != {ref.num_paths}"
    np.testing.assert_array_equal(gpu.path_offsets.astype(np.uint64), ref.path_off, err_msg=f"{what}: segments per path")
    np.testing.assert_array_equal(gpu.segment_offsets.astype(np.uint64), ref.seg_off, err_msg=f"{what}: waypoints per segment")
    np.testing.assert_array_equal(gpu.path_plane, ref.path_plane, err_msg=f"{what}: lattice index of paths")
    if gpu.edge_lo is not None:
        np.testing.assert_array_equal(gpu.edge_lo, ref.edge_lo, err_msg=f"{what}: edge sequence (lo)")
        np.testing.assert_array_equal(gpu.edge_hi, ref.edge_hi, err_msg=f"{what}: edge sequence (hi)")
    if exact:
        np.testing.assert_array_equal(bits(gpu.positions), bits(ref.pos), err_msg=f"{what}: positions (bits)")
        np.testing.assert_array_equal(bits(gpu.normals), bits(ref.nrm), err_msg=f"{what}: normals (bits)")
    else:
        diag = float(np.linalg.norm(ref.pos.max(0) - ref.pos.min(0))) if len(ref.pos) else 1.0
        assert np.abs(gpu.positions - ref.pos).max(initial=0.0) <= 1e-6 * diag
        a = gpu.normals / np.maximum(np.linalg.norm(gpu.normals, axis=1, keepdims=True), 1e-30)
        b = ref.nrm / np.maximum(np.linalg.norm(ref.nrm, axis=1, keepdims=True), 1e-30)
        ang = np.arccos(np.clip((a * b).sum(1), -1, 1))
        assert ang.max(initial=0.0) <= 1e-5
    if poses and gpu.n_waypoints:
        gp = gpu.poses()
        np.testing.assert_array_equal(gp, ref.pose, err_msg=f"{what}: 4x4 poses")


def capi_lattice_from_oracle(lat):
    """oracle.Lattice -> noether_b200._capi.Lattice (same field layout, distinct ctypes classes)"""
    from noether_b200 import _capi
    out = _capi.Lattice()
    C.memmove(C.byref(out), C.byref(lat), C.sizeof(out))
    return out


def oracle_pca_from_capi(orc, p):
    """noether_b200._capi.Pca -> oracle Pca (mean, evecs, scales)"""
    ev = np.array(p.evecs, dtype=np.float32).reshape(3, 3).T
    return orc.pca_from_parts(np.array(p.mean, np.float32), ev, np.array(p.scales, np.float32))


def canonical_segments(flat):
    """Connectivity canonicalised for direction: per path a sorted list of edge sequences (each sequence oriented so
    that it is lexicographically <= its reverse)."""
    out = []
    for p in range(len(flat.path_off) - 1):
        segs = []
        for s in range(int(flat.path_off[p]), int(flat.path_off[p + 1])):
            b, e = int(flat.seg_off[s]), int(flat.seg_off[s + 1])
            seq = [tuple(sorted((int(flat.edge_lo[i]), int(flat.edge_hi[i])))) for i in range(b, e)]
            rev = seq[::-1]
            segs.append(min(seq, rev))
        out.append(sorted(segs))
    return out
