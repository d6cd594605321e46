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


_PLY_NP = {"char": "i1", "uchar": "u1", "short": "<i2", "ushort": "<u2", "int": "<i4", "uint": "<u4", "float": "<f4",
           "double": "<f8", "int8": "i1", "uint8": "u1", "int16": "<i2", "uint16": "<u2", "int32": "<i4",
           "uint32": "<u4", "float32": "<f4", "float64": "<f8"}


def write_ply(xyz, tri, normals=None, vertex_props=None, count_type="uchar", index_type="int", face_pre=(), face_post=(),
              comments=(), fmt="binary_little_endian", faces=None, extra_elements_before=()):
    """A PLY image (bytes).  vertex_props: ordered [(type, name)]; names x y z nx ny nz take the mesh data, anything else
    a deterministic filler.  face_pre / face_post: [(type, name)] scalar properties around the vertex_indices list.
    faces: explicit list of index lists (polygons of any size) instead of `tri`.  extra_elements_before: [(name, count,
    [(type, name)])] scalar-only elements stored ahead of the vertex element."""
    xyz = np.asarray(xyz, np.float32)
    n = len(xyz)
    if vertex_props is None:
        vertex_props = [("float", a) for a in "xyz"] + ([("float", a) for a in ("nx", "ny", "nz")] if normals is not None else [])
    polys = [list(map(int, f)) for f in (faces if faces is not None else np.asarray(tri).reshape(-1, 3))]
    head = ["ply", f"format {fmt} 1.0"] + [f"comment {c}" for c in comments]
    for name, count, props in extra_elements_before:
        head.append(f"element {name} {count}")
        head += [f"property {t} {p}" for t, p in props]
    head.append(f"element vertex {n}")
    head += [f"property {t} {p}" for t, p in vertex_props]
    head.append(f"element face {len(polys)}")
    head += [f"property {t} {p}" for t, p in face_pre]
    head.append(f"property list {count_type} {index_type} vertex_indices")
    head += [f"property {t} {p}" for t, p in face_post]
    head.append("end_header")
    out = bytearray(("\n".join(head) + "\n").encode("ascii"))
    src = {"x": xyz[:, 0], "y": xyz[:, 1], "z": xyz[:, 2]}
    if normals is not None:
        nn = np.asarray(normals, np.float32)
        src.update({"nx": nn[:, 0], "ny": nn[:, 1], "nz": nn[:, 2]})

    def filler(k, count, t):
        return ((np.arange(count) * (7 + k)) % 113).astype(_PLY_NP[t])

    if fmt == "ascii":
        lines = []
        for name, count, props in extra_elements_before:
            cols = [filler(k, count, t) for k, (t, _) in enumerate(props)]
            lines += [" ".join(repr(c[i].item()) for c in cols) for i in range(count)]
        cols = [src[p].astype(_PLY_NP[t]) if p in src else filler(k, n, t) for k, (t, p) in enumerate(vertex_props)]
        lines += [" ".join(repr(float(c[i])) if c.dtype.kind == "f" else str(int(c[i])) for c in cols) for i in range(n)]
        for i, f in enumerate(polys):
            pre = [str(int(filler(k, len(polys), t)[i])) for k, (t, _) in enumerate(face_pre)]
            post = [str(int(filler(k, len(polys), t)[i])) for k, (t, _) in enumerate(face_post)]
            lines.append(" ".join(pre + [str(len(f))] + [str(v) for v in f] + post))
        return bytes(out) + ("\n".join(lines) + "\n").encode("ascii")
    swap = fmt == "binary_big_endian"

    def packed(arr, t):
        a = np.asarray(arr).astype(_PLY_NP[t])
        return a.byteswap().tobytes() if swap else a.tobytes()

    for name, count, props in extra_elements_before:
        dt = np.dtype([(p, _PLY_NP[t]) for t, p in props])
        rec = np.zeros(count, dt)
        for k, (t, p) in enumerate(props):
            rec[p] = filler(k, count, t)
        out += rec.byteswap().tobytes() if swap else rec.tobytes()
    dt = np.dtype([(p, _PLY_NP[t]) for t, p in vertex_props])
    rec = np.zeros(n, dt)
    for k, (t, p) in enumerate(vertex_props):
        rec[p] = src[p].astype(_PLY_NP[t]) if p in src else filler(k, n, t)
    out += rec.byteswap().tobytes() if swap else rec.tobytes()
    uniform = len({len(f) for f in polys}) <= 1 and polys
    if uniform:
        k = len(polys[0])
        fields = [(p, _PLY_NP[t]) for t, p in face_pre] + [("_n", _PLY_NP[count_type]), ("_i", _PLY_NP[index_type], (k,))] + \
                 [(p, _PLY_NP[t]) for t, p in face_post]
        rec = np.zeros(len(polys), np.dtype(fields))
        for j, (t, p) in enumerate(face_pre):
            rec[p] = filler(j, len(polys), t)
        for j, (t, p) in enumerate(face_post):
            rec[p] = filler(j, len(polys), t)
        rec["_n"] = k
        rec["_i"] = np.asarray(polys).astype(_PLY_NP[index_type])
        out += rec.byteswap().tobytes() if swap else rec.tobytes()
    else:
        for i, f in enumerate(polys):
            for j, (t, _) in enumerate(face_pre):
                out += packed(filler(j, len(polys), t)[i:i + 1], t)
            out += packed([len(f)], count_type) + packed(f, index_type)
            for j, (t, _) in enumerate(face_post):
                out += packed(filler(j, len(polys), t)[i:i + 1], t)
    return bytes(out)
