"""Mesh containers and synthetic inputs for the raster-slicing path.

`MeshBlob` mirrors the part of ``pcl::PolygonMesh`` the path reads: an interleaved ``PCLPointCloud2`` byte blob whose
``x,y,z`` / ``normal_x,normal_y,normal_z`` float32 fields are located by offset (noether_tpp/src/utils.cpp:74-117) and
a flat ``uint32[3F]`` triangle list (``pcl::Vertices`` flattened).  The synthetic generators are the ones SURVEY.md §8(d)
defines for the benchmark configurations.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class MeshBlob:
    data: np.ndarray        # (V * point_step,) uint8 — PCLPointCloud2::data
    n_vertices: int
    point_step: int
    off_xyz: int            # byte offset of field "x" (y, z follow contiguously)
    off_normal: int         # byte offset of field "normal_x", or -1 when the cloud has no normals
    tri: np.ndarray         # (F, 3) uint32

    @property
    def n_triangles(self) -> int:
        return int(self.tri.shape[0])

    def xyz(self) -> np.ndarray:
        return _field3(self.data, self.n_vertices, self.point_step, self.off_xyz)

    def normals(self) -> np.ndarray | None:
        if self.off_normal < 0:
            return None
        return _field3(self.data, self.n_vertices, self.point_step, self.off_normal)


def _field3(data, n, step, off):
    rec = np.ndarray(shape=(n, 3), dtype=np.float32, buffer=data, offset=off, strides=(step, 4))
    return np.ascontiguousarray(rec)


def pack_blob(xyz, normals, tri, layout: str = "PointNormal") -> MeshBlob:
    """Interleave into one of the PCL layouts the path meets.

    PointXYZ: step 16 (x0 y4 z8); PointNormal / PointXYZRGBNormal: step 48 (x0.., normal_x16..);
    NormalsFirst: output of NormalsFromMeshFaces = pcl::Normal (32 B: normal_x0.. curvature16) ++ PointXYZ (x32..).
    """
    xyz = np.ascontiguousarray(xyz, np.float32)
    n = xyz.shape[0]
    tri = np.ascontiguousarray(tri, np.uint32).reshape(-1, 3)
    if layout == "PointXYZ":
        step, ox, on = 16, 0, -1
    elif layout in ("PointNormal", "PointXYZRGBNormal"):
        step, ox, on = 48, 0, 16
    elif layout == "NormalsFirst":
        step, ox, on = 48, 32, 0
    elif layout == "Packed24":
        step, ox, on = 24, 0, 12
    else:
        raise ValueError(layout)
    data = np.zeros(n * step, np.uint8)
    view = np.ndarray(shape=(n, 3), dtype=np.float32, buffer=data, offset=ox, strides=(step, 4))
    view[:] = xyz
    if on >= 0:
        if normals is None:
            raise ValueError("layout needs normals")
        nv = np.ndarray(shape=(n, 3), dtype=np.float32, buffer=data, offset=on, strides=(step, 4))
        nv[:] = np.asarray(normals, np.float32)
    return MeshBlob(data, n, step, ox, on, tri)


def read_ply(path: str) -> MeshBlob:
    """Minimal binary-little-endian PLY reader (vertex float properties + uchar colours, face lists of int)."""
    with open(path, "rb") as f:
        raw = f.read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    header = raw[:end].decode("ascii").splitlines()
    if "format binary_little_endian 1.0" not in header:
        raise ValueError("only binary_little_endian PLY is supported")
    types = {"float": "<f4", "float32": "<f4", "uchar": "u1", "uint8": "u1", "int": "<i4", "int32": "<i4",
             "uint": "<u4", "double": "<f8", "short": "<i2", "ushort": "<u2", "char": "i1"}
    elems = []
    for line in header:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "element":
            elems.append([tok[1], int(tok[2]), []])
        elif tok[0] == "property":
            elems[-1][2].append(tok[1:])
    pos = end
    xyz = nrm = tri = None
    for name, count, props in elems:
        if name == "vertex":
            dt = np.dtype([(p[-1], types[p[0]]) for p in props])
            v = np.frombuffer(raw, dt, count, pos)
            pos += dt.itemsize * count
            xyz = np.stack([v["x"], v["y"], v["z"]], 1).astype(np.float32)
            if "nx" in v.dtype.names:
                nrm = np.stack([v["nx"], v["ny"], v["nz"]], 1).astype(np.float32)
        elif name == "face":
            p = props[0]
            assert p[0] == "list"
            cdt, idt = np.dtype(types[p[1]]), np.dtype(types[p[2]])
            faces = []
            for _ in range(count):
                k = int(np.frombuffer(raw, cdt, 1, pos)[0])
                pos += cdt.itemsize
                idx = np.frombuffer(raw, idt, k, pos)
                pos += idt.itemsize * k
                if k != 3:
                    raise ValueError("only triangle meshes are supported")
                faces.append(idx)
            tri = np.asarray(faces, dtype=np.uint32)
        else:
            raise ValueError(f"unsupported PLY element {name}")
    layout = "PointXYZRGBNormal" if nrm is not None else "PointXYZ"
    return pack_blob(xyz, nrm, tri, layout)


def wavy(nx: int, ny: int, lx: float = 1.0, ly: float = 0.8, amp: float = 0.02, lam: float = 0.25):
    """SURVEY.md §8(d) wavy(nx, ny, Lx, Ly, A, lambda): (xyz float32 (V,3), tri uint32 (F,3)).

    Vertex i = col + (nx+1)*row; z = A sin(2 pi x / lambda) sin(2 pi y / lambda) evaluated in double;
    all "lower" triangles (i, i+1, i+nx+2) first, then all "upper" ones (i, i+nx+2, i+nx+1).
    """
    col = np.arange(nx + 1, dtype=np.float64) * (lx / nx)
    row = np.arange(ny + 1, dtype=np.float64) * (ly / ny)
    x, y = np.meshgrid(col, row)
    z = amp * np.sin(2 * np.pi * x / lam) * np.sin(2 * np.pi * y / lam)
    xyz = np.stack([x.ravel(), y.ravel(), z.ravel()], 1).astype(np.float32)
    c, r = np.meshgrid(np.arange(nx, dtype=np.int64), np.arange(ny, dtype=np.int64))
    i = (c + (nx + 1) * r).ravel()
    lower = np.stack([i, i + 1, i + nx + 2], 1)
    upper = np.stack([i, i + nx + 2, i + nx + 1], 1)
    tri = np.concatenate([lower, upper], 0).astype(np.uint32)
    return xyz, tri


def cut_hole(xyz, tri, centre_frac=(0.5, 0.5), radius_frac=0.15):
    """Hole variant: drop triangles whose centroid lies within radius_frac * Ly of the centre (connectivity tests)."""
    lo, hi = xyz.min(0), xyz.max(0)
    c = lo[:2] + np.asarray(centre_frac) * (hi[:2] - lo[:2])
    r = radius_frac * (hi[1] - lo[1])
    cen = xyz[tri.astype(np.int64)].mean(1)[:, :2]
    keep = np.linalg.norm(cen - c, axis=1) > r
    return xyz, np.ascontiguousarray(tri[keep])


def permute(xyz, tri, seed=42, normals=None):
    """Randomly permute vertex and face order (scanner-like ordering, cfg4)."""
    rng = np.random.Generator(np.random.MT19937(seed))
    vperm = rng.permutation(xyz.shape[0])           # new position j holds old vertex vperm[j]
    inv = np.empty_like(vperm)
    inv[vperm] = np.arange(vperm.size)
    fperm = rng.permutation(tri.shape[0])
    xyz2 = np.ascontiguousarray(xyz[vperm])
    tri2 = np.ascontiguousarray(inv[tri.astype(np.int64)][fperm].astype(np.uint32))
    if normals is not None:
        return xyz2, tri2, np.ascontiguousarray(normals[vperm])
    return xyz2, tri2


def rigid_transform(xyz, T):
    """Apply a 4x4 (row, col) transform to float32 points in double, store float32."""
    T = np.asarray(T, np.float64)
    p = xyz.astype(np.float64) @ T[:3, :3].T + T[:3, 3]
    return p.astype(np.float32)


def cfg5_mesh(k: int):
    """SURVEY.md §8(d) cfg5, mesh k of the batch: wavy(nx_k, ny_k) with nx_k = round(1250 f_k), ny_k = round(800 f_k),
    f_k = sqrt(0.75 + 0.5 u_k) (1.5 - 2.5 M triangles), amplitude A_k = 0.01 + 0.03 u'_k, moved by a seeded random rigid
    transform (translation within +-1 m, any rotation).  Returns (xyz, tri, T)."""
    rng = np.random.Generator(np.random.MT19937(1234 + k))
    u, u2 = rng.random(), rng.random()
    f = np.sqrt(0.75 + 0.5 * u)
    nx, ny = int(round(1250 * f)), int(round(800 * f))
    xyz, tri = wavy(nx, ny, amp=0.01 + 0.03 * u2)
    # rotation from a random unit quaternion, translation in [-1, 1]^3
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = rng.uniform(-1.0, 1.0, size=3)
    return rigid_transform(xyz, T), tri, T


def cfg5_triangle_counts(n_meshes: int = 64) -> list[int]:
    """Triangle count of every mesh of the cfg5 batch without building it (the LPT assignment needs only these)."""
    out = []
    for k in range(n_meshes):
        rng = np.random.Generator(np.random.MT19937(1234 + k))
        f = np.sqrt(0.75 + 0.5 * rng.random())
        out.append(2 * int(round(1250 * f)) * int(round(800 * f)))
    return out
