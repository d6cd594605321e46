"""createTransform / DirectionOfTravelOrientationModifier on the host cores (noether_b200/csrc/toolpath.cpp): the AVX2 form
(four waypoints per step) must equal the scalar form bit for bit (the scalar form is compared with the oracle's convertToPoses,
plane_slicer_raster_planner.cpp:380-394, by the GPU parity tests).  No GPU involved: nb2_expand_poses works on arrays the caller holds."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r"""
import ctypes as C, sys, numpy as np
sys.path.insert(0, %r)
from noether_b200 import _capi
L = _capi.load()
d = np.load(sys.argv[1])
pos, nrm, off = d["pos"], d["nrm"], d["off"]
out = np.zeros((pos.shape[0], 16), np.float64)
for post in (0, 1):
    rc = L.nb2_expand_poses(pos.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                            C.c_uint64(off.size - 1), C.c_int(post), out.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    np.save(sys.argv[2] + str(post) + ".npy", out)
"""


def _segments(rng):
    # segments of every length from 1 up (tails of 0 - 3 waypoints after the groups of four), degenerate inputs included
    lens = list(range(1, 23)) + [int(x) for x in rng.integers(1, 300, 40)]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    W = int(off[-1])
    pos = rng.standard_normal((W, 3)).astype(np.float32)
    nrm = rng.standard_normal((W, 3)).astype(np.float32)
    nrm[5] = 0.0                      # zero normal: normalized() leaves it
    pos[40] = pos[41]                 # zero travel direction
    nrm[60] = (pos[61] - pos[60])     # normal parallel to the travel direction: y = 0
    pos[100:104] *= 1e-30             # tiny squared norms
    return pos, nrm, off


def test_avx2_equals_scalar_bit_for_bit(tmp_path):
    rng = np.random.default_rng(7)
    pos, nrm, off = _segments(rng)
    np.savez(tmp_path / "in.npz", pos=pos, nrm=nrm, off=off)
    out = {}
    for name, env in (("simd", {}), ("scalar", {"NB2_NO_AVX2": "1"})):
        e = dict(os.environ)
        e.pop("NB2_NO_AVX2", None)
        e.update(env)
        r = subprocess.run([sys.executable, "-c", _CHILD % ROOT, str(tmp_path / "in.npz"), str(tmp_path / name)], env=e,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        out[name] = [np.load(str(tmp_path / name) + f"{p}.npy") for p in (0, 1)]
    for p in (0, 1):
        a, b = out["simd"][p], out["scalar"][p]
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), f"post-ops {p}: AVX2 and scalar poses differ"
        assert np.all(a[:, [3, 7, 11]] == 0.0) and np.all(a[:, 15] == 1.0)
        assert np.array_equal(a[:, 12:15], pos.astype(np.float64))
