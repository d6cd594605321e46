"""The C-ABI library without a GPU: it loads, exports every symbol include/noether_b200.h declares, keeps torch/C++
types out of the signatures, refuses to run without a device, and its host-only arithmetic matches the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "noether_b200.h")


@pytest.fixture(scope="module")
def capi():
    from noether_b200 import _capi
    from noether_b200.build import build_library
    build_library()
    _capi.load()
    return _capi


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(nb2_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(capi):
    declared = _declared_functions()
    assert set(declared) == set(capi.DECLARED_SYMBOLS)
    L = capi.load()
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/noether_b200.h but not exported"
    # exported dynamic symbols carry C linkage (no mangled nb2_ names)
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(declared) <= exported
    assert L.nb2_abi_version() == 3


def test_header_is_plain_c(tmp_path):
    # the boundary must be consumable from C: no C++, no torch / PCL / Eigen types
    src = tmp_path / "t.c"
    src.write_text('#include "noether_b200.h"\nint main(void){ nb2_plan_params p; nb2_plan_params_default(&p); return (int)sizeof(p) == 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(HEADER).read()
    for forbidden in ("torch", "std::", "Eigen::", "pcl::", "template"):
        assert forbidden not in re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def test_struct_layouts_match_the_header(capi, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "noether_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(nb2_mesh_view),sizeof(nb2_pca),sizeof(nb2_lattice),sizeof(nb2_plan_params),sizeof(nb2_result_view),'
                   'offsetof(nb2_plan_params,origin),offsetof(nb2_plan_params,download));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    assert sizes[:5] == [C.sizeof(capi.MeshView), C.sizeof(capi.Pca), C.sizeof(capi.Lattice), C.sizeof(capi.PlanParams),
                         C.sizeof(capi.ResultView)]
    assert sizes[5] == capi.PlanParams.origin.offset and sizes[6] == capi.PlanParams.download.offset


def test_no_cpu_fallback(capi):
    """Without a usable CUDA device the product refuses to run (it must never compute on the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = capi.load().nb2_context_create(0, C.byref(h))
    assert rc == capi.NB2_ERR_CUDA and not h.value
    assert b"no CPU fallback" in capi.load().nb2_last_error()
    import noether_b200 as nb
    with pytest.raises(capi.Nb2Error):
        nb.Context(0)


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may use oracle/."""
    pkg = os.path.join(ROOT, "noether_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f
                assert "oracle/" not in text and "oracle_" not in text, f
    plugin = os.path.join(ROOT, "plugin")
    for dirpath, _, files in os.walk(plugin):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            text = open(os.path.join(dirpath, f), errors="ignore").read()
            assert "oracle" not in text.lower(), f
    out = subprocess.run(["ldd", os.path.join(pkg, "libnoether_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_plan_params_default(capi):
    p = capi.PlanParams()
    capi.load().nb2_plan_params_default(C.byref(p))
    assert p.bidirectional == 1 and p.raster_post_ops == 1 and p.download == 1 and p.shard_count == 1
    assert p.direction_mode == capi.NB2_DIRECTION_PRINCIPAL_AXIS and p.origin_mode == capi.NB2_ORIGIN_CENTROID
    assert p.line_spacing == pytest.approx(0.1)


@pytest.mark.parametrize("seed", range(6))
def test_host_lattice_matches_oracle_bit_for_bit(capi, orc, seed):
    """nb2_make_lattice is pure host arithmetic (no device): same bytes as the oracle's restatement of
    plane_slicer_raster_planner.cpp:550-586 for the same PCA frame, including the reference's quirks."""
    from noether_b200 import make_lattice
    rng = np.random.default_rng(seed)
    pts = ((rng.random((500, 3)) - 0.5) * np.array([3.0, 1.5, 0.2]) + rng.normal(size=3) * 4).astype(np.float32)
    o = orc.pca(pts, orc.MOMENTS_F32_SEQUENTIAL)
    g = capi.Pca()
    for i in range(3):
        g.mean[i] = float(o.mean[i])
        g.scales[i] = float(o.scales[i])
    for i, v in enumerate(o.evecs.T.ravel()):
        g.evecs[i] = float(v)
    for bidir in (True, False):
        d = rng.normal(size=3)
        org = rng.normal(size=3) * 3
        spacing = float(rng.uniform(0.01, 0.5))
        a = make_lattice(g, d, org, spacing, bidir)
        b = orc.make_lattice(o, d, org, spacing, bidir)
        assert bytes(a) == bytes(b)
    with pytest.raises(capi.Nb2Error):
        make_lattice(g, [1, 0, 0], [0, 0, 0], -1.0, True)


def test_entry_points_refuse_a_null_context(capi):
    """argument checks happen before any device work: a NULL context is a status + message, never a crash"""
    L = capi.load()
    info = capi.MeshFileInfo()
    for call in (lambda: L.nb2_face_midpoint_subdivision(None, 1, -1, C.byref(info)),
                 lambda: L.nb2_face_subdivision_by_area(None, C.c_float(1.0), -1, 0, C.byref(info)),
                 lambda: L.nb2_mesh_download_cloud(None, None, 0),
                 lambda: L.nb2_mesh_record_layout(None, None, None, None),
                 lambda: L.nb2_mesh_download_triangles(None, None, 0),
                 lambda: L.nb2_euclidean_cluster_labels(None, C.c_double(1.0), None, 0)):
        assert call() == capi.NB2_ERR_INVALID_ARGUMENT
        assert L.nb2_last_error()
