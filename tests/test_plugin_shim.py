"""The C++ host side: plugin/b200_plugins.cpp (the boost_plugin_loader plugins that bind the C ABI) compiled against
plugin/compat, and the reference's own planner / modifier tests driven through it (tests/cpp/plugin_shim_test.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "plugin", "build")
PLUGIN = os.path.join(OUT, "libnoether_b200_plugins.so")


def _build():
    from noether_b200.build import build_library
    build_library()
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "plugin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_plugin_library_builds_and_exports_the_reference_aliases():
    _build()
    syms = subprocess.run(["nm", "-D", "--defined-only", PLUGIN], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in syms.splitlines() if line.strip()}
    # aliases of noether_tpp/src/plugins.cpp:62,89,111,179,278
    for alias in ("PlaneSlicer", "PlaneSlicerLegacy", "PrincipalAxis", "PCARotated", "Centroid", "AABBCenter", "NormalsFromMeshFaces",
                  "FaceMidpointSubdivision", "FaceSubdivisionByArea", "EuclideanClustering",  # plugins.cpp:57-59
                  # tool path modifiers that run on the device (plugins.cpp:182-199,219)
                  "SnakeOrganization", "LinearApproach", "LinearDeparture", "Offset", "FixedOrientation", "UniformOrientation",
                  "Concatenate", "ToolDragOrientation", "BiasedToolDragOrientation", "DirectionOfTravelOrientation",
                  "MovingAverageOrientationSmoothing", "CircularLeadIn", "CircularLeadOut", "RasterOrganization", "JoinCloseSegments", "MinimumSegmentLength",
                  "CompoundToolPathModifier"):
        assert alias in exported
    # each plugin object sits in the ELF section boost_plugin_loader enumerates for its kind
    # (NOETHER_*_SECTION, noether_tpp/CMakeLists.txt:126-131)
    sections = subprocess.run(["readelf", "-S", "-W", PLUGIN], capture_output=True, text=True, check=True).stdout
    names = {tok for line in sections.splitlines() for tok in line.replace("]", " ").split()}
    for section in ("tpp", "dg", "og", "mesh", "mod"):
        assert section in names
    # every symbol the plugin library needs resolves at load time (boost_plugin_loader dlopens it; a template the stand-in
    # headers declare but do not define would only show up there)
    import ctypes
    ctypes.CDLL(PLUGIN, mode=os.RTLD_NOW)
    # the shim reaches the product through the C ABI only
    undefined = subprocess.run(["nm", "-D", "--undefined-only", PLUGIN], capture_output=True, text=True, check=True).stdout
    used = {line.split()[-1] for line in undefined.splitlines() if " nb2_" in line}
    from noether_b200 import _capi
    assert used and used <= set(_capi.DECLARED_SYMBOLS)


@pytest.mark.gpu
def test_reference_tests_through_the_plugin_boundary(golden_dir):
    _build()
    r = subprocess.run([os.path.join(OUT, "plugin_shim_test"), PLUGIN, os.path.join(golden_dir, "wavy_mesh_with_hole.ply")],
                       capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


@pytest.mark.gpu
def test_clustering_through_the_plugin_boundary(golden_dir):
    _build()
    r = subprocess.run([os.path.join(OUT, "plugin_shim_test"), PLUGIN, os.path.join(golden_dir, "wavy_mesh_with_hole.ply"),
                        "--clustering"], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0


def test_clustering_host_side_of_the_shim(golden_dir):
    """size filter, PCL's size sort and sub-mesh cutting of the plugin library's EuclideanClustering, driven through its test
    hook with the labels the device would return: no GPU involved"""
    _build()
    r = subprocess.run([os.path.join(OUT, "plugin_shim_test"), PLUGIN, os.path.join(golden_dir, "wavy_mesh_with_hole.ply"),
                        "--cpu-only"], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "all checks passed" in r.stdout
