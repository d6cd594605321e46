import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    import oracle
    oracle.build()
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def ply_mesh(golden_dir):
    from noether_b200 import meshes
    return meshes.read_ply(os.path.join(golden_dir, "wavy_mesh_with_hole.ply"))


@pytest.fixture(scope="session")
def ctx():
    """A GPU context; only -m gpu tests may request it."""
    import noether_b200 as nb
    c = nb.Context(0)
    yield c
    c.close()
