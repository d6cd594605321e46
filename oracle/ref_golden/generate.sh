#!/usr/bin/env bash
# Generate tests/golden/ref/*.bin from the UNMODIFIED reference, inside the reference's own docker image.
# Needs docker and network access (apt / rosdep / vcs pull VTK 9.1, PCL, Eigen 3.4, yaml-cpp, boost_plugin_loader 0.4.3):
# neither exists in this repository's development image, so this script is the committed recipe, not something CI runs.
#
#   oracle/ref_golden/generate.sh /path/to/noether-checkout [ubuntu tag: jammy | noble]
#
# 1. builds the reference's image from ITS Dockerfile (docker/Dockerfile: vcs import dependencies.repos, rosdep install,
#    colcon build -DCMAKE_BUILD_TYPE=Release) — the same recipe its CI uses (.github/workflows/ubuntu.yml);
# 2. compiles dump_golden.cpp against the installed noether_tpp inside that image;
# 3. runs it on noether_tpp/test/meshes/wavy_mesh_with_hole.ply and copies the .bin files to tests/golden/ref/.
# Afterwards `python -m pytest tests/test_ref_golden.py` compares the oracle (and, with -m gpu, the CUDA path) with them;
# commit the .bin files together with the versions file this script writes next to them.
set -euo pipefail
REF=${1:?path to a checkout of ros-industrial/noether (the version under test: noether_tpp 0.16.0)}
TAG=${2:-jammy}
HERE=$(cd "$(dirname "$0")" && pwd)
REPO=$(cd "$HERE/../.." && pwd)
OUT="$REPO/tests/golden/ref"
mkdir -p "$OUT"
docker build -f "$REF/docker/Dockerfile" --build-arg TAG="$TAG" -t noether-ref:"$TAG" "$REF"
docker run --rm -v "$HERE":/golden_src:ro -v "$REF":/noether_src:ro -v "$OUT":/golden_out --entrypoint /bin/bash noether-ref:"$TAG" -c '
  set -euo pipefail
  source /opt/noether/install/setup.bash
  cmake -S /golden_src -B /tmp/golden_build -DCMAKE_BUILD_TYPE=Release >/dev/null
  cmake --build /tmp/golden_build -j >/dev/null
  /tmp/golden_build/dump_golden /noether_src/noether_tpp/test/meshes/wavy_mesh_with_hole.ply /golden_out
  { echo "ubuntu: $(. /etc/os-release && echo $VERSION_ID)";
    dpkg-query -W -f="\${Package} \${Version}\n" "libvtk9*" "libpcl-common*" libeigen3-dev "libyaml-cpp0*" 2>/dev/null || true;
    echo "noether: $(grep -m1 "<version>" /noether_src/noether_tpp/package.xml)"; } > /golden_out/VERSIONS.txt
'
ls -l "$OUT"
