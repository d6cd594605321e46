"""In-tree build of libnoether_b200.so (hand-written sm_100a CUDA + the C ABI in include/noether_b200.h).

    python -m noether_b200.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU; the .so lands next to this file so that it travels to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnoether_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SOURCES = ["api.cu", "kernels_mesh.cu", "kernels_slice.cu", "kernels_normals.cu", "kernels_frame.cu", "kernels_legacy.cu", "kernels_ply.cu", "kernels_stl.cu", "kernels_subdivide.cu", "kernels_cluster.cu", "modifiers.cu", "modifier_plan.cpp", "ply.cpp", "frame.cpp", "toolpath.cpp",
           "comm.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # parity: never contract mul+add into FMA, neither on the device nor in the host arithmetic
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA toolkit is required to build noether_b200 (there is no CPU fallback)")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "noether_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    objs = []
    build_dir = os.path.join(HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    logs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, cmd, p in procs:
        out, _ = p.communicate()
        logs.append(f"$ {' '.join(cmd)}\n{out}")
        if p.returncode != 0:
            failed = True
    log_text = "\n".join(logs)
    with open(os.path.join(build_dir, "build.log"), "w") as f:
        f.write(log_text)
    if failed:
        raise RuntimeError("nvcc failed:\n" + log_text)
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-ldl", "-lpthread"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    if verbose:
        print(log_text)
    return LIB


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
