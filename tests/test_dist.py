"""Host logic of the N > 1 path on CPU: world_size-2 (and 3) process groups over gloo.

Each rank plans its slab of lattice planes (here with the CPU oracle standing in for the device, since this container
has no GPU), the slabs are gathered to rank 0 through torch.distributed and concatenated in rank order, the deferred
raster post-ops run on the whole set, and the result must equal the unsharded plan.  Timing is reduced with MAX.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist

    import oracle as orc
    from noether_b200 import meshes
    from noether_b200.dist import RankInfo, SlabPaths, gather_slabs, init_process_group, max_over_ranks, slab_range

    info = RankInfo.from_env()
    init_process_group(info, "gloo")
    xyz, tri = meshes.wavy(40, 32)
    xyz, tri = meshes.cut_hole(xyz, tri)
    nrm = orc.normals_from_mesh_faces(xyz, tri)
    p = orc.pca(xyz, orc.MOMENTS_F64)
    lat = orc.make_lattice(p, orc.principal_axis_direction(p, 0.0), orc.centroid(xyz, orc.MOMENTS_F64), 0.017, True)
    kb, ke = slab_range(int(lat.num_planes), info.rank, info.world)
    f = orc.slice_mesh(xyz, nrm, tri, lat, orc.SLICE_INTENT, orc.SCAN_BINNED, kb, ke).flat()
    local = SlabPaths(f.path_off, f.seg_off, f.path_plane, f.pos, f.nrm, f.edge_lo, f.edge_hi)
    whole = gather_slabs(local, info, 0)
    t = max_over_ranks(float(info.rank + 1), info.world)
    assert t == float(info.world)
    if info.rank == 0:
        ref = orc.slice_mesh(xyz, nrm, tri, lat, orc.SLICE_INTENT, orc.SCAN_BINNED).flat()
        ok = (np.array_equal(whole.path_offsets, ref.path_off) and np.array_equal(whole.segment_offsets, ref.seg_off)
              and np.array_equal(whole.path_plane, ref.path_plane) and np.array_equal(whole.positions, ref.pos)
              and np.array_equal(whole.normals, ref.nrm) and np.array_equal(whole.edge_lo, ref.edge_lo)
              and np.array_equal(whole.edge_hi, ref.edge_hi))
        q.put(("ok" if ok else "mismatch", int(lat.num_planes), int(ref.num_paths)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_plane_slab_sharding_over_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    status = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert status[0] == "ok", status
    assert status[1] > world and status[2] > 10


def test_slab_ranges_partition_the_lattice():
    from noether_b200.dist import slab_range
    for P in (0, 1, 7, 1000, 1001):
        for world in (1, 2, 3, 4, 8):
            edges = [slab_range(P, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == P
            for a, b in zip(edges[:-1], edges[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1


def test_lpt_assignment_of_the_mesh_batch():
    """cfg5: 64 meshes of 1.5 - 2.5 M triangles over 8 GPUs, whole meshes, longest-processing-time first."""
    from noether_b200 import meshes
    from noether_b200.dist import lpt_assign
    sizes = meshes.cfg5_triangle_counts(64)
    assert len(sizes) == 64 and min(sizes) >= 1_450_000 and max(sizes) <= 2_550_000
    shares = lpt_assign(sizes, 8)
    flat = sorted(i for s in shares for i in s)
    assert flat == list(range(64))                       # every mesh planned exactly once
    loads = [sum(sizes[i] for i in s) for s in shares]
    assert max(loads) - min(loads) <= max(sizes)         # LPT bound
    assert max(loads) <= 1.05 * sum(sizes) / 8           # within 5 % of perfect balance on this batch
    assert lpt_assign(sizes, 8) == shares                # deterministic: every rank derives the same assignment
    assert lpt_assign([5, 5, 5], 1) == [[0, 1, 2]]


def test_bench_reference_arm_runs_without_a_gpu():
    """`bench.py --impl reference` (the reference's CPU algorithm through the oracle) needs no device and prints the
    contract's keys"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "cfg2", "--steps", "1",
                        "--warmup", "1", "--cpu-planes", "8"], capture_output=True, text=True, timeout=600, check=True)
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "triangles sliced/sec" and line["unit"] == "triangles/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "triangles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the whole plan measured on every host core rides along (the fairer figure next to the single-threaded port)
    opt = line["cpu_baseline_optimised"]
    assert opt["kind"] == "port" and opt["cores"] >= 1 and opt["value"] > 0 and "whole plan measured" in opt["sample"]


def test_bench_reference_arm_of_the_normals_workload():
    """`bench.py --impl reference --workload cfg4` (NormalsFromMeshFaces + PCA frame through the oracle, on a tenth of the
    mesh): the same contract keys, no device"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "cfg4", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, check=True)
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "triangles/s" and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "cfg4" in line["config"]["workload"]
