"""Worker of tests/test_multi_gpu.py: one rank per GPU (torchrun).  Every rank plans its plane slab of the same mesh,
the slabs are gathered to rank 0 over NCCL, and rank 0 checks the gathered tool paths, array by array and bit for bit,
against the ORACLE's plan (CPU restatement of the reference; handed the GPU's 15 frame floats like every parity test) and
against its own unsharded GPU plan.  argv[1] (optional): "cfg3" runs the 10 M-triangle headline mesh instead of the small
one (plain planner only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import noether_b200 as nb
    from noether_b200 import meshes
    from noether_b200.dist import RankInfo, broadcast_bytes, init_process_group

    info = RankInfo.from_env()
    torch.cuda.set_device(info.local_rank)
    init_process_group(info, "nccl")
    ctx = nb.Context(info.local_rank)
    uid = broadcast_bytes(ctx.comm_unique_id() if info.rank == 0 else None, 0, info.world)
    ctx.comm_init(uid, info.rank, info.world)

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import assert_same_paths, oracle_pca_from_capi
    big = len(sys.argv) > 1 and sys.argv[1] == "cfg3"
    if big:
        xyz, tri = meshes.wavy(2500, 2000)
        spacing = 0.001
    else:
        xyz, tri = meshes.wavy(300, 240)
        xyz, tri = meshes.cut_hole(xyz, tri)
        spacing = 0.0031
    ctx.upload(meshes.pack_blob(xyz, None, tri, "PointXYZ"), force=True)
    nrm = ctx.normals_from_mesh_faces(xyz.shape[0])
    blob = meshes.pack_blob(xyz, nrm, tri)
    # PCIe-parallel upload: each rank copies 1 / world of the bytes, NVLink all-gathers the shares; every rank must end
    # up with exactly the mesh an ordinary upload gives it
    ctx.upload(blob, force=True, sharded=True)
    gx, gn = ctx.download(blob.n_vertices)
    assert np.array_equal(gx.view(np.uint32), xyz.view(np.uint32)) and np.array_equal(gn.view(np.uint32), nrm.view(np.uint32))
    for legacy in ((False,) if big else (False, True)):
        if legacy:
            planner = nb.PlaneSlicerLegacyRasterPlanner(nb.PrincipalAxisDirectionGenerator(0.0, ctx),
                                                        nb.CentroidOriginGenerator(ctx), 0.004, 0.01, 0.0, ctx)
        else:
            planner = nb.PlaneSlicerRasterPlanner(nb.PrincipalAxisDirectionGenerator(0.0, ctx),
                                                  nb.CentroidOriginGenerator(ctx), ctx)
        planner.set_line_spacing(spacing)
        planner.generate_rasters_bidirectionally(True)
        ctx._mesh_key = None
        local = planner.plan(blob, info.rank, info.world, sharded_upload=True)
        assert local.n_waypoints > 0 and local.positions.shape[0] == 0, "a slab stays on the device until the gather"
        got = ctx.gather(local, 0)
        local.free()
        if info.rank == 0:
            ctx._mesh_key = None
            ref = planner.plan(blob)
            assert got.n_paths == ref.n_paths > 100 and got.n_segments == ref.n_segments >= got.n_paths
            assert big or got.n_segments > got.n_paths
            for name in ("path_offsets", "segment_offsets", "path_plane", "positions", "normals"):
                a, b = getattr(got, name), getattr(ref, name)
                assert a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8)), name
            assert np.array_equal(got.poses(), ref.poses())
            # ... and against the oracle: the reference's algorithm on the CPU, the whole lattice in one piece
            import oracle as orc
            orc.build()
            frame = oracle_pca_from_capi(orc, ctx.pca())
            if legacy:
                paths, _, _ = orc.plan_plane_slicer_legacy(xyz, nrm, tri, spacing, 0.004, 0.01, 0.0, True, slice_mode=orc.SLICE_INTENT,
                                                           scan_mode=orc.SCAN_BINNED, frame=frame)
                f = paths.flat()
                assert got.n_paths == f.num_paths
                assert np.array_equal(got.segment_offsets.astype(np.uint64), f.seg_off)
                assert np.array_equal(got.poses(), f.pose), "legacy poses differ from the oracle's"
            else:
                paths, _, _ = orc.plan_plane_slicer(xyz, nrm, tri, spacing, True, slice_mode=orc.SLICE_INTENT,
                                                    scan_mode=orc.SCAN_BINNED, frame=frame)
                assert_same_paths(got, paths.flat(), f"{info.world} slabs", poses=not big)
            print(f"gather ok (legacy={legacy}): {got.n_paths} paths, {got.n_segments} segments, {got.n_waypoints} waypoints "
                  f"from {info.world} slabs", flush=True)
        else:
            assert got is None
    torch.distributed.barrier()
    ctx.close()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
