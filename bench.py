#!/usr/bin/env python
"""Benchmark of the raster-slicing hot path (BASELINE.json metric: triangles sliced / s and ms per plan).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 2 --warmup 1      # the reference's CPU algorithm on the host cores

A step = one plan of the workload mesh: PlaneSlicerRasterPlanner(PrincipalAxis(0), Centroid) at the workload's line
spacing.  Workload `cfg3` (default) is SURVEY.md §8(d) cfg3: wavy(2500, 2000) = 10,000,000 triangles / 5,004,501
vertices, 1 mm spacing (BASELINE.json configs[2], the mesh the metric is quoted on; it fits one GPU).

  value  device throughput: the interleaved cloud + triangle list are resident in HBM when the timed region starts; each
         step runs ingest + moments + extents + lattice classification + hit count / scan / bucket fill + per-plane
         linking, tool paths left in HBM.  Timed with CUDA events on the library's stream; max over ranks.
  e2e    the same plan through the reference-facing planner call (noether_b200.PlaneSlicerRasterPlanner.plan) from
         pinned HOST buffers: H2D of the mesh, the plan, D2H of the tool paths, raster post-ops (and, for N > 1, the
         NCCL gather of the slabs to rank 0) inside the timed region.  Wall clock around synchronous calls.

For N > 1 the lattice planes are split into N contiguous slabs (one rank = one GPU = one slab); total work is fixed, so
scaling is "strong".  torch is plumbing only (pinned/device buffers, rendezvous, barriers).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nx, ny, line_spacing, description)
    "cfg3": (2500, 2000, 0.001, "wavy(2500,2000) 10M tris / 5.0M verts, 1 mm spacing, PrincipalAxis(0)+Centroid, bidirectional"),
    "cfg2": (800, 625, 0.005, "wavy(800,625) 1M tris / 0.5M verts, 5 mm spacing, PrincipalAxis(0)+Centroid, bidirectional"),
    "tiny": (80, 64, 0.02, "wavy(80,64) smoke-sized mesh"),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index  # index, comma-separated list of indices, or None (this rank does not sample)
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False) if gpu_index is not None else None
        self.proc = None

    def start(self):
        if self.gpu is None:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.gpu is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampled by rank 0"]}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        sm, smax, reasons = [], [], set()
        for line in self.file.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.file.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


def build_workload(name: str):
    from noether_b200 import meshes
    nx, ny, spacing, desc = WORKLOADS[name]
    xyz, tri = meshes.wavy(nx, ny)
    return xyz, tri, spacing, desc


def workload_config(name: str, desc: str, world: int, V: int, F: int) -> dict:
    """`config` of the bench line: the workload and how each arm measures it.  Both arms print the very same object, so that a
    comparison of the two lines sees one configuration; what a run found out (planes, waypoints, wall times) is in `result`."""
    return {
        "workload": f"{name}: {desc}",
        "parallelism": f"plane slabs x{world}" if world > 1 else "single GPU",
        "l2": ("GPU arm: flushed between timed iterations (256 MB write outside the timed intervals); inputs are %d MB, resident as "
               "24-byte {xyz, normal} cloud records (what this library's upload leaves in HBM) and uint32 x 3 triangles; reference arm: "
               "host memory" % ((24 * V + 12 * F) // 2**20)),
        "timing": ("GPU arm: CUDA events on the library stream (entry of nb2_plan_mesh .. behind the plan's last kernel), max over "
                   "ranks; reference arm: wall clock of a bounded plane sample, extrapolated linearly to the full lattice"),
    }


def cpu_reference_sample(xyz, nrm, tri, spacing, n_planes: int):
    """The reference's algorithm (oracle, faithful mode, every vertex + every cell per plane, one thread) on a bounded
    sample: the plane-independent prologue plus `n_planes` evenly spaced lattice planes, extrapolated linearly in the
    plane count (iterations of the reference's loop are independent).  Returns a dict."""
    import oracle as orc
    t0 = time.perf_counter()
    p = orc.pca(xyz, orc.MOMENTS_F32_SEQUENTIAL)
    d = orc.principal_axis_direction(p, 0.0)
    o = orc.centroid(xyz, orc.MOMENTS_F32_SEQUENTIAL)
    lat = orc.make_lattice(p, d, o, spacing, True)
    t_pro = time.perf_counter() - t0
    P = int(lat.num_planes)
    ks = np.unique(np.linspace(0, P - 1, min(n_planes, P)).round().astype(int))
    t_planes = []
    for k in ks:
        t1 = time.perf_counter()
        orc.slice_mesh(xyz, nrm, tri, lat, orc.SLICE_FAITHFUL, orc.SCAN_ALL_CELLS, int(k), int(k) + 1)
        t_planes.append(time.perf_counter() - t1)
    per_plane = float(np.mean(t_planes))
    t_full = t_pro + per_plane * P
    return {"t_prologue_s": t_pro, "t_per_plane_s": per_plane, "planes_sampled": int(len(ks)), "planes": P,
            "t_plan_extrapolated_s": t_full, "t_measured_s": t_pro + float(np.sum(t_planes)),
            "tri_per_s": tri.shape[0] / t_full}


def ncu_traffic(workload: str, stage: str):
    """DRAM bytes per launch of the stage's kernel from the committed `ncu --set full` capture of this workload
    (profiles/ncu_traffic_<workload>.json, written by tools/ncu_summary.py), or None."""
    kernel = {"ingest+moments": "k_ingest_moments", "extents": "k_extents", "extents+lattice": "k_extents",
              "classify": "k_classify", "count": "k_count_hits", "count+scan": "k_count_hits", "emit": "k_emit_hits",
              "link": "k_link_planes", "waypoints": "k_waypoints", "lattice": "k_frame_lattice"}.get(stage)
    path = os.path.join(ROOT, "profiles", f"ncu_traffic_{workload}.json")
    if kernel is None or not os.path.exists(path):
        return None
    with open(path) as f:
        table = json.load(f)
    # (template instances and kernel variants carry suffixes: k_count_hits<0>, k_link_planes_wide, k_waypoints<1>)
    entry = table.get(kernel) or next((v for k, v in table.items() if k.startswith(kernel)), None)
    return entry["dram_bytes_per_launch"] if entry else None


def algorithmic_bytes(stage: str, V: int, F: int, H: int, W: int, S: int, P: int, point_step: int) -> float | None:
    """Compulsory bytes of each kernel: what it must read and write once."""
    table = {
        "ingest+moments": 12 * V + 16 * V,       # xyz in, float4 position out (normals are gathered in place later)
        "validate": 12 * F,
        "extents": 12 * V,
        "extents+lattice": 12 * V,
        "classify": 12 * V + 4 * V,
        "count": 12 * F + 4 * V,
        "count+scan": 12 * F + 4 * V + 8 * P,
        "emit": 12 * F + 4 * V + 32 * H,          # triangles, vertex classes, two 16-byte endpoint records per hit
        "link": 32 * H + 8 * W + 4 * S,            # endpoint records in, one {rank, plane} word pair per waypoint out
        "waypoints": 8 * W + 12 * W + 24 * W + 24 * W,  # tags, first-inserter triangles, cut-edge vertex data, waypoints out
    }
    return table.get(stage)


def side_workload(args, info, emit):
    """BASELINE.json configs[3] and [4] behind the same command line: cfg4 = NormalsFromMeshFaces + PCA frame on the 50 M-triangle
    mesh with scanner-like (permuted) vertex / face order, device-resident; cfg5 = the batch of 64 part meshes, whole meshes
    per GPU (LPT on triangle counts), end to end from pinned host memory.  One JSON line with the contract's keys."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    peak, peak_src = measured_peaks()
    if args.workload == "cfg4":
        if info.rank != 0:
            return
        if args.impl == "reference":
            import oracle as orc
            from noether_b200 import meshes
            orc.build()
            # bounded sample: the same generator at 1/10 of the triangles (5 M), permuted; the passes are linear in the mesh size
            xyz, tri = meshes.permute(*meshes.wavy(1976, 1265), seed=42)
            t0 = time.perf_counter()
            for _ in range(max(1, args.steps)):
                orc.normals_from_mesh_faces(xyz, tri)
                orc.pca(xyz, orc.MOMENTS_F32_SEQUENTIAL)
            dt = (time.perf_counter() - t0) / max(1, args.steps)
            value = tri.shape[0] / dt
            emit({"impl": "reference", "metric": "triangles/s (NormalsFromMeshFaces + PCA frame)", "value": value, "unit": "triangles/s",
                  "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt * 50e6 / tri.shape[0],
                  "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                  "config": {"workload": "cfg4: wavy(6250,4000) 50M tris / 25M verts, permuted vertex and face order"},
                  "cpu_baseline": {"value": value, "unit": "triangles/s", "cores": 1, "kind": "port",
                                   "sample": f"the same generator at {tri.shape[0]} triangles (1/10), permuted; single-thread oracle, "
                                             "rate carried over (the passes are linear in the mesh size)"},
                  "e2e": {"value": value, "unit": "triangles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
            return
        import profile_normals
        sampler = ClockSampler("0")
        sampler.start()
        r = profile_normals.measure("cfg4", iters=max(args.steps, 3) + 1)
        clocks = sampler.stop()
        ms = r["normals_ms_mean"] + r["pca_frame_ms_mean"]
        B = 12 * r["F"] + 24 * r["V"]
        emit({"metric": "triangles/s (NormalsFromMeshFaces + PCA frame)", "value": r["F"] / (ms / 1e3), "unit": "triangles/s",
              "n_gpus": 1, "steps": max(args.steps, 3), "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
              "vs_baseline": None, "dtype": "f32", "data": "synthetic",
              "config": {"workload": "cfg4: wavy(6250,4000) 50M tris / 25M verts, permuted vertex and face order (scan-like), "
                                     "NormalsFromMeshFaces then the PrincipalAxis PCA frame, mesh resident in HBM",
                         "l2": "inputs are 900 MB, larger than L2", "timing": "CUDA events on the library stream"},
              "clocks": clocks, "e2e": None, "gpu_launches": int(r["launches_per_iteration"]),
              "roofline": {"bound": "hbm", "kernel": "NormalsFromMeshFaces (4 kernels)", "achieved": B / r["normals_ms_mean"] / 1e6,
                           "peak": peak, "unit": "GB/s", "frac": B / r["normals_ms_mean"] / 1e6 / peak, "traffic": None,
                           "algorithmic_bytes_per_launch": B, "launch_ms": r["normals_ms_mean"], "peak_source": peak_src,
                           "stage_ms": r["stages_ms"], "pca_frame_ms": r["pca_frame_ms_mean"], "pca_frac": r["pca_frac_of_peak"]},
              "cpu_baseline": None})
        return
    # cfg5
    if args.impl == "reference":
        if info.rank != 0:
            return
        import oracle as orc
        from noether_b200 import meshes
        orc.build()
        xyz, tri, _ = meshes.cfg5_mesh(0)
        nrm = orc.normals_from_mesh_faces(xyz, tri)
        s = cpu_reference_sample(xyz, nrm, tri, 0.002, 24)
        emit({"impl": "reference", "metric": "triangles sliced/sec (batch of 64 part meshes)", "value": s["tri_per_s"], "unit": "triangles/s",
              "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": 1e3 * s["t_plan_extrapolated_s"] * 64,
              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
              "config": {"workload": "cfg5: 64 part meshes (1.5 - 2.5 M triangles), 2 mm spacing"},
              "cpu_baseline": {"value": s["tri_per_s"], "unit": "triangles/s", "cores": 1, "kind": "port",
                               "sample": f"mesh 0 of the batch, prologue + {s['planes_sampled']} of {s['planes']} planes, extrapolated"},
              "e2e": {"value": s["tri_per_s"], "unit": "triangles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
        return
    import bench_batch
    sampler = ClockSampler(",".join(str(i) for i in range(info.world)) if info.rank == 0 else None)
    sampler.start()
    r = bench_batch.main(["--meshes", "64", "--threads", "2", "--repeat", str(max(args.steps, 3)), "--quiet"])
    clocks = sampler.stop()
    if info.rank != 0 or r is None:
        return
    e2e = {"value": r["triangles"] / r["batch_s_mean"], "unit": "triangles/s", "ms_per_step": 1e3 * r["batch_s_mean"],
           "h2d_bytes_per_step": r["h2d_bytes"] * info.world, "d2h_bytes_per_step": None,
           "boundary": "C ABI from Python, pinned host meshes in, compact SoA tool paths out; two contexts per GPU overlap copies and kernels"}
    emit({"metric": "triangles sliced/sec (batch of 64 part meshes)", "value": e2e["value"], "unit": "triangles/s", "n_gpus": info.world,
          "steps": r["repeat"], "warmup": 1, "ms_per_step": e2e["ms_per_step"], "higher_is_better": True, "scaling": "strong",
          "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "config": {"workload": r["workload"], "parallelism": f"whole meshes x{info.world} (LPT), no exchange between ranks",
                     "timing": "wall clock around the batch, max over ranks (the step is end to end: every mesh is uploaded, planned, downloaded)",
                     "l2": "every mesh is a different input (56 MB each), 3.6 GB per batch"},
          "clocks": clocks, "e2e": e2e, "gpu_launches": r["rank0_kernel_launches"], "roofline": None, "cpu_baseline": None})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS) + ["cfg4", "cfg5"],
                    help="cfg3 (default) is the configuration BASELINE.json's metric is quoted on; cfg2, cfg4 (NormalsFromMeshFaces "
                         "+ PCA on the 50 M-triangle permuted mesh) and cfg5 (batch of 64 part meshes, whole meshes per GPU) are the "
                         "other BASELINE.json configs")
    ap.add_argument("--cpu-planes", type=int, default=160,
                    help="lattice planes sampled by the CPU baseline (cfg3: about 76 ms each => ~12 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: everything libraries print meanwhile (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)

    from noether_b200.dist import RankInfo
    info = RankInfo.from_env()
    if info.world != args.gpus and info.world > 1:
        print(f"warning: WORLD_SIZE={info.world} but --gpus {args.gpus}", file=sys.stderr)

    if args.workload in ("cfg4", "cfg5"):
        side_workload(args, info, emit)
        return
    xyz, tri, spacing, desc = build_workload(args.workload)
    V, F = xyz.shape[0], tri.shape[0]

    # -----------------------------------------------------------------------------------------------------------------
    # reference arm: the reference's own CPU algorithm on the host cores (rank 0 only)
    # -----------------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if info.rank != 0:
            return
        import oracle as orc
        orc.build()
        nrm = orc.normals_from_mesh_faces(xyz, tri)
        per_step = max(2, args.cpu_planes // 2)  # cfg3: 80 planes x 76 ms = ~6 s of single-thread work per step
        for _ in range(args.warmup):
            cpu_reference_sample(xyz, nrm, tri, spacing, 2)
        t0 = time.perf_counter()
        runs = [cpu_reference_sample(xyz, nrm, tri, spacing, per_step) for _ in range(args.steps)]
        wall = time.perf_counter() - t0
        t_plan = float(np.mean([r["t_plan_extrapolated_s"] for r in runs]))
        value = F / t_plan
        line = {
            "impl": "reference", "metric": "triangles sliced/sec", "value": value, "unit": "triangles/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_plan,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, desc, args.gpus, V, F),
            "cpu_baseline": {"value": value, "unit": "triangles/s", "cores": 1, "kind": "port",
                             "sample": (f"plane-independent prologue + {runs[0]['planes_sampled']} of {runs[0]['planes']} "
                                        "lattice planes per step (every vertex and every cell visited per plane, as "
                                        "vtkCutter does), extrapolated linearly in the plane count; the reference is "
                                        "single-threaded so 1 core is all it can use; wall time of the sampled work "
                                        f"{wall:.1f} s")},
            "e2e": {"value": value, "unit": "triangles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        # next to it, a whole plan actually measured on every host core: the same algorithm with the cells binned by lattice
        # interval and the lattice cut into slabs over forked workers (tools/cpu_optimised.py) — not what the reference does
        # (it is single-threaded and visits every cell per plane), but the fairer figure to hold a GPU number against
        if args.workload in ("cfg2", "cfg3"):
            try:
                r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "cpu_optimised.py"), args.workload],
                                   capture_output=True, text=True, timeout=180)
                o = json.loads(r.stdout.strip().splitlines()[-1])
                line["cpu_baseline_optimised"] = {
                    "value": o["tri_per_s"], "unit": "triangles/s", "cores": o["workers"], "kind": "port",
                    "sample": f"whole plan measured: prologue + all {o['planes']} lattice planes in {o['t_plan_s']:.2f} s"}
            except Exception as e:
                line["cpu_baseline_optimised"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        emit(line)
        return

    # -----------------------------------------------------------------------------------------------------------------
    # our arm
    # -----------------------------------------------------------------------------------------------------------------
    import torch

    import noether_b200 as nb
    from noether_b200 import _capi, meshes
    from noether_b200.dist import broadcast_bytes, init_process_group, max_over_ranks

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: noether_b200 has no CPU fallback")
    torch.cuda.set_device(info.local_rank)
    dev = torch.device("cuda", info.local_rank)
    init_process_group(info, "nccl")
    ctx = nb.Context(info.local_rank)
    if info.world > 1:
        uid = broadcast_bytes(ctx.comm_unique_id() if info.rank == 0 else None, 0, info.world)
        ctx.comm_init(uid, info.rank, info.world)

    def barrier():
        if info.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # vertex normals from our own NormalsFromMeshFaces (timed separately; not part of the plan)
    blob0 = meshes.pack_blob(xyz, None, tri, "PointXYZ")
    ctx.upload(blob0, force=True)
    nrm = ctx.normals_from_mesh_faces(V)
    ctx.normals_from_mesh_faces(V, download=False)
    normals_ms = {k: v for k, v in ctx.timings().items() if k in ("face normals + valence", "scan", "csr fill", "vertex gather")}
    blob = meshes.pack_blob(xyz, nrm, tri, "PointNormal")

    # pinned host copies (e2e input) and device-resident copies (value input); torch only allocates the memory
    h_cloud = torch.empty(blob.data.size, dtype=torch.uint8, pin_memory=True)
    h_cloud.numpy()[:] = blob.data
    h_tri = torch.empty(tri.size, dtype=torch.int32, pin_memory=True)
    h_tri.numpy()[:] = tri.reshape(-1).view(np.int32)
    # value leg: the mesh as this library's own upload leaves it in HBM — {xyz, normal} records of 24 bytes
    # (nb2_mesh_upload with NB2_MESH_GEOMETRY_ONLY / nb2_mesh_upload_stream compact the host cloud's 48-byte records)
    blob_dev = meshes.pack_blob(xyz, nrm, tri, "Packed24")
    d_cloud = torch.from_numpy(blob_dev.data).to(dev)
    d_tri = h_tri.to(dev)
    torch.cuda.synchronize()
    pinned_mesh = meshes.MeshBlob(h_cloud.numpy(), V, blob.point_step, blob.off_xyz, blob.off_normal,
                                  h_tri.numpy().view(np.uint32).reshape(-1, 3))

    prm = _capi.PlanParams()
    _capi.load().nb2_plan_params_default(prm)
    prm.line_spacing = spacing
    prm.bidirectional = 1
    prm.direction_mode = _capi.NB2_DIRECTION_PRINCIPAL_AXIS
    prm.origin_mode = _capi.NB2_ORIGIN_CENTROID
    prm.emit_edges = 0
    prm.shard_rank, prm.shard_count = info.rank, info.world
    prm.download = 0

    def device_step():
        r = ctx.plan_mesh_raw(d_cloud.data_ptr(), V, blob_dev.point_step, blob_dev.off_xyz, blob_dev.off_normal, d_tri.data_ptr(), F, prm)
        counts = (r.n_paths, r.n_segments, r.n_waypoints)
        r.free()
        return counts

    planner = nb.PlaneSlicerRasterPlanner(nb.PrincipalAxisDirectionGenerator(0.0, ctx), nb.CentroidOriginGenerator(ctx), ctx)
    planner.set_line_spacing(spacing)
    planner.generate_rasters_bidirectionally(True)
    planner.emit_edges = False

    trace = os.environ.get("NB2_BENCH_TRACE") == "1"

    def e2e_step():
        ctx._mesh_key = None  # every step uploads the mesh again from the pinned host buffers
        # N > 1: every rank copies 1 / N of the mesh over its own PCIe link, NVLink all-gathers the shares
        t0 = time.perf_counter()
        t1 = t0
        if trace and os.environ.get("NB2_BENCH_SPLIT") == "1":
            ctx.upload(pinned_mesh, force=True, sharded=info.world > 1)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
        local = planner.plan(pinned_mesh, info.rank, info.world, sharded_upload=info.world > 1)
        t2 = time.perf_counter()
        if info.world > 1:
            out = ctx.gather(local, 0)
            if trace:
                print(f"[rank {info.rank}] upload {1e3 * (t1 - t0):.3f} ms, plan {1e3 * (t2 - t1):.3f} ms, "
                      f"gather {1e3 * (time.perf_counter() - t2):.3f} ms", file=sys.stderr)
            n = (out.n_paths, out.n_segments, out.n_waypoints) if out is not None else (0, 0, 0)
            # the slabs travel GPU to GPU; only the root copies tool paths to the host (all of them)
            d2h = (out.n_waypoints * 24 + out.n_segments * 4) if out is not None else 0
            local.free()
            if out is not None:
                out.free()
            return n, d2h
        n = (local.n_paths, local.n_segments, local.n_waypoints)
        d2h = local.n_waypoints * 24 + local.n_segments * 4
        local.free()
        return n, d2h

    # ---- value: device-resident ------------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        device_step()
    # one sampler for the whole job (rank 0 polls every GPU of the job): a poller per rank makes eight nvidia-smi
    # processes contend for the driver with the ranks they are supposed to observe
    sampler = ClockSampler(",".join(str(i) for i in range(info.world)) if info.rank == 0 else None)
    stage_acc: dict[str, list[float]] = {}
    launches0 = ctx.kernel_launches()
    # L2 is flushed between timed iterations (a 256 MB write, outside the timed intervals): exactly K steps are timed,
    # each bracketed by CUDA events on the library's stream; the K intervals are summed.
    flush = torch.empty(256 * 2**20, dtype=torch.uint8, device=dev)

    def l2_flush():
        flush.zero_()
        torch.cuda.synchronize()

    barrier()
    sampler.start()
    counts = None
    dev_ms = wall_ms = bracket_ms = 0.0
    for _ in range(args.steps):
        l2_flush()
        t0 = time.perf_counter()
        ctx.timer_start()
        counts = device_step()
        bracket_ms += ctx.timer_stop()
        # the step's device time: CUDA events recorded by the library on its stream at the entry of nb2_plan_mesh and behind
        # the plan's last kernel (the bracket above additionally holds this interpreter's time between the calls)
        dev_ms += ctx.last_plan_span_ms()
        wall_ms += 1e3 * (time.perf_counter() - t0)
        for k, v in ctx.timings().items():  # stages of this step's upload and plan (read outside the timed interval)
            stage_acc.setdefault(k, []).append(v)
    barrier()
    clocks = sampler.stop()
    launches = ctx.kernel_launches() - launches0
    dev_ms = max_over_ranks(dev_ms, info.world, dev)
    wall_ms = max_over_ranks(wall_ms, info.world, dev)
    bracket_ms = max_over_ranks(bracket_ms, info.world, dev)
    ms_per_step = dev_ms / args.steps

    # ---- e2e ---------------------------------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        e2e_step()
    barrier()
    totals, d2h_bytes = None, 0
    e2e_total = 0.0
    for _ in range(args.steps):
        l2_flush()
        barrier()
        t0 = time.perf_counter()
        totals, d2h_bytes = e2e_step()
        torch.cuda.synchronize()
        e2e_total += 1e3 * (time.perf_counter() - t0)
        if trace:
            print(f"[rank {info.rank}] e2e step {1e3 * (time.perf_counter() - t0):.3f} ms", file=sys.stderr)
    barrier()
    e2e_ms = max_over_ranks(e2e_total, info.world, dev) / args.steps
    h2d_bytes = int(blob.data.size + tri.size * 4)  # per job: with N ranks each copies 1 / N of it

    if info.rank != 0:
        return

    # ---- e2e through the reference-facing boundary (N = 1): the C++ plugin, created by alias through the factory, planning a
    # host pcl::PolygonMesh into nested noether::ToolPaths (plugin/b200_plugins.cpp; driver: tests/cpp/plugin_shim_test.cpp
    # --bench).  Its own process: this one's context is closed first so that the two do not share the device.
    plugin_e2e = None
    if info.world == 1 and args.workload in WORKLOADS and os.environ.get("NB2_BENCH_NO_PLUGIN") != "1":
        nx, ny = WORKLOADS[args.workload][:2]
        exe = os.path.join(ROOT, "plugin", "build", "plugin_shim_test")
        lib = os.path.join(ROOT, "plugin", "build", "libnoether_b200_plugins.so")
        try:
            if not (os.path.exists(exe) and os.path.exists(lib)):
                subprocess.run(["make", "-C", os.path.join(ROOT, "plugin")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            r = subprocess.run([exe, lib, "-", "--bench", str(nx), str(ny), repr(spacing), str(args.steps), str(max(args.warmup, 3))],
                               capture_output=True, text=True, timeout=600)
            plugin_e2e = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
        except Exception as e:
            plugin_e2e = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    # ---- roofline of the dominant kernel (live CUDA-event stage times averaged over the timed steps) -------------
    peak, peak_src = measured_peaks()
    lat = ctx.last_lattice()
    P_planes = int(lat.num_planes)
    n_paths, n_segments, n_waypoints = totals if info.world == 1 else totals
    stage_ms = {k: float(np.mean(v)) for k, v in stage_acc.items()}
    kernel_stages = {k: v for k, v in stage_ms.items() if k not in ("h2d", "d2h", "setup", "validate") and not k.startswith("host:")}
    dominant = max(kernel_stages, key=kernel_stages.get)
    # per-rank quantities for the dominant kernel's launch on this rank
    W_rank, S_rank = counts[2], counts[1]
    H_rank = max(W_rank - S_rank, 0)
    B = algorithmic_bytes(dominant, V, F, H_rank, W_rank, S_rank, P_planes, blob.point_step)
    achieved = (B / 1e9) / (kernel_stages[dominant] / 1e3) if B else None
    B_plan = 12 * F + 24 * V + 24 * n_waypoints + 4 * (n_segments + 1) + 4 * (n_paths + 1)
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None,
                "traffic": ncu_traffic(args.workload, dominant) if info.world == 1 else None,
                "algorithmic_bytes_per_launch": B, "launch_ms": kernel_stages[dominant], "peak_source": peak_src,
                "stage_ms": stage_ms,
                "plan": {"algorithmic_bytes": B_plan, "achieved": (B_plan / 1e9) / (ms_per_step / 1e3),
                         "frac": (B_plan / 1e9) / (ms_per_step / 1e3) / peak,
                         "note": "B_plan = 12F + 24V + 24W + 4(S+1) + 4(P+1) over the whole device step"}}

    cpu = None
    if not args.no_cpu_baseline:
        import oracle as orc
        orc.build()
        s = cpu_reference_sample(xyz, nrm, tri, spacing, args.cpu_planes)
        cpu = {"value": s["tri_per_s"], "unit": "triangles/s", "cores": 1, "kind": "port",
               "host_cores_available": os.cpu_count(),
               "sample": (f"prologue + {s['planes_sampled']} of {s['planes']} lattice planes "
                          f"({s['t_measured_s']:.1f} s measured), extrapolated linearly to the full lattice "
                          f"({s['t_plan_extrapolated_s']:.1f} s per plan); reference-structured single-thread oracle")}

    # second, fairer CPU figure (SURVEY §8 d): the same algorithm with the cells binned by lattice interval and the lattice
    # cut into plane slabs over all host cores; its own process (it forks), whole plan, N = 1 only
    cpu_opt = None
    if cpu is not None and info.world == 1 and args.workload in ("cfg2", "cfg3"):
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "cpu_optimised.py"), args.workload],
                               capture_output=True, text=True, timeout=180)
            o = json.loads(r.stdout.strip().splitlines()[-1])
            cpu_opt = {"value": o["tri_per_s"], "unit": "triangles/s", "cores": o["workers"], "kind": "port",
                       "sample": (f"whole plan: prologue + all {o['planes']} lattice planes in {o['t_plan_s']:.2f} s; oracle with "
                                  "the cells binned by lattice interval (each plane visits only the triangles that can "
                                  "touch it; same output as the every-cell scan), contiguous plane slabs over forked workers")}
        except Exception as e:  # a reported baseline must not take the bench line down with it
            cpu_opt = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    # e2e_compact: the C ABI called from Python with pinned host arrays in, the compact 24-byte-per-waypoint SoA result in
    # pinned host memory out.  e2e (the headline): at N = 1 the C++ plugin call a noether application makes — pageable
    # pcl::PolygonMesh in (one std::vector per face), nested std::vector<Eigen::Isometry3d> out (128 bytes per waypoint),
    # every plan on a mesh the device has not seen; at N > 1 the sharded C-ABI path (the plugin drives one GPU).
    e2e_compact = {"value": F / (e2e_ms / 1e3), "unit": "triangles/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(d2h_bytes),
                   "boundary": "C ABI (nb2_plan_mesh) from Python: pinned host cloud + triangles in, SoA positions / normals / segment lengths out"}
    if plugin_e2e and "ms_per_plan" in plugin_e2e:
        has_nrm = blob.off_normal >= 0
        e2e = {"value": F / (plugin_e2e["ms_per_plan"] / 1e3), "unit": "triangles/s", "ms_per_step": plugin_e2e["ms_per_plan"],
               "ms_best": plugin_e2e["ms_best"],
               # what crosses PCIe: the cloud compacted to {xyz, normal} records by the staging threads, the triangles, and back
               # the compact result; the host-side bytes the call consumes and produces are listed next to it
               "h2d_bytes_per_step": int(V * (24 if has_nrm else 12) + 12 * F),
               "d2h_bytes_per_step": int(plugin_e2e["waypoints"] * 24 + plugin_e2e["segments"] * 4),
               "host_in_bytes_per_step": plugin_e2e["host_in_bytes"], "host_out_bytes_per_step": plugin_e2e["host_out_bytes"],
               "waypoints": plugin_e2e["waypoints"], "tool_paths": plugin_e2e["tool_paths"],
               "boundary": "C++ plugin: Factory::createToolPathPlanner('PlaneSlicer')->plan(pcl::PolygonMesh) -> noether::ToolPaths "
                           "(nested Eigen::Isometry3d vectors); two meshes with different fingerprints alternate, so every plan "
                           "flattens, fingerprints, uploads, plans, downloads and expands; wall clock"}
    else:
        e2e = dict(e2e_compact)
        if plugin_e2e:
            e2e["plugin_boundary"] = plugin_e2e

    line = {
        "metric": "triangles sliced/sec", "value": F / (ms_per_step / 1e3), "unit": "triangles/s",
        "n_gpus": info.world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, desc, info.world, V, F),
        "result": {"planes": P_planes, "tool_paths": int(n_paths), "segments": int(n_segments), "waypoints": int(n_waypoints),
                   "events_recorded_from_the_interpreter_ms_per_step": bracket_ms / args.steps,
                   "wall_clock_ms_per_step": wall_ms / args.steps},
        "clocks": clocks,
        "e2e": e2e,
        "e2e_compact": e2e_compact,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "cpu_baseline_optimised": cpu_opt,
        "normals_from_mesh_faces_ms": normals_ms,
    }
    emit(line)


if __name__ == "__main__":
    main()
