#!/usr/bin/env python
"""bench.py — headline benchmark of the bevy_terrain hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # B200 arm (this framework)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on the host cores

metric  : Mverts/s generated (strip-ordered vertex stream = the reference's indirect vertex_count)
workload: BASELINE.json configs[1] — planar 4096^2 heightmap, 8 LODs, grid 64, bilinear, normals on:
          one step = one frame of one view = select_tiles + gen_vertices (~41 M vertices)
value   : device-resident throughput (uniforms + atlas already in HBM), all ranks (N>1: independent
          replicas over different views — the per-frame path does not shard)
e2e     : same metric through the public API with host inputs: per step the host computes the view's
          Taylor approximation, H2Ds uniforms + tile-tree table from pinned memory, runs both kernels
          and D2Hs the tile list + indirect record (the 1.5 GB vertex stream stays on the device, where
          its consumer — the draw — reads it, as in the reference)
preprocess (sub-object): tiles/s of the sharded config-5-shaped pyramid (height + albedo) at N GPUs, run by
          bt_preprocessor_run_sharded (in-library NCCL, overlapped), with a device-side hash proof that the
          sharded result equals the single-GPU pyramid
extra   : Mverts/s + roofline of configs[2] (16384^2, 12 LODs, morph+blend) and configs[3] (spherical)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALGORITHMIC_BYTES_PER_VERTEX = 80  # 40 B written + 5 bilinear taps x 4 texels x 2 B
COMPULSORY_BYTES_PER_VERTEX = 40   # what must reach DRAM: the record (texel reads are L1/L2 hits, < 0.3 B/vertex)
NUM_VIEWS = 8
VIEW_SEED = 0x5EED0010
VPT = 2 * 64 * 66


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--source", type=int, default=4096, help="config-2 heightmap size")
    ap.add_argument("--lods", type=int, default=8)
    ap.add_argument("--preprocess-source", type=int, default=65536, help="config-5 source size (0 = skip)")
    ap.add_argument("--preprocess-lods", type=int, default=8)
    ap.add_argument("--cpu-sample-mverts", type=float, default=64.0, help="cpu_baseline: vertices per pass (million; capped at view 0's count)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="cpu_baseline: CPU time budget of the vertex sample")
    ap.add_argument("--ref-sample-mverts", type=float, default=0.0, help="reference arm: vertices per step (millions); 0 = size from --ref-budget-s")
    ap.add_argument("--ref-budget-s", type=float, default=100.0, help="reference arm: target wall time of the whole --steps/--warmup run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="all", help="extra workloads beside the headline configs[1]: all | none | configs[2] | configs[3]")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# host threads: decided HERE, not inherited (torch.distributed.run exports OMP_NUM_THREADS=1)
# ------------------------------------------------------------------------------------------------
_HOST_THREADS = None
_HOST_AFFINITY = None


def host_threads() -> int:
    """the cores this process may run on, sampled ONCE at start-up: after libgomp has initialised with OMP_PROC_BIND the initial
    thread is bound to a single place and sched_getaffinity would report 1"""
    global _HOST_THREADS
    if _HOST_THREADS is None:
        global _HOST_AFFINITY
        try:
            _HOST_AFFINITY = set(os.sched_getaffinity(0))
            _HOST_THREADS = max(1, len(_HOST_AFFINITY))
        except AttributeError:
            _HOST_THREADS = max(1, os.cpu_count() or 1)
    return _HOST_THREADS


def unbind_main_thread():
    """OMP_PROC_BIND (wanted for the cpu_baseline leg) makes libgomp bind the INITIAL thread - this Python thread, which drives the
    GPU frames - to hardware thread 0 as soon as torch loads libgomp; the frame loop then shares that core with whatever else runs
    there (measured: e2e 0.440 ms per frame bound, 0.420 ms unbound).  The OpenMP workers keep their places."""
    if _HOST_AFFINITY:
        try:
            os.sched_setaffinity(0, _HOST_AFFINITY)
        except (AttributeError, OSError):
            pass


def pin_openmp_env(bind: bool):
    """must run before anything loads libgomp (torch, the oracle): thread count = the cores this process may run on; with `bind`
    the threads are bound to hardware threads (close) so that the CPU arm is reproducible from box to box.  Ranks of a multi-GPU
    run must NOT bind: every rank's initial thread would land on the same first place and the ranks would share one core."""
    n = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(n)
    if bind:
        os.environ["OMP_PROC_BIND"] = "close"
        os.environ["OMP_PLACES"] = "threads"
    else:
        os.environ["OMP_PROC_BIND"] = "false"
        os.environ.pop("OMP_PLACES", None)
    return n


def spread(values):
    v = np.asarray(values, dtype=np.float64)
    return {"median": float(np.median(v)), "min": float(v.min()), "max": float(v.max()), "n": int(len(v))}


# ------------------------------------------------------------------------------------------------
# workload description shared by both arms
# ------------------------------------------------------------------------------------------------
def view_positions(rank: int):
    """8 seeded camera positions 20 above the approximate surface height"""
    rng = np.random.default_rng(VIEW_SEED + rank)
    out = []
    for _ in range(NUM_VIEWS):
        out.append((float(rng.uniform(-800, 800)), -100.0 + 250.0 + 20.0, float(rng.uniform(-800, 800))))
    return out


def workload_config(args, n_gpus):
    return {
        "workload": f"configs[1]: planar {args.source}^2 R16 heightmap, {args.lods} LODs, texture 512 border 2, grid 64, "
                    "tree_size 8, bilinear, NORMALS on, MORPH/BLEND off; step = select_tiles + gen_vertices for one of "
                    f"{NUM_VIEWS} seeded views (camera 20 above the surface, ~4.8k tiles, ~41M strip vertices)",
        "views_per_rank": NUM_VIEWS,
        "parallelism": "single-GPU per view" if n_gpus == 1 else f"{n_gpus} independent replicas (per-frame path does not shard)",
        "l2": "inputs larger than L2: each step reads a different view's 256 MiB atlas and writes ~1.6 GB of vertices (L2 = 126 MB)",
        "algorithmic_bytes_per_vertex": ALGORITHMIC_BYTES_PER_VERTEX,
    }


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons DURING the timed region"""

    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [s for t, s in self.samples if t0 - 0.05 <= t <= t1 + 0.15] or [s for _, s in self.samples[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def build_terrain(api, ctx, args):
    """config-2 terrain: synthetic source -> GPU pyramid (also timed: single-GPU preprocess of config 2)"""
    import torch

    from bevy_terrain_b200 import _ffi as F
    model = api.TerrainModel.planar((0.0, -100.0, 0.0), 2000.0, 0.0, 500.0)            # examples/planar.rs:6-9,53
    cfg = api.TerrainConfig(lod_count=args.lods, model=model, atlas_size=8 * 64)       # tree_size^2 tiles per LOD resident
    cfg.add_attachment(api.AttachmentConfig(name="height", texture_size=512, border_size=2, mip_level_count=4,
                                            format=api.AttachmentFormat.R16))          # examples/planar.rs:57-63
    src = ctx.synth_source(F.BT_FORMAT_R16, args.source, args.source, seed=0x5EED0001)
    pre = api.Preprocessor(ctx, cfg)
    pre.preprocess_tile(api.PreprocessDataset(device_pixels=src, width=args.source, height=args.source, lod_range=range(0, args.lods)))
    pre.reserve()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pre.run()
    e1.record()
    torch.cuda.synchronize()
    n_tiles = len(pre.tiles())
    info = {"tiles": n_tiles, "ms": e0.elapsed_time(e1), "tiles_per_s": n_tiles / (e0.elapsed_time(e1) * 1e-3)}
    return cfg, pre, src, info


def time_views(torch, views, flags, rounds, select=True):
    """device-timed select + gen over `views`, `rounds` passes after one warm-up pass; also per-launch gen_vertices events"""
    for v in views:
        v["tree"].select_tiles()
        v["tree"].gen_vertices(flags)
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gen_events, nv = [], 0
    g0.record()
    for _ in range(rounds):
        for v in views:
            if select:
                v["tree"].select_tiles()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            v["tree"].gen_vertices(flags)
            b.record()
            gen_events.append((a, b, v["verts"]))
            nv += v["verts"]
    g1.record()
    torch.cuda.synchronize()
    ms = g0.elapsed_time(g1)
    gen_ms = float(np.mean([a.elapsed_time(b) for a, b, _ in gen_events]))
    gen_verts = float(np.mean([n for _, _, n in gen_events]))
    return {"Mverts_per_s": nv / (ms * 1e-3) / 1e6, "ms_per_step": ms / (rounds * len(views)), "gen_vertices_ms": gen_ms,
            "verts_per_step": gen_verts, "steps": rounds * len(views)}


def roofline_block(gen_ms, gen_verts, traffic_per_vertex):
    peak, peak_src = hbm_peak()
    achieved = ALGORITHMIC_BYTES_PER_VERTEX * gen_verts / (gen_ms * 1e-3) / 1e9
    compulsory = COMPULSORY_BYTES_PER_VERTEX * gen_verts
    out = {"kernel": "gen_vertices_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
           "peak_source": peak_src, "avg_launch_ms": gen_ms, "algorithmic_bytes_per_launch": ALGORITHMIC_BYTES_PER_VERTEX * gen_verts,
           "compulsory_bytes": compulsory, "frac_dram_compulsory": compulsory / (gen_ms * 1e-3) / 1e9 / peak,
           "traffic": None, "frac_dram": None}
    if traffic_per_vertex:
        out["traffic"] = traffic_per_vertex["bytes_per_vertex"] * gen_verts
        out["frac_dram"] = out["traffic"] / (gen_ms * 1e-3) / 1e9 / peak
        out["traffic_source"] = traffic_per_vertex["source"]
    out["note"] = ("achieved/frac use SURVEY 8d's ALGORITHMIC 80 B/vertex (40 written + 40 of bilinear-tap reads, which L1/L2 serve); "
                   "frac_dram_compulsory = 40 B/vertex x vertices / launch time / peak is the DRAM-side fraction the north-star target "
                   "(>= 0.70) is held against; traffic = ncu dram__bytes per vertex of the committed capture x this run's vertices")
    return out


def load_traffic():
    for name in ("gen_vertices_traffic_r02.json", "gen_vertices_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            d = json.load(open(p))
            if "dram_bytes_per_vertex" in d:
                return {"bytes_per_vertex": float(d["dram_bytes_per_vertex"]), "source": f"profiles/{name}"}
            if "dram_bytes_per_launch" in d:   # round-1 capture: launches 5/6 = views 5/6 of rank 0 (36.0 M vertices each on average)
                return {"bytes_per_vertex": float(d["dram_bytes_per_launch"]) / float(d.get("verts_per_launch", 36.0e6)), "source": f"profiles/{name}"}
    return None


def run_b200(args):
    import torch
    import torch.distributed as dist

    from bevy_terrain_b200 import _ffi as F
    from bevy_terrain_b200 import api

    rank, world, local = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    unbind_main_thread()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = api.Context(local)
    flags = F.BT_FLAG_NORMALS
    default_flags = F.BT_FLAG_MORPH | F.BT_FLAG_BLEND | F.BT_FLAG_NORMALS
    cfg, pre, src_ptr, pre2 = build_terrain(api, ctx, args)
    vcfg = api.TerrainViewConfig(grid_size=64)
    views = []
    for pos in view_positions(rank):
        atlas = api.TileAtlas(ctx, cfg)
        tree = api.TileTree(atlas, vcfg, max_vertices=6400 * VPT)
        tree.add_existing_tiles(pre.tiles())
        api.load_resident_tiles(tree, pre, pos)
        tree.prepare()
        tree.select_tiles()
        tiles, ind = tree.readback()
        views.append({"tree": tree, "atlas": atlas, "pos": pos, "verts": int(ind.vertex_count), "tiles": len(tiles)})
    pre.close()
    ctx.free(src_ptr)

    def step_resident(i, gen_events=None):
        v = views[i % NUM_VIEWS]
        v["tree"].select_tiles()
        if gen_events is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            v["tree"].gen_vertices(flags)
            b.record()
            gen_events.append((a, b, v["verts"]))
        else:
            v["tree"].gen_vertices(flags)
        return v["verts"]

    def step_e2e(i):
        v = views[i % NUM_VIEWS]
        t = v["tree"]
        # bt_view_run_frame: host f64 Taylor approximation + H2D (pinned) of uniforms and tile-tree table, select_tiles,
        # gen_vertices, and the D2H of tile list + indirect record (blocking for the host; overlaps gen_vertices on the device)
        tiles, ind, _ = t.frame_step(flags)
        return int(ind.vertex_count), len(tiles)

    n_warm = max(args.warmup, NUM_VIEWS)      # EVERY view is warm before either timed region (lazy state, page tables, caches)

    # ---- device-resident timed region --------------------------------------------------------------
    for i in range(n_warm):
        step_resident(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gen_events, verts = [], 0
    barrier()
    t_wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        verts += step_resident(i, gen_events)
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # ---- end-to-end timed region ----------------------------------------------------------------------
    for i in range(n_warm):
        step_e2e(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_verts, e2e_tiles = 0, 0
    barrier()
    w0 = time.time()
    f0.record()
    for i in range(args.steps):
        nv, nt = step_e2e(i)
        e2e_verts += nv
        e2e_tiles += nt
    f1.record()
    barrier()
    w1 = time.time()
    e2e_ms = max(f0.elapsed_time(f1), (w1 - w0) * 1e3)   # the blocking readback makes wall >= device time

    # ---- flag variants (device-resident, every view, 2 passes) ---------------------------------------------
    variants = {}
    for name, vflags in (("reference_default_MORPH_BLEND_plus_NORMALS", default_flags),
                         ("vertex_stage_only_no_normals", 0), ("MORPH_BLEND_no_normals", F.BT_FLAG_MORPH | F.BT_FLAG_BLEND)):
        variants[name] = time_views(torch, views, vflags, 2)

    # ---- reduce over ranks (max time, sum work) ---------------------------------------------------------
    def reduce(vals, op):
        if world == 1:
            return vals
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return t.tolist()

    dv = variants["reference_default_MORPH_BLEND_plus_NORMALS"]
    dv_ms = dv["ms_per_step"] * dv["steps"]
    ms_max, e2e_ms_max, dv_ms_max = reduce([ms, e2e_ms, dv_ms], dist.ReduceOp.MAX if world > 1 else None)
    verts_all, e2e_verts_all, dv_verts_all = reduce([float(verts), float(e2e_verts), dv["verts_per_step"] * dv["steps"]],
                                                    dist.ReduceOp.SUM if world > 1 else None)

    gen_ms = [a.elapsed_time(b) for a, b, _ in gen_events]
    gen_verts = [n for _, _, n in gen_events]
    gen_avg_ms = float(np.mean(gen_ms))
    traffic = load_traffic()
    roofline = roofline_block(gen_avg_ms, float(np.mean(gen_verts)), traffic)
    roofline["reference_default_flags"] = roofline_block(dv["gen_vertices_ms"], dv["verts_per_step"], None)

    # ---- free the per-frame state, then the other workloads ---------------------------------------------
    frame_state = {"tiles_per_view": [v["tiles"] for v in views], "verts_per_view": [v["verts"] for v in views]}
    host_inputs = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        host_inputs = capture_host_inputs(ctx, views[0], cfg, vcfg)
    h2d_bytes = 1168 + views[0]["tree"].entry_count * 8   # sizeof(ViewUniforms) + tile-tree table
    for v in views:
        v["tree"].close(); v["atlas"].close()
    views.clear()
    torch.cuda.empty_cache()

    extra = {}
    if world == 1 and args.workload != "none":
        if args.workload in ("all", "configs[2]"):
            extra["configs[2]"] = run_config2(torch, api, F, ctx)
        if args.workload in ("all", "configs[3]"):
            extra["configs[3]"] = run_config3(torch, api, F, ctx)
        torch.cuda.empty_cache()

    preprocess = None
    if args.preprocess_source:
        preprocess = run_preprocess_bench(api, F, ctx, args, rank, world, barrier, reduce)
        if preprocess is not None:
            preprocess["config2_single_gpu"] = pre2

    cpu_baseline = None
    if host_inputs is not None:
        cpu_baseline = cpu_vertex_baseline(host_inputs, args.cpu_sample_mverts, args.cpu_seconds)
        if preprocess is not None:
            preprocess["cpu_baseline"] = cpu_preprocess_baseline(api, F, ctx)

    if rank == 0:
        out = {
            "metric": "Mverts/s generated (strip order) — select_tiles + gen_vertices per frame",
            "value": verts_all / (ms_max * 1e-3) / 1e6,
            "unit": "Mverts/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_verts_all / (e2e_ms_max * 1e-3) / 1e6, "unit": "Mverts/s",
                    "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(16 * e2e_tiles / args.steps + 32),
                    "ms_per_step": e2e_ms_max / args.steps,
                    "note": "bt_view_run_frame per step; the H2D bytes (uniform block + tile-tree table) are written to pinned host memory every step and pulled over PCIe by the select kernel itself (BT_FRAME_FETCH=0: two cudaMemcpyAsync instead); the D2H is the tile list + indirect record, the vertex stream stays in HBM for its on-device consumer (as the reference's draw)"},
            "gpu_launches": 2 * args.steps,
            # the reference's DEFAULT vertex flags (MORPH | BLEND, terrain_material.rs:404-409) + the fused normals: first-class beside `value`
            "value_reference_default_flags": {"value": dv_verts_all / (dv_ms_max * 1e-3) / 1e6, "unit": "Mverts/s", "flags": "MORPH|BLEND|NORMALS",
                                              "ms_per_step": dv["ms_per_step"], "gen_vertices_ms": dv["gen_vertices_ms"]},
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "frame": frame_state,
            "variants_rank0": variants,
            "extra": extra,
            "preprocess": preprocess,
        }
        emit(out)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# extra workloads: configs[2] and configs[3] at full size (the shapes of tests/test_gpu_fullsize_config2.py / _spherical.py)
# ------------------------------------------------------------------------------------------------
def frame_views(api, ctx, cfg, vcfg, pre, positions, max_tiles, attachments=(0,)):
    views = []
    for pos in positions:
        atlas = api.TileAtlas(ctx, cfg)
        tree = api.TileTree(atlas, vcfg, max_vertices=max_tiles * VPT)
        tree.add_existing_tiles(pre.tiles(0))
        api.load_resident_tiles(tree, pre, pos, attachments=attachments)
        tree.prepare()
        tree.select_tiles()
        tiles, ind = tree.readback()
        views.append({"tree": tree, "atlas": atlas, "pos": pos, "verts": int(ind.vertex_count), "tiles": len(tiles)})
    return views


def run_config2(torch, api, F, ctx):
    """planar 16384^2 height + albedo, terrain with 12 LODs (data resolves LODs 0..5), camera 1 above the surface, MORPH + BLEND + NORMALS"""
    S, L, DATA_LODS = 16384, 12, 6
    model = api.TerrainModel.planar((0.0, -100.0, 0.0), 2000.0, 0.0, 500.0)
    cfg = api.TerrainConfig(lod_count=L, model=model, atlas_size=DATA_LODS * 64)
    cfg.add_attachment(api.AttachmentConfig(name="height", texture_size=512, border_size=2, mip_level_count=4, format=api.AttachmentFormat.R16))
    cfg.add_attachment(api.AttachmentConfig(name="albedo", texture_size=512, border_size=2, mip_level_count=4, format=api.AttachmentFormat.Rgba8))
    srcs = [ctx.synth_source(F.BT_FORMAT_R16, S, S, seed=0x5EED0001), ctx.synth_source(F.BT_FORMAT_RGBA8, S, S, seed=0x5EED0002)]
    pre = api.Preprocessor(ctx, cfg)
    for a in (0, 1):
        pre.preprocess_tile(api.PreprocessDataset(attachment_index=a, device_pixels=srcs[a], width=S, height=S, lod_range=range(0, DATA_LODS)))
    pre.reserve()
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(); pre.run(); p1.record(); torch.cuda.synchronize()
    rng = np.random.default_rng(VIEW_SEED + 2)
    positions = [(float(rng.uniform(-700, 700)), -100.0 + 250.0 + 1.0, float(rng.uniform(-700, 700))) for _ in range(4)]
    vcfg = api.TerrainViewConfig(grid_size=64)
    views = frame_views(api, ctx, cfg, vcfg, pre, positions, 12000, attachments=(0, 1))
    flags = F.BT_FLAG_MORPH | F.BT_FLAG_BLEND | F.BT_FLAG_NORMALS
    r = time_views(torch, views, flags, 4)
    # the "trilinear + aniso" sampler of the config's name + the per-vertex albedo stream (44 B / vertex written)
    rf = time_views(torch, views, flags | F.BT_FLAG_FILTER_TRILINEAR_ANISO | F.BT_FLAG_ALBEDO, 2)
    out = {"workload": "configs[2]: planar 16384^2 height R16 + albedo Rgba8, 12 LODs (data LODs 0..5), grid 64, camera 1 above the surface, "
                       "MORPH + BLEND + NORMALS; 4 seeded views", "tiles_per_view": [v["tiles"] for v in views],
           "verts_per_view": [v["verts"] for v in views], "flags": "MORPH|BLEND|NORMALS", **r,
           "roofline": roofline_block(r["gen_vertices_ms"], r["verts_per_step"], None),
           "trilinear_aniso_plus_albedo": {"flags": "MORPH|BLEND|NORMALS|FILTER_TRILINEAR_ANISO|ALBEDO", "Mverts_per_s": rf["Mverts_per_s"],
                                           "gen_vertices_ms": rf["gen_vertices_ms"], "bytes_written_per_vertex": 44,
                                           "frac_dram_compulsory": 44.0 * rf["verts_per_step"] / (rf["gen_vertices_ms"] * 1e-3) / 1e9 / hbm_peak()[0]},
           "preprocess_ms": p0.elapsed_time(p1), "preprocess_tiles": int(len(pre.tiles(0)) + len(pre.tiles(1)))}
    for v in views:
        v["tree"].close(); v["atlas"].close()
    pre.close()
    for p in srcs:
        ctx.free(p)
    return out


def run_config3(torch, api, F, ctx):
    """spherical, six 8192^2 faces, terrain with 10 LODs (data resolves LODs 0..4), the example's camera altitudes + corner / edge views"""
    S, L, DATA_LODS = 8192, 10, 5
    R, MIN_H, MAX_H = 6371000.0, -12000.0, 9000.0
    model = api.TerrainModel.sphere((0.0, 0.0, 0.0), R, MIN_H, MAX_H)
    cfg = api.TerrainConfig(lod_count=L, model=model, atlas_size=6 * DATA_LODS * 64)
    cfg.add_attachment(api.AttachmentConfig(name="height", texture_size=512, border_size=2, mip_level_count=4, format=api.AttachmentFormat.R16))
    faces = [ctx.synth_source(F.BT_FORMAT_R16, S, S, seed=0x5EED0001 + 17 * s) for s in range(6)]
    pre = api.Preprocessor(ctx, cfg)
    pre.preprocess_spherical(api.SphericalDataset(device_faces=faces, width=S, height=S, lod_range=range(0, DATA_LODS)))
    pre.reserve()
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(); pre.run(); p1.record(); torch.cuda.synchronize()

    def direction(v):
        v = np.asarray(v, dtype=np.float64)
        return v / np.linalg.norm(v)

    ALL = F.BT_FLAG_NORMALS | F.BT_FLAG_MORPH | F.BT_FLAG_BLEND
    named = [("corner_30km", direction([0.62, 0.55, 0.56]) * (R + 30000.0), ALL),
             ("minus_x_10km", direction([-1.0, 0.003, 0.002]) * (R + 10000.0), ALL),
             ("face_edge_100km", direction([1.0, 0.02, 1.0]) * (R + 100000.0), ALL),
             ("minus_x_3km_high_precision", direction([-1.0, 0.003, 0.002]) * (R + 3000.0), ALL | F.BT_FLAG_HIGH_PRECISION)]
    vcfg = api.TerrainViewConfig(grid_size=64)
    views = frame_views(api, ctx, cfg, vcfg, pre, [tuple(float(x) for x in p) for _, p, _ in named], 20000)
    per_view = {}
    for (name, _, vflags), v in zip(named, views):
        r = time_views(torch, [v], vflags, 6)
        per_view[name] = {"tiles": v["tiles"], "verts": v["verts"], "Mverts_per_s": r["Mverts_per_s"], "gen_vertices_ms": r["gen_vertices_ms"],
                          "frac_dram_compulsory": roofline_block(r["gen_vertices_ms"], r["verts_per_step"], None)["frac_dram_compulsory"]}
    r = time_views(torch, views[:3], ALL, 4)
    out = {"workload": "configs[3]: spherical, six 8192^2 R16 faces, 10 LODs (data LODs 0..4), grid 64, MORPH + BLEND + NORMALS "
                       "(+ HIGH_PRECISION for the 3 km view); aggregate over the three non-HIGH_PRECISION views", **r,
           "roofline": roofline_block(r["gen_vertices_ms"], r["verts_per_step"], None), "per_view": per_view,
           "preprocess_ms": p0.elapsed_time(p1), "preprocess_tiles": int(len(pre.tiles(0)))}
    for v in views:
        v["tree"].close(); v["atlas"].close()
    pre.close()
    for p in faces:
        ctx.free(p)
    return out


def capture_host_inputs(ctx, view, cfg, vcfg):
    """downloads one view's device inputs so the CPU oracle runs on EXACTLY the same data"""
    tree, atlas = view["tree"], view["atlas"]
    a = cfg.attachments[0]
    host_atlas = np.empty((cfg.atlas_size, a.texture_size, a.texture_size), dtype=np.uint16)
    ctx.d2h(host_atlas, atlas.device_ptr(0))
    tiles, ind = tree.readback()
    return {"tree_frame": bytes(tree.frame.view_config), "approx": bytes(tree.frame.approximation), "culling": bytes(tree.frame.culling),
            "entries": np.frombuffer(tree.entries, dtype=np.uint32).reshape(-1, 2).copy(), "atlas": host_atlas, "tiles": tiles,
            "pos": view["pos"], "lod_count": cfg.lod_count, "T": a.texture_size, "B": a.border_size,
            "mesh": bytes(cfg.model.mesh_transform()), "min_h": cfg.model.c.min_height, "max_h": cfg.model.c.max_height,
            "scale": cfg.model.scale(), "verts": view["verts"]}


def oracle_vertex_inputs(h, flags):
    from bevy_terrain_b200 import _ffi as F
    from tests import oracle as O
    vi = O.VertexInputs()
    C.memmove(C.byref(vi.view_config), h["tree_frame"], len(h["tree_frame"]))
    C.memmove(C.byref(vi.approximation), h["approx"], len(h["approx"]))
    C.memmove(C.byref(vi.mesh), h["mesh"], len(h["mesh"]))
    vi.terrain_config = F.TerrainConfigUniform(h["lod_count"], h["min_h"], h["max_h"], float(np.float32(h["scale"])))
    c = np.float32(h["T"] - 2 * h["B"])
    vi.attachments[0].size = float(c)
    vi.attachments[0].scale = float(c / np.float32(h["T"]))
    vi.attachments[0].offset = float(np.float32(h["B"]) / np.float32(h["T"]))
    vi.view_world_position[:] = [float(np.float32(v)) for v in h["pos"]]
    vi.spherical, vi.texture_size, vi.atlas_size = 0, h["T"], len(h["atlas"])
    vi.tile_tree = h["entries"].ctypes.data_as(C.POINTER(F.TileTreeEntry))
    vi.height_atlas = h["atlas"].ctypes.data
    vi.tiles = h["tiles"].ctypes.data_as(C.POINTER(F.Tile))
    vi.tile_count, vi.flags = len(h["tiles"]), flags
    return vi


def cpu_vertex_baseline(h, sample_mverts, seconds):
    """cpu_baseline leg: the oracle port (OpenMP over vertices) on a bounded sample of view 0, same inputs; >= 3 timed repeats"""
    from bevy_terrain_b200 import _ffi as F
    from tests import oracle as O
    vi = oracle_vertex_inputs(h, F.BT_FLAG_NORMALS)
    n = int(min(h["verts"], sample_mverts * 1e6))
    threads = host_threads()
    out = np.empty(n, dtype=O.VERTEX_DTYPE)
    O.lib().bto_gen_vertices(C.byref(vi), 0, n, out.ctypes.data, threads)  # warm-up pass: page faults, thread pool, caches
    rates, t_all = [], time.time()
    while True:   # about `seconds` of CPU work, at least three passes
        t0 = time.time()
        O.lib().bto_gen_vertices(C.byref(vi), 0, n, out.ctypes.data, threads)
        rates.append(n / (time.time() - t0) / 1e6)
        if (time.time() - t_all >= seconds and len(rates) >= 3) or len(rates) >= 64:
            break
    return {"value": float(np.median(rates)), "unit": "Mverts/s", "cores": threads, "kind": "port", "repeats": spread(rates),
            "omp": {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OMP_PROC_BIND", "OMP_PLACES")}, "os_cpu_count": os.cpu_count(),
            "sample": f"median of {len(rates)} timed passes over the first {n} strip vertices of view 0 of the same workload, after one warm-up pass "
                      f"(oracle/bt_oracle.c gen_vertices, OpenMP, {threads} threads = the cores this process may run on, bound close)"}


def cpu_preprocess_baseline(api, F, ctx):
    """the config-5 8192^2 two-attachment sub-problem (5 LODs, 2 x 341 tiles), CPU oracle and GPU, same process"""
    import torch

    from tests import oracle as O
    S, L = 8192, 5
    threads = host_threads()
    fmts = ((F.BT_FORMAT_R16, 0x5EED0001, api.AttachmentFormat.R16, "height"), (F.BT_FORMAT_RGBA8, 0x5EED0002, api.AttachmentFormat.Rgba8, "albedo"))
    srcs = [O.synth_source(fmt, S, S, seed=seed) for fmt, seed, _, _ in fmts]

    def cpu_once():
        t0, n = time.time(), 0
        for (fmt, _, _, _), src in zip(fmts, srcs):
            pyr = O.Pyramid(F.BT_KIND_PLANAR, fmt, 512, 2, L)
            pyr.add_datasets([O.make_dataset(src, lod_range=(0, L))])
            pyr.run(threads)
            n += len(pyr.tiles())
            pyr.close()
        return n, time.time() - t0

    cpu_once()   # warm-up (page faults of the 0.7 GB of tiles)
    runs = [cpu_once() for _ in range(3)]
    n_tiles = runs[0][0]
    cpu_rates = [n / s for n, s in runs]
    cfg = api.TerrainConfig(lod_count=L, model=api.TerrainModel.planar((0.0, -100.0, 0.0), 2000.0, 0.0, 500.0))
    for _, _, afmt, name in fmts:
        cfg.add_attachment(api.AttachmentConfig(name=name, texture_size=512, border_size=2, format=afmt))
    dev = [ctx.synth_source(fmt, S, S, seed=seed) for fmt, seed, _, _ in fmts]
    gpu_ms = []
    for _ in range(3):
        pre = api.Preprocessor(ctx, cfg)
        for a in (0, 1):
            pre.preprocess_tile(api.PreprocessDataset(attachment_index=a, device_pixels=dev[a], width=S, height=S, lod_range=range(0, L)))
        pre.reserve()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pre.run(); e1.record(); torch.cuda.synchronize()
        gpu_ms.append(e0.elapsed_time(e1))
        pre.close()
    for p in dev:
        ctx.free(p)
    gpu_rate = n_tiles / (float(np.median(gpu_ms)) * 1e-3)
    return {"value": float(np.median(cpu_rates)), "unit": "tiles/s", "cores": threads, "kind": "port", "repeats": spread(cpu_rates),
            "sample": f"config-5 sub-problem: {S}^2 height R16 + albedo Rgba8 -> {L} LODs = {n_tiles} tiles, oracle/bt_oracle_pre.c (OpenMP, {threads} threads), "
                      "median of 3 runs after one warm-up, same process and same synthetic sources as the GPU number beside it",
            "gpu_same_subproblem": {"tiles_per_s": gpu_rate, "ms": float(np.median(gpu_ms)), "ratio_gpu_over_cpu": gpu_rate / float(np.median(cpu_rates))}}


def run_preprocess_bench(api, F, ctx, args, rank, world, barrier, reduce):
    """config-5-shaped pyramid: S^2 height (R16) + albedo (Rgba8) -> full tile pyramid, rows sharded over ranks; the whole sharded
    job runs inside the library (bt_preprocessor_run_sharded: ONE grouped NCCL all-gather on a side stream, overlapped with the
    interior stitch).  Outside the timed region a device-side hash proves sharded == single-GPU."""
    import torch
    import torch.distributed as dist
    S, L = args.preprocess_source, args.preprocess_lods
    model = api.TerrainModel.planar((0.0, -100.0, 0.0), 2000.0, 0.0, 500.0)
    cfg = api.TerrainConfig(lod_count=L, model=model)
    cfg.add_attachment(api.AttachmentConfig(name="height", texture_size=512, border_size=2, format=api.AttachmentFormat.R16))
    cfg.add_attachment(api.AttachmentConfig(name="albedo", texture_size=512, border_size=2, format=api.AttachmentFormat.Rgba8))
    n_rows = 1 << (L - 1)
    if world > n_rows:
        return None
    comm = api.Comm.from_torch_distributed(ctx) if world > 1 else None
    # this rank's band of source rows (+1 row of bilinear footprint on each side)
    y0 = (rank * n_rows // world) * (S // n_rows)
    y1 = ((rank + 1) * n_rows // world) * (S // n_rows)
    r0, r1 = max(y0 - 1, 0), min(y1 + 1, S)
    fmts = ((F.BT_FORMAT_R16, 0x5EED0001), (F.BT_FORMAT_RGBA8, 0x5EED0002))
    srcs = [ctx.synth_source(fmt, S, r1 - r0, x0=0, y0=r0, seed=seed) for fmt, seed in fmts]
    torch.cuda.synchronize()

    def one_run(keep=False):
        pre = api.Preprocessor(ctx, cfg, rank=rank, world=world)
        for a in (0, 1):
            kw = dict(row_begin=r0, row_count=r1 - r0) if world > 1 else {}
            pre.preprocess_tile(api.PreprocessDataset(attachment_index=a, device_pixels=srcs[a], width=S, height=S, lod_range=range(0, L), **kw))
        pre.reserve()   # device store, exchange buffers and copy tables allocated outside the timed region (one-time cost, like the atlas itself)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if world == 1:
            pre.run()
        else:
            pre.run_sharded(comm)
        e1.record()
        barrier()
        n = len(pre.tiles(0)) + len(pre.tiles(1))
        ms = e0.elapsed_time(e1)
        if keep:
            return ms, n, pre
        pre.close()
        return ms, n, None

    one_run()                       # warm-up (allocations, NCCL channels)
    times = []
    for _ in range(2):
        ms, n_tiles, _ = one_run()
        times.append(ms)
    ms, n_tiles, pre = one_run(keep=True)
    times.append(ms)
    ms = float(np.median(times))
    # ---- parity proof outside the timed region: sum of the ranks' device hashes == the single-GPU pyramid's hash
    hashes = [pre.content_hash(0), pre.content_hash(1)]
    pre.close()
    sharded_equals_single = None
    if world > 1:
        sums = comm.all_reduce_sum_u64(hashes)
        verdict = torch.zeros(1, dtype=torch.int64, device="cuda")
        if rank == 0:   # rank 0 recomputes the WHOLE pyramid on one GPU from the whole source (24 GiB + 32 GiB fit a B200)
            for p in srcs:
                ctx.free(p)
            srcs.clear()
            torch.cuda.empty_cache()
            whole = [ctx.synth_source(fmt, S, S, seed=seed) for fmt, seed in fmts]
            single = api.Preprocessor(ctx, cfg)
            for a in (0, 1):
                single.preprocess_tile(api.PreprocessDataset(attachment_index=a, device_pixels=whole[a], width=S, height=S, lod_range=range(0, L)))
            single.run()
            ok = sums == [single.content_hash(0), single.content_hash(1)]
            single.close()
            for p in whole:
                ctx.free(p)
            verdict[0] = 1 if ok else 0
        dist.broadcast(verdict, 0)
        sharded_equals_single = bool(int(verdict.item()))
        if rank == 0:
            print(f"MULTI_GPU_PREPROCESS_HASH world={world} source={S} sharded_equals_single={sharded_equals_single} "
                  f"hash_height={sums[0]:#018x} hash_albedo={sums[1]:#018x}", file=sys.stderr)
    for p in srcs:
        ctx.free(p)
    info = comm.info() if comm else None
    if comm:
        comm.close()
    ms_max, = reduce([ms], dist.ReduceOp.MAX if world > 1 else None)
    tiles_all, = reduce([float(n_tiles)], dist.ReduceOp.SUM if world > 1 else None)
    bytes_per_tile = {"R16": 2 * (512 * 512 + S * S / ((4 ** L - 1) / 3)), "Rgba8": 4 * (512 * 512 + S * S / ((4 ** L - 1) / 3))}
    alg = (bytes_per_tile["R16"] + bytes_per_tile["Rgba8"]) * (tiles_all / 2)
    peak, _ = hbm_peak()
    return {"metric": "tiles/s preprocessed (split + downsample + stitch, height R16 + albedo Rgba8)", "source": f"{S}^2", "lods": L,
            "tiles": int(tiles_all), "ms": ms_max, "ms_repeats": spread(times), "tiles_per_s": tiles_all / (ms_max * 1e-3), "n_gpus": world, "scaling": "strong",
            "algorithmic_GBps_per_gpu": alg / (ms_max * 1e-3) / 1e9 / world, "roofline_frac": alg / (ms_max * 1e-3) / 1e9 / world / peak,
            "algorithmic_bytes_per_tile": bytes_per_tile, "sharded_equals_single": sharded_equals_single,
            "exchange": "none" if world == 1 else f"in-library: bt_preprocessor_run_sharded, ONE grouped ncclAllGather (sub-root tiles + border strips of both "
                                                  f"attachments) on the communicator's stream, overlapped with the interior stitch; nccl {info['nccl_version']}"}


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU port of the reference path (oracle) on the host cores
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank, world, _ = dist_setup(args.gpus)
    if rank != 0:
        return
    from bevy_terrain_b200 import _ffi as F
    from tests import oracle as O
    from tests.clipmap_ref import ClipmapRef
    threads = host_threads()
    t_setup = time.time()
    model = O.terrain_model(F.BT_KIND_PLANAR, (0.0, -100.0, 0.0), 2000.0, 0.0, 500.0)
    vcfg = O.default_view_config(grid_size=64)
    vc = O.view_config_uniform(vcfg, model)
    mesh = O.mesh_transform(model)
    src = O.synth_source(F.BT_FORMAT_R16, args.source, args.source, seed=0x5EED0001)
    pyr = O.Pyramid(F.BT_KIND_PLANAR, F.BT_FORMAT_R16, 512, 2, args.lods)
    pyr.add_datasets([O.make_dataset(src, lod_range=(0, args.lods))])
    t0 = time.time()
    pyr.run(threads)
    pre_s = time.time() - t0
    existing = {tuple(int(v) for v in t) for t in pyr.tiles()}
    views = []
    for pos in view_positions(0):
        clip = ClipmapRef(model, vcfg, args.lods, 8 * 64, existing)
        atlas = np.zeros((8 * 64, 512, 512), dtype=np.uint16)
        for tc, idx in clip.update(pos):
            atlas[idx] = pyr.tile(*tc)
            clip.mark_loaded(tc)
        views.append({"pos": pos, "entries": np.ascontiguousarray(clip.entries()), "atlas": atlas, "h": clip.approximate_height})
    n_pyr = len(existing)
    pyr.close()
    setup_s = time.time() - t_setup
    max_verts = int(vc.vertices_per_tile) * 6000      # buffer for the largest view (~4.8 k tiles)

    def step(i):
        v = views[i % NUM_VIEWS]
        approx = O.model_approximation(model, v["pos"], vcfg.origin_lod, v["h"])
        tiles, ind, _ = O.select_tiles(vc, approx, O.culling(v["pos"]), mesh, False)
        h = {"tree_frame": bytes(vc), "approx": bytes(approx), "mesh": bytes(mesh), "lod_count": args.lods, "min_h": 0.0, "max_h": 500.0,
             "scale": 1000.0, "T": 512, "B": 2, "pos": v["pos"], "entries": v["entries"], "atlas": v["atlas"], "tiles": tiles}
        vi = oracle_vertex_inputs(h, F.BT_FLAG_NORMALS)
        n = min(step.sample, int(ind.vertex_count), max_verts)
        out = step.buf[:n]
        t_gen = time.time()
        O.lib().bto_gen_vertices(C.byref(vi), 0, n, out.ctypes.data, threads)
        step.gen_s = time.time() - t_gen
        return n

    step.buf = np.empty(max_verts, dtype=O.VERTEX_DTYPE)
    step.buf[:] = 0                                    # touch the pages once, outside every timed region
    # bounded sample per step: calibrate on untimed steps (the second one is warm), then size the vertex sample so that the
    # whole --steps/--warmup run takes about --ref-budget-s seconds (capped at the whole view = the GPU arm's step)
    step.sample = 4_000_000
    for i in range(2):
        t0 = time.time()
        n_cal = step(i)
        cal_s = time.time() - t0
    rate, overhead = n_cal / max(step.gen_s, 1e-6), max(cal_s - step.gen_s, 0.0)
    if args.ref_sample_mverts > 0:
        step.sample = int(args.ref_sample_mverts * 1e6)
    else:
        per_step = args.ref_budget_s / max(args.steps + args.warmup, 1)
        step.sample = int(min(max((per_step - overhead) * rate, 1e6), max_verts))
    sample = step.sample
    for i in range(args.warmup):
        step(i)
    t0 = time.time()
    verts, per_step_rates = 0, []
    for i in range(args.steps):
        ts = time.time()
        n = step(i)
        per_step_rates.append(n / (time.time() - ts) / 1e6)
        verts += n
    dt = time.time() - t0
    value = verts / dt / 1e6
    out = {
        "impl": "reference",
        "metric": "Mverts/s generated (strip order) — select_tiles + gen_vertices per frame",
        "value": value, "unit": "Mverts/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": "Mverts/s", "cores": threads, "kind": "port", "per_step_Mverts_s": spread(per_step_rates),
                         "omp": {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OMP_PROC_BIND", "OMP_PLACES")}, "os_cpu_count": os.cpu_count(),
                         "sample": f"per step: select (1 view) + up to {sample} strip vertices of that view (sized from --ref-budget-s; a whole view has 36-41 M); oracle/ (C port of the reference WGSL), OpenMP {threads} threads (= the cores this process may run on; an inherited OMP_NUM_THREADS is ignored), bound close"},
        "e2e": {"value": value, "unit": "Mverts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference ships no CPU implementation of this path (SURVEY F3) and cannot be built here (no Rust/WGSL toolchain): this arm times the CPU restatement under oracle/",
        "preprocess": {"metric": "tiles/s preprocessed (oracle, CPU)", "tiles": n_pyr, "s": pre_s, "tiles_per_s": n_pyr / pre_s, "cores": threads,
                       "source": f"{args.source}^2", "lods": args.lods, "attachments": "height R16 only (the setup pyramid of this arm; the like-for-like "
                       "two-attachment 8192^2 comparison is preprocess.cpu_baseline of the GPU arm)"},
        "setup_s": setup_s,
    }
    emit(out)


_REAL_STDOUT = None


def emit(obj):
    """the ONE JSON line goes to the real stdout; everything else (NCCL banners, library chatter) went to stderr"""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


def main():
    global _REAL_STDOUT
    args = parse_args()
    # before torch / the oracle load libgomp; only a process that is alone on the box (the CPU arm, or the 1-GPU arm with its cpu_baseline) binds
    pin_openmp_env(bind=args.impl == "reference" or int(os.environ.get("WORLD_SIZE", "1")) == 1)
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)   # anything a library prints to fd 1 from here on lands on stderr
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
