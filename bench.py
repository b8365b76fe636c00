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
          and D2Hs the tile list + indirect record
preprocess (sub-object): tiles/s of the sharded config-5-shaped pyramid (height + albedo) at N GPUs
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
NUM_VIEWS = 8
VIEW_SEED = 0x5EED0010


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
    ap.add_argument("--ref-sample-mverts", type=float, default=0.0, help="reference arm: vertices per step (millions); 0 = size from --ref-budget-s")
    ap.add_argument("--ref-budget-s", type=float, default=100.0, help="reference arm: target wall time of the whole --steps/--warmup run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


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


def run_b200(args):
    import torch
    import torch.distributed as dist

    from bevy_terrain_b200 import _ffi as F
    from bevy_terrain_b200 import api

    rank, world, local = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = api.Context(local)
    flags = F.BT_FLAG_NORMALS
    cfg, pre, src_ptr, pre2 = build_terrain(api, ctx, args)
    vcfg = api.TerrainViewConfig(grid_size=64)
    vpt = 2 * 64 * 66
    views = []
    for pos in view_positions(rank):
        atlas = api.TileAtlas(ctx, cfg)
        tree = api.TileTree(atlas, vcfg, max_vertices=6400 * vpt)
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

    # ---- device-resident timed region --------------------------------------------------------------
    for i in range(args.warmup):
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
    for i in range(args.warmup):
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

    # ---- flag variants (short device-resident runs; informational) --------------------------------------
    variants = {}
    for name, vflags in (("reference_default_MORPH_BLEND_plus_NORMALS", F.BT_FLAG_MORPH | F.BT_FLAG_BLEND | F.BT_FLAG_NORMALS),
                         ("vertex_stage_only_no_normals", 0), ("MORPH_BLEND_no_normals", F.BT_FLAG_MORPH | F.BT_FLAG_BLEND)):
        for i in range(3):
            views[i % NUM_VIEWS]["tree"].gen_vertices(vflags)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nv = 0
        g0.record()
        for i in range(16):
            views[i % NUM_VIEWS]["tree"].select_tiles()
            views[i % NUM_VIEWS]["tree"].gen_vertices(vflags)
            nv += views[i % NUM_VIEWS]["verts"]
        g1.record()
        torch.cuda.synchronize()
        variants[name] = {"Mverts_per_s": nv / (g0.elapsed_time(g1) * 1e-3) / 1e6, "ms_per_step": g0.elapsed_time(g1) / 16}

    # ---- reduce over ranks (max time, sum work) ---------------------------------------------------------
    def reduce(vals, op):
        if world == 1:
            return vals
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return t.tolist()

    ms_max, e2e_ms_max = reduce([ms, e2e_ms], dist.ReduceOp.MAX if world > 1 else None)
    verts_all, e2e_verts_all = reduce([float(verts), float(e2e_verts)], dist.ReduceOp.SUM if world > 1 else None)

    gen_ms = [a.elapsed_time(b) for a, b, _ in gen_events]
    gen_verts = [n for _, _, n in gen_events]
    gen_avg_ms = float(np.mean(gen_ms))
    achieved = ALGORITHMIC_BYTES_PER_VERTEX * float(np.mean(gen_verts)) / (gen_avg_ms * 1e-3) / 1e9

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gen_vertices_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")   # dram__bytes_read + write per launch, one ncu --set full capture

    # ---- free the per-frame state, then the sharded preprocess benchmark --------------------------------
    frame_state = {"tiles_per_view": [v["tiles"] for v in views], "verts_per_view": [v["verts"] for v in views]}
    host_inputs = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        host_inputs = capture_host_inputs(ctx, views[0], cfg, vcfg)
    h2d_bytes = 1168 + views[0]["tree"].entry_count * 8   # sizeof(ViewUniforms) + tile-tree table
    for v in views:
        v["tree"].close(); v["atlas"].close()
    views.clear()
    torch.cuda.empty_cache()

    preprocess = None
    if args.preprocess_source:
        preprocess = run_preprocess_bench(api, F, ctx, args, rank, world, barrier, reduce)
        if preprocess is not None:
            preprocess["config2_single_gpu"] = pre2

    cpu_baseline = None
    if host_inputs is not None:
        cpu_baseline = cpu_vertex_baseline(host_inputs, args.cpu_sample_mverts)

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
                    "ms_per_step": e2e_ms_max / args.steps},
            "gpu_launches": 2 * args.steps,
            "roofline": {"kernel": "gen_vertices_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "avg_launch_ms": gen_avg_ms, "algorithmic_bytes_per_launch": ALGORITHMIC_BYTES_PER_VERTEX * float(np.mean(gen_verts)),
                         "write_only_GBps": 40.0 * float(np.mean(gen_verts)) / (gen_avg_ms * 1e-3) / 1e9,
                         "dram_frac": (traffic / (gen_avg_ms * 1e-3) / 1e9 / peak) if traffic else None,
                         "note": "achieved/frac use SURVEY 8d's ALGORITHMIC 80 B/vertex (40 written + 40 of bilinear-tap reads); the tap reads are "
                                 "served by L1/L2 (traffic = ncu DRAM bytes, ~40 B/vertex), so frac can exceed 1 while dram_frac = traffic/time/peak "
                                 "stays below it; the kernel is instruction-issue bound (profiles/README_r01.md)"},
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "frame": frame_state,
            "variants_rank0": variants,
            "preprocess": preprocess,
        }
        emit(out)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


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


def cpu_vertex_baseline(h, sample_mverts):
    """cpu_baseline leg: the oracle port (OpenMP over vertices) on a bounded sample of view 0, same inputs"""
    from bevy_terrain_b200 import _ffi as F
    from tests import oracle as O
    vi = oracle_vertex_inputs(h, F.BT_FLAG_NORMALS)
    n = int(min(h["verts"], sample_mverts * 1e6))
    threads = O.lib().bto_max_threads()
    out = np.empty(n, dtype=O.VERTEX_DTYPE)
    O.lib().bto_gen_vertices(C.byref(vi), 0, n, out.ctypes.data, threads)  # warm-up pass: page faults, thread pool, caches
    passes, t0 = 0, time.time()
    while True:   # about 10 s of CPU work (at least two passes)
        O.lib().bto_gen_vertices(C.byref(vi), 0, n, out.ctypes.data, threads)
        passes += 1
        dt = time.time() - t0
        if dt >= 10.0 and passes >= 2 or passes >= 64:
            break
    return {"value": passes * n / dt / 1e6, "unit": "Mverts/s", "cores": threads, "kind": "port",
            "sample": f"{passes} passes over the first {n} strip vertices of view 0 of the same workload, after one warm-up pass "
                      f"(oracle/bt_oracle.c gen_vertices, OpenMP, {threads} threads, {dt:.1f} s)"}


def run_preprocess_bench(api, F, ctx, args, rank, world, barrier, reduce):
    """config-5-shaped pyramid: S^2 height (R16) + albedo (Rgba8) -> full tile pyramid, rows sharded over ranks,
    sub-root gather + halo exchange with NCCL all_gather between the kernel phases"""
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
    # this rank's band of source rows (+1 row of bilinear footprint on each side)
    y0 = (rank * n_rows // world) * (S // n_rows)
    y1 = ((rank + 1) * n_rows // world) * (S // n_rows)
    r0, r1 = max(y0 - 1, 0), min(y1 + 1, S)
    srcs = []
    for fmt, seed in ((F.BT_FORMAT_R16, 0x5EED0001), (F.BT_FORMAT_RGBA8, 0x5EED0002)):
        srcs.append(ctx.synth_source(fmt, S, r1 - r0, x0=0, y0=r0, seed=seed))
    torch.cuda.synchronize()

    def one_run():
        pre = api.Preprocessor(ctx, cfg, rank=rank, world=world)
        for a in (0, 1):
            kw = dict(row_begin=r0, row_count=r1 - r0) if world > 1 else {}
            pre.preprocess_tile(api.PreprocessDataset(attachment_index=a, device_pixels=srcs[a], width=S, height=S, lod_range=range(0, L), **kw))
        pre.reserve()   # device store allocated outside the timed region (one-time cost, like the atlas itself)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if world == 1:
            pre.run()
        else:
            for a in (0, 1):   # jobs run in order: one attachment after the other
                pre.run_split_downsample()
                for nbytes, pack, unpack in ((pre.roots_bytes(), pre.roots_pack, pre.roots_unpack_and_finish),
                                             (pre.halo_bytes(), pre.halo_pack, pre.halo_unpack)):
                    if nbytes:
                        send = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
                        recv = torch.empty(nbytes * world, dtype=torch.uint8, device="cuda")
                        pack(send.data_ptr())
                        dist.all_gather_into_tensor(recv, send)
                        unpack(recv.data_ptr())
                    else:
                        unpack(0)
                pre.run_stitch()
        e1.record()
        barrier()
        n = len(pre.tiles(0)) + len(pre.tiles(1))
        pre.close()
        return e0.elapsed_time(e1), n

    one_run()                       # warm-up (allocations, NCCL channels)
    ms, n_tiles = one_run()
    for p in srcs:
        ctx.free(p)
    ms_max, = reduce([ms], dist.ReduceOp.MAX if world > 1 else None)
    tiles_all, = reduce([float(n_tiles)], dist.ReduceOp.SUM if world > 1 else None)
    bytes_per_tile = {"R16": 2 * (512 * 512 + S * S / ((4 ** L - 1) / 3)), "Rgba8": 4 * (512 * 512 + S * S / ((4 ** L - 1) / 3))}
    alg = (bytes_per_tile["R16"] + bytes_per_tile["Rgba8"]) * (tiles_all / 2)
    return {"metric": "tiles/s preprocessed (split + downsample + stitch, height R16 + albedo Rgba8)", "source": f"{S}^2", "lods": L,
            "tiles": int(tiles_all), "ms": ms_max, "tiles_per_s": tiles_all / (ms_max * 1e-3), "n_gpus": world, "scaling": "strong",
            "algorithmic_GBps_per_gpu": alg / (ms_max * 1e-3) / 1e9 / world, "algorithmic_bytes_per_tile": bytes_per_tile,
            "exchange": "none" if world == 1 else "NCCL all_gather of sub-root tiles + border strips (torch.distributed)"}


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
    threads = O.lib().bto_max_threads()
    t_setup = time.time()
    model = O.terrain_model(F.BT_KIND_PLANAR, (0.0, -100.0, 0.0), 2000.0, 0.0, 500.0)
    vcfg = O.default_view_config(grid_size=64)
    vc = O.view_config_uniform(vcfg, model)
    mesh = O.mesh_transform(model)
    src = O.synth_source(F.BT_FORMAT_R16, args.source, args.source, seed=0x5EED0001)
    pyr = O.Pyramid(F.BT_KIND_PLANAR, F.BT_FORMAT_R16, 512, 2, args.lods)
    pyr.add_datasets([O.make_dataset(src, lod_range=(0, args.lods))])
    t0 = time.time()
    pyr.run()
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
    # bounded sample per step: calibrate on one untimed step, then size the vertex sample so that the whole
    # --steps/--warmup run takes about --ref-budget-s seconds (capped at the whole view = the GPU arm's step)
    step.sample = 4_000_000
    t0 = time.time()
    n_cal = step(0)
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
    verts = 0
    for i in range(args.steps):
        verts += step(i)
    dt = time.time() - t0
    value = verts / dt / 1e6
    out = {
        "impl": "reference",
        "metric": "Mverts/s generated (strip order) — select_tiles + gen_vertices per frame",
        "value": value, "unit": "Mverts/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": "Mverts/s", "cores": threads, "kind": "port",
                         "sample": f"per step: select (1 view) + up to {sample} strip vertices of that view (sized from --ref-budget-s; a whole view has 36-41 M); oracle/ (C port of the reference WGSL), OpenMP {threads} threads"},
        "e2e": {"value": value, "unit": "Mverts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference ships no CPU implementation of this path (SURVEY F3) and cannot be built here (no Rust/WGSL toolchain): this arm times the CPU restatement under oracle/",
        "preprocess": {"metric": "tiles/s preprocessed (oracle, CPU)", "tiles": n_pyr, "s": pre_s, "tiles_per_s": n_pyr / pre_s, "cores": threads,
                       "source": f"{args.source}^2", "lods": args.lods, "attachments": "height R16 only"},
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
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)   # anything a library prints to fd 1 from here on lands on stderr
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
