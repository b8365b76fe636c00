"""Times bt_select_tiles alone on the views of configs[1] (device events around 200 back-to-back launches per view).
   BT_LIB_PATH=... python tools/select_probe.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import bench
    from bevy_terrain_b200 import _ffi as F, api
    sys.argv = sys.argv[:1]
    args = bench.parse_args()
    ctx = api.Context(0)
    sys.argv = sys.argv[:1]
    cfg, pre, src, _ = bench.build_terrain(api, ctx, args)
    vcfg = api.TerrainViewConfig(grid_size=64)
    views = bench.frame_views(api, ctx, cfg, vcfg, pre, bench.view_positions(0), 8192)
    out = {"lib": os.environ.get("BT_LIB_PATH", "default"), "us": [], "tiles": []}
    for v in views:
        t = v["tree"]
        for _ in range(5):
            t.select_tiles()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            t.select_tiles()
        e1.record()
        torch.cuda.synchronize()
        out["us"].append(round(e0.elapsed_time(e1) * 1000 / 200, 2))
        tiles, _ = t.readback()
        out["tiles"].append(len(tiles))
    out["mean_us"] = round(sum(out["us"]) / len(out["us"]), 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
