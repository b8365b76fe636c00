"""Times the config-5 preprocess (65536^2 -> 8 LODs of 512^2 tiles) per attachment on one GPU and prints the device content hash,
so that kernel variants (BT_SPLIT_LEGACY=1, BT_SPLIT_STAGES, BT_SPLIT_BANDS: read once per process) can be compared inside one
gpurun call:  python tools/pre_probe.py [--source 65536] [--lods 8] [--tag name]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--source", type=int, default=65536)
    ap.add_argument("--lods", type=int, default=8)
    ap.add_argument("--tag", default="")
    ap.add_argument("--repeats", type=int, default=3)
    args = ap.parse_args()
    import torch
    from bevy_terrain_b200 import _ffi as F, api
    ctx = api.Context(0)
    S, L = args.source, args.lods
    model = api.TerrainModel.planar((0.0, -100.0, 0.0), 2000.0, 0.0, 500.0)
    out = {"tag": args.tag, "source": S, "lods": L, "env": {k: v for k, v in os.environ.items() if k.startswith("BT_SPLIT")}}
    for name, fmt, afmt, seed in (("height", F.BT_FORMAT_R16, api.AttachmentFormat.R16, 0x5EED0001), ("albedo", F.BT_FORMAT_RGBA8, api.AttachmentFormat.Rgba8, 0x5EED0002)):
        cfg = api.TerrainConfig(lod_count=L, model=model)
        cfg.add_attachment(api.AttachmentConfig(name=name, texture_size=512, border_size=2, format=afmt))
        src = ctx.synth_source(fmt, S, S, seed=seed)
        torch.cuda.synchronize()
        times, h = [], None
        for rep in range(args.repeats + 1):
            pre = api.Preprocessor(ctx, cfg)
            pre.preprocess_tile(api.PreprocessDataset(attachment_index=0, device_pixels=src, width=S, height=S, lod_range=range(0, L)))
            pre.reserve()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pre.run()
            e1.record()
            torch.cuda.synchronize()
            if rep or args.repeats == 0:
                times.append(e0.elapsed_time(e1))
            h = pre.content_hash(0)
            pre.close()
        ctx.free(src)
        out[name] = {"ms_median": float(np.median(times)), "ms_min": float(min(times)), "hash": f"{h:#018x}"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
