"""BASELINE.json configs[4] at full size on one GPU: preprocess of a 65536^2 synthetic height (R16) + albedo (Rgba8) source into
the full 8-LOD tile pyramid (2 x 21 845 tiles of 512^2).  The sharded (NCCL) form of the same job is compared with this
single-GPU form and with the oracle at reduced size in tests/test_gpu_multi_rank.py / tests/test_multi_rank_gloo.py; here the
size-independent properties are checked at full size, plus an exact comparison with the oracle on a sub-tree."""
import numpy as np
import pytest

from bevy_terrain_b200 import _ffi as F
from tests import oracle as O

pytestmark = pytest.mark.gpu
T, B, L, S = 512, 2, 8, 65536
Cn = T - 2 * B


@pytest.mark.parametrize("fmt,afmt,seed", [(F.BT_FORMAT_R16, "R16", 0x5EED0001), (F.BT_FORMAT_RGBA8, "Rgba8", 0x5EED0002)])
def test_config4_full_pyramid(ctx, fmt, afmt, seed):
    from bevy_terrain_b200 import api
    model = api.TerrainModel.planar((0.0, -100.0, 0.0), 2000.0, 0.0, 500.0)
    cfg = api.TerrainConfig(lod_count=L, model=model)
    cfg.add_attachment(api.AttachmentConfig(name="a", texture_size=T, border_size=B, format=getattr(api.AttachmentFormat, afmt)))
    src = ctx.synth_source(fmt, S, S, seed=seed)
    pre = api.Preprocessor(ctx, cfg)
    try:
        pre.preprocess_tile(api.PreprocessDataset(device_pixels=src, width=S, height=S, lod_range=range(0, L))).run()
        assert len(pre.tiles()) == (4 ** L - 1) // 3 == 21845
        rng = np.random.default_rng(7)
        for lod in (7, 6, 4, 1):
            n = 1 << lod
            for _ in range(4):
                x, y = int(rng.integers(0, n - 1)), int(rng.integers(0, n - 1))
                a, right, below, diag = (pre.get_tile(0, (0, lod, x + dx, y + dy)) for dx, dy in ((0, 0), (1, 0), (0, 1), (1, 1)))
                assert np.array_equal(a[B:B + Cn, B + Cn:], right[B:B + Cn, B:2 * B])      # stitched border == neighbour centre
                assert np.array_equal(a[B + Cn:, B:B + Cn], below[B:2 * B, B:B + Cn])
                assert np.array_equal(a[B + Cn:, B + Cn:], diag[B:2 * B, B:2 * B])
                assert a.reshape(T, T, -1)[..., 0].all()                                    # no holes in the synthetic source
        corner = pre.get_tile(0, (0, 7, 0, 0))      # terrain corner: the border repeats the own centre (no neighbour there)
        assert np.array_equal(corner[:B, B:B + Cn], np.repeat(corner[B:B + 1, B:B + Cn], B, axis=0))
        # a sub-tree through BOTH implementations: the source window of lod-5 tile (9, 20) -> that tile and its 20 descendants
        w = S // 32
        crop = O.synth_source(fmt, w + 1, w + 1, x0=9 * w, y0=20 * w, seed=seed)      # the same noise field, generated on the host
        tl, br = (9 / 32, 20 / 32), (10 / 32, 21 / 32)
        pre2 = api.Preprocessor(ctx, cfg)
        pre2.preprocess_tile(api.PreprocessDataset(pixels=crop, lod_range=range(5, 8), top_left=tl, bottom_right=br)).run()
        ref = O.Pyramid(F.BT_KIND_PLANAR, fmt, T, B, L)
        ref.add_datasets([O.make_dataset(crop, top_left=tl, bottom_right=br, lod_range=(5, 8))])
        ref.run()
        got = {tuple(int(v) for v in t) for t in pre2.tiles()}
        exp = {tuple(int(v) for v in t) for t in ref.tiles()}
        assert got == exp and len(exp) == 21
        for tc in sorted(exp):
            assert np.array_equal(pre2.get_tile(0, tc), ref.tile(*tc)), tc
        pre2.close()
    finally:
        pre.close()
        ctx.free(src)
