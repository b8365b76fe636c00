"""SECOND, independent restatement of the reference's preprocess and select shaders — TEST INFRASTRUCTURE.

Written from the WGSL alone (src/shaders/preprocess/{preprocessing,split,downsample,stitch}.wgsl,
src/shaders/tiling_prepass/{refine_tiles,prepare_prepass}.wgsl, src/shaders/functions.wgsl), in numpy, WITHOUT reading
oracle/*.c: whole tiles at a time, float32 array arithmetic (every numpy f32 operation rounds once, which is the numerics
contract of DESIGN.md section 2).  tests/test_independent_twin.py diffs it against the C oracle: a misreading of the shaders would
have to be made twice, in two different formulations, to go unnoticed.

Conventions fixed by the WGSL spec / DESIGN.md section 2 (the only things taken from the project's documents rather than from the
shaders): textureSampleLevel(level 0) = bilinear with texel centres at i + 0.5, exact f32 weights, clamp-to-edge, combined in
split as  mix(mix(t00, t10, fx), mix(t01, t11, fx), fy)  with WGSL's mix(a, b, t) = a * (1 - t) + b * t; textureGather = the
same four texels; pack*unorm(v) = floor(0.5 + M * clamp(v, 0, 1)); 0 / 0 in downsample := 0 (quirk Q3).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


# ------------------------------------------------------------------------------------------------------------------------
# texel <-> float
# ------------------------------------------------------------------------------------------------------------------------
def decode(tile: np.ndarray) -> np.ndarray:
    """texture_2d_array<f32> view of a tile: (T, T) uint16 -> (T, T, 4) with g = b = 0, a = 1; (T, T, 4) uint8 -> /255"""
    if tile.dtype == np.uint16:
        out = np.zeros(tile.shape + (4,), dtype=f32)
        out[..., 0] = tile.astype(f32) / f32(65535.0)
        out[..., 3] = f32(1.0)
        return out
    return tile.astype(f32) / f32(255.0)


def pack(values: np.ndarray, like: np.dtype) -> np.ndarray:
    """process_entry: pack2x16unorm(.x) for R16, pack4x8unorm for Rgba8 (preprocessing.wgsl:73-90)"""
    m = f32(65535.0) if like == np.uint16 else f32(255.0)
    with np.errstate(invalid="ignore"):
        c = np.minimum(np.maximum(values, f32(0.0)), f32(1.0))      # clamp(v, 0, 1); NaN handled by the caller (Q3)
        q = np.floor(f32(0.5) + m * c)
    if like == np.uint16:
        return q[..., 0].astype(np.uint16)
    return q.astype(np.uint8)


def is_border_mask(T: int, B: int) -> np.ndarray:
    """is_border(coords) for every pixel: !inside(coords, (B, B, C, C))"""
    C = T - 2 * B
    y, x = np.mgrid[0:T, 0:T]
    return ~((x >= B) & (x < B + C) & (y >= B) & (y < B + C))


# ------------------------------------------------------------------------------------------------------------------------
# split.wgsl:18-43
# ------------------------------------------------------------------------------------------------------------------------
def split_tile(source: np.ndarray, tile_xy, lod: int, T: int, B: int, top_left=(0.0, 0.0), bottom_right=(1.0, 1.0),
               existing: np.ndarray | None = None) -> np.ndarray:
    """one output tile of the split pass; `existing` = the atlas texels already there (an earlier dataset), zeros if None"""
    C = T - 2 * B
    like = source.dtype
    src = decode(source)
    H, W = source.shape[0], source.shape[1]
    coords = np.arange(T, dtype=np.int64)
    # tile_coords = vec2<f32>(coords - border) / f32(center_size): u32 arithmetic wraps for coords < border, but those are border pixels
    tc = (coords - B).astype(f32) / f32(C)
    scale = f32(1 << lod)
    out_val = np.zeros((T, T, 4), dtype=f32)
    sx = (f32(tile_xy[0]) + tc) / scale
    sy = (f32(tile_xy[1]) + tc) / scale
    sx = (sx - f32(top_left[0])) / (f32(bottom_right[0]) - f32(top_left[0]))       # inverse_mix
    sy = (sy - f32(top_left[1])) / (f32(bottom_right[1]) - f32(top_left[1]))
    px = sx * f32(W) - f32(0.5)
    py = sy * f32(H) - f32(0.5)
    fx, fy = np.floor(px), np.floor(py)
    rx, ry = (px - fx).astype(f32), (py - fy).astype(f32)
    ix, iy = fx.astype(np.int64), fy.astype(np.int64)
    x0, x1 = np.clip(ix, 0, W - 1), np.clip(ix + 1, 0, W - 1)
    y0, y1 = np.clip(iy, 0, H - 1), np.clip(iy + 1, 0, H - 1)
    t00 = src[y0[:, None], x0[None, :]]
    t01 = src[y1[:, None], x0[None, :]]     # (x, y + 1)
    t10 = src[y0[:, None], x1[None, :]]
    t11 = src[y1[:, None], x1[None, :]]
    wy = ry[:, None, None]
    wx = rx[None, :, None]
    # the bilinear convention of the PREPROCESS path: WGSL mix(a, b, t) = a * (1 - t) + b * t, x first, then y
    # (the vertex path uses the order of the reference's CPU sampler instead; both are inside the latitude WGSL leaves to the sampler)
    one = f32(1.0)
    top = t00 * (one - wx) + t10 * wx
    bottom = t01 * (one - wx) + t11 * wx
    value = top * (one - wy) + bottom * wy
    valid = (t00[..., 0] != 0) & (t01[..., 0] != 0) & (t10[..., 0] != 0) & (t11[..., 0] != 0)    # all(gather(0) != 0)
    inside_1d = (tc >= f32(0.0)) & (tc <= f32(1.0))                                               # step(0, p) * step(p, 1)
    inside = inside_1d[:, None] & inside_1d[None, :]
    keep = decode(existing) if existing is not None else np.zeros((T, T, 4), dtype=f32)
    out_val = np.where((valid & inside)[..., None], value, keep)
    out_val[is_border_mask(T, B)] = f32(0.0)
    return pack(out_val, like)


# ------------------------------------------------------------------------------------------------------------------------
# downsample.wgsl:12-40
# ------------------------------------------------------------------------------------------------------------------------
def downsample_tile(children, T: int, B: int, like) -> np.ndarray:
    """children[i] for child index i = x + 2 * y (None = the child does not exist: its layer reads as zeros)"""
    C = T - 2 * B
    child_size = C // 2
    dec = [decode(c) if c is not None else np.zeros((T, T, 4), dtype=f32) for c in children]
    coords = np.arange(T, dtype=np.int64)
    tcd = coords - B
    cc = 2 * (tcd % child_size) + B
    which = tcd // child_size
    value = np.zeros((T, T, 4), dtype=f32)
    count = np.zeros((T, T), dtype=f32)
    centre = ~is_border_mask(T, B)
    Y, X = np.mgrid[0:T, 0:T]
    for ox, oy in ((0, 0), (0, 1), (1, 0), (1, 1)):
        cx = np.clip(cc[X] + ox, 0, T - 1)
        cy = np.clip(cc[Y] + oy, 0, T - 1)
        child_index = np.clip(which[X] + 2 * which[Y], 0, 3)
        texel = np.zeros((T, T, 4), dtype=f32)
        for i in range(4):
            sel = (child_index == i) & centre
            texel[sel] = dec[i][cy[sel], cx[sel]]
        in_range = (cc[X] + ox < T) & (cc[Y] + oy < T)           # an out-of-range textureLoad returns zeros
        texel[~in_range] = f32(0.0)
        ok = (texel[..., 0] != 0) | (texel[..., 1] != 0) | (texel[..., 2] != 0)
        value = np.where(ok[..., None], value + texel, value)
        count = np.where(ok, count + f32(1.0), count)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = value / count[..., None]
    mean[count == 0] = f32(0.0)                                   # 0 / 0 := 0 (Q3)
    mean[~centre] = f32(0.0)
    return pack(mean, like)


# ------------------------------------------------------------------------------------------------------------------------
# stitch.wgsl
# ------------------------------------------------------------------------------------------------------------------------
_EVEN = [("PS", "PT"), ("PS", "PT"), ("NT", "PS"), ("NT", "NS"), ("PT", "NS"), ("PS", "PT")]
_ODD = [("PS", "PT"), ("PS", "PT"), ("PT", "NS"), ("PT", "PS"), ("NT", "PS"), ("PS", "PT")]


def _project(x, y, T, own_side, other_side):
    index = (6 + other_side - own_side) % 6
    info = _EVEN[index] if own_side % 2 == 0 else _ODD[index]
    pick = {"PS": x, "PT": y, "NS": T - 1 - x, "NT": T - 1 - y}
    return pick[info[0]], pick[info[1]]


def stitch_tile(tile: np.ndarray, own_side: int, neighbours, T: int, B: int) -> np.ndarray:
    """neighbours: 8 entries (texels, side) or None, in TileCoordinate::neighbours order: top, right, bottom, left, TL, TR, BR, BL"""
    C = T - 2 * B
    O = B + C
    out = tile.copy()
    bounds = [(B, 0, C, B), (O, B, B, C), (B, O, C, B), (0, B, B, C), (0, 0, B, B), (O, 0, B, B), (O, O, B, B), (0, O, B, B)]
    offsets = [(0, C), (-C, 0), (0, -C), (C, 0), (C, C), (-C, C), (-C, -C), (C, -C)]
    for k, (bx, by, bw, bh) in enumerate(bounds):
        ys, xs = np.mgrid[by:by + bh, bx:bx + bw]
        if neighbours[k] is None:
            out[ys, xs] = tile[np.clip(ys, B, B + C - 1), np.clip(xs, B, B + C - 1)]          # repeat_data
        else:
            texels, side = neighbours[k]
            nx, ny = _project(xs + offsets[k][0], ys + offsets[k][1], T, own_side, side)
            ok = (nx >= 0) & (nx < T) & (ny >= 0) & (ny < T)
            vals = texels[np.clip(ny, 0, T - 1), np.clip(nx, 0, T - 1)]
            vals = np.where(ok[..., None], vals, 0) if vals.ndim == 3 else np.where(ok, vals, 0)
            out[ys, xs] = vals
    return out


# ------------------------------------------------------------------------------------------------------------------------
# tiling prepass: prepare_root / refine_tiles / prepare_next / prepare_render, executed sequentially (invocations in id order)
# ------------------------------------------------------------------------------------------------------------------------
C_SQR = f32(0.87 * 0.87)
ONE_PLUS_C_SQR = f32(1.0 + 0.87 * 0.87)


def _change_lod(lod, xy, uv, new_lod):
    d = int(new_lod) - int(lod)
    if d == 0:
        return lod, xy, uv
    count = 1 << abs(d)
    size = np.ldexp(f32(1.0), d)                   # pow(2.0, f32(d)), exact
    if d > 0:
        s = (uv[0] * size, uv[1] * size)
        xy = (xy[0] * count + int(s[0]), xy[1] * count + int(s[1]))
        uv = (f32(s[0] - np.trunc(s[0])), f32(s[1] - np.trunc(s[1])))
    else:
        uv = (f32((f32(xy[0] % count) + uv[0]) * size), f32((f32(xy[1] % count) + uv[1]) * size))
        xy = (xy[0] // count, xy[1] // count)
    return new_lod, xy, uv


def _normalize(v):
    n = np.sqrt(f32(f32(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]))
    return (f32(v[0] / n), f32(v[1] / n), f32(v[2] / n))


def _local_position(spherical, side, lod, xy, uv):
    tcnt = f32(1 << lod)
    u = f32((f32(xy[0]) + uv[0]) / tcnt)
    v = f32((f32(xy[1]) + uv[1]) / tcnt)
    if not spherical:
        return (f32(u - f32(0.5)), f32(0.0), f32(v - f32(0.5)))
    u = f32((u - f32(0.5)) / f32(0.5))
    v = f32((v - f32(0.5)) / f32(0.5))
    u = f32(u / np.sqrt(f32(ONE_PLUS_C_SQR - f32(f32(C_SQR * u) * u))))
    v = f32(v / np.sqrt(f32(ONE_PLUS_C_SQR - f32(f32(C_SQR * v) * v))))
    one = f32(1.0)
    p = [(-one, -v, u), (u, -v, one), (u, one, v), (one, -u, v), (v, -u, -one), (v, -one, u)][side]
    return _normalize(p)


def select_tiles(view_config, approximation, view_world_position, mesh, spherical: bool):
    """returns the final tile list [(side, lod, x, y)] in the order a sequential execution of the shaders produces, and the
    number of parents dropped after the last pass.  `mesh`: world_from_local[12] + normal_matrix[9] (column-major) as in
    bt_mesh_transform; the other arguments are the ctypes structs of bevy_terrain_b200._ffi."""
    a = [f32(x) for x in mesh.world_from_local]
    nm = [f32(x) for x in mesh.normal_matrix]
    h = f32(approximation.approximate_height)
    e = [f32(x) for x in view_world_position]
    sub = f32(view_config.subdivision_distance)

    def divided(side, lod, x, y):
        p = approximation.sides[side]
        vl, vxy, vuv = _change_lod(approximation.origin_lod, (int(np.uint32(p.view_xy[0])), int(np.uint32(p.view_xy[1]))),
                                   (f32(p.view_uv[0]), f32(p.view_uv[1])), lod)
        # vec2<i32>(view.xy) - vec2<i32>(tile.xy)
        ox = int(np.int32(np.uint32(vxy[0] & 0xFFFFFFFF))) - x
        oy = int(np.int32(np.uint32(vxy[1] & 0xFFFFFFFF))) - y
        u = f32(0.0) if ox < 0 else (f32(1.0) if ox > 0 else vuv[0])
        v = f32(0.0) if oy < 0 else (f32(1.0) if oy > 0 else vuv[1])
        lp = _local_position(spherical, side, lod, (x, y), (u, v))
        wp = [f32(f32(f32(f32(a[k] * lp[0]) + f32(a[3 + k] * lp[1])) + f32(a[6 + k] * lp[2])) + a[9 + k]) for k in range(3)]
        n_local = lp if spherical else (f32(0.0), f32(1.0), f32(0.0))
        wn = _normalize([f32(f32(f32(nm[k] * n_local[0]) + f32(nm[3 + k] * n_local[1])) + f32(nm[6 + k] * n_local[2])) for k in range(3)])
        q = [f32(f32(wp[k] + f32(h * wn[k])) - e[k]) for k in range(3)]
        dist = np.sqrt(f32(f32(q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]))
        return bool(dist < f32(sub / f32(1 << lod)))

    frontier = [(s, 0, 0, 0) for s in range(6 if spherical else 1)]       # prepare_root
    final, dropped = [], 0
    for refinement in range(view_config.refinement_count + 1):
        nxt = []
        for side, lod, x, y in frontier:                                   # refine_tiles, invocations in id order
            if divided(side, lod, x, y):
                for i in range(4):                                         # subdivide
                    nxt.append((side, lod + 1, (x << 1) + (i & 1), (y << 1) + ((i >> 1) & 1)))
            else:
                final.append((side, lod, x, y))
        if refinement == view_config.refinement_count:
            dropped = len(nxt) // 4                                        # children of the last pass are never read (quirk Q2)
        frontier = nxt
        if not frontier:
            break
    return final, dropped


# ------------------------------------------------------------------------------------------------------------------------
# TileCoordinate::neighbours (src/math/coordinate.rs:9-53,208-279) and the task order of Preprocessor (preprocessor.rs:234-343)
# ------------------------------------------------------------------------------------------------------------------------
_NEIGHBOURING_SIDES = [[0, 4, 2, 1, 5], [1, 0, 2, 3, 5], [2, 0, 4, 3, 1], [3, 2, 4, 5, 1], [4, 2, 0, 5, 3], [5, 4, 0, 1, 3]]
_SI_EVEN = [("S", "T"), ("0", "T"), ("0", "S"), ("T", "S"), ("T", "0"), ("S", "0")]
_SI_ODD = [("S", "T"), ("S", "1"), ("T", "1"), ("T", "S"), ("1", "S"), ("1", "T")]
_OFFSETS = [(0, -1), (1, 0), (0, 1), (-1, 0), (-1, -1), (1, -1), (1, 1), (-1, 1)]


def neighbours(side: int, lod: int, x: int, y: int, spherical: bool):
    """the eight neighbour coordinates (side, lod, x, y) or None (INVALID), in the reference's order"""
    n = 1 << lod
    out = []
    for dx, dy in _OFFSETS:
        nx, ny = x + dx, y + dy
        if not spherical:
            out.append((side, lod, nx, ny) if 0 <= nx < n and 0 <= ny < n else None)
            continue
        xo, yo = nx < 0 or nx >= n, ny < 0 or ny >= n
        if xo and yo:
            out.append(None)
            continue
        edge = 1 if nx < 0 else 2 if ny < 0 else 3 if nx >= n else 4 if ny >= n else 0
        px, py = min(max(nx, 0), n - 1), min(max(ny, 0), n - 1)
        other = _NEIGHBOURING_SIDES[side][edge]
        info = (_SI_EVEN if side % 2 == 0 else _SI_ODD)[(6 + other - side) % 6]
        pick = {"0": 0, "1": n - 1, "S": px, "T": py}
        out.append((other, lod, pick[info[0]], pick[info[1]]))
    return out


def build_pyramid(sources: dict, T: int, B: int, lod_count: int, spherical: bool) -> dict:
    """full-coverage datasets {side: source image} -> {(side, lod, x, y): texels}: split the deepest LOD, downsample every
    coarser one from its four children, then stitch every tile from its neighbours' (pre-stitch == post-stitch) centres"""
    tiles = {}
    deepest = lod_count - 1
    for side, src in sources.items():
        n = 1 << deepest
        for x in range(n):
            for y in range(n):
                tiles[(side, deepest, x, y)] = split_tile(src, (x, y), deepest, T, B)
    like = next(iter(sources.values())).dtype
    for lod in range(deepest - 1, -1, -1):
        for side in sources:
            for x in range(1 << lod):
                for y in range(1 << lod):
                    kids = [tiles.get((side, lod + 1, 2 * x + i % 2, 2 * y + i // 2)) for i in range(4)]
                    tiles[(side, lod, x, y)] = downsample_tile(kids, T, B, like)
    stitched = {}
    for (side, lod, x, y), texels in tiles.items():
        nb = []
        for c in neighbours(side, lod, x, y, spherical):
            nb.append((tiles[c], c[0]) if c is not None and c in tiles else None)
        stitched[(side, lod, x, y)] = stitch_tile(texels, side, nb, T, B)
    return stitched
