"""Pins the oracle's numerics primitives (WGSL semantics table) against independent
numpy formulations.  The reference has no tests: PARITY UNPINNED by reference vectors."""
import ctypes as C

import numpy as np

from tests import oracle as O


def test_log2_within_2ulp_of_libm():
    L = O.lib()
    rng = np.random.default_rng(1)
    xs = np.concatenate([np.float32(2.0) ** np.arange(-40, 40, dtype=np.float32),
                         rng.uniform(1e-6, 1e6, 4000).astype(np.float32),
                         np.exp(rng.uniform(-30, 30, 4000)).astype(np.float32),
                         (1.0 + rng.uniform(-1e-3, 1e-3, 2000)).astype(np.float32)])
    got = np.array([L.bto_log2f(float(x)) for x in xs], dtype=np.float32)
    ref = np.log2(xs.astype(np.float64))
    ulp = np.spacing(np.maximum(np.abs(ref), 1e-30).astype(np.float32)).astype(np.float64)
    err = np.abs(got.astype(np.float64) - ref) / ulp
    assert err.max() <= 2.0, err.max()
    # exact powers of two are exact (the blend/morph LOD decisions sit on these)
    for k in range(-20, 20):
        assert L.bto_log2f(float(np.float32(2.0) ** k)) == float(k)


def test_f32_to_u32_saturates_like_wgsl():
    L = O.lib()
    assert L.bto_f32_to_u32(-1.5) == 0
    assert L.bto_f32_to_u32(float("nan")) == 0
    assert L.bto_f32_to_u32(3.99) == 3
    assert L.bto_f32_to_u32(5e9) == 0xFFFFFFFF


def _change(lod, xy, uv, new):
    L = O.lib()
    l = C.c_uint32(lod)
    x = (C.c_uint32 * 2)(*xy)
    u = (C.c_float * 2)(*uv)
    L.bto_coordinate_change_lod(C.byref(l), x, u, new)
    return l.value, (x[0], x[1]), (u[0], u[1])


def test_coordinate_change_lod_matches_f64_position():
    rng = np.random.default_rng(2)
    for _ in range(500):
        lod = int(rng.integers(0, 12))
        new = int(rng.integers(0, 12))
        xy = rng.integers(0, 1 << lod, 2)
        uv = rng.uniform(0, 1, 2).astype(np.float32)
        l2, xy2, uv2 = _change(lod, xy, uv, new)
        assert l2 == new
        for k in range(2):
            before = (float(xy[k]) + float(uv[k])) / (1 << lod)
            after = (float(xy2[k]) + float(uv2[k])) / (1 << new)
            assert abs(before - after) < 2e-6
            assert 0.0 <= uv2[k] <= 1.0


def test_change_lod_uv_one_rolls_into_next_tile():
    """quirk Q5: uv == 1.0 rolls into tile x+1 when refining"""
    l, xy, uv = _change(2, (1, 1), (1.0, 0.5), 4)
    assert xy == (1 * 4 + 4, 1 * 4 + 2) and uv == (0.0, 0.0)


def test_local_position_cube_sphere_conventions():
    """uv -> sphere mapping of functions.wgsl:73-96 vs an f64 numpy restatement of coordinate.rs:110-130"""
    L = O.lib()
    csq = 0.87 * 0.87
    rng = np.random.default_rng(3)
    for side in range(6):
        for _ in range(50):
            lod = int(rng.integers(0, 8))
            xy = rng.integers(0, 1 << lod, 2)
            uv = rng.uniform(0, 1, 2).astype(np.float32)
            out = (C.c_float * 3)()
            L.bto_local_position(1, side, lod, (C.c_uint32 * 2)(*xy), (C.c_float * 2)(*uv), out)
            st = (xy + uv.astype(np.float64)) / (1 << lod)
            w = (st - 0.5) / 0.5
            q = w / np.sqrt(1 + csq - csq * w * w)
            p = {0: (-1, -q[1], q[0]), 1: (q[0], -q[1], 1), 2: (q[0], 1, q[1]), 3: (1, -q[0], q[1]), 4: (q[1], -q[0], -1), 5: (q[1], -1, q[0])}[side]
            p = np.array(p) / np.linalg.norm(p)
            assert np.allclose(np.array(out[:]), p, atol=3e-7)
    # planar: (u - 1/2, 0, v - 1/2)
    out = (C.c_float * 3)()
    L.bto_local_position(0, 0, 1, (C.c_uint32 * 2)(1, 0), (C.c_float * 2)(0.5, 0.25), out)
    assert tuple(out[:]) == (0.25, 0.0, -0.375)


def test_sampler_twin_matches_numpy_bilinear():
    """AttachmentData::sample (terrain_data/mod.rs:221-264) restated in numpy f32"""
    L = O.lib()
    rng = np.random.default_rng(4)
    T = 32
    tile = rng.integers(0, 65536, (T, T)).astype(np.uint16)
    for _ in range(300):
        u, v = rng.uniform(0.02, 0.98, 2).astype(np.float32)
        got = L.bto_sample_r16(tile.ctypes.data, T, float(u), float(v))
        px, py = np.float32(u * np.float32(T)) - np.float32(0.5), np.float32(v * np.float32(T)) - np.float32(0.5)
        ix, iy = int(px), int(py)
        rx, ry = np.float32(px - np.float32(ix)), np.float32(py - np.float32(iy))
        val = lambda x, y: np.float32(tile[y, x]) / np.float32(65535)
        a = val(ix, iy) + (val(ix, iy + 1) - val(ix, iy)) * ry
        b = val(ix + 1, iy) + (val(ix + 1, iy + 1) - val(ix + 1, iy)) * ry
        exp = np.float32(a + np.float32(b - a) * rx)
        assert got == float(exp)
