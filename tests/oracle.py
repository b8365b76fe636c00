"""ctypes wrapper around oracle/libbt_oracle.so — the CPU restatement of the reference algorithm.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  Struct *layouts* are shared with the product
binding (bevy_terrain_b200._ffi) because both sides speak the reference's POD formats.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from bevy_terrain_b200 import _ffi as F

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(_ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libbt_oracle.so")

u32, i32, f32, f64, u64 = C.c_uint32, C.c_int32, C.c_float, C.c_double, C.c_uint64


class SelectResult(C.Structure):
    _fields_ = [("final_count", u32), ("dropped_parents", u32), ("max_frontier", u32), ("overflow", u32)]


class VertexInputs(C.Structure):
    _fields_ = [("view_config", F.ViewConfigUniform), ("approximation", F.ModelApproximation),
                ("terrain_config", F.TerrainConfigUniform), ("attachments", F.AttachmentUniform * 8),
                ("mesh", F.MeshTransform), ("view_world_position", f32 * 3), ("spherical", i32),
                ("texture_size", u32), ("atlas_size", u32), ("tile_tree", C.POINTER(F.TileTreeEntry)),
                ("height_atlas", C.c_void_p), ("tiles", C.POINTER(F.Tile)), ("tile_count", u32), ("flags", u32)]


class FilterInputs(C.Structure):   # bto_filter_inputs: the configs[2] sampler / albedo extras
    _fields_ = [("height_mips", C.c_void_p), ("albedo_atlas", C.c_void_p), ("albedo_mips", C.c_void_p),
                ("mip_level_count", u32), ("albedo_attachment", u32)]


_lib = None


def build() -> None:
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


def _stale() -> bool:
    """the built checker is older than its sources (an edit without `make -C oracle`): rebuild instead of testing stale code"""
    try:
        d = os.path.dirname(ORACLE_LIB)
        built = os.path.getmtime(ORACLE_LIB)
        return any(os.path.getmtime(os.path.join(d, f)) > built for f in os.listdir(d) if f.endswith((".c", ".h")))
    except OSError:
        return False


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(ORACLE_LIB) or _stale():
        build()
    L = C.CDLL(ORACLE_LIB)
    P, vp = C.POINTER, C.c_void_p
    sigs = {
        "bto_log2f": (f32, [f32]),
        "bto_f32_to_u32": (u32, [f32]),
        "bto_coordinate_change_lod": (None, [P(u32), P(u32), P(f32), u32]),
        "bto_local_position": (None, [i32, u32, u32, P(u32), P(f32), P(f32)]),
        "bto_select_tiles": (None, [P(F.ViewConfigUniform), P(F.ModelApproximation), P(F.CullingUniform),
                                    P(F.MeshTransform), i32, P(F.Tile), u32, P(F.Indirect), P(SelectResult)]),
        "bto_select_tiles_ex": (None, [P(F.ViewConfigUniform), P(F.ModelApproximation), P(F.CullingUniform), P(F.MeshTransform), i32, i32,
                                       f32, f32, P(F.Tile), u32, P(F.Indirect), P(SelectResult)]),
        "bto_gen_vertices": (None, [P(VertexInputs), u64, u64, vp, i32]),
        "bto_tile_minmax": (None, [vp, u32, u32, vp]),
        "bto_select_tiles_minmax": (None, [P(F.ViewConfigUniform), P(F.ModelApproximation), P(F.CullingUniform), P(F.MeshTransform), i32, i32,
                                           f32, f32, vp, vp, u32, u32, P(F.Tile), u32, P(F.Indirect), P(SelectResult)]),
        "bto_pyramid_create": (vp, [u32, u32, u32, u32, u32]),
        "bto_pyramid_destroy": (None, [vp]),
        "bto_pyramid_add_dataset": (None, [vp, P(F.Dataset)]),
        "bto_pyramid_add_datasets": (None, [vp, P(F.Dataset), u32]),
        "bto_pyramid_run": (None, [vp, i32]),
        "bto_pyramid_run_phases": (None, [vp, i32, i32]),
        "bto_stitch_tile_raw": (None, [u32, u32, u32, P(F.Tile), vp, P(F.Tile), P(vp)]),
        "bto_downsample_tile_raw": (None, [u32, u32, u32, vp, P(vp)]),
        "bto_pyramid_tile_count": (u32, [vp]),
        "bto_pyramid_list_tiles": (u32, [vp, P(F.Tile), u32]),
        "bto_pyramid_tile": (vp, [vp, P(F.Tile)]),
        "bto_tile_neighbours": (None, [P(F.Tile), i32, P(F.Tile)]),
        "bto_generate_mipmaps": (None, [u32, u32, u32, vp, vp]),
        "bto_gen_vertices_filtered": (None, [P(VertexInputs), P(FilterInputs), u64, u64, vp, vp, i32]),
        "bto_model_approximation": (None, [P(F.TerrainModel), P(f64), u32, f32, P(F.ModelApproximation)]),
        "bto_coordinate_from_world_position": (None, [P(F.TerrainModel), P(f64), P(u32), P(f64)]),
        "bto_project_point_ellipsoid": (None, [f64, f64, P(f64), P(f64)]),
        "bto_coordinate_world_position": (None, [P(F.TerrainModel), u32, P(f64), f32, P(f64)]),
        "bto_tile_tree_origin": (None, [u32, u32, P(f64), P(u32)]),
        "bto_project_to_side": (None, [P(F.TerrainModel), u32, P(f64), u32, P(f64)]),
        "bto_tile_distance": (f64, [P(F.TerrainModel), P(F.Tile), u32, P(f64), f32, P(f64)]),
        "bto_sample_r16": (f32, [vp, u32, f32, f32]),
        "bto_synth_source": (None, [u32, u32, u32, u32, u32, u32, u32, vp, i32]),
        "bto_max_threads": (i32, []),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


# ------------------------------------------------------------------------------------------------
# host-side model helpers (restating src/math/terrain_model.rs constructors; oracle side)
# ------------------------------------------------------------------------------------------------

def terrain_model(kind: int, position, size: float, min_height: float, max_height: float, minor_axis: float | None = None) -> F.TerrainModel:
    """TerrainModel::planar(position, side_length, ..) / ::sphere(position, radius, ..) / ::ellipsoid(position, major, minor, ..)
    (terrain_model.rs:84-128)."""
    m = F.TerrainModel()
    m.kind = kind
    m.position[:] = list(position)
    m.scale[:] = [size, size if minor_axis is None else minor_axis, size]
    m.min_height, m.max_height = min_height, max_height
    return m


def model_scale(m: F.TerrainModel) -> float:
    """TerrainModel::scale (terrain_model.rs:183-193)."""
    if m.kind == F.BT_KIND_ELLIPSOIDAL:
        return (m.scale[0] + m.scale[1]) / 2.0
    return m.scale[0] / 2.0 if m.kind == F.BT_KIND_PLANAR else m.scale[0]


def mesh_transform(m: F.TerrainModel) -> F.MeshTransform:
    """TerrainModel::transform -> MeshUniform (terrain_model.rs:196-203, terrain_bind_group.rs:185-205)."""
    t = F.MeshTransform()
    s = [np.float32(m.scale[k]) for k in range(3)]
    wfl = [0.0] * 12
    wfl[0], wfl[4], wfl[8] = s[0], s[1], s[2]
    wfl[9], wfl[10], wfl[11] = [np.float32(m.position[k]) for k in range(3)]
    t.world_from_local[:] = [float(v) for v in wfl]
    nm = [0.0] * 9
    nm[0], nm[4], nm[8] = [float(np.float32(1.0) / s[k]) for k in range(3)]
    t.normal_matrix[:] = nm
    return t


def view_config_uniform(cfg: F.TerrainViewConfig, m: F.TerrainModel) -> F.ViewConfigUniform:
    """TileTree::new + TerrainViewConfigUniform::from_tile_tree (tile_tree.rs:135-153, terrain_view_bind_group.rs:98-115)."""
    scale = model_scale(m)
    u = F.ViewConfigUniform()
    u.tree_size = cfg.tree_size
    u.geometry_tile_count = cfg.geometry_tile_count
    u.refinement_count = cfg.refinement_count
    u.grid_size = float(cfg.grid_size)
    u.vertices_per_row = 2 * (cfg.grid_size + 2)
    u.vertices_per_tile = 2 * cfg.grid_size * (cfg.grid_size + 2)
    u.morph_distance = cfg.morph_distance * scale
    u.blend_distance = cfg.blend_distance * scale
    u.load_distance = cfg.load_distance * scale
    u.subdivision_distance = cfg.morph_distance * scale * (1.0 + cfg.subdivision_tolerance)
    u.precision_threshold_distance = cfg.precision_threshold_distance * scale
    u.morph_range = cfg.morph_range
    u.blend_range = cfg.blend_range
    return u


def default_view_config(**overrides) -> F.TerrainViewConfig:
    """TerrainViewConfig::default (terrain_view.rs:47-64)."""
    c = F.TerrainViewConfig(tree_size=8, geometry_tile_count=1000000, refinement_count=30, grid_size=16,
                            subdivision_tolerance=0.1, precision_threshold_distance=0.001, load_distance=2.5,
                            morph_distance=16.0, blend_distance=2.0, morph_range=0.2, blend_range=0.2, origin_lod=10)
    for k, v in overrides.items():
        setattr(c, k, v)
    return c


def model_approximation(m: F.TerrainModel, view_pos, origin_lod: int, approximate_height: float) -> F.ModelApproximation:
    out = F.ModelApproximation()
    vp = (f64 * 3)(*view_pos)
    lib().bto_model_approximation(C.byref(m), vp, origin_lod, approximate_height, C.byref(out))
    return out


def culling(view_pos) -> F.CullingUniform:
    """CullingUniform::from(&ExtractedView) (culling_bind_group.rs:47-55): only world_position is consumed."""
    c = F.CullingUniform()
    c.world_position[:] = [float(np.float32(v)) for v in view_pos]
    return c


def select_tiles(vc, approx, cull, mesh, spherical: bool, capacity: int | None = None, frustum_cull: bool = False,
                 min_height: float = 0.0, max_height: float = 0.0):
    cap = capacity if capacity is not None else vc.geometry_tile_count
    tiles = (F.Tile * cap)()
    ind = F.Indirect()
    res = SelectResult()
    lib().bto_select_tiles_ex(C.byref(vc), C.byref(approx), C.byref(cull), C.byref(mesh), int(spherical), int(frustum_cull),
                              min_height, max_height, tiles, cap, C.byref(ind), C.byref(res))
    arr = np.frombuffer(tiles, dtype=np.uint32).reshape(-1, 4)[: res.final_count].copy()
    return arr, ind, res


def tile_minmax(tile: np.ndarray, B: int) -> np.ndarray:
    out = np.zeros((65, 2), dtype=np.uint16)
    tile = np.ascontiguousarray(tile)
    lib().bto_tile_minmax(tile.ctypes.data, tile.shape[0], B, out.ctypes.data)
    return out


def select_tiles_minmax(vc, approx, cull, mesh, spherical: bool, min_height: float, max_height: float, entries: np.ndarray, table: np.ndarray,
                        lod_count: int, capacity: int | None = None):
    """frustum cull with per-tile height bounds from the (min, max) table of the resident atlas tiles"""
    cap = capacity if capacity is not None else vc.geometry_tile_count
    tiles = (F.Tile * cap)()
    ind, res = F.Indirect(), SelectResult()
    entries, table = np.ascontiguousarray(entries, dtype=np.uint32), np.ascontiguousarray(table, dtype=np.uint16)
    lib().bto_select_tiles_minmax(C.byref(vc), C.byref(approx), C.byref(cull), C.byref(mesh), int(spherical), 1, min_height, max_height,
                                  entries.ctypes.data, table.ctypes.data, lod_count, len(table), tiles, cap, C.byref(ind), C.byref(res))
    return np.frombuffer(tiles, dtype=np.uint32).reshape(-1, 4)[: res.final_count].copy(), ind, res


VERTEX_DTYPE = np.dtype([("position", "<f4", 4), ("normal", "<f4", 3), ("uv", "<f4", 2), ("tile_index", "<u4")])
assert VERTEX_DTYPE.itemsize == 40


def mip_chain_texels(T: int, levels: int) -> int:
    return sum((T >> k) * (T >> k) for k in range(1, levels))


def generate_mipmaps(tile: np.ndarray, levels: int) -> np.ndarray:
    """levels 1.. of one tile appended (AttachmentData::generate_mipmaps); uint16 (T, T) or uint8 (T, T, 4)"""
    T = tile.shape[0]
    fmt = F.BT_FORMAT_R16 if tile.dtype == np.uint16 else F.BT_FORMAT_RGBA8
    n = mip_chain_texels(T, levels)
    out = np.zeros(n if fmt == F.BT_FORMAT_R16 else (n, 4), dtype=tile.dtype)
    tile = np.ascontiguousarray(tile)
    if n:
        lib().bto_generate_mipmaps(fmt, T, levels, tile.ctypes.data, out.ctypes.data)
    return out


def gen_vertices_filtered(inputs: VertexInputs, height_mips, albedo_atlas, albedo_mips, mip_level_count: int, first: int, count: int,
                          albedo_attachment: int = 1, threads: int = 0):
    """vertices + per-vertex albedo with the configs[2] sampler (flags in inputs.flags); arrays may be None"""
    fi = FilterInputs()
    fi.height_mips = height_mips.ctypes.data if height_mips is not None else None
    fi.albedo_atlas = albedo_atlas.ctypes.data if albedo_atlas is not None else None
    fi.albedo_mips = albedo_mips.ctypes.data if albedo_mips is not None else None
    fi.mip_level_count, fi.albedo_attachment = mip_level_count, albedo_attachment
    out = np.empty(count, dtype=VERTEX_DTYPE)
    albedo = np.zeros(count, dtype=np.uint32)
    lib().bto_gen_vertices_filtered(C.byref(inputs), C.byref(fi), first, count, out.ctypes.data, albedo.ctypes.data, threads)
    return out, albedo


def gen_vertices(inputs: VertexInputs, first: int, count: int, threads: int = 0) -> np.ndarray:
    out = np.empty(count, dtype=VERTEX_DTYPE)
    lib().bto_gen_vertices(C.byref(inputs), first, count, out.ctypes.data, threads)
    return out


def synth_source(fmt: int, width: int, height: int, x0: int = 0, y0: int = 0, seed: int = 0x5EED0001, holes: bool = False) -> np.ndarray:
    if fmt == F.BT_FORMAT_R16:
        out = np.empty((height, width), dtype=np.uint16)
    else:
        out = np.empty((height, width, 4), dtype=np.uint8)
    lib().bto_synth_source(fmt, width, height, x0, y0, seed, int(holes), out.ctypes.data, 0)
    return out


class Pyramid:
    """Oracle twin of Preprocessor + the atlas it fills (preprocessor.rs:214-343)."""

    def __init__(self, kind: int, fmt: int, texture_size: int, border_size: int, lod_count: int):
        self.kind, self.fmt, self.T, self.B, self.lod_count = kind, fmt, texture_size, border_size, lod_count
        self._p = lib().bto_pyramid_create(kind, fmt, texture_size, border_size, lod_count)
        self._keep = []

    def add_datasets(self, datasets) -> None:
        arr = (F.Dataset * len(datasets))(*datasets)
        self._keep.append((arr, datasets))
        lib().bto_pyramid_add_datasets(self._p, arr, len(datasets))

    def run(self, threads: int = 0, phases: int = 3) -> None:
        lib().bto_pyramid_run_phases(self._p, threads, phases)

    def tiles(self) -> np.ndarray:
        n = lib().bto_pyramid_tile_count(self._p)
        buf = (F.Tile * max(n, 1))()
        lib().bto_pyramid_list_tiles(self._p, buf, n)
        return np.frombuffer(buf, dtype=np.uint32).reshape(-1, 4)[:n].copy()

    def tile(self, side: int, lod: int, x: int, y: int):
        t = F.Tile(side, lod, x, y)
        ptr = lib().bto_pyramid_tile(self._p, C.byref(t))
        if not ptr:
            return None
        if self.fmt == F.BT_FORMAT_R16:
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint16)), shape=(self.T, self.T)).copy()
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(self.T, self.T, 4)).copy()

    def close(self) -> None:
        if self._p:
            lib().bto_pyramid_destroy(self._p)
            self._p = None

    def __del__(self):
        self.close()


def make_dataset(pixels: np.ndarray, attachment_index=0, side=0, top_left=(0.0, 0.0), bottom_right=(1.0, 1.0),
                 lod_range=(0, 1)) -> F.Dataset:
    """PreprocessDataset (preprocessor.rs:36-55) with the decoded source image instead of a path."""
    d = F.Dataset()
    d.attachment_index, d.side = attachment_index, side
    d.top_left[:] = list(top_left)
    d.bottom_right[:] = list(bottom_right)
    d.lod_begin, d.lod_end = lod_range
    d.height, d.width = pixels.shape[0], pixels.shape[1]
    d.pixels_on_device = 0
    d.pixels = pixels.ctypes.data
    d._keepalive = pixels
    return d
