//! Safe wrappers with the reference's names (src/lib.rs:60-86 prelude) over the C ABI.
//! NOT COMPILED in this repository (no Rust toolchain in the build image); see INTEGRATION.md for the
//! call-point mapping.  Bodies only marshal plain data: all computation happens in libbevy_terrain_b200.so.
pub mod ffi;

use anyhow::{anyhow, Result};
use std::ffi::{CStr, CString};
use std::ops::Range;
use std::ptr;

fn check(ctx: *const ffi::bt_context, status: ffi::bt_status) -> Result<()> {
    if status == ffi::BT_OK { return Ok(()); }
    let msg = unsafe { CStr::from_ptr(ffi::bt_last_error(ctx)) }.to_string_lossy().into_owned();
    Err(anyhow!("bevy_terrain_b200 status {status}: {msg}"))
}

pub struct Context { raw: *mut ffi::bt_context }
impl Context {
    pub fn new(device: i32) -> Result<Self> {
        let mut raw = ptr::null_mut();
        let st = unsafe { ffi::bt_context_create(device, &mut raw) };
        if st != ffi::BT_OK { return Err(anyhow!("no usable sm_100a CUDA device (there is no CPU fallback)")); }
        Ok(Self { raw })
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { ffi::bt_context_destroy(self.raw) } } }

/// `AttachmentFormat` (src/terrain_data/mod.rs:38-57); Rgb8/Rg16 have no shader path in the reference either.
#[derive(Clone, Copy, Debug)] pub enum AttachmentFormat { Rgba8, R16 }
/// `AttachmentConfig` (src/terrain_data/mod.rs:88-109)
#[derive(Clone, Debug)]
pub struct AttachmentConfig { pub name: String, pub texture_size: u32, pub border_size: u32, pub mip_level_count: u32, pub format: AttachmentFormat }
impl AttachmentConfig {
    fn raw(&self) -> ffi::bt_attachment_config {
        ffi::bt_attachment_config { texture_size: self.texture_size, border_size: self.border_size, mip_level_count: self.mip_level_count,
            format: match self.format { AttachmentFormat::Rgba8 => ffi::BT_FORMAT_RGBA8, AttachmentFormat::R16 => ffi::BT_FORMAT_R16 } }
    }
}

/// `TerrainModel` (src/math/terrain_model.rs:40-128): the plain-data model every host-math entry point takes.
#[derive(Clone, Copy)] pub struct TerrainModel(pub ffi::bt_terrain_model);
impl TerrainModel {
    pub fn planar(position: [f64; 3], side_length: f64, min_height: f32, max_height: f32) -> Self {
        let mut m = ffi::bt_terrain_model::default();
        unsafe { ffi::bt_terrain_model_planar(position.as_ptr(), side_length, min_height, max_height, &mut m) };
        Self(m)
    }
    pub fn sphere(position: [f64; 3], radius: f64, min_height: f32, max_height: f32) -> Self {
        let mut m = ffi::bt_terrain_model::default();
        unsafe { ffi::bt_terrain_model_sphere(position.as_ptr(), radius, min_height, max_height, &mut m) };
        Self(m)
    }
    pub fn ellipsoid(position: [f64; 3], major_axis: f64, minor_axis: f64, min_height: f32, max_height: f32) -> Self {
        let mut m = ffi::bt_terrain_model::default();
        unsafe { ffi::bt_terrain_model_ellipsoid(position.as_ptr(), major_axis, minor_axis, min_height, max_height, &mut m) };
        Self(m)
    }
    pub fn is_spherical(&self) -> bool { self.0.kind != ffi::BT_KIND_PLANAR }
    pub fn scale(&self) -> f64 { unsafe { ffi::bt_terrain_model_scale(&self.0) } }
}

/// A decoded source image (`PreprocessDataset::path` of the reference): PNG / TIFF through `bt_image_load`.
pub struct SourceImage { pub width: u32, pub height: u32, pub format: AttachmentFormat, pub pixels: Vec<u8> }
impl SourceImage {
    pub fn load(path: &str) -> Result<Self> {
        let c_path = CString::new(path)?;
        let mut image = ffi::bt_image { width: 0, height: 0, format: 0, _pad: 0, pixels: ptr::null_mut() };
        let mut error = [0 as std::os::raw::c_char; 256];
        let st = unsafe { ffi::bt_image_load(c_path.as_ptr(), &mut image, error.as_mut_ptr(), error.len()) };
        if st != ffi::BT_OK { return Err(anyhow!("{}", unsafe { CStr::from_ptr(error.as_ptr()) }.to_string_lossy())); }
        let (format, px) = if image.format == ffi::BT_FORMAT_R16 { (AttachmentFormat::R16, 2) } else { (AttachmentFormat::Rgba8, 4) };
        let bytes = image.width as usize * image.height as usize * px;
        let pixels = unsafe { std::slice::from_raw_parts(image.pixels as *const u8, bytes) }.to_vec();
        unsafe { ffi::bt_image_free(&mut image) };
        Ok(Self { width: image.width, height: image.height, format, pixels })
    }
}

/// `PreprocessDataset` (src/preprocess/preprocessor.rs:36-55); `pixels` come from `SourceImage::load(path)` or from the `image`/`tiff` crates as in the reference.
pub struct PreprocessDataset<'a> {
    pub attachment_index: u32, pub pixels: &'a [u8], pub width: u32, pub height: u32, pub side: u32,
    pub top_left: [f32; 2], pub bottom_right: [f32; 2], pub lod_range: Range<u32>,
}

/// `Preprocessor` (src/preprocess/preprocessor.rs:214-343): same builder calls, GPU work in `run`.
pub struct Preprocessor<'c> { ctx: &'c Context, raw: *mut ffi::bt_preprocessor }
impl<'c> Preprocessor<'c> {
    pub fn new(ctx: &'c Context, spherical: bool, attachments: &[AttachmentConfig], lod_count: u32) -> Result<Self> {
        let raws: Vec<_> = attachments.iter().map(|a| a.raw()).collect();
        let mut raw = ptr::null_mut();
        let kind = if spherical { ffi::BT_KIND_SPHERICAL } else { ffi::BT_KIND_PLANAR };
        check(ctx.raw, unsafe { ffi::bt_preprocessor_create(ctx.raw, kind, raws.as_ptr(), raws.len() as u32, lod_count, ffi::bt_shard { rank: 0, world: 1 }, &mut raw) })?;
        Ok(Self { ctx, raw })
    }
    fn dataset(d: &PreprocessDataset) -> ffi::bt_dataset {
        ffi::bt_dataset { attachment_index: d.attachment_index, side: d.side, top_left: d.top_left, bottom_right: d.bottom_right,
            lod_begin: d.lod_range.start, lod_end: d.lod_range.end, width: d.width, height: d.height, pixels_on_device: 0,
            row_begin: 0, row_count: 0, _pad: 0, pixels: d.pixels.as_ptr() as *const _ }
    }
    /// `Preprocessor::preprocess_tile` (preprocessor.rs:298-312)
    pub fn preprocess_tile(self, dataset: PreprocessDataset) -> Result<Self> {
        let raw = Self::dataset(&dataset);
        check(self.ctx.raw, unsafe { ffi::bt_preprocessor_add_datasets(self.raw, &raw, 1) })?;
        Ok(self)
    }
    /// `Preprocessor::preprocess_spherical` (preprocessor.rs:314-343): six faces queued as one job
    pub fn preprocess_spherical(self, faces: [PreprocessDataset; 6]) -> Result<Self> {
        let raws: Vec<_> = faces.iter().map(Self::dataset).collect();
        check(self.ctx.raw, unsafe { ffi::bt_preprocessor_add_datasets(self.raw, raws.as_ptr(), 6) })?;
        Ok(self)
    }
    /// replaces the frame-paced `TerrainPreprocessNode::run` loop (src/preprocess/mod.rs:143-217)
    pub fn run(&mut self) -> Result<()> { check(self.ctx.raw, unsafe { ffi::bt_preprocessor_run(self.raw, ptr::null_mut()) }) }
    /// Save tasks + `save_tile_config` (tile_atlas.rs:77-116,605-613)
    pub fn save(&mut self, attachment: u32, directory: &str) -> Result<()> {
        let dir = CString::new(directory)?;
        check(self.ctx.raw, unsafe { ffi::bt_preprocessor_save(self.raw, attachment, dir.as_ptr()) })
    }
}
impl Drop for Preprocessor<'_> { fn drop(&mut self) { unsafe { ffi::bt_preprocessor_destroy(self.raw) } } }

/// `TerrainConfig` (src/terrain.rs:27-56): same fields, same builder.
#[derive(Clone)]
pub struct TerrainConfig { pub lod_count: u32, pub model: TerrainModel, pub atlas_size: u32, pub path: String, pub attachments: Vec<AttachmentConfig> }
impl Default for TerrainConfig {
    fn default() -> Self { Self { lod_count: 1, model: TerrainModel::sphere([0.0; 3], 1.0, 0.0, 1.0), atlas_size: 1024, path: String::new(), attachments: vec![] } }
}
impl TerrainConfig {
    pub fn add_attachment(mut self, attachment_config: AttachmentConfig) -> Self { self.attachments.push(attachment_config); self }
}

/// `TerrainViewConfig` (src/terrain_view.rs:19-64): same twelve fields and defaults.
#[derive(Clone, Copy, Debug)]
pub struct TerrainViewConfig {
    pub tree_size: u32, pub geometry_tile_count: u32, pub refinement_count: u32, pub grid_size: u32, pub subdivision_tolerance: f64,
    pub precision_threshold_distance: f64, pub load_distance: f64, pub morph_distance: f64, pub blend_distance: f64,
    pub morph_range: f32, pub blend_range: f32, pub origin_lod: u32,
}
impl Default for TerrainViewConfig {
    fn default() -> Self {
        let mut c = ffi::bt_terrain_view_config::default();
        unsafe { ffi::bt_terrain_view_config_default(&mut c) };
        Self { tree_size: c.tree_size, geometry_tile_count: c.geometry_tile_count, refinement_count: c.refinement_count, grid_size: c.grid_size,
               subdivision_tolerance: c.subdivision_tolerance, precision_threshold_distance: c.precision_threshold_distance, load_distance: c.load_distance,
               morph_distance: c.morph_distance, blend_distance: c.blend_distance, morph_range: c.morph_range, blend_range: c.blend_range, origin_lod: c.origin_lod }
    }
}
impl TerrainViewConfig {
    fn raw(&self) -> ffi::bt_terrain_view_config {
        ffi::bt_terrain_view_config { tree_size: self.tree_size, geometry_tile_count: self.geometry_tile_count, refinement_count: self.refinement_count,
            grid_size: self.grid_size, subdivision_tolerance: self.subdivision_tolerance, precision_threshold_distance: self.precision_threshold_distance,
            load_distance: self.load_distance, morph_distance: self.morph_distance, blend_distance: self.blend_distance, morph_range: self.morph_range,
            blend_range: self.blend_range, origin_lod: self.origin_lod, _pad: 0 }
    }
}

/// `TileAtlas::new` (src/terrain_data/tile_atlas.rs:531-551) + its GPU half (`GpuTileAtlas`, gpu_tile_atlas.rs:208-223): the device atlas
/// arrays and the per-terrain `TileAtlasState` every view's `TileTree` shares.
pub struct TileAtlas<'c> { ctx: &'c Context, pub(crate) raw: *mut ffi::bt_terrain, pub(crate) state: *mut ffi::bt_atlas_state, pub config: TerrainConfig }
impl<'c> TileAtlas<'c> {
    pub fn new(ctx: &'c Context, config: &TerrainConfig) -> Result<Self> {
        let mut d: ffi::bt_terrain_desc = unsafe { std::mem::zeroed() };
        d.kind = if config.model.is_spherical() { ffi::BT_KIND_SPHERICAL } else { ffi::BT_KIND_PLANAR };
        d.lod_count = config.lod_count;
        d.atlas_size = config.atlas_size;
        d.attachment_count = config.attachments.len() as u32;
        for (i, a) in config.attachments.iter().enumerate().take(8) { d.attachments[i] = a.raw(); }
        d.min_height = config.model.0.min_height;
        d.max_height = config.model.0.max_height;
        d.scale = config.model.scale();
        unsafe { ffi::bt_terrain_model_mesh_transform(&config.model.0, &mut d.transform) };
        let (mut raw, mut state) = (ptr::null_mut(), ptr::null_mut());
        check(ctx.raw, unsafe { ffi::bt_terrain_create(ctx.raw, &d, &mut raw) })?;
        let st = unsafe { ffi::bt_atlas_state_create(config.atlas_size, &mut state) };
        if st != ffi::BT_OK { unsafe { ffi::bt_terrain_destroy(raw) }; return Err(anyhow!("bt_atlas_state_create failed")); }
        Ok(Self { ctx, raw, state, config: config.clone() })
    }
    /// `GpuTileAtlas::upload_tiles` (gpu_tile_atlas.rs:309-336) + the mip generation and height bounds the loader derives per tile
    pub fn upload_tile(&mut self, attachment: u32, atlas_index: u32, texels: &[u8]) -> Result<()> {
        check(self.ctx.raw, unsafe { ffi::bt_atlas_upload_tile(self.raw, attachment, atlas_index, texels.as_ptr() as *const _, texels.len(), ptr::null_mut()) })?;
        check(self.ctx.raw, unsafe { ffi::bt_atlas_generate_mips(self.raw, attachment, atlas_index, ptr::null_mut()) })?;
        if attachment == 0 { check(self.ctx.raw, unsafe { ffi::bt_atlas_update_minmax(self.raw, atlas_index, ptr::null_mut()) })?; }
        Ok(())
    }
}
impl Drop for TileAtlas<'_> { fn drop(&mut self) { unsafe { ffi::bt_atlas_state_destroy(self.state); ffi::bt_terrain_destroy(self.raw) } } }

/// `TileTree::new` (src/terrain_data/tile_tree.rs:135-173): the clipmap residency of one view (f64 host math in the library).
pub struct TileTree { raw: *mut ffi::bt_tile_tree, pub entries: Vec<ffi::bt_tile_tree_entry>, pub origins: Vec<u32>, pub view_config: TerrainViewConfig, model: TerrainModel }
impl TileTree {
    pub fn new(tile_atlas: &TileAtlas, view_config: &TerrainViewConfig) -> Result<Self> {
        let mut raw = ptr::null_mut();
        let c = view_config.raw();
        let st = unsafe { ffi::bt_tile_tree_create_shared(&tile_atlas.config.model.0, &c, tile_atlas.config.lod_count, tile_atlas.state, &mut raw) };
        if st != ffi::BT_OK { return Err(anyhow!("bt_tile_tree_create_shared: status {st}")); }
        let sides = if tile_atlas.config.model.is_spherical() { 6 } else { 1 };
        let n = (sides * tile_atlas.config.lod_count * view_config.tree_size * view_config.tree_size) as usize;
        Ok(Self { raw, entries: vec![ffi::bt_tile_tree_entry { atlas_index: u32::MAX, atlas_lod: u32::MAX }; n],
                  origins: vec![0; (sides * tile_atlas.config.lod_count * 2) as usize], view_config: *view_config, model: tile_atlas.config.model })
    }
    /// `TileTree::compute_requests` + `TileAtlas::update` (tile_tree.rs:268-333, tile_atlas.rs:574-601): the tiles to load this frame
    pub fn compute_requests(&mut self, view_world_position: [f64; 3], atlas_size: u32) -> Result<Vec<ffi::bt_tile_load>> {
        let mut loads = vec![ffi::bt_tile_load::default(); atlas_size as usize];
        let mut n = 0u32;
        let st = unsafe { ffi::bt_tile_tree_update(self.raw, view_world_position.as_ptr(), loads.as_mut_ptr(), atlas_size, &mut n) };
        if st != ffi::BT_OK { return Err(anyhow!("bt_tile_tree_update: status {st} (atlas out of indices?)")); }
        loads.truncate(n as usize);
        Ok(loads)
    }
    pub fn mark_loaded(&mut self, tile: &ffi::bt_tile) { unsafe { ffi::bt_tile_tree_mark_loaded(self.raw, tile) }; }
    /// `TileTree::adjust_to_tile_atlas` (tile_tree.rs:363-374)
    pub fn adjust_to_tile_atlas(&mut self) -> Result<()> {
        let st = unsafe { ffi::bt_tile_tree_entries(self.raw, self.entries.as_mut_ptr(), self.entries.len() as u32, self.origins.as_mut_ptr(), self.origins.len() as u32) };
        if st == ffi::BT_OK { Ok(()) } else { Err(anyhow!("bt_tile_tree_entries: status {st}")) }
    }
    /// the Prepare-stage payload of one frame (`TerrainViewConfigUniform`, `TerrainModelApproximation::compute`, culling, table)
    pub fn frame(&self, view_world_position: [f64; 3], approximate_height: f32) -> Result<ffi::bt_view_frame> {
        let mut f: ffi::bt_view_frame = unsafe { std::mem::zeroed() };
        let c = self.view_config.raw();
        unsafe { ffi::bt_view_config_uniform_from(&c, &self.model.0, &mut f.view_config) };
        let st = unsafe { ffi::bt_model_approximation_compute(&self.model.0, view_world_position.as_ptr(), self.view_config.origin_lod, approximate_height, &mut f.approximation) };
        if st != ffi::BT_OK { return Err(anyhow!("bt_model_approximation_compute: status {st}")); }
        let p = [view_world_position[0] as f32, view_world_position[1] as f32, view_world_position[2] as f32];
        f.culling.world_position = p;
        f.view_world_position = p;
        f.tile_tree_entry_count = self.entries.len() as u32;
        f.tile_tree = self.entries.as_ptr();
        f.origins = self.origins.as_ptr();
        Ok(f)
    }
}
impl Drop for TileTree { fn drop(&mut self) { unsafe { ffi::bt_tile_tree_destroy(self.raw) } } }

/// `SphericalDataset` (src/preprocess/preprocessor.rs:29-33): one image per cube face.
pub struct SphericalDataset<'a> { pub attachment_index: u32, pub faces: [&'a [u8]; 6], pub width: u32, pub height: u32, pub lod_range: Range<u32> }

/// The multi-GPU communicator of the sharded preprocessor (one process per GPU; NCCL over NVLink).  Bootstrap like any NCCL
/// program: rank 0 calls `Comm::unique_id`, the host distributes the 128 bytes, every rank calls `Comm::new`.
pub struct Comm<'c> { ctx: &'c Context, raw: *mut ffi::bt_comm }
impl<'c> Comm<'c> {
    pub fn unique_id() -> Result<[u8; ffi::BT_COMM_UNIQUE_ID_BYTES]> {
        let mut id = [0u8; ffi::BT_COMM_UNIQUE_ID_BYTES];
        let st = unsafe { ffi::bt_comm_unique_id(id.as_mut_ptr() as *mut _, id.len()) };
        if st == ffi::BT_OK { Ok(id) } else { Err(anyhow!("bt_comm_unique_id: {}", unsafe { CStr::from_ptr(ffi::bt_comm_load_error()) }.to_string_lossy())) }
    }
    pub fn new(ctx: &'c Context, rank: u32, world: u32, unique_id: &[u8; ffi::BT_COMM_UNIQUE_ID_BYTES]) -> Result<Self> {
        let mut raw = ptr::null_mut();
        check(ctx.raw, unsafe { ffi::bt_comm_create(ctx.raw, rank, world, unique_id.as_ptr() as *const _, unique_id.len(), &mut raw) })?;
        Ok(Self { ctx, raw })
    }
}
impl Drop for Comm<'_> { fn drop(&mut self) { unsafe { ffi::bt_comm_destroy(self.raw) } } }

impl<'c> Preprocessor<'c> {
    /// a preprocessor that owns rank `rank` of `world` row bands of the deepest LOD; run it with `run_sharded`
    pub fn new_sharded(ctx: &'c Context, spherical: bool, attachments: &[AttachmentConfig], lod_count: u32, rank: u32, world: u32) -> Result<Self> {
        let raws: Vec<_> = attachments.iter().map(|a| a.raw()).collect();
        let mut raw = ptr::null_mut();
        let kind = if spherical { ffi::BT_KIND_SPHERICAL } else { ffi::BT_KIND_PLANAR };
        check(ctx.raw, unsafe { ffi::bt_preprocessor_create(ctx.raw, kind, raws.as_ptr(), raws.len() as u32, lod_count, ffi::bt_shard { rank, world }, &mut raw) })?;
        Ok(Self { ctx, raw })
    }
    /// `Preprocessor::preprocess_spherical` with the reference's `SphericalDataset`
    pub fn preprocess_spherical_dataset(self, dataset: SphericalDataset) -> Result<Self> {
        let faces: [PreprocessDataset; 6] = std::array::from_fn(|side| PreprocessDataset { attachment_index: dataset.attachment_index, pixels: dataset.faces[side],
            width: dataset.width, height: dataset.height, side: side as u32, top_left: [0.0, 0.0], bottom_right: [1.0, 1.0], lod_range: dataset.lod_range.clone() });
        self.preprocess_spherical(faces)
    }
    /// the whole sharded job in the library: kernels + one grouped NCCL all-gather per wave, overlapped with the interior stitch
    pub fn run_sharded(&mut self, comm: &Comm) -> Result<()> {
        check(self.ctx.raw, unsafe { ffi::bt_preprocessor_reserve(self.raw, ptr::null_mut()) })?;
        check(self.ctx.raw, unsafe { ffi::bt_preprocessor_run_sharded(self.raw, comm.raw, ptr::null_mut()) })
    }
}

/// Per-(terrain, view) device state = `TerrainViewData` (src/render/terrain_view_bind_group.rs:118-187).
pub struct TerrainView { pub(crate) ctx: *const ffi::bt_context, pub(crate) raw: *mut ffi::bt_view }
impl TerrainView {
    /// `TerrainViewData::new` (terrain_view_bind_group.rs:118-165): final / temporary tile buffers, indirect record, uniforms, tile-tree table,
    /// the vertex stream (`max_vertices` records) and the pinned read-back buffers of `run_frame`; nothing is allocated per frame afterwards
    pub fn new(tile_atlas: &TileAtlas, view_config: &TerrainViewConfig, max_vertices: u64) -> Result<Self> {
        let desc = ffi::bt_view_desc { geometry_tile_count: view_config.geometry_tile_count, tree_size: view_config.tree_size, grid_size: view_config.grid_size,
                                       readback_tile_capacity: 0, max_vertices };
        let mut raw = ptr::null_mut();
        check(tile_atlas.ctx.raw, unsafe { ffi::bt_view_create(tile_atlas.raw, &desc, &mut raw) })?;
        Ok(Self { ctx: tile_atlas.ctx.raw, raw })
    }
    /// Prepare stage: the reference's `write_buffer`s (terrain_view_bind_group.rs:238-262, gpu_tile_tree.rs:84-95)
    pub fn update(&mut self, frame: &ffi::bt_view_frame) -> Result<()> { check(self.ctx, unsafe { ffi::bt_view_update(self.raw, frame, ptr::null_mut()) }) }
    /// `TilingPrepassNode::run` (src/render/tiling_prepass.rs:204-271) in one launch
    pub fn select_tiles(&mut self) -> Result<()> { check(self.ctx, unsafe { ffi::bt_select_tiles(self.raw, 0, ptr::null_mut()) }) }
    /// the vertex stage of `DrawTerrainCommand`'s indirect draw, written to a device vertex buffer
    pub fn gen_vertices(&mut self, flags: u32) -> Result<ffi::bt_vertex_buffer> {
        let mut vb = ffi::bt_vertex_buffer { vertices: ptr::null_mut(), capacity: 0 };
        check(self.ctx, unsafe { ffi::bt_gen_vertices(self.raw, flags, ptr::null_mut(), &mut vb) })?;
        Ok(vb)
    }
}
impl TerrainView {
    /// One frame in one call (update + select + generate + overlapped tile-list read-back): the path `bench.py` times as `e2e`.
    pub fn run_frame(&mut self, frame: &ffi::bt_view_frame, select_flags: u32, gen_flags: u32, tiles: &mut Vec<ffi::bt_tile>)
        -> Result<(ffi::bt_indirect, ffi::bt_vertex_buffer)> {
        let mut count = 0u32;
        let mut indirect = ffi::bt_indirect::default();
        let mut vb = ffi::bt_vertex_buffer { vertices: ptr::null_mut(), capacity: 0 };
        let cap = tiles.capacity() as u32;
        check(self.ctx, unsafe {
            ffi::bt_view_run_frame(self.raw, frame, select_flags, gen_flags, ptr::null_mut(), tiles.as_mut_ptr(), cap, &mut count, &mut indirect, &mut vb)
        })?;
        unsafe { tiles.set_len(count.min(cap) as usize) };
        Ok((indirect, vb))
    }
}

impl Drop for TerrainView { fn drop(&mut self) { unsafe { ffi::bt_view_destroy(self.raw) } } }

/// The device twin of the per-frame half of `TileTree` (`bt_clipmap`): slots, origins and the best-tile table stay in GPU
/// memory; `update` = `TileTree::compute_requests` + `TileAtlas::update`, `adjust_to_tile_atlas` writes the view's table in place.
pub struct DeviceClipmap { ctx: *const ffi::bt_context, raw: *mut ffi::bt_clipmap }
impl DeviceClipmap {
    pub fn new(view: &TerrainView, tree: &TileTree) -> Result<Self> {
        let mut raw = ptr::null_mut();
        check(view.ctx, unsafe { ffi::bt_clipmap_create(view.raw, tree.raw, &mut raw) })?;
        Ok(Self { ctx: view.ctx, raw })
    }
    pub fn update(&mut self, view_world_position: [f64; 3], atlas_size: u32) -> Result<Vec<ffi::bt_tile_load>> {
        let mut loads = vec![ffi::bt_tile_load::default(); atlas_size as usize];
        let mut n = 0u32;
        check(self.ctx, unsafe { ffi::bt_clipmap_update(self.raw, view_world_position.as_ptr(), loads.as_mut_ptr(), atlas_size, &mut n, ptr::null_mut()) })?;
        loads.truncate(n as usize);
        Ok(loads)
    }
    pub fn adjust_to_tile_atlas(&mut self) -> Result<()> { check(self.ctx, unsafe { ffi::bt_clipmap_entries(self.raw, ptr::null_mut()) }) }
}
impl Drop for DeviceClipmap { fn drop(&mut self) { unsafe { ffi::bt_clipmap_destroy(self.raw) } } }

/// `TilingPrepassNode::run` (src/render/tiling_prepass.rs:204-271) for every (terrain, view) pair of a frame: the reference loops on the
/// host with 63 dispatches per pair; here all views are refined by ONE launch (one CTA per view).
pub struct TilingPrepass { ctx: *const ffi::bt_context, raw: *mut ffi::bt_view_batch }
impl TilingPrepass {
    pub fn new(views: &[&TerrainView]) -> Result<Self> {
        let mut ptrs: Vec<*mut ffi::bt_view> = views.iter().map(|v| v.raw).collect();
        let mut raw = ptr::null_mut();
        let ctx = views.first().map(|v| v.ctx).unwrap_or(ptr::null());
        check(ctx, unsafe { ffi::bt_view_batch_create(ptrs.as_mut_ptr(), ptrs.len() as u32, &mut raw) })?;
        Ok(Self { ctx, raw })
    }
    pub fn run(&mut self, flags: u32) -> Result<()> { check(self.ctx, unsafe { ffi::bt_select_tiles_batch(self.raw, flags, ptr::null_mut()) }) }
}
impl Drop for TilingPrepass { fn drop(&mut self) { unsafe { ffi::bt_view_batch_destroy(self.raw) } } }

/// The reference's prelude (src/lib.rs:60-86), for `use bevy_terrain_b200::prelude::*`.  `TerrainPlugin`, `TerrainBundle`,
/// `TerrainMaterialPlugin` and `TerrainPreprocessPlugin` are Bevy ECS glue (plugin.rs:32, terrain.rs:62-98, terrain_material.rs:437,
/// preprocess/mod.rs:255): they stay in the reference crate and call the types above at the points INTEGRATION.md lists.
pub mod prelude {
    pub use crate::{AttachmentConfig, AttachmentFormat, Comm, Context, DeviceClipmap, PreprocessDataset, Preprocessor, SourceImage, SphericalDataset, TerrainConfig,
                    TerrainModel, TerrainView, TerrainViewConfig, TileAtlas, TileTree, TilingPrepass};
}
