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

/// Per-(terrain, view) device state = `TerrainViewData` (src/render/terrain_view_bind_group.rs:118-187).
pub struct TerrainView { pub(crate) ctx: *const ffi::bt_context, pub(crate) raw: *mut ffi::bt_view }
impl TerrainView {
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
