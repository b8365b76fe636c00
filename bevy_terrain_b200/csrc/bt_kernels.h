// bt_kernels.h — internal launch interface between the C-ABI runtime (bt_api.cu) and the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bevy_terrain_b200.h"

namespace btd {

// written by select_tiles_kernel (device), read back by bt_view_select_status / bt_view_readback
struct SelectStatus {
    uint32_t final_count;
    uint32_t dropped_parents;  // parents that still wanted subdivision in the last pass (reference Q2: dropped)
    uint32_t max_frontier;
    uint32_t overflow;         // 1 = a tile list exceeded its capacity (unchecked in the reference)
};

// device-resident launch record of one view for select_tiles_kernel (one CTA per entry)
struct SelectViewParams {
    const void* uniforms;      // ViewUniforms
    bt_tile* final_tiles;
    bt_tile* temporary_tiles;
    bt_indirect* indirect;
    SelectStatus* status;
    uint32_t capacity;
    uint32_t atlas_size;
    const bt_tile_tree_entry* tile_tree;   // the view's table (BT_FLAG_MINMAX_CULL)
    const uint16_t* minmax;                // the terrain's [atlas_size][65][2] (min, max) table, or nullptr
    // bt_view_run_frame: the frame's inputs are pulled by the select kernel itself from the view's pinned (mapped) staging block instead
    // of being pushed by two copy-engine operations in front of it (kSelectFetch* bits of the launch flags)
    const void* host_uniforms;             // device alias of the pinned ViewUniforms staging
    const void* host_table;                // device alias of the pinned tile-tree staging
    uint32_t table_entries;
    uint32_t _pad;
};
constexpr uint32_t kSelectFetchUniforms = 1u << 30;   // internal launch flags (masked off the public flags)
constexpr uint32_t kSelectFetchTable = 1u << 31;

struct GenStatus {
    uint32_t vertex_count_lo, vertex_count_hi;
    uint32_t overflow;         // vertex buffer too small for indirect.vertex_count
    uint32_t _pad;
};

// one preprocess task = one output tile of one phase (SplitData / DownsampleData / StitchData,
// src/preprocess/gpu_preprocessor.rs:31-50); slots index the device pyramid store
constexpr uint32_t kTaskFresh = 1u;

struct PyramidTask {
    bt_tile  tile;
    uint32_t slot;          // output tile slot
    uint32_t dataset;       // split: index into the dataset table
    uint32_t dep[8];        // downsample: 4 child slots; stitch: 8 neighbour slots; 0xFFFFFFFF = invalid
    uint32_t dep_side[8];   // stitch: neighbour side (for project_to_side)
    uint32_t flags;         // kTaskFresh: the output slot has never been written (reads as the zeroed atlas)
    uint32_t _pad;
};

// one contiguous row copy of the halo / root exchange (multi-GPU preprocess)
struct RowCopy {
    const void* src;
    void* dst;
    uint64_t bytes;   // multiple of 2
};

// extra inputs of gen_vertices for BT_FLAG_FILTER_TRILINEAR_ANISO / BT_FLAG_ALBEDO (all device pointers)
struct GenFilterInputs {
    const uint16_t* height_mips;
    const void* albedo_atlas;
    const void* albedo_mips;
    const float* srgb_lut;
    uint32_t* albedo_out;
    uint32_t mip_level_count, mip_chain_texels, albedo_attachment, _pad;
};

// ---- device clipmap (clipmap.cu) ----
struct ClipmapParams {   // one frame's inputs of clipmap_update_kernel
    double position[3], scale[3];      // TerrainModel world_from_local (scale, translation) — terrain_model.rs:61-82
    double view_world_position[3];
    double view_uv[2];                 // Coordinate::from_world_position(view) — coordinate.rs:69-108 (one f64 evaluation, done by the host)
    double load_distance;              // pre-multiplied by the model scale (tile_tree.rs:148)
    float approximate_height;
    uint32_t view_side, kind, tree_size, lod_count, side_count, _pad;
};
struct ClipmapSlot { uint32_t x, y, requested, valid; };                 // TileState of one (side, lod, x % tree, y % tree) slot
struct ClipmapLists { uint32_t n_released, n_requested, _pad[2]; };        // header of the list buffer: [header | released | requested]
struct ClipmapLoaded { uint32_t side_lod, x, y, atlas_index; };           // open-addressing set of the loaded tiles; side_lod 0xFFFFFFFF = empty

struct TileHashTask { uint64_t key; uint32_t slot, _pad; };   // key = hash of (attachment-independent) tile coordinate

struct DatasetDevice {
    const void* pixels;     // device pointer to source row `row_begin`, row-major
    uint32_t width, height; // of the WHOLE image
    float top_left[2], bottom_right[2];
    uint32_t row_begin, row_count;  // rows present behind `pixels` (row_count == height for a whole image)
};

}  // namespace btd

extern "C" {
cudaError_t bt_launch_select_tiles(const btd::SelectViewParams* device_views, uint32_t n_views, int spherical, uint32_t flags,
                                   cudaStream_t stream);
cudaError_t bt_launch_gen_vertices(const void* uniforms, const bt_tile* final_tiles, const bt_indirect* indirect,
                                   const bt_tile_tree_entry* tile_tree, const uint16_t* height_atlas,
                                   bt_vertex* vertices, uint64_t vertex_capacity, btd::GenStatus* status,
                                   int spherical, uint32_t flags, uint32_t grid_size, int sm_count, const btd::GenFilterInputs* filter,
                                   cudaStream_t stream);
}
