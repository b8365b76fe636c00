// bt_runtime.h — internal definitions of the opaque C-ABI handles.
#pragma once

#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/bevy_terrain_b200.h"
#include "bt_kernels.h"

struct bt_context {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0;
    std::mutex error_mutex;
    std::string error;
    void set_error(const char* fmt, ...);
};

#define BT_CUDA(ctx, expr)                                                                         \
    do {                                                                                           \
        cudaError_t bt_e_ = (expr);                                                                \
        if (bt_e_ != cudaSuccess) {                                                                \
            (ctx)->set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(bt_e_)); \
            (void)cudaGetLastError();                                                              \
            return bt_e_ == cudaErrorMemoryAllocation ? BT_ERR_OUT_OF_MEMORY : BT_ERR_CUDA;        \
        }                                                                                          \
    } while (0)

struct bt_terrain {
    bt_context* ctx = nullptr;
    bt_terrain_desc desc{};
    void* atlas[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t tile_bytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    void* atlas_mips[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // levels 1.. of every slot (generate_mipmaps layout)
    size_t mip_chain_bytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t mip_chain_texels[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint16_t* d_minmax = nullptr;   // [atlas_size][65][2]: (min, max) of every resident height tile: whole tile + 8 x 8 cells
    float* d_srgb_lut = nullptr;    // 256 floats: the sRGB EOTF (Rgba8UnormSrgb decodes per texel before filtering)
    void* h_stage = nullptr;        // pinned staging for file loads
    size_t h_stage_bytes = 0;
    cudaEvent_t stage_free = nullptr;
};

struct bt_view {
    bt_terrain* terrain = nullptr;
    bt_view_desc desc{};
    uint32_t tile_tree_entries = 0;
    void* d_uniforms = nullptr;
    bt_tile* d_final_tiles = nullptr;
    bt_tile* d_temporary_tiles = nullptr;
    bt_indirect* d_indirect = nullptr;
    btd::SelectStatus* d_select_status = nullptr;
    btd::GenStatus* d_gen_status = nullptr;
    btd::SelectViewParams* d_select_params = nullptr;   // this view's launch record (a batch of one)
    bt_tile_tree_entry* d_tile_tree = nullptr;
    bt_vertex* d_vertices = nullptr;
    uint64_t vertex_capacity = 0;
    uint32_t* d_albedo = nullptr;   // per-vertex pack4x8unorm albedo stream (BT_FLAG_ALBEDO), same capacity as d_vertices
    uint64_t albedo_capacity = 0;
    void* h_uniforms = nullptr;
    void* h_tile_tree = nullptr;
    void* d_host_uniforms = nullptr;   // device aliases of the two pinned staging blocks (mapped host memory)
    void* d_host_table = nullptr;
    void* h_readback = nullptr;
    cudaEvent_t staging_free = nullptr;
    // bt_view_frame: the tile-list readback runs on its own stream so that it overlaps gen_vertices
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t select_done = nullptr, copy_done = nullptr;
    bt_tile* h_tiles = nullptr;      // pinned
    uint32_t h_tiles_cap = 0;
    size_t h2d_bytes_last_update = 0;
    bool updated = false;
    bool device_table = false;       // a bt_clipmap writes d_tile_tree on the device: bt_view_update may omit the host table
};

// several views of one frame refined by ONE select_tiles launch (gridDim.x = views)
struct bt_view_batch {
    bt_context* ctx = nullptr;
    std::vector<bt_view*> views;
    btd::SelectViewParams* d_params = nullptr;
    int spherical = 0;
};

extern "C" {
cudaError_t bt_launch_clipmap_update(const btd::ClipmapParams* params, btd::ClipmapSlot* slots, uint32_t* origins, btd::ClipmapLists* lists,
                                     bt_tile* released, bt_tile* requested, cudaStream_t stream);
cudaError_t bt_launch_clipmap_entries(const btd::ClipmapSlot* slots, uint32_t n, uint32_t tree_size, uint32_t lod_count,
                                      const btd::ClipmapLoaded* loaded, uint32_t hash_capacity, bt_tile_tree_entry* entries, cudaStream_t stream);
uint32_t bt_clipmap_hash(uint32_t side_lod, uint32_t x, uint32_t y);
cudaError_t bt_launch_split(void* arena, uint64_t tile_bytes, uint32_t T, uint32_t B, uint32_t format,
                            const btd::PyramidTask* tasks, uint32_t n_tasks, const btd::DatasetDevice* datasets,
                            float src_per_px_x, float src_per_px_y, cudaStream_t stream);
cudaError_t bt_launch_downsample(void* arena, uint64_t tile_bytes, uint32_t T, uint32_t B, uint32_t format,
                                 const btd::PyramidTask* tasks, uint32_t n_tasks, cudaStream_t stream);
cudaError_t bt_launch_stitch(void* arena, uint64_t tile_bytes, uint32_t T, uint32_t B, uint32_t format,
                             const btd::PyramidTask* tasks, uint32_t n_tasks, cudaStream_t stream);
cudaError_t bt_launch_mipmaps(uint32_t format, uint32_t texture_size, uint32_t mip_level_count, const void* mip0,
                              void* mips_out, cudaStream_t stream);
cudaError_t bt_launch_synth(uint32_t format, uint32_t width, uint32_t height, uint32_t x0, uint32_t y0, uint32_t seed,
                            uint32_t holes, void* out, cudaStream_t stream);
cudaError_t bt_launch_selftest(const float* x, float* log2_out, uint32_t n, uint32_t* mismatches, cudaStream_t stream);
cudaError_t bt_launch_row_copy(const btd::RowCopy* copies, uint32_t n, cudaStream_t stream);
cudaError_t bt_launch_tile_minmax(const uint16_t* tile, uint32_t T, uint32_t B, uint16_t* out130, cudaStream_t stream);
cudaError_t bt_launch_tile_hash(const void* arena, uint64_t tile_bytes, const btd::TileHashTask* tasks, uint32_t n,
                                unsigned long long* out, cudaStream_t stream);
}
