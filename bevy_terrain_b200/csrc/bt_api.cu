// bt_api.cu — the C-ABI runtime declared in include/bevy_terrain_b200.h: contexts, the atlas
// (GpuTileAtlas mirror), views (TerrainViewData mirror) and the per-frame entry points.
// The preprocessor lives in bt_preprocess.cu, the f64 host math in host_math.cpp.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cstdlib>

#include "bt_device.cuh"
#include "bt_kernels.h"
#include "bt_runtime.h"

using btd::ViewUniforms;

void bt_context::set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    std::lock_guard<std::mutex> g(error_mutex);
    error = buf;
}

extern "C" {

uint32_t bt_abi_version(void) { return BT_ABI_VERSION; }

bt_status bt_context_create(int32_t device, bt_context** out) {
    if (!out) return BT_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) {
        // no CPU fallback: the caller gets an error, never silently different results
        (void)cudaGetLastError();
        return BT_ERR_CUDA;
    }
    bt_context* ctx = new bt_context();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return BT_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    if (prop.major < 10) {  // kernels are built for sm_100a only
        delete ctx;
        return BT_ERR_CUDA;
    }
    *out = ctx;
    return BT_OK;
}

void bt_context_destroy(bt_context* ctx) { delete ctx; }

const char* bt_last_error(const bt_context* ctx) { return ctx ? ctx->error.c_str() : "null context"; }

bt_status bt_context_synchronize(bt_context* ctx) {
    if (!ctx) return BT_ERR_INVALID_ARG;
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    BT_CUDA(ctx, cudaDeviceSynchronize());
    return BT_OK;
}

// ------------------------------------------------------------------------------------------------
// terrain / atlas
// ------------------------------------------------------------------------------------------------
bt_status bt_terrain_create(bt_context* ctx, const bt_terrain_desc* desc, bt_terrain** out) {
    if (!ctx || !desc || !out) return BT_ERR_INVALID_ARG;
    if (desc->attachment_count == 0 || desc->attachment_count > 8 || desc->lod_count == 0 || desc->lod_count > 31 || desc->atlas_size == 0) {
        ctx->set_error("bt_terrain_create: invalid attachment_count/lod_count/atlas_size");
        return BT_ERR_INVALID_ARG;
    }
    for (uint32_t i = 0; i < desc->attachment_count; ++i) {
        const bt_attachment_config& a = desc->attachments[i];
        if ((a.format != BT_FORMAT_R16 && a.format != BT_FORMAT_RGBA8) || a.texture_size <= 2 * a.border_size) {
            ctx->set_error("bt_terrain_create: attachment %u: unsupported format %u or texture_size <= 2*border", i, a.format);
            return BT_ERR_INVALID_ARG;
        }
    }
    if (desc->attachments[0].format != BT_FORMAT_R16) {
        ctx->set_error("bt_terrain_create: attachment 0 must be the R16 height attachment");
        return BT_ERR_INVALID_ARG;
    }
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    bt_terrain* t = new bt_terrain();
    t->ctx = ctx;
    t->desc = *desc;
    for (uint32_t i = 0; i < desc->attachment_count; ++i) {
        const bt_attachment_config& a = desc->attachments[i];
        const size_t px = a.format == BT_FORMAT_R16 ? 2 : 4;
        t->tile_bytes[i] = (size_t)a.texture_size * a.texture_size * px;
        cudaError_t e = cudaMalloc(&t->atlas[i], t->tile_bytes[i] * desc->atlas_size);
        if (e == cudaSuccess) e = cudaMemset(t->atlas[i], 0, t->tile_bytes[i] * desc->atlas_size);  // wgpu zero-initialises textures
        if (e != cudaSuccess) {
            ctx->set_error("bt_terrain_create: atlas allocation of %zu bytes failed: %s", t->tile_bytes[i] * desc->atlas_size, cudaGetErrorString(e));
            bt_terrain_destroy(t);
            return e == cudaErrorMemoryAllocation ? BT_ERR_OUT_OF_MEMORY : BT_ERR_CUDA;
        }
        // mip levels 1.. of every slot (AttachmentData::generate_mipmaps appends them after mip 0, terrain_data/mod.rs:143-219):
        // filled by bt_atlas_generate_mips, consumed by BT_FLAG_FILTER_TRILINEAR_ANISO
        uint32_t chain = 0;
        for (uint32_t level = 1; level < a.mip_level_count; ++level) chain += (a.texture_size >> level) * (a.texture_size >> level);
        t->mip_chain_texels[i] = chain;
        t->mip_chain_bytes[i] = (size_t)chain * px;
        if (chain) {
            e = cudaMalloc(&t->atlas_mips[i], t->mip_chain_bytes[i] * desc->atlas_size);
            if (e == cudaSuccess) e = cudaMemset(t->atlas_mips[i], 0, t->mip_chain_bytes[i] * desc->atlas_size);
            if (e != cudaSuccess) { ctx->set_error("bt_terrain_create: mip allocation failed: %s", cudaGetErrorString(e)); bt_terrain_destroy(t); return BT_ERR_OUT_OF_MEMORY; }
        }
    }
    {   // per-slot (min, max) height table for BT_FLAG_MINMAX_CULL: zero = "nothing valid" = the terrain-wide range
        const size_t bytes = (size_t)desc->atlas_size * 65 * 2 * sizeof(uint16_t);
        cudaError_t e = cudaMalloc((void**)&t->d_minmax, bytes);
        if (e == cudaSuccess) e = cudaMemset(t->d_minmax, 0, bytes);
        if (e != cudaSuccess) { ctx->set_error("bt_terrain_create: %s", cudaGetErrorString(e)); bt_terrain_destroy(t); return BT_ERR_CUDA; }
    }
    {   // the sRGB EOTF table (IEC 61966-2-1), computed in double and rounded once: the same formula as the oracle's
        float lut[256];
        for (int i = 0; i < 256; ++i) {
            const double c = (double)i / 255.0;
            lut[i] = (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
        }
        cudaError_t e = cudaMalloc((void**)&t->d_srgb_lut, sizeof(lut));
        if (e == cudaSuccess) e = cudaMemcpy(t->d_srgb_lut, lut, sizeof(lut), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { ctx->set_error("bt_terrain_create: %s", cudaGetErrorString(e)); bt_terrain_destroy(t); return BT_ERR_CUDA; }
    }
    *out = t;
    return BT_OK;
}

// GPU replacement of the CPU mip generation at tile load (tile_atlas.rs:141 -> AttachmentData::generate_mipmaps): levels 1.. of
// one atlas slot from its mip 0, same integer rules (R16: mean of the non-zero texels; Rgba8: sum / 4)
bt_status bt_atlas_generate_mips(bt_terrain* t, uint32_t attachment, uint32_t atlas_index, void* stream) {
    if (!t || attachment >= t->desc.attachment_count || atlas_index >= t->desc.atlas_size) return BT_ERR_INVALID_ARG;
    if (!t->atlas_mips[attachment]) return BT_OK;     // mip_level_count == 1
    const bt_attachment_config& a = t->desc.attachments[attachment];
    BT_CUDA(t->ctx, cudaSetDevice(t->ctx->device));
    BT_CUDA(t->ctx, bt_launch_mipmaps(a.format, a.texture_size, a.mip_level_count,
                                      static_cast<const unsigned char*>(t->atlas[attachment]) + t->tile_bytes[attachment] * atlas_index,
                                      static_cast<unsigned char*>(t->atlas_mips[attachment]) + t->mip_chain_bytes[attachment] * atlas_index, (cudaStream_t)stream));
    return BT_OK;
}

// (min, max) of the height tile in one atlas slot: whole centre + 8 x 8 cells, each grown by a texel, zeros ignored (BT_FLAG_MINMAX_CULL)
bt_status bt_atlas_update_minmax(bt_terrain* t, uint32_t atlas_index, void* stream) {
    if (!t || atlas_index >= t->desc.atlas_size) return BT_ERR_INVALID_ARG;
    const bt_attachment_config& a = t->desc.attachments[0];
    BT_CUDA(t->ctx, cudaSetDevice(t->ctx->device));
    BT_CUDA(t->ctx, bt_launch_tile_minmax(static_cast<const uint16_t*>(t->atlas[0]) + (size_t)a.texture_size * a.texture_size * atlas_index, a.texture_size,
                                          a.border_size, t->d_minmax + (size_t)atlas_index * 65 * 2, (cudaStream_t)stream));
    return BT_OK;
}

bt_status bt_atlas_minmax_device_ptr(bt_terrain* t, void** out) {
    if (!t || !out) return BT_ERR_INVALID_ARG;
    *out = t->d_minmax;
    return BT_OK;
}

bt_status bt_atlas_mips_device_ptr(bt_terrain* t, uint32_t attachment, void** out, size_t* chain_bytes_per_slot) {
    if (!t || !out || attachment >= t->desc.attachment_count) return BT_ERR_INVALID_ARG;
    *out = t->atlas_mips[attachment];
    if (chain_bytes_per_slot) *chain_bytes_per_slot = t->mip_chain_bytes[attachment];
    return BT_OK;
}

void bt_terrain_destroy(bt_terrain* t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    for (int i = 0; i < 8; ++i) { if (t->atlas[i]) cudaFree(t->atlas[i]); if (t->atlas_mips[i]) cudaFree(t->atlas_mips[i]); }
    cudaFree(t->d_srgb_lut);
    cudaFree(t->d_minmax);
    if (t->h_stage) cudaFreeHost(t->h_stage);
    if (t->stage_free) cudaEventDestroy(t->stage_free);
    delete t;
}

static bt_status atlas_slot(bt_terrain* t, uint32_t attachment, uint32_t atlas_index, size_t bytes, unsigned char** ptr) {
    if (!t) return BT_ERR_INVALID_ARG;
    if (attachment >= t->desc.attachment_count || atlas_index >= t->desc.atlas_size || bytes != t->tile_bytes[attachment]) {
        t->ctx->set_error("atlas access: attachment %u index %u bytes %zu out of range (tile is %zu bytes)", attachment, atlas_index,
                          bytes, attachment < 8 ? t->tile_bytes[attachment] : 0);
        return BT_ERR_INVALID_ARG;
    }
    *ptr = static_cast<unsigned char*>(t->atlas[attachment]) + t->tile_bytes[attachment] * atlas_index;
    return BT_OK;
}

bt_status bt_atlas_upload_tile(bt_terrain* t, uint32_t attachment, uint32_t atlas_index, const void* texels, size_t bytes, void* stream) {
    unsigned char* dst;
    if (!texels) return BT_ERR_INVALID_ARG;
    bt_status s = atlas_slot(t, attachment, atlas_index, bytes, &dst);
    if (s != BT_OK) return s;
    BT_CUDA(t->ctx, cudaSetDevice(t->ctx->device));
    BT_CUDA(t->ctx, cudaMemcpyAsync(dst, texels, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return BT_OK;
}

bt_status bt_atlas_copy_tile(bt_terrain* t, uint32_t attachment, uint32_t atlas_index, const void* device_texels, size_t bytes, void* stream) {
    unsigned char* dst;
    if (!device_texels) return BT_ERR_INVALID_ARG;
    bt_status s = atlas_slot(t, attachment, atlas_index, bytes, &dst);
    if (s != BT_OK) return s;
    BT_CUDA(t->ctx, cudaSetDevice(t->ctx->device));
    BT_CUDA(t->ctx, cudaMemcpyAsync(dst, device_texels, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return BT_OK;
}

bt_status bt_atlas_download_tile(bt_terrain* t, uint32_t attachment, uint32_t atlas_index, void* texels, size_t bytes) {
    unsigned char* src;
    if (!texels) return BT_ERR_INVALID_ARG;
    bt_status s = atlas_slot(t, attachment, atlas_index, bytes, &src);
    if (s != BT_OK) return s;
    BT_CUDA(t->ctx, cudaSetDevice(t->ctx->device));
    BT_CUDA(t->ctx, cudaMemcpy(texels, src, bytes, cudaMemcpyDeviceToHost));
    return BT_OK;
}

bt_status bt_atlas_load_tile_file(bt_terrain* t, uint32_t attachment, uint32_t atlas_index, const char* directory, const bt_tile* tile, void* stream) {
    if (!t || !directory || !tile || attachment >= t->desc.attachment_count) return BT_ERR_INVALID_ARG;
    unsigned char* dst;
    const size_t bytes = t->tile_bytes[attachment];
    bt_status s = atlas_slot(t, attachment, atlas_index, bytes, &dst);
    if (s != BT_OK) return s;
    bt_context* ctx = t->ctx;
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    if (t->h_stage_bytes < bytes) {
        if (t->h_stage) BT_CUDA(ctx, cudaFreeHost(t->h_stage));
        t->h_stage = nullptr;
        BT_CUDA(ctx, cudaHostAlloc(&t->h_stage, bytes, cudaHostAllocDefault));
        t->h_stage_bytes = bytes;
    }
    if (!t->stage_free) BT_CUDA(ctx, cudaEventCreateWithFlags(&t->stage_free, cudaEventDisableTiming));
    BT_CUDA(ctx, cudaEventSynchronize(t->stage_free));  // the previous copy out of the staging buffer has finished
    const std::string path = std::string(directory) + "/" + std::to_string(tile->side) + "_" + std::to_string(tile->lod) + "_" +
                             std::to_string(tile->x) + "_" + std::to_string(tile->y) + ".bin";   // TileCoordinate::path, coordinate.rs:184-186,282-285
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { ctx->set_error("bt_atlas_load_tile_file: cannot open %s", path.c_str()); return BT_ERR_IO; }
    const size_t got = std::fread(t->h_stage, 1, bytes, f);
    const bool extra = std::fgetc(f) != EOF;
    std::fclose(f);
    if (got != bytes || extra) { ctx->set_error("bt_atlas_load_tile_file: %s holds %zu%s bytes, expected %zu", path.c_str(), got, extra ? "+" : "", bytes); return BT_ERR_IO; }
    BT_CUDA(ctx, cudaMemcpyAsync(dst, t->h_stage, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    BT_CUDA(ctx, cudaEventRecord(t->stage_free, (cudaStream_t)stream));
    return BT_OK;
}

bt_status bt_tile_config_load(const char* path, bt_tile* tiles, uint32_t capacity, uint32_t* count) {
    if (!path || !count) return BT_ERR_INVALID_ARG;
    FILE* f = std::fopen(path, "rb");
    if (!f) return BT_ERR_IO;
    std::vector<unsigned char> buf;
    unsigned char chunk[4096];
    size_t n;
    while ((n = std::fread(chunk, 1, sizeof(chunk), f)) > 0) buf.insert(buf.end(), chunk, chunk + n);
    std::fclose(f);
    size_t pos = 0;
    auto varint = [&](uint64_t* out) -> bool {  // bincode 2 "standard" varint
        if (pos >= buf.size()) return false;
        const unsigned char b = buf[pos++];
        if (b < 251) { *out = b; return true; }
        const size_t len = b == 251 ? 2 : b == 252 ? 4 : b == 253 ? 8 : 0;
        if (len == 0 || pos + len > buf.size()) return false;
        uint64_t v = 0;
        for (size_t i = 0; i < len; ++i) v |= (uint64_t)buf[pos + i] << (8 * i);
        pos += len;
        *out = v;
        return true;
    };
    uint64_t total;
    if (!varint(&total) || total > 0xFFFFFFFFull) return BT_ERR_IO;
    for (uint64_t i = 0; i < total; ++i) {
        uint64_t v[4];
        for (int k = 0; k < 4; ++k) if (!varint(&v[k]) || v[k] > 0xFFFFFFFFull) return BT_ERR_IO;
        if (tiles && i < capacity) tiles[i] = bt_tile{(uint32_t)v[0], (uint32_t)v[1], (uint32_t)v[2], (uint32_t)v[3]};
    }
    if (pos != buf.size()) return BT_ERR_IO;
    *count = (uint32_t)total;
    return BT_OK;
}

bt_status bt_atlas_sample(bt_terrain* t, uint32_t attachment, const bt_tile_lookup* lookup, float out[4]) {
    if (!t || !lookup || !out || attachment >= t->desc.attachment_count) return BT_ERR_INVALID_ARG;
    out[0] = out[1] = out[2] = out[3] = 0.0f;
    if (lookup->atlas_index == 0xFFFFFFFFu) return BT_OK;   // AtlasAttachment::sample: "Todo: Handle this better" -> zeros
    bt_context* ctx = t->ctx;
    const bt_attachment_config& a = t->desc.attachments[attachment];
    if (lookup->atlas_index >= t->desc.atlas_size) { ctx->set_error("bt_atlas_sample: atlas_index out of range"); return BT_ERR_INVALID_ARG; }
    const float size = (float)a.texture_size;
    const float scale = (float)(a.texture_size - 2 * a.border_size) / size, offset = (float)a.border_size / size;   // AtlasAttachment::new, tile_atlas.rs:172-193
    // AttachmentData::sample — terrain_data/mod.rs:221-264
    const float ux = (lookup->atlas_uv[0] * scale + offset) * size - 0.5f, uy = (lookup->atlas_uv[1] * scale + offset) * size - 0.5f;
    const float rx = std::fmod(ux, 1.0f), ry = std::fmod(uy, 1.0f);
    const int ix = (int)ux, iy = (int)uy;
    if (ix < 0 || iy < 0 || ix + 1 >= (int)a.texture_size || iy + 1 >= (int)a.texture_size) {
        ctx->set_error("bt_atlas_sample: sample footprint leaves the tile (the reference would index out of bounds)");
        return BT_ERR_INVALID_ARG;
    }
    const size_t px = a.format == BT_FORMAT_R16 ? 2 : 4;
    unsigned char texels[2][8];
    const unsigned char* base = static_cast<const unsigned char*>(t->atlas[attachment]) + t->tile_bytes[attachment] * lookup->atlas_index;
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int y = 0; y < 2; ++y)
        BT_CUDA(ctx, cudaMemcpy(texels[y], base + ((size_t)(iy + y) * a.texture_size + ix) * px, 2 * px, cudaMemcpyDeviceToHost));
    float v[2][2][4];   // [x][y][channel]
    for (int x = 0; x < 2; ++x) for (int y = 0; y < 2; ++y) {
        if (a.format == BT_FORMAT_R16) {
            uint16_t w; std::memcpy(&w, texels[y] + x * 2, 2);
            v[x][y][0] = (float)w / 65535.0f; v[x][y][1] = v[x][y][2] = v[x][y][3] = 0.0f;
        } else {
            for (int c = 0; c < 4; ++c) v[x][y][c] = (float)texels[y][x * 4 + c] / 255.0f;
        }
    }
    for (int c = 0; c < 4; ++c) {  // Vec4::lerp(lerp(v00, v01, ry), lerp(v10, v11, ry), rx), glam: a + (b - a) * s
        const float l0 = v[0][0][c] + (v[0][1][c] - v[0][0][c]) * ry;
        const float l1 = v[1][0][c] + (v[1][1][c] - v[1][0][c]) * ry;
        out[c] = l0 + (l1 - l0) * rx;
    }
    return BT_OK;
}

bt_status bt_atlas_device_ptr(bt_terrain* t, uint32_t attachment, void** out) {
    if (!t || !out || attachment >= t->desc.attachment_count) return BT_ERR_INVALID_ARG;
    *out = t->atlas[attachment];
    return BT_OK;
}

// ------------------------------------------------------------------------------------------------
// view
// ------------------------------------------------------------------------------------------------
static btd::SelectViewParams select_params_of(const bt_view* v) {
    return btd::SelectViewParams{v->d_uniforms, v->d_final_tiles, v->d_temporary_tiles, v->d_indirect, v->d_select_status, v->desc.geometry_tile_count,
                                 v->terrain->desc.atlas_size, v->d_tile_tree, v->terrain->d_minmax,
                                 v->d_host_uniforms, v->d_host_table, v->tile_tree_entries, 0u};
}

bt_status bt_view_create(bt_terrain* t, const bt_view_desc* desc, bt_view** out) {
    if (!t || !desc || !out) return BT_ERR_INVALID_ARG;
    bt_context* ctx = t->ctx;
    if (desc->geometry_tile_count == 0 || desc->tree_size == 0 || desc->grid_size == 0) {
        ctx->set_error("bt_view_create: geometry_tile_count, tree_size and grid_size must be non-zero");
        return BT_ERR_INVALID_ARG;
    }
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    bt_view* v = new bt_view();
    v->terrain = t;
    v->desc = *desc;
    const uint32_t sides = t->desc.kind == BT_KIND_SPHERICAL ? 6u : 1u;
    v->tile_tree_entries = sides * t->desc.lod_count * desc->tree_size * desc->tree_size;
    const size_t cap = desc->geometry_tile_count;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    alloc((void**)&v->d_uniforms, sizeof(ViewUniforms));
    alloc((void**)&v->d_final_tiles, cap * sizeof(bt_tile));          // final_tiles — terrain_view_bind_group.rs:139-142
    alloc((void**)&v->d_temporary_tiles, 2 * cap * sizeof(bt_tile));  // temporary_tiles (two halves instead of a double-ended stack)
    alloc((void**)&v->d_indirect, sizeof(bt_indirect));
    alloc((void**)&v->d_select_status, sizeof(btd::SelectStatus));
    alloc((void**)&v->d_gen_status, sizeof(btd::GenStatus));
    alloc((void**)&v->d_select_params, sizeof(btd::SelectViewParams));
    alloc((void**)&v->d_tile_tree, (size_t)v->tile_tree_entries * sizeof(bt_tile_tree_entry));
    if (e == cudaSuccess) e = cudaMemset(v->d_indirect, 0, sizeof(bt_indirect));
    if (e == cudaSuccess) e = cudaMemset(v->d_select_status, 0, sizeof(btd::SelectStatus));
    if (e == cudaSuccess) e = cudaMemset(v->d_gen_status, 0, sizeof(btd::GenStatus));
    if (e == cudaSuccess) e = cudaMemset(v->d_tile_tree, 0xFF, (size_t)v->tile_tree_entries * sizeof(bt_tile_tree_entry));
    // pinned staging so bt_view_update is a pair of truly asynchronous copies; mapped, so that bt_view_run_frame's select kernel
    // can fetch the frame's inputs itself
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&v->h_uniforms, sizeof(ViewUniforms), cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&v->h_tile_tree, (size_t)v->tile_tree_entries * sizeof(bt_tile_tree_entry), cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer(&v->d_host_uniforms, v->h_uniforms, 0);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer(&v->d_host_table, v->h_tile_tree, 0);
    if (e == cudaSuccess) {
        const btd::SelectViewParams sp = select_params_of(v);
        e = cudaMemcpy(v->d_select_params, &sp, sizeof(sp), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&v->h_readback, sizeof(btd::SelectStatus) + sizeof(bt_indirect) + sizeof(btd::GenStatus), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&v->staging_free, cudaEventDisableTiming);
    // everything bt_view_run_frame needs is created HERE: no allocation on the per-frame path (SURVEY 8b ownership rule)
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&v->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&v->select_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&v->copy_done, cudaEventDisableTiming);
    if (e == cudaSuccess) {
        const uint32_t want = desc->readback_tile_capacity ? desc->readback_tile_capacity : 65536u;
        v->h_tiles_cap = want < desc->geometry_tile_count ? want : desc->geometry_tile_count;
        e = cudaHostAlloc((void**)&v->h_tiles, (size_t)v->h_tiles_cap * sizeof(bt_tile), cudaHostAllocDefault);
    }
    if (e == cudaSuccess && desc->max_vertices) {
        e = cudaMalloc((void**)&v->d_vertices, desc->max_vertices * sizeof(bt_vertex));
        v->vertex_capacity = desc->max_vertices;
    }
    if (e != cudaSuccess) {
        ctx->set_error("bt_view_create: %s", cudaGetErrorString(e));
        bt_view_destroy(v);
        return e == cudaErrorMemoryAllocation ? BT_ERR_OUT_OF_MEMORY : BT_ERR_CUDA;
    }
    *out = v;
    return BT_OK;
}

void bt_view_destroy(bt_view* v) {
    if (!v) return;
    cudaSetDevice(v->terrain->ctx->device);
    cudaDeviceSynchronize();
    cudaFree(v->d_uniforms); cudaFree(v->d_final_tiles); cudaFree(v->d_temporary_tiles); cudaFree(v->d_indirect);
    cudaFree(v->d_select_status); cudaFree(v->d_gen_status); cudaFree(v->d_tile_tree); cudaFree(v->d_vertices); cudaFree(v->d_select_params);
    cudaFree(v->d_albedo);
    if (v->h_uniforms) cudaFreeHost(v->h_uniforms);
    if (v->h_tile_tree) cudaFreeHost(v->h_tile_tree);
    if (v->h_readback) cudaFreeHost(v->h_readback);
    if (v->staging_free) cudaEventDestroy(v->staging_free);
    if (v->select_done) cudaEventDestroy(v->select_done);
    if (v->copy_done) cudaEventDestroy(v->copy_done);
    if (v->copy_stream) cudaStreamDestroy(v->copy_stream);
    if (v->h_tiles) cudaFreeHost(v->h_tiles);
    delete v;
}

// The Prepare-stage payload of a frame: validated and written into the view's pinned staging blocks.  push = true: two stream-ordered
// host-to-device copies follow (bt_view_update); push = false: the caller's select kernel fetches the blocks itself (bt_view_run_frame).
static bt_status view_stage(bt_view* v, const bt_view_frame* f, void* stream, bool push, bool* table_staged) {
    if (!v || !f) return BT_ERR_INVALID_ARG;
    bt_terrain* t = v->terrain;
    bt_context* ctx = t->ctx;
    const bool host_table = f->tile_tree != nullptr;
    if (table_staged) *table_staged = host_table;
    if (!host_table && !v->device_table) {
        ctx->set_error("bt_view_update: tile_tree is NULL and no bt_clipmap maintains this view's table on the device");
        return BT_ERR_INVALID_ARG;
    }
    if (host_table && f->tile_tree_entry_count != v->tile_tree_entries) {
        ctx->set_error("bt_view_update: tile_tree_entry_count %u, expected sides*lod_count*tree_size^2 = %u", f->tile_tree_entry_count, v->tile_tree_entries);
        return BT_ERR_INVALID_ARG;
    }
    if (f->view_config.tree_size != v->desc.tree_size || (uint32_t)f->view_config.grid_size != v->desc.grid_size ||
        f->view_config.vertices_per_row != 2 * (v->desc.grid_size + 2) ||
        f->view_config.vertices_per_tile != 2 * v->desc.grid_size * (v->desc.grid_size + 2)) {
        ctx->set_error("bt_view_update: view_config tree_size/grid_size/vertices_per_* disagree with the view");
        return BT_ERR_INVALID_ARG;
    }
    if (f->approximation.origin_lod > 30) {
        ctx->set_error("bt_view_update: origin_lod %u > 30", f->approximation.origin_lod);
        return BT_ERR_INVALID_ARG;
    }
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    BT_CUDA(ctx, cudaEventSynchronize(v->staging_free));  // previous async copies out of the staging buffers are done
    ViewUniforms& U = *static_cast<ViewUniforms*>(v->h_uniforms);
    std::memset(&U, 0, sizeof(U));
    U.view_config = f->view_config;
    U.spherical = t->desc.kind == BT_KIND_SPHERICAL;
    U.texture_size = t->desc.attachments[0].texture_size;
    U.tile_tree_entries = v->tile_tree_entries;
    U.atlas_size = t->desc.atlas_size;
    U.approximation = f->approximation;
    U.culling = f->culling;
    // TerrainConfigUniform::from_tile_atlas — terrain_bind_group.rs:84-93
    U.terrain_config.lod_count = t->desc.lod_count;
    U.terrain_config.min_height = t->desc.min_height;
    U.terrain_config.max_height = t->desc.max_height;
    U.terrain_config.scale = (float)t->desc.scale;
    for (uint32_t i = 0; i < t->desc.attachment_count; ++i) {  // AttachmentUniform::new — terrain_bind_group.rs:60-72
        const bt_attachment_config& a = t->desc.attachments[i];
        const float center = (float)(a.texture_size - 2 * a.border_size);
        U.attachments[i].size = center;
        U.attachments[i].scale = center / (float)a.texture_size;
        U.attachments[i].offset = (float)a.border_size / (float)a.texture_size;
    }
    U.mesh = t->desc.transform;
    std::memcpy(U.view_world_position, f->view_world_position, sizeof(float) * 3);
    cudaStream_t s = (cudaStream_t)stream;
    if (host_table) std::memcpy(v->h_tile_tree, f->tile_tree, (size_t)v->tile_tree_entries * sizeof(bt_tile_tree_entry));
    v->h2d_bytes_last_update = sizeof(ViewUniforms) + (host_table ? (size_t)v->tile_tree_entries * sizeof(bt_tile_tree_entry) : 0);
    if (push) {
        BT_CUDA(ctx, cudaMemcpyAsync(v->d_uniforms, v->h_uniforms, sizeof(ViewUniforms), cudaMemcpyHostToDevice, s));
        if (host_table)   // (with a device clipmap the table is already where the kernels read it)
            BT_CUDA(ctx, cudaMemcpyAsync(v->d_tile_tree, v->h_tile_tree, (size_t)v->tile_tree_entries * sizeof(bt_tile_tree_entry), cudaMemcpyHostToDevice, s));
        BT_CUDA(ctx, cudaEventRecord(v->staging_free, s));
    }
    v->updated = true;
    return BT_OK;
}

bt_status bt_view_update(bt_view* v, const bt_view_frame* f, void* stream) { return view_stage(v, f, stream, true, nullptr); }

bt_status bt_select_tiles(bt_view* v, uint32_t flags, void* stream) {
    if (!v) return BT_ERR_INVALID_ARG;
    bt_context* ctx = v->terrain->ctx;
    if (!v->updated) { ctx->set_error("bt_select_tiles: bt_view_update has not been called"); return BT_ERR_INVALID_ARG; }
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    BT_CUDA(ctx, bt_launch_select_tiles(v->d_select_params, 1u, v->terrain->desc.kind == BT_KIND_SPHERICAL,
                                        flags & ~(btd::kSelectFetchUniforms | btd::kSelectFetchTable), (cudaStream_t)stream));
    return BT_OK;
}

// ---- the views of one frame in ONE select launch (TilingPrepassNode::run iterates (terrain, view) pairs, tiling_prepass.rs:224-263) ----
bt_status bt_view_batch_create(bt_view* const* views, uint32_t count, bt_view_batch** out) {
    if (!views || !out || count == 0 || !views[0]) return BT_ERR_INVALID_ARG;
    bt_context* ctx = views[0]->terrain->ctx;
    const int spherical = views[0]->terrain->desc.kind == BT_KIND_SPHERICAL;
    std::vector<btd::SelectViewParams> params;
    for (uint32_t i = 0; i < count; ++i) {
        if (!views[i] || views[i]->terrain->ctx != ctx || (views[i]->terrain->desc.kind == BT_KIND_SPHERICAL) != spherical) {
            ctx->set_error("bt_view_batch_create: the views of a batch must share one context and one terrain kind");
            return BT_ERR_INVALID_ARG;
        }
        params.push_back(select_params_of(views[i]));
    }
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    bt_view_batch* b = new bt_view_batch();
    b->ctx = ctx;
    b->spherical = spherical;
    b->views.assign(views, views + count);
    cudaError_t e = cudaMalloc((void**)&b->d_params, params.size() * sizeof(btd::SelectViewParams));
    if (e == cudaSuccess) e = cudaMemcpy(b->d_params, params.data(), params.size() * sizeof(btd::SelectViewParams), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(b->d_params); delete b; ctx->set_error("bt_view_batch_create: %s", cudaGetErrorString(e)); return BT_ERR_CUDA; }
    *out = b;
    return BT_OK;
}

void bt_view_batch_destroy(bt_view_batch* b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaFree(b->d_params);
    delete b;
}

bt_status bt_select_tiles_batch(bt_view_batch* b, uint32_t flags, void* stream) {
    if (!b) return BT_ERR_INVALID_ARG;
    bt_context* ctx = b->ctx;
    for (bt_view* v : b->views)
        if (!v->updated) { ctx->set_error("bt_select_tiles_batch: bt_view_update has not been called on every view"); return BT_ERR_INVALID_ARG; }
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    BT_CUDA(ctx, bt_launch_select_tiles(b->d_params, (uint32_t)b->views.size(), b->spherical,
                                        flags & ~(btd::kSelectFetchUniforms | btd::kSelectFetchTable), (cudaStream_t)stream));
    return BT_OK;
}

bt_status bt_gen_vertices(bt_view* v, uint32_t flags, void* stream, bt_vertex_buffer* out) {
    if (!v) return BT_ERR_INVALID_ARG;
    bt_context* ctx = v->terrain->ctx;
    if (!v->updated) { ctx->set_error("bt_gen_vertices: bt_view_update has not been called"); return BT_ERR_INVALID_ARG; }
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (v->desc.max_vertices == 0) {
        // grow-on-demand mode: needs the tile count on the host (one small blocking readback);
        // views created with max_vertices stay fully asynchronous
        bt_indirect ind;
        BT_CUDA(ctx, cudaMemcpyAsync(&ind, v->d_indirect, sizeof(ind), cudaMemcpyDeviceToHost, s));
        BT_CUDA(ctx, cudaStreamSynchronize(s));
        if ((uint64_t)ind.vertex_count > v->vertex_capacity) {
            if (v->d_vertices) BT_CUDA(ctx, cudaFree(v->d_vertices));
            v->d_vertices = nullptr;
            v->vertex_capacity = 0;
            cudaError_t e = cudaMalloc((void**)&v->d_vertices, (size_t)ind.vertex_count * sizeof(bt_vertex));
            if (e != cudaSuccess) { ctx->set_error("bt_gen_vertices: vertex buffer allocation failed: %s", cudaGetErrorString(e)); return BT_ERR_OUT_OF_MEMORY; }
            v->vertex_capacity = ind.vertex_count;
        }
    }
    btd::GenFilterInputs filter{};
    const btd::GenFilterInputs* filter_ptr = nullptr;
    if (flags & (BT_FLAG_FILTER_TRILINEAR_ANISO | BT_FLAG_ALBEDO)) {
        bt_terrain* t = v->terrain;
        const uint32_t aa = 1u;   // the albedo attachment of the reference's examples (planar.rs:64-70): attachment 1, Rgba8
        if (flags & BT_FLAG_ALBEDO) {
            if (t->desc.attachment_count <= aa || t->desc.attachments[aa].format != BT_FORMAT_RGBA8 ||
                t->desc.attachments[aa].texture_size != t->desc.attachments[0].texture_size) {
                ctx->set_error("bt_gen_vertices: BT_FLAG_ALBEDO needs attachment 1 = Rgba8 with the height attachment's texture size");
                return BT_ERR_INVALID_ARG;
            }
            if (v->albedo_capacity < v->vertex_capacity) {   // one-time growth on the first ALBEDO frame (like the grow-on-demand vertex buffer)
                if (v->d_albedo) BT_CUDA(ctx, cudaFree(v->d_albedo));
                v->d_albedo = nullptr;
                v->albedo_capacity = 0;
                cudaError_t e = cudaMalloc((void**)&v->d_albedo, (size_t)v->vertex_capacity * sizeof(uint32_t));
                if (e != cudaSuccess) { ctx->set_error("bt_gen_vertices: albedo stream allocation failed: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return BT_ERR_OUT_OF_MEMORY; }
                v->albedo_capacity = v->vertex_capacity;
            }
        }
        if ((flags & BT_FLAG_FILTER_TRILINEAR_ANISO) && t->desc.attachments[0].mip_level_count > 1 && (flags & BT_FLAG_ALBEDO) &&
            t->desc.attachments[aa].mip_level_count != t->desc.attachments[0].mip_level_count) {
            ctx->set_error("bt_gen_vertices: the height and albedo attachments must have the same mip_level_count for the filtered sampler");
            return BT_ERR_INVALID_ARG;
        }
        filter.height_mips = static_cast<const uint16_t*>(t->atlas_mips[0]);
        filter.albedo_atlas = (flags & BT_FLAG_ALBEDO) ? t->atlas[aa] : nullptr;
        filter.albedo_mips = (flags & BT_FLAG_ALBEDO) ? t->atlas_mips[aa] : nullptr;
        filter.srgb_lut = t->d_srgb_lut;
        filter.albedo_out = v->d_albedo;
        filter.mip_level_count = t->desc.attachments[0].mip_level_count ? t->desc.attachments[0].mip_level_count : 1u;
        filter.mip_chain_texels = t->mip_chain_texels[0];
        filter.albedo_attachment = aa;
        filter_ptr = &filter;
    }
    BT_CUDA(ctx, bt_launch_gen_vertices(v->d_uniforms, v->d_final_tiles, v->d_indirect, v->d_tile_tree,
                                        static_cast<const uint16_t*>(v->terrain->atlas[0]), v->d_vertices, v->vertex_capacity,
                                        v->d_gen_status, v->terrain->desc.kind == BT_KIND_SPHERICAL, flags, v->desc.grid_size,
                                        ctx->sm_count, filter_ptr, s));
    if (out) { out->vertices = v->d_vertices; out->capacity = v->vertex_capacity; }
    return BT_OK;
}

bt_status bt_view_readback(bt_view* v, bt_tile* tiles, uint32_t capacity, uint32_t* count, bt_indirect* indirect, void* stream) {
    if (!v) return BT_ERR_INVALID_ARG;
    bt_context* ctx = v->terrain->ctx;
    cudaStream_t s = (cudaStream_t)stream;
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned char* hb = static_cast<unsigned char*>(v->h_readback);
    BT_CUDA(ctx, cudaMemcpyAsync(hb, v->d_select_status, sizeof(btd::SelectStatus), cudaMemcpyDeviceToHost, s));
    BT_CUDA(ctx, cudaMemcpyAsync(hb + sizeof(btd::SelectStatus), v->d_indirect, sizeof(bt_indirect), cudaMemcpyDeviceToHost, s));
    BT_CUDA(ctx, cudaStreamSynchronize(s));
    btd::SelectStatus st;
    std::memcpy(&st, hb, sizeof(st));
    if (indirect) std::memcpy(indirect, hb + sizeof(btd::SelectStatus), sizeof(bt_indirect));
    if (count) *count = st.final_count;
    if (tiles) {
        const uint32_t n = st.final_count < capacity ? st.final_count : capacity;
        if (n) {
            BT_CUDA(ctx, cudaMemcpyAsync(tiles, v->d_final_tiles, (size_t)n * sizeof(bt_tile), cudaMemcpyDeviceToHost, s));
            BT_CUDA(ctx, cudaStreamSynchronize(s));
        }
    }
    if (st.overflow) { ctx->set_error("select_tiles: tile list exceeded geometry_tile_count = %u", v->desc.geometry_tile_count); return BT_ERR_CAPACITY; }
    return BT_OK;
}

// One frame of one (terrain, view) pair in a single call: Prepare-stage uploads, TilingPrepassNode::run, the vertex stage,
// and the host read-back of the tile list + indirect record.  The read-back is issued on an internal copy stream right after
// select_tiles and overlaps gen_vertices; the call returns when the tile list is on the host while gen_vertices may still be
// running on `stream` (stream-ordered: the next frame's work queues behind it).  The vertex buffer stays on the device.
bt_status bt_view_run_frame(bt_view* v, const bt_view_frame* frame, uint32_t select_flags, uint32_t gen_flags, void* stream,
                        bt_tile* tiles, uint32_t capacity, uint32_t* count, bt_indirect* indirect, bt_vertex_buffer* vertices) {
    if (!v || !frame) return BT_ERR_INVALID_ARG;
    bt_context* ctx = v->terrain->ctx;
    if (v->desc.max_vertices == 0) { ctx->set_error("bt_view_run_frame: needs a view created with max_vertices (no blocking growth on the frame path)"); return BT_ERR_INVALID_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    // the pinned read-back buffer was sized by bt_view_create (bt_view_desc.readback_tile_capacity): nothing is allocated here
    const uint32_t want = tiles ? (capacity < v->h_tiles_cap ? capacity : v->h_tiles_cap) : 0u;
    // The frame's inputs go into the pinned staging blocks and the select kernel - the first kernel of the frame - pulls them over PCIe
    // itself (BT_FRAME_FETCH=0 falls back to two copy-engine operations in front of it, the bt_view_update path): same bytes, but no
    // engine hand-overs on the critical path.
    static const bool kernel_fetch = [] { const char* e = getenv("BT_FRAME_FETCH"); return !(e && e[0] == '0'); }();
    bool table_staged = false;
    bt_status st = view_stage(v, frame, stream, !kernel_fetch, &table_staged);
    if (st != BT_OK) return st;
    if (kernel_fetch) {
        const uint32_t fetch = btd::kSelectFetchUniforms | (table_staged ? btd::kSelectFetchTable : 0u);
        BT_CUDA(ctx, bt_launch_select_tiles(v->d_select_params, 1u, v->terrain->desc.kind == BT_KIND_SPHERICAL,
                                            (select_flags & ~(btd::kSelectFetchUniforms | btd::kSelectFetchTable)) | fetch, s));
        BT_CUDA(ctx, cudaEventRecord(v->staging_free, s));   // the staging blocks may be rewritten once this kernel has run
    } else {
        st = bt_select_tiles(v, select_flags, stream);
        if (st != BT_OK) return st;
    }
    BT_CUDA(ctx, cudaEventRecord(v->select_done, s));
    BT_CUDA(ctx, cudaStreamWaitEvent(v->copy_stream, v->select_done, 0));
    unsigned char* hb = static_cast<unsigned char*>(v->h_readback);
    BT_CUDA(ctx, cudaMemcpyAsync(hb, v->d_select_status, sizeof(btd::SelectStatus), cudaMemcpyDeviceToHost, v->copy_stream));
    BT_CUDA(ctx, cudaMemcpyAsync(hb + sizeof(btd::SelectStatus), v->d_indirect, sizeof(bt_indirect), cudaMemcpyDeviceToHost, v->copy_stream));
    if (want) BT_CUDA(ctx, cudaMemcpyAsync(v->h_tiles, v->d_final_tiles, (size_t)want * sizeof(bt_tile), cudaMemcpyDeviceToHost, v->copy_stream));
    BT_CUDA(ctx, cudaEventRecord(v->copy_done, v->copy_stream));
    st = bt_gen_vertices(v, gen_flags, stream, vertices);
    if (st != BT_OK) return st;
    // the next select on `stream` must not overwrite final_tiles / status while this copy is in flight
    BT_CUDA(ctx, cudaStreamWaitEvent(s, v->copy_done, 0));
    BT_CUDA(ctx, cudaEventSynchronize(v->copy_done));
    btd::SelectStatus sst;
    std::memcpy(&sst, hb, sizeof(sst));
    if (indirect) std::memcpy(indirect, hb + sizeof(btd::SelectStatus), sizeof(bt_indirect));
    if (count) *count = sst.final_count;
    if (tiles) std::memcpy(tiles, v->h_tiles, (size_t)(sst.final_count < want ? sst.final_count : want) * sizeof(bt_tile));
    if (sst.overflow) { ctx->set_error("select_tiles: tile list exceeded geometry_tile_count = %u", v->desc.geometry_tile_count); return BT_ERR_CAPACITY; }
    return BT_OK;
}

bt_status bt_view_select_status(bt_view* v, uint32_t* dropped_parents, void* stream) {
    if (!v) return BT_ERR_INVALID_ARG;
    bt_context* ctx = v->terrain->ctx;
    cudaStream_t s = (cudaStream_t)stream;
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    btd::SelectStatus st;
    BT_CUDA(ctx, cudaMemcpyAsync(&st, v->d_select_status, sizeof(st), cudaMemcpyDeviceToHost, s));
    BT_CUDA(ctx, cudaStreamSynchronize(s));
    if (dropped_parents) *dropped_parents = st.dropped_parents;
    return st.overflow ? BT_ERR_CAPACITY : BT_OK;
}

bt_status bt_view_download_vertices(bt_view* v, bt_vertex* out, uint64_t first, uint64_t count, void* stream) {
    if (!v || !out) return BT_ERR_INVALID_ARG;
    bt_context* ctx = v->terrain->ctx;
    if (first + count > v->vertex_capacity) { ctx->set_error("bt_view_download_vertices: range exceeds the vertex buffer"); return BT_ERR_INVALID_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    btd::GenStatus gs;
    BT_CUDA(ctx, cudaMemcpyAsync(&gs, v->d_gen_status, sizeof(gs), cudaMemcpyDeviceToHost, s));
    if (count) BT_CUDA(ctx, cudaMemcpyAsync(out, v->d_vertices + first, count * sizeof(bt_vertex), cudaMemcpyDeviceToHost, s));
    BT_CUDA(ctx, cudaStreamSynchronize(s));
    if (gs.overflow) { ctx->set_error("gen_vertices: vertex buffer (%llu vertices) too small", (unsigned long long)v->vertex_capacity); return BT_ERR_CAPACITY; }
    return BT_OK;
}

// the per-vertex albedo stream of the last bt_gen_vertices with BT_FLAG_ALBEDO (device pointer, strip order, pack4x8unorm)
bt_status bt_view_albedo_buffer(bt_view* v, uint32_t** device_albedo, uint64_t* capacity) {
    if (!v || !device_albedo) return BT_ERR_INVALID_ARG;
    *device_albedo = v->d_albedo;
    if (capacity) *capacity = v->albedo_capacity;
    return BT_OK;
}

bt_status bt_view_download_albedo(bt_view* v, uint32_t* out, uint64_t first, uint64_t count, void* stream) {
    if (!v || !out) return BT_ERR_INVALID_ARG;
    bt_context* ctx = v->terrain->ctx;
    if (!v->d_albedo || first + count > v->albedo_capacity) { ctx->set_error("bt_view_download_albedo: no albedo stream, or the range exceeds it"); return BT_ERR_INVALID_ARG; }
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    if (count) BT_CUDA(ctx, cudaMemcpyAsync(out, v->d_albedo + first, count * sizeof(uint32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    BT_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    return BT_OK;
}

bt_status bt_view_device_ptrs(bt_view* v, bt_tile** final_tiles, bt_indirect** indirect) {
    if (!v) return BT_ERR_INVALID_ARG;
    if (final_tiles) *final_tiles = v->d_final_tiles;
    if (indirect) *indirect = v->d_indirect;
    return BT_OK;
}

// ------------------------------------------------------------------------------------------------
// utilities
// ------------------------------------------------------------------------------------------------
// planes() — src/render/culling_bind_group.rs:25-38 (Gribb-Hartmann rows of clip_from_world, column-major input)
void bt_culling_planes(const float clip_from_world[16], float planes[5][4]) {
    auto row = [&](int r, int c) { return clip_from_world[c * 4 + r]; };
    for (int i = 0; i < 5; ++i) {
        const int r = i / 2;
        const bool add = (i & 1) == 0 && i != 4;
        for (int c = 0; c < 4; ++c) planes[i][c] = add ? row(3, c) + row(r, c) : row(3, c) - row(r, c);
    }
}

bt_status bt_generate_mipmaps(bt_context* ctx, uint32_t format, uint32_t texture_size, uint32_t mip_level_count,
                              const void* device_mip0, void* device_mips_out, void* stream) {
    if (!ctx || !device_mip0 || !device_mips_out || (format != BT_FORMAT_R16 && format != BT_FORMAT_RGBA8)) return BT_ERR_INVALID_ARG;
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    BT_CUDA(ctx, bt_launch_mipmaps(format, texture_size, mip_level_count, device_mip0, device_mips_out, (cudaStream_t)stream));
    return BT_OK;
}

bt_status bt_synth_source(bt_context* ctx, uint32_t format, uint32_t width, uint32_t height, uint32_t x0, uint32_t y0,
                          uint32_t seed, uint32_t holes, void* device_out, void* stream) {
    if (!ctx || !device_out || (format != BT_FORMAT_R16 && format != BT_FORMAT_RGBA8)) return BT_ERR_INVALID_ARG;
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    BT_CUDA(ctx, bt_launch_synth(format, width, height, x0, y0, seed, holes, device_out, (cudaStream_t)stream));
    return BT_OK;
}

bt_status bt_selftest_numerics(bt_context* ctx, const float* x, float* log2_out, uint32_t n, uint32_t mismatches[4]) {
    if (!ctx || !mismatches || (n && (!x || !log2_out))) return BT_ERR_INVALID_ARG;
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    float *dx = nullptr, *dy = nullptr;
    uint32_t* dm = nullptr;
    BT_CUDA(ctx, cudaMalloc((void**)&dx, (size_t)(n ? n : 1) * sizeof(float)));
    BT_CUDA(ctx, cudaMalloc((void**)&dy, (size_t)(n ? n : 1) * sizeof(float)));
    BT_CUDA(ctx, cudaMalloc((void**)&dm, 4 * sizeof(uint32_t)));
    BT_CUDA(ctx, cudaMemset(dm, 0, 4 * sizeof(uint32_t)));
    if (n) BT_CUDA(ctx, cudaMemcpy(dx, x, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    BT_CUDA(ctx, bt_launch_selftest(dx, dy, n, dm, nullptr));
    if (n) BT_CUDA(ctx, cudaMemcpy(log2_out, dy, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    BT_CUDA(ctx, cudaMemcpy(mismatches, dm, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dy); cudaFree(dm);
    return BT_OK;
}

bt_status bt_device_alloc(bt_context* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return BT_ERR_INVALID_ARG;
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) { ctx->set_error("bt_device_alloc(%zu): %s", bytes, cudaGetErrorString(e)); return BT_ERR_OUT_OF_MEMORY; }
    return BT_OK;
}

bt_status bt_device_free(bt_context* ctx, void* ptr) {
    if (!ctx) return BT_ERR_INVALID_ARG;
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    BT_CUDA(ctx, cudaFree(ptr));
    return BT_OK;
}

bt_status bt_memcpy_h2d(bt_context* ctx, void* dst_device, const void* src_host, size_t bytes, void* stream) {
    if (!ctx || (!dst_device && bytes) || (!src_host && bytes)) return BT_ERR_INVALID_ARG;
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    BT_CUDA(ctx, cudaMemcpyAsync(dst_device, src_host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    BT_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    return BT_OK;
}

bt_status bt_memcpy_d2h(bt_context* ctx, void* dst_host, const void* src_device, size_t bytes, void* stream) {
    if (!ctx || (!dst_host && bytes) || (!src_device && bytes)) return BT_ERR_INVALID_ARG;
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    BT_CUDA(ctx, cudaMemcpyAsync(dst_host, src_device, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    BT_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    return BT_OK;
}

}  // extern "C"
