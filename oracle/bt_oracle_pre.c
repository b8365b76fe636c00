/*
 * bt_oracle_pre.c — CPU oracle: tile-pyramid preprocessor, mip generation, synthetic sources.
 * TEST INFRASTRUCTURE ONLY (see bt_oracle.h).  PARITY UNPINNED (no reference tests exist).
 *
 * Restates src/shaders/preprocess/{preprocessing,split,downsample,stitch}.wgsl and the task
 * graph of src/preprocess/preprocessor.rs:58-66,234-343.  Numerics contract: bt_oracle.c header.
 */
#include "bt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ============================================================================================ */
/* tile coordinate helpers — src/math/coordinate.rs                                              */
/* ============================================================================================ */

static const bt_tile TILE_INVALID = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};

/* NEIGHBOURING_SIDES — coordinate.rs:9-16 */
static const uint32_t NEIGHBOURING_SIDES[6][5] = {
    {0, 4, 2, 1, 5}, {1, 0, 2, 3, 5}, {2, 0, 4, 3, 1}, {3, 2, 4, 5, 1}, {4, 2, 0, 5, 3}, {5, 4, 0, 1, 3}};

enum { SI_FIXED0, SI_FIXED1, SI_POS_S, SI_POS_T };
/* SideInfo::EVEN_LIST / ODD_LIST — coordinate.rs:27-42 */
static const uint8_t SI_EVEN[6][2] = {{SI_POS_S, SI_POS_T}, {SI_FIXED0, SI_POS_T}, {SI_FIXED0, SI_POS_S},
                                      {SI_POS_T, SI_POS_S}, {SI_POS_T, SI_FIXED0}, {SI_POS_S, SI_FIXED0}};
static const uint8_t SI_ODD[6][2] = {{SI_POS_S, SI_POS_T}, {SI_POS_S, SI_FIXED1}, {SI_POS_T, SI_FIXED1},
                                     {SI_POS_T, SI_POS_S}, {SI_FIXED1, SI_POS_S}, {SI_FIXED1, SI_POS_T}};

static int tile_is_invalid(const bt_tile* t) { return t->side == 0xFFFFFFFFu && t->lod == 0xFFFFFFFFu && t->x == 0xFFFFFFFFu && t->y == 0xFFFFFFFFu; }

/* TileCoordinate::neighbour_coordinate — coordinate.rs:227-279 */
static bt_tile neighbour_coordinate(const bt_tile* self, int32_t nx, int32_t ny, int spherical) {
    int32_t tile_count = (int32_t)(1u << self->lod);
    if (spherical) {
        int edge_index;
        if ((nx < 0 && ny < 0) || (nx < 0 && ny >= tile_count) || (nx >= tile_count && ny < 0) || (nx >= tile_count && ny >= tile_count))
            return TILE_INVALID;
        else if (nx < 0) edge_index = 1;
        else if (ny < 0) edge_index = 2;
        else if (nx >= tile_count) edge_index = 3;
        else if (ny >= tile_count) edge_index = 4;
        else edge_index = 0;
        uint32_t px = (uint32_t)(nx < 0 ? 0 : (nx > tile_count - 1 ? tile_count - 1 : nx));
        uint32_t py = (uint32_t)(ny < 0 ? 0 : (ny > tile_count - 1 ? tile_count - 1 : ny));
        uint32_t neighbour_side = NEIGHBOURING_SIDES[self->side][edge_index];
        uint32_t index = (6 + neighbour_side - self->side) % 6;
        const uint8_t* info = (self->side % 2 == 0) ? SI_EVEN[index] : SI_ODD[index];
        uint32_t xy[2];
        for (int k = 0; k < 2; ++k) {
            switch (info[k]) {
                case SI_FIXED0: xy[k] = 0; break;
                case SI_FIXED1: xy[k] = (uint32_t)tile_count - 1; break;
                case SI_POS_S: xy[k] = px; break;
                default: xy[k] = py; break;
            }
        }
        bt_tile out = {neighbour_side, self->lod, xy[0], xy[1]};
        return out;
    }
    if (nx < 0 || ny < 0 || nx >= tile_count || ny >= tile_count) return TILE_INVALID;
    bt_tile out = {self->side, self->lod, (uint32_t)nx, (uint32_t)ny};
    return out;
}

/* TileCoordinate::neighbours — coordinate.rs:208-225 */
void bto_tile_neighbours(const bt_tile* tile, int spherical, bt_tile out[8]) {
    static const int32_t OFFSETS[8][2] = {{0, -1}, {1, 0}, {0, 1}, {-1, 0}, {-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
    for (int i = 0; i < 8; ++i)
        out[i] = neighbour_coordinate(tile, (int32_t)tile->x + OFFSETS[i][0], (int32_t)tile->y + OFFSETS[i][1], spherical);
}

/* ============================================================================================ */
/* pyramid store: hash map TileCoordinate -> texels                                             */
/* ============================================================================================ */

typedef struct slot { bt_tile key; void* data; uint32_t exists_from_job; int used; } slot;

typedef struct job { bt_dataset* datasets; uint32_t count; } job;

struct bto_pyramid {
    uint32_t kind, format, T, B, C, lod_count, pixel_size;
    slot* table; size_t table_cap; size_t n_tiles;
    bt_tile* order; /* insertion order */
    job* jobs; uint32_t n_jobs;
};

static uint64_t tile_hash(const bt_tile* t) {
    uint64_t h = 0xcbf29ce484222325ull;
    const uint32_t w[4] = {t->side, t->lod, t->x, t->y};
    for (int i = 0; i < 4; ++i) { h ^= w[i]; h *= 0x100000001b3ull; h ^= h >> 29; }
    return h;
}

static slot* table_find(const bto_pyramid* p, const bt_tile* t) {
    size_t mask = p->table_cap - 1;
    size_t i = (size_t)tile_hash(t) & mask;
    for (;;) {
        slot* s = &p->table[i];
        if (!s->used) return NULL;
        if (s->key.side == t->side && s->key.lod == t->lod && s->key.x == t->x && s->key.y == t->y) return s;
        i = (i + 1) & mask;
    }
}

static void table_grow(bto_pyramid* p) {
    size_t old_cap = p->table_cap; slot* old = p->table;
    p->table_cap = old_cap ? old_cap * 2 : 1024;
    p->table = (slot*)calloc(p->table_cap, sizeof(slot));
    size_t mask = p->table_cap - 1;
    for (size_t j = 0; j < old_cap; ++j) if (old[j].used) {
        size_t i = (size_t)tile_hash(&old[j].key) & mask;
        while (p->table[i].used) i = (i + 1) & mask;
        p->table[i] = old[j];
    }
    free(old);
}

/* TileAtlasState::get_or_allocate_tile — tile_atlas.rs:391-416 (atlas textures are zero-initialised) */
static slot* get_or_allocate(bto_pyramid* p, const bt_tile* t, uint32_t job_index) {
    slot* s = table_find(p, t);
    if (s) return s;
    if ((p->n_tiles + 1) * 2 > p->table_cap) table_grow(p);
    size_t mask = p->table_cap - 1;
    size_t i = (size_t)tile_hash(t) & mask;
    while (p->table[i].used) i = (i + 1) & mask;
    p->table[i].used = 1; p->table[i].key = *t; p->table[i].exists_from_job = job_index;
    p->table[i].data = calloc((size_t)p->T * p->T, p->pixel_size);
    p->order = (bt_tile*)realloc(p->order, (p->n_tiles + 1) * sizeof(bt_tile));
    p->order[p->n_tiles++] = *t;
    return &p->table[i];
}

/* TileAtlasState::get_tile — tile_atlas.rs:369-381; `job_index` = queue-build time of the task */
static void* get_tile(const bto_pyramid* p, const bt_tile* t, uint32_t job_index) {
    if (tile_is_invalid(t)) return NULL;
    slot* s = table_find(p, t);
    if (!s || s->exists_from_job > job_index) return NULL;
    return s->data;
}

bto_pyramid* bto_pyramid_create(uint32_t kind, uint32_t format, uint32_t texture_size, uint32_t border_size, uint32_t lod_count) {
    bto_pyramid* p = (bto_pyramid*)calloc(1, sizeof(bto_pyramid));
    p->kind = kind; p->format = format; p->T = texture_size; p->B = border_size;
    p->C = texture_size - 2 * border_size; p->lod_count = lod_count;
    p->pixel_size = (format == BT_FORMAT_R16) ? 2 : 4;
    table_grow(p);
    return p;
}

void bto_pyramid_destroy(bto_pyramid* p) {
    if (!p) return;
    for (size_t i = 0; i < p->table_cap; ++i) if (p->table[i].used) free(p->table[i].data);
    for (uint32_t j = 0; j < p->n_jobs; ++j) free(p->jobs[j].datasets);
    free(p->jobs); free(p->table); free(p->order); free(p);
}

/* n = 1: Preprocessor::preprocess_tile (preprocessor.rs:298-312);
 * n > 1: Preprocessor::preprocess_spherical ordering (:314-343) — all splits/downsamples first,
 * then per LOD, per dataset, stitch.                                                          */
void bto_pyramid_add_datasets(bto_pyramid* p, const bt_dataset* datasets, uint32_t n) {
    p->jobs = (job*)realloc(p->jobs, (p->n_jobs + 1) * sizeof(job));
    job* j = &p->jobs[p->n_jobs++];
    j->count = n;
    j->datasets = (bt_dataset*)malloc(n * sizeof(bt_dataset));
    memcpy(j->datasets, datasets, n * sizeof(bt_dataset));
}

void bto_pyramid_add_dataset(bto_pyramid* p, const bt_dataset* dataset) { bto_pyramid_add_datasets(p, dataset, 1); }

uint32_t bto_pyramid_tile_count(const bto_pyramid* p) { return (uint32_t)p->n_tiles; }
uint32_t bto_pyramid_list_tiles(const bto_pyramid* p, bt_tile* out, uint32_t cap) {
    uint32_t n = (uint32_t)p->n_tiles;
    for (uint32_t i = 0; i < n && i < cap; ++i) out[i] = p->order[i];
    return n;
}
const void* bto_pyramid_tile(const bto_pyramid* p, const bt_tile* tile) {
    slot* s = table_find(p, tile);
    return s ? s->data : NULL;
}

/* PreprocessDataset::overlapping_tiles — preprocessor.rs:58-66 (f32 arithmetic) */
static void overlapping_range(const bt_dataset* d, uint32_t lod, uint32_t lower[2], uint32_t upper[2]) {
    float tile_count = (float)(1u << lod);
    for (int k = 0; k < 2; ++k) {
        lower[k] = bto_f32_to_u32(d->top_left[k] * tile_count);
        upper[k] = bto_f32_to_u32(ceilf(d->bottom_right[k] * tile_count));
    }
}

/* ============================================================================================ */
/* pixel helpers — preprocessing.wgsl                                                           */
/* ============================================================================================ */

typedef struct px4 { float v[4]; } px4;

static inline px4 load_px(const bto_pyramid* p, const void* tile, uint32_t x, uint32_t y) {
    px4 r;
    if (!tile) { r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0.0f; return r; } /* invalid layer: defined as zero */
    if (p->format == BT_FORMAT_R16) {
        uint16_t t = ((const uint16_t*)tile)[(size_t)y * p->T + x];
        r.v[0] = (float)t / 65535.0f; r.v[1] = 0.0f; r.v[2] = 0.0f; r.v[3] = 1.0f;
    } else {
        const uint8_t* t = (const uint8_t*)tile + ((size_t)y * p->T + x) * 4;
        for (int k = 0; k < 4; ++k) r.v[k] = (float)t[k] / 255.0f;
    }
    return r;
}

/* pack2x16unorm / pack4x8unorm — preprocessing.wgsl:73-90 */
static inline uint32_t pack_unorm(float v, float maxv) {
    float c = fminf(fmaxf(v, 0.0f), 1.0f);
    if (v != v) c = 0.0f; /* NaN (downsample 0/0, Q3) is defined as 0 */
    return (uint32_t)floorf(0.5f + maxv * c);
}

static inline void store_px(const bto_pyramid* p, void* tile, uint32_t x, uint32_t y, const px4* v) {
    if (p->format == BT_FORMAT_R16) {
        ((uint16_t*)tile)[(size_t)y * p->T + x] = (uint16_t)pack_unorm(v->v[0], 65535.0f);
    } else {
        uint8_t* t = (uint8_t*)tile + ((size_t)y * p->T + x) * 4;
        for (int k = 0; k < 4; ++k) t[k] = (uint8_t)pack_unorm(v->v[k], 255.0f);
    }
}

/* is_border — preprocessing.wgsl:55-57 */
static inline int is_border(const bto_pyramid* p, uint32_t x, uint32_t y) {
    return !(x >= p->B && x < p->B + p->C && y >= p->B && y < p->B + p->C);
}

static inline float o_mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }

/* ---- split — split.wgsl:18-43 ---- */
static inline px4 source_texel(const bt_dataset* d, uint32_t format, int32_t x, int32_t y) {
    int32_t mx = (int32_t)d->width - 1, my = (int32_t)d->height - 1;
    x = x < 0 ? 0 : (x > mx ? mx : x);
    y = y < 0 ? 0 : (y > my ? my : y);
    px4 r;
    if (format == BT_FORMAT_R16) {
        uint16_t t = ((const uint16_t*)d->pixels)[(size_t)y * d->width + (size_t)x];
        r.v[0] = (float)t / 65535.0f; r.v[1] = 0.0f; r.v[2] = 0.0f; r.v[3] = 1.0f;
    } else {
        const uint8_t* t = (const uint8_t*)d->pixels + ((size_t)y * d->width + (size_t)x) * 4;
        for (int k = 0; k < 4; ++k) r.v[k] = (float)t[k] / 255.0f;
    }
    return r;
}

static void split_tile(bto_pyramid* p, const bt_dataset* d, const bt_tile* tile, void* data) {
    float tile_scale = (float)(1u << tile->lod);
    for (uint32_t y = 0; y < p->T; ++y) for (uint32_t x = 0; x < p->T; ++x) {
        px4 out;
        if (is_border(p, x, y)) {
            out.v[0] = out.v[1] = out.v[2] = out.v[3] = 0.0f;
            store_px(p, data, x, y, &out);
            continue;
        }
        float tcx = (float)(x - p->B) / (float)p->C;
        float tcy = (float)(y - p->B) / (float)p->C;
        float sx = ((float)tile->x + tcx) / tile_scale;
        float sy = ((float)tile->y + tcy) / tile_scale;
        sx = (sx - d->top_left[0]) / (d->bottom_right[0] - d->top_left[0]);
        sy = (sy - d->top_left[1]) / (d->bottom_right[1] - d->top_left[1]);
        /* textureSampleLevel(linear, clamp-to-edge), level 0 */
        float px = sx * (float)d->width - 0.5f, py = sy * (float)d->height - 0.5f;
        float fx = floorf(px), fy = floorf(py);
        float rx = px - fx, ry = py - fy;
        int32_t ix = (int32_t)fx, iy = (int32_t)fy;
        px4 v00 = source_texel(d, p->format, ix, iy), v10 = source_texel(d, p->format, ix + 1, iy);
        px4 v01 = source_texel(d, p->format, ix, iy + 1), v11 = source_texel(d, p->format, ix + 1, iy + 1);
        px4 value;
        for (int k = 0; k < 4; ++k) value.v[k] = o_mix(o_mix(v00.v[k], v10.v[k], rx), o_mix(v01.v[k], v11.v[k], rx), ry);
        /* textureGather(0, ...) = the same four texels' red channel */
        int is_valid = v00.v[0] != 0.0f && v10.v[0] != 0.0f && v01.v[0] != 0.0f && v11.v[0] != 0.0f;
        /* inside_square(tile_coords, 0, 1) — functions.wgsl:158-162 */
        int is_inside = (tcx >= 0.0f && tcx <= 1.0f) && (tcy >= 0.0f && tcy <= 1.0f);
        if (is_valid && is_inside) store_px(p, data, x, y, &value);
        /* else: textureLoad(atlas, coords) re-packed = existing texel unchanged */
    }
}

/* ---- downsample — downsample.wgsl:12-40 ---- */
static void downsample_tile(bto_pyramid* p, void* data, void* const child[4]) {
    static const uint32_t OFFSETS[4][2] = {{0, 0}, {0, 1}, {1, 0}, {1, 1}};
    uint32_t child_size = p->C / 2u;
    for (uint32_t y = 0; y < p->T; ++y) for (uint32_t x = 0; x < p->T; ++x) {
        px4 out;
        out.v[0] = out.v[1] = out.v[2] = out.v[3] = 0.0f;
        if (is_border(p, x, y)) { store_px(p, data, x, y, &out); continue; }
        uint32_t tx = x - p->B, ty = y - p->B;
        uint32_t cx = 2u * (tx % child_size) + p->B, cy = 2u * (ty % child_size) + p->B;
        uint32_t child_index = tx / child_size + 2u * (ty / child_size);
        /* C odd (not the case for T=512,B=2) would index child 2/3 out of range; clamp like an array */
        if (child_index > 3u) child_index = 3u;
        float count = 0.0f;
        for (int i = 0; i < 4; ++i) {
            uint32_t lx = cx + OFFSETS[i][0], ly = cy + OFFSETS[i][1];
            px4 cv;
            if (lx >= p->T || ly >= p->T) { cv.v[0] = cv.v[1] = cv.v[2] = cv.v[3] = 0.0f; }
            else cv = load_px(p, child[child_index], lx, ly);
            int is_valid = cv.v[0] != 0.0f || cv.v[1] != 0.0f || cv.v[2] != 0.0f;
            if (is_valid) { for (int k = 0; k < 4; ++k) out.v[k] += cv.v[k]; count += 1.0f; }
        }
        if (count == 0.0f) { out.v[0] = out.v[1] = out.v[2] = out.v[3] = 0.0f; } /* Q3: 0/0 := 0 */
        else for (int k = 0; k < 4; ++k) out.v[k] = out.v[k] / count;
        store_px(p, data, x, y, &out);
    }
}

/* ---- stitch — stitch.wgsl:12-118 ---- */
static void stitch_project_to_side(const bto_pyramid* p, uint32_t cx, uint32_t cy, uint32_t original_side, uint32_t projected_side, uint32_t out[2]) {
    enum { PS, PT, NS, NT };
    static const uint8_t EVEN_LIST[6][2] = {{PS, PT}, {PS, PT}, {NT, PS}, {NT, NS}, {PT, NS}, {PS, PT}};
    static const uint8_t ODD_LIST[6][2] = {{PS, PT}, {PS, PT}, {PT, NS}, {PT, PS}, {NT, PS}, {PS, PT}};
    uint32_t index = (6u + projected_side - original_side) % 6u;
    const uint8_t* info = (original_side % 2u == 0u) ? EVEN_LIST[index] : ODD_LIST[index];
    for (int k = 0; k < 2; ++k) {
        switch (info[k]) {
            case PS: out[k] = cx; break;
            case PT: out[k] = cy; break;
            case NS: out[k] = p->T - 1u - cx; break;
            default: out[k] = p->T - 1u - cy; break;
        }
    }
}

static inline int inside_bounds(uint32_t x, uint32_t y, const uint32_t b[4]) {
    return x >= b[0] && x < b[0] + b[2] && y >= b[1] && y < b[1] + b[3];
}

static void stitch_tile(bto_pyramid* p, const bt_tile* tile, void* data, const bt_tile nb_tiles[8], void* const nb[8]) {
    uint32_t B = p->B, C = p->C, O = p->B + p->C;
    const uint32_t bounds[8][4] = {{B, 0, C, B}, {O, B, B, C}, {B, O, C, B}, {0, B, B, C},
                                   {0, 0, B, B}, {O, 0, B, B}, {O, O, B, B}, {0, O, B, B}};
    int32_t c = (int32_t)C;
    const int32_t offsets[8][2] = {{0, c}, {-c, 0}, {0, -c}, {c, 0}, {c, c}, {-c, c}, {-c, -c}, {c, -c}};
    size_t px_bytes = p->pixel_size;
    /* the centre is passed through unchanged; borders read only neighbour CENTRES, so the tile
     * can be stitched in place without ordering hazards between stitch tasks                  */
    for (uint32_t y = 0; y < p->T; ++y) for (uint32_t x = 0; x < p->T; ++x) {
        if (!is_border(p, x, y)) continue;
        uint32_t ni = 0;
        for (uint32_t k = 0; k < 8; ++k) if (inside_bounds(x, y, bounds[k])) { ni = k; break; }
        const uint8_t* src; uint32_t sx, sy;
        if (!nb[ni]) { /* repeat_data — stitch.wgsl:98-103 */
            sx = x < B ? B : (x > B + C - 1u ? B + C - 1u : x);
            sy = y < B ? B : (y > B + C - 1u ? B + C - 1u : y);
            src = (const uint8_t*)data;
        } else { /* neighbour_data — stitch.wgsl:76-96 */
            uint32_t qx = (uint32_t)((int32_t)x + offsets[ni][0]);
            uint32_t qy = (uint32_t)((int32_t)y + offsets[ni][1]);
            uint32_t o[2];
            stitch_project_to_side(p, qx, qy, tile->side, nb_tiles[ni].side, o);
            sx = o[0]; sy = o[1];
            src = (const uint8_t*)nb[ni];
        }
        uint8_t* dst = (uint8_t*)data + ((size_t)y * p->T + x) * px_bytes;
        if (sx >= p->T || sy >= p->T) memset(dst, 0, px_bytes); /* out-of-range textureLoad := 0 */
        else memcpy(dst, src + ((size_t)sy * p->T + sx) * px_bytes, px_bytes);
    }
}

/* ============================================================================================ */
/* task graph — preprocessor.rs:234-343                                                         */
/* ============================================================================================ */

typedef struct task { bt_tile tile; void* data; void* dep[8]; bt_tile dep_tiles[8]; const bt_dataset* ds; } task;

static task* collect(bto_pyramid* p, const bt_dataset* d, uint32_t lod, uint32_t job_index, size_t* n_out) {
    uint32_t lo[2], up[2];
    overlapping_range(d, lod, lo, up);
    size_t n = (up[0] > lo[0] && up[1] > lo[1]) ? (size_t)(up[0] - lo[0]) * (up[1] - lo[1]) : 0;
    task* t = (task*)calloc(n ? n : 1, sizeof(task));
    size_t i = 0;
    /* iproduct!(lower.x..upper.x, lower.y..upper.y): x outer, y inner */
    for (uint32_t x = lo[0]; x < up[0]; ++x) for (uint32_t y = lo[1]; y < up[1]; ++y) {
        bt_tile tc = {d->side, lod, x, y};
        t[i].tile = tc; t[i].ds = d;
        t[i].data = get_or_allocate(p, &tc, job_index)->data;
        ++i;
    }
    *n_out = n;
    return t;
}

static void pyramid_run_phases(bto_pyramid* p, int threads, int phases);
void bto_pyramid_run(bto_pyramid* p, int threads) { pyramid_run_phases(p, threads, 3); }
/* phases: bit 0 = split + downsample, bit 1 = stitch (tests of the sharded exchange stop after bit 0) */
void bto_pyramid_run_phases(bto_pyramid* p, int threads, int phases) { pyramid_run_phases(p, threads, phases); }

/* raw per-tile entry points (no pyramid store): used by the world_size-2 gloo test to stitch / downsample
 * rank-local tiles whose neighbours come from halo exchanges */
void bto_stitch_tile_raw(uint32_t format, uint32_t T, uint32_t B, const bt_tile* tile, void* data,
                         const bt_tile nb_tiles[8], void* const nb[8]) {
    bto_pyramid tmp;
    memset(&tmp, 0, sizeof(tmp));
    tmp.format = format; tmp.T = T; tmp.B = B; tmp.C = T - 2 * B; tmp.pixel_size = (format == BT_FORMAT_R16) ? 2 : 4;
    stitch_tile(&tmp, tile, data, nb_tiles, nb);
}
void bto_downsample_tile_raw(uint32_t format, uint32_t T, uint32_t B, void* data, void* const child[4]) {
    bto_pyramid tmp;
    memset(&tmp, 0, sizeof(tmp));
    tmp.format = format; tmp.T = T; tmp.B = B; tmp.C = T - 2 * B; tmp.pixel_size = (format == BT_FORMAT_R16) ? 2 : 4;
    downsample_tile(&tmp, data, child);
}

static void pyramid_run_phases(bto_pyramid* p, int threads, int phases) {
    (void)threads;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#endif
    int spherical = p->kind == BT_KIND_SPHERICAL;
    for (uint32_t ji = 0; ji < p->n_jobs; ++ji) {
        job* j = &p->jobs[ji];
        /* queue build: every tile of every LOD of every dataset of this job exists before any
         * stitch/downsample task resolves its dependencies (split_and_downsample is called for all
         * datasets first, preprocessor.rs:330-332)                                              */
        for (uint32_t di = 0; di < j->count; ++di) {
            const bt_dataset* d = &j->datasets[di];
            for (uint32_t lod = d->lod_begin; lod < d->lod_end; ++lod) { size_t n; free(collect(p, d, lod, ji, &n)); }
        }
        /* split_and_downsample — preprocessor.rs:234-269 */
        for (uint32_t di = 0; (phases & 1) && di < j->count; ++di) {
            const bt_dataset* d = &j->datasets[di];
            if (d->lod_end <= d->lod_begin) continue;
            size_t n; task* t = collect(p, d, d->lod_end - 1u, ji, &n);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
            for (int64_t i = 0; i < (int64_t)n; ++i) split_tile(p, d, &t[i].tile, t[i].data);
            free(t);
            for (int32_t lod = (int32_t)d->lod_end - 2; lod >= (int32_t)d->lod_begin; --lod) {
                t = collect(p, d, (uint32_t)lod, ji, &n);
                for (size_t i = 0; i < n; ++i) /* TileCoordinate::children — coordinate.rs:197-206 */
                    for (uint32_t c = 0; c < 4; ++c) {
                        bt_tile ch = {t[i].tile.side, t[i].tile.lod + 1u, (t[i].tile.x << 1) + c % 2u, (t[i].tile.y << 1) + c / 2u};
                        t[i].dep[c] = get_tile(p, &ch, ji);
                    }
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
                for (int64_t i = 0; i < (int64_t)n; ++i) downsample_tile(p, t[i].data, t[i].dep);
                free(t);
            }
        }
        /* stitch_and_save_layer — preprocessor.rs:271-288, order :306-310 / :334-340 */
        uint32_t lod_begin = 0xFFFFFFFFu, lod_end = 0;
        for (uint32_t di = 0; di < j->count; ++di) {
            if (j->datasets[di].lod_begin < lod_begin) lod_begin = j->datasets[di].lod_begin;
            if (j->datasets[di].lod_end > lod_end) lod_end = j->datasets[di].lod_end;
        }
        for (uint32_t lod = lod_begin; (phases & 2) && lod < lod_end; ++lod) for (uint32_t di = 0; di < j->count; ++di) {
            const bt_dataset* d = &j->datasets[di];
            if (lod < d->lod_begin || lod >= d->lod_end) continue;
            size_t n; task* t = collect(p, d, lod, ji, &n);
            for (size_t i = 0; i < n; ++i) {
                bto_tile_neighbours(&t[i].tile, spherical, t[i].dep_tiles);
                for (int k = 0; k < 8; ++k) t[i].dep[k] = get_tile(p, &t[i].dep_tiles[k], ji);
            }
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
            for (int64_t i = 0; i < (int64_t)n; ++i) stitch_tile(p, &t[i].tile, t[i].data, t[i].dep_tiles, t[i].dep);
            free(t);
        }
    }
}

/* ============================================================================================ */
/* mip generation at tile load — AttachmentData::generate_mipmaps, terrain_data/mod.rs:143-219   */
/* ============================================================================================ */
void bto_generate_mipmaps(uint32_t format, uint32_t texture_size, uint32_t mip_level_count, const void* mip0, void* out) {
    size_t parent_size = texture_size;
    const uint8_t* parent = (const uint8_t*)mip0;
    uint8_t* dst = (uint8_t*)out;
    for (uint32_t level = 1; level < mip_level_count; ++level) {
        size_t child_size = parent_size >> 1;
        if (format == BT_FORMAT_R16) {
            const uint16_t* ps = (const uint16_t*)parent; uint16_t* cd = (uint16_t*)dst;
            for (size_t cy = 0; cy < child_size; ++cy) for (size_t cx = 0; cx < child_size; ++cx) {
                uint32_t value = 0, count = 0;
                for (size_t x = 0; x < 2; ++x) for (size_t y = 0; y < 2; ++y) { /* generate_mipmap_r16 :168-198 */
                    uint32_t v = ps[((cy << 1) + y) * parent_size + ((cx << 1) + x)];
                    if (v != 0) { value += v; count += 1; }
                }
                cd[cy * child_size + cx] = count == 0 ? 0 : (uint16_t)(value / count);
            }
            parent = dst; dst += child_size * child_size * 2;
        } else {
            for (size_t cy = 0; cy < child_size; ++cy) for (size_t cx = 0; cx < child_size; ++cx) {
                uint64_t value[4] = {0, 0, 0, 0};
                for (size_t i = 0; i < 4; ++i) { /* generate_mipmap_rgba8 :144-166 */
                    size_t px = (cx << 1) + (i >> 1), py = (cy << 1) + (i & 1);
                    const uint8_t* t = parent + (py * parent_size + px) * 4;
                    for (int k = 0; k < 4; ++k) value[k] += t[k];
                }
                uint8_t* o = dst + (cy * child_size + cx) * 4;
                for (int k = 0; k < 4; ++k) o[k] = (uint8_t)(value[k] / 4);
            }
            parent = dst; dst += child_size * child_size * 4;
        }
        parent_size = child_size;
    }
}

/* ============================================================================================ */
/* synthetic sources: integer-only hash value-noise (NOT part of the reference)                 */
/* ============================================================================================ */
static inline uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
static inline uint32_t lattice(uint32_t ix, uint32_t iy, uint32_t seed) {
    return lowbias32(ix + lowbias32(iy + lowbias32(seed))) >> 16; /* 16-bit value */
}
/* 6 octaves, cell size 2^(8-o); value in [0, 65535) */
static uint32_t synth_value(uint32_t x, uint32_t y, uint32_t seed) {
    uint32_t sum = 0;
    for (uint32_t o = 0; o < 6; ++o) {
        uint32_t k = 8u - o;
        uint32_t ix = x >> k, iy = y >> k;
        uint64_t fx = x & ((1u << k) - 1u), fy = y & ((1u << k) - 1u), one = 1u << k;
        uint64_t v00 = lattice(ix, iy, seed + o), v10 = lattice(ix + 1, iy, seed + o);
        uint64_t v01 = lattice(ix, iy + 1, seed + o), v11 = lattice(ix + 1, iy + 1, seed + o);
        uint64_t top = v00 * (one - fx) + v10 * fx, bot = v01 * (one - fx) + v11 * fx;
        uint32_t v = (uint32_t)((top * (one - fy) + bot * fy) >> (2u * k));
        sum += v >> (o + 1u);
    }
    return sum;
}
static inline int synth_hole(uint32_t x, uint32_t y, uint32_t seed) {
    return lowbias32((x >> 7) + lowbias32((y >> 7) + lowbias32(seed ^ 0x5EED0003u))) % 50u == 0u;
}

void bto_synth_source(uint32_t format, uint32_t width, uint32_t height, uint32_t x0, uint32_t y0,
                      uint32_t seed, uint32_t holes, void* out, int threads) {
    (void)threads;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
    for (int64_t yy = 0; yy < (int64_t)height; ++yy) for (uint32_t xx = 0; xx < width; ++xx) {
        uint32_t x = x0 + xx, y = y0 + (uint32_t)yy;
        int hole = holes && synth_hole(x, y, seed);
        if (format == BT_FORMAT_R16) {
            uint32_t v = synth_value(x, y, seed);
            v = v < 1u ? 1u : (v > 65535u ? 65535u : v);
            ((uint16_t*)out)[(size_t)yy * width + xx] = hole ? 0 : (uint16_t)v;
        } else {
            uint8_t* o = (uint8_t*)out + ((size_t)yy * width + xx) * 4;
            for (uint32_t c = 0; c < 3; ++c) {
                uint32_t v = synth_value(x, y, seed + 0x100u * (c + 1u)) >> 8;
                if (c == 0) v = v < 1u ? 1u : v;
                o[c] = hole ? 0 : (uint8_t)(v > 255u ? 255u : v);
            }
            o[3] = hole ? 0 : 255;
        }
    }
}
