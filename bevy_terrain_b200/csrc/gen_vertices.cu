// gen_vertices.cu — per-tile heightmap-sampled vertex-grid generation (positions + normals).
//
// Replaces the vertex stage of the reference's indirect draw (src/shaders/render/vertex.wgsl:85-98,
// driven by DrawTerrainCommand, src/render/terrain_view_bind_group.rs:274-298) and fuses the
// height-derived normal of the fragment stage (sample_normal, src/shaders/attachments.wgsl:51-107,
// blended as src/shaders/render/fragment.wgsl:99-107).
//
// Design (B200):
//   * persistent CTAs (one per SM, warp-specialised: 19 compute warps + 1 TMA store warp) stride over work items
//     = (tile, block of CB column strips); image hand-over by mbarriers, no CTA-wide barrier in steady state
//   * the reference's triangle-strip index stream visits every grid vertex ~2x (8448 strip
//     vertices for 65x65 = 4225 unique ones at grid 64, compute_tile_uv functions.wgsl:64-71);
//     each unique vertex is computed ONCE and scattered to its 2-3 strip positions in a
//     shared-memory image of the output block — bit-identical to recomputation because the
//     vertex is a pure function of (tile, column, row)
//   * the image is written to HBM with TMA 1-D bulk stores (cp.async.bulk.global.shared::cta,
//     one per strip), double-buffered: the stores of item i drain while item i+1 is computed;
//     HBM sees only full-line, fully coalesced writes of the 40-B interleaved records
//   * the uniform block is staged into shared memory with one TMA bulk load + mbarrier
//   * height texels are gathered with read-only (ld.global.nc) 16-bit loads; lanes walk along
//     texel rows so one request touches 1-3 cache lines; exact-f32 bilinear weights (no texture
//     unit: its fixed-point weights are not reproducible)
#include "bt_device.cuh"
#include "bt_kernels.h"

namespace btd {

#ifndef BT_GEN_COMPUTE_WARPS
#define BT_GEN_COMPUTE_WARPS 19
#endif
#ifndef BT_GEN_ROTATION
#define BT_GEN_ROTATION 11
#endif
constexpr uint32_t GEN_ROTATION = BT_GEN_ROTATION;
constexpr int GEN_COMPUTE_WARPS = BT_GEN_COMPUTE_WARPS;  // compute warps per CTA (+ 1 store warp), one CTA per SM
constexpr int GEN_THREADS = (GEN_COMPUTE_WARPS + 1) * 32;   // + the auxiliary warp (per-item tables, TMA stores)
constexpr int GEN_CTAS_PER_SM = 1;
#ifndef BT_GEN_BUFFERS
#define BT_GEN_BUFFERS 2
#endif
#ifndef BT_GEN_MAX_CB
#define BT_GEN_MAX_CB 16
#endif
#ifndef BT_GEN_FULL_SLEEP_NS
#define BT_GEN_FULL_SLEEP_NS 256
#endif
#ifndef BT_GEN_EMPTY_SLEEP_NS
#define BT_GEN_EMPTY_SLEEP_NS 64
#endif
constexpr int GEN_BUFFERS = BT_GEN_BUFFERS;   // output images per CTA: compute of item i+1 overlaps the TMA stores of item i
// Strips are laid out in the smem image with a stride that is 16 (mod 128) bytes: lanes of a warp write records of 32 different
// strips, and with that stride each quarter-warp of a 128-bit store (and each half-warp of a 64-bit one) covers all 32 banks
// once (the 40-byte records themselves are 8-byte aligned; the stride keeps every strip 16-byte aligned for the TMA).
__host__ __device__ constexpr uint32_t strip_stride_for(uint32_t strip_bytes) { return strip_bytes + (144u - strip_bytes % 128u) % 128u; }

struct GenParams {
    const ViewUniforms* uniforms;
    const bt_tile* tiles;
    const bt_indirect* indirect;
    const bt_tile_tree_entry* tile_tree;
    const uint16_t* height_atlas;
    bt_vertex* vertices;
    uint64_t vertex_capacity;
    GenStatus* status;
    uint32_t flags;
    uint32_t grid_size;        // g
    uint32_t strips_per_item;  // CB
    uint32_t items_per_tile;   // ceil(g / CB)
    uint32_t image_bytes;      // CB * strip_stride, per buffer
    uint32_t strip_stride;     // strip_stride_for(vertices_per_row * sizeof(bt_vertex))
    uint32_t magic_items;      // ceil(2^32 / items_per_tile): w / items_per_tile == umulhi(w, magic_items) (w < 2^32 / items_per_tile)
    uint32_t magic_cols;       // ceil(2^32 / (CB + 1)) for full-width items
    uint32_t magic_cols_ragged;  // the same for the narrower last item of a tile (g % CB strips), if any
    uint32_t table_bytes;      // one slot of per-item tables: sizeof(ItemTables) + (g + 1) row entries, 128-byte multiple
    // configs[2] sampler + per-vertex albedo (BT_FLAG_FILTER_TRILINEAR_ANISO / BT_FLAG_ALBEDO; generic kernel only)
    const uint16_t* height_mips;   // [atlas][chain]: levels 1.. of every atlas slot (generate_mipmaps layout), or nullptr
    const uchar4* albedo_atlas;    // [atlas][T][T]
    const uchar4* albedo_mips;
    const float* srgb_lut;         // 256 entries: the sRGB EOTF, decoded per texel before filtering
    uint32_t* albedo_out;          // strip-ordered pack4x8unorm colours, one per vertex
    uint32_t mip_level_count, mip_chain_texels, albedo_attachment;
    uint32_t albedo_image_bytes;   // per buffer: strips_per_item * albedo_stride (0 without BT_FLAG_ALBEDO)
    uint32_t albedo_stride;        // vertices_per_row * 4
};

struct Record {  // one bt_vertex as five 8-byte words
    float2 w[5];
};

// Loop-invariant values of one launch, hoisted into registers once per thread.  Every entry is the
// result of exactly the operations the per-vertex formula would perform, so hoisting is bit-neutral.
struct FrameConsts {
    float g, inv_g;            // grid_size and, when it is a power of two, its exact reciprocal
    bool  g_pow2;
    float Tf;                  // f32(texture_size)
    int   T, Tm1;
    uint32_t interior;         // texel (i, j) with i-1, j-1 < interior has its 4x4 neighbourhood inside the texture
    float min_h, max_h;
    float a_scale, a_offset;   // attachment_uv (attachments.wgsl:7-10)
    float normal_offset;       // 0.5 / attachments[0].size (attachments.wgsl:85)
    float side_length;         // attachments.wgsl:71,78
    float size0;               // attachments[0].size
    Vec3  planar_normal;       // normal_local_to_world for the planar model: a constant (functions.wgsl:14-24)
    float lod_cap;             // f32(lod_count) - 0.00001 (functions.wgsl:52)
    uint32_t tree_size, tree_mask, lod_count;  // tree_mask = tree_size - 1 when tree_size is a power of two, else 0
    uint32_t last_layer;       // atlas_size - 1: texture-array indices clamp to it
    uint32_t morph_mant, blend_mant;   // mantissa thresholds of the certain-zero tests of the morph / blend ratio (range_mantissa_threshold)
    uint32_t blend_window, lod_max;    // 0x7fe000 - blend_mant (0 = shortcut disabled); u32(lod_cap)
};

// The morph and blend ratios are  inverse_mix(lod + range, lod, t) = saturate((t - (lod + range)) / (lod - (lod + range)))  with
// t = log2_det(x); they are exactly 0 whenever t >= fl(lod + range).  For a positive normal x = m * 2^k (1 <= m < 2) the value
// log2_det(x) differs from k + log2(m) by less than 1e-5 (a few ulp of a number < 64, tests/test_oracle_numerics.py), so
//     k == lod  and  log2(m) >= range + 2e-4      (or k > lod, given range <= 0.99)
// guarantees t >= fl(lod + range) without evaluating log2_det.  The threshold on the 23 mantissa bits of m is computed once per
// thread; 0xFFFFFFFF disables the shortcut (range outside (0, 0.99]: every vertex takes the full evaluation).
__device__ __forceinline__ uint32_t range_mantissa_threshold(float range) {
    if (!(range > 0.0f) || range > 0.99f) return 0xFFFFFFFFu;
    const float m = exp2f(range + 2e-4f);                        // in (1, 2); exp2f is accurate to 2 ulp, the 2e-4 margin covers it
    const uint32_t t = (uint32_t)ceilf((m - 1.0f) * 8388608.0f) + 2u;
    return t < 0x007fe000u ? t : 0xFFFFFFFFu;
}
// x = m * 2^k positive and normal?  Splits it into k and the mantissa bits.
__device__ __forceinline__ bool split_normal(float x, int& k, uint32_t& mant) {
    const uint32_t bits = __float_as_uint(x), expo = bits >> 23;   // sign + exponent
    k = (int)expo - 127;
    mant = bits & 0x007fffffu;
    return (expo - 1u) < 253u;
}

template <bool SPHERICAL>
__device__ __forceinline__ FrameConsts make_frame_consts(const ViewUniforms& U) {
    FrameConsts F;
    F.g = U.view_config.grid_size;
    const uint32_t gi = (uint32_t)F.g;
    F.g_pow2 = ((float)gi == F.g) && gi != 0u && (gi & (gi - 1u)) == 0u;
    F.inv_g = 1.0f / F.g;      // exact when g is a power of two: x * inv_g == x / g bit for bit
    F.T = (int)U.texture_size;
    F.Tm1 = F.T - 1;
    F.interior = F.T >= 4 ? (uint32_t)(F.T - 3) : 0u;   // 1 <= i <= T-3  <=>  i-1 >= 0 and i+2 <= T-1
    F.Tf = (float)U.texture_size;
    F.min_h = U.terrain_config.min_height;
    F.max_h = U.terrain_config.max_height;
    F.a_scale = U.attachments[0].scale;
    F.a_offset = U.attachments[0].offset;
    F.size0 = U.attachments[0].size;
    F.normal_offset = 0.5f / F.size0;
    F.side_length = SPHERICAL ? (float)(3.14159265359 / 4.0) * U.terrain_config.scale : U.terrain_config.scale;
    F.planar_normal = normal_local_to_world<false>(U.mesh, Vec3{0.0f, 0.0f, 0.0f});
    F.lod_cap = (float)U.terrain_config.lod_count - 0.00001f;
    F.tree_size = U.view_config.tree_size;
    F.tree_mask = (F.tree_size & (F.tree_size - 1u)) == 0u ? F.tree_size - 1u : 0u;
    F.lod_count = U.terrain_config.lod_count;
    F.last_layer = U.atlas_size ? U.atlas_size - 1u : 0u;
    F.morph_mant = range_mantissa_threshold(U.view_config.morph_range);
    F.blend_mant = range_mantissa_threshold(U.view_config.blend_range);
    F.blend_window = F.blend_mant == 0xFFFFFFFFu ? 0u : 0x007fe000u - F.blend_mant;
    F.lod_max = f2u(F.lod_cap);
    return F;
}

// bilinear R16 fetch — exact f32 weights, texel centres at i+0.5, clamp-to-edge; operation order
// of the CPU twin AttachmentData::sample (src/terrain_data/mod.rs:221-264): lerp y, then x
__device__ __forceinline__ float sample_r16(const FrameConsts& F, const uint16_t* __restrict__ layer, float u, float v) {
    float px = u * F.Tf - 0.5f, py = v * F.Tf - 0.5f;
    float fx = floorf(px), fy = floorf(py);
    float rx = px - fx, ry = py - fy;
    int ix = (int)fx, iy = (int)fy;
    int x0 = min(max(ix, 0), F.Tm1), x1 = min(max(ix + 1, 0), F.Tm1);
    int y0 = min(max(iy, 0), F.Tm1), y1 = min(max(iy + 1, 0), F.Tm1);
    const uint16_t* r0 = layer + y0 * F.T;
    const uint16_t* r1 = layer + y1 * F.T;
    float v00 = unorm16_to_float(__ldg(r0 + x0));
    float v01 = unorm16_to_float(__ldg(r1 + x0));
    float v10 = unorm16_to_float(__ldg(r0 + x1));
    float v11 = unorm16_to_float(__ldg(r1 + x1));
    float a = v00 + (v01 - v00) * ry;
    float b = v10 + (v11 - v10) * ry;
    return a + (b - a) * rx;
}

struct AtlasTile {
    uint32_t index;
    Coord c;
};

// lookup_tile (default variant) — functions.wgsl:232-246,198-206
__device__ __forceinline__ AtlasTile lookup_tile(const FrameConsts& F, const bt_tile_tree_entry* __restrict__ tile_tree,
                                                 Coord c, uint32_t lod) {
    change_lod(c, lod);
    const uint32_t ts = F.tree_size;
    const uint32_t tx = F.tree_mask ? (c.x & F.tree_mask) : c.x % ts;   // xy % tree_size (functions.wgsl:199)
    const uint32_t ty = F.tree_mask ? (c.y & F.tree_mask) : c.y % ts;
    const uint32_t idx = ((c.side * F.lod_count + c.lod) * ts + tx) * ts + ty;
    const uint2 e = __ldg(reinterpret_cast<const uint2*>(tile_tree) + idx);
    change_lod(c, e.y);
    return AtlasTile{e.x, c};
}

// sample_height — attachments.wgsl:7-24,45-49
__device__ __forceinline__ float sample_height_uv(const FrameConsts& F, const uint16_t* __restrict__ layer, float u, float v) {
    const float s = sample_r16(F, layer, u, v);
    return mix(F.min_h, F.max_h, s);
}

// The five taps of sample_height + sample_normal (centre, left, up, right, down) land in a "plus" of at most
// 12 distinct texels: rows iy..iy+1 x columns ix-1..ix+2 and rows iy-1..iy+2 x columns ix..ix+1.  This routine
// loads and decodes each texel once and evaluates every tap with exactly the per-tap arithmetic of sample_r16
// (same operands, same order), so it is bit-identical to five independent taps.  `v + 0.0f == v` and
// `u + 0.0f == u` hold bit for bit because attachment_uv yields values >= +0.
struct FiveTaps { float centre, left, up, right, down; bool ok; };

__device__ __forceinline__ float lerp2(float v00, float v01, float v10, float v11, float rx, float ry) {
    const float a = v00 + (v01 - v00) * ry;
    const float b = v10 + (v11 - v10) * ry;
    return a + (b - a) * rx;
}

// The five taps in three steps, so that the caller can put the loads of TWO vertices in flight before consuming either
// (the kernel is latency-bound at 5 warps per scheduler: two independent vertices per lane double the work each warp can issue
// while its texel loads are outstanding):
//   prepare_taps : tap coordinates, weights, which-texel selectors, and whether the whole plus lies inside the texture
//   load_taps    : the 12 raw texels (interior: one base address, +-T row pointers, immediate column offsets; else clamped)
//   finish_taps  : decode + the five bilinear results
struct TapSetup {
    const uint16_t* layer;
    float rxc, ryc, rxl, rxr, ryu, ryd;
    int ix, iy;
    bool jl, jr, ju, jd;      // left / right / up / down tap starts one texel further (offsets of ~0.5 texel: 0 or 1)
    bool ok;                  // the selectors are all 0 / 1 (always, for offsets < 1 texel; guarded anyway)
    bool interior;
};
struct TapTexels { uint32_t r10, r11, r12, r13, r20, r21, r22, r23, r01, r02, r31, r32; };   // raw texel r<row><col>

__device__ __forceinline__ TapSetup prepare_taps(const FrameConsts& F, const uint16_t* __restrict__ layer, float u, float v) {
    const float o = F.normal_offset;
    // tap coordinates: (centre x, centre y), (left x, right x), (up y, down y) as packed pairs; each product is followed by a
    // scalar "- 0.5" (same two roundings per value as sample_r16)
    const float2 Tf2 = make_float2(F.Tf, F.Tf);
    const float2 pc = __fmul2_rn(make_float2(u, v), Tf2);
    const float2 plr = __fmul2_rn(__fadd2_rn(make_float2(u, u), make_float2(-o, o)), Tf2);
    const float2 pud = __fmul2_rn(__fadd2_rn(make_float2(v, v), make_float2(-o, o)), Tf2);
    const float pxc = pc.x - 0.5f, pyc = pc.y - 0.5f, pxl = plr.x - 0.5f, pxr = plr.y - 0.5f, pyu = pud.x - 0.5f, pyd = pud.y - 0.5f;
    // texel coordinates are far inside floor_small's range (|p| < 2^22): floors and the two integer indices without FRND / F2I
    const FloorInt fcx = floor_small(pxc), fcy = floor_small(pyc);
    const float fxc = fcx.f, fyc = fcy.f, fxl = floor_small_f(pxl), fxr = floor_small_f(pxr), fyu = floor_small_f(pyu), fyd = floor_small_f(pyd);
    const float2 rc = __fadd2_rn(make_float2(pxc, pyc), make_float2(-fxc, -fyc));
    const float2 rlr = __fadd2_rn(make_float2(pxl, pxr), make_float2(-fxl, -fxr));
    const float2 rud = __fadd2_rn(make_float2(pyu, pyd), make_float2(-fyu, -fyd));
    TapSetup s;
    s.layer = layer;
    s.rxc = rc.x; s.ryc = rc.y; s.rxl = rlr.x; s.rxr = rlr.y; s.ryu = rud.x; s.ryd = rud.y;
    s.ix = fcx.i;
    s.iy = fcy.i;
    // the floors are integers in float form: the selectors come from float compares (no float->int conversions on the XU pipe)
    s.jl = fxl == fxc;             // (int)fxl - ix + 1 == 1
    s.jr = fxr != fxc;             // (int)fxr - ix == 1
    s.ju = fyu == fyc;
    s.jd = fyd != fyc;
    s.ok = (s.jl || fxl == fxc - 1.0f) && (!s.jr || fxr == fxc + 1.0f) && (s.ju || fyu == fyc - 1.0f) && (!s.jd || fyd == fyc + 1.0f);
    s.interior = (uint32_t)(s.ix - 1) < F.interior && (uint32_t)(s.iy - 1) < F.interior;
    return s;
}

__device__ __forceinline__ TapTexels load_taps_interior(const FrameConsts& F, const TapSetup& s) {
    // the whole plus lies inside the texture (always the case for uv inside a tile with border_size >= 2)
    const uint16_t* p1 = s.layer + (s.iy * F.T + s.ix);
    const uint16_t* p0 = p1 - F.T;
    const uint16_t* p2 = p1 + F.T;
    const uint16_t* p3 = p2 + F.T;
    TapTexels t;
    t.r10 = __ldg(p1 - 1); t.r11 = __ldg(p1); t.r12 = __ldg(p1 + 1); t.r13 = __ldg(p1 + 2);
    t.r20 = __ldg(p2 - 1); t.r21 = __ldg(p2); t.r22 = __ldg(p2 + 1); t.r23 = __ldg(p2 + 2);
    t.r01 = __ldg(p0); t.r02 = __ldg(p0 + 1);
    t.r31 = __ldg(p3); t.r32 = __ldg(p3 + 1);
    return t;
}

__device__ __forceinline__ TapTexels load_taps_clamped(const FrameConsts& F, const TapSetup& s) {
    const int ix = s.ix, iy = s.iy;
    const int c0 = min(max(ix - 1, 0), F.Tm1), c1 = min(max(ix, 0), F.Tm1), c2 = min(max(ix + 1, 0), F.Tm1), c3 = min(max(ix + 2, 0), F.Tm1);
    const uint16_t* q0 = s.layer + min(max(iy - 1, 0), F.Tm1) * F.T;
    const uint16_t* q1 = s.layer + min(max(iy, 0), F.Tm1) * F.T;
    const uint16_t* q2 = s.layer + min(max(iy + 1, 0), F.Tm1) * F.T;
    const uint16_t* q3 = s.layer + min(max(iy + 2, 0), F.Tm1) * F.T;
    TapTexels t;
    t.r10 = __ldg(q1 + c0); t.r11 = __ldg(q1 + c1); t.r12 = __ldg(q1 + c2); t.r13 = __ldg(q1 + c3);
    t.r20 = __ldg(q2 + c0); t.r21 = __ldg(q2 + c1); t.r22 = __ldg(q2 + c2); t.r23 = __ldg(q2 + c3);
    t.r01 = __ldg(q0 + c1); t.r02 = __ldg(q0 + c2);
    t.r31 = __ldg(q3 + c1); t.r32 = __ldg(q3 + c2);
    return t;
}

__device__ __forceinline__ FiveTaps finish_taps(const TapSetup& s, const TapTexels& t) {
    FiveTaps out;
    out.ok = s.ok;
    const float rxc = s.rxc, ryc = s.ryc, rxl = s.rxl, rxr = s.rxr, ryu = s.ryu, ryd = s.ryd;
    const bool jl = s.jl, jr = s.jr, ju = s.ju, jd = s.jd;
    // decode in pairs of adjacent columns of one row — the layout the packed y-lerps consume without re-pairing
    const float2 A01 = unorm16x2_to_float(t.r10, t.r11), A23 = unorm16x2_to_float(t.r12, t.r13);   // row 1: (t10, t11), (t12, t13)
    const float2 B01 = unorm16x2_to_float(t.r20, t.r21), B23 = unorm16x2_to_float(t.r22, t.r23);   // row 2
    const float2 E0 = unorm16x2_to_float(t.r01, t.r02), E3 = unorm16x2_to_float(t.r31, t.r32);     // rows 0 and 3: columns 1, 2
    // y-lerps of rows (1, 2) with the centre's ry:  L<col> = t1c + (t2c - t1c) * ryc  — lerp2's `a` / `b` for centre, left, right.
    // Packed: difference and product for two columns per instruction; the final additions stay scalar (a packed mul feeding a
    // packed add would be contracted into FFMA2 by ptxas, which the numerics contract forbids).
    const float2 ryc2 = make_float2(ryc, ryc);
    const float2 d01 = __fmul2_rn(__fadd2_rn(B01, make_float2(-A01.x, -A01.y)), ryc2);
    const float2 d23 = __fmul2_rn(__fadd2_rn(B23, make_float2(-A23.x, -A23.y)), ryc2);
    const float L0 = A01.x + d01.x, L1 = A01.y + d01.y, L2 = A23.x + d23.x, L3 = A23.y + d23.y;
    // x-stage: a + (b - a) * rx
    out.centre = L1 + (L2 - L1) * rxc;
    // left: columns (ix-1+jl, ix+jl); right: columns (ix+jr, ix+1+jr); rows iy, iy+1 (py == centre's)
    {
        const float la = jl ? L1 : L0, lb = jl ? L2 : L1, ra = jr ? L2 : L1, rb = jr ? L3 : L2;
        const float2 m = __fmul2_rn(__fadd2_rn(make_float2(lb, rb), make_float2(-la, -ra)), make_float2(rxl, rxr));
        out.left = la + m.x;
        out.right = ra + m.y;
    }
    // up: rows (iy-1+ju, iy+ju); down: rows (iy+jd, iy+1+jd); columns ix, ix+1 (px == centre's)
    {
        const float2 M1 = make_float2(A01.y, A23.x), M2 = make_float2(B01.y, B23.x);   // (t11, t12), (t21, t22)
        // (v00, v10) / (v01, v11): up = ju ? rows (1, 2) : rows (0, 1);  down = jd ? rows (2, 3) : rows (1, 2)
        const float2 ulo = ju ? M1 : E0, uhi = ju ? M2 : M1;
        const float2 wlo = jd ? M2 : M1, whi = jd ? E3 : M2;
        const float2 mu = __fmul2_rn(__fadd2_rn(uhi, make_float2(-ulo.x, -ulo.y)), make_float2(ryu, ryu));
        const float2 md = __fmul2_rn(__fadd2_rn(whi, make_float2(-wlo.x, -wlo.y)), make_float2(ryd, ryd));
        const float ua = ulo.x + mu.x, ub = ulo.y + mu.y, da = wlo.x + md.x, db = wlo.y + md.y;
        const float2 m = __fmul2_rn(__fadd2_rn(make_float2(ub, db), make_float2(-ua, -da)), make_float2(rxc, rxc));
        out.up = ua + m.x;
        out.down = da + m.y;
    }
    return out;
}

__device__ __forceinline__ FiveTaps sample_five(const FrameConsts& F, const uint16_t* __restrict__ layer, float u, float v) {
    const TapSetup s = prepare_taps(F, layer, u, v);
    const TapTexels t = s.interior ? load_taps_interior(F, s) : load_taps_clamped(F, s);
    return finish_taps(s, t);
}

// sample_normal — attachments.wgsl:51-107 (textureSampleLevel branch); `pre` carries the shared taps, `distance_between_samples`
// = side_length / (attachments[0].size * tile_count(tile.lod)) comes from the per-item LOD record (or the generic lookup)
template <bool SPHERICAL>
__device__ __forceinline__ Vec3 sample_normal(const FrameConsts& F, const uint16_t* __restrict__ layer, uint32_t side,
                                              float distance_between_samples, float u, float v, Vec3 vertex_normal, const FiveTaps& pre) {
    const float offset = F.normal_offset;
    float left, up, right, down;
    if (pre.ok) {
        // mix(min_h, max_h, s) = min_h * (1 - s) + max_h * s for two taps at a time; the sum of the two products stays scalar
        const float2 lo = make_float2(F.min_h, F.min_h), hi = make_float2(F.max_h, F.max_h), one = make_float2(1.0f, 1.0f);
        const float2 s_lr = make_float2(pre.left, pre.right), s_ud = make_float2(pre.up, pre.down);
        const float2 a_lr = __fmul2_rn(lo, __fadd2_rn(one, make_float2(-s_lr.x, -s_lr.y))), b_lr = __fmul2_rn(hi, s_lr);
        const float2 a_ud = __fmul2_rn(lo, __fadd2_rn(one, make_float2(-s_ud.x, -s_ud.y))), b_ud = __fmul2_rn(hi, s_ud);
        left = a_lr.x + b_lr.x;
        right = a_lr.y + b_lr.y;
        up = a_ud.x + b_ud.x;
        down = a_ud.y + b_ud.y;
    } else {
        left = sample_height_uv(F, layer, u + -offset, v + 0.0f);
        up = sample_height_uv(F, layer, u + 0.0f, v + -offset);
        right = sample_height_uv(F, layer, u + offset, v + 0.0f);
        down = sample_height_uv(F, layer, u + 0.0f, v + offset);
    }
    const Vec3 sn = normalize3(Vec3{left - right, down - up, distance_between_samples});
    Vec3 r;
    if (SPHERICAL) {
        Vec3 face_up;
        switch (side) {
            case 0: case 1: face_up = Vec3{0.0f, 1.0f, 0.0f}; break;
            case 2: case 3: face_up = Vec3{0.0f, 0.0f, -1.0f}; break;
            default: face_up = Vec3{-1.0f, 0.0f, 0.0f}; break;
        }
        const Vec3 n = normalize3(vertex_normal);
        const Vec3 tg = cross3(face_up, n);
        const Vec3 bt = cross3(n, tg);
        r = Vec3{(tg.x * sn.x + bt.x * sn.y) + n.x * sn.z, (tg.y * sn.x + bt.y * sn.y) + n.y * sn.z,
                 (tg.z * sn.x + bt.z * sn.y) + n.z * sn.z};
    } else {
        r = Vec3{sn.x, sn.z, sn.y};  // TBN columns (1,0,0),(0,0,1),(0,1,0)
    }
    return normalize3(r);
}

// ------------------------------------------------------------------------------------------------
// Per-item tables (shared memory), produced by the SETUP warp a few items ahead of the compute warps.
//
// Everything in vertex_info / lookup_tile that depends only on (tile, grid column), (tile, grid row) or
// (tile, blend LOD) is evaluated ONCE per item with exactly the operations the per-vertex formula performs
// (so the per-vertex result is bit-identical), instead of ~150 instructions per vertex:
//   col[lx]  : coordinate u of the grid column and the column part of compute_local_position /
//              position_local_to_world (functions.wgsl:73-96,26-29)
//   row[gy]  : the same for grid rows
//   lod[L]   : lookup_tile(coordinate, L) for every blend LOD L <= tile.lod: both coordinate_change_lod steps of
//              functions.wgsl:232-246 collapse to  uv' = ((f1 + uv) * s1 + f2) * s2  with per-(item, L) constants — a step with
//              delta 0 is encoded as f = 0, s = 1, which leaves uv unchanged bit for bit (uv >= +0) — plus the
//              tile-tree entry's atlas layer and sample_normal's distance_between_samples.
// Blend LODs above the tile's own LOD (a geometry tile coarser than the data it samples: lookups then differ
// per vertex) and table entries with atlas_lod > lod take the generic per-vertex path.
// ------------------------------------------------------------------------------------------------
constexpr int GEN_TABLE_SLOTS = 2 * GEN_BUFFERS;   // tables of up to 2 * GEN_BUFFERS items in flight
constexpr int GEN_MAX_COLS = 32;        // strips per item + 1 <= 17
constexpr uint32_t LODREC_SLOW = 0x80000000u;

struct alignas(16) LodRecord {
    float f1x, f1y, f2x, f2y;           // offsets of the two change_lod steps
    float s1, s2;                       // their scales (2^delta)
    float dbs;                          // distance_between_samples of the resolved atlas tile
    uint32_t layer;                     // min(atlas_index, atlas_size - 1) | LODREC_SLOW
};

struct alignas(16) ItemTables {
    float4 col[GEN_MAX_COLS];
    LodRecord lod[32];
    uint4 tile;                         // side, lod, x, y
    uint4 meta;                         // tile_index, n_strips, magic (floor(i / n_cols) == umulhi(i, magic)), bits of morph_zero_from
    // float4 row[grid_size + 1] follows
};

template <bool SPHERICAL>
__device__ __forceinline__ float4 axis_terms(const ViewUniforms& U, const FrameConsts& F, uint32_t tile_xy, float inv_tc, uint32_t g_index, int axis) {
    // compute_tile_uv — functions.wgsl:64-71 (x / g == x * (1/g) exactly when g is a power of two)
    const float t = F.g_pow2 ? (float)g_index * F.inv_g : (float)g_index / F.g;
    float w = ((float)tile_xy + t) * inv_tc;        // compute_local_position: (xy + uv) / tile_count(lod)
    if (!SPHERICAL) {
        const float p = w - 0.5f;                   // local x (axis 0) or z (axis 1); local y = 0
        const float* a = U.mesh.world_from_local;
        if (axis == 0)   // position_local_to_world: (a[k] * p.x + a[3 + k] * p.y) — the first sum of every row, with p.y = 0
            return make_float4(t, a[0] * p + a[3] * 0.0f, a[1] * p + a[4] * 0.0f, a[2] * p + a[5] * 0.0f);
        return make_float4(t, a[6] * p, a[7] * p, a[8] * p);   // ... + a[6 + k] * p.z
    }
    w = (w - 0.5f) * 2.0f;                          // (uv - 0.5) / 0.5, exact
    const float q = div_rn(w, sqrt_rn(BT_ONE_PLUS_C_SQR_F - (BT_C_SQR_F * w) * w));   // the cube-to-sphere warp of one axis
    return make_float4(t, q, 0.0f, 0.0f);
}

// One item's tables, computed by GEN_BUILDER_LANES lanes (index `lane` = 0..30) of the auxiliary warp.
constexpr uint32_t GEN_BUILDER_LANES = 31u;
template <bool SPHERICAL>
__device__ __forceinline__ void build_item_tables(const ViewUniforms& U, const FrameConsts& F, const bt_tile_tree_entry* __restrict__ tile_tree,
                                                  ItemTables& T, float4* rows, const bt_tile& tile, uint32_t tile_index, uint32_t magic,
                                                  uint32_t c0, uint32_t n_cols, uint32_t n_rows, int lane) {
    const float inv_tc = exp2i(-(int)(tile.lod & 31u));
    for (uint32_t lx = (uint32_t)lane; lx < n_cols; lx += GEN_BUILDER_LANES) T.col[lx] = axis_terms<SPHERICAL>(U, F, tile.x, inv_tc, c0 + lx, 0);
    for (uint32_t gy = (uint32_t)lane; gy < n_rows; gy += GEN_BUILDER_LANES) rows[gy] = axis_terms<SPHERICAL>(U, F, tile.y, inv_tc, gy, 1);
    if (lane == 0) {
        T.tile = make_uint4(tile.side, tile.lod, tile.x, tile.y);
        // compute_morph's ratio is exactly 0 for every x = 2 * morph_distance / d >= morph_zero_from (range_mantissa_threshold):
        // the float with exponent tile.lod and mantissa bits morph_mant; LOD 0 never morphs (-inf); +inf disables the shortcut
        const uint32_t zero_from = tile.lod == 0u ? 0xff800000u : (F.morph_mant == 0xFFFFFFFFu ? 0x7f800000u : (((127u + (tile.lod & 31u)) << 23) | F.morph_mant));
        T.meta = make_uint4(tile_index, n_cols - 1u, magic, zero_from);
        T.lod[31] = LodRecord{0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 1.0f, 0.0f, LODREC_SLOW};   // LOD 31 does not exist (lod_count <= 31): generic path
    }
    const uint32_t L = (uint32_t)lane;
    LodRecord r{0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 1.0f, 0.0f, LODREC_SLOW};
    if (L < F.lod_count && L <= tile.lod) {
        // first coordinate_change_lod: tile.lod -> L (down or equal)
        uint32_t x = tile.x, y = tile.y;
        const int d1 = (int)L - (int)tile.lod;
        if (d1 != 0) {
            const uint32_t shift = (uint32_t)(-d1) & 31u, mask = (1u << shift) - 1u;
            r.f1x = (float)(x & mask);
            r.f1y = (float)(y & mask);
            r.s1 = exp2i(d1);
            x >>= shift;
            y >>= shift;
        }
        const uint32_t ts = F.tree_size;
        const uint32_t tx = F.tree_mask ? (x & F.tree_mask) : x % ts, ty = F.tree_mask ? (y & F.tree_mask) : y % ts;
        const uint2 e = __ldg(reinterpret_cast<const uint2*>(tile_tree) + (((tile.side * F.lod_count + L) * ts + tx) * ts + ty));
        // second coordinate_change_lod: L -> entry.atlas_lod (down or equal for a valid table; INVALID = 0xFFFFFFFF is "down" too)
        const int d2 = (int)e.y - (int)L;
        if (d2 <= 0) {
            if (d2 != 0) {
                const uint32_t shift = (uint32_t)(-d2) & 31u, mask = (1u << shift) - 1u;
                r.f2x = (float)(x & mask);
                r.f2y = (float)(y & mask);
                r.s2 = exp2i(d2);
            }
            r.dbs = div_rn(F.side_length, F.size0 * tile_count(e.y));
            r.layer = min(e.x, F.last_layer);
        }
    }
    T.lod[lane] = r;
}

struct Lookup {
    const uint16_t* layer;
    float u, v;        // attachment_uv of the resolved atlas coordinate
    float dbs;
};

// lookup_tile + attachment_uv through the item's LOD record; generic per-vertex evaluation when the record says so
template <bool SPHERICAL>
__device__ __forceinline__ Lookup lookup_sample_point(const FrameConsts& F, const ItemTables& T, const bt_tile_tree_entry* __restrict__ tile_tree,
                                                      const uint16_t* __restrict__ atlas, const bt_tile& tile, float cu, float cv, uint32_t lod) {
    Lookup out;
    const uint32_t tile_texels = (uint32_t)(F.T * F.T);
    const LodRecord* rec = &T.lod[lod & 31u];
    const float4 ra = *reinterpret_cast<const float4*>(rec);
    const float4 rb = *reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(rec) + 16);
    const uint32_t layer_bits = __float_as_uint(rb.w);
    float u, v;
    if (lod < 32u && !(layer_bits & LODREC_SLOW)) {
        const float2 t1 = __fmul2_rn(__fadd2_rn(make_float2(ra.x, ra.y), make_float2(cu, cv)), make_float2(rb.x, rb.x));
        const float2 t2 = __fmul2_rn(make_float2(ra.z + t1.x, ra.w + t1.y), make_float2(rb.y, rb.y));
        u = t2.x;
        v = t2.y;
        out.layer = atlas + (size_t)layer_bits * tile_texels;
        out.dbs = rb.z;
    } else {
        const AtlasTile t = lookup_tile(F, tile_tree, Coord{tile.side, tile.lod, tile.x, tile.y, cu, cv}, lod);
        u = t.c.u;
        v = t.c.v;
        out.layer = atlas + (size_t)min(t.index, F.last_layer) * tile_texels;
        out.dbs = div_rn(F.side_length, F.size0 * tile_count(t.c.lod));
    }
    // attachment_uv — attachments.wgsl:7-10
    const float2 s = __fmul2_rn(make_float2(u, v), make_float2(F.a_scale, F.a_scale));
    out.u = s.x + F.a_offset;
    out.v = s.y + F.a_offset;
    return out;
}

// floor(log2_det(x)) clamped like compute_blend does, WITHOUT evaluating log2_det when the result is certain:
// for a positive normal x = m * 2^k with m <= 2 - 2^-10, log2_det(x) lies in [k, k + 1 - 7e-4] (its error is < 1e-5), so
// u32(min(log2_det(x), cap)) == clamp(k, 0, u32(cap)) exactly.  Everything else (m within 2^-10 of 2, denormals, inf, nan,
// non-positive) evaluates the full function.
__device__ __forceinline__ uint32_t blend_lod_of(float x, float lod_cap) {
    int k;
    uint32_t mant;
    if (split_normal(x, k, mant) && mant < 0x007fe000u) return (uint32_t)min(max(k, 0), (int)f2u(lod_cap));
    return f2u(fminf(log2_det(x), lod_cap));
}

// vertex_info + vertex + vertex_output — vertex.wgsl:30-98, in two halves around the texel loads of the first sample:
//   vertex_front : vertex_info (position, distance, morph, blend) + lookup_tile + the tap coordinates of the first sample
//   vertex_back  : the samples' arithmetic, the blend with the LOD below, vertex_output
// so that the compute loop can run the fronts of two vertices, issue both sets of texel loads, and only then consume them.
struct Front {
    Vec3 wp, wn;
    float cu, cv;
    uint32_t blend_lod;
    float blend_ratio;
    Lookup la;
    TapSetup taps;     // NORMALS only
};

template <bool SPHERICAL>
__device__ __forceinline__ Front vertex_front(const ViewUniforms& U, const FrameConsts& F, const ItemTables& T, const float4* __restrict__ rows,
                                              const bt_tile_tree_entry* __restrict__ tile_tree, const uint16_t* __restrict__ atlas,
                                              uint32_t flags, const bt_tile& tile, float morph_zero_from, uint32_t lx, uint32_t gy) {
    const bt_view_config_uniform& vc = U.view_config;
    const bt_model_approximation& ap = U.approximation;
    const float g = F.g;
    const float4 cr = T.col[lx], rr = rows[gy];
    const float u0 = cr.x, v0 = rr.x;          // compute_tile_uv
    Vec3 lp, wp, wn;
    if (!SPHERICAL) {
        const float* a = U.mesh.world_from_local;
        wp = Vec3{(cr.y + rr.y) + a[9], (cr.z + rr.z) + a[10], (cr.w + rr.w) + a[11]};   // position_local_to_world, column + row + translation
        wn = F.planar_normal;
        lp = Vec3{0.0f, 0.0f, 0.0f};
    } else {
        const float qu = cr.y, qv = rr.y;
        Vec3 p;
        switch (tile.side) {
            case 0: p = Vec3{-1.0f, -qv, qu}; break;
            case 1: p = Vec3{qu, -qv, 1.0f}; break;
            case 2: p = Vec3{qu, 1.0f, qv}; break;
            case 3: p = Vec3{1.0f, -qu, qv}; break;
            case 4: p = Vec3{qv, -qu, -1.0f}; break;
            case 5: p = Vec3{qv, -1.0f, qu}; break;
            default: p = Vec3{0.0f, 0.0f, 0.0f}; break;
        }
        lp = normalize3(p);
        wp = position_local_to_world(U.mesh, lp);
        wn = normal_local_to_world<SPHERICAL>(U.mesh, lp);
    }
    const float h0 = ap.approximate_height;
    const float* e = U.view_world_position;
    float d = length3(Vec3{(wp.x + h0 * wn.x) - e[0], (wp.y + h0 * wn.y) - e[1], (wp.z + h0 * wn.z) - e[2]});
    const Coord c0{tile.side, tile.lod, tile.x, tile.y, u0, v0};
    const bool high_precision = (flags & BT_FLAG_HIGH_PRECISION) && d < vc.precision_threshold_distance;
    if (high_precision) {
        const Vec3 rel = relative_position(ap, c0);
        d = length3(Vec3{rel.x + h0 * wn.x, rel.y + h0 * wn.y, rel.z + h0 * wn.z});
    }
    // compute_morph — functions.wgsl:35-49
    float cu = u0, cv = v0;
    bool morphed = false;    // false: the morph ratio is exactly 0, so mix(uv, even_uv, 0) == uv bit for bit and nothing moves
    if (flags & BT_FLAG_MORPH) {
        const float xm = div_rn(2.0f * vc.morph_distance, d);
        if (!(xm >= morph_zero_from)) {      // (positive floats order like their (exponent, mantissa) pairs; +inf gives ratio 0 as well)
            const float target_lod = log2_det(xm);
            const float ratio = (tile.lod == 0u) ? 0.0f : inverse_mix_rn((float)tile.lod + vc.morph_range, (float)tile.lod, target_lod);
            const float even_u = F.g_pow2 ? (float)(f2u(u0 * g) & ~1u) * F.inv_g : (float)(f2u(u0 * g) & ~1u) / g;
            const float even_v = F.g_pow2 ? (float)(f2u(v0 * g) & ~1u) * F.inv_g : (float)(f2u(v0 * g) & ~1u) / g;
            cu = mix(u0, even_u, ratio);
            cv = mix(v0, even_v, ratio);
            morphed = true;
        }
    }
    if (high_precision) {
        const Vec3 rel = relative_position(ap, Coord{tile.side, tile.lod, tile.x, tile.y, cu, cv});
        wp = Vec3{e[0] + rel.x, e[1] + rel.y, e[2] + rel.z};
    } else if (morphed) {
        // without MORPH the coordinate is unchanged and the recomputation would reproduce lp/wp/wn bit for bit
        lp = local_position<SPHERICAL>(Coord{tile.side, tile.lod, tile.x, tile.y, cu, cv});
        wp = position_local_to_world(U.mesh, lp);
        if (SPHERICAL) wn = normal_local_to_world<SPHERICAL>(U.mesh, lp);
    }
    // compute_blend — functions.wgsl:51-62
    uint32_t blend_lod;
    float blend_ratio = 0.0f;
    const float blend_x = div_rn(vc.blend_distance, d);
    if (flags & BT_FLAG_BLEND) {
        const uint32_t bits = __float_as_uint(blend_x), expo = bits >> 23;   // sign + exponent
        if ((expo - 1u) < 126u) {
            blend_lod = 0u;                        // positive normal below 1: log2 < 0, u32() clamps to 0, and LOD 0 never blends
        } else if ((expo - 127u) <= F.lod_max && ((bits & 0x007fffffu) - F.blend_mant) < F.blend_window) {
            blend_lod = expo - 127u;               // floor is certain and the ratio is certainly 0 (see range_mantissa_threshold)
        } else {
            const float target_lod = fminf(log2_det(blend_x), F.lod_cap);
            blend_lod = f2u(target_lod);
            if (blend_lod != 0u) blend_ratio = inverse_mix_rn((float)blend_lod + vc.blend_range, (float)blend_lod, target_lod);
        }
    } else {
        blend_lod = blend_lod_of(blend_x, F.lod_cap);   // only the integer part is consumed
    }

    Front f;
    f.wp = wp; f.wn = wn; f.cu = cu; f.cv = cv; f.blend_lod = blend_lod; f.blend_ratio = blend_ratio;
    f.la = lookup_sample_point<SPHERICAL>(F, T, tile_tree, atlas, tile, cu, cv, blend_lod);
    if (flags & BT_FLAG_NORMALS) f.taps = prepare_taps(F, f.la.layer, f.la.u, f.la.v);
    return f;
}

template <bool SPHERICAL>
__device__ __forceinline__ Record vertex_back(const FrameConsts& F, const ItemTables& T, const bt_tile_tree_entry* __restrict__ tile_tree,
                                              const uint16_t* __restrict__ atlas, uint32_t flags, const bt_tile& tile, uint32_t tile_index,
                                              const Front& f, const TapTexels* texels) {
    const bool normals = (flags & BT_FLAG_NORMALS) != 0u;
    const Vec3 wp = f.wp, wn = f.wn;
    const float cu = f.cu, cv = f.cv, blend_ratio = f.blend_ratio;
    const Lookup& la = f.la;
    float height;
    Vec3 nrm = wn;
    if (normals) {
        const FiveTaps taps = finish_taps(f.taps, *texels);
        height = mix(F.min_h, F.max_h, taps.centre);
        nrm = sample_normal<SPHERICAL>(F, la.layer, tile.side, la.dbs, la.u, la.v, wn, taps);
    } else {
        height = sample_height_uv(F, la.layer, la.u, la.v);
    }
    if (blend_ratio > 0.0f) {
        const Lookup lb = lookup_sample_point<SPHERICAL>(F, T, tile_tree, atlas, tile, cu, cv, f.blend_lod - 1u);
        if (normals) {  // fragment.wgsl:103-107
            const FiveTaps taps = sample_five(F, lb.layer, lb.u, lb.v);
            height = mix(height, mix(F.min_h, F.max_h, taps.centre), blend_ratio);
            const Vec3 n2 = sample_normal<SPHERICAL>(F, lb.layer, tile.side, lb.dbs, lb.u, lb.v, wn, taps);
            nrm = Vec3{mix(nrm.x, n2.x, blend_ratio), mix(nrm.y, n2.y, blend_ratio), mix(nrm.z, n2.z, blend_ratio)};
        } else {
            height = mix(height, sample_height_uv(F, lb.layer, lb.u, lb.v), blend_ratio);
        }
    }
    Record r;  // vertex_output — vertex.wgsl:73-83
    const float2 hn = __fmul2_rn(make_float2(height, height), make_float2(wn.x, wn.y));
    r.w[0] = make_float2(wp.x + hn.x, wp.y + hn.y);
    r.w[1] = make_float2(wp.z + height * wn.z, 1.0f);
    r.w[2] = make_float2(nrm.x, nrm.y);
    r.w[3] = make_float2(nrm.z, cu);
    r.w[4] = make_float2(cv, __uint_as_float(tile_index));
    return r;
}

// One 40-byte record at dst (8-byte aligned): 16-byte aligned records take two 128-bit stores + one 64-bit store, the others
// 64 + 128 + 128.  ODD is static at every call site.
template <int ODD>
__device__ __forceinline__ void store_record_at(unsigned char* dst, const Record& r) {
    if (ODD) {
        // (five 64-bit stores would save the register moves that the two quad layouts of one record cost, but 64-bit stores are 2-way
        //  bank-conflicted at the image's strip stride: measured slower)
        *reinterpret_cast<float2*>(dst) = r.w[0];
        *reinterpret_cast<float4*>(dst + 8) = make_float4(r.w[1].x, r.w[1].y, r.w[2].x, r.w[2].y);
        *reinterpret_cast<float4*>(dst + 24) = make_float4(r.w[3].x, r.w[3].y, r.w[4].x, r.w[4].y);
    } else {
        *reinterpret_cast<float4*>(dst) = make_float4(r.w[0].x, r.w[0].y, r.w[1].x, r.w[1].y);
        *reinterpret_cast<float4*>(dst + 16) = make_float4(r.w[2].x, r.w[2].y, r.w[3].x, r.w[3].y);
        *reinterpret_cast<float2*>(dst + 32) = r.w[4];
    }
}

// ------------------------------------------------------------------------------------------------
// BASELINE configs[2]: the reference's sampler (linear / linear / linear, anisotropy clamp 16 — terrain_bind_group.rs:120-126;
// textureSampleGrad in attachments.wgsl:16-20) evaluated per vertex with the footprint of one grid step, and the per-vertex
// albedo.  Same operations in the same order as the oracle's sample_filtered / gen_vertex_filtered.
// Generic kernel only (FLAGS_CT < 0): an opt-in variant, not the headline path.
// ------------------------------------------------------------------------------------------------
struct FilterCtx {
    const void* mip0;
    const void* mips;
    const float* lut;      // non-null: Rgba8 texels (sRGB rgb through the table, alpha / 255)
    uint32_t T, mip_level_count;
    bool trilinear;
};

__device__ __forceinline__ uint32_t mip_offset_texels(uint32_t T, uint32_t level) {
    uint32_t off = 0;
    for (uint32_t j = 1; j < level; ++j) off += (T >> j) * (T >> j);
    return off;
}

template <bool RGBA>
__device__ __forceinline__ void bilinear_level(const void* __restrict__ texels, const float* __restrict__ lut, uint32_t size, float u, float v, float (&out)[4]) {
    const float px = u * (float)size - 0.5f, py = v * (float)size - 0.5f;
    const float fx = floorf(px), fy = floorf(py);
    const float rx = px - fx, ry = py - fy;
    const int ix = (int)fx, iy = (int)fy, m = (int)size - 1;
    const int x0 = min(max(ix, 0), m), x1 = min(max(ix + 1, 0), m), y0 = min(max(iy, 0), m), y1 = min(max(iy + 1, 0), m);
    if (!RGBA) {
        const uint16_t* t = static_cast<const uint16_t*>(texels);
        const float v00 = (float)__ldg(t + y0 * size + x0) / 65535.0f, v01 = (float)__ldg(t + y1 * size + x0) / 65535.0f;
        const float v10 = (float)__ldg(t + y0 * size + x1) / 65535.0f, v11 = (float)__ldg(t + y1 * size + x1) / 65535.0f;
        const float l0 = v00 + (v01 - v00) * ry, l1 = v10 + (v11 - v10) * ry;
        out[0] = l0 + (l1 - l0) * rx;
    } else {
        const uchar4* t = static_cast<const uchar4*>(texels);
        const uchar4 q00 = __ldg(t + y0 * size + x0), q01 = __ldg(t + y1 * size + x0), q10 = __ldg(t + y0 * size + x1), q11 = __ldg(t + y1 * size + x1);
        const unsigned char c00[4] = {q00.x, q00.y, q00.z, q00.w}, c01[4] = {q01.x, q01.y, q01.z, q01.w};
        const unsigned char c10[4] = {q10.x, q10.y, q10.z, q10.w}, c11[4] = {q11.x, q11.y, q11.z, q11.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float v00 = c < 3 ? __ldg(lut + c00[c]) : (float)c00[c] / 255.0f, v01 = c < 3 ? __ldg(lut + c01[c]) : (float)c01[c] / 255.0f;
            const float v10 = c < 3 ? __ldg(lut + c10[c]) : (float)c10[c] / 255.0f, v11 = c < 3 ? __ldg(lut + c11[c]) : (float)c11[c] / 255.0f;
            const float l0 = v00 + (v01 - v00) * ry, l1 = v10 + (v11 - v10) * ry;
            out[c] = l0 + (l1 - l0) * rx;
        }
    }
}

template <bool RGBA>
__device__ __noinline__ void sample_filtered(const FilterCtx& fc, float u, float v, float step_u, float step_v, float (&out)[4]) {
    constexpr int NCH = RGBA ? 4 : 1;
    if (!fc.trilinear) { bilinear_level<RGBA>(fc.mip0, fc.lut, fc.T, u, v, out); return; }
    const float Px = step_u * (float)fc.T, Py = step_v * (float)fc.T;
    const float Pmax = fmaxf(Px, Py), Pmin = fminf(Px, Py);
    const float ratio = ceilf(Pmax / Pmin);
    const uint32_t N = (ratio >= 16.0f) ? 16u : (ratio >= 1.0f ? (uint32_t)ratio : 1u);   // NaN (0 / 0) -> 1
    const float lambda = log2_det(Pmax / (float)N);
    const float d = fminf(fmaxf(lambda, 0.0f), (float)(fc.mip_level_count - 1u));
    const uint32_t lo = (uint32_t)floorf(d), hi = lo + 1u < fc.mip_level_count ? lo + 1u : lo;
    const float w = d - (float)lo;
    const bool x_major = Px >= Py;
    constexpr uint32_t PX = RGBA ? 4u : 2u;
    const unsigned char* base = static_cast<const unsigned char*>(fc.mips);
    const void* lo_texels = lo == 0u ? fc.mip0 : static_cast<const void*>(base + (size_t)mip_offset_texels(fc.T, lo) * PX);
    const void* hi_texels = base + (size_t)mip_offset_texels(fc.T, hi) * PX;
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (uint32_t i = 0; i < N; ++i) {
        const float t = ((float)i + 0.5f) / (float)N - 0.5f;
        const float uu = x_major ? u + t * step_u : u, vv = x_major ? v : v + t * step_v;
        float a[4], b[4];
        bilinear_level<RGBA>(lo_texels, fc.lut, fc.T >> lo, uu, vv, a);
        if (hi != lo) bilinear_level<RGBA>(hi_texels, fc.lut, fc.T >> hi, uu, vv, b);
#pragma unroll
        for (int c = 0; c < NCH; ++c) acc[c] += (hi != lo) ? a[c] * (1.0f - w) + b[c] * w : a[c];
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) out[c] = acc[c] / (float)N;
}

struct FilteredVertex { Record r; uint32_t albedo; };

// the sampling half of the vertex with the configs[2] sampler (vertex_front supplies position, morphed uv, blend LOD / ratio)
template <bool SPHERICAL>
__device__ __noinline__ FilteredVertex vertex_back_filtered(const ViewUniforms& U, const FrameConsts& F, const GenParams& P, uint32_t flags,
                                                            const bt_tile& tile, uint32_t tile_index, const Front& f) {
    const Vec3 wp = f.wp, wn = f.wn;
    const uint32_t tile_texels = (uint32_t)(F.T * F.T);
    const uint32_t aa = P.albedo_attachment;
    float height = 0.0f, color[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    Vec3 nrm = wn;
    const uint32_t passes = f.blend_ratio > 0.0f ? 2u : 1u;
    for (uint32_t pass = 0; pass < passes; ++pass) {
        const AtlasTile t = lookup_tile(F, P.tile_tree, Coord{tile.side, tile.lod, tile.x, tile.y, f.cu, f.cv}, f.blend_lod - pass);
        const uint32_t layer = min(t.index, F.last_layer);
        const float u = t.c.u * F.a_scale + F.a_offset, v = t.c.v * F.a_scale + F.a_offset;
        const float inv_g = 1.0f / F.g;
        const float rel = exp2i((int)t.c.lod - (int)tile.lod);
        const float step = (inv_g * rel) * F.a_scale;
        FilterCtx fh{P.height_atlas + (size_t)layer * tile_texels, P.height_mips ? P.height_mips + (size_t)layer * P.mip_chain_texels : nullptr, nullptr,
                     (uint32_t)F.T, P.mip_level_count, (flags & BT_FLAG_FILTER_TRILINEAR_ANISO) != 0u};
        float s[4];
        sample_filtered<false>(fh, u, v, step, step, s);
        const float h = mix(F.min_h, F.max_h, s[0]);
        Vec3 n2 = wn;
        if (flags & BT_FLAG_NORMALS) {
            const float dbs = F.side_length / (F.size0 * tile_count(t.c.lod));
            const float o = F.normal_offset;
            float taps[4];
            const float du[4] = {-o, 0.0f, o, 0.0f}, dv[4] = {0.0f, -o, 0.0f, o};   // left, up, right, down
            for (int k = 0; k < 4; ++k) {
                sample_filtered<false>(fh, u + du[k], v + dv[k], step, step, s);
                taps[k] = mix(F.min_h, F.max_h, s[0]);
            }
            const Vec3 sn = normalize3(Vec3{taps[0] - taps[2], taps[3] - taps[1], dbs});
            Vec3 r;
            if (SPHERICAL) {
                Vec3 face_up;
                switch (t.c.side) {
                    case 0: case 1: face_up = Vec3{0.0f, 1.0f, 0.0f}; break;
                    case 2: case 3: face_up = Vec3{0.0f, 0.0f, -1.0f}; break;
                    default: face_up = Vec3{-1.0f, 0.0f, 0.0f}; break;
                }
                const Vec3 n = normalize3(wn);
                const Vec3 tg = cross3(face_up, n);
                const Vec3 bt = cross3(n, tg);
                r = Vec3{(tg.x * sn.x + bt.x * sn.y) + n.x * sn.z, (tg.y * sn.x + bt.y * sn.y) + n.y * sn.z, (tg.z * sn.x + bt.z * sn.y) + n.z * sn.z};
            } else {
                r = Vec3{sn.x, sn.z, sn.y};
            }
            n2 = normalize3(r);
        }
        float col[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (flags & BT_FLAG_ALBEDO) {
            const float a_scale = U.attachments[aa].scale, a_offset = U.attachments[aa].offset;
            const float ua = t.c.u * a_scale + a_offset, va = t.c.v * a_scale + a_offset;
            const float step_a = (inv_g * rel) * a_scale;
            FilterCtx fa{P.albedo_atlas + (size_t)layer * tile_texels, P.albedo_mips ? P.albedo_mips + (size_t)layer * P.mip_chain_texels : nullptr, P.srgb_lut,
                         (uint32_t)F.T, P.mip_level_count, (flags & BT_FLAG_FILTER_TRILINEAR_ANISO) != 0u};
            sample_filtered<true>(fa, ua, va, step_a, step_a, col);
        }
        if (pass == 0u) {
            height = h;
            nrm = n2;
#pragma unroll
            for (int k = 0; k < 4; ++k) color[k] = col[k];
        } else {
            height = mix(height, h, f.blend_ratio);
            if (flags & BT_FLAG_NORMALS) nrm = Vec3{mix(nrm.x, n2.x, f.blend_ratio), mix(nrm.y, n2.y, f.blend_ratio), mix(nrm.z, n2.z, f.blend_ratio)};
#pragma unroll
            for (int k = 0; k < 4; ++k) color[k] = mix(color[k], col[k], f.blend_ratio);
        }
    }
    FilteredVertex out;
    out.r.w[0] = make_float2(wp.x + height * wn.x, wp.y + height * wn.y);
    out.r.w[1] = make_float2(wp.z + height * wn.z, 1.0f);
    out.r.w[2] = make_float2(nrm.x, nrm.y);
    out.r.w[3] = make_float2(nrm.z, f.cu);
    out.r.w[4] = make_float2(f.cv, __uint_as_float(tile_index));
    uint32_t packed = 0;   // pack4x8unorm of the linear colour
#pragma unroll
    for (int k = 0; k < 4; ++k) packed |= (uint32_t)floorf(0.5f + 255.0f * saturate(color[k])) << (8 * k);
    out.albedo = packed;
    return out;
}

struct GenSmemHeader {
    ViewUniforms U;
    uint64_t bar;                   // uniform-block TMA load
    uint64_t full[GEN_BUFFERS];     // image b completely written by the compute warps (count = compute warps)
    uint64_t empty[GEN_BUFFERS];    // image b completely read by the TMA stores (count = 1, the store warp)
    uint64_t tab_ready[GEN_TABLE_SLOTS];   // tables of slot s written by the auxiliary warp (count = 1)
};

// Warp-specialised pipeline, one CTA per SM:
//   * GEN_COMPUTE_WARPS warps compute vertices and scatter them into output image (it & 1); there is NO CTA-wide barrier
//     on the steady-state path: a warp that has finished its chunks of item i arrives on full[b] and starts on its
//     chunks of item i+1 in the other image, so the ragged last round of an item (1105 vertices = 34.5 warp-chunks) is
//     filled with the next item's work
//   * one auxiliary warp builds the per-item tables (above) GEN_TABLE_LEAD items ahead and, per item, waits on full[b], issues the
//     item's TMA bulk stores, waits until they have READ the image (cp.async.bulk.wait_group.read) and arrives on empty[b],
//     which the compute warps wait on before reusing image b
// FLAGS_CT == -2: generic flags + the configs[2] sampler / albedo stream (vertex_back_filtered).
// FLAGS_CT >= 0: the flag word is a compile-time constant (the common sets get their own instantiation, so the flag
// branches and the code of the disabled features disappear); FLAGS_CT < 0: generic kernel, flags read from the parameters.
template <bool SPHERICAL, int FLAGS_CT>
__global__ void __launch_bounds__(GEN_THREADS, 1) gen_vertices_kernel(const GenParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    GenSmemHeader& H = *reinterpret_cast<GenSmemHeader*>(smem_raw);
    unsigned char* tables = smem_raw + ((sizeof(GenSmemHeader) + 127) & ~size_t(127));
    unsigned char* images = tables + (size_t)GEN_TABLE_SLOTS * P.table_bytes;
    unsigned char* aimages = images + (size_t)GEN_BUFFERS * P.image_bytes;      // albedo stream images (BT_FLAG_ALBEDO only)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&H.bar, 1);
        for (int b = 0; b < GEN_BUFFERS; ++b) { mbar_init(&H.full[b], GEN_COMPUTE_WARPS); mbar_init(&H.empty[b], 1); }
        for (int s = 0; s < GEN_TABLE_SLOTS; ++s) mbar_init(&H.tab_ready[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&H.bar, (uint32_t)sizeof(ViewUniforms));
        bulk_load(&H.U, P.uniforms, (uint32_t)sizeof(ViewUniforms), &H.bar);
    }
    mbar_wait_cta(&H.bar, 0);
    const ViewUniforms& U = H.U;

    const uint32_t g = P.grid_size;
    const uint32_t vpr = U.view_config.vertices_per_row, vpt = U.view_config.vertices_per_tile;
    const uint32_t strip_bytes = vpr * (uint32_t)sizeof(bt_vertex);
    const uint32_t strip_stride = P.strip_stride;
    uint32_t n_tiles = __ldg(&P.indirect->vertex_count) / vpt;
    const uint64_t cap_tiles = P.vertex_capacity / vpt;
    if (n_tiles > cap_tiles) {
        if (blockIdx.x == 0 && tid == 0) P.status->overflow = 1u;
        n_tiles = (uint32_t)cap_tiles;
    } else if (blockIdx.x == 0 && tid == 0) {
        P.status->overflow = 0u;
    }
    if (blockIdx.x == 0 && tid == 0) {
        const uint64_t n = (uint64_t)n_tiles * vpt;
        P.status->vertex_count_lo = (uint32_t)n;
        P.status->vertex_count_hi = (uint32_t)(n >> 32);
    }
    const uint32_t n_items = n_tiles * P.items_per_tile;

    const FrameConsts F = make_frame_consts<SPHERICAL>(U);

    if (warp == GEN_COMPUTE_WARPS) {
        // ---------------- auxiliary warp: TMA stores (lane 0) and per-item tables (lanes 1-31), two independent loops ----------------
        const uint32_t n_local = blockIdx.x < n_items ? (n_items - blockIdx.x + gridDim.x - 1u) / gridDim.x : 0u;
        if (lane != 0) {
            // Builder lanes.  The tables of item `it` go to slot it % GEN_TABLE_SLOTS, which last held item it - 2 * GEN_BUFFERS; they are
            // built as soon as full[] of item it - GEN_BUFFERS has completed (every compute warp is then past that item), i.e. before
            // the image of item `it` can even be free.  (full[] of item it cannot complete before its tables exist: no parity aliasing.)
            const int bl = lane - 1;
            for (uint32_t it = 0; it < n_local; ++it) {
                if (it >= (uint32_t)GEN_BUFFERS) mbar_wait_parked(&H.full[it % GEN_BUFFERS], ((it - GEN_BUFFERS) / GEN_BUFFERS) & 1u, 128u);
                const uint32_t w = blockIdx.x + it * gridDim.x;
                const uint32_t tile_index = P.items_per_tile == 1u ? w : __umulhi(w, P.magic_items);
                const uint32_t c0 = (w - tile_index * P.items_per_tile) * P.strips_per_item;
                const uint32_t n_strips = min(P.strips_per_item, g - c0);
                const uint4 tq = __ldg(reinterpret_cast<const uint4*>(P.tiles) + tile_index);
                const bt_tile tile{tq.x, tq.y, tq.z, tq.w};
                ItemTables& T = *reinterpret_cast<ItemTables*>(tables + (size_t)(it % GEN_TABLE_SLOTS) * P.table_bytes);
                float4* rows = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(&T) + sizeof(ItemTables));
                build_item_tables<SPHERICAL>(U, F, P.tile_tree, T, rows, tile, tile_index, n_strips == P.strips_per_item ? P.magic_cols : P.magic_cols_ragged,
                                             c0, n_strips + 1u, g + 1u, bl);
                __syncwarp(0xFFFFFFFEu);
                if (bl == 0) mbar_arrive(&H.tab_ready[it % GEN_TABLE_SLOTS]);
            }
            return;
        }
        // Store scheduling (lane 0), adaptive so that BOTH regimes run without bubbles:
        //   * an image is handed back as soon as the TMA has read it (the compute warps may then start the item after next), and
        //   * if the NEXT item's image is already full at that point in time, its stores are issued FIRST, so that two bulk groups are
        //     in flight and the DRAM write stream never waits for the ~1 us TMA start-up of the following item.
        // issued / released count this CTA's items; item k lives in image k % GEN_BUFFERS; at most GEN_BUFFERS groups are in flight.
        uint32_t issued = 0, released = 0;
        while (released < n_local) {
            const uint32_t in_flight = issued - released;
            bool can_issue = issued < n_local && in_flight < (uint32_t)GEN_BUFFERS;
            if (can_issue && in_flight != 0u)       // something is draining: only issue the next item if its image is full NOW
                can_issue = mbar_try_wait(&H.full[issued % GEN_BUFFERS], (issued / GEN_BUFFERS) & 1u);
            if (can_issue) {
                const uint32_t it = issued, b = it % GEN_BUFFERS, use = it / GEN_BUFFERS;
                const uint32_t w = blockIdx.x + it * gridDim.x;
                const uint32_t tile_index = P.items_per_tile == 1u ? w : __umulhi(w, P.magic_items);
                const uint32_t c0 = (w - tile_index * P.items_per_tile) * P.strips_per_item;
                const uint32_t n_strips = min(P.strips_per_item, g - c0);
                unsigned char* image = images + (size_t)b * P.image_bytes;
                mbar_wait_parked(&H.full[b], use & 1u, BT_GEN_FULL_SLEEP_NS);   // every compute warp has written (and proxy-fenced) its records
                fence_proxy_async();
                unsigned char* out = reinterpret_cast<unsigned char*>(P.vertices) +
                                     ((size_t)tile_index * vpt + (size_t)c0 * vpr) * sizeof(bt_vertex);
#ifndef BT_GEN_SKIP_STORE   // (compute-only measurement variant of tools/gen_probe.py: NOT the product path)
                for (uint32_t s = 0; s < n_strips; ++s)
                    bulk_store(out + (size_t)s * strip_bytes, image + (size_t)s * strip_stride, strip_bytes);
                if (FLAGS_CT == -2 && P.albedo_image_bytes != 0u) {    // the 4-byte-per-vertex albedo stream of the same strips
                    const unsigned char* aimage = aimages + (size_t)b * P.albedo_image_bytes;
                    unsigned char* aout = reinterpret_cast<unsigned char*>(P.albedo_out) + ((size_t)tile_index * vpt + (size_t)c0 * vpr) * sizeof(uint32_t);
                    for (uint32_t s = 0; s < n_strips; ++s)
                        bulk_store(aout + (size_t)s * P.albedo_stride, aimage + (size_t)s * P.albedo_stride, P.albedo_stride);
                }
#endif
                bulk_commit();
                ++issued;
            } else {
                // the oldest group in flight: wait until the TMA has READ its image, then hand the image back
                if (in_flight >= 3u) bulk_wait_read<2>(); else if (in_flight == 2u) bulk_wait_read<1>(); else bulk_wait_read<0>();
                static_assert(GEN_BUFFERS <= 3, "the wait_group.read immediates above cover three groups in flight");
                mbar_arrive(&H.empty[released % GEN_BUFFERS]);
                ++released;
            }
        }
        bulk_wait<0>();                                   // all global writes complete before the CTA exits
        return;
    }

    // ---------------- compute warps ----------------
    uint32_t it = 0, rot = 0;
    for (uint32_t w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const uint32_t b = it % GEN_BUFFERS, use = it / GEN_BUFFERS;
        const uint32_t ts = it % GEN_TABLE_SLOTS, tuse = it / GEN_TABLE_SLOTS;
        unsigned char* image = images + (size_t)b * P.image_bytes;
        const ItemTables& T = *reinterpret_cast<const ItemTables*>(tables + (size_t)ts * P.table_bytes);
        const float4* rows = reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(&T) + sizeof(ItemTables));
        mbar_wait_parked(&H.tab_ready[ts], tuse & 1u, 32u);                       // this item's tables are written
        if (use > 0u) mbar_wait_parked(&H.empty[b], (use - 1u) & 1u, BT_GEN_EMPTY_SLEEP_NS);        // the stores of the previous user of image b have read it

        // the item's scalars come from its tables (computed once by the builder lanes instead of by every compute warp)
        const uint4 tq = T.tile, mq = T.meta;
        const bt_tile tile{tq.x, tq.y, tq.z, tq.w};
        const uint32_t tile_index = mq.x, n_strips = mq.y, magic = mq.z;
        const float morph_zero_from = __uint_as_float(mq.w);
        const uint32_t n_cols = n_strips + 1u, n_rows = g + 1u;
        const uint32_t n_verts = n_cols * n_rows;
        const uint32_t n_chunks = (n_verts + 31u) >> 5;
        uint32_t first = (uint32_t)warp + GEN_COMPUTE_WARPS - rot;   // chunk c of this item goes to warp (rot + c) mod GEN_COMPUTE_WARPS
        if (first >= GEN_COMPUTE_WARPS) first -= GEN_COMPUTE_WARPS;
        rot += GEN_ROTATION % GEN_COMPUTE_WARPS;                     // spreads the ragged tail (measured: DESIGN.md section 6)
        if (rot >= GEN_COMPUTE_WARPS) rot -= GEN_COMPUTE_WARPS;
        const uint32_t flags = FLAGS_CT >= 0 ? (uint32_t)FLAGS_CT : P.flags;
        // scatter of one computed vertex: strip `lx` holds it as (column, row 2*gy) -> record 2*gy+1, strip `lx-1` as (column+1,
        // row 2*gy+1) -> record 2*gy+2; the duplicated first / last index of a strip (degenerate triangles) only concerns warps that
        // touch row 0 or row g (full-warp votes: every lane computes, lanes past the end of a ragged chunk store nothing)
        auto scatter = [&](const Record& r, uint32_t lx, uint32_t gy, bool live) {
            unsigned char* d1 = image + lx * strip_stride + (2u * gy + 1u) * (uint32_t)sizeof(bt_vertex);
            unsigned char* d2 = d1 - strip_stride + (uint32_t)sizeof(bt_vertex);
            const bool in1 = live && lx < n_strips, in2 = live && lx >= 1u;
            if (in1) store_record_at<1>(d1, r);
            if (in2) store_record_at<0>(d2, r);
            if (__any_sync(0xffffffffu, gy == 0u) && in1 && gy == 0u) store_record_at<0>(d1 - sizeof(bt_vertex), r);
            if (__any_sync(0xffffffffu, gy == g) && in2 && gy == g) store_record_at<1>(d2 + sizeof(bt_vertex), r);   // record vpr - 1 (vpr is even)
        };
#if defined(BT_GEN_SKIP_COMPUTE)   // store-pipeline floor measurement only (tools/gen_probe.py): NOT the product path
        for (uint32_t c = first; c < n_chunks; c += GEN_COMPUTE_WARPS) {
            const uint32_t i_raw = (c << 5) + (uint32_t)lane;
            const bool live = i_raw < n_verts;
            const uint32_t i = live ? i_raw : n_verts - 1u;
            const uint32_t gy = __umulhi(i, magic), lx = i - gy * n_cols;
            Record r;
            r.w[0] = make_float2((float)lx, (float)gy); r.w[1] = make_float2(0.0f, 1.0f); r.w[2] = r.w[0]; r.w[3] = r.w[1];
            r.w[4] = make_float2(0.0f, __uint_as_float(tile_index));
            scatter(r, lx, gy, live);
        }
#else
        uint32_t c = first;
#if defined(BT_GEN_PAIRS)   // (experiment, off: measured no faster for NORMALS and slower, through spills, with MORPH + BLEND)
        // TWO chunks per iteration (a warp owns chunks first, first + NW, ...: usually exactly two per item): both fronts, then both
        // sets of texel loads, then the backs — two independent vertices per lane in flight (the kernel is latency-bound)
        if ((flags & BT_FLAG_NORMALS) != 0u) {
            for (; c + GEN_COMPUTE_WARPS < n_chunks; c += 2u * GEN_COMPUTE_WARPS) {
                const uint32_t ia = (c << 5) + (uint32_t)lane;                       // chunk c is never the ragged last one here
                const uint32_t ib_raw = ((c + GEN_COMPUTE_WARPS) << 5) + (uint32_t)lane;
                const bool live_b = ib_raw < n_verts;
                const uint32_t ib = live_b ? ib_raw : n_verts - 1u;
                const uint32_t gya = __umulhi(ia, magic), lxa = ia - gya * n_cols;
                const uint32_t gyb = __umulhi(ib, magic), lxb = ib - gyb * n_cols;
                const Front fa = vertex_front<SPHERICAL>(U, F, T, rows, P.tile_tree, P.height_atlas, flags, tile, morph_zero_from, lxa, gya);
                const Front fb = vertex_front<SPHERICAL>(U, F, T, rows, P.tile_tree, P.height_atlas, flags, tile, morph_zero_from, lxb, gyb);
                TapTexels ta, tb;
                if (__all_sync(0xffffffffu, fa.taps.interior && fb.taps.interior)) {
                    ta = load_taps_interior(F, fa.taps);
                    tb = load_taps_interior(F, fb.taps);
                } else {
                    ta = fa.taps.interior ? load_taps_interior(F, fa.taps) : load_taps_clamped(F, fa.taps);
                    tb = fb.taps.interior ? load_taps_interior(F, fb.taps) : load_taps_clamped(F, fb.taps);
                }
                const Record ra = vertex_back<SPHERICAL>(F, T, P.tile_tree, P.height_atlas, flags, tile, tile_index, fa, &ta);
                scatter(ra, lxa, gya, true);
                const Record rb = vertex_back<SPHERICAL>(F, T, P.tile_tree, P.height_atlas, flags, tile, tile_index, fb, &tb);
                scatter(rb, lxb, gyb, live_b);
            }
        }
#endif
        for (; c < n_chunks; c += GEN_COMPUTE_WARPS) {
            // every lane computes a vertex (the lanes past the end of a ragged last chunk recompute the last one and store nothing)
            const uint32_t i_raw = (c << 5) + (uint32_t)lane;
            const bool live = i_raw < n_verts;
            const uint32_t i = live ? i_raw : n_verts - 1u;
            const uint32_t gy = __umulhi(i, magic), lx = i - gy * n_cols;  // lanes walk along texel rows
            const Front f = vertex_front<SPHERICAL>(U, F, T, rows, P.tile_tree, P.height_atlas, flags, tile, morph_zero_from, lx, gy);
            if (FLAGS_CT == -2) {
                // configs[2] variant: trilinear + anisotropic sampler and / or the per-vertex albedo stream
                const FilteredVertex fv = vertex_back_filtered<SPHERICAL>(U, F, P, flags, tile, tile_index, f);
                scatter(fv.r, lx, gy, live);
                if (P.albedo_image_bytes != 0u) {
                    unsigned char* aimage = aimages + (size_t)b * P.albedo_image_bytes;
                    uint32_t* a1 = reinterpret_cast<uint32_t*>(aimage + lx * P.albedo_stride) + (2u * gy + 1u);
                    uint32_t* a2 = reinterpret_cast<uint32_t*>(aimage + (lx - 1u) * P.albedo_stride) + (2u * gy + 2u);
                    if (live && lx < n_strips) { *a1 = fv.albedo; if (gy == 0u) a1[-1] = fv.albedo; }
                    if (live && lx >= 1u) { *a2 = fv.albedo; if (gy == g) a2[1] = fv.albedo; }
                }
                continue;
            }
            TapTexels t;
            if (flags & BT_FLAG_NORMALS) t = f.taps.interior ? load_taps_interior(F, f.taps) : load_taps_clamped(F, f.taps);
            const Record r = vertex_back<SPHERICAL>(F, T, P.tile_tree, P.height_atlas, flags, tile, tile_index, f, &t);
            scatter(r, lx, gy, live);
        }
#endif
        fence_proxy_async();   // this lane's generic-proxy smem writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&H.full[b]);
    }
}

}  // namespace btd

extern "C" cudaError_t bt_launch_gen_vertices(const void* uniforms, const bt_tile* final_tiles, const bt_indirect* indirect,
                                              const bt_tile_tree_entry* tile_tree, const uint16_t* height_atlas,
                                              bt_vertex* vertices, uint64_t vertex_capacity, btd::GenStatus* status,
                                              int spherical, uint32_t flags, uint32_t grid_size, int sm_count,
                                              const btd::GenFilterInputs* filter, cudaStream_t stream) {
    using namespace btd;
    if (grid_size == 0) return cudaErrorInvalidValue;
    const uint32_t strip_stride = strip_stride_for(2u * (grid_size + 2u) * (uint32_t)sizeof(bt_vertex));
    const size_t header = (sizeof(GenSmemHeader) + 127) & ~size_t(127);
    const size_t table_bytes = (sizeof(ItemTables) + (size_t)(grid_size + 1u) * sizeof(float4) + 127) & ~size_t(127);
#ifndef BT_GEN_SMEM_KB
#define BT_GEN_SMEM_KB 208
#endif
    if (header + GEN_TABLE_SLOTS * table_bytes + 2048 > (size_t)BT_GEN_SMEM_KB * 1024) return cudaErrorInvalidValue;   // grid_size too large for the row tables
    const size_t budget = ((size_t)BT_GEN_SMEM_KB * 1024) / GEN_CTAS_PER_SM - header - GEN_TABLE_SLOTS * table_bytes - 1024;  // per CTA; leaves L1 some room out of the 227 KiB
    uint32_t cb = (uint32_t)(budget / ((size_t)GEN_BUFFERS * strip_stride));
    if (cb > (uint32_t)BT_GEN_MAX_CB) cb = BT_GEN_MAX_CB;
    if (cb + 1u > (uint32_t)GEN_MAX_COLS) cb = GEN_MAX_COLS - 1u;
    if (cb > grid_size) cb = grid_size;
    if (cb == 0) return cudaErrorInvalidValue;  // grid_size too large for one strip per CTA
    GenParams P;
    P.uniforms = static_cast<const ViewUniforms*>(uniforms);
    P.tiles = final_tiles;
    P.indirect = indirect;
    P.tile_tree = tile_tree;
    P.height_atlas = height_atlas;
    P.vertices = vertices;
    P.vertex_capacity = vertex_capacity;
    P.status = status;
    P.flags = flags;
    P.grid_size = grid_size;
    P.strips_per_item = cb;
    P.items_per_tile = (grid_size + cb - 1u) / cb;
    P.image_bytes = cb * strip_stride;
    P.strip_stride = strip_stride;
    P.magic_items = P.items_per_tile == 1u ? 0u : 0xFFFFFFFFu / P.items_per_tile + 1u;
    P.magic_cols = 0xFFFFFFFFu / (cb + 1u) + 1u;
    {
        const uint32_t last_strips = grid_size - (P.items_per_tile - 1u) * cb;   // 1..cb
        P.magic_cols_ragged = 0xFFFFFFFFu / (last_strips + 1u) + 1u;
    }
    P.table_bytes = (uint32_t)table_bytes;
    P.height_mips = nullptr; P.albedo_atlas = nullptr; P.albedo_mips = nullptr; P.srgb_lut = nullptr; P.albedo_out = nullptr;
    P.mip_level_count = 1u; P.mip_chain_texels = 0u; P.albedo_attachment = 1u; P.albedo_image_bytes = 0u; P.albedo_stride = 0u;
    if (flags & (BT_FLAG_FILTER_TRILINEAR_ANISO | BT_FLAG_ALBEDO)) {
        if (!filter) return cudaErrorInvalidValue;
        P.height_mips = filter->height_mips;
        P.albedo_atlas = static_cast<const uchar4*>(filter->albedo_atlas);
        P.albedo_mips = static_cast<const uchar4*>(filter->albedo_mips);
        P.srgb_lut = filter->srgb_lut;
        P.albedo_out = filter->albedo_out;
        P.mip_level_count = filter->mip_level_count;
        P.mip_chain_texels = filter->mip_chain_texels;
        P.albedo_attachment = filter->albedo_attachment;
        if (flags & BT_FLAG_ALBEDO) {
            P.albedo_stride = 2u * (grid_size + 2u) * (uint32_t)sizeof(uint32_t);
            P.albedo_image_bytes = cb * P.albedo_stride;
        }
    }
    const size_t smem = header + GEN_TABLE_SLOTS * table_bytes + (size_t)GEN_BUFFERS * (P.image_bytes + P.albedo_image_bytes);
    if (smem > (size_t)227 * 1024) return cudaErrorInvalidValue;
    // kernels specialised for the common flag sets (measured: 6 % faster than the generic kernel for NORMALS, 3 % for the
    // reference default + NORMALS); everything else (HIGH_PRECISION, partial sets) takes the generic one
    constexpr int kNone = 0, kNormals = BT_FLAG_NORMALS, kMorphBlend = BT_FLAG_MORPH | BT_FLAG_BLEND, kDefault = kMorphBlend | BT_FLAG_NORMALS;
    const uint32_t vertex_flags = flags & (BT_FLAG_MORPH | BT_FLAG_BLEND | BT_FLAG_HIGH_PRECISION | BT_FLAG_NORMALS | BT_FLAG_FILTER_TRILINEAR_ANISO | BT_FLAG_ALBEDO);
    void (*kernel)(const GenParams);
    switch (vertex_flags) {
        case kNone: kernel = spherical ? gen_vertices_kernel<true, kNone> : gen_vertices_kernel<false, kNone>; break;
        case kNormals: kernel = spherical ? gen_vertices_kernel<true, kNormals> : gen_vertices_kernel<false, kNormals>; break;
        case kMorphBlend: kernel = spherical ? gen_vertices_kernel<true, kMorphBlend> : gen_vertices_kernel<false, kMorphBlend>; break;
        case kDefault: kernel = spherical ? gen_vertices_kernel<true, kDefault> : gen_vertices_kernel<false, kDefault>; break;
        default: kernel = spherical ? gen_vertices_kernel<true, -1> : gen_vertices_kernel<false, -1>; break;
    }
    if (flags & (BT_FLAG_FILTER_TRILINEAR_ANISO | BT_FLAG_ALBEDO))    // the configs[2] sampler / albedo variant: its own generic instantiation
        kernel = spherical ? gen_vertices_kernel<true, -2> : gen_vertices_kernel<false, -2>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (sm_count > 0 ? sm_count : 148) * GEN_CTAS_PER_SM;
    kernel<<<grid, GEN_THREADS, smem, stream>>>(P);
    return cudaGetLastError();
}
