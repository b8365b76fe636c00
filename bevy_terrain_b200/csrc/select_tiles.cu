// select_tiles.cu — per-frame quadtree refinement + tile-list / indirect-draw emit, ONE launch.
//
// Replaces the reference's 2*refinement_count+3 = 63 compute dispatches per view
// (TilingPrepassNode::run, src/render/tiling_prepass.rs:248-263) that run prepare_root /
// refine_tiles / prepare_next / prepare_render (src/shaders/tiling_prepass/*.wgsl):
//
//   * one persistent CTA per view; the level loop runs in-kernel
//   * the uniform block (view config, Taylor approximation, culling, mesh) is staged into shared
//     memory with one TMA 1-D bulk copy (cp.async.bulk + mbarrier)
//   * the frontier ping-pongs between two shared-memory buffers and spills to the global
//     `temporary_tiles` buffer only when a level could exceed them
//   * warp-ballot + block prefix-sum compaction replaces the two global atomics
//     (refine_tiles.wgsl:5-15); the emission order is the order a sequential execution of the
//     reference shader (invocations in id order) produces, so results compare bit-exactly AND
//     in order, not only as sets
//   * capacity overflow (unchecked in the reference) is detected and reported
#include "bt_device.cuh"
#include "bt_kernels.h"

namespace btd {

constexpr int SELECT_THREADS = 1024;
constexpr int SELECT_WARPS = SELECT_THREADS / 32;
constexpr uint32_t SELECT_SMEM_TILES = 4096;  // per frontier buffer (2 x 64 KiB)

// compute_subdivision_coordinate + approximate_view_distance + should_be_divided
// (functions.wgsl:133-154,117-131; refine_tiles.wgsl:17-22).  HIGH_PRECISION is compiled out in the
// prepass (tiling_prepass.rs:61-78 only sets SPHERICAL/TEST*).
// `planar_normal`: normal_local_to_world of a planar terrain does not depend on the tile (normalize(normal_matrix * +Y)); the
// caller evaluates it once per thread, which takes a 40-instruction dependent chain (sqrt, reciprocal, three quotients) out of
// every level of this latency-bound kernel.  Same operations on the same operands, so the same bits.
template <bool SPHERICAL>
__device__ __forceinline__ bool should_be_divided(const ViewUniforms& U, const bt_tile& tile, const Vec3& planar_normal) {
    const bt_side_parameter& p = U.approximation.sides[tile.side];
    Coord v{tile.side, U.approximation.origin_lod, (uint32_t)p.view_xy[0], (uint32_t)p.view_xy[1], p.view_uv[0], p.view_uv[1]};
    change_lod(v, tile.lod);
    int ox = (int)v.x - (int)tile.x;
    int oy = (int)v.y - (int)tile.y;
    float u = v.u, w = v.v;
    if (ox < 0) u = 0.0f; else if (ox > 0) u = 1.0f;
    if (oy < 0) w = 0.0f; else if (oy > 0) w = 1.0f;
    Coord c{tile.side, tile.lod, tile.x, tile.y, u, w};
    Vec3 lp = local_position<SPHERICAL>(c);
    Vec3 wp = position_local_to_world(U.mesh, lp);
    Vec3 wn = SPHERICAL ? normal_local_to_world<SPHERICAL>(U.mesh, lp) : planar_normal;
    float h = U.approximation.approximate_height;
    const float* e = U.culling.world_position;
    Vec3 q{(wp.x + h * wn.x) - e[0], (wp.y + h * wn.y) - e[1], (wp.z + h * wn.z) - e[2]};
    float d = length3(q);
    return d < U.view_config.subdivision_distance * exp2i(-(int)(tile.lod & 31u));
}

// EXTENSION (BT_FLAG_FRUSTUM_CULL, default off — the reference does not cull at this commit): ten bounding
// points (four corners + centre, at min_height and max_height) against the five planes of culling_bind_group.rs:25-38;
// culled when all ten lie on the negative side of one plane.  Same arithmetic as the oracle's tile_outside_frustum.
// EXTENSION (BT_FLAG_MINMAX_CULL, default off): the height range of the bounding points comes from the best resident atlas tile that
// contains the geometry tile (docs/implementation.md:69: "a 2D min-max height data attachment to approximate the bounding volumes"):
// the (min, max) table of bt_atlas_update_minmax — whole tile + 8 x 8 cells — looked up through the view's tile-tree table.
// Same arithmetic as the oracle's tile_height_bounds.
constexpr uint32_t MINMAX_CELLS = 8u, MINMAX_ENTRIES = 1u + MINMAX_CELLS * MINMAX_CELLS;
__device__ __forceinline__ void tile_height_bounds(const ViewUniforms& U, const SelectViewParams& V, const bt_tile& tile, float& hmin, float& hmax) {
    hmin = U.terrain_config.min_height;
    hmax = U.terrain_config.max_height;
    if (!V.minmax || !V.tile_tree) return;
    const uint32_t lod_count = U.terrain_config.lod_count, ts = U.view_config.tree_size;
    Coord c{tile.side, tile.lod, tile.x, tile.y, 0.5f, 0.5f};
    const uint32_t L = min(tile.lod, lod_count - 1u);
    change_lod(c, L);
    const uint2 e = __ldg(reinterpret_cast<const uint2*>(V.tile_tree) + (((c.side * lod_count + c.lod) * ts + c.x % ts) * ts + c.y % ts));
    if (e.x >= V.atlas_size || e.y > L) return;            // nothing resident (INVALID) or a malformed entry
    change_lod(c, e.y);
    uint32_t entry = 0u;
    if (tile.lod - e.y >= 3u) {
        const uint32_t cx = min(f2u(c.u * (float)MINMAX_CELLS), MINMAX_CELLS - 1u), cy = min(f2u(c.v * (float)MINMAX_CELLS), MINMAX_CELLS - 1u);
        entry = 1u + cy * MINMAX_CELLS + cx;
    }
    const ushort2 mm = __ldg(reinterpret_cast<const ushort2*>(V.minmax) + ((size_t)e.x * MINMAX_ENTRIES + entry));
    if (mm.y == 0) return;                                 // no valid texel in the region
    hmin = mix(U.terrain_config.min_height, U.terrain_config.max_height, (float)mm.x / 65535.0f);
    hmax = mix(U.terrain_config.min_height, U.terrain_config.max_height, (float)mm.y / 65535.0f);
}

template <bool SPHERICAL>
__device__ __forceinline__ bool tile_outside_frustum(const ViewUniforms& U, const bt_tile& tile, float hmin, float hmax) {
    uint32_t outside = 0x1Fu;  // bit p: all points so far are outside plane p
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float u = i == 4 ? 0.5f : (float)(i & 1), v = i == 4 ? 0.5f : (float)(i >> 1);
        const Coord c{tile.side, tile.lod, tile.x, tile.y, u, v};
        const Vec3 lp = local_position<SPHERICAL>(c);
        const Vec3 wp = position_local_to_world(U.mesh, lp);
        const Vec3 wn = normal_local_to_world<SPHERICAL>(U.mesh, lp);
#pragma unroll
        for (int hsel = 0; hsel < 2; ++hsel) {
            const float h = hsel ? hmax : hmin;
            const Vec3 q{wp.x + h * wn.x, wp.y + h * wn.y, wp.z + h * wn.z};
#pragma unroll
            for (int p = 0; p < 5; ++p) {
                const float* pl = U.culling.planes[p];
                const float dist = ((pl[0] * q.x + pl[1] * q.y) + pl[2] * q.z) + pl[3];
                if (!(dist < 0.0f)) outside &= ~(1u << p);
            }
        }
    }
    return outside != 0u;
}

struct SelectSmem {
    ViewUniforms U;
    uint64_t bar;
    uint32_t warp_div[SELECT_WARPS];
    uint32_t warp_fin[SELECT_WARPS];
    uint32_t total_div, total_fin;
    bt_tile frontier[2][SELECT_SMEM_TILES];
};

// gridDim.x = number of views: CTA v refines view v (the views of a frame are independent; the reference loops over
// (terrain, view) pairs on the host, tiling_prepass.rs:224-263, with 63 dispatches each)
template <bool SPHERICAL>
__global__ void __launch_bounds__(SELECT_THREADS, 1)
select_tiles_kernel(const SelectViewParams* __restrict__ views, uint32_t flags) {
    const SelectViewParams V = views[blockIdx.x];
    const ViewUniforms* __restrict__ uniforms = static_cast<const ViewUniforms*>(V.uniforms);
    bt_tile* __restrict__ final_tiles = V.final_tiles;
    bt_tile* temporary_tiles = V.temporary_tiles;
    const uint32_t capacity = V.capacity;
    bt_indirect* __restrict__ indirect = V.indirect;
    SelectStatus* __restrict__ status = V.status;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SelectSmem& S = *reinterpret_cast<SelectSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        mbar_init(&S.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // bt_view_run_frame: this kernel is the first of the frame and fetches the frame's inputs itself - the uniform block by the TMA load
    // below (source = the pinned staging block, mapped into the device address space), the tile-tree table with volatile loads - and
    // leaves both in device memory for gen_vertices.  That removes two copy-engine operations (and their engine hand-overs) from
    // the frame's critical path; the bytes that cross PCIe are the same.
    const bool fetch_uniforms = (flags & kSelectFetchUniforms) != 0u, fetch_table = (flags & kSelectFetchTable) != 0u;
    if (tid == 0) {
        mbar_expect_tx(&S.bar, (uint32_t)sizeof(ViewUniforms));
        bulk_load(&S.U, fetch_uniforms ? static_cast<const ViewUniforms*>(V.host_uniforms) : uniforms, (uint32_t)sizeof(ViewUniforms), &S.bar);
    }
    if (fetch_table) {
        const uint2* src = static_cast<const uint2*>(V.host_table);
        uint2* dst = reinterpret_cast<uint2*>(const_cast<bt_tile_tree_entry*>(V.tile_tree));
        for (uint32_t i = tid; i < V.table_entries; i += SELECT_THREADS) dst[i] = __ldcv(src + i);
    }
    // prepare_root — prepare_prepass.wgsl:4-23 (overlaps the bulk copy)
    const uint32_t n_roots = SPHERICAL ? 6u : 1u;
    if (tid < (int)n_roots) S.frontier[0][tid] = bt_tile{(uint32_t)tid, 0u, 0u, 0u};
    mbar_wait_cta(&S.bar, 0);   // (its block barrier also orders the table stores above before this CTA's reads of the table)
    if (fetch_uniforms) {
        static_assert(sizeof(ViewUniforms) % 16 == 0, "uniform block is copied in 16-byte pieces");
        uint4* dst = reinterpret_cast<uint4*>(const_cast<void*>(V.uniforms));
        const uint4* src = reinterpret_cast<const uint4*>(&S.U);
        for (uint32_t i = tid; i < (uint32_t)(sizeof(ViewUniforms) / 16); i += SELECT_THREADS) dst[i] = src[i];
    }

    const ViewUniforms& U = S.U;
    const Vec3 planar_normal = SPHERICAL ? Vec3{0.0f, 0.0f, 0.0f} : normal_local_to_world<false>(U.mesh, Vec3{0.0f, 0.0f, 0.0f});
    const uint32_t refinement_count = U.view_config.refinement_count;
    bt_tile* cur = S.frontier[0];
    uint32_t n_cur = n_roots;
    uint32_t n_final = 0, max_frontier = n_roots, overflow = 0, dropped = 0;

    for (uint32_t pass = 0; pass <= refinement_count && n_cur > 0; ++pass) {
        // children of this level go to shared memory when 4*n_cur fits, else to a global half
        bt_tile* nxt;
        uint32_t nxt_cap;
        if (4ull * n_cur <= SELECT_SMEM_TILES) {
            nxt = (cur == S.frontier[0]) ? S.frontier[1] : S.frontier[0];
            nxt_cap = SELECT_SMEM_TILES;
        } else {
            nxt = (cur == temporary_tiles) ? temporary_tiles + capacity : temporary_tiles;
            nxt_cap = capacity;
        }
        uint32_t n_nxt = 0;
        const bool last_pass = (pass == refinement_count);

        for (uint32_t base = 0; base < n_cur; base += SELECT_THREADS) {
            const uint32_t i = base + tid;
            bt_tile tile{0, 0, 0, 0};
            bool valid = i < n_cur, divide = false;
            if (valid) {
                tile = cur[i];
                bool culled = false;
                if (flags & (BT_FLAG_FRUSTUM_CULL | BT_FLAG_MINMAX_CULL)) {
                    float hmin = U.terrain_config.min_height, hmax = U.terrain_config.max_height;
                    if (flags & BT_FLAG_MINMAX_CULL) tile_height_bounds(U, V, tile, hmin, hmax);
                    culled = tile_outside_frustum<SPHERICAL>(U, tile, hmin, hmax);
                }
                if (culled) valid = false;  // neither refined nor emitted
                else divide = should_be_divided<SPHERICAL>(U, tile, planar_normal);
            }
            const uint32_t div_mask = __ballot_sync(0xffffffffu, valid && divide);
            const uint32_t fin_mask = __ballot_sync(0xffffffffu, valid && !divide);
            if (lane == 0) { S.warp_div[warp] = __popc(div_mask); S.warp_fin[warp] = __popc(fin_mask); }
            __syncthreads();
            if (warp == 0) {  // exclusive scan of the 32 warp totals
                uint32_t d = S.warp_div[lane], f = S.warp_fin[lane];
                uint32_t ds = d, fs = f;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t td = __shfl_up_sync(0xffffffffu, ds, o), tf = __shfl_up_sync(0xffffffffu, fs, o);
                    if (lane >= o) { ds += td; fs += tf; }
                }
                S.warp_div[lane] = ds - d;
                S.warp_fin[lane] = fs - f;
                if (lane == 31) { S.total_div = ds; S.total_fin = fs; }
            }
            __syncthreads();
            const uint32_t lt = (1u << lane) - 1u;
            if (valid) {
                if (divide) {
                    if (!last_pass) {  // subdivide — refine_tiles.wgsl:24-31; children of the last pass are never read (Q2)
                        const uint32_t off = n_nxt + 4u * (S.warp_div[warp] + __popc(div_mask & lt));
                        if (off + 4u <= nxt_cap) {
#pragma unroll
                            for (uint32_t c = 0; c < 4; ++c)
                                nxt[off + c] = bt_tile{tile.side, tile.lod + 1u, (tile.x << 1) + (c & 1u), (tile.y << 1) + ((c >> 1) & 1u)};
                        }
                    }
                } else {
                    const uint32_t off = n_final + S.warp_fin[warp] + __popc(fin_mask & lt);
                    if (off < capacity) final_tiles[off] = tile;
                }
            }
            const uint32_t td = S.total_div, tf = S.total_fin;
            __syncthreads();  // totals consumed before the next chunk overwrites them
            n_final += tf;
            if (last_pass) dropped += td; else n_nxt += 4u * td;
        }
        if (n_nxt > nxt_cap) { overflow = 1; n_nxt = nxt_cap & ~3u; }
        if (!last_pass) {
            cur = nxt;
            n_cur = n_nxt;
            if (n_cur > max_frontier) max_frontier = n_cur;
        }
        // writes to the next frontier (shared or global, same CTA) are ordered by the barrier above
    }

    if (tid == 0) {
        if (n_final > capacity) { overflow = 1; n_final = capacity; }
        // prepare_render — prepare_prepass.wgsl:38-43
        indirect->vertex_count = U.view_config.vertices_per_tile * n_final;
        indirect->instance_count = 1u;
        indirect->base_vertex = 0u;
        indirect->base_instance = 0u;
        status->final_count = n_final;
        status->dropped_parents = dropped;
        status->max_frontier = max_frontier;
        status->overflow = overflow;
    }
}

}  // namespace btd

extern "C" cudaError_t bt_launch_select_tiles(const btd::SelectViewParams* device_views, uint32_t n_views, int spherical, uint32_t flags,
                                              cudaStream_t stream) {
    using namespace btd;
    if (n_views == 0) return cudaSuccess;
    const size_t smem = sizeof(SelectSmem);
    cudaError_t e = cudaFuncSetAttribute(spherical ? select_tiles_kernel<true> : select_tiles_kernel<false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (spherical) select_tiles_kernel<true><<<n_views, SELECT_THREADS, smem, stream>>>(device_views, flags);
    else select_tiles_kernel<false><<<n_views, SELECT_THREADS, smem, stream>>>(device_views, flags);
    return cudaGetLastError();
}
