// clipmap.cu — clipmap residency and best-tile table on the device.
//
// Device restatement of TileTree::update (src/terrain_data/tile_tree.rs:268-333: compute_origin :180-191,
// compute_tile_distance :193-221, the request / release decisions) and of TileTree::adjust_to_tile_atlas + TileAtlasState::
// get_best_tile (tile_tree.rs:363-374, tile_atlas.rs:477-503), in f64 like the reference's main-world code and like
// host_math.cpp.  The reference walks side x lod x tree_size^2 tiles on the CPU every frame and re-uploads the whole table
// (gpu_tile_tree.rs:84-95); here the slots, the origins and the table live in device memory, a frame uploads one parameter
// block, and what comes back is the (ordered) list of tiles to request / release, which only the host can act on (disk).
//
//   clipmap_update_kernel   one CTA: origins, distances, slot state machine, request / release lists compacted in the
//                           reference's loop order (side, lod, x, y) by a block scan, so that the host's TileAtlasState
//                           hands out the same atlas indices as with bt_tile_tree_update
//   clipmap_entries_kernel  best resident ancestor of every slot through a device hash set of the loaded tiles
//
// f64 note: Coordinate::world_position evaluates powf(x, 0.5) (coordinate.rs:116); the host code calls std::pow, the device
// calls sqrt (correctly rounded).  The two can differ in the last bit of a distance; the kernel's outputs are the integer
// decisions `distance < load_distance`, which such a difference flips only when the two sides agree to ~1e-16 relative.
#include <cuda_runtime.h>
#include <stdint.h>

#include "bt_kernels.h"

namespace btd {

constexpr double kCSqrD = 0.87 * 0.87;   // C_SQR — src/math/mod.rs:13

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ double length(D3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
__device__ __forceinline__ D3 normalized(D3 a) { const double l = length(a); return {a.x / l, a.y / l, a.z / l}; }

// Coordinate::world_position — coordinate.rs:110-130, TerrainModel::position_local_to_world — terrain_model.rs:131-143
__device__ D3 coordinate_world_position(const ClipmapParams& P, uint32_t side, double cu, double cv) {
    D3 lp;
    const bool spherical = P.kind != 0u;
    if (spherical) {
        const double wu = (cu - 0.5) / 0.5, wv = (cv - 0.5) / 0.5;
        const double u = wu / sqrt(1.0 + kCSqrD - kCSqrD * wu * wu), v = wv / sqrt(1.0 + kCSqrD - kCSqrD * wv * wv);
        switch (side) {
            case 0: lp = {-1.0, -v, u}; break;
            case 1: lp = {u, -v, 1.0}; break;
            case 2: lp = {u, 1.0, v}; break;
            case 3: lp = {1.0, -u, v}; break;
            case 4: lp = {v, -u, -1.0}; break;
            default: lp = {v, -1.0, u}; break;
        }
        lp = normalized(lp);
    } else {
        lp = {cu - 0.5, 0.0, cv - 0.5};
    }
    const D3 wp{P.scale[0] * lp.x + P.position[0], P.scale[1] * lp.y + P.position[1], P.scale[2] * lp.z + P.position[2]};
    const D3 nv = spherical ? lp : D3{0.0, 1.0, 0.0};
    const D3 wn = normalized(D3{P.scale[0] * nv.x, P.scale[1] * nv.y, P.scale[2] * nv.z});
    const double h = (double)P.approximate_height;
    return wp + D3{h * wn.x, h * wn.y, h * wn.z};
}

// SideInfo::project_to_side — coordinate.rs:19-53,134-151
__device__ void project_to_side(const ClipmapParams& P, uint32_t side, double& u, double& v) {
    u = P.view_uv[0];
    v = P.view_uv[1];
    if (P.kind == 0u) return;
    // 0 = Fixed0, 1 = Fixed1, 2 = PositiveS, 3 = PositiveT
    const unsigned char even[6][2] = {{2, 3}, {0, 3}, {0, 2}, {3, 2}, {3, 0}, {2, 0}};
    const unsigned char odd[6][2] = {{2, 3}, {2, 1}, {3, 1}, {3, 2}, {1, 2}, {1, 3}};
    const uint32_t index = (6u + side - P.view_side) % 6u;
    const unsigned char* info = (P.view_side % 2u == 0u) ? even[index] : odd[index];
    auto pick = [&](unsigned char i) { return i == 0 ? 0.0 : i == 1 ? 1.0 : i == 2 ? P.view_uv[0] : P.view_uv[1]; };
    u = pick(info[0]);
    v = pick(info[1]);
}

__device__ __forceinline__ uint32_t to_u32_saturating(double x) { return x != x ? 0u : (x <= 0.0 ? 0u : (x >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)x)); }
__device__ __forceinline__ int32_t to_i32_saturating(double x) {
    return x != x ? 0 : (x <= -2147483648.0 ? (int32_t)0x80000000 : (x >= 2147483647.0 ? 0x7FFFFFFF : (int32_t)x));
}

constexpr int CLIP_THREADS = 1024;

// exclusive block scan of one 0/1 flag per thread; returns this thread's offset and the block total
__device__ uint32_t block_scan_flag(uint32_t flag, uint32_t* warp_sums, uint32_t& total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, flag != 0u);
    const uint32_t in_warp = __popc(ballot & ((1u << lane) - 1u));
    __syncthreads();                                   // warp_sums of the previous scan are no longer read
    if (lane == 0) warp_sums[warp] = __popc(ballot);
    __syncthreads();
    uint32_t base = 0, sum = 0;
    for (uint32_t w = 0; w < CLIP_THREADS / 32; ++w) {
        const uint32_t s = warp_sums[w];
        if (w < warp) base += s;
        sum += s;
    }
    total = sum;
    return base + in_warp;
}

__global__ void __launch_bounds__(CLIP_THREADS, 1)
clipmap_update_kernel(const ClipmapParams* __restrict__ params, ClipmapSlot* __restrict__ slots, uint32_t* __restrict__ origins,
                      ClipmapLists* __restrict__ lists, bt_tile* __restrict__ released, bt_tile* __restrict__ requested) {
    __shared__ ClipmapParams P;
    __shared__ uint32_t s_origin[6 * 32 * 2];
    __shared__ double s_view_uv[6 * 2];
    __shared__ uint32_t warp_sums[CLIP_THREADS / 32];
    if (threadIdx.x == 0) P = *params;
    __syncthreads();
    const uint32_t ts = P.tree_size, L = P.lod_count, S = P.side_count;
    if (threadIdx.x < S) {
        double u, v;
        project_to_side(P, threadIdx.x, u, v);
        s_view_uv[2 * threadIdx.x] = u;
        s_view_uv[2 * threadIdx.x + 1] = v;
    }
    __syncthreads();
    // TileTree::compute_origin — tile_tree.rs:180-191
    for (uint32_t i = threadIdx.x; i < S * L; i += CLIP_THREADS) {
        const uint32_t side = i / L, lod = i - side * L;
        const double tile_count = (double)(1u << lod), m = tile_count - 0.000001, hi = tile_count - (double)ts;
        for (int a = 0; a < 2; ++a) {
            double xy = fmin(s_view_uv[2 * side + a] * tile_count, m);      // compute_tree_xy — tile_tree.rs:175-178
            xy = round(xy - 0.5 * (double)ts);
            xy = fmin(fmax(xy, 0.0), hi);
            const uint32_t o = to_u32_saturating(xy);
            s_origin[2 * i + a] = o;
            origins[2 * i + a] = o;
        }
    }
    __syncthreads();
    const uint32_t n = S * L * ts * ts;
    uint32_t n_released = 0, n_requested = 0;
    for (uint32_t first = 0; first < n; first += CLIP_THREADS) {
        const uint32_t e = first + threadIdx.x;
        uint32_t rel = 0, req = 0;
        bt_tile rel_tile{}, req_tile{};
        if (e < n) {
            const uint32_t y = e % ts, x = (e / ts) % ts, sl = e / (ts * ts), side = sl / L, lod = sl - side * L;
            const uint32_t tx = s_origin[2 * sl] + x, ty = s_origin[2 * sl + 1] + y;
            // TileTree::compute_tile_distance — tile_tree.rs:193-221
            const double tile_count = (double)(1u << lod), m = tile_count - 0.000001;
            const double vx = fmin(s_view_uv[2 * side] * tile_count, m), vy = fmin(s_view_uv[2 * side + 1] * tile_count, m);
            const int32_t ox = to_i32_saturating(vx) - (int32_t)tx, oy = to_i32_saturating(vy) - (int32_t)ty;
            double fx = fmod(vx, 1.0), fy = fmod(vy, 1.0);
            if (ox < 0) fx = 0.0; else if (ox > 0) fx = 1.0;
            if (oy < 0) fy = 0.0; else if (oy > 0) fy = 1.0;
            const D3 wp = coordinate_world_position(P, side, ((double)(int32_t)tx + fx) / tile_count, ((double)(int32_t)ty + fy) / tile_count);
            const D3 view{P.view_world_position[0], P.view_world_position[1], P.view_world_position[2]};
            const double tile_distance = length(wp - view);
            const bool want = lod == 0u || tile_distance < P.load_distance / tile_count;
            // the slot state machine of TileTree::update — tile_tree.rs:300-331
            ClipmapSlot& slot = slots[(size_t)sl * ts * ts + (size_t)(tx % ts) * ts + (ty % ts)];
            ClipmapSlot s = slot;
            if (!(s.valid && s.x == tx && s.y == ty)) {
                if (s.requested) { rel = 1; rel_tile = bt_tile{side, lod, s.x, s.y}; s.requested = 0; }
                s.x = tx; s.y = ty; s.valid = 1;
            }
            if (!s.requested && want) { s.requested = 1; req = 1; req_tile = bt_tile{side, lod, tx, ty}; }
            else if (s.requested && !want) { s.requested = 0; rel = 1; rel_tile = bt_tile{side, lod, tx, ty}; }
            slot = s;
        }
        uint32_t total;
        const uint32_t r_off = block_scan_flag(rel, warp_sums, total);
        if (rel) released[n_released + r_off] = rel_tile;
        n_released += total;
        const uint32_t q_off = block_scan_flag(req, warp_sums, total);
        if (req) requested[n_requested + q_off] = req_tile;
        n_requested += total;
    }
    if (threadIdx.x == 0) { lists->n_released = n_released; lists->n_requested = n_requested; }
}

__host__ __device__ inline uint32_t clipmap_hash(uint32_t side_lod, uint32_t x, uint32_t y) {
    uint32_t h = side_lod * 0x9E3779B1u;
    h ^= x + 0x7F4A7C15u + (h << 6) + (h >> 2);
    h ^= y + 0x85EBCA6Bu + (h << 6) + (h >> 2);
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 12;
    return h;
}

// TileTree::adjust_to_tile_atlas + TileAtlasState::get_best_tile — tile_tree.rs:363-374, tile_atlas.rs:477-503
__global__ void clipmap_entries_kernel(const ClipmapSlot* __restrict__ slots, uint32_t n, uint32_t tree_size, uint32_t lod_count,
                                       const ClipmapLoaded* __restrict__ loaded, uint32_t hash_mask, bt_tile_tree_entry* __restrict__ entries) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const ClipmapSlot s = slots[e];
    bt_tile_tree_entry out{0xFFFFFFFFu, 0xFFFFFFFFu};
    if (s.valid) {
        const uint32_t sl = e / (tree_size * tree_size), side = sl / lod_count;
        uint32_t lod = sl - side * lod_count, x = s.x, y = s.y;
        for (;;) {
            const uint32_t key = (side << 8) | lod;
            uint32_t i = clipmap_hash(key, x, y) & hash_mask;
            bool found = false;
            for (;;) {   // linear probing; the table is at most 1/4 full
                const ClipmapLoaded c = loaded[i];
                if (c.side_lod == 0xFFFFFFFFu) break;
                if (c.side_lod == key && c.x == x && c.y == y) { out = bt_tile_tree_entry{c.atlas_index, lod}; found = true; break; }
                i = (i + 1u) & hash_mask;
            }
            if (found || lod == 0u) break;   // parent of a lod-0 tile is invalid (wrapping_sub): no ancestor is loaded
            lod -= 1u; x >>= 1; y >>= 1;
        }
    }
    entries[e] = out;
}

}  // namespace btd

using namespace btd;

extern "C" cudaError_t bt_launch_clipmap_update(const ClipmapParams* params, ClipmapSlot* slots, uint32_t* origins, ClipmapLists* lists,
                                                bt_tile* released, bt_tile* requested, cudaStream_t stream) {
    clipmap_update_kernel<<<1, CLIP_THREADS, 0, stream>>>(params, slots, origins, lists, released, requested);
    return cudaGetLastError();
}

extern "C" cudaError_t bt_launch_clipmap_entries(const ClipmapSlot* slots, uint32_t n, uint32_t tree_size, uint32_t lod_count,
                                                 const ClipmapLoaded* loaded, uint32_t hash_capacity, bt_tile_tree_entry* entries, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    clipmap_entries_kernel<<<(n + 255u) / 256u, 256, 0, stream>>>(slots, n, tree_size, lod_count, loaded, hash_capacity - 1u, entries);
    return cudaGetLastError();
}

extern "C" uint32_t bt_clipmap_hash(uint32_t side_lod, uint32_t x, uint32_t y) { return clipmap_hash(side_lod, x, y); }
