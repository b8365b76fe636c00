// bt_preprocess.cu — the preprocessor runtime behind bt_preprocessor_* (C ABI).
//
// C++ counterpart of Preprocessor (src/preprocess/preprocessor.rs:214-343: task graph
// Split -> Barrier -> Downsample per level -> Barrier -> per-LOD Stitch -> Save), of
// GpuPreprocessor::prepare (src/preprocess/gpu_preprocessor.rs:120-223: per-task uniforms) and of
// TerrainPreprocessNode::run (src/preprocess/mod.rs:143-217: the dispatch loop).  Differences by
// design: all tiles of a phase go in ONE launch (no 32-tile write-section cap, no frame pacing),
// tiles live in a device-resident pyramid store instead of atlas slots + staging buffers, and
// the deepest-LOD tile rows can be sharded over ranks (bt_shard) with a halo + sub-root exchange
// that the host performs with NCCL between the exposed phases.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "bt_comm.h"
#include "bt_kernels.h"
#include "bt_runtime.h"

using btd::DatasetDevice;
using btd::PyramidTask;
using btd::RowCopy;

#define BT_SET_ERROR(ctx_, ...) do { if (ctx_) (ctx_)->set_error(__VA_ARGS__); } while (0)

namespace {

constexpr uint32_t kInvalid = 0xFFFFFFFFu;

struct TileKey {
    uint32_t side, lod, x, y;
    bool operator==(const TileKey& o) const { return side == o.side && lod == o.lod && x == o.x && y == o.y; }
};
struct TileKeyHash {
    size_t operator()(const TileKey& k) const {
        uint64_t h = 0xcbf29ce484222325ull;
        for (uint32_t w : {k.side, k.lod, k.x, k.y}) { h ^= w; h *= 0x100000001b3ull; h ^= h >> 31; }
        return (size_t)h;
    }
};

enum SlotKind : uint8_t { kOwned = 0, kGhost = 1, kReplica = 2 };

struct Store {  // one attachment's device-resident tile pyramid
    bt_attachment_config cfg{};
    size_t tile_bytes = 0;
    uint32_t px = 0;
    std::unordered_map<TileKey, uint32_t, TileKeyHash> slot_of;
    std::vector<TileKey> keys;
    std::vector<uint32_t> exists_from_job;
    std::vector<uint8_t> kind;
    unsigned char* arena = nullptr;
    std::vector<uint8_t> planned;  // slot is the target of an already planned (earlier) job
    std::vector<uint8_t> written;  // slot has been produced by a kernel (else its content is "the zero-initialised atlas texture")
    uint32_t arena_slots = 0;    // slots the arena can hold
    uint32_t cleared_slots = 0;  // slots whose first-use handling (ghost zeroing) has been done
};

struct Range { uint32_t begin = 0, count = 0; };
struct Level { uint32_t lod; Range range; };
struct JobPlan {  // all task tables of one job, concatenated: [split | downsample levels (deepest first) | stitch]
    std::vector<PyramidTask> tasks;
    std::vector<DatasetDevice> datasets;
    Range split, stitch;
    float split_src_per_px[2] = {0.0f, 0.0f};   // densest source of the split tasks, in source texels per tile pixel (x, y): sizes the kernel's stage buffers
    uint32_t stitch_local = 0;   // the first stitch_local tasks of `stitch` need nothing from another rank (run while the exchange is in flight)
    std::vector<Level> levels;
    PyramidTask* d_tasks = nullptr;
    DatasetDevice* d_datasets = nullptr;
};

// device-resident copy tables + buffer segments of one job's exchange (kind 0 = halo strips, 1 = sub-root tiles), built once
struct ExchangeTables {
    RowCopy* d_copies = nullptr;          // [pack kind 0 | pack kind 1 | unpack kind 0 | unpack kind 1]
    uint32_t pack_begin[2] = {0, 0}, pack_count[2] = {0, 0}, unpack_begin[2] = {0, 0}, unpack_count[2] = {0, 0};
    uint64_t send_off[2] = {0, 0}, recv_off[2] = {0, 0}, bytes[2] = {0, 0};
    bool built = false;
};

struct Job {
    uint32_t attachment = 0;
    std::vector<bt_dataset> datasets;
    std::vector<void*> owned_pixels;  // device copies of host-provided sources
    JobPlan plan;
    ExchangeTables xt;
    bool prepared = false, split_done = false, roots_done = false, halo_done = false, stitched = false;
};

// NEIGHBOURING_SIDES / SideInfo — src/math/coordinate.rs:9-53
constexpr uint32_t kNeighbouringSides[6][5] = {{0, 4, 2, 1, 5}, {1, 0, 2, 3, 5}, {2, 0, 4, 3, 1}, {3, 2, 4, 5, 1}, {4, 2, 0, 5, 3}, {5, 4, 0, 1, 3}};
enum SideInfo : uint8_t { kFixed0, kFixed1, kPosS, kPosT };
constexpr SideInfo kEven[6][2] = {{kPosS, kPosT}, {kFixed0, kPosT}, {kFixed0, kPosS}, {kPosT, kPosS}, {kPosT, kFixed0}, {kPosS, kFixed0}};
constexpr SideInfo kOdd[6][2] = {{kPosS, kPosT}, {kPosS, kFixed1}, {kPosT, kFixed1}, {kPosT, kPosS}, {kFixed1, kPosS}, {kFixed1, kPosT}};
constexpr TileKey kInvalidTile{kInvalid, kInvalid, kInvalid, kInvalid};

// TileCoordinate::neighbour_coordinate — coordinate.rs:227-279
TileKey neighbour_coordinate(const TileKey& self, int32_t nx, int32_t ny, bool spherical) {
    const int32_t n = (int32_t)(1u << self.lod);
    if (!spherical) {
        if (nx < 0 || ny < 0 || nx >= n || ny >= n) return kInvalidTile;
        return TileKey{self.side, self.lod, (uint32_t)nx, (uint32_t)ny};
    }
    const bool xl = nx < 0, xh = nx >= n, yl = ny < 0, yh = ny >= n;
    if ((xl || xh) && (yl || yh)) return kInvalidTile;
    const int edge = xl ? 1 : yl ? 2 : xh ? 3 : yh ? 4 : 0;
    const uint32_t px = (uint32_t)std::min(std::max(nx, 0), n - 1), py = (uint32_t)std::min(std::max(ny, 0), n - 1);
    const uint32_t side = kNeighbouringSides[self.side][edge];
    const uint32_t index = (6 + side - self.side) % 6;
    const SideInfo* info = (self.side % 2 == 0) ? kEven[index] : kOdd[index];
    auto pick = [&](SideInfo i) { return i == kFixed0 ? 0u : i == kFixed1 ? (uint32_t)n - 1u : i == kPosS ? px : py; };
    return TileKey{side, self.lod, pick(info[0]), pick(info[1])};
}

uint32_t f32_to_u32(float x) {
    if (!(x > 0.0f)) return 0u;
    if (x >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)x;
}

uint32_t ilog2(uint32_t v) { uint32_t l = 0; while ((1u << (l + 1)) <= v) ++l; return l; }

// TileCoordinate::neighbours order — coordinate.rs:209-218: top, right, bottom, left, TL, TR, BR, BL
constexpr int32_t kNeighbourOffsets[8][2] = {{0, -1}, {1, 0}, {0, 1}, {-1, 0}, {-1, -1}, {1, -1}, {1, 1}, {-1, 1}};

}  // namespace

struct bt_preprocessor {
    bt_context* ctx = nullptr;
    uint32_t kind = BT_KIND_PLANAR;
    uint32_t lod_count = 0;
    bt_shard shard{0, 1};
    uint32_t root_lod = 0;  // L0 = log2(world): levels below it are computed redundantly on every rank
    std::vector<Store> stores;
    std::vector<Job> jobs;
    // scratch device buffers (grown on demand)
    RowCopy* d_copies = nullptr; size_t d_copies_cap = 0;
    // exchange buffers of the in-library sharded run (bt_preprocessor_run_sharded): all pending jobs' segments back to back
    unsigned char* d_send = nullptr; unsigned char* d_recv = nullptr; uint64_t send_cap = 0, recv_cap = 0;

    bool spherical() const { return kind == BT_KIND_SPHERICAL; }
    bool plan_only() const { return ctx == nullptr; }  // created without a context: host-side planning only

    // PreprocessDataset::overlapping_tiles — preprocessor.rs:58-66 (f32), intersected with the rows
    // this rank owns (world == 1: everything)
    void tile_range(const bt_dataset& d, uint32_t lod, uint32_t lo[2], uint32_t hi[2]) const {
        const float tc = (float)(1u << lod);
        for (int k = 0; k < 2; ++k) {
            lo[k] = f32_to_u32(d.top_left[k] * tc);
            hi[k] = f32_to_u32(std::ceil(d.bottom_right[k] * tc));
        }
    }
    // rows [y0, y1) of `lod` owned by this rank; levels above the root level are owned by everyone
    void owned_rows(uint32_t lod, uint32_t* y0, uint32_t* y1) const {
        const uint32_t n = 1u << lod;
        if (shard.world == 1 || lod < root_lod) { *y0 = 0; *y1 = n; return; }
        *y0 = (uint32_t)((uint64_t)shard.rank * n / shard.world);
        *y1 = (uint32_t)((uint64_t)(shard.rank + 1) * n / shard.world);
    }

    // rank that owns tile row y of `lod` (lod >= root_lod; the band split is the same on every cube face)
    uint32_t owner_of_row(uint32_t lod, uint32_t y) const { return (uint32_t)((uint64_t)y * shard.world >> lod); }

    uint32_t get_or_allocate(Store& s, const TileKey& k, uint32_t job, uint8_t kind_) {
        auto it = s.slot_of.find(k);
        if (it != s.slot_of.end()) {
            if (kind_ == kOwned) s.kind[it->second] = kOwned;
            return it->second;
        }
        const uint32_t slot = (uint32_t)s.keys.size();
        s.slot_of.emplace(k, slot);
        s.keys.push_back(k);
        s.exists_from_job.push_back(job);
        s.kind.push_back(kind_);
        s.written.push_back(0);
        return slot;
    }
    // TileAtlasState::get_tile — tile_atlas.rs:369-381, evaluated at queue-build time of `job`
    uint32_t get_tile(const Store& s, const TileKey& k, uint32_t job) const {
        if (k == kInvalidTile) return kInvalid;
        auto it = s.slot_of.find(k);
        if (it == s.slot_of.end() || s.exists_from_job[it->second] > job) return kInvalid;
        return it->second;
    }

    bt_status ensure_arena(Store& s, cudaStream_t stream) {
        const uint32_t need = (uint32_t)s.keys.size();
        if (need > s.arena_slots) {
            unsigned char* bigger = nullptr;
            cudaError_t e = cudaMalloc((void**)&bigger, (size_t)need * s.tile_bytes);
            if (e != cudaSuccess) {
                BT_SET_ERROR(ctx, "preprocessor: pyramid store of %u tiles (%zu bytes) does not fit: %s", need, (size_t)need * s.tile_bytes, cudaGetErrorString(e));
                (void)cudaGetLastError();
                return BT_ERR_OUT_OF_MEMORY;
            }
            if (s.arena) {
                BT_CUDA(ctx, cudaMemcpyAsync(bigger, s.arena, (size_t)s.arena_slots * s.tile_bytes, cudaMemcpyDeviceToDevice, stream));
                BT_CUDA(ctx, cudaStreamSynchronize(stream));
                BT_CUDA(ctx, cudaFree(s.arena));
            }
            s.arena = bigger;
            s.arena_slots = need;
        }
        // The reference's atlas textures start zeroed (wgpu) and split falls through to that content.  Instead of
        // clearing the whole store (32 GiB for config 5), split tasks on never-written slots carry a "fresh" flag and
        // substitute zero for the read; downsample overwrites every texel; only ghost tiles (written strip-wise by
        // the halo unpack) are cleared here.
        for (uint32_t i = s.cleared_slots; i < need; ++i)
            if (s.kind[i] == kGhost) BT_CUDA(ctx, cudaMemsetAsync(s.arena + (size_t)i * s.tile_bytes, 0, s.tile_bytes, stream));
        s.cleared_slots = need;
        return BT_OK;
    }

    template <typename T>
    bt_status upload(T** dev, size_t* cap, const std::vector<T>& host, cudaStream_t stream) {
        if (host.size() > *cap) {
            if (*dev) { BT_CUDA(ctx, cudaStreamSynchronize(stream)); BT_CUDA(ctx, cudaFree(*dev)); *dev = nullptr; }
            BT_CUDA(ctx, cudaMalloc((void**)dev, host.size() * sizeof(T)));
            *cap = host.size();
        }
        if (!host.empty()) {
            BT_CUDA(ctx, cudaMemcpyAsync(*dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
            BT_CUDA(ctx, cudaStreamSynchronize(stream));  // `host` is a temporary: make the pageable copy complete
        }
        return BT_OK;
    }

    // ---- task builders -------------------------------------------------------------------------
    // all tiles (every LOD, every dataset) of a job exist before any task resolves its dependencies:
    // split_and_downsample is queued for all datasets before the first stitch (preprocessor.rs:298-343)
    void allocate_job_tiles(uint32_t ji) {
        Job& j = jobs[ji];
        Store& s = stores[j.attachment];
        for (const bt_dataset& d : j.datasets)
            for (uint32_t lod = d.lod_begin; lod < d.lod_end; ++lod) {
                uint32_t lo[2], hi[2], y0, y1;
                tile_range(d, lod, lo, hi);
                owned_rows(lod, &y0, &y1);
                for (uint32_t x = lo[0]; x < hi[0]; ++x)
                    for (uint32_t y = std::max(lo[1], y0); y < std::min(hi[1], y1); ++y)
                        get_or_allocate(s, TileKey{d.side, lod, x, y}, ji, (lod < root_lod && shard.rank != 0) ? kReplica : kOwned);
                if (shard.world > 1 && lod >= root_lod) {
                    if (lod == root_lod) {  // the whole root level is replicated by the sub-root gather
                        for (uint32_t x = lo[0]; x < hi[0]; ++x)
                            for (uint32_t y = lo[1]; y < hi[1]; ++y)
                                if (y < y0 || y >= y1) get_or_allocate(s, TileKey{d.side, lod, x, y}, ji, kReplica);
                    } else if (!spherical()) {  // ghost rows that receive the halo strips of the adjacent ranks
                        for (uint32_t x = lo[0]; x < hi[0]; ++x) {
                            if (y0 > lo[1]) get_or_allocate(s, TileKey{d.side, lod, x, y0 - 1}, ji, kGhost);
                            if (y1 < hi[1]) get_or_allocate(s, TileKey{d.side, lod, x, y1}, ji, kGhost);
                        }
                    } else {  // cube: every neighbour (same face or across a face edge) of an owned tile that another rank owns
                        for (uint32_t x = lo[0]; x < hi[0]; ++x)
                            for (uint32_t y = std::max(lo[1], y0); y < std::min(hi[1], y1); ++y)
                                for (int k = 0; k < 8; ++k) {
                                    const TileKey self{d.side, lod, x, y};
                                    const TileKey nb = neighbour_coordinate(self, (int32_t)x + kNeighbourOffsets[k][0], (int32_t)y + kNeighbourOffsets[k][1], true);
                                    if (!(nb == kInvalidTile) && owner_of_row(lod, nb.y) != shard.rank) get_or_allocate(s, nb, ji, kGhost);
                                }
                    }
                }
            }
    }

    void collect(uint32_t ji, const bt_dataset& d, uint32_t di, uint32_t lod, std::vector<PyramidTask>& out) {
        Store& s = stores[jobs[ji].attachment];
        uint32_t lo[2], hi[2], y0, y1;
        tile_range(d, lod, lo, hi);
        owned_rows(lod, &y0, &y1);
        for (uint32_t x = lo[0]; x < hi[0]; ++x)
            for (uint32_t y = std::max(lo[1], y0); y < std::min(hi[1], y1); ++y) {
                PyramidTask t;
                std::memset(&t, 0xFF, sizeof(t));
                t.tile = bt_tile{d.side, lod, x, y};
                t.slot = s.slot_of.at(TileKey{d.side, lod, x, y});
                t.dataset = di;
                t.flags = 0u;
                out.push_back(t);
            }
    }

    // ---- job plan: every task table of a job is built once, uploaded with ONE copy, and the kernels are then
    // launched back to back without host synchronisation (the reference rebuilds a uniform buffer + bind group
    // per task per frame, gpu_preprocessor.rs:120-223)
    bt_status prepare_job(uint32_t ji, cudaStream_t stream) {
        Job& j = jobs[ji];
        if (j.prepared) return BT_OK;
        Store& s = stores[j.attachment];
        allocate_job_tiles(ji);
        if (s.planned.size() < s.keys.size()) s.planned.resize(s.keys.size(), 0);
        uint32_t lod_end = 0, lod_begin = kInvalid;
        for (const bt_dataset& d : j.datasets) { lod_end = std::max(lod_end, d.lod_end); lod_begin = std::min(lod_begin, d.lod_begin); }
        JobPlan& P = j.plan;
        std::vector<PyramidTask>& tasks = P.tasks;
        tasks.clear();
        P.datasets.clear();
        // split of the deepest level of every dataset
        for (uint32_t di = 0; di < j.datasets.size(); ++di) {
            const bt_dataset& d = j.datasets[di];
            DatasetDevice x{};
            x.pixels = d.pixels;
            x.width = d.width; x.height = d.height;
            x.top_left[0] = d.top_left[0]; x.top_left[1] = d.top_left[1];
            x.bottom_right[0] = d.bottom_right[0]; x.bottom_right[1] = d.bottom_right[1];
            x.row_begin = d.row_count ? d.row_begin : 0;
            x.row_count = d.row_count ? d.row_count : d.height;
            P.datasets.push_back(x);
            const double px = (double)(s.cfg.texture_size - 2u * s.cfg.border_size) * (double)(1u << ((d.lod_end - 1u) & 31u));   // centre pixels across the unit square
            const double rx = (double)d.bottom_right[0] - (double)d.top_left[0], ry = (double)d.bottom_right[1] - (double)d.top_left[1];
            const float qx = rx > 0.0 ? (float)((double)d.width / (px * rx)) : 1e9f, qy = ry > 0.0 ? (float)((double)d.height / (px * ry)) : 1e9f;
            P.split_src_per_px[0] = di == 0 ? qx : std::max(P.split_src_per_px[0], qx);
            P.split_src_per_px[1] = di == 0 ? qy : std::max(P.split_src_per_px[1], qy);
            collect(ji, d, di, d.lod_end - 1, tasks);
        }
        P.split = Range{0u, (uint32_t)tasks.size()};
        for (uint32_t i = P.split.begin; i < P.split.begin + P.split.count; ++i) {
            tasks[i].flags = (s.written[tasks[i].slot] || s.planned[tasks[i].slot]) ? 0u : btd::kTaskFresh;
            s.planned[tasks[i].slot] = 1;
        }
        // downsample levels, deepest first; sharded: levels below the root level run after the sub-root gather
        P.levels.clear();
        for (int32_t lod = (int32_t)lod_end - 2; lod >= (int32_t)lod_begin; --lod) {
            const uint32_t first = (uint32_t)tasks.size();
            for (uint32_t di = 0; di < j.datasets.size(); ++di) {
                const bt_dataset& d = j.datasets[di];
                if ((uint32_t)lod + 1 >= d.lod_end || (uint32_t)lod < d.lod_begin) continue;  // the deepest level is split, not downsampled
                const size_t f = tasks.size();
                collect(ji, d, di, (uint32_t)lod, tasks);
                for (size_t i = f; i < tasks.size(); ++i) {  // TileCoordinate::children — coordinate.rs:197-206
                    for (uint32_t c = 0; c < 4; ++c)
                        tasks[i].dep[c] = get_tile(s, TileKey{d.side, (uint32_t)lod + 1, (tasks[i].tile.x << 1) + c % 2, (tasks[i].tile.y << 1) + c / 2}, ji);
                    s.planned[tasks[i].slot] = 1;
                }
            }
            if (tasks.size() > first) P.levels.push_back(Level{(uint32_t)lod, Range{first, (uint32_t)tasks.size() - first}});
        }
        // stitch: every LOD in one launch (reads only neighbour CENTRES, all final after the downsample phase)
        const uint32_t stitch_first = (uint32_t)tasks.size();
        for (uint32_t di = 0; di < j.datasets.size(); ++di) {
            const bt_dataset& d = j.datasets[di];
            for (uint32_t lod = d.lod_begin; lod < d.lod_end; ++lod) {
                const size_t f = tasks.size();
                collect(ji, d, di, lod, tasks);
                for (size_t i = f; i < tasks.size(); ++i) {
                    const TileKey self{tasks[i].tile.side, tasks[i].tile.lod, tasks[i].tile.x, tasks[i].tile.y};
                    for (int k = 0; k < 8; ++k) {  // TileCoordinate::neighbours — coordinate.rs:208-225
                        const TileKey nb = neighbour_coordinate(self, (int32_t)self.x + kNeighbourOffsets[k][0], (int32_t)self.y + kNeighbourOffsets[k][1], spherical());
                        tasks[i].dep[k] = get_tile(s, nb, ji);
                        tasks[i].dep_side[k] = nb.side;
                    }
                }
            }
        }
        P.stitch = Range{stitch_first, (uint32_t)tasks.size() - stitch_first};
        // sharded: tasks that read nothing another rank produces (no ghost / replica neighbour, not one of the replicated top
        // levels) go first, so that bt_preprocessor_run_sharded can stitch them while the exchange is still in flight
        P.stitch_local = P.stitch.count;
        if (shard.world > 1) {
            auto local = [&](const PyramidTask& t) {
                if (t.tile.lod < root_lod) return false;
                for (int k = 0; k < 8; ++k) if (t.dep[k] != kInvalid && s.kind[t.dep[k]] != kOwned) return false;
                return true;
            };
            const auto mid = std::stable_partition(tasks.begin() + stitch_first, tasks.end(), local);
            P.stitch_local = (uint32_t)(mid - (tasks.begin() + stitch_first));
        }
        j.prepared = true;
        if (plan_only()) return BT_OK;
        // device copies (one allocation + one copy per job)
        bt_status st = ensure_arena(s, stream);
        if (st != BT_OK) return st;
        BT_CUDA(ctx, cudaMalloc((void**)&P.d_tasks, std::max<size_t>(tasks.size(), 1) * sizeof(PyramidTask)));
        BT_CUDA(ctx, cudaMalloc((void**)&P.d_datasets, std::max<size_t>(P.datasets.size(), 1) * sizeof(DatasetDevice)));
        BT_CUDA(ctx, cudaMemcpyAsync(P.d_tasks, tasks.data(), tasks.size() * sizeof(PyramidTask), cudaMemcpyHostToDevice, stream));
        BT_CUDA(ctx, cudaMemcpyAsync(P.d_datasets, P.datasets.data(), P.datasets.size() * sizeof(DatasetDevice), cudaMemcpyHostToDevice, stream));
        BT_CUDA(ctx, cudaStreamSynchronize(stream));  // pageable sources: the copies are complete when this returns
        return BT_OK;
    }

    bt_status run_levels(uint32_t ji, uint32_t lod_hi_exclusive, uint32_t lod_lo, bool with_split, cudaStream_t stream) {
        // downsample levels lod_hi_exclusive-1 ... lod_lo (and the split of the deepest level if asked); launches only
        Job& j = jobs[ji];
        Store& s = stores[j.attachment];
        const JobPlan& P = j.plan;
        if (with_split) {
            BT_CUDA(ctx, bt_launch_split(s.arena, s.tile_bytes, s.cfg.texture_size, s.cfg.border_size, s.cfg.format, P.d_tasks + P.split.begin, P.split.count, P.d_datasets,
                                         P.split_src_per_px[0], P.split_src_per_px[1], stream));
            for (uint32_t i = P.split.begin; i < P.split.begin + P.split.count; ++i) s.written[P.tasks[i].slot] = 1;
        }
        for (const Level& lv : P.levels) {
            if (lv.lod >= lod_hi_exclusive || lv.lod < lod_lo) continue;
            BT_CUDA(ctx, bt_launch_downsample(s.arena, s.tile_bytes, s.cfg.texture_size, s.cfg.border_size, s.cfg.format, P.d_tasks + lv.range.begin, lv.range.count, stream));
            for (uint32_t i = lv.range.begin; i < lv.range.begin + lv.range.count; ++i) s.written[P.tasks[i].slot] = 1;
        }
        return BT_OK;
    }

    bt_status run_split_downsample_job(uint32_t ji, cudaStream_t stream) {
        Job& j = jobs[ji];
        bt_status st = prepare_job(ji, stream);
        if (st != BT_OK) return st;
        uint32_t lod_end = 0, lod_begin = kInvalid;
        for (const bt_dataset& d : j.datasets) { lod_end = std::max(lod_end, d.lod_end); lod_begin = std::min(lod_begin, d.lod_begin); }
        // sharded: stop at the root level; the levels above it follow the sub-root gather
        const uint32_t lo = shard.world > 1 ? std::max(lod_begin, root_lod) : lod_begin;
        st = run_levels(ji, lod_end - 1, lo, true, stream);
        if (st == BT_OK) j.split_done = true;
        return st;
    }

    // part 0 = the tasks that need nothing from another rank, 1 = the rest, 2 = all
    bt_status run_stitch_job(uint32_t ji, cudaStream_t stream, int part = 2) {
        Job& j = jobs[ji];
        Store& s = stores[j.attachment];
        const JobPlan& P = j.plan;
        const uint32_t first = part == 1 ? P.stitch_local : 0u;
        const uint32_t count = part == 0 ? P.stitch_local : P.stitch.count - first;
        BT_CUDA(ctx, bt_launch_stitch(s.arena, s.tile_bytes, s.cfg.texture_size, s.cfg.border_size, s.cfg.format, P.d_tasks + P.stitch.begin + first, count, stream));
        if (part != 0) j.stitched = true;
        return BT_OK;
    }

    int pending_job() const {
        for (size_t i = 0; i < jobs.size(); ++i) if (!jobs[i].stitched) return (int)i;
        return -1;
    }

    // ---- multi-GPU exchange layouts ---------------------------------------------------------------
    // halo: for every dataset LOD above the root level: [top strips of my first owned row][bottom
    // strips of my last owned row], each strip = border_size rows x center_size pixels of one tile
    uint64_t halo_bytes(const Job& j) const {
        if (shard.world == 1) return 0;
        const Store& s = stores[j.attachment];
        const bt_dataset& d = j.datasets[0];
        const uint32_t C = s.cfg.texture_size - 2 * s.cfg.border_size;
        uint64_t bytes = 0;
        for (uint32_t lod = std::max(d.lod_begin, root_lod + 1); lod < d.lod_end; ++lod) {
            if (!spherical()) bytes += 2ull * (1u << lod) * s.cfg.border_size * C * s.px;
            else bytes += (uint64_t)j.datasets.size() * cube_strips_per_side(lod) * s.cfg.border_size * C * s.px;
        }
        return bytes;
    }
    uint64_t roots_bytes(const Job& j) const {
        if (shard.world == 1) return 0;
        const bt_dataset& d = j.datasets[0];
        if (d.lod_begin > root_lod) return 0;  // no level above the shard boundary is requested... but root tiles still needed by nobody
        return (uint64_t)j.datasets.size() * (1u << root_lod) * stores[j.attachment].tile_bytes;  // one row of the root level per dataset (cube face)
    }
    // cube halo: per LOD and face a rank sends the outward strips of its band-boundary and face-edge tiles, in this order:
    // [top strips of row y0, x ascending][bottom strips of row y1-1][left strips of column 0, y ascending][right strips of
    // column n-1]; every strip is border_size x center_size pixels, so an entry's offset is its index times one strip size.
    uint64_t cube_strips_per_side(uint32_t lod) const { const uint64_t n = 1ull << lod; return 2 * n + 2 * (n / shard.world); }
};

extern "C" {

bt_status bt_preprocessor_create(bt_context* ctx, uint32_t kind, const bt_attachment_config* attachments, uint32_t attachment_count,
                                 uint32_t lod_count, bt_shard shard, bt_preprocessor** out) {
    if (!attachments || !out || attachment_count == 0 || attachment_count > 8 || lod_count == 0 || lod_count > 24) return BT_ERR_INVALID_ARG;
    if (kind != BT_KIND_PLANAR && kind != BT_KIND_SPHERICAL) return BT_ERR_INVALID_ARG;
    if (shard.world == 0 || shard.rank >= shard.world || (shard.world & (shard.world - 1)) != 0) {
        BT_SET_ERROR(ctx, "bt_preprocessor_create: shard world must be a power of two and rank < world");
        return BT_ERR_INVALID_ARG;
    }
    bt_preprocessor* p = new bt_preprocessor();
    p->ctx = ctx;
    p->kind = kind;
    p->lod_count = lod_count;
    p->shard = shard;
    p->root_lod = ilog2(shard.world);
    for (uint32_t i = 0; i < attachment_count; ++i) {
        const bt_attachment_config& a = attachments[i];
        if ((a.format != BT_FORMAT_R16 && a.format != BT_FORMAT_RGBA8) || a.texture_size <= 2 * a.border_size) {
            BT_SET_ERROR(ctx, "bt_preprocessor_create: attachment %u: unsupported format %u (only R16/Rgba8 have shader paths in the reference) or texture_size <= 2*border", i, a.format);
            delete p;
            return BT_ERR_INVALID_ARG;
        }
        Store s;
        s.cfg = a;
        s.px = a.format == BT_FORMAT_R16 ? 2 : 4;
        s.tile_bytes = (size_t)a.texture_size * a.texture_size * s.px;
        p->stores.push_back(std::move(s));
    }
    *out = p;
    return BT_OK;
}

void bt_preprocessor_destroy(bt_preprocessor* p) {
    if (!p) return;
    if (p->plan_only()) { delete p; return; }
    cudaSetDevice(p->ctx->device);
    cudaDeviceSynchronize();
    for (Store& s : p->stores) if (s.arena) cudaFree(s.arena);
    for (Job& j : p->jobs) { for (void* px : j.owned_pixels) cudaFree(px); cudaFree(j.plan.d_tasks); cudaFree(j.plan.d_datasets); cudaFree(j.xt.d_copies); }
    cudaFree(p->d_copies);
    cudaFree(p->d_send);
    cudaFree(p->d_recv);
    delete p;
}

bt_status bt_preprocessor_add_datasets(bt_preprocessor* p, const bt_dataset* datasets, uint32_t count) {
    if (!p || !datasets || count == 0) return BT_ERR_INVALID_ARG;
    bt_context* ctx = p->ctx;
    Job j;
    j.attachment = datasets[0].attachment_index;
    if (j.attachment >= p->stores.size()) { BT_SET_ERROR(ctx, "add_datasets: attachment_index %u out of range", j.attachment); return BT_ERR_INVALID_ARG; }
    if (p->shard.world > 1 && count != (p->spherical() ? 6u : 1u)) {
        BT_SET_ERROR(ctx, "add_datasets: sharded preprocessing takes one dataset per job (planar) or the six cube faces of one attachment (spherical)");
        return BT_ERR_INVALID_ARG;
    }
    uint32_t side_mask = 0;
    for (uint32_t i = 0; i < count; ++i) {
        const bt_dataset& d = datasets[i];
        if (d.attachment_index != j.attachment || d.side >= 6 || d.lod_begin >= d.lod_end || d.lod_end > p->lod_count || (!d.pixels && !p->plan_only()) || d.width == 0 || d.height == 0) {
            BT_SET_ERROR(ctx, "add_datasets: dataset %u invalid (attachment/side/lod_range %u..%u of %u/pixels)", i, d.lod_begin, d.lod_end, p->lod_count);
            return BT_ERR_INVALID_ARG;
        }
        if (!(d.bottom_right[0] > d.top_left[0]) || !(d.bottom_right[1] > d.top_left[1])) { BT_SET_ERROR(ctx, "add_datasets: empty extent"); return BT_ERR_INVALID_ARG; }
        if (d.row_count && (uint64_t)d.row_begin + d.row_count > d.height) { BT_SET_ERROR(ctx, "add_datasets: row window exceeds the image"); return BT_ERR_INVALID_ARG; }
        if (d.row_count && d.row_count < d.height) {
            // the band's deepest-LOD tiles may only sample rows inside the window (bilinear footprint: +-1 row)
            if (p->shard.world == 1) { BT_SET_ERROR(ctx, "add_datasets: a row window needs a sharded preprocessor"); return BT_ERR_INVALID_ARG; }
            const uint32_t n = 1u << (d.lod_end - 1);
            const double rows_per_tile = (double)d.height / n;
            const double first = std::floor((double)((uint64_t)p->shard.rank * n / p->shard.world) * rows_per_tile) - 1.0;
            const double last = std::ceil((double)((uint64_t)(p->shard.rank + 1) * n / p->shard.world) * rows_per_tile);
            if ((double)d.row_begin > std::fmax(first, 0.0) || (double)(d.row_begin + d.row_count) < std::fmin(last, (double)d.height)) {
                BT_SET_ERROR(ctx, "add_datasets: row window [%u,%u) does not cover rows [%.0f,%.0f) needed by rank %u", d.row_begin, d.row_begin + d.row_count, std::fmax(first, 0.0), std::fmin(last, (double)d.height), p->shard.rank);
                return BT_ERR_INVALID_ARG;
            }
        }
        if (count > 1 && (side_mask & (1u << d.side))) { BT_SET_ERROR(ctx, "add_datasets: datasets of one job must lie on distinct sides"); return BT_ERR_INVALID_ARG; }
        side_mask |= 1u << d.side;
        if (p->shard.world > 1) {
            if (d.top_left[0] != 0.0f || d.top_left[1] != 0.0f || d.bottom_right[0] != 1.0f || d.bottom_right[1] != 1.0f || d.lod_end - 1 < p->root_lod) {
                BT_SET_ERROR(ctx, "add_datasets: sharded preprocessing needs a full-coverage dataset whose deepest LOD has >= world tile rows");
                return BT_ERR_INVALID_ARG;
            }
        }
    }
    if (p->shard.world > 1)
        for (uint32_t i = 1; i < count; ++i)
            if (datasets[i].lod_begin != datasets[0].lod_begin || datasets[i].lod_end != datasets[0].lod_end) {
                BT_SET_ERROR(ctx, "add_datasets: the six faces of a sharded job must share one lod_range");
                return BT_ERR_INVALID_ARG;
            }
    if (!p->plan_only()) BT_CUDA(ctx, cudaSetDevice(ctx->device));
    const Store& s = p->stores[j.attachment];
    for (uint32_t i = 0; i < count; ++i) {
        bt_dataset d = datasets[i];
        if (!d.pixels_on_device && !p->plan_only()) {  // stage the host image on the device (the reference uploads the Image asset, preprocessor.rs:241-246)
            void* dev = nullptr;
            const size_t bytes = (size_t)d.width * (d.row_count ? d.row_count : d.height) * s.px;
            BT_CUDA(ctx, cudaMalloc(&dev, bytes));
            cudaError_t e = cudaMemcpy(dev, d.pixels, bytes, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { cudaFree(dev); BT_CUDA(ctx, e); }
            j.owned_pixels.push_back(dev);
            d.pixels = dev;
            d.pixels_on_device = 1;
        }
        j.datasets.push_back(d);
    }
    p->jobs.push_back(std::move(j));
    return BT_OK;
}

static std::vector<uint32_t> next_wave(const bt_preprocessor* p);
static bt_status build_exchange_tables(bt_preprocessor* p, const std::vector<uint32_t>& wave, cudaStream_t stream);

// Allocates (and first-use-initialises) the device pyramid store for every pending job, so that a subsequent
// run contains no cudaMalloc.  Optional: run_* do the same lazily.
bt_status bt_preprocessor_reserve(bt_preprocessor* p, void* stream) {
    if (!p || p->plan_only()) return BT_ERR_INVALID_ARG;
    BT_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
    for (size_t ji = 0; ji < p->jobs.size(); ++ji) p->allocate_job_tiles((uint32_t)ji);   // size every store for all jobs first
    for (size_t ji = 0; ji < p->jobs.size(); ++ji) {
        if (p->jobs[ji].stitched) continue;
        bt_status st = p->prepare_job((uint32_t)ji, (cudaStream_t)stream);
        if (st != BT_OK) return st;
    }
    if (p->shard.world > 1) {   // the exchange buffers and copy tables of the first wave (bt_preprocessor_run_sharded)
        const std::vector<uint32_t> wave = next_wave(p);
        if (!wave.empty()) { bt_status st = build_exchange_tables(p, wave, (cudaStream_t)stream); if (st != BT_OK) return st; }
    }
    return BT_OK;
}

bt_status bt_preprocessor_run(bt_preprocessor* p, void* stream) {
    if (!p || p->plan_only()) return BT_ERR_INVALID_ARG;
    if (p->shard.world != 1) { BT_SET_ERROR(p->ctx, "bt_preprocessor_run: sharded runs use the phase API (split_downsample / roots / halo / stitch)"); return BT_ERR_INVALID_ARG; }
    BT_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
    for (int ji; (ji = p->pending_job()) >= 0;) {
        bt_status st = p->run_split_downsample_job((uint32_t)ji, (cudaStream_t)stream);
        if (st == BT_OK) st = p->run_stitch_job((uint32_t)ji, (cudaStream_t)stream);
        if (st != BT_OK) return st;
    }
    return BT_OK;
}

bt_status bt_preprocessor_run_split_downsample(bt_preprocessor* p, void* stream) {
    if (!p || p->plan_only()) return BT_ERR_INVALID_ARG;
    const int ji = p->pending_job();
    if (ji < 0 || p->jobs[ji].split_done) { BT_SET_ERROR(p->ctx, "run_split_downsample: no pending job"); return BT_ERR_INVALID_ARG; }
    BT_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
    return p->run_split_downsample_job((uint32_t)ji, (cudaStream_t)stream);
}

bt_status bt_preprocessor_run_stitch(bt_preprocessor* p, void* stream) {
    if (!p || p->plan_only()) return BT_ERR_INVALID_ARG;
    const int ji = p->pending_job();
    if (ji < 0 || !p->jobs[ji].split_done) { BT_SET_ERROR(p->ctx, "run_stitch: run_split_downsample has not run"); return BT_ERR_INVALID_ARG; }
    const Job& j = p->jobs[ji];
    if (p->shard.world > 1 && ((p->roots_bytes(j) && !j.roots_done) || (p->halo_bytes(j) && !j.halo_done))) {
        BT_SET_ERROR(p->ctx, "run_stitch: the sub-root gather / halo exchange has not been completed on this rank");
        return BT_ERR_INVALID_ARG;
    }
    BT_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
    return p->run_stitch_job((uint32_t)ji, (cudaStream_t)stream);
}

// ---- multi-GPU exchange: ONE host-side plan drives both the CUDA pack/unpack and the CPU (gloo) tests -----
// kind 0 = halo strips, 1 = sub-root tiles.  pack: my tiles -> my send buffer (offsets relative to it);
// unpack: gathered buffer [world][bytes_per_rank] -> my ghost / replica tiles (absolute offsets).
static void exchange_plan(const bt_preprocessor* p, const Job& j, uint32_t kind, bool unpack, std::vector<bt_exchange_copy>& out) {
    const Store& s = p->stores[j.attachment];
    const bt_dataset& d = j.datasets[0];
    const uint32_t T = s.cfg.texture_size, B = s.cfg.border_size, C = T - 2 * B;
    const uint32_t rank = p->shard.rank, world = p->shard.world;
    if (kind == 1) {  // one row of the root level per rank and dataset (cube face), tiles x-major
        const uint64_t per_rank = p->roots_bytes(j);
        if (per_rank == 0) return;
        const uint32_t n = 1u << p->root_lod;
        for (uint32_t r = 0; r < world; ++r) {
            if (unpack ? r == rank : r != rank) continue;
            for (uint32_t di = 0; di < j.datasets.size(); ++di)
                for (uint32_t x = 0; x < n; ++x)
                    out.push_back(bt_exchange_copy{bt_tile{j.datasets[di].side, p->root_lod, x, r}, 0u, 0u, T * T, r,
                                                   (unpack ? (uint64_t)r * per_rank : 0ull) + ((uint64_t)di * n + x) * s.tile_bytes});
        }
        return;
    }
    const uint64_t per_rank = p->halo_bytes(j);
    if (per_rank == 0) return;
    const uint64_t row_bytes = (uint64_t)C * s.px;
    uint64_t offset = 0;
    if (p->spherical()) {
        // cube: strips of band-boundary and face-edge tiles (layout: bt_preprocessor::cube_strips_per_side)
        const uint64_t strip = (uint64_t)B * row_bytes;
        uint32_t dataset_of_side[6] = {0, 0, 0, 0, 0, 0};
        for (uint32_t di = 0; di < j.datasets.size(); ++di) dataset_of_side[j.datasets[di].side] = di;
        // emits the row segments of strip `which` (0 top, 1 bottom, 2 left, 3 right) of `tile`, stored at `at` in the buffer
        auto emit = [&](const TileKey& t, uint32_t which, uint32_t src_rank, uint64_t at) {
            const bt_tile bt{t.side, t.lod, t.x, t.y};
            if (which < 2)
                for (uint32_t b = 0; b < B; ++b) out.push_back(bt_exchange_copy{bt, (which == 0 ? B : C) + b, B, C, src_rank, at + (uint64_t)b * row_bytes});
            else
                for (uint32_t i = 0; i < C; ++i) out.push_back(bt_exchange_copy{bt, B + i, which == 2 ? B : C, B, src_rank, at + (uint64_t)i * B * s.px});
        };
        for (uint32_t lod = std::max(d.lod_begin, p->root_lod + 1); lod < d.lod_end; ++lod) {
            const uint32_t n = 1u << lod, h = n / world;
            const uint64_t per_side = p->cube_strips_per_side(lod);
            if (!unpack) {
                const uint32_t y0 = rank * h, y1 = y0 + h;
                for (uint32_t di = 0; di < j.datasets.size(); ++di) {
                    const uint32_t side = j.datasets[di].side;
                    const uint64_t base = offset + (uint64_t)di * per_side * strip;
                    for (uint32_t x = 0; x < n; ++x) {
                        emit(TileKey{side, lod, x, y0}, 0, rank, base + (uint64_t)x * strip);
                        emit(TileKey{side, lod, x, y1 - 1}, 1, rank, base + (uint64_t)(n + x) * strip);
                    }
                    for (uint32_t y = y0; y < y1; ++y) {
                        emit(TileKey{side, lod, 0, y}, 2, rank, base + (uint64_t)(2 * n + (y - y0)) * strip);
                        emit(TileKey{side, lod, n - 1, y}, 3, rank, base + (uint64_t)(2 * n + h + (y - y0)) * strip);
                    }
                }
            } else {
                for (uint32_t slot = 0; slot < s.keys.size(); ++slot) {   // every ghost tile of this LOD: all strips its owner sends
                    const TileKey& g = s.keys[slot];
                    if (s.kind[slot] != kGhost || g.lod != lod) continue;
                    const uint32_t q = p->owner_of_row(lod, g.y), qy0 = q * h;
                    const uint64_t base = (uint64_t)q * per_rank + offset + (uint64_t)dataset_of_side[g.side] * per_side * strip;
                    if (g.y == qy0) emit(g, 0, q, base + (uint64_t)g.x * strip);
                    if (g.y == qy0 + h - 1) emit(g, 1, q, base + (uint64_t)(n + g.x) * strip);
                    if (g.x == 0) emit(g, 2, q, base + (uint64_t)(2 * n + (g.y - qy0)) * strip);
                    if (g.x == n - 1) emit(g, 3, q, base + (uint64_t)(2 * n + h + (g.y - qy0)) * strip);
                }
            }
            offset += (uint64_t)j.datasets.size() * per_side * strip;
        }
        return;
    }
    for (uint32_t lod = std::max(d.lod_begin, p->root_lod + 1); lod < d.lod_end; ++lod) {
        const uint32_t n = 1u << lod;
        uint32_t y0, y1;
        p->owned_rows(lod, &y0, &y1);
        const uint64_t strip = (uint64_t)B * row_bytes, half = (uint64_t)n * strip;
        for (uint32_t x = 0; x < n; ++x)
            for (uint32_t b = 0; b < B; ++b) {
                const uint64_t in_strip = (uint64_t)x * strip + (uint64_t)b * row_bytes;
                if (!unpack) {
                    // top strip = first B centre rows of my first row; bottom strip = last B centre rows of my last row
                    out.push_back(bt_exchange_copy{bt_tile{d.side, lod, x, y0}, B + b, B, C, rank, offset + in_strip});
                    out.push_back(bt_exchange_copy{bt_tile{d.side, lod, x, y1 - 1}, C + b, B, C, rank, offset + half + in_strip});
                } else {
                    if (y0 > 0)   // ghost row above me <- bottom strip of rank-1 (its last centre rows)
                        out.push_back(bt_exchange_copy{bt_tile{d.side, lod, x, y0 - 1}, C + b, B, C, rank - 1, (uint64_t)(rank - 1) * per_rank + offset + half + in_strip});
                    if (y1 < n)   // ghost row below me <- top strip of rank+1 (its first centre rows)
                        out.push_back(bt_exchange_copy{bt_tile{d.side, lod, x, y1}, B + b, B, C, rank + 1, (uint64_t)(rank + 1) * per_rank + offset + in_strip});
                }
            }
        offset += 2 * half;
    }
}

constexpr uint64_t kCopyChunk = 8192;

static bt_status run_exchange(bt_preprocessor* p, const Job& j, uint32_t kind, bool unpack, unsigned char* buffer, cudaStream_t stream) {
    Store& s = p->stores[j.attachment];
    std::vector<bt_exchange_copy> plan;
    exchange_plan(p, j, kind, unpack, plan);
    std::vector<RowCopy> copies;
    copies.reserve(plan.size());
    for (const bt_exchange_copy& c : plan) {
        unsigned char* tile = s.arena + (size_t)s.slot_of.at(TileKey{c.tile.side, c.tile.lod, c.tile.x, c.tile.y}) * s.tile_bytes +
                              ((size_t)c.tile_row * s.cfg.texture_size + c.tile_col) * s.px;
        unsigned char* buf = buffer + c.buffer_offset;
        const uint64_t bytes = (uint64_t)c.pixels * s.px;
        for (uint64_t o = 0; o < bytes; o += kCopyChunk) {   // one warp copies one chunk (whole root tiles are split)
            const uint64_t nb = std::min<uint64_t>(kCopyChunk, bytes - o);
            copies.push_back(unpack ? RowCopy{buf + o, tile + o, nb} : RowCopy{tile + o, buf + o, nb});
        }
    }
    BT_CUDA(p->ctx, cudaStreamSynchronize(stream));  // an earlier launch may still read d_copies
    bt_status st = p->upload(&p->d_copies, &p->d_copies_cap, copies, stream);
    if (st != BT_OK) return st;
    BT_CUDA(p->ctx, bt_launch_row_copy(p->d_copies, (uint32_t)copies.size(), stream));
    return BT_OK;
}

// ---- in-library sharded run: split/downsample -> pack -> ONE grouped NCCL all-gather on the communicator's stream, overlapped
// with the stitch of every tile that needs no remote data -> unpack -> replicated top levels -> boundary stitch -----------------
static bt_status build_exchange_tables(bt_preprocessor* p, const std::vector<uint32_t>& wave, cudaStream_t stream) {
    bt_context* ctx = p->ctx;
    bool all_built = true;
    for (uint32_t ji : wave) all_built = all_built && p->jobs[ji].xt.built;
    if (all_built) return BT_OK;
    // buffer segments: per job [halo | roots], back to back; the gathered buffer holds world x the same segments
    uint64_t send_total = 0;
    for (uint32_t ji : wave) {
        Job& j = p->jobs[ji];
        j.xt.bytes[0] = p->halo_bytes(j);
        j.xt.bytes[1] = p->roots_bytes(j);
        for (int k = 0; k < 2; ++k) { j.xt.send_off[k] = send_total; j.xt.recv_off[k] = send_total * p->shard.world; send_total += (j.xt.bytes[k] + 255) & ~255ull; }
    }
    BT_CUDA(ctx, cudaStreamSynchronize(stream));
    if (send_total > p->send_cap) {
        if (p->d_send) BT_CUDA(ctx, cudaFree(p->d_send));
        if (p->d_recv) BT_CUDA(ctx, cudaFree(p->d_recv));
        p->d_send = p->d_recv = nullptr;
        BT_CUDA(ctx, cudaMalloc((void**)&p->d_send, send_total));
        BT_CUDA(ctx, cudaMalloc((void**)&p->d_recv, send_total * p->shard.world));
        p->send_cap = send_total;
        p->recv_cap = send_total * p->shard.world;
    }
    for (uint32_t ji : wave) {
        Job& j = p->jobs[ji];
        Store& s = p->stores[j.attachment];
        std::vector<RowCopy> copies;
        for (int unpack = 0; unpack < 2; ++unpack)
            for (uint32_t kind = 0; kind < 2; ++kind) {
                std::vector<bt_exchange_copy> plan;
                exchange_plan(p, j, kind, unpack != 0, plan);
                const uint32_t begin = (uint32_t)copies.size();
                unsigned char* buffer = unpack ? p->d_recv + j.xt.recv_off[kind] : p->d_send + j.xt.send_off[kind];
                for (const bt_exchange_copy& c : plan) {
                    unsigned char* tile = s.arena + (size_t)s.slot_of.at(TileKey{c.tile.side, c.tile.lod, c.tile.x, c.tile.y}) * s.tile_bytes +
                                          ((size_t)c.tile_row * s.cfg.texture_size + c.tile_col) * s.px;
                    unsigned char* buf = buffer + c.buffer_offset;
                    const uint64_t bytes = (uint64_t)c.pixels * s.px;
                    for (uint64_t o = 0; o < bytes; o += kCopyChunk) {
                        const uint64_t nb = std::min<uint64_t>(kCopyChunk, bytes - o);
                        copies.push_back(unpack ? RowCopy{buf + o, tile + o, nb} : RowCopy{tile + o, buf + o, nb});
                    }
                }
                (unpack ? j.xt.unpack_begin : j.xt.pack_begin)[kind] = begin;
                (unpack ? j.xt.unpack_count : j.xt.pack_count)[kind] = (uint32_t)copies.size() - begin;
            }
        if (j.xt.d_copies) { BT_CUDA(ctx, cudaFree(j.xt.d_copies)); j.xt.d_copies = nullptr; }
        BT_CUDA(ctx, cudaMalloc((void**)&j.xt.d_copies, std::max<size_t>(copies.size(), 1) * sizeof(RowCopy)));
        if (!copies.empty()) BT_CUDA(ctx, cudaMemcpy(j.xt.d_copies, copies.data(), copies.size() * sizeof(RowCopy), cudaMemcpyHostToDevice));
        j.xt.built = true;
    }
    return BT_OK;
}

// pending jobs are processed in waves of jobs on DISTINCT attachments (successive datasets of one attachment overlay each other
// and must stay ordered); config 5 (height + albedo) is one wave = one collective
static std::vector<uint32_t> next_wave(const bt_preprocessor* p) {
    std::vector<uint32_t> wave;
    uint32_t seen = 0;
    for (uint32_t ji = 0; ji < p->jobs.size(); ++ji) {
        if (p->jobs[ji].stitched) continue;
        const uint32_t bit = 1u << p->jobs[ji].attachment;
        if (seen & bit) break;
        seen |= bit;
        wave.push_back(ji);
    }
    return wave;
}

bt_status bt_preprocessor_run_sharded(bt_preprocessor* p, bt_comm* comm, void* stream_) {
    if (!p || p->plan_only() || !comm) return BT_ERR_INVALID_ARG;
    bt_context* ctx = p->ctx;
    if (comm->world != p->shard.world || comm->rank != p->shard.rank || comm->ctx != ctx) {
        BT_SET_ERROR(ctx, "bt_preprocessor_run_sharded: the communicator (rank %u of %u) does not match the preprocessor's shard (rank %u of %u)",
                     comm->rank, comm->world, p->shard.rank, p->shard.world);
        return BT_ERR_INVALID_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    if (p->shard.world == 1) return bt_preprocessor_run(p, stream_);
    // BT_SHARDED_TIMING=1: per-phase device times of every wave on stderr (diagnostics; synchronises at the end of the call)
    static const bool timing = [] { const char* e = getenv("BT_SHARDED_TIMING"); return e && e[0] == '1'; }();
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    auto mark = [&](const char* name) {
        if (!timing) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, stream);
        marks.emplace_back(name, e);
    };
    for (;;) {
        const std::vector<uint32_t> wave = next_wave(p);
        if (wave.empty()) break;
        bt_status st;
        mark("start");
        for (uint32_t ji : wave) if ((st = p->prepare_job(ji, stream)) != BT_OK) return st;
        if ((st = build_exchange_tables(p, wave, stream)) != BT_OK) return st;
        // 1. communication-free part: split + downsample of this rank's band, down to the shard boundary level
        // (the jobs of a wave run one after the other: launching the R16 and the Rgba8 split side by side on two streams was
        //  measured 12 % SLOWER at N = 2 - their CTAs partition an SM's shared memory and registers badly)
        for (uint32_t ji : wave) if ((st = p->run_split_downsample_job(ji, stream)) != BT_OK) return st;
        mark("split+downsample");
        // 2. pack every job's sub-root tiles and border strips
        std::vector<bt_comm_segment> segs;
        for (uint32_t ji : wave) {
            Job& j = p->jobs[ji];
            for (int k = 0; k < 2; ++k) {
                if (!j.xt.bytes[k]) continue;
                BT_CUDA(ctx, bt_launch_row_copy(j.xt.d_copies + j.xt.pack_begin[k], j.xt.pack_count[k], stream));
                segs.push_back(bt_comm_segment{p->d_send + j.xt.send_off[k], p->d_recv + j.xt.recv_off[k], (size_t)j.xt.bytes[k]});
            }
        }
        mark("pack");
        // 3. ONE grouped collective for the whole wave, on the communicator's stream ...
        BT_CUDA(ctx, cudaEventRecord(comm->ready, stream));
        BT_CUDA(ctx, cudaStreamWaitEvent(comm->stream, comm->ready, 0));
        if ((st = bt_comm_all_gather_group(comm, segs.data(), (uint32_t)segs.size(), comm->stream)) != BT_OK) return st;
        BT_CUDA(ctx, cudaEventRecord(comm->done, comm->stream));
        // 4. ... while the compute stream stitches every tile whose neighbours are all local
        for (uint32_t ji : wave) if ((st = p->run_stitch_job(ji, stream, 0)) != BT_OK) return st;
        mark("interior stitch");
        // 5. behind the collective: unpack, the replicated top levels, and the boundary tiles' stitch
        BT_CUDA(ctx, cudaStreamWaitEvent(stream, comm->done, 0));
        mark("wait for all-gather");
        for (uint32_t ji : wave) {
            Job& j = p->jobs[ji];
            if (j.xt.bytes[1]) {
                BT_CUDA(ctx, bt_launch_row_copy(j.xt.d_copies + j.xt.unpack_begin[1], j.xt.unpack_count[1], stream));
                if ((st = p->run_levels(ji, p->root_lod, j.datasets[0].lod_begin, false, stream)) != BT_OK) return st;
            }
            j.roots_done = true;
            if (j.xt.bytes[0]) BT_CUDA(ctx, bt_launch_row_copy(j.xt.d_copies + j.xt.unpack_begin[0], j.xt.unpack_count[0], stream));
            j.halo_done = true;
            if ((st = p->run_stitch_job(ji, stream, 1)) != BT_OK) return st;
        }
        mark("unpack + top levels + boundary stitch");
    }
    if (timing && !marks.empty()) {
        cudaStreamSynchronize(stream);
        std::string line = "bt_preprocessor_run_sharded rank " + std::to_string(comm->rank) + ":";
        for (size_t i = 1; i < marks.size(); ++i) {
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
            char buf[96];
            snprintf(buf, sizeof(buf), " %s %.3f ms;", marks[i].first, ms);
            line += buf;
        }
        fprintf(stderr, "%s\n", line.c_str());
        for (auto& m : marks) cudaEventDestroy(m.second);
    }
    return BT_OK;
}

// host-side planning entry points (work on a plan-only preprocessor too)
bt_status bt_preprocessor_plan(bt_preprocessor* p) {
    if (!p) return BT_ERR_INVALID_ARG;
    const int ji = p->pending_job();
    if (ji < 0) return BT_ERR_INVALID_ARG;
    return p->prepare_job((uint32_t)ji, nullptr);
}

bt_status bt_preprocessor_plan_tiles(bt_preprocessor* p, uint32_t attachment, bt_tile* tiles, uint32_t* kinds, uint32_t cap, uint32_t* count) {
    if (!p || attachment >= p->stores.size()) return BT_ERR_INVALID_ARG;
    const Store& s = p->stores[attachment];
    for (uint32_t i = 0; i < s.keys.size() && i < cap; ++i) {
        if (tiles) tiles[i] = bt_tile{s.keys[i].side, s.keys[i].lod, s.keys[i].x, s.keys[i].y};
        if (kinds) kinds[i] = s.kind[i];
    }
    if (count) *count = (uint32_t)s.keys.size();
    return BT_OK;
}

bt_status bt_preprocessor_plan_exchange(bt_preprocessor* p, uint32_t kind, uint32_t unpack, bt_exchange_copy* out, uint32_t cap, uint32_t* count) {
    if (!p || kind > 1) return BT_ERR_INVALID_ARG;
    const int ji = p->pending_job();
    if (ji < 0) return BT_ERR_INVALID_ARG;
    std::vector<bt_exchange_copy> plan;
    exchange_plan(p, p->jobs[ji], kind, unpack != 0, plan);
    if (out) for (size_t i = 0; i < plan.size() && i < cap; ++i) out[i] = plan[i];
    if (count) *count = (uint32_t)plan.size();
    return BT_OK;
}

bt_status bt_preprocessor_roots_bytes(bt_preprocessor* p, uint64_t* send_bytes_per_rank) {
    if (!p || !send_bytes_per_rank) return BT_ERR_INVALID_ARG;
    const int ji = p->pending_job();
    *send_bytes_per_rank = ji < 0 ? 0 : p->roots_bytes(p->jobs[ji]);
    return BT_OK;
}

bt_status bt_preprocessor_halo_bytes(bt_preprocessor* p, uint64_t* send_bytes_per_rank) {
    if (!p || !send_bytes_per_rank) return BT_ERR_INVALID_ARG;
    const int ji = p->pending_job();
    *send_bytes_per_rank = ji < 0 ? 0 : p->halo_bytes(p->jobs[ji]);
    return BT_OK;
}

static bt_status exchange_entry(bt_preprocessor* p, uint32_t kind, bool unpack, void* buffer, void* stream) {
    if (!p || p->plan_only()) return BT_ERR_INVALID_ARG;
    const int ji = p->pending_job();
    if (ji < 0 || !p->jobs[ji].split_done) { BT_SET_ERROR(p->ctx, "exchange: run_split_downsample has not run"); return BT_ERR_INVALID_ARG; }
    Job& j = p->jobs[ji];
    const uint64_t bytes = kind == 1 ? p->roots_bytes(j) : p->halo_bytes(j);
    if (bytes) {
        if (!buffer) return BT_ERR_INVALID_ARG;
        BT_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
        bt_status st = run_exchange(p, j, kind, unpack, static_cast<unsigned char*>(buffer), (cudaStream_t)stream);
        if (st != BT_OK) return st;
    }
    if (unpack && kind == 1) {
        if (bytes) {  // the levels above the shard boundary: every rank computes them (21 tiles at world = 8)
            bt_status st = p->run_levels((uint32_t)ji, p->root_lod, j.datasets[0].lod_begin, false, (cudaStream_t)stream);   // (all datasets of a sharded job share one lod_range)
            if (st != BT_OK) return st;
        }
        j.roots_done = true;
    }
    if (unpack && kind == 0) j.halo_done = true;
    return BT_OK;
}

bt_status bt_preprocessor_roots_pack(bt_preprocessor* p, void* device_send, void* stream) { return exchange_entry(p, 1, false, device_send, stream); }
bt_status bt_preprocessor_roots_unpack_and_finish(bt_preprocessor* p, const void* device_gathered, void* stream) {
    return exchange_entry(p, 1, true, const_cast<void*>(device_gathered), stream);
}
bt_status bt_preprocessor_halo_pack(bt_preprocessor* p, void* device_send, void* stream) { return exchange_entry(p, 0, false, device_send, stream); }
bt_status bt_preprocessor_halo_unpack(bt_preprocessor* p, const void* device_gathered, void* stream) {
    return exchange_entry(p, 0, true, const_cast<void*>(device_gathered), stream);
}

// ---- results --------------------------------------------------------------------------------------
static bool is_output(const bt_preprocessor* p, const Store& s, uint32_t slot) { (void)p; return s.kind[slot] == kOwned; }

bt_status bt_preprocessor_tile_count(bt_preprocessor* p, uint32_t attachment, uint32_t* count) {
    if (!p || !count || attachment >= p->stores.size()) return BT_ERR_INVALID_ARG;
    const Store& s = p->stores[attachment];
    uint32_t n = 0;
    for (uint32_t i = 0; i < s.keys.size(); ++i) n += is_output(p, s, i);
    *count = n;
    return BT_OK;
}

bt_status bt_preprocessor_list_tiles(bt_preprocessor* p, uint32_t attachment, bt_tile* tiles, uint32_t cap, uint32_t* count) {
    if (!p || attachment >= p->stores.size()) return BT_ERR_INVALID_ARG;
    const Store& s = p->stores[attachment];
    uint32_t n = 0;
    for (uint32_t i = 0; i < s.keys.size(); ++i) {
        if (!is_output(p, s, i)) continue;
        if (tiles && n < cap) tiles[n] = bt_tile{s.keys[i].side, s.keys[i].lod, s.keys[i].x, s.keys[i].y};
        ++n;
    }
    if (count) *count = n;
    return BT_OK;
}

bt_status bt_preprocessor_tile_device_ptr(bt_preprocessor* p, uint32_t attachment, const bt_tile* tile, void** out) {
    if (!p || !tile || !out || attachment >= p->stores.size()) return BT_ERR_INVALID_ARG;
    const Store& s = p->stores[attachment];
    auto it = s.slot_of.find(TileKey{tile->side, tile->lod, tile->x, tile->y});
    if (it == s.slot_of.end() || !s.arena || it->second >= s.arena_slots) return BT_ERR_NOT_FOUND;
    *out = s.arena + (size_t)it->second * s.tile_bytes;
    return BT_OK;
}

bt_status bt_preprocessor_get_tile(bt_preprocessor* p, uint32_t attachment, const bt_tile* tile, void* host_out, size_t bytes) {
    void* dev = nullptr;
    if (!host_out || !p || p->plan_only()) return BT_ERR_INVALID_ARG;
    bt_status st = bt_preprocessor_tile_device_ptr(p, attachment, tile, &dev);
    if (st != BT_OK) return st;
    if (bytes != p->stores[attachment].tile_bytes) { BT_SET_ERROR(p->ctx, "get_tile: expected %zu bytes", p->stores[attachment].tile_bytes); return BT_ERR_INVALID_ARG; }
    BT_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
    BT_CUDA(p->ctx, cudaMemcpy(host_out, dev, bytes, cudaMemcpyDeviceToHost));
    return BT_OK;
}

// Order-independent 64-bit content hash of this rank's OUTPUT tiles (coordinate-keyed): the sum over ranks of a sharded run
// equals the hash of the single-GPU pyramid iff every tile is byte-identical (up to 2^-64 collisions).  Blocking.
bt_status bt_preprocessor_hash(bt_preprocessor* p, uint32_t attachment, uint64_t* out, void* stream) {
    if (!p || p->plan_only() || !out || attachment >= p->stores.size()) return BT_ERR_INVALID_ARG;
    bt_context* ctx = p->ctx;
    const Store& s = p->stores[attachment];
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<btd::TileHashTask> tasks;
    for (uint32_t i = 0; i < s.keys.size(); ++i) {
        if (!is_output(p, s, i)) continue;
        const TileKey& k = s.keys[i];
        uint64_t key = 0x9e3779b97f4a7c15ull;
        for (uint32_t w : {k.side, k.lod, k.x, k.y}) { key ^= w; key *= 0x100000001b3ull; key ^= key >> 29; }
        tasks.push_back(btd::TileHashTask{key, i, 0u});
    }
    *out = 0;
    if (tasks.empty()) return BT_OK;
    btd::TileHashTask* d_tasks = nullptr;
    unsigned long long* d_out = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    BT_CUDA(ctx, cudaMalloc((void**)&d_tasks, tasks.size() * sizeof(btd::TileHashTask)));
    cudaError_t e = cudaMalloc((void**)&d_out, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_out, 0, sizeof(unsigned long long), st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_tasks, tasks.data(), tasks.size() * sizeof(btd::TileHashTask), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = bt_launch_tile_hash(s.arena, s.tile_bytes, d_tasks, (uint32_t)tasks.size(), d_out, st);
    unsigned long long h = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_tasks);
    cudaFree(d_out);
    BT_CUDA(ctx, e);
    *out = (uint64_t)h;
    return BT_OK;
}

// bincode 2 "standard" varint (formats/mod.rs:14-24)
static void put_varint(std::vector<unsigned char>& out, uint64_t v) {
    if (v < 251) out.push_back((unsigned char)v);
    else if (v < (1ull << 16)) { out.push_back(251); for (int i = 0; i < 2; ++i) out.push_back((unsigned char)(v >> (8 * i))); }
    else if (v < (1ull << 32)) { out.push_back(252); for (int i = 0; i < 4; ++i) out.push_back((unsigned char)(v >> (8 * i))); }
    else { out.push_back(253); for (int i = 0; i < 8; ++i) out.push_back((unsigned char)(v >> (8 * i))); }
}

// Save tasks + save_tile_config (tile_atlas.rs:77-116,605-613): raw mip-0 texels per tile and the
// bincode tile list.  Files go to <directory>/<side>_<lod>_<x>_<y>.bin and <directory>/config.tc.
bt_status bt_preprocessor_save(bt_preprocessor* p, uint32_t attachment, const char* directory) {
    if (!p || p->plan_only() || !directory || attachment >= p->stores.size()) return BT_ERR_INVALID_ARG;
    bt_context* ctx = p->ctx;
    const Store& s = p->stores[attachment];
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    BT_CUDA(ctx, cudaDeviceSynchronize());
    std::vector<unsigned char> host(s.tile_bytes), tc;
    uint64_t n = 0;
    for (uint32_t i = 0; i < s.keys.size(); ++i) n += is_output(p, s, i);
    put_varint(tc, n);
    for (uint32_t i = 0; i < s.keys.size(); ++i) {
        if (!is_output(p, s, i)) continue;
        const TileKey& k = s.keys[i];
        BT_CUDA(ctx, cudaMemcpy(host.data(), s.arena + (size_t)i * s.tile_bytes, s.tile_bytes, cudaMemcpyDeviceToHost));
        const std::string path = std::string(directory) + "/" + std::to_string(k.side) + "_" + std::to_string(k.lod) + "_" + std::to_string(k.x) + "_" + std::to_string(k.y) + ".bin";
        FILE* f = std::fopen(path.c_str(), "wb");
        if (!f || std::fwrite(host.data(), 1, host.size(), f) != host.size()) {
            if (f) std::fclose(f);
            BT_SET_ERROR(ctx, "bt_preprocessor_save: cannot write %s", path.c_str());
            return BT_ERR_IO;
        }
        std::fclose(f);
        put_varint(tc, k.side); put_varint(tc, k.lod); put_varint(tc, k.x); put_varint(tc, k.y);
    }
    const std::string path = std::string(directory) + "/config.tc";
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f || std::fwrite(tc.data(), 1, tc.size(), f) != tc.size()) {
        if (f) std::fclose(f);
        BT_SET_ERROR(ctx, "bt_preprocessor_save: cannot write %s", path.c_str());
        return BT_ERR_IO;
    }
    std::fclose(f);
    return BT_OK;
}

}  // extern "C"
