// host_math.cpp — f64 host-side producers of the hot path's per-frame inputs (no GPU needed).
//
// C++ counterparts of the reference's main-world Rust code that feeds the prepass and the vertex
// stage: TerrainModel (src/math/terrain_model.rs:40-203), Coordinate (src/math/coordinate.rs:56-152),
// TerrainModelApproximation::compute (terrain_model.rs:263-360), TileTree
// (src/terrain_data/tile_tree.rs:94-386) and the request/release/best-tile bookkeeping of
// TileAtlasState (src/terrain_data/tile_atlas.rs:369-503).  Exposed through the C ABI so a Rust /
// Python / C++ host can drive the kernels with exactly the inputs the reference would upload.
#include <cmath>
#include <cstring>
#include <deque>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/bevy_terrain_b200.h"
#include "bt_runtime.h"   // only the device clipmap at the end of this file touches the GPU

namespace {

constexpr double kCSqr = 0.87 * 0.87;  // src/math/mod.rs:13

struct D2 { double x, y; };
struct D3 { double x, y, z; };

inline D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline D3 operator*(double s, D3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline D3 operator/(D3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline double length(D3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline D3 normalized(D3 a) { return a / length(a); }

inline bool spherical(const bt_terrain_model& m) { return m.kind != BT_KIND_PLANAR; }

// world_from_local = from_scale_rotation_translation(scale, IDENTITY, translation) — terrain_model.rs:61-82
inline D3 transform_point(const bt_terrain_model& m, D3 p) {
    return {m.scale[0] * p.x + m.position[0], m.scale[1] * p.y + m.position[1], m.scale[2] * p.z + m.position[2]};
}
inline D3 transform_vector(const bt_terrain_model& m, D3 p) { return {m.scale[0] * p.x, m.scale[1] * p.y, m.scale[2] * p.z}; }
inline D3 inverse_transform_point(const bt_terrain_model& m, D3 p) {
    return {(p.x - m.position[0]) / m.scale[0], (p.y - m.position[1]) / m.scale[1], (p.z - m.position[2]) / m.scale[2]};
}

// ---- closest point on an axis-aligned ellipsoid (src/math/ellipsoid.rs; the algorithm of D. Eberly, "Distance from a Point to
// an Ellipse, an Ellipsoid, or a Hyperellipsoid", Geometric Tools): work in the first octant with semi-axes e0 >= e1 >= e2, find
// the root t of  sum_i (r_i z_i / (t + r_i))^2 = 1  by bisection (r_i = e_i^2 / e_last^2, z_i = y_i / e_i), then
// x_i = r_i y_i / (t + r_i).  The iteration count and the termination tests follow the reference so that the f64 results agree.
constexpr int kEllipsoidIterations = 1074;   // ellipsoid.rs:8-10: after that many halvings s0 == s1

template <int N>
double ellipsoid_root(const double (&r)[N], const double (&z)[N], double g) {
    double n[N], nn = 0.0;
    for (int i = 0; i < N; ++i) { n[i] = r[i] * z[i]; nn += n[i] * n[i]; }
    double s0 = z[N - 1] - 1.0, s1 = g < 0.0 ? 0.0 : std::sqrt(nn) - 1.0, s = 0.0;
    for (int it = 0; it < kEllipsoidIterations; ++it) {
        s = (s0 + s1) / 2.0;
        if (s == s0 || s == s1) break;
        double q = 0.0;
        for (int i = 0; i < N; ++i) { const double ratio = n[i] / (s + r[i]); q += ratio * ratio; }
        const double gg = q - 1.0;
        if (gg < 0.0) s1 = s; else if (gg > 0.0) s0 = s; else break;
    }
    return s;
}

// ellipse with semi-axes e0 >= e1, query (y0, y1) in the first quadrant
void project_point_ellipse(double e0, double e1, double y0, double y1, double out[2]) {
    if (y1 > 0.0) {
        if (y0 > 0.0) {
            const double z[2] = {y0 / e0, y1 / e1};
            const double g = (z[0] * z[0] + z[1] * z[1]) - 1.0;
            if (g != 0.0) {
                const double r[2] = {(e0 * e0) / (e1 * e1), 1.0};
                const double t = ellipsoid_root<2>(r, z, g);
                out[0] = r[0] * y0 / (t + r[0]);
                out[1] = r[1] * y1 / (t + r[1]);
            } else { out[0] = y0; out[1] = y1; }
        } else { out[0] = 0.0; out[1] = e1; }
    } else {
        const double numer0 = e0 * y0, denom0 = e0 * e0 - e1 * e1;
        if (numer0 < denom0) {
            const double xde0 = numer0 / denom0;
            out[0] = e0 * xde0;
            out[1] = e1 * std::sqrt(1.0 - xde0 * xde0);
        } else { out[0] = e0; out[1] = 0.0; }
    }
}

inline double signum(double v) { return std::isnan(v) ? v : (std::signbit(v) ? -1.0 : 1.0); }   // f64::signum: signum(+-0) = +-1

// project_point_ellipsoid(e, y) — ellipsoid.rs:12-63; e = (e0, e1, e2) descending, the query is swizzled (x, z, y) so that the
// terrain's polar axis (world y, the minor axis) lands in the last slot
D3 project_point_ellipsoid(const double e[3], D3 query) {
    const double sign[3] = {signum(query.x), signum(query.y), signum(query.z)};
    const double y[3] = {std::fabs(query.x), std::fabs(query.z), std::fabs(query.y)};
    double x[3];
    if (y[2] > 0.0) {
        if (y[1] > 0.0) {
            if (y[0] > 0.0) {
                const double z[3] = {y[0] / e[0], y[1] / e[1], y[2] / e[2]};
                const double g = ((z[0] * z[0] + z[1] * z[1]) + z[2] * z[2]) - 1.0;
                if (g != 0.0) {
                    const double r[3] = {(e[0] * e[0]) / (e[2] * e[2]), (e[1] * e[1]) / (e[2] * e[2]), 1.0};
                    const double t = ellipsoid_root<3>(r, z, g);
                    for (int i = 0; i < 3; ++i) x[i] = r[i] * y[i] / (t + r[i]);
                } else { x[0] = y[0]; x[1] = y[1]; x[2] = y[2]; }
            } else {
                double p[2];
                project_point_ellipse(e[1], e[2], y[1], y[2], p);
                x[0] = 0.0; x[1] = p[0]; x[2] = p[1];
            }
        } else if (y[0] > 0.0) {
            double p[2];
            project_point_ellipse(e[0], e[2], y[0], y[2], p);
            x[0] = p[0]; x[1] = 0.0; x[2] = p[1];
        } else { x[0] = 0.0; x[1] = 0.0; x[2] = e[2]; }
    } else {
        const double denom0 = e[0] * e[0] - e[2] * e[2], denom1 = e[1] * e[1] - e[2] * e[2];
        const double numer0 = e[0] * y[0], numer1 = e[1] * y[1];
        bool done = false;
        if (numer0 < denom0 && numer1 < denom1) {
            const double xde0 = numer0 / denom0, xde1 = numer1 / denom1;
            const double discr = (1.0 - xde0 * xde0) - xde1 * xde1;
            if (discr > 0.0) { x[0] = e[0] * xde0; x[1] = e[1] * xde1; x[2] = e[2] * std::sqrt(discr); done = true; }
        }
        if (!done) {
            double p[2];
            project_point_ellipse(e[0], e[1], y[0], y[1], p);
            x[0] = p[0]; x[1] = p[1]; x[2] = 0.0;
        }
    }
    return {sign[0] * x[0], sign[1] * x[2], sign[2] * x[1]};   // swizzle back: (x, z, y)
}

// TerrainModel::position_world_to_local — terrain_model.rs:145-168
D3 position_world_to_local(const bt_terrain_model& m, D3 world) {
    if (m.kind == BT_KIND_ELLIPSOIDAL) {
        // ellipsoid_from_world = inverse(rotation * translation) with the identity rotation: subtract the position
        const D3 rel{world.x - m.position[0], world.y - m.position[1], world.z - m.position[2]};
        const double e[3] = {m.scale[0], m.scale[0], m.scale[1]};   // (major, major, minor)
        // (literal reference behaviour: local_from_world is applied to the point in the ELLIPSOID frame, terrain_model.rs:163-166,
        //  i.e. the translation is subtracted twice — invisible for the reference's own use, an ellipsoid at the origin)
        return normalized(inverse_transform_point(m, project_point_ellipsoid(e, rel)));
    }
    D3 l = inverse_transform_point(m, world);
    if (m.kind == BT_KIND_PLANAR) return {l.x, 0.0 * l.y, l.z};
    return normalized(l);
}

// TerrainModel::position_local_to_world — terrain_model.rs:131-143
D3 position_local_to_world(const bt_terrain_model& m, D3 local, double height) {
    const D3 wp = transform_point(m, local);
    const D3 wn = normalized(transform_vector(m, spherical(m) ? local : D3{0.0, 1.0, 0.0}));
    return wp + height * wn;
}

struct Coordinate { uint32_t side; D2 uv; };

// Coordinate::from_world_position — coordinate.rs:69-108
Coordinate coordinate_from_world(const bt_terrain_model& m, D3 world) {
    const D3 n = position_world_to_local(m, world);
    if (!spherical(m)) {
        auto clamp01 = [](double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); };
        return {0u, {clamp01(n.x + 0.5), clamp01(n.z + 0.5)}};
    }
    const double ax = std::fabs(n.x), ay = std::fabs(n.y), az = std::fabs(n.z);
    uint32_t side;
    D2 q;
    if (ax > ay && ax > az) {
        if (n.x < 0.0) { side = 0; q = {-n.z / n.x, n.y / n.x}; }
        else { side = 3; q = {-n.y / n.x, n.z / n.x}; }
    } else if (az > ay) {
        if (n.z > 0.0) { side = 1; q = {n.x / n.z, -n.y / n.z}; }
        else { side = 4; q = {n.y / n.z, -n.x / n.z}; }
    } else {
        if (n.y > 0.0) { side = 2; q = {n.x / n.y, n.z / n.y}; }
        else { side = 5; q = {-n.z / n.y, -n.x / n.y}; }
    }
    auto warp = [](double v) { return 0.5 * (v * std::pow((1.0 + kCSqr) / (1.0 + kCSqr * v * v), 0.5)) + 0.5; };
    return {side, {warp(q.x), warp(q.y)}};
}

// Coordinate::world_position — coordinate.rs:110-130
D3 coordinate_world_position(const bt_terrain_model& m, Coordinate c, float height) {
    D3 lp;
    if (spherical(m)) {
        auto unwarp = [](double v) { const double w = (v - 0.5) / 0.5; return w / std::pow(1.0 + kCSqr - kCSqr * w * w, 0.5); };
        const double u = unwarp(c.uv.x), v = unwarp(c.uv.y);
        switch (c.side) {
            case 0: lp = {-1.0, -v, u}; break;
            case 1: lp = {u, -v, 1.0}; break;
            case 2: lp = {u, 1.0, v}; break;
            case 3: lp = {1.0, -u, v}; break;
            case 4: lp = {v, -u, -1.0}; break;
            default: lp = {v, -1.0, u}; break;
        }
        lp = normalized(lp);
    } else {
        lp = {c.uv.x - 0.5, 0.0, c.uv.y - 0.5};
    }
    return position_local_to_world(m, lp, (double)height);
}

// SideInfo::project_to_side + Coordinate::project_to_side — coordinate.rs:19-53,134-151
enum SideInfo : uint8_t { kFixed0, kFixed1, kPosS, kPosT };
constexpr SideInfo kEven[6][2] = {{kPosS, kPosT}, {kFixed0, kPosT}, {kFixed0, kPosS}, {kPosT, kPosS}, {kPosT, kFixed0}, {kPosS, kFixed0}};
constexpr SideInfo kOdd[6][2] = {{kPosS, kPosT}, {kPosS, kFixed1}, {kPosT, kFixed1}, {kPosT, kPosS}, {kFixed1, kPosS}, {kFixed1, kPosT}};

Coordinate project_to_side(const bt_terrain_model& m, Coordinate c, uint32_t side) {
    if (!spherical(m)) return c;
    const uint32_t index = (6 + side - c.side) % 6;
    const SideInfo* info = (c.side % 2 == 0) ? kEven[index] : kOdd[index];
    auto pick = [&](SideInfo i) { return i == kFixed0 ? 0.0 : i == kFixed1 ? 1.0 : i == kPosS ? c.uv.x : c.uv.y; };
    return {side, {pick(info[0]), pick(info[1])}};
}

// integer power by squaring, the evaluation f64::powi lowers to
double powi(double a, int b) {
    double r = 1.0;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return r;
}

// SIDE_MATRICES — terrain_model.rs:14-21 (column-major)
constexpr double kSideMatrices[6][9] = {
    {-1, 0, 0, 0, 0, 1, 0, -1, 0}, {0, 0, 1, 1, 0, 0, 0, -1, 0}, {0, 1, 0, 1, 0, 0, 0, 0, 1},
    {1, 0, 0, 0, -1, 0, 0, 0, 1},  {0, 0, -1, 0, -1, 0, 1, 0, 0}, {0, -1, 0, 0, 0, 1, 1, 0, 0}};
D3 side_mul(uint32_t side, D3 v) {
    const double* m = kSideMatrices[side];
    return {m[0] * v.x + m[3] * v.y + m[6] * v.z, m[1] * v.x + m[4] * v.y + m[7] * v.z, m[2] * v.x + m[5] * v.y + m[8] * v.z};
}

int32_t to_i32_saturating(double x) {
    if (x != x) return 0;
    if (x <= -2147483648.0) return INT32_MIN;
    if (x >= 2147483647.0) return INT32_MAX;
    return (int32_t)x;
}
uint32_t to_u32_saturating(double x) {
    if (!(x > 0.0)) return 0u;
    if (x >= 4294967295.0) return 0xFFFFFFFFu;
    return (uint32_t)x;
}

void store3(float out[3], D3 v) { out[0] = (float)v.x; out[1] = (float)v.y; out[2] = (float)v.z; }

}  // namespace

extern "C" {

void bt_terrain_model_planar(const double position[3], double side_length, float min_height, float max_height, bt_terrain_model* out) {
    std::memset(out, 0, sizeof(*out));
    out->kind = BT_KIND_PLANAR;
    for (int k = 0; k < 3; ++k) { out->position[k] = position[k]; out->scale[k] = side_length; }
    out->min_height = min_height;
    out->max_height = max_height;
}

void bt_terrain_model_sphere(const double position[3], double radius, float min_height, float max_height, bt_terrain_model* out) {
    std::memset(out, 0, sizeof(*out));
    out->kind = BT_KIND_SPHERICAL;
    for (int k = 0; k < 3; ++k) { out->position[k] = position[k]; out->scale[k] = radius; }
    out->min_height = min_height;
    out->max_height = max_height;
}

void bt_terrain_model_ellipsoid(const double position[3], double major_axis, double minor_axis, float min_height, float max_height,
                                bt_terrain_model* out) {   // terrain_model.rs:106-128: scale = (major, minor, major), identity rotation
    std::memset(out, 0, sizeof(*out));
    out->kind = BT_KIND_ELLIPSOIDAL;
    for (int k = 0; k < 3; ++k) { out->position[k] = position[k]; out->scale[k] = k == 1 ? minor_axis : major_axis; }
    out->min_height = min_height;
    out->max_height = max_height;
}

void bt_terrain_view_config_default(bt_terrain_view_config* out) {  // terrain_view.rs:47-64
    std::memset(out, 0, sizeof(*out));
    out->tree_size = 8;
    out->geometry_tile_count = 1000000;
    out->refinement_count = 30;
    out->grid_size = 16;
    out->subdivision_tolerance = 0.1;
    out->load_distance = 2.5;
    out->morph_distance = 16.0;
    out->blend_distance = 2.0;
    out->morph_range = 0.2f;
    out->blend_range = 0.2f;
    out->precision_threshold_distance = 0.001;
    out->origin_lod = 10;
}

double bt_terrain_model_scale(const bt_terrain_model* model) {  // terrain_model.rs:183-193
    if (model->kind == BT_KIND_ELLIPSOIDAL) return (model->scale[0] + model->scale[1]) / 2.0;   // (major + minor) / 2
    return model->kind == BT_KIND_PLANAR ? model->scale[0] / 2.0 : model->scale[0];
}

void bt_terrain_model_mesh_transform(const bt_terrain_model* model, bt_mesh_transform* out) {
    // TerrainModel::transform (f32 translation/scale, terrain_model.rs:196-203) -> MeshUniform:
    // affine columns + inverse-transpose of the upper 3x3 (diagonal for a scale-only transform)
    std::memset(out, 0, sizeof(*out));
    for (int k = 0; k < 3; ++k) {
        const float s = (float)model->scale[k];
        out->world_from_local[4 * k] = s;  // (0,0) (1,1) (2,2) of the column-major 3x4
        out->world_from_local[9 + k] = (float)model->position[k];
        out->normal_matrix[4 * k] = 1.0f / s;
    }
}

void bt_view_config_uniform_from(const bt_terrain_view_config* cfg, const bt_terrain_model* model, bt_view_config_uniform* out) {
    const double scale = bt_terrain_model_scale(model);  // TileTree::new — tile_tree.rs:135-153
    out->tree_size = cfg->tree_size;
    out->geometry_tile_count = cfg->geometry_tile_count;
    out->refinement_count = cfg->refinement_count;
    out->grid_size = (float)cfg->grid_size;
    out->vertices_per_row = 2 * (cfg->grid_size + 2);
    out->vertices_per_tile = 2 * cfg->grid_size * (cfg->grid_size + 2);
    out->morph_distance = (float)(cfg->morph_distance * scale);
    out->blend_distance = (float)(cfg->blend_distance * scale);
    out->load_distance = (float)(cfg->load_distance * scale);
    out->subdivision_distance = (float)(cfg->morph_distance * scale * (1.0 + cfg->subdivision_tolerance));
    out->morph_range = cfg->morph_range;
    out->blend_range = cfg->blend_range;
    out->precision_threshold_distance = (float)(cfg->precision_threshold_distance * scale);
}

// TerrainModelApproximation::compute — terrain_model.rs:263-360
bt_status bt_model_approximation_compute(const bt_terrain_model* model, const double view_world_position[3],
                                         uint32_t origin_lod, float approximate_height, bt_model_approximation* out) {
    if (!model || !view_world_position || !out || origin_lod > 30) return BT_ERR_INVALID_ARG;
    std::memset(out, 0, sizeof(*out));
    const D3 view{view_world_position[0], view_world_position[1], view_world_position[2]};
    const double origin_count = (double)(1u << origin_lod);
    const Coordinate view_coordinate = coordinate_from_world(*model, view);
    for (uint32_t side = 0; side < 6; ++side) {
        const Coordinate vc = project_to_side(*model, view_coordinate, side);
        bt_side_parameter& sp = out->sides[side];
        const double sx = vc.uv.x * origin_count, sy = vc.uv.y * origin_count;
        sp.view_xy[0] = to_i32_saturating(sx);
        sp.view_xy[1] = to_i32_saturating(sy);
        sp.view_uv[0] = (float)(sx - std::floor(sx));
        sp.view_uv[1] = (float)(sy - std::floor(sy));

        const double s = vc.uv.x, t = vc.uv.y;
        const double u_denom = std::sqrt(1.0 - 4.0 * kCSqr * s * (s - 1.0));
        const double u = (2.0 * s - 1.0) / u_denom;
        const double u_ds = 2.0 * (kCSqr + 1.0) / powi(u_denom, 3);
        const double u_dss = 12.0 * kCSqr * (kCSqr + 1.0) * (2.0 * s - 1.0) / powi(u_denom, 5);
        const double v_denom = std::sqrt(1.0 - 4.0 * kCSqr * t * (t - 1.0));
        const double v = (2.0 * t - 1.0) / v_denom;
        const double v_dt = 2.0 * (kCSqr + 1.0) / powi(v_denom, 3);
        const double v_dtt = 12.0 * kCSqr * (kCSqr + 1.0) * (2.0 * t - 1.0) / powi(v_denom, 5);

        const double l = std::sqrt(1.0 + u * u + v * v);
        const double l_ds = u * u_ds / l;
        const double l_dt = v * v_dt / l;
        const double l_dss = (u * u_dss * l * l + (v * v + 1.0) * u_ds * u_ds) / powi(l, 3);
        const double l_dst = -(u * v * u_ds * v_dt) / powi(l, 3);
        const double l_dtt = (v * v_dtt * l * l + (u * u + 1.0) * v_dt * v_dt) / powi(l, 3);

        const D3 p0{1.0, u, v};
        const D3 p_ds{-l_ds, -u * l_ds + l * u_ds, -v * l_ds};
        const D3 p_dt{-l_dt, -u * l_dt, -v * l_dt + l * v_dt};
        const D3 p_dss{2.0 * l_ds * l_ds - l * l_dss,
                       2.0 * u * l_ds * l_ds - l * (2.0 * u_ds * l_ds + u * l_dss) + u_dss * l * l,
                       2.0 * v * l_ds * l_ds - l * v * l_dss};
        const D3 p_dst{2.0 * l_ds * l_dt - l * l_dst,
                       2.0 * u * l_ds * l_dt - l * (u_ds * l_dt + u * l_dst),
                       2.0 * v * l_ds * l_dt - l * (v_dt * l_ds + v * l_dst)};
        const D3 p_dtt{2.0 * l_dt * l_dt - l * l_dtt,
                       2.0 * u * l_dt * l_dt - l * u * l_dtt,
                       2.0 * v * l_dt * l_dt - l * (2.0 * v_dt * l_dt + v * l_dtt) + v_dtt * l * l};

        const double l2 = powi(l, 2), l3 = powi(l, 3);
        const D3 p = transform_point(*model, side_mul(side, p0) / l);
        store3(sp.c, p - view);
        store3(sp.c_s, transform_vector(*model, side_mul(side, p_ds) / l2));
        store3(sp.c_t, transform_vector(*model, side_mul(side, p_dt) / l2));
        store3(sp.c_ss, transform_vector(*model, side_mul(side, p_dss) / l3) / 2.0);
        store3(sp.c_st, transform_vector(*model, side_mul(side, p_dst) / l3));
        store3(sp.c_tt, transform_vector(*model, side_mul(side, p_dtt) / l3) / 2.0);
    }
    out->origin_lod = origin_lod;
    out->approximate_height = approximate_height;
    return BT_OK;
}

}  // extern "C"

// ================================================================================================
// TileTree (clipmap residency) + the TileAtlasState bookkeeping it drives
// ================================================================================================
namespace {

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
constexpr TileKey kInvalidTile{0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};

struct AtlasTileState { uint32_t requests; bool loaded; uint32_t atlas_index; };
struct TreeSlot { TileKey coordinate = kInvalidTile; bool requested = false; };

}  // namespace

// TileAtlasState — tile_atlas.rs:272-300.  Per TERRAIN in the reference and shared by every view's TileTree: slot allocation,
// the LRU of unused slots, the set of tiles that exist on disk and the load queue.  Several bt_tile_tree handles created with
// bt_tile_tree_create_shared reference one state; bt_tile_tree_create makes a private one (single-view terrains).
struct bt_atlas_state {
    uint32_t atlas_size = 0;
    uint32_t users = 0;            // trees referencing this state (a private state dies with its tree)
    bool owned_by_tree = false;
    std::unordered_map<TileKey, AtlasTileState, TileKeyHash> tile_states;
    std::deque<std::pair<TileKey, uint32_t>> unused_tiles;  // (coordinate, atlas_index), LRU order
    std::unordered_set<TileKey, TileKeyHash> existing_tiles;
    std::vector<bt_tile_load> pending_loads;
    uint64_t loaded_version = 0;   // bumped whenever the set of loaded tiles may have changed (device clipmaps re-upload their copy)

    explicit bt_atlas_state(uint32_t n) : atlas_size(n) {
        for (uint32_t i = 0; i < n; ++i) unused_tiles.emplace_back(kInvalidTile, i);  // TileAtlasState::new — tile_atlas.rs:302-325
    }
    // TileAtlasState::allocate_tile — tile_atlas.rs:383-389
    bool allocate_tile(uint32_t* atlas_index) {
        if (unused_tiles.empty()) return false;  // reference: expect("Atlas out of indices")
        auto t = unused_tiles.front();
        unused_tiles.pop_front();
        tile_states.erase(t.first);
        loaded_version += 1;
        *atlas_index = t.second;
        return true;
    }
    // TileAtlasState::request_tile — tile_atlas.rs:418-457
    bool request_tile(const TileKey& k) {
        if (!existing_tiles.count(k)) return true;
        auto it = tile_states.find(k);
        if (it != tile_states.end()) {
            if (it->second.requests == 0) {
                const uint32_t idx = it->second.atlas_index;
                for (auto u = unused_tiles.begin(); u != unused_tiles.end();)
                    u = (u->second == idx) ? unused_tiles.erase(u) : u + 1;
            }
            it->second.requests += 1;
        } else {
            uint32_t atlas_index;
            if (!allocate_tile(&atlas_index)) return false;
            tile_states[k] = AtlasTileState{1, false, atlas_index};
            bt_tile_load l{};
            l.tile = bt_tile{k.side, k.lod, k.x, k.y};
            l.atlas_index = atlas_index;
            pending_loads.push_back(l);
        }
        return true;
    }
    // TileAtlasState::release_tile — tile_atlas.rs:459-475
    void release_tile(const TileKey& k) {
        if (!existing_tiles.count(k)) return;
        auto it = tile_states.find(k);
        if (it == tile_states.end()) return;  // reference: expect("Tried releasing a tile, which is not present.")
        if (it->second.requests > 0) it->second.requests -= 1;
        if (it->second.requests == 0) unused_tiles.emplace_back(k, it->second.atlas_index);
    }
    // TileAtlasState::get_best_tile — tile_atlas.rs:477-503
    bt_tile_tree_entry best_tile(TileKey k) const {
        for (;;) {
            if (k == kInvalidTile || k.lod == 0xFFFFFFFFu) return bt_tile_tree_entry{0xFFFFFFFFu, 0xFFFFFFFFu};
            auto it = tile_states.find(k);
            if (it != tile_states.end() && it->second.loaded) return bt_tile_tree_entry{it->second.atlas_index, k.lod};
            k = TileKey{k.side, k.lod - 1u, k.x >> 1, k.y >> 1};  // TileCoordinate::parent — coordinate.rs:188-195 (wrapping_sub)
        }
    }
};

struct bt_tile_tree {
    bt_terrain_model model;
    bt_terrain_view_config cfg;
    uint32_t lod_count, atlas_size, side_count;
    double load_distance;  // pre-multiplied by the model scale (tile_tree.rs:148)
    float approximate_height;
    D3 view_world_position{0, 0, 0};
    std::vector<TreeSlot> tiles;            // [side][lod][x % tree][y % tree]
    std::vector<uint32_t> origins;          // [side][lod] uvec2
    bt_atlas_state* state = nullptr;        // the terrain's TileAtlasState (shared by all views of the terrain)

    size_t slot_index(uint32_t side, uint32_t lod, uint32_t x, uint32_t y) const {
        const uint32_t ts = cfg.tree_size;
        return (((size_t)side * lod_count + lod) * ts + x % ts) * ts + y % ts;
    }
    bt_tile_tree_entry best_tile(TileKey k) const { return state->best_tile(k); }
};

namespace {

// TileTree::compute_tree_xy — tile_tree.rs:175-178
D2 compute_tree_xy(D2 uv, double tile_count) {
    const double m = tile_count - 0.000001;
    return {std::fmin(uv.x * tile_count, m), std::fmin(uv.y * tile_count, m)};
}

// TileTree::compute_origin — tile_tree.rs:180-191
void compute_origin(const bt_tile_tree& t, D2 uv, uint32_t lod, uint32_t out[2]) {
    const double tile_count = (double)(1u << lod);
    const D2 xy = compute_tree_xy(uv, tile_count);
    const double hi = tile_count - (double)t.cfg.tree_size;
    auto f = [&](double v) {
        v = std::round(v - 0.5 * (double)t.cfg.tree_size);
        v = std::fmax(v, 0.0);
        v = std::fmin(v, hi);
        return to_u32_saturating(v);
    };
    out[0] = f(xy.x);
    out[1] = f(xy.y);
}

// TileTree::compute_tile_distance — tile_tree.rs:193-221
double compute_tile_distance(const bt_tile_tree& t, const TileKey& tile, Coordinate view) {
    const double tile_count = (double)(1u << tile.lod);
    const D2 vxy = compute_tree_xy(view.uv, tile_count);
    const int32_t ox = to_i32_saturating(vxy.x) - (int32_t)tile.x, oy = to_i32_saturating(vxy.y) - (int32_t)tile.y;
    D2 o{std::fmod(vxy.x, 1.0), std::fmod(vxy.y, 1.0)};
    if (ox < 0) o.x = 0.0; else if (ox > 0) o.x = 1.0;
    if (oy < 0) o.y = 0.0; else if (oy > 0) o.y = 1.0;
    const Coordinate c{tile.side, {((double)(int32_t)tile.x + o.x) / tile_count, ((double)(int32_t)tile.y + o.y) / tile_count}};
    return length(coordinate_world_position(t.model, c, t.approximate_height) - t.view_world_position);
}

}  // namespace

extern "C" {

bt_status bt_atlas_state_create(uint32_t atlas_size, bt_atlas_state** out) {
    if (!out || atlas_size == 0) return BT_ERR_INVALID_ARG;
    *out = new bt_atlas_state(atlas_size);
    return BT_OK;
}

void bt_atlas_state_destroy(bt_atlas_state* state) {
    if (!state) return;
    if (state->users == 0) delete state;
    else state->owned_by_tree = true;   // still referenced: the last tree that goes away frees it
}

bt_status bt_tile_tree_create(const bt_terrain_model* model, const bt_terrain_view_config* cfg, uint32_t lod_count,
                              uint32_t atlas_size, bt_tile_tree** out) {
    if (!out || atlas_size == 0) return BT_ERR_INVALID_ARG;
    bt_atlas_state* state = new bt_atlas_state(atlas_size);
    state->owned_by_tree = true;
    const bt_status st = bt_tile_tree_create_shared(model, cfg, lod_count, state, out);
    if (st != BT_OK) delete state;
    return st;
}

bt_status bt_tile_tree_create_shared(const bt_terrain_model* model, const bt_terrain_view_config* cfg, uint32_t lod_count,
                                     bt_atlas_state* state, bt_tile_tree** out) {
    if (!model || !cfg || !out || !state || lod_count == 0 || lod_count > 31 || cfg->tree_size == 0) return BT_ERR_INVALID_ARG;
    const uint32_t atlas_size = state->atlas_size;
    bt_tile_tree* t = new bt_tile_tree();
    t->state = state;
    state->users += 1;
    t->model = *model;
    t->cfg = *cfg;
    t->lod_count = lod_count;
    t->atlas_size = atlas_size;
    t->side_count = spherical(*model) ? 6u : 1u;  // TerrainModel::side_count — terrain_model.rs:174-180
    t->load_distance = cfg->load_distance * bt_terrain_model_scale(model);
    t->approximate_height = (model->min_height + model->max_height) / 2.0f;  // tile_tree.rs:155
    t->tiles.resize((size_t)t->side_count * lod_count * cfg->tree_size * cfg->tree_size);
    t->origins.assign((size_t)t->side_count * lod_count * 2, 0u);
    *out = t;
    return BT_OK;
}

void bt_tile_tree_destroy(bt_tile_tree* tree) {
    if (!tree) return;
    bt_atlas_state* st = tree->state;
    // the view's outstanding requests go back to the shared state (a despawned view releases its tiles)
    for (TreeSlot& slot : tree->tiles) if (slot.requested) st->release_tile(slot.coordinate);
    if (st->users) st->users -= 1;
    if (st->owned_by_tree && st->users == 0) delete st;
    delete tree;
}

bt_status bt_tile_tree_add_existing(bt_tile_tree* tree, const bt_tile* tiles, uint32_t count) {
    if (!tree || (!tiles && count)) return BT_ERR_INVALID_ARG;
    for (uint32_t i = 0; i < count; ++i) tree->state->existing_tiles.insert(TileKey{tiles[i].side, tiles[i].lod, tiles[i].x, tiles[i].y});
    return BT_OK;
}

void bt_tile_tree_set_approximate_height(bt_tile_tree* tree, float h) { if (tree) tree->approximate_height = h; }

// sample_attachment's host half — terrain_data/mod.rs:267-295 with TileTree::compute_blend / lookup_tile (tile_tree.rs:223-266)
bt_status bt_tile_tree_sample_lookup(bt_tile_tree* tree, const double sample_world_position[3], bt_tile_lookup lookups[2],
                                     uint32_t* lookup_count, float* blend_ratio) {
    if (!tree || !sample_world_position || !lookups || !lookup_count || !blend_ratio) return BT_ERR_INVALID_ARG;
    const bt_tile_tree& t = *tree;
    // TerrainModel::surface_position — terrain_model.rs:170-172
    const D3 local = position_world_to_local(t.model, D3{sample_world_position[0], sample_world_position[1], sample_world_position[2]});
    const D3 surface = position_local_to_world(t.model, local, (double)t.approximate_height);
    // TileTree::compute_blend — tile_tree.rs:223-237
    const double blend_distance = t.cfg.blend_distance * bt_terrain_model_scale(&t.model);
    const double view_distance = length(t.view_world_position - surface);
    const float target_lod = (float)std::fmin(std::log2(blend_distance / view_distance), (double)t.lod_count - 0.00001);
    const uint32_t lod = target_lod > 0.0f ? (target_lod >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)target_lod) : 0u;   // `as u32` saturates
    float ratio = 0.0f;
    if (lod != 0) {  // util::inverse_mix — util.rs:8-10
        const float a = (float)lod + t.cfg.blend_range, b = (float)lod;
        const float v = (target_lod - a) / (b - a);
        ratio = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    }
    auto lookup = [&](uint32_t tree_lod) {  // TileTree::lookup_tile — tile_tree.rs:239-266
        const Coordinate c = coordinate_from_world(t.model, surface);
        const double tile_count = (double)(1u << tree_lod);
        const double m = tile_count - 0.000001;
        const D2 xy{std::fmin(c.uv.x * tile_count, m), std::fmin(c.uv.y * tile_count, m)};
        const TreeSlot& slot = t.tiles[t.slot_index(c.side, tree_lod, (uint32_t)(size_t)xy.x, (uint32_t)(size_t)xy.y)];
        const bt_tile_tree_entry e = t.best_tile(slot.coordinate);
        bt_tile_lookup out{0xFFFFFFFFu, 0xFFFFFFFFu, {0.0f, 0.0f}};   // TileLookup::INVALID
        if (e.atlas_lod == 0xFFFFFFFFu) return out;
        const double div = (double)(1u << (tree_lod - e.atlas_lod));
        out.atlas_index = e.atlas_index;
        out.atlas_lod = e.atlas_lod;
        out.atlas_uv[0] = (float)std::fmod(xy.x / div, 1.0);
        out.atlas_uv[1] = (float)std::fmod(xy.y / div, 1.0);
        return out;
    };
    lookups[0] = lookup(lod);
    *lookup_count = 1;
    if (ratio > 0.0f) { lookups[1] = lookup(lod - 1); *lookup_count = 2; }
    *blend_ratio = ratio;
    return BT_OK;
}

// TileTree::update (tile_tree.rs:268-333) followed by TileAtlas::update's request/release pass
// (tile_atlas.rs:574-601: released tiles first, then requested ones)
bt_status bt_tile_tree_update(bt_tile_tree* tree, const double view_world_position[3], bt_tile_load* loads, uint32_t cap, uint32_t* count) {
    if (!tree || !view_world_position) return BT_ERR_INVALID_ARG;
    bt_tile_tree& t = *tree;
    t.view_world_position = D3{view_world_position[0], view_world_position[1], view_world_position[2]};
    const Coordinate view = coordinate_from_world(t.model, t.view_world_position);
    std::vector<TileKey> released, requested;
    const uint32_t ts = t.cfg.tree_size;
    for (uint32_t side = 0; side < t.side_count; ++side) {
        const Coordinate vc = project_to_side(t.model, view, side);
        for (uint32_t lod = 0; lod < t.lod_count; ++lod) {
            uint32_t origin[2];
            compute_origin(t, vc.uv, lod, origin);
            t.origins[((size_t)side * t.lod_count + lod) * 2 + 0] = origin[0];
            t.origins[((size_t)side * t.lod_count + lod) * 2 + 1] = origin[1];
            for (uint32_t x = 0; x < ts; ++x) for (uint32_t y = 0; y < ts; ++y) {
                const TileKey tc{side, lod, origin[0] + x, origin[1] + y};
                const double tile_distance = compute_tile_distance(t, tc, vc);
                const double load_distance = t.load_distance / (double)(1u << lod);
                const bool want = lod == 0 || tile_distance < load_distance;
                TreeSlot& slot = t.tiles[t.slot_index(side, lod, tc.x, tc.y)];
                if (!(tc == slot.coordinate)) {
                    if (slot.requested) { slot.requested = false; released.push_back(slot.coordinate); }
                    slot.coordinate = tc;
                }
                if (!slot.requested && want) { slot.requested = true; requested.push_back(tc); }
                else if (slot.requested && !want) { slot.requested = false; released.push_back(tc); }
            }
        }
    }
    bt_atlas_state& as = *t.state;
    for (const TileKey& k : released) as.release_tile(k);
    bt_status st = BT_OK;
    for (const TileKey& k : requested) if (!as.request_tile(k)) st = BT_ERR_CAPACITY;  // atlas out of indices
    const uint32_t n = (uint32_t)as.pending_loads.size();   // the terrain's load queue: requests of every view that shares the state
    if (count) *count = n;
    if (loads) for (uint32_t i = 0; i < n && i < cap; ++i) loads[i] = as.pending_loads[i];
    if (loads && cap >= n) as.pending_loads.clear();
    return st;
}

// TileAtlasState::loaded_tile_attachment once per attachment collapses to one call per tile
bt_status bt_tile_tree_mark_loaded(bt_tile_tree* tree, const bt_tile* tile) {
    if (!tree || !tile) return BT_ERR_INVALID_ARG;
    auto it = tree->state->tile_states.find(TileKey{tile->side, tile->lod, tile->x, tile->y});
    if (it == tree->state->tile_states.end()) return BT_ERR_NOT_FOUND;
    it->second.loaded = true;
    tree->state->loaded_version += 1;
    return BT_OK;
}

// TileTree::adjust_to_tile_atlas — tile_tree.rs:363-374
bt_status bt_tile_tree_entries(bt_tile_tree* tree, bt_tile_tree_entry* entries, uint32_t entry_cap, uint32_t* origins_xy, uint32_t origin_cap) {
    if (!tree || !entries || entry_cap < tree->tiles.size()) return BT_ERR_INVALID_ARG;
    for (size_t i = 0; i < tree->tiles.size(); ++i) entries[i] = tree->best_tile(tree->tiles[i].coordinate);
    if (origins_xy) {
        if (origin_cap < tree->origins.size()) return BT_ERR_INVALID_ARG;
        std::memcpy(origins_xy, tree->origins.data(), tree->origins.size() * sizeof(uint32_t));
    }
    return BT_OK;
}


// ================================================================================================
// Device clipmap: the per-frame half of TileTree on the GPU (clipmap.cu).  The slots, the origins and the
// best-tile table stay in device memory; the host keeps what only it can do: Coordinate::from_world_position of the view (one
// f64 evaluation, with the ellipsoid projection), the TileAtlasState bookkeeping and the loads themselves.
// ================================================================================================
struct bt_clipmap {
    bt_view* view = nullptr;
    bt_tile_tree* tree = nullptr;
    bt_context* ctx = nullptr;
    uint32_t entries = 0, hash_capacity = 0;
    btd::ClipmapParams* d_params = nullptr;
    btd::ClipmapParams* h_params = nullptr;      // pinned
    btd::ClipmapSlot* d_slots = nullptr;
    uint32_t* d_origins = nullptr;
    unsigned char* d_lists = nullptr;            // [ClipmapLists | released[entries] | requested[entries]]
    unsigned char* h_lists = nullptr;            // pinned
    size_t lists_bytes = 0;
    btd::ClipmapLoaded* d_loaded = nullptr;
    btd::ClipmapLoaded* h_loaded = nullptr;      // pinned
    uint64_t uploaded_version = ~0ull;
    size_t h2d_bytes_last_update = 0, d2h_bytes_last_update = 0;
};

bt_status bt_clipmap_create(bt_view* view, bt_tile_tree* tree, bt_clipmap** out) {
    if (!view || !tree || !out) return BT_ERR_INVALID_ARG;
    bt_context* ctx = view->terrain->ctx;
    const uint32_t entries = tree->side_count * tree->lod_count * tree->cfg.tree_size * tree->cfg.tree_size;
    if (entries != view->tile_tree_entries || tree->cfg.tree_size != view->desc.tree_size) {
        ctx->set_error("bt_clipmap_create: the tree has %u entries (tree_size %u), the view %u (tree_size %u)", entries, tree->cfg.tree_size,
                       view->tile_tree_entries, view->desc.tree_size);
        return BT_ERR_INVALID_ARG;
    }
    for (const TreeSlot& slot : tree->tiles) if (slot.requested) {
        ctx->set_error("bt_clipmap_create: the tree has already been updated on the host; create the clipmap first");
        return BT_ERR_INVALID_ARG;
    }
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    bt_clipmap* c = new bt_clipmap();
    c->view = view; c->tree = tree; c->ctx = ctx; c->entries = entries;
    c->hash_capacity = 64;
    while (c->hash_capacity < 4u * tree->atlas_size) c->hash_capacity *= 2u;
    c->lists_bytes = sizeof(btd::ClipmapLists) + 2 * (size_t)entries * sizeof(bt_tile);
    auto fail = [&](cudaError_t e) { ctx->set_error("bt_clipmap_create: %s", cudaGetErrorString(e)); bt_clipmap_destroy(c); return e == cudaErrorMemoryAllocation ? BT_ERR_OUT_OF_MEMORY : BT_ERR_CUDA; };
    cudaError_t e;
    if ((e = cudaMalloc((void**)&c->d_params, sizeof(btd::ClipmapParams))) != cudaSuccess) return fail(e);
    if ((e = cudaMallocHost((void**)&c->h_params, sizeof(btd::ClipmapParams))) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc((void**)&c->d_slots, (size_t)entries * sizeof(btd::ClipmapSlot))) != cudaSuccess) return fail(e);
    if ((e = cudaMemset(c->d_slots, 0, (size_t)entries * sizeof(btd::ClipmapSlot))) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc((void**)&c->d_origins, (size_t)tree->side_count * tree->lod_count * 2 * sizeof(uint32_t))) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc((void**)&c->d_lists, c->lists_bytes)) != cudaSuccess) return fail(e);
    if ((e = cudaMallocHost((void**)&c->h_lists, c->lists_bytes)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc((void**)&c->d_loaded, (size_t)c->hash_capacity * sizeof(btd::ClipmapLoaded))) != cudaSuccess) return fail(e);
    if ((e = cudaMallocHost((void**)&c->h_loaded, (size_t)c->hash_capacity * sizeof(btd::ClipmapLoaded))) != cudaSuccess) return fail(e);
    view->device_table = true;
    *out = c;
    return BT_OK;
}

void bt_clipmap_destroy(bt_clipmap* c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    // the view's outstanding requests go back to the shared state, as bt_tile_tree_destroy does for host-side slots
    if (c->d_slots && c->tree) {
        std::vector<btd::ClipmapSlot> slots(c->entries);
        if (cudaMemcpy(slots.data(), c->d_slots, slots.size() * sizeof(btd::ClipmapSlot), cudaMemcpyDeviceToHost) == cudaSuccess) {
            const uint32_t per = c->tree->cfg.tree_size * c->tree->cfg.tree_size;
            for (uint32_t i = 0; i < c->entries; ++i) if (slots[i].valid && slots[i].requested) {
                const uint32_t sl = i / per, side = sl / c->tree->lod_count, lod = sl - side * c->tree->lod_count;
                c->tree->state->release_tile(TileKey{side, lod, slots[i].x, slots[i].y});
            }
        }
    }
    if (c->view) c->view->device_table = false;
    cudaFree(c->d_params); cudaFreeHost(c->h_params); cudaFree(c->d_slots); cudaFree(c->d_origins); cudaFree(c->d_lists);
    cudaFreeHost(c->h_lists); cudaFree(c->d_loaded); cudaFreeHost(c->h_loaded);
    delete c;
}

// Twin of bt_tile_tree_update: one parameter block up, one kernel, the request / release lists down; the TileAtlasState
// bookkeeping (same order: released first, then requested — tile_atlas.rs:574-601) runs on the host and reports the loads.
bt_status bt_clipmap_update(bt_clipmap* c, const double view_world_position[3], bt_tile_load* loads, uint32_t cap, uint32_t* count, void* stream) {
    if (!c || !view_world_position) return BT_ERR_INVALID_ARG;
    bt_context* ctx = c->ctx;
    bt_tile_tree& t = *c->tree;
    cudaStream_t s = (cudaStream_t)stream;
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    t.view_world_position = D3{view_world_position[0], view_world_position[1], view_world_position[2]};
    const Coordinate view = coordinate_from_world(t.model, t.view_world_position);
    btd::ClipmapParams& P = *c->h_params;
    for (int i = 0; i < 3; ++i) { P.position[i] = t.model.position[i]; P.scale[i] = t.model.scale[i]; P.view_world_position[i] = view_world_position[i]; }
    P.view_uv[0] = view.uv.x; P.view_uv[1] = view.uv.y;
    P.load_distance = t.load_distance;
    P.approximate_height = t.approximate_height;
    P.view_side = view.side; P.kind = t.model.kind; P.tree_size = t.cfg.tree_size; P.lod_count = t.lod_count; P.side_count = t.side_count; P._pad = 0;
    BT_CUDA(ctx, cudaMemcpyAsync(c->d_params, c->h_params, sizeof(P), cudaMemcpyHostToDevice, s));
    bt_tile* d_released = reinterpret_cast<bt_tile*>(c->d_lists + sizeof(btd::ClipmapLists));
    BT_CUDA(ctx, bt_launch_clipmap_update(c->d_params, c->d_slots, c->d_origins, reinterpret_cast<btd::ClipmapLists*>(c->d_lists), d_released,
                                          d_released + c->entries, s));
    BT_CUDA(ctx, cudaMemcpyAsync(c->h_lists, c->d_lists, c->lists_bytes, cudaMemcpyDeviceToHost, s));
    BT_CUDA(ctx, cudaStreamSynchronize(s));
    c->h2d_bytes_last_update = sizeof(P);
    c->d2h_bytes_last_update = c->lists_bytes;
    const btd::ClipmapLists& H = *reinterpret_cast<const btd::ClipmapLists*>(c->h_lists);
    const bt_tile* released = reinterpret_cast<const bt_tile*>(c->h_lists + sizeof(btd::ClipmapLists));
    const bt_tile* requested = released + c->entries;
    bt_atlas_state& as = *t.state;
    for (uint32_t i = 0; i < H.n_released; ++i) as.release_tile(TileKey{released[i].side, released[i].lod, released[i].x, released[i].y});
    bt_status st = BT_OK;
    for (uint32_t i = 0; i < H.n_requested; ++i)
        if (!as.request_tile(TileKey{requested[i].side, requested[i].lod, requested[i].x, requested[i].y})) st = BT_ERR_CAPACITY;
    const uint32_t n = (uint32_t)as.pending_loads.size();
    if (count) *count = n;
    if (loads) for (uint32_t i = 0; i < n && i < cap; ++i) loads[i] = as.pending_loads[i];
    if (loads && cap >= n) as.pending_loads.clear();
    if (st != BT_OK) ctx->set_error("bt_clipmap_update: atlas out of indices");
    return st;
}

// Twin of bt_tile_tree_entries (TileTree::adjust_to_tile_atlas): writes the view's table in place.  The set of loaded tiles
// goes up only when it changed since the last call.
bt_status bt_clipmap_entries(bt_clipmap* c, void* stream) {
    if (!c) return BT_ERR_INVALID_ARG;
    bt_context* ctx = c->ctx;
    cudaStream_t s = (cudaStream_t)stream;
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    bt_atlas_state& as = *c->tree->state;
    c->h2d_bytes_last_update = 0;
    if (as.loaded_version != c->uploaded_version) {
        BT_CUDA(ctx, cudaStreamSynchronize(s));   // the previous upload out of h_loaded has finished
        const uint32_t mask = c->hash_capacity - 1u;
        std::memset(c->h_loaded, 0xFF, (size_t)c->hash_capacity * sizeof(btd::ClipmapLoaded));
        for (const auto& kv : as.tile_states) {
            if (!kv.second.loaded) continue;
            const uint32_t key = (kv.first.side << 8) | kv.first.lod;
            uint32_t i = bt_clipmap_hash(key, kv.first.x, kv.first.y) & mask;
            while (c->h_loaded[i].side_lod != 0xFFFFFFFFu) i = (i + 1u) & mask;
            c->h_loaded[i] = btd::ClipmapLoaded{key, kv.first.x, kv.first.y, kv.second.atlas_index};
        }
        BT_CUDA(ctx, cudaMemcpyAsync(c->d_loaded, c->h_loaded, (size_t)c->hash_capacity * sizeof(btd::ClipmapLoaded), cudaMemcpyHostToDevice, s));
        c->uploaded_version = as.loaded_version;
        c->h2d_bytes_last_update = (size_t)c->hash_capacity * sizeof(btd::ClipmapLoaded);
    }
    BT_CUDA(ctx, bt_launch_clipmap_entries(c->d_slots, c->entries, c->tree->cfg.tree_size, c->tree->lod_count, c->d_loaded, c->hash_capacity,
                                           c->view->d_tile_tree, s));
    return BT_OK;
}

// test / debugging read-back of what the kernels maintain (table + origins)
bt_status bt_clipmap_download(bt_clipmap* c, bt_tile_tree_entry* entries, uint32_t entry_cap, uint32_t* origins_xy, uint32_t origin_cap, void* stream) {
    if (!c || !entries || entry_cap < c->entries) return BT_ERR_INVALID_ARG;
    bt_context* ctx = c->ctx;
    cudaStream_t s = (cudaStream_t)stream;
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    BT_CUDA(ctx, cudaMemcpyAsync(entries, c->view->d_tile_tree, (size_t)c->entries * sizeof(bt_tile_tree_entry), cudaMemcpyDeviceToHost, s));
    if (origins_xy) {
        const uint32_t n = c->tree->side_count * c->tree->lod_count * 2u;
        if (origin_cap < n) return BT_ERR_INVALID_ARG;
        BT_CUDA(ctx, cudaMemcpyAsync(origins_xy, c->d_origins, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    }
    BT_CUDA(ctx, cudaStreamSynchronize(s));
    return BT_OK;
}

void bt_clipmap_transfer_bytes(const bt_clipmap* c, uint64_t* h2d, uint64_t* d2h) {
    if (h2d) *h2d = c ? c->h2d_bytes_last_update : 0;
    if (d2h) *d2h = c ? c->d2h_bytes_last_update : 0;
}

}  // extern "C"
