/*
 * bt_oracle_host.c — CPU oracle: f64 host math that produces the hot path's per-frame inputs.
 * TEST INFRASTRUCTURE ONLY (see bt_oracle.h).  PARITY UNPINNED (no reference tests exist).
 *
 * Restates src/math/coordinate.rs:69-151, src/math/terrain_model.rs:14-21,131-193,263-360 and
 * src/terrain_data/tile_tree.rs:175-221 and src/math/ellipsoid.rs (closest point on an ellipsoid).  Planar, sphere and
 * ellipsoid models with the identity rotation, as every TerrainModel constructor produces (terrain_model.rs:84-128).
 */
#include "bt_oracle.h"

#include <math.h>
#include <string.h>

static const double C_SQR = 0.87 * 0.87; /* src/math/mod.rs:13 */

static int is_spherical(const bt_terrain_model* m) { return m->kind != BT_KIND_PLANAR; }

/* DMat4::from_scale_rotation_translation(scale, IDENTITY, translation).transform_point3 */
static void world_from_local_point(const bt_terrain_model* m, const double p[3], double out[3]) {
    for (int k = 0; k < 3; ++k) out[k] = m->scale[k] * p[k] + m->position[k];
}
static void world_from_local_vector(const bt_terrain_model* m, const double p[3], double out[3]) {
    for (int k = 0; k < 3; ++k) out[k] = m->scale[k] * p[k];
}
/* local_from_world = world_from_local.inverse() */
static void local_from_world_point(const bt_terrain_model* m, const double p[3], double out[3]) {
    for (int k = 0; k < 3; ++k) out[k] = (p[k] - m->position[k]) / m->scale[k];
}
static void normalize3(double v[3]) {
    double l = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    v[0] /= l; v[1] /= l; v[2] /= l;
}

/* ---- src/math/ellipsoid.rs: closest point on an axis-aligned ellipse / ellipsoid (first quadrant / octant, semi-axes in
 * descending order), Eberly's bisection on  F(t) = sum (r_i z_i / (t + r_i))^2 - 1  ------------------------------------ */
#define BTO_ELLIPSOID_MAX_ITERATIONS 1074 /* ellipsoid.rs:10 */

/* get_root_2d / get_root_3d — ellipsoid.rs:92-142: dims = 2 or 3 */
static double ellipsoid_get_root(int dims, const double r[3], const double z[3], double g) {
    double n[3] = {0.0, 0.0, 0.0}, len2 = 0.0, s0, s1, s = 0.0;
    for (int i = 0; i < dims; ++i) { n[i] = r[i] * z[i]; len2 += n[i] * n[i]; }
    s0 = z[dims - 1] - 1.0;
    s1 = g < 0.0 ? 0.0 : sqrt(len2) - 1.0;
    for (int iteration = 0; iteration < BTO_ELLIPSOID_MAX_ITERATIONS; ++iteration) {
        double sum = 0.0, gi;
        s = (s0 + s1) / 2.0;
        if (s == s0 || s == s1) break;
        for (int i = 0; i < dims; ++i) { double ratio = n[i] / (s + r[i]); sum += ratio * ratio; }
        gi = sum - 1.0;
        if (gi < 0.0) s1 = s;
        else if (gi > 0.0) s0 = s;
        else break;
    }
    return s;
}

/* project_point_ellipse — ellipsoid.rs:65-90 */
static void ellipse_project(double ea, double eb, double ya, double yb, double* xa, double* xb) {
    if (yb > 0.0) {
        if (ya > 0.0) {
            double z[3] = {ya / ea, yb / eb, 0.0};
            double g = (z[0] * z[0] + z[1] * z[1]) - 1.0;
            if (g != 0.0) {
                double r[3] = {(ea * ea) / (eb * eb), 1.0, 0.0};
                double t = ellipsoid_get_root(2, r, z, g);
                *xa = r[0] * ya / (t + r[0]);
                *xb = r[1] * yb / (t + r[1]);
            } else { *xa = ya; *xb = yb; }
        } else { *xa = 0.0; *xb = eb; }
    } else {
        double numer0 = ea * ya, denom0 = ea * ea - eb * eb;
        if (numer0 < denom0) {
            double xde0 = numer0 / denom0;
            *xa = ea * xde0;
            *xb = eb * sqrt(1.0 - xde0 * xde0);
        } else { *xa = ea; *xb = 0.0; }
    }
}

static double f64_signum(double v) { return v != v ? v : (signbit(v) ? -1.0 : 1.0); }

/* project_point_ellipsoid — ellipsoid.rs:12-63; e = (major, major, minor); the query's y and z are exchanged on the way in and
 * on the way out (`y.xzy()`, `x.xzy()`) because the terrain's minor axis is world y */
static void ellipsoid_project(const double e[3], const double query[3], double out[3]) {
    const double sx = f64_signum(query[0]), sy = f64_signum(query[1]), sz = f64_signum(query[2]);
    const double y[3] = {fabs(query[0]), fabs(query[2]), fabs(query[1])};
    double x[3];
    if (y[2] > 0.0) {
        if (y[1] > 0.0) {
            if (y[0] > 0.0) {
                double z[3] = {y[0] / e[0], y[1] / e[1], y[2] / e[2]};
                double g = ((z[0] * z[0] + z[1] * z[1]) + z[2] * z[2]) - 1.0;
                if (g != 0.0) {
                    double r[3] = {(e[0] * e[0]) / (e[2] * e[2]), (e[1] * e[1]) / (e[2] * e[2]), 1.0};
                    double t = ellipsoid_get_root(3, r, z, g);
                    for (int i = 0; i < 3; ++i) x[i] = r[i] * y[i] / (t + r[i]);
                } else { x[0] = y[0]; x[1] = y[1]; x[2] = y[2]; }
            } else { x[0] = 0.0; ellipse_project(e[1], e[2], y[1], y[2], &x[1], &x[2]); }
        } else if (y[0] > 0.0) { x[1] = 0.0; ellipse_project(e[0], e[2], y[0], y[2], &x[0], &x[2]); }
        else { x[0] = 0.0; x[1] = 0.0; x[2] = e[2]; }
    } else {
        double denom0 = e[0] * e[0] - e[2] * e[2], denom1 = e[1] * e[1] - e[2] * e[2];
        double numer0 = e[0] * y[0], numer1 = e[1] * y[1];
        int have = 0;
        if (numer0 < denom0 && numer1 < denom1) {
            double xde0 = numer0 / denom0, xde1 = numer1 / denom1;
            double discr = (1.0 - xde0 * xde0) - xde1 * xde1;
            if (discr > 0.0) { x[0] = e[0] * xde0; x[1] = e[1] * xde1; x[2] = e[2] * sqrt(discr); have = 1; }
        }
        if (!have) { x[2] = 0.0; ellipse_project(e[0], e[1], y[0], y[1], &x[0], &x[1]); }
    }
    out[0] = sx * x[0];
    out[1] = sy * x[2];
    out[2] = sz * x[1];
}

void bto_project_point_ellipsoid(double major_axis, double minor_axis, const double query[3], double out[3]) {
    const double e[3] = {major_axis, major_axis, minor_axis};
    ellipsoid_project(e, query, out);
}

/* TerrainModel::position_world_to_local — terrain_model.rs:145-168 */
static void position_world_to_local(const bt_terrain_model* m, const double world[3], double out[3]) {
    if (m->kind == BT_KIND_ELLIPSOIDAL) {
        /* ellipsoid_from_world = (rotation * translation)^-1 with the identity rotation; NOTE the reference then applies
         * local_from_world to the point of the ELLIPSOID frame (terrain_model.rs:163-166): reproduced literally */
        const double e[3] = {m->scale[0], m->scale[0], m->scale[1]};
        double rel[3] = {world[0] - m->position[0], world[1] - m->position[1], world[2] - m->position[2]}, surface[3];
        ellipsoid_project(e, rel, surface);
        local_from_world_point(m, surface, out);
        normalize3(out);
        return;
    }
    local_from_world_point(m, world, out);
    if (m->kind == BT_KIND_PLANAR) { out[1] = 0.0 * out[1]; }
    else normalize3(out);
}

/* TerrainModel::position_local_to_world — terrain_model.rs:131-143 */
static void position_local_to_world(const bt_terrain_model* m, const double local[3], double height, double out[3]) {
    double wp[3], wn[3];
    world_from_local_point(m, local, wp);
    if (is_spherical(m)) world_from_local_vector(m, local, wn);
    else { const double up[3] = {0.0, 1.0, 0.0}; world_from_local_vector(m, up, wn); }
    normalize3(wn);
    for (int k = 0; k < 3; ++k) out[k] = wp[k] + height * wn[k];
}

/* Coordinate::from_world_position — coordinate.rs:69-108 */
void bto_coordinate_from_world_position(const bt_terrain_model* model, const double world[3], uint32_t* side, double uv[2]) {
    double lp[3];
    position_world_to_local(model, world, lp);
    if (is_spherical(model)) {
        double ax = fabs(lp[0]), ay = fabs(lp[1]), az = fabs(lp[2]);
        double q[2];
        if (ax > ay && ax > az) {
            if (lp[0] < 0.0) { *side = 0; q[0] = -lp[2] / lp[0]; q[1] = lp[1] / lp[0]; }
            else { *side = 3; q[0] = -lp[1] / lp[0]; q[1] = lp[2] / lp[0]; }
        } else if (az > ay) {
            if (lp[2] > 0.0) { *side = 1; q[0] = lp[0] / lp[2]; q[1] = -lp[1] / lp[2]; }
            else { *side = 4; q[0] = lp[1] / lp[2]; q[1] = -lp[0] / lp[2]; }
        } else {
            if (lp[1] > 0.0) { *side = 2; q[0] = lp[0] / lp[1]; q[1] = lp[2] / lp[1]; }
            else { *side = 5; q[0] = -lp[2] / lp[1]; q[1] = -lp[0] / lp[1]; }
        }
        for (int k = 0; k < 2; ++k) {
            double w = q[k] * pow((1.0 + C_SQR) / (1.0 + C_SQR * q[k] * q[k]), 0.5);
            uv[k] = 0.5 * w + 0.5;
        }
    } else {
        *side = 0;
        double u = lp[0] + 0.5, v = lp[2] + 0.5;
        uv[0] = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
        uv[1] = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
    }
}

/* Coordinate::world_position — coordinate.rs:110-130 */
void bto_coordinate_world_position(const bt_terrain_model* model, uint32_t side, const double uv[2], float height, double out[3]) {
    double lp[3];
    if (is_spherical(model)) {
        double q[2];
        for (int k = 0; k < 2; ++k) {
            double w = (uv[k] - 0.5) / 0.5;
            q[k] = w / pow(1.0 + C_SQR - C_SQR * w * w, 0.5);
        }
        switch (side) {
            case 0: lp[0] = -1.0; lp[1] = -q[1]; lp[2] = q[0]; break;
            case 1: lp[0] = q[0]; lp[1] = -q[1]; lp[2] = 1.0; break;
            case 2: lp[0] = q[0]; lp[1] = 1.0; lp[2] = q[1]; break;
            case 3: lp[0] = 1.0; lp[1] = -q[0]; lp[2] = q[1]; break;
            case 4: lp[0] = q[1]; lp[1] = -q[0]; lp[2] = -1.0; break;
            default: lp[0] = q[1]; lp[1] = -1.0; lp[2] = q[0]; break;
        }
        normalize3(lp);
    } else {
        lp[0] = uv[0] - 0.5; lp[1] = 0.0; lp[2] = uv[1] - 0.5;
    }
    position_local_to_world(model, lp, (double)height, out);
}

/* SideInfo::project_to_side + Coordinate::project_to_side — coordinate.rs:19-53,134-151 */
static void project_to_side(const bt_terrain_model* model, uint32_t self_side, const double self_uv[2], uint32_t side, double out_uv[2]) {
    enum { F0, F1, PS, PT };
    static const unsigned char EVEN[6][2] = {{PS, PT}, {F0, PT}, {F0, PS}, {PT, PS}, {PT, F0}, {PS, F0}};
    static const unsigned char ODD[6][2] = {{PS, PT}, {PS, F1}, {PT, F1}, {PT, PS}, {F1, PS}, {F1, PT}};
    if (!is_spherical(model)) { out_uv[0] = self_uv[0]; out_uv[1] = self_uv[1]; return; }
    uint32_t index = (6 + side - self_side) % 6;
    const unsigned char* info = (self_side % 2 == 0) ? EVEN[index] : ODD[index];
    for (int k = 0; k < 2; ++k) {
        switch (info[k]) {
            case F0: out_uv[k] = 0.0; break;
            case F1: out_uv[k] = 1.0; break;
            case PS: out_uv[k] = self_uv[0]; break;
            default: out_uv[k] = self_uv[1]; break;
        }
    }
}

/* f64::powi as compiler-rt's __powidf2 evaluates it (binary exponentiation) */
static double powi(double a, int b) {
    double r = 1.0;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return r;
}

/* SIDE_MATRICES — terrain_model.rs:14-21 (column-major) applied to a vector */
static void side_matrix_mul(uint32_t side, const double v[3], double out[3]) {
    static const double SM[6][9] = {
        {-1, 0, 0, 0, 0, 1, 0, -1, 0}, {0, 0, 1, 1, 0, 0, 0, -1, 0}, {0, 1, 0, 1, 0, 0, 0, 0, 1},
        {1, 0, 0, 0, -1, 0, 0, 0, 1}, {0, 0, -1, 0, -1, 0, 1, 0, 0}, {0, -1, 0, 0, 0, 1, 1, 0, 0}};
    const double* m = SM[side];
    for (int r = 0; r < 3; ++r) out[r] = m[0 + r] * v[0] + m[3 + r] * v[1] + m[6 + r] * v[2];
}

static int32_t f64_to_i32_sat(double x) {
    if (x != x) return 0;
    if (x <= -2147483648.0) return (int32_t)0x80000000;
    if (x >= 2147483647.0) return 0x7fffffff;
    return (int32_t)x;
}

/* TerrainModelApproximation::compute — terrain_model.rs:263-360 */
void bto_model_approximation(const bt_terrain_model* model, const double view_world_position[3],
                             uint32_t origin_lod, float approximate_height, bt_model_approximation* out) {
    memset(out, 0, sizeof(*out));
    double origin_count = (double)(1u << origin_lod);
    uint32_t view_side; double view_uv[2];
    bto_coordinate_from_world_position(model, view_world_position, &view_side, view_uv);
    for (uint32_t side = 0; side < 6; ++side) {
        double uv[2];
        project_to_side(model, view_side, view_uv, side, uv);
        bt_side_parameter* sp = &out->sides[side];
        for (int k = 0; k < 2; ++k) {
            double scaled = uv[k] * origin_count;
            sp->view_xy[k] = f64_to_i32_sat(scaled);
            sp->view_uv[k] = (float)(scaled - floor(scaled));
        }
        double s = uv[0], t = uv[1];
        double u_denom = sqrt(1.0 - 4.0 * C_SQR * s * (s - 1.0));
        double u = (2.0 * s - 1.0) / u_denom;
        double u_ds = 2.0 * (C_SQR + 1.0) / powi(u_denom, 3);
        double u_dss = 12.0 * C_SQR * (C_SQR + 1.0) * (2.0 * s - 1.0) / powi(u_denom, 5);
        double v_denom = sqrt(1.0 - 4.0 * C_SQR * t * (t - 1.0));
        double v = (2.0 * t - 1.0) / v_denom;
        double v_dt = 2.0 * (C_SQR + 1.0) / powi(v_denom, 3);
        double v_dtt = 12.0 * C_SQR * (C_SQR + 1.0) * (2.0 * t - 1.0) / powi(v_denom, 5);

        double l = sqrt(1.0 + u * u + v * v);
        double l_ds = u * u_ds / l;
        double l_dt = v * v_dt / l;
        double l_dss = (u * u_dss * l * l + (v * v + 1.0) * u_ds * u_ds) / powi(l, 3);
        double l_dst = -(u * v * u_ds * v_dt) / powi(l, 3);
        double l_dtt = (v * v_dtt * l * l + (u * u + 1.0) * v_dt * v_dt) / powi(l, 3);

        double a = 1.0;
        double a_ds = -l_ds;
        double a_dt = -l_dt;
        double a_dss = 2.0 * l_ds * l_ds - l * l_dss;
        double a_dst = 2.0 * l_ds * l_dt - l * l_dst;
        double a_dtt = 2.0 * l_dt * l_dt - l * l_dtt;

        double b = u;
        double b_ds = -u * l_ds + l * u_ds;
        double b_dt = -u * l_dt;
        double b_dss = 2.0 * u * l_ds * l_ds - l * (2.0 * u_ds * l_ds + u * l_dss) + u_dss * l * l;
        double b_dst = 2.0 * u * l_ds * l_dt - l * (u_ds * l_dt + u * l_dst);
        double b_dtt = 2.0 * u * l_dt * l_dt - l * u * l_dtt;

        double c = v;
        double c_ds = -v * l_ds;
        double c_dt = -v * l_dt + l * v_dt;
        double c_dss = 2.0 * v * l_ds * l_ds - l * v * l_dss;
        double c_dst = 2.0 * v * l_ds * l_dt - l * (v_dt * l_ds + v * l_dst);
        double c_dtt = 2.0 * v * l_dt * l_dt - l * (2.0 * v_dt * l_dt + v * l_dtt) + v_dtt * l * l;

        double l2 = powi(l, 2), l3 = powi(l, 3);
        double in[6][3] = {{a, b, c}, {a_ds, b_ds, c_ds}, {a_dt, b_dt, c_dt}, {a_dss, b_dss, c_dss}, {a_dst, b_dst, c_dst}, {a_dtt, b_dtt, c_dtt}};
        double div[6] = {l, l2, l2, l3, l3, l3};
        double res[6][3];
        for (int i = 0; i < 6; ++i) {
            double sm[3];
            side_matrix_mul(side, in[i], sm);
            for (int k = 0; k < 3; ++k) sm[k] = sm[k] / div[i];
            if (i == 0) world_from_local_point(model, sm, res[i]);
            else world_from_local_vector(model, sm, res[i]);
        }
        for (int k = 0; k < 3; ++k) {
            sp->c[k] = (float)(res[0][k] - view_world_position[k]);
            sp->c_s[k] = (float)res[1][k];
            sp->c_t[k] = (float)res[2][k];
            sp->c_ss[k] = (float)(res[3][k] / 2.0);
            sp->c_st[k] = (float)res[4][k];
            sp->c_tt[k] = (float)(res[5][k] / 2.0);
        }
    }
    out->origin_lod = origin_lod;
    out->approximate_height = approximate_height;
}

/* TileTree::compute_tree_xy — tile_tree.rs:175-178 */
static void compute_tree_xy(const double uv[2], double tile_count, double out[2]) {
    for (int k = 0; k < 2; ++k) {
        double v = uv[k] * tile_count, m = tile_count - 0.000001;
        out[k] = v < m ? v : m;
    }
}

/* TileTree::compute_origin — tile_tree.rs:180-191 */
void bto_tile_tree_origin(uint32_t tree_size, uint32_t lod, const double uv[2], uint32_t out[2]) {
    double tile_count = (double)(1u << lod);
    double xy[2];
    compute_tree_xy(uv, tile_count, xy);
    for (int k = 0; k < 2; ++k) {
        double v = round(xy[k] - 0.5 * (double)tree_size); /* f64::round: half away from zero */
        double lo = 0.0, hi = tile_count - (double)tree_size;
        /* DVec2::clamp(min, max) = self.max(min).min(max) */
        v = v > lo ? v : lo;
        v = v < hi ? v : hi;
        out[k] = v <= 0.0 ? 0u : (v >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)v);
    }
}

/* TileTree::compute_tile_distance — tile_tree.rs:193-221 */
double bto_tile_distance(const bt_terrain_model* model, const bt_tile* tile, uint32_t view_side,
                         const double view_uv[2], float approximate_height, const double view_world_position[3]) {
    (void)view_side;
    double tile_count = (double)(1u << tile->lod);
    double vxy[2];
    compute_tree_xy(view_uv, tile_count, vxy);
    int32_t off[2] = {f64_to_i32_sat(vxy[0]) - (int32_t)tile->x, f64_to_i32_sat(vxy[1]) - (int32_t)tile->y};
    double o[2] = {fmod(vxy[0], 1.0), fmod(vxy[1], 1.0)};
    for (int k = 0; k < 2; ++k) { if (off[k] < 0) o[k] = 0.0; else if (off[k] > 0) o[k] = 1.0; }
    double uv[2] = {((double)(int32_t)tile->x + o[0]) / tile_count, ((double)(int32_t)tile->y + o[1]) / tile_count};
    double wp[3];
    bto_coordinate_world_position(model, tile->side, uv, approximate_height, wp);
    double d[3] = {wp[0] - view_world_position[0], wp[1] - view_world_position[1], wp[2] - view_world_position[2]};
    return sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
}

/* exposed so tests can project the view coordinate per side (TileTree::update, tile_tree.rs:274-275) */
void bto_project_to_side(const bt_terrain_model* model, uint32_t self_side, const double self_uv[2], uint32_t side, double out_uv[2]) {
    project_to_side(model, self_side, self_uv, side, out_uv);
}
