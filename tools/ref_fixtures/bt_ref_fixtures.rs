//! Dropped into the reference crate as `src/bt_ref_fixtures.rs` by tools/ref_fixtures/apply.sh (`#[cfg(test)] mod bt_ref_fixtures;`).
//! Every test calls one CPU function of the reference's hot path on seeded inputs and writes inputs + outputs as JSON.
//! Floats are written as IEEE bit patterns (u32 / u64) so that the consumer compares exactly.
use crate::math::{Coordinate, TerrainModel, TerrainModelApproximation, TileCoordinate};
use crate::terrain::TerrainConfig;
use crate::terrain_data::{tile_atlas::TileAtlas, tile_tree::TileTree, AttachmentData, AttachmentFormat};
use crate::terrain_view::TerrainViewConfig;
use bevy::math::{DVec3, Vec2};
use std::fmt::Write as _;

fn out_path(name: &str) -> String {
    let dir = std::env::var("BT_REF_FIXTURES_DIR").unwrap_or_else(|_| "target/bt_ref_fixtures".to_string());
    std::fs::create_dir_all(&dir).unwrap();
    format!("{dir}/{name}")
}

/// the generator the Python consumer re-implements: x_{n+1} = x_n * 6364136223846793005 + 1442695040888963407 (mod 2^64)
struct Lcg(u64);
impl Lcg {
    fn next(&mut self) -> u64 {
        self.0 = self.0.wrapping_mul(6364136223846793005).wrapping_add(1442695040888963407);
        self.0 >> 11
    }
    fn unit(&mut self) -> f64 {
        self.next() as f64 / (1u64 << 53) as f64
    }
}

fn models() -> Vec<(&'static str, TerrainModel)> {
    vec![
        ("planar", TerrainModel::planar(DVec3::new(0.0, -100.0, 0.0), 2000.0, 0.0, 500.0)),
        ("sphere", TerrainModel::sphere(DVec3::ZERO, 6371000.0, -12000.0, 9000.0)),
        ("ellipsoid", TerrainModel::ellipsoid(DVec3::ZERO, 6378137.0, 6356752.314245, -12000.0, 9000.0)),
    ]
}

fn positions(name: &str, rng: &mut Lcg, n: usize) -> Vec<DVec3> {
    (0..n)
        .map(|_| {
            if name == "planar" {
                DVec3::new(rng.unit() * 1800.0 - 900.0, 100.0 + rng.unit() * 600.0, rng.unit() * 1800.0 - 900.0)
            } else {
                let d = DVec3::new(rng.unit() * 2.0 - 1.0, rng.unit() * 2.0 - 1.0, rng.unit() * 2.0 - 1.0).normalize();
                d * (6371000.0 + 1000.0 + rng.unit() * 400000.0)
            }
        })
        .collect()
}

#[test]
fn bt_ref_fixtures_sample() {
    // AttachmentData::sample — terrain_data/mod.rs:221-264
    let size = 8u32;
    let mut rng = Lcg(1);
    let texels: Vec<u16> = (0..size * size).map(|_| (rng.next() % 65536) as u16).collect();
    let data = AttachmentData::from_bytes(bytemuck::cast_slice(&texels), AttachmentFormat::R16);
    let mut s = String::from("{\"size\":8,\"texels\":[");
    for (i, t) in texels.iter().enumerate() { write!(s, "{}{}", if i > 0 { "," } else { "" }, t).unwrap(); }
    s.push_str("],\"samples\":[");
    for i in 0..96 {
        // uv such that the 2 x 2 footprint stays inside the tile (the reference indexes out of bounds otherwise)
        let uv = Vec2::new((0.07 + 0.86 * rng.unit()) as f32, (0.07 + 0.86 * rng.unit()) as f32);
        let v = data.sample(uv, size);
        write!(s, "{}{{\"uv\":[{},{}],\"value\":[{},{},{},{}]}}", if i > 0 { "," } else { "" }, uv.x.to_bits(), uv.y.to_bits(),
               v.x.to_bits(), v.y.to_bits(), v.z.to_bits(), v.w.to_bits()).unwrap();
    }
    s.push_str("]}");
    std::fs::write(out_path("sample.json"), s).unwrap();
}

#[test]
fn bt_ref_fixtures_mipmaps() {
    // AttachmentData::generate_mipmaps — terrain_data/mod.rs:143-219
    let size = 16u32;
    let mut rng = Lcg(2);
    let mut s = String::from("{\"size\":16,\"mip_level_count\":4,\"cases\":[");
    for (k, format) in [AttachmentFormat::R16, AttachmentFormat::Rgba8].into_iter().enumerate() {
        let bytes: Vec<u8> = match format {
            AttachmentFormat::R16 => {
                let t: Vec<u16> = (0..size * size).map(|_| { let v = (rng.next() % 65536) as u16; if rng.next() % 5 == 0 { 0 } else { v } }).collect();
                bytemuck::cast_slice(&t).to_vec()
            }
            _ => (0..size * size * 4).map(|_| (rng.next() % 256) as u8).collect(),
        };
        let mut data = AttachmentData::from_bytes(&bytes, format);
        data.generate_mipmaps(size, 4);
        let all = data.bytes();
        write!(s, "{}{{\"format\":\"{}\",\"mip0\":[", if k > 0 { "," } else { "" }, if k == 0 { "R16" } else { "Rgba8" }).unwrap();
        for (i, b) in bytes.iter().enumerate() { write!(s, "{}{}", if i > 0 { "," } else { "" }, b).unwrap(); }
        s.push_str("],\"with_mips\":[");
        for (i, b) in all.iter().enumerate() { write!(s, "{}{}", if i > 0 { "," } else { "" }, b).unwrap(); }
        s.push_str("]}");
    }
    s.push_str("]}");
    std::fs::write(out_path("mipmaps.json"), s).unwrap();
}

#[test]
fn bt_ref_fixtures_neighbours() {
    // TileCoordinate::neighbours — math/coordinate.rs:208-279
    let mut s = String::from("{\"cases\":[");
    let mut first = true;
    for spherical in [false, true] {
        for side in 0..(if spherical { 6 } else { 1 }) {
            for lod in 0..4u32 {
                let n = 1u32 << lod;
                for &(x, y) in &[(0, 0), (n - 1, 0), (0, n - 1), (n - 1, n - 1), (n / 2, n / 2), (n / 3, n - 1)] {
                    let t = TileCoordinate::new(side, lod, x, y);
                    write!(s, "{}{{\"spherical\":{},\"tile\":[{},{},{},{}],\"neighbours\":[", if first { "" } else { "," }, spherical, side, lod, x, y).unwrap();
                    first = false;
                    for (i, nb) in t.neighbours(spherical).enumerate() {
                        write!(s, "{}[{},{},{},{}]", if i > 0 { "," } else { "" }, nb.side, nb.lod, nb.x, nb.y).unwrap();
                    }
                    s.push_str("]}");
                }
            }
        }
    }
    s.push_str("]}");
    std::fs::write(out_path("neighbours.json"), s).unwrap();
}

#[test]
fn bt_ref_fixtures_from_world_position() {
    // Coordinate::from_world_position — math/coordinate.rs:69-108 (incl. TerrainModel::position_world_to_local, ellipsoid projection)
    let mut rng = Lcg(3);
    let mut s = String::from("{\"cases\":[");
    let mut first = true;
    for (name, model) in models() {
        for p in positions(name, &mut rng, 24) {
            let c = Coordinate::from_world_position(p, &model);
            write!(s, "{}{{\"model\":\"{}\",\"position\":[{},{},{}],\"side\":{},\"uv\":[{},{}]}}", if first { "" } else { "," }, name,
                   p.x.to_bits(), p.y.to_bits(), p.z.to_bits(), c.side, c.uv.x.to_bits(), c.uv.y.to_bits()).unwrap();
            first = false;
        }
    }
    s.push_str("]}");
    std::fs::write(out_path("from_world_position.json"), s).unwrap();
}

#[test]
fn bt_ref_fixtures_approximation() {
    // TerrainModelApproximation::compute — math/terrain_model.rs:263-360
    let mut rng = Lcg(4);
    let mut s = String::from("{\"origin_lod\":10,\"cases\":[");
    let mut first = true;
    for (name, model) in models() {
        let config = TerrainConfig { lod_count: 8, model: model.clone(), atlas_size: 64, path: "bt_ref_fixtures_none".to_string(), attachments: vec![] };
        let atlas = TileAtlas::new(&config);
        let mut tree = TileTree::new(&atlas, &TerrainViewConfig::default());
        for p in positions(name, &mut rng, 12) {
            tree.view_world_position = p;
            tree.approximate_height = 123.5;
            let a = TerrainModelApproximation::compute(&tree, &atlas);
            write!(s, "{}{{\"model\":\"{}\",\"position\":[{},{},{}],\"approximate_height\":{},\"sides\":[", if first { "" } else { "," }, name,
                   p.x.to_bits(), p.y.to_bits(), p.z.to_bits(), 123.5f32.to_bits()).unwrap();
            first = false;
            for (i, sd) in a.sides.iter().enumerate() {
                let v3 = |v: bevy::math::Vec3| format!("[{},{},{}]", v.x.to_bits(), v.y.to_bits(), v.z.to_bits());
                write!(s, "{}{{\"origin_xy\":[{},{}],\"origin_uv\":[{},{}],\"c\":{},\"c_s\":{},\"c_t\":{},\"c_ss\":{},\"c_st\":{},\"c_tt\":{}}}",
                       if i > 0 { "," } else { "" }, sd.origin_xy.x, sd.origin_xy.y, sd.origin_uv.x.to_bits(), sd.origin_uv.y.to_bits(),
                       v3(sd.c), v3(sd.c_s), v3(sd.c_t), v3(sd.c_ss), v3(sd.c_st), v3(sd.c_tt)).unwrap();
            }
            s.push_str("]}");
        }
    }
    s.push_str("]}");
    std::fs::write(out_path("approximation.json"), s).unwrap();
}
