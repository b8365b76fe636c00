
// ---- appended to src/terrain_data/tile_tree.rs by tools/ref_fixtures/apply.sh (compute_origin / compute_tile_distance are private here) ----
#[cfg(test)]
mod bt_ref_fixtures_tile_tree {
    use super::*;
    use crate::terrain::TerrainConfig;
    use std::fmt::Write as _;

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

    #[test]
    fn bt_ref_fixtures_tile_tree() {
        // TileTree::compute_origin / compute_tile_distance — terrain_data/tile_tree.rs:180-221
        let dir = std::env::var("BT_REF_FIXTURES_DIR").unwrap_or_else(|_| "target/bt_ref_fixtures".to_string());
        std::fs::create_dir_all(&dir).unwrap();
        let mut rng = Lcg(5);
        let mut s = String::from("{\"tree_size\":8,\"cases\":[");
        let mut first = true;
        for (name, model) in [("planar", TerrainModel::planar(DVec3::new(0.0, -100.0, 0.0), 2000.0, 0.0, 500.0)),
                              ("sphere", TerrainModel::sphere(DVec3::ZERO, 6371000.0, -12000.0, 9000.0))] {
            let config = TerrainConfig { lod_count: 8, model: model.clone(), atlas_size: 64, path: "bt_ref_fixtures_none".to_string(), attachments: vec![] };
            let atlas = TileAtlas::new(&config);
            let mut tree = TileTree::new(&atlas, &TerrainViewConfig::default());
            for _ in 0..16 {
                let p = if name == "planar" {
                    DVec3::new(rng.unit() * 1800.0 - 900.0, 100.0 + rng.unit() * 600.0, rng.unit() * 1800.0 - 900.0)
                } else {
                    DVec3::new(rng.unit() * 2.0 - 1.0, rng.unit() * 2.0 - 1.0, rng.unit() * 2.0 - 1.0).normalize() * (6371000.0 + 1000.0 + rng.unit() * 400000.0)
                };
                tree.view_world_position = p;
                tree.approximate_height = 77.25;
                let view = Coordinate::from_world_position(p, &model);
                let lod = (rng.next() % 8) as u32;
                let origin = tree.compute_origin(view, lod);
                let n = 1u32 << lod;
                let tile = TileCoordinate::new(view.side, lod, (rng.next() % n as u64) as u32, (rng.next() % n as u64) as u32);
                let d = tree.compute_tile_distance(tile, view, &model);
                write!(s, "{}{{\"model\":\"{}\",\"position\":[{},{},{}],\"approximate_height\":{},\"view_side\":{},\"view_uv\":[{},{}],\"lod\":{},\"origin\":[{},{}],\"tile\":[{},{},{},{}],\"distance\":{}}}",
                       if first { "" } else { "," }, name, p.x.to_bits(), p.y.to_bits(), p.z.to_bits(), 77.25f32.to_bits(), view.side, view.uv.x.to_bits(), view.uv.y.to_bits(),
                       lod, origin.x, origin.y, tile.side, tile.lod, tile.x, tile.y, d.to_bits()).unwrap();
                first = false;
            }
        }
        s.push_str("]}");
        std::fs::write(format!("{dir}/tile_tree.json"), s).unwrap();
    }
}
