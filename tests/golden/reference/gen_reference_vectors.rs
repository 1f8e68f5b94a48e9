// gen_reference_vectors -- dumps what the REAL crates compute for the pieces of the hot path that the oracle
// (oracle/snes_oracle.c) restates from their published algorithms: ssimulacra2 0.5.1 (+ yuvxyb 0.4.2), palette 0.7.6,
// cogset 0.2.0, at the versions aexoden/snesimage pins (Cargo.toml:12-25).  Every call below is spelled the way
// src/lib.rs spells it (line numbers in the comments), so the vectors pin the reference's own call sites.
//
//   python tests/golden/write_reference_inputs.py            # writes tests/golden/reference/inputs/*
//   cargo run --release --manifest-path tests/golden/reference/Cargo.toml -- tests/golden/reference/inputs \
//         > tests/golden/reference_vectors.json
//   python -m pytest tests/test_reference_vectors.py         # the oracle (and, with -m gpu, the GPU path) against them
//
// Floats are written as bit patterns (u32 / u64) next to their decimal form, so that "bit-exact" can be checked.
// This file has never been compiled in the repository's own image (no Rust toolchain there); it is ~150 lines of
// straight-line calls, kept simple on purpose.
use std::fs;

use cogset::{Euclid, Kmeans};
use palette::{color_difference::Ciede2000, FromColor, IntoColor, Lab, Srgb};
use serde_json::json;
use ssimulacra2::{compute_frame_ssimulacra2, ColorPrimaries, LinearRgb, Rgb, TransferCharacteristic};

fn f32v(v: f32) -> serde_json::Value { json!({"bits": v.to_bits(), "value": v}) }
fn f64v(v: f64) -> serde_json::Value { json!({"bits": v.to_bits().to_string(), "value": v}) }

fn read_rgba(path: &str) -> Vec<u8> {
    let d = fs::read(path).unwrap_or_else(|e| panic!("{path}: {e}"));
    assert_eq!(d.len(), 256 * 256 * 4, "{path}: expected a raw 256x256 RGBA8 file");
    d
}

// lib.rs:506-525: r, g, b / 255 as f32 (alpha ignored), sRGB transfer, BT.709 primaries
fn to_rgb(rgba: &[u8]) -> Rgb {
    let data = rgba.chunks_exact(4)
        .map(|p| [f32::from(p[0]) / 255.0, f32::from(p[1]) / 255.0, f32::from(p[2]) / 255.0])
        .collect::<Vec<_>>();
    Rgb::new(data, 256, 256, TransferCharacteristic::SRGB, ColorPrimaries::BT709).expect("Rgb::new")
}

// lib.rs:1092-1099
fn cielab(c1: [u8; 3], c2: [u8; 3]) -> f64 {
    let a: Lab = Srgb::new(c1[0], c1[1], c1[2]).into_format().into_color();
    let b: Lab = Srgb::new(c2[0], c2[1], c2[2]).into_format().into_color();
    a.difference(b).into()
}

fn main() {
    let dir = std::env::args().nth(1).unwrap_or_else(|| "tests/golden/reference/inputs".to_string());

    // 1. yuvxyb's sRGB -> linear for the 256 8-bit values, the way compute_frame_ssimulacra2 sees them (lib.rs:506-547)
    let ramp = (0..256).map(|v| [v as f32 / 255.0; 3]).collect::<Vec<_>>();
    let rgb = Rgb::new(ramp, 16, 16, TransferCharacteristic::SRGB, ColorPrimaries::BT709).expect("Rgb::new");
    let lin = LinearRgb::try_from(rgb).expect("LinearRgb::try_from");
    let eotf_yuvxyb = lin.data().iter().map(|p| f32v(p[0])).collect::<Vec<_>>();

    // 2. palette's Srgb<u8> -> linear f32 (the first half of lib.rs:101-103)
    let eotf_palette = (0..=255u8)
        .map(|v| f32v(Srgb::new(v, v, v).into_format::<f32>().into_linear().red))
        .collect::<Vec<_>>();

    // 3. Srgb<u8> -> Lab<D65, f32> on a colour grid (lib.rs:101-103, 344-346, 1092-1097)
    let mut lab_grid = Vec::new();
    for r in (0..=255u16).step_by(51) { for g in (0..=255u16).step_by(51) { for b in (0..=255u16).step_by(51) {
        let lab: Lab = Srgb::<u8>::new(r as u8, g as u8, b as u8).into_format().into_color();
        lab_grid.push(json!({"rgb": [r, g, b], "lab": [f32v(lab.l), f32v(lab.a), f32v(lab.b)]}));
    }}}

    // 4. color_distance_cielab on colour pairs (lib.rs:1090-1100); pairs.u8 holds 6 bytes per pair
    let pairs = fs::read(format!("{dir}/pairs.u8")).expect("pairs.u8");
    let ciede = pairs.chunks_exact(6)
        .map(|p| json!({"a": [p[0], p[1], p[2]], "b": [p[3], p[4], p[5]], "d": f64v(cielab([p[0], p[1], p[2]], [p[3], p[4], p[5]]))}))
        .collect::<Vec<_>>();

    // 5. Lab<f64> -> Srgb<u8> (lib.rs:141-142, 369-371); labs.f64 holds 3 little-endian f64 per colour
    let labs = fs::read(format!("{dir}/labs.f64")).expect("labs.f64");
    let lab_to_srgb = labs.chunks_exact(24).map(|c| {
        let v: Vec<f64> = c.chunks_exact(8).map(|b| f64::from_le_bytes(b.try_into().unwrap())).collect();
        let rgb: Srgb<u8> = Srgb::from_format(Srgb::from_color(Lab::new(v[0], v[1], v[2])));
        json!({"lab": [f64v(v[0]), f64v(v[1]), f64v(v[2])], "rgb": [rgb.red, rgb.green, rgb.blue]})
    }).collect::<Vec<_>>();

    // 6. cogset Kmeans::new(&points, k) (lib.rs:130, 366): points.f64 holds 3 little-endian f64 per point; k from k.txt
    let pts = fs::read(format!("{dir}/points.f64")).expect("points.f64");
    let points = pts.chunks_exact(24)
        .map(|c| { let v: Vec<f64> = c.chunks_exact(8).map(|b| f64::from_le_bytes(b.try_into().unwrap())).collect(); Euclid([v[0], v[1], v[2]]) })
        .collect::<Vec<_>>();
    let k: usize = fs::read_to_string(format!("{dir}/k.txt")).expect("k.txt").trim().parse().expect("k");
    let kmeans = Kmeans::new(&points, k);
    let clusters = kmeans.clusters().iter()
        .map(|(centre, members)| json!({"centre": [f64v(centre.0[0]), f64v(centre.0[1]), f64v(centre.0[2])], "members": members}))
        .collect::<Vec<_>>();

    // 7. compute_frame_ssimulacra2(src, dst) (lib.rs:547) on pairs of raw images: pairs.txt lists "src.rgba dst.rgba" per line
    let list = fs::read_to_string(format!("{dir}/image_pairs.txt")).expect("image_pairs.txt");
    let scores = list.lines().filter(|l| !l.trim().is_empty()).map(|l| {
        let mut it = l.split_whitespace();
        let (s, d) = (it.next().unwrap(), it.next().unwrap());
        let score = compute_frame_ssimulacra2(to_rgb(&read_rgba(&format!("{dir}/{s}"))), to_rgb(&read_rgba(&format!("{dir}/{d}"))))
            .expect("compute_frame_ssimulacra2");
        json!({"src": s, "dst": d, "ssimulacra2": f64v(score), "error": f64v(100.0 - score)})   // lib.rs:547
    }).collect::<Vec<_>>();

    let doc = json!({
        "crates": {"ssimulacra2": "0.5.1", "palette": "0.7.6", "cogset": "0.2.0"},
        "eotf_yuvxyb": eotf_yuvxyb, "eotf_palette": eotf_palette, "lab_grid": lab_grid, "ciede2000": ciede,
        "lab_to_srgb8": lab_to_srgb, "kmeans": {"k": k, "clusters": clusters}, "ssimulacra2": scores,
    });
    println!("{}", serde_json::to_string(&doc).expect("json"));
}
