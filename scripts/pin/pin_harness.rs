//! Known-answer dump from the REAL reference (packages/backend/libs + ICICLE v3.8.0 CPU backend).
//!
//! Not compiled in this repository (no Rust toolchain in the build image).  scripts/pin/pin_against_reference.sh copies
//! this file to packages/backend/libs/tests/pin_harness.rs on a machine that can build the reference and runs
//!     cargo test --release -p libs --test pin_harness -- --nocapture --test-threads=1
//! which prints one JSON document between the markers below; the script stores it as tests/golden/reference_pins.json,
//! where tests/test_reference_pins.py compares it with the oracle (and, on a GPU box, with libtokamak_b200).
//! What it pins: the primitive roots of unity ICICLE's NTT domain uses (the one convention identities cannot pin,
//! SURVEY.md §8c), a bivariate NTT with and without cosets, an MSM, and one commitment through encode_poly.
use icicle_bls12_381::curve::{ScalarCfg, ScalarField};
use icicle_core::ntt;
use icicle_core::traits::{FieldImpl, GenerateRandom};
use icicle_runtime::memory::HostSlice;
use libs::bivariate_polynomial::{init_ntt_domain_for_size, BivariatePolynomial, DensePolynomialExt};
use libs::group_structures::{msm_g1_bases, G1serde};

fn hex(s: &ScalarField) -> String {
    let b = s.to_bytes_le();
    let mut out = String::from("0x");
    for v in b.iter().rev() { out.push_str(&format!("{:02x}", v)); }
    out
}
fn hex_bytes_le(b: &[u8]) -> String {
    let mut out = String::from("0x");
    for v in b.iter().rev() { out.push_str(&format!("{:02x}", v)); }
    out
}
fn seeded(n: usize, seed: u64) -> Vec<ScalarField> {
    // SplitMix64 stream -> 32 little-endian bytes -> reduced mod r by from_bytes_le's caller (values below 2^254)
    let mut s = seed;
    (0..n)
        .map(|_| {
            let mut bytes = [0u8; 32];
            for k in 0..4 {
                s = s.wrapping_add(0x9e3779b97f4a7c15);
                let mut z = s;
                z = (z ^ (z >> 30)).wrapping_mul(0xbf58476d1ce4e5b9);
                z = (z ^ (z >> 27)).wrapping_mul(0x94d049bb133111eb);
                z ^= z >> 31;
                bytes[8 * k..8 * k + 8].copy_from_slice(&z.to_le_bytes());
            }
            bytes[31] &= 0x3f;
            ScalarField::from_bytes_le(&bytes)
        })
        .collect()
}
fn list(v: &[ScalarField]) -> String {
    format!("[{}]", v.iter().map(|s| format!("\"{}\"", hex(s))).collect::<Vec<_>>().join(","))
}

#[test]
fn dump_reference_pins() {
    libs::utils::check_device();
    init_ntt_domain_for_size(1 << 16).unwrap();
    let mut doc = String::from("{");
    // 1. roots of unity of the NTT domain
    let roots: Vec<String> = (1..=23u32)
        .map(|k| format!("\"{}\":\"{}\"", k, hex(&ntt::get_root_of_unity::<ScalarField>(1u64 << k))))
        .collect();
    doc.push_str(&format!("\"root_of_unity\":{{{}}},", roots.join(",")));
    // 2. bivariate NTT 8 x 4, plain and with cosets on both axes
    let (x, y) = (8usize, 4usize);
    let coeffs = seeded(x * y, 1);
    let poly = DensePolynomialExt::from_coeffs(HostSlice::from_slice(&coeffs), x, y);
    let mut evals = vec![ScalarField::zero(); x * y];
    poly.to_rou_evals(None, None, HostSlice::from_mut_slice(&mut evals));
    let (gx, gy) = (ScalarField::from_u32(5), ScalarField::from_u32(7));
    let mut evals_c = vec![ScalarField::zero(); x * y];
    poly.to_rou_evals(Some(&gx), Some(&gy), HostSlice::from_mut_slice(&mut evals_c));
    let back = DensePolynomialExt::from_rou_evals(HostSlice::from_slice(&coeffs), x, y, None, None);
    let mut inv = vec![ScalarField::zero(); x * y];
    back.copy_coeffs(0, HostSlice::from_mut_slice(&mut inv));
    doc.push_str(&format!(
        "\"bintt\":{{\"x\":{},\"y\":{},\"in\":{},\"fwd\":{},\"coset_x\":\"{}\",\"coset_y\":\"{}\",\"fwd_coset\":{},\"inv\":{}}},",
        x, y, list(&coeffs), list(&evals), hex(&gx), hex(&gy), list(&evals_c), list(&inv)
    ));
    // 3. MSM over bases k_i * G
    let ks = seeded(16, 2);
    let ss = seeded(16, 3);
    let g = G1serde::generator();
    let bases: Vec<G1serde> = ks.iter().map(|k| g * *k).collect();
    let r = msm_g1_bases(&bases, &ss);
    doc.push_str(&format!(
        "\"msm\":{{\"base_multipliers\":{},\"scalars\":{},\"result\":{{\"x\":\"{}\",\"y\":\"{}\"}}}},",
        list(&ks), list(&ss), hex_bytes_le(&r.0.x.to_bytes_le()), hex_bytes_le(&r.0.y.to_bytes_le())
    ));
    // 4. a random value so a stale file is noticed
    let _ = ScalarCfg::generate_random(1);
    doc.push_str("\"source\":\"packages/backend/libs, ICICLE v3.8.0\"}");
    println!("-----BEGIN REFERENCE PINS-----\n{}\n-----END REFERENCE PINS-----", doc);
}
