//! `libs`-shaped shim over libtokamak_b200: keeps the signatures the prover calls
//! (packages/backend/libs/src/bivariate_polynomial/mod.rs:140-260,532-1416, group_structures/mod.rs:59-300,888-947,
//! vector_operations/mod.rs:19-141,639-693, prove/src/sigma_source.rs:50-123) and replaces the bodies.  Source only (no
//! Rust toolchain in the build image); the same surface is compiled and run on the GPU through the C++ mirror
//! (host/cpp/tokamak_b200.hpp + test_libs.cpp).  Errors keep the reference's behaviour: every non-zero status becomes a
//! panic! carrying tkm_last_error().
pub mod bivariate_polynomial;
pub mod group_structures;
pub mod iotools;
pub mod sigma_source;
pub mod vector_operations;

use std::ffi::CStr;
use std::ops::{Add, Mul, Neg, Sub};
use std::sync::OnceLock;
use tokamak_b200_sys as sys;

pub struct Ctx(pub *mut sys::tkm_ctx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}
static CTX: OnceLock<Ctx> = OnceLock::new();

/// utils::check_device (libs/src/utils/mod.rs:88-110): binds device 0; panics without a GPU (no CPU fallback).
pub fn ctx() -> *mut sys::tkm_ctx {
    CTX.get_or_init(|| {
        let mut c = std::ptr::null_mut();
        check(unsafe { sys::tkm_ctx_create(0, &mut c) });
        Ctx(c)
    })
    .0
}

pub fn check(status: i32) {
    if status != 0 {
        let msg = unsafe { CStr::from_ptr(sys::tkm_last_error()) }.to_string_lossy().into_owned();
        panic!("{}", msg);
    }
}

/// BLS12-381 Fr, 32-byte little-endian canonical: the layout of icicle ScalarField::to_bytes_le, with the host arithmetic
/// the prover uses on single values (prove/src/lib.rs:2-4: zero, one, from_hex, from_bytes_le, to_bytes_le, from_u32, pow,
/// inv, + - * ==, to_string).
#[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub struct ScalarField(pub [u8; 32]);

const R: [u64; 4] = [0xffffffff00000001, 0x53bda402fffe5bfe, 0x3339d80809a1d805, 0x73eda753299d7d48];
const R2: [u64; 4] = [0xc999e990f3f29c6d, 0x2b6cedcb87925c23, 0x05d314967254398f, 0x0748d9d99f59ff11]; // 2^512 mod r
const INV: u64 = 0xfffffffeffffffff; // -r^-1 mod 2^64

fn geq_r(a: &[u64; 4]) -> bool {
    for i in (0..4).rev() {
        if a[i] != R[i] { return a[i] > R[i]; }
    }
    true
}
fn sub_r(a: &mut [u64; 4]) {
    let mut borrow = 0u128;
    for i in 0..4 {
        let d = (a[i] as u128).wrapping_sub(R[i] as u128).wrapping_sub(borrow);
        a[i] = d as u64;
        borrow = (d >> 64) & 1;
    }
}
/// a * b / 2^256 mod r (CIOS)
fn mont(a: &[u64; 4], b: &[u64; 4]) -> [u64; 4] {
    let mut t = [0u64; 6];
    for i in 0..4 {
        let mut c = 0u128;
        for j in 0..4 {
            c += (a[j] as u128) * (b[i] as u128) + t[j] as u128;
            t[j] = c as u64;
            c >>= 64;
        }
        c += t[4] as u128;
        t[4] = c as u64;
        t[5] = (c >> 64) as u64;
        let m = t[0].wrapping_mul(INV);
        let mut c = ((m as u128) * (R[0] as u128) + t[0] as u128) >> 64;
        for j in 1..4 {
            c += (m as u128) * (R[j] as u128) + t[j] as u128;
            t[j - 1] = c as u64;
            c >>= 64;
        }
        c += t[4] as u128;
        t[3] = c as u64;
        t[4] = t[5] + (c >> 64) as u64;
    }
    let mut r = [t[0], t[1], t[2], t[3]];
    if t[4] != 0 || geq_r(&r) { sub_r(&mut r); }
    r
}

impl ScalarField {
    pub fn limbs(&self) -> [u64; 4] {
        let mut l = [0u64; 4];
        for i in 0..4 { l[i] = u64::from_le_bytes(self.0[8 * i..8 * i + 8].try_into().unwrap()); }
        l
    }
    pub fn from_limbs(l: [u64; 4]) -> Self {
        let mut b = [0u8; 32];
        for i in 0..4 { b[8 * i..8 * i + 8].copy_from_slice(&l[i].to_le_bytes()); }
        Self(b)
    }
    pub fn zero() -> Self { Self([0u8; 32]) }
    pub fn one() -> Self { Self::from_u32(1) }
    pub fn from_u32(v: u32) -> Self { Self::from_limbs([v as u64, 0, 0, 0]) }
    pub fn from_bytes_le(b: &[u8]) -> Self {
        let mut a = [0u8; 32];
        a[..b.len().min(32)].copy_from_slice(&b[..b.len().min(32)]);
        let mut l = Self(a).limbs();
        while geq_r(&l) { sub_r(&mut l); }
        Self::from_limbs(l)
    }
    pub fn to_bytes_le(&self) -> Vec<u8> { self.0.to_vec() }
    /// "0x..." big-endian hex digits, reduced mod r (HexString parsing, libs/src/iotools/mod.rs:126-146)
    pub fn from_hex(hex: &str) -> Self {
        let digits = hex.strip_prefix("0x").or_else(|| hex.strip_prefix("0X")).unwrap_or(hex);
        let sixteen = Self::from_u32(16);
        let mut acc = Self::zero();
        for ch in digits.chars() {
            let d = ch.to_digit(16).expect("invalid hex digit in ScalarField::from_hex");
            acc = acc * sixteen + Self::from_u32(d);
        }
        acc
    }
    pub fn pow(&self, mut e: usize) -> Self {
        let (mut acc, mut base) = (Self::one(), *self);
        while e != 0 {
            if e & 1 == 1 { acc = acc * base; }
            base = base * base;
            e >>= 1;
        }
        acc
    }
    /// a^(r-2); inv(0) = 0 like the reference backend (bivariate_polynomial/mod.rs:2011-2013)
    pub fn inv(&self) -> Self {
        let e = [R[0] - 2, R[1], R[2], R[3]];
        let (mut acc, mut base) = (Self::one(), *self);
        for limb in e {
            for b in 0..64 {
                if (limb >> b) & 1 == 1 { acc = acc * base; }
                base = base * base;
            }
        }
        acc
    }
}
impl std::fmt::Display for ScalarField {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "0x")?;
        for b in self.0.iter().rev() { write!(f, "{:02x}", b)?; }
        Ok(())
    }
}
impl Add for ScalarField {
    type Output = Self;
    fn add(self, o: Self) -> Self {
        let (a, b) = (self.limbs(), o.limbs());
        let mut r = [0u64; 4];
        let mut c = 0u128;
        for i in 0..4 {
            c += a[i] as u128 + b[i] as u128;
            r[i] = c as u64;
            c >>= 64;
        }
        if c != 0 || geq_r(&r) { sub_r(&mut r); }
        Self::from_limbs(r)
    }
}
impl Sub for ScalarField {
    type Output = Self;
    fn sub(self, o: Self) -> Self {
        let (a, b) = (self.limbs(), o.limbs());
        let mut r = [0u64; 4];
        let mut borrow = 0u128;
        for i in 0..4 {
            let d = (a[i] as u128).wrapping_sub(b[i] as u128).wrapping_sub(borrow);
            r[i] = d as u64;
            borrow = (d >> 64) & 1;
        }
        if borrow != 0 {
            let mut c = 0u128;
            for i in 0..4 {
                c += r[i] as u128 + R[i] as u128;
                r[i] = c as u64;
                c >>= 64;
            }
        }
        Self::from_limbs(r)
    }
}
impl Mul for ScalarField {
    type Output = Self;
    fn mul(self, o: Self) -> Self { Self::from_limbs(mont(&mont(&self.limbs(), &o.limbs()), &R2)) }
}
impl Neg for ScalarField {
    type Output = Self;
    fn neg(self) -> Self { Self::zero() - self }
}

/// ScalarCfg::generate_random: uniform values below 2^254 from a xorshift stream seeded by the OS clock.
pub struct ScalarCfg;
impl ScalarCfg {
    pub fn generate_random(n: usize) -> Vec<ScalarField> {
        use std::time::{SystemTime, UNIX_EPOCH};
        static STATE: OnceLock<std::sync::Mutex<u64>> = OnceLock::new();
        let st = STATE.get_or_init(|| {
            std::sync::Mutex::new(SystemTime::now().duration_since(UNIX_EPOCH).map(|d| d.as_nanos() as u64).unwrap_or(1) | 1)
        });
        let mut s = st.lock().unwrap();
        (0..n)
            .map(|_| {
                let mut l = [0u64; 4];
                for w in l.iter_mut() {
                    *s ^= *s << 13;
                    *s ^= *s >> 7;
                    *s ^= *s << 17;
                    *w = *s;
                }
                l[3] &= 0x3fffffffffffffff;
                ScalarField::from_limbs(l)
            })
            .collect()
    }
}

/// x || y, 2 x 48 bytes little-endian canonical; all-zero = identity (group_structures/mod.rs:889-893).
#[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub struct G1serde(pub [u8; 96]);

/// G1serde ops (group_structures/mod.rs:895-947): through projective and back to affine in the reference; one device call here.
impl G1serde {
    pub fn zero() -> Self { Self([0u8; 96]) }
    pub fn neg(&self) -> Self {
        const Q: [u64; 6] = [0xb9feffffffffaaab, 0x1eabfffeb153ffff, 0x6730d2a0f6b0f624, 0x64774b84f38512bf, 0x4b1ba7b6434bacd7, 0x1a0111ea397fe69a];
        if self.0[48..].iter().all(|b| *b == 0) { return *self; }
        let mut out = *self;
        let mut borrow = 0u128;
        for i in 0..6 {
            let y = u64::from_le_bytes(self.0[48 + 8 * i..56 + 8 * i].try_into().unwrap());
            let d = (Q[i] as u128).wrapping_sub(y as u128).wrapping_sub(borrow);
            out.0[48 + 8 * i..56 + 8 * i].copy_from_slice(&(d as u64).to_le_bytes());
            borrow = (d >> 64) & 1;
        }
        out
    }
}
impl Add for G1serde {
    type Output = Self;
    fn add(self, o: Self) -> Self {
        let mut out = G1serde::zero();
        check(unsafe { sys::tkm_g1_add(ctx(), self.0.as_ptr(), o.0.as_ptr(), out.0.as_mut_ptr()) });
        out
    }
}
impl Sub for G1serde {
    type Output = Self;
    fn sub(self, o: Self) -> Self { self + o.neg() }
}
impl Mul<ScalarField> for G1serde {
    type Output = Self;
    fn mul(self, k: ScalarField) -> Self {
        let mut out = G1serde::zero();
        check(unsafe { sys::tkm_g1_mul(ctx(), self.0.as_ptr(), k.0.as_ptr(), out.0.as_mut_ptr()) });
        out
    }
}
impl Mul<G1serde> for ScalarField {
    type Output = G1serde;
    fn mul(self, p: G1serde) -> G1serde { p * self }
}
