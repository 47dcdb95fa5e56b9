//! `libs`-shaped shim over libtokamak_b200: keeps the signatures the prover calls
//! (packages/backend/libs/src/bivariate_polynomial/mod.rs:1283-1416, group_structures/mod.rs:59-143,
//! prove/src/sigma_source.rs:50-123) and replaces the bodies.  Source only (no Rust toolchain in the build image).
//! Errors keep the reference's behaviour: every non-zero status becomes a panic! with tkm_last_error().
pub mod bivariate_polynomial;
pub mod group_structures;

use std::ffi::CStr;
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

/// 32-byte little-endian canonical scalar: the layout of icicle ScalarField::to_bytes_le.
#[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub struct ScalarField(pub [u8; 32]);

/// x || y, 2 x 48 bytes little-endian canonical; all-zero = identity (group_structures/mod.rs:889-893).
#[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub struct G1serde(pub [u8; 96]);
