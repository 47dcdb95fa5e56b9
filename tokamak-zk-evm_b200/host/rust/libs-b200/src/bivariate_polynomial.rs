//! DensePolynomialExt over a device-resident tkm_poly handle (RAII: Drop frees the device buffer).
use crate::{check, ctx, ScalarField};
use std::ops::{Add, Mul, Neg, Sub};
use tokamak_b200_sys as sys;

pub struct DensePolynomialExt {
    h: *mut sys::tkm_poly,
    pub x_degree: i64,
    pub y_degree: i64,
    pub x_size: usize,
    pub y_size: usize,
}

impl Drop for DensePolynomialExt {
    fn drop(&mut self) {
        unsafe { sys::tkm_poly_free(ctx(), self.h) };
    }
}

impl Clone for DensePolynomialExt {
    fn clone(&self) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::tkm_poly_clone(ctx(), self.h, &mut h) });
        Self { h, ..*self }
    }
}

/// init_ntt_domain_for_size (bivariate_polynomial/mod.rs:33-55)
pub fn init_ntt_domain_for_size(size: usize) -> Result<(), ()> {
    if size == 0 { panic!("NTT domain size must be non-zero."); }
    if !size.is_power_of_two() { panic!("NTT domain size must be a power of two."); }
    check(unsafe { sys::tkm_ntt_domain_init(ctx(), size.trailing_zeros()) });
    Ok(())
}

impl DensePolynomialExt {
    fn wrap(h: *mut sys::tkm_poly) -> Self {
        let (mut x, mut y) = (0usize, 0usize);
        check(unsafe { sys::tkm_poly_shape(h, &mut x, &mut y) });
        Self { h, x_degree: x as i64 - 1, y_degree: y as i64 - 1, x_size: x, y_size: y }
    }
    pub fn handle(&self) -> *mut sys::tkm_poly { self.h }

    /// from_coeffs (:1527-1551) for a host slice of canonical scalars
    pub fn from_coeffs(coeffs: &[ScalarField], x_size: usize, y_size: usize) -> Self {
        if x_size * y_size != coeffs.len() { panic!("Mismatch between the coefficient vector and the polynomial size") }
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::tkm_poly_from_coeffs_host(ctx(), coeffs.as_ptr() as *const u8, x_size, y_size, &mut h) });
        Self::wrap(h)
    }
    /// from_rou_evals (:1615-1644)
    pub fn from_rou_evals(evals: &[ScalarField], x_size: usize, y_size: usize, coset_x: Option<&ScalarField>, coset_y: Option<&ScalarField>) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::tkm_poly_from_evals_host(ctx(), evals.as_ptr() as *const u8, x_size, y_size, opt(coset_x), opt(coset_y), &mut h) });
        Self::wrap(h)
    }
    /// to_rou_evals (:1646-1674) -- no D->H->D round trip of the coefficients
    pub fn to_rou_evals(&self, coset_x: Option<&ScalarField>, coset_y: Option<&ScalarField>, evals: &mut [ScalarField]) {
        if evals.len() < self.x_size * self.y_size { panic!("Insufficient buffer length for to_rou_evals") }
        check(unsafe { sys::tkm_poly_to_evals_host(ctx(), self.h, opt(coset_x), opt(coset_y), evals.as_mut_ptr() as *mut u8) });
    }
    pub fn find_degree(&self) -> (i64, i64) {
        let (mut x, mut y) = (0i64, 0i64);
        check(unsafe { sys::tkm_poly_find_degree(ctx(), self.h, &mut x, &mut y) });
        (x, y)
    }
    pub fn resize(&mut self, tx: usize, ty: usize) {
        check(unsafe { sys::tkm_poly_resize(ctx(), self.h, tx, ty) });
        check(unsafe { sys::tkm_poly_shape(self.h, &mut self.x_size, &mut self.y_size) });
    }
    pub fn optimize_size(&mut self) {
        let (xd, yd) = self.find_degree();
        self.x_degree = xd;
        self.y_degree = yd;
        check(unsafe { sys::tkm_poly_optimize_size(ctx(), self.h) });
        check(unsafe { sys::tkm_poly_shape(self.h, &mut self.x_size, &mut self.y_size) });
    }
    pub fn mul_monomial(&self, ex: usize, ey: usize) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::tkm_poly_mul_monomial(ctx(), self.h, ex, ey, &mut h) });
        Self::wrap(h)
    }
    pub fn eval(&self, x: &ScalarField, y: &ScalarField) -> ScalarField {
        let mut out = ScalarField([0u8; 32]);
        check(unsafe { sys::tkm_poly_eval(ctx(), self.h, x.0.as_ptr(), y.0.as_ptr(), out.0.as_mut_ptr()) });
        out
    }
    pub fn scale_coeffs_x(&self, s: &ScalarField) -> Self { self.scale(Some(s), None) }
    pub fn scale_coeffs_y(&self, s: &ScalarField) -> Self { self.scale(None, Some(s)) }
    fn scale(&self, sx: Option<&ScalarField>, sy: Option<&ScalarField>) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::tkm_poly_scale_coeffs(ctx(), self.h, opt(sx), opt(sy), &mut h) });
        Self::wrap(h)
    }
    /// div_by_vanishing_opt (:2284-2410)
    pub fn div_by_vanishing_opt(&mut self, x_degree: i64, y_degree: i64) -> (Self, Self) {
        let (mut qx, mut qy) = (std::ptr::null_mut(), std::ptr::null_mut());
        check(unsafe { sys::tkm_poly_div_by_vanishing(ctx(), self.h, x_degree as usize, y_degree as usize, &mut qx, &mut qy) });
        check(unsafe { sys::tkm_poly_shape(self.h, &mut self.x_size, &mut self.y_size) });
        (Self::wrap(qx), Self::wrap(qy))
    }
    /// div_by_ruffini (:2412-2477)
    pub fn div_by_ruffini(&self, x: &ScalarField, y: &ScalarField) -> (Self, Self, ScalarField) {
        let (mut qx, mut qy) = (std::ptr::null_mut(), std::ptr::null_mut());
        let mut r = ScalarField([0u8; 32]);
        check(unsafe { sys::tkm_poly_div_by_ruffini(ctx(), self.h, x.0.as_ptr(), y.0.as_ptr(), &mut qx, &mut qy, r.0.as_mut_ptr()) });
        (Self::wrap(qx), Self::wrap(qy), r)
    }
    fn axpby(&self, ca: Option<&ScalarField>, b: Option<&Self>, cb: Option<&ScalarField>) -> Self {
        let mut h = std::ptr::null_mut();
        let bh = b.map(|p| p.h as *const sys::tkm_poly).unwrap_or(std::ptr::null());
        check(unsafe { sys::tkm_poly_axpby(ctx(), self.h, opt(ca), bh, opt(cb), &mut h) });
        Self::wrap(h)
    }
}

fn opt(s: Option<&ScalarField>) -> *const u8 { s.map(|v| v.0.as_ptr()).unwrap_or(std::ptr::null()) }
/// r - 1 = -1 mod r, little-endian
const MINUS_ONE: ScalarField = ScalarField([
    0x00, 0x00, 0x00, 0x00, 0xff, 0xff, 0xff, 0xff, 0xfe, 0x5b, 0xfe, 0xff, 0x02, 0xa4, 0xbd, 0x53,
    0x05, 0xd8, 0xa1, 0x09, 0x08, 0xd8, 0x39, 0x33, 0x48, 0x7d, 0x9d, 0x29, 0x53, 0xa7, 0xed, 0x73]);

impl<'a> Add<&'a DensePolynomialExt> for &'a DensePolynomialExt {
    type Output = DensePolynomialExt;
    fn add(self, rhs: Self) -> DensePolynomialExt { self.axpby(None, Some(rhs), None) }
}
impl<'a> Sub<&'a DensePolynomialExt> for &'a DensePolynomialExt {
    type Output = DensePolynomialExt;
    fn sub(self, rhs: Self) -> DensePolynomialExt { self.axpby(None, Some(rhs), Some(&MINUS_ONE)) }
}
impl<'a> Mul<&'a DensePolynomialExt> for &'a DensePolynomialExt {
    type Output = DensePolynomialExt;
    fn mul(self, rhs: Self) -> DensePolynomialExt {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::tkm_poly_mul(ctx(), self.h, rhs.h, &mut h) });
        DensePolynomialExt::wrap(h)
    }
}
impl<'a> Mul<&'a ScalarField> for &'a DensePolynomialExt {
    type Output = DensePolynomialExt;
    fn mul(self, rhs: &ScalarField) -> DensePolynomialExt { self.axpby(Some(rhs), None, None) }
}
impl<'a> Neg for &'a DensePolynomialExt {
    type Output = DensePolynomialExt;
    fn neg(self) -> DensePolynomialExt { self.axpby(Some(&MINUS_ONE), None, None) }
}
