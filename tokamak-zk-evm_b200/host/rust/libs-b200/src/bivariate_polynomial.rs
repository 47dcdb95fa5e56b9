//! DensePolynomialExt over a device-resident tkm_poly handle (RAII: Drop frees the device buffer).
use crate::{check, ctx, ScalarField};
use std::ops::{Add, AddAssign, Mul, Neg, Sub, SubAssign};
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

/// ntt::get_root_of_unity (uses at bivariate_polynomial/mod.rs:505, prove/src/lib.rs:1971-1972): the primitive n-th root
pub fn get_root_of_unity(n: u64) -> ScalarField {
    assert!(n.is_power_of_two(), "root order must be a power of two");
    let mut out = [0u8; 32];
    check(unsafe { sys::tkm_root_of_unity(n.trailing_zeros(), out.as_mut_ptr()) });
    ScalarField::from_bytes_le(&out)
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
    /// zero (:1283-1416): the 1 x 1 zero polynomial
    pub fn zero() -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::tkm_poly_zero(ctx(), 1, 1, &mut h) });
        Self::wrap(h)
    }
    pub fn is_zero(&self) -> bool { self.find_degree().0 < 0 }
    pub fn degree(&self) -> (i64, i64) { (self.x_degree, self.y_degree) }
    /// from_coeffs with a DeviceSlice of Montgomery-form elements (:1527-1551)
    pub fn from_device(dev_coeffs: *const std::ffi::c_void, x_size: usize, y_size: usize) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::tkm_poly_from_device(ctx(), dev_coeffs, x_size, y_size, &mut h) });
        Self::wrap(h)
    }
    /// copy_coeffs (:1676-1682) into a host slice of canonical scalars
    pub fn copy_coeffs(&self, _start_idx: u64, coeffs: &mut [ScalarField]) {
        if coeffs.len() < self.x_size * self.y_size { panic!("Insufficient buffer length for copy_coeffs") }
        check(unsafe { sys::tkm_poly_copy_coeffs_host(ctx(), self.h, coeffs.as_mut_ptr() as *mut u8) });
    }
    pub fn get_coeff(&self, idx_x: u64, idx_y: u64) -> ScalarField {
        if idx_x as usize >= self.x_size || idx_y as usize >= self.y_size { panic!("The index exceeds polynomial size.") }
        let mut all = vec![ScalarField::zero(); self.x_size * self.y_size];
        self.copy_coeffs(0, &mut all);
        all[idx_x as usize * self.y_size + idx_y as usize]
    }
    /// eval_x / eval_y (:1719-1740): partial evaluations, shapes 1 x y_size and x_size x 1
    pub fn eval_x(&self, x: &ScalarField) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::tkm_poly_eval_x(ctx(), self.h, x.0.as_ptr(), &mut h) });
        Self::wrap(h)
    }
    pub fn eval_y(&self, y: &ScalarField) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::tkm_poly_eval_y(ctx(), self.h, y.0.as_ptr(), &mut h) });
        Self::wrap(h)
    }
    /// get_univariate_polynomial_x / _y (:1752-1782): one row / column as a univariate polynomial
    pub fn get_univariate_polynomial_x(&self, idx_y: u64) -> Self {
        let mut all = vec![ScalarField::zero(); self.x_size * self.y_size];
        self.copy_coeffs(0, &mut all);
        let col: Vec<ScalarField> = (0..self.x_size).map(|i| all[i * self.y_size + idx_y as usize]).collect();
        Self::from_coeffs(&col, self.x_size, 1)
    }
    pub fn get_univariate_polynomial_y(&self, idx_x: u64) -> Self {
        let mut all = vec![ScalarField::zero(); self.x_size * self.y_size];
        self.copy_coeffs(0, &mut all);
        let s = idx_x as usize * self.y_size;
        Self::from_coeffs(&all[s..s + self.y_size], 1, self.y_size)
    }
    /// divide_x / divide_y (:1998-2094): line-wise long division by a univariate denominator -> (quotient, remainder)
    pub fn divide_x(&self, denominator: &Self) -> (Self, Self) { self.divide_uni(denominator, false) }
    pub fn divide_y(&self, denominator: &Self) -> (Self, Self) { self.divide_uni(denominator, true) }
    fn divide_uni(&self, denom: &Self, y_dir: bool) -> (Self, Self) {
        let (mut q, mut r) = (std::ptr::null_mut(), std::ptr::null_mut());
        check(unsafe { sys::tkm_poly_divide_uni(ctx(), self.h, denom.h, y_dir as i32, &mut q, &mut r) });
        (Self::wrap(q), Self::wrap(r))
    }
    /// div_by_vanishing (legacy coset formulation, :2096-2282): the same unique decomposition as div_by_vanishing_opt
    /// (Q_Y's X-degree stays below c); the cache of inverted denominators is not needed on this path.
    pub fn div_by_vanishing(&mut self, x_degree: i64, y_degree: i64, _cache: &mut DivByVanishingCache) -> (Self, Self) {
        self.div_by_vanishing_opt(x_degree, y_degree)
    }
    /// poly_comb! (prove/src/lib.rs:30-38) and its shifted helpers (:48-124) in one pass: sum of c * X^sx * Y^sy * p
    pub fn lincomb(terms: &[(ScalarField, &Self, u32, u32)]) -> Self {
        if terms.is_empty() { panic!("empty polynomial combination") }
        let hs: Vec<*const sys::tkm_poly> = terms.iter().map(|t| t.1.h as *const sys::tkm_poly).collect();
        let cs: Vec<ScalarField> = terms.iter().map(|t| t.0).collect();
        let sx: Vec<u32> = terms.iter().map(|t| t.2).collect();
        let sy: Vec<u32> = terms.iter().map(|t| t.3).collect();
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::tkm_poly_lincomb(ctx(), terms.len() as u32, hs.as_ptr(), cs.as_ptr() as *const u8, sx.as_ptr(), sy.as_ptr(), &mut h) });
        Self::wrap(h)
    }
    fn add_scalar(&self, s: &ScalarField) -> Self {
        let out = self.clone();
        check(unsafe { sys::tkm_poly_add_scalar(ctx(), out.h, s.0.as_ptr()) });
        out
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

// ---- the remaining operator impls of :532-1281: += / -=, poly +/- scalar, scalar +/- poly, scalar * poly
impl<'a> AddAssign<&'a DensePolynomialExt> for DensePolynomialExt {
    fn add_assign(&mut self, rhs: &'a DensePolynomialExt) { *self = &*self + rhs; }
}
impl<'a> SubAssign<&'a DensePolynomialExt> for DensePolynomialExt {
    fn sub_assign(&mut self, rhs: &'a DensePolynomialExt) { *self = &*self - rhs; }
}
impl<'a> Add<&'a ScalarField> for &'a DensePolynomialExt {
    type Output = DensePolynomialExt;
    fn add(self, rhs: &ScalarField) -> DensePolynomialExt { self.add_scalar(rhs) }
}
impl<'a> Sub<&'a ScalarField> for &'a DensePolynomialExt {
    type Output = DensePolynomialExt;
    fn sub(self, rhs: &ScalarField) -> DensePolynomialExt { self.add_scalar(&(-*rhs)) }
}
impl<'a> Add<&'a DensePolynomialExt> for &'a ScalarField {
    type Output = DensePolynomialExt;
    fn add(self, rhs: &DensePolynomialExt) -> DensePolynomialExt { rhs.add_scalar(self) }
}
impl<'a> Sub<&'a DensePolynomialExt> for &'a ScalarField {
    type Output = DensePolynomialExt;
    fn sub(self, rhs: &DensePolynomialExt) -> DensePolynomialExt { (-rhs).add_scalar(self) }
}
impl<'a> Mul<&'a DensePolynomialExt> for &'a ScalarField {
    type Output = DensePolynomialExt;
    fn mul(self, rhs: &DensePolynomialExt) -> DensePolynomialExt { rhs * self }
}

/// DivByVanishingCache (:88-110): kept for signature compatibility; the device path needs no cached denominators.
#[derive(Default)]
pub struct DivByVanishingCache {
    pub denom_x_eval_inv: Box<[ScalarField]>,
    pub denom_y_eval_inv: Box<[ScalarField]>,
    pub denom_x_axis_inv: Box<[ScalarField]>,
    pub denom_y_axis_inv: Box<[ScalarField]>,
}

/// domain_size_for_degree (:438-444)
fn domain_size_for_degree(degree: i64) -> usize {
    if degree < 0 { 1 } else { ((degree + 1) as usize).next_power_of_two() }
}

/// PolyExpr (:140-260): expression DAG over borrowed polynomials.  evaluate_fused(_with_domain) compiles the DAG to a
/// postfix program and hands it to tkm_polyexpr_eval: one forward biNTT per distinct leaf (pointer-keyed like the
/// reference's leaf cache :459-502), ONE pointwise kernel for the whole DAG, one inverse biNTT.
#[derive(Clone)]
pub enum PolyExpr<'a> {
    Poly(&'a DensePolynomialExt),
    Scalar(ScalarField),
    Add(Box<PolyExpr<'a>>, Box<PolyExpr<'a>>),
    Sub(Box<PolyExpr<'a>>, Box<PolyExpr<'a>>),
    Mul(Box<PolyExpr<'a>>, Box<PolyExpr<'a>>),
    Scale(ScalarField, Box<PolyExpr<'a>>),
    MulXMinusOne(Box<PolyExpr<'a>>),
    Sum(Vec<PolyExpr<'a>>),
    /// p(X / w_mx, Y / w_my), w_m the primitive m-th root of unity (0 = axis not scaled): an addition to the reference's
    /// constructors -- the leaf of `p.scale_coeffs_x(w_mx^-1).scale_coeffs_y(w_my^-1)` read rotated out of p's transform
    /// (TKM_PEX_LEAF_SHIFT), as prove2 needs for r(X/w, Y) and r(X/w, Y/w) (prove/src/lib.rs:2110-2146)
    PolyOverRoots(&'a DensePolynomialExt, u64, u64),
}

#[derive(Default)]
struct Program {
    leaves: Vec<*const sys::tkm_poly>,
    consts: Vec<ScalarField>,
    ops: Vec<u32>,
}
impl Program {
    fn leaf(&mut self, p: &DensePolynomialExt) -> u32 {
        let h = p.h as *const sys::tkm_poly;
        if let Some(i) = self.leaves.iter().position(|q| *q == h) { return i as u32; }
        self.leaves.push(h);
        self.leaves.len() as u32 - 1
    }
    fn konst(&mut self, s: ScalarField) -> u32 {
        if let Some(i) = self.consts.iter().position(|q| *q == s) { return i as u32; }
        self.consts.push(s);
        self.consts.len() as u32 - 1
    }
}

impl<'a> PolyExpr<'a> {
    pub fn poly(poly: &'a DensePolynomialExt) -> Self { Self::Poly(poly) }
    pub fn poly_over_roots(poly: &'a DensePolynomialExt, mx: u64, my: u64) -> Self {
        assert!(mx & mx.wrapping_sub(1) == 0 && my & my.wrapping_sub(1) == 0, "the root's order must be a power of two");
        Self::PolyOverRoots(poly, mx, my)
    }
    pub fn scalar(scalar: ScalarField) -> Self { Self::Scalar(scalar) }
    pub fn add(lhs: Self, rhs: Self) -> Self { Self::Add(Box::new(lhs), Box::new(rhs)) }
    pub fn sub(lhs: Self, rhs: Self) -> Self { Self::Sub(Box::new(lhs), Box::new(rhs)) }
    pub fn mul(lhs: Self, rhs: Self) -> Self { Self::Mul(Box::new(lhs), Box::new(rhs)) }
    pub fn scale(scalar: ScalarField, expr: Self) -> Self { Self::Scale(scalar, Box::new(expr)) }
    pub fn mul_x_minus_one(expr: Self) -> Self { Self::MulXMinusOne(Box::new(expr)) }
    pub fn weighted_sum(terms: Vec<(ScalarField, Self)>) -> Self {
        Self::Sum(terms.into_iter().map(|(s, e)| Self::scale(s, e)).collect())
    }

    /// evaluate_coeffs (:190-218): the coefficient-domain operators, node by node
    pub fn evaluate_coeffs(&self) -> DensePolynomialExt {
        match self {
            Self::Poly(p) => (*p).clone(),
            Self::PolyOverRoots(p, mx, my) => {
                let mut q = (*p).clone();
                if *mx != 0 { q = q.scale_coeffs_x(&get_root_of_unity(*mx).inv()); }
                if *my != 0 { q = q.scale_coeffs_y(&get_root_of_unity(*my).inv()); }
                q
            }
            Self::Scalar(s) => DensePolynomialExt::from_coeffs(&[*s], 1, 1),
            Self::Add(l, r) => &l.evaluate_coeffs() + &r.evaluate_coeffs(),
            Self::Sub(l, r) => &l.evaluate_coeffs() - &r.evaluate_coeffs(),
            Self::Mul(l, r) => &l.evaluate_coeffs() * &r.evaluate_coeffs(),
            Self::Scale(s, e) => &e.evaluate_coeffs() * s,
            Self::MulXMinusOne(e) => {
                let p = e.evaluate_coeffs();
                &p.mul_monomial(1, 0) - &p
            }
            Self::Sum(terms) => {
                let mut it = terms.iter();
                let Some(first) = it.next() else { return DensePolynomialExt::zero() };
                let mut acc = first.evaluate_coeffs();
                for t in it { acc += &t.evaluate_coeffs(); }
                acc
            }
        }
    }

    /// degree bound (:262-309); (-1, -1) = the zero polynomial
    pub fn degree_bound(&self) -> (i64, i64) {
        match self {
            Self::Poly(p) | Self::PolyOverRoots(p, _, _) => p.find_degree(),
            Self::Scalar(s) => if *s == ScalarField::zero() { (-1, -1) } else { (0, 0) },
            Self::Add(l, r) | Self::Sub(l, r) => {
                let (a, b) = (l.degree_bound(), r.degree_bound());
                (a.0.max(b.0), a.1.max(b.1))
            }
            Self::Mul(l, r) => {
                let (a, b) = (l.degree_bound(), r.degree_bound());
                if a.0 < 0 || a.1 < 0 || b.0 < 0 || b.1 < 0 { (-1, -1) } else { (a.0 + b.0, a.1 + b.1) }
            }
            Self::Scale(s, e) => if *s == ScalarField::zero() { (-1, -1) } else { e.degree_bound() },
            Self::MulXMinusOne(e) => {
                let d = e.degree_bound();
                if d.0 < 0 || d.1 < 0 { (-1, -1) } else { (d.0 + 1, d.1) }
            }
            Self::Sum(terms) => terms.iter().map(|t| t.degree_bound()).fold((-1, -1), |a, b| (a.0.max(b.0), a.1.max(b.1))),
        }
    }

    pub fn evaluate_fused(&self) -> DensePolynomialExt {
        let (xd, yd) = self.degree_bound();
        self.evaluate_fused_with_domain(domain_size_for_degree(xd), domain_size_for_degree(yd))
    }

    /// evaluate_fused_with_domain (:227-260)
    pub fn evaluate_fused_with_domain(&self, target_x_size: usize, target_y_size: usize) -> DensePolynomialExt {
        if !target_x_size.is_power_of_two() || !target_y_size.is_power_of_two() {
            panic!("Fused polynomial expression domains must be powers of two.");
        }
        let (xd, yd) = self.degree_bound();
        if domain_size_for_degree(xd) > target_x_size || domain_size_for_degree(yd) > target_y_size {
            panic!("Fused polynomial expression domain is too small for the expression degree.");
        }
        let mut pr = Program::default();
        self.emit(&mut pr);
        if pr.consts.is_empty() { pr.consts.push(ScalarField::zero()); }
        let mut h = std::ptr::null_mut();
        check(unsafe {
            sys::tkm_polyexpr_eval(ctx(), pr.leaves.as_ptr(), pr.leaves.len() as u32, pr.ops.as_ptr(), pr.ops.len() as u32,
                                   pr.consts.as_ptr() as *const u8, pr.consts.len() as u32, target_x_size, target_y_size, &mut h)
        });
        DensePolynomialExt::wrap(h)
    }

    fn emit(&self, pr: &mut Program) {
        match self {
            Self::Poly(p) => { let i = pr.leaf(p); pr.ops.push(sys::TKM_PEX_LEAF | i << 8); }
            Self::PolyOverRoots(p, mx, my) => {
                let field = |m: u64| 64 - m.leading_zeros();  // log2(m) + 1, 0 = axis not scaled
                let i = pr.leaf(p);
                pr.ops.push(sys::TKM_PEX_LEAF_SHIFT | (i | field(*mx) << 4 | field(*my) << 10) << 8);
            }
            Self::Scalar(s) => { let i = pr.konst(*s); pr.ops.push(sys::TKM_PEX_CONST | i << 8); }
            Self::Add(l, r) => { l.emit(pr); r.emit(pr); pr.ops.push(sys::TKM_PEX_ADD); }
            Self::Sub(l, r) => { l.emit(pr); r.emit(pr); pr.ops.push(sys::TKM_PEX_SUB); }
            Self::Mul(l, r) => { l.emit(pr); r.emit(pr); pr.ops.push(sys::TKM_PEX_MUL); }
            Self::Scale(s, e) => {
                e.emit(pr);
                if *s != ScalarField::one() { let i = pr.konst(*s); pr.ops.push(sys::TKM_PEX_SCALE | i << 8); }
            }
            Self::MulXMinusOne(e) => { e.emit(pr); pr.ops.push(sys::TKM_PEX_XM1); }
            Self::Sum(terms) => {
                if terms.is_empty() { let i = pr.konst(ScalarField::zero()); pr.ops.push(sys::TKM_PEX_CONST | i << 8); }
                for (k, t) in terms.iter().enumerate() {
                    t.emit(pr);
                    if k > 0 { pr.ops.push(sys::TKM_PEX_ADD); }
                }
            }
        }
    }
}
