//! Commitments: Sigma1::encode_poly (group_structures/mod.rs:59-119, iotools/mod.rs:2041-2113) and msm_g1_bases
//! (:127-143) over a device-resident CRS uploaded once.
use crate::bivariate_polynomial::DensePolynomialExt;
use crate::{check, ctx, G1serde, ScalarField};
use tokamak_b200_sys as sys;

pub struct Sigma1Device {
    h: *mut sys::tkm_crs,
    pub rs_x_size: usize,
    pub rs_y_size: usize,
}

impl Drop for Sigma1Device {
    fn drop(&mut self) { unsafe { sys::tkm_crs_free(ctx(), self.h) }; }
}

impl Sigma1Device {
    /// `xy_powers` as stored in the archived CRS: 2 x 48 little-endian bytes per point, index rs_y*h + i <-> x^h y^i
    /// (iotools/mod.rs:1701-1706; prove/src/sigma_source.rs:22-32 mmaps the file, so this is one cudaMemcpy of the mapping).
    pub fn upload(xy_powers: &[G1serde], rs_x_size: usize, rs_y_size: usize) -> Self {
        assert_eq!(xy_powers.len(), rs_x_size * rs_y_size);
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::tkm_crs_upload(ctx(), xy_powers.as_ptr() as *const u8, rs_x_size, rs_y_size, &mut h) });
        Self { h, rs_x_size, rs_y_size }
    }
    pub fn encode_poly(&self, poly: &mut DensePolynomialExt) -> G1serde {
        let mut out = G1serde([0u8; 96]);
        check(unsafe { sys::tkm_poly_commit(ctx(), poly.handle(), self.h, out.0.as_mut_ptr()) });
        out
    }
}

/// msm_g1_bases (group_structures/mod.rs:127-143): host scalars and host bases, one affine result.
pub fn msm_g1_bases(scalars: &[ScalarField], bases: &[G1serde]) -> G1serde {
    if scalars.len() != bases.len() { panic!("msm input length mismatch"); }
    let mut out = G1serde([0u8; 96]);
    check(unsafe { sys::tkm_msm_g1_host(ctx(), scalars.as_ptr() as *const u8, bases.as_ptr() as *const u8, scalars.len(), out.0.as_mut_ptr()) });
    out
}

/// A second resident table for the sparse encoders (gamma_inv_o_inst, eta_inv_li_o_inter_alpha4_kj, delta_inv_li_o_prv;
/// group_structures/mod.rs:361-394), rows x cols like the reference's boxed 2-D arrays, uploaded once from the archive.
pub struct G1Table {
    h: *mut sys::tkm_crs,
    pub rows: usize,
    pub cols: usize,
}
impl Drop for G1Table {
    fn drop(&mut self) { unsafe { sys::tkm_crs_free(ctx(), self.h) }; }
}
impl G1Table {
    pub fn upload(points: &[G1serde], rows: usize, cols: usize) -> Self {
        assert_eq!(points.len(), rows * cols);
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::tkm_crs_upload(ctx(), points.as_ptr() as *const u8, rows, cols, &mut h) });
        Self { h, rows, cols }
    }
    /// msm_g1_bases over gathered entries (encode_o_pub_fix_common / encode_o_pub_free_common / encode_statement_common,
    /// :145-300): sum_k scalars[k] * table[idx[k]]; empty input gives the identity (:131-133).
    pub fn msm_indexed(&self, scalars: &[ScalarField], idx: &[u32]) -> G1serde {
        if scalars.len() != idx.len() { panic!("Mismatch between the numbers of bases and scalars"); }
        let mut out = G1serde::zero();
        if scalars.is_empty() { return out; }
        unsafe {
            let (mut ds, mut di, mut dt) = (std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut());
            check(sys::tkm_crs_device_ptr(self.h, &mut dt, std::ptr::null_mut(), std::ptr::null_mut()));
            check(sys::tkm_dev_alloc(ctx(), scalars.len() * 32, &mut ds));
            check(sys::tkm_dev_alloc(ctx(), idx.len() * 4, &mut di));
            check(sys::tkm_memcpy_h2d(ctx(), ds, scalars.as_ptr() as *const _, scalars.len() * 32));
            check(sys::tkm_memcpy_h2d(ctx(), di, idx.as_ptr() as *const _, idx.len() * 4));
            let st = sys::tkm_msm_g1_indexed(ctx(), ds, 0, dt, di, scalars.len(), out.0.as_mut_ptr());
            sys::tkm_dev_free(ctx(), ds);
            sys::tkm_dev_free(ctx(), di);
            check(st);
        }
        out
    }
    /// The same MSM queued (tkm_msm_g1_indexed_begin): returns at once with a ticket; `PendingG1::get` resolves it
    /// (tkm_commit_end).  Prover::init's O_pub_free / O_mid / O_prv and their blinding sums (prove/src/lib.rs:1092-1176) are
    /// independent, so a caller queues them all and collects them together: their serial tails overlap.
    pub fn msm_indexed_begin(&self, scalars: &[ScalarField], idx: &[u32]) -> PendingG1 {
        if scalars.len() != idx.len() { panic!("Mismatch between the numbers of bases and scalars"); }
        let mut ticket: i32 = -1;
        unsafe {
            let (mut ds, mut di, mut dt) = (std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut());
            check(sys::tkm_crs_device_ptr(self.h, &mut dt, std::ptr::null_mut(), std::ptr::null_mut()));
            check(sys::tkm_dev_alloc(ctx(), scalars.len().max(1) * 32, &mut ds));
            check(sys::tkm_dev_alloc(ctx(), idx.len().max(1) * 4, &mut di));
            check(sys::tkm_memcpy_h2d(ctx(), ds, scalars.as_ptr() as *const _, scalars.len() * 32));
            check(sys::tkm_memcpy_h2d(ctx(), di, idx.as_ptr() as *const _, idx.len() * 4));
            let st = sys::tkm_msm_g1_indexed_begin(ctx(), ds, 0, dt, di, scalars.len(), &mut ticket);
            sys::tkm_dev_free(ctx(), ds);  // stream-ordered: released after the queued kernels have read them
            sys::tkm_dev_free(ctx(), di);
            check(st);
        }
        PendingG1 { ticket }
    }
}

/// A queued commitment or MSM (tkm_poly_commit_begin, tkm_msm_g1_begin, tkm_msm_g1_indexed_begin).
pub struct PendingG1 { ticket: i32 }
impl PendingG1 {
    pub fn get(self) -> G1serde {
        let mut out = G1serde::zero();
        check(unsafe { sys::tkm_commit_end(ctx(), self.ticket, out.0.as_mut_ptr()) });
        out
    }
}
