//! vector_operations (libs/src/vector_operations/mod.rs:19-141,639-693): the host-slice helpers the prover imports
//! (prove/src/lib.rs:16-17).  Element-wise products and quotients run on the device (batched inversion for the division).
use crate::{check, ctx, ScalarField};
use tokamak_b200_sys as sys;

fn pointwise(op: i32, lhs: &[ScalarField], rhs: &[ScalarField], res: &mut [ScalarField]) {
    if lhs.len() != rhs.len() || lhs.len() != res.len() { panic!("Mismatch of sizes of vectors to be pointwise operated"); }
    if lhs.is_empty() { return; }
    check(unsafe { sys::tkm_fr_vec_op_host(ctx(), op, lhs.as_ptr() as *const u8, rhs.as_ptr() as *const u8, res.as_mut_ptr() as *mut u8, lhs.len()) });
}
pub fn point_mul_two_vecs(lhs: &[ScalarField], rhs: &[ScalarField], res: &mut [ScalarField]) { pointwise(sys::TKM_OP_MUL, lhs, rhs, res) }
pub fn point_div_two_vecs(lhs: &[ScalarField], rhs: &[ScalarField], res: &mut [ScalarField]) { pointwise(sys::TKM_OP_DIV, lhs, rhs, res) }
pub fn point_add_two_vecs(lhs: &[ScalarField], rhs: &[ScalarField], res: &mut [ScalarField]) { pointwise(sys::TKM_OP_ADD, lhs, rhs, res) }
pub fn point_sub_two_vecs(lhs: &[ScalarField], rhs: &[ScalarField], res: &mut [ScalarField]) { pointwise(sys::TKM_OP_SUB, lhs, rhs, res) }

/// transpose_inplace (:139-141): row_size x col_size row-major -> col_size x row_size
pub fn transpose_inplace(a: &mut [ScalarField], row_size: usize, col_size: usize) {
    if a.len() != row_size * col_size { panic!("Error in transpose"); }
    let src = a.to_vec();
    for i in 0..row_size {
        for j in 0..col_size { a[j * row_size + i] = src[i * col_size + j]; }
    }
}

/// resize (:639-672): copy the overlapping rectangle of a curr_row x curr_col matrix into target_row x target_col
pub fn resize(mat: &[ScalarField], curr_row: usize, curr_col: usize, target_row: usize, target_col: usize, zero: ScalarField) -> Vec<ScalarField> {
    let mut out = vec![zero; target_row * target_col];
    for i in 0..curr_row.min(target_row) {
        for j in 0..curr_col.min(target_col) { out[i * target_col + j] = mat[i * curr_col + j]; }
    }
    out
}

/// scale_vec (:90-100): res = scaler * vec
pub fn scale_vec(scaler: ScalarField, vec: &[ScalarField], res: &mut [ScalarField]) {
    for (r, v) in res.iter_mut().zip(vec.iter()) { *r = *v * scaler; }
}
