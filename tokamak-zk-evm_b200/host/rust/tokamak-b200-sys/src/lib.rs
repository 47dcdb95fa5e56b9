//! 1:1 declarations of include/tokamak_b200.h.  Every function returns a status code
//! (0 = OK); `tkm_last_error()` gives the message.  Byte formats are the reference's:
//! 32-byte little-endian canonical Fr, 96-byte x||y little-endian canonical affine G1.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

#[repr(C)] pub struct tkm_ctx { _p: [u8; 0] }
#[repr(C)] pub struct tkm_poly { _p: [u8; 0] }
#[repr(C)] pub struct tkm_crs { _p: [u8; 0] }

pub const TKM_FORWARD: i32 = 0;
pub const TKM_INVERSE: i32 = 1;
pub const TKM_COMM_ID_BYTES: usize = 128;
pub const TKM_OP_ADD: i32 = 0;
pub const TKM_OP_SUB: i32 = 1;
pub const TKM_OP_MUL: i32 = 2;
pub const TKM_OP_DIV: i32 = 3;
/// Opcodes of the postfix programs `tkm_polyexpr_eval` runs (word = opcode | operand << 8).
pub const TKM_PEX_LEAF: u32 = 0;
pub const TKM_PEX_CONST: u32 = 1;
pub const TKM_PEX_ADD: u32 = 2;
pub const TKM_PEX_SUB: u32 = 3;
pub const TKM_PEX_MUL: u32 = 4;
pub const TKM_PEX_SCALE: u32 = 5;
pub const TKM_PEX_XM1: u32 = 6;
pub const TKM_PEX_LEAF_SHIFT: u32 = 7;

extern "C" {
    pub fn tkm_last_error() -> *const c_char;
    pub fn tkm_version() -> *const c_char;
    pub fn tkm_ctx_create(device_ordinal: i32, out: *mut *mut tkm_ctx) -> i32;
    pub fn tkm_ctx_destroy(ctx: *mut tkm_ctx) -> i32;
    pub fn tkm_ctx_set_stream(ctx: *mut tkm_ctx, cuda_stream: *mut c_void) -> i32;
    pub fn tkm_ctx_sync(ctx: *mut tkm_ctx) -> i32;
    pub fn tkm_dev_alloc(ctx: *mut tkm_ctx, bytes: usize, out_dev: *mut *mut c_void) -> i32;
    pub fn tkm_dev_free(ctx: *mut tkm_ctx, dev: *mut c_void) -> i32;
    pub fn tkm_memcpy_h2d(ctx: *mut tkm_ctx, dev: *mut c_void, host: *const c_void, bytes: usize) -> i32;
    pub fn tkm_memcpy_d2h(ctx: *mut tkm_ctx, host: *mut c_void, dev: *const c_void, bytes: usize) -> i32;
    pub fn tkm_ntt_domain_init(ctx: *mut tkm_ctx, log2_size: u32) -> i32;
    pub fn tkm_ntt_domain_release(ctx: *mut tkm_ctx) -> i32;
    pub fn tkm_ntt_domain_log2(ctx: *mut tkm_ctx, out_log2: *mut i32) -> i32;
    pub fn tkm_root_of_unity(log2_n: u32, out32: *mut u8) -> i32;
    pub fn tkm_fr_to_mont(ctx: *mut tkm_ctx, dev_in: *const c_void, dev_out: *mut c_void, n: usize) -> i32;
    pub fn tkm_fr_from_mont(ctx: *mut tkm_ctx, dev_in: *const c_void, dev_out: *mut c_void, n: usize) -> i32;
    pub fn tkm_fr_vec_op(ctx: *mut tkm_ctx, op: i32, a: *const c_void, b: *const c_void, out: *mut c_void, n: usize) -> i32;
    pub fn tkm_fr_vec_scale(ctx: *mut tkm_ctx, s32: *const u8, a: *const c_void, out: *mut c_void, n: usize) -> i32;
    pub fn tkm_fr_vec_inv(ctx: *mut tkm_ctx, a: *const c_void, out: *mut c_void, n: usize) -> i32;
    pub fn tkm_fr_vec_fill(ctx: *mut tkm_ctx, s32: *const u8, dev_out: *mut c_void, n: usize) -> i32;
    pub fn tkm_fr_mul_x_minus_one(ctx: *mut tkm_ctx, dev_in: *const c_void, dev_out: *mut c_void, x_size: usize, y_size: usize) -> i32;
    pub fn tkm_fr_suffix_product(ctx: *mut tkm_ctx, dev_in: *const c_void, dev_out: *mut c_void, n: usize) -> i32;
    pub fn tkm_fr_vec_reduce(ctx: *mut tkm_ctx, op: i32, dev_a: *const c_void, dev_b: *const c_void, n: usize, out32: *mut u8) -> i32;
    pub fn tkm_fr_outer_product(ctx: *mut tkm_ctx, dev_col: *const c_void, dev_row: *const c_void, dev_out: *mut c_void, rows: usize, cols: usize) -> i32;
    pub fn tkm_fr_transpose(ctx: *mut tkm_ctx, dev_in: *const c_void, dev_out: *mut c_void, rows: usize, cols: usize) -> i32;
    pub fn tkm_fr_vec_op_host(ctx: *mut tkm_ctx, op: i32, a: *const u8, b: *const u8, out: *mut u8, n: usize) -> i32;
    pub fn tkm_bintt(ctx: *mut tkm_ctx, dev_in: *const c_void, dev_out: *mut c_void, x_size: usize, y_size: usize, dir: i32,
                     coset_x32: *const u8, coset_y32: *const u8) -> i32;
    pub fn tkm_bintt_host(ctx: *mut tkm_ctx, input: *const u8, out: *mut u8, x_size: usize, y_size: usize, dir: i32,
                          coset_x32: *const u8, coset_y32: *const u8) -> i32;
    pub fn tkm_ntt_batch(ctx: *mut tkm_ctx, dev_in: *const c_void, dev_out: *mut c_void, n: usize, batch: usize,
                         columns_batch: i32, dir: i32, coset32: *const u8) -> i32;
    pub fn tkm_msm_g1_host(ctx: *mut tkm_ctx, scalars: *const u8, bases: *const u8, n: usize, out96: *mut u8) -> i32;
    pub fn tkm_msm_g1(ctx: *mut tkm_ctx, dev_scalars: *const c_void, scalars_mont: i32, dev_bases_mont: *const c_void, n: usize,
                      out96: *mut u8) -> i32;
    pub fn tkm_g1_bases_to_mont(ctx: *mut tkm_ctx, dev_in: *const c_void, dev_out: *mut c_void, n: usize) -> i32;
    pub fn tkm_msm_g1_rect(ctx: *mut tkm_ctx, dev_scalars: *const c_void, scalars_mont: i32, scalar_row_stride: usize,
                           dev_bases_mont: *const c_void, base_row_stride: usize, rows: usize, cols: usize, out96: *mut u8) -> i32;
    pub fn tkm_msm_g1_indexed(ctx: *mut tkm_ctx, dev_scalars: *const c_void, scalars_mont: i32, dev_bases_mont: *const c_void,
                              dev_idx: *const c_void, n: usize, out96: *mut u8) -> i32;
    pub fn tkm_g1_fixed_base_mul(ctx: *mut tkm_ctx, base96: *const u8, dev_scalars: *const c_void, scalars_mont: i32, n: usize,
                                 dev_out_affine: *mut c_void) -> i32;
    pub fn tkm_g1_add(ctx: *mut tkm_ctx, a96: *const u8, b96: *const u8, out96: *mut u8) -> i32;
    pub fn tkm_g1_sum(ctx: *mut tkm_ctx, points96: *const u8, n: usize, out96: *mut u8) -> i32;
    pub fn tkm_g1_mul(ctx: *mut tkm_ctx, a96: *const u8, k32: *const u8, out96: *mut u8) -> i32;
    pub fn tkm_crs_upload(ctx: *mut tkm_ctx, points96: *const u8, rows: usize, cols: usize, out: *mut *mut tkm_crs) -> i32;
    pub fn tkm_crs_from_device(ctx: *mut tkm_ctx, dev_points: *mut c_void, rows: usize, cols: usize, take_ownership: i32,
                               out: *mut *mut tkm_crs) -> i32;
    pub fn tkm_crs_precompute(ctx: *mut tkm_ctx, crs: *mut tkm_crs, window_bits: u32) -> i32;
    pub fn tkm_crs_free(ctx: *mut tkm_ctx, crs: *mut tkm_crs) -> i32;
    pub fn tkm_crs_device_ptr(crs: *mut tkm_crs, out_dev: *mut *mut c_void, rows: *mut usize, cols: *mut usize) -> i32;
    pub fn tkm_poly_from_coeffs_host(ctx: *mut tkm_ctx, coeffs: *const u8, x_size: usize, y_size: usize, out: *mut *mut tkm_poly) -> i32;
    pub fn tkm_poly_from_evals_host(ctx: *mut tkm_ctx, evals: *const u8, x_size: usize, y_size: usize, coset_x32: *const u8,
                                    coset_y32: *const u8, out: *mut *mut tkm_poly) -> i32;
    pub fn tkm_poly_from_device(ctx: *mut tkm_ctx, dev_coeffs: *const c_void, x_size: usize, y_size: usize, out: *mut *mut tkm_poly) -> i32;
    pub fn tkm_poly_zero(ctx: *mut tkm_ctx, x_size: usize, y_size: usize, out: *mut *mut tkm_poly) -> i32;
    pub fn tkm_poly_clone(ctx: *mut tkm_ctx, p: *const tkm_poly, out: *mut *mut tkm_poly) -> i32;
    pub fn tkm_poly_free(ctx: *mut tkm_ctx, p: *mut tkm_poly) -> i32;
    pub fn tkm_poly_shape(p: *const tkm_poly, x_size: *mut usize, y_size: *mut usize) -> i32;
    pub fn tkm_poly_device_ptr(p: *mut tkm_poly, out_dev: *mut *mut c_void) -> i32;
    pub fn tkm_poly_copy_coeffs_host(ctx: *mut tkm_ctx, p: *const tkm_poly, out: *mut u8) -> i32;
    pub fn tkm_poly_to_evals_host(ctx: *mut tkm_ctx, p: *const tkm_poly, coset_x32: *const u8, coset_y32: *const u8, out: *mut u8) -> i32;
    pub fn tkm_poly_ntt_inplace(ctx: *mut tkm_ctx, p: *mut tkm_poly, dir: i32, coset_x32: *const u8, coset_y32: *const u8) -> i32;
    pub fn tkm_poly_find_degree(ctx: *mut tkm_ctx, p: *const tkm_poly, x_degree: *mut i64, y_degree: *mut i64) -> i32;
    pub fn tkm_poly_resize(ctx: *mut tkm_ctx, p: *mut tkm_poly, target_x: usize, target_y: usize) -> i32;
    pub fn tkm_poly_optimize_size(ctx: *mut tkm_ctx, p: *mut tkm_poly) -> i32;
    pub fn tkm_poly_mul_monomial(ctx: *mut tkm_ctx, p: *const tkm_poly, ex: usize, ey: usize, out: *mut *mut tkm_poly) -> i32;
    pub fn tkm_poly_axpby(ctx: *mut tkm_ctx, a: *const tkm_poly, ca32: *const u8, b: *const tkm_poly, cb32: *const u8,
                          out: *mut *mut tkm_poly) -> i32;
    pub fn tkm_poly_add_scalar(ctx: *mut tkm_ctx, p: *mut tkm_poly, s32: *const u8) -> i32;
    pub fn tkm_poly_mul(ctx: *mut tkm_ctx, a: *const tkm_poly, b: *const tkm_poly, out: *mut *mut tkm_poly) -> i32;
    pub fn tkm_poly_scale_coeffs(ctx: *mut tkm_ctx, p: *const tkm_poly, sx32: *const u8, sy32: *const u8, out: *mut *mut tkm_poly) -> i32;
    pub fn tkm_poly_eval(ctx: *mut tkm_ctx, p: *const tkm_poly, x32: *const u8, y32: *const u8, out32: *mut u8) -> i32;
    pub fn tkm_poly_eval_x(ctx: *mut tkm_ctx, p: *const tkm_poly, x32: *const u8, out: *mut *mut tkm_poly) -> i32;
    pub fn tkm_poly_eval_y(ctx: *mut tkm_ctx, p: *const tkm_poly, y32: *const u8, out: *mut *mut tkm_poly) -> i32;
    pub fn tkm_poly_div_by_vanishing(ctx: *mut tkm_ctx, p: *mut tkm_poly, c: usize, d: usize, out_qx: *mut *mut tkm_poly,
                                     out_qy: *mut *mut tkm_poly) -> i32;
    pub fn tkm_poly_div_by_ruffini(ctx: *mut tkm_ctx, p: *const tkm_poly, x32: *const u8, y32: *const u8, out_qx: *mut *mut tkm_poly,
                                   out_qy: *mut *mut tkm_poly, out_r32: *mut u8) -> i32;
    pub fn tkm_poly_commit(ctx: *mut tkm_ctx, p: *mut tkm_poly, crs: *const tkm_crs, out96: *mut u8) -> i32;
    pub fn tkm_poly_divide_uni(ctx: *mut tkm_ctx, p: *const tkm_poly, denom: *const tkm_poly, y_dir: i32, out_q: *mut *mut tkm_poly,
                               out_r: *mut *mut tkm_poly) -> i32;
    pub fn tkm_poly_commit_begin(ctx: *mut tkm_ctx, p: *mut tkm_poly, crs: *const tkm_crs, out_ticket: *mut i32) -> i32;
    pub fn tkm_commit_end(ctx: *mut tkm_ctx, ticket: i32, out96: *mut u8) -> i32;
    pub fn tkm_msm_g1_begin(ctx: *mut tkm_ctx, dev_scalars: *const c_void, scalars_mont: i32, dev_bases_mont: *const c_void, n: usize,
                            out_ticket: *mut i32) -> i32;
    pub fn tkm_msm_g1_indexed_begin(ctx: *mut tkm_ctx, dev_scalars: *const c_void, scalars_mont: i32, dev_bases_mont: *const c_void,
                                    dev_idx: *const c_void, n: usize, out_ticket: *mut i32) -> i32;
    pub fn tkm_r1cs_uvw_polys(ctx: *mut tkm_ctx, s_d: u32, n_rows: *const u32, rp_base: *const u64, row_ptr: *const u32, row_ptr_len: usize,
                              wire: *const u32, coeff32: *const u8, nnz: usize, sub_of_col: *const u32, var_off: *const u64,
                              witness32: *const u8, n_vars: usize, n: usize, s_max: usize, out_u: *mut *mut tkm_poly,
                              out_v: *mut *mut tkm_poly, out_w: *mut *mut tkm_poly) -> i32;
    pub fn tkm_ntt_batch_scatter(ctx: *mut tkm_ctx, dev_in: *const c_void, n: usize, batch: usize, columns_batch: i32, dir: i32,
                                 coset32: *const u8, peer_out: *const *mut c_void, n_peers: u32, stride_a: u64, stride_b: u64, b0: u64) -> i32;
    pub fn tkm_g1_bases_from_mont(ctx: *mut tkm_ctx, dev_in: *const c_void, dev_out: *mut c_void, n: usize) -> i32;
    pub fn tkm_kernel_time_last(ctx: *mut tkm_ctx, out_ms: *mut f32) -> i32;
    pub fn tkm_host_parse_hex_scalars(text: *const c_char, len: usize, out32: *mut u8, capacity: usize, out_count: *mut usize) -> i32;
    pub fn tkm_host_parse_r1cs(data: *const u8, len: usize, n_wires: *mut u32, n_constraints: *mut u32, nnz: *mut usize, row_ptr: *mut u32,
                               wire: *mut u32, coeff32: *mut u8) -> i32;
    pub fn tkm_event_time_begin(ctx: *mut tkm_ctx) -> i32;
    pub fn tkm_event_time_end(ctx: *mut tkm_ctx, out_ms: *mut f32) -> i32;
    pub fn tkm_launch_count(ctx: *mut tkm_ctx, out: *mut u64) -> i32;
    pub fn tkm_host_keccak256(data: *const u8, len: usize, out32: *mut u8) -> i32;
    pub fn tkm_microbench(ctx: *mut tkm_ctx, kind: i32, out_ops_per_s: *mut f64) -> i32;
    pub fn tkm_poly_lincomb(ctx: *mut tkm_ctx, k: u32, polys: *const *const tkm_poly, coeffs32: *const u8, shift_x: *const u32,
                            shift_y: *const u32, out: *mut *mut tkm_poly) -> i32;
    pub fn tkm_polyexpr_eval(ctx: *mut tkm_ctx, leaves: *const *const tkm_poly, n_leaves: u32, program: *const u32, n_ops: u32,
                             consts32: *const u8, n_consts: u32, target_x: usize, target_y: usize, out: *mut *mut tkm_poly) -> i32;
    pub fn tkm_poly_kernel_time_last(ctx: *mut tkm_ctx, out_ms: *mut f32) -> i32;
    pub fn tkm_crs_upload_mont(ctx: *mut tkm_ctx, points96_mont: *const u8, rows: usize, cols: usize, out: *mut *mut tkm_crs) -> i32;
    pub fn tkm_msm_tree_stats(ctx: *mut tkm_ctx, out_levels: *mut u32, out_counts: *mut u64) -> i32;
    pub fn tkm_comm_unique_id(out_id: *mut u8) -> i32;
    pub fn tkm_comm_init(ctx: *mut tkm_ctx, id: *const u8, rank: i32, world: i32) -> i32;
    pub fn tkm_comm_destroy(ctx: *mut tkm_ctx) -> i32;
    pub fn tkm_comm_rank(ctx: *mut tkm_ctx, out_rank: *mut i32, out_world: *mut i32) -> i32;
    pub fn tkm_msm_g1_sharded(ctx: *mut tkm_ctx, dev_scalars: *const c_void, scalars_mont: i32, dev_bases_mont: *const c_void, n_local: usize,
                              out96: *mut u8) -> i32;
    pub fn tkm_bintt_sharded(ctx: *mut tkm_ctx, dev_in: *const c_void, dev_out: *mut c_void, x_size: usize, y_size: usize, dir: i32,
                             coset_x32: *const u8, coset_y32: *const u8) -> i32;
    pub fn tkm_fr_powers(ctx: *mut tkm_ctx, base32: *const u8, dev_out: *mut c_void, n: usize) -> i32;
    pub fn tkm_fr_gather(ctx: *mut tkm_ctx, dev_table: *const c_void, table_len: usize, dev_idx: *const c_void, n: usize, dev_out: *mut c_void) -> i32;
    pub fn tkm_fr_scatter_from_table(ctx: *mut tkm_ctx, dev_dst: *mut c_void, dst_len: usize, dev_dst_idx: *const c_void, dev_table: *const c_void,
                                     table_len: usize, dev_src_idx: *const c_void, n: usize) -> i32;
}
