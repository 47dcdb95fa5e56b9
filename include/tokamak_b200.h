/* libtokamak_b200 -- C-ABI of the B200-native proving core for Tokamak zk-EVM.
 *
 * Drop-in boundary for the data-parallel hot path of packages/backend (reference paths below are
 * relative to /root/reference/packages/backend/): the BLS12-381 G1 MSM behind every commitment
 * and the bivariate polynomial engine of libs (DensePolynomialExt).  Each entry point names the
 * reference interface it replaces.  A thin Rust `-sys` crate binds these symbols 1:1
 * (see INTEGRATION.md); tests bind them with ctypes.
 *
 * Conventions
 *  - every function returns int32_t status: 0 = OK, negative = tkm_status; tkm_last_error()
 *    returns a thread-local message.  No C++ exception or abort crosses this boundary
 *    (the reference panics/unwraps eIcicleError; the Rust shim turns a non-zero status into panic!).
 *  - byte formats at the boundary are the reference's: Fr = 32-byte little-endian canonical,
 *    G1 affine = x||y, 2 x 48-byte little-endian canonical, all-zero = identity
 *    (libs/src/iotools/mod.rs:1786-1815,1969-1980; group_structures/mod.rs:889-893).
 *  - "host" pointers are ordinary host memory; "dev" pointers are CUDA device pointers owned by
 *    the caller (or by an opaque handle).  Device-resident field data is in Montgomery form.
 *  - a context is bound to one device and one stream and is not thread-safe (the reference has
 *    the same rule: .cargo/config.toml:5 RUST_TEST_THREADS=1).
 *  - the CUDA kernels are the only implementation: there is no CPU fallback.
 */
#ifndef TOKAMAK_B200_H
#define TOKAMAK_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  TKM_OK = 0,
  TKM_ERR_INVALID_ARGUMENT = -1, /* eIcicleError::InvalidArgument */
  TKM_ERR_ALLOCATION = -2,       /* eIcicleError::AllocationFailed / OutOfMemory */
  TKM_ERR_DOMAIN = -3,           /* NTT domain missing or too small (bivariate_polynomial/mod.rs:1437-1445) */
  TKM_ERR_CUDA = -4,             /* any CUDA runtime failure */
  TKM_ERR_NO_DEVICE = -5,        /* no usable CUDA device: there is deliberately no CPU fallback */
  TKM_ERR_INTERNAL = -6
} tkm_status;

typedef struct tkm_ctx tkm_ctx;   /* device + stream + NTT domain + scratch */
typedef struct tkm_poly tkm_poly; /* device-resident bivariate polynomial (DensePolynomialExt) */
typedef struct tkm_crs tkm_crs;   /* device-resident G1 base table (sigma_1.xy_powers or a sparse table) */

enum { TKM_FORWARD = 0, TKM_INVERSE = 1 };                  /* NTTDir::kForward / kInverse */
enum { TKM_OP_ADD = 0, TKM_OP_SUB = 1, TKM_OP_MUL = 2, TKM_OP_DIV = 3 }; /* VecOps::add/sub/mul/div */

const char *tkm_last_error(void);
const char *tkm_version(void);

/* ---- context: replaces utils::check_device (libs/src/utils/mod.rs:88-110) ------------------ */
int32_t tkm_ctx_create(int32_t device_ordinal, tkm_ctx **out);
int32_t tkm_ctx_destroy(tkm_ctx *ctx);
/* Use a caller-owned cudaStream_t (e.g. torch's current stream); NULL restores the context's own. */
int32_t tkm_ctx_set_stream(tkm_ctx *ctx, void *cuda_stream);
int32_t tkm_ctx_sync(tkm_ctx *ctx);

/* ---- raw device memory (DeviceVec::device_malloc / copy_from_host / copy_to_host) ----------- */
int32_t tkm_dev_alloc(tkm_ctx *ctx, size_t bytes, void **out_dev);
int32_t tkm_dev_free(tkm_ctx *ctx, void *dev);
int32_t tkm_memcpy_h2d(tkm_ctx *ctx, void *dev, const void *host, size_t bytes);
int32_t tkm_memcpy_d2h(tkm_ctx *ctx, void *host, const void *dev, size_t bytes);

/* ---- NTT domain: init_ntt_domain_for_size / ntt::initialize_domain / release_domain /
 *      get_root_of_unity (libs/src/bivariate_polynomial/mod.rs:33-55, :505) --------------------- */
int32_t tkm_ntt_domain_init(tkm_ctx *ctx, uint32_t log2_size);
int32_t tkm_ntt_domain_release(tkm_ctx *ctx);
int32_t tkm_ntt_domain_log2(tkm_ctx *ctx, int32_t *out_log2); /* -1 when not initialised */
int32_t tkm_root_of_unity(uint32_t log2_n, uint8_t out32[32]);

/* ---- Fr vectors on the device (Montgomery form) ------------------------------------------------ */
/* canonical <-> Montgomery, in place allowed.  (ICICLE hides this inside ScalarField.) */
int32_t tkm_fr_to_mont(tkm_ctx *ctx, const void *dev_in, void *dev_out, size_t n);
int32_t tkm_fr_from_mont(tkm_ctx *ctx, const void *dev_in, void *dev_out, size_t n);
/* VecOps::{add,sub,mul,div} (libs/src/vector_operations/mod.rs:19-141); div uses inv(0)=0. */
int32_t tkm_fr_vec_op(tkm_ctx *ctx, int32_t op, const void *dev_a, const void *dev_b, void *dev_out, size_t n);
/* VecOps::scalar_mul: out = s * a, s = 32-byte canonical host scalar. */
int32_t tkm_fr_vec_scale(tkm_ctx *ctx, const uint8_t s32[32], const void *dev_a, void *dev_out, size_t n);
/* VecOps::inv, batched (Montgomery trick in-kernel), inv(0)=0 (bivariate_polynomial/mod.rs:2180). */
int32_t tkm_fr_vec_inv(tkm_ctx *ctx, const void *dev_a, void *dev_out, size_t n);
/* device_vec_from_scalar (bivariate_polynomial/mod.rs:452-457): out[k] = s for k < n. */
int32_t tkm_fr_vec_fill(tkm_ctx *ctx, const uint8_t s32[32], void *dev_out, size_t n);
/* Pointwise product with the evaluations of (X - 1) on the x_size-th roots of unity, i.e.
 * out[i*y_size + j] = in[i*y_size + j] * (omega_x^i - 1): PolyExpr::MulXMinusOne in the evaluation domain
 * (x_minus_one_evals, bivariate_polynomial/mod.rs:504-518) without materialising the factor matrix. */
int32_t tkm_fr_mul_x_minus_one(tkm_ctx *ctx, const void *dev_in, void *dev_out, size_t x_size, size_t y_size);
/* Exclusive suffix product out[i] = prod_{k>i} in[k], out[n-1] = 1: the recursion-polynomial scan of prove1
 * (prove/src/lib.rs:1858-1867, a serial 2^20-step loop in the reference).  in may equal out. */
int32_t tkm_fr_suffix_product(tkm_ctx *ctx, const void *dev_in, void *dev_out, size_t n);
/* Reductions to one host scalar (32-byte canonical): op 0 = VecOps::sum (vector_operations/mod.rs:124,336),
 * op 1 = VecOps::product (prove/src/lib.rs:1005-1016), op 2 = inner_product_two_vecs sum_k a[k]*b[k]
 * (vector_operations/mod.rs:100-141; dev_b ignored for ops 0/1). */
int32_t tkm_fr_vec_reduce(tkm_ctx *ctx, int32_t op, const void *dev_a, const void *dev_b, size_t n, uint8_t out32[32]);
/* outer_product_two_vecs (vector_operations/mod.rs:551-600): out[i*cols + j] = col[i] * row[j]. */
int32_t tkm_fr_outer_product(tkm_ctx *ctx, const void *dev_col, const void *dev_row, void *dev_out, size_t rows, size_t cols);
/* out[k] = base^k, k < n, in Montgomery form on the device (the power vectors behind scale_coeffs, the vanishing and
 * permutation tables; resize_monomial_vec / extend_monomial_vec, vector_operations/mod.rs:674-693). */
int32_t tkm_fr_powers(tkm_ctx *ctx, const uint8_t base32[32], void *dev_out, size_t n);
/* tkm_fr_gather: out[k] = table[idx[k]], k < n (u32 indices on the device, range-checked) -- the scalar vector of a
 * sparse-gather MSM picked out of the device-resident placement variables (encode_statement_common collects
 * placement_variables[i].variables[j] the same way on the host, group_structures/mod.rs:266-300).
 * tkm_fr_scatter_from_table:
 * dst[dst_idx[k]] = table[src_idx[k]] for k < n (u32 indices on the device, range-checked): the sparse overrides of the
 * permutation evaluation tables s0, s1 (Permutation::to_poly, libs/src/iotools/mod.rs:419-455) without a host round trip. */
int32_t tkm_fr_gather(tkm_ctx *ctx, const void *dev_table, size_t table_len, const void *dev_idx, size_t n, void *dev_out);
int32_t tkm_fr_scatter_from_table(tkm_ctx *ctx, void *dev_dst, size_t dst_len, const void *dev_dst_idx, const void *dev_table, size_t table_len,
                                  const void *dev_src_idx, size_t n);
/* VecOps::transpose (vector_operations/mod.rs:139,168): rows x cols -> cols x rows, out != in. */
int32_t tkm_fr_transpose(tkm_ctx *ctx, const void *dev_in, void *dev_out, size_t rows, size_t cols);
/* Host-buffer forms of the same ops (HostSlice in, HostSlice out; canonical bytes). */
int32_t tkm_fr_vec_op_host(tkm_ctx *ctx, int32_t op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n);

/* ---- bivariate NTT: DensePolynomialExt::_biNTT (libs/src/bivariate_polynomial/mod.rs:1422-1478)
 * Row-major, X = row, Y = contiguous column; natural order in and out; inverse includes 1/(x*y);
 * coset_x / coset_y: NULL (= 1) or a 32-byte canonical host scalar per axis; forward scales
 * coefficient (i,j) by gx^i gy^j first, inverse scales by gx^-i gy^-j last (libs/src/tests.rs:134-180).
 * dev_in may equal dev_out. */
int32_t tkm_bintt(tkm_ctx *ctx, const void *dev_in, void *dev_out, size_t x_size, size_t y_size, int32_t dir,
                  const uint8_t *coset_x32, const uint8_t *coset_y32);
/* Same with host buffers in canonical form (the HostSlice path of ntt::ntt): H2D, transform, D2H. */
int32_t tkm_bintt_host(tkm_ctx *ctx, const uint8_t *in, uint8_t *out, size_t x_size, size_t y_size, int32_t dir,
                       const uint8_t *coset_x32, const uint8_t *coset_y32);
/* ntt::ntt with NTTConfig{batch_size, columns_batch} (used by tests.rs:519-588): `batch` vectors of
 * length n, row batch (element j of vector b at b*n+j) or column batch (at j*batch+b). */
int32_t tkm_ntt_batch(tkm_ctx *ctx, const void *dev_in, void *dev_out, size_t n, size_t batch, int32_t columns_batch,
                      int32_t dir, const uint8_t *coset32);

/* ---- G1 MSM: msm::msm with MSMConfig::default() (libs/src/iotools/mod.rs:2093-2099;
 *      group_structures/mod.rs:108-114,135-141) ---------------------------------------------------- */
/* Host scalars (n x 32 B canonical), host affine bases (n x 96 B canonical) -> one affine point (96 B).
 * From 2^19 points up the host-to-device copies are pipelined with the accumulation (pinned buffers make them asynchronous):
 * the point range is cut into pieces, scalars travel before bases, and a piece's digit decomposition and sort run while its
 * bases are still in flight; the piece layout follows the share of the previous call its copies took. */
int32_t tkm_msm_g1_host(tkm_ctx *ctx, const uint8_t *scalars, const uint8_t *bases, size_t n, uint8_t out96[96]);
/* Device-resident: scalars n x 8 u32 (Montgomery if scalars_mont != 0, else canonical),
 * bases = device table in Montgomery form (from tkm_g1_bases_to_mont or a tkm_crs). */
int32_t tkm_msm_g1(tkm_ctx *ctx, const void *dev_scalars, int32_t scalars_mont, const void *dev_bases_mont, size_t n,
                   uint8_t out96[96]);
/* tkm_ntt_batch fused with the multi-GPU re-sharding exchange: the last pass stores straight into the
 * peers' buffers over NVLink instead of writing a local result, so the X<->Y transpose of a row-sharded bivariate NTT
 * needs no separate transpose kernel and no separate all-to-all.  The element at axis position a of batch lane b is
 * written to peer_out[a / (n / n_peers)] + ((a mod (n / n_peers)) * stride_a + (b + b0) * stride_b) elements.
 * peer_out[] (host array) holds n_peers peer-mapped device pointers (a power of two <= 16); the caller orders the
 * launch against the peers with its own barriers. */
int32_t tkm_ntt_batch_scatter(tkm_ctx *ctx, const void *dev_in, size_t n, size_t batch, int32_t columns_batch, int32_t dir,
                              const uint8_t *coset32, void *const *peer_out, uint32_t n_peers, uint64_t stride_a, uint64_t stride_b,
                              uint64_t b0);
/* Canonical affine bytes on the device -> Montgomery-form base table (in place allowed). */
int32_t tkm_g1_bases_to_mont(tkm_ctx *ctx, const void *dev_in, void *dev_out, size_t n);
/* The inverse conversion (device Montgomery affine -> canonical), e.g. to write a generated CRS table out. */
int32_t tkm_g1_bases_from_mont(tkm_ctx *ctx, const void *dev_in, void *dev_out, size_t n);
/* Strided rectangle (the trimmed (deg_x+1) x (deg_y+1) rectangle encode_poly commits,
 * iotools/mod.rs:2061-2088): scalar (i,j) at dev_scalars[i*scalar_row_stride + j],
 * base (i,j) at bases[i*base_row_stride + j], i < rows, j < cols. */
int32_t tkm_msm_g1_rect(tkm_ctx *ctx, const void *dev_scalars, int32_t scalars_mont, size_t scalar_row_stride,
                        const void *dev_bases_mont, size_t base_row_stride, size_t rows, size_t cols,
                        uint8_t out96[96]);
/* Sparse-gather MSM (msm_g1_bases over gathered CRS rows, group_structures/mod.rs:127-300):
 * result = sum_k scalars[k] * table[idx[k]].  dev_idx: n x u32. */
int32_t tkm_msm_g1_indexed(tkm_ctx *ctx, const void *dev_scalars, int32_t scalars_mont, const void *dev_bases_mont,
                           const void *dev_idx, size_t n, uint8_t out96[96]);
/* N independent scalar multiples of one base: msm::msm with batch_size = N over a generator vector
 * (from_coef_vec_to_g1serde_vec, iotools/mod.rs:1113-1135).  out: n x 96 B canonical on the device. */
int32_t tkm_g1_fixed_base_mul(tkm_ctx *ctx, const uint8_t base96[96], const void *dev_scalars, int32_t scalars_mont,
                              size_t n, void *dev_out_affine);
/* G1serde ops (group_structures/mod.rs:895-947): out = a + b ; out = k * a.  Host bytes. */
int32_t tkm_g1_add(tkm_ctx *ctx, const uint8_t a96[96], const uint8_t b96[96], uint8_t out96[96]);
int32_t tkm_g1_mul(tkm_ctx *ctx, const uint8_t a96[96], const uint8_t k32[32], uint8_t out96[96]);
/* Sum of n affine points (host bytes): combines the per-GPU partial sums of a point-range-sharded MSM in one launch. */
int32_t tkm_g1_sum(tkm_ctx *ctx, const uint8_t *points96, size_t n, uint8_t out96[96]);

/* ---- CRS tables: Sigma1.xy_powers and friends (libs/src/group_structures/mod.rs:361-394;
 *      archived form iotools/mod.rs:1701-1783) ------------------------------------------------------- */
/* Upload rows*cols canonical affine points (row-major, index cols*h + i <-> x^h y^i) once. */
int32_t tkm_crs_upload(tkm_ctx *ctx, const uint8_t *points96, size_t rows, size_t cols, tkm_crs **out);
/* The same for points that are already in the device layout: x || y with 48-byte little-endian MONTGOMERY coordinates
 * (R = 2^384) -- the `ffjs-g1-affine-96` encoding of the reference's flat TZBWASM1 CRS container
 * (packages/backend-wasm/src/artifacts/binary/binary-format.ts:24-31, specs/prover-crs.v1.json), so a section of a
 * memory-mapped prover_crs file is uploaded with one copy and no conversion. */
int32_t tkm_crs_upload_mont(tkm_ctx *ctx, const uint8_t *points96_mont, size_t rows, size_t cols, tkm_crs **out);
/* Wrap points already on the device in canonical form (converted in place to Montgomery). */
int32_t tkm_crs_from_device(tkm_ctx *ctx, void *dev_points, size_t rows, size_t cols, int32_t take_ownership,
                            tkm_crs **out);
/* Optional, for provers that keep one CRS across many proofs: build fixed-base tables 2^(c*w) * P for every CRS point
 * (W = ceil(256/c) tables, e.g. 13 x 384 MiB at c = 20).  tkm_poly_commit then folds all digit windows into one shared
 * bucket set: fewer additions per point, no per-window reduction, no Horner tail.  Results are identical. */
int32_t tkm_crs_precompute(tkm_ctx *ctx, tkm_crs *crs, uint32_t window_bits);
int32_t tkm_crs_free(tkm_ctx *ctx, tkm_crs *crs);
int32_t tkm_crs_device_ptr(tkm_crs *crs, void **out_dev, size_t *rows, size_t *cols);

/* ---- device-resident bivariate polynomial: DensePolynomialExt
 *      (libs/src/bivariate_polynomial/mod.rs:112-118 and the BivariatePolynomial trait :1283-1416) -- */
/* from_coeffs (:1527-1551): sizes must be powers of two; coeffs canonical host bytes. */
int32_t tkm_poly_from_coeffs_host(tkm_ctx *ctx, const uint8_t *coeffs, size_t x_size, size_t y_size, tkm_poly **out);
/* from_rou_evals (:1615-1644). */
int32_t tkm_poly_from_evals_host(tkm_ctx *ctx, const uint8_t *evals, size_t x_size, size_t y_size,
                                 const uint8_t *coset_x32, const uint8_t *coset_y32, tkm_poly **out);
/* from_coeffs with a DeviceSlice (:1527-1551): copies x_size*y_size Montgomery-form elements from dev_coeffs. */
int32_t tkm_poly_from_device(tkm_ctx *ctx, const void *dev_coeffs, size_t x_size, size_t y_size, tkm_poly **out);
/* read_R1CS_gen_uvwXY (libs/src/iotools/mod.rs:1287-1420): the sparse R1CS x witness products of every placement, as
 * the three witness polynomials u, v, w of shape n x s_max (evaluation tables [row][placement] built on the device, then
 * one inverse biNTT each).  All pointers are HOST arrays: the library's constraints as concatenated CSR (row_ptr /
 * wire / coeff, rp_base[3*s + m] = start of (subcircuit s, matrix m)'s row pointers, n_rows[s] constraints), the
 * placement list (sub_of_col[c] = subcircuit of column c or 0xffffffff, var_off[c] = offset of its variables in
 * witness32) and the variables as 32-byte canonical scalars. */
int32_t tkm_r1cs_uvw_polys(tkm_ctx *ctx, uint32_t s_D, const uint32_t *n_rows, const uint64_t *rp_base, const uint32_t *row_ptr,
                           size_t row_ptr_len, const uint32_t *wire, const uint8_t *coeff32, size_t nnz, const uint32_t *sub_of_col,
                           const uint64_t *var_off, const uint8_t *witness32, size_t n_vars, size_t n, size_t s_max, tkm_poly **out_u,
                           tkm_poly **out_v, tkm_poly **out_w);
int32_t tkm_poly_zero(tkm_ctx *ctx, size_t x_size, size_t y_size, tkm_poly **out);
int32_t tkm_poly_clone(tkm_ctx *ctx, const tkm_poly *p, tkm_poly **out); /* Clone (:520-530) */
int32_t tkm_poly_free(tkm_ctx *ctx, tkm_poly *p);
int32_t tkm_poly_shape(const tkm_poly *p, size_t *x_size, size_t *y_size);
int32_t tkm_poly_device_ptr(tkm_poly *p, void **out_dev); /* Montgomery form, row-major */
/* copy_coeffs (:1676-1682) / to_rou_evals (:1646-1674) into canonical host bytes. */
int32_t tkm_poly_copy_coeffs_host(tkm_ctx *ctx, const tkm_poly *p, uint8_t *out);
int32_t tkm_poly_to_evals_host(tkm_ctx *ctx, const tkm_poly *p, const uint8_t *coset_x32, const uint8_t *coset_y32,
                               uint8_t *out);
/* In-place transforms between coefficient and evaluation form on the device buffer. */
int32_t tkm_poly_ntt_inplace(tkm_ctx *ctx, tkm_poly *p, int32_t dir, const uint8_t *coset_x32,
                             const uint8_t *coset_y32);
/* find_degree (:1480-1515): (-1,-1) for the zero polynomial. */
int32_t tkm_poly_find_degree(tkm_ctx *ctx, const tkm_poly *p, int64_t *x_degree, int64_t *y_degree);
/* resize (:1784-1806): crop / zero-pad to the next powers of two of (target_x, target_y). */
int32_t tkm_poly_resize(tkm_ctx *ctx, tkm_poly *p, size_t target_x, size_t target_y);
/* optimize_size (:1808-1818). */
int32_t tkm_poly_optimize_size(tkm_ctx *ctx, tkm_poly *p);
/* mul_monomial (:1820-1844): multiply by X^ex Y^ey (new polynomial). */
int32_t tkm_poly_mul_monomial(tkm_ctx *ctx, const tkm_poly *p, size_t ex, size_t ey, tkm_poly **out);
/* poly_comb! (prove/src/lib.rs:30-38) and the helper polynomials built from it (:48-124) in ONE pass:
 * out = sum_{t<k} c_t * X^shift_x[t] * Y^shift_y[t] * polys[t] on the union shape (a shifted term has mul_monomial's
 * shape, bivariate_polynomial/mod.rs:1820-1844).  coeffs32: k canonical 32-byte scalars or NULL (all 1);
 * shift_x / shift_y: k monomial exponents or NULL (no shift).  k <= 16. */
int32_t tkm_poly_lincomb(tkm_ctx *ctx, uint32_t k, const tkm_poly *const *polys, const uint8_t *coeffs32, const uint32_t *shift_x,
                         const uint32_t *shift_y, tkm_poly **out);
/* PolyExpr::evaluate_fused_with_domain (bivariate_polynomial/mod.rs:227-260, evaluate_on_domain :311-435): the
 * expression is handed over as a postfix program over `leaves` (the distinct polynomials of the DAG: one forward
 * biNTT each, the reference's pointer-keyed leaf cache :459-502) and `consts32` (canonical scalars); the whole
 * pointwise DAG runs as one kernel between the leaf transforms and the single inverse biNTT.  Program words are
 * opcode | operand << 8.  Errors mirror the reference's panics ("Fused polynomial expression domains must be powers
 * of two.", "... domain is too small for the expression degree.").  Limits: 16 leaves, 96 ops, 24 constants, depth 8. */
enum {
  TKM_PEX_LEAF = 0,  /* push the evaluations of leaves[operand]                    (PolyExpr::Poly)          */
  TKM_PEX_CONST = 1, /* push the constant consts32[operand]                         (PolyExpr::Scalar)        */
  TKM_PEX_ADD = 2,   /* b = pop, a = pop, push a + b                                (Add, Sum)                */
  TKM_PEX_SUB = 3,   /* push a - b                                                  (Sub)                     */
  TKM_PEX_MUL = 4,   /* push a * b                                                  (Mul)                     */
  TKM_PEX_SCALE = 5, /* top *= consts32[operand]                                    (Scale)                   */
  TKM_PEX_XM1 = 6,   /* top *= (omega_x^i - 1), the evaluations of X - 1            (MulXMinusOne, :504-518)  */
  /* push the evaluations of leaves[l](X / w_mx, Y / w_my), w_m the primitive m-th root of unity (a power of two dividing the
   * domain's extent): operand = l | (log2(mx) + 1) << 4 | (log2(my) + 1) << 10, a zero field = that axis is not scaled.  On the
   * evaluation grid this is leaves[l]'s table rotated by x_size/mx rows and y_size/my columns, so r(X/w, Y) and r(X/w, Y/w)
   * (prove/src/lib.rs:2110-2146: scale_coeffs_x / _y of r, then PolyExpr::poly of each) cost no transform of their own. */
  TKM_PEX_LEAF_SHIFT = 7
};
int32_t tkm_polyexpr_eval(tkm_ctx *ctx, const tkm_poly *const *leaves, uint32_t n_leaves, const uint32_t *program, uint32_t n_ops,
                          const uint8_t *consts32, uint32_t n_consts, size_t target_x, size_t target_y, tkm_poly **out);
/* out = a*ca + b*cb on the max power-of-two shape (operator impls :532-1281 and poly_comb!,
 * prove/src/lib.rs:30-38, fused: no host resize/clone).  ca/cb: 32-byte canonical or NULL (= 1). */
int32_t tkm_poly_axpby(tkm_ctx *ctx, const tkm_poly *a, const uint8_t *ca32, const tkm_poly *b, const uint8_t *cb32,
                       tkm_poly **out);
/* p(0,0) += s  (poly +/- scalar, :975-1020). */
int32_t tkm_poly_add_scalar(tkm_ctx *ctx, tkm_poly *p, const uint8_t s32[32]);
/* _mul (:1846-1996): product via pad -> 2 x NTT -> pointwise -> INTT; scalar fast paths included. */
int32_t tkm_poly_mul(tkm_ctx *ctx, const tkm_poly *a, const tkm_poly *b, tkm_poly **out);
/* scale_coeffs_x / scale_coeffs_y (:1553-1613): c_ij * sx^i * sy^j; NULL = 1. */
int32_t tkm_poly_scale_coeffs(tkm_ctx *ctx, const tkm_poly *p, const uint8_t *sx32, const uint8_t *sy32,
                              tkm_poly **out);
/* eval (:1742-1750): P(x, y). */
int32_t tkm_poly_eval(tkm_ctx *ctx, const tkm_poly *p, const uint8_t x32[32], const uint8_t y32[32], uint8_t out32[32]);
/* eval_x / eval_y (:1719-1740): partial evaluation, result shape 1 x y_size / x_size x 1. */
int32_t tkm_poly_eval_x(tkm_ctx *ctx, const tkm_poly *p, const uint8_t x32[32], tkm_poly **out);
int32_t tkm_poly_eval_y(tkm_ctx *ctx, const tkm_poly *p, const uint8_t y32[32], tkm_poly **out);
/* div_by_vanishing_opt (:2284-2410): P = Qx (X^c - 1) + Qy (Y^d - 1); p is optimize_size'd first
 * (it takes &mut self in the reference).  Qx shape = p shape, Qy shape = c x y_size. */
int32_t tkm_poly_div_by_vanishing(tkm_ctx *ctx, tkm_poly *p, size_t c, size_t d, tkm_poly **out_qx, tkm_poly **out_qy);
/* div_by_ruffini (:2412-2477): P = Qx (X - x) + Qy (Y - y) + r.  Qx shape = p shape, Qy = 1 x y_size. */
int32_t tkm_poly_div_by_ruffini(tkm_ctx *ctx, const tkm_poly *p, const uint8_t x32[32], const uint8_t y32[32],
                                tkm_poly **out_qx, tkm_poly **out_qy, uint8_t out_r32[32]);
/* divide_x / divide_y (:1998-2094): univariate long division of every line along X (y_dir = 0) or Y (y_dir = 1) by an
 * X- (Y-) univariate denominator; P = Q * D + R with deg R < deg D along that axis.  Q and R have the shape of p (a constant
 * denominator gives Q = p / c and the 1 x 1 zero remainder).  The reference's panics become TKM_ERR_INVALID_ARGUMENT with the
 * same texts ("Denominator for divide_x must be X-univariate", "Numer.degree < Denom.degree for divide_x", "Divide by zero"). */
int32_t tkm_poly_divide_uni(tkm_ctx *ctx, const tkm_poly *p, const tkm_poly *denom, int32_t y_dir, tkm_poly **out_q, tkm_poly **out_r);
/* encode_poly (iotools/mod.rs:2041-2113; group_structures/mod.rs:59-119): optimize_size, bounds
 * check against the CRS grid, commit the trimmed rectangle.  Zero polynomial -> identity. */
int32_t tkm_poly_commit(tkm_ctx *ctx, tkm_poly *p, const tkm_crs *crs, uint8_t out96[96]);
/* The same commitment as a begin/end pair: begin queues the MSM and returns a ticket at once; its serial recombination
 * tail (one warp, 1-2 ms) runs on a side stream, so a caller that queues the next commitment before calling end overlaps
 * that tail with the next accumulation (prove0..4 produce up to six commitments before the transcript needs any of them).
 * The polynomial may be freed or modified after begin returns.  At most 32 tickets in flight. */
int32_t tkm_poly_commit_begin(tkm_ctx *ctx, tkm_poly *p, const tkm_crs *crs, int32_t *out_ticket);
int32_t tkm_commit_end(tkm_ctx *ctx, int32_t ticket, uint8_t out96[96]);
/* The begin half for the plain and the sparse-gather MSM (tkm_msm_g1, tkm_msm_g1_indexed; tkm_commit_end resolves the
 * ticket): Prover::init's binding commitments O_pub_free / O_mid / O_prv and their blinding terms (prove/src/lib.rs:1092-1176,
 * group_structures/mod.rs:145-300) are independent of each other, so their tails overlap.  The device buffers may be
 * released with tkm_dev_free (stream-ordered) as soon as begin returns. */
int32_t tkm_msm_g1_begin(tkm_ctx *ctx, const void *dev_scalars, int32_t scalars_mont, const void *dev_bases_mont, size_t n,
                         int32_t *out_ticket);
int32_t tkm_msm_g1_indexed_begin(tkm_ctx *ctx, const void *dev_scalars, int32_t scalars_mont, const void *dev_bases_mont,
                                 const void *dev_idx, size_t n, int32_t *out_ticket);

/* ---- multi-GPU (the reference pins device 0, libs/src/utils/mod.rs:90-96, so this is new surface) --------
 * One context per GPU -- one process per GPU, or several contexts in one process.  NCCL (over NVLink/NVSwitch) is the
 * plumbing and is loaded at run time; single-GPU use needs none of this.  Bootstrap like ncclCommInitRank: rank 0 obtains
 * an id, hands it to every rank out of band (MPI, a file, torch.distributed, ...), and every rank calls tkm_comm_init
 * (collective).  The sharded entry points are collective too: every rank of the communicator must call them. */
#define TKM_COMM_ID_BYTES 128
int32_t tkm_comm_unique_id(uint8_t out_id[TKM_COMM_ID_BYTES]);
int32_t tkm_comm_init(tkm_ctx *ctx, const uint8_t id[TKM_COMM_ID_BYTES], int32_t rank, int32_t world);
int32_t tkm_comm_destroy(tkm_ctx *ctx);
int32_t tkm_comm_rank(tkm_ctx *ctx, int32_t *out_rank, int32_t *out_world);
/* msm::msm over a point range per GPU: every rank passes its slice (device-resident, like tkm_msm_g1); the 96-byte partial
 * sums are all-gathered (the only collective) and summed, and every rank receives the total. */
int32_t tkm_msm_g1_sharded(tkm_ctx *ctx, const void *dev_scalars, int32_t scalars_mont, const void *dev_bases_mont, size_t n_local,
                           uint8_t out96[96]);
/* _biNTT of an x_size x y_size polynomial sharded by rows.  Forward: dev_in = this rank's rows [x/G][y] (Montgomery form),
 * dev_out = its column shard of the evaluations [x][y/G] (entry (k, l) is the value at (w_x^k, w_y^(rank*y/G + l))); local
 * Y pass, one all-to-all of (x/G) x (y/G) tiles, local X pass.  Inverse: dev_in = column shard, dev_out = row shard of the
 * coefficients.  Both buffers hold x*y/G elements and must not alias; the call is asynchronous on the context stream. */
int32_t tkm_bintt_sharded(tkm_ctx *ctx, const void *dev_in, void *dev_out, size_t x_size, size_t y_size, int32_t dir,
                          const uint8_t *coset_x32, const uint8_t *coset_y32);

/* Keccak-256 with the original 0x01 padding (tiny_keccak::Keccak::v256): the hash of RollingKeccakTranscript
 * (prove/src/lib.rs:3211-3519).  Host only, no device needed. */
int32_t tkm_host_keccak256(const uint8_t *data, size_t len, uint8_t out32[32]);
/* ---- host-side data loader (no device work) -------------------------------------------------------
 * Every "0x..." string of a JSON text, in file order, as 32-byte canonical little-endian scalars reduced mod r:
 * the HexString -> ScalarField::from_hex parsing of placementVariables.json / instance.json
 * (libs/src/iotools/mod.rs:126-146,367-372,1582-1588).  out_count receives the number of scalars written. */
int32_t tkm_host_parse_hex_scalars(const char *text, size_t len, uint8_t *out32, size_t capacity, size_t *out_count);
/* iden3 .r1cs binary -> CSR (R1csBinary::read / scan_constraints, libs/src/iotools/mod.rs:505-650).  Call once with
 * row_ptr = NULL to get n_wires, n_constraints and nnz[3] (entries of A, B, C); call again with row_ptr[3 * (n_constraints + 1)],
 * wire[nnz total] and coeff32[32 * nnz total] to fill them (matrix-major entry order; nnz[] as returned by the first call). */
int32_t tkm_host_parse_r1cs(const uint8_t *data, size_t len, uint32_t *n_wires, uint32_t *n_constraints, size_t nnz[3], uint32_t *row_ptr,
                            uint32_t *wire, uint8_t *coeff32);

/* ---- instrumentation used by bench.py (not part of the reference API) ---------------------- */
/* Times one device-resident launch sequence with CUDA events on the context stream; ms out. */
int32_t tkm_event_time_begin(tkm_ctx *ctx);
int32_t tkm_event_time_end(tkm_ctx *ctx, float *out_ms);
/* Duration (CUDA events on the context stream) of the most recent dominant-kernel launch: k_accumulate of the last MSM,
 * or all k_ntt_pass launches of the last (bi)NTT.  Used by bench.py for the per-kernel roofline. */
int32_t tkm_kernel_time_last(tkm_ctx *ctx, float *out_ms);
/* Work accounting of the most recent MSM accumulation pass: the number L of affine pair-tree levels it ran (0 = chained
 * XYZZ additions only) and the entry counts n_0 .. n_L of the levels (n_l - n_{l+1} affine additions at level l; the
 * n_L remaining entries go through XYZZ mixed additions).  bench.py's roofline counts the issued multiplications with it. */
int32_t tkm_msm_tree_stats(tkm_ctx *ctx, uint32_t *out_levels, uint64_t out_counts[9]);
/* The same for the fused expression kernel of the most recent tkm_polyexpr_eval (bench.py's poly_engine GB/s). */
int32_t tkm_poly_kernel_time_last(tkm_ctx *ctx, float *out_ms);
/* Kernel launches issued by this library on this context since creation. */
int32_t tkm_launch_count(tkm_ctx *ctx, uint64_t *out);
/* Micro-benchmarks: integer pipe peak (dependent-free IMAD / IMAD.WIDE streams) and field-mul rate.
 * kind: 0 = IMAD.U32, 1 = IMAD.WIDE.U32 (64-bit addend), 2 = Fr mul, 3 = Fq mul, 4 = XYZZ mixed add,
 * 5 = IMAD.WIDE.U32.X carry chains (the form the field multiplier issues), 6 = Fr NTT butterfly (product + add + sub),
 * 7 = the same butterfly on the round-1 reduction (add-with-carry chains; kept for comparison),
 * 8 / 9 / 10 = ONE thread's chain of Fq inversions (binary Euclid / binary GCD on approximations / Fermat): 1 / latency.
 * out = ops/s. */
int32_t tkm_microbench(tkm_ctx *ctx, int32_t kind, double *out_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* TOKAMAK_B200_H */
