/* ORACLE -- CPU restatement of the Tokamak zk-EVM hot path (TEST INFRASTRUCTURE ONLY).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (libtokamak_b200.so) never
 * links, loads or calls it.
 *
 * PARITY UNPINNED: the reference has no golden vectors for this path and its
 * arithmetic lives in un-vendored crates (icicle-* v3.8.0, Cargo.toml:20-23), so
 * this file restates the published algorithms (Montgomery fields, short
 * Weierstrass Jacobian arithmetic, bucket-method MSM, radix-2 NTT) and the
 * reference's own call-site semantics.  It is cross-checked against the
 * independent big-integer restatement in oracle/pyref.py and against the
 * in-tree constants (fixed-tau generator on curve, root of unity order).
 *
 * Paths cited are relative to /root/reference/packages/backend/.
 * All interface buffers are canonical little-endian: Fr = 4 x u64 (32 B),
 * Fq = 6 x u64 (48 B), G1 affine = x || y (96 B), all-zero = identity.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ Fr */
static const uint64_t FR_MOD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
                                   0x73eda753299d7d48ULL};
static const uint64_t FR_R2[4] = {0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL,
                                  0x0748d9d99f59ff11ULL};
#define FNAME(x) fr_##x
#define NL 4
#define MODULUS FR_MOD
#define INV64 0xfffffffeffffffffULL
#define R2 FR_R2
#include "mont_tmpl.h"

/* ------------------------------------------------------------------ Fq */
static const uint64_t FQ_MOD[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                                   0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const uint64_t FQ_R2[6] = {0xf4df1f341c341746ULL, 0x0a76e6a609d104f1ULL, 0x8de5476c4c95b6d5ULL,
                                  0x67eb88a9939d83c0ULL, 0x9a793e85b519952dULL, 0x11988fe592cae3aaULL};
#define FNAME(x) fq_##x
#define NL 6
#define MODULUS FQ_MOD
#define INV64 0x89f3fffcfffcfffdULL
#define R2 FQ_R2
#include "mont_tmpl.h"

/* 5^((r-1)/2^32): the 2^32-th root ICICLE's domain is built from. */
static const uint64_t FR_ROU[4] = {0x1b788f500b912f1fULL, 0xc4024ff270b3e094ULL, 0x0fd56dc8d168d6c0ULL,
                                   0x0212d79e5b416b6fULL};

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static inline void fr_load(fr_t *r, const uint64_t *src) {
  fr_t t;
  memcpy(&t, src, 32);
  fr_to_mont(r, &t);
}
static inline void fr_store(uint64_t *dst, const fr_t *a) {
  fr_t t;
  fr_from_mont(&t, a);
  memcpy(dst, &t, 32);
}
static void fr_from_u64(fr_t *r, uint64_t v) {
  fr_t t;
  memset(&t, 0, sizeof t);
  t.l[0] = v;
  fr_to_mont(r, &t);
}

/* ---------------------------------------------------- element-wise Fr ops
 * VecOps::{add,sub,mul} (vector_operations/mod.rs:19-141; bivariate_polynomial/mod.rs:1974). */
void orc_fr_vec_op(int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    fr_t x, y, z;
    memcpy(&x, a + 4 * i, 32);
    memcpy(&y, b + 4 * i, 32);
    if (op == 0)
      fr_add(&z, &x, &y);
    else if (op == 1)
      fr_sub(&z, &x, &y);
    else {
      fr_t xm;
      fr_to_mont(&xm, &x);
      fr_mul(&z, &xm, &y); /* (xR)*y/R = xy */
    }
    memcpy(out + 4 * i, &z, 32);
  }
}
/* VecOps::inv with inv(0)=0 (bivariate_polynomial/mod.rs:2180). */
void orc_fr_vec_inv(const uint64_t *a, uint64_t *out, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    fr_t x, y;
    fr_load(&x, a + 4 * i);
    fr_inv(&y, &x);
    fr_store(out + 4 * i, &y);
  }
}

/* ntt::get_root_of_unity(2^log_n) (bivariate_polynomial/mod.rs:50). */
static void root_of_unity_mont(fr_t *w, unsigned log_n) {
  fr_t t;
  memcpy(&t, FR_ROU, 32);
  fr_to_mont(w, &t);
  for (unsigned i = log_n; i < 32; i++) fr_sqr(w, w);
}
void orc_root_of_unity(unsigned log_n, uint64_t *out) {
  fr_t w;
  root_of_unity_mont(&w, log_n);
  fr_store(out, &w);
}

/* ------------------------------------------------------------------ NTT
 * ICICLE ntt::ntt semantics (libs/src/tests.rs:107-180):
 * natural order in/out, inverse scales by 1/n, coset forward pre-scales by g^i,
 * coset inverse post-scales by g^-i. */
static unsigned ilog2(size_t n) {
  unsigned l = 0;
  while (((size_t)1 << l) < n) l++;
  return l;
}
/* tw[k] = w^k, k < n/2 (Montgomery). */
static fr_t *make_twiddles(size_t n, int inverse) {
  fr_t *tw = (fr_t *)malloc(sizeof(fr_t) * (n / 2 ? n / 2 : 1));
  fr_t w;
  root_of_unity_mont(&w, ilog2(n));
  if (inverse) fr_inv(&w, &w);
  fr_one(&tw[0]);
  for (size_t k = 1; k < n / 2; k++) fr_mul(&tw[k], &tw[k - 1], &w);
  return tw;
}
static void ntt_core(fr_t *a, size_t n, const fr_t *tw) {
  unsigned logn = ilog2(n);
  for (size_t i = 0, j = 0; i < n; i++) {
    if (i < j) {
      fr_t t = a[i];
      a[i] = a[j];
      a[j] = t;
    }
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
  }
  for (unsigned s = 1; s <= logn; s++) {
    size_t len = (size_t)1 << s, half = len >> 1, step = n / len;
    for (size_t st = 0; st < n; st += len)
      for (size_t k = 0; k < half; k++) {
        fr_t v, u = a[st + k];
        fr_mul(&v, &a[st + k + half], &tw[k * step]);
        fr_add(&a[st + k], &u, &v);
        fr_sub(&a[st + k + half], &u, &v);
      }
  }
}
/* One vector, Montgomery in/out, with pre/post scaling tables prepared by caller. */
static void ntt_vec(fr_t *a, size_t n, int inverse, const fr_t *tw, const fr_t *coset_pows, const fr_t *n_inv) {
  if (n == 1) return;
  if (!inverse && coset_pows)
    for (size_t i = 0; i < n; i++) fr_mul(&a[i], &a[i], &coset_pows[i]);
  ntt_core(a, n, tw);
  if (inverse) {
    for (size_t i = 0; i < n; i++) {
      fr_mul(&a[i], &a[i], n_inv);
      if (coset_pows) fr_mul(&a[i], &a[i], &coset_pows[i]);
    }
  }
}
static fr_t *make_coset_pows(size_t n, const uint64_t *coset, int inverse) {
  if (!coset) return NULL;
  fr_t g, one;
  fr_load(&g, coset);
  fr_one(&one);
  if (fr_eq(&g, &one)) return NULL;
  if (inverse) fr_inv(&g, &g);
  fr_t *p = (fr_t *)malloc(sizeof(fr_t) * n);
  p[0] = one;
  for (size_t i = 1; i < n; i++) fr_mul(&p[i], &p[i - 1], &g);
  return p;
}
/* Row batch: `batch` contiguous vectors of length n (NTTConfig.batch_size, columns_batch=false). */
static void ntt_rows_mont(fr_t *data, size_t n, size_t batch, int inverse, const uint64_t *coset) {
  if (n == 1) return;
  fr_t *tw = make_twiddles(n, inverse);
  fr_t *cp = make_coset_pows(n, coset, inverse);
  fr_t n_inv;
  fr_from_u64(&n_inv, (uint64_t)n);
  fr_inv(&n_inv, &n_inv);
#pragma omp parallel for schedule(static)
  for (size_t b = 0; b < batch; b++) ntt_vec(data + b * n, n, inverse, tw, cp, &n_inv);
  free(tw);
  free(cp);
}
/* Column batch: `batch` interleaved vectors, element j of vector b at j*batch+b (columns_batch=true). */
static void ntt_cols_mont(fr_t *data, size_t n, size_t batch, int inverse, const uint64_t *coset) {
  if (n == 1) return;
  fr_t *tw = make_twiddles(n, inverse);
  fr_t *cp = make_coset_pows(n, coset, inverse);
  fr_t n_inv;
  fr_from_u64(&n_inv, (uint64_t)n);
  fr_inv(&n_inv, &n_inv);
#pragma omp parallel
  {
    fr_t *col = (fr_t *)malloc(sizeof(fr_t) * n);
#pragma omp for schedule(static)
    for (size_t b = 0; b < batch; b++) {
      for (size_t j = 0; j < n; j++) col[j] = data[j * batch + b];
      ntt_vec(col, n, inverse, tw, cp, &n_inv);
      for (size_t j = 0; j < n; j++) data[j * batch + b] = col[j];
    }
    free(col);
  }
  free(tw);
  free(cp);
}
static fr_t *load_vec(const uint64_t *src, size_t n) {
  fr_t *v = (fr_t *)malloc(sizeof(fr_t) * (n ? n : 1));
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) fr_load(&v[i], src + 4 * i);
  return v;
}
static void store_vec(uint64_t *dst, const fr_t *v, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) fr_store(dst + 4 * i, &v[i]);
}
int orc_ntt(const uint64_t *in, uint64_t *out, size_t n, size_t batch, int columns_batch, int inverse,
            const uint64_t *coset) {
  if (n == 0 || (n & (n - 1))) return -1;
  fr_t *v = load_vec(in, n * batch);
  if (columns_batch)
    ntt_cols_mont(v, n, batch, inverse, coset);
  else
    ntt_rows_mont(v, n, batch, inverse, coset);
  store_vec(out, v, n * batch);
  free(v);
  return 0;
}
/* DensePolynomialExt::_biNTT (bivariate_polynomial/mod.rs:1422-1478). */
static void bintt_mont(fr_t *v, size_t x, size_t y, int inverse, const uint64_t *cx, const uint64_t *cy) {
  if (x == 1) {
    ntt_rows_mont(v, y, 1, inverse, cy);
  } else if (y == 1) {
    ntt_rows_mont(v, x, 1, inverse, cx);
  } else {
    ntt_rows_mont(v, y, x, inverse, cy);
    ntt_cols_mont(v, x, y, inverse, cx);
  }
}
int orc_bintt(const uint64_t *in, uint64_t *out, size_t x, size_t y, int inverse, const uint64_t *cx,
              const uint64_t *cy) {
  if (x == 0 || y == 0 || (x & (x - 1)) || (y & (y - 1))) return -1;
  fr_t *v = load_vec(in, x * y);
  bintt_mont(v, x, y, inverse, cx, cy);
  store_vec(out, v, x * y);
  free(v);
  return 0;
}
/* _mul generic path (bivariate_polynomial/mod.rs:1880-1995) on already-padded operands of shape x*y. */
int orc_poly_mul_padded(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t x, size_t y) {
  if (x == 0 || y == 0 || (x & (x - 1)) || (y & (y - 1))) return -1;
  size_t n = x * y;
  fr_t *fa = load_vec(a, n), *fb = load_vec(b, n);
  bintt_mont(fa, x, y, 0, NULL, NULL);
  bintt_mont(fb, x, y, 0, NULL, NULL);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) fr_mul(&fa[i], &fa[i], &fb[i]);
  bintt_mont(fa, x, y, 1, NULL, NULL);
  store_vec(out, fa, n);
  free(fa);
  free(fb);
  return 0;
}

/* -------------------------------------------------- bivariate poly helpers */
/* scale_coeffs_x / scale_coeffs_y (bivariate_polynomial/mod.rs:1553-1613): c_ij * sx^i * sy^j. */
void orc_scale_coeffs(const uint64_t *in, uint64_t *out, size_t x, size_t y, const uint64_t *sx, const uint64_t *sy) {
  fr_t *px = (fr_t *)malloc(sizeof(fr_t) * x), *py = (fr_t *)malloc(sizeof(fr_t) * y);
  fr_t gx, gy;
  fr_one(&gx);
  fr_one(&gy);
  if (sx) fr_load(&gx, sx);
  if (sy) fr_load(&gy, sy);
  fr_one(&px[0]);
  fr_one(&py[0]);
  for (size_t i = 1; i < x; i++) fr_mul(&px[i], &px[i - 1], &gx);
  for (size_t j = 1; j < y; j++) fr_mul(&py[j], &py[j - 1], &gy);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < x; i++)
    for (size_t j = 0; j < y; j++) {
      fr_t c;
      fr_load(&c, in + 4 * (i * y + j));
      fr_mul(&c, &c, &px[i]);
      fr_mul(&c, &c, &py[j]);
      fr_store(out + 4 * (i * y + j), &c);
    }
  free(px);
  free(py);
}
/* eval (bivariate_polynomial/mod.rs:1719-1750): P(px, py). */
void orc_eval(const uint64_t *in, size_t x, size_t y, const uint64_t *px, const uint64_t *py, uint64_t *out) {
  fr_t ax, ay, acc;
  fr_load(&ax, px);
  fr_load(&ay, py);
  fr_t *rows = (fr_t *)malloc(sizeof(fr_t) * x);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < x; i++) {
    fr_t r, c;
    memset(&r, 0, sizeof r);
    for (size_t j = y; j-- > 0;) {
      fr_mul(&r, &r, &ay);
      fr_load(&c, in + 4 * (i * y + j));
      fr_add(&r, &r, &c);
    }
    rows[i] = r;
  }
  memset(&acc, 0, sizeof acc);
  for (size_t i = x; i-- > 0;) {
    fr_mul(&acc, &acc, &ax);
    fr_add(&acc, &acc, &rows[i]);
  }
  fr_store(out, &acc);
  free(rows);
}
/* div_by_vanishing_opt (bivariate_polynomial/mod.rs:2284-2410); p is x*y with c | x, d | y.
 * qx: x*y, qy: c*y. Add/sub only, so it runs on canonical values directly. */
int orc_div_by_vanishing_opt(const uint64_t *p, size_t x, size_t y, size_t c, size_t d, uint64_t *qx, uint64_t *qy) {
  if (c == 0 || d == 0 || x % c || y % d) return -1;
  size_t m = x / c;
  const fr_t *P = (const fr_t *)p;
  fr_t *QX = (fr_t *)qx, *QY = (fr_t *)qy;
  fr_t *acc = (fr_t *)calloc(c * y, sizeof(fr_t));
  fr_t *b = (fr_t *)malloc(sizeof(fr_t) * x * y);
  memcpy(b, P, sizeof(fr_t) * x * y);
  memset(QX, 0, sizeof(fr_t) * x * y);
  memset(QY, 0, sizeof(fr_t) * c * y);
  for (size_t bx = 0; bx < m; bx++)
    for (size_t lx = 0; lx < c; lx++)
      for (size_t j = 0; j < y; j++) fr_add(&acc[lx * y + j], &acc[lx * y + j], &P[(bx * c + lx) * y + j]);
  if (y > d) {
    for (size_t i = 0; i < c; i++)
      for (size_t j = 0; j < y - d; j++) {
        fr_t prev;
        memset(&prev, 0, sizeof prev);
        if (j >= d) prev = QY[i * y + j - d];
        fr_sub(&QY[i * y + j], &prev, &acc[i * y + j]);
      }
    for (size_t i = 0; i < c; i++)
      for (size_t j = 0; j < y - d; j++) {
        fr_t co = QY[i * y + j];
        fr_add(&b[i * y + j], &b[i * y + j], &co);
        fr_sub(&b[i * y + j + d], &b[i * y + j + d], &co);
      }
  }
  if (x > c) {
    for (size_t i = 0; i < x - c; i++)
      for (size_t j = 0; j < y; j++) {
        fr_t prev;
        memset(&prev, 0, sizeof prev);
        if (i >= c) prev = QX[(i - c) * y + j];
        fr_sub(&QX[i * y + j], &prev, &b[i * y + j]);
      }
  }
  free(acc);
  free(b);
  return 0;
}
/* div_by_ruffini (bivariate_polynomial/mod.rs:2412-2477): P = Qx (X-px) + Qy (Y-py) + r.
 * qx: x*y, qy: y, r: 1. */
void orc_div_by_ruffini(const uint64_t *p, size_t x, size_t y, const uint64_t *px, const uint64_t *py, uint64_t *qx,
                        uint64_t *qy, uint64_t *r) {
  fr_t ax, ay;
  fr_load(&ax, px);
  fr_load(&ay, py);
  fr_t *rx = (fr_t *)malloc(sizeof(fr_t) * y);
  memset(qx, 0, 32 * x * y);
  memset(qy, 0, 32 * y);
#pragma omp parallel for schedule(static)
  for (size_t j = 0; j < y; j++) {
    fr_t b, c;
    if (x < 2) {
      fr_load(&rx[j], p + 4 * j);
      continue;
    }
    fr_load(&b, p + 4 * ((x - 1) * y + j));
    fr_store(qx + 4 * ((x - 2) * y + j), &b);
    for (size_t i = 3; i <= x; i++) {
      fr_mul(&b, &b, &ax);
      fr_load(&c, p + 4 * ((x - i + 1) * y + j));
      fr_add(&b, &b, &c);
      fr_store(qx + 4 * ((x - i) * y + j), &b);
    }
    fr_mul(&b, &b, &ax);
    fr_load(&c, p + 4 * j);
    fr_add(&rx[j], &b, &c);
  }
  if (y < 2) {
    fr_store(r, &rx[0]);
  } else {
    fr_t b = rx[y - 1];
    fr_store(qy + 4 * (y - 2), &b);
    for (size_t i = 3; i <= y; i++) {
      fr_mul(&b, &b, &ay);
      fr_add(&b, &b, &rx[y - i + 1]);
      fr_store(qy + 4 * (y - i), &b);
    }
    fr_mul(&b, &b, &ay);
    fr_add(&b, &b, &rx[0]);
    fr_store(r, &b);
  }
  free(rx);
}

/* ------------------------------------------------------------------ G1
 * y^2 = x^3 + 4.  Jacobian (X,Y,Z), Z = 0 is the identity.  Restates the
 * G1Projective arithmetic behind G1serde ops (group_structures/mod.rs:888-947)
 * and msm::msm (iotools/mod.rs:2093-2099). */
typedef struct { fq_t X, Y, Z; } jac_t;
typedef struct { fq_t x, y; int inf; } aff_t;

static void jac_set_inf(jac_t *p) { memset(p, 0, sizeof *p); }
static int jac_is_inf(const jac_t *p) { return fq_is_zero(&p->Z); }
static void jac_double(jac_t *r, const jac_t *p) {
  if (jac_is_inf(p)) {
    *r = *p;
    return;
  }
  fq_t A, B, C, D, E, F, t, X3, Y3, Z3;
  fq_sqr(&A, &p->X);
  fq_sqr(&B, &p->Y);
  fq_sqr(&C, &B);
  fq_add(&t, &p->X, &B);
  fq_sqr(&t, &t);
  fq_sub(&t, &t, &A);
  fq_sub(&t, &t, &C);
  fq_add(&D, &t, &t);
  fq_add(&E, &A, &A);
  fq_add(&E, &E, &A);
  fq_sqr(&F, &E);
  fq_sub(&X3, &F, &D);
  fq_sub(&X3, &X3, &D);
  fq_sub(&t, &D, &X3);
  fq_mul(&Y3, &E, &t);
  fq_add(&C, &C, &C);
  fq_add(&C, &C, &C);
  fq_add(&C, &C, &C);
  fq_sub(&Y3, &Y3, &C);
  fq_mul(&Z3, &p->Y, &p->Z);
  fq_add(&Z3, &Z3, &Z3);
  r->X = X3;
  r->Y = Y3;
  r->Z = Z3;
}
static void jac_add_affine(jac_t *r, const jac_t *p, const aff_t *q) {
  if (q->inf) {
    *r = *p;
    return;
  }
  if (jac_is_inf(p)) {
    r->X = q->x;
    r->Y = q->y;
    fq_one(&r->Z);
    return;
  }
  fq_t Z1Z1, U2, S2, H, Rr, HH, HHH, V, t, X3, Y3, Z3;
  fq_sqr(&Z1Z1, &p->Z);
  fq_mul(&U2, &q->x, &Z1Z1);
  fq_mul(&S2, &q->y, &p->Z);
  fq_mul(&S2, &S2, &Z1Z1);
  fq_sub(&H, &U2, &p->X);
  fq_sub(&Rr, &S2, &p->Y);
  if (fq_is_zero(&H)) {
    if (fq_is_zero(&Rr))
      jac_double(r, p);
    else
      jac_set_inf(r);
    return;
  }
  fq_sqr(&HH, &H);
  fq_mul(&HHH, &H, &HH);
  fq_mul(&V, &p->X, &HH);
  fq_sqr(&X3, &Rr);
  fq_sub(&X3, &X3, &HHH);
  fq_sub(&X3, &X3, &V);
  fq_sub(&X3, &X3, &V);
  fq_sub(&t, &V, &X3);
  fq_mul(&Y3, &Rr, &t);
  fq_mul(&t, &p->Y, &HHH);
  fq_sub(&Y3, &Y3, &t);
  fq_mul(&Z3, &p->Z, &H);
  r->X = X3;
  r->Y = Y3;
  r->Z = Z3;
}
static void jac_add(jac_t *r, const jac_t *p, const jac_t *q) {
  if (jac_is_inf(q)) {
    *r = *p;
    return;
  }
  if (jac_is_inf(p)) {
    *r = *q;
    return;
  }
  fq_t Z1Z1, Z2Z2, U1, U2, S1, S2, H, Rr, HH, HHH, V, t, X3, Y3, Z3;
  fq_sqr(&Z1Z1, &p->Z);
  fq_sqr(&Z2Z2, &q->Z);
  fq_mul(&U1, &p->X, &Z2Z2);
  fq_mul(&U2, &q->X, &Z1Z1);
  fq_mul(&S1, &p->Y, &q->Z);
  fq_mul(&S1, &S1, &Z2Z2);
  fq_mul(&S2, &q->Y, &p->Z);
  fq_mul(&S2, &S2, &Z1Z1);
  fq_sub(&H, &U2, &U1);
  fq_sub(&Rr, &S2, &S1);
  if (fq_is_zero(&H)) {
    if (fq_is_zero(&Rr))
      jac_double(r, p);
    else
      jac_set_inf(r);
    return;
  }
  fq_sqr(&HH, &H);
  fq_mul(&HHH, &H, &HH);
  fq_mul(&V, &U1, &HH);
  fq_sqr(&X3, &Rr);
  fq_sub(&X3, &X3, &HHH);
  fq_sub(&X3, &X3, &V);
  fq_sub(&X3, &X3, &V);
  fq_sub(&t, &V, &X3);
  fq_mul(&Y3, &Rr, &t);
  fq_mul(&t, &S1, &HHH);
  fq_sub(&Y3, &Y3, &t);
  fq_mul(&Z3, &p->Z, &q->Z);
  fq_mul(&Z3, &Z3, &H);
  r->X = X3;
  r->Y = Y3;
  r->Z = Z3;
}
static void aff_load(aff_t *a, const uint64_t *src) {
  fq_t x, y;
  memcpy(&x, src, 48);
  memcpy(&y, src + 6, 48);
  a->inf = fq_is_zero(&x) && fq_is_zero(&y);
  fq_to_mont(&a->x, &x);
  fq_to_mont(&a->y, &y);
}
static void jac_store_affine(uint64_t *dst, const jac_t *p) {
  if (jac_is_inf(p)) {
    memset(dst, 0, 96);
    return;
  }
  fq_t zi, zi2, x, y;
  fq_inv(&zi, &p->Z);
  fq_sqr(&zi2, &zi);
  fq_mul(&x, &p->X, &zi2);
  fq_mul(&y, &p->Y, &zi2);
  fq_mul(&y, &y, &zi);
  fq_from_mont(&x, &x);
  fq_from_mont(&y, &y);
  memcpy(dst, &x, 48);
  memcpy(dst + 6, &y, 48);
}
int orc_g1_is_on_curve(const uint64_t *pt) {
  aff_t a;
  aff_load(&a, pt);
  if (a.inf) return 1;
  fq_t l, rr, four, t;
  fq_sqr(&l, &a.y);
  fq_sqr(&rr, &a.x);
  fq_mul(&rr, &rr, &a.x);
  memset(&t, 0, sizeof t);
  t.l[0] = 4;
  fq_to_mont(&four, &t);
  fq_add(&rr, &rr, &four);
  return fq_eq(&l, &rr);
}
void orc_g1_add(const uint64_t *a, const uint64_t *b, uint64_t *out) {
  aff_t A, B;
  jac_t J;
  aff_load(&A, a);
  aff_load(&B, b);
  jac_set_inf(&J);
  jac_add_affine(&J, &J, &A);
  jac_add_affine(&J, &J, &B);
  jac_store_affine(out, &J);
}
static void jac_mul_scalar(jac_t *r, const aff_t *base, const uint64_t *k /* 4 limbs canonical */) {
  jac_t acc;
  jac_set_inf(&acc);
  for (int i = 255; i >= 0; i--) {
    jac_double(&acc, &acc);
    if ((k[i >> 6] >> (i & 63)) & 1) jac_add_affine(&acc, &acc, base);
  }
  *r = acc;
}
/* G1serde * ScalarField (group_structures/mod.rs:929-947). */
void orc_g1_mul(const uint64_t *pt, const uint64_t *k, uint64_t *out) {
  aff_t A;
  jac_t J;
  aff_load(&A, pt);
  jac_mul_scalar(&J, &A, k);
  jac_store_affine(out, &J);
}
/* N independent 1-point MSMs with one shared base = fixed-base batch scalar-mul
 * (from_coef_vec_to_g1serde_vec, iotools/mod.rs:1113-1135).  8-bit fixed windows. */
void orc_g1_fixed_base_mul_batch(const uint64_t *base, const uint64_t *scalars, size_t n, uint64_t *out) {
  aff_t B;
  aff_load(&B, base);
  /* table[w][d] = d * 2^(8w) * B, d = 1..255, as affine (via per-entry inversion; setup cost only) */
  const int W = 32;
  aff_t *tab = (aff_t *)malloc(sizeof(aff_t) * W * 256);
  jac_t cur;
  jac_set_inf(&cur);
  jac_add_affine(&cur, &cur, &B);
  for (int w = 0; w < W; w++) {
    jac_t acc;
    jac_set_inf(&acc);
    uint64_t tmp[12];
    aff_t curA;
    jac_store_affine(tmp, &cur);
    aff_load(&curA, tmp);
    for (int d = 1; d < 256; d++) {
      jac_add_affine(&acc, &acc, &curA);
      jac_store_affine(tmp, &acc);
      aff_load(&tab[w * 256 + d], tmp);
    }
    for (int k = 0; k < 8; k++) jac_double(&cur, &cur);
  }
  /* blocks of FB_BLK points share one field inversion for their conversions to affine (Montgomery's trick);
   * bench.py's reference arm generates 2^22 bases with this */
  enum { FB_BLK = 256 };
  const size_t nblk = (n + FB_BLK - 1) / FB_BLK;
#pragma omp parallel for schedule(static)
  for (size_t blk = 0; blk < nblk; blk++) {
    const size_t lo = blk * FB_BLK, hi = lo + FB_BLK < n ? lo + FB_BLK : n;
    jac_t acc[FB_BLK];
    fq_t pref[FB_BLK], run, one_m, tmp1;
    memset(&tmp1, 0, sizeof tmp1);
    ((uint64_t *)&tmp1)[0] = 1;
    fq_to_mont(&one_m, &tmp1);
    run = one_m;
    for (size_t i = lo; i < hi; i++) {
      jac_t *a = &acc[i - lo];
      jac_set_inf(a);
      const uint8_t *kb = (const uint8_t *)(scalars + 4 * i);
      for (int w = 0; w < W; w++)
        if (kb[w]) jac_add_affine(a, a, &tab[w * 256 + kb[w]]);
      pref[i - lo] = run;
      if (!jac_is_inf(a)) fq_mul(&run, &run, &a->Z);
    }
    fq_t inv;
    fq_inv(&inv, &run);
    for (size_t i = hi; i-- > lo;) {
      const jac_t *a = &acc[i - lo];
      if (jac_is_inf(a)) {
        memset(out + 12 * i, 0, 96);
        continue;
      }
      fq_t zi, zi2, x, y;
      fq_mul(&zi, &inv, &pref[i - lo]);
      fq_mul(&inv, &inv, &a->Z);
      fq_sqr(&zi2, &zi);
      fq_mul(&x, &a->X, &zi2);
      fq_mul(&y, &a->Y, &zi2);
      fq_mul(&y, &y, &zi);
      fq_from_mont(&x, &x);
      fq_from_mont(&y, &y);
      memcpy(out + 12 * i, &x, 48);
      memcpy(out + 12 * i + 6, &y, 48);
    }
  }
  free(tab);
}
/* Reference-semantics MSM (double-and-add per point): the slow, obviously-correct form. */
void orc_msm_g1_naive(const uint64_t *scalars, const uint64_t *bases, size_t n, uint64_t *out) {
  jac_t acc;
  jac_set_inf(&acc);
  for (size_t i = 0; i < n; i++) {
    aff_t A;
    jac_t J;
    aff_load(&A, bases + 12 * i);
    jac_mul_scalar(&J, &A, scalars + 4 * i);
    jac_add(&acc, &acc, &J);
  }
  jac_store_affine(out, &acc);
}
/* Bucket-method (Pippenger) MSM, the algorithm class of ICICLE's CPU backend for msm::msm
 * (call sites iotools/mod.rs:2093-2099; group_structures/mod.rs:108-114,135-141).
 * Unsigned c-bit windows; tasks = windows x point-chunks spread over OpenMP threads.
 * `stride_*` let the caller address a strided rectangle (encode_poly's trimmed rectangle of the
 * CRS grid, iotools/mod.rs:2075-2088): element k=(i,j) lives at i*row_stride + j, j < cols. */
int orc_msm_g1_rect(const uint64_t *scalars, size_t s_row_stride, const uint64_t *bases, size_t b_row_stride,
                    size_t rows, size_t cols, uint64_t *out) {
  size_t n = rows * cols;
  if (n == 0) {
    memset(out, 0, 96);
    return 0;
  }
  unsigned c = 1;
  while (((size_t)1 << (c + 4)) < n && c < 16) c++; /* c ~ log2(n) - 4 */
  if (c < 4) c = 4;
  unsigned W = (255 + c - 1) / c;
  int threads = orc_num_threads();
  size_t nchunks = (size_t)((2 * threads + W - 1) / W);
  if (nchunks < 1) nchunks = 1;
  if (nchunks > n) nchunks = n;
  size_t chunk = (n + nchunks - 1) / nchunks;
  size_t ntasks = (size_t)W * nchunks;
  size_t nb = ((size_t)1 << c);
  jac_t *partial = (jac_t *)malloc(sizeof(jac_t) * ntasks);
  aff_t *pts = (aff_t *)malloc(sizeof(aff_t) * n);
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < n; k++) aff_load(&pts[k], bases + 12 * ((k / cols) * b_row_stride + (k % cols)));
#pragma omp parallel
  {
    jac_t *buckets = (jac_t *)malloc(sizeof(jac_t) * nb);
#pragma omp for schedule(dynamic, 1)
    for (size_t t = 0; t < ntasks; t++) {
      unsigned w = (unsigned)(t / nchunks);
      size_t lo = (t % nchunks) * chunk, hi = lo + chunk;
      if (hi > n) hi = n;
      for (size_t b = 0; b < nb; b++) jac_set_inf(&buckets[b]);
      unsigned bit0 = w * c;
      for (size_t k = lo; k < hi; k++) {
        const uint64_t *s = scalars + 4 * ((k / cols) * s_row_stride + (k % cols));
        unsigned limb = bit0 >> 6, sh = bit0 & 63;
        uint64_t d = s[limb] >> sh;
        if (sh + c > 64 && limb + 1 < 4) d |= s[limb + 1] << (64 - sh);
        d &= (nb - 1);
        if (d) jac_add_affine(&buckets[d], &buckets[d], &pts[k]);
      }
      jac_t run, sum;
      jac_set_inf(&run);
      jac_set_inf(&sum);
      for (size_t b = nb - 1; b >= 1; b--) {
        jac_add(&run, &run, &buckets[b]);
        jac_add(&sum, &sum, &run);
      }
      partial[t] = sum;
    }
    free(buckets);
  }
  jac_t acc;
  jac_set_inf(&acc);
  for (int w = (int)W - 1; w >= 0; w--) {
    for (unsigned k = 0; k < c; k++) jac_double(&acc, &acc);
    for (size_t ch = 0; ch < nchunks; ch++) jac_add(&acc, &acc, &partial[(size_t)w * nchunks + ch]);
  }
  jac_store_affine(out, &acc);
  free(partial);
  free(pts);
  return 0;
}
int orc_msm_g1(const uint64_t *scalars, const uint64_t *bases, size_t n, uint64_t *out) {
  return orc_msm_g1_rect(scalars, n, bases, n, 1, n, out);
}
/* sum_i a_i*b_i mod r: lets tests check MSM over bases k_i*G via (sum s_i k_i)*G. */
void orc_fr_inner_product(const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out) {
  fr_t acc;
  memset(&acc, 0, sizeof acc);
#pragma omp parallel
  {
    fr_t loc;
    memset(&loc, 0, sizeof loc);
#pragma omp for schedule(static) nowait
    for (size_t i = 0; i < n; i++) {
      fr_t x, y;
      fr_load(&x, a + 4 * i);
      fr_load(&y, b + 4 * i);
      fr_mul(&x, &x, &y);
      fr_add(&loc, &loc, &x);
    }
#pragma omp critical
    fr_add(&acc, &acc, &loc);
  }
  fr_store(out, &acc);
}
/* Deterministic inputs: SplitMix64 -> 256-bit draw reduced mod r.
 * Matches oracle/pyref.py SplitMix64.fr(). Element i uses its own stream seeded seed + i*0x1000193
 * so generation parallelises and any sub-range can be regenerated. */
static uint64_t splitmix_next(uint64_t *s) {
  *s += 0x9E3779B97F4A7C15ULL;
  uint64_t z = *s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
void orc_random_fr(uint64_t seed, size_t n, uint64_t *out) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    uint64_t s = seed + (uint64_t)i * 0x1000193ULL;
    uint64_t v[4];
    for (int k = 0; k < 4; k++) v[k] = splitmix_next(&s);
    v[3] &= 0x7fffffffffffffffULL; /* < 2^255 < 2r: one conditional subtract reduces */
    if (fr_geq_mod(v)) fr_sub_mod_raw(v);
    if (fr_geq_mod(v)) fr_sub_mod_raw(v);
    memcpy(out + 4 * i, v, 32);
  }
}
