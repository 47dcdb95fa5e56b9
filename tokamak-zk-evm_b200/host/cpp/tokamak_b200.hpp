// Host-side C++ mirror of the reference's `libs` API over the C-ABI of libtokamak_b200 (include/tokamak_b200.h).
//
// The reference's host code is Rust (absent from this build image); this header is the compiled-language counterpart of
// the Rust shim sketched in INTEGRATION.md: same names and argument meaning as libs::bivariate_polynomial::
// DensePolynomialExt (libs/src/bivariate_polynomial/mod.rs:1283-1416), ScalarField, Sigma1::encode_poly
// (libs/src/group_structures/mod.rs:59-119) and msm_g1_bases (:127-143), RAII ownership like the reference's DeviceVec-
// backed values, `&`-style operators that never mutate their inputs, and panics mapped to exceptions (every non-zero
// status throws std::runtime_error carrying tkm_last_error()).  Header-only; link with -ltokamak_b200.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/tokamak_b200.h"

namespace tokamak_b200 {

inline void check(int32_t st) {
  if (st != TKM_OK) throw std::runtime_error(std::string(tkm_last_error()));
}

// ---------------------------------------------------------------- ScalarField (BLS12-381 Fr), canonical 4 x u64 LE
struct ScalarField {
  uint64_t l[4];
  static constexpr uint64_t R[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
  static constexpr uint64_t R2[4] = {0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full, 0x0748d9d99f59ff11ull};  // 2^512 mod r
  static constexpr uint64_t INV = 0xfffffffeffffffffull;  // -r^-1 mod 2^64
  static ScalarField zero() { return ScalarField{{0, 0, 0, 0}}; }
  static ScalarField one() { return ScalarField{{1, 0, 0, 0}}; }
  static ScalarField from_u32(uint32_t v) { return ScalarField{{v, 0, 0, 0}}; }
  static ScalarField from_u64(uint64_t v) { return ScalarField{{v, 0, 0, 0}}; }
  const uint8_t *bytes() const { return reinterpret_cast<const uint8_t *>(l); }
  uint8_t *bytes() { return reinterpret_cast<uint8_t *>(l); }
  bool operator==(const ScalarField &o) const { return std::memcmp(l, o.l, 32) == 0; }
  bool operator!=(const ScalarField &o) const { return !(*this == o); }
  static bool geq_r(const uint64_t *a) {
    for (int i = 3; i >= 0; i--) {
      if (a[i] != R[i]) return a[i] > R[i];
    }
    return true;
  }
  static void sub_r(uint64_t *a) {
    unsigned __int128 borrow = 0;
    for (int i = 0; i < 4; i++) {
      unsigned __int128 d = (unsigned __int128)a[i] - R[i] - borrow;
      a[i] = (uint64_t)d;
      borrow = (d >> 64) & 1;
    }
  }
  ScalarField operator+(const ScalarField &o) const {
    ScalarField r;
    unsigned __int128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (unsigned __int128)l[i] + o.l[i];
      r.l[i] = (uint64_t)c;
      c >>= 64;
    }
    if (c || geq_r(r.l)) sub_r(r.l);
    return r;
  }
  ScalarField operator-(const ScalarField &o) const {
    ScalarField r;
    unsigned __int128 borrow = 0;
    for (int i = 0; i < 4; i++) {
      unsigned __int128 d = (unsigned __int128)l[i] - o.l[i] - borrow;
      r.l[i] = (uint64_t)d;
      borrow = (d >> 64) & 1;
    }
    if (borrow) {
      unsigned __int128 c = 0;
      for (int i = 0; i < 4; i++) {
        c += (unsigned __int128)r.l[i] + R[i];
        r.l[i] = (uint64_t)c;
        c >>= 64;
      }
    }
    return r;
  }
  static ScalarField mont(const ScalarField &a, const ScalarField &b) {  // a * b / 2^256 mod r (CIOS)
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
      unsigned __int128 c = 0;
      for (int j = 0; j < 4; j++) {
        c += (unsigned __int128)a.l[j] * b.l[i] + t[j];
        t[j] = (uint64_t)c;
        c >>= 64;
      }
      c += t[4];
      t[4] = (uint64_t)c;
      t[5] = (uint64_t)(c >> 64);
      const uint64_t m = t[0] * INV;
      c = (unsigned __int128)m * R[0] + t[0];
      c >>= 64;
      for (int j = 1; j < 4; j++) {
        c += (unsigned __int128)m * R[j] + t[j];
        t[j - 1] = (uint64_t)c;
        c >>= 64;
      }
      c += t[4];
      t[3] = (uint64_t)c;
      t[4] = t[5] + (uint64_t)(c >> 64);
    }
    ScalarField r{{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq_r(r.l)) sub_r(r.l);
    return r;
  }
  ScalarField operator*(const ScalarField &o) const {
    ScalarField r2;
    std::memcpy(r2.l, R2, 32);
    return mont(mont(*this, o), r2);
  }
  // ScalarField::from_hex ("0x..." big-endian hex digits, reduced mod r), to_string (the reference prints 0x + 64 digits),
  // from_bytes_le / to_bytes_le (libs/src/iotools/mod.rs:126-146,1786-1815)
  static ScalarField from_hex(const std::string &hex) {
    size_t i = (hex.size() >= 2 && hex[0] == '0' && (hex[1] == 'x' || hex[1] == 'X')) ? 2 : 0;
    ScalarField r = zero();
    const ScalarField sixteen = from_u32(16);
    for (; i < hex.size(); i++) {
      const char ch = hex[i];
      int d = (ch >= '0' && ch <= '9') ? ch - '0' : (ch >= 'a' && ch <= 'f') ? ch - 'a' + 10 : (ch >= 'A' && ch <= 'F') ? ch - 'A' + 10 : -1;
      if (d < 0) throw std::runtime_error("invalid hex digit in ScalarField::from_hex");
      r = r * sixteen + from_u32((uint32_t)d);
    }
    return r;
  }
  std::string to_string() const {
    static const char *dig = "0123456789abcdef";
    std::string out = "0x";
    for (int i = 3; i >= 0; i--)
      for (int b = 60; b >= 0; b -= 4) out.push_back(dig[(l[i] >> b) & 15]);
    return out;
  }
  static ScalarField from_bytes_le(const uint8_t *b32) {
    ScalarField r;
    std::memcpy(r.l, b32, 32);
    while (geq_r(r.l)) sub_r(r.l);
    return r;
  }
  void to_bytes_le(uint8_t *out32) const { std::memcpy(out32, l, 32); }
  ScalarField pow(uint64_t e) const {
    ScalarField acc = one(), base = *this;
    while (e) {
      if (e & 1) acc = acc * base;
      base = base * base;
      e >>= 1;
    }
    return acc;
  }
  ScalarField inv() const {  // a^(r-2); inv(0) = 0 like the reference backend
    ScalarField acc = one(), base = *this;
    uint64_t e[4] = {R[0] - 2, R[1], R[2], R[3]};
    for (int i = 0; i < 4; i++)
      for (int b = 0; b < 64; b++) {
        if ((e[i] >> b) & 1) acc = acc * base;
        base = base * base;
      }
    return acc;
  }
};

// ScalarCfg::generate_random: deterministic xorshift stream of values below 2^254 (< r)
struct ScalarCfg {
  uint64_t s;
  explicit ScalarCfg(uint64_t seed) : s(seed * 0x9e3779b97f4a7c15ull + 1) {}
  uint64_t next() {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    return s;
  }
  std::vector<ScalarField> generate_random(size_t n) {
    std::vector<ScalarField> v(n);
    for (auto &x : v) {
      for (int i = 0; i < 4; i++) x.l[i] = next();
      x.l[3] &= 0x3fffffffffffffffull;
    }
    return v;
  }
};

struct G1Affine {
  uint8_t b[96];  // x || y, 48-byte little-endian canonical each; all-zero = identity
  static G1Affine zero() {
    G1Affine g;
    std::memset(g.b, 0, 96);
    return g;
  }
  bool operator==(const G1Affine &o) const { return std::memcmp(b, o.b, 96) == 0; }
};

// ---------------------------------------------------------------- context (check_device + NTT domain)
class Context {
 public:
  explicit Context(int device = 0) { check(tkm_ctx_create(device, &h_)); }
  ~Context() {
    if (h_) tkm_ctx_destroy(h_);
  }
  Context(const Context &) = delete;
  Context &operator=(const Context &) = delete;
  tkm_ctx *raw() const { return h_; }
  // init_ntt_domain_for_size (bivariate_polynomial/mod.rs:33-55)
  void init_ntt_domain_for_size(size_t size) {
    uint32_t lg = 0;
    while (((size_t)1 << lg) < size) lg++;
    check(tkm_ntt_domain_init(h_, lg));
  }
  ScalarField get_root_of_unity(uint64_t n) const {
    uint32_t lg = 0;
    while (((uint64_t)1 << lg) < n) lg++;
    ScalarField w;
    check(tkm_root_of_unity(lg, w.bytes()));
    return w;
  }
  G1Affine g1_mul(const G1Affine &a, const ScalarField &k) const {
    G1Affine r;
    check(tkm_g1_mul(h_, a.b, k.bytes(), r.b));
    return r;
  }
  G1Affine g1_add(const G1Affine &a, const G1Affine &c) const {
    G1Affine r;
    check(tkm_g1_add(h_, a.b, c.b, r.b));
    return r;
  }
  // msm_g1_bases (group_structures/mod.rs:127-143): panics on a length mismatch, identity for empty input
  G1Affine msm_g1_bases(const std::vector<ScalarField> &scalars, const std::vector<G1Affine> &bases) const {
    if (scalars.size() != bases.size()) throw std::runtime_error("msm input length mismatch");
    G1Affine r;
    check(tkm_msm_g1_host(h_, scalars.empty() ? nullptr : scalars[0].bytes(), bases.empty() ? nullptr : bases[0].b, scalars.size(), r.b));
    return r;
  }

 private:
  tkm_ctx *h_ = nullptr;
};

// ---------------------------------------------------------------- DensePolynomialExt
class DensePolynomialExt {
 public:
  DensePolynomialExt(const Context &c, tkm_poly *h) : c_(&c), h_(h) {}
  ~DensePolynomialExt() {
    if (h_) tkm_poly_free(c_->raw(), h_);
  }
  DensePolynomialExt(DensePolynomialExt &&o) noexcept : c_(o.c_), h_(o.h_) { o.h_ = nullptr; }
  DensePolynomialExt &operator=(DensePolynomialExt &&o) noexcept {
    if (this != &o) {
      if (h_) tkm_poly_free(c_->raw(), h_);
      c_ = o.c_;
      h_ = o.h_;
      o.h_ = nullptr;
    }
    return *this;
  }
  DensePolynomialExt(const DensePolynomialExt &o) : c_(o.c_) { check(tkm_poly_clone(c_->raw(), o.h_, &h_)); }  // Clone (:520-530)
  DensePolynomialExt &operator=(const DensePolynomialExt &o) {
    if (this != &o) *this = DensePolynomialExt(o);
    return *this;
  }
  tkm_poly *raw() const { return h_; }

  static DensePolynomialExt from_coeffs(const Context &c, const std::vector<ScalarField> &coeffs, size_t x_size, size_t y_size) {
    if (x_size * y_size != coeffs.size()) throw std::runtime_error("Mismatch between the coefficient vector and the polynomial size");
    tkm_poly *h = nullptr;
    check(tkm_poly_from_coeffs_host(c.raw(), coeffs[0].bytes(), x_size, y_size, &h));
    return DensePolynomialExt(c, h);
  }
  static DensePolynomialExt from_rou_evals(const Context &c, const std::vector<ScalarField> &evals, size_t x_size, size_t y_size,
                                           const ScalarField *coset_x = nullptr, const ScalarField *coset_y = nullptr) {
    if (x_size * y_size != evals.size()) throw std::runtime_error("Mismatch between the evaluation vector and the polynomial size");
    tkm_poly *h = nullptr;
    check(tkm_poly_from_evals_host(c.raw(), evals[0].bytes(), x_size, y_size, coset_x ? coset_x->bytes() : nullptr,
                                   coset_y ? coset_y->bytes() : nullptr, &h));
    return DensePolynomialExt(c, h);
  }
  std::vector<ScalarField> to_rou_evals(const ScalarField *coset_x = nullptr, const ScalarField *coset_y = nullptr) const {
    std::vector<ScalarField> out(x_size() * y_size());
    check(tkm_poly_to_evals_host(c_->raw(), h_, coset_x ? coset_x->bytes() : nullptr, coset_y ? coset_y->bytes() : nullptr, out[0].bytes()));
    return out;
  }
  std::vector<ScalarField> copy_coeffs() const {
    std::vector<ScalarField> out(x_size() * y_size());
    check(tkm_poly_copy_coeffs_host(c_->raw(), h_, out[0].bytes()));
    return out;
  }
  ScalarField get_coeff(size_t idx_x, size_t idx_y) const { return copy_coeffs()[idx_x * y_size() + idx_y]; }
  size_t x_size() const { return shape().first; }
  size_t y_size() const { return shape().second; }
  std::pair<size_t, size_t> shape() const {
    size_t x, y;
    check(tkm_poly_shape(h_, &x, &y));
    return {x, y};
  }
  std::pair<int64_t, int64_t> find_degree() const {
    int64_t xd, yd;
    check(tkm_poly_find_degree(c_->raw(), h_, &xd, &yd));
    return {xd, yd};
  }
  void resize(size_t tx, size_t ty) { check(tkm_poly_resize(c_->raw(), h_, tx, ty)); }
  void optimize_size() { check(tkm_poly_optimize_size(c_->raw(), h_)); }
  DensePolynomialExt mul_monomial(size_t ex, size_t ey) const {
    tkm_poly *h = nullptr;
    check(tkm_poly_mul_monomial(c_->raw(), h_, ex, ey, &h));
    return DensePolynomialExt(*c_, h);
  }
  DensePolynomialExt axpby(const ScalarField *ca, const DensePolynomialExt *other, const ScalarField *cb) const {
    tkm_poly *h = nullptr;
    check(tkm_poly_axpby(c_->raw(), h_, ca ? ca->bytes() : nullptr, other ? other->h_ : nullptr, cb ? cb->bytes() : nullptr, &h));
    return DensePolynomialExt(*c_, h);
  }
  DensePolynomialExt operator+(const DensePolynomialExt &o) const { return axpby(nullptr, &o, nullptr); }
  DensePolynomialExt operator-(const DensePolynomialExt &o) const {
    const ScalarField m1 = ScalarField::zero() - ScalarField::one();
    return axpby(nullptr, &o, &m1);
  }
  DensePolynomialExt operator-() const {
    const ScalarField m1 = ScalarField::zero() - ScalarField::one();
    return axpby(&m1, nullptr, nullptr);
  }
  DensePolynomialExt operator*(const DensePolynomialExt &o) const {
    tkm_poly *h = nullptr;
    check(tkm_poly_mul(c_->raw(), h_, o.h_, &h));
    return DensePolynomialExt(*c_, h);
  }
  DensePolynomialExt operator*(const ScalarField &s) const { return axpby(&s, nullptr, nullptr); }
  DensePolynomialExt operator+(const ScalarField &s) const {
    DensePolynomialExt r(*this);
    check(tkm_poly_add_scalar(c_->raw(), r.h_, s.bytes()));
    return r;
  }
  DensePolynomialExt operator-(const ScalarField &s) const { return *this + (ScalarField::zero() - s); }
  DensePolynomialExt scale_coeffs_x(const ScalarField &s) const { return scale(&s, nullptr); }
  DensePolynomialExt scale_coeffs_y(const ScalarField &s) const { return scale(nullptr, &s); }
  ScalarField eval(const ScalarField &x, const ScalarField &y) const {
    ScalarField r;
    check(tkm_poly_eval(c_->raw(), h_, x.bytes(), y.bytes(), r.bytes()));
    return r;
  }
  // zero / is_zero / degree (trait :1283-1416)
  static DensePolynomialExt zero(const Context &c, size_t x_size = 1, size_t y_size = 1) {
    tkm_poly *h = nullptr;
    check(tkm_poly_zero(c.raw(), x_size, y_size, &h));
    return DensePolynomialExt(c, h);
  }
  bool is_zero() const { return find_degree().first < 0; }
  std::pair<int64_t, int64_t> degree() const { return {(int64_t)x_size() - 1, (int64_t)y_size() - 1}; }  // upper bounds, like the reference's fields
  // += / -= (AddAssign / SubAssign, :532-1281): same union-shape semantics as the binary operators
  DensePolynomialExt &operator+=(const DensePolynomialExt &o) { return *this = *this + o; }
  DensePolynomialExt &operator-=(const DensePolynomialExt &o) { return *this = *this - o; }
  // eval_x / eval_y (:1719-1740): partial evaluations (1 x y_size and x_size x 1)
  DensePolynomialExt eval_x(const ScalarField &x) const {
    tkm_poly *h = nullptr;
    check(tkm_poly_eval_x(c_->raw(), h_, x.bytes(), &h));
    return DensePolynomialExt(*c_, h);
  }
  DensePolynomialExt eval_y(const ScalarField &y) const {
    tkm_poly *h = nullptr;
    check(tkm_poly_eval_y(c_->raw(), h_, y.bytes(), &h));
    return DensePolynomialExt(*c_, h);
  }
  // divide_x / divide_y (:1998-2094): (quotient, remainder) of the line-wise long division by a univariate denominator
  std::pair<DensePolynomialExt, DensePolynomialExt> divide_x(const DensePolynomialExt &denom) const { return divide(denom, 0); }
  std::pair<DensePolynomialExt, DensePolynomialExt> divide_y(const DensePolynomialExt &denom) const { return divide(denom, 1); }
  // poly_comb! (prove/src/lib.rs:30-38) and the shifted helpers (:48-124): sum of c * X^sx * Y^sy * p in one pass
  struct Term {
    ScalarField c;
    const DensePolynomialExt *p;
    uint32_t sx = 0, sy = 0;
  };
  static DensePolynomialExt lincomb(const std::vector<Term> &terms) {
    if (terms.empty()) throw std::runtime_error("empty polynomial combination");
    std::vector<const tkm_poly *> hs;
    std::vector<ScalarField> cs;
    std::vector<uint32_t> sx, sy;
    for (const Term &t : terms) {
      hs.push_back(t.p->h_);
      cs.push_back(t.c);
      sx.push_back(t.sx);
      sy.push_back(t.sy);
    }
    tkm_poly *h = nullptr;
    check(tkm_poly_lincomb(terms[0].p->c_->raw(), (uint32_t)terms.size(), hs.data(), cs[0].bytes(), sx.data(), sy.data(), &h));
    return DensePolynomialExt(*terms[0].p->c_, h);
  }
  const Context &context() const { return *c_; }
  // div_by_vanishing (legacy coset formulation, :2096-2282): the same decomposition P = Q_X (X^c - 1) + Q_Y (Y^d - 1) as
  // div_by_vanishing_opt (both fix Q_Y's X-degree below c, which makes the pair unique); the prover only calls _opt.
  std::pair<DensePolynomialExt, DensePolynomialExt> div_by_vanishing(size_t c, size_t d) { return div_by_vanishing_opt(c, d); }
  // div_by_vanishing_opt (:2284-2410): P = Q_X (X^c - 1) + Q_Y (Y^d - 1)
  std::pair<DensePolynomialExt, DensePolynomialExt> div_by_vanishing_opt(size_t c, size_t d) {
    tkm_poly *qx = nullptr, *qy = nullptr;
    check(tkm_poly_div_by_vanishing(c_->raw(), h_, c, d, &qx, &qy));
    return {DensePolynomialExt(*c_, qx), DensePolynomialExt(*c_, qy)};
  }
  // div_by_ruffini (:2412-2458): P = Q_X (X - x) + Q_Y (Y - y) + r
  struct Ruffini;
  Ruffini div_by_ruffini(const ScalarField &x, const ScalarField &y) const;

 private:
  std::pair<DensePolynomialExt, DensePolynomialExt> divide(const DensePolynomialExt &denom, int y_dir) const {
    tkm_poly *q = nullptr, *r = nullptr;
    check(tkm_poly_divide_uni(c_->raw(), h_, denom.h_, y_dir, &q, &r));
    return {DensePolynomialExt(*c_, q), DensePolynomialExt(*c_, r)};
  }
  DensePolynomialExt scale(const ScalarField *sx, const ScalarField *sy) const {
    tkm_poly *h = nullptr;
    check(tkm_poly_scale_coeffs(c_->raw(), h_, sx ? sx->bytes() : nullptr, sy ? sy->bytes() : nullptr, &h));
    return DensePolynomialExt(*c_, h);
  }
  const Context *c_;
  tkm_poly *h_ = nullptr;
};
struct DensePolynomialExt::Ruffini {
  DensePolynomialExt q_x, q_y;
  ScalarField remainder;
};
inline DensePolynomialExt::Ruffini DensePolynomialExt::div_by_ruffini(const ScalarField &x, const ScalarField &y) const {
  tkm_poly *qx = nullptr, *qy = nullptr;
  ScalarField r;
  check(tkm_poly_div_by_ruffini(c_->raw(), h_, x.bytes(), y.bytes(), &qx, &qy, r.bytes()));
  return Ruffini{DensePolynomialExt(*c_, qx), DensePolynomialExt(*c_, qy), r};
}
inline DensePolynomialExt operator*(const ScalarField &s, const DensePolynomialExt &p) { return p * s; }
inline DensePolynomialExt operator+(const ScalarField &s, const DensePolynomialExt &p) { return p + s; }
inline DensePolynomialExt operator-(const ScalarField &s, const DensePolynomialExt &p) { return (-p) + s; }

// ---------------------------------------------------------------- PolyExpr (libs/src/bivariate_polynomial/mod.rs:140-260)
// Expression DAG over borrowed polynomials.  evaluate_coeffs walks it with the coefficient-domain operators;
// evaluate_fused(_with_domain) compiles it to a postfix program and hands it to tkm_polyexpr_eval: one forward biNTT per
// distinct leaf (pointer-keyed, like the reference's leaf cache :459-502), ONE pointwise kernel, one inverse biNTT.
class PolyExpr {
 public:
  enum Kind { Poly, Scalar, Add, Sub, Mul, Scale, MulXMinusOne, Sum, PolyOverRoots };
  static PolyExpr poly(const DensePolynomialExt &p) {
    PolyExpr e(Poly);
    e.n_->p = &p;
    return e;
  }
  // p(X / w_mx, Y / w_my), w_m the primitive m-th root of unity (m a power of two; 0 = axis not scaled): what
  // scale_coeffs_x / _y by an inverse root produce (r(X/w, Y), r(X/w, Y/w) of prove2, prove/src/lib.rs:2110-2146) as a leaf that
  // shares p's transform (TKM_PEX_LEAF_SHIFT: p's evaluation table read rotated).  An addition to the reference's constructors.
  static PolyExpr poly_over_roots(const DensePolynomialExt &p, uint64_t mx, uint64_t my) {
    if ((mx & (mx - 1)) || (my & (my - 1))) throw std::runtime_error("the root's order must be a power of two");
    PolyExpr e(PolyOverRoots);
    e.n_->p = &p;
    e.n_->mx = mx;
    e.n_->my = my;
    return e;
  }
  static PolyExpr scalar(const ScalarField &s) {
    PolyExpr e(Scalar);
    e.n_->s = s;
    return e;
  }
  static PolyExpr add(const PolyExpr &l, const PolyExpr &r) { return binary(Add, l, r); }
  static PolyExpr sub(const PolyExpr &l, const PolyExpr &r) { return binary(Sub, l, r); }
  static PolyExpr mul(const PolyExpr &l, const PolyExpr &r) { return binary(Mul, l, r); }
  static PolyExpr scale(const ScalarField &s, const PolyExpr &x) {
    PolyExpr e(Scale);
    e.n_->s = s;
    e.n_->kids = {x.n_};
    return e;
  }
  static PolyExpr mul_x_minus_one(const PolyExpr &x) {
    PolyExpr e(MulXMinusOne);
    e.n_->kids = {x.n_};
    return e;
  }
  static PolyExpr weighted_sum(const std::vector<std::pair<ScalarField, PolyExpr>> &terms) {
    PolyExpr e(Sum);
    for (const auto &t : terms) e.n_->kids.push_back(scale(t.first, t.second).n_);
    return e;
  }

  // evaluate_coeffs (:190-218)
  DensePolynomialExt evaluate_coeffs(const Context &c) const { return coeffs(*n_, c); }
  // degree bound (:262-309); (-1, -1) = the zero polynomial
  std::pair<int64_t, int64_t> degree_bound() const { return bound(*n_); }
  // evaluate_fused (:220-225) / evaluate_fused_with_domain (:227-260)
  DensePolynomialExt evaluate_fused(const Context &c) const {
    const auto d = degree_bound();
    return evaluate_fused_with_domain(c, domain_size_for_degree(d.first), domain_size_for_degree(d.second));
  }
  DensePolynomialExt evaluate_fused_with_domain(const Context &c, size_t target_x_size, size_t target_y_size) const {
    if ((target_x_size & (target_x_size - 1)) || (target_y_size & (target_y_size - 1)) || !target_x_size || !target_y_size)
      throw std::runtime_error("Fused polynomial expression domains must be powers of two.");
    const auto d = degree_bound();
    if (domain_size_for_degree(d.first) > target_x_size || domain_size_for_degree(d.second) > target_y_size)
      throw std::runtime_error("Fused polynomial expression domain is too small for the expression degree.");
    Program pr;
    emit(*n_, pr);
    std::vector<const tkm_poly *> hs;
    for (const DensePolynomialExt *p : pr.leaves) hs.push_back(p->raw());
    if (pr.consts.empty()) pr.consts.push_back(ScalarField::zero());
    tkm_poly *h = nullptr;
    check(tkm_polyexpr_eval(c.raw(), hs.empty() ? nullptr : hs.data(), (uint32_t)hs.size(), pr.ops.data(), (uint32_t)pr.ops.size(), pr.consts[0].bytes(),
                            (uint32_t)pr.consts.size(), target_x_size, target_y_size, &h));
    return DensePolynomialExt(c, h);
  }
  static size_t domain_size_for_degree(int64_t degree) {  // :438-444
    if (degree < 0) return 1;
    size_t n = 1;
    while (n < (size_t)degree + 1) n <<= 1;
    return n;
  }

 private:
  struct Node {
    Kind k;
    const DensePolynomialExt *p = nullptr;
    ScalarField s = ScalarField::zero();
    uint64_t mx = 0, my = 0;  // PolyOverRoots
    std::vector<std::shared_ptr<Node>> kids;
  };
  struct Program {
    std::vector<const DensePolynomialExt *> leaves;
    std::vector<ScalarField> consts;
    std::vector<uint32_t> ops;
    uint32_t leaf(const DensePolynomialExt *p) {
      for (size_t i = 0; i < leaves.size(); i++)
        if (leaves[i] == p) return (uint32_t)i;
      leaves.push_back(p);
      return (uint32_t)leaves.size() - 1;
    }
    uint32_t konst(const ScalarField &s) {
      for (size_t i = 0; i < consts.size(); i++)
        if (consts[i] == s) return (uint32_t)i;
      consts.push_back(s);
      return (uint32_t)consts.size() - 1;
    }
  };
  explicit PolyExpr(Kind k) : n_(std::make_shared<Node>()) { n_->k = k; }
  static PolyExpr binary(Kind k, const PolyExpr &l, const PolyExpr &r) {
    PolyExpr e(k);
    e.n_->kids = {l.n_, r.n_};
    return e;
  }
  static void emit(const Node &n, Program &pr) {
    switch (n.k) {
      case Poly: pr.ops.push_back(TKM_PEX_LEAF | pr.leaf(n.p) << 8); break;
      case PolyOverRoots: {
        auto field = [](uint64_t m) { uint32_t f = 0; while (m) { f++; m >>= 1; } return f; };  // log2(m) + 1, 0 = not scaled
        pr.ops.push_back(TKM_PEX_LEAF_SHIFT | (pr.leaf(n.p) | field(n.mx) << 4 | field(n.my) << 10) << 8);
        break;
      }
      case Scalar: pr.ops.push_back(TKM_PEX_CONST | pr.konst(n.s) << 8); break;
      case Add: case Sub: case Mul:
        emit(*n.kids[0], pr);
        emit(*n.kids[1], pr);
        pr.ops.push_back(n.k == Add ? TKM_PEX_ADD : n.k == Sub ? TKM_PEX_SUB : TKM_PEX_MUL);
        break;
      case Scale:
        emit(*n.kids[0], pr);
        if (n.s != ScalarField::one()) pr.ops.push_back(TKM_PEX_SCALE | pr.konst(n.s) << 8);
        break;
      case MulXMinusOne:
        emit(*n.kids[0], pr);
        pr.ops.push_back(TKM_PEX_XM1);
        break;
      case Sum:
        if (n.kids.empty()) pr.ops.push_back(TKM_PEX_CONST | pr.konst(ScalarField::zero()) << 8);
        for (size_t i = 0; i < n.kids.size(); i++) {
          emit(*n.kids[i], pr);
          if (i) pr.ops.push_back(TKM_PEX_ADD);
        }
        break;
    }
  }
  static DensePolynomialExt coeffs(const Node &n, const Context &c) {
    switch (n.k) {
      case Poly: return DensePolynomialExt(*n.p);
      case PolyOverRoots: {
        DensePolynomialExt q(*n.p);
        if (n.mx) q = q.scale_coeffs_x(c.get_root_of_unity(n.mx).inv());
        if (n.my) q = q.scale_coeffs_y(c.get_root_of_unity(n.my).inv());
        return q;
      }
      case Scalar: return DensePolynomialExt::from_coeffs(c, {n.s}, 1, 1);
      case Add: return coeffs(*n.kids[0], c) + coeffs(*n.kids[1], c);
      case Sub: return coeffs(*n.kids[0], c) - coeffs(*n.kids[1], c);
      case Mul: return coeffs(*n.kids[0], c) * coeffs(*n.kids[1], c);
      case Scale: return coeffs(*n.kids[0], c) * n.s;
      case MulXMinusOne: {
        DensePolynomialExt p = coeffs(*n.kids[0], c);
        return p.mul_monomial(1, 0) - p;
      }
      case Sum: {
        if (n.kids.empty()) return DensePolynomialExt::zero(c);
        DensePolynomialExt acc = coeffs(*n.kids[0], c);
        for (size_t i = 1; i < n.kids.size(); i++) acc += coeffs(*n.kids[i], c);
        return acc;
      }
    }
    throw std::logic_error("unreachable");
  }
  static std::pair<int64_t, int64_t> bound(const Node &n) {
    switch (n.k) {
      case Poly: case PolyOverRoots: return n.p->find_degree();
      case Scalar: return n.s == ScalarField::zero() ? std::make_pair<int64_t, int64_t>(-1, -1) : std::make_pair<int64_t, int64_t>(0, 0);
      case Add: case Sub: {
        auto l = bound(*n.kids[0]), r = bound(*n.kids[1]);
        return {std::max(l.first, r.first), std::max(l.second, r.second)};
      }
      case Mul: {
        auto l = bound(*n.kids[0]), r = bound(*n.kids[1]);
        if (l.first < 0 || l.second < 0 || r.first < 0 || r.second < 0) return {-1, -1};
        return {l.first + r.first, l.second + r.second};
      }
      case Scale: return n.s == ScalarField::zero() ? std::make_pair<int64_t, int64_t>(-1, -1) : bound(*n.kids[0]);
      case MulXMinusOne: {
        auto d = bound(*n.kids[0]);
        if (d.first < 0 || d.second < 0) return {-1, -1};
        return {d.first + 1, d.second};
      }
      case Sum: {
        std::pair<int64_t, int64_t> out{-1, -1};
        for (const auto &k : n.kids) {
          auto d = bound(*k);
          out = {std::max(out.first, d.first), std::max(out.second, d.second)};
        }
        return out;
      }
    }
    throw std::logic_error("unreachable");
  }
  std::shared_ptr<Node> n_;
};

// ---------------------------------------------------------------- G1serde ops (libs/src/group_structures/mod.rs:888-947)
inline G1Affine g1_add(const Context &c, const G1Affine &a, const G1Affine &b) {
  G1Affine r;
  check(tkm_g1_add(c.raw(), a.b, b.b, r.b));
  return r;
}
inline G1Affine g1_mul(const Context &c, const G1Affine &a, const ScalarField &k) {
  G1Affine r;
  check(tkm_g1_mul(c.raw(), a.b, k.bytes(), r.b));
  return r;
}
inline G1Affine g1_neg(const G1Affine &a) {  // (x, -y); the identity stays (0, 0)
  static const uint64_t Q[6] = {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull, 0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};
  G1Affine r = a;
  uint64_t y[6];
  std::memcpy(y, a.b + 48, 48);
  uint64_t any = 0;
  for (int i = 0; i < 6; i++) any |= y[i];
  if (!any) return r;
  unsigned __int128 borrow = 0;
  for (int i = 0; i < 6; i++) {
    unsigned __int128 d = (unsigned __int128)Q[i] - y[i] - borrow;
    y[i] = (uint64_t)d;
    borrow = (d >> 64) & 1;
  }
  std::memcpy(r.b + 48, y, 48);
  return r;
}
inline G1Affine g1_sub(const Context &c, const G1Affine &a, const G1Affine &b) { return g1_add(c, a, g1_neg(b)); }
// msm_g1_bases (:127-143): empty input -> identity
inline G1Affine msm_g1_bases(const Context &c, const std::vector<G1Affine> &bases, const std::vector<ScalarField> &scalars) {
  if (bases.size() != scalars.size()) throw std::runtime_error("Mismatch between the numbers of bases and scalars");
  G1Affine r = G1Affine::zero();
  if (!bases.empty()) check(tkm_msm_g1_host(c.raw(), scalars[0].bytes(), bases[0].b, bases.size(), r.b));
  return r;
}

// ---------------------------------------------------------------- vector_operations (libs/src/vector_operations/mod.rs:19-141,639-693)
namespace vector_operations {
inline std::vector<ScalarField> pointwise(const Context &c, int32_t op, const std::vector<ScalarField> &a, const std::vector<ScalarField> &b) {
  if (a.size() != b.size()) throw std::runtime_error("Mismatch of sizes of vectors to be pointwise operated");
  std::vector<ScalarField> out(a.size());
  if (!a.empty()) check(tkm_fr_vec_op_host(c.raw(), op, a[0].bytes(), b[0].bytes(), out[0].bytes(), a.size()));
  return out;
}
inline std::vector<ScalarField> point_mul_two_vecs(const Context &c, const std::vector<ScalarField> &a, const std::vector<ScalarField> &b) {
  return pointwise(c, TKM_OP_MUL, a, b);
}
inline std::vector<ScalarField> point_div_two_vecs(const Context &c, const std::vector<ScalarField> &a, const std::vector<ScalarField> &b) {
  return pointwise(c, TKM_OP_DIV, a, b);
}
inline std::vector<ScalarField> point_add_two_vecs(const Context &c, const std::vector<ScalarField> &a, const std::vector<ScalarField> &b) {
  return pointwise(c, TKM_OP_ADD, a, b);
}
// transpose_inplace (:139-141): rows x cols row-major -> cols x rows
inline void transpose_inplace(std::vector<ScalarField> &v, size_t rows, size_t cols) {
  std::vector<ScalarField> out(v.size());
  for (size_t i = 0; i < rows; i++)
    for (size_t j = 0; j < cols; j++) out[j * rows + i] = v[i * cols + j];
  v.swap(out);
}
// resize (:639-672): copy the overlapping rectangle of a rows x cols matrix into target_rows x target_cols, zero elsewhere
inline std::vector<ScalarField> resize(const std::vector<ScalarField> &m, size_t rows, size_t cols, size_t target_rows, size_t target_cols) {
  std::vector<ScalarField> out(target_rows * target_cols, ScalarField::zero());
  for (size_t i = 0; i < std::min(rows, target_rows); i++)
    for (size_t j = 0; j < std::min(cols, target_cols); j++) out[i * target_cols + j] = m[i * cols + j];
  return out;
}
}  // namespace vector_operations

// ---------------------------------------------------------------- Sigma1 (xy_powers resident on the device)
class Sigma1 {
 public:
  Sigma1(const Context &c, const std::vector<G1Affine> &xy_powers, size_t rs_x, size_t rs_y) : c_(&c) {
    if (xy_powers.size() != rs_x * rs_y) throw std::runtime_error("xy_powers size mismatch");
    check(tkm_crs_upload(c.raw(), xy_powers[0].b, rs_x, rs_y, &h_));
  }
  ~Sigma1() {
    if (h_) tkm_crs_free(c_->raw(), h_);
  }
  Sigma1(const Sigma1 &) = delete;
  Sigma1 &operator=(const Sigma1 &) = delete;
  // A second resident table (gamma_inv_o_inst, eta_inv_li_o_inter_alpha4_kj, delta_inv_li_o_prv: group_structures/mod.rs:
  // 361-394) for the sparse encoders; rows x cols like the reference's boxed 2-D arrays.
  Sigma1(const Context &c, const G1Affine *points, size_t rows, size_t cols) : c_(&c) { check(tkm_crs_upload(c.raw(), points[0].b, rows, cols, &h_)); }
  // msm_g1_bases over gathered entries of this table (encode_o_pub_fix_common / encode_o_pub_free_common /
  // encode_statement_common, group_structures/mod.rs:145-300): sum_k scalars[k] * table[idx[k]]
  G1Affine msm_indexed(const std::vector<ScalarField> &scalars, const std::vector<uint32_t> &idx) const {
    if (scalars.size() != idx.size()) throw std::runtime_error("Mismatch between the numbers of bases and scalars");
    G1Affine r = G1Affine::zero();
    if (scalars.empty()) return r;
    void *ds = nullptr, *di = nullptr, *dt = nullptr;
    size_t rows, cols;
    check(tkm_crs_device_ptr(h_, &dt, &rows, &cols));
    for (uint32_t i : idx)
      if (i >= rows * cols) throw std::runtime_error("CRS index out of range");
    check(tkm_dev_alloc(c_->raw(), scalars.size() * 32, &ds));
    check(tkm_dev_alloc(c_->raw(), idx.size() * 4, &di));
    check(tkm_memcpy_h2d(c_->raw(), ds, scalars[0].bytes(), scalars.size() * 32));
    check(tkm_memcpy_h2d(c_->raw(), di, idx.data(), idx.size() * 4));
    const int32_t st = tkm_msm_g1_indexed(c_->raw(), ds, 0, dt, di, scalars.size(), r.b);
    tkm_dev_free(c_->raw(), ds);
    tkm_dev_free(c_->raw(), di);
    check(st);
    return r;
  }
  // encode_poly(&mut poly): may shrink the polynomial (optimize_size), panics if the CRS is too small
  G1Affine encode_poly(DensePolynomialExt &poly) const {
    G1Affine r;
    check(tkm_poly_commit(c_->raw(), poly.raw(), h_, r.b));
    return r;
  }

 private:
  const Context *c_;
  tkm_crs *h_ = nullptr;
};

}  // namespace tokamak_b200
