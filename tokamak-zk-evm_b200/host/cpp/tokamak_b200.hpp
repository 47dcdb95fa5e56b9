// Host-side C++ mirror of the reference's `libs` API over the C-ABI of libtokamak_b200 (include/tokamak_b200.h).
//
// The reference's host code is Rust (absent from this build image); this header is the compiled-language counterpart of
// the Rust shim sketched in INTEGRATION.md: same names and argument meaning as libs::bivariate_polynomial::
// DensePolynomialExt (libs/src/bivariate_polynomial/mod.rs:1283-1416), ScalarField, Sigma1::encode_poly
// (libs/src/group_structures/mod.rs:59-119) and msm_g1_bases (:127-143), RAII ownership like the reference's DeviceVec-
// backed values, `&`-style operators that never mutate their inputs, and panics mapped to exceptions (every non-zero
// status throws std::runtime_error carrying tkm_last_error()).  Header-only; link with -ltokamak_b200.
#pragma once
#include <cstdint>
#include <cstring>
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
