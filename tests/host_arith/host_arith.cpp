// Host-side harness: compiles the device arithmetic headers (ff.cuh, g1.cuh) with g++ through their
// host emulation of the PTX carry chains, so the exact algorithm text the kernels run is checked on the
// CPU against the oracle (tests/test_host_arith.py).  Test infrastructure only.
#include <cstring>
#include "../../tokamak-zk-evm_b200/csrc/g1.cuh"
#include "../../tokamak-zk-evm_b200/csrc/glv.cuh"
using namespace tkm;
template <class F> static F ld(const uint32_t *p) { F r; memcpy(r.v, p, sizeof(r.v)); return r.to_mont(); }
template <class F> static void st(uint32_t *p, const F &a) { F r = a.from_mont(); memcpy(p, r.v, sizeof(r.v)); }
static G1Affine lda(const uint32_t *p) { G1Affine a; a.x = ld<Fq>(p); a.y = ld<Fq>(p + 12); if (a.x.is_zero() && a.y.is_zero()) return G1Affine::identity(); return a; }
static void sta(uint32_t *p, const G1Affine &a) { st(p, a.x); st(p + 12, a.y); }
extern "C" {
// op: 0 add 1 sub 2 mul 3 inv(a) 4 raw mont mul (no conversion) 5 sqr(a) 6 raw mont sqr 7 inv_bgcd(a) 8 inv_fast(a) 9 (Fq) inv_pornin(a) alone, 0 when its own check fails
void h_fr_op(int op, const uint32_t *a, const uint32_t *b, uint32_t *o) {
  if (op == 4) { Fr x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32); Fr z = x * y; memcpy(o, z.v, 32); return; }
  if (op == 6) { Fr x; memcpy(x.v, a, 32); Fr z = x.sqr(); memcpy(o, z.v, 32); return; }
  Fr x = ld<Fr>(a), y = ld<Fr>(b);
  Fr z = op == 0 ? x + y : op == 1 ? x - y : op == 2 ? x * y : op == 5 ? x.sqr() : op == 7 ? x.inv_bgcd() : op == 8 ? x.inv_fast() : x.inv();
  st(o, z);
}
void h_fq_op(int op, const uint32_t *a, const uint32_t *b, uint32_t *o) {
  if (op == 4) { Fq x, y; memcpy(x.v, a, 48); memcpy(y.v, b, 48); Fq z = x * y; memcpy(o, z.v, 48); return; }
  if (op == 6) { Fq x; memcpy(x.v, a, 48); Fq z = x.sqr(); memcpy(o, z.v, 48); return; }
  Fq x = ld<Fq>(a), y = ld<Fq>(b);
  Fq z = op == 0 ? x + y : op == 1 ? x - y : op == 2 ? x * y : op == 5 ? x.sqr() : op == 7 ? x.inv_bgcd() : op == 8 ? x.inv_fast() : op == 9 ? [&] { bool ok; Fq r = x.inv_pornin(&ok); return ok ? r : Fq::zero(); }() : x.inv();
  st(o, z);
}
// acc = sum of n affine points via madd into an XYZZ accumulator (exercises all branches)
void h_g1_sum_madd(const uint32_t *pts, int n, uint32_t *o) {
  G1Xyzz acc = G1Xyzz::identity();
  for (int i = 0; i < n; i++) g1_madd(acc, lda(pts + 24 * i));
  sta(o, g1_to_affine(acc));
}
// tree sum with full XYZZ adds (exercises g1_add / g1_dbl): ((p0+p1)+(p2+p3))...
void h_g1_sum_add(const uint32_t *pts, int n, uint32_t *o) {
  G1Xyzz acc = G1Xyzz::identity();
  for (int i = 0; i + 1 < n; i += 2) {
    G1Xyzz t = G1Xyzz::from_affine(lda(pts + 24 * i));
    g1_madd(t, lda(pts + 24 * (i + 1)));
    g1_add(acc, t);
  }
  if (n & 1) g1_add(acc, G1Xyzz::from_affine(lda(pts + 24 * (n - 1))));
  sta(o, g1_to_affine(acc));
}
void h_g1_mul(const uint32_t *pt, const uint32_t *k, uint32_t *o) { sta(o, g1_to_affine(g1_mul_scalar(lda(pt), k, 8))); }
void h_g1_dbl_n(const uint32_t *pt, int n, uint32_t *o) {
  G1Xyzz a = G1Xyzz::from_affine(lda(pt));
  for (int i = 0; i < n; i++) a = g1_dbl(a);
  sta(o, g1_to_affine(a));
}
// GLV split: out[0..3] = |k1|, out[4..7] = |k2|, out[8] = sign k1, out[9] = sign k2
void h_glv_split(const uint32_t *k, uint32_t *out) {
  GlvSplit s = glv_split(k);
  for (int i = 0; i < 4; i++) { out[i] = s.mag[0][i]; out[4 + i] = s.mag[1][i]; }
  out[8] = s.neg[0]; out[9] = s.neg[1];
}
// The digit stream k_decompose emits for one scalar (same helper calls): 2*wh signed digits (GLV) or wd digits.
// digits[i] = signed value of window i; returns the number of windows, or -1 if a carry is left over.
int h_msm_digits(const uint32_t *k, uint32_t c, int use_glv, int32_t *digits) {
  if (use_glv) {
    const uint32_t wh = (128 + c - 1) / c;
    GlvSplit s = glv_split(k);
    for (int h = 0; h < 2; h++) {
      uint32_t carry = 0;
      for (uint32_t w = 0; w < wh; w++) {
        uint32_t mag, neg;
        signed_digit(s.mag[h], 4, w, c, carry, mag, neg);
        if (mag) neg ^= s.neg[h];
        digits[h * wh + w] = neg ? -(int32_t)mag : (int32_t)mag;
      }
      if (carry) return -1;
    }
    return (int)(2 * wh);
  }
  const uint32_t wd = (256 + c - 1) / c;
  uint32_t carry = 0;
  for (uint32_t w = 0; w < wd; w++) {
    uint32_t mag, neg;
    signed_digit(k, 8, w, c, carry, mag, neg);
    digits[w] = neg ? -(int32_t)mag : (int32_t)mag;
  }
  return carry ? -1 : (int)wd;
}
// phi(P) = (beta*x, y)
void h_g1_phi(const uint32_t *pt, uint32_t *o) {
  G1Affine a = lda(pt);
  Fq b;
  for (int i = 0; i < 12; i++) b.v[i] = glv::beta(i);
  a.x = a.x * b.to_mont();
  sta(o, a);
}
// raw Montgomery a*b + c*d with one reduction: (ab + cd)/R mod q
void h_fq_dot2(const uint32_t *a, const uint32_t *b, const uint32_t *c, const uint32_t *d, uint32_t *o) {
  Fq x, y, z, w; memcpy(x.v, a, 48); memcpy(y.v, b, 48); memcpy(z.v, c, 48); memcpy(w.v, d, 48);
  Fq r = Fq::dot2(x, y, z, w); memcpy(o, r.v, 48);
}
}
