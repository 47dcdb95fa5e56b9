// Host-side harness: compiles the device arithmetic headers (ff.cuh, g1.cuh) with g++ through their
// host emulation of the PTX carry chains, so the exact algorithm text the kernels run is checked on the
// CPU against the oracle (tests/test_host_arith.py).  Test infrastructure only.
#include <cstring>
#include "../../tokamak-zk-evm_b200/csrc/g1.cuh"
using namespace tkm;
template <class F> static F ld(const uint32_t *p) { F r; memcpy(r.v, p, sizeof(r.v)); return r.to_mont(); }
template <class F> static void st(uint32_t *p, const F &a) { F r = a.from_mont(); memcpy(p, r.v, sizeof(r.v)); }
static G1Affine lda(const uint32_t *p) { G1Affine a; a.x = ld<Fq>(p); a.y = ld<Fq>(p + 12); if (a.x.is_zero() && a.y.is_zero()) return G1Affine::identity(); return a; }
static void sta(uint32_t *p, const G1Affine &a) { st(p, a.x); st(p + 12, a.y); }
extern "C" {
// op: 0 add 1 sub 2 mul 3 inv(a) 4 raw mont mul (no conversion) 5 sqr(a) 6 raw mont sqr
void h_fr_op(int op, const uint32_t *a, const uint32_t *b, uint32_t *o) {
  if (op == 4) { Fr x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32); Fr z = x * y; memcpy(o, z.v, 32); return; }
  if (op == 6) { Fr x; memcpy(x.v, a, 32); Fr z = x.sqr(); memcpy(o, z.v, 32); return; }
  Fr x = ld<Fr>(a), y = ld<Fr>(b);
  Fr z = op == 0 ? x + y : op == 1 ? x - y : op == 2 ? x * y : op == 5 ? x.sqr() : x.inv();
  st(o, z);
}
void h_fq_op(int op, const uint32_t *a, const uint32_t *b, uint32_t *o) {
  if (op == 4) { Fq x, y; memcpy(x.v, a, 48); memcpy(y.v, b, 48); Fq z = x * y; memcpy(o, z.v, 48); return; }
  if (op == 6) { Fq x; memcpy(x.v, a, 48); Fq z = x.sqr(); memcpy(o, z.v, 48); return; }
  Fq x = ld<Fq>(a), y = ld<Fq>(b);
  Fq z = op == 0 ? x + y : op == 1 ? x - y : op == 2 ? x * y : op == 5 ? x.sqr() : x.inv();
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
}
