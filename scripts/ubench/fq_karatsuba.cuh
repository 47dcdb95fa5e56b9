// Costed-and-rejected experiment: one level of Karatsuba for the 12-limb Fq product plus a
// reduction-only Montgomery pass.  Correct (fq_karatsuba_test.cpp, host emulation of the carry chains), but its SASS is
// 240 IMAD.WIDE + 73 IMAD + 261 IADD3 + 53 SEL against 279 + 24 + 57 + 12 for Fp::operator*: 5 % fewer FMA-pipe cycles for
// 70 % more instructions.  Not part of the library.
//   g++ -O2 -std=c++17 -o /tmp/kt scripts/ubench/fq_karatsuba_test.cpp && /tmp/kt
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -cubin scripts/ubench/fq_karatsuba_sass.cu && cuobjdump -sass ...
// Measured on B200 (fq_karatsuba_bench.cu, profiles/r01_fq_karatsuba_bench.log): 29.2 G products/s against 30.3 for Fp::operator*.
#pragma once
#include "../../tokamak-zk-evm_b200/csrc/ff.cuh"
namespace tkm {
// t[0..11] = a[0..5] * b[0..5]: even-aligned accumulator E (pairs (0,1),(2,3),..) and odd-aligned O (O[k] = column k+1).
TKM_HD void mul6(const uint32_t *a, const uint32_t *b, uint32_t *t) {
  uint32_t E[12], O[12];
#pragma unroll
  for (int k = 0; k < 12; k++) { E[k] = 0; O[k] = 0; }
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const uint32_t bi = b[i];
    if ((i & 1) == 0) {
      // even limbs of a -> even columns i+j (E at i+j); odd limbs -> odd columns i+j (O at i+j-1)
      E[i] = mad_lo_cc(a[0], bi, E[i]);         E[i + 1] = madc_hi_cc(a[0], bi, E[i + 1]);
      E[i + 2] = madc_lo_cc(a[2], bi, E[i + 2]); E[i + 3] = madc_hi_cc(a[2], bi, E[i + 3]);
      E[i + 4] = madc_lo_cc(a[4], bi, E[i + 4]); E[i + 5] = madc_hi_cc(a[4], bi, E[i + 5]);
      if (i + 6 < 12) E[i + 6] = addc(E[i + 6], 0);
      O[i] = mad_lo_cc(a[1], bi, O[i]);         O[i + 1] = madc_hi_cc(a[1], bi, O[i + 1]);
      O[i + 2] = madc_lo_cc(a[3], bi, O[i + 2]); O[i + 3] = madc_hi_cc(a[3], bi, O[i + 3]);
      O[i + 4] = madc_lo_cc(a[5], bi, O[i + 4]); O[i + 5] = madc_hi_cc(a[5], bi, O[i + 5]);
      if (i + 6 < 12) O[i + 6] = addc(O[i + 6], 0);
    } else {
      // even limbs of a -> odd columns i+j (O at i+j-1); odd limbs -> even columns i+j (E at i+j)
      O[i - 1] = mad_lo_cc(a[0], bi, O[i - 1]);   O[i] = madc_hi_cc(a[0], bi, O[i]);
      O[i + 1] = madc_lo_cc(a[2], bi, O[i + 1]);  O[i + 2] = madc_hi_cc(a[2], bi, O[i + 2]);
      O[i + 3] = madc_lo_cc(a[4], bi, O[i + 3]);  O[i + 4] = madc_hi_cc(a[4], bi, O[i + 4]);
      if (i + 5 < 12) O[i + 5] = addc(O[i + 5], 0);
      E[i + 1] = mad_lo_cc(a[1], bi, E[i + 1]);   E[i + 2] = madc_hi_cc(a[1], bi, E[i + 2]);
      E[i + 3] = madc_lo_cc(a[3], bi, E[i + 3]);  E[i + 4] = madc_hi_cc(a[3], bi, E[i + 4]);
      E[i + 5] = madc_lo_cc(a[5], bi, E[i + 5]);  E[i + 6] = madc_hi_cc(a[5], bi, E[i + 6]);
      if (i + 7 < 12) E[i + 7] = addc(E[i + 7], 0);
    }
  }
  t[0] = E[0];
  t[1] = add_cc(E[1], O[0]);
#pragma unroll
  for (int k = 2; k < 11; k++) t[k] = addc_cc(E[k], O[k - 1]);
  t[11] = addc(E[11], O[10]);
}
// t[0..23] = a * b by one level of Karatsuba: z0 = a0 b0, z2 = a1 b1, zm = (a0 + a1)(b0 + b1) - z0 - z2.
TKM_HD void mul12_karatsuba(const uint32_t *a, const uint32_t *b, uint32_t *t) {
  uint32_t z0[12], z2[12], sa[6], sb[6], zm[13];
  mul6(a, b, z0);
  mul6(a + 6, b + 6, z2);
  sa[0] = add_cc(a[0], a[6]);
#pragma unroll
  for (int k = 1; k < 6; k++) sa[k] = addc_cc(a[k], a[k + 6]);
  const uint32_t ca = addc(0, 0);
  sb[0] = add_cc(b[0], b[6]);
#pragma unroll
  for (int k = 1; k < 6; k++) sb[k] = addc_cc(b[k], b[k + 6]);
  const uint32_t cb = addc(0, 0);
  mul6(sa, sb, zm);
  // (sa + ca 2^192)(sb + cb 2^192) = sa sb + (ca sb + cb sa) 2^192 + ca cb 2^384
  const uint32_t ma = 0u - ca, mb = 0u - cb;
  zm[6] = add_cc(zm[6], sb[0] & ma);
#pragma unroll
  for (int k = 1; k < 6; k++) zm[6 + k] = addc_cc(zm[6 + k], sb[k] & ma);
  zm[12] = addc(ca & cb, 0);
  zm[6] = add_cc(zm[6], sa[0] & mb);
#pragma unroll
  for (int k = 1; k < 6; k++) zm[6 + k] = addc_cc(zm[6 + k], sa[k] & mb);
  zm[12] = addc(zm[12], 0);
  // zm -= z0 ; zm -= z2   (the middle term is non-negative and below 2^(32*13))
  zm[0] = sub_cc(zm[0], z0[0]);
#pragma unroll
  for (int k = 1; k < 12; k++) zm[k] = subc_cc(zm[k], z0[k]);
  zm[12] = subc(zm[12], 0);
  zm[0] = sub_cc(zm[0], z2[0]);
#pragma unroll
  for (int k = 1; k < 12; k++) zm[k] = subc_cc(zm[k], z2[k]);
  zm[12] = subc(zm[12], 0);
  // t = z0 + z2 2^384 + zm 2^192
#pragma unroll
  for (int k = 0; k < 6; k++) t[k] = z0[k];
  t[6] = add_cc(z0[6], zm[0]);
#pragma unroll
  for (int k = 1; k < 6; k++) t[6 + k] = addc_cc(z0[6 + k], zm[k]);
#pragma unroll
  for (int k = 0; k < 6; k++) t[12 + k] = addc_cc(z2[k], zm[6 + k]);
  t[18] = addc_cc(z2[6], zm[12]);
#pragma unroll
  for (int k = 7; k < 11; k++) t[12 + k] = addc_cc(z2[k], 0);
  t[23] = addc(z2[11], 0);
}
// Montgomery reduction of a 24-limb value below p * 2^384: the row structure of Fp::operator* with the products removed;
// row i injects the limb (row 11: the two limbs) of t that enters the 13-column window.
TKM_HD Fq redc24(const uint32_t *t) {
  constexpr int N = 12;
  uint32_t X[N], Y[N];
#pragma unroll
  for (int k = 0; k < N; k++) { X[k] = t[k]; Y[k] = 0; }
  Fq::reduce_row(X, Y);
#pragma unroll
  for (int i = 1; i < N; i++) {
    uint32_t *E = (i & 1) ? X : Y;
    uint32_t *O = (i & 1) ? Y : X;
    O[0] = add_cc(O[0], E[1]);
#pragma unroll
    for (int j = 1; j < N - 1; j += 2) {
      E[j - 1] = addc_cc(E[j + 1], 0);
      E[j] = addc_cc(E[j + 2], 0);
    }
    E[N - 2] = addc_cc(t[N - 1 + i], 0);
    E[N - 1] = addc(i == N - 1 ? t[2 * N - 1] : 0u, 0);
    Fq::reduce_row(O, E);
  }
  Fq r;
  r.v[0] = add_cc(X[0], Y[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) r.v[k] = addc_cc(X[k], Y[k + 1]);
  r.v[N - 1] = addc(X[N - 1], 0);
  Fq::final_sub(r.v);
  return r;
}
TKM_HD Fq mul_karatsuba(const Fq &a, const Fq &b) {
  uint32_t t[24];
  mul12_karatsuba(a.v, b.v, t);
  return redc24(t);
}
}  // namespace tkm
