// GLV split of a BLS12-381 scalar and the signed-digit recoding of the MSM (host + device text).
//
// BLS12-381 G1 has the endomorphism phi(x, y) = (beta*x, y) = lambda*(x, y) with lambda = z^2 - 1
// (z = -0xd201000000010000, the curve parameter) and r = lambda^2 + lambda + 1.  Writing a scalar as
// k = k1 + k2*lambda (mod r) with |k1|, |k2| < 2^127 turns sum k_i P_i into sum k1_i P_i + phi(sum k2_i P_i):
// the same number of bucket additions (2 x 128 bits instead of 256 bits of digits), but the serial tail of
// the MSM (Horner over windows, msm.cu k_final) becomes two independent chains of half the depth, and phi is
// applied once to the second chain's sum.  The MSM value is unchanged (it is a group element), so this is
// invisible at the boundary: msm::msm(..) as called at libs/src/iotools/mod.rs:2093-2099 and
// libs/src/group_structures/mod.rs:108-114,135-141 returns the same point.
//
// Everything here is plain integer code (no carry-chain PTX): it runs once per scalar in k_decompose, which is
// well under 1 % of an MSM, and is unit-tested on the CPU through tests/host_arith (tests/test_host_arith.py).
#pragma once
#include <cstdint>

#include "ff.cuh"

namespace tkm {

struct GlvSplit {
  uint32_t mag[2][4];  // |k1|, |k2| < 2^127, little-endian limbs
  uint32_t neg[2];     // 1 if k1 (k2) is negative
};

namespace glv {
// lambda = 0xac45a4010001a40200000000ffffffff
TKM_HD constexpr uint32_t lambda(int i) {
  constexpr uint32_t L[5] = {0xffffffffu, 0x00000000u, 0x0001a402u, 0xac45a401u, 0u};
  return L[i];
}
// mu = floor(2^256 / lambda) = 0x1_7c6becf1_e01faadd_63f6e522_f6cfee30 (129 bits)
TKM_HD constexpr uint32_t mu(int i) {
  constexpr uint32_t M[5] = {0xf6cfee30u, 0x63f6e522u, 0xe01faaddu, 0x7c6becf1u, 0x00000001u};
  return M[i];
}
// floor(lambda / 2) and floor((lambda + 1) / 2)
TKM_HD constexpr uint32_t half_lambda(int i) {
  constexpr uint32_t H[5] = {0x7fffffffu, 0x00000000u, 0x8000d201u, 0x5622d200u, 0u};
  return H[i];
}
TKM_HD constexpr uint32_t half_lambda1(int i) {
  constexpr uint32_t H[5] = {0x80000000u, 0x00000000u, 0x8000d201u, 0x5622d200u, 0u};
  return H[i];
}
// beta with phi(P) = (beta*x, y) = lambda*P, canonical limbs (a primitive cube root of unity in Fq)
TKM_HD constexpr uint32_t beta(int i) {
  constexpr uint32_t Bq[12] = {0x0000aaacu, 0x8bfd0000u, 0x4f49fffdu, 0x409427ebu, 0x0fb85f9bu, 0x897d2965u,
                               0x89759ad4u, 0xaa0d857du, 0x63d4de85u, 0xec024086u, 0x397fe699u, 0x1a0111eau};
  return Bq[i];
}

// 160-bit two's-complement helpers (5 limbs): a += b, a -= b, sign, magnitude.
TKM_HD void add5(uint32_t *a, const uint32_t *b) {
  uint64_t c = 0;
  for (int i = 0; i < 5; i++) {
    c += (uint64_t)a[i] + b[i];
    a[i] = (uint32_t)c;
    c >>= 32;
  }
}
TKM_HD void sub5(uint32_t *a, const uint32_t *b) {
  uint64_t br = 0;
  for (int i = 0; i < 5; i++) {
    uint64_t d = (uint64_t)a[i] - b[i] - br;
    a[i] = (uint32_t)d;
    br = (d >> 63) & 1;
  }
}
// a > b for non-negative 5-limb values
TKM_HD bool gt5(const uint32_t *a, const uint32_t *b) {
  for (int i = 4; i >= 0; i--) {
    if (a[i] != b[i]) return a[i] > b[i];
  }
  return false;
}
}  // namespace glv

// k (8 limbs, canonical) -> (k1, k2) with k1 + k2*lambda = k (mod r), |k1| <= lambda/2 + 1, |k2| <= lambda/2 + 2.
TKM_HD GlvSplit glv_split(const uint32_t *k_in) {
  using namespace glv;
  // The boundary promises canonical scalars; a value in [r, 2^256) is still folded into [0, r) first (2^256 < 3r), so the
  // bounds below hold for any 256-bit input and the result is the same group element.
  uint32_t k[8];
  for (int i = 0; i < 8; i++) k[i] = k_in[i];
  for (int it = 0; it < 2; it++) {
    uint32_t d[8];
    uint64_t br = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint64_t t = (uint64_t)k[i] - FrParams::mod(i) - br;
      d[i] = (uint32_t)t;
      br = (t >> 63) & 1;
    }
    if (!br)
      for (int i = 0; i < 8; i++) k[i] = d[i];
  }
  // q = floor(k*mu / 2^256): Barrett estimate of floor(k / lambda), at most 2 below it.
  uint32_t prod[13];
  for (int i = 0; i < 13; i++) prod[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
      c += (uint64_t)k[i] * mu(j) + prod[i + j];
      prod[i + j] = (uint32_t)c;
      c >>= 32;
    }
    prod[i + 5] = (uint32_t)c;  // column i+5 is untouched by the earlier rows
  }
  uint32_t q[5] = {prod[8], prod[9], prod[10], prod[11], prod[12]};
  // rem = k - q*lambda, exact modulo 2^160 (the true value is below 3*lambda < 2^130)
  uint32_t ql[5] = {0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
      if (i + j >= 5) break;
      c += (uint64_t)q[i] * lambda(j) + ql[i + j];
      ql[i + j] = (uint32_t)c;
      c >>= 32;
    }
  }
  uint32_t rem[5] = {k[0], k[1], k[2], k[3], k[4]};
  sub5(rem, ql);
  uint32_t lam[5], one[5] = {1, 0, 0, 0, 0}, h0[5], h1[5];
#pragma unroll
  for (int i = 0; i < 5; i++) {
    lam[i] = lambda(i);
    h0[i] = half_lambda(i);
    h1[i] = half_lambda1(i);
  }
  for (int it = 0; it < 3; it++) {
    if (!gt5(lam, rem)) {  // rem >= lambda
      sub5(rem, lam);
      add5(q, one);
    }
  }
  // now k = q*lambda + rem, 0 <= rem < lambda, 0 <= q <= lambda + 1.  Balance with the lattice vectors
  // (1, lambda + 1) [1 + (lambda + 1)*lambda = r] and (lambda, -1).
  if (gt5(q, h1)) {  // q > (lambda + 1)/2: (k1, k2) -= (1, lambda + 1)
    sub5(q, lam);
    sub5(q, one);
    sub5(rem, one);
  }
  const bool rem_neg = (rem[4] >> 31) != 0;  // only -1 is possible here
  if (!rem_neg && gt5(rem, h0)) {  // k1 > lambda/2: (k1, k2) += (-lambda, 1)
    sub5(rem, lam);
    add5(q, one);
  }
  GlvSplit s;
  const uint32_t *src[2] = {rem, q};
  for (int h = 0; h < 2; h++) {
    uint32_t t[5];
    const uint32_t neg = src[h][4] >> 31;
    if (neg) {
      for (int i = 0; i < 5; i++) t[i] = 0;
      sub5(t, src[h]);
    } else {
      for (int i = 0; i < 5; i++) t[i] = src[h][i];
    }
    for (int i = 0; i < 4; i++) s.mag[h][i] = t[i];
    s.neg[h] = neg;
  }
  return s;
}

// Signed c-bit recoding, one digit per call, low window first: digit value in [-(2^(c-1) - 1), 2^(c-1)].
// `limbs` holds `nl` little-endian limbs; windows past the top read zero.  carry: in/out (0 before the first window).
TKM_HD void signed_digit(const uint32_t *limbs, uint32_t nl, uint32_t w, uint32_t c, uint32_t &carry, uint32_t &mag,
                         uint32_t &neg) {
  const uint32_t bit = w * c;
  const uint32_t limb = bit >> 5, sh = bit & 31;
  uint32_t raw = 0;
  if (limb < nl) {
    uint64_t two = limbs[limb];
    if (limb + 1 < nl) two |= (uint64_t)limbs[limb + 1] << 32;
    raw = (uint32_t)(two >> sh) & ((1u << c) - 1);
  }
  raw += carry;
  mag = raw;
  neg = 0;
  carry = 0;
  if (raw > (1u << (c - 1))) {
    mag = (1u << c) - raw;
    neg = 1;
    carry = 1;
  }
}

}  // namespace tkm
