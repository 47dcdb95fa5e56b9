// BLS12-381 G1 (y^2 = x^3 + 4) point arithmetic in extended Jacobian "XYZZ" coordinates.
//
// Replaces the G1Affine / G1Projective arithmetic the reference reaches through ICICLE
// (libs/src/group_structures/mod.rs:888-947 G1serde ops; msm::msm at iotools/mod.rs:2093-2099).
// A point is (X, Y, ZZ, ZZZ) with x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; ZZ == 0 is the identity.
// Affine points (x, y) use the reference's convention (0, 0) = identity
// (group_structures/mod.rs:889-893).  All coordinates are in Montgomery form on the device.
//
// Formulas: Explicit-Formulas Database, short Weierstrass a = 0, "xyzz": madd-2008-s (8M+2S),
// add-2008-s (12M+2S), dbl-2008-s-1 (6M+4S... 9 products here), mdbl-2008-s-1.
// Every exceptional case (identity operands, P + P, P + (-P)) is handled so results are exact.
#pragma once
#include "ff.cuh"

namespace tkm {

struct alignas(16) G1Affine {
  Fq x, y;
  TKM_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
  TKM_HD static G1Affine identity() { return G1Affine{Fq::zero(), Fq::zero()}; }
  TKM_HD G1Affine neg() const { return is_identity() ? *this : G1Affine{x, y.neg()}; }
};

struct alignas(16) G1Xyzz {
  Fq X, Y, ZZ, ZZZ;
  TKM_HD bool is_identity() const { return ZZ.is_zero(); }
  TKM_HD static G1Xyzz identity() { return G1Xyzz{Fq::zero(), Fq::zero(), Fq::zero(), Fq::zero()}; }
  TKM_HD static G1Xyzz from_affine(const G1Affine &p) {
    if (p.is_identity()) return identity();
    return G1Xyzz{p.x, p.y, Fq::one(), Fq::one()};
  }
  TKM_HD G1Xyzz neg() const { return G1Xyzz{X, Y.neg(), ZZ, ZZZ}; }
};

// 2 * (x, y) for an affine, non-identity point (mdbl-2008-s-1).  y != 0 on G1 (odd prime order).
TKM_HD G1Xyzz g1_mdbl(const G1Affine &p) {
  Fq U = p.y.dbl();
  Fq V = U.sqr();
  Fq W = U * V;
  Fq S = p.x * V;
  Fq XX = p.x.sqr();
  Fq M = XX.dbl() + XX;
  G1Xyzz r;
  r.X = M.sqr() - S.dbl();
  r.Y = Fq::dot2(M, S - r.X, W, p.y.neg());
  r.ZZ = V;
  r.ZZZ = W;
  return r;
}

// 2 * P (dbl-2008-s-1).
TKM_HD G1Xyzz g1_dbl(const G1Xyzz &p) {
  if (p.is_identity()) return p;
  Fq U = p.Y.dbl();
  Fq V = U.sqr();
  Fq W = U * V;
  Fq S = p.X * V;
  Fq XX = p.X.sqr();
  Fq M = XX.dbl() + XX;
  G1Xyzz r;
  r.X = M.sqr() - S.dbl();
  r.Y = Fq::dot2(M, S - r.X, W, p.Y.neg());
  r.ZZ = V * p.ZZ;
  r.ZZZ = W * p.ZZZ;
  return r;
}

#if defined(__CUDACC__)
// Latency-oriented doubling for the single-chain tails (Horner over windows): four lanes hold identical copies
// of P and split the nine products into three dependency levels (2 + 3 + 4), exchanging results with shuffles,
// so a doubling costs three product latencies instead of nine.  Every lane of the warp must call it (groups are
// lanes {4g..4g+3}); all lanes return the full result.
__device__ __forceinline__ Fq g1_bcast4(const Fq &v, int src_sub) {
  Fq r;
  const int src = (threadIdx.x & 28) + src_sub;  // lane id within the warp: (lane & ~3) + src_sub
#pragma unroll
  for (int i = 0; i < Fq::N; i++) r.v[i] = __shfl_sync(0xffffffffu, v.v[i], src);
  return r;
}
__device__ __forceinline__ G1Xyzz g1_dbl_coop4(const G1Xyzz &p) {
  // No early return for the identity: groups of one warp may carry different accumulators (the two GLV chains of
  // k_final) and every lane must reach the full-mask shuffles.  The formulas map the all-zero identity to itself
  // (every product has a zero factor; ZZ3 = V*ZZ = 0 for any representation with ZZ = 0).
  if (__all_sync(0xffffffffu, p.is_identity())) return p;  // warp-uniform: nothing to double anywhere
  const int sub = threadIdx.x & 3;
  const Fq U = p.Y.dbl();
  // level 1: V = U^2 | XX = X^2
  Fq r1 = ((sub & 1) ? p.X : U).sqr();
  const Fq V = g1_bcast4(r1, 0), XX = g1_bcast4(r1, 1);
  const Fq M = XX.dbl() + XX;
  // level 2: W = U*V | S = X*V | MM = M^2
  const Fq a2 = sub == 0 ? U : (sub == 1 ? p.X : M);
  const Fq b2 = sub >= 2 ? M : V;
  Fq r2 = a2 * b2;
  const Fq W = g1_bcast4(r2, 0), S = g1_bcast4(r2, 1), MM = g1_bcast4(r2, 2);
  G1Xyzz r;
  r.X = MM - S.dbl();
  // level 3: M*(S - X3) | W*Y | V*ZZ | W*ZZZ
  const Fq a3 = sub == 0 ? M : (sub == 2 ? V : W);
  const Fq b3 = sub == 0 ? (S - r.X) : (sub == 1 ? p.Y : (sub == 2 ? p.ZZ : p.ZZZ));
  Fq r3 = a3 * b3;
  r.Y = g1_bcast4(r3, 0) - g1_bcast4(r3, 1);
  r.ZZ = g1_bcast4(r3, 2);
  r.ZZZ = g1_bcast4(r3, 3);
  return r;
}
// Affine conversion for replicated inputs (every lane of the warp calls it with the same point): a single latency-bound
// chain, so the inverse is the binary extended Euclid (Fp::inv_bgcd, ~0.8 k shift/subtract steps) rather than the Fermat
// chain of ~570 dependent Montgomery products.  The replicas run the same data-dependent loops: no divergence.
__device__ __forceinline__ G1Affine g1_to_affine_coop(const G1Xyzz &p) {
  if (p.is_identity()) return G1Affine::identity();
  Fq t = (p.ZZ * p.ZZZ).inv_fast();
  Fq zz_inv = t * p.ZZZ;
  Fq zzz_inv = t * p.ZZ;
  return G1Affine{p.X * zz_inv, p.Y * zzz_inv};
}
#endif

// acc += (x, y)  (madd-2008-s: 7 products + 2 squarings + one fused two-product Y3), all exceptional cases handled.
TKM_HD void g1_madd(G1Xyzz &acc, const G1Affine &p) {
  if (p.is_identity()) return;
  if (acc.is_identity()) {
    acc = G1Xyzz{p.x, p.y, Fq::one(), Fq::one()};
    return;
  }
  Fq U2 = p.x * acc.ZZ;
  Fq S2 = p.y * acc.ZZZ;
  Fq Pd = U2 - acc.X;
  Fq Rd = S2 - acc.Y;
  if (Pd.is_zero()) {
    if (Rd.is_zero())
      acc = g1_mdbl(p);
    else
      acc = G1Xyzz::identity();
    return;
  }
  Fq PP = Pd.sqr();
  Fq PPP = Pd * PP;
  Fq Q = acc.X * PP;
  Fq X3 = Rd.sqr() - PPP - Q.dbl();
  acc.ZZ = acc.ZZ * PP;
  acc.ZZZ = acc.ZZZ * PPP;
  acc.Y = Fq::dot2(Rd, Q - X3, acc.Y.neg(), PPP);  // R*(Q - X3) - Y*PPP under one Montgomery reduction
  acc.X = X3;
}

// acc += q  (add-2008-s), all exceptional cases handled.
TKM_HD void g1_add(G1Xyzz &acc, const G1Xyzz &q) {
  if (q.is_identity()) return;
  if (acc.is_identity()) {
    acc = q;
    return;
  }
  Fq U1 = acc.X * q.ZZ;
  Fq U2 = q.X * acc.ZZ;
  Fq S1 = acc.Y * q.ZZZ;
  Fq S2 = q.Y * acc.ZZZ;
  Fq Pd = U2 - U1;
  Fq Rd = S2 - S1;
  if (Pd.is_zero()) {
    if (Rd.is_zero())
      acc = g1_dbl(acc);
    else
      acc = G1Xyzz::identity();
    return;
  }
  Fq PP = Pd.sqr();
  Fq PPP = Pd * PP;
  Fq Q = U1 * PP;
  Fq X3 = Rd.sqr() - PPP - Q.dbl();
  Fq Y3 = Fq::dot2(Rd, Q - X3, S1.neg(), PPP);
  acc.X = X3;
  acc.Y = Y3;
  acc.ZZ = acc.ZZ * q.ZZ * PP;
  acc.ZZZ = acc.ZZZ * q.ZZZ * PPP;
}

// Affine (x, y) = (X/ZZ, Y/ZZZ) with one field inversion; identity -> (0, 0).
TKM_HD G1Affine g1_to_affine(const G1Xyzz &p) {
  if (p.is_identity()) return G1Affine::identity();
  Fq t = (p.ZZ * p.ZZZ).inv();
  Fq zz_inv = t * p.ZZZ;
  Fq zzz_inv = t * p.ZZ;
  return G1Affine{p.X * zz_inv, p.Y * zzz_inv};
}

// The same conversion for a single value (one thread, or replicas): the inverse is one latency-bound chain, so the binary
// extended Euclid (Fp::inv_bgcd) beats the Fermat chain by several times.  Not for lanes holding different values (the trip
// counts depend on the value).
TKM_HD G1Affine g1_to_affine_single(const G1Xyzz &p) {
  if (p.is_identity()) return G1Affine::identity();
  Fq t = (p.ZZ * p.ZZZ).inv_fast();
  Fq zz_inv = t * p.ZZZ;
  Fq zzz_inv = t * p.ZZ;
  return G1Affine{p.X * zz_inv, p.Y * zzz_inv};
}

// k * P by left-to-right double-and-add, k canonical (non-Montgomery) little-endian limbs.
// Used for the handful of single scalar multiplications of the prover
// (G1serde * ScalarField, group_structures/mod.rs:929-947).
TKM_HD G1Xyzz g1_mul_scalar(const G1Affine &p, const uint32_t *k, int nlimbs) {
  G1Xyzz acc = G1Xyzz::identity();
  for (int i = nlimbs * 32 - 1; i >= 0; i--) {
    acc = g1_dbl(acc);
    if ((k[i >> 5] >> (i & 31)) & 1) g1_madd(acc, p);
  }
  return acc;
}

}  // namespace tkm
