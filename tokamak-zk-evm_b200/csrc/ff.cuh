// Montgomery prime fields on 32-bit limbs for sm_100a: BLS12-381 Fr (8 limbs) and Fq (12 limbs).
//
// Replaces the field arithmetic the reference obtains from ICICLE's ScalarField / BaseField
// (icicle-bls12-381 v3.8.0; imports at libs/src/bivariate_polynomial/mod.rs:2-8 and
// libs/src/group_structures/mod.rs:12-15).  Device values are kept in Montgomery form
// (R = 2^(32*N)); the C-ABI converts at the boundary, where the reference's canonical
// little-endian layout is kept.
//
// Multiplication: two column-interleaved accumulators ("even"/"odd" columns) so that every
// 32x32 partial product is one mad.lo.cc/madc.hi.cc pair on an aligned register pair, which
// ptxas fuses into a single IMAD.WIDE.U32.X with predicate carry-in/out (checked with
// cuobjdump -sass): N^2 wide IMADs for a*b plus N^2 for the interleaved reduction.
//
// The carry-chain primitives have a host emulation (thread-local carry flag) so the exact same
// algorithm text is unit-tested on the CPU against the oracle (tests/test_host_arith.py).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define TKM_HD __host__ __device__ __forceinline__
#define TKM_D __device__ __forceinline__
#else
#define TKM_HD inline
#define TKM_D inline
#endif

namespace tkm {

// ---------------------------------------------------------------- carry-chain primitives
#if defined(__CUDA_ARCH__)
TKM_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
TKM_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
TKM_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
TKM_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
TKM_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
TKM_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
TKM_D uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
TKM_D uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
TKM_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
TKM_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
TKM_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
TKM_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
// Host emulation of the PTX condition-code register.
inline uint32_t &cc_flag() { static thread_local uint32_t cc = 0; return cc; }
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; cc_flag() = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + cc_flag(); cc_flag() = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + cc_flag(); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b; cc_flag() = (uint32_t)(s >> 63); return (uint32_t)s; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b - cc_flag(); cc_flag() = (uint32_t)(s >> 63); return (uint32_t)s; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - cc_flag(); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)(a * b) + c; cc_flag() = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)(a * b) + c + cc_flag(); cc_flag() = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (((uint64_t)a * b) >> 32) + c + cc_flag(); cc_flag() = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)((((uint64_t)a * b) >> 32) + c + cc_flag()); }
#endif

// -x mod 2^32 as a product ptxas cannot turn back into a negation (the factor lives in constant memory); see Fp::reduce_row.
#if defined(__CUDA_ARCH__)
static __constant__ uint32_t c_all_ones = 0xffffffffu;
TKM_D uint32_t neg_opaque(uint32_t x) { return x * c_all_ones; }
#elif defined(__CUDACC__)
static __constant__ uint32_t c_all_ones = 0xffffffffu;
inline uint32_t neg_opaque(uint32_t x) { return 0u - x; }
#else
inline uint32_t neg_opaque(uint32_t x) { return 0u - x; }
#endif

// ---------------------------------------------------------------- field parameters
struct FrParams {
  static constexpr int N = 8;
  static constexpr uint32_t INV = 0xffffffffu;  // -r^-1 mod 2^32
  // r = ... ffffffff 00000001: the two low limbs are 1 and 2^32-1, so m*(p0 + p1*2^32) = m*2^64 - m*2^32 + m
  // needs no multiplier at all (see Fp::reduce_row).
  static constexpr bool LOW_LIMBS_ONE_MINUS_ONE = true;
  // 1: reduction rows as fused multiply-add carry chains started by the borrow of 0 - E[0] (see Fp::reduce_row);
  // 0: the round-1 form (plain 64-bit products folded in with add-with-carry chains on the ALU pipe).
  static constexpr int REDUCE_FORM = 1;
  // r uses 255 of the 256 container bits: the doubled rows of the dedicated squaring (see Fp::sqr) would overflow the
  // 2^(32(N+1)) accumulator bound, so Fr squares through the general product.
  static constexpr bool DEDICATED_SQR = false;
  TKM_HD static constexpr uint32_t mod(int i) {
    constexpr uint32_t M[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
    return M[i];
  }
  TKM_HD static constexpr uint32_t r2(int i) {  // R^2 mod r
    constexpr uint32_t M[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
    return M[i];
  }
  TKM_HD static constexpr uint32_t one(int i) {  // R mod r
    constexpr uint32_t M[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
    return M[i];
  }
};
struct FrParamsAddChains : FrParams {
  static constexpr int REDUCE_FORM = 0;
};
struct FqParams {
  static constexpr int N = 12;
  static constexpr int REDUCE_FORM = 0;
  static constexpr uint32_t INV = 0xfffcfffdu;  // -q^-1 mod 2^32
  static constexpr bool LOW_LIMBS_ONE_MINUS_ONE = false;
  static constexpr bool DEDICATED_SQR = true;  // q < 2^381 leaves three spare bits in the 384-bit container
  TKM_HD static constexpr uint32_t mod(int i) {
    constexpr uint32_t M[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u, 0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
    return M[i];
  }
  TKM_HD static constexpr uint32_t r2(int i) {
    constexpr uint32_t M[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu, 0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
    return M[i];
  }
  TKM_HD static constexpr uint32_t one(int i) {
    constexpr uint32_t M[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u, 0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
    return M[i];
  }
};

// ---------------------------------------------------------------- field element
template <class P>
struct alignas(16) Fp {
  static constexpr int N = P::N;
  uint32_t v[N];

  TKM_HD static Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = 0;
    return r;
  }
  TKM_HD static Fp one() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::one(i);
    return r;
  }
  TKM_HD static Fp r2() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::r2(i);
    return r;
  }
  TKM_HD bool is_zero() const {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < N; i++) acc |= v[i];
    return acc == 0;
  }
  TKM_HD bool operator==(const Fp &o) const {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < N; i++) acc |= v[i] ^ o.v[i];
    return acc == 0;
  }
  TKM_HD bool operator!=(const Fp &o) const { return !(*this == o); }

  // r = t - p if t >= p else t   (t < 2p)
  TKM_HD static void final_sub(uint32_t *t) {
    uint32_t d[N];
    d[0] = sub_cc(t[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < N; i++) d[i] = subc_cc(t[i], P::mod(i));
    uint32_t borrow = subc(0, 0);  // 0xffffffff if t < p
#pragma unroll
    for (int i = 0; i < N; i++) t[i] = borrow ? t[i] : d[i];
  }

  TKM_HD friend Fp operator+(const Fp &a, const Fp &b) {
    Fp r;
    r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
    r.v[N - 1] = addc(a.v[N - 1], b.v[N - 1]);  // both moduli leave headroom in the top limb
    final_sub(r.v);
    return r;
  }
  TKM_HD friend Fp operator-(const Fp &a, const Fp &b) {
    Fp r;
    r.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.v[i] = subc_cc(a.v[i], b.v[i]);
    uint32_t mask = subc(0, 0);  // all-ones if a < b
    r.v[0] = add_cc(r.v[0], P::mod(0) & mask);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(r.v[i], P::mod(i) & mask);
    r.v[N - 1] = addc(r.v[N - 1], P::mod(N - 1) & mask);
    return r;
  }
  TKM_HD Fp neg() const { return zero() - *this; }
  TKM_HD Fp dbl() const { return *this + *this; }

  // ---- Montgomery product.  E = accumulator aligned at column 0, O = aligned at column 1:
  // running total T = E + 2^32 * O.  See the derivation in DESIGN.md ("Field multiplication").
  // (lo,hi)(acc[j],acc[j+1]) += a[j]*b for j = 0,2,..,N-2, one carry chain; returns with CC = carry out.
  TKM_HD static void chain_mad(uint32_t *acc, const uint32_t *a, uint32_t b) {
    acc[0] = mad_lo_cc(a[0], b, acc[0]);
    acc[1] = madc_hi_cc(a[0], b, acc[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      acc[j] = madc_lo_cc(a[j], b, acc[j]);
      acc[j + 1] = madc_hi_cc(a[j], b, acc[j + 1]);
    }
  }
  TKM_HD static void chain_mad_mod(uint32_t *acc, int first, uint32_t m) {
    acc[0] = mad_lo_cc(P::mod(first), m, acc[0]);
    acc[1] = madc_hi_cc(P::mod(first), m, acc[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      acc[j] = madc_lo_cc(P::mod(first + j), m, acc[j]);
      acc[j + 1] = madc_hi_cc(P::mod(first + j), m, acc[j + 1]);
    }
  }
  // Reduction half-row: make column 0 of T vanish.  E aligned at column 0, O at column 1.
  TKM_HD static void reduce_row(uint32_t *E, uint32_t *O) {
    if (P::LOW_LIMBS_ONE_MINUS_ONE && P::REDUCE_FORM == 1) {
      // p0 = 1 and -p^-1 = -1 mod 2^32: m = -E[0], and column 0 becomes E[0] + m = 2^32*[E[0] != 0].  That carry is
      // the carry out of E[0] + 0xffffffff, so that addition (result discarded) leaves it in the carry flag and the odd-limb
      // chain starts with it: (O[0],O[1]) += m*p1 + carry, (O[2],O[3]) += m*p3, ...  (Not sub.cc 0 - E[0]: the hardware
      // carry after a subtraction is the inverted borrow and ptxas hands it to a following madc unchanged.)  The p0 product is never formed:
      // 7 wide IMADs per row instead of 8 (Fr: 64 + 56 = 120 per product) and no separate add-with-carry chains.
      // m is taken as E[0] * 0xffffffff with the factor read from constant memory: when ptxas sees the negation it
      // splits every m*p[j] multiply-add into an IMAD.X + IMAD.HI.U32.X pair (6 FMA-pipe cycles instead of 4); a product
      // it cannot strength-reduce keeps the pairs fused into IMAD.WIDE.U32.X (checked with cuobjdump -sass,
      // profiles/r02_sass_histograms.txt: 280 instructions per NTT butterfly instead of 408).
      const uint32_t m = neg_opaque(E[0]);
      (void)add_cc(E[0], 0xffffffffu);  // carry flag = [E[0] != 0]
      O[0] = madc_lo_cc(P::mod(1), m, O[0]);
      O[1] = madc_hi_cc(P::mod(1), m, O[1]);
#pragma unroll
      for (int j = 3; j < N; j += 2) {
        O[j - 1] = madc_lo_cc(P::mod(j), m, O[j - 1]);
        O[j] = madc_hi_cc(P::mod(j), m, O[j]);
      }
      E[2] = mad_lo_cc(P::mod(2), m, E[2]);
      E[3] = madc_hi_cc(P::mod(2), m, E[3]);
#pragma unroll
      for (int j = 4; j < N; j += 2) {
        E[j] = madc_lo_cc(P::mod(j), m, E[j]);
        E[j + 1] = madc_hi_cc(P::mod(j), m, E[j + 1]);
      }
      O[N - 1] = addc(O[N - 1], 0);
      return;  // E[0] is now (logically) zero and is never read again; E[1] is untouched
    }
    if (P::LOW_LIMBS_ONE_MINUS_ONE) {
      // (round-1 form, kept for the microbenchmark comparison: FrParamsAddChains)
      // p0 = 1, p1 = 2^32 - 1, -p^-1 = -1 mod 2^32:  m = -E[0].  Column 0: E[0] + m = 2^32*[m != 0].
      // Columns (1,2) receive that carry plus m*p1 = m*2^32 - m, i.e. the 64-bit value
      //   V = m ? (m << 32) - (m - 1) : 0   ->  lo = m ? 1 - m : 0,  hi = m - [m > 1]
      // added to the aligned pair (O[0], O[1]); the carry continues into the odd-limb products.
      uint32_t m = 0u - E[0];
      uint32_t lo = m ? (1u - m) : 0u;
      uint32_t hi = m - (m > 1u ? 1u : 0u);
      O[0] = add_cc(O[0], lo);
      O[1] = addc_cc(O[1], hi);
      // The remaining products m*p[j] are formed as plain 64-bit products (one IMAD.WIDE each, no carry
      // operand) and folded in with add-with-carry chains: those IADD3.X run on the ALU pipe, which is idle
      // while the FMA pipe is the bottleneck.  (With madc chains ptxas splits each of these products into an
      // IMAD.X + IMAD.HI.U32.X pair for the 8-limb field: 6 FMA-pipe cycles instead of 4 per product.)
      uint32_t pl[N], ph[N];
#pragma unroll
      for (int j = 2; j < N; j++) {
        uint64_t pr = (uint64_t)P::mod(j) * (uint64_t)m;
        pl[j] = (uint32_t)pr;
        ph[j] = (uint32_t)(pr >> 32);
      }
#pragma unroll
      for (int j = 3; j < N; j += 2) {
        O[j - 1] = addc_cc(O[j - 1], pl[j]);
        O[j] = addc_cc(O[j], ph[j]);
      }
      E[2] = add_cc(E[2], pl[2]);
      E[3] = addc_cc(E[3], ph[2]);
#pragma unroll
      for (int j = 4; j < N; j += 2) {
        E[j] = addc_cc(E[j], pl[j]);
        E[j + 1] = addc_cc(E[j + 1], ph[j]);
      }
      O[N - 1] = addc(O[N - 1], 0);
      return;  // E[0] is now (logically) zero and is never read again; E[1] is untouched
    }
    uint32_t m = mul_lo(E[0], P::INV);
    chain_mad_mod(O, 1, m);  // odd limbs of p land on columns (1,2),(3,4),..; no carry out (T bound)
    chain_mad_mod(E, 0, m);  // even limbs of p land on columns (0,1),(2,3),..
    O[N - 1] = addc(O[N - 1], 0);  // carry out of column N-1 -> column N = O[N-1]
  }
  TKM_HD friend Fp operator*(const Fp &a, const Fp &b) {
    uint32_t X[N], Y[N];
    // row 0: X holds columns of even limbs of a, Y (shifted by one column) the odd limbs.
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      X[j] = mul_lo(a.v[j], b.v[0]);
      X[j + 1] = mul_hi(a.v[j], b.v[0]);
      Y[j] = mul_lo(a.v[j + 1], b.v[0]);
      Y[j + 1] = mul_hi(a.v[j + 1], b.v[0]);
    }
    reduce_row(X, Y);
#pragma unroll
    for (int i = 1; i < N; i++) {
      // Entering: T = E + 2^32*O with E[0] == 0.  T/2^32 = O + (E >> 32): O becomes the column-0
      // accumulator (plus the stray word E[1]); E >> 64 becomes the column-1 accumulator.
      uint32_t *E = (i & 1) ? X : Y;
      uint32_t *O = (i & 1) ? Y : X;
      uint32_t bi = b.v[i];
      O[0] = add_cc(O[0], E[1]);
#pragma unroll
      for (int j = 1; j < N - 1; j += 2) {
        E[j - 1] = madc_lo_cc(a.v[j], bi, E[j + 1]);
        E[j] = madc_hi_cc(a.v[j], bi, E[j + 2]);
      }
      E[N - 2] = madc_lo_cc(a.v[N - 1], bi, 0);
      E[N - 1] = madc_hi(a.v[N - 1], bi, 0);
      chain_mad(O, a.v, bi);
      E[N - 1] = addc(E[N - 1], 0);
      reduce_row(O, E);
    }
    // N is even: after the last row E = Y (column 0, Y[0] == 0), O = X.  Result = O + (E >> 32).
    Fp r;
    r.v[0] = add_cc(X[0], Y[1]);
#pragma unroll
    for (int k = 1; k < N - 1; k++) r.v[k] = addc_cc(X[k], Y[k + 1]);
    r.v[N - 1] = addc(X[N - 1], 0);
    final_sub(r.v);
    return r;
  }
  // (lo,hi)(acc[j-1],acc[j]) += c[j]*d for the odd limbs j = 1,3,..,N-1, one carry chain (no carry out by the T bound).
  TKM_HD static void chain_mad_odd(uint32_t *acc, const uint32_t *c, uint32_t d) {
    acc[0] = mad_lo_cc(c[1], d, acc[0]);
    acc[1] = madc_hi_cc(c[1], d, acc[1]);
#pragma unroll
    for (int j = 3; j < N; j += 2) {
      acc[j - 1] = madc_lo_cc(c[j], d, acc[j - 1]);
      acc[j] = madc_hi_cc(c[j], d, acc[j]);
    }
  }
  // a*b + c*d with ONE interleaved reduction: 2 N^2 product IMADs + N^2 reduction IMADs instead of 4 N^2 for two
  // Montgomery products (Fq: 432 instead of 576).  The point formulas end in Y3 = R*(Q - X3) - Y*PPP, which is this with
  // c = -Y.  Row i adds a*b_i + c*d_i + m_i*p before the division by 2^32, so T < 3p + 3p*2^32 must stay below
  // 2^(32(N+1)): true for Fq (2^414.6 < 2^416), false for Fr -- the same spare-bit condition as the dedicated squaring.
  // The result (ab + cd + Mp)/R < p(2p/R + 1) < 2p, so one final subtraction is enough.
  TKM_HD static Fp dot2(const Fp &a, const Fp &b, const Fp &c, const Fp &d) {
    static_assert(P::DEDICATED_SQR, "dot2 needs the spare bits of the container (Fq only)");
    uint32_t X[N], Y[N];
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      X[j] = mul_lo(a.v[j], b.v[0]);
      X[j + 1] = mul_hi(a.v[j], b.v[0]);
      Y[j] = mul_lo(a.v[j + 1], b.v[0]);
      Y[j + 1] = mul_hi(a.v[j + 1], b.v[0]);
    }
    chain_mad(X, c.v, d.v[0]);          // even limbs of c: columns (0,1),(2,3),..; carry out -> column N = Y[N-1]
    Y[N - 1] = addc(Y[N - 1], 0);
    chain_mad_odd(Y, c.v, d.v[0]);      // odd limbs of c: columns (1,2),(3,4),..
    reduce_row(X, Y);
#pragma unroll
    for (int i = 1; i < N; i++) {
      uint32_t *E = (i & 1) ? X : Y;
      uint32_t *O = (i & 1) ? Y : X;
      const uint32_t bi = b.v[i], di = d.v[i];
      O[0] = add_cc(O[0], E[1]);
#pragma unroll
      for (int j = 1; j < N - 1; j += 2) {
        E[j - 1] = madc_lo_cc(a.v[j], bi, E[j + 1]);
        E[j] = madc_hi_cc(a.v[j], bi, E[j + 2]);
      }
      E[N - 2] = madc_lo_cc(a.v[N - 1], bi, 0);
      E[N - 1] = madc_hi(a.v[N - 1], bi, 0);
      chain_mad(O, a.v, bi);
      E[N - 1] = addc(E[N - 1], 0);
      chain_mad(O, c.v, di);
      E[N - 1] = addc(E[N - 1], 0);
      chain_mad_odd(E, c.v, di);
      reduce_row(O, E);
    }
    Fp r;
    r.v[0] = add_cc(X[0], Y[1]);
#pragma unroll
    for (int k = 1; k < N - 1; k++) r.v[k] = addc_cc(X[k], Y[k + 1]);
    r.v[N - 1] = addc(X[N - 1], 0);
    final_sub(r.v);
    return r;
  }
  // Dedicated Montgomery squaring: a^2 = sum_i a_i * c^(i) * 2^(32 i) with c^(i) = a_i 2^(32 i) + 2 * sum_{j>i} a_j 2^(32 j),
  // so row i needs only the N - i products with j >= i: N(N+1)/2 wide IMADs for the product instead of N^2 (Fq: 78 + 144
  // for the interleaved reduction = 222 instead of 288).  The doubled multiplicand is taken limb-wise: position i is a_i,
  // position i+1 is a_{i+1} << 1 (the bit that would shift in from a_i is not part of the sum), positions j >= i+2 are
  // d_j = (a_j << 1) | (a_{j-1} >> 31); d_N = a_{N-1} >> 31 is zero for both fields (top limbs < 2^31).  Same row structure
  // as operator*: skipped products of the shifting odd-column chain become carry-propagating moves (IADD3.X on the idle
  // ALU pipe), the even-column chain simply starts at the first column that has a product.  Bound: the partial sums are
  // L (L + 2H 2^(32(i+1))) < 2a 2^(32(i+1)), so before each division T < 3p + 3 2^32 p, which must stay below 2^(32(N+1)) as
  // operator* requires: true for Fq (p < 2^381: 2^414.6 < 2^416), false for Fr (2^288.4 > 2^288) -> P::DEDICATED_SQR.
  TKM_HD Fp sqr() const {
    if (!P::DEDICATED_SQR) return *this * *this;
    const uint32_t *a = v;
    uint32_t e[N], d[N];
    e[0] = a[0] << 1;
    d[0] = e[0];
#pragma unroll
    for (int j = 1; j < N; j++) {
      e[j] = a[j] << 1;
      d[j] = e[j] | (a[j - 1] >> 31);
    }
#define TKM_SQR_V(i, j) ((j) == (i) ? a[(j)] : ((j) == (i) + 1 ? e[(j)] : d[(j)]))
    uint32_t X[N], Y[N];
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      X[j] = mul_lo(TKM_SQR_V(0, j), a[0]);
      X[j + 1] = mul_hi(TKM_SQR_V(0, j), a[0]);
      Y[j] = mul_lo(TKM_SQR_V(0, j + 1), a[0]);
      Y[j + 1] = mul_hi(TKM_SQR_V(0, j + 1), a[0]);
    }
    reduce_row(X, Y);
#pragma unroll
    for (int i = 1; i < N; i++) {
      uint32_t *E = (i & 1) ? X : Y;
      uint32_t *O = (i & 1) ? Y : X;
      const uint32_t bi = a[i];
      O[0] = add_cc(O[0], E[1]);
#pragma unroll
      for (int j = 1; j < N - 1; j += 2) {
        if (j >= i) {
          E[j - 1] = madc_lo_cc(TKM_SQR_V(i, j), bi, E[j + 1]);
          E[j] = madc_hi_cc(TKM_SQR_V(i, j), bi, E[j + 2]);
        } else {  // no product in this column: shift down two limbs and keep the carry moving
          E[j - 1] = addc_cc(E[j + 1], 0);
          E[j] = addc_cc(E[j + 2], 0);
        }
      }
      E[N - 2] = madc_lo_cc(TKM_SQR_V(i, N - 1), bi, 0);  // N - 1 is odd and >= i for every row
      E[N - 1] = madc_hi(TKM_SQR_V(i, N - 1), bi, 0);
      const int j0 = (i + 1) & ~1;  // first even column with a product
      if (j0 < N) {
        O[j0] = mad_lo_cc(TKM_SQR_V(i, j0), bi, O[j0]);
        O[j0 + 1] = madc_hi_cc(TKM_SQR_V(i, j0), bi, O[j0 + 1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
          if (j > j0) {
            O[j] = madc_lo_cc(TKM_SQR_V(i, j), bi, O[j]);
            O[j + 1] = madc_hi_cc(TKM_SQR_V(i, j), bi, O[j + 1]);
          }
        }
        E[N - 1] = addc(E[N - 1], 0);
      }
      reduce_row(O, E);
    }
#undef TKM_SQR_V
    Fp r;
    r.v[0] = add_cc(X[0], Y[1]);
#pragma unroll
    for (int k = 1; k < N - 1; k++) r.v[k] = addc_cc(X[k], Y[k + 1]);
    r.v[N - 1] = addc(X[N - 1], 0);
    final_sub(r.v);
    return r;
  }

  TKM_HD Fp to_mont() const { return *this * r2(); }
  TKM_HD Fp from_mont() const {
    Fp o = zero();
    o.v[0] = 1;
    return *this * o;
  }
  // a^e for a small public exponent.
  TKM_HD Fp pow_u64(uint64_t e) const {
    Fp acc = one(), base = *this;
    while (e) {
      if (e & 1) acc = acc * base;
      base = base.sqr();
      e >>= 1;
    }
    return acc;
  }
  // Fermat inverse a^(p-2); inv(0) = 0 like the reference backend (bivariate_polynomial/mod.rs:2011-2013).
  TKM_HD Fp inv() const {
    Fp acc = one(), base = *this;
    uint32_t borrow = 2;  // exponent limbs of p - 2, borrow rippling up (r ends in ...00000001)
    for (int i = 0; i < N; i++) {
      uint32_t m = P::mod(i);
      uint32_t e = m - borrow;
      borrow = (m < borrow) ? 1u : 0u;
      for (int b = 0; b < 32; b++) {
        if ((e >> b) & 1) acc = acc * base;
        base = base.sqr();
      }
    }
    return acc;
  }
  // The same inverse by the binary extended Euclid (shifts and subtractions only): ~2*bits cheap steps instead of ~1.5*bits
  // dependent Montgomery products.  For the latency-bound single-chain tails (one warp converting one point to affine) it is
  // several times faster than the Fermat chain; the loop trip counts depend on the value, so it is for replicated or scalar
  // use, not for lanes that hold different values.  Invariants: a*x1 = u, a*x2 = v (mod p) with a the limbs of *this read
  // as an integer; they end at u = 1 or v = 1.  *this = A*R gives (A*R)^-1; two Montgomery products by R^2 restore A^-1*R.
  TKM_HD Fp inv_bgcd() const {
    if (is_zero()) return zero();
    uint32_t u[N], w[N];
    Fp x1 = zero(), x2 = zero();
    x1.v[0] = 1;
#pragma unroll
    for (int i = 0; i < N; i++) {
      u[i] = v[i];
      w[i] = P::mod(i);
    }
    for (;;) {
      while (!(u[0] & 1)) {
        shr1(u, 0);
        x1.halve();
      }
      if (is_one(u)) break;
      while (!(w[0] & 1)) {
        shr1(w, 0);
        x2.halve();
      }
      if (is_one(w)) break;
      uint32_t d[N];
      d[0] = sub_cc(u[0], w[0]);
#pragma unroll
      for (int i = 1; i < N; i++) d[i] = subc_cc(u[i], w[i]);
      const uint32_t borrow = subc(0, 0);  // all-ones if u < w
      if (!borrow) {
#pragma unroll
        for (int i = 0; i < N; i++) u[i] = d[i];
        x1 = x1 - x2;
      } else {
        w[0] = sub_cc(w[0], u[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) w[i] = subc_cc(w[i], u[i]);
        w[N - 1] = subc(w[N - 1], u[N - 1]);
        x2 = x2 - x1;
      }
    }
    const Fp t = is_one(u) ? x1 : x2;
    return (t * r2()) * r2();
  }
  // ---- The same inverse by the binary GCD on approximations (Pornin, "Optimized Binary GCD for Modular Inversion",
  // ePrint 2020/972, algorithm 2 with k = 32): every round takes 31 steps of the binary GCD on 64-bit stand-ins for (a, b)
  // -- their low 31 bits, which decide the parities exactly, under their top 33 bits, which decide the comparisons almost
  // always -- and records them as a 2x2 matrix of factors |f|,|g| <= 2^31; the matrix is then applied once to the full-width
  // (a, b) (exact division by 2^31; a wrong comparison only makes a value negative, which is undone by negating its row) and,
  // with a Montgomery step by 2^31, to the cofactors (u, v) mod p.  ceil((2*32N - 1)/31) rounds reach (a, b) = (0, 1) for
  // every input, after which v = (limbs of *this)^-1.  About a third of the instructions of inv_bgcd (one 12-limb update
  // per 31 steps instead of per step): the one-thread inversions on the MSM's critical path -- one per pair-tree level, one in
  // the Horner tail -- take 25-30 us instead of 85.  inv_fast() checks the product and falls back to inv_bgcd, so a wrong
  // value can not leave this function.
  TKM_HD static void lin_comb(const uint32_t *X, const uint32_t *Y, int64_t f, int64_t g, uint32_t *S) {  // S: N + 2 limbs, two's complement
    const uint32_t fs = f < 0, gs = g < 0;
    const uint32_t fa = (uint32_t)(fs ? -f : f), ga = (uint32_t)(gs ? -g : g);  // <= 2^31
    const uint32_t mf = 0u - fs, mg = 0u - gs;
    uint64_t cp = 0, cq = 0, c = (uint64_t)fs + gs;
#pragma unroll
    for (int i = 0; i < N; i++) {
      cp += (uint64_t)X[i] * fa;
      cq += (uint64_t)Y[i] * ga;
      c += (uint64_t)((uint32_t)cp ^ mf) + ((uint32_t)cq ^ mg);
      S[i] = (uint32_t)c;
      c >>= 32;
      cp >>= 32;
      cq >>= 32;
    }
    c += (uint64_t)((uint32_t)cp ^ mf) + ((uint32_t)cq ^ mg);
    S[N] = (uint32_t)c;
    c >>= 32;
    S[N + 1] = (uint32_t)(c + mf + mg);
  }
  // the 64-bit stand-in of x: low 31 bits under the 33 bits below position 32*(hi + 1) - lz (hi >= 1: the top non-zero limb of a | b)
  TKM_HD static uint64_t bingcd_approx(const uint32_t *x, int hi, uint32_t lz) {
    uint32_t x2 = x[1], x1 = x[0], x0 = 0;
#pragma unroll
    for (int i = 2; i < N; i++)
      if (i == hi) {
        x2 = x[i];
        x1 = x[i - 1];
        x0 = x[i - 2];
      }
    const uint32_t sh = 63u - lz;  // 31..63
    const uint64_t low = ((uint64_t)x1 << 32) | x0;
    uint64_t top = (low >> sh) | ((uint64_t)x2 << (64u - sh));
    top &= ((uint64_t)1 << 33) - 1;
    return (uint64_t)(x[0] & 0x7fffffffu) | (top << 31);
  }
  TKM_HD Fp inv_pornin(bool *ok) const {
    uint32_t a[N], b[N], u[N], w[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
      a[i] = v[i];
      b[i] = P::mod(i);
      u[i] = 0;
      w[i] = 0;
    }
    u[0] = 1;
    constexpr int ROUNDS = (2 * 32 * N - 1 + 30) / 31;
    for (int round = 0; round < ROUNDS; round++) {
      int hi = 1;
#pragma unroll
      for (int i = 2; i < N; i++)
        if ((a[i] | b[i]) != 0) hi = i;
      uint32_t chi = a[1] | b[1];
#pragma unroll
      for (int i = 2; i < N; i++)
        if (i == hi) chi = a[i] | b[i];
      uint32_t lz = 32;
      if (chi) {
        lz = 0;
        while (!((chi << lz) & 0x80000000u)) lz++;
      }
      uint64_t aa = bingcd_approx(a, hi, lz), bb = bingcd_approx(b, hi, lz);
      int64_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
      for (int i = 0; i < 31; i++) {
        if (aa & 1) {
          if (aa < bb) {
            const uint64_t t = aa; aa = bb; bb = t;
            int64_t s = f0; f0 = f1; f1 = s;
            s = g0; g0 = g1; g1 = s;
          }
          aa = (aa - bb) >> 1;
          f0 -= f1;
          g0 -= g1;
        } else {
          aa >>= 1;
        }
        f1 <<= 1;
        g1 <<= 1;
      }
      uint32_t Sa[N + 2], Sb[N + 2];
      lin_comb(a, b, f0, g0, Sa);
      lin_comb(a, b, f1, g1, Sb);
      // (a, b) <- |S >> 31|; a negated row negates its factors
      const uint32_t na = Sa[N + 1] >> 31, nb = Sb[N + 1] >> 31;
      {
        const uint32_t ma = 0u - na, mb = 0u - nb;
        uint64_t ca = na, cb = nb;
#pragma unroll
        for (int i = 0; i < N; i++) {
          const uint32_t xa = (Sa[i] >> 31) | (Sa[i + 1] << 1), xb = (Sb[i] >> 31) | (Sb[i + 1] << 1);
          ca += (uint32_t)(xa ^ ma);
          cb += (uint32_t)(xb ^ mb);
          a[i] = (uint32_t)ca;
          b[i] = (uint32_t)cb;
          ca >>= 32;
          cb >>= 32;
        }
      }
      if (na) { f0 = -f0; g0 = -g0; }
      if (nb) { f1 = -f1; g1 = -g1; }
      // (u, w) <- ((u f + w g) + t p) / 2^31 mod p, t = -(u f + w g) / p mod 2^31
      uint32_t Su[N + 2], Sw[N + 2];
      lin_comb(u, w, f0, g0, Su);
      lin_comb(u, w, f1, g1, Sw);
#pragma unroll
      for (int which = 0; which < 2; which++) {
        uint32_t *S = which ? Sw : Su;
        uint32_t *dst = which ? w : u;
        const uint32_t t = (S[0] * P::INV) & 0x7fffffffu;
        uint64_t c = 0, cm = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
          cm += (uint64_t)P::mod(i) * t;
          c += (uint64_t)S[i] + (uint32_t)cm;
          S[i] = (uint32_t)c;
          c >>= 32;
          cm >>= 32;
        }
        c += (uint64_t)S[N] + (uint32_t)cm;
        S[N] = (uint32_t)c;
        c >>= 32;
        S[N + 1] = (uint32_t)(S[N + 1] + c);
        // r = S >> 31 in [-p, 2p): N limbs + a sign / overflow limb
        uint32_t r[N + 1];
#pragma unroll
        for (int i = 0; i <= N; i++) r[i] = (S[i] >> 31) | (S[i + 1] << 1);
        if (r[N] >> 31) {  // negative: add p
          uint64_t cc = 0;
#pragma unroll
          for (int i = 0; i < N; i++) {
            cc += (uint64_t)r[i] + P::mod(i);
            r[i] = (uint32_t)cc;
            cc >>= 32;
          }
        } else {  // subtract p when r >= p (r < 2p)
          uint32_t d[N];
          uint64_t bw = 0;
#pragma unroll
          for (int i = 0; i < N; i++) {
            const uint64_t df = (uint64_t)r[i] - P::mod(i) - bw;
            d[i] = (uint32_t)df;
            bw = (df >> 63) & 1;
          }
          const bool ge = r[N] != 0 || bw == 0;
          if (ge) {
#pragma unroll
            for (int i = 0; i < N; i++) r[i] = d[i];
          }
        }
#pragma unroll
        for (int i = 0; i < N; i++) dst[i] = r[i];
      }
    }
    // (a, b) must be (0, 1)
    uint32_t chk = b[0] ^ 1u;
#pragma unroll
    for (int i = 0; i < N; i++) chk |= a[i] | (i ? b[i] : 0u);
    *ok = chk == 0;
    Fp t;
#pragma unroll
    for (int i = 0; i < N; i++) t.v[i] = w[i];
    return (t * r2()) * r2();
  }
  // Inverse for single chains: the fast form, checked; inv(0) = 0.
  TKM_HD Fp inv_fast() const {
    if (is_zero()) return zero();
    bool ok;
    const Fp r = inv_pornin(&ok);
    if (ok && (r * *this) == one()) return r;
    return inv_bgcd();
  }
  // helpers of inv_bgcd
  TKM_HD static bool is_one(const uint32_t *a) {
    uint32_t acc = a[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < N; i++) acc |= a[i];
    return acc == 0;
  }
  TKM_HD static void shr1(uint32_t *a, uint32_t top_bit) {
#pragma unroll
    for (int i = 0; i < N - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
    a[N - 1] = (a[N - 1] >> 1) | (top_bit << 31);
  }
  // x/2 mod p for x in [0, p): (x odd ? x + p : x) >> 1.  x + p < 2^(32N) for both fields (no carry out of the container).
  TKM_HD void halve() {
    const uint32_t mask = 0u - (v[0] & 1u);
    v[0] = add_cc(v[0], P::mod(0) & mask);
#pragma unroll
    for (int i = 1; i < N - 1; i++) v[i] = addc_cc(v[i], P::mod(i) & mask);
    v[N - 1] = addc(v[N - 1], P::mod(N - 1) & mask);
    shr1(v, 0);
  }
};

using Fr = Fp<FrParams>;
using Fq = Fp<FqParams>;

}  // namespace tkm
