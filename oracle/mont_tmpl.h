/* ORACLE (test infrastructure only) -- Montgomery field template, 64-bit limbs.
 *
 * Included twice by oracle.c with
 *   FNAME(x)  name-mangling macro, NL number of 64-bit limbs,
 *   MODULUS[] little-endian limbs, INV64 = -p^-1 mod 2^64, R2[] = R^2 mod p.
 * Plain CIOS Montgomery multiplication with unsigned __int128; no assembly.
 * Restates the field arithmetic the reference gets from the un-vendored
 * icicle-bls12-381 v3.8.0 crate (packages/backend/Cargo.toml:20-23):
 * ScalarField / BaseField with canonical little-endian host representation
 * (SURVEY.md Appendix C).
 */
#ifndef FNAME
#error "define FNAME, NL, MODULUS, INV64, R2 before including"
#endif

typedef struct { uint64_t l[NL]; } FNAME(t);

static inline int FNAME(is_zero)(const FNAME(t) *a) {
  uint64_t acc = 0;
  for (int i = 0; i < NL; i++) acc |= a->l[i];
  return acc == 0;
}
static inline int FNAME(eq)(const FNAME(t) *a, const FNAME(t) *b) {
  uint64_t acc = 0;
  for (int i = 0; i < NL; i++) acc |= a->l[i] ^ b->l[i];
  return acc == 0;
}
static inline int FNAME(geq_mod)(const uint64_t *a) {
  for (int i = NL - 1; i >= 0; i--) {
    if (a[i] > MODULUS[i]) return 1;
    if (a[i] < MODULUS[i]) return 0;
  }
  return 1;
}
static inline void FNAME(sub_mod_raw)(uint64_t *a) {
  unsigned __int128 br = 0;
  for (int i = 0; i < NL; i++) {
    unsigned __int128 d = (unsigned __int128)a[i] - MODULUS[i] - (uint64_t)br;
    a[i] = (uint64_t)d;
    br = (d >> 64) & 1;
  }
}
static inline void FNAME(add)(FNAME(t) *r, const FNAME(t) *a, const FNAME(t) *b) {
  unsigned __int128 c = 0;
  uint64_t t[NL];
  for (int i = 0; i < NL; i++) {
    c += (unsigned __int128)a->l[i] + b->l[i];
    t[i] = (uint64_t)c;
    c >>= 64;
  }
  if (c || FNAME(geq_mod)(t)) FNAME(sub_mod_raw)(t);
  for (int i = 0; i < NL; i++) r->l[i] = t[i];
}
static inline void FNAME(sub)(FNAME(t) *r, const FNAME(t) *a, const FNAME(t) *b) {
  unsigned __int128 br = 0;
  uint64_t t[NL];
  for (int i = 0; i < NL; i++) {
    unsigned __int128 d = (unsigned __int128)a->l[i] - b->l[i] - (uint64_t)br;
    t[i] = (uint64_t)d;
    br = (d >> 64) & 1;
  }
  if (br) {
    unsigned __int128 c = 0;
    for (int i = 0; i < NL; i++) {
      c += (unsigned __int128)t[i] + MODULUS[i];
      t[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  for (int i = 0; i < NL; i++) r->l[i] = t[i];
}
static inline void FNAME(neg)(FNAME(t) *r, const FNAME(t) *a) {
  FNAME(t) z;
  memset(&z, 0, sizeof z);
  FNAME(sub)(r, &z, a);
}
/* r = a*b/R mod p (CIOS). */
static inline void FNAME(mul)(FNAME(t) *r, const FNAME(t) *a, const FNAME(t) *b) {
  uint64_t t[NL + 2];
  memset(t, 0, sizeof t);
  for (int i = 0; i < NL; i++) {
    unsigned __int128 c = 0;
    for (int j = 0; j < NL; j++) {
      c += (unsigned __int128)a->l[j] * b->l[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[NL];
    t[NL] = (uint64_t)c;
    t[NL + 1] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * INV64;
    c = (unsigned __int128)m * MODULUS[0] + t[0];
    c >>= 64;
    for (int j = 1; j < NL; j++) {
      c += (unsigned __int128)m * MODULUS[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[NL];
    t[NL - 1] = (uint64_t)c;
    t[NL] = t[NL + 1] + (uint64_t)(c >> 64);
  }
  if (t[NL] || FNAME(geq_mod)(t)) FNAME(sub_mod_raw)(t);
  for (int i = 0; i < NL; i++) r->l[i] = t[i];
}
static inline void FNAME(sqr)(FNAME(t) *r, const FNAME(t) *a) { FNAME(mul)(r, a, a); }
static inline void FNAME(to_mont)(FNAME(t) *r, const FNAME(t) *a) {
  FNAME(t) r2;
  for (int i = 0; i < NL; i++) r2.l[i] = R2[i];
  FNAME(mul)(r, a, &r2);
}
static inline void FNAME(from_mont)(FNAME(t) *r, const FNAME(t) *a) {
  FNAME(t) one;
  memset(&one, 0, sizeof one);
  one.l[0] = 1;
  FNAME(mul)(r, a, &one);
}
static inline void FNAME(one)(FNAME(t) *r) {
  FNAME(t) one;
  memset(&one, 0, sizeof one);
  one.l[0] = 1;
  FNAME(to_mont)(r, &one);
}
/* r = a^e, e given as ne little-endian limbs (plain integer). */
static inline void FNAME(pow)(FNAME(t) *r, const FNAME(t) *a, const uint64_t *e, int ne) {
  FNAME(t) acc, base = *a;
  FNAME(one)(&acc);
  for (int i = 0; i < ne; i++)
    for (int b = 0; b < 64; b++) {
      if ((e[i] >> b) & 1) FNAME(mul)(&acc, &acc, &base);
      FNAME(sqr)(&base, &base);
    }
  *r = acc;
}
/* Fermat inverse; inv(0) = 0 (ICICLE convention, bivariate_polynomial/mod.rs:2011-2013). */
static inline void FNAME(inv)(FNAME(t) *r, const FNAME(t) *a) {
  uint64_t e[NL];
  for (int i = 0; i < NL; i++) e[i] = MODULUS[i];
  e[0] -= 2; /* both moduli end in ...01 / ...ab: no borrow */
  FNAME(pow)(r, a, e, NL);
}

#undef FNAME
#undef NL
#undef MODULUS
#undef INV64
#undef R2
