#include "fq_karatsuba.cuh"
#include <cstdio>
#include <random>
using namespace tkm;
int main() {
  std::mt19937_64 rng(5);
  int bad = 0;
  for (int it = 0; it < 200000; it++) {
    Fq a, b;
    for (int i = 0; i < 12; i++) { a.v[i] = (uint32_t)rng(); b.v[i] = (uint32_t)rng(); }
    if (it % 7 == 0) for (int i = 0; i < 12; i++) a.v[i] = 0xffffffffu;
    if (it % 11 == 0) for (int i = 0; i < 12; i++) b.v[i] = 0xffffffffu;
    if (it % 13 == 0) for (int i = 6; i < 12; i++) a.v[i] = 0;
    a.v[11] &= 0x1fffffffu; b.v[11] &= 0x1fffffffu;  // keep a, b < 2^381 (container values of the formulas)
    if (it % 5 == 0) { // reduce below p by a multiplication
      a = a * Fq::one(); b = b * Fq::one();
    }
    Fq r1 = a * b, r2 = mul_karatsuba(a, b);
    if (!(r1 == r2)) { if (bad < 5) printf("mismatch at %d\n", it); bad++; }
  }
  printf("bad=%d\n", bad);
  return bad != 0;
}
