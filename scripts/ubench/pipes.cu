// Instruction-throughput probes for the integer pipes of sm_100a (developer tool).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../tokamak-zk-evm_b200/csrc/g1.cuh"
using namespace tkm;

#define ITERS 2048
template <int KIND>
__global__ void __launch_bounds__(256) probe(uint32_t *sink, uint32_t seed) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x + seed;
  uint32_t r = 0;
  if (KIND == 0) {  // 16 independent IMAD.WIDE chains reg*reg + pair
    uint64_t a[16];
    uint32_t m = t | 1u;
    for (int i = 0; i < 16; i++) a[i] = t * (i + 3);
    for (int it = 0; it < ITERS; it++)
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = (uint64_t)(uint32_t)(a[i] >> 7) * m + a[i];
    for (int i = 0; i < 16; i++) r ^= (uint32_t)a[i] ^ (uint32_t)(a[i] >> 32);
  } else if (KIND == 1) {  // carry-chained wide mads: 8 pairs per chain, 2 chains
    uint32_t E[16], O[16], a[8], b = t | 1u;
    for (int i = 0; i < 16; i++) { E[i] = t + i; O[i] = t ^ i; }
    for (int i = 0; i < 8; i++) a[i] = t * (2 * i + 1);
    for (int it = 0; it < ITERS; it++) {
      E[0] = mad_lo_cc(a[0], b, E[0]); E[1] = madc_hi_cc(a[0], b, E[1]);
#pragma unroll
      for (int j = 1; j < 8; j++) { E[2 * j] = madc_lo_cc(a[j], b, E[2 * j]); E[2 * j + 1] = madc_hi_cc(a[j], b, E[2 * j + 1]); }
      O[0] = mad_lo_cc(a[0], b, O[0]); O[1] = madc_hi_cc(a[0], b, O[1]);
#pragma unroll
      for (int j = 1; j < 8; j++) { O[2 * j] = madc_lo_cc(a[j], b, O[2 * j]); O[2 * j + 1] = madc_hi_cc(a[j], b, O[2 * j + 1]); }
      b += E[3];
    }
    for (int i = 0; i < 16; i++) r ^= E[i] ^ O[i];
  } else if (KIND == 2) {  // IADD3.X carry chains only: 2 chains of 16
    uint32_t E[16], O[16], a[16];
    for (int i = 0; i < 16; i++) { E[i] = t + i; O[i] = t ^ i; a[i] = t * (i + 5); }
    for (int it = 0; it < ITERS; it++) {
      E[0] = add_cc(E[0], a[0]);
#pragma unroll
      for (int j = 1; j < 16; j++) E[j] = addc_cc(E[j], a[j]);
      O[0] = add_cc(O[0], a[1]);
#pragma unroll
      for (int j = 1; j < 16; j++) O[j] = addc_cc(O[j], a[(j + 1) & 15]);
      a[0] ^= E[15];
    }
    for (int i = 0; i < 16; i++) r ^= E[i] ^ O[i];
  } else if (KIND == 3) {  // IMAD.HI chains (independent)
    uint32_t a[16], m = t | 1u;
    for (int i = 0; i < 16; i++) a[i] = t * (i + 3);
    for (int it = 0; it < ITERS; it++)
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = __umulhi(a[i], m) + a[(i + 1) & 15];
    for (int i = 0; i < 16; i++) r ^= a[i];
  } else if (KIND == 4) {  // independent plain wide multiplies (no addend) + IADD3.X accumulate
    uint32_t E[16], a[8], b = t | 1u;
    for (int i = 0; i < 16; i++) E[i] = t + i;
    for (int i = 0; i < 8; i++) a[i] = t * (2 * i + 1);
    for (int it = 0; it < ITERS; it++) {
      uint32_t pl[8], ph[8];
#pragma unroll
      for (int j = 0; j < 8; j++) { uint64_t pr = (uint64_t)a[j] * b; pl[j] = (uint32_t)pr; ph[j] = (uint32_t)(pr >> 32); }
      E[0] = add_cc(E[0], pl[0]); E[1] = addc_cc(E[1], ph[0]);
#pragma unroll
      for (int j = 1; j < 8; j++) { E[2 * j] = addc_cc(E[2 * j], pl[j]); E[2 * j + 1] = addc_cc(E[2 * j + 1], ph[j]); }
      b += E[3];
    }
    for (int i = 0; i < 16; i++) r ^= E[i];
  } else if (KIND == 5) {  // Fr mul
    Fr x = Fr::one(), y = Fr::r2(); x.v[0] ^= t;
    for (int it = 0; it < ITERS / 8; it++) { x = x * y; y = y * x; }
    r = x.v[0] ^ y.v[1];
  } else if (KIND == 6) {  // Fr add/sub
    Fr x = Fr::one(), y = Fr::r2(); x.v[0] ^= t;
    for (int it = 0; it < ITERS; it++) { x = x + y; y = y - x; }
    r = x.v[0] ^ y.v[1];
  } else if (KIND == 7) {  // Fq mul
    Fq x = Fq::one(), y = Fq::r2(); x.v[0] ^= t;
    for (int it = 0; it < ITERS / 16; it++) { x = x * y; y = y * x; }
    r = x.v[0] ^ y.v[1];
  } else if (KIND == 8) {  // NTT-like butterfly: 1 mul + add + sub
    Fr a = Fr::one(), b = Fr::r2(), w = Fr::r2(); a.v[0] ^= t; w.v[1] ^= t;
    for (int it = 0; it < ITERS / 8; it++) { Fr s = a + b; Fr d = (a - b) * w; a = s; b = d; }
    r = a.v[0] ^ b.v[1];
  }
  if (r == 0x12345678u) sink[0] = r;
}

template <int KIND>
double run(const char *name, double ops_per_thread, uint32_t *sink) {
  int blocks = 148 * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<KIND><<<blocks, threads>>>(sink, 1);
  cudaEventRecord(e0);
  probe<KIND><<<blocks, threads>>>(sink, 2);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double rate = (double)blocks * threads * ops_per_thread / (ms * 1e-3);
  // cycles per warp-instruction per SMSP at 1.965 GHz over 592 SMSPs
  double cyc = 592.0 * 1.965e9 / (rate / 32.0);
  printf("%-44s %10.2f Gops/s   %6.2f SMSP-cycles per warp-op\n", name, rate / 1e9, cyc);
  return rate;
}

int main() {
  uint32_t *sink; cudaMalloc(&sink, 64);
  run<0>("IMAD.WIDE reg*reg+pair, 16 indep chains", 16.0 * ITERS, sink);
  run<1>("IMAD.WIDE.X carry chains (2 x 8)", 16.0 * ITERS, sink);
  run<2>("IADD3.X carry chains (2 x 16)", 32.0 * ITERS, sink);
  run<3>("IMAD.HI + IADD", 16.0 * ITERS, sink);
  run<4>("IMAD.WIDE (no addend) + IADD3.X chain (8+16)", 8.0 * ITERS, sink);
  run<5>("Fr mul", 2.0 * (ITERS / 8), sink);
  run<6>("Fr add+sub pair", 1.0 * ITERS, sink);
  run<7>("Fq mul", 2.0 * (ITERS / 16), sink);
  run<8>("Fr butterfly (mul+add+sub)", 1.0 * (ITERS / 8), sink);
  cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
