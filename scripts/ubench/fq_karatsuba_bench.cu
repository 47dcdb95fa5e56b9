// Throughput of Fp::operator* against the Karatsuba variant (fq_karatsuba.cuh), both as a stream of dependent products
// with two independent chains per thread, 12 warps per SM like k_accumulate.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o /tmp/kb scripts/ubench/fq_karatsuba_bench.cu && /tmp/kb
#include <cstdio>
#include "fq_karatsuba.cuh"
using namespace tkm;
template <int KIND>
__global__ void __launch_bounds__(128, 3) k_stream(Fq *io, int iters) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  Fq x = io[2 * t], y = io[2 * t + 1];
  for (int i = 0; i < iters; i++) {
    if (KIND == 0) { x = x * y; y = y * x; } else { x = mul_karatsuba(x, y); y = mul_karatsuba(y, x); }
  }
  io[2 * t] = x;
  io[2 * t + 1] = y;
}
int main() {
  const int blocks = 148 * 3 * 4, threads = 128, n = blocks * threads, iters = 2000;
  Fq *h = new Fq[2 * n];
  for (int i = 0; i < 2 * n; i++)
    for (int k = 0; k < 12; k++) h[i].v[k] = (k == 11) ? (uint32_t)(i * 2654435761u) & 0x0fffffffu : (uint32_t)(i * 2246822519u + k * 3266489917u);
  Fq *d0, *d1;
  cudaMalloc(&d0, 2 * n * sizeof(Fq));
  cudaMalloc(&d1, 2 * n * sizeof(Fq));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms[2];
  Fq *out[2] = {new Fq[2 * n], new Fq[2 * n]};
  for (int kind = 0; kind < 2; kind++) {
    Fq *d = kind ? d1 : d0;
    for (int rep = 0; rep < 2; rep++) {  // first repetition warms up
      cudaMemcpy(d, h, 2 * n * sizeof(Fq), cudaMemcpyHostToDevice);
      cudaEventRecord(e0);
      if (kind == 0) k_stream<0><<<blocks, threads>>>(d, iters); else k_stream<1><<<blocks, threads>>>(d, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms[kind], e0, e1);
    }
    cudaMemcpy(out[kind], d, 2 * n * sizeof(Fq), cudaMemcpyDeviceToHost);
    printf("%s: %.3f ms, %.2f G mul/s\n", kind ? "karatsuba" : "operator*", ms[kind], 2.0 * n * iters / ms[kind] / 1e6);
  }
  int bad = 0;
  for (int i = 0; i < 2 * n; i++) bad += !(out[0][i] == out[1][i]);
  printf("mismatches: %d, cuda: %s\n", bad, cudaGetErrorString(cudaGetLastError()));
  return bad != 0;
}
