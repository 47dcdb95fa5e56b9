// Probe: XYZZ mixed-addition stream (library g1_madd) at 3 CTAs/SM (168 registers, what k_accumulate runs at) against
// 4 CTAs/SM (<= 128 registers, spills) and 2 CTAs/SM (<= 255 registers).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o scripts/ubench/madd_occupancy_bench scripts/ubench/madd_occupancy_bench.cu
#include <cstdio>
#include "../../tokamak-zk-evm_b200/csrc/g1.cuh"
using namespace tkm;
__device__ __forceinline__ void body(const G1Affine *pts, G1Xyzz *out, int iters) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  G1Xyzz acc = G1Xyzz::identity();
  G1Affine p = pts[t];
  for (int i = 0; i < iters; i++) {
    g1_madd(acc, p);
    p.x = p.x + acc.X;
  }
  out[t] = acc;
}
__global__ void __launch_bounds__(128, 2) k2(const G1Affine *pts, G1Xyzz *out, int iters) { body(pts, out, iters); }
__global__ void __launch_bounds__(128, 3) k3(const G1Affine *pts, G1Xyzz *out, int iters) { body(pts, out, iters); }
__global__ void __launch_bounds__(128, 4) k4(const G1Affine *pts, G1Xyzz *out, int iters) { body(pts, out, iters); }
int main() {
  const int blocks = 148 * 12, threads = 128, n = blocks * threads, iters = 400;
  G1Affine *h = new G1Affine[n];
  for (int i = 0; i < n; i++)
    for (int k = 0; k < 12; k++) { h[i].x.v[k] = (k == 11) ? (i * 2654435761u) & 0x0fffffffu : i * 2246822519u + k * 3266489917u; h[i].y.v[k] = (k == 11) ? (i * 40503u) & 0x0fffffffu : i * 668265263u + k * 374761393u; }
  G1Affine *d; G1Xyzz *o;
  cudaMalloc(&d, n * sizeof(G1Affine)); cudaMalloc(&o, n * sizeof(G1Xyzz));
  cudaMemcpy(d, h, n * sizeof(G1Affine), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int kind = 2; kind <= 4; kind++)
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      if (kind == 2) k2<<<blocks, threads>>>(d, o, iters); else if (kind == 3) k3<<<blocks, threads>>>(d, o, iters); else k4<<<blocks, threads>>>(d, o, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("%d CTAs/SM: %.3f ms, %.3f G madd/s\n", kind, ms, 1.0 * n * iters / ms / 1e6);
    }
  printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
