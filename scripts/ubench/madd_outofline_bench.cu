// Probe: XYZZ mixed-addition stream with the Fq products inlined (the library) against out-of-line products (smaller hot
// loop, call ABI moving the operands).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr
//   -o scripts/ubench/madd_outofline_bench scripts/ubench/madd_outofline_bench.cu
#include <cstdio>
#include "../../tokamak-zk-evm_b200/csrc/g1.cuh"
using namespace tkm;
__device__ __noinline__ Fq mul_ni(Fq a, Fq b) { return a * b; }
__device__ __noinline__ Fq sqr_ni(Fq a) { return a.sqr(); }
__device__ __noinline__ Fq dot2_ni(Fq a, Fq b, Fq c, Fq d) { return Fq::dot2(a, b, c, d); }
__device__ __forceinline__ void madd_ni(G1Xyzz &acc, const G1Affine &p) {
  if (p.is_identity()) return;
  if (acc.is_identity()) { acc = G1Xyzz{p.x, p.y, Fq::one(), Fq::one()}; return; }
  Fq U2 = mul_ni(p.x, acc.ZZ);
  Fq S2 = mul_ni(p.y, acc.ZZZ);
  Fq Pd = U2 - acc.X;
  Fq Rd = S2 - acc.Y;
  if (Pd.is_zero()) { acc = Rd.is_zero() ? g1_mdbl(p) : G1Xyzz::identity(); return; }
  Fq PP = sqr_ni(Pd);
  Fq PPP = mul_ni(Pd, PP);
  Fq Q = mul_ni(acc.X, PP);
  Fq X3 = sqr_ni(Rd) - PPP - Q.dbl();
  acc.ZZ = mul_ni(acc.ZZ, PP);
  acc.ZZZ = mul_ni(acc.ZZZ, PPP);
  acc.Y = dot2_ni(Rd, Q - X3, acc.Y.neg(), PPP);
  acc.X = X3;
}
template <int KIND>
__global__ void __launch_bounds__(128, 3) k_stream(const G1Affine *pts, G1Xyzz *out, int iters) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  G1Xyzz acc = G1Xyzz::identity();
  G1Affine p = pts[t];
  for (int i = 0; i < iters; i++) {
    if (KIND == 0) g1_madd(acc, p); else madd_ni(acc, p);
    p.x = p.x + acc.X;  // a different (not on-curve) operand every time: the formulas are exercised as arithmetic
  }
  out[t] = acc;
}
int main() {
  const int blocks = 148 * 3 * 2, threads = 128, n = blocks * threads, iters = 400;
  G1Affine *h = new G1Affine[n];
  for (int i = 0; i < n; i++)
    for (int k = 0; k < 12; k++) { h[i].x.v[k] = (k == 11) ? (i * 2654435761u) & 0x0fffffffu : i * 2246822519u + k * 3266489917u; h[i].y.v[k] = (k == 11) ? (i * 40503u) & 0x0fffffffu : i * 668265263u + k * 374761393u; }
  G1Affine *d; G1Xyzz *o0, *o1;
  cudaMalloc(&d, n * sizeof(G1Affine)); cudaMalloc(&o0, n * sizeof(G1Xyzz)); cudaMalloc(&o1, n * sizeof(G1Xyzz));
  cudaMemcpy(d, h, n * sizeof(G1Affine), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int kind = 0; kind < 2; kind++)
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      if (kind == 0) k_stream<0><<<blocks, threads>>>(d, o0, iters); else k_stream<1><<<blocks, threads>>>(d, o1, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("%s: %.3f ms, %.3f G madd/s\n", kind ? "out-of-line products" : "inlined (library)", ms, 1.0 * n * iters / ms / 1e6);
    }
  G1Xyzz *a = new G1Xyzz[n], *b = new G1Xyzz[n];
  cudaMemcpy(a, o0, n * sizeof(G1Xyzz), cudaMemcpyDeviceToHost); cudaMemcpy(b, o1, n * sizeof(G1Xyzz), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < n; i++) bad += !(a[i].X == b[i].X && a[i].Y == b[i].Y && a[i].ZZ == b[i].ZZ && a[i].ZZZ == b[i].ZZZ);
  printf("mismatches: %d, cuda: %s\n", bad, cudaGetErrorString(cudaGetLastError()));
  return bad != 0;
}
