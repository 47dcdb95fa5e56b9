#include "fq_karatsuba.cuh"
using namespace tkm;
__global__ void k_kara(const Fq *a, const Fq *b, Fq *o) { Fq x = a[threadIdx.x], y = b[threadIdx.x]; o[threadIdx.x] = mul_karatsuba(x, y); }
__global__ void k_mul(const Fq *a, const Fq *b, Fq *o) { Fq x = a[threadIdx.x], y = b[threadIdx.x]; o[threadIdx.x] = x * y; }
