// C-ABI entry points of libtokamak_b200 (see include/tokamak_b200.h for the reference interface each
// one replaces).  Validation + plumbing only; the kernels live in vec_ops.cu, ntt.cu, msm.cu, poly.cu.
#include <cstdlib>

#include "common.cuh"

namespace tkm {
Fr root_of_unity_host(uint32_t log_n);
int32_t domain_init(tkm_ctx *ctx, uint32_t log2_size);
int32_t domain_release(tkm_ctx *ctx);
int32_t fill_powers_public(tkm_ctx *ctx, Fr *out, const Fr &base, const Fr &scale, size_t count);

// ---- single-point helpers (G1serde ops, group_structures/mod.rs:895-947) and fixed-base batch mul
__global__ void k_g1_add_single(const G1Affine *a, const G1Affine *b, uint32_t *out_canonical) {
  if (threadIdx.x || blockIdx.x) return;
  G1Affine pa = *a, pb = *b;
  pa.x = pa.x.to_mont(); pa.y = pa.y.to_mont();
  pb.x = pb.x.to_mont(); pb.y = pb.y.to_mont();
  G1Xyzz acc = G1Xyzz::identity();
  g1_madd(acc, pa);
  g1_madd(acc, pb);
  G1Affine r = g1_to_affine_single(acc);
  Fq x = r.x.from_mont(), y = r.y.from_mont();
  for (int i = 0; i < 12; i++) { out_canonical[i] = x.v[i]; out_canonical[12 + i] = y.v[i]; }
}

// Sum of n canonical affine points (partial sums of the point-range shards): one warp, every lane
// carries a replica (g1_to_affine_coop).
__global__ void __launch_bounds__(32) k_g1_sum_canonical(const G1Affine *pts, size_t n, uint32_t *out_canonical) {
  G1Xyzz acc = G1Xyzz::identity();
  for (size_t i = 0; i < n; i++) {
    G1Affine p = pts[i];
    p.x = p.x.to_mont();
    p.y = p.y.to_mont();
    g1_madd(acc, p);
  }
  G1Affine r = g1_to_affine_coop(acc);
  if (threadIdx.x == 0) {
    Fq x = r.x.from_mont(), y = r.y.from_mont();
    for (int i = 0; i < 12; i++) { out_canonical[i] = x.v[i]; out_canonical[12 + i] = y.v[i]; }
  }
}

// out[i] = k_i * base for n scalars: 4-bit fixed windows over a 64 x 15 table of affine multiples held in
// global memory (built by one small kernel), one thread per scalar, then a batched conversion to affine.
constexpr int FB_WINDOWS = 64;
__global__ void k_fixed_base_table(G1Affine base_canonical, G1Affine *table /* [64][16] */) {
  // one warp; lane w handles windows w and w+32: entry [w][d] = d * 16^w * base
  const int lane = threadIdx.x;
  G1Affine b = base_canonical;
  b.x = b.x.to_mont(); b.y = b.y.to_mont();
  for (int w = lane; w < FB_WINDOWS; w += 32) {
    G1Xyzz cur = G1Xyzz::from_affine(b);
    for (int k = 0; k < 4 * w; k++) cur = g1_dbl(cur);
    G1Affine step = g1_to_affine(cur);
    G1Xyzz acc = G1Xyzz::identity();
    table[w * 16] = G1Affine::identity();
    for (int d = 1; d < 16; d++) {
      g1_madd(acc, step);
      table[w * 16 + d] = g1_to_affine(acc);
    }
  }
}
__global__ void __launch_bounds__(128) k_fixed_base_mul(const G1Affine *__restrict__ table, const Fr *__restrict__ scalars, int scalars_mont,
                                                       size_t n, G1Affine *__restrict__ out_canonical) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fr s = scalars[i];
    if (scalars_mont) s = s.from_mont();
    G1Xyzz acc = G1Xyzz::identity();
    for (int w = 0; w < FB_WINDOWS; w++) {
      uint32_t d = (s.v[w >> 3] >> ((w & 7) * 4)) & 15u;
      if (d) {
        G1Affine t = table[w * 16 + d];
        g1_madd(acc, t);
      }
    }
    G1Affine r = g1_to_affine(acc);
    r.x = r.x.from_mont();
    r.y = r.y.from_mont();
    out_canonical[i] = r;
  }
}

__global__ void __launch_bounds__(256) k_check_indices(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, size_t n, size_t a_lim, size_t b_lim,
                                                       uint32_t *__restrict__ bad) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    if (a[k] >= a_lim || b[k] >= b_lim) *bad = 1;
}
// dst[dst_idx[k]] = table[src_idx[k]]: the sparse overrides of an otherwise regular evaluation table
// (Permutation::to_poly, libs/src/iotools/mod.rs:419-455).
__global__ void __launch_bounds__(256) k_scatter_from_table(Fr *__restrict__ dst, const uint32_t *__restrict__ dst_idx, const Fr *__restrict__ table,
                                                            const uint32_t *__restrict__ src_idx, size_t n) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) dst[dst_idx[k]] = table[src_idx[k]];
}

// out[k] = table[idx[k]] (out-of-range indices are flagged and read nothing): the scalar side of the sparse-gather MSMs.
__global__ void __launch_bounds__(256) k_fr_gather(Fr *__restrict__ out, const Fr *__restrict__ table, size_t table_len, const uint32_t *__restrict__ idx, size_t n,
                                                   uint32_t *__restrict__ bad) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    const uint32_t i = idx[k];
    if (i >= table_len) {
      *bad = 1;
      out[k] = Fr::zero();
    } else {
      out[k] = table[i];
    }
  }
}

// ---- micro-benchmarks: dependent-free integer streams and field-op rates (ops/s over the whole GPU)
template <int KIND>
__global__ void __launch_bounds__(256) k_microbench(uint32_t *sink, int iters) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (KIND == 0) {  // 8 independent IMAD chains
    uint32_t a[8];
    for (int i = 0; i < 8; i++) a[i] = t + i;
    uint32_t m = t | 1u;
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = a[i] * m + a[(i + 1) & 7];
    uint32_t r = 0;
    for (int i = 0; i < 8; i++) r ^= a[i];
    if (r == 0x12345678u) sink[0] = r;
  } else if (KIND == 1) {  // 16 independent IMAD.WIDE.U32 accumulators (64-bit addend, no carry flag); the multiplier changes
                           // every iteration so the product can not be hoisted out of the loop
    uint64_t a[16];
    uint32_t m[16];
    for (int i = 0; i < 16; i++) { a[i] = (uint64_t)t * (i + 3); m[i] = (t * 2654435761u + i) | 1u; }
    uint32_t q = t | 3u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[i]) : "r"(m[i]), "r"(q));
      q ^= (uint32_t)it;
    }
    uint64_t r = 0;
    for (int i = 0; i < 16; i++) r ^= a[i];
    if (r == 0x12345678u) sink[0] = (uint32_t)r;
  } else if (KIND == 5) {  // the form the field multiplier issues: mad.lo.cc/madc.hi.cc pairs = IMAD.WIDE.U32.X carry chains,
                           // two independent chains of 8 wide multiply-adds per iteration
    uint32_t E[16], O[16], a[8], b = t | 1u;
    for (int i = 0; i < 16; i++) { E[i] = t + i; O[i] = t ^ i; }
    for (int i = 0; i < 8; i++) a[i] = t * (2 * i + 1);
    for (int it = 0; it < iters; it++) {
      E[0] = mad_lo_cc(a[0], b, E[0]); E[1] = madc_hi_cc(a[0], b, E[1]);
#pragma unroll
      for (int j = 1; j < 8; j++) { E[2 * j] = madc_lo_cc(a[j], b, E[2 * j]); E[2 * j + 1] = madc_hi_cc(a[j], b, E[2 * j + 1]); }
      O[0] = mad_lo_cc(a[0], b, O[0]); O[1] = madc_hi_cc(a[0], b, O[1]);
#pragma unroll
      for (int j = 1; j < 8; j++) { O[2 * j] = madc_lo_cc(a[j], b, O[2 * j]); O[2 * j + 1] = madc_hi_cc(a[j], b, O[2 * j + 1]); }
      b ^= (uint32_t)it;
    }
    uint32_t r = 0;
    for (int i = 0; i < 16; i++) r ^= E[i] ^ O[i];
    if (r == 0x12345678u) sink[0] = r;
  } else if (KIND == 6) {  // the NTT butterfly: one Fr product, one addition, one subtraction
    Fr x = Fr::one(), y = Fr::r2(), w = Fr::r2();
    x.v[0] ^= t;
    w.v[1] ^= t;
    for (int it = 0; it < iters; it++) {
      Fr s = x + y;
      Fr d = (x - y) * w;
      x = s;
      y = d;
    }
    if (x.v[0] == 0x12345678u && y.v[1] == 1) sink[0] = x.v[1];
  } else if (KIND == 7) {  // the same butterfly on the round-1 reduction (add-with-carry chains on the ALU pipe), for comparison
    using FrOld = Fp<FrParamsAddChains>;
    FrOld x = FrOld::one(), y = FrOld::r2(), w = FrOld::r2();
    x.v[0] ^= t;
    w.v[1] ^= t;
    for (int it = 0; it < iters; it++) {
      FrOld s = x + y;
      FrOld d = (x - y) * w;
      x = s;
      y = d;
    }
    if (x.v[0] == 0x12345678u && y.v[1] == 1) sink[0] = x.v[1];
  } else if (KIND == 8 || KIND == 9 || KIND == 10) {  // latency of ONE thread's Fq inversion chain (what a pair-tree level waits for)
    if (t != 0) return;
    Fq x = Fq::r2();
    x.v[0] ^= (uint32_t)iters;
    for (int it = 0; it < iters; it++) {
      Fq y = KIND == 8 ? x.inv_bgcd() : KIND == 9 ? x.inv_fast() : x.inv();
      x = y + Fq::one();
    }
    if (x.v[0] == 0x12345678u) sink[0] = x.v[1];
  } else if (KIND == 2) {
    Fr x = Fr::one(), y = Fr::r2();
    x.v[0] ^= t;
    for (int it = 0; it < iters; it++) { x = x * y; y = y * x; }
    if (x.v[0] == 0x12345678u && y.v[1] == 1) sink[0] = x.v[1];
  } else if (KIND == 3) {
    Fq x = Fq::one(), y = Fq::r2();
    x.v[0] ^= t;
    for (int it = 0; it < iters; it++) { x = x * y; y = y * x; }
    if (x.v[0] == 0x12345678u && y.v[1] == 1) sink[0] = x.v[1];
  } else {
    G1Xyzz acc = G1Xyzz::identity();
    G1Affine p;
    p.x = Fq::one(); p.y = Fq::r2();
    p.x.v[0] ^= t;
    acc.X = Fq::r2(); acc.Y = Fq::one(); acc.ZZ = Fq::one(); acc.ZZZ = Fq::one();
    for (int it = 0; it < iters; it++) { g1_madd(acc, p); p.x.v[1] ^= acc.X.v[0]; }
    if (acc.X.v[0] == 0x12345678u && acc.Y.v[1] == 1) sink[0] = acc.ZZ.v[1];
  }
}
}  // namespace tkm

using namespace tkm;

#define API_BEGIN                                                  \
  if (!ctx) return fail(TKM_ERR_INVALID_ARGUMENT, "null context"); \
  cudaSetDevice(ctx->device);

extern "C" {

const char *tkm_last_error(void) { return last_error().c_str(); }
const char *tkm_version(void) { return "tokamak_b200 0.1.0 (sm_100a)"; }

int32_t tkm_ctx_create(int32_t device_ordinal, tkm_ctx **out) {
  if (!out) return fail(TKM_ERR_INVALID_ARGUMENT, "null out pointer");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(TKM_ERR_NO_DEVICE, "no CUDA device available (%s); libtokamak_b200 has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device_ordinal < 0 || device_ordinal >= count) return fail(TKM_ERR_INVALID_ARGUMENT, "device ordinal %d out of range [0,%d)", device_ordinal, count);
  TKM_CUDA(cudaSetDevice(device_ordinal));
  tkm_ctx *ctx = new (std::nothrow) tkm_ctx();
  if (!ctx) return fail(TKM_ERR_ALLOCATION, "out of host memory");
  ctx->device = device_ordinal;
  cudaDeviceProp prop;
  TKM_CUDA(cudaGetDeviceProperties(&prop, device_ordinal));
  ctx->sm_count = prop.multiProcessorCount;
  TKM_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  TKM_CUDA(cudaEventCreate(&ctx->ev0));
  TKM_CUDA(cudaEventCreate(&ctx->ev1));
  TKM_CUDA(cudaEventCreate(&ctx->kev0));
  TKM_CUDA(cudaEventCreate(&ctx->kev1));
  TKM_CUDA(cudaMallocHost((void **)&ctx->tree_counts, 16 * sizeof(uint32_t)));
  memset(ctx->tree_counts, 0, 16 * sizeof(uint32_t));
  TKM_CUDA(cudaEventCreate(&ctx->pev0));
  TKM_CUDA(cudaEventCreate(&ctx->pev1));
  // keep freed scratch in the stream-ordered pool instead of returning it to the driver
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device_ordinal) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  Fr two = Fr::one() + Fr::one();
  Fr inv2 = two.inv();
  ctx->inv_pow2[0] = Fr::one();
  for (int k = 1; k <= 32; k++) ctx->inv_pow2[k] = ctx->inv_pow2[k - 1] * inv2;
  *out = ctx;
  return TKM_OK;
}

int32_t tkm_ctx_destroy(tkm_ctx *ctx) {
  if (!ctx) return TKM_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->twiddles) cudaFree(ctx->twiddles);
  if (ctx->fb_table) cudaFree(ctx->fb_table);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->kev0) cudaEventDestroy(ctx->kev0);
  if (ctx->kev1) cudaEventDestroy(ctx->kev1);
  if (ctx->tree_counts) cudaFreeHost(ctx->tree_counts);
  if (ctx->tree_stream) {
    cudaStreamSynchronize(ctx->tree_stream);
    cudaStreamDestroy(ctx->tree_stream);
    cudaEventDestroy(ctx->tree_fork);
    cudaEventDestroy(ctx->tree_join);
  }
  if (ctx->pev0) cudaEventDestroy(ctx->pev0);
  if (ctx->pev1) cudaEventDestroy(ctx->pev1);
  if (ctx->side_stream) {
    cudaStreamSynchronize(ctx->side_stream);
    cudaStreamDestroy(ctx->side_stream);
  }
  for (tkm_ctx::Ticket &k : ctx->tickets) {
    if (k.parts) cudaFree(k.parts);
    if (k.host) cudaFreeHost(k.host);
    if (k.ready) cudaEventDestroy(k.ready);
    if (k.done) cudaEventDestroy(k.done);
  }
  if (ctx->copy_stream) {
    cudaStreamDestroy(ctx->copy_stream);
    for (cudaEvent_t e : ctx->copy_ev)
      if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->copy_t)
      if (e) cudaEventDestroy(e);
  }
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return TKM_OK;
}

int32_t tkm_ctx_set_stream(tkm_ctx *ctx, void *cuda_stream) {
  API_BEGIN
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return TKM_OK;
}

int32_t tkm_ctx_sync(tkm_ctx *ctx) {
  API_BEGIN
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  return TKM_OK;
}

int32_t tkm_dev_alloc(tkm_ctx *ctx, size_t bytes, void **out_dev) {
  API_BEGIN
  TKM_REQUIRE(out_dev, "null out pointer");
  // stream-ordered pool allocation (release threshold = infinity): no driver round trip after warm-up.  The
  // stream is synchronised so the pointer is immediately usable from any stream.
  cudaError_t e = cudaMallocAsync(out_dev, bytes ? bytes : 1, ctx->stream);
  if (e != cudaSuccess) return fail(TKM_ERR_ALLOCATION, "cudaMallocAsync(%zu) failed: %s", bytes, cudaGetErrorString(e));
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  return TKM_OK;
}
int32_t tkm_dev_free(tkm_ctx *ctx, void *dev) {
  API_BEGIN
  if (!dev) return TKM_OK;
  TKM_CUDA(cudaFreeAsync(dev, ctx->stream));
  return TKM_OK;
}
int32_t tkm_memcpy_h2d(tkm_ctx *ctx, void *dev, const void *host, size_t bytes) {
  API_BEGIN
  TKM_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  return TKM_OK;
}
int32_t tkm_memcpy_d2h(tkm_ctx *ctx, void *host, const void *dev, size_t bytes) {
  API_BEGIN
  TKM_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  return TKM_OK;
}

int32_t tkm_ntt_domain_init(tkm_ctx *ctx, uint32_t log2_size) {
  API_BEGIN
  return domain_init(ctx, log2_size);
}
int32_t tkm_ntt_domain_release(tkm_ctx *ctx) {
  API_BEGIN
  return domain_release(ctx);
}
int32_t tkm_ntt_domain_log2(tkm_ctx *ctx, int32_t *out_log2) {
  API_BEGIN
  TKM_REQUIRE(out_log2, "null out pointer");
  *out_log2 = ctx->domain_log2;
  return TKM_OK;
}
int32_t tkm_root_of_unity(uint32_t log2_n, uint8_t out32[32]) {
  if (log2_n > 32) return fail(TKM_ERR_INVALID_ARGUMENT, "no 2^%u-th root of unity in Fr (2-adicity 32)", log2_n);
  fr_to_bytes_host(root_of_unity_host(log2_n), out32);
  return TKM_OK;
}

int32_t tkm_fr_to_mont(tkm_ctx *ctx, const void *in, void *out, size_t n) {
  API_BEGIN
  return vec_to_mont(ctx, (const Fr *)in, (Fr *)out, n);
}
int32_t tkm_fr_from_mont(tkm_ctx *ctx, const void *in, void *out, size_t n) {
  API_BEGIN
  return vec_from_mont(ctx, (const Fr *)in, (Fr *)out, n);
}
int32_t tkm_fr_vec_op(tkm_ctx *ctx, int32_t op, const void *a, const void *b, void *out, size_t n) {
  API_BEGIN
  return vec_op(ctx, op, (const Fr *)a, (const Fr *)b, (Fr *)out, n);
}
int32_t tkm_fr_vec_scale(tkm_ctx *ctx, const uint8_t s32[32], const void *a, void *out, size_t n) {
  API_BEGIN
  TKM_REQUIRE(s32, "null scalar");
  return vec_scale(ctx, fr_from_bytes_host(s32), (const Fr *)a, (Fr *)out, n);
}
int32_t tkm_fr_vec_inv(tkm_ctx *ctx, const void *a, void *out, size_t n) {
  API_BEGIN
  return vec_inv(ctx, (const Fr *)a, (Fr *)out, n);
}
int32_t tkm_fr_vec_fill(tkm_ctx *ctx, const uint8_t s32[32], void *out, size_t n) {
  API_BEGIN
  TKM_REQUIRE(s32 && (out || n == 0), "null argument");
  return vec_fill(ctx, fr_from_bytes_host(s32), (Fr *)out, n);
}
int32_t tkm_fr_mul_x_minus_one(tkm_ctx *ctx, const void *in, void *out, size_t x_size, size_t y_size) {
  API_BEGIN
  TKM_REQUIRE(in && out, "null argument");
  return vec_mul_x_minus_one(ctx, (const Fr *)in, (Fr *)out, x_size, y_size);
}
int32_t tkm_fr_suffix_product(tkm_ctx *ctx, const void *in, void *out, size_t n) {
  API_BEGIN
  TKM_REQUIRE((in && out) || n == 0, "null argument");
  return vec_suffix_product(ctx, (const Fr *)in, (Fr *)out, n);
}
int32_t tkm_fr_vec_reduce(tkm_ctx *ctx, int32_t op, const void *a, const void *b, size_t n, uint8_t out32[32]) {
  API_BEGIN
  TKM_REQUIRE(out32 && (a || n == 0), "null argument");
  TKM_REQUIRE(op >= 0 && op <= 2, "unknown reduction %d", op);
  TKM_REQUIRE(op != 2 || b || n == 0, "inner product needs two vectors");
  Fr r;
  TKM_TRY(vec_reduce(ctx, op, (const Fr *)a, (const Fr *)b, n, &r));
  fr_to_bytes_host(r, out32);
  return TKM_OK;
}
int32_t tkm_fr_powers(tkm_ctx *ctx, const uint8_t base32[32], void *dev_out, size_t n) {
  API_BEGIN
  TKM_REQUIRE(base32 && (dev_out || n == 0), "null argument");
  if (n == 0) return TKM_OK;
  return fill_powers_public(ctx, (Fr *)dev_out, fr_from_bytes_host(base32), Fr::one(), n);
}
int32_t tkm_fr_scatter_from_table(tkm_ctx *ctx, void *dev_dst, size_t dst_len, const void *dev_dst_idx, const void *dev_table, size_t table_len,
                                  const void *dev_src_idx, size_t n) {
  API_BEGIN
  if (n == 0) return TKM_OK;
  TKM_REQUIRE(dev_dst && dev_dst_idx && dev_table && dev_src_idx, "null argument");
  // the index arrays live on the device: bounds are checked there before the scatter
  Scratch<uint32_t> bad;
  TKM_TRY(bad.alloc(ctx, 1));
  TKM_CUDA(cudaMemsetAsync(bad.p, 0, 4, ctx->stream));
  k_check_indices<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>((const uint32_t *)dev_dst_idx, (const uint32_t *)dev_src_idx, n, dst_len, table_len, bad.p);
  TKM_TRY(launch_check(ctx, "k_check_indices"));
  uint32_t h_bad = 0;
  TKM_CUDA(cudaMemcpyAsync(&h_bad, bad.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  TKM_REQUIRE(!h_bad, "scatter index out of range");
  k_scatter_from_table<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>((Fr *)dev_dst, (const uint32_t *)dev_dst_idx, (const Fr *)dev_table,
                                                                                (const uint32_t *)dev_src_idx, n);
  return launch_check(ctx, "k_scatter_from_table");
}
int32_t tkm_fr_gather(tkm_ctx *ctx, const void *dev_table, size_t table_len, const void *dev_idx, size_t n, void *dev_out) {
  API_BEGIN
  if (n == 0) return TKM_OK;
  TKM_REQUIRE(dev_table && dev_idx && dev_out, "null argument");
  Scratch<uint32_t> bad;
  TKM_TRY(bad.alloc(ctx, 1));
  TKM_CUDA(cudaMemsetAsync(bad.p, 0, 4, ctx->stream));
  k_fr_gather<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>((Fr *)dev_out, (const Fr *)dev_table, table_len, (const uint32_t *)dev_idx, n, bad.p);
  TKM_TRY(launch_check(ctx, "k_fr_gather"));
  uint32_t h_bad = 0;
  TKM_CUDA(cudaMemcpyAsync(&h_bad, bad.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  TKM_REQUIRE(!h_bad, "gather index out of range");
  return TKM_OK;
}
int32_t tkm_fr_outer_product(tkm_ctx *ctx, const void *col, const void *row, void *out, size_t rows, size_t cols) {
  API_BEGIN
  TKM_REQUIRE((col && row && out) || rows * cols == 0, "null argument");
  return vec_outer_product(ctx, (const Fr *)col, (const Fr *)row, (Fr *)out, rows, cols);
}
int32_t tkm_fr_transpose(tkm_ctx *ctx, const void *in, void *out, size_t rows, size_t cols) {
  API_BEGIN
  TKM_REQUIRE(in && out, "null argument");
  return vec_transpose(ctx, (const Fr *)in, (Fr *)out, rows, cols);
}
int32_t tkm_fr_vec_op_host(tkm_ctx *ctx, int32_t op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n) {
  API_BEGIN
  TKM_REQUIRE(a && b && out, "null argument");
  Scratch<Fr> da, db;
  TKM_TRY(da.alloc(ctx, n));
  TKM_TRY(db.alloc(ctx, n));
  TKM_CUDA(cudaMemcpyAsync(da.p, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  TKM_CUDA(cudaMemcpyAsync(db.p, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (op == TKM_OP_MUL || op == TKM_OP_DIV) {
    // (aR)*b/R = ab ; (aR) * (bR)^-1 ... keep it simple: both to Montgomery, result back
    TKM_TRY(vec_to_mont(ctx, da.p, da.p, n));
    TKM_TRY(vec_to_mont(ctx, db.p, db.p, n));
    TKM_TRY(vec_op(ctx, op, da.p, db.p, da.p, n));
    TKM_TRY(vec_from_mont(ctx, da.p, da.p, n));
  } else {
    TKM_TRY(vec_op(ctx, op, da.p, db.p, da.p, n));  // add/sub are form-agnostic
  }
  TKM_CUDA(cudaMemcpyAsync(out, da.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  return TKM_OK;
}

int32_t tkm_bintt(tkm_ctx *ctx, const void *in, void *out, size_t x_size, size_t y_size, int32_t dir, const uint8_t *cx, const uint8_t *cy) {
  API_BEGIN
  TKM_REQUIRE(in && out, "null argument");
  Fr gx, gy;
  if (cx) gx = fr_from_bytes_host(cx);
  if (cy) gy = fr_from_bytes_host(cy);
  return bintt_dev(ctx, (const Fr *)in, (Fr *)out, x_size, y_size, dir, cx ? &gx : nullptr, cy ? &gy : nullptr);
}

int32_t tkm_bintt_host(tkm_ctx *ctx, const uint8_t *in, uint8_t *out, size_t x_size, size_t y_size, int32_t dir, const uint8_t *cx,
                       const uint8_t *cy) {
  API_BEGIN
  TKM_REQUIRE(in && out, "null argument");
  size_t n = x_size * y_size;
  Scratch<Fr> d;
  TKM_TRY(d.alloc(ctx, n));
  TKM_CUDA(cudaMemcpyAsync(d.p, in, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  TKM_TRY(vec_to_mont(ctx, d.p, d.p, n));
  TKM_TRY(tkm_bintt(ctx, d.p, d.p, x_size, y_size, dir, cx, cy));
  TKM_TRY(vec_from_mont(ctx, d.p, d.p, n));
  TKM_CUDA(cudaMemcpyAsync(out, d.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  return TKM_OK;
}

int32_t tkm_ntt_batch(tkm_ctx *ctx, const void *in, void *out, size_t n, size_t batch, int32_t columns_batch, int32_t dir,
                      const uint8_t *coset32) {
  API_BEGIN
  TKM_REQUIRE(in && out, "null argument");
  Fr g;
  if (coset32) g = fr_from_bytes_host(coset32);
  if (columns_batch) return ntt_axis(ctx, (const Fr *)in, (Fr *)out, 1, n, batch, dir, coset32 ? &g : nullptr);
  return ntt_axis(ctx, (const Fr *)in, (Fr *)out, batch, n, 1, dir, coset32 ? &g : nullptr);
}

int32_t tkm_ntt_batch_scatter(tkm_ctx *ctx, const void *in, size_t n, size_t batch, int32_t columns_batch, int32_t dir, const uint8_t *coset32,
                              void *const *peer_out, uint32_t n_peers, uint64_t stride_a, uint64_t stride_b, uint64_t b0) {
  API_BEGIN
  TKM_REQUIRE(in && peer_out, "null argument");
  TKM_REQUIRE(dir == TKM_FORWARD || dir == TKM_INVERSE, "bad direction");
  for (uint32_t i = 0; i < n_peers; i++) TKM_REQUIRE(peer_out[i], "null peer buffer");
  Fr g;
  if (coset32) g = fr_from_bytes_host(coset32);
  if (columns_batch) return ntt_axis_scatter(ctx, (const Fr *)in, 1, n, batch, dir, coset32 ? &g : nullptr, peer_out, n_peers, stride_a, stride_b, b0);
  return ntt_axis_scatter(ctx, (const Fr *)in, batch, n, 1, dir, coset32 ? &g : nullptr, peer_out, n_peers, stride_a, stride_b, b0);
}

int32_t tkm_g1_bases_to_mont(tkm_ctx *ctx, const void *in, void *out, size_t n) {
  API_BEGIN
  return g1_to_mont_dev(ctx, (const G1Affine *)in, (G1Affine *)out, n);
}

int32_t tkm_g1_bases_from_mont(tkm_ctx *ctx, const void *in, void *out, size_t n) {
  API_BEGIN
  return g1_from_mont_dev(ctx, (const G1Affine *)in, (G1Affine *)out, n);
}

int32_t tkm_msm_g1_rect(tkm_ctx *ctx, const void *scalars, int32_t scalars_mont, size_t s_stride, const void *bases, size_t b_stride,
                        size_t rows, size_t cols, uint8_t out96[96]) {
  API_BEGIN
  TKM_REQUIRE(out96, "null out pointer");
  TKM_REQUIRE(rows * cols == 0 || (scalars && bases), "null argument");
  MsmInput in;
  in.scalars = (const Fr *)scalars;
  in.scalars_mont = scalars_mont != 0;
  in.scalar_row_stride = s_stride;
  in.bases = (const G1Affine *)bases;
  in.base_row_stride = b_stride;
  in.rows = rows;
  in.cols = cols;
  in.idx = nullptr;
  return msm_run(ctx, in, out96);
}
int32_t tkm_msm_g1(tkm_ctx *ctx, const void *scalars, int32_t scalars_mont, const void *bases, size_t n, uint8_t out96[96]) {
  return tkm_msm_g1_rect(ctx, scalars, scalars_mont, n, bases, n, 1, n, out96);
}
int32_t tkm_msm_g1_indexed(tkm_ctx *ctx, const void *scalars, int32_t scalars_mont, const void *bases, const void *idx, size_t n,
                           uint8_t out96[96]) {
  API_BEGIN
  TKM_REQUIRE(out96, "null out pointer");
  TKM_REQUIRE(n == 0 || (scalars && bases && idx), "null argument");
  MsmInput in;
  in.scalars = (const Fr *)scalars;
  in.scalars_mont = scalars_mont != 0;
  in.scalar_row_stride = n;
  in.bases = (const G1Affine *)bases;
  in.base_row_stride = n;
  in.rows = 1;
  in.cols = n;
  in.idx = (const uint32_t *)idx;
  return msm_run(ctx, in, out96);
}
static int32_t msm_begin_common(tkm_ctx *ctx, const void *scalars, int32_t scalars_mont, const void *bases, const void *idx, size_t n, int32_t *out_ticket) {
  MsmInput in;
  in.scalars = (const Fr *)scalars;
  in.scalars_mont = scalars_mont != 0;
  in.scalar_row_stride = n;
  in.bases = (const G1Affine *)bases;
  in.base_row_stride = n;
  in.rows = 1;
  in.cols = n;
  in.idx = (const uint32_t *)idx;
  return msm_run_async(ctx, in, out_ticket);
}
int32_t tkm_msm_g1_begin(tkm_ctx *ctx, const void *scalars, int32_t scalars_mont, const void *bases, size_t n, int32_t *out_ticket) {
  API_BEGIN
  TKM_REQUIRE(out_ticket, "null out pointer");
  TKM_REQUIRE(n == 0 || (scalars && bases), "null argument");
  return msm_begin_common(ctx, scalars, scalars_mont, bases, nullptr, n, out_ticket);
}
int32_t tkm_msm_g1_indexed_begin(tkm_ctx *ctx, const void *scalars, int32_t scalars_mont, const void *bases, const void *idx, size_t n,
                                 int32_t *out_ticket) {
  API_BEGIN
  TKM_REQUIRE(out_ticket, "null out pointer");
  TKM_REQUIRE(n == 0 || (scalars && bases && idx), "null argument");
  return msm_begin_common(ctx, scalars, scalars_mont, bases, idx, n, out_ticket);
}
int32_t tkm_msm_g1_host(tkm_ctx *ctx, const uint8_t *scalars, const uint8_t *bases, size_t n, uint8_t out96[96]) {
  API_BEGIN
  TKM_REQUIRE(out96, "null out pointer");
  if (n == 0) {
    memset(out96, 0, 96);
    return TKM_OK;
  }
  TKM_REQUIRE(scalars && bases, "null argument");
  // Large inputs: overlap the host->device copies with the bucket accumulation, one point range at a time; the piece layout
  // (msm_host_pipelined) follows the share of the previous call its copies took.
  uint32_t pieces = 0;
  if (const char *e = getenv("TKM_MSM_HOST_PIECES")) pieces = (uint32_t)atoi(e);  // developer knob
  return msm_host_pipelined(ctx, scalars, bases, n, pieces, out96);
}

int32_t tkm_g1_fixed_base_mul(tkm_ctx *ctx, const uint8_t base96[96], const void *scalars, int32_t scalars_mont, size_t n, void *out) {
  API_BEGIN
  TKM_REQUIRE(base96 && out, "null argument");
  if (n == 0) return TKM_OK;
  TKM_REQUIRE(scalars, "null scalars");
  // the 64 x 16 window table of the most recent base stays on the context: setup multiplies one generator millions of
  // times across dozens of calls (Sigma1::gen), and building the table is a serial 20 ms job
  if (!ctx->fb_table) TKM_CUDA(cudaMalloc((void **)&ctx->fb_table, FB_WINDOWS * 16 * sizeof(G1Affine)));
  if (!ctx->fb_valid || memcmp(ctx->fb_base, base96, 96) != 0) {
    G1Affine b;
    memcpy(&b, base96, 96);
    ctx->fb_valid = false;
    k_fixed_base_table<<<1, 32, 0, ctx->stream>>>(b, ctx->fb_table);
    TKM_TRY(launch_check(ctx, "k_fixed_base_table"));
    memcpy(ctx->fb_base, base96, 96);
    ctx->fb_valid = true;
  }
  k_fixed_base_mul<<<grid_for(n, 128, ctx->sm_count), 128, 0, ctx->stream>>>(ctx->fb_table, (const Fr *)scalars, scalars_mont, n,
                                                                            (G1Affine *)out);
  return launch_check(ctx, "k_fixed_base_mul");
}

int32_t tkm_g1_add(tkm_ctx *ctx, const uint8_t a96[96], const uint8_t b96[96], uint8_t out96[96]) {
  API_BEGIN
  TKM_REQUIRE(a96 && b96 && out96, "null argument");
  Scratch<G1Affine> d;
  Scratch<uint32_t> r;
  TKM_TRY(d.alloc(ctx, 2));
  TKM_TRY(r.alloc(ctx, 24));
  TKM_CUDA(cudaMemcpyAsync(d.p, a96, 96, cudaMemcpyHostToDevice, ctx->stream));
  TKM_CUDA(cudaMemcpyAsync(d.p + 1, b96, 96, cudaMemcpyHostToDevice, ctx->stream));
  k_g1_add_single<<<1, 32, 0, ctx->stream>>>(d.p, d.p + 1, r.p);
  TKM_TRY(launch_check(ctx, "k_g1_add_single"));
  TKM_CUDA(cudaMemcpyAsync(out96, r.p, 96, cudaMemcpyDeviceToHost, ctx->stream));
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  return TKM_OK;
}

int32_t tkm_g1_sum(tkm_ctx *ctx, const uint8_t *points96, size_t n, uint8_t out96[96]) {
  API_BEGIN
  TKM_REQUIRE(out96 && (points96 || n == 0), "null argument");
  if (n == 0) {
    memset(out96, 0, 96);
    return TKM_OK;
  }
  Scratch<G1Affine> d;
  Scratch<uint32_t> r;
  TKM_TRY(d.alloc(ctx, n));
  TKM_TRY(r.alloc(ctx, 24));
  TKM_CUDA(cudaMemcpyAsync(d.p, points96, n * 96, cudaMemcpyHostToDevice, ctx->stream));
  k_g1_sum_canonical<<<1, 32, 0, ctx->stream>>>(d.p, n, r.p);
  TKM_TRY(launch_check(ctx, "k_g1_sum_canonical"));
  TKM_CUDA(cudaMemcpyAsync(out96, r.p, 96, cudaMemcpyDeviceToHost, ctx->stream));
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  return TKM_OK;
}

int32_t tkm_g1_mul(tkm_ctx *ctx, const uint8_t a96[96], const uint8_t k32[32], uint8_t out96[96]) {
  API_BEGIN
  TKM_REQUIRE(a96 && k32 && out96, "null argument");
  // A base whose window table is already on the context (setup multiplies one generator many times) uses it; any other
  // base is a one-point MSM (GLV halves, ~1.5 ms) -- building a 64 x 16 table for one product would be a serial 20 ms job,
  // and would evict the generator's table.  (G1serde * ScalarField, group_structures/mod.rs:929-947.)
  if (!ctx->fb_valid || memcmp(ctx->fb_base, a96, 96) != 0) return tkm_msm_g1_host(ctx, k32, a96, 1, out96);
  Scratch<Fr> s;
  Scratch<G1Affine> r;
  TKM_TRY(s.alloc(ctx, 1));
  TKM_TRY(r.alloc(ctx, 1));
  TKM_CUDA(cudaMemcpyAsync(s.p, k32, 32, cudaMemcpyHostToDevice, ctx->stream));
  TKM_TRY(tkm_g1_fixed_base_mul(ctx, a96, s.p, 0, 1, r.p));
  TKM_CUDA(cudaMemcpyAsync(out96, r.p, 96, cudaMemcpyDeviceToHost, ctx->stream));
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  return TKM_OK;
}

int32_t tkm_crs_from_device(tkm_ctx *ctx, void *dev_points, size_t rows, size_t cols, int32_t take_ownership, tkm_crs **out) {
  API_BEGIN
  TKM_REQUIRE(dev_points && out, "null argument");
  tkm_crs *c = new (std::nothrow) tkm_crs();
  if (!c) return fail(TKM_ERR_ALLOCATION, "out of host memory");
  c->d = (G1Affine *)dev_points;
  c->rows = rows;
  c->cols = cols;
  c->owned = take_ownership != 0;
  int32_t st = g1_to_mont_dev(ctx, c->d, c->d, rows * cols);
  if (st != TKM_OK) {
    delete c;
    return st;
  }
  *out = c;
  return TKM_OK;
}
int32_t tkm_crs_upload(tkm_ctx *ctx, const uint8_t *points96, size_t rows, size_t cols, tkm_crs **out) {
  API_BEGIN
  TKM_REQUIRE(points96 && out, "null argument");
  void *d = nullptr;
  cudaError_t e = cudaMalloc(&d, rows * cols * 96 + 16);
  if (e != cudaSuccess) return fail(TKM_ERR_ALLOCATION, "cudaMalloc(%zu) failed: %s", rows * cols * 96, cudaGetErrorString(e));
  e = cudaMemcpyAsync(d, points96, rows * cols * 96, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) {
    cudaFree(d);
    return fail(TKM_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
  }
  int32_t st = tkm_crs_from_device(ctx, d, rows, cols, 1, out);
  if (st != TKM_OK) cudaFree(d);
  return st;
}
int32_t tkm_crs_upload_mont(tkm_ctx *ctx, const uint8_t *points96_mont, size_t rows, size_t cols, tkm_crs **out) {
  API_BEGIN
  TKM_REQUIRE(points96_mont && out, "null argument");
  void *d = nullptr;
  cudaError_t e = cudaMalloc(&d, rows * cols * 96 + 16);
  if (e != cudaSuccess) return fail(TKM_ERR_ALLOCATION, "cudaMalloc(%zu) failed: %s", rows * cols * 96, cudaGetErrorString(e));
  e = cudaMemcpyAsync(d, points96_mont, rows * cols * 96, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    cudaFree(d);
    return fail(TKM_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
  }
  tkm_crs *c = new (std::nothrow) tkm_crs();
  if (!c) {
    cudaFree(d);
    return fail(TKM_ERR_ALLOCATION, "out of host memory");
  }
  c->d = (G1Affine *)d;
  c->rows = rows;
  c->cols = cols;
  c->owned = true;
  *out = c;
  return TKM_OK;
}
int32_t tkm_crs_precompute(tkm_ctx *ctx, tkm_crs *crs, uint32_t window_bits) {
  API_BEGIN
  TKM_REQUIRE(crs, "null crs");
  if (crs->pre) {
    if (crs->pre_c == window_bits) return TKM_OK;
    TKM_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(crs->pre);
    crs->pre = nullptr;
    if (crs->xpad) cudaFree(crs->xpad);
    crs->xpad = nullptr;
  }
  TKM_TRY(crs_precompute(ctx, crs->d, crs->rows * crs->cols, window_bits, &crs->pre, &crs->pre_W));
  crs->pre_c = window_bits;
  return TKM_OK;
}
int32_t tkm_crs_free(tkm_ctx *ctx, tkm_crs *crs) {
  API_BEGIN
  if (!crs) return TKM_OK;
  cudaStreamSynchronize(ctx->stream);
  if (crs->pre) cudaFree(crs->pre);
  if (crs->xpad) cudaFree(crs->xpad);
  if (crs->owned && crs->d) cudaFree(crs->d);
  delete crs;
  return TKM_OK;
}
int32_t tkm_crs_device_ptr(tkm_crs *crs, void **out_dev, size_t *rows, size_t *cols) {
  if (!crs) return fail(TKM_ERR_INVALID_ARGUMENT, "null crs");
  if (out_dev) *out_dev = crs->d;
  if (rows) *rows = crs->rows;
  if (cols) *cols = crs->cols;
  return TKM_OK;
}

// Host-side data loader: every "0x..." string of a JSON text (placementVariables.json is 40 MB of them) as 32-byte
// canonical little-endian scalars, in file order.  Pure host code, no context: the counterpart of the reference's
// HexString -> ScalarField::from_hex parsing (libs/src/iotools/mod.rs:126-146,1582-1588), which dominates its witness load.
int32_t tkm_host_parse_hex_scalars(const char *text, size_t len, uint8_t *out32, size_t capacity, size_t *out_count) {
  if (!text || !out_count || (capacity && !out32)) return fail(TKM_ERR_INVALID_ARGUMENT, "null argument");
  static const uint64_t R[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
  static const int8_t HEX[256] = {
      -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
      -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  -1, -1, -1, -1, -1, -1,
      -1, 10, 11, 12, 13, 14, 15, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
      -1, 10, 11, 12, 13, 14, 15, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
      -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
      -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
      -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
      -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
  size_t count = 0;
  const char *end = text + len;
  for (const char *q = (const char *)memchr(text, '"', len); q && q + 3 < end; q = (const char *)memchr(q, '"', (size_t)(end - q))) {
    if (q[1] != '0' || (q[2] != 'x' && q[2] != 'X')) {
      q++;  // some other string (a key): skip to its closing quote on the next iterations
      continue;
    }
    const char *first = q + 3;
    const char *close = (const char *)memchr(first, '"', (size_t)(end - first));
    if (!close) return fail(TKM_ERR_INVALID_ARGUMENT, "unterminated string at byte %zu", (size_t)(q - text));
    while (first < close && *first == '0') first++;  // leading zeros carry no information
    const size_t digits = (size_t)(close - first);
    if (digits > 64) return fail(TKM_ERR_INVALID_ARGUMENT, "hex scalar at byte %zu exceeds 256 bits", (size_t)(q - text));
    uint64_t v[4] = {0, 0, 0, 0};
    for (size_t t = 0; t < digits; t++) {
      const int d = HEX[(unsigned char)first[t]];
      if (d < 0) return fail(TKM_ERR_INVALID_ARGUMENT, "invalid hex digit at byte %zu", (size_t)(first + t - text));
      const size_t pos = digits - 1 - t;  // nibble index from the least significant end
      v[pos >> 4] |= (uint64_t)d << ((pos & 15) * 4);
    }
    const size_t i = (size_t)(q - text);
    (void)i;
    // reduce into [0, r): at most two subtractions for a 256-bit value
    for (int pass = 0; pass < 3; pass++) {
      bool ge = true;
      for (int k = 3; k >= 0; k--) {
        if (v[k] != R[k]) {
          ge = v[k] > R[k];
          break;
        }
      }
      if (!ge) break;
      unsigned __int128 borrow = 0;
      for (int k = 0; k < 4; k++) {
        unsigned __int128 dd = (unsigned __int128)v[k] - R[k] - borrow;
        v[k] = (uint64_t)dd;
        borrow = (dd >> 64) & 1;
      }
    }
    if (count >= capacity) return fail(TKM_ERR_INVALID_ARGUMENT, "more than %zu hex scalars in the text", capacity);
    memcpy(out32 + count * 32, v, 32);
    count++;
    q = close + 1;
  }
  *out_count = count;
  return TKM_OK;
}

// iden3 .r1cs binary -> CSR of the A, B, C matrices (R1csBinary::read + scan_constraints, libs/src/iotools/mod.rs:505-650).
// Host only.  Pass 1 (row_ptr == NULL) reports the header and the entry counts; pass 2 fills row_ptr[3][n_constraints + 1]
// (entry indices into wire / coeff32, matrix-major: all of A, then B, then C), wire[] and the canonical 32-byte coefficients.
int32_t tkm_host_parse_r1cs(const uint8_t *data, size_t len, uint32_t *n_wires, uint32_t *n_constraints, size_t nnz[3], uint32_t *row_ptr,
                            uint32_t *wire, uint8_t *coeff32) {
  if (!data || !n_wires || !n_constraints || !nnz) return fail(TKM_ERR_INVALID_ARGUMENT, "null argument");
  auto rd32 = [&](size_t off, uint32_t *v) {
    if (off + 4 > len) return false;
    memcpy(v, data + off, 4);
    return true;
  };
  auto rd64 = [&](size_t off, uint64_t *v) {
    if (off + 8 > len) return false;
    memcpy(v, data + off, 8);
    return true;
  };
  if (len < 12 || memcmp(data, "r1cs", 4) != 0) return fail(TKM_ERR_INVALID_ARGUMENT, "invalid R1CS magic");
  uint32_t version, nsec;
  rd32(4, &version);
  rd32(8, &nsec);
  if (version != 1) return fail(TKM_ERR_INVALID_ARGUMENT, "unsupported R1CS version %u", version);
  size_t off = 12, hdr = 0, cons = 0, cons_size = 0;
  bool have_h = false, have_c = false;
  for (uint32_t k = 0; k < nsec; k++) {
    uint32_t typ;
    uint64_t size;
    if (!rd32(off, &typ) || !rd64(off + 4, &size)) return fail(TKM_ERR_INVALID_ARGUMENT, "truncated R1CS section table");
    off += 12;
    if (size > len - off) return fail(TKM_ERR_INVALID_ARGUMENT, "R1CS section extends past end of file");
    if (typ == 1 && !have_h) { hdr = off; have_h = true; }
    if (typ == 2 && !have_c) { cons = off; cons_size = (size_t)size; have_c = true; }
    off += (size_t)size;
  }
  if (!have_h || !have_c) return fail(TKM_ERR_INVALID_ARGUMENT, "missing R1CS header or constraints section");
  uint32_t fs;
  if (!rd32(hdr, &fs) || fs == 0 || fs % 8 || fs > 32) return fail(TKM_ERR_INVALID_ARGUMENT, "invalid R1CS field size");
  uint32_t nw, nc;
  if (!rd32(hdr + 4 + fs, &nw) || !rd32(hdr + 4 + fs + 4 * 4 + 8, &nc)) return fail(TKM_ERR_INVALID_ARGUMENT, "truncated R1CS header");
  *n_wires = nw;
  *n_constraints = nc;
  static const uint64_t R[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
  size_t cnt[3] = {0, 0, 0};
  size_t base[3] = {0, 0, 0};
  if (row_ptr) {
    base[1] = nnz[0];
    base[2] = nnz[0] + nnz[1];
  }
  size_t c = cons;
  const size_t cend = cons + cons_size;
  for (uint32_t row = 0; row < nc; row++) {
    for (int m = 0; m < 3; m++) {
      uint32_t n_ent;
      if (c + 4 > cend || !rd32(c, &n_ent)) return fail(TKM_ERR_INVALID_ARGUMENT, "truncated R1CS constraints");
      c += 4;
      if ((size_t)n_ent * (4 + fs) > cend - c) return fail(TKM_ERR_INVALID_ARGUMENT, "truncated R1CS constraints");
      if (row_ptr) row_ptr[(size_t)m * (nc + 1) + row] = (uint32_t)(base[m] + cnt[m]);
      for (uint32_t e = 0; e < n_ent; e++) {
        uint32_t w;
        rd32(c, &w);
        if (w >= nw) return fail(TKM_ERR_INVALID_ARGUMENT, "R1CS wire index %u exceeds nWires %u", w, nw);
        if (row_ptr) {
          const size_t slot = base[m] + cnt[m];
          if (cnt[m] >= nnz[m]) return fail(TKM_ERR_INVALID_ARGUMENT, "R1CS entry counts changed between passes");
          wire[slot] = w;
          uint64_t v[4] = {0, 0, 0, 0};
          memcpy(v, data + c + 4, fs);
          for (int pass = 0; pass < 3; pass++) {
            bool ge = true;
            for (int k = 3; k >= 0; k--)
              if (v[k] != R[k]) {
                ge = v[k] > R[k];
                break;
              }
            if (!ge) break;
            unsigned __int128 borrow = 0;
            for (int k = 0; k < 4; k++) {
              unsigned __int128 dd = (unsigned __int128)v[k] - R[k] - borrow;
              v[k] = (uint64_t)dd;
              borrow = (dd >> 64) & 1;
            }
          }
          memcpy(coeff32 + slot * 32, v, 32);
        }
        cnt[m]++;
        c += 4 + fs;
      }
    }
  }
  if (c != cend) return fail(TKM_ERR_INVALID_ARGUMENT, "R1CS constraints section has %zu trailing bytes", cend - c);
  if (row_ptr) {
    for (int m = 0; m < 3; m++) {
      if (cnt[m] != nnz[m]) return fail(TKM_ERR_INVALID_ARGUMENT, "R1CS entry counts changed between passes");
      row_ptr[(size_t)m * (nc + 1) + nc] = (uint32_t)(base[m] + cnt[m]);
    }
  } else {
    for (int m = 0; m < 3; m++) nnz[m] = cnt[m];
  }
  return TKM_OK;
}

int32_t tkm_event_time_begin(tkm_ctx *ctx) {
  API_BEGIN
  TKM_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  return TKM_OK;
}
int32_t tkm_event_time_end(tkm_ctx *ctx, float *out_ms) {
  API_BEGIN
  TKM_REQUIRE(out_ms, "null out pointer");
  TKM_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  TKM_CUDA(cudaEventSynchronize(ctx->ev1));
  TKM_CUDA(cudaEventElapsedTime(out_ms, ctx->ev0, ctx->ev1));
  return TKM_OK;
}
int32_t tkm_kernel_time_last(tkm_ctx *ctx, float *out_ms) {
  API_BEGIN
  TKM_REQUIRE(out_ms, "null out pointer");
  TKM_REQUIRE(ctx->kernel_timed, "no dominant-kernel launch has been timed on this context yet");
  TKM_CUDA(cudaEventSynchronize(ctx->kev1));
  TKM_CUDA(cudaEventElapsedTime(out_ms, ctx->kev0, ctx->kev1));
  return TKM_OK;
}
int32_t tkm_msm_tree_stats(tkm_ctx *ctx, uint32_t *out_levels, uint64_t out_counts[9]) {
  API_BEGIN
  TKM_REQUIRE(out_levels && out_counts, "null out pointer");
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  *out_levels = ctx->tree_levels;
  for (uint32_t l = 0; l < 9; l++) out_counts[l] = l <= ctx->tree_levels ? ctx->tree_counts[l] : 0;
  return TKM_OK;
}
int32_t tkm_poly_kernel_time_last(tkm_ctx *ctx, float *out_ms) {
  API_BEGIN
  TKM_REQUIRE(out_ms, "null out pointer");
  TKM_REQUIRE(ctx->poly_kernel_timed, "no polynomial-engine kernel has been timed on this context yet");
  TKM_CUDA(cudaEventSynchronize(ctx->pev1));
  TKM_CUDA(cudaEventElapsedTime(out_ms, ctx->pev0, ctx->pev1));
  return TKM_OK;
}
int32_t tkm_launch_count(tkm_ctx *ctx, uint64_t *out) {
  API_BEGIN
  TKM_REQUIRE(out, "null out pointer");
  *out = ctx->launches;
  return TKM_OK;
}

// Keccak-256 with the original 0x01 padding (tiny_keccak::Keccak::v256, the hash of the reference's Fiat-Shamir transcript,
// prove/src/lib.rs:3211-3519), on the host: a proof takes ~90 hashes of 100 bytes between its stages, and while the host
// hashes the device has nothing queued.
static void keccak_f1600(uint64_t a[25]) {
  static const uint64_t RC[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull,
                                  0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull,
                                  0x0000000080008009ull, 0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,
                                  0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
                                  0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};  // [x + 5y]
  auto rol = [](uint64_t v, int n) { return n ? (v << n) | (v >> (64 - n)) : v; };
  for (int round = 0; round < 24; round++) {
    uint64_t c[5], b[25];
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
    for (int x = 0; x < 5; x++) {
      const uint64_t d = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
      for (int y = 0; y < 5; y++) a[x + 5 * y] ^= d;
    }
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol(a[x + 5 * y], ROT[x + 5 * y]);
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    a[0] ^= RC[round];
  }
}
int32_t tkm_host_keccak256(const uint8_t *data, size_t len, uint8_t out32[32]) {
  if ((!data && len) || !out32) return fail(TKM_ERR_INVALID_ARGUMENT, "null argument");
  const size_t rate = 136;
  uint64_t a[25] = {0};
  auto absorb = [&](const uint8_t *blk) {
    for (size_t i = 0; i < rate / 8; i++) {
      uint64_t w = 0;
      for (int k = 7; k >= 0; k--) w = (w << 8) | blk[8 * i + k];
      a[i] ^= w;
    }
    keccak_f1600(a);
  };
  size_t off = 0;
  for (; off + rate <= len; off += rate) absorb(data + off);
  uint8_t last[136] = {0};
  if (len > off) memcpy(last, data + off, len - off);
  last[len - off] ^= 0x01;
  last[rate - 1] ^= 0x80;
  absorb(last);
  for (int i = 0; i < 4; i++)
    for (int k = 0; k < 8; k++) out32[8 * i + k] = (uint8_t)(a[i] >> (8 * k));
  return TKM_OK;
}

int32_t tkm_microbench(tkm_ctx *ctx, int32_t kind, double *out_ops_per_s) {
  API_BEGIN
  TKM_REQUIRE(out_ops_per_s, "null out pointer");
  Scratch<uint32_t> sink;
  TKM_TRY(sink.alloc(ctx, 4));
  const int blocks = ctx->sm_count * 8, threads = 256;
  int iters;
  double ops_per_iter;
  for (int rep = 0; rep < 2; rep++) {  // rep 0 = warm-up
    TKM_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    switch (kind) {
      case 0: iters = 4096; ops_per_iter = 8; k_microbench<0><<<blocks, threads, 0, ctx->stream>>>(sink.p, iters); break;
      case 1: iters = 4096; ops_per_iter = 16; k_microbench<1><<<blocks, threads, 0, ctx->stream>>>(sink.p, iters); break;
      case 2: iters = 512; ops_per_iter = 2; k_microbench<2><<<blocks, threads, 0, ctx->stream>>>(sink.p, iters); break;
      case 3: iters = 256; ops_per_iter = 2; k_microbench<3><<<blocks, threads, 0, ctx->stream>>>(sink.p, iters); break;
      case 4: iters = 64; ops_per_iter = 1; k_microbench<4><<<blocks, threads, 0, ctx->stream>>>(sink.p, iters); break;
      case 5: iters = 4096; ops_per_iter = 16; k_microbench<5><<<blocks, threads, 0, ctx->stream>>>(sink.p, iters); break;
      case 6: iters = 512; ops_per_iter = 1; k_microbench<6><<<blocks, threads, 0, ctx->stream>>>(sink.p, iters); break;
      case 7: iters = 512; ops_per_iter = 1; k_microbench<7><<<blocks, threads, 0, ctx->stream>>>(sink.p, iters); break;
      case 8: iters = 64; ops_per_iter = 1; k_microbench<8><<<1, 32, 0, ctx->stream>>>(sink.p, iters); break;
      case 9: iters = 64; ops_per_iter = 1; k_microbench<9><<<1, 32, 0, ctx->stream>>>(sink.p, iters); break;
      case 10: iters = 16; ops_per_iter = 1; k_microbench<10><<<1, 32, 0, ctx->stream>>>(sink.p, iters); break;
      default: return fail(TKM_ERR_INVALID_ARGUMENT, "unknown microbench kind %d", kind);
    }
    TKM_TRY(launch_check(ctx, "k_microbench"));
    TKM_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    TKM_CUDA(cudaEventSynchronize(ctx->ev1));
  }
  float ms = 0;
  TKM_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *out_ops_per_s = (kind >= 8 ? 1.0 : (double)blocks * threads) * iters * ops_per_iter / (ms * 1e-3);
  return TKM_OK;
}

}  // extern "C"
