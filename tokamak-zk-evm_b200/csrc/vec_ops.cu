// Element-wise Fr kernels: the device side of ICICLE's VecOps::{add,sub,mul,div,scalar_mul,inv}
// as the reference uses them (libs/src/vector_operations/mod.rs:19-141;
// libs/src/bivariate_polynomial/mod.rs:332-435,1974,2180).  HBM-bound: one 32-byte element per
// thread per array, 128-bit loads/stores, grid-stride loops sized to the SM count.
#include "common.cuh"

namespace tkm {

std::string &last_error() {
  static thread_local std::string e;
  return e;
}
int32_t fail(int32_t code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}
int32_t launch_check(tkm_ctx *ctx, const char *what) {
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TKM_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  return TKM_OK;
}

Fr fr_from_bytes_host(const uint8_t *b32) {
  Fr t;
  memcpy(t.v, b32, 32);
  return t.to_mont();
}
void fr_to_bytes_host(const Fr &a, uint8_t *b32) {
  Fr t = a.from_mont();
  memcpy(b32, t.v, 32);
}

enum { K_TO_MONT = 0, K_FROM_MONT = 1 };

template <int KIND>
__global__ void __launch_bounds__(256) k_mont_convert(const Fr *__restrict__ in, Fr *__restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fr a = in[i];
    out[i] = (KIND == K_TO_MONT) ? a.to_mont() : a.from_mont();
  }
}

// Single-element inverse helper shared by div and inv: Fermat, inv(0) = 0.
template <int OP>
__global__ void __launch_bounds__(256) k_vec_op(const Fr *__restrict__ a, const Fr *__restrict__ b, Fr *__restrict__ out,
                                                size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fr x = a[i], y = b[i], z;
    if (OP == TKM_OP_ADD) z = x + y;
    if (OP == TKM_OP_SUB) z = x - y;
    if (OP == TKM_OP_MUL) z = x * y;
    if (OP == TKM_OP_DIV) z = x * y.inv();
    out[i] = z;
  }
}

__global__ void __launch_bounds__(256) k_vec_scale(Fr s, const Fr *__restrict__ a, Fr *__restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fr x = a[i];
    out[i] = x * s;
  }
}

// Batched inversion: each thread inverts CH consecutive-by-stride elements with one Fermat inverse
// (Montgomery's trick).  Zeros are skipped and map to zero (ICICLE convention inv(0) = 0).
constexpr int INV_CH = 8;
__global__ void __launch_bounds__(128) k_vec_inv(const Fr *__restrict__ a, Fr *__restrict__ out, size_t n) {
  size_t nthreads = (size_t)gridDim.x * blockDim.x;
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t groups = (n + INV_CH - 1) / INV_CH;
  for (size_t g = t; g < groups; g += nthreads) {
    // element k of group g lives at g + k*groups: coalesced across the warp for every k
    Fr v[INV_CH], pre[INV_CH];
    Fr acc = Fr::one();
#pragma unroll
    for (int k = 0; k < INV_CH; k++) {
      size_t idx = g + (size_t)k * groups;
      v[k] = (idx < n) ? a[idx] : Fr::zero();
      pre[k] = acc;
      if (!v[k].is_zero()) acc = acc * v[k];
    }
    Fr inv = acc.inv();
#pragma unroll
    for (int k = INV_CH - 1; k >= 0; k--) {
      size_t idx = g + (size_t)k * groups;
      Fr r = Fr::zero();
      if (!v[k].is_zero()) {
        r = inv * pre[k];
        inv = inv * v[k];
      }
      if (idx < n) out[idx] = r;
    }
  }
}

__global__ void __launch_bounds__(256) k_vec_fill(Fr s, Fr *__restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = s;
}

// out[i][j] = in[i][j] * (omega_x^i - 1); omega_x^i read from the domain table tw[k] = omega_M^k, k <= M/2
// (second half of the circle: omega^(M/2 + k) = -omega^k).
__global__ void __launch_bounds__(256) k_mul_x_minus_one(const Fr *__restrict__ in, Fr *__restrict__ out, size_t x_size, size_t y_size,
                                                         const Fr *__restrict__ tw, uint32_t log_stride, size_t half_m) {
  size_t total = x_size * y_size;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
    size_t i = k / y_size;
    size_t e = i << log_stride;  // exponent of omega_M
    Fr w;
    if (e <= half_m) {
      w = tw[e];
    } else {
      Fr t = tw[e - half_m];
      w = t.neg();
    }
    Fr v = in[k];
    out[k] = v * (w - Fr::one());
  }
}

// 32x32 tiles through shared memory, 128-bit accesses on both sides (two uint4 planes, +1 padding).
__global__ void __launch_bounds__(256) k_transpose(const Fr *__restrict__ in, Fr *__restrict__ out, size_t rows, size_t cols) {
  __shared__ uint4 lo[32][33], hi[32][33];
  const size_t c0 = (size_t)blockIdx.x * 32, r0 = (size_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const uint4 *gin = reinterpret_cast<const uint4 *>(in);
  uint4 *gout = reinterpret_cast<uint4 *>(out);
  for (int k = ty; k < 32; k += 8) {
    size_t r = r0 + k, c = c0 + tx;
    if (r < rows && c < cols) {
      lo[k][tx] = gin[2 * (r * cols + c)];
      hi[k][tx] = gin[2 * (r * cols + c) + 1];
    }
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    size_t c = c0 + k, r = r0 + tx;  // output row = input column
    if (r < rows && c < cols) {
      gout[2 * (c * rows + r)] = lo[tx][k];
      gout[2 * (c * rows + r) + 1] = hi[tx][k];
    }
  }
}

// Exclusive suffix product out[i] = prod_{k > i} in[k] (out[n-1] = 1): the recursion polynomial's evaluations
// in prove1 (prove/src/lib.rs:1858-1867 computes it with a serial loop over 2^20 elements).
// Chunked three-phase scan: per-chunk products, recursive scan of the chunk products, per-chunk apply.
constexpr size_t SCAN_CHUNK = 64;
__global__ void __launch_bounds__(128) k_scan_chunk_products(const Fr *__restrict__ in, Fr *__restrict__ cp, size_t n) {
  size_t nchunks = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= nchunks) return;
  size_t lo = t * SCAN_CHUNK, hi = lo + SCAN_CHUNK;
  if (hi > n) hi = n;
  Fr acc = Fr::one();
  for (size_t i = lo; i < hi; i++) {
    Fr v = in[i];
    acc = acc * v;
  }
  cp[t] = acc;
}
__global__ void __launch_bounds__(128) k_scan_apply(const Fr *__restrict__ in, const Fr *__restrict__ cs, Fr *__restrict__ out, size_t n) {
  size_t nchunks = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= nchunks) return;
  size_t lo = t * SCAN_CHUNK, hi = lo + SCAN_CHUNK;
  if (hi > n) hi = n;
  Fr acc = cs[t];
  for (size_t i = hi; i-- > lo;) {
    Fr v = in[i];  // read before the write: in == out is allowed
    out[i] = acc;
    acc = acc * v;
  }
}
__global__ void k_scan_serial(const Fr *__restrict__ in, Fr *__restrict__ out, size_t n) {
  if (threadIdx.x || blockIdx.x) return;
  Fr acc = Fr::one();
  for (size_t i = n; i-- > 0;) {
    Fr v = in[i];
    out[i] = acc;
    acc = acc * v;
  }
}
int32_t vec_suffix_product(tkm_ctx *ctx, const Fr *in, Fr *out, size_t n) {
  if (n == 0) return TKM_OK;
  if (n <= 256) {
    k_scan_serial<<<1, 32, 0, ctx->stream>>>(in, out, n);
    return launch_check(ctx, "k_scan_serial");
  }
  size_t nchunks = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
  Scratch<Fr> cp;
  TKM_TRY(cp.alloc(ctx, nchunks));
  k_scan_chunk_products<<<(unsigned)((nchunks + 127) / 128), 128, 0, ctx->stream>>>(in, cp.p, n);
  TKM_TRY(launch_check(ctx, "k_scan_chunk_products"));
  TKM_TRY(vec_suffix_product(ctx, cp.p, cp.p, nchunks));
  k_scan_apply<<<(unsigned)((nchunks + 127) / 128), 128, 0, ctx->stream>>>(in, cp.p, out, n);
  return launch_check(ctx, "k_scan_apply");
}

// ---- reductions: VecOps::sum / VecOps::product (vector_operations/mod.rs:124,336; prove/src/lib.rs:1005-1016) and
// inner_product_two_vecs (vector_operations/mod.rs:100-141).  Grid-stride partials, warp shuffle tree, one partial per warp,
// then the same kernel reduces the partials.
template <int OP>  // 0 = sum(a), 1 = product(a), 2 = sum(a*b)
__global__ void __launch_bounds__(256) k_reduce(const Fr *__restrict__ a, const Fr *__restrict__ b, size_t n, Fr *__restrict__ partial) {
  Fr acc = (OP == 1) ? Fr::one() : Fr::zero();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fr v = a[i];
    if (OP == 2) { Fr w = b[i]; v = v * w; }
    acc = (OP == 1) ? acc * v : acc + v;
  }
  for (int d = 16; d > 0; d >>= 1) {
    Fr o;
#pragma unroll
    for (int k = 0; k < 8; k++) o.v[k] = __shfl_xor_sync(0xffffffffu, acc.v[k], d);
    acc = (OP == 1) ? acc * o : acc + o;
  }
  if ((threadIdx.x & 31) == 0) partial[(blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5] = acc;
}
int32_t vec_reduce(tkm_ctx *ctx, int op, const Fr *a, const Fr *b, size_t n, Fr *host_out) {
  if (n == 0) {
    *host_out = (op == 1) ? Fr::one() : Fr::zero();
    return TKM_OK;
  }
  unsigned blocks = grid_for(n, 256, ctx->sm_count, 2);
  size_t nwarps = (size_t)blocks * 8;
  Scratch<Fr> p1, p2;
  TKM_TRY(p1.alloc(ctx, nwarps));
  TKM_TRY(p2.alloc(ctx, 8));
  if (op == 0) k_reduce<0><<<blocks, 256, 0, ctx->stream>>>(a, nullptr, n, p1.p);
  else if (op == 1) k_reduce<1><<<blocks, 256, 0, ctx->stream>>>(a, nullptr, n, p1.p);
  else k_reduce<2><<<blocks, 256, 0, ctx->stream>>>(a, b, n, p1.p);
  TKM_TRY(launch_check(ctx, "k_reduce"));
  if (op == 1) k_reduce<1><<<1, 32, 0, ctx->stream>>>(p1.p, nullptr, nwarps, p2.p);
  else k_reduce<0><<<1, 32, 0, ctx->stream>>>(p1.p, nullptr, nwarps, p2.p);
  TKM_TRY(launch_check(ctx, "k_reduce"));
  TKM_CUDA(cudaMemcpyAsync(host_out, p2.p, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  return TKM_OK;
}
// outer_product_two_vecs (vector_operations/mod.rs:551-600): out[i*cols + j] = col[i] * row[j]
__global__ void __launch_bounds__(256) k_outer_product(const Fr *__restrict__ col, const Fr *__restrict__ row, Fr *__restrict__ out, size_t rows,
                                                       size_t cols) {
  size_t total = rows * cols;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
    Fr u = col[k / cols], v = row[k % cols];
    out[k] = u * v;
  }
}
int32_t vec_outer_product(tkm_ctx *ctx, const Fr *col, const Fr *row, Fr *out, size_t rows, size_t cols) {
  if (rows * cols == 0) return TKM_OK;
  k_outer_product<<<grid_for(rows * cols, 256, ctx->sm_count), 256, 0, ctx->stream>>>(col, row, out, rows, cols);
  return launch_check(ctx, "k_outer_product");
}

int32_t vec_fill(tkm_ctx *ctx, const Fr &s, Fr *out, size_t n) {
  if (n == 0) return TKM_OK;
  k_vec_fill<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(s, out, n);
  return launch_check(ctx, "k_vec_fill");
}
int32_t vec_mul_x_minus_one(tkm_ctx *ctx, const Fr *in, Fr *out, size_t x_size, size_t y_size) {
  if (!is_pow2(x_size)) return fail(TKM_ERR_INVALID_ARGUMENT, "x_size must be a power of two");
  if (ctx->domain_log2 < 0 || log2_exact(x_size) > (uint32_t)ctx->domain_log2)
    return fail(TKM_ERR_DOMAIN, "NTT domain is not initialized or smaller than x_size = %zu", x_size);
  size_t n = x_size * y_size;
  if (n == 0) return TKM_OK;
  if (x_size == 1) return vec_fill(ctx, Fr::zero(), out, n);  // X - 1 vanishes on the trivial domain {1}
  uint32_t log_stride = (uint32_t)ctx->domain_log2 - log2_exact(x_size);
  k_mul_x_minus_one<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(in, out, x_size, y_size, ctx->twiddles, log_stride,
                                                                             (size_t)1 << (ctx->domain_log2 - 1));
  return launch_check(ctx, "k_mul_x_minus_one");
}
int32_t vec_transpose(tkm_ctx *ctx, const Fr *in, Fr *out, size_t rows, size_t cols) {
  if (rows * cols == 0) return TKM_OK;
  if (in == out) return fail(TKM_ERR_INVALID_ARGUMENT, "transpose must be out of place");
  dim3 g((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  k_transpose<<<g, 256, 0, ctx->stream>>>(in, out, rows, cols);
  return launch_check(ctx, "k_transpose");
}

int32_t vec_to_mont(tkm_ctx *ctx, const Fr *in, Fr *out, size_t n) {
  if (n == 0) return TKM_OK;
  k_mont_convert<K_TO_MONT><<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(in, out, n);
  return launch_check(ctx, "k_mont_convert<to>");
}
int32_t vec_from_mont(tkm_ctx *ctx, const Fr *in, Fr *out, size_t n) {
  if (n == 0) return TKM_OK;
  k_mont_convert<K_FROM_MONT><<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(in, out, n);
  return launch_check(ctx, "k_mont_convert<from>");
}
int32_t vec_op(tkm_ctx *ctx, int op, const Fr *a, const Fr *b, Fr *out, size_t n) {
  if (n == 0) return TKM_OK;
  unsigned g = grid_for(n, 256, ctx->sm_count);
  switch (op) {
    case TKM_OP_ADD: k_vec_op<TKM_OP_ADD><<<g, 256, 0, ctx->stream>>>(a, b, out, n); break;
    case TKM_OP_SUB: k_vec_op<TKM_OP_SUB><<<g, 256, 0, ctx->stream>>>(a, b, out, n); break;
    case TKM_OP_MUL: k_vec_op<TKM_OP_MUL><<<g, 256, 0, ctx->stream>>>(a, b, out, n); break;
    case TKM_OP_DIV: {
      // a / b = a * inv(b) with the batched inverse (one Fermat inverse per 8 elements) instead of one per element
      Scratch<Fr> binv;
      TKM_TRY(binv.alloc(ctx, n));
      TKM_TRY(vec_inv(ctx, b, binv.p, n));
      k_vec_op<TKM_OP_MUL><<<g, 256, 0, ctx->stream>>>(a, binv.p, out, n);
      break;
    }
    default: return fail(TKM_ERR_INVALID_ARGUMENT, "unknown vector op %d", op);
  }
  return launch_check(ctx, "k_vec_op");
}
int32_t vec_scale(tkm_ctx *ctx, const Fr &s, const Fr *a, Fr *out, size_t n) {
  if (n == 0) return TKM_OK;
  k_vec_scale<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(s, a, out, n);
  return launch_check(ctx, "k_vec_scale");
}
int32_t vec_inv(tkm_ctx *ctx, const Fr *a, Fr *out, size_t n) {
  if (n == 0) return TKM_OK;
  size_t groups = (n + INV_CH - 1) / INV_CH;
  k_vec_inv<<<grid_for(groups, 128, ctx->sm_count), 128, 0, ctx->stream>>>(a, out, n);
  return launch_check(ctx, "k_vec_inv");
}

}  // namespace tkm
