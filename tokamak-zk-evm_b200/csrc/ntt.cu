// Bivariate NTT over BLS12-381 Fr for sm_100a.
//
// Replaces ntt::initialize_domain / ntt::ntt as the reference calls them from
// DensePolynomialExt::_biNTT (libs/src/bivariate_polynomial/mod.rs:33-55,1422-1478):
// natural order in and out, inverse includes 1/N, per-axis coset generators with the semantics
// pinned by libs/src/tests.rs:134-180.
//
// One generic kernel transforms a tile of a strided axis entirely in shared memory:
//   array [outer][n][inner], transform along n.  A length-n axis is split n = n1*n2 into at most two
//   radix-2 DIF passes:  pass 1 runs the first log2(n1) stages on the n1 elements {l*n2 + g} (twiddles
//   omega_n^(2^u (j*n2 + g)) are exact table look-ups, so the split costs no extra multiplications);
//   pass 2 finishes the n1 independent length-n2 blocks and stores in natural order by applying the
//   bit-reversal in the store addresses.  A tile is L x C elements (L = sub-transform length, C = batch of
//   adjacent columns or rows, L*C <= 2048) staged in two 128-bit planes so every butterfly access is
//   a conflict-free LDS.128/STS.128; the per-stage twiddles of the tile are staged in shared memory
//   once and shared by the C batch lanes.  Global traffic is 128-bit, coalesced along whichever of
//   (axis, batch) is contiguous.
//
// Twiddle domain: tw[k] = omega_M^k, k = 0..M/2 (tw[M/2] = -1).  Inverse twiddles are read as
// omega^-e = -tw[M/2 - e] and the butterfly computes (b - a) * tw[M/2 - e], so one table serves both
// directions.
#include <cstdlib>

#include "common.cuh"

namespace tkm {

// 5^((r-1)/2^32), canonical limbs (the root ICICLE's domain is built from).
static const uint32_t FR_ROU32[8] = {0x0b912f1fu, 0x1b788f50u, 0x70b3e094u, 0xc4024ff2u,
                                     0xd168d6c0u, 0x0fd56dc8u, 0x5b416b6fu, 0x0212d79eu};

Fr root_of_unity_host(uint32_t log_n) {
  Fr w;
  memcpy(w.v, FR_ROU32, 32);
  w = w.to_mont();
  for (uint32_t i = log_n; i < 32; i++) w = w.sqr();
  return w;
}

struct PowTable {
  Fr p[32];  // p[b] = base^(2^b)
};

// out[k] = scale * base^k for k < count (binary exponentiation against a 2^b power table).
__global__ void __launch_bounds__(256) k_powers(Fr *__restrict__ out, PowTable tab, Fr scale, size_t count) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < count; k += (size_t)gridDim.x * blockDim.x) {
    Fr acc = scale;
    size_t e = k;
    for (int b = 0; e; b++, e >>= 1)
      if (e & 1) acc = acc * tab.p[b];
    out[k] = acc;
  }
}

static int32_t fill_powers(tkm_ctx *ctx, Fr *out, const Fr &base, const Fr &scale, size_t count) {
  PowTable tab;
  tab.p[0] = base;
  for (int b = 1; b < 32; b++) tab.p[b] = tab.p[b - 1].sqr();
  k_powers<<<grid_for(count, 256, ctx->sm_count), 256, 0, ctx->stream>>>(out, tab, scale, count);
  return launch_check(ctx, "k_powers");
}

int32_t fill_powers_public(tkm_ctx *ctx, Fr *out, const Fr &base, const Fr &scale, size_t count) {
  return fill_powers(ctx, out, base, scale, count);
}

constexpr int NTT_MAX_PEERS = 16;
struct NttPass {
  const Fr *in;
  Fr *out;
  const Fr *tw;
  const Fr *pre;   // multiplier by axis position applied after load (or null)
  const Fr *post;  // multiplier by output axis position applied before store (or null)
  Fr post_scalar;  // extra scalar applied before store when has_post_scalar
  uint64_t inner, outer;
  uint32_t logM, logn, logn1, logn2, logL, logC;
  uint32_t pass;         // 1 = first pass of a split, 2 = last (or only) pass
  uint32_t batch_inner;  // 1: batch lanes walk `inner`; 0: batch lanes walk `outer` (inner == 1)
  uint32_t has_post_scalar;
  // Fused exchange (multi-GPU row<->column re-sharding): when sc_on, the last pass does not write `out`;
  // the element at axis position a of batch lane b goes to sc_peer[a >> sc_logblk] + (a & (blk-1))*sc_sA + (b + sc_b0)*sc_sB,
  // where sc_peer[] are peer-mapped device buffers (NVLink P2P stores): the all-to-all transpose happens in the store.
  uint32_t sc_on, sc_logblk;
  uint64_t sc_sA, sc_sB, sc_b0;
  Fr *sc_peer[NTT_MAX_PEERS];
};

constexpr uint32_t NTT_TILE_LOG = 11;  // 2048 elements = 64 KiB of tile data
constexpr uint32_t NTT_MAX_LOGL = 10;
constexpr uint32_t NTT_THREADS = 128;
#ifndef TKM_NTT_MIN_CTAS
#define TKM_NTT_MIN_CTAS 6
#endif
constexpr int NTT_MIN_CTAS = TKM_NTT_MIN_CTAS;

// Shared-memory index of tile element i.  Tiles whose batch lanes walk `outer` (C = 2 columns per 512-point row) are
// accessed with power-of-two strides in the last butterfly stages and in the bit-reversed store; XOR-ing the folded
// higher index bits into bits 1-2 spreads every such 8-lane group over the 8 bank groups of 16 bytes (a bijection within
// each aligned block of 8 elements).  Tiles with C >= 8 keep the identity: their fastest index is the batch lane.
template <bool SWZ>
__device__ __forceinline__ uint32_t sidx(uint32_t i) {
  if (!SWZ) return i;
  return i ^ ((((i >> 3) ^ (i >> 5) ^ (i >> 7) ^ (i >> 9)) & 3u) << 1);
}
__device__ __forceinline__ Fr lds_fr(const uint4 *lo, const uint4 *hi, uint32_t i) {
  Fr r;
  uint4 a = lo[i], b = hi[i];
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void sts_fr(uint4 *lo, uint4 *hi, uint32_t i, const Fr &r) {
  lo[i] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  hi[i] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ Fr ldg_fr(const Fr *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

template <bool INVERSE, bool SWZ>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_CTAS) k_ntt_pass(NttPass p) {
  extern __shared__ uint4 smem[];
  const uint32_t L = 1u << p.logL, C = 1u << p.logC, TILE = L * C;
  // the two 128-bit planes sit one 16-byte slot apart modulo the 128-byte bank row (SWZ: lanes alternate between them)
  uint4 *d_lo = smem, *d_hi = smem + TILE + (SWZ ? 1 : 0), *t_lo = smem + 2 * TILE + (SWZ ? 8 : 0), *t_hi = t_lo + L;
#define DI(i) sidx<SWZ>(i)
  const uint32_t tid = threadIdx.x;
  const uint64_t n = 1ull << p.logn;
  const uint32_t n1 = 1u << p.logn1, n2 = 1u << p.logn2;

  // tile -> (batch block cb, group g, outer index o)
  const uint64_t batch_total = p.batch_inner ? p.inner : p.outer;
  const uint64_t nb = batch_total >> p.logC;
  const uint32_t G = (p.pass == 1) ? n2 : n1;
  uint64_t bid = blockIdx.x;
  const uint64_t cb = bid % nb;
  bid /= nb;
  const uint32_t g = (uint32_t)(bid % G);
  const uint64_t o = bid / G;

  // element (l, c) of the tile sits at base + pos(l)*pos_stride + c*c_stride
  const uint64_t pos_stride = p.inner;
  const uint64_t c_stride = p.batch_inner ? 1 : n;  // batch over outer requires inner == 1
  const uint64_t base = p.batch_inner ? (o * n * p.inner + cb * C) : (cb * C * n);
  const uint32_t pos_mul = (p.pass == 1) ? n2 : 1;
  const uint32_t pos_add = (p.pass == 1) ? g : g * n2;

  // ---- stage the tile's twiddles: stage u holds L >> (u+1) entries at offset L - (L >> u)
  {
    const uint32_t logA = p.logM - p.logL;
    const uint64_t B = (p.pass == 1) ? ((uint64_t)g << (p.logM - p.logn)) : 0;
    const uint64_t half_m = 1ull << (p.logM - 1);
    for (uint32_t e = tid; e + 1 < L; e += NTT_THREADS) {
      uint32_t r = L - e;                               // 2..L
      uint32_t u = p.logL - (32 - __clz(r - 1));        // largest u with r <= L >> u
      uint32_t j = e - (L - (L >> u));
      uint64_t idx = (((uint64_t)j << logA) + B) << u;  // exponent of omega_M, < M/2
      if (INVERSE) idx = half_m - idx;
      Fr w = ldg_fr(p.tw + idx);
      sts_fr(t_lo, t_hi, e, w);
    }
  }
  // ---- load the tile (128-bit, two lanes per element)
  {
    const uint4 *gin = reinterpret_cast<const uint4 *>(p.in);
    for (uint32_t idx = tid; idx < 2 * TILE; idx += NTT_THREADS) {
      uint32_t half = idx & 1, e = idx >> 1, l, c;
      if (p.batch_inner) { c = e & (C - 1); l = e >> p.logC; }
      else { l = e & (L - 1); c = e >> p.logL; }
      uint64_t addr = base + (uint64_t)(l * pos_mul + pos_add) * pos_stride + (uint64_t)c * c_stride;
      uint4 v = gin[2 * addr + half];
      (half ? d_hi : d_lo)[DI(l * C + c)] = v;
    }
  }
  __syncthreads();
  if (p.pre) {
    for (uint32_t e = tid; e < TILE; e += NTT_THREADS) {
      uint32_t l = e >> p.logC;
      Fr x = lds_fr(d_lo, d_hi, DI(e));
      Fr s = ldg_fr(p.pre + (l * pos_mul + pos_add));
      sts_fr(d_lo, d_hi, DI(e), x * s);
    }
    __syncthreads();
  }
  // ---- radix-2 DIF stages in shared memory
  // In the last (or only) pass the final two stages have twiddles {1, omega_4} and {1}: they are fused into
  // one radix-4 register butterfly with a single multiplication per four elements.
  const bool radix4_tail = (p.pass == 2 && p.logL >= 2);
  const uint32_t r2_stages = radix4_tail ? p.logL - 2 : p.logL;
  // Stages are taken two at a time as radix-4 register butterflies (4 elements, 4 products, 3 twiddles): half the
  // shared-memory round trips and half the barriers of a radix-2 sweep; an odd stage count starts with one radix-2 stage.
  uint32_t u = 0;
  if (r2_stages & 1) {
    const uint32_t logh = p.logL - 1, h = 1u << logh;
    for (uint32_t w = tid; w < (TILE >> 1); w += NTT_THREADS) {
      uint32_t c = w & (C - 1), bidx = w >> p.logC;
      uint32_t j = bidx & (h - 1);
      uint32_t l0 = ((bidx >> logh) << (logh + 1)) + j;
      uint32_t i0 = l0 * C + c, i1 = i0 + h * C;
      Fr a = lds_fr(d_lo, d_hi, DI(i0));
      Fr b = lds_fr(d_lo, d_hi, DI(i1));
      Fr tw = lds_fr(t_lo, t_hi, j);  // stage 0 twiddles start at offset 0
      Fr sm = a + b;
      Fr df = INVERSE ? (b - a) : (a - b);
      sts_fr(d_lo, d_hi, DI(i0), sm);
      sts_fr(d_lo, d_hi, DI(i1), df * tw);
    }
    __syncthreads();
    u = 1;
  }
  for (; u + 1 < r2_stages; u += 2) {
    const uint32_t logh = p.logL - 1 - u, h = 1u << logh, logh2 = logh - 1, h2 = h >> 1;
    const uint32_t toff0 = L - (L >> u), toff1 = L - (L >> (u + 1));
    for (uint32_t w = tid; w < (TILE >> 2); w += NTT_THREADS) {
      const uint32_t c = w & (C - 1), q = w >> p.logC;
      const uint32_t j = q & (h2 - 1);
      const uint32_t l0 = ((q >> logh2) << (logh + 1)) + j;
      const uint32_t i0 = l0 * C + c, i1 = i0 + h2 * C, i2 = i0 + h * C, i3 = i2 + h2 * C;
      Fr x0 = lds_fr(d_lo, d_hi, DI(i0)), x2 = lds_fr(d_lo, d_hi, DI(i2));
      Fr a0 = x0 + x2;
      Fr a2 = (INVERSE ? (x2 - x0) : (x0 - x2)) * lds_fr(t_lo, t_hi, toff0 + j);
      Fr x1 = lds_fr(d_lo, d_hi, DI(i1)), x3 = lds_fr(d_lo, d_hi, DI(i3));
      Fr a1 = x1 + x3;
      Fr a3 = (INVERSE ? (x3 - x1) : (x1 - x3)) * lds_fr(t_lo, t_hi, toff0 + j + h2);
      const Fr tw1 = lds_fr(t_lo, t_hi, toff1 + j);
      sts_fr(d_lo, d_hi, DI(i0), a0 + a1);
      sts_fr(d_lo, d_hi, DI(i1), (INVERSE ? (a1 - a0) : (a0 - a1)) * tw1);
      sts_fr(d_lo, d_hi, DI(i2), a2 + a3);
      sts_fr(d_lo, d_hi, DI(i3), (INVERSE ? (a3 - a2) : (a2 - a3)) * tw1);
    }
    __syncthreads();
  }
  if (radix4_tail) {
    const Fr w4 = lds_fr(t_lo, t_hi, (L - 4) + 1);  // stage logL-2 (offset L - 4), j = 1: omega_4 (or its stand-in for the inverse)
    for (uint32_t w = tid; w < (TILE >> 2); w += NTT_THREADS) {
      uint32_t c = w & (C - 1), q = w >> p.logC;
      uint32_t i0 = (q << 2) * C + c, i1 = i0 + C, i2 = i1 + C, i3 = i2 + C;
      Fr x0 = lds_fr(d_lo, d_hi, DI(i0)), x1 = lds_fr(d_lo, d_hi, DI(i1)), x2 = lds_fr(d_lo, d_hi, DI(i2)), x3 = lds_fr(d_lo, d_hi, DI(i3));
      Fr y0 = x0 + x2, y1 = x1 + x3;
      Fr y2 = x0 - x2;
      Fr y3 = (INVERSE ? (x3 - x1) : (x1 - x3)) * w4;
      sts_fr(d_lo, d_hi, DI(i0), y0 + y1);
      sts_fr(d_lo, d_hi, DI(i1), y0 - y1);
      sts_fr(d_lo, d_hi, DI(i2), y2 + y3);
      sts_fr(d_lo, d_hi, DI(i3), y2 - y3);
    }
    __syncthreads();
  }
  // ---- optional output scaling (coset^-i and/or 1/N), by output axis position
  const uint32_t brev_g = (p.logn1 == 0) ? 0 : (__brev(g) >> (32 - p.logn1));
  if (p.pass == 2 && (p.post || p.has_post_scalar)) {
    for (uint32_t e = tid; e < TILE; e += NTT_THREADS) {
      uint32_t l = e >> p.logC;
      uint32_t k2 = (p.logL == 0) ? 0 : (__brev(l) >> (32 - p.logL));
      Fr x = lds_fr(d_lo, d_hi, DI(e));
      if (p.post) x = x * ldg_fr(p.post + ((uint64_t)k2 * n1 + brev_g));
      if (p.has_post_scalar) x = x * p.post_scalar;
      sts_fr(d_lo, d_hi, DI(e), x);
    }
    __syncthreads();
  }
  // ---- store
  {
    uint4 *gout = reinterpret_cast<uint4 *>(p.out);
    for (uint32_t idx = tid; idx < 2 * TILE; idx += NTT_THREADS) {
      uint32_t half = idx & 1, e = idx >> 1, k, c;
      if (p.batch_inner) { c = e & (C - 1); k = e >> p.logC; }
      else { k = e & (L - 1); c = e >> p.logL; }
      uint64_t pos;
      uint32_t l;
      if (p.pass == 1) {
        l = k;
        pos = (uint64_t)k * n2 + g;
      } else {
        l = (p.logL == 0) ? 0 : (__brev(k) >> (32 - p.logL));
        pos = (uint64_t)k * n1 + brev_g;
      }
      const uint4 v = (half ? d_hi : d_lo)[DI(l * C + c)];
      if (p.sc_on) {
        const uint64_t a_loc = pos & ((1ull << p.sc_logblk) - 1);
        uint4 *gp = reinterpret_cast<uint4 *>(p.sc_peer[pos >> p.sc_logblk]);
        const uint64_t dst = a_loc * p.sc_sA + (cb * C + c + p.sc_b0) * p.sc_sB;
        gp[2 * dst + half] = v;
      } else {
        uint64_t addr = base + pos * pos_stride + (uint64_t)c * c_stride;
        gout[2 * addr + half] = v;
      }
    }
  }
}

#undef DI

static int32_t launch_pass(tkm_ctx *ctx, const NttPass &p, bool inverse) {
  const uint32_t L = 1u << p.logL, C = 1u << p.logC;
  // Swizzled tiles remove the shared-memory bank conflicts of the row-batched (C = 2) pass -- measured on B200: the
  // 16384 x 512 transform takes 1.875 ms with them against 1.858 ms without (the index arithmetic costs more ALU issue slots
  // than the conflicts cost LSU cycles; the kernel is bound by the integer pipes).  Opt-in for profiling.
  static const bool swz_on = getenv("TKM_NTT_SWIZZLE") != nullptr;
  const bool swz = swz_on && !p.batch_inner && C < 8 && L >= 16;
  size_t smem = (size_t)(2 * L * C + 2 * L + (swz ? 16 : 0)) * sizeof(uint4);
  bool *attr_set = ctx->ntt_attr_set;  // per context = per device (cudaFuncSetAttribute applies to the current device only)
  const int variant = (inverse ? 1 : 0) + (swz ? 2 : 0);
  if (!attr_set[variant]) {
    size_t max_smem = (size_t)(2 * (1u << NTT_TILE_LOG) + 2 * (1u << NTT_MAX_LOGL) + 16) * sizeof(uint4);
    const void *fn = variant == 0 ? (const void *)k_ntt_pass<false, false> : variant == 1 ? (const void *)k_ntt_pass<true, false>
                     : variant == 2 ? (const void *)k_ntt_pass<false, true> : (const void *)k_ntt_pass<true, true>;
    TKM_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    attr_set[variant] = true;
  }
  const uint64_t batch_total = p.batch_inner ? p.inner : p.outer;
  const uint64_t G = (p.pass == 1) ? (1ull << p.logn2) : (1ull << p.logn1);
  const uint64_t tiles = (batch_total >> p.logC) * G * (p.batch_inner ? p.outer : 1);
  if (tiles == 0 || tiles > 0x7fffffffull) return fail(TKM_ERR_INVALID_ARGUMENT, "NTT tile count out of range");
  if (variant == 0) k_ntt_pass<false, false><<<(unsigned)tiles, NTT_THREADS, smem, ctx->stream>>>(p);
  else if (variant == 1) k_ntt_pass<true, false><<<(unsigned)tiles, NTT_THREADS, smem, ctx->stream>>>(p);
  else if (variant == 2) k_ntt_pass<false, true><<<(unsigned)tiles, NTT_THREADS, smem, ctx->stream>>>(p);
  else k_ntt_pass<true, true><<<(unsigned)tiles, NTT_THREADS, smem, ctx->stream>>>(p);
  return launch_check(ctx, "k_ntt_pass");
}

// Transform along the middle axis of [outer][n][inner].  coset: Montgomery scalar or null.
// extra_scalar (inverse only): folded into the output scaling of this axis (used for the 1/N of the
// other axis so a 2-D inverse multiplies each element by 1/(x*y) once).
struct NttScatter {
  void *const *peers;  // n_peers peer-mapped device buffers
  uint32_t n_peers;
  uint64_t stride_a, stride_b, b0;
};
static int32_t ntt_axis_impl(tkm_ctx *ctx, const Fr *in, Fr *out, size_t outer, size_t n, size_t inner, int dir,
                             const Fr *coset, const Fr *extra_scalar, bool defer_scale = false, const NttScatter *scatter = nullptr) {
  if (!is_pow2(n) || !is_pow2(outer) || !is_pow2(inner)) return fail(TKM_ERR_INVALID_ARGUMENT, "NTT sizes must be powers of two");
  const bool inverse = dir == TKM_INVERSE;
  const size_t total = outer * n * inner;
  const uint32_t logn = log2_exact(n);
  if (ctx->domain_log2 < 0) return fail(TKM_ERR_DOMAIN, "NTT domain is not initialized. Call tkm_ntt_domain_init first.");
  if ((int32_t)logn > ctx->domain_log2)
    return fail(TKM_ERR_DOMAIN, "NTT domain size too small: initialized 2^%d but axis 2^%u", ctx->domain_log2, logn);
  if (logn > 2 * NTT_MAX_LOGL) return fail(TKM_ERR_INVALID_ARGUMENT, "axis length 2^%u exceeds the two-pass limit 2^%u", logn, 2 * NTT_MAX_LOGL);

  bool coset_on = coset && !(*coset == Fr::one());
  if (n == 1) {
    // length-1 transform is the identity (1/1 scaling); only a pending extra scalar applies
    if (extra_scalar) return vec_scale(ctx, *extra_scalar, in, out, total);
    if (in != out) TKM_CUDA(cudaMemcpyAsync(out, in, total * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
    return TKM_OK;
  }

  // scaling tables
  Scratch<Fr> table;
  const Fr *pre = nullptr, *post = nullptr;
  Fr post_scalar = Fr::one();
  bool has_post_scalar = false;
  if (inverse) {
    Fr sc = ctx->inv_pow2[logn];  // 1/n
    if (extra_scalar) sc = sc * *extra_scalar;
    if (defer_scale && !coset_on) {
      // the caller folds this axis' 1/n into the other axis' output scaling
    } else if (coset_on) {
      TKM_TRY(table.alloc(ctx, n));
      TKM_TRY(fill_powers(ctx, table.p, coset->inv(), sc, n));
      post = table.p;
    } else {
      post_scalar = sc;
      has_post_scalar = true;
    }
  } else if (coset_on) {
    TKM_TRY(table.alloc(ctx, n));
    TKM_TRY(fill_powers(ctx, table.p, *coset, Fr::one(), n));
    pre = table.p;
  }

  NttPass p;
  memset(&p, 0, sizeof p);
  p.tw = ctx->twiddles;
  p.inner = inner;
  p.outer = outer;
  p.logM = (uint32_t)ctx->domain_log2;
  p.logn = logn;
  p.batch_inner = (inner > 1 || outer == 1) ? 1 : 0;
  if (!p.batch_inner && inner != 1) return fail(TKM_ERR_INTERNAL, "batch-over-outer needs inner == 1");
  const size_t batch_total = p.batch_inner ? inner : outer;
  // 512-element tiles (16 KiB + twiddles) on CTAs of 128 threads, six resident per SM (80 registers): the finer grain lets the
  // load/store and barrier phases of one tile hide under the butterflies of five others (16384 x 512 forward: 1.598 ms with
  // 1024-element tiles on 256-thread CTAs, 1.552 ms with 1024 / 128 threads, 1.538 ms with 512 / 128; 64-thread CTAs: 1.59 ms)
  uint32_t tile_log = 9;
  uint32_t tile_log_long = 9;  // tiles of sub-transforms with L >= 512 (one row per tile at 9: the twiddles are not shared)
  if (const char *e = getenv("TKM_NTT_TILE_LOG")) {  // developer knob
    uint32_t v = (uint32_t)atoi(e);
    if (v >= 8 && v <= NTT_TILE_LOG) tile_log = tile_log_long = v;
  }
  if (const char *e = getenv("TKM_NTT_TILE_LOG_LONG")) {  // developer knob
    uint32_t v = (uint32_t)atoi(e);
    if (v >= 8 && v <= NTT_TILE_LOG) tile_log_long = v;
  }
  auto pick_logC = [&](uint32_t logL) {
    const uint32_t tl = logL >= 9 ? tile_log_long : tile_log;
    uint32_t lc = tl > logL ? tl - logL : 0;
    uint32_t lb = log2_exact(batch_total);
    return lc < lb ? lc : lb;
  };

  auto arm_scatter = [&]() {
    if (!scatter) return;
    p.sc_on = 1;
    p.sc_logblk = logn - log2_exact(scatter->n_peers);
    p.sc_sA = scatter->stride_a;
    p.sc_sB = scatter->stride_b;
    p.sc_b0 = scatter->b0;
    for (uint32_t i = 0; i < scatter->n_peers; i++) p.sc_peer[i] = (Fr *)scatter->peers[i];
  };
  if (scatter) {
    if (scatter->n_peers == 0 || scatter->n_peers > NTT_MAX_PEERS || !is_pow2(scatter->n_peers) || scatter->n_peers > n)
      return fail(TKM_ERR_INVALID_ARGUMENT, "fused exchange needs a power-of-two peer count <= min(%d, axis length)", NTT_MAX_PEERS);
    if (p.batch_inner && outer != 1) return fail(TKM_ERR_INVALID_ARGUMENT, "fused exchange over a strided axis needs outer == 1");
  }
  if (logn <= NTT_MAX_LOGL) {
    p.in = in;
    p.out = out;
    arm_scatter();
    p.logn1 = 0;
    p.logn2 = logn;
    p.logL = logn;
    p.logC = pick_logC(logn);
    p.pass = 2;
    p.pre = pre;
    p.post = post;
    p.post_scalar = post_scalar;
    p.has_post_scalar = has_post_scalar;
    return launch_pass(ctx, p, inverse);
  }
  // two passes through a scratch buffer (pass 2 scatters across tiles, so it cannot run in place)
  Scratch<Fr> mid;
  TKM_TRY(mid.alloc(ctx, total));
  p.logn1 = (logn + 1) / 2;
  p.logn2 = logn - p.logn1;
  p.in = in;
  p.out = mid.p;
  p.logL = p.logn1;
  p.logC = pick_logC(p.logL);
  p.pass = 1;
  p.pre = pre;
  p.post = nullptr;
  p.has_post_scalar = 0;
  TKM_TRY(launch_pass(ctx, p, inverse));
  p.in = mid.p;
  p.out = out;
  arm_scatter();
  p.logL = p.logn2;
  p.logC = pick_logC(p.logL);
  p.pass = 2;
  p.pre = nullptr;
  p.post = post;
  p.post_scalar = post_scalar;
  p.has_post_scalar = has_post_scalar;
  return launch_pass(ctx, p, inverse);
}

int32_t ntt_axis(tkm_ctx *ctx, const Fr *in, Fr *out, size_t outer, size_t n, size_t inner, int dir, const Fr *coset) {
  return ntt_axis_impl(ctx, in, out, outer, n, inner, dir, coset, nullptr);
}

// Batched 1-D transform whose output is re-sharded across peers in the store of its last pass (no separate transpose,
// no separate all-to-all).  The 1/n of an inverse transform is applied here.
int32_t ntt_axis_scatter(tkm_ctx *ctx, const Fr *in, size_t outer, size_t n, size_t inner, int dir, const Fr *coset, void *const *peers,
                         uint32_t n_peers, uint64_t stride_a, uint64_t stride_b, uint64_t b0) {
  NttScatter sc{peers, n_peers, stride_a, stride_b, b0};
  return ntt_axis_impl(ctx, in, nullptr, outer, n, inner, dir, coset, nullptr, false, &sc);
}

// DensePolynomialExt::_biNTT (bivariate_polynomial/mod.rs:1422-1478).
int32_t bintt_dev(tkm_ctx *ctx, const Fr *in, Fr *out, size_t x, size_t y, int dir, const Fr *coset_x, const Fr *coset_y) {
  if (!is_pow2(x) || !is_pow2(y)) return fail(TKM_ERR_INVALID_ARGUMENT, "biNTT sizes must be powers of two (got %zu x %zu)", x, y);
  if (ctx->domain_log2 < 0) return fail(TKM_ERR_DOMAIN, "NTT domain is not initialized. Call tkm_ntt_domain_init first.");
  if (log2_exact(x) + log2_exact(y) > (uint32_t)ctx->domain_log2)
    return fail(TKM_ERR_DOMAIN, "NTT domain size too small: initialized size 2^%d but input size %zu", ctx->domain_log2, x * y);
  if (x == 1) return ntt_axis_impl(ctx, in, out, 1, y, 1, dir, coset_y, nullptr);
  if (y == 1) return ntt_axis_impl(ctx, in, out, 1, x, 1, dir, coset_x, nullptr);
  // Y pass over contiguous rows, then X pass over strided columns (same order as the reference).
  // Inverse: unless the Y axis needs a coset table anyway, its 1/y rides on the X axis' output scaling
  // so every element is multiplied by 1/(x*y) exactly once.
  const bool inverse = dir == TKM_INVERSE;
  const bool coset_y_on = coset_y && !(*coset_y == Fr::one());
  const bool defer = inverse && !coset_y_on;
  TKM_CUDA(cudaEventRecord(ctx->kev0, ctx->stream));
  TKM_TRY(ntt_axis_impl(ctx, in, out, x, y, 1, dir, coset_y, nullptr, defer));
  const Fr *extra = defer ? &ctx->inv_pow2[log2_exact(y)] : nullptr;
  TKM_TRY(ntt_axis_impl(ctx, out, out, 1, x, y, dir, coset_x, extra));
  TKM_CUDA(cudaEventRecord(ctx->kev1, ctx->stream));
  ctx->kernel_timed = true;
  return TKM_OK;
}

// ---- domain ---------------------------------------------------------------------------------
int32_t domain_init(tkm_ctx *ctx, uint32_t log2_size) {
  if (log2_size > 32) return fail(TKM_ERR_INVALID_ARGUMENT, "NTT domain 2^%u exceeds the 2-adicity of Fr (32)", log2_size);
  if (log2_size > 28) return fail(TKM_ERR_ALLOCATION, "NTT domain 2^%u would need a %llu MiB twiddle table", log2_size,
                                  (unsigned long long)((1ull << (log2_size - 1)) * 32 >> 20));
  if (ctx->domain_log2 >= (int32_t)log2_size) return TKM_OK;  // bivariate_polynomial/mod.rs:43-46
  if (ctx->twiddles) {
    TKM_CUDA(cudaStreamSynchronize(ctx->stream));
    TKM_CUDA(cudaFree(ctx->twiddles));
    ctx->twiddles = nullptr;
    ctx->domain_log2 = -1;
  }
  size_t count = (log2_size == 0) ? 1 : ((size_t)1 << (log2_size - 1)) + 1;
  TKM_CUDA(cudaMalloc((void **)&ctx->twiddles, count * sizeof(Fr)));
  Fr w = root_of_unity_host(log2_size);
  int32_t st = fill_powers(ctx, ctx->twiddles, w, Fr::one(), count);
  if (st != TKM_OK) return st;
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->domain_log2 = (int32_t)log2_size;
  return TKM_OK;
}

int32_t domain_release(tkm_ctx *ctx) {
  if (ctx->twiddles) {
    TKM_CUDA(cudaStreamSynchronize(ctx->stream));
    TKM_CUDA(cudaFree(ctx->twiddles));
  }
  ctx->twiddles = nullptr;
  ctx->domain_log2 = -1;
  return TKM_OK;
}

}  // namespace tkm
