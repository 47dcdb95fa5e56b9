// BLS12-381 G1 multi-scalar multiplication for sm_100a: signed-digit Pippenger.
//
// Replaces msm::msm(scalars, bases, MSMConfig::default(), out[1]) as the reference calls it for every
// commitment (libs/src/iotools/mod.rs:2093-2099; libs/src/group_structures/mod.rs:108-114,135-141).
//
// Pipeline (all on the context stream, no host synchronisation until the 96-byte result is read):
//   1. k_decompose     one thread per scalar: (optional from-Montgomery,) GLV split k = k1 + k2*lambda into
//                      two signed 127-bit halves (glv.cuh), signed c-bit digits of each half;
//                      emits (bucket key, base index | sign) pairs, window-major, zero digits keyed
//                      to a trash bucket that sorts last.
//   2. cub radix sort  of the pairs by bucket key (only the significant key bits).
//   3. k_accumulate    one thread per fixed-length chunk of the sorted list, so work is balanced for
//                      ANY scalar distribution: XYZZ mixed additions of gathered affine bases; runs
//                      that lie strictly inside a chunk are final and go straight to their bucket,
//                      the first/last run of every chunk go to a (key, point) partial list.
//   4. k_segreduce     warp-cooperative segmented reduction of the partial list (shuffle tree of
//                      full XYZZ additions, 32 entries per warp), repeated until one warp remains.
//   5. k_bucket_seg /  parallel window reduction: running sums over 16-bucket segments, then per
//      k_bucket_bits / window a masked tree-sum per index bit (sum_d d*B_d = sum_k 2^k sum_{d: bit k} B_d), each
//      k_window_sums   part weighted by its 2^k where it is produced, window sums as shuffle tree-sums,
//   6. k_final         Horner over the windows of each half as two concurrent chains, sum = chain1 + phi(chain2),
//                      one inversion (binary extended Euclid) to affine.
// Integer-pipe bound: N*W mixed additions of ~10 Fq products each; no tensor cores.
#include <cub/device/device_radix_sort.cuh>

#include <cstdlib>

#include "common.cuh"
#include "glv.cuh"

namespace tkm {

struct MsmGeom {
  uint32_t c;        // window bits
  uint32_t Wd;       // digit windows of the scalar decomposition (GLV: 2*Wh, half h owns windows [h*Wh, h*Wh + Wh))
  uint32_t glv;      // 1: scalars are split k = k1 + k2*lambda (glv.cuh); 0: plain 256-bit digits (fixed-base tables)
  uint32_t Wh;       // windows per chain of the Horner tail (= Wd/2 with GLV, Wd without)
  uint32_t val_stride;  // 0, or (fixed-base tables) offset of window w's table: base index += w*val_stride
  uint32_t W;        // bucket windows (= Wd; 1 when precomputed tables fold every window into one bucket set)
  uint32_t B;        // buckets per window = 2^(c-1)
  uint32_t logB;
  uint32_t nbuckets; // W*B (+1 trash bucket at index nbuckets)
  uint32_t g;        // bucket segment length for the window reduction
  uint32_t logg;
  uint32_t nseg;     // segments per window = B/g
  uint32_t nbits;    // log2(nseg)
};

static MsmGeom pick_geom(size_t n, uint32_t fixed_c = 0, uint32_t table_stride = 0) {
  // minimise W*(n + 3*2^(c-1)) -- mixed adds plus ~3 madd-equivalents per bucket of reduction
  double best = 1e300;
  uint32_t bc = 8;
  // GLV halves cover 128 bits each (|k1|, |k2| < 2^127 plus the carry bit of the signed recoding).
  static const bool glv_off = getenv("TKM_MSM_NO_GLV") != nullptr;  // developer knob: plain 256-bit digits
  const bool use_glv = !fixed_c && !glv_off;
  for (uint32_t c = 4; c <= 20; c++) {
    uint32_t W = use_glv ? 2 * ((128 + c - 1) / c) : (256 + c - 1) / c;
    double cost = (double)W * ((double)n + 3.0 * (double)(1u << (c - 1)));
    if (cost < best) {
      best = cost;
      bc = c;
    }
  }
  if (fixed_c) bc = fixed_c;
  else if (const char *e = getenv("TKM_MSM_C")) {  // developer knob: force the window width (scripts/msm_c_sweep.py)
    const uint32_t v = (uint32_t)atoi(e);
    if (v >= 4 && v <= 20) bc = v;
  }
  MsmGeom m;
  m.c = bc;
  m.glv = use_glv ? 1 : 0;
  m.Wh = use_glv ? (128 + bc - 1) / bc : (256 + bc - 1) / bc;
  m.Wd = use_glv ? 2 * m.Wh : m.Wh;
  m.val_stride = table_stride;
  m.W = fixed_c ? 1 : m.Wd;
  m.logB = bc - 1;
  m.B = 1u << m.logB;
  m.nbuckets = m.W * m.B;
  uint32_t want_logg = n < ((size_t)1 << 18) ? 3 : 4;  // segment length 8 / 16 (measured: shorter chains win when few buckets)
  if (const char *e = getenv("TKM_MSM_LOGG")) want_logg = (uint32_t)atoi(e);  // developer knob
  m.logg = m.logB < want_logg ? m.logB : want_logg;
  m.g = 1u << m.logg;
  m.nseg = m.B >> m.logg;
  m.nbits = m.logB - m.logg;
  return m;
}

__device__ __forceinline__ G1Affine ldg_affine(const G1Affine *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 r[6];
#pragma unroll
  for (int i = 0; i < 6; i++) r[i] = __ldg(q + i);
  G1Affine a;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    a.x.v[4 * i + 0] = r[i].x; a.x.v[4 * i + 1] = r[i].y; a.x.v[4 * i + 2] = r[i].z; a.x.v[4 * i + 3] = r[i].w;
    a.y.v[4 * i + 0] = r[i + 3].x; a.y.v[4 * i + 1] = r[i + 3].y; a.y.v[4 * i + 2] = r[i + 3].z; a.y.v[4 * i + 3] = r[i + 3].w;
  }
  return a;
}

// ---------------------------------------------------------------- 1. digit decomposition
__global__ void __launch_bounds__(256) k_decompose(const Fr *__restrict__ scalars, int scalars_mont, size_t s_row_stride,
                                                   size_t b_row_stride, const uint32_t *__restrict__ gather,
                                                   uint32_t rows, uint32_t cols, MsmGeom m, uint32_t *__restrict__ keys,
                                                   uint32_t *__restrict__ vals) {
  const size_t n = (size_t)rows * cols;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    uint32_t i = (uint32_t)(k / cols), j = (uint32_t)(k % cols);
    Fr s = scalars[(size_t)i * s_row_stride + j];
    if (scalars_mont) s = s.from_mont();
    uint32_t base_idx = gather ? gather[k] : (uint32_t)((size_t)i * b_row_stride + j);
    GlvSplit sp;
    if (m.glv) sp = glv_split(s.v);
    const uint32_t halves = m.glv ? 2 : 1;
    for (uint32_t h = 0; h < halves; h++) {
      const uint32_t *limbs = m.glv ? sp.mag[h] : s.v;
      const uint32_t nl = m.glv ? 4 : 8;
      const uint32_t flip = m.glv ? sp.neg[h] : 0;
      uint32_t carry = 0;
      for (uint32_t wi = 0; wi < m.Wh; wi++) {
        uint32_t mag, neg;
        signed_digit(limbs, nl, wi, m.c, carry, mag, neg);
        const uint32_t w = h * m.Wh + wi;
        size_t slot = (size_t)w * n + k;
        keys[slot] = mag ? ((m.W == 1 ? 0u : w * m.B) + mag - 1) : m.nbuckets;
        vals[slot] = (base_idx + w * m.val_stride) | ((neg ^ flip) << 31);
      }
    }
  }
}

// ---------------------------------------------------------------- 3. chunked bucket accumulation
__device__ __forceinline__ void store_xyzz(G1Xyzz *dst, const G1Xyzz &p) {
  uint4 *q = reinterpret_cast<uint4 *>(dst);
  const Fq *f[4] = {&p.X, &p.Y, &p.ZZ, &p.ZZZ};
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int i = 0; i < 3; i++) q[3 * c + i] = make_uint4(f[c]->v[4 * i], f[c]->v[4 * i + 1], f[c]->v[4 * i + 2], f[c]->v[4 * i + 3]);
}
__device__ __forceinline__ G1Xyzz load_xyzz(const G1Xyzz *src) {
  const uint4 *q = reinterpret_cast<const uint4 *>(src);
  G1Xyzz p;
  Fq *f[4] = {&p.X, &p.Y, &p.ZZ, &p.ZZZ};
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int i = 0; i < 3; i++) {
      uint4 v = q[3 * c + i];
      f[c]->v[4 * i] = v.x; f[c]->v[4 * i + 1] = v.y; f[c]->v[4 * i + 2] = v.z; f[c]->v[4 * i + 3] = v.w;
    }
  return p;
}

constexpr int ACC_THREADS = 128;
// Resident CTAs per SM the accumulation kernel is compiled for (register cap 168 at 3).  Overridable for the occupancy
// probe of scripts/ubench (a mixed-addition stream alone is 3 % faster at 2 CTAs/SM with 206 registers).
#ifndef TKM_ACC_MIN_BLOCKS
#define TKM_ACC_MIN_BLOCKS 3
#endif

__global__ void __launch_bounds__(ACC_THREADS, TKM_ACC_MIN_BLOCKS) k_accumulate(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                                           size_t M, uint32_t chunk, const G1Affine *__restrict__ bases,
                                                           uint32_t invalid_key, G1Xyzz *__restrict__ buckets,
                                                           uint32_t *__restrict__ pkeys, G1Xyzz *__restrict__ ppts,
                                                           size_t nthreads) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= nthreads) return;
  size_t start = t * chunk, end = start + chunk;
  if (end > M) end = M;
  uint32_t cur = keys[start];
  G1Xyzz acc = G1Xyzz::identity();
  bool first_run = true;
  // software pipeline: the next base is in flight while the current addition runs
  uint32_t nk = cur, nv = vals[start];
  G1Affine npt = (nk != invalid_key) ? ldg_affine(bases + (nv & 0x7fffffffu)) : G1Affine::identity();
  for (size_t i = start; i < end; i++) {
    uint32_t k = nk, v = nv;
    G1Affine pt = npt;
    if (k == invalid_key && cur == invalid_key) break;  // sorted last: nothing but zero digits from here on
    if (i + 1 < end) {
      nk = keys[i + 1];
      nv = vals[i + 1];
      if (nk != invalid_key) npt = ldg_affine(bases + (nv & 0x7fffffffu));
    }
    if (k != cur) {
      if (first_run) {
        pkeys[2 * t] = cur;
        store_xyzz(ppts + 2 * t, acc);
        first_run = false;
      } else {
        store_xyzz(buckets + cur, acc);
      }
      acc = G1Xyzz::identity();
      cur = k;
      if (k == invalid_key) break;
    }
    if (v >> 31) pt.y = pt.y.neg();
    g1_madd(acc, pt);
  }
  if (first_run) {
    pkeys[2 * t] = cur;
    store_xyzz(ppts + 2 * t, acc);
    pkeys[2 * t + 1] = cur;
    store_xyzz(ppts + 2 * t + 1, G1Xyzz::identity());
  } else {
    pkeys[2 * t + 1] = cur;
    store_xyzz(ppts + 2 * t + 1, acc);
  }
}

// ---------------------------------------------------------------- 4. warp-cooperative segmented reduction
__device__ __forceinline__ Fq shfl_down_fq(const Fq &a, int d) {
  Fq r;
#pragma unroll
  for (int i = 0; i < Fq::N; i++) r.v[i] = __shfl_down_sync(0xffffffffu, a.v[i], d);
  return r;
}

__device__ __forceinline__ Fq shfl_fq(const Fq &a, int src_lane) {
  Fq r;
#pragma unroll
  for (int i = 0; i < Fq::N; i++) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src_lane);
  return r;
}

constexpr int SEG_THREADS = 128;

// Entries [32*w, 32*w+32) belong to warp w.  Sorted keys: equal keys are contiguous.  First run of a
// warp -> slot 2w, last run -> slot 2w+1 (identity if the warp holds one run), runs strictly inside
// the warp are final.  When `last` (one warp left) every run is final.
__global__ void __launch_bounds__(SEG_THREADS) k_segreduce(const uint32_t *__restrict__ keys_in, const G1Xyzz *__restrict__ pts_in,
                                                          size_t P, G1Xyzz *__restrict__ buckets, uint32_t *__restrict__ keys_out,
                                                          G1Xyzz *__restrict__ pts_out, int last, uint32_t pad_key) {
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const size_t nwarps = (P + 31) >> 5;
  if (warp >= nwarps) return;
  const size_t i = warp * 32 + lane;
  uint32_t key = pad_key;
  G1Xyzz pt = G1Xyzz::identity();
  if (i < P) {
    key = keys_in[i];
    pt = load_xyzz(pts_in + i);
  }
  const uint32_t key_prev = __shfl_up_sync(0xffffffffu, key, 1);
  const bool head = (lane == 0) || (key_prev != key);
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t ok = __shfl_down_sync(0xffffffffu, key, d);
    G1Xyzz o;
    o.X = shfl_down_fq(pt.X, d);
    o.Y = shfl_down_fq(pt.Y, d);
    o.ZZ = shfl_down_fq(pt.ZZ, d);
    o.ZZZ = shfl_down_fq(pt.ZZZ, d);
    if (lane + d < 32 && ok == key) g1_add(pt, o);
  }
  const uint32_t key_first = __shfl_sync(0xffffffffu, key, 0);
  const uint32_t key_last = __shfl_sync(0xffffffffu, key, 31);
  if (head) {
    if (last) {
      store_xyzz(buckets + key, pt);
    } else if (key == key_first) {
      keys_out[2 * warp] = key;
      store_xyzz(pts_out + 2 * warp, pt);
      if (key_last == key_first) {
        keys_out[2 * warp + 1] = key;
        store_xyzz(pts_out + 2 * warp + 1, G1Xyzz::identity());
      }
    } else if (key == key_last) {
      keys_out[2 * warp + 1] = key;
      store_xyzz(pts_out + 2 * warp + 1, pt);
    } else {
      store_xyzz(buckets + key, pt);
    }
  }
}

// ---------------------------------------------------------------- 5. window reduction
// Segment s of window w covers digits d = s*g+1 .. s*g+g (bucket slots w*B + s*g .. +g-1).
// run = sum B_d, acc = sum (d - s*g) B_d.  Then  sum_d d*B_d = sum_s acc_s + g * sum_s s*run_s.
__global__ void __launch_bounds__(128) k_bucket_seg(const G1Xyzz *__restrict__ buckets, MsmGeom m, G1Xyzz *__restrict__ seg_acc,
                                                   G1Xyzz *__restrict__ seg_run) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t total = (size_t)m.W * m.nseg;
  if (t >= total) return;
  const G1Xyzz *b = buckets + t * m.g;
  G1Xyzz run = G1Xyzz::identity(), acc = G1Xyzz::identity();
  for (int k = (int)m.g - 1; k >= 0; k--) {
    G1Xyzz p = load_xyzz(b + k);
    g1_add(run, p);
    g1_add(acc, run);
  }
  store_xyzz(seg_acc + t, acc);
  store_xyzz(seg_run + t, run);
}

constexpr int BITS_THREADS = 128;
// Part (w, k) of the window reduction enters the window sum as g * 2^k * M_{w,k} (k < nbits) or as A_w (k == nbits):
// the k + log2(g) doublings are done here, by the first warp of the block that finished the part (cooperative doubling on
// replicas), so that all parts are weighted concurrently on different SMs and the window sum is a plain tree-sum.
__device__ __forceinline__ void weigh_and_store_part(const G1Xyzz &part, uint32_t k, uint32_t nbits, uint32_t logg, G1Xyzz *dst) {
  G1Xyzz a = part;
  const uint32_t nd = (k == nbits) ? 0 : k + logg;
  for (uint32_t i = 0; i < nd; i++) a = g1_dbl_coop4(a);
  if ((threadIdx.x & 31) == 0) store_xyzz(dst, a);
}
// Block (w, k, slice): k < nbits -> partial of M_k = sum over segments s with bit k set of run_s ; k == nbits -> partial of
// A = sum_s acc_s.  Each block tree-sums one slice of the window's segments (`splits` slices per (w, k)) so the
// reduction stays parallel when there are few windows (fixed-base tables: one window, 2^19 buckets).
__global__ void __launch_bounds__(BITS_THREADS) k_bucket_bits(const G1Xyzz *__restrict__ seg_acc, const G1Xyzz *__restrict__ seg_run,
                                                             MsmGeom m, uint32_t splits, G1Xyzz *__restrict__ out) {
  __shared__ G1Xyzz sh[BITS_THREADS];
  const uint32_t slice = blockIdx.x % splits, wk = blockIdx.x / splits;
  const uint32_t w = wk / (m.nbits + 1), k = wk % (m.nbits + 1);
  const G1Xyzz *src = (k == m.nbits ? seg_acc : seg_run) + (size_t)w * m.nseg;
  const uint32_t per = (m.nseg + splits - 1) / splits;
  const uint32_t lo = slice * per, hi = (lo + per < m.nseg) ? lo + per : m.nseg;
  G1Xyzz acc = G1Xyzz::identity();
  for (uint32_t s = lo + threadIdx.x; s < hi; s += BITS_THREADS) {
    if (k == m.nbits || ((s >> k) & 1)) {
      G1Xyzz p = load_xyzz(src + s);
      g1_add(acc, p);
    }
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int stride = BITS_THREADS / 2; stride > 0; stride >>= 1) {
    if ((int)threadIdx.x < stride) {
      G1Xyzz a = sh[threadIdx.x], b = sh[threadIdx.x + stride];
      g1_add(a, b);
      sh[threadIdx.x] = a;
    }
    __syncthreads();
  }
  if (splits > 1) {
    if (threadIdx.x == 0) store_xyzz(out + blockIdx.x, sh[0]);
    return;
  }
  if (threadIdx.x < 32) weigh_and_store_part(sh[0], k, m.nbits, m.logg, out + blockIdx.x);
}
// out[g] = sum of the `count` consecutive points in[g*count ..] (second stage of the sliced reduction), weighted like
// k_bucket_bits' single-slice result.
__global__ void __launch_bounds__(32) k_sum_groups(const G1Xyzz *__restrict__ in, uint32_t count, uint32_t nbits, uint32_t logg,
                                                  G1Xyzz *__restrict__ out) {
  __shared__ G1Xyzz sh[32];
  G1Xyzz acc = G1Xyzz::identity();
  for (uint32_t s = threadIdx.x; s < count; s += 32) {
    G1Xyzz p = load_xyzz(in + (size_t)blockIdx.x * count + s);
    g1_add(acc, p);
  }
  sh[threadIdx.x] = acc;
  __syncwarp();
  for (int stride = 16; stride > 0; stride >>= 1) {
    if ((int)threadIdx.x < stride) {
      G1Xyzz a = sh[threadIdx.x], b = sh[threadIdx.x + stride];
      g1_add(a, b);
      sh[threadIdx.x] = a;
    }
    __syncwarp();
  }
  weigh_and_store_part(sh[0], blockIdx.x % (nbits + 1), nbits, logg, out + blockIdx.x);
}

// Window sums S_w = A_w + g * sum_k 2^k M_{w,k} from the weighted parts: block w, one lane per part, shuffle tree.
__global__ void __launch_bounds__(32) k_window_sums(const G1Xyzz *__restrict__ parts, uint32_t nparts, G1Xyzz *__restrict__ window_sums) {
  const uint32_t lane = threadIdx.x;
  G1Xyzz pt = G1Xyzz::identity();
  if (lane < nparts) pt = load_xyzz(parts + (size_t)blockIdx.x * nparts + lane);  // nparts = nbits + 1 <= 22
#pragma unroll 1
  for (int d = 16; d > 0; d >>= 1) {
    G1Xyzz o;
    o.X = shfl_down_fq(pt.X, d);
    o.Y = shfl_down_fq(pt.Y, d);
    o.ZZ = shfl_down_fq(pt.ZZ, d);
    o.ZZZ = shfl_down_fq(pt.ZZZ, d);
    if (lane + d < 32) g1_add(pt, o);
  }
  if (lane == 0) store_xyzz(window_sums + blockIdx.x, pt);
}

// ---------------------------------------------------------------- 6. recombination
// One warp: Horner over the window sums, conversion to affine.
__global__ void __launch_bounds__(32) k_final(const G1Xyzz *__restrict__ window_sums, MsmGeom m, G1Affine *__restrict__ out_mont,
                                             uint32_t *__restrict__ out_canonical) {
  const uint32_t lane = threadIdx.x;
  // Horner over windows: a dependency chain of Wh*c doublings per half.  Every lane carries a replica of an accumulator
  // and groups of four lanes split each doubling's products (g1_dbl_coop4) to cut the chain latency.  With GLV the
  // groups with (lane>>2) even run the k1 chain (windows 0..Wh-1) and the odd groups the k2 chain (windows Wh..2Wh-1)
  // at the same time; the result is chain1 + phi(chain2), phi(X, Y, ZZ, ZZZ) = (beta*X, Y, ZZ, ZZZ).
  const uint32_t chain = m.glv ? ((lane >> 2) & 1) : 0;
  const uint32_t Wc = m.glv ? m.Wh : m.W;  // bucket windows per chain (W = 1 with fixed-base tables)
  G1Xyzz acc = G1Xyzz::identity();
  for (int w = (int)Wc - 1; w >= 0; w--) {
    for (uint32_t k = 0; k < m.c; k++) acc = g1_dbl_coop4(acc);
    G1Xyzz s = load_xyzz(window_sums + chain * Wc + w);
    g1_add(acc, s);
    __syncwarp();
  }
  if (m.glv) {
    G1Xyzz a, b;
    a.X = shfl_fq(acc.X, 0); a.Y = shfl_fq(acc.Y, 0); a.ZZ = shfl_fq(acc.ZZ, 0); a.ZZZ = shfl_fq(acc.ZZZ, 0);
    b.X = shfl_fq(acc.X, 4); b.Y = shfl_fq(acc.Y, 4); b.ZZ = shfl_fq(acc.ZZ, 4); b.ZZZ = shfl_fq(acc.ZZZ, 4);
    Fq beta;
#pragma unroll
    for (int i = 0; i < Fq::N; i++) beta.v[i] = glv::beta(i);
    b.X = b.X * beta.to_mont();
    acc = a;
    g1_add(acc, b);
  }
  G1Affine r = g1_to_affine_coop(acc);
  if (lane == 0) {
    if (out_mont) *out_mont = r;
    Fq x = r.x.from_mont(), y = r.y.from_mont();
    for (int i = 0; i < 12; i++) {
      out_canonical[i] = x.v[i];
      out_canonical[12 + i] = y.v[i];
    }
  }
}

__global__ void __launch_bounds__(256) k_fill_identity(G1Xyzz *p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    store_xyzz(p + i, G1Xyzz::identity());
}

// sets[0][i] += sets[1][i] + .. + sets[count-1][i]: bucket sets of the point ranges of one pipelined host MSM.
__global__ void __launch_bounds__(128) k_bucket_merge(G1Xyzz *__restrict__ sets, size_t set_stride, uint32_t count, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Xyzz acc = load_xyzz(sets + i);
  for (uint32_t k = 1; k < count; k++) {
    G1Xyzz p = load_xyzz(sets + k * set_stride + i);
    g1_add(acc, p);
  }
  store_xyzz(sets + i, acc);
}

__global__ void __launch_bounds__(256) k_g1_to_mont(const G1Affine *__restrict__ in, G1Affine *__restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    G1Affine a = in[i];
    a.x = a.x.to_mont();
    a.y = a.y.to_mont();
    out[i] = a;
  }
}

__global__ void __launch_bounds__(256) k_g1_from_mont(const G1Affine *__restrict__ in, G1Affine *__restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    G1Affine a = in[i];
    a.x = a.x.from_mont();
    a.y = a.y.from_mont();
    out[i] = a;
  }
}

// Fixed-base tables for a resident CRS: out[w*n + i] = 2^(c*w) * P_i in affine form, w < W.  With them every digit
// window of a commitment lands in ONE shared bucket set (no per-window reduction, no Horner tail) and the window can be
// wider (fewer additions per point).  One thread per base: c*(W-1) doublings, then all W-1 conversions to affine share a
// single inversion (Montgomery's trick over the thread's own points).  HBM capacity is what makes this affordable:
// 13 tables of the 8192 x 512 grid are 5.2 GB.
constexpr int PRE_MAX_W = 32;
__global__ void __launch_bounds__(128) k_crs_precompute(const G1Affine *__restrict__ base, size_t n, uint32_t c, uint32_t W,
                                                       G1Xyzz *__restrict__ tmp, G1Affine *__restrict__ out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine P = base[i];
  out[i] = P;
  G1Xyzz Q = G1Xyzz::from_affine(P);
  Fq pref[PRE_MAX_W];
  Fq acc = Fq::one();
  for (uint32_t w = 1; w < W; w++) {
    for (uint32_t k = 0; k < c; k++) Q = g1_dbl(Q);
    store_xyzz(tmp + (size_t)(w - 1) * n + i, Q);
    pref[w - 1] = acc;
    if (!Q.is_identity()) acc = acc * (Q.ZZ * Q.ZZZ);
  }
  Fq inv = acc.inv();
  for (uint32_t w = W - 1; w >= 1; w--) {
    G1Xyzz R = load_xyzz(tmp + (size_t)(w - 1) * n + i);
    G1Affine a = G1Affine::identity();
    if (!R.is_identity()) {
      Fq zinv = inv * pref[w - 1];  // 1 / (ZZ*ZZZ)
      inv = inv * (R.ZZ * R.ZZZ);
      a.x = R.X * (zinv * R.ZZZ);
      a.y = R.Y * (zinv * R.ZZ);
    }
    out[(size_t)w * n + i] = a;
  }
}

int32_t crs_precompute(tkm_ctx *ctx, const G1Affine *base, size_t n, uint32_t c, G1Affine **out_table, uint32_t *out_W) {
  if (c < 4 || c > 22) return fail(TKM_ERR_INVALID_ARGUMENT, "window bits %u out of range [4,22]", c);
  const uint32_t W = (256 + c - 1) / c;
  if (W > PRE_MAX_W) return fail(TKM_ERR_INVALID_ARGUMENT, "too many windows");
  if ((size_t)W * n >= ((size_t)1 << 31)) return fail(TKM_ERR_INVALID_ARGUMENT, "table of %u x %zu points exceeds the 31-bit index space", W, n);
  G1Affine *table = nullptr;
  cudaError_t e = cudaMalloc((void **)&table, (size_t)W * n * sizeof(G1Affine));
  if (e != cudaSuccess) return fail(TKM_ERR_ALLOCATION, "cudaMalloc(%zu) failed: %s", (size_t)W * n * sizeof(G1Affine), cudaGetErrorString(e));
  Scratch<G1Xyzz> tmp;
  int32_t st = tmp.alloc(ctx, (size_t)(W - 1) * n);
  if (st == TKM_OK) {
    k_crs_precompute<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(base, n, c, W, tmp.p, table);
    st = launch_check(ctx, "k_crs_precompute");
  }
  if (st == TKM_OK) {
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    if (e2 != cudaSuccess) st = fail(TKM_ERR_CUDA, "k_crs_precompute failed: %s", cudaGetErrorString(e2));
  }
  if (st != TKM_OK) {
    cudaFree(table);
    return st;
  }
  *out_table = table;
  *out_W = W;
  return TKM_OK;
}

int32_t g1_to_mont_dev(tkm_ctx *ctx, const G1Affine *in, G1Affine *out, size_t n) {
  if (n == 0) return TKM_OK;
  k_g1_to_mont<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(in, out, n);
  return launch_check(ctx, "k_g1_to_mont");
}

// One accumulation pass over `in`: digits -> sort -> chunked bucket accumulation -> segmented reduction of the partial
// list.  `buckets` must hold identities on entry (every bucket is written at most once per pass).
static int32_t msm_accumulate_pass(tkm_ctx *ctx, const MsmInput &in, const MsmGeom &m, G1Xyzz *buckets) {
  const size_t n = in.rows * in.cols;
  const size_t M = n * m.Wd;
  const uint32_t invalid = m.nbuckets;
  uint32_t key_bits = 1;
  while ((1ull << key_bits) <= invalid) key_bits++;

  Scratch<uint32_t> keys, vals, keys_s, vals_s;
  TKM_TRY(keys.alloc(ctx, M));
  TKM_TRY(vals.alloc(ctx, M));
  TKM_TRY(keys_s.alloc(ctx, M));
  TKM_TRY(vals_s.alloc(ctx, M));
  k_decompose<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(in.scalars, in.scalars_mont ? 1 : 0, in.scalar_row_stride,
                                                                        in.base_row_stride, in.idx, (uint32_t)in.rows,
                                                                        (uint32_t)in.cols, m, keys.p, vals.p);
  TKM_TRY(launch_check(ctx, "k_decompose"));

  size_t temp_bytes = 0;
  TKM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys.p, keys_s.p, vals.p, vals_s.p, (int)M, 0, (int)key_bits,
                                           ctx->stream));
  Scratch<uint8_t> temp;
  TKM_TRY(temp.alloc(ctx, temp_bytes));
  TKM_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, temp_bytes, keys.p, keys_s.p, vals.p, vals_s.p, (int)M, 0, (int)key_bits,
                                           ctx->stream));
  ctx->launches += 4;  // cub's histogram + onesweep passes (approximate; they are library launches)

  // Chunk length.  Every thread does the same amount of work (one chunk), so the launch runs in lock-step waves of
  // `cap` resident threads: pick the number of waves for chunks of at most ~256 entries (2 partial-list entries per chunk:
  // longer chunks shrink the segmented-reduction levels), then size the chunk so that the waves are full.
  int &occ = ctx->acc_occ;  // per context = per device
  if (!occ) {
    TKM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_accumulate, ACC_THREADS, 0));
    if (occ < 1) occ = 1;
  }
  const size_t cap = (size_t)ctx->sm_count * occ * ACC_THREADS;
  uint32_t target = 256;
  if (const char *e = getenv("TKM_MSM_CHUNK")) target = (uint32_t)atoi(e);  // developer knob
  if (target < 8) target = 8;
  const size_t waves = (M + cap * target - 1) / (cap * target);
  uint32_t chunk = (uint32_t)((M + waves * cap - 1) / (waves * cap));
  if (chunk < 8) chunk = 8;
  const size_t T = (M + chunk - 1) / chunk;
  size_t P = 2 * T;
  Scratch<uint32_t> pk_a, pk_b;
  Scratch<G1Xyzz> pp_a, pp_b;
  TKM_TRY(pk_a.alloc(ctx, P));
  TKM_TRY(pp_a.alloc(ctx, P));
  const size_t P2 = 2 * ((P + 31) / 32);
  TKM_TRY(pk_b.alloc(ctx, P2));
  TKM_TRY(pp_b.alloc(ctx, P2));
  const unsigned acc_grid = (unsigned)((T + ACC_THREADS - 1) / ACC_THREADS);
  TKM_CUDA(cudaEventRecord(ctx->kev0, ctx->stream));
  k_accumulate<<<acc_grid, ACC_THREADS, 0, ctx->stream>>>(keys_s.p, vals_s.p, M, chunk, in.bases, invalid, buckets, pk_a.p, pp_a.p, T);
  TKM_CUDA(cudaEventRecord(ctx->kev1, ctx->stream));
  ctx->kernel_timed = true;
  TKM_TRY(launch_check(ctx, "k_accumulate"));

  uint32_t *kin = pk_a.p, *kout = pk_b.p;
  G1Xyzz *pin = pp_a.p, *pout = pp_b.p;
  for (;;) {
    const size_t nwarps = (P + 31) / 32;
    const int last = nwarps == 1;
    const size_t threads = nwarps * 32;
    k_segreduce<<<(unsigned)((threads + SEG_THREADS - 1) / SEG_THREADS), SEG_THREADS, 0, ctx->stream>>>(kin, pin, P, buckets, kout, pout,
                                                                                                  last, invalid);
    TKM_TRY(launch_check(ctx, "k_segreduce"));
    if (last) break;
    P = 2 * nwarps;
    uint32_t *tk = kin; kin = kout; kout = tk;
    G1Xyzz *tp = pin; pin = pout; pout = tp;
  }
  return TKM_OK;
}

// Window reduction of a filled bucket set into weighted `parts` and window sums on the context stream, then the
// recombination kernel (one warp, latency-bound: < 1 ms) on `final_stream`.  When that is a side stream the caller can already queue the next MSM: the
// serial tail overlaps the next accumulation instead of idling 147 SMs.
static int32_t msm_reduce_to(tkm_ctx *ctx, const MsmGeom &m, const G1Xyzz *buckets, G1Xyzz *parts, G1Xyzz *wsum, uint32_t *res_dev,
                             cudaStream_t final_stream, cudaEvent_t ready) {
  const size_t nsegs = (size_t)m.W * m.nseg;
  Scratch<G1Xyzz> seg_acc, seg_run;
  TKM_TRY(seg_acc.alloc(ctx, nsegs));
  TKM_TRY(seg_run.alloc(ctx, nsegs));
  k_bucket_seg<<<(unsigned)((nsegs + 127) / 128), 128, 0, ctx->stream>>>(buckets, m, seg_acc.p, seg_run.p);
  TKM_TRY(launch_check(ctx, "k_bucket_seg"));
  {
    // Slices per (window, bit).  The blocks are latency-bound tree-sums, so what matters is that the launch is ONE wave:
    // as many slices as fit the resident block slots (2^22 points: 192 groups -> 1 slice, no second stage; fixed-base
    // tables: 16 groups -> 18 slices), never more than one slice per BITS_THREADS segments.
    const uint32_t groups = m.W * (m.nbits + 1);
    int &bits_occ = ctx->bits_occ;  // per context = per device
    if (!bits_occ) {
      TKM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bits_occ, k_bucket_bits, BITS_THREADS, 0));
      if (bits_occ < 1) bits_occ = 1;
    }
    uint32_t splits = (uint32_t)(ctx->sm_count * bits_occ) / groups;
    const uint32_t max_useful = (m.nseg + BITS_THREADS - 1) / BITS_THREADS;
    if (splits > max_useful) splits = max_useful;
    if (const char *e = getenv("TKM_MSM_SPLITS")) splits = (uint32_t)atoi(e);  // developer knob
    if (splits < 1) splits = 1;
    if (splits > 64) splits = 64;
    if (splits == 1) {
      k_bucket_bits<<<groups, BITS_THREADS, 0, ctx->stream>>>(seg_acc.p, seg_run.p, m, 1, parts);
      TKM_TRY(launch_check(ctx, "k_bucket_bits"));
    } else {
      Scratch<G1Xyzz> sliced;
      TKM_TRY(sliced.alloc(ctx, (size_t)groups * splits));
      k_bucket_bits<<<groups * splits, BITS_THREADS, 0, ctx->stream>>>(seg_acc.p, seg_run.p, m, splits, sliced.p);
      TKM_TRY(launch_check(ctx, "k_bucket_bits"));
      k_sum_groups<<<groups, 32, 0, ctx->stream>>>(sliced.p, splits, m.nbits, m.logg, parts);
      TKM_TRY(launch_check(ctx, "k_sum_groups"));
    }
  }
  k_window_sums<<<m.W, 32, 0, ctx->stream>>>(parts, m.nbits + 1, wsum);
  TKM_TRY(launch_check(ctx, "k_window_sums"));
  if (final_stream != ctx->stream) {
    TKM_CUDA(cudaEventRecord(ready, ctx->stream));
    TKM_CUDA(cudaStreamWaitEvent(final_stream, ready, 0));
  }
  k_final<<<1, 32, 0, final_stream>>>(wsum, m, nullptr, res_dev);
  return launch_check(ctx, "k_final");
}

// Synchronous form: reads back the 96-byte canonical affine result.
static int32_t msm_reduce(tkm_ctx *ctx, const MsmGeom &m, const G1Xyzz *buckets, uint8_t out96[96]) {
  Scratch<G1Xyzz> parts, wsum;
  Scratch<uint32_t> res;
  TKM_TRY(parts.alloc(ctx, (size_t)m.W * (m.nbits + 1)));
  TKM_TRY(wsum.alloc(ctx, m.W));
  TKM_TRY(res.alloc(ctx, 24));
  TKM_TRY(msm_reduce_to(ctx, m, buckets, parts.p, wsum.p, res.p, ctx->stream, nullptr));
  TKM_CUDA(cudaMemcpyAsync(out96, res.p, 96, cudaMemcpyDeviceToHost, ctx->stream));
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  return TKM_OK;
}

// ---- asynchronous MSMs: tickets own the small buffers the side-stream tail reads
constexpr size_t TICKET_PARTS = 2048;  // >= W * (nbits + 1) for every geometry (W <= 64, nbits <= 21)
static int32_t ticket_acquire(tkm_ctx *ctx, int32_t *out) {
  if (!ctx->side_stream) {
    // highest priority: the one-warp tail should get an SM slot as soon as a CTA of the next accumulation retires
    int lo = 0, hi = 0;
    TKM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    TKM_CUDA(cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, hi));
  }
  for (int t = 0; t < TKM_MAX_TICKETS; t++) {
    tkm_ctx::Ticket &k = ctx->tickets[t];
    if (k.busy) continue;
    if (!k.parts) {
      TKM_CUDA(cudaMalloc((void **)&k.parts, (TICKET_PARTS + 64) * sizeof(G1Xyzz) + 128));
      TKM_CUDA(cudaMallocHost((void **)&k.host, 96));
      TKM_CUDA(cudaEventCreateWithFlags(&k.ready, cudaEventDisableTiming));
      TKM_CUDA(cudaEventCreateWithFlags(&k.done, cudaEventDisableTiming));
    }
    k.busy = true;
    k.zero = false;
    *out = t;
    return TKM_OK;
  }
  return fail(TKM_ERR_INVALID_ARGUMENT, "too many commitments in flight (%d): call tkm_commit_end first", TKM_MAX_TICKETS);
}

int32_t msm_run_async(tkm_ctx *ctx, const MsmInput &in, int32_t *out_ticket) {
  const size_t n = in.rows * in.cols;
  int32_t t;
  TKM_TRY(ticket_acquire(ctx, &t));
  tkm_ctx::Ticket &k = ctx->tickets[t];
  *out_ticket = t;
  if (n == 0) {
    k.zero = true;
    return TKM_OK;
  }
  int32_t st = TKM_OK;
  if (n > 0x7fffffffull / 32) st = fail(TKM_ERR_INVALID_ARGUMENT, "MSM size %zu too large", n);
  if (st == TKM_OK && in.idx && in.rows != 1) st = fail(TKM_ERR_INVALID_ARGUMENT, "indexed MSM must be one row");
  if (st == TKM_OK) {
    const MsmGeom m = pick_geom(n, in.pre_c, in.pre_stride);
    if ((size_t)m.W * (m.nbits + 1) > TICKET_PARTS || m.W > 64) st = fail(TKM_ERR_INTERNAL, "ticket buffers too small for this geometry");
    Scratch<G1Xyzz> buckets;
    if (st == TKM_OK) st = buckets.alloc(ctx, (size_t)m.nbuckets + 1);
    if (st == TKM_OK) {
      k_fill_identity<<<grid_for((size_t)m.nbuckets + 1, 256, ctx->sm_count), 256, 0, ctx->stream>>>(buckets.p, (size_t)m.nbuckets + 1);
      st = launch_check(ctx, "k_fill_identity");
    }
    if (st == TKM_OK) st = msm_accumulate_pass(ctx, in, m, buckets.p);
    G1Xyzz *wsum = k.parts + TICKET_PARTS;
    uint32_t *res = reinterpret_cast<uint32_t *>(k.parts + TICKET_PARTS + 64);
    if (st == TKM_OK) st = msm_reduce_to(ctx, m, buckets.p, k.parts, wsum, res, ctx->side_stream, k.ready);
    if (st == TKM_OK) {
      cudaError_t e = cudaMemcpyAsync(k.host, res, 96, cudaMemcpyDeviceToHost, ctx->side_stream);
      if (e == cudaSuccess) e = cudaEventRecord(k.done, ctx->side_stream);
      if (e != cudaSuccess) st = fail(TKM_ERR_CUDA, "queuing the MSM tail failed: %s", cudaGetErrorString(e));
    }
  }
  if (st != TKM_OK) k.busy = false;
  return st;
}

int32_t msm_wait(tkm_ctx *ctx, int32_t ticket, uint8_t out96[96]) {
  if (ticket < 0 || ticket >= TKM_MAX_TICKETS || !ctx->tickets[ticket].busy) return fail(TKM_ERR_INVALID_ARGUMENT, "invalid commitment ticket %d", ticket);
  tkm_ctx::Ticket &k = ctx->tickets[ticket];
  if (k.zero) {
    memset(out96, 0, 96);
  } else {
    cudaError_t e = cudaEventSynchronize(k.done);
    if (e != cudaSuccess) {
      k.busy = false;
      return fail(TKM_ERR_CUDA, "MSM tail failed: %s", cudaGetErrorString(e));
    }
    memcpy(out96, k.host, 96);
  }
  k.busy = false;
  return TKM_OK;
}

int32_t g1_from_mont_dev(tkm_ctx *ctx, const G1Affine *in, G1Affine *out, size_t n) {
  if (n == 0) return TKM_OK;
  k_g1_from_mont<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(in, out, n);
  return launch_check(ctx, "k_g1_from_mont");
}

int32_t msm_run(tkm_ctx *ctx, const MsmInput &in, uint8_t out96[96]) {
  const size_t n = in.rows * in.cols;
  if (n == 0) {  // msm_g1_bases returns the identity for empty input (group_structures/mod.rs:131-133)
    memset(out96, 0, 96);
    return TKM_OK;
  }
  if (n > 0x7fffffffull / 32) return fail(TKM_ERR_INVALID_ARGUMENT, "MSM size %zu too large", n);
  if (in.idx && in.rows != 1) return fail(TKM_ERR_INVALID_ARGUMENT, "indexed MSM must be one row");
  const MsmGeom m = pick_geom(n, in.pre_c, in.pre_stride);
  Scratch<G1Xyzz> buckets;
  TKM_TRY(buckets.alloc(ctx, (size_t)m.nbuckets + 1));
  k_fill_identity<<<grid_for((size_t)m.nbuckets + 1, 256, ctx->sm_count), 256, 0, ctx->stream>>>(buckets.p, (size_t)m.nbuckets + 1);
  TKM_TRY(launch_check(ctx, "k_fill_identity"));
  TKM_TRY(msm_accumulate_pass(ctx, in, m, buckets.p));
  return msm_reduce(ctx, m, buckets.p, out96);
}

// Host-buffer MSM (msm::msm with HostSlice scalars and bases, libs/src/iotools/mod.rs:2093-2099) as a pipeline: the
// point range is cut into `pieces`; piece k's scalars and bases travel on the copy stream while piece k-1 is decomposed,
// sorted and accumulated into its own bucket set on the compute stream (no read-modify-write in the hot loop); the sets
// are summed bucket-wise and reduced once at the end.
int32_t msm_host_pipelined(tkm_ctx *ctx, const uint8_t *scalars, const uint8_t *bases, size_t n, uint32_t pieces, uint8_t out96[96]) {
  if (n > 0x7fffffffull / 32) return fail(TKM_ERR_INVALID_ARGUMENT, "MSM size %zu too large", n);
  if (pieces < 1) pieces = 1;
  if (pieces > 16) pieces = 16;
  if (pieces > n) pieces = (uint32_t)n;
  if (!ctx->copy_stream) {
    TKM_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 17; i++) TKM_CUDA(cudaEventCreateWithFlags(&ctx->copy_ev[i], cudaEventDisableTiming));
  }
  // Piece boundaries.  The copy engine outruns the accumulation (~2.4 ms vs ~7 ms per 2^20 points), so only the first
  // piece's copy is exposed: cut the range in growing pieces (weights 1, 3, 4, 4, ..) -- a small first piece starts the
  // compute early, few large later pieces keep the per-piece overhead (sort, chunk tails, bucket merge) low.
  size_t bound[17];
  {
    uint32_t wsum = 0, acc = 0;
    for (uint32_t k = 0; k < pieces; k++) wsum += k == 0 ? 1 : (k == 1 ? 3 : 4);
    bound[0] = 0;
    for (uint32_t k = 0; k < pieces; k++) {
      acc += k == 0 ? 1 : (k == 1 ? 3 : 4);
      bound[k + 1] = k + 1 == pieces ? n : (size_t)((unsigned __int128)n * acc / wsum);
      if (bound[k + 1] <= bound[k]) bound[k + 1] = bound[k] + 1;  // n >= pieces keeps every piece non-empty
      if (bound[k + 1] > n) bound[k + 1] = n;
    }
  }
  const MsmGeom m = pick_geom(n);
  Scratch<Fr> ds;
  Scratch<G1Affine> db;
  Scratch<G1Xyzz> buckets;
  TKM_TRY(ds.alloc(ctx, n));
  TKM_TRY(db.alloc(ctx, n));
  const size_t set_stride = (size_t)m.nbuckets + 1;
  TKM_TRY(buckets.alloc(ctx, set_stride * pieces));
  // the staging buffers come from the compute stream's pool: the copy stream may touch them only after this point
  TKM_CUDA(cudaEventRecord(ctx->copy_ev[16], ctx->stream));
  TKM_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[16], 0));
  int32_t st = TKM_OK;
  for (uint32_t k = 0; k < pieces && st == TKM_OK; k++) {
    const size_t off = bound[k], cnt = bound[k + 1] - bound[k];
    cudaError_t e = cudaMemcpyAsync(ds.p + off, scalars + off * 32, cnt * 32, cudaMemcpyHostToDevice, ctx->copy_stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(db.p + off, bases + off * 96, cnt * 96, cudaMemcpyHostToDevice, ctx->copy_stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->copy_ev[k], ctx->copy_stream);
    if (e != cudaSuccess) st = fail(TKM_ERR_CUDA, "host-to-device copy of MSM piece %u failed: %s", k, cudaGetErrorString(e));
  }
  if (st == TKM_OK) {
    k_fill_identity<<<grid_for(set_stride * pieces, 256, ctx->sm_count), 256, 0, ctx->stream>>>(buckets.p, set_stride * pieces);
    st = launch_check(ctx, "k_fill_identity");
  }
  for (uint32_t k = 0; k < pieces && st == TKM_OK; k++) {
    const size_t off = bound[k], cnt = bound[k + 1] - bound[k];
    cudaError_t e = cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[k], 0);
    if (e != cudaSuccess) {
      st = fail(TKM_ERR_CUDA, "cudaStreamWaitEvent failed: %s", cudaGetErrorString(e));
      break;
    }
    st = g1_to_mont_dev(ctx, db.p + off, db.p + off, cnt);
    if (st != TKM_OK) break;
    MsmInput in;
    in.scalars = ds.p + off;
    in.scalars_mont = false;
    in.scalar_row_stride = cnt;
    in.bases = db.p + off;
    in.base_row_stride = cnt;
    in.rows = 1;
    in.cols = cnt;
    in.idx = nullptr;
    st = msm_accumulate_pass(ctx, in, m, buckets.p + k * set_stride);
  }
  if (st == TKM_OK && pieces > 1) {
    k_bucket_merge<<<(unsigned)((m.nbuckets + 127) / 128), 128, 0, ctx->stream>>>(buckets.p, set_stride, pieces, m.nbuckets);
    st = launch_check(ctx, "k_bucket_merge");
  }
  if (st != TKM_OK) {
    // the staging buffers are freed on the compute stream when this scope ends: drain the copies first
    cudaStreamSynchronize(ctx->copy_stream);
    return st;
  }
  return msm_reduce(ctx, m, buckets.p, out96);
}

}  // namespace tkm
