// BLS12-381 G1 multi-scalar multiplication for sm_100a: signed-digit Pippenger.
//
// Replaces msm::msm(scalars, bases, MSMConfig::default(), out[1]) as the reference calls it for every
// commitment (libs/src/iotools/mod.rs:2093-2099; libs/src/group_structures/mod.rs:108-114,135-141).
//
// Pipeline (no host synchronisation until the 96-byte result is read):
//   1. k_decompose     one thread per scalar: (optional from-Montgomery,) GLV split k = k1 + k2*lambda into
//                      two signed 127-bit halves (glv.cuh), signed c-bit digits of each half;
//                      emits (bucket key, base index | sign) pairs, window-major, zero digits keyed
//                      to a trash bucket that sorts last.  Bucket (window w, digit d) has key (d << log2 W) | w.
//   2. radix sort      stable, over the digit bits of the key only: the window bits are already in order in the
//                      window-major list (two 8-bit cub onesweep passes at c = 16).
//   3a. pair tree      (>= 2^24 digit entries) levels of run-aligned pairwise AFFINE additions with batched inversion:
//                      k_run_bounds / k_scan_* (all levels' offsets), then per level k_tree_fwd (denominators, prefix
//                      products), k_tree_l2_up / k_tree_l3 / k_tree_l2_down (one inversion per level) and
//                      k_tree_apply; the two halves of the bucket range run on two streams.
//   3b. k_accumulate   one thread per fixed-length chunk of the (remaining) sorted list, so work is balanced for
//                      ANY scalar distribution: XYZZ mixed additions; runs that lie strictly inside a chunk are final
//                      and go straight to their bucket, the first/last run of every chunk go to a (key, point) partial list.
//   4. k_segreduce     warp-cooperative segmented reduction of the partial list (shuffle tree of
//                      full XYZZ additions, 32 entries per warp), repeated until one warp remains.
//   5. k_bucket_seg /  parallel window reduction: running sums over 16-bucket segments, then per
//      k_bucket_bits / window a masked tree-sum per index bit (sum_d d*B_d = sum_k 2^k sum_{d: bit k} B_d), each
//      k_window_sums   part weighted by its 2^k where it is produced, window sums as shuffle tree-sums,
//   6. k_final         Horner over the windows of each half as two concurrent chains, sum = chain1 + phi(chain2),
//                      one inversion (binary extended Euclid) to affine.
//   Queued MSMs (tickets) run 5-6 on a side stream under the next MSM's 1-3.
// Integer-pipe bound: N*W bucket additions of 5 products + 1 squaring (affine, tree) or ~10 products (XYZZ) in Fq
// (SURVEY.md 8d); no tensor cores.
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
  uint32_t logWp;    // log2 of W rounded up to a power of two (0 when W == 1): bucket (w, d) lives at index (d << logWp) | w
  uint32_t nbuckets; // B << logWp (+1 trash bucket at index nbuckets); slots with w >= W stay empty when W is not a power of two
  uint32_t g;        // bucket segment length for the window reduction
  uint32_t logg;
  uint32_t nseg;     // segments per window = B/g
  uint32_t nbits;    // log2(nseg)
};

static MsmGeom pick_geom(size_t n, uint32_t fixed_c = 0, uint32_t table_stride = 0) {
  // Cost of a window width in wide-IMAD units.  Without the pair tree: W*(n + 3*2^(c-1)) chained XYZZ mixed additions
  // (2604 each; ~3 madd-equivalents per bucket of window reduction).  With it (>= 2^24 digit entries, >= 16 per bucket): the
  // levels take all but 8..16 entries per bucket as affine additions (1662), the rest stays XYZZ, every level has a fixed
  // cost (one inversion latency), the sort is 2 or 3 passes and the window reduction costs what was measured (2.1 ns per
  // bucket).  Measured on B200 (profiles/r02_msm_c_level_sweep.log): c = 16 beats the old model's c = 19 by 16 % at 2^23
  // (38.4 vs 45.8 ms) and 9 % at 2^24 (72.5 vs 79.3 ms) -- deep buckets are what the tree is good at.
  double best = 1e300;
  uint32_t bc = 8;
  // GLV halves cover 128 bits each (|k1|, |k2| < 2^127 plus the carry bit of the signed recoding).
  static const bool glv_off = getenv("TKM_MSM_NO_GLV") != nullptr;  // developer knob: plain 256-bit digits
  const bool use_glv = !fixed_c && !glv_off;
  for (uint32_t c = 4; c <= 20; c++) {
    const uint32_t W = use_glv ? 2 * ((128 + c - 1) / c) : (256 + c - 1) / c;
    const double B = (double)(1u << (c - 1)), M = (double)W * (double)n, avg = (double)n / B;
    uint32_t L = 0;
    if (avg >= 16.0 && M >= (double)(1u << 24))
      while ((8u << (L + 1)) <= avg && L < 8) L++;
    double cost;
    if (L == 0) {
      cost = 2604.0 * (double)W * ((double)n + 3.0 * B);
    } else {
      const double R = M / (double)(1u << L);
      const double sort_passes = (c + 7) / 8;  // 8-bit digits over the c key bits that are not pre-sorted
      cost = (M - R) * 1662.0 + R * 2604.0 + (double)W * B * 18000.0 + sort_passes * M * 75.0 + (double)L * 3.0e9;
    }
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
  // Digit-major bucket numbering, window in the LOW bits of the key.  k_decompose emits the digit list window-major, so
  // the list is already sorted by the low logWp key bits; a stable LSD radix sort over the remaining bits alone (the digit
  // and the trash flag: c bits) leaves it sorted by the whole key -- 16 bits = two 8-bit passes at c = 16 instead of the three
  // a 20-bit key needs.
  m.logWp = 0;
  while ((1u << m.logWp) < m.W) m.logWp++;
  m.nbuckets = m.B << m.logWp;
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
        keys[slot] = mag ? (((mag - 1) << m.logWp) | (m.W == 1 ? 0u : w)) : m.nbuckets;
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

// ---------------------------------------------------------------- 3a. affine pair tree with batched inversion
// The additions of one bucket form a chain acc += P in k_accumulate (XYZZ mixed additions, 10 products each).  They are
// associative, so the same sum can be taken as a binary tree: level l pairs the entries (2t, 2t+1) of every bucket's run
// of the sorted list and replaces them by their AFFINE sum -- lambda = (y2 - y1)/(x2 - x1), x3 = lambda^2 - x1 - x2,
// y3 = lambda (x1 - x3) - y1: 2 products + 1 squaring + one inversion.  All additions of a level are independent, so
// their denominators are inverted together with Montgomery's trick (3 products each): 5 products + 1 squaring per
// addition instead of 10, and the trick's single inversion per batch is amortised over millions of additions by a
// second batch level (k_tree_fwd2 / k_tree_inv / k_tree_bwd2).  Run-aligned pairing keeps every level a compact sorted
// list: bucket b holds m_l(b) = ceil(count(b) / 2^l) entries at offset off_l[b] (exclusive scans of m_l, all levels
// computed up front from the bucket counts), element j of level l+1 with t = j - off_{l+1}[b] is
// in_l[off_l[b] + 2t] (+ in_l[off_l[b] + 2t + 1] when 2t + 1 < m_l(b)).  After L levels the short remaining runs go
// through the chunked XYZZ accumulation below, which also absorbs any skew (hot buckets).  Exceptional pairs (P = Q:
// tangent slope with denominator 2y; P = -Q: identity, kept as (0, 0); identity operands) are exact.
// Thread t of T handles elements j = s*T + t, s < TREE_B: consecutive lanes touch consecutive elements (coalesced),
// and the prefix-product chain of a thread runs over its strided set.
constexpr uint32_t TREE_MAX_LEVELS = 8;
constexpr int TREE_THREADS = 128;
constexpr int TREE2_LEAVES = 256;        // thread products per block of the second batch level (a product tree in shared memory)
constexpr int TREE3_LEAVES = 512;        // block products the single top block can take
constexpr uint32_t TREE_MAX_T = TREE2_LEAVES * TREE2_LEAVES * TREE3_LEAVES;  // first-level threads two block tiers + the top block can serve

struct TreeLevel {
  // input of the level: level 0 reads the sorted digit list and gathers bases, deeper levels read the previous SoA output
  const uint32_t *vals;      // level 0: base index | sign << 31 per sorted entry
  const G1Affine *bases;     // level 0
  const uint4 *xpad;         // level 0, forward pass: the bases' x coordinates alone, one 64-byte slot each (one DRAM granule per gather)
  const Fq *in_x, *in_y;     // level >= 1
  const uint32_t *off_in;    // off_l[b], b <= nbuckets
  const uint32_t *off_out;   // off_{l+1}[b]
  const uint32_t *off_next;  // off_{l+2}[b] or null (no further level): keys of the next level are written by the apply pass
  const uint32_t *keys_out;  // bucket of every output element
  uint32_t *keys_next;
  const uint32_t *src_out;   // per output element: input index of its first operand | (has a partner) << 31
  uint32_t *src_next;
  Fq *out_x, *out_y;
  Fq *pref;                  // prefix products, one per output element
  Fq *tot;                   // product of every thread's denominators
  const Fq *inv_tot;         // their inverses (second batch level)
  uint32_t nbuckets;
  uint32_t b_lo, b_hi;       // this launch handles the elements of buckets [b_lo, b_hi): positions [off_out[b_lo], off_out[b_hi]) of the level
  uint32_t T;                // threads of the first batch level (stride of the element assignment)
  uint32_t level0;
};

struct TreePair {
  G1Affine p1, p2;
  uint32_t kind;  // 0: copy p1, 1: copy p2, 2: identity, 3: chord (d = x2 - x1), 4: tangent (d = 2 y1)
};

__device__ __forceinline__ Fq ldg_fq(const Fq *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 r0 = __ldg(q), r1 = __ldg(q + 1), r2 = __ldg(q + 2);
  Fq a;
  a.v[0] = r0.x; a.v[1] = r0.y; a.v[2] = r0.z; a.v[3] = r0.w;
  a.v[4] = r1.x; a.v[5] = r1.y; a.v[6] = r1.z; a.v[7] = r1.w;
  a.v[8] = r2.x; a.v[9] = r2.y; a.v[10] = r2.z; a.v[11] = r2.w;
  return a;
}
__device__ __forceinline__ void st_fq(Fq *p, const Fq &a) {
  uint4 *q = reinterpret_cast<uint4 *>(p);
  q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
  q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
  q[2] = make_uint4(a.v[8], a.v[9], a.v[10], a.v[11]);
}

// Operands of output element j and the case it falls in.  WITH_Y = false loads only what the denominator needs (the y
// coordinates are fetched on demand in the rare cases that need them).
template <bool WITH_Y>
__device__ __forceinline__ void tree_classify(const TreeLevel &L, uint32_t i, bool partner, uint32_t v1, uint32_t v2, TreePair &pr, Fq &d);

// Operands of one output element (input entries i and, when it has a partner, i + 1; v1, v2 = their digit-list values at
// level 0) and the case it falls in.  WITH_Y = false loads only what the denominator needs (the y coordinates are fetched on
// demand in the rare cases that need them).
template <bool WITH_Y>
__device__ __forceinline__ void tree_load(const TreeLevel &L, uint32_t i, bool partner, uint32_t v1, uint32_t v2, TreePair &pr, Fq &d) {
  if (L.level0) {
    const G1Affine *q1 = L.bases + (v1 & 0x7fffffffu);
    if (WITH_Y) {
      pr.p1 = ldg_affine(q1);
      if (v1 >> 31) pr.p1.y = pr.p1.y.neg();
    } else {
      pr.p1.x = L.xpad ? ldg_fq(reinterpret_cast<const Fq *>(L.xpad + 4 * (size_t)(v1 & 0x7fffffffu))) : ldg_fq(&q1->x);
    }
    if (partner) {
      const G1Affine *q2 = L.bases + (v2 & 0x7fffffffu);
      if (WITH_Y) {
        pr.p2 = ldg_affine(q2);
        if (v2 >> 31) pr.p2.y = pr.p2.y.neg();
      } else {
        pr.p2.x = L.xpad ? ldg_fq(reinterpret_cast<const Fq *>(L.xpad + 4 * (size_t)(v2 & 0x7fffffffu))) : ldg_fq(&q2->x);
      }
    }
  } else {
    pr.p1.x = ldg_fq(L.in_x + i);
    if (WITH_Y) pr.p1.y = ldg_fq(L.in_y + i);
    if (partner) {
      pr.p2.x = ldg_fq(L.in_x + i + 1);
      if (WITH_Y) pr.p2.y = ldg_fq(L.in_y + i + 1);
    }
  }
  tree_classify<WITH_Y>(L, i, partner, v1, v2, pr, d);
}

template <bool WITH_Y>
__device__ __forceinline__ void tree_classify(const TreeLevel &L, uint32_t i, bool partner, uint32_t v1, uint32_t v2, TreePair &pr, Fq &d) {
  d = Fq::one();
  if (!partner) {
    pr.kind = 0;
    return;
  }
  const bool x_equal = pr.p1.x == pr.p2.x;
  if (!x_equal && !pr.p1.x.is_zero() && !pr.p2.x.is_zero()) {  // the common case: no coordinate is needed beyond x
    pr.kind = 3;
    d = pr.p2.x - pr.p1.x;
    return;
  }
  // rare: an operand may be the identity (0, 0), or the points share their x
  if (!WITH_Y) {
    if (L.level0) {
      pr.p1.y = ldg_fq(&L.bases[v1 & 0x7fffffffu].y);
      pr.p2.y = ldg_fq(&L.bases[v2 & 0x7fffffffu].y);
      if (v1 >> 31) pr.p1.y = pr.p1.y.neg();
      if (v2 >> 31) pr.p2.y = pr.p2.y.neg();
    } else {
      pr.p1.y = ldg_fq(L.in_y + i);
      pr.p2.y = ldg_fq(L.in_y + i + 1);
    }
  }
  if (pr.p1.is_identity()) {
    pr.kind = 1;
  } else if (pr.p2.is_identity()) {
    pr.kind = 0;
  } else if (!x_equal) {
    pr.kind = 3;
    d = pr.p2.x - pr.p1.x;
  } else if (pr.p1.y == pr.p2.y && !pr.p1.y.is_zero()) {
    pr.kind = 4;
    d = pr.p1.y.dbl();
  } else {
    pr.kind = 2;  // P + (-P)
  }
}

#ifndef TKM_TREE_PREFETCH
#define TKM_TREE_PREFETCH 0
#endif
// (Measured on B200: explicit L2 prefetches of the next element's operands make the tree SLOWER -- 2^22: 22.5 ms with them
// against 21.9 ms without, 2^24: 84.6 against 79.5 -- the line-granular prefetch over-fetches around the 96-byte points and
// the passes are already limited by DRAM traffic at level 0.  Kept behind a build knob.)
__device__ __forceinline__ void prefetch_l2(const void *p) {
#if TKM_TREE_PREFETCH
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

// Software pipeline of a thread's chain: the element's source descriptor (src: input index | partner << 31, written when the
// level's keys were) is loaded three elements ahead, its digit values (level 0) two ahead, and the operands it names are
// prefetched into L2 one element ahead -- the dependent gathers key -> index -> value -> point never stall the chain.
struct TreeFetch {
  uint32_t src[3];   // descriptors of the next three elements of the chain
  uint32_t v[2][2];  // level 0: digit values of the next two elements
};
template <bool WITH_Y>
__device__ __forceinline__ void tree_prefetch(const TreeLevel &L, uint32_t src, uint32_t v1, uint32_t v2, const Fq *pref_next) {
  const uint32_t i = src & 0x7fffffffu;
  const bool partner = src >> 31;
  if (L.level0) {
    if (WITH_Y || !L.xpad) {
      const char *q1 = reinterpret_cast<const char *>(L.bases + (v1 & 0x7fffffffu));
      prefetch_l2(q1);
      if (partner) {
        const char *q2 = reinterpret_cast<const char *>(L.bases + (v2 & 0x7fffffffu));
        prefetch_l2(q2);
      }
    } else {
      prefetch_l2(L.xpad + 4 * (size_t)(v1 & 0x7fffffffu));
      if (partner) prefetch_l2(L.xpad + 4 * (size_t)(v2 & 0x7fffffffu));
    }
  } else {
    const char *qx = reinterpret_cast<const char *>(L.in_x + i);
    prefetch_l2(qx);
    if (WITH_Y) prefetch_l2(reinterpret_cast<const char *>(L.in_y + i));
  }
  if (pref_next) prefetch_l2(pref_next);
}

// x coordinates of the bases an MSM touches, one 64-byte slot per base: the forward pass of level 0 gathers only x, and from
// the 96-byte (x, y) entries a 48-byte x costs one or two 64-byte DRAM granules plus the neighbouring sector.
__global__ void __launch_bounds__(256) k_tree_xpad(const G1Affine *__restrict__ bases, size_t row_stride, uint32_t rows, uint32_t cols, const uint32_t *__restrict__ gather,
                                                   uint32_t tables, uint32_t table_stride, uint4 *__restrict__ xpad) {
  // addresses the same index space as k_decompose's base indices: i*row_stride + j (or gather[k]) + w*table_stride
  const size_t n = (size_t)rows * cols;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n * tables; e += (size_t)gridDim.x * blockDim.x) {
    const size_t k = e % n, w = e / n;
    const size_t idx = (gather ? gather[k] : (k / cols) * row_stride + (k % cols)) + w * table_stride;
    const uint4 *q = reinterpret_cast<const uint4 *>(&bases[idx].x);
    uint4 *o = xpad + 4 * idx;
    o[0] = __ldg(q);
    o[1] = __ldg(q + 1);
    o[2] = __ldg(q + 2);
  }
}

__global__ void __launch_bounds__(TREE_THREADS) k_tree_fwd(const __grid_constant__ TreeLevel L) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L.T) return;
  const uint32_t j0 = L.off_out[L.b_lo], n = L.off_out[L.b_hi];
  const uint32_t B = (n - j0 + L.T - 1) / L.T;  // chain length from the actual element count: T is sized for the expected share of the bucket range
  Fq acc = Fq::one();
  auto pos = [&](uint32_t s) { return (uint64_t)j0 + (uint64_t)s * L.T + t; };  // element s of this thread's chain
  auto live = [&](uint32_t s) { return s < B && pos(s) < n; };
  auto load_vals = [&](uint32_t src, uint32_t *v) {
    if (!L.level0) return;
    const uint32_t i = src & 0x7fffffffu;
    v[0] = L.vals[i];
    v[1] = (src >> 31) ? L.vals[i + 1] : 0u;
  };
  // prime the pipeline: descriptors two elements ahead, digit values one ahead (loading the operands themselves one element
  // ahead into registers was measured too: 118 registers, no gain)
  uint32_t sc = live(0) ? L.src_out[pos(0)] : 0, s1 = live(1) ? L.src_out[pos(1)] : 0, s2 = live(2) ? L.src_out[pos(2)] : 0;
  uint32_t vc[2] = {0, 0}, v1[2] = {0, 0};
  if (live(0)) load_vals(sc, vc);
  if (live(1)) load_vals(s1, v1);
  for (uint32_t s = 0; s < B; s++) {
    const uint64_t j = pos(s);
    if (j >= n) break;
    if (live(s + 1)) tree_prefetch<false>(L, s1, v1[0], v1[1], nullptr);  // (compiled out unless TKM_TREE_PREFETCH)
    uint32_t v2[2] = {0, 0};
    if (live(s + 2)) load_vals(s2, v2);                                    // digit values two ahead
    const uint32_t s3 = live(s + 3) ? L.src_out[pos(s + 3)] : 0;           // descriptor three ahead
    TreePair pr;
    Fq d;
    tree_load<false>(L, sc & 0x7fffffffu, sc >> 31, vc[0], vc[1], pr, d);
    st_fq(L.pref + j, acc);
    if (pr.kind >= 3) acc = acc * d;
    sc = s1; s1 = s2; s2 = s3;
    vc[0] = v1[0]; vc[1] = v1[1];
    v1[0] = v2[0]; v1[1] = v2[1];
  }
  st_fq(L.tot + t, acc);
}

// Second batch level over the T thread products: a product tree per block of TREE2_LEAVES in shared memory (log depth: a
// few product latencies instead of a serial chain), the block roots reduced the same way by ONE top block, whose single
// thread-0 inversion (binary extended Euclid) is the only inversion of the whole level; the inverses flow back down the
// same trees.  tree[LEAVES + i] = leaf i, tree[k] = tree[2k] * tree[2k+1]; going down, a node's slot is overwritten by its
// inverse: inv(left) = inv(node) * right, inv(right) = inv(node) * left.
template <int LEAVES>
__device__ __forceinline__ void tree_up(Fq *tree, int tid) {
  for (int w = LEAVES / 2; w >= 1; w >>= 1) {
    __syncthreads();
    if (tid < w) tree[w + tid] = tree[2 * (w + tid)] * tree[2 * (w + tid) + 1];
  }
  __syncthreads();
}
template <int LEAVES>
__device__ __forceinline__ void tree_down(Fq *tree, int tid) {  // tree[1] holds the inverse of the root product on entry
  for (int w = 1; w < LEAVES; w <<= 1) {
    __syncthreads();
    if (tid < w) {
      const Fq iv = tree[w + tid], l = tree[2 * (w + tid)], r = tree[2 * (w + tid) + 1];
      tree[2 * (w + tid)] = iv * r;
      tree[2 * (w + tid) + 1] = iv * l;
    }
  }
  __syncthreads();
}
__global__ void __launch_bounds__(TREE2_LEAVES) k_tree_l2_up(const Fq *__restrict__ tot, uint32_t T, Fq *__restrict__ blk) {
  __shared__ Fq tree[2 * TREE2_LEAVES];
  const int tid = threadIdx.x;
  const uint32_t v = blockIdx.x * TREE2_LEAVES + tid;
  tree[TREE2_LEAVES + tid] = v < T ? ldg_fq(tot + v) : Fq::one();
  tree_up<TREE2_LEAVES>(tree, tid);
  if (tid == 0) st_fq(blk + blockIdx.x, tree[1]);
}
__global__ void __launch_bounds__(TREE3_LEAVES) k_tree_l3(Fq *__restrict__ blk, uint32_t nblk) {  // in place: products -> inverses
  __shared__ Fq tree[2 * TREE3_LEAVES];
  const int tid = threadIdx.x;
  tree[TREE3_LEAVES + tid] = (uint32_t)tid < nblk ? blk[tid] : Fq::one();
  tree_up<TREE3_LEAVES>(tree, tid);
  if (tid == 0) tree[1] = tree[1].inv_fast();
  tree_down<TREE3_LEAVES>(tree, tid);
  if ((uint32_t)tid < nblk) st_fq(blk + tid, tree[TREE3_LEAVES + tid]);
}
// (inv_tot may alias tot: a block reads its own segment before the first barrier and writes it after the last)
__global__ void __launch_bounds__(TREE2_LEAVES) k_tree_l2_down(const Fq *tot, const Fq *blk_inv, uint32_t T, Fq *inv_tot) {
  __shared__ Fq tree[2 * TREE2_LEAVES];
  const int tid = threadIdx.x;
  const uint32_t v = blockIdx.x * TREE2_LEAVES + tid;
  tree[TREE2_LEAVES + tid] = v < T ? tot[v] : Fq::one();
  tree_up<TREE2_LEAVES>(tree, tid);
  if (tid == 0) tree[1] = blk_inv[blockIdx.x];
  tree_down<TREE2_LEAVES>(tree, tid);
  if (v < T) st_fq(inv_tot + v, tree[TREE2_LEAVES + tid]);
}

__global__ void __launch_bounds__(TREE_THREADS, 4) k_tree_apply(const __grid_constant__ TreeLevel L) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L.T) return;
  const uint32_t j0 = L.off_out[L.b_lo], n = L.off_out[L.b_hi];
  const uint32_t B = (n - j0 + L.T - 1) / L.T;
  uint32_t cnt = B;
  while (cnt && (uint64_t)j0 + (uint64_t)(cnt - 1) * L.T + t >= n) cnt--;
  if (!cnt) return;
  Fq run = ldg_fq(L.inv_tot + t);
  // the chain runs backwards: element cnt-1 first.  k counts elements already consumed; element index s = cnt - 1 - k.
  auto pos = [&](uint32_t k) { return (uint32_t)((uint64_t)j0 + (uint64_t)(cnt - 1 - k) * L.T + t); };
  auto live = [&](uint32_t k) { return k < cnt; };
  auto load_vals = [&](uint32_t src, uint32_t *v) {
    if (!L.level0) return;
    const uint32_t i = src & 0x7fffffffu;
    v[0] = L.vals[i];
    v[1] = (src >> 31) ? L.vals[i + 1] : 0u;
  };
  uint32_t sc = L.src_out[pos(0)], s1 = live(1) ? L.src_out[pos(1)] : 0, s2 = live(2) ? L.src_out[pos(2)] : 0;
  uint32_t vc[2] = {0, 0}, v1[2] = {0, 0};
  load_vals(sc, vc);
  if (live(1)) load_vals(s1, v1);
  for (uint32_t k = 0; k < cnt; k++) {
    const uint32_t j = pos(k);
    if (live(k + 1)) tree_prefetch<true>(L, s1, v1[0], v1[1], L.pref + pos(k + 1));
    uint32_t v2[2] = {0, 0};
    if (live(k + 2)) load_vals(s2, v2);
    const uint32_t s3 = live(k + 3) ? L.src_out[pos(k + 3)] : 0;
    // loads that feed the end of the iteration are issued first: the bucket of this element (for the next level's
    // descriptors) and its prefix product; the offsets that depend on the bucket follow once the operands have arrived
    const uint32_t b = L.off_next ? L.keys_out[j] : 0u;
    const Fq pf = ldg_fq(L.pref + j);
    TreePair pr;
    Fq d;
    tree_load<true>(L, sc & 0x7fffffffu, sc >> 31, vc[0], vc[1], pr, d);
    uint32_t o_b = 0, o_b1 = 0, o_n = 0;
    if (L.off_next) {
      o_b = L.off_out[b];
      o_b1 = L.off_out[b + 1];
      o_n = L.off_next[b];
    }
    G1Affine r;
    if (pr.kind >= 3) {
      const Fq dinv = run * pf;
      run = run * d;
      Fq num;
      if (pr.kind == 3) {
        num = pr.p2.y - pr.p1.y;
      } else {
        const Fq xx = pr.p1.x.sqr();
        num = xx.dbl() + xx;
      }
      const Fq lam = num * dinv;
      r.x = lam.sqr() - pr.p1.x - pr.p2.x;
      r.y = lam * (pr.p1.x - r.x) - pr.p1.y;
    } else if (pr.kind == 0) {
      r = pr.p1;
    } else if (pr.kind == 1) {
      r = pr.p2;
    } else {
      r = G1Affine::identity();
    }
    st_fq(L.out_x + j, r.x);
    st_fq(L.out_y + j, r.y);
    if (L.off_next) {  // descriptors and keys of the next level: even positions of this level's runs name their pair
      const uint32_t tp = j - o_b;
      if (!(tp & 1)) {
        const uint32_t jn = o_n + (tp >> 1);
        L.keys_next[jn] = b;
        L.src_next[jn] = j | ((tp + 1 < o_b1 - o_b) ? 0x80000000u : 0u);
      }
    }
    sc = s1; s1 = s2; s2 = s3;
    vc[0] = v1[0]; vc[1] = v1[1];
    v1[0] = v2[0]; v1[1] = v2[1];
  }
}

// Bucket counts of the sorted key list (run boundaries), then all levels' offsets in one three-phase scan:
// off[l][b] = sum_{b' < b} ceil(count(b') / 2^l), b <= nbuckets.
__global__ void __launch_bounds__(256) k_run_bounds(const uint32_t *__restrict__ keys, size_t M, uint32_t nbuckets, uint32_t *__restrict__ start,
                                                    uint32_t *__restrict__ end) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < M; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t k = keys[i];
    if (k >= nbuckets) continue;  // zero digits (trash key, sorted last)
    if (i == 0 || keys[i - 1] != k) start[k] = (uint32_t)i;
    if (i + 1 == M || keys[i + 1] != k) end[k] = (uint32_t)i + 1;
  }
}
constexpr uint32_t SCAN_ITEMS = 16, SCAN_THREADS = 256, SCAN_TILE = SCAN_ITEMS * SCAN_THREADS;
__device__ __forceinline__ uint32_t tree_len(uint32_t cnt, uint32_t l) { return (cnt + (1u << l) - 1) >> l; }
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const uint32_t *__restrict__ start, const uint32_t *__restrict__ end, uint32_t nbuckets,
                                                                uint32_t levels, uint32_t ntiles, uint32_t *__restrict__ tile_sums) {
  __shared__ uint32_t sh[SCAN_THREADS / 32];
  const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  for (uint32_t l = 0; l < levels; l++) {
    uint32_t sum = 0;
    for (uint32_t k = 0; k < SCAN_ITEMS; k++) {
      const uint32_t b = base + k;
      if (b < nbuckets) sum += tree_len(end[b] - start[b], l);
    }
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t tot = 0;
      for (uint32_t w = 0; w < SCAN_THREADS / 32; w++) tot += sh[w];
      tile_sums[l * ntiles + blockIdx.x] = tot;
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(1024) k_scan_tiles(uint32_t *__restrict__ tile_sums, uint32_t ntiles, uint32_t levels) {
  // one block per level; exclusive scan of the tile sums in place (ntiles is a few thousand)
  __shared__ uint32_t sh[1024];
  const uint32_t l = blockIdx.x;
  if (l >= levels) return;
  uint32_t *a = tile_sums + l * ntiles;
  uint32_t carry = 0;
  for (uint32_t base = 0; base < ntiles; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < ntiles ? a[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
      const uint32_t add = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
      __syncthreads();
      sh[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < ntiles) a[i] = carry + sh[threadIdx.x] - v;
    carry += sh[1023];
    __syncthreads();
  }
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_write(const uint32_t *__restrict__ start, const uint32_t *__restrict__ end, uint32_t nbuckets,
                                                            uint32_t levels, uint32_t ntiles, const uint32_t *__restrict__ tile_sums,
                                                            uint32_t *__restrict__ off /* [levels][nbuckets + 1] */) {
  __shared__ uint32_t sh[SCAN_THREADS];
  const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  for (uint32_t l = 0; l < levels; l++) {
    uint32_t v[SCAN_ITEMS], sum = 0;
    for (uint32_t k = 0; k < SCAN_ITEMS; k++) {
      const uint32_t b = base + k;
      v[k] = b < nbuckets ? tree_len(end[b] - start[b], l) : 0;
      sum += v[k];
    }
    sh[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < SCAN_THREADS; d <<= 1) {
      const uint32_t add = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
      __syncthreads();
      sh[threadIdx.x] += add;
      __syncthreads();
    }
    uint32_t run = tile_sums[l * ntiles + blockIdx.x] + sh[threadIdx.x] - sum;
    uint32_t *o = off + (size_t)l * (nbuckets + 1);
    for (uint32_t k = 0; k < SCAN_ITEMS; k++) {
      const uint32_t b = base + k;
      if (b <= nbuckets) o[b] = run;  // index nbuckets receives the total
      run += v[k];
    }
    __syncthreads();
  }
}
// keys of level 1: every even-position entry of a level-0 run names its bucket at off_1[b] + t/2 (buckets >= split go to the
// second half's buffer)
__global__ void __launch_bounds__(256) k_tree_keys1(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ off0, const uint32_t *__restrict__ off1,
                                                    uint32_t nbuckets, uint32_t split, uint32_t *__restrict__ keys1_lo, uint32_t *__restrict__ keys1_hi,
                                                    uint32_t *__restrict__ src1_lo, uint32_t *__restrict__ src1_hi) {
  const size_t n0 = off0[nbuckets];
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n0; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t b = keys[i];
    const uint32_t o = off0[b], t = (uint32_t)i - o;
    if (!(t & 1)) {
      const uint32_t m = off0[b + 1] - o, j = off1[b] + (t >> 1);
      (b < split ? keys1_lo : keys1_hi)[j] = b;
      (b < split ? src1_lo : src1_hi)[j] = (uint32_t)i | ((t + 1 < m) ? 0x80000000u : 0u);
    }
  }
}
// the second half's final level, copied next to the first half's (positions [off[split], off[nbuckets]) are its own)
__global__ void __launch_bounds__(256) k_tree_merge(const uint32_t *__restrict__ off, uint32_t split, uint32_t nbuckets, const uint4 *__restrict__ sx,
                                                    const uint4 *__restrict__ sy, const uint32_t *__restrict__ sk, uint4 *__restrict__ dx, uint4 *__restrict__ dy,
                                                    uint32_t *__restrict__ dk) {
  const size_t j0 = off[split], j1 = off[nbuckets];
  for (size_t e = (size_t)3 * j0 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)3 * j1; e += (size_t)gridDim.x * blockDim.x) {
    dx[e] = sx[e];
    dy[e] = sy[e];
    if (e % 3 == 0) dk[e / 3] = sk[e / 3];
  }
}

constexpr int ACC_THREADS = 128;
// Resident CTAs per SM the accumulation kernel is compiled for (register cap 168 at 3).  Overridable for the occupancy
// probe of scripts/ubench (a mixed-addition stream alone is 3 % faster at 2 CTAs/SM with 206 registers).
#ifndef TKM_ACC_MIN_BLOCKS
#define TKM_ACC_MIN_BLOCKS 3
#endif

// DIRECT = false: entry i is the base vals[i] (sign in bit 31) gathered from `bases`.  DIRECT = true: entry i is the affine
// point (px[i], py[i]) left by the pair tree, and the list length is read from *n_ptr (the launch is sized for its bound).
template <bool DIRECT>
__global__ void __launch_bounds__(ACC_THREADS, TKM_ACC_MIN_BLOCKS) k_accumulate(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                                           size_t M, uint32_t chunk, const G1Affine *__restrict__ bases,
                                                           const Fq *__restrict__ px, const Fq *__restrict__ py, const uint32_t *__restrict__ n_ptr,
                                                           uint32_t invalid_key, G1Xyzz *__restrict__ buckets,
                                                           uint32_t *__restrict__ pkeys, G1Xyzz *__restrict__ ppts,
                                                           size_t nthreads) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= nthreads) return;
  if (DIRECT) M = *n_ptr;
  size_t start = t * chunk, end = start + chunk;
  if (end > M) end = M;
  if (start >= M) {  // beyond the actual list (DIRECT: the launch covers the upper bound): two padding entries
    pkeys[2 * t] = invalid_key;
    pkeys[2 * t + 1] = invalid_key;
    store_xyzz(ppts + 2 * t, G1Xyzz::identity());
    store_xyzz(ppts + 2 * t + 1, G1Xyzz::identity());
    return;
  }
  auto fetch = [&](size_t i, uint32_t v) {
    if (DIRECT) return G1Affine{ldg_fq(px + i), ldg_fq(py + i)};
    return ldg_affine(bases + (v & 0x7fffffffu));
  };
  uint32_t cur = keys[start];
  G1Xyzz acc = G1Xyzz::identity();
  bool first_run = true;
  // software pipeline: the next base is in flight while the current addition runs
  uint32_t nk = cur, nv = DIRECT ? 0u : vals[start];
  G1Affine npt = (nk != invalid_key) ? fetch(start, nv) : G1Affine::identity();
  for (size_t i = start; i < end; i++) {
    uint32_t k = nk, v = nv;
    G1Affine pt = npt;
    if (k == invalid_key && cur == invalid_key) break;  // sorted last: nothing but zero digits from here on
    if (i + 1 < end) {
      nk = keys[i + 1];
      nv = DIRECT ? 0u : vals[i + 1];
      if (nk != invalid_key) npt = fetch(i + 1, nv);
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
// Segment s of window w covers digits d = s*g+1 .. s*g+g (bucket slots ((s*g + k) << logWp) | w, k < g).
// run = sum B_d, acc = sum (d - s*g) B_d.  Then  sum_d d*B_d = sum_s acc_s + g * sum_s s*run_s.
__global__ void __launch_bounds__(128) k_bucket_seg(const G1Xyzz *__restrict__ buckets, MsmGeom m, G1Xyzz *__restrict__ seg_acc,
                                                   G1Xyzz *__restrict__ seg_run) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t total = (size_t)m.W * m.nseg;
  if (t >= total) return;
  const uint32_t w = (uint32_t)(t / m.nseg), sgm = (uint32_t)(t % m.nseg);
  const G1Xyzz *b = buckets + (((size_t)sgm * m.g) << m.logWp) + w;
  G1Xyzz run = G1Xyzz::identity(), acc = G1Xyzz::identity();
  for (int k = (int)m.g - 1; k >= 0; k--) {
    G1Xyzz p = load_xyzz(b + ((size_t)k << m.logWp));
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

int32_t msm_build_xpad(tkm_ctx *ctx, const G1Affine *bases, size_t row_stride, size_t rows, size_t cols, uint32_t tables, size_t table_stride, uint4 *xpad) {
  if (rows * cols == 0) return TKM_OK;
  k_tree_xpad<<<grid_for(rows * cols * tables, 256, ctx->sm_count), 256, 0, ctx->stream>>>(bases, row_stride, (uint32_t)rows, (uint32_t)cols, nullptr, tables,
                                                                                         (uint32_t)table_stride, xpad);
  return launch_check(ctx, "k_tree_xpad");
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
  auto await_bases = [&]() -> int32_t {  // see MsmInput::bases_ready
    if (in.bases_ready) TKM_CUDA(cudaStreamWaitEvent(ctx->stream, in.bases_ready, 0));
    if (in.bases_canonical) TKM_TRY(g1_to_mont_dev(ctx, in.bases, const_cast<G1Affine *>(in.bases), n));
    return TKM_OK;
  };
  // the list leaves k_decompose sorted by the low logWp key bits (window-major); the stable sort covers the rest: digit + trash flag
  const int sort_lo = (int)m.logWp, sort_hi = (int)(m.logWp + m.logB + 1);

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
  TKM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys.p, keys_s.p, vals.p, vals_s.p, (int)M, sort_lo, sort_hi,
                                           ctx->stream));
  Scratch<uint8_t> temp;
  TKM_TRY(temp.alloc(ctx, temp_bytes));
  TKM_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, temp_bytes, keys.p, keys_s.p, vals.p, vals_s.p, (int)M, sort_lo, sort_hi,
                                           ctx->stream));
  ctx->launches += 4;  // cub's histogram + onesweep passes (approximate; they are library launches)

  // ---- affine pair tree (see "3a"): L levels of run-aligned pairwise affine additions with batched inversion, when the
  // buckets are deep enough to pay for the per-level fixed cost
  uint32_t L = 0;
  {
    const double avg = (double)M / ((double)m.W * m.B);  // entries per populated bucket slot
    if (avg >= 16.0 && M >= ((size_t)1 << 24)) {  // below ~16 M entries (2^20 points) the levels' fixed cost (one inversion latency each) eats the saving
      while ((8u << (L + 1)) <= avg && L < TREE_MAX_LEVELS) L++;  // leaves runs of 8..16 entries for the XYZZ pass (measured best at 2^22)
      // many buckets (fixed-base tables: one set of 2^19): the remaining list is still millions of entries, so a further level's
      // affine additions save more than its fixed cost (tables c = 20 at 2^22: 19.36 -> 19.09 ms with a fourth level)
      while (L < TREE_MAX_LEVELS && (M >> L) > ((size_t)6 << 20) && avg / (double)(1u << L) >= 4.0) L++;
    }
    if (const char *e = getenv("TKM_MSM_TREE_LEVELS")) {  // developer knob (0 = chained XYZZ additions only)
      const uint32_t v = (uint32_t)atoi(e);
      L = v <= TREE_MAX_LEVELS ? v : TREE_MAX_LEVELS;
    }
  }
  Scratch<uint32_t> t_start, t_end, t_off, t_tiles, t_keys[2][2], t_src[2][2];  // [part][side]
  Scratch<Fq> t_x[2][2], t_y[2][2], t_pref[2], t_tot, t_invtot, t_tot2, t_tot3;
  Scratch<uint4> t_xpad;
  const uint4 *xpad = in.xpad;
  const uint32_t nb = m.nbuckets;
  auto bound = [&](uint32_t l) { return (size_t)((M + ((size_t)1 << l) - 1) >> l) + nb; };  // >= entries of level l
  TKM_CUDA(cudaEventRecord(ctx->kev0, ctx->stream));
  if (L > 0) {
    if (M >= 0x7fffffffull) return fail(TKM_ERR_INVALID_ARGUMENT, "MSM too large for the 31-bit element indices of the pair tree");
    const uint32_t levels = L + 1;
    TKM_TRY(t_start.alloc(ctx, nb));
    TKM_TRY(t_end.alloc(ctx, nb));
    TKM_TRY(t_off.alloc(ctx, (size_t)levels * (nb + 1)));
    const uint32_t ntiles = (nb + 1 + SCAN_TILE - 1) / SCAN_TILE;
    TKM_TRY(t_tiles.alloc(ctx, (size_t)levels * ntiles));
    TKM_CUDA(cudaMemsetAsync(t_start.p, 0, (size_t)nb * 4, ctx->stream));
    TKM_CUDA(cudaMemsetAsync(t_end.p, 0, (size_t)nb * 4, ctx->stream));
    k_run_bounds<<<grid_for(M, 256, ctx->sm_count), 256, 0, ctx->stream>>>(keys_s.p, M, nb, t_start.p, t_end.p);
    TKM_TRY(launch_check(ctx, "k_run_bounds"));
    k_scan_tile_sums<<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(t_start.p, t_end.p, nb, levels, ntiles, t_tiles.p);
    TKM_TRY(launch_check(ctx, "k_scan_tile_sums"));
    k_scan_tiles<<<levels, 1024, 0, ctx->stream>>>(t_tiles.p, ntiles, levels);
    TKM_TRY(launch_check(ctx, "k_scan_tiles"));
    k_scan_write<<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(t_start.p, t_end.p, nb, levels, ntiles, t_tiles.p, t_off.p);
    TKM_TRY(launch_check(ctx, "k_scan_write"));
    // The levels of two halves of the bucket range are independent, so they run on two streams: while one half sits in
    // the latency-bound second batch level (a handful of small kernels and ONE inversion per level), the other half's
    // forward/apply kernels keep the SMs busy.  Each half has its own level buffers (a half running ahead must not overwrite
    // what the other still reads), sized for ALL entries: thread counts assume an even split, the chain length is derived on
    // the device from the actual count, so a skewed split only shifts time, never correctness.
    static const bool tree_one_stream = getenv("TKM_MSM_TREE_ONE_STREAM") != nullptr;  // developer knob
    const uint32_t parts = (!tree_one_stream && nb >= 2 && M >= ((size_t)1 << 22)) ? 2 : 1;
    for (uint32_t part = 0; part < parts; part++) {
      for (int side = 0; side < 2; side++) {
        const size_t cap = bound(1 + side);
        if (side == 1 && L < 2) break;
        TKM_TRY(t_x[part][side].alloc(ctx, cap));
        TKM_TRY(t_y[part][side].alloc(ctx, cap));
        TKM_TRY(t_keys[part][side].alloc(ctx, cap));
        TKM_TRY(t_src[part][side].alloc(ctx, cap));
      }
      TKM_TRY(t_pref[part].alloc(ctx, bound(1)));
    }
    // threads per level: about eight resident waves (4 CTAs of 128 per SM) so that blocks are balanced dynamically
    static const uint32_t tree_waves = getenv("TKM_MSM_TREE_WAVES") ? (uint32_t)atoi(getenv("TKM_MSM_TREE_WAVES")) : 8;  // developer knob
    auto pick_T = [&](size_t cap_elems, uint32_t *B_out) {
      size_t target = (size_t)ctx->sm_count * 4 * TREE_THREADS * (tree_waves ? tree_waves : 1);
      if (target > TREE_MAX_T) target = TREE_MAX_T;
      size_t B = (cap_elems + target - 1) / target;
      static const uint32_t min_b = getenv("TKM_MSM_TREE_MIN_B") ? (uint32_t)atoi(getenv("TKM_MSM_TREE_MIN_B")) : 32;  // developer knob
      if (B < min_b) B = min_b;  // small levels: fewer, longer chains keep the second batch level (its cost grows with T) small
      *B_out = (uint32_t)B;
      return (cap_elems + B - 1) / B;
    };
    size_t T1 = 0;  // the largest thread count of any level (the rounding of B makes it non-monotonic in the level)
    for (uint32_t l = 0; l < L; l++) {
      uint32_t Bl;
      const size_t Tl = pick_T((bound(l + 1) + parts - 1) / parts, &Bl);
      if (Tl > T1) T1 = Tl;
    }
    const size_t n2 = (T1 + TREE2_LEAVES - 1) / TREE2_LEAVES, n3 = (n2 + TREE2_LEAVES - 1) / TREE2_LEAVES;
    TKM_TRY(t_tot.alloc(ctx, T1 * parts));
    TKM_TRY(t_invtot.alloc(ctx, T1 * parts));
    TKM_TRY(t_tot2.alloc(ctx, n2 * parts));
    TKM_TRY(t_tot3.alloc(ctx, n3 * parts));
    auto off = [&](uint32_t l) { return t_off.p + (size_t)l * (nb + 1); };
    k_tree_keys1<<<grid_for(M, 256, ctx->sm_count), 256, 0, ctx->stream>>>(keys_s.p, off(0), off(1), nb, parts == 2 ? nb / 2 : nb, t_keys[0][0].p,
                                                                           t_keys[parts - 1][0].p, t_src[0][0].p, t_src[parts - 1][0].p);
    TKM_TRY(launch_check(ctx, "k_tree_keys1"));
    TKM_TRY(await_bases());  // everything above needed only the scalars
    if (!xpad && !in.idx && !in.pre_c) {  // plain bases (dense or a strided rectangle): a per-call x table pays for itself over the windows
      const size_t span = (in.rows - 1) * in.base_row_stride + in.cols;
      TKM_TRY(t_xpad.alloc(ctx, span * 4));
      TKM_TRY(msm_build_xpad(ctx, in.bases, in.base_row_stride, in.rows, in.cols, 1, 0, t_xpad.p));
      xpad = t_xpad.p;
    }
    if (parts == 2) {
      if (!ctx->tree_stream) {
        TKM_CUDA(cudaStreamCreateWithFlags(&ctx->tree_stream, cudaStreamNonBlocking));
        TKM_CUDA(cudaEventCreateWithFlags(&ctx->tree_fork, cudaEventDisableTiming));
        TKM_CUDA(cudaEventCreateWithFlags(&ctx->tree_join, cudaEventDisableTiming));
      }
      TKM_CUDA(cudaEventRecord(ctx->tree_fork, ctx->stream));
      TKM_CUDA(cudaStreamWaitEvent(ctx->tree_stream, ctx->tree_fork, 0));
    }
    for (uint32_t l = 0; l < L; l++) {
      const int o = l & 1, i = o ^ 1;  // level l+1 lands in side o; level l (l >= 1) sits in side i
      for (uint32_t part = 0; part < parts; part++) {
        cudaStream_t st = part ? ctx->tree_stream : ctx->stream;
        TreeLevel tl;
        memset(&tl, 0, sizeof tl);
        tl.level0 = l == 0;
        tl.vals = vals_s.p;
        tl.bases = in.bases;
        tl.xpad = xpad;
        tl.in_x = l ? t_x[part][i].p : nullptr;
        tl.in_y = l ? t_y[part][i].p : nullptr;
        tl.off_in = off(l);
        tl.off_out = off(l + 1);
        tl.off_next = l + 2 <= L ? off(l + 2) : nullptr;
        tl.keys_out = t_keys[part][o].p;
        tl.keys_next = t_keys[part][i].p;
        tl.src_out = t_src[part][o].p;
        tl.src_next = t_src[part][i].p;
        tl.out_x = t_x[part][o].p;
        tl.out_y = t_y[part][o].p;
        tl.pref = t_pref[part].p;
        Fq *tot = t_tot.p + (size_t)part * T1, *invtot = t_invtot.p + (size_t)part * T1, *tot2 = t_tot2.p + (size_t)part * n2, *tot3 = t_tot3.p + (size_t)part * n3;
        tl.tot = tot;
        tl.inv_tot = invtot;
        tl.nbuckets = nb;
        tl.b_lo = part ? nb / 2 : 0;
        tl.b_hi = (parts == 2 && part == 0) ? nb / 2 : nb;
        uint32_t B;
        const size_t T = pick_T((bound(l + 1) + parts - 1) / parts, &B);
        const uint32_t nblk = (uint32_t)((T + TREE2_LEAVES - 1) / TREE2_LEAVES);
        tl.T = (uint32_t)T;
        const unsigned g1 = (unsigned)((T + TREE_THREADS - 1) / TREE_THREADS);
        // (staggering the halves with an event -- second half's forward pass after the first half's -- was measured: 22.38 ms
        // against 22.19 ms unstaggered at 2^22, so the streams are left to interleave freely)
        k_tree_fwd<<<g1, TREE_THREADS, 0, st>>>(tl);
        TKM_TRY(launch_check(ctx, "k_tree_fwd"));
        // second batch level: products of 256 thread products per block, then (when there are more than 512 of those) a
        // second block tier, the top block with the level's only inversion, and back down
        k_tree_l2_up<<<nblk, TREE2_LEAVES, 0, st>>>(tot, (uint32_t)T, tot2);
        TKM_TRY(launch_check(ctx, "k_tree_l2_up"));
        if (nblk <= (uint32_t)TREE3_LEAVES) {
          k_tree_l3<<<1, TREE3_LEAVES, 0, st>>>(tot2, nblk);
          TKM_TRY(launch_check(ctx, "k_tree_l3"));
        } else {
          const uint32_t nblk2 = (nblk + TREE2_LEAVES - 1) / TREE2_LEAVES;
          k_tree_l2_up<<<nblk2, TREE2_LEAVES, 0, st>>>(tot2, nblk, tot3);
          TKM_TRY(launch_check(ctx, "k_tree_l2_up"));
          k_tree_l3<<<1, TREE3_LEAVES, 0, st>>>(tot3, nblk2);
          TKM_TRY(launch_check(ctx, "k_tree_l3"));
          k_tree_l2_down<<<nblk2, TREE2_LEAVES, 0, st>>>(tot2, tot3, nblk, tot2);
          TKM_TRY(launch_check(ctx, "k_tree_l2_down"));
        }
        k_tree_l2_down<<<nblk, TREE2_LEAVES, 0, st>>>(tot, tot2, (uint32_t)T, invtot);
        TKM_TRY(launch_check(ctx, "k_tree_l2_down"));
        k_tree_apply<<<g1, TREE_THREADS, 0, st>>>(tl);
        TKM_TRY(launch_check(ctx, "k_tree_apply"));
      }
    }
    if (parts == 2) {
      TKM_CUDA(cudaEventRecord(ctx->tree_join, ctx->tree_stream));
      TKM_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->tree_join, 0));
      const int f = (L & 1) ? 0 : 1;
      k_tree_merge<<<grid_for(bound(L) * 3 / 2, 256, ctx->sm_count), 256, 0, ctx->stream>>>(off(L), nb / 2, nb, (const uint4 *)t_x[1][f].p, (const uint4 *)t_y[1][f].p,
                                                                                           t_keys[1][f].p, (uint4 *)t_x[0][f].p, (uint4 *)t_y[0][f].p, t_keys[0][f].p);
      TKM_TRY(launch_check(ctx, "k_tree_merge"));
    }
    // per-level entry counts for bench.py's work accounting (read back lazily by tkm_msm_tree_stats)
    ctx->tree_levels = L;
    for (uint32_t l = 0; l <= L; l++)
      TKM_CUDA(cudaMemcpyAsync(ctx->tree_counts + l, off(l) + nb, 4, cudaMemcpyDeviceToHost, ctx->stream));
  } else {
    ctx->tree_levels = 0;
    TKM_TRY(await_bases());
  }
  const int fin = (L & 1) ? 0 : 1;  // side holding level L (L >= 1)
  const size_t Macc = L ? bound(L) : M;

  // Chunk length.  Every thread does the same amount of work (one chunk), so the launch runs in lock-step waves of
  // `cap` resident threads: pick the number of waves for chunks of at most ~256 entries (2 partial-list entries per chunk:
  // longer chunks shrink the segmented-reduction levels), then size the chunk so that the waves are full.
  int &occ = ctx->acc_occ;  // per context = per device
  if (!occ) {
    TKM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_accumulate<false>, ACC_THREADS, 0));
    if (occ < 1) occ = 1;
  }
  const size_t cap = (size_t)ctx->sm_count * occ * ACC_THREADS;
  uint32_t target = 256;
  if (const char *e = getenv("TKM_MSM_CHUNK")) target = (uint32_t)atoi(e);  // developer knob
  if (target < 8) target = 8;
  const size_t waves = (Macc + cap * target - 1) / (cap * target);
  uint32_t chunk = (uint32_t)((Macc + waves * cap - 1) / (waves * cap));
  if (chunk < 8) chunk = 8;
  const size_t T = (Macc + chunk - 1) / chunk;
  size_t P = 2 * T;
  Scratch<uint32_t> pk_a, pk_b;
  Scratch<G1Xyzz> pp_a, pp_b;
  TKM_TRY(pk_a.alloc(ctx, P));
  TKM_TRY(pp_a.alloc(ctx, P));
  const size_t P2 = 2 * ((P + 31) / 32);
  TKM_TRY(pk_b.alloc(ctx, P2));
  TKM_TRY(pp_b.alloc(ctx, P2));
  const unsigned acc_grid = (unsigned)((T + ACC_THREADS - 1) / ACC_THREADS);
  if (L)
    k_accumulate<true><<<acc_grid, ACC_THREADS, 0, ctx->stream>>>(t_keys[0][fin].p, nullptr, Macc, chunk, nullptr, t_x[0][fin].p, t_y[0][fin].p,
                                                                   t_off.p + (size_t)L * (nb + 1) + nb, invalid, buckets, pk_a.p, pp_a.p, T);
  else
    k_accumulate<false><<<acc_grid, ACC_THREADS, 0, ctx->stream>>>(keys_s.p, vals_s.p, M, chunk, in.bases, nullptr, nullptr, nullptr, invalid, buckets,
                                                                    pk_a.p, pp_a.p, T);
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

// Window reduction of a filled bucket set into weighted `parts` and window sums, then the recombination kernel (one warp).
// All of it is latency-bound (a few warps per scheduler, chains of full XYZZ additions: 1.7 ms at c = 16) and runs on
// `final_stream`.  When that is a side stream the caller can already queue the next MSM on the context stream: the whole
// reduction overlaps the next decomposition, sort and pair tree instead of leaving most of the device idle.  The scratch
// buffers are allocated on the context stream and handed to `final_stream` for their stream-ordered release.
static int32_t msm_reduce_to(tkm_ctx *ctx, const MsmGeom &m, const G1Xyzz *buckets, G1Xyzz *parts, G1Xyzz *wsum, uint32_t *res_dev,
                             cudaStream_t final_stream, cudaEvent_t ready) {
  const size_t nsegs = (size_t)m.W * m.nseg;
  Scratch<G1Xyzz> seg_acc, seg_run, sliced;
  TKM_TRY(seg_acc.alloc(ctx, nsegs));
  TKM_TRY(seg_run.alloc(ctx, nsegs));
  const uint32_t groups_all = m.W * (m.nbits + 1);
  TKM_TRY(sliced.alloc(ctx, (size_t)groups_all * 64));
  const cudaStream_t rs = final_stream;
  if (rs != ctx->stream) {
    TKM_CUDA(cudaEventRecord(ready, ctx->stream));  // buckets filled, scratch allocated
    TKM_CUDA(cudaStreamWaitEvent(rs, ready, 0));
    seg_acc.s = seg_run.s = sliced.s = rs;
  }
  k_bucket_seg<<<(unsigned)((nsegs + 127) / 128), 128, 0, rs>>>(buckets, m, seg_acc.p, seg_run.p);
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
      k_bucket_bits<<<groups, BITS_THREADS, 0, rs>>>(seg_acc.p, seg_run.p, m, 1, parts);
      TKM_TRY(launch_check(ctx, "k_bucket_bits"));
    } else {
      k_bucket_bits<<<groups * splits, BITS_THREADS, 0, rs>>>(seg_acc.p, seg_run.p, m, splits, sliced.p);
      TKM_TRY(launch_check(ctx, "k_bucket_bits"));
      k_sum_groups<<<groups, 32, 0, rs>>>(sliced.p, splits, m.nbits, m.logg, parts);
      TKM_TRY(launch_check(ctx, "k_sum_groups"));
    }
  }
  k_window_sums<<<m.W, 32, 0, rs>>>(parts, m.nbits + 1, wsum);
  TKM_TRY(launch_check(ctx, "k_window_sums"));
  k_final<<<1, 32, 0, rs>>>(wsum, m, nullptr, res_dev);
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
    if (st == TKM_OK) buckets.s = ctx->side_stream;  // the reduction reads the bucket set there: release it after that
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
  if (pieces > 16) pieces = 16;  // 0 = automatic layout
  if (pieces > n) pieces = (uint32_t)n;
  if (n < 4) pieces = 1;
  if (!ctx->copy_stream) {
    TKM_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 34; i++) TKM_CUDA(cudaEventCreateWithFlags(&ctx->copy_ev[i], cudaEventDisableTiming));
    for (int i = 0; i < 4; i++) TKM_CUDA(cudaEventCreate(&ctx->copy_t[i]));
  }
  // Piece boundaries.  On a box of its own the copy engine outruns the accumulation (512 MiB in 9 ms against 21 ms at 2^22), so
  // only the first piece's copy is exposed and every further piece costs a less efficient accumulation pass: two pieces, a
  // quarter of the points first.  When the copies take most of the call (eight ranks sharing the host's memory and PCIe roots:
  // 23 ms), what is exposed is the work left after the LAST bases have arrived: four pieces ending in a small one
  // (2 : 3 : 2 : 1).  The layout follows the copy share measured in the previous call on this context.
  size_t bound[17];
  static const uint32_t first_permille = getenv("TKM_MSM_HOST_FIRST") ? (uint32_t)atoi(getenv("TKM_MSM_HOST_FIRST")) : 0;  // developer knob: size of the first piece
  uint32_t weight[16];
  if (pieces == 0) {  // automatic
    if (ctx->h2d_share > 0.55f) ctx->h2d_copy_bound = true;
    else if (ctx->h2d_share > 0.f && ctx->h2d_share < 0.40f) ctx->h2d_copy_bound = false;
    bool copy_bound = ctx->h2d_copy_bound;
    if (const char *e = getenv("TKM_MSM_HOST_COPY_BOUND")) copy_bound = atoi(e) != 0;  // developer knob: force a layout
    if (n < ((size_t)1 << 19)) {
      pieces = 1;
      weight[0] = 1;
    } else if (copy_bound && n >= ((size_t)1 << 21)) {
      pieces = 4;
      weight[0] = 2; weight[1] = 3; weight[2] = 2; weight[3] = 1;
    } else {
      pieces = 2;
      weight[0] = 1; weight[1] = 3;
    }
  } else {
    for (uint32_t k = 0; k < pieces; k++) weight[k] = k == 0 ? 1 : (k == 1 ? 3 : 4);
  }
  {
    uint32_t wsum = 0, acc = 0;
    for (uint32_t k = 0; k < pieces; k++) wsum += weight[k];
    bound[0] = 0;
    for (uint32_t k = 0; k < pieces; k++) {
      acc += weight[k];
      bound[k + 1] = k + 1 == pieces ? n : (size_t)((unsigned __int128)n * acc / wsum);
      if (k == 0 && first_permille && pieces > 1) bound[1] = (size_t)((unsigned __int128)n * first_permille / 1000);
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
  TKM_CUDA(cudaEventRecord(ctx->copy_ev[32], ctx->stream));
  TKM_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[32], 0));
  TKM_CUDA(cudaEventRecord(ctx->copy_t[2], ctx->stream));
  TKM_CUDA(cudaEventRecord(ctx->copy_t[0], ctx->copy_stream));
  int32_t st = TKM_OK;
  for (uint32_t k = 0; k < pieces && st == TKM_OK; k++) {
    const size_t off = bound[k], cnt = bound[k + 1] - bound[k];
    // scalars first: the piece's digit decomposition, sort and tree offsets start while its bases are still in flight
    cudaError_t e = cudaMemcpyAsync(ds.p + off, scalars + off * 32, cnt * 32, cudaMemcpyHostToDevice, ctx->copy_stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->copy_ev[2 * k], ctx->copy_stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(db.p + off, bases + off * 96, cnt * 96, cudaMemcpyHostToDevice, ctx->copy_stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->copy_ev[2 * k + 1], ctx->copy_stream);
    if (e != cudaSuccess) st = fail(TKM_ERR_CUDA, "host-to-device copy of MSM piece %u failed: %s", k, cudaGetErrorString(e));
  }
  if (st == TKM_OK && cudaEventRecord(ctx->copy_t[1], ctx->copy_stream) != cudaSuccess) st = fail(TKM_ERR_CUDA, "cudaEventRecord failed");
  if (st == TKM_OK) {
    k_fill_identity<<<grid_for(set_stride * pieces, 256, ctx->sm_count), 256, 0, ctx->stream>>>(buckets.p, set_stride * pieces);
    st = launch_check(ctx, "k_fill_identity");
  }
  for (uint32_t k = 0; k < pieces && st == TKM_OK; k++) {
    const size_t off = bound[k], cnt = bound[k + 1] - bound[k];
    cudaError_t e = cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[2 * k], 0);
    if (e != cudaSuccess) {
      st = fail(TKM_ERR_CUDA, "cudaStreamWaitEvent failed: %s", cudaGetErrorString(e));
      break;
    }
    MsmInput in;
    in.scalars = ds.p + off;
    in.scalars_mont = false;
    in.scalar_row_stride = cnt;
    in.bases = db.p + off;
    in.base_row_stride = cnt;
    in.rows = 1;
    in.cols = cnt;
    in.idx = nullptr;
    in.bases_ready = ctx->copy_ev[2 * k + 1];  // waited for inside the pass, right before the first kernel that reads a base
    in.bases_canonical = true;                 // ... and converted to Montgomery form there
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
  st = msm_reduce(ctx, m, buckets.p, out96);  // synchronises the compute stream
  if (st == TKM_OK && cudaEventRecord(ctx->copy_t[3], ctx->stream) == cudaSuccess && cudaEventSynchronize(ctx->copy_t[3]) == cudaSuccess) {
    float t_copy = 0.f, t_call = 0.f;
    if (cudaEventElapsedTime(&t_copy, ctx->copy_t[0], ctx->copy_t[1]) == cudaSuccess &&
        cudaEventElapsedTime(&t_call, ctx->copy_t[2], ctx->copy_t[3]) == cudaSuccess && t_call > 0.f)
      ctx->h2d_share = t_copy / t_call;
  }
  return st;
}

}  // namespace tkm
