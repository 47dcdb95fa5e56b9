// Device-resident bivariate polynomial engine: the DensePolynomialExt operations of the reference
// (libs/src/bivariate_polynomial/mod.rs) without its host round-trips.  Coefficients live on the
// device in Montgomery form, row-major [x_size][y_size] (X = row index, Y contiguous;
// bivariate_polynomial/mod.rs:1756).  Every kernel here is HBM-bound (a few field ops per 32-byte
// element): 128-bit accesses, threads walk the contiguous Y axis.
#include <vector>

#include "common.cuh"

namespace tkm {

int32_t fill_powers_public(tkm_ctx *ctx, Fr *out, const Fr &base, const Fr &scale, size_t count);

// ---------------------------------------------------------------- shape management
// dst[(i+ox)*dy + (j+oy)] = src[i*sy + j], i < rows, j < cols   (resize / mul_monomial / clone)
__global__ void __launch_bounds__(256) k_copy_rect(Fr *__restrict__ dst, size_t dy, size_t ox, size_t oy, const Fr *__restrict__ src,
                                                   size_t sy, size_t rows, size_t cols) {
  size_t total = rows * cols;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
    size_t i = k / cols, j = k % cols;
    dst[(i + ox) * dy + (j + oy)] = src[i * sy + j];
  }
}

// find_degree (bivariate_polynomial/mod.rs:1480-1515): highest row / column holding a non-zero coefficient.
__global__ void __launch_bounds__(256) k_find_degree(const Fr *__restrict__ c, size_t x_size, size_t y_size, int *__restrict__ deg) {
  size_t total = x_size * y_size;
  int mx = -1, my = -1;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
    Fr v = c[k];
    if (!v.is_zero()) {
      int i = (int)(k / y_size), j = (int)(k % y_size);
      mx = max(mx, i);
      my = max(my, j);
    }
  }
  for (int d = 16; d > 0; d >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    my = max(my, __shfl_xor_sync(0xffffffffu, my, d));
  }
  if ((threadIdx.x & 31) == 0) {
    if (mx >= 0) atomicMax(deg, mx);
    if (my >= 0) atomicMax(deg + 1, my);
  }
}

// out = ca*a + cb*b on the union shape (operator impls, bivariate_polynomial/mod.rs:532-763, and poly_comb!).
__global__ void __launch_bounds__(256) k_axpby(Fr *__restrict__ out, size_t ox, size_t oy, const Fr *__restrict__ a, size_t ax, size_t ay,
                                               Fr ca, int ca_one, const Fr *__restrict__ b, size_t bx, size_t by, Fr cb, int cb_one) {
  size_t total = ox * oy;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
    size_t i = k / oy, j = k % oy;
    Fr r = Fr::zero();
    if (i < ax && j < ay) {
      Fr v = a[i * ay + j];
      r = ca_one ? v : v * ca;
    }
    if (b && i < bx && j < by) {
      Fr v = b[i * by + j];
      r = r + (cb_one ? v : v * cb);
    }
    out[k] = r;
  }
}

__global__ void k_add_scalar(Fr *c, Fr s) { c[0] = c[0] + s; }

// c_ij * px[i] * py[j]  (scale_coeffs_x / _y, bivariate_polynomial/mod.rs:1553-1613)
__global__ void __launch_bounds__(256) k_scale_coeffs(Fr *__restrict__ out, const Fr *__restrict__ in, size_t x_size, size_t y_size,
                                                      const Fr *__restrict__ px, const Fr *__restrict__ py) {
  size_t total = x_size * y_size;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
    size_t i = k / y_size, j = k % y_size;
    Fr v = in[k];
    if (px) { Fr s = px[i]; v = v * s; }
    if (py) { Fr s = py[j]; v = v * s; }
    out[k] = v;
  }
}

// ---------------------------------------------------------------- k-ary linear combination (poly_comb!, prove/src/lib.rs:30-38)
// out[i][j] = sum_t c_t * p_t[i - sx_t][j - sy_t] on the union shape: the whole chain of scalar products, monomial shifts
// (mul_monomial), clones, resizes and additions of a `poly_comb!` line in ONE pass -- every operand is read once, the
// result written once.
constexpr int LINCOMB_MAX = 16;
struct LincombArgs {
  const Fr *p[LINCOMB_MAX];
  uint32_t x[LINCOMB_MAX], y[LINCOMB_MAX], sx[LINCOMB_MAX], sy[LINCOMB_MAX];
  Fr c[LINCOMB_MAX];
  uint32_t one[LINCOMB_MAX];  // coefficient is 1: skip the product
  uint32_t k;
};
__global__ void __launch_bounds__(256) k_lincomb(Fr *__restrict__ out, size_t ox, size_t oy, const __grid_constant__ LincombArgs a) {
  const size_t total = ox * oy;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const uint32_t i = (uint32_t)(e / oy), j = (uint32_t)(e % oy);
    Fr acc = Fr::zero();
    for (uint32_t t = 0; t < a.k; t++) {
      const uint32_t ii = i - a.sx[t], jj = j - a.sy[t];  // wraps to a huge value when the shift exceeds the index
      if (ii < a.x[t] && jj < a.y[t]) {
        const uint4 *q = reinterpret_cast<const uint4 *>(a.p[t] + (size_t)ii * a.y[t] + jj);
        uint4 lo = __ldg(q), hi = __ldg(q + 1);
        Fr v;
        v.v[0] = lo.x; v.v[1] = lo.y; v.v[2] = lo.z; v.v[3] = lo.w;
        v.v[4] = hi.x; v.v[5] = hi.y; v.v[6] = hi.z; v.v[7] = hi.w;
        acc = acc + (a.one[t] ? v : v * a.c[t]);
      }
    }
    out[e] = acc;
  }
}

// ---------------------------------------------------------------- fused expression evaluation (PolyExpr, bivariate_polynomial/mod.rs:140-435)
// The pointwise part of evaluate_on_domain as ONE kernel: a postfix program over the leaves' evaluation vectors, run per
// element with the top of the stack in registers.  All threads execute the same program, so the interpreter's branches are
// uniform.  HBM traffic: every leaf read once, the result written once (the reference allocates a fresh 256 MiB vector
// per node, prove/src/lib.rs:2110-2146).
constexpr int PEX_MAX_LEAVES = 16, PEX_MAX_OPS = 96, PEX_MAX_CONSTS = 24, PEX_STACK = 8;
struct PexProgram {
  const Fr *leaf[PEX_MAX_LEAVES];
  Fr konst[PEX_MAX_CONSTS];
  uint32_t op[PEX_MAX_OPS];  // opcode | operand << 8
  uint32_t n_ops;
};
__global__ void __launch_bounds__(256) k_polyexpr(Fr *__restrict__ out, size_t x_size, size_t y_size, const __grid_constant__ PexProgram pr,
                                                  const Fr *__restrict__ tw, uint32_t log_stride, size_t half_m) {
  const size_t total = x_size * y_size;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    Fr st[PEX_STACK];
    Fr top = Fr::zero();
    int sp = 0;  // entries below the top
    for (uint32_t pc = 0; pc < pr.n_ops; pc++) {
      const uint32_t w = pr.op[pc], code = w & 0xffu, arg = w >> 8;
      switch (code) {
        case TKM_PEX_LEAF: {
          st[sp++] = top;
          const uint4 *q = reinterpret_cast<const uint4 *>(pr.leaf[arg] + e);
          uint4 lo = __ldg(q), hi = __ldg(q + 1);
          top.v[0] = lo.x; top.v[1] = lo.y; top.v[2] = lo.z; top.v[3] = lo.w;
          top.v[4] = hi.x; top.v[5] = hi.y; top.v[6] = hi.z; top.v[7] = hi.w;
          break;
        }
        case TKM_PEX_LEAF_SHIFT: {  // the leaf's table rotated by x_size/mx rows and y_size/my columns (see the header)
          st[sp++] = top;
          const uint32_t lx = (arg >> 4) & 63u, ly = (arg >> 10) & 63u;
          size_t i = e / y_size, j = e % y_size;
          if (lx) i = (i + x_size - (x_size >> (lx - 1))) & (x_size - 1);
          if (ly) j = (j + y_size - (y_size >> (ly - 1))) & (y_size - 1);
          const uint4 *q = reinterpret_cast<const uint4 *>(pr.leaf[arg & 15u] + i * y_size + j);
          uint4 lo = __ldg(q), hi = __ldg(q + 1);
          top.v[0] = lo.x; top.v[1] = lo.y; top.v[2] = lo.z; top.v[3] = lo.w;
          top.v[4] = hi.x; top.v[5] = hi.y; top.v[6] = hi.z; top.v[7] = hi.w;
          break;
        }
        case TKM_PEX_CONST:
          st[sp++] = top;
          top = pr.konst[arg];
          break;
        case TKM_PEX_ADD: top = st[--sp] + top; break;
        case TKM_PEX_SUB: top = st[--sp] - top; break;
        case TKM_PEX_MUL: top = st[--sp] * top; break;
        case TKM_PEX_SCALE: top = top * pr.konst[arg]; break;
        default: {  // TKM_PEX_XM1: times (omega_x^i - 1), omega_x^i from the domain table (second half: -omega^(k - M/2))
          const size_t ex = (e / y_size) << log_stride;
          Fr wv;
          if (x_size == 1) wv = Fr::one();
          else if (ex <= half_m) wv = tw[ex];
          else wv = tw[ex - half_m].neg();
          top = top * (wv - Fr::one());
          break;
        }
      }
    }
    out[e] = top;
  }
}

// ---------------------------------------------------------------- evaluation
__device__ __forceinline__ Fr warp_sum(Fr v) {
  for (int d = 16; d > 0; d >>= 1) {
    Fr o;
#pragma unroll
    for (int i = 0; i < 8; i++) o.v[i] = __shfl_xor_sync(0xffffffffu, v.v[i], d);
    v = v + o;
  }
  return v;
}
// out[i] = sum_j c[i][j] * w[j]: one warp per row (eval_y, bivariate_polynomial/mod.rs:1731-1740).
__global__ void __launch_bounds__(256) k_row_dot(Fr *__restrict__ out, const Fr *__restrict__ c, size_t rows, size_t cols,
                                                 const Fr *__restrict__ w) {
  size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  int lane = threadIdx.x & 31;
  for (size_t i = warp; i < rows; i += nwarps) {
    Fr acc = Fr::zero();
    for (size_t j = lane; j < cols; j += 32) {
      Fr v = c[i * cols + j], s = w[j];
      acc = acc + v * s;
    }
    acc = warp_sum(acc);
    if (lane == 0) out[i] = acc;
  }
}
// eval (bivariate_polynomial/mod.rs:1742-1750) in one pass over the coefficients: partial[block] = sum over the block's rows i
// of wx[i] * sum_j c[i][j] * wy[j] (one warp per row, eight rows in flight per block); k_sum_partials adds the partials.  No
// serial chain longer than a row's 1/32nd: the old form finished with ONE warp walking all x_size row values.
__global__ void __launch_bounds__(256) k_eval_partial(Fr *__restrict__ partial, const Fr *__restrict__ c, size_t rows, size_t cols,
                                                      const Fr *__restrict__ wy, const Fr *__restrict__ wx) {
  __shared__ Fr sh[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Fr wacc = Fr::zero();
  for (size_t i = blockIdx.x * (size_t)8 + warp; i < rows; i += (size_t)gridDim.x * 8) {
    Fr acc = Fr::zero();
    for (size_t j = lane; j < cols; j += 32) {
      Fr v = c[i * cols + j], s = wy[j];
      acc = acc + v * s;
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      Fr s = wx[i];
      wacc = wacc + acc * s;
    }
  }
  if (lane == 0) sh[warp] = wacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    Fr t = sh[0];
    for (int k = 1; k < 8; k++) t = t + sh[k];
    partial[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) k_sum_partials(Fr *__restrict__ out, const Fr *__restrict__ partial, size_t n) {
  __shared__ Fr sh[256];
  Fr acc = Fr::zero();
  for (size_t k = threadIdx.x; k < n; k += 256) acc = acc + partial[k];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int stride = 128; stride > 0; stride >>= 1) {
    if ((int)threadIdx.x < stride) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + stride];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}
// partial[chunk][j] = sum_{i in chunk} c[i][j] * w[i]   (eval_x, bivariate_polynomial/mod.rs:1719-1729)
__global__ void __launch_bounds__(128) k_col_dot_partial(Fr *__restrict__ partial, const Fr *__restrict__ c, size_t rows, size_t cols,
                                                         const Fr *__restrict__ w, size_t rows_per_chunk) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t chunk = blockIdx.y;
  if (j >= cols) return;
  size_t lo = chunk * rows_per_chunk, hi = lo + rows_per_chunk;
  if (hi > rows) hi = rows;
  Fr acc = Fr::zero();
  for (size_t i = lo; i < hi; i++) {
    Fr v = c[i * cols + j];
    if (w) { Fr s = w[i]; v = v * s; }
    acc = acc + v;
  }
  partial[chunk * cols + j] = acc;
}

// ---------------------------------------------------------------- division by vanishing polynomials
// div_by_vanishing_opt (bivariate_polynomial/mod.rs:2284-2410), restated as two chain kernels.
// Kernel 1: thread (lx < c, yy < d) folds the m X-blocks and walks its Y chain (stride d):
//   qy[lx][y] = qy[lx][y-d] - acc[lx][y] for y < y_size - d, else 0.
__global__ void __launch_bounds__(256) k_vanish_qy(Fr *__restrict__ qy, const Fr *__restrict__ p, size_t x_size, size_t y_size, size_t c,
                                                   size_t d) {
  size_t total = c * d;
  size_t m = x_size / c, n = y_size / d;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    size_t lx = t / d, yy = t % d;
    Fr prev = Fr::zero();
    for (size_t k = 0; k < n; k++) {
      size_t y = yy + k * d;
      Fr q = Fr::zero();
      if (k + 1 < n) {
        Fr acc = Fr::zero();
        for (size_t bx = 0; bx < m; bx++) {
          Fr v = p[(bx * c + lx) * y_size + y];
          acc = acc + v;
        }
        q = prev - acc;
        prev = q;
      }
      qy[lx * y_size + y] = q;
    }
  }
}
// Kernel 2: thread (lx < c, y) walks its X chain (stride c) over B = P + (Y^d - 1) qy:
//   qx[x][y] = qx[x-c][y] - B[x][y] for x < x_size - c, else 0.
__global__ void __launch_bounds__(256) k_vanish_qx(Fr *__restrict__ qx, const Fr *__restrict__ p, const Fr *__restrict__ qy, size_t x_size,
                                                   size_t y_size, size_t c, size_t d) {
  size_t total = c * y_size;
  size_t m = x_size / c;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    size_t lx = t / y_size, y = t % y_size;
    Fr prev = Fr::zero();
    for (size_t bx = 0; bx < m; bx++) {
      size_t x = bx * c + lx;
      Fr q = Fr::zero();
      if (bx + 1 < m) {
        Fr b = p[x * y_size + y];
        if (bx == 0 && y_size > d) {
          if (y < y_size - d) { Fr t1 = qy[lx * y_size + y]; b = b + t1; }
          if (y >= d) { Fr t2 = qy[lx * y_size + y - d]; b = b - t2; }
        }
        q = prev - b;
        prev = q;
      }
      qx[x * y_size + y] = q;
    }
  }
}

// ---------------------------------------------------------------- Ruffini division
// div_by_ruffini (bivariate_polynomial/mod.rs:2412-2477).  One thread per Y column runs the
// synthetic division along X (a Horner chain); remainders go to rx[y].
__global__ void __launch_bounds__(128) k_ruffini_x(Fr *__restrict__ qx, Fr *__restrict__ rx, const Fr *__restrict__ p, size_t x_size,
                                                   size_t y_size, Fr pt) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= y_size) return;
  if (x_size < 2) {
    rx[j] = p[j];
    qx[j] = Fr::zero();
    return;
  }
  Fr b = p[(x_size - 1) * y_size + j];
  qx[(x_size - 1) * y_size + j] = Fr::zero();
  qx[(x_size - 2) * y_size + j] = b;
  for (size_t i = x_size - 2; i >= 1; i--) {
    Fr v = p[i * y_size + j];
    b = v + b * pt;
    qx[(i - 1) * y_size + j] = b;
  }
  Fr v0 = p[j];
  rx[j] = v0 + b * pt;
}
// Segmented form for long X axes: the chain b_i = p_i + b_{i+1}*x is split into segments of `seg` rows.
//   pass 1 (thread per (segment, column)): local Horner value of the segment with zero incoming carry
//   pass 2 (thread per column): carry into every segment, top down: in_s = c_{s+1} + in_{s+1} * x^seg
//   pass 3 (thread per (segment, column)): re-run the segment with its carry and write the quotient rows
__global__ void __launch_bounds__(128) k_ruffini_seg_local(Fr *__restrict__ segc, const Fr *__restrict__ p, size_t x_size, size_t y_size,
                                                         size_t seg, Fr pt) {
  size_t nseg = x_size / seg;
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= nseg * y_size) return;
  size_t s = t / y_size, j = t % y_size;
  Fr c = Fr::zero();
  for (size_t k = seg; k-- > 0;) {
    Fr v = p[(s * seg + k) * y_size + j];
    c = v + c * pt;
  }
  segc[t] = c;
}
__global__ void __launch_bounds__(128) k_ruffini_seg_carry(Fr *__restrict__ carry, const Fr *__restrict__ segc, size_t nseg, size_t y_size,
                                                         Fr pt_pow_seg) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= y_size) return;
  Fr in = Fr::zero();
  carry[(nseg - 1) * y_size + j] = in;
  for (size_t s = nseg - 1; s-- > 0;) {
    Fr c = segc[(s + 1) * y_size + j];
    in = c + in * pt_pow_seg;
    carry[s * y_size + j] = in;
  }
}
// The carries of one column as a block-wide suffix scan over its segments (one block per column, one thread per segment,
// nseg <= 1024): log2(nseg) steps instead of nseg dependent products.
__global__ void __launch_bounds__(1024) k_ruffini_seg_carry_scan(Fr *__restrict__ carry, const Fr *__restrict__ segc, uint32_t nseg, size_t y_size,
                                                                Fr pt_pow_seg) {
  __shared__ Fr a[1024];
  const uint32_t s = threadIdx.x;
  const size_t j = blockIdx.x;
  if (s < nseg) a[s] = segc[(size_t)s * y_size + j];
  __syncthreads();
  Fr pw = pt_pow_seg;
  for (uint32_t d = 1; d < nseg; d <<= 1) {
    Fr t = Fr::zero();
    const bool on = s < nseg && s + d < nseg;
    if (on) t = a[s + d];
    __syncthreads();
    if (on) a[s] = a[s] + t * pw;
    __syncthreads();
    pw = pw * pw;
  }
  if (s < nseg) carry[(size_t)s * y_size + j] = (s + 1 < nseg) ? a[s + 1] : Fr::zero();
}
__global__ void __launch_bounds__(128) k_ruffini_seg_apply(Fr *__restrict__ qx, Fr *__restrict__ rx, const Fr *__restrict__ p,
                                                         const Fr *__restrict__ carry, size_t x_size, size_t y_size, size_t seg, Fr pt) {
  size_t nseg = x_size / seg;
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= nseg * y_size) return;
  size_t s = t / y_size, j = t % y_size;
  Fr b = carry[t];
  if (s == nseg - 1) qx[(x_size - 1) * y_size + j] = Fr::zero();
  for (size_t k = seg; k-- > 0;) {
    size_t i = s * seg + k;
    Fr v = p[i * y_size + j];
    b = v + b * pt;
    if (i >= 1)
      qx[(i - 1) * y_size + j] = b;
    else
      rx[j] = b;
  }
}
// Single chain along Y on the remainders: qy[0..y_size), r.
__global__ void k_ruffini_y(Fr *__restrict__ qy, Fr *__restrict__ r, const Fr *__restrict__ rx, size_t y_size, Fr pt) {
  if (threadIdx.x || blockIdx.x) return;
  if (y_size < 2) {
    qy[0] = Fr::zero();
    r[0] = rx[0];
    return;
  }
  Fr b = rx[y_size - 1];
  qy[y_size - 1] = Fr::zero();
  qy[y_size - 2] = b;
  for (size_t i = y_size - 2; i >= 1; i--) {
    Fr v = rx[i];
    b = v + b * pt;
    qy[i - 1] = b;
  }
  Fr v0 = rx[0];
  r[0] = v0 + b * pt;
}
// The same chain as a block-wide suffix scan for y_size <= 1024: b_i = r_i + y*b_{i+1} is the composition of affine maps with
// one common slope, so step s adds y^(2^s) * a[i + 2^s] to a[i]: log2(y_size) steps of one product each instead of y_size
// dependent ones (512 columns: ~10 us instead of 0.26 ms, four calls per prove).
__global__ void __launch_bounds__(1024) k_ruffini_y_scan(Fr *__restrict__ qy, Fr *__restrict__ r, const Fr *__restrict__ rx, uint32_t y_size, Fr pt) {
  __shared__ Fr a[1024];
  const uint32_t i = threadIdx.x;
  if (i < y_size) a[i] = rx[i];
  __syncthreads();
  Fr pw = pt;  // y^(2^s)
  for (uint32_t d = 1; d < y_size; d <<= 1) {
    Fr t = Fr::zero();
    const bool on = i < y_size && i + d < y_size;
    if (on) t = a[i + d];
    __syncthreads();
    if (on) a[i] = a[i] + t * pw;
    __syncthreads();
    pw = pw * pw;
  }
  if (i < y_size) {
    if (i == 0) r[0] = a[0];
    else qy[i - 1] = a[i];
    if (i == y_size - 1) qy[i] = Fr::zero();
  }
}


// ---------------------------------------------------------------- sparse R1CS x witness (read_R1CS_gen_uvwXY)
// One thread per (constraint row r, placement column c): the three dot products  sum_e coeff[e] * var[wire[e]]  over the
// sparse rows of the placement's subcircuit (eval_sparse_rows, libs/src/iotools/mod.rs:1589-1608), written straight into
// the [row][placement] evaluation tables of u, v, w (the reference's CPU loop + transpose, :1325-1365).  Canonical inputs
// are converted on the fly; outputs are in Montgomery form.
struct R1csView {
  const uint32_t *row_ptr;   // concatenated CSR row pointers of every (subcircuit, matrix)
  const uint32_t *wire;      // local wire index of every entry
  const Fr *coeff;           // entry coefficients, canonical
  const uint64_t *rp_base;   // [s_D * 3]: where (subcircuit, matrix)'s row pointers start in row_ptr
  const uint32_t *n_rows;    // [s_D]: constraints of each subcircuit
  const uint32_t *sub_of_col;  // [s_max]: subcircuit placed in each column (0xffffffff = empty)
  const uint64_t *var_off;   // [s_max]: where the column's variables start in `witness`
  const Fr *witness;         // all placement variables, canonical
};
__global__ void __launch_bounds__(128) k_r1cs_uvw(R1csView v, size_t n, size_t s_max, Fr *__restrict__ u, Fr *__restrict__ vv,
                                                  Fr *__restrict__ w) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= n * s_max) return;
  const size_t r = t / s_max, c = t % s_max;
  Fr acc[3] = {Fr::zero(), Fr::zero(), Fr::zero()};
  const uint32_t s = v.sub_of_col[c];
  if (s != 0xffffffffu && r < v.n_rows[s]) {
    const Fr *var = v.witness + v.var_off[c];
    for (int m = 0; m < 3; m++) {
      const uint32_t *rp = v.row_ptr + v.rp_base[3 * s + m];
      Fr a = Fr::zero();
      for (uint32_t e = rp[r]; e < rp[r + 1]; e++) a = a + v.coeff[e].to_mont() * var[v.wire[e]].to_mont();
      acc[m] = a;
    }
  }
  u[t] = acc[0];
  vv[t] = acc[1];
  w[t] = acc[2];
}


// ---------------------------------------------------------------- univariate long division along one axis (divide_x / divide_y)
// One thread per line of the sweep direction: schoolbook division of the line (length len, stride es) by the univariate
// denominator den[0..dd] (lead_inv = 1 / den[dd]).  rem holds a copy of the numerator on entry and the remainder on exit.
// (_divide_uni, libs/src/bivariate_polynomial/mod.rs:2052-2094: the reference slices every line into a DensePolynomial
// and calls ICICLE's divide on it; tests only, so clarity over speed.)
__global__ void __launch_bounds__(128) k_divide_uni(Fr *__restrict__ rem, Fr *__restrict__ quo, const Fr *__restrict__ den, size_t den_stride,
                                                    uint32_t dd, Fr lead_inv, size_t len, size_t lines, size_t es, size_t ls) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= lines) return;
  Fr *r = rem + t * ls;
  Fr *q = quo + t * ls;
  for (size_t i = len - dd; i-- > 0;) {
    const Fr c = r[(i + dd) * es];
    if (c.is_zero()) continue;
    const Fr f = c * lead_inv;
    q[i * es] = f;
    for (uint32_t k = 0; k < dd; k++) r[(i + k) * es] = r[(i + k) * es] - f * den[k * den_stride];
    r[(i + dd) * es] = Fr::zero();
  }
}

// ---------------------------------------------------------------- host orchestration
static int32_t poly_alloc(tkm_ctx *ctx, size_t x, size_t y, tkm_poly **out) {
  if (!is_pow2(x) || !is_pow2(y)) return fail(TKM_ERR_INVALID_ARGUMENT, "The input sizes must be powers of two (got %zu x %zu).", x, y);
  tkm_poly *p = new (std::nothrow) tkm_poly();
  if (!p) return fail(TKM_ERR_ALLOCATION, "out of host memory");
  cudaError_t e = cudaMallocAsync((void **)&p->d, x * y * sizeof(Fr), ctx->stream);
  if (e != cudaSuccess) {
    delete p;
    return fail(TKM_ERR_ALLOCATION, "cudaMallocAsync(%zu bytes) failed: %s", x * y * sizeof(Fr), cudaGetErrorString(e));
  }
  p->x_size = x;
  p->y_size = y;
  *out = p;
  return TKM_OK;
}
static void poly_release(tkm_ctx *ctx, tkm_poly *p) {
  if (!p) return;
  if (p->d) cudaFreeAsync(p->d, ctx->stream);
  delete p;
}

int32_t poly_find_degree(tkm_ctx *ctx, const tkm_poly *p, int64_t *xd, int64_t *yd) {
  Scratch<int> deg;
  TKM_TRY(deg.alloc(ctx, 2));
  TKM_CUDA(cudaMemsetAsync(deg.p, 0xff, 2 * sizeof(int), ctx->stream));
  size_t total = p->x_size * p->y_size;
  k_find_degree<<<grid_for(total, 256, ctx->sm_count), 256, 0, ctx->stream>>>(p->d, p->x_size, p->y_size, deg.p);
  TKM_TRY(launch_check(ctx, "k_find_degree"));
  int h[2];
  TKM_CUDA(cudaMemcpyAsync(h, deg.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  *xd = h[0];
  *yd = h[1];
  // the reference reports (-1,-1) for the zero polynomial (is_zero, bivariate_polynomial/mod.rs:134-137)
  if (h[0] < 0 || h[1] < 0) *xd = *yd = -1;
  return TKM_OK;
}

// New buffer of shape nx x ny holding src placed at offset (ox, oy), cropped to fit.
static int32_t poly_reshape_into(tkm_ctx *ctx, const Fr *src, size_t sx, size_t sy, size_t nx, size_t ny, size_t ox, size_t oy, Fr *dst) {
  TKM_CUDA(cudaMemsetAsync(dst, 0, nx * ny * sizeof(Fr), ctx->stream));
  if (ox >= nx || oy >= ny) return TKM_OK;
  size_t rows = sx < nx - ox ? sx : nx - ox, cols = sy < ny - oy ? sy : ny - oy;
  if (rows * cols == 0) return TKM_OK;
  k_copy_rect<<<grid_for(rows * cols, 256, ctx->sm_count), 256, 0, ctx->stream>>>(dst, ny, ox, oy, src, sy, rows, cols);
  return launch_check(ctx, "k_copy_rect");
}

int32_t poly_resize(tkm_ctx *ctx, tkm_poly *p, size_t tx, size_t ty) {
  if (tx == 0 || ty == 0) return fail(TKM_ERR_INVALID_ARGUMENT, "Invalid target sizes for resize");
  size_t nx = next_pow2(tx), ny = next_pow2(ty);
  if (nx == p->x_size && ny == p->y_size) return TKM_OK;
  Fr *nd = nullptr;
  cudaError_t e = cudaMallocAsync((void **)&nd, nx * ny * sizeof(Fr), ctx->stream);
  if (e != cudaSuccess) return fail(TKM_ERR_ALLOCATION, "cudaMallocAsync failed: %s", cudaGetErrorString(e));
  int32_t st = poly_reshape_into(ctx, p->d, p->x_size, p->y_size, nx, ny, 0, 0, nd);
  if (st != TKM_OK) {
    cudaFreeAsync(nd, ctx->stream);
    return st;
  }
  cudaFreeAsync(p->d, ctx->stream);
  p->d = nd;
  p->x_size = nx;
  p->y_size = ny;
  return TKM_OK;
}

// p[i][j] *= v[i] (by_row) or v[j]: the pointwise step of a product with a univariate factor.
__global__ void __launch_bounds__(256) k_axis_scale(Fr *__restrict__ p, const Fr *__restrict__ v, size_t x_size, size_t y_size, int by_row) {
  const size_t total = x_size * y_size;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const Fr s = v[by_row ? e / y_size : e % y_size];
    p[e] = p[e] * s;
  }
}

// _mul with a univariate factor (K0(X) * d(X, Y), t(X) * q(X, Y), L(X) * L(Y) in prove2/prove4): the product is a
// convolution along ONE axis, so only that axis is transformed -- the other axis' passes of all three transforms and the whole
// 2-D transform of the univariate factor are not needed.  uni: the univariate factor (degree 0 along the other axis), its
// coefficients along `x_axis ? X : Y`; other: any polynomial.  Same result as the general path (exact arithmetic).
static int32_t poly_mul_univariate(tkm_ctx *ctx, const tkm_poly *uni, int64_t udeg, bool x_axis, const tkm_poly *other, int64_t odx, int64_t ody,
                                   tkm_poly **out) {
  const size_t nx = next_pow2((size_t)(odx + (x_axis ? udeg : 0) + 1)), ny = next_pow2((size_t)(ody + (x_axis ? 0 : udeg) + 1));
  const size_t na = x_axis ? nx : ny;
  if (ctx->domain_log2 < 0) return fail(TKM_ERR_DOMAIN, "NTT domain is not initialized. Call tkm_ntt_domain_init first.");
  if (log2_exact(nx) + log2_exact(ny) > (uint32_t)ctx->domain_log2)  // the general path's condition, kept (bivariate_polynomial/mod.rs:1440-1445)
    return fail(TKM_ERR_DOMAIN, "NTT domain size too small: initialized size 2^%d but input size %zu", ctx->domain_log2, nx * ny);
  tkm_poly *res = nullptr;
  TKM_TRY(poly_alloc(ctx, nx, ny, &res));
  Scratch<Fr> vec;
  int32_t st = vec.alloc(ctx, na);
  if (st == TKM_OK) st = poly_reshape_into(ctx, other->d, other->x_size, other->y_size, nx, ny, 0, 0, res->d);
  // the factor's coefficients: column 0 of uni (x axis) or row 0 (y axis), zero-padded to the axis length
  if (st == TKM_OK) st = x_axis ? poly_reshape_into(ctx, uni->d, uni->x_size, uni->y_size, na, 1, 0, 0, vec.p)
                                : poly_reshape_into(ctx, uni->d, uni->x_size, uni->y_size, 1, na, 0, 0, vec.p);
  const size_t outer = x_axis ? 1 : nx, inner = x_axis ? ny : 1;
  if (st == TKM_OK) st = ntt_axis(ctx, res->d, res->d, outer, na, inner, TKM_FORWARD, nullptr);
  if (st == TKM_OK) st = ntt_axis(ctx, vec.p, vec.p, 1, na, 1, TKM_FORWARD, nullptr);
  if (st == TKM_OK) {
    k_axis_scale<<<grid_for(nx * ny, 256, ctx->sm_count), 256, 0, ctx->stream>>>(res->d, vec.p, nx, ny, x_axis ? 1 : 0);
    st = launch_check(ctx, "k_axis_scale");
  }
  if (st == TKM_OK) st = ntt_axis(ctx, res->d, res->d, outer, na, inner, TKM_INVERSE, nullptr);
  if (st != TKM_OK) {
    poly_release(ctx, res);
    return st;
  }
  *out = res;
  return TKM_OK;
}

int32_t poly_mul(tkm_ctx *ctx, const tkm_poly *a, const tkm_poly *b, tkm_poly **out) {
  int64_t adx, ady, bdx, bdy;
  TKM_TRY(poly_find_degree(ctx, a, &adx, &ady));
  TKM_TRY(poly_find_degree(ctx, b, &bdx, &bdy));
  const bool a_zero = adx < 0, b_zero = bdx < 0;
  // scalar fast paths (bivariate_polynomial/mod.rs:1867-1877); a zero operand gives the zero polynomial
  if (a_zero || b_zero) {
    TKM_TRY(poly_alloc(ctx, 1, 1, out));
    TKM_CUDA(cudaMemsetAsync((*out)->d, 0, sizeof(Fr), ctx->stream));
    return TKM_OK;
  }
  const bool a_const = adx + ady == 0, b_const = bdx + bdy == 0;
  if (a_const || b_const) {
    const tkm_poly *cpoly = a_const ? a : b, *other = a_const ? b : a;
    Fr s;
    TKM_CUDA(cudaMemcpyAsync(&s, cpoly->d, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    TKM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (a_const && b_const) {
      TKM_TRY(poly_alloc(ctx, 1, 1, out));
    } else {
      TKM_TRY(poly_alloc(ctx, other->x_size, other->y_size, out));
    }
    size_t cnt = (*out)->x_size * (*out)->y_size;
    return vec_scale(ctx, s, other->d, (*out)->d, cnt);
  }
  // a univariate factor: one-axis convolution (a or b; X- or Y-univariate)
  if (ady == 0) return poly_mul_univariate(ctx, a, adx, true, b, bdx, bdy, out);
  if (bdy == 0) return poly_mul_univariate(ctx, b, bdx, true, a, adx, ady, out);
  if (adx == 0) return poly_mul_univariate(ctx, a, ady, false, b, bdx, bdy, out);
  if (bdx == 0) return poly_mul_univariate(ctx, b, bdy, false, a, adx, ady, out);
  const size_t tx = (size_t)(adx + bdx + 1), ty = (size_t)(ady + bdy + 1);
  const size_t nx = next_pow2(tx), ny = next_pow2(ty);
  const size_t total = nx * ny;
  tkm_poly *res = nullptr;
  TKM_TRY(poly_alloc(ctx, nx, ny, &res));
  Scratch<Fr> rhs;
  int32_t st = rhs.alloc(ctx, total);
  if (st == TKM_OK) st = poly_reshape_into(ctx, a->d, a->x_size, a->y_size, nx, ny, 0, 0, res->d);
  if (st == TKM_OK) st = poly_reshape_into(ctx, b->d, b->x_size, b->y_size, nx, ny, 0, 0, rhs.p);
  if (st == TKM_OK) st = bintt_dev(ctx, res->d, res->d, nx, ny, TKM_FORWARD, nullptr, nullptr);
  if (st == TKM_OK) st = bintt_dev(ctx, rhs.p, rhs.p, nx, ny, TKM_FORWARD, nullptr, nullptr);
  if (st == TKM_OK) st = vec_op(ctx, TKM_OP_MUL, res->d, rhs.p, res->d, total);
  if (st == TKM_OK) st = bintt_dev(ctx, res->d, res->d, nx, ny, TKM_INVERSE, nullptr, nullptr);
  if (st != TKM_OK) {
    poly_release(ctx, res);
    return st;
  }
  *out = res;
  return TKM_OK;
}

// sum_i tmp[i] * w[i] style reductions built from the two dot kernels.
static int32_t row_dot(tkm_ctx *ctx, Fr *out, const Fr *c, size_t rows, size_t cols, const Fr *w) {
  size_t warps = rows;
  unsigned blocks = grid_for(warps * 32, 256, ctx->sm_count);
  k_row_dot<<<blocks, 256, 0, ctx->stream>>>(out, c, rows, cols, w);
  return launch_check(ctx, "k_row_dot");
}
static int32_t col_dot(tkm_ctx *ctx, Fr *out, const Fr *c, size_t rows, size_t cols, const Fr *w) {
  // two levels: chunks of rows -> partial[chunk][j] -> out[j]
  size_t rows_per_chunk = 64;
  size_t chunks = (rows + rows_per_chunk - 1) / rows_per_chunk;
  if (chunks == 1) {
    dim3 g((unsigned)((cols + 127) / 128), 1);
    k_col_dot_partial<<<g, 128, 0, ctx->stream>>>(out, c, rows, cols, w, rows);
    return launch_check(ctx, "k_col_dot_partial");
  }
  Scratch<Fr> partial;
  TKM_TRY(partial.alloc(ctx, chunks * cols));
  dim3 g((unsigned)((cols + 127) / 128), (unsigned)chunks);
  k_col_dot_partial<<<g, 128, 0, ctx->stream>>>(partial.p, c, rows, cols, w, rows_per_chunk);
  TKM_TRY(launch_check(ctx, "k_col_dot_partial"));
  dim3 g2((unsigned)((cols + 127) / 128), 1);
  k_col_dot_partial<<<g2, 128, 0, ctx->stream>>>(out, partial.p, chunks, cols, nullptr, chunks);
  return launch_check(ctx, "k_col_dot_partial");
}

}  // namespace tkm

using namespace tkm;

#define API_BEGIN                 \
  if (!ctx) return fail(TKM_ERR_INVALID_ARGUMENT, "null context"); \
  cudaSetDevice(ctx->device);

extern "C" {

int32_t tkm_poly_from_coeffs_host(tkm_ctx *ctx, const uint8_t *coeffs, size_t x_size, size_t y_size, tkm_poly **out) {
  API_BEGIN
  TKM_REQUIRE(coeffs && out, "null argument");
  TKM_TRY(poly_alloc(ctx, x_size, y_size, out));
  size_t n = x_size * y_size;
  int32_t st = TKM_OK;
  cudaError_t e = cudaMemcpyAsync((*out)->d, coeffs, n * 32, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) st = fail(TKM_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
  if (st == TKM_OK) st = vec_to_mont(ctx, (*out)->d, (*out)->d, n);
  if (st != TKM_OK) {
    poly_release(ctx, *out);
    *out = nullptr;
  }
  return st;
}

int32_t tkm_poly_from_evals_host(tkm_ctx *ctx, const uint8_t *evals, size_t x_size, size_t y_size, const uint8_t *cx, const uint8_t *cy,
                                 tkm_poly **out) {
  API_BEGIN
  TKM_TRY(tkm_poly_from_coeffs_host(ctx, evals, x_size, y_size, out));
  int32_t st = tkm_poly_ntt_inplace(ctx, *out, TKM_INVERSE, cx, cy);
  if (st != TKM_OK) {
    poly_release(ctx, *out);
    *out = nullptr;
  }
  return st;
}

int32_t tkm_poly_from_device(tkm_ctx *ctx, const void *dev_coeffs, size_t x_size, size_t y_size, tkm_poly **out) {
  API_BEGIN
  TKM_REQUIRE(dev_coeffs && out, "null argument");
  TKM_TRY(poly_alloc(ctx, x_size, y_size, out));
  cudaError_t e = cudaMemcpyAsync((*out)->d, dev_coeffs, x_size * y_size * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream);
  if (e != cudaSuccess) {
    poly_release(ctx, *out);
    *out = nullptr;
    return fail(TKM_ERR_CUDA, "D2D copy failed: %s", cudaGetErrorString(e));
  }
  return TKM_OK;
}

int32_t tkm_r1cs_uvw_polys(tkm_ctx *ctx, uint32_t s_D, const uint32_t *n_rows, const uint64_t *rp_base, const uint32_t *row_ptr,
                           size_t row_ptr_len, const uint32_t *wire, const uint8_t *coeff32, size_t nnz, const uint32_t *sub_of_col,
                           const uint64_t *var_off, const uint8_t *witness32, size_t n_vars, size_t n, size_t s_max, tkm_poly **out_u,
                           tkm_poly **out_v, tkm_poly **out_w) {
  API_BEGIN
  TKM_REQUIRE(n_rows && rp_base && row_ptr && sub_of_col && var_off && witness32 && out_u && out_v && out_w, "null argument");
  TKM_REQUIRE(nnz == 0 || (wire && coeff32), "null sparse entries");
  TKM_REQUIRE(is_pow2(n) && is_pow2(s_max), "n and s_max must be powers of two");
  // validate the host-side metadata so the kernel can not read out of bounds
  for (uint32_t s = 0; s < s_D; s++) {
    TKM_REQUIRE(n_rows[s] <= n, "n is smaller than the actual number of constraints.");
    for (int m = 0; m < 3; m++) TKM_REQUIRE(rp_base[3 * s + m] + n_rows[s] + 1 <= row_ptr_len, "row pointer table too short");
  }
  for (size_t i = 0; i < row_ptr_len; i++) TKM_REQUIRE(row_ptr[i] <= nnz, "row pointer exceeds the number of entries");
  // CSR rows must be non-decreasing inside every (subcircuit, matrix) block, and every wire an entry of subcircuit s names
  // must exist in each column that places s: wire_hi[s] = 1 + the largest local wire index of s (0 = no entries).
  std::vector<uint64_t> wire_hi(s_D, 0);
  for (uint32_t s = 0; s < s_D; s++) {
    for (int m = 0; m < 3; m++) {
      const uint32_t *rp = row_ptr + rp_base[3 * s + m];
      for (uint32_t r = 0; r < n_rows[s]; r++) TKM_REQUIRE(rp[r] <= rp[r + 1], "row pointers of subcircuit %u are not non-decreasing", s);
      for (uint32_t e = rp[0]; e < rp[n_rows[s]]; e++)
        if ((uint64_t)wire[e] + 1 > wire_hi[s]) wire_hi[s] = (uint64_t)wire[e] + 1;
    }
  }
  for (size_t c = 0; c < s_max; c++) {
    if (sub_of_col[c] == 0xffffffffu) continue;
    TKM_REQUIRE(sub_of_col[c] < s_D, "invalid placement column");
    TKM_REQUIRE(var_off[c] <= n_vars && wire_hi[sub_of_col[c]] <= n_vars - var_off[c],
                "placement column %zu: subcircuit %u reads wire %llu but only %llu variables follow its offset", c, sub_of_col[c],
                (unsigned long long)(wire_hi[sub_of_col[c]] - 1), (unsigned long long)(n_vars - var_off[c]));
  }
  Scratch<uint32_t> d_rp, d_wire, d_nrows, d_sub;
  Scratch<uint64_t> d_base, d_off;
  Scratch<Fr> d_coeff, d_wit;
  TKM_TRY(d_rp.alloc(ctx, row_ptr_len));
  TKM_TRY(d_wire.alloc(ctx, nnz));
  TKM_TRY(d_coeff.alloc(ctx, nnz));
  TKM_TRY(d_nrows.alloc(ctx, s_D));
  TKM_TRY(d_base.alloc(ctx, (size_t)s_D * 3));
  TKM_TRY(d_sub.alloc(ctx, s_max));
  TKM_TRY(d_off.alloc(ctx, s_max));
  TKM_TRY(d_wit.alloc(ctx, n_vars));
  TKM_CUDA(cudaMemcpyAsync(d_rp.p, row_ptr, row_ptr_len * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (nnz) {
    TKM_CUDA(cudaMemcpyAsync(d_wire.p, wire, nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
    TKM_CUDA(cudaMemcpyAsync(d_coeff.p, coeff32, nnz * 32, cudaMemcpyHostToDevice, ctx->stream));
  }
  TKM_CUDA(cudaMemcpyAsync(d_nrows.p, n_rows, (size_t)s_D * 4, cudaMemcpyHostToDevice, ctx->stream));
  TKM_CUDA(cudaMemcpyAsync(d_base.p, rp_base, (size_t)s_D * 3 * 8, cudaMemcpyHostToDevice, ctx->stream));
  TKM_CUDA(cudaMemcpyAsync(d_sub.p, sub_of_col, s_max * 4, cudaMemcpyHostToDevice, ctx->stream));
  TKM_CUDA(cudaMemcpyAsync(d_off.p, var_off, s_max * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (n_vars) TKM_CUDA(cudaMemcpyAsync(d_wit.p, witness32, n_vars * 32, cudaMemcpyHostToDevice, ctx->stream));
  tkm_poly *p[3] = {nullptr, nullptr, nullptr};
  int32_t st = TKM_OK;
  for (int i = 0; i < 3 && st == TKM_OK; i++) st = poly_alloc(ctx, n, s_max, &p[i]);
  if (st == TKM_OK) {
    R1csView v{d_rp.p, d_wire.p, d_coeff.p, d_base.p, d_nrows.p, d_sub.p, d_off.p, d_wit.p};
    const size_t total = n * s_max;
    k_r1cs_uvw<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(v, n, s_max, p[0]->d, p[1]->d, p[2]->d);
    st = launch_check(ctx, "k_r1cs_uvw");
  }
  for (int i = 0; i < 3 && st == TKM_OK; i++) st = tkm_poly_ntt_inplace(ctx, p[i], TKM_INVERSE, nullptr, nullptr);
  if (st != TKM_OK) {
    for (int i = 0; i < 3; i++) poly_release(ctx, p[i]);
    return st;
  }
  // the host staging buffers above are freed in stream order; the caller may reuse its arrays after this returns
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  *out_u = p[0];
  *out_v = p[1];
  *out_w = p[2];
  return TKM_OK;
}

int32_t tkm_poly_zero(tkm_ctx *ctx, size_t x_size, size_t y_size, tkm_poly **out) {
  API_BEGIN
  TKM_REQUIRE(out, "null argument");
  TKM_TRY(poly_alloc(ctx, x_size, y_size, out));
  TKM_CUDA(cudaMemsetAsync((*out)->d, 0, x_size * y_size * sizeof(Fr), ctx->stream));
  return TKM_OK;
}

int32_t tkm_poly_clone(tkm_ctx *ctx, const tkm_poly *p, tkm_poly **out) {
  API_BEGIN
  TKM_REQUIRE(p && out, "null argument");
  TKM_TRY(poly_alloc(ctx, p->x_size, p->y_size, out));
  TKM_CUDA(cudaMemcpyAsync((*out)->d, p->d, p->x_size * p->y_size * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
  return TKM_OK;
}

int32_t tkm_poly_free(tkm_ctx *ctx, tkm_poly *p) {
  API_BEGIN
  poly_release(ctx, p);
  return TKM_OK;
}

int32_t tkm_poly_shape(const tkm_poly *p, size_t *x_size, size_t *y_size) {
  if (!p) return fail(TKM_ERR_INVALID_ARGUMENT, "null polynomial");
  if (x_size) *x_size = p->x_size;
  if (y_size) *y_size = p->y_size;
  return TKM_OK;
}

int32_t tkm_poly_device_ptr(tkm_poly *p, void **out_dev) {
  if (!p || !out_dev) return fail(TKM_ERR_INVALID_ARGUMENT, "null argument");
  *out_dev = p->d;
  return TKM_OK;
}

int32_t tkm_poly_copy_coeffs_host(tkm_ctx *ctx, const tkm_poly *p, uint8_t *out) {
  API_BEGIN
  TKM_REQUIRE(p && out, "null argument");
  size_t n = p->x_size * p->y_size;
  Scratch<Fr> tmp;
  TKM_TRY(tmp.alloc(ctx, n));
  TKM_TRY(vec_from_mont(ctx, p->d, tmp.p, n));
  TKM_CUDA(cudaMemcpyAsync(out, tmp.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  return TKM_OK;
}

int32_t tkm_poly_to_evals_host(tkm_ctx *ctx, const tkm_poly *p, const uint8_t *cx, const uint8_t *cy, uint8_t *out) {
  API_BEGIN
  TKM_REQUIRE(p && out, "null argument");
  size_t n = p->x_size * p->y_size;
  Scratch<Fr> tmp;
  TKM_TRY(tmp.alloc(ctx, n));
  Fr gx, gy;
  if (cx) gx = fr_from_bytes_host(cx);
  if (cy) gy = fr_from_bytes_host(cy);
  TKM_TRY(bintt_dev(ctx, p->d, tmp.p, p->x_size, p->y_size, TKM_FORWARD, cx ? &gx : nullptr, cy ? &gy : nullptr));
  TKM_TRY(vec_from_mont(ctx, tmp.p, tmp.p, n));
  TKM_CUDA(cudaMemcpyAsync(out, tmp.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  return TKM_OK;
}

int32_t tkm_poly_ntt_inplace(tkm_ctx *ctx, tkm_poly *p, int32_t dir, const uint8_t *cx, const uint8_t *cy) {
  API_BEGIN
  TKM_REQUIRE(p, "null polynomial");
  Fr gx, gy;
  if (cx) gx = fr_from_bytes_host(cx);
  if (cy) gy = fr_from_bytes_host(cy);
  return bintt_dev(ctx, p->d, p->d, p->x_size, p->y_size, dir, cx ? &gx : nullptr, cy ? &gy : nullptr);
}

int32_t tkm_poly_find_degree(tkm_ctx *ctx, const tkm_poly *p, int64_t *xd, int64_t *yd) {
  API_BEGIN
  TKM_REQUIRE(p && xd && yd, "null argument");
  return poly_find_degree(ctx, p, xd, yd);
}

int32_t tkm_poly_resize(tkm_ctx *ctx, tkm_poly *p, size_t tx, size_t ty) {
  API_BEGIN
  TKM_REQUIRE(p, "null polynomial");
  return poly_resize(ctx, p, tx, ty);
}

int32_t tkm_poly_optimize_size(tkm_ctx *ctx, tkm_poly *p) {
  API_BEGIN
  TKM_REQUIRE(p, "null polynomial");
  int64_t xd, yd;
  TKM_TRY(poly_find_degree(ctx, p, &xd, &yd));
  if (xd < 0 || yd < 0) return TKM_OK;  // zero polynomial keeps its shape (bivariate_polynomial/mod.rs:1814-1816)
  return poly_resize(ctx, p, (size_t)xd + 1, (size_t)yd + 1);
}

int32_t tkm_poly_mul_monomial(tkm_ctx *ctx, const tkm_poly *p, size_t ex, size_t ey, tkm_poly **out) {
  API_BEGIN
  TKM_REQUIRE(p && out, "null argument");
  size_t nx = next_pow2(p->x_size + ex), ny = next_pow2(p->y_size + ey);
  TKM_TRY(poly_alloc(ctx, nx, ny, out));
  int32_t st = poly_reshape_into(ctx, p->d, p->x_size, p->y_size, nx, ny, ex, ey, (*out)->d);
  if (st != TKM_OK) {
    poly_release(ctx, *out);
    *out = nullptr;
  }
  return st;
}

int32_t tkm_poly_axpby(tkm_ctx *ctx, const tkm_poly *a, const uint8_t *ca32, const tkm_poly *b, const uint8_t *cb32, tkm_poly **out) {
  API_BEGIN
  TKM_REQUIRE(a && out, "null argument");
  size_t ox = a->x_size, oy = a->y_size;
  if (b) {
    if (b->x_size > ox) ox = b->x_size;
    if (b->y_size > oy) oy = b->y_size;
  }
  Fr ca = Fr::one(), cb = Fr::one();
  if (ca32) ca = fr_from_bytes_host(ca32);
  if (cb32) cb = fr_from_bytes_host(cb32);
  TKM_TRY(poly_alloc(ctx, ox, oy, out));
  k_axpby<<<grid_for(ox * oy, 256, ctx->sm_count), 256, 0, ctx->stream>>>((*out)->d, ox, oy, a->d, a->x_size, a->y_size, ca,
                                                                          ca32 ? 0 : 1, b ? b->d : nullptr, b ? b->x_size : 0,
                                                                          b ? b->y_size : 0, cb, cb32 ? 0 : 1);
  return launch_check(ctx, "k_axpby");
}

int32_t tkm_poly_lincomb(tkm_ctx *ctx, uint32_t k, const tkm_poly *const *polys, const uint8_t *coeffs32, const uint32_t *shift_x,
                         const uint32_t *shift_y, tkm_poly **out) {
  API_BEGIN
  TKM_REQUIRE(polys && out && k >= 1, "null argument or empty combination");
  TKM_REQUIRE(k <= (uint32_t)LINCOMB_MAX, "at most %d terms per combination (got %u)", LINCOMB_MAX, k);
  static const uint8_t ONE32[32] = {1};
  LincombArgs a;
  memset(&a, 0, sizeof a);
  a.k = k;
  size_t ox = 1, oy = 1;
  for (uint32_t t = 0; t < k; t++) {
    TKM_REQUIRE(polys[t], "null polynomial in term %u", t);
    const size_t sx = shift_x ? shift_x[t] : 0, sy = shift_y ? shift_y[t] : 0;
    // a shifted term has mul_monomial's shape (:1820-1844): the next powers of two of (size + shift)
    const size_t tx = (sx || sy) ? next_pow2(polys[t]->x_size + sx) : polys[t]->x_size;
    const size_t ty = (sx || sy) ? next_pow2(polys[t]->y_size + sy) : polys[t]->y_size;
    if (tx > ox) ox = tx;
    if (ty > oy) oy = ty;
    a.p[t] = polys[t]->d;
    a.x[t] = (uint32_t)polys[t]->x_size;
    a.y[t] = (uint32_t)polys[t]->y_size;
    a.sx[t] = (uint32_t)sx;
    a.sy[t] = (uint32_t)sy;
    const uint8_t *c = coeffs32 ? coeffs32 + 32 * (size_t)t : ONE32;
    a.one[t] = memcmp(c, ONE32, 32) == 0;
    a.c[t] = fr_from_bytes_host(c);
  }
  TKM_REQUIRE(ox < ((size_t)1 << 31) && oy < ((size_t)1 << 31), "combination shape out of range");
  TKM_TRY(poly_alloc(ctx, ox, oy, out));
  k_lincomb<<<grid_for(ox * oy, 256, ctx->sm_count), 256, 0, ctx->stream>>>((*out)->d, ox, oy, a);
  return launch_check(ctx, "k_lincomb");
}

int32_t tkm_polyexpr_eval(tkm_ctx *ctx, const tkm_poly *const *leaves, uint32_t n_leaves, const uint32_t *program, uint32_t n_ops,
                          const uint8_t *consts32, uint32_t n_consts, size_t target_x, size_t target_y, tkm_poly **out) {
  API_BEGIN
  TKM_REQUIRE(program && out && n_ops >= 1, "null argument or empty program");
  TKM_REQUIRE(n_leaves == 0 || leaves, "null leaves");
  TKM_REQUIRE(n_consts == 0 || consts32, "null constants");
  TKM_REQUIRE(n_leaves <= (uint32_t)PEX_MAX_LEAVES && n_ops <= (uint32_t)PEX_MAX_OPS && n_consts <= (uint32_t)PEX_MAX_CONSTS,
              "expression too large (limits: %d leaves, %d ops, %d constants)", PEX_MAX_LEAVES, PEX_MAX_OPS, PEX_MAX_CONSTS);
  if (!is_pow2(target_x) || !is_pow2(target_y)) return fail(TKM_ERR_INVALID_ARGUMENT, "Fused polynomial expression domains must be powers of two.");
  if (ctx->domain_log2 < 0) return fail(TKM_ERR_DOMAIN, "NTT domain is not initialized. Call tkm_ntt_domain_init first.");
  if (log2_exact(target_x) + log2_exact(target_y) > (uint32_t)ctx->domain_log2)
    return fail(TKM_ERR_DOMAIN, "NTT domain size too small: initialized size 2^%d but the expression domain is %zu x %zu", ctx->domain_log2, target_x, target_y);
  // validate the program: operands in range, the stack never underflows or overflows, exactly one value is left
  int depth = 0;
  for (uint32_t pc = 0; pc < n_ops; pc++) {
    const uint32_t code = program[pc] & 0xffu, arg = program[pc] >> 8;
    switch (code) {
      case TKM_PEX_LEAF: TKM_REQUIRE(arg < n_leaves && leaves[arg], "op %u: leaf %u out of range", pc, arg); depth++; break;
      case TKM_PEX_CONST: TKM_REQUIRE(arg < n_consts, "op %u: constant %u out of range", pc, arg); depth++; break;
      case TKM_PEX_ADD: case TKM_PEX_SUB: case TKM_PEX_MUL: TKM_REQUIRE(depth >= 2, "op %u: stack underflow", pc); depth--; break;
      case TKM_PEX_SCALE: TKM_REQUIRE(arg < n_consts && depth >= 1, "op %u: bad scale", pc); break;
      case TKM_PEX_XM1: TKM_REQUIRE(depth >= 1, "op %u: stack underflow", pc); break;
      case TKM_PEX_LEAF_SHIFT: {
        const uint32_t l = arg & 15u, lx = (arg >> 4) & 63u, ly = (arg >> 10) & 63u;
        TKM_REQUIRE((arg >> 16) == 0 && l < n_leaves && leaves[l], "op %u: leaf %u out of range", pc, l);
        TKM_REQUIRE((lx == 0 || ((size_t)1 << (lx - 1)) <= target_x) && (ly == 0 || ((size_t)1 << (ly - 1)) <= target_y),
                    "op %u: the root's order must divide the domain's extent", pc);
        depth++;
        break;
      }
      default: return fail(TKM_ERR_INVALID_ARGUMENT, "op %u: unknown opcode %u", pc, code);
    }
    TKM_REQUIRE(depth <= PEX_STACK, "op %u: expression needs more than %d stack entries", pc, PEX_STACK);
  }
  TKM_REQUIRE(depth == 1, "the program leaves %d values on the stack (expected 1)", depth);
  const size_t n = target_x * target_y;
  PexProgram pr;
  memset(&pr, 0, sizeof pr);
  pr.n_ops = n_ops;
  memcpy(pr.op, program, n_ops * sizeof(uint32_t));
  for (uint32_t c = 0; c < n_consts; c++) pr.konst[c] = fr_from_bytes_host(consts32 + 32 * (size_t)c);
  // one forward transform per distinct leaf (eval_poly_leaf, :459-502): zero-pad to the domain, NTT in place
  Scratch<Fr> ev[PEX_MAX_LEAVES];
  for (uint32_t l = 0; l < n_leaves; l++) {
    const tkm_poly *p = leaves[l];
    if (p->x_size > target_x || p->y_size > target_y)
      return fail(TKM_ERR_INVALID_ARGUMENT, "Fused polynomial expression domain is too small for the expression degree.");
    TKM_TRY(ev[l].alloc(ctx, n));
    TKM_TRY(poly_reshape_into(ctx, p->d, p->x_size, p->y_size, target_x, target_y, 0, 0, ev[l].p));
    TKM_TRY(bintt_dev(ctx, ev[l].p, ev[l].p, target_x, target_y, TKM_FORWARD, nullptr, nullptr));
    pr.leaf[l] = ev[l].p;
  }
  tkm_poly *res = nullptr;
  TKM_TRY(poly_alloc(ctx, target_x, target_y, &res));
  const uint32_t log_stride = (uint32_t)ctx->domain_log2 - log2_exact(target_x);
  TKM_CUDA(cudaEventRecord(ctx->pev0, ctx->stream));
  k_polyexpr<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(res->d, target_x, target_y, pr, ctx->twiddles, log_stride,
                                                                       (size_t)1 << (ctx->domain_log2 > 0 ? ctx->domain_log2 - 1 : 0));
  TKM_CUDA(cudaEventRecord(ctx->pev1, ctx->stream));
  ctx->poly_kernel_timed = true;
  int32_t st = launch_check(ctx, "k_polyexpr");
  if (st == TKM_OK) st = bintt_dev(ctx, res->d, res->d, target_x, target_y, TKM_INVERSE, nullptr, nullptr);
  if (st != TKM_OK) {
    poly_release(ctx, res);
    return st;
  }
  *out = res;
  return TKM_OK;
}

int32_t tkm_poly_add_scalar(tkm_ctx *ctx, tkm_poly *p, const uint8_t s32[32]) {
  API_BEGIN
  TKM_REQUIRE(p && s32, "null argument");
  k_add_scalar<<<1, 1, 0, ctx->stream>>>(p->d, fr_from_bytes_host(s32));
  return launch_check(ctx, "k_add_scalar");
}

int32_t tkm_poly_mul(tkm_ctx *ctx, const tkm_poly *a, const tkm_poly *b, tkm_poly **out) {
  API_BEGIN
  TKM_REQUIRE(a && b && out, "null argument");
  return poly_mul(ctx, a, b, out);
}

int32_t tkm_poly_scale_coeffs(tkm_ctx *ctx, const tkm_poly *p, const uint8_t *sx32, const uint8_t *sy32, tkm_poly **out) {
  API_BEGIN
  TKM_REQUIRE(p && out, "null argument");
  Scratch<Fr> px, py;
  if (sx32) {
    TKM_TRY(px.alloc(ctx, p->x_size));
    TKM_TRY(fill_powers_public(ctx, px.p, fr_from_bytes_host(sx32), Fr::one(), p->x_size));
  }
  if (sy32) {
    TKM_TRY(py.alloc(ctx, p->y_size));
    TKM_TRY(fill_powers_public(ctx, py.p, fr_from_bytes_host(sy32), Fr::one(), p->y_size));
  }
  TKM_TRY(poly_alloc(ctx, p->x_size, p->y_size, out));
  size_t n = p->x_size * p->y_size;
  k_scale_coeffs<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>((*out)->d, p->d, p->x_size, p->y_size, sx32 ? px.p : nullptr,
                                                                          sy32 ? py.p : nullptr);
  return launch_check(ctx, "k_scale_coeffs");
}

int32_t tkm_poly_eval_y(tkm_ctx *ctx, const tkm_poly *p, const uint8_t y32[32], tkm_poly **out) {
  API_BEGIN
  TKM_REQUIRE(p && y32 && out, "null argument");
  Scratch<Fr> py;
  TKM_TRY(py.alloc(ctx, p->y_size));
  TKM_TRY(fill_powers_public(ctx, py.p, fr_from_bytes_host(y32), Fr::one(), p->y_size));
  TKM_TRY(poly_alloc(ctx, p->x_size, 1, out));
  return row_dot(ctx, (*out)->d, p->d, p->x_size, p->y_size, py.p);
}

int32_t tkm_poly_eval_x(tkm_ctx *ctx, const tkm_poly *p, const uint8_t x32[32], tkm_poly **out) {
  API_BEGIN
  TKM_REQUIRE(p && x32 && out, "null argument");
  Scratch<Fr> px;
  TKM_TRY(px.alloc(ctx, p->x_size));
  TKM_TRY(fill_powers_public(ctx, px.p, fr_from_bytes_host(x32), Fr::one(), p->x_size));
  TKM_TRY(poly_alloc(ctx, 1, p->y_size, out));
  return col_dot(ctx, (*out)->d, p->d, p->x_size, p->y_size, px.p);
}

int32_t tkm_poly_eval(tkm_ctx *ctx, const tkm_poly *p, const uint8_t x32[32], const uint8_t y32[32], uint8_t out32[32]) {
  API_BEGIN
  TKM_REQUIRE(p && x32 && y32 && out32, "null argument");
  Scratch<Fr> px, py, partial, res;
  TKM_TRY(px.alloc(ctx, p->x_size));
  TKM_TRY(py.alloc(ctx, p->y_size));
  TKM_TRY(res.alloc(ctx, 1));
  TKM_TRY(fill_powers_public(ctx, px.p, fr_from_bytes_host(x32), Fr::one(), p->x_size));
  TKM_TRY(fill_powers_public(ctx, py.p, fr_from_bytes_host(y32), Fr::one(), p->y_size));
  size_t nblk = (p->x_size + 7) / 8;
  const size_t cap = (size_t)ctx->sm_count * 8;
  if (nblk > cap) nblk = cap;
  TKM_TRY(partial.alloc(ctx, nblk));
  k_eval_partial<<<(unsigned)nblk, 256, 0, ctx->stream>>>(partial.p, p->d, p->x_size, p->y_size, py.p, px.p);
  TKM_TRY(launch_check(ctx, "k_eval_partial"));
  k_sum_partials<<<1, 256, 0, ctx->stream>>>(res.p, partial.p, nblk);
  TKM_TRY(launch_check(ctx, "k_sum_partials"));
  Fr h;
  TKM_CUDA(cudaMemcpyAsync(&h, res.p, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  fr_to_bytes_host(h, out32);
  return TKM_OK;
}

int32_t tkm_poly_div_by_vanishing(tkm_ctx *ctx, tkm_poly *p, size_t c, size_t d, tkm_poly **out_qx, tkm_poly **out_qy) {
  API_BEGIN
  TKM_REQUIRE(p && out_qx && out_qy, "null argument");
  if (!is_pow2(c) || !is_pow2(d)) return fail(TKM_ERR_INVALID_ARGUMENT, "The denominators must have degress as powers of two.");
  int64_t xd, yd;
  TKM_TRY(poly_find_degree(ctx, p, &xd, &yd));
  if (xd >= 0 && yd >= 0) TKM_TRY(poly_resize(ctx, p, (size_t)xd + 1, (size_t)yd + 1));  // optimize_size (:2290)
  if (xd < (int64_t)c || yd < (int64_t)d) return fail(TKM_ERR_INVALID_ARGUMENT, "The numerator must have grater degrees than denominators.");
  const size_t x = p->x_size, y = p->y_size;
  tkm_poly *qx = nullptr, *qy = nullptr;
  TKM_TRY(poly_alloc(ctx, x, y, &qx));
  int32_t st = poly_alloc(ctx, c, y, &qy);
  if (st == TKM_OK) {
    k_vanish_qy<<<grid_for(c * d, 256, ctx->sm_count), 256, 0, ctx->stream>>>(qy->d, p->d, x, y, c, d);
    st = launch_check(ctx, "k_vanish_qy");
  }
  if (st == TKM_OK) {
    k_vanish_qx<<<grid_for(c * y, 256, ctx->sm_count), 256, 0, ctx->stream>>>(qx->d, p->d, qy->d, x, y, c, d);
    st = launch_check(ctx, "k_vanish_qx");
  }
  if (st != TKM_OK) {
    poly_release(ctx, qx);
    poly_release(ctx, qy);
    return st;
  }
  *out_qx = qx;
  *out_qy = qy;
  return TKM_OK;
}

int32_t tkm_poly_div_by_ruffini(tkm_ctx *ctx, const tkm_poly *p, const uint8_t x32[32], const uint8_t y32[32], tkm_poly **out_qx,
                                tkm_poly **out_qy, uint8_t out_r32[32]) {
  API_BEGIN
  TKM_REQUIRE(p && x32 && y32 && out_qx && out_qy && out_r32, "null argument");
  const size_t x = p->x_size, y = p->y_size;
  tkm_poly *qx = nullptr, *qy = nullptr;
  Scratch<Fr> rx, r;
  TKM_TRY(rx.alloc(ctx, y));
  TKM_TRY(r.alloc(ctx, 1));
  TKM_TRY(poly_alloc(ctx, x, y, &qx));
  int32_t st = poly_alloc(ctx, 1, y, &qy);
  const Fr ptx = fr_from_bytes_host(x32);
  const size_t seg = 64;
  if (st == TKM_OK && x >= 4 * seg) {
    const size_t nseg = x / seg;
    Scratch<Fr> segc, carry;
    st = segc.alloc(ctx, nseg * y);
    if (st == TKM_OK) st = carry.alloc(ctx, nseg * y);
    if (st == TKM_OK) {
      k_ruffini_seg_local<<<(unsigned)((nseg * y + 127) / 128), 128, 0, ctx->stream>>>(segc.p, p->d, x, y, seg, ptx);
      st = launch_check(ctx, "k_ruffini_seg_local");
    }
    if (st == TKM_OK) {
      if (nseg <= 1024 && y <= 0x7fffffffull) {
        unsigned bt = 32;
        while (bt < nseg) bt <<= 1;
        k_ruffini_seg_carry_scan<<<(unsigned)y, bt, 0, ctx->stream>>>(carry.p, segc.p, (uint32_t)nseg, y, ptx.pow_u64(seg));
      } else {
        k_ruffini_seg_carry<<<(unsigned)((y + 127) / 128), 128, 0, ctx->stream>>>(carry.p, segc.p, nseg, y, ptx.pow_u64(seg));
      }
      st = launch_check(ctx, "k_ruffini_seg_carry");
    }
    if (st == TKM_OK) {
      k_ruffini_seg_apply<<<(unsigned)((nseg * y + 127) / 128), 128, 0, ctx->stream>>>(qx->d, rx.p, p->d, carry.p, x, y, seg, ptx);
      st = launch_check(ctx, "k_ruffini_seg_apply");
    }
  } else if (st == TKM_OK) {
    k_ruffini_x<<<(unsigned)((y + 127) / 128), 128, 0, ctx->stream>>>(qx->d, rx.p, p->d, x, y, ptx);
    st = launch_check(ctx, "k_ruffini_x");
  }
  if (st == TKM_OK) {
    if (y >= 2 && y <= 1024) k_ruffini_y_scan<<<1, 1024, 0, ctx->stream>>>(qy->d, r.p, rx.p, (uint32_t)y, fr_from_bytes_host(y32));
    else k_ruffini_y<<<1, 32, 0, ctx->stream>>>(qy->d, r.p, rx.p, y, fr_from_bytes_host(y32));
    st = launch_check(ctx, "k_ruffini_y");
  }
  Fr h;
  if (st == TKM_OK) {
    cudaError_t e = cudaMemcpyAsync(&h, r.p, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) st = fail(TKM_ERR_CUDA, "D2H copy failed: %s", cudaGetErrorString(e));
  }
  if (st != TKM_OK) {
    poly_release(ctx, qx);
    poly_release(ctx, qy);
    return st;
  }
  fr_to_bytes_host(h, out_r32);
  *out_qx = qx;
  *out_qy = qy;
  return TKM_OK;
}

static int32_t commit_input(tkm_ctx *ctx, tkm_poly *p, const tkm_crs *crs, MsmInput *in, bool *zero) {
  int64_t xd, yd;
  TKM_TRY(poly_find_degree(ctx, p, &xd, &yd));
  *zero = xd < 0 || yd < 0;  // the zero polynomial commits to the identity (iotools/mod.rs:2057-2059)
  if (*zero) return TKM_OK;
  const size_t tx = (size_t)xd + 1, ty = (size_t)yd + 1;
  if (tx > crs->rows || ty > crs->cols) return fail(TKM_ERR_INVALID_ARGUMENT, "Insufficient length of sigma.sigma_1.xy_powers");
  in->scalars = p->d;
  in->scalars_mont = true;
  in->scalar_row_stride = p->y_size;
  in->bases = crs->pre ? crs->pre : crs->d;
  in->base_row_stride = crs->cols;
  in->rows = tx;
  in->cols = ty;
  in->idx = nullptr;
  if (crs->pre) {  // fixed-base tables (tkm_crs_precompute): one shared bucket set, no Horner tail
    in->pre_c = crs->pre_c;
    in->pre_stride = (uint32_t)(crs->rows * crs->cols);
  }
  // the CRS's x-only table for the pair tree's forward pass: built once per CRS (and again if tables appear or vanish)
  tkm_crs *mc = const_cast<tkm_crs *>(crs);
  const bool want_pre = crs->pre != nullptr;
  if (tx * ty >= ((size_t)1 << 20)) {  // only commitments large enough to run the tree need it
    if (mc->xpad && mc->xpad_for_pre != want_pre) {
      TKM_CUDA(cudaStreamSynchronize(ctx->stream));
      cudaFree(mc->xpad);
      mc->xpad = nullptr;
    }
    if (!mc->xpad) {
      const size_t n_all = crs->rows * crs->cols, tables = want_pre ? crs->pre_W : 1;
      if (cudaMalloc((void **)&mc->xpad, n_all * tables * 64) == cudaSuccess) {
        int32_t st = msm_build_xpad(ctx, in->bases, crs->cols, crs->rows, crs->cols, (uint32_t)tables, n_all, mc->xpad);
        if (st != TKM_OK) return st;
        mc->xpad_for_pre = want_pre;
      } else {
        cudaGetLastError();  // not enough memory for the optional table: the tree falls back to gathering x from the points
        mc->xpad = nullptr;
      }
    }
    in->xpad = mc->xpad;
  }
  return TKM_OK;
}

int32_t tkm_poly_divide_uni(tkm_ctx *ctx, const tkm_poly *p, const tkm_poly *denom, int32_t y_dir, tkm_poly **out_q, tkm_poly **out_r) {
  API_BEGIN
  TKM_REQUIRE(p && denom && out_q && out_r, "null argument");
  int64_t nx, ny, dx, dy;
  TKM_TRY(poly_find_degree(ctx, p, &nx, &ny));
  TKM_TRY(poly_find_degree(ctx, denom, &dx, &dy));
  const char *axis = y_dir ? "divide_y" : "divide_x";
  if (dx < 0) return fail(TKM_ERR_INVALID_ARGUMENT, "Divide by zero");
  if (y_dir ? dx != 0 : dy != 0) return fail(TKM_ERR_INVALID_ARGUMENT, "Denominator for %s must be %s-univariate", axis, y_dir ? "Y" : "X");
  const int64_t nd = y_dir ? ny : nx, dd = y_dir ? dy : dx;
  if (nd < dd) return fail(TKM_ERR_INVALID_ARGUMENT, "Numer.degree < Denom.degree for %s", axis);
  // leading coefficient of the denominator and its inverse (host side, one element)
  const size_t den_stride = y_dir ? 1 : denom->y_size;
  Fr lead;
  TKM_CUDA(cudaMemcpyAsync(&lead, denom->d + (size_t)dd * den_stride, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  const Fr lead_inv = lead.inv();
  tkm_poly *q = nullptr, *r = nullptr;
  if (dd == 0) {  // constant denominator: quotient = p / c, remainder = the zero constant (:2010-2020)
    TKM_TRY(poly_alloc(ctx, p->x_size, p->y_size, &q));
    int32_t st = vec_scale(ctx, lead_inv, p->d, q->d, p->x_size * p->y_size);
    if (st == TKM_OK) st = tkm_poly_zero(ctx, 1, 1, &r);
    if (st != TKM_OK) {
      poly_release(ctx, q);
      return st;
    }
    *out_q = q;
    *out_r = r;
    return TKM_OK;
  }
  int32_t st = tkm_poly_zero(ctx, p->x_size, p->y_size, &q);
  if (st == TKM_OK) st = tkm_poly_clone(ctx, p, &r);
  if (st == TKM_OK) {
    const size_t len = y_dir ? p->y_size : p->x_size, lines = y_dir ? p->x_size : p->y_size;
    const size_t es = y_dir ? 1 : p->y_size, ls = y_dir ? p->y_size : 1;
    k_divide_uni<<<(unsigned)((lines + 127) / 128), 128, 0, ctx->stream>>>(r->d, q->d, denom->d, den_stride, (uint32_t)dd, lead_inv, len, lines, es, ls);
    st = launch_check(ctx, "k_divide_uni");
  }
  if (st != TKM_OK) {
    poly_release(ctx, q);
    poly_release(ctx, r);
    return st;
  }
  *out_q = q;
  *out_r = r;
  return TKM_OK;
}

int32_t tkm_poly_commit(tkm_ctx *ctx, tkm_poly *p, const tkm_crs *crs, uint8_t out96[96]) {
  API_BEGIN
  TKM_REQUIRE(p && crs && out96, "null argument");
  MsmInput in;
  bool zero;
  TKM_TRY(commit_input(ctx, p, crs, &in, &zero));
  if (zero) {
    memset(out96, 0, 96);
    return TKM_OK;
  }
  return msm_run(ctx, in, out96);
}

int32_t tkm_poly_commit_begin(tkm_ctx *ctx, tkm_poly *p, const tkm_crs *crs, int32_t *out_ticket) {
  API_BEGIN
  TKM_REQUIRE(p && crs && out_ticket, "null argument");
  MsmInput in;
  bool zero;
  TKM_TRY(commit_input(ctx, p, crs, &in, &zero));
  if (zero) in.rows = in.cols = 0;
  return msm_run_async(ctx, in, out_ticket);
}

int32_t tkm_commit_end(tkm_ctx *ctx, int32_t ticket, uint8_t out96[96]) {
  API_BEGIN
  TKM_REQUIRE(out96, "null out pointer");
  return msm_wait(ctx, ticket, out96);
}

}  // extern "C"
