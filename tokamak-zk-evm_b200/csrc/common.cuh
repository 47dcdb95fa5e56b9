// Shared host-side plumbing of libtokamak_b200: status/error handling, the context object,
// stream-ordered scratch allocation and launch accounting.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/tokamak_b200.h"
#include "g1.cuh"

namespace tkm {

// Thread-local last-error text returned by tkm_last_error().
std::string &last_error();
int32_t fail(int32_t code, const char *fmt, ...);

#define TKM_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) return ::tkm::fail(TKM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,         \
                                               cudaGetErrorString(_e), __FILE__, __LINE__);         \
  } while (0)

#define TKM_TRY(expr)                 \
  do {                                \
    int32_t _s = (expr);              \
    if (_s != TKM_OK) return _s;      \
  } while (0)

#define TKM_REQUIRE(cond, ...)                                           \
  do {                                                                   \
    if (!(cond)) return ::tkm::fail(TKM_ERR_INVALID_ARGUMENT, __VA_ARGS__); \
  } while (0)

}  // namespace tkm

constexpr int TKM_MAX_TICKETS = 32;
struct tkm_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  // NTT domain: tw[k] = omega_M^k in Montgomery form, k = 0..M/2 (tw[M/2] = -1), M = 2^domain_log2.
  int32_t domain_log2 = -1;
  tkm::Fr *twiddles = nullptr;
  tkm::Fr inv_pow2[33];  // 2^-k in Montgomery form (1/n factors of the inverse transforms)
  uint64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // events bracketing the most recent launch of a dominant kernel (k_accumulate, the last k_ntt_pass of a transform):
  // bench.py's roofline divides the kernel's algorithmic work by this duration (tkm_kernel_time_last)
  cudaEvent_t kev0 = nullptr, kev1 = nullptr;
  bool kernel_timed = false;
  // the same for the most recent polynomial-engine kernel of interest (k_polyexpr): bench.py's poly_engine GB/s
  cudaEvent_t pev0 = nullptr, pev1 = nullptr;
  bool poly_kernel_timed = false;
  // window table (64 x 16 affine multiples) of the most recent tkm_g1_fixed_base_mul base
  tkm::G1Affine *fb_table = nullptr;
  uint8_t fb_base[96] = {};
  bool fb_valid = false;
  // asynchronous commitments: the serial recombination tail (k_final) of MSM k runs on side_stream while MSM k+1 accumulates
  struct Ticket {
    tkm::G1Xyzz *parts = nullptr;  // per-ticket device buffers read by the tail (parts | window sums | result)
    uint8_t *host = nullptr;       // pinned 96-byte result slot
    cudaEvent_t ready = nullptr, done = nullptr;
    bool busy = false, zero = false;
  };
  cudaStream_t side_stream = nullptr;
  Ticket tickets[TKM_MAX_TICKETS];
  // copy engine side of the pipelined host-buffer MSM (created on first use)
  cudaStream_t copy_stream = nullptr;
  // share of the previous host-buffer MSM's duration its host-to-device copies took (0 = not measured yet): picks the piece layout
  float h2d_share = 0.f;
  bool h2d_copy_bound = false;  // with hysteresis: set above 0.55, cleared below 0.40
  cudaEvent_t copy_t[4] = {};  // timing events: copies begin / end, call begin / end
  cudaEvent_t copy_ev[34] = {};  // [2k] scalars of piece k arrived, [2k+1] its bases; [32] staging buffers allocated
  // per-device launch state (function attributes and occupancy are properties of the device the context lives on, so they
  // are cached here and not in process-wide statics: a second context on another GPU must get its own shared-memory opt-in)
  bool ntt_attr_set[4] = {false, false, false, false};
  int acc_occ = 0, bits_occ = 0;
  // entry counts of the affine pair tree's levels in the most recent MSM accumulation pass (pinned; written by async copies)
  uint32_t *tree_counts = nullptr;  // [9]
  uint32_t tree_levels = 0;
  // second stream of the MSM pair tree (the two halves of the bucket range run concurrently), created on first use
  cudaStream_t tree_stream = nullptr;
  cudaEvent_t tree_fork = nullptr, tree_join = nullptr;
  // multi-GPU: the NCCL communicator this context belongs to (comm.cu); null = single GPU
  void *comm = nullptr;
  int32_t comm_rank = 0, comm_world = 1;
};

struct tkm_poly {
  tkm::Fr *d = nullptr;  // row-major [x_size][y_size], Montgomery form
  size_t x_size = 0, y_size = 0;
};

struct tkm_crs {
  tkm::G1Affine *d = nullptr;  // row-major [rows][cols], Montgomery form
  size_t rows = 0, cols = 0;
  bool owned = true;
  tkm::G1Affine *pre = nullptr;  // optional fixed-base tables [pre_W][rows*cols]: 2^(pre_c*w) * P
  uint32_t pre_c = 0, pre_W = 0;
  // x coordinates alone in 64-byte slots (same index space as d, or as pre when tables exist): what the forward pass of the
  // MSM's affine pair tree gathers.  Built on the first commitment against this CRS.
  uint4 *xpad = nullptr;
  bool xpad_for_pre = false;
};

namespace tkm {

// Stream-ordered scratch buffer (cudaMallocAsync pool; freed on the same stream).
template <class T>
struct Scratch {
  T *p = nullptr;
  cudaStream_t s = nullptr;
  Scratch() = default;
  Scratch(const Scratch &) = delete;
  Scratch &operator=(const Scratch &) = delete;
  int32_t alloc(tkm_ctx *ctx, size_t count) {
    s = ctx->stream;
    if (count == 0) count = 1;
    cudaError_t e = cudaMallocAsync((void **)&p, count * sizeof(T), s);
    if (e != cudaSuccess) {
      p = nullptr;
      return fail(TKM_ERR_ALLOCATION, "cudaMallocAsync(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
    }
    return TKM_OK;
  }
  ~Scratch() {
    if (p) cudaFreeAsync(p, s);
  }
};

inline bool is_pow2(size_t v) { return v && !(v & (v - 1)); }
inline uint32_t log2_exact(size_t v) {
  uint32_t l = 0;
  while (((size_t)1 << l) < v) l++;
  return l;
}
inline size_t next_pow2(size_t v) {
  // _find_size_as_twopower (bivariate_polynomial/mod.rs:72-86)
  if (is_pow2(v)) return v;
  size_t r = 1;
  while (r <= v) r <<= 1;
  return r;
}
inline unsigned grid_for(size_t work, unsigned block, unsigned sm_count, unsigned waves = 8) {
  size_t blocks = (work + block - 1) / block;
  size_t cap = (size_t)sm_count * waves;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  return (unsigned)blocks;
}

// Host-side scalar helpers (tiny, run on the device via one-thread kernels is overkill; these are
// exact big-integer-free Montgomery ops through the host emulation of ff.cuh).
Fr fr_from_bytes_host(const uint8_t *b32);  // canonical bytes -> Montgomery
void fr_to_bytes_host(const Fr &a, uint8_t *b32);

// Internal cross-file entry points ------------------------------------------------------------
int32_t launch_check(tkm_ctx *ctx, const char *what);
int32_t vec_to_mont(tkm_ctx *ctx, const Fr *in, Fr *out, size_t n);
int32_t vec_from_mont(tkm_ctx *ctx, const Fr *in, Fr *out, size_t n);
int32_t vec_op(tkm_ctx *ctx, int op, const Fr *a, const Fr *b, Fr *out, size_t n);
int32_t vec_scale(tkm_ctx *ctx, const Fr &s, const Fr *a, Fr *out, size_t n);
int32_t vec_inv(tkm_ctx *ctx, const Fr *a, Fr *out, size_t n);
int32_t vec_fill(tkm_ctx *ctx, const Fr &s, Fr *out, size_t n);
int32_t vec_reduce(tkm_ctx *ctx, int op, const Fr *a, const Fr *b, size_t n, Fr *host_out);
int32_t vec_outer_product(tkm_ctx *ctx, const Fr *col, const Fr *row, Fr *out, size_t rows, size_t cols);
int32_t vec_suffix_product(tkm_ctx *ctx, const Fr *in, Fr *out, size_t n);
int32_t vec_mul_x_minus_one(tkm_ctx *ctx, const Fr *in, Fr *out, size_t x_size, size_t y_size);
int32_t vec_transpose(tkm_ctx *ctx, const Fr *in, Fr *out, size_t rows, size_t cols);
int32_t bintt_dev(tkm_ctx *ctx, const Fr *in, Fr *out, size_t x, size_t y, int dir, const Fr *coset_x,
                  const Fr *coset_y);
int32_t ntt_axis(tkm_ctx *ctx, const Fr *in, Fr *out, size_t outer, size_t n, size_t inner, int dir, const Fr *coset);
int32_t ntt_axis_scatter(tkm_ctx *ctx, const Fr *in, size_t outer, size_t n, size_t inner, int dir, const Fr *coset, void *const *peers,
                         uint32_t n_peers, uint64_t stride_a, uint64_t stride_b, uint64_t b0);
int32_t g1_to_mont_dev(tkm_ctx *ctx, const G1Affine *in, G1Affine *out, size_t n);
int32_t g1_from_mont_dev(tkm_ctx *ctx, const G1Affine *in, G1Affine *out, size_t n);
struct MsmInput {
  const Fr *scalars;
  bool scalars_mont;
  size_t scalar_row_stride;
  const G1Affine *bases;
  size_t base_row_stride;
  size_t rows, cols;
  const uint32_t *idx;  // optional gather indices into bases (rows must be 1)
  uint32_t pre_c = 0;       // fixed-base tables: window bits the tables were built for (0 = plain bases)
  uint32_t pre_stride = 0;  // fixed-base tables: points per table (table w starts at bases + w*pre_stride)
  const uint4 *xpad = nullptr;  // optional: 64-byte x slots over the same index space as bases (see tkm_crs::xpad)
  // Host pipeline: the bases may still be in flight when the pass starts.  Digit decomposition, sort and the tree's offset
  // tables need only the scalars; the pass waits for this event right before the first kernel that reads a base, then (when
  // bases_canonical) converts them to Montgomery form in place.
  cudaEvent_t bases_ready = nullptr;
  bool bases_canonical = false;
};
int32_t msm_build_xpad(tkm_ctx *ctx, const G1Affine *bases, size_t row_stride, size_t rows, size_t cols, uint32_t tables, size_t table_stride, uint4 *xpad);
int32_t crs_precompute(tkm_ctx *ctx, const G1Affine *base, size_t n, uint32_t c, G1Affine **out_table, uint32_t *out_W);
int32_t msm_run(tkm_ctx *ctx, const MsmInput &in, uint8_t out96[96]);
int32_t msm_run_async(tkm_ctx *ctx, const MsmInput &in, int32_t *out_ticket);
int32_t msm_wait(tkm_ctx *ctx, int32_t ticket, uint8_t out96[96]);
int32_t msm_host_pipelined(tkm_ctx *ctx, const uint8_t *scalars, const uint8_t *bases, size_t n, uint32_t pieces, uint8_t out96[96]);

}  // namespace tkm
