// Multi-GPU entry points of libtokamak_b200: one context per GPU (one process per GPU, or several
// contexts in one process), NCCL over NVLink/NVSwitch as plumbing.  The reference has no multi-device code at all
// (libs/src/utils/mod.rs:90-96 pins device 0); this is the C-ABI a Rust prover binds for the sharded paths:
//
//   * G1 MSM shards by point range: every rank runs the full Pippenger on its slice; the 96-byte affine partial sums are
//     all-gathered (the only collective) and every rank adds them up, so the result is available everywhere.
//   * the bivariate NTT shards by rows (X index): local Y pass, ONE all-to-all (grouped ncclSend/ncclRecv of the
//     (x/G) x (y/G) tiles) that re-shards rows -> columns, local X pass on whole columns.  The inverse runs the mirror image.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2), so single-GPU users of the library need no NCCL at all and a
// process that already carries an NCCL (e.g. under PyTorch) shares it.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace tkm {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return nullptr;
  bool ok = true;
  auto sym = [&](const char *name) {
    void *p = dlsym(h, name);
    if (!p) ok = false;
    return p;
  };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
  api.Send = (decltype(api.Send))sym("ncclSend");
  api.Recv = (decltype(api.Recv))sym("ncclRecv");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  if (!ok) return nullptr;
  api.handle = h;
  return &api;
}

#define TKM_NCCL(api, expr)                                                                                    \
  do {                                                                                                         \
    ncclResult_t _r = (expr);                                                                                  \
    if (_r != ncclSuccess) return fail(TKM_ERR_CUDA, "%s failed: %s", #expr, (api)->GetErrorString(_r));       \
  } while (0)

// [rows][G][tile] <-> [G][rows][tile] (tile = contiguous run of `tile` field elements): the packing either side of the
// all-to-all, 128-bit accesses.
__global__ void __launch_bounds__(256) k_permute_tiles(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t rows, size_t G, size_t tile16,
                                                       int to_peer_major) {
  const size_t total = rows * G * tile16;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t w = e % tile16, rg = e / tile16;
    size_t r, g;
    if (to_peer_major) {  // e enumerates the output [G][rows][tile]
      r = rg % rows;
      g = rg / rows;
      out[e] = in[(r * G + g) * tile16 + w];
    } else {  // e enumerates the output [rows][G][tile]
      g = rg % G;
      r = rg / G;
      out[e] = in[(g * rows + r) * tile16 + w];
    }
  }
}

static int32_t all_to_all(tkm_ctx *ctx, NcclApi *api, const Fr *send, Fr *recv, size_t chunk_elems) {
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  TKM_NCCL(api, api->GroupStart());
  for (int p = 0; p < ctx->comm_world; p++) {
    TKM_NCCL(api, api->Send(send + (size_t)p * chunk_elems, chunk_elems * sizeof(Fr), ncclUint8, p, comm, ctx->stream));
    TKM_NCCL(api, api->Recv(recv + (size_t)p * chunk_elems, chunk_elems * sizeof(Fr), ncclUint8, p, comm, ctx->stream));
  }
  TKM_NCCL(api, api->GroupEnd());
  return TKM_OK;
}

}  // namespace tkm

using namespace tkm;

#define API_BEGIN                                                  \
  if (!ctx) return fail(TKM_ERR_INVALID_ARGUMENT, "null context"); \
  cudaSetDevice(ctx->device);

extern "C" {

int32_t tkm_comm_unique_id(uint8_t out_id[TKM_COMM_ID_BYTES]) {
  if (!out_id) return fail(TKM_ERR_INVALID_ARGUMENT, "null out pointer");
  NcclApi *api = nccl_api();
  if (!api) return fail(TKM_ERR_INTERNAL, "NCCL (libnccl.so.2) could not be loaded: %s", dlerror());
  static_assert(sizeof(ncclUniqueId) <= TKM_COMM_ID_BYTES, "unique id does not fit");
  ncclUniqueId id;
  TKM_NCCL(api, api->GetUniqueId(&id));
  memset(out_id, 0, TKM_COMM_ID_BYTES);
  memcpy(out_id, &id, sizeof id);
  return TKM_OK;
}

int32_t tkm_comm_init(tkm_ctx *ctx, const uint8_t id[TKM_COMM_ID_BYTES], int32_t rank, int32_t world) {
  API_BEGIN
  TKM_REQUIRE(id, "null id");
  TKM_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank %d out of range [0,%d)", rank, world);
  TKM_REQUIRE(!ctx->comm, "this context already belongs to a communicator");
  NcclApi *api = nccl_api();
  if (!api) return fail(TKM_ERR_INTERNAL, "NCCL (libnccl.so.2) could not be loaded: %s", dlerror());
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof uid);
  ncclComm_t comm = nullptr;
  TKM_NCCL(api, api->CommInitRank(&comm, world, uid, rank));
  ctx->comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  return TKM_OK;
}

int32_t tkm_comm_destroy(tkm_ctx *ctx) {
  API_BEGIN
  if (!ctx->comm) return TKM_OK;
  NcclApi *api = nccl_api();
  cudaStreamSynchronize(ctx->stream);
  if (api) api->CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
  ctx->comm_world = 1;
  ctx->comm_rank = 0;
  return TKM_OK;
}

int32_t tkm_comm_rank(tkm_ctx *ctx, int32_t *out_rank, int32_t *out_world) {
  API_BEGIN
  if (out_rank) *out_rank = ctx->comm_rank;
  if (out_world) *out_world = ctx->comm_world;
  return TKM_OK;
}

int32_t tkm_msm_g1_sharded(tkm_ctx *ctx, const void *dev_scalars, int32_t scalars_mont, const void *dev_bases_mont, size_t n_local,
                           uint8_t out96[96]) {
  API_BEGIN
  TKM_REQUIRE(out96, "null out pointer");
  uint8_t part[96];
  TKM_TRY(tkm_msm_g1(ctx, dev_scalars, scalars_mont, dev_bases_mont, n_local, part));
  if (!ctx->comm || ctx->comm_world == 1) {
    memcpy(out96, part, 96);
    return TKM_OK;
  }
  NcclApi *api = nccl_api();
  Scratch<uint8_t> mine, all;
  TKM_TRY(mine.alloc(ctx, 96));
  TKM_TRY(all.alloc(ctx, 96 * (size_t)ctx->comm_world));
  TKM_CUDA(cudaMemcpyAsync(mine.p, part, 96, cudaMemcpyHostToDevice, ctx->stream));
  TKM_NCCL(api, api->AllGather(mine.p, all.p, 96, ncclUint8, (ncclComm_t)ctx->comm, ctx->stream));
  std::string host(96 * (size_t)ctx->comm_world, '\0');
  TKM_CUDA(cudaMemcpyAsync(&host[0], all.p, host.size(), cudaMemcpyDeviceToHost, ctx->stream));
  TKM_CUDA(cudaStreamSynchronize(ctx->stream));
  return tkm_g1_sum(ctx, (const uint8_t *)host.data(), (size_t)ctx->comm_world, out96);
}

int32_t tkm_bintt_sharded(tkm_ctx *ctx, const void *dev_in, void *dev_out, size_t x_size, size_t y_size, int32_t dir, const uint8_t *cx,
                          const uint8_t *cy) {
  API_BEGIN
  TKM_REQUIRE(dev_in && dev_out && dev_in != dev_out, "null or aliased buffers");
  TKM_REQUIRE(dir == TKM_FORWARD || dir == TKM_INVERSE, "bad direction");
  const size_t G = ctx->comm ? (size_t)ctx->comm_world : 1;
  TKM_REQUIRE(is_pow2(x_size) && is_pow2(y_size), "biNTT sizes must be powers of two (got %zu x %zu)", x_size, y_size);
  TKM_REQUIRE(x_size % G == 0 && y_size % G == 0, "both extents must be divisible by the number of ranks (%zu)", G);
  Fr gx, gy;
  if (cx) gx = fr_from_bytes_host(cx);
  if (cy) gy = fr_from_bytes_host(cy);
  const size_t xb = x_size / G, yb = y_size / G, n_local = xb * y_size;
  const Fr *in = (const Fr *)dev_in;
  Fr *out = (Fr *)dev_out;
  if (G == 1) return bintt_dev(ctx, in, out, x_size, y_size, dir, cx ? &gx : nullptr, cy ? &gy : nullptr);
  NcclApi *api = nccl_api();
  Scratch<Fr> tmp;
  TKM_TRY(tmp.alloc(ctx, n_local));
  const unsigned grid = grid_for(n_local * 2, 256, ctx->sm_count);
  if (dir == TKM_FORWARD) {
    // rows [xb][y] -> Y pass -> [xb][G][yb] packed peer-major -> all-to-all -> [G*xb][yb] = columns [x][yb] -> X pass
    TKM_TRY(ntt_axis(ctx, in, out, xb, y_size, 1, TKM_FORWARD, cy ? &gy : nullptr));
    k_permute_tiles<<<grid, 256, 0, ctx->stream>>>((const uint4 *)out, (uint4 *)tmp.p, xb, G, yb * 2, 1);
    TKM_TRY(launch_check(ctx, "k_permute_tiles"));
    TKM_TRY(all_to_all(ctx, api, tmp.p, out, xb * yb));
    return ntt_axis(ctx, out, out, 1, x_size, yb, TKM_FORWARD, cx ? &gx : nullptr);
  }
  // columns [x][yb] -> inverse X pass -> chunk p = rows of peer p -> all-to-all -> [G][xb][yb] -> unpack to rows [xb][y] -> inverse Y pass
  TKM_TRY(ntt_axis(ctx, in, tmp.p, 1, x_size, yb, TKM_INVERSE, cx ? &gx : nullptr));
  TKM_TRY(all_to_all(ctx, api, tmp.p, out, xb * yb));
  k_permute_tiles<<<grid, 256, 0, ctx->stream>>>((const uint4 *)out, (uint4 *)tmp.p, xb, G, yb * 2, 0);
  TKM_TRY(launch_check(ctx, "k_permute_tiles"));
  return ntt_axis(ctx, tmp.p, out, xb, y_size, 1, TKM_INVERSE, cy ? &gy : nullptr);
}

}  // extern "C"
