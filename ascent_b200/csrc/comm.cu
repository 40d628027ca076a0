// comm.cu -- multi-GPU sort-last exchange over NVLink peer memory (one process per GPU).
//
// Replaces DirectSendCompositor::CompositeVolume (src/libs/vtkh/compositing/
// DirectSendCompositor.cpp:121-181: DIY all-to-all of sub-tiles, host-side fold, MPI_Barrier,
// DIY gather to rank 0) with ONE kernel per rank: each rank owns a contiguous pixel range, reads
// that range of every rank's quantised image straight out of the peers' HBM (NVLink P2P loads,
// 16 bytes per lane), folds the N layers front-to-back in visibility order with the truncating
// uint8 over operator (ImageCompositor.hpp:15-47), and stores the folded range straight into rank
// 0's result image (P2P stores).  Exchange, fold and gather overlap tile by tile inside the
// kernel; cross-GPU ordering uses epoch flags in the peers' arenas (system-scope release/acquire),
// no host round trip and no MPI barrier.
//
// Arena layout per rank (one cudaMalloc, exported through CUDA IPC):
//   [flags 4 KiB][img rgba8 x2][img depth x2][result rgba8 x2][result depth x2][partial area]
// Images are double-buffered by epoch parity so a rank may start quantising frame e+1 while a
// slower peer still reads frame e.
#include <cstdio>
#include <cstring>

#include "vr_internal.h"

namespace vr
{

namespace
{
constexpr size_t kFlagBytes = 4096;
constexpr int kMaxRanks = 16;

struct Flags
{
  unsigned int ready[kMaxRanks]; // ready[r] = last epoch for which rank r's image is complete
  unsigned int done[kMaxRanks];  // (rank 0 only) done[r] = last epoch rank r finished storing
  unsigned int cta_done;         // local: CTAs of the current fold kernel that have finished
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v)
{
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p)
{
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned blend_u8x4(unsigned front, unsigned back)
{
  const unsigned opacity = 255u - (front >> 24);
  unsigned r = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c)
  {
    const unsigned f = (front >> (8 * c)) & 0xffu;
    const unsigned b = (back >> (8 * c)) & 0xffu;
    r |= ((f + ((opacity * b / 255u) & 0xffu)) & 0xffu) << (8 * c);
  }
  return r;
}
__device__ __forceinline__ float std_min(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float blend_depth(float f, float b)
{
  return std_min(std_min(f, 1.001f), std_min(b, 1.001f));
}

// grid: persistent, multiple of the SM count.  16 bytes (4 pixels) per lane per layer.
__global__ void __launch_bounds__(256) fold_p2p_kernel(const __grid_constant__ FoldP2PParams P)
{
  Flags* my_flags = reinterpret_cast<Flags*>(P.peers[P.rank] + P.off_flags);
  // ---- announce: my image for this epoch is complete (previous kernel on this stream wrote it)
  if (blockIdx.x == 0 && threadIdx.x < P.size)
  {
    Flags* f = reinterpret_cast<Flags*>(P.peers[threadIdx.x] + P.off_flags);
    __threadfence_system();
    st_release_sys(&f->ready[P.rank], P.epoch);
  }
  // ---- wait until every rank's image is complete
  if (threadIdx.x < P.size)
    while (ld_acquire_sys(&my_flags->ready[threadIdx.x]) < P.epoch) __nanosleep(64);
  __syncthreads();

  // my pixel range (in units of 4 pixels)
  const size_t n4 = (P.n_pixels + 3) / 4;
  const size_t chunk4 = (n4 + P.size - 1) / P.size;
  const size_t lo = chunk4 * P.rank;
  const size_t hi = lo + chunk4 < n4 ? lo + chunk4 : n4;

  const uint4* layer_rgba[kMaxRanks];
  const float4* layer_depth[kMaxRanks];
#pragma unroll
  for (int l = 0; l < kMaxRanks; ++l)
    if (l < P.size)
    {
      layer_rgba[l] = reinterpret_cast<const uint4*>(P.peers[P.order[l]] + P.off_img_rgba);
      layer_depth[l] = reinterpret_cast<const float4*>(P.peers[P.order[l]] + P.off_img_depth);
    }
  uint4* out_rgba = reinterpret_cast<uint4*>(P.peers[0] + P.off_res_rgba);
  float4* out_depth = reinterpret_cast<float4*>(P.peers[0] + P.off_res_depth);

  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride)
  {
    // issue all layer loads first (they are independent: N outstanding 16-byte NVLink reads)
    uint4 c[kMaxRanks];
    float4 d[kMaxRanks];
#pragma unroll
    for (int l = 0; l < kMaxRanks; ++l)
      if (l < P.size)
      {
        c[l] = layer_rgba[l][i];
        d[l] = layer_depth[l][i];
      }
    uint4 f = c[0];
    float4 fd = d[0];
#pragma unroll
    for (int l = 1; l < kMaxRanks; ++l)
      if (l < P.size)
      {
        f.x = blend_u8x4(f.x, c[l].x); f.y = blend_u8x4(f.y, c[l].y);
        f.z = blend_u8x4(f.z, c[l].z); f.w = blend_u8x4(f.w, c[l].w);
        fd.x = blend_depth(fd.x, d[l].x); fd.y = blend_depth(fd.y, d[l].y);
        fd.z = blend_depth(fd.z, d[l].z); fd.w = blend_depth(fd.w, d[l].w);
      }
    out_rgba[i] = f;
    out_depth[i] = fd;
  }

  // ---- last CTA out tells rank 0 that my range has landed
  __syncthreads();
  if (threadIdx.x == 0)
  {
    __threadfence_system();
    const unsigned prev = atomicAdd(&my_flags->cta_done, 1u);
    if (prev == gridDim.x - 1)
    {
      my_flags->cta_done = 0;
      Flags* root = reinterpret_cast<Flags*>(P.peers[0] + P.off_flags);
      __threadfence_system();
      st_release_sys(&root->done[P.rank], P.epoch);
    }
  }
}

__global__ void wait_done_kernel(const unsigned int* done, int size, unsigned int epoch)
{
  if (threadIdx.x < size)
    while (ld_acquire_sys(done + threadIdx.x) < epoch) __nanosleep(64);
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Layout
{
  size_t off_flags, off_img_rgba[2], off_img_depth[2], off_res_rgba[2], off_res_depth[2], total;
};
Layout make_layout(size_t max_pixels)
{
  Layout L;
  const size_t px = align_up(max_pixels, 64);
  size_t o = 0;
  L.off_flags = o; o += kFlagBytes;
  for (int b = 0; b < 2; ++b) { L.off_img_rgba[b] = o; o += px * 4; }
  for (int b = 0; b < 2; ++b) { L.off_img_depth[b] = o; o += px * 4; }
  for (int b = 0; b < 2; ++b) { L.off_res_rgba[b] = o; o += px * 4; }
  for (int b = 0; b < 2; ++b) { L.off_res_depth[b] = o; o += px * 4; }
  L.total = align_up(o, 2 << 20);
  return L;
}

} // namespace

cudaError_t launch_fold_p2p(const FoldP2PParams& p, int sm_count, cudaStream_t s)
{
  fold_p2p_kernel<<<sm_count * 2, 256, 0, s>>>(p);
  return cudaGetLastError();
}

void comm_destroy(vr_ctx* ctx)
{
  Comm& c = ctx->comm;
  if (!c.on) return;
  for (int r = 0; r < (int)c.peer.size(); ++r)
    if (r != c.rank && c.peer[r]) cudaIpcCloseMemHandle(c.peer[r]);
  cudaFree(c.peer_dev);
  cudaFree(c.arena);
  c.on = false;
}

// point the context's image/result buffers at the arena halves of the NEXT epoch's parity
vr_status comm_bind_frame(vr_ctx* ctx, size_t n_pixels)
{
  Comm& c = ctx->comm;
  if (n_pixels > c.max_pixels)
  {
    ctx->err = "image larger than the max_pixels given to vr_comm_init";
    return VR_ERR_INVALID;
  }
  const Layout L = make_layout(c.max_pixels);
  const int b = (c.epoch + 1) & 1;
  ctx->img_rgba = reinterpret_cast<uchar4*>(c.arena + L.off_img_rgba[b]);
  ctx->img_depth = reinterpret_cast<float*>(c.arena + L.off_img_depth[b]);
  ctx->res_rgba = reinterpret_cast<uchar4*>(c.arena + L.off_res_rgba[b]);
  ctx->res_depth = reinterpret_cast<float*>(c.arena + L.off_res_depth[b]);
  return VR_OK;
}

} // namespace vr

using namespace vr;

static vr_status cfail(vr_ctx* ctx, vr_status st, const char* what, cudaError_t e)
{
  char buf[256];
  snprintf(buf, sizeof(buf), "%s: %s", what, e == cudaSuccess ? "invalid use" : cudaGetErrorString(e));
  ctx->err = buf;
  return st;
}

extern "C" vr_status vr_comm_init(vr_ctx* ctx, int rank, int n_ranks, size_t max_pixels,
                                  size_t max_partials, void* handle_out)
{
  if (!ctx) return VR_ERR_INVALID;
  if (ctx->comm.on) return cfail(ctx, VR_ERR_STATE, "vr_comm_init: already initialised", cudaSuccess);
  if (rank < 0 || n_ranks < 1 || rank >= n_ranks || n_ranks > kMaxRanks || !handle_out || max_pixels == 0)
    return cfail(ctx, VR_ERR_INVALID, "vr_comm_init: bad rank/size (max 16 ranks) or NULL handle", cudaSuccess);
  cudaSetDevice(ctx->device);
  Comm& c = ctx->comm;
  c.rank = rank;
  c.size = n_ranks;
  c.max_pixels = max_pixels;
  c.max_partials = max_partials;
  const Layout L = make_layout(max_pixels);
  c.arena_bytes = L.total;
  cudaError_t e = cudaMalloc(&c.arena, c.arena_bytes);
  if (e != cudaSuccess) return cfail(ctx, VR_ERR_NOMEM, "vr_comm_init: arena", e);
  cudaMemset(c.arena, 0, kFlagBytes);
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, c.arena);
  if (e != cudaSuccess) return cfail(ctx, VR_ERR_CUDA, "cudaIpcGetMemHandle", e);
  static_assert(sizeof(h) == VR_IPC_HANDLE_BYTES, "CUDA IPC handle size");
  std::memcpy(handle_out, &h, sizeof(h));
  // the context's own image/result buffers move into the arena
  cudaStreamSynchronize(ctx->stream);
  if (!ctx->img_in_arena)
  {
    cudaFree(ctx->img_rgba); cudaFree(ctx->img_depth); cudaFree(ctx->res_rgba); cudaFree(ctx->res_depth);
  }
  ctx->img_in_arena = true;
  c.peer.assign(n_ranks, nullptr);
  c.peer[rank] = c.arena;
  c.epoch = 0;
  c.on = true; // connected == peer_dev != nullptr
  comm_bind_frame(ctx, 1);
  cudaDeviceSynchronize();
  return VR_OK;
}

extern "C" vr_status vr_comm_connect(vr_ctx* ctx, const void* all_handles)
{
  if (!ctx) return VR_ERR_INVALID;
  Comm& c = ctx->comm;
  if (!c.on || !all_handles) return cfail(ctx, VR_ERR_STATE, "vr_comm_connect: call vr_comm_init first", cudaSuccess);
  cudaSetDevice(ctx->device);
  const unsigned char* hb = static_cast<const unsigned char*>(all_handles);
  for (int r = 0; r < c.size; ++r)
  {
    if (r == c.rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, hb + (size_t)r * VR_IPC_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return cfail(ctx, VR_ERR_CUDA, "cudaIpcOpenMemHandle (peer access over NVLink)", e);
    c.peer[r] = static_cast<unsigned char*>(p);
  }
  cudaError_t e = cudaMalloc(&c.peer_dev, sizeof(unsigned char*) * kMaxRanks);
  if (e != cudaSuccess) return cfail(ctx, VR_ERR_NOMEM, "vr_comm_connect", e);
  unsigned char* table[kMaxRanks] = { nullptr };
  for (int r = 0; r < c.size; ++r) table[r] = c.peer[r];
  e = cudaMemcpy(c.peer_dev, table, sizeof(table), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return cfail(ctx, VR_ERR_CUDA, "vr_comm_connect", e);
  return VR_OK;
}

extern "C" vr_status vr_comm_composite_images(vr_ctx* ctx, const int* vis_order)
{
  if (!ctx) return VR_ERR_INVALID;
  Comm& c = ctx->comm;
  if (!c.on || !c.peer_dev) return cfail(ctx, VR_ERR_STATE, "vr_comm_composite_images: not connected", cudaSuccess);
  if (!vis_order || ctx->W <= 0) return cfail(ctx, VR_ERR_INVALID, "vr_comm_composite_images: no image / NULL order", cudaSuccess);
  cudaSetDevice(ctx->device);
  const Layout L = make_layout(c.max_pixels);
  c.epoch += 1; // the image was quantised into parity (epoch+1)&1 by vr_image_from_canvas
  const int b = c.epoch & 1;
  FoldP2PParams p;
  std::memset(&p, 0, sizeof(p));
  p.peers = c.peer_dev;
  p.rank = c.rank;
  p.size = c.size;
  p.epoch = c.epoch;
  p.n_pixels = (size_t)ctx->W * ctx->H;
  p.off_img_rgba = L.off_img_rgba[b];
  p.off_img_depth = L.off_img_depth[b];
  p.off_res_rgba = L.off_res_rgba[b];
  p.off_res_depth = L.off_res_depth[b];
  p.off_flags = L.off_flags;
  // fold order: ranks sorted by ascending vis_order (stable)
  int idx[kMaxRanks];
  for (int i = 0; i < c.size; ++i) idx[i] = i;
  for (int i = 1; i < c.size; ++i)
  {
    int k = idx[i], j = i - 1;
    while (j >= 0 && vis_order[idx[j]] > vis_order[k]) { idx[j + 1] = idx[j]; --j; }
    idx[j + 1] = k;
  }
  for (int i = 0; i < c.size; ++i) p.order[i] = idx[i];
  cudaError_t e = launch_fold_p2p(p, ctx->sm_count, ctx->stream);
  if (e != cudaSuccess) return cfail(ctx, VR_ERR_CUDA, "fold_p2p launch", e);
  ctx->launches++;
  if (c.rank == 0)
  {
    const Flags* f = reinterpret_cast<const Flags*>(c.arena + L.off_flags);
    wait_done_kernel<<<1, 32, 0, ctx->stream>>>(f->done, c.size, c.epoch);
    ctx->launches++;
    // result of this epoch (res_* currently points at parity b)
    ctx->res_rgba = reinterpret_cast<uchar4*>(c.arena + L.off_res_rgba[b]);
    ctx->res_depth = reinterpret_cast<float*>(c.arena + L.off_res_depth[b]);
  }
  // next frame quantises into the other parity
  const int nb = (c.epoch + 1) & 1;
  ctx->img_rgba = reinterpret_cast<uchar4*>(c.arena + L.off_img_rgba[nb]);
  ctx->img_depth = reinterpret_cast<float*>(c.arena + L.off_img_depth[nb]);
  return VR_OK;
}

extern "C" vr_status vr_image_result_download(vr_ctx* ctx, uint8_t* rgba, float* depth)
{
  if (!ctx) return VR_ERR_INVALID;
  if (ctx->W <= 0 || !ctx->res_rgba) return cfail(ctx, VR_ERR_STATE, "vr_image_result_download: no result", cudaSuccess);
  const size_t n = (size_t)ctx->W * ctx->H;
  cudaError_t e = cudaSuccess;
  if (rgba) e = cudaMemcpyAsync(rgba, ctx->res_rgba, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && depth) e = cudaMemcpyAsync(depth, ctx->res_depth, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cfail(ctx, VR_ERR_CUDA, "vr_image_result_download", e);
  return VR_OK;
}

extern "C" vr_status vr_image_result_to_canvas(vr_ctx* ctx)
{
  if (!ctx) return VR_ERR_INVALID;
  if (ctx->W <= 0 || !ctx->res_rgba) return cfail(ctx, VR_ERR_STATE, "vr_image_result_to_canvas: no result", cudaSuccess);
  return vr_image_to_canvas_dev(ctx, reinterpret_cast<const uint8_t*>(ctx->res_rgba), ctx->res_depth);
}

extern "C" vr_status vr_comm_composite_partials(vr_ctx* ctx)
{
  if (!ctx) return VR_ERR_INVALID;
  return cfail(ctx, VR_ERR_STATE, "vr_comm_composite_partials: multi-GPU partial exchange not built yet",
               cudaSuccess);
}
