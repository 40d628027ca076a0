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
// The float-partial path (vtkh_diy_partial_redistribute.hpp:58-152 all-to-all, PartialCompositor.cpp
// :329-488 serial std::sort + fold, vtkh_diy_partial_collect.hpp:57-141 gather) becomes: every rank
// orders its own list by (pixel, depth) in its arena, then ONE kernel per rank pulls, for each pixel
// of the range it owns, that pixel's run from every rank's list over NVLink, merges the N runs by
// depth in registers, folds front to back and appends the result to rank 0's list.  No receive
// buffers, no second sort, no host in the loop.
//
// Arena layout per rank (one cudaMalloc, exported through CUDA IPC):
//   [flags 4 KiB][img rgba8 x2][img depth x2][result rgba8 x2][result depth x2]
//   [partial offsets x2][sorted partials x2][composited partials (rank 0)]
// Buffers are double-buffered by epoch parity so a rank may start producing frame e+1 while a
// slower peer still reads frame e.
#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "vr_internal.h"
#include "vr_radixk.hpp"

namespace vr
{
vr_status ensure_frame_pub(vr_ctx* ctx, int W, int H); // vr_api.cu

namespace
{
constexpr size_t kFlagBytes = 4096;
constexpr int kMaxRanks = 16;
// A rank's quantised image of frame e lives in ring slot e % kImgRing (its own arena when pulled, the owners'
// receive slots when pushed).  Two slots would do for strictly alternating render/exchange; the third
// lets a rank trace frame e+1 BEFORE it issues the exchange of frame e (VR_FRAME_AHEAD); the fourth lets
// that trace start while the rank's own exchange of frame e-1 is still RUNNING on the exchange stream:
// frame e+1 reuses the slot of frame e-3, and every peer has finished reading e-3 once this rank's
// exchange of e-2 -- the one before the latest, which vr_trace_to_image waits for -- has completed.
constexpr int kImgRing = 6;
constexpr int kLayerRing = 3; // ray layers of frame e live in buffer e % 3 (same argument, no frame is traced ahead)

struct Flags
{
  unsigned int ready[kMaxRanks]; // ready[r] = last epoch for which rank r's image is complete
  unsigned int done[kMaxRanks];  // (rank 0 only) done[r] = last epoch rank r finished storing
  unsigned int cta_done;         // local: CTAs of the current fold kernel that have finished
  // partial path
  unsigned int p_ready[kMaxRanks]; // p_ready[r] = last epoch for which rank r's sorted list is complete
  unsigned int p_done[kMaxRanks];  // (rank 0 only) ranks whose composited range has landed
  int p_minmax[2][kMaxRanks][2];   // [epoch parity][rank] = {min,max} pixel id of that rank's list
  unsigned int p_cta_done;
  unsigned int p_overflow;         // set when a list did not fit max_partials
  unsigned long long p_out_count;  // (rank 0) length of the composited list
  // image path: the rectangle {x0,y0,x1,y1} (x in multiples of 4) outside of which rank r's image
  // of that epoch parity is empty (colour 0, depth 1.001) and need not be read
  int img_rect[2][kMaxRanks][4];
  // geometry of this arena: every rank must have been initialised with the same values, because
  // a rank addresses its peers' arenas with its own layout (checked in vr_comm_connect)
  unsigned long long cfg_max_pixels, cfg_max_partials;
  // depth broadcast (Scene::SynchDepths): rank 0's staging copy is complete for epoch s_ready;
  // (rank 0 only) s_done[r] = last epoch rank r has finished pulling
  unsigned int s_ready;
  unsigned int s_done[kMaxRanks];
  // (rank 0 only) outside these rectangles {x0,y0,x1,y1} the result image of each parity / the canvas
  // hold the cleared value, as left by the last exchange kernel (host decides whether that still holds)
  int clean_res[2][4];
  int clean_canvas[4];
  // image path: rank r PUSHED its image of that epoch parity into the owners' receive slots (sampler
  // mode 5) instead of leaving it in its own arena
  int img_pushed[2][kMaxRanks];
  // a rank that hit a rank-local error publishes "I abort exchange <epoch>" here (per path: image,
  // partial list, layers) together with its ready flag, so that nobody waits for data that never comes
  unsigned int aborted[3][kMaxRanks];
  // (local) first exchange of this rank that did not complete normally: epoch, path, reason
  // (1 = a wait ran into the time limit, 2 = a peer aborted), for the host to report
  unsigned int err_epoch, err_path, err_reason, err_peer;
  // (diagnostics, VR_TIMELINE=1) globaltimer stamps of the last image exchange on this GPU:
  // [0] fold kernel entered  [1] all ranks ready  [2] last fold CTA out ("done" released)
  // [3] to-canvas kernel entered (rank 0)  [4] all ranks done  [5] its CTA 0 finished  [6] end of the last trace
  // CTA 0 of the fold kernel: [8] prologue done  [9] own chunks folded  [10] rank 0's clears done  [11] fenced + counted
  unsigned long long timeline[16];
};
static_assert(sizeof(Flags) <= 4096, "flag block");
enum { kPathImage = 0, kPathPartials = 1, kPathLayers = 2 };

__device__ __forceinline__ ExchangeError* err_of(Flags* f) { return reinterpret_cast<ExchangeError*>(&f->err_epoch); }

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
// Ownership is round-robin in chunks of 1024 pixels (balanced even when the covered region is a
// band of rows).  A rank's image is only read inside the rectangle it announced: outside it the
// layer is colour 0 / depth 1.001, which ImageCompositor::Blend treats as the identity for colour
// and as the constant 1.001 for depth -- folded in without touching NVLink.  Groups no rank covers
// are written by rank 0 itself.
constexpr int kChunkGroups = 256; // 4-pixel groups per ownership chunk

// TO_CANVAS: rank 0 also produces its float canvas (Renderer::ImageToCanvas, Renderer.cpp:265-283):
// the groups no rank covers are written here, while the peers' pixels are still in flight; the
// covered ones are converted by covered_to_canvas_kernel once they have landed.
// ZBUF: the opaque-surface mode of the same exchange (Compositor Z_BUFFER_SURFACE ->
// RadixKCompositor::CompositeSurface, RadixKCompositor.cpp:35-180): the per-pixel operator is
// ImageCompositor::ZBufferComposite (ImageCompositor.hpp:49-76) -- nearest fragment wins, fragments
// with depth > 1 never replace, an incoming fragment at EQUAL depth does.  Select-nearest is associative,
// so the reference's multi-round tree (k = 8) and one round of k = N give the same image provided equal-depth
// fragments are visited in the tree's order: zselect_group takes that order, piece by piece, from the
// closed-form schedule of vr_radixk.hpp (pinned against the reference's own reduce_images + DIY,
// tests/test_oracle_radixk.py).  One round is the right radix on NVSwitch: every peer at full bandwidth.
// NR: number of ranks rounded up to a power of two (the per-layer registers are fully unrolled).
// Four pixels (one 16-byte group, linear index 4 i .. 4 i + 3) of the radix-k z-select: the nearest fragment
// with depth <= 1 wins; among equal depths the one the reference's tree composites LAST into the piece that
// ends up owning the pixel (ImageCompositor.hpp:65: `front.depth < depth` keeps, so <= replaces); if no
// fragment is <= 1 the piece owner's own pixel stays (the first of the sequence is always the owner).
template <int NR>
__device__ __forceinline__ void zselect_group(const FoldP2PParams& P, const uint4 (&c)[NR], const float4 (&d)[NR],
                                              unsigned cover, size_t i, uint4& f, float4& fd)
{
  unsigned oc[4];
  float od[4];
#pragma unroll
  for (int q = 0; q < 4; ++q)
  {
    const unsigned idx = (unsigned)i * 4u + (unsigned)q;
    const unsigned y = idx / (unsigned)P.W, x = idx - y * (unsigned)P.W;
    int cx = 0, cy = 0;
#pragma unroll
    for (int s = 1; s < 16; ++s) // (unused entries are INT_MAX; equal starts = empty pieces: the last one wins)
    {
      if ((int)x >= P.zs.lo_x[s]) cx = s;
      if ((int)y >= P.zs.lo_y[s]) cy = s;
    }
    const int g = cx + P.zs.div_x * cy;
    int best = -1, bp = -1;
    float bd = 0.f;
    unsigned bc = 0u;
#pragma unroll
    for (int l = 0; l < NR; ++l)
      if (l < P.size && ((cover >> l) & 1u))
      {
        const float dl = (&d[l].x)[q];
        if (!(dl > 1.f))
        {
          const int p = P.zs.pos[g][l];
          if (best < 0 || dl < bd || (dl == bd && p > bp))
          {
            best = l; bd = dl; bp = p; bc = (&c[l].x)[q];
          }
        }
      }
    if (best < 0)
    {
      bc = 0u; bd = 1.001f; // (outside its rectangle a layer is colour 0, depth 1.001)
#pragma unroll
      for (int l = 0; l < NR; ++l)
        if (l == g && ((cover >> l) & 1u)) { bc = (&c[l].x)[q]; bd = (&d[l].x)[q]; }
    }
    oc[q] = bc;
    od[q] = bd;
  }
  f = make_uint4(oc[0], oc[1], oc[2], oc[3]);
  fd = make_float4(od[0], od[1], od[2], od[3]);
}

template <int NR, bool TO_CANVAS, bool ZBUF>
__global__ void __launch_bounds__(256) fold_p2p_kernel(const __grid_constant__ FoldP2PParams P)
{
  Flags* my_flags = reinterpret_cast<Flags*>(P.peers[P.rank] + P.off_flags);
  const int par = P.epoch & 1;
  // ---- announce: my image for this epoch is complete (previous kernel on this stream wrote it -- into
  // my own arena, or, pushed, into the owners' receive slots)
  if (blockIdx.x == 0 && threadIdx.x < P.size)
  {
    Flags* f = reinterpret_cast<Flags*>(P.peers[threadIdx.x] + P.off_flags);
#pragma unroll
    for (int k = 0; k < 4; ++k) ((volatile int*)f->img_rect[par][P.rank])[k] = P.rect[k];
    ((volatile int*)f->img_pushed[par])[P.rank] = P.pushed;
    __threadfence_system();
    st_release_sys(&f->ready[P.rank], P.epoch);
  }
  if (P.timeline && blockIdx.x == 0 && threadIdx.x == 0) my_flags->timeline[0] = global_ns();
  // ---- wait until every rank's image is complete (bounded: a peer that bailed out must not hang us)
  const bool go = wait_all_ready(err_of(my_flags), my_flags->ready, my_flags->aborted[kPathImage], kPathImage, P.size,
                                 P.epoch, P.timeout_ns);

  if (P.timeline && blockIdx.x == 0 && threadIdx.x == 0) my_flags->timeline[1] = global_ns();
  // ---- everything this CTA needs from the flag block, fetched by different threads at once (these are
  // system-coherent loads of ~1 us each: one after the other they used to cost more than the fold)
  __shared__ int s_rect[kMaxRanks][4]; // in fold order
  __shared__ int s_pushed[kMaxRanks];  // in fold order
  __shared__ int s_clean[2][4];        // rank 0: [0] result image of this parity, [1] canvas
  if (threadIdx.x < P.size * 4)
  {
    const int l = threadIdx.x >> 2, k = threadIdx.x & 3;
    s_rect[l][k] = go ? ((volatile int*)my_flags->img_rect[par][P.order[l]])[k] : 0; // aborted: nothing to fold
  }
  else if (threadIdx.x >= 64 && threadIdx.x < 64 + P.size)
    s_pushed[threadIdx.x - 64] = ((volatile int*)my_flags->img_pushed[par])[P.order[threadIdx.x - 64]];
  else if (threadIdx.x >= 96 && threadIdx.x < 104 && P.rank == 0)
  {
    const int k = threadIdx.x - 96;
    s_clean[k >> 2][k & 3] = k < 4 ? ((volatile int*)my_flags->clean_res[par])[k] : ((volatile int*)my_flags->clean_canvas)[k - 4];
  }
  __syncthreads();

  // a layer that was PUSHED sits in my own arena, in the receive slot of its source rank, indexed by
  // my local chunk number; one that was not is pulled out of its owner's arena at the global index
  const uint4* layer_rgba[NR];
  const float4* layer_depth[NR];
  unsigned local_mask = 0;
#pragma unroll
  for (int l = 0; l < NR; ++l)
    if (l < P.size)
    {
      const int src = P.order[l];
      if (s_pushed[l])
      {
        local_mask |= 1u << l;
        layer_rgba[l] = reinterpret_cast<const uint4*>(P.peers[P.rank] + P.off_recv_rgba) + (size_t)src * P.share_groups;
        layer_depth[l] = reinterpret_cast<const float4*>(P.peers[P.rank] + P.off_recv_depth) + (size_t)src * P.share_groups;
      }
      else
      {
        layer_rgba[l] = reinterpret_cast<const uint4*>(P.peers[src] + P.off_img_rgba);
        layer_depth[l] = reinterpret_cast<const float4*>(P.peers[src] + P.off_img_depth);
      }
    }
  uint4* out_rgba = reinterpret_cast<uint4*>(P.peers[0] + P.off_res_rgba);
  float4* out_depth = reinterpret_cast<float4*>(P.peers[0] + P.off_res_depth);

  // when W % 4 != 0 the host announces an unbounded rectangle, so x/y below are never compared
  const size_t n4 = (P.n_pixels + 3) / 4;
  const int w4 = P.W >= 4 ? P.W / 4 : 1;
  const size_t n_chunks = (n4 + kChunkGroups - 1) / kChunkGroups;

  // group index -> (row, first column); 32-bit arithmetic (n_pixels < 2^31): a 64-bit divide per group
  // used to cost more than the fold itself
  const unsigned w4u = (unsigned)w4;
  auto coverage = [&](unsigned y, unsigned x) -> unsigned {
    unsigned cover = 0;
#pragma unroll
    for (int l = 0; l < NR; ++l)
      if (l < P.size && (int)y >= s_rect[l][1] && (int)y < s_rect[l][3] && (int)x >= s_rect[l][0] && (int)x < s_rect[l][2])
        cover |= 1u << l;
    return cover;
  };
  // ---- rank 0: what may be dirty.  The previous exchange left the result image of this parity (and the
  // canvas) cleared outside the rectangles in the flags; if the host vouches that nobody wrote them
  // since, only groups inside those rectangles that no rank covers now need the cleared value again.
  __shared__ int s_dirty[2][4]; // [0] result image, [1] canvas
  __shared__ int s_union[4];    // bounding box of this frame's rectangles
  if (P.rank == 0 && threadIdx.x == 0)
  {
    int u[4] = { 0x7fffffff, 0x7fffffff, 0, 0 };
    for (int l = 0; l < P.size; ++l)
      if (s_rect[l][2] > s_rect[l][0] && s_rect[l][3] > s_rect[l][1])
      {
        u[0] = min(u[0], s_rect[l][0]); u[1] = min(u[1], s_rect[l][1]);
        u[2] = max(u[2], s_rect[l][2]); u[3] = max(u[3], s_rect[l][3]);
      }
    if (u[2] <= u[0]) u[0] = u[1] = u[2] = u[3] = 0;
    for (int k = 0; k < 4; ++k)
    {
      s_union[k] = u[k];
      s_dirty[0][k] = P.track_res ? s_clean[0][k] : (k < 2 ? 0 : 0x7fffffff);
      s_dirty[1][k] = P.track_canvas ? s_clean[1][k] : (k < 2 ? 0 : 0x7fffffff);
    }
  }
  __syncthreads();
  auto write_empty = [&](size_t i, unsigned uy, unsigned ux) {
    const int y = (int)uy, x = (int)ux;
    if (y >= s_dirty[0][1] && y < s_dirty[0][3] && x >= s_dirty[0][0] && x < s_dirty[0][2])
    {
      out_rgba[i] = make_uint4(0u, 0u, 0u, 0u);
      // N >= 2 all-clear layers fold to min(min(1.001,1.001), ...) = 1.001; one layer is copied
      out_depth[i] = make_float4(1.001f, 1.001f, 1.001f, 1.001f);
    }
    if (TO_CANVAS && y >= s_dirty[1][1] && y < s_dirty[1][3] && x >= s_dirty[1][0] && x < s_dirty[1][2])
    {
#pragma unroll
      for (int k = 0; k < 4; ++k) P.canvas_rgba[4 * i + k] = make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(P.canvas_depth)[i] = make_float4(1.001f, 1.001f, 1.001f, 1.001f);
    }
  };

  if (P.timeline && blockIdx.x == 0 && threadIdx.x == 0) my_flags->timeline[8] = global_ns();
  // ---- my chunks: rank, rank + size, rank + 2 size, ... dealt to the CTAs one after the other
  const size_t n_mine = n_chunks > (size_t)P.rank ? (n_chunks - 1 - (size_t)P.rank) / (size_t)P.size + 1 : 0;
  for (size_t k = blockIdx.x; k < n_mine; k += gridDim.x)
  {
    const size_t chunk = k * (size_t)P.size + (size_t)P.rank;
    const size_t i = chunk * kChunkGroups + threadIdx.x;
    if (i >= n4) continue;
    const unsigned gy = (unsigned)i / w4u, gx = ((unsigned)i - gy * w4u) * 4u;
    const unsigned cover = coverage(gy, gx);
    if (cover == 0)
    {
      if (P.rank == 0) write_empty(i, gy, gx);
      continue;
    }
    // issue all covering layer loads first (independent 16-byte NVLink reads)
    uint4 c[NR];
    float4 d[NR];
#pragma unroll
    for (int l = 0; l < NR; ++l)
      if (l < P.size)
      {
        if (cover & (1u << l))
        {
          const size_t at = (local_mask & (1u << l)) ? k * kChunkGroups + threadIdx.x : i;
          c[l] = layer_rgba[l][at];
          d[l] = layer_depth[l][at];
        }
        else
        {
          c[l] = make_uint4(0u, 0u, 0u, 0u);
          d[l] = make_float4(1.001f, 1.001f, 1.001f, 1.001f);
        }
      }
    uint4 f = c[0];
    float4 fd = d[0];
    if (ZBUF) zselect_group<NR>(P, c, d, cover, i, f, fd);
#pragma unroll
    for (int l = 1; l < NR; ++l)
      if (!ZBUF && l < P.size)
      {
        if (cover & (1u << l))
        {
          f.x = blend_u8x4(f.x, c[l].x); f.y = blend_u8x4(f.y, c[l].y);
          f.z = blend_u8x4(f.z, c[l].z); f.w = blend_u8x4(f.w, c[l].w);
        }
        fd.x = blend_depth(fd.x, d[l].x); fd.y = blend_depth(fd.y, d[l].y);
        fd.z = blend_depth(fd.z, d[l].z); fd.w = blend_depth(fd.w, d[l].w);
      }
    out_rgba[i] = f;
    out_depth[i] = fd;
  }
  if (P.timeline && blockIdx.x == 0 && threadIdx.x == 0) my_flags->timeline[9] = global_ns();
  // ---- rank 0 also writes the groups NO rank covers inside the other ranks' chunks (their owners skip
  // them); streaming stores into local HBM while the peers' folded pixels are still in flight
  if (P.rank == 0 && P.size > 1)
  {
    // only the rows the dirty rectangles span
    const int y_lo = min(s_dirty[0][1], TO_CANVAS ? s_dirty[1][1] : 0x7fffffff);
    const int y_hi = max(s_dirty[0][3], TO_CANVAS ? s_dirty[1][3] : 0);
    const size_t c_lo = y_lo <= 0 ? 0 : ((size_t)y_lo * (size_t)w4) / kChunkGroups;
    const size_t c_end = y_hi >= (int)(n4 / (size_t)w4) ? n_chunks
                                                          : min(n_chunks, ((size_t)y_hi * (size_t)w4) / kChunkGroups + 1);
    for (size_t chunk = c_lo + blockIdx.x; chunk < c_end; chunk += gridDim.x)
    {
      if (chunk % (size_t)P.size == 0) continue;
      const size_t i = chunk * kChunkGroups + threadIdx.x;
      if (i >= n4) continue;
      const unsigned gy = (unsigned)i / w4u, gx = ((unsigned)i - gy * w4u) * 4u;
      if (coverage(gy, gx) == 0) write_empty(i, gy, gx);
    }
  }

  // ---- last CTA out tells rank 0 that my range has landed
  __syncthreads();
  if (threadIdx.x == 0)
  {
    if (P.timeline && blockIdx.x == 0) my_flags->timeline[10] = global_ns();
    __threadfence_system();
    const unsigned prev = atomicAdd(&my_flags->cta_done, 1u);
    if (P.timeline && blockIdx.x == 0) my_flags->timeline[11] = global_ns();
    if (prev == gridDim.x - 1)
    {
      my_flags->cta_done = 0;
      if (P.rank == 0)
      {
        // every CTA has read the old rectangles: what this frame leaves cleared
        for (int k = 0; k < 4; ++k)
        {
          my_flags->clean_res[par][k] = s_union[k];
          if (TO_CANVAS) my_flags->clean_canvas[k] = s_union[k];
        }
      }
      Flags* root = reinterpret_cast<Flags*>(P.peers[0] + P.off_flags);
      __threadfence_system();
      st_release_sys(&root->done[P.rank], P.epoch);
      if (P.timeline) my_flags->timeline[2] = global_ns();
    }
  }
}

// The same exchange with CTAs of 128 threads and <= 72 registers: exactly the footprint of ONE sampler CTA.
// The exchange of frame k runs (exchange stream, highest priority) while the sampler -- a persistent grid that
// holds every register of every SM until its tile counter runs dry -- is busy with frame k+1.  A 256-thread x
// 128-register CTA (fold_p2p_kernel<8>) needs the slots of four sampler CTAs of one SM to be free at once,
// which only happens at the very end of a trace; this one is placed as soon as ANY sampler CTA of any launch
// retires.  The layers are folded in batches of four (8 + 8 registers each in flight) instead of all at once.
// Same ownership, same fold order and operator as fold_p2p_kernel: the same bits.
template <int kLightThreads, bool TO_CANVAS>
__global__ void __launch_bounds__(kLightThreads, kLightThreads == 128 ? 7 : 3) fold_p2p_light_kernel(const __grid_constant__ FoldP2PParams P)
{
  Flags* my_flags = reinterpret_cast<Flags*>(P.peers[P.rank] + P.off_flags);
  const int par = P.epoch & 1;
  if (blockIdx.x == 0 && threadIdx.x < P.size)
  {
    Flags* f = reinterpret_cast<Flags*>(P.peers[threadIdx.x] + P.off_flags);
#pragma unroll
    for (int k = 0; k < 4; ++k) ((volatile int*)f->img_rect[par][P.rank])[k] = P.rect[k];
    ((volatile int*)f->img_pushed[par])[P.rank] = P.pushed;
    __threadfence_system();
    st_release_sys(&f->ready[P.rank], P.epoch);
  }
  if (P.timeline && blockIdx.x == 0 && threadIdx.x == 0) my_flags->timeline[0] = global_ns();
  const bool go = wait_all_ready(err_of(my_flags), my_flags->ready, my_flags->aborted[kPathImage], kPathImage, P.size,
                                 P.epoch, P.timeout_ns);
  if (P.timeline && blockIdx.x == 0 && threadIdx.x == 0) my_flags->timeline[1] = global_ns();

  __shared__ int s_rect[kMaxRanks][4]; // in fold order
  __shared__ int s_pushed[kMaxRanks];  // in fold order
  __shared__ int s_clean[2][4];        // rank 0: [0] result image of this parity, [1] canvas
  __shared__ const uint4* s_lrgba[kMaxRanks];
  __shared__ const float4* s_ldepth[kMaxRanks];
  if (threadIdx.x < P.size * 4)
  {
    const int l = threadIdx.x >> 2, k = threadIdx.x & 3;
    s_rect[l][k] = go ? ((volatile int*)my_flags->img_rect[par][P.order[l]])[k] : 0;
  }
  else if (threadIdx.x >= 64 && threadIdx.x < 64 + P.size)
    s_pushed[threadIdx.x - 64] = ((volatile int*)my_flags->img_pushed[par])[P.order[threadIdx.x - 64]];
  else if (threadIdx.x >= 96 && threadIdx.x < 104 && P.rank == 0)
  {
    const int k = threadIdx.x - 96;
    s_clean[k >> 2][k & 3] = k < 4 ? ((volatile int*)my_flags->clean_res[par])[k] : ((volatile int*)my_flags->clean_canvas)[k - 4];
  }
  __syncthreads();
  if (threadIdx.x < P.size)
  {
    const int l = threadIdx.x, src = P.order[l];
    if (s_pushed[l])
    {
      s_lrgba[l] = reinterpret_cast<const uint4*>(P.peers[P.rank] + P.off_recv_rgba) + (size_t)src * P.share_groups;
      s_ldepth[l] = reinterpret_cast<const float4*>(P.peers[P.rank] + P.off_recv_depth) + (size_t)src * P.share_groups;
    }
    else
    {
      s_lrgba[l] = reinterpret_cast<const uint4*>(P.peers[src] + P.off_img_rgba);
      s_ldepth[l] = reinterpret_cast<const float4*>(P.peers[src] + P.off_img_depth);
    }
  }
  uint4* out_rgba = reinterpret_cast<uint4*>(P.peers[0] + P.off_res_rgba);
  float4* out_depth = reinterpret_cast<float4*>(P.peers[0] + P.off_res_depth);

  const size_t n4 = (P.n_pixels + 3) / 4;
  const int w4 = P.W >= 4 ? P.W / 4 : 1;
  const size_t n_chunks = (n4 + kChunkGroups - 1) / kChunkGroups;
  const unsigned w4u = (unsigned)w4;
  const int size = P.size;
  auto coverage = [&](unsigned y, unsigned x) -> unsigned {
    unsigned cover = 0;
    for (int l = 0; l < size; ++l)
      if ((int)y >= s_rect[l][1] && (int)y < s_rect[l][3] && (int)x >= s_rect[l][0] && (int)x < s_rect[l][2])
        cover |= 1u << l;
    return cover;
  };
  __shared__ int s_dirty[2][4]; // [0] result image, [1] canvas
  __shared__ int s_union[4];    // bounding box of this frame's rectangles
  if (P.rank == 0 && threadIdx.x == 0)
  {
    int u[4] = { 0x7fffffff, 0x7fffffff, 0, 0 };
    for (int l = 0; l < P.size; ++l)
      if (s_rect[l][2] > s_rect[l][0] && s_rect[l][3] > s_rect[l][1])
      {
        u[0] = min(u[0], s_rect[l][0]); u[1] = min(u[1], s_rect[l][1]);
        u[2] = max(u[2], s_rect[l][2]); u[3] = max(u[3], s_rect[l][3]);
      }
    if (u[2] <= u[0]) u[0] = u[1] = u[2] = u[3] = 0;
    for (int k = 0; k < 4; ++k)
    {
      s_union[k] = u[k];
      s_dirty[0][k] = P.track_res ? s_clean[0][k] : (k < 2 ? 0 : 0x7fffffff);
      s_dirty[1][k] = P.track_canvas ? s_clean[1][k] : (k < 2 ? 0 : 0x7fffffff);
    }
  }
  __syncthreads();
  auto write_empty = [&](size_t i, unsigned uy, unsigned ux) {
    const int y = (int)uy, x = (int)ux;
    if (y >= s_dirty[0][1] && y < s_dirty[0][3] && x >= s_dirty[0][0] && x < s_dirty[0][2])
    {
      out_rgba[i] = make_uint4(0u, 0u, 0u, 0u);
      out_depth[i] = make_float4(1.001f, 1.001f, 1.001f, 1.001f);
    }
    if (TO_CANVAS && y >= s_dirty[1][1] && y < s_dirty[1][3] && x >= s_dirty[1][0] && x < s_dirty[1][2])
    {
#pragma unroll
      for (int k = 0; k < 4; ++k) P.canvas_rgba[4 * i + k] = make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(P.canvas_depth)[i] = make_float4(1.001f, 1.001f, 1.001f, 1.001f);
    }
  };

  if (P.timeline && blockIdx.x == 0 && threadIdx.x == 0) my_flags->timeline[8] = global_ns();
  // ---- my chunks (rank, rank + size, ...), dealt to the CTAs one after the other; a thread folds two of
  // the chunk's 256 four-pixel groups
  const size_t n_mine = n_chunks > (size_t)P.rank ? (n_chunks - 1 - (size_t)P.rank) / (size_t)P.size + 1 : 0;
  for (size_t k = blockIdx.x; k < n_mine; k += gridDim.x)
  {
    const size_t chunk = k * (size_t)P.size + (size_t)P.rank;
#pragma unroll 1
    for (int h = 0; h < kChunkGroups / kLightThreads; ++h)
    {
      const unsigned g = threadIdx.x + h * kLightThreads;
      const size_t i = chunk * kChunkGroups + g;
      if (i >= n4) continue;
      const unsigned gy = (unsigned)i / w4u, gx = ((unsigned)i - gy * w4u) * 4u;
      const unsigned cover = coverage(gy, gx);
      if (cover == 0)
      {
        if (P.rank == 0) write_empty(i, gy, gx);
        continue;
      }
      const size_t at_local = k * kChunkGroups + g;
      uint4 f = make_uint4(0u, 0u, 0u, 0u);
      float4 fd = make_float4(1.001f, 1.001f, 1.001f, 1.001f);
      for (int l0 = 0; l0 < size; l0 += 4)
      {
        uint4 c[4];
        float4 d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
        {
          const int l = l0 + u;
          if (l < size && (cover & (1u << l)))
          {
            const size_t at = s_pushed[l] ? at_local : i;
            c[u] = s_lrgba[l][at];
            d[u] = s_ldepth[l][at];
          }
          else
          {
            c[u] = make_uint4(0u, 0u, 0u, 0u);
            d[u] = make_float4(1.001f, 1.001f, 1.001f, 1.001f);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
        {
          const int l = l0 + u;
          if (l >= size) break;
          if (l == 0) { f = c[0]; fd = d[0]; continue; }
          if (cover & (1u << l))
          {
            f.x = blend_u8x4(f.x, c[u].x); f.y = blend_u8x4(f.y, c[u].y);
            f.z = blend_u8x4(f.z, c[u].z); f.w = blend_u8x4(f.w, c[u].w);
          }
          fd.x = blend_depth(fd.x, d[u].x); fd.y = blend_depth(fd.y, d[u].y);
          fd.z = blend_depth(fd.z, d[u].z); fd.w = blend_depth(fd.w, d[u].w);
        }
      }
      out_rgba[i] = f;
      out_depth[i] = fd;
    }
  }
  if (P.timeline && blockIdx.x == 0 && threadIdx.x == 0) my_flags->timeline[9] = global_ns();
  // ---- rank 0: the groups NO rank covers inside the other ranks' chunks, where they may be dirty
  if (P.rank == 0 && P.size > 1)
  {
    const int y_lo = min(s_dirty[0][1], TO_CANVAS ? s_dirty[1][1] : 0x7fffffff);
    const int y_hi = max(s_dirty[0][3], TO_CANVAS ? s_dirty[1][3] : 0);
    const size_t c_lo = y_lo <= 0 ? 0 : ((size_t)y_lo * (size_t)w4) / kChunkGroups;
    const size_t c_end = y_hi >= (int)(n4 / (size_t)w4) ? n_chunks
                                                          : min(n_chunks, ((size_t)y_hi * (size_t)w4) / kChunkGroups + 1);
    for (size_t chunk = c_lo + blockIdx.x; chunk < c_end; chunk += gridDim.x)
    {
      if (chunk % (size_t)P.size == 0) continue;
      for (int h = 0; h < kChunkGroups / kLightThreads; ++h)
      {
        const size_t i = chunk * kChunkGroups + threadIdx.x + h * kLightThreads;
        if (i >= n4) continue;
        const unsigned gy = (unsigned)i / w4u, gx = ((unsigned)i - gy * w4u) * 4u;
        if (coverage(gy, gx) == 0) write_empty(i, gy, gx);
      }
    }
  }

  // ---- last CTA out tells rank 0 that my range has landed
  __syncthreads();
  if (threadIdx.x == 0)
  {
    if (P.timeline && blockIdx.x == 0) my_flags->timeline[10] = global_ns();
    __threadfence_system();
    const unsigned prev = atomicAdd(&my_flags->cta_done, 1u);
    if (P.timeline && blockIdx.x == 0) my_flags->timeline[11] = global_ns();
    if (prev == gridDim.x - 1)
    {
      my_flags->cta_done = 0;
      if (P.rank == 0)
      {
        for (int k = 0; k < 4; ++k)
        {
          my_flags->clean_res[par][k] = s_union[k];
          if (TO_CANVAS) my_flags->clean_canvas[k] = s_union[k];
        }
      }
      Flags* root = reinterpret_cast<Flags*>(P.peers[0] + P.off_flags);
      __threadfence_system();
      st_release_sys(&root->done[P.rank], P.epoch);
      if (P.timeline) my_flags->timeline[2] = global_ns();
    }
  }
}

// rank 0, after every rank's range has landed: ImageToCanvas for the groups some rank covers
template <int THREADS>
__global__ void __launch_bounds__(THREADS) covered_to_canvas_kernel(const __grid_constant__ FoldP2PParams P)
{
  const Flags* my_flags = reinterpret_cast<const Flags*>(P.peers[0] + P.off_flags);
  const int par = P.epoch & 1;
  if (P.timeline && blockIdx.x == 0 && threadIdx.x == 0) const_cast<Flags*>(my_flags)->timeline[3] = global_ns();
  // every rank's folded range has landed in my result image (what wait_done_kernel does, without
  // the extra launch)
  if (threadIdx.x < P.size && !wait_epoch(&my_flags->done[threadIdx.x], P.epoch, P.timeout_ns))
    report_error(err_of(const_cast<Flags*>(my_flags)), P.epoch, kPathImage, 1u, threadIdx.x);
  __syncthreads();
  if (P.timeline && blockIdx.x == 0 && threadIdx.x == 0) const_cast<Flags*>(my_flags)->timeline[4] = global_ns();
  __shared__ int s_rect[kMaxRanks][4];
  __shared__ int s_box[4];
  if (threadIdx.x < P.size * 4) s_rect[threadIdx.x >> 2][threadIdx.x & 3] = my_flags->img_rect[par][threadIdx.x >> 2][threadIdx.x & 3];
  __syncthreads();
  const int w4 = P.W >= 4 ? P.W / 4 : 1;
  const size_t n4 = (P.n_pixels + 3) / 4;
  const int h_img = (int)(n4 / (size_t)w4) + ((n4 % (size_t)w4) ? 1 : 0);
  if (threadIdx.x == 0)
  {
    // bounding box of the rectangles, in groups of 4 pixels: nothing outside of it is covered
    int u[4] = { 0x7fffffff, 0x7fffffff, 0, 0 };
    for (int l = 0; l < P.size; ++l)
      if (s_rect[l][2] > s_rect[l][0] && s_rect[l][3] > s_rect[l][1])
      {
        u[0] = min(u[0], s_rect[l][0]); u[1] = min(u[1], s_rect[l][1]);
        u[2] = max(u[2], s_rect[l][2]); u[3] = max(u[3], s_rect[l][3]);
      }
    if (u[2] <= u[0]) u[0] = u[1] = u[2] = u[3] = 0;
    s_box[0] = max(u[0], 0) / 4; s_box[1] = max(u[1], 0);
    s_box[2] = min((min(u[2], P.W) + 3) / 4, w4); s_box[3] = min(u[3], h_img);
    if (P.W % 4 != 0) { s_box[0] = 0; s_box[1] = 0; s_box[2] = w4; s_box[3] = h_img; } // (unbounded rectangles)
  }
  __syncthreads();
  const uint4* res_rgba = reinterpret_cast<const uint4*>(P.peers[0] + P.off_res_rgba);
  const float4* res_depth = reinterpret_cast<const float4*>(P.peers[0] + P.off_res_depth);
  const float k = 1.f / 255.f;
  const int bw = s_box[2] - s_box[0], bh = s_box[3] - s_box[1];
  const unsigned n_box = bw > 0 && bh > 0 ? (unsigned)bw * (unsigned)bh : 0u;
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < n_box; t += stride)
  {
    const unsigned by = t / (unsigned)bw;
    const int y = s_box[1] + (int)by, x4 = s_box[0] + (int)(t - by * (unsigned)bw), x = x4 * 4;
    const size_t i = (size_t)y * (size_t)w4 + (size_t)x4;
    if (i >= n4) continue;
    bool cover = false;
    for (int l = 0; l < P.size; ++l)
      cover = cover || (y >= s_rect[l][1] && y < s_rect[l][3] && x >= s_rect[l][0] && x < s_rect[l][2]);
    if (!cover) continue;
    const uint4 c = res_rgba[i];
    const unsigned w[4] = { c.x, c.y, c.z, c.w };
#pragma unroll
    for (int q = 0; q < 4; ++q)
      P.canvas_rgba[4 * i + q] = make_float4((float)(w[q] & 0xffu) * k, (float)((w[q] >> 8) & 0xffu) * k,
                                             (float)((w[q] >> 16) & 0xffu) * k, (float)(w[q] >> 24) * k);
    reinterpret_cast<float4*>(P.canvas_depth)[i] = res_depth[i];
  }
  if (P.timeline && blockIdx.x == 0 && threadIdx.x == 0) const_cast<Flags*>(my_flags)->timeline[5] = global_ns();
}

__global__ void wait_done_kernel(Flags* mine, const unsigned int* done, int size, unsigned int epoch, unsigned path,
                                 unsigned long long timeout_ns)
{
  if (threadIdx.x < size && !wait_epoch(done + threadIdx.x, epoch, timeout_ns, 64))
    report_error(err_of(mine), epoch, path, 1u, threadIdx.x);
}

// A rank that cannot take part in exchange <epoch> (rank-local error: list too long for the arena, too
// many layers, ...) still has to release its peers: it raises its ready flag for the epoch together
// with an abort mark, and reports "done" to rank 0.  The peers' kernels see the mark, skip the work and
// leave an error for their hosts (vr_status of the next synchronising call).
__global__ void abort_announce_kernel(unsigned char* const* peers, int rank, int size, size_t off_flags, unsigned path,
                                      unsigned int epoch)
{
  const int t = threadIdx.x;
  if (t < size)
  {
    Flags* f = reinterpret_cast<Flags*>(peers[t] + off_flags);
    ((volatile unsigned int*)f->aborted[path])[rank] = epoch;
    __threadfence_system();
    st_release_sys(path == kPathImage ? &f->ready[rank] : &f->p_ready[rank], epoch);
  }
  if (t == 0)
  {
    Flags* root = reinterpret_cast<Flags*>(peers[0] + off_flags);
    __threadfence_system();
    st_release_sys(path == kPathImage ? &root->done[rank] : &root->p_done[rank], epoch);
    report_error(err_of(reinterpret_cast<Flags*>(peers[rank] + off_flags)), epoch, path, 2u, rank);
  }
}


// ------------------------------------------------------------------ Scene::SynchDepths (Scene.cpp:249-264)
// MPI_Bcast of rank 0's canvas depth: rank 0 stages its depth buffer in its arena and raises
// s_ready; every other rank pulls it over NVLink into its own canvas and reports s_done, which
// rank 0 waits on before it stages the next one.
__global__ void sync_post_kernel(Flags* root_flags, unsigned int epoch)
{
  if (threadIdx.x == 0)
  {
    __threadfence_system();
    root_flags->s_done[0] = epoch;
    st_release_sys(&root_flags->s_ready, epoch);
  }
}
// (`staged` is rank 0's peer memory, rewritten by rank 0 for every broadcast: no const/__restrict__,
// so that the loads cannot be turned into non-coherent ones)
__global__ void __launch_bounds__(256) sync_pull_kernel(Flags* root_flags, Flags* my_flags, float* staged,
                                                        float* depth, size_t n, int rank, unsigned int epoch,
                                                        unsigned long long timeout_ns)
{
  if (threadIdx.x == 0 && !wait_epoch(&root_flags->s_ready, epoch, timeout_ns, 128))
    report_error(err_of(my_flags), epoch, kPathImage, 1u, 0u);
  __syncthreads();
  const size_t n4 = n / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = t; i < n4; i += stride)
    reinterpret_cast<float4*>(depth)[i] = __ldcv(reinterpret_cast<const float4*>(staged) + i);
  for (size_t i = n4 * 4 + t; i < n; i += stride) depth[i] = __ldcv(staged + i);
  __syncthreads();
  if (threadIdx.x == 0)
  {
    __threadfence_system();
    const unsigned prev = atomicAdd(&my_flags->cta_done, 1u);
    if (prev == gridDim.x - 1)
    {
      my_flags->cta_done = 0;
      __threadfence_system();
      st_release_sys(&root_flags->s_done[rank], epoch);
    }
  }
}

// ------------------------------------------------------------------ partial path: pull + merge + fold
struct MergeP2PParams
{
  unsigned char* const* peers;
  int rank, size;
  unsigned int epoch;
  size_t n_pixels;
  size_t off_flags, off_poff, off_psorted, off_pout;
  size_t sorted_cap, out_cap;
  unsigned long long timeout_ns;
  // fused partials_to_canvas (frame starts from a cleared canvas): owners write rank 0's canvas
  size_t off_canvas_rgba, off_canvas_depth;
  ToCanvasParams tp;
};

__device__ __forceinline__ void partial_blend(vr_partial& a, const vr_partial& o)
{
  // VolumePartial::blend, VolumePartial.hpp:86-95
  if (a.alpha >= 1.f || o.alpha == 0.f) return;
  const float opacity = (1.f - a.alpha);
  a.rgb[0] += opacity * o.rgb[0];
  a.rgb[1] += opacity * o.rgb[1];
  a.rgb[2] += opacity * o.rgb[2];
  a.alpha += opacity * o.alpha;
  a.alpha = a.alpha > 1.f ? 1.f : a.alpha;
}

// vr_partial is 24 bytes, 8-byte aligned in the arena: three 8-byte loads
__device__ __forceinline__ vr_partial load_partial(const vr_partial* p)
{
  const uint2* q = reinterpret_cast<const uint2*>(p);
  const uint2 a = q[0], b = q[1], c = q[2];
  vr_partial r;
  r.pixel_id = (int)a.x; r.depth = __uint_as_float(a.y);
  r.rgb[0] = __uint_as_float(b.x); r.rgb[1] = __uint_as_float(b.y);
  r.rgb[2] = __uint_as_float(c.x); r.alpha = __uint_as_float(c.y);
  return r;
}

// One thread per owned pixel.  NR = number of ranks rounded up to a power of two: the per-source
// cursors live in registers (all loops over sources are fully unrolled).
template <int NR, bool TO_CANVAS>
__global__ void __launch_bounds__(256) merge_fold_p2p_kernel(const __grid_constant__ MergeP2PParams P)
{
  Flags* my_flags = reinterpret_cast<Flags*>(P.peers[P.rank] + P.off_flags);
  const int par = P.epoch & 1;
  // ---- announce my sorted list (written by the previous kernels on this stream), then wait for all
  if (blockIdx.x == 0 && threadIdx.x < P.size)
  {
    Flags* f = reinterpret_cast<Flags*>(P.peers[threadIdx.x] + P.off_flags);
    __threadfence_system();
    st_release_sys(&f->p_ready[P.rank], P.epoch);
  }
  const bool go = wait_all_ready(err_of(my_flags), my_flags->p_ready, my_flags->aborted[kPathPartials], kPathPartials,
                                 P.size, P.epoch, P.timeout_ns, 64);

  // ---- global pixel bounds and my range: RegularDecomposer<DiscreteBounds>, 1-D
  // (vtkh_diy_partial_redistribute.hpp:133-150, decomposition.hpp:37-46,648-666)
  int gmin = 0x7fffffff, gmax = -1;
  for (int r = 0; r < P.size && go; ++r) // (an aborted exchange folds nothing, but still reports "done")
  {
    const int lo = ((volatile int*)my_flags->p_minmax[par][r])[0];
    const int hi = ((volatile int*)my_flags->p_minmax[par][r])[1];
    if (hi >= 0) { gmin = min(gmin, lo); gmax = max(gmax, hi); }
  }
  long long lo_px = 0, hi_px = 0; // [lo, hi)
  if (gmax >= 0)
  {
    long long width = ((long long)gmax - gmin + 1) / P.size;
    if (width < 1) width = 1;
    lo_px = (long long)gmin + width * P.rank;
    hi_px = (P.rank == P.size - 1) ? (long long)gmax + 1 : lo_px + width;
    if (lo_px > (long long)gmax + 1) lo_px = (long long)gmax + 1;
    if (hi_px > (long long)gmax + 1) hi_px = (long long)gmax + 1;
  }

  const int* off[NR];
  const vr_partial* list[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r)
  {
    const int rr = r < P.size ? r : 0;
    off[r] = reinterpret_cast<const int*>(P.peers[rr] + P.off_poff);
    list[r] = reinterpret_cast<const vr_partial*>(P.peers[rr] + P.off_psorted);
  }
  Flags* root = reinterpret_cast<Flags*>(P.peers[0] + P.off_flags);
  vr_partial* out = reinterpret_cast<vr_partial*>(P.peers[0] + P.off_pout);

  __shared__ unsigned long long s_base;
  __shared__ int s_warp[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long base = lo_px + (long long)blockIdx.x * blockDim.x; base < hi_px; base += stride)
  {
    const long long px = base + threadIdx.x;
    int cur[NR], end[NR];
    float head[NR];
    int total = 0;
#pragma unroll
    for (int r = 0; r < NR; ++r)
    {
      cur[r] = end[r] = 0;
      if (r < P.size && px < hi_px)
      {
        cur[r] = px > 0 ? off[r][px - 1] : 0; // the arrays hold each pixel's END offset
        end[r] = off[r][px];
        if ((size_t)end[r] > P.sorted_cap) end[r] = cur[r]; // overflowed list: dropped, flagged by its owner
      }
      total += end[r] - cur[r];
    }
    vr_partial result;
    if (total > 0)
    {
#pragma unroll
      for (int r = 0; r < NR; ++r)
        head[r] = cur[r] < end[r] ? list[r][cur[r]].depth : 0.f;
      bool first = true;
      for (int k = 0; k < total; ++k)
      {
        // next partial in (depth, rank) order; within a rank the run is already ordered
        int sel = -1;
        float best = 0.f;
#pragma unroll
        for (int r = 0; r < NR; ++r)
          if (cur[r] < end[r] && (sel < 0 || head[r] < best)) { sel = r; best = head[r]; }
        vr_partial q;
#pragma unroll
        for (int r = 0; r < NR; ++r)
          if (r == sel)
          {
            q = load_partial(list[r] + cur[r]);
            cur[r] += 1;
            if (cur[r] < end[r]) head[r] = list[r][cur[r]].depth;
          }
        if (first) { result = q; first = false; }
        else partial_blend(result, q);
      }
    }
    if (TO_CANVAS)
    {
      // partials_to_canvas over a cleared canvas (in = 0), stored into rank 0's canvas: 20 bytes per
      // covered pixel over NVLink instead of a 24-byte partial, no list, no atomics
      if (total > 0)
      {
        float4 o;
        float d;
        partial_to_canvas(result, P.tp, make_float4(0.f, 0.f, 0.f, 0.f), o, d);
        reinterpret_cast<float4*>(P.peers[0] + P.off_canvas_rgba)[px] = o;
        reinterpret_cast<float*>(P.peers[0] + P.off_canvas_depth)[px] = d;
      }
      continue;
    }
    // block-aggregated append to rank 0's list: one NVLink atomic per CTA iteration
    const unsigned mask = __ballot_sync(0xffffffffu, total > 0);
    if (lane == 0) s_warp[warp] = __popc(mask);
    __syncthreads();
    if (threadIdx.x == 0)
    {
      int sum = 0;
      for (int w = 0; w < 8; ++w) { const int c = s_warp[w]; s_warp[w] = sum; sum += c; }
      s_base = sum ? atomicAdd(&root->p_out_count, (unsigned long long)sum) : 0ull;
    }
    __syncthreads();
    if (total > 0)
    {
      const unsigned long long slot = s_base + s_warp[warp] + __popc(mask & ((1u << lane) - 1u));
      if (slot < P.out_cap) out[slot] = result;
    }
    __syncthreads();
  }

  // ---- last CTA out tells rank 0 that my range has landed
  __syncthreads();
  if (threadIdx.x == 0)
  {
    __threadfence_system();
    const unsigned prev = atomicAdd(&my_flags->p_cta_done, 1u);
    if (prev == gridDim.x - 1)
    {
      my_flags->p_cta_done = 0;
      __threadfence_system();
      st_release_sys(&root->p_done[P.rank], P.epoch);
    }
  }
}

// publish {min,max} of my list into every peer's flag block (parity slot of this epoch) and reset
// rank 0's output counter; runs after the pixel sort, before the merge kernel, on the same stream
__global__ void publish_minmax_kernel(unsigned char* const* peers, int rank, int size, size_t off_flags,
                                      int par, const int* minmax, const unsigned long long* count_dev,
                                      size_t sorted_cap)
{
  const int t = threadIdx.x;
  if (t < size)
  {
    Flags* f = reinterpret_cast<Flags*>(peers[t] + off_flags);
    ((volatile int*)f->p_minmax[par][rank])[0] = minmax[0];
    ((volatile int*)f->p_minmax[par][rank])[1] = minmax[1];
  }
  if (t == 0)
  {
    Flags* mine = reinterpret_cast<Flags*>(peers[rank] + off_flags);
    if (*count_dev > sorted_cap) mine->p_overflow = 1u;
    if (rank == 0) mine->p_out_count = 0ull;
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Layout
{
  size_t off_flags, off_img_rgba[kImgRing], off_img_depth[kImgRing], off_res_rgba[2], off_res_depth[2];
  size_t off_poff[2], off_psorted[2], off_pout, off_canvas_rgba, off_canvas_depth, total;
  size_t off_lflags, off_ltab[kLayerRing], off_lpool_rgba[kLayerRing], off_lpool_depth[kLayerRing];
  size_t off_sync_depth;
  // receive ring of pushed frames: slot e % 3 holds, for every source rank, the pixels of frame e that
  // THIS rank owns (round-robin 1024-pixel chunks), written by the sources' samplers over NVLink
  size_t off_recv_rgba[kImgRing], off_recv_depth[kImgRing];
  // receive pools of pushed ray layers: buffer e % 3 holds, per source rank, max_partials entries at the
  // source's own pool indices -- only those whose screen tile this rank folds are ever written
  size_t off_lrecv_rgba[kLayerRing], off_lrecv_depth[kLayerRing];
};
Layout make_layout(size_t max_pixels, size_t max_partials, bool is_root, int n_ranks)
{
  Layout L;
  const size_t px = align_up(max_pixels, 64);
  size_t o = 0;
  L.off_flags = o; o += kFlagBytes;
  for (int b = 0; b < kImgRing; ++b) { L.off_img_rgba[b] = o; o += px * 4; }
  for (int b = 0; b < kImgRing; ++b) { L.off_img_depth[b] = o; o += px * 4; }
  for (int b = 0; b < 2; ++b) { L.off_res_rgba[b] = o; o += px * 4; }
  for (int b = 0; b < 2; ++b) { L.off_res_depth[b] = o; o += px * 4; }
  // `ranks` receive slots of ceil(chunks / ranks) whole chunks each: up to one extra chunk per rank
  const size_t px_recv = px + (size_t)kMaxRanks * 1024;
  for (int b = 0; b < kImgRing; ++b) { L.off_recv_rgba[b] = o; o += px_recv * 4; }
  for (int b = 0; b < kImgRing; ++b) { L.off_recv_depth[b] = o; o += px_recv * 4; }
  for (int b = 0; b < 2; ++b) { L.off_poff[b] = o; o += max_partials ? align_up(partial_scan_padded(max_pixels) * 4, 256) : 0; }
  for (int b = 0; b < 2; ++b) { L.off_psorted[b] = o; o += align_up(max_partials * sizeof(vr_partial), 256); }
  // dense ray layers (layers.cu): flags, layer tables and pools, double-buffered
  L.off_lflags = o; o += max_partials ? kFlagBytes : 0;
  for (int b = 0; b < kLayerRing; ++b) { L.off_ltab[b] = o; o += max_partials ? align_up(sizeof(LayerTable), 256) : 0; }
  for (int b = 0; b < kLayerRing; ++b) { L.off_lpool_rgba[b] = o; o += align_up(max_partials * sizeof(float4), 256); }
  for (int b = 0; b < kLayerRing; ++b) { L.off_lpool_depth[b] = o; o += align_up(max_partials * sizeof(float), 256); }
  const size_t recv_entries = n_ranks > 1 ? max_partials * (size_t)n_ranks : 0;
  for (int b = 0; b < kLayerRing; ++b) { L.off_lrecv_rgba[b] = o; o += align_up(recv_entries * sizeof(float4), 256); }
  for (int b = 0; b < kLayerRing; ++b) { L.off_lrecv_depth[b] = o; o += align_up(recv_entries * sizeof(float), 256); }
  // regions only rank 0 allocates; their OFFSETS are the same on every rank (peers address them)
  const size_t common_end = o;
  L.off_pout = o;
  if (max_partials) o += align_up(max_pixels * sizeof(vr_partial), 256); // <= 1 per pixel
  // rank 0's canvas for the fused partial path: owners store finished pixels straight into it
  L.off_canvas_rgba = o;
  if (max_partials) o += align_up(max_pixels * sizeof(float4), 256);
  L.off_canvas_depth = o;
  if (max_partials) o += align_up(max_pixels * sizeof(float), 256);
  // rank 0's staging copy of its canvas depth for the depth broadcast
  L.off_sync_depth = o;
  o += align_up(max_pixels * sizeof(float), 256);
  if (!is_root) o = common_end;
  L.total = align_up(o, 2 << 20);
  return L;
}

} // namespace

template <int NR>
static cudaError_t launch_fold_p2p_nr(const FoldP2PParams& p, int sm_count, cudaStream_t s)
{
  auto go = [&](auto kernel) {
    // One CTA per SM: the exchange overlaps the next frame's trace (exchange stream), so what counts is how
    // few SM resources it holds while its stores drain over NVLink -- and every CTA ends with one
    // system-scope fence, which is the expensive part -- not how fast it would run alone.
    const size_t n4 = (p.n_pixels + 3) / 4;
    const size_t n_chunks = (n4 + kChunkGroups - 1) / kChunkGroups;
    size_t grid = (size_t)sm_count * (size_t)(p.grid_per_sm > 0 ? p.grid_per_sm : 1);
    const size_t want = p.rank == 0 ? n_chunks : (n_chunks + p.size - 1) / p.size;
    if (grid > want) grid = want ? want : 1;
    if (p.max_ctas > 0 && grid > (size_t)p.max_ctas) grid = (size_t)p.max_ctas;
    kernel<<<(unsigned)grid, 256, 0, s>>>(p);
  };
  if (p.zbuffer) go(fold_p2p_kernel<NR, false, true>);
  else if (p.canvas_rgba) go(fold_p2p_kernel<NR, true, false>);
  else go(fold_p2p_kernel<NR, false, false>);
  return cudaGetLastError();
}

cudaError_t launch_fold_p2p(const FoldP2PParams& p, int sm_count, cudaStream_t s)
{
  if (p.light && !p.zbuffer) // (the z-select lives in fold_p2p_kernel only)
  {
    const size_t n4 = (p.n_pixels + 3) / 4;
    const size_t n_chunks = (n4 + kChunkGroups - 1) / kChunkGroups;
    size_t grid = (size_t)sm_count * (size_t)(p.grid_per_sm > 0 ? p.grid_per_sm : 1);
    const size_t want = p.rank == 0 ? n_chunks : (n_chunks + p.size - 1) / p.size;
    if (grid > want) grid = want ? want : 1;
    if (p.max_ctas > 0 && grid > (size_t)p.max_ctas) grid = (size_t)p.max_ctas;
    auto go = [&](auto k128, auto k256) {
      if (p.light == 1) k128<<<(unsigned)grid, 128, 0, s>>>(p);
      else k256<<<(unsigned)grid, 256, 0, s>>>(p);
    };
    if (p.canvas_rgba) go(fold_p2p_light_kernel<128, true>, fold_p2p_light_kernel<256, true>);
    else go(fold_p2p_light_kernel<128, false>, fold_p2p_light_kernel<256, false>);
    return cudaGetLastError();
  }
  if (p.force_nr8 && p.size <= 8) return launch_fold_p2p_nr<8>(p, sm_count, s);
  if (p.size <= 2) return launch_fold_p2p_nr<2>(p, sm_count, s);
  if (p.size <= 4) return launch_fold_p2p_nr<4>(p, sm_count, s);
  if (p.size <= 8) return launch_fold_p2p_nr<8>(p, sm_count, s);
  return launch_fold_p2p_nr<16>(p, sm_count, s);
}

void preload_comm_kernels()
{
  auto folds = [](auto nr) {
    constexpr int NR = decltype(nr)::value;
    preload_kernel(fold_p2p_kernel<NR, false, false>);
    preload_kernel(fold_p2p_kernel<NR, true, false>);
    preload_kernel(fold_p2p_kernel<NR, false, true>);
    preload_kernel(merge_fold_p2p_kernel<NR, true>);
    preload_kernel(merge_fold_p2p_kernel<NR, false>);
  };
  folds(std::integral_constant<int, 2>());
  folds(std::integral_constant<int, 4>());
  folds(std::integral_constant<int, 8>());
  folds(std::integral_constant<int, 16>());
  preload_kernel(merge_fold_p2p_kernel<1, true>);
  preload_kernel(merge_fold_p2p_kernel<1, false>);
  preload_kernel(fold_p2p_light_kernel<128, true>);
  preload_kernel(fold_p2p_light_kernel<128, false>);
  preload_kernel(fold_p2p_light_kernel<256, true>);
  preload_kernel(fold_p2p_light_kernel<256, false>);
  preload_kernel(covered_to_canvas_kernel<128>);
  preload_kernel(covered_to_canvas_kernel<256>);
  preload_kernel(wait_done_kernel);
  preload_kernel(abort_announce_kernel);
  preload_kernel(sync_post_kernel);
  preload_kernel(sync_pull_kernel);
  preload_kernel(publish_minmax_kernel);
}

void comm_destroy(vr_ctx* ctx)
{
  Comm& c = ctx->comm;
  // (the streams belong to the context, not to the arena: created in vr_create)
  if (c.xstream) { cudaStreamSynchronize(c.xstream); cudaStreamDestroy(c.xstream); c.xstream = nullptr; }
  for (int k = 0; k < Comm::kMaxTraceStreams; ++k)
  {
    if (c.tstream[k]) { cudaStreamSynchronize(c.tstream[k]); cudaStreamDestroy(c.tstream[k]); c.tstream[k] = nullptr; }
    if (c.ev_t[k]) cudaEventDestroy(c.ev_t[k]);
  }
  if (c.ev_trace) cudaEventDestroy(c.ev_trace);
  if (c.ev_main) cudaEventDestroy(c.ev_main);
  for (int k = 0; k < 8; ++k)
    if (c.ev_x[k]) cudaEventDestroy(c.ev_x[k]);
  if (!c.on) return;
  for (int r = 0; r < (int)c.peer.size(); ++r)
    if (r != c.rank && c.peer[r] && !c.local_peers) cudaIpcCloseMemHandle(c.peer[r]);
  cudaFree(c.peer_dev);
  cudaFree(c.minmax_dev);
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
  const Layout L = make_layout(c.max_pixels, c.max_partials, c.rank == 0, c.size);
  const int b = (c.epoch + 1) & 1;
  const int slot = (int)((c.epoch + 1) % kImgRing);
  ctx->img_rgba = reinterpret_cast<uchar4*>(c.arena + L.off_img_rgba[slot]);
  ctx->img_depth = reinterpret_cast<float*>(c.arena + L.off_img_depth[slot]);
  ctx->res_rgba = reinterpret_cast<uchar4*>(c.arena + L.off_res_rgba[b]);
  ctx->res_depth = reinterpret_cast<float*>(c.arena + L.off_res_depth[b]);
  if (c.rank == 0 && c.max_partials)
  {
    // rank 0's canvas lives in the arena so that peers can store finished pixels into it
    if (!ctx->canvas_in_arena)
    {
      cudaStreamSynchronize(ctx->stream);
      cudaFree(ctx->canvas_rgba);
      cudaFree(ctx->canvas_depth);
      ctx->canvas_in_arena = true;
    }
    ctx->canvas_rgba = reinterpret_cast<float4*>(c.arena + L.off_canvas_rgba);
    ctx->canvas_depth = reinterpret_cast<float*>(c.arena + L.off_canvas_depth);
    if (ctx->cap_pixels < c.max_pixels) ctx->cap_pixels = c.max_pixels;
  }
  return VR_OK;
}

// where an image traced one frame AHEAD of the pending exchange goes (VR_FRAME_AHEAD)
vr_status comm_ahead_image(vr_ctx* ctx, uchar4** rgba, float** depth)
{
  Comm& c = ctx->comm;
  if (!c.on)
  {
    ctx->err = "VR_FRAME_AHEAD needs the exchange arena (vr_comm_init)";
    return VR_ERR_STATE;
  }
  const Layout L = make_layout(c.max_pixels, c.max_partials, c.rank == 0, c.size);
  const int slot = (int)((c.epoch + 2) % kImgRing);
  *rgba = reinterpret_cast<uchar4*>(c.arena + L.off_img_rgba[slot]);
  *depth = reinterpret_cast<float*>(c.arena + L.off_img_depth[slot]);
  return VR_OK;
}

// vr_trace_to_image without VR_FRAME_WRITE_CANVAS, and the layer tracing calls without a canvas-depth
// clamp, touch only the new frame's ring slot / layer buffer: they may start while the latest exchange is
// still running, but not before the one before it has completed
void comm_join_previous_exchange(vr_ctx* ctx)
{
  Comm& c = ctx->comm;
  if (!c.xstream || c.xserial < 2) return;
  cudaStreamWaitEvent(ctx->stream, c.ev_x[(c.xserial - 1) & 7], 0);
}

// An image-only trace of image frame e (= epoch + 1, or + 2 when traced ahead) writes ring slot e % R
// (R = kImgRing) -- in this rank's arena, or, pushed, in every owner's receive ring -- which frame e - R
// used.  Every rank has finished reading frame e - R once THIS rank's exchange of frame e - R + 1 has
// completed: that exchange passed the all-ranks-ready barrier of e - R + 1, which a rank only joins after
// its own exchange of e - R (same stream).  So the trace waits for exchange e - R + 1 -- issued several
// calls ago, i.e. normally long done -- and the exchanges after it may still be running while it goes.
void comm_join_for_image_trace(vr_ctx* ctx, bool ahead, cudaStream_t s)
{
  Comm& c = ctx->comm;
  if (!c.on || !c.xstream) return;
  const unsigned int e = c.epoch + (ahead ? 2u : 1u);
  if (e < (unsigned)kImgRing) return;
  const unsigned int xs = c.x_of_img_epoch[(e - (kImgRing - 1)) & 7];
  if (xs == 0) return; // that exchange ran on the context's stream: ordered already
  // (an event slot is reused every 8 exchanges: if image and layer exchanges were mixed in between, fall
  // back to the latest exchange)
  const unsigned int w = (c.xserial - xs >= 8) ? c.xserial : xs;
  cudaStreamWaitEvent(s, c.ev_x[w & 7], 0);
}

// receive-slot geometry of a pushed frame (sampler mode 5)
vr_status comm_push_target(vr_ctx* ctx, bool ahead, int width, int height, TraceParams& p)
{
  Comm& c = ctx->comm;
  if (!c.on || !c.peer_dev)
  {
    ctx->err = "VR_FRAME_PUSH needs a connected exchange (vr_comm_init + vr_comm_connect)";
    return VR_ERR_STATE;
  }
  if ((size_t)width * height > c.max_pixels)
  {
    ctx->err = "image larger than the max_pixels given to vr_comm_init";
    return VR_ERR_INVALID;
  }
  const Layout L = make_layout(c.max_pixels, c.max_partials, c.rank == 0, c.size);
  const int slot = (int)((c.epoch + (ahead ? 2 : 1)) % kImgRing);
  const size_t n4 = ((size_t)width * height + 3) / 4;
  const size_t n_chunks = (n4 + kChunkGroups - 1) / kChunkGroups;
  p.push_peers = c.peer_dev;
  p.push_off_rgba = L.off_recv_rgba[slot];
  p.push_off_depth = L.off_recv_depth[slot];
  p.push_share_px = (unsigned int)(((n_chunks + c.size - 1) / c.size) * 1024);
  p.push_rank = c.rank;
  p.push_size = c.size;
  return VR_OK;
}

// diagnostics: the globaltimer stamps of the last image exchange (VR_TIMELINE=1) and where the sampler
// should leave its end-of-kernel stamp
unsigned long long* comm_timeline_slot(vr_ctx* ctx, int k)
{
  Comm& c = ctx->comm;
  if (!c.on || !c.timeline) return nullptr;
  const Layout L = make_layout(c.max_pixels, c.max_partials, c.rank == 0, c.size);
  return reinterpret_cast<unsigned long long*>(c.arena + L.off_flags + offsetof(Flags, timeline)) + k;
}

// what the exchange kernels left for the host (a wait that hit the time limit, a peer that aborted):
// reported once, by the next synchronising entry point
vr_status comm_check_errors(vr_ctx* ctx)
{
  Comm& c = ctx->comm;
  if (!c.on) return VR_OK;
  const Layout L = make_layout(c.max_pixels, c.max_partials, c.rank == 0, c.size);
  unsigned int e[4] = { 0, 0, 0, 0 }, le[4] = { 0, 0, 0, 0 };
  if (cudaMemcpy(e, c.arena + L.off_flags + offsetof(Flags, err_epoch), sizeof(e), cudaMemcpyDeviceToHost) != cudaSuccess)
    return VR_OK;
  if (c.max_partials)
    cudaMemcpy(le, c.arena + L.off_lflags + offsetof(LayerFlags, err), sizeof(le), cudaMemcpyDeviceToHost);
  const unsigned int* w = e[0] ? e : (le[0] ? le : nullptr);
  if (!w) return VR_OK;
  static const char* path[3] = { "image", "partial-list", "ray-layer" };
  char buf[240];
  snprintf(buf, sizeof(buf), "%s exchange %u did not complete on rank %d: %s (rank %u)", path[w[1] < 3 ? w[1] : 0], w[0],
           c.rank, w[2] == 2 ? "aborted by a rank-local error" : "timed out waiting for a peer", w[3]);
  ctx->err = buf;
  cudaMemset(c.arena + L.off_flags + offsetof(Flags, err_epoch), 0, sizeof(e));
  if (c.max_partials) cudaMemset(c.arena + L.off_lflags + offsetof(LayerFlags, err), 0, sizeof(le));
  return VR_ERR_STATE;
}

// point the context's layer table/pools at the arena halves of the NEXT layer frame's parity
vr_status comm_bind_layers(vr_ctx* ctx)
{
  Comm& c = ctx->comm;
  const Layout L = make_layout(c.max_pixels, c.max_partials, c.rank == 0, c.size);
  const int b = (int)((c.lepoch + 1) % kLayerRing);
  if (!ctx->layers_in_arena)
  {
    cudaStreamSynchronize(ctx->stream);
    for (int k = 0; k < vr_ctx::kOwnLayerRing; ++k)
    {
      cudaFree(ctx->own_ltab[k]); cudaFree(ctx->own_lpool_rgba[k]); cudaFree(ctx->own_lpool_depth[k]);
      ctx->own_ltab[k] = nullptr; ctx->own_lpool_rgba[k] = nullptr; ctx->own_lpool_depth[k] = nullptr;
      ctx->own_lpool_cap[k] = 0;
    }
    ctx->layers_in_arena = true;
  }
  ctx->ltab = reinterpret_cast<LayerTable*>(c.arena + L.off_ltab[b]);
  ctx->lpool_rgba = reinterpret_cast<float4*>(c.arena + L.off_lpool_rgba[b]);
  ctx->lpool_depth = reinterpret_cast<float*>(c.arena + L.off_lpool_depth[b]);
  ctx->lpool_cap = c.max_partials;
  ctx->layers_pushed = c.layer_push && c.peer_dev && c.size > 1;
  return VR_OK;
}

// where the sampler pushes the entries of the layer frame being traced (see TraceParams::lpush_*)
void comm_layer_push_target(vr_ctx* ctx, TraceParams& p)
{
  Comm& c = ctx->comm;
  p.lpush_peers = nullptr;
  if (!ctx->layers_pushed || !c.on || !c.peer_dev || c.size < 2) return;
  const Layout L = make_layout(c.max_pixels, c.max_partials, c.rank == 0, c.size);
  const int b = (int)((c.lepoch + 1) % kLayerRing);
  p.lpush_peers = c.peer_dev;
  p.lpush_off_rgba = L.off_lrecv_rgba[b];
  p.lpush_off_depth = L.off_lrecv_depth[b];
  p.lpush_stride = c.max_partials;
  p.lpush_rank = c.rank;
  p.lpush_size = c.size;
  p.lpush_tpr = (p.W + 31) / 32;
}

vr_status upload_layer_table_pub(vr_ctx* ctx);

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
  VR_ENTER(ctx);
  if (ctx->comm.on) return cfail(ctx, VR_ERR_STATE, "vr_comm_init: already initialised", cudaSuccess);
  if (rank < 0 || n_ranks < 1 || rank >= n_ranks || n_ranks > kMaxRanks || !handle_out || max_pixels == 0)
    return cfail(ctx, VR_ERR_INVALID, "vr_comm_init: bad rank/size (max 16 ranks) or NULL handle", cudaSuccess);
  cudaSetDevice(ctx->device);
  Comm& c = ctx->comm;
  c.rank = rank;
  c.size = n_ranks;
  c.max_pixels = max_pixels;
  c.max_partials = max_partials;
  {
    // every cross-rank wait inside the kernels is bounded (default 20 s; VR_COMM_TIMEOUT_MS=0: unbounded)
    const char* t = std::getenv("VR_COMM_TIMEOUT_MS");
    const long long ms = t ? std::atoll(t) : 20000;
    c.timeout_ns = ms > 0 ? (unsigned long long)ms * 1000000ull : 0ull;
    if (const char* e = std::getenv("VR_TIMELINE")) c.timeline = std::atoi(e) != 0;
  }
  const Layout L = make_layout(max_pixels, max_partials, rank == 0, n_ranks);
  c.arena_bytes = L.total;
  cudaError_t e = cudaMalloc(&c.arena, c.arena_bytes);
  if (e != cudaSuccess) return cfail(ctx, VR_ERR_NOMEM, "vr_comm_init: arena", e);
  cudaMemset(c.arena, 0, kFlagBytes);
  if (max_partials) cudaMemset(c.arena + L.off_lflags, 0, kFlagBytes);
  {
    const unsigned long long cfg[2] = { max_pixels, max_partials };
    cudaMemcpy(c.arena + offsetof(Flags, cfg_max_pixels), cfg, sizeof(cfg), cudaMemcpyHostToDevice);
  }
  if (cudaMalloc(&c.minmax_dev, 2 * sizeof(int)) != cudaSuccess) return cfail(ctx, VR_ERR_NOMEM, "vr_comm_init", cudaErrorMemoryAllocation);
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
  VR_ENTER(ctx);
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
    unsigned long long cfg[2] = { 0, 0 };
    e = cudaMemcpy(cfg, c.peer[r] + offsetof(Flags, cfg_max_pixels), sizeof(cfg), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return cfail(ctx, VR_ERR_CUDA, "vr_comm_connect: reading a peer arena", e);
    if (cfg[0] != c.max_pixels || cfg[1] != c.max_partials)
    {
      char buf[200];
      snprintf(buf, sizeof(buf), "vr_comm_connect: rank %d was initialised with max_pixels/max_partials = %llu/%llu, "
               "this rank with %zu/%zu; they must be identical on all ranks", r, cfg[0], cfg[1], c.max_pixels, c.max_partials);
      ctx->err = buf;
      return VR_ERR_INVALID;
    }
  }
  cudaError_t e = cudaMalloc(&c.peer_dev, sizeof(unsigned char*) * kMaxRanks);
  if (e != cudaSuccess) return cfail(ctx, VR_ERR_NOMEM, "vr_comm_connect", e);
  unsigned char* table[kMaxRanks] = { nullptr };
  for (int r = 0; r < c.size; ++r) table[r] = c.peer[r];
  e = cudaMemcpy(c.peer_dev, table, sizeof(table), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return cfail(ctx, VR_ERR_CUDA, "vr_comm_connect", e);
  return VR_OK;
}

// Deployment shape (i) of SURVEY 8(b): ONE process drives every GPU of the node, one context per GPU.  No IPC:
// the arenas are this process's own allocations; a context on another device reaches them through plain
// peer access (cudaDeviceEnablePeerAccess over NVLink).  Everything else -- kernels, flags, rings -- is the
// one-process-per-GPU path.  Collective entry points must then be ISSUED on every context before any of
// them is synchronised (the kernels of one context wait for the others' flags).
extern "C" vr_status vr_comm_connect_local(vr_ctx* const* ctxs, int n_ranks)
{
  if (!ctxs || n_ranks < 1 || n_ranks > kMaxRanks) return VR_ERR_INVALID;
  for (int r = 0; r < n_ranks; ++r)
  {
    vr_ctx* ctx = ctxs[r];
    if (!ctx) return VR_ERR_INVALID;
    Comm& c = ctx->comm;
    if (!c.on || c.rank != r || c.size != n_ranks)
      return cfail(ctx, VR_ERR_STATE, "vr_comm_connect_local: context r must have been vr_comm_init'ed as rank r of n_ranks", cudaSuccess);
    if (c.peer_dev) return cfail(ctx, VR_ERR_STATE, "vr_comm_connect_local: already connected", cudaSuccess);
    if (c.max_pixels != ctxs[0]->comm.max_pixels || c.max_partials != ctxs[0]->comm.max_partials)
      return cfail(ctx, VR_ERR_INVALID, "vr_comm_connect_local: max_pixels/max_partials must be identical on all ranks", cudaSuccess);
  }
  for (int r = 0; r < n_ranks; ++r)
  {
    vr_ctx* ctx = ctxs[r];
    Comm& c = ctx->comm;
    cudaSetDevice(ctx->device);
    for (int q = 0; q < n_ranks; ++q)
    {
      if (ctxs[q]->device != ctx->device)
      {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, ctx->device, ctxs[q]->device);
        if (!can) return cfail(ctx, VR_ERR_CUDA, "vr_comm_connect_local: no peer access between two of the devices", cudaSuccess);
        cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) return cfail(ctx, VR_ERR_CUDA, "cudaDeviceEnablePeerAccess", e);
      }
      c.peer[q] = ctxs[q]->comm.arena;
    }
    cudaError_t e = cudaMalloc(&c.peer_dev, sizeof(unsigned char*) * kMaxRanks);
    if (e != cudaSuccess) return cfail(ctx, VR_ERR_NOMEM, "vr_comm_connect_local", e);
    unsigned char* table[kMaxRanks] = { nullptr };
    for (int q = 0; q < n_ranks; ++q) table[q] = c.peer[q];
    e = cudaMemcpy(c.peer_dev, table, sizeof(table), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cfail(ctx, VR_ERR_CUDA, "vr_comm_connect_local", e);
    c.local_peers = true;
  }
  return VR_OK;
}

static vr_status comm_composite_images_impl(vr_ctx* ctx, const int* vis_order, bool to_canvas, bool zbuffer = false)
{
  Comm& c = ctx->comm;
  if (!c.on || !c.peer_dev) return cfail(ctx, VR_ERR_STATE, "vr_comm_composite_images: not connected", cudaSuccess);
  if (!vis_order || ctx->W <= 0) return cfail(ctx, VR_ERR_INVALID, "vr_comm_composite_images: no image / NULL order", cudaSuccess);
  radixk::Schedule zsched;
  if (zbuffer && !radixk::make_schedule(c.size, ctx->W, ctx->H, zsched))
    // (every rank computes the same schedule from the same n, W, H: all fail alike, nobody waits.  The
    // reference throws "Unable to decompose domain into N blocks" from RegularDecomposer::fill_divisions.)
    return cfail(ctx, VR_ERR_INVALID, "vr_comm_composite_zbuffer: unable to decompose the frame into one block per rank", cudaSuccess);
  cudaSetDevice(ctx->device);
  const Layout L = make_layout(c.max_pixels, c.max_partials, c.rank == 0, c.size);
  c.epoch += 1; // the image was quantised into parity (epoch+1)&1 by vr_image_from_canvas
  const int b = c.epoch & 1;
  FoldP2PParams p;
  std::memset(&p, 0, sizeof(p));
  p.peers = c.peer_dev;
  p.rank = c.rank;
  p.size = c.size;
  p.epoch = c.epoch;
  p.n_pixels = (size_t)ctx->W * ctx->H;
  p.W = ctx->W;
  for (int k = 0; k < 4; ++k) p.rect[k] = ctx->img_rect[k];
  if (ctx->W % 4 != 0) { p.rect[0] = p.rect[1] = 0; p.rect[2] = p.rect[3] = 0x7fffffff; }
  p.off_img_rgba = L.off_img_rgba[c.epoch % kImgRing];
  p.off_img_depth = L.off_img_depth[c.epoch % kImgRing];
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
  p.zbuffer = zbuffer ? 1 : 0;
  if (zbuffer)
  {
    for (int k = 0; k < 16; ++k) { p.zs.lo_x[k] = zsched.lo[0][k]; p.zs.lo_y[k] = zsched.lo[1][k]; }
    p.zs.div_x = zsched.divisions[0];
    std::memcpy(p.zs.pos, zsched.pos, sizeof(p.zs.pos));
  }
  p.timeout_ns = c.timeout_ns;
  p.timeline = c.timeline ? 1 : 0;
  p.pushed = ctx->img_pushed ? 1 : 0;
  {
    const size_t n4 = (p.n_pixels + 3) / 4;
    const size_t n_chunks = (n4 + kChunkGroups - 1) / kChunkGroups;
    p.share_groups = ((n_chunks + c.size - 1) / c.size) * kChunkGroups;
    p.off_recv_rgba = L.off_recv_rgba[c.epoch % kImgRing];
    p.off_recv_depth = L.off_recv_depth[c.epoch % kImgRing];
  }
  if (to_canvas && !zbuffer && c.rank == 0 && ctx->W % 4 == 0)
  {
    p.canvas_rgba = ctx->canvas_rgba;
    p.canvas_depth = ctx->canvas_depth;
  }
  if (c.rank == 0)
  {
    // may the kernel trust the "cleared outside" rectangles its predecessor left in the flags?
    auto still = [&](const Comm::Clean& k) {
      return k.valid && k.serial == ctx->api_serial && k.W == ctx->W && k.H == ctx->H;
    };
    p.track_res = still(c.clean_res[b]) ? 1 : 0;
    p.track_canvas = (p.canvas_rgba && !ctx->canvas_exposed && still(c.clean_canvas)) ? 1 : 0;
  }
  // the exchange stream: ordered after the trace that produced this image (everything queued on the
  // context's stream so far), not before anything the caller queues next
  cudaStream_t xs = c.xstream ? c.xstream : ctx->stream;
  const bool fused_tail = c.rank != 0 || p.canvas_rgba != nullptr || (!to_canvas);
  if (!fused_tail) xs = ctx->stream; // (rank 0's unfused ImageToCanvas below runs on the context's stream)
  if (xs != ctx->stream)
  {
    cudaEventRecord(c.ev_trace, ctx->stream);
    cudaStreamWaitEvent(xs, c.ev_trace, 0);
    // this frame's image was traced on side stream (epoch % streams): the exchange is what waits for it, and whoever
    // joins the exchange has thereby joined the trace (a frame traced AHEAD sits on the other stream)
    const int side = c.trace_streams > 0 ? (int)(c.epoch % (unsigned)c.trace_streams) : 0;
    if (c.t_pending[side])
    {
      cudaStreamWaitEvent(xs, c.ev_t[side], 0);
      c.t_pending[side] = false;
    }
  }
  else
    VR_JOIN(ctx);
  p.light = c.fold_light;
  p.grid_per_sm = c.fold_grid;
  p.force_nr8 = c.fold_nr8 ? 1 : 0;
  p.max_ctas = c.exchange_max_ctas;
  cudaError_t e = launch_fold_p2p(p, ctx->sm_count, xs);
  if (e != cudaSuccess) return cfail(ctx, VR_ERR_CUDA, "fold_p2p launch", e);
  ctx->launches++;
  if (c.rank == 0)
  {
    const Flags* f = reinterpret_cast<const Flags*>(c.arena + L.off_flags);
    // result of this epoch (res_* currently points at parity b)
    ctx->res_rgba = reinterpret_cast<uchar4*>(c.arena + L.off_res_rgba[b]);
    ctx->res_depth = reinterpret_cast<float*>(c.arena + L.off_res_depth[b]);
    if (p.canvas_rgba)
    {
      // waits for every rank's "done" itself, then converts the covered groups
      const int cgrid = c.exchange_max_ctas > 0 ? std::min(ctx->sm_count * 2, c.exchange_max_ctas) : ctx->sm_count * 2;
      if (p.light == 1) covered_to_canvas_kernel<128><<<cgrid, 128, 0, xs>>>(p);
      else covered_to_canvas_kernel<256><<<cgrid, 256, 0, xs>>>(p);
      ctx->launches++;
    }
    else
    {
      wait_done_kernel<<<1, 32, 0, xs>>>(const_cast<Flags*>(f), f->done, c.size, c.epoch, kPathImage, c.timeout_ns);
      ctx->launches++;
      if (to_canvas)
      {
        vr_status st = vr_image_result_to_canvas(ctx);
        if (st != VR_OK) return st;
      }
    }
  }
  if (c.rank == 0)
  {
    // (a canvas not written by this call keeps whatever state it had: its serial no longer matches
    // only if some other entry point has been called since)
    c.clean_res[b].valid = true; c.clean_res[b].serial = ctx->api_serial; c.clean_res[b].W = ctx->W; c.clean_res[b].H = ctx->H;
    if (p.canvas_rgba)
    {
      c.clean_canvas.valid = true; c.clean_canvas.serial = ctx->api_serial; c.clean_canvas.W = ctx->W; c.clean_canvas.H = ctx->H;
    }
  }
  c.x_of_img_epoch[c.epoch & 7] = 0;
  if (xs != ctx->stream)
  {
    c.xserial += 1;
    cudaEventRecord(c.ev_x[c.xserial & 7], xs);
    c.x_pending = true;
    c.x_of_img_epoch[c.epoch & 7] = c.xserial;
  }
  // the next frame's image: the following ring slot -- where a frame traced ahead already sits
  const int ns = (int)((c.epoch + 1) % kImgRing);
  ctx->img_rgba = reinterpret_cast<uchar4*>(c.arena + L.off_img_rgba[ns]);
  ctx->img_depth = reinterpret_cast<float*>(c.arena + L.off_img_depth[ns]);
  ctx->img_pushed = false;
  if (ctx->img_ahead)
  {
    for (int k = 0; k < 4; ++k) ctx->img_rect[k] = ctx->img_rect_ahead[k];
    ctx->img_pushed = ctx->img_pushed_ahead;
    ctx->img_ahead = false;
    ctx->img_pushed_ahead = false;
  }
  return VR_OK;
}

extern "C" vr_status vr_image_result_download(vr_ctx* ctx, uint8_t* rgba, float* depth)
{
  VR_ENTER_RO(ctx);
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
  VR_ENTER(ctx);
  if (ctx->W <= 0 || !ctx->res_rgba) return cfail(ctx, VR_ERR_STATE, "vr_image_result_to_canvas: no result", cudaSuccess);
  return vr_image_to_canvas_dev(ctx, reinterpret_cast<const uint8_t*>(ctx->res_rgba), ctx->res_depth);
}

namespace vr
{
vr_status ensure_partial_scratch_pub(vr_ctx* ctx, size_t n_pixels, size_t n_parts);
void fill_to_canvas_params_pub(const vr_camera* cam, int W, int H, ToCanvasParams& tp);
}

static vr_status comm_composite_partials_impl(vr_ctx* ctx, const vr_camera* cam)
{
  Comm& c = ctx->comm;
  if (!c.on || !c.peer_dev) return cfail(ctx, VR_ERR_STATE, "vr_comm_composite_partials: not connected", cudaSuccess);
  if (ctx->pW <= 0) return cfail(ctx, VR_ERR_STATE, "vr_comm_composite_partials: call vr_partials_begin first", cudaSuccess);
  const size_t n_pixels = (size_t)ctx->pW * ctx->pH;
  if (c.max_partials == 0 || n_pixels > c.max_pixels)
    return cfail(ctx, VR_ERR_INVALID, "vr_comm_composite_partials: frame larger than the max_pixels/max_partials given to vr_comm_init", cudaSuccess);
  cudaSetDevice(ctx->device);
  if (ctx->n_partials_host > c.max_partials)
  {
    // rank-local: the peers are (or will be) inside this exchange already -- release them
    const Layout La = make_layout(c.max_pixels, c.max_partials, c.rank == 0, c.size);
    c.pepoch += 1;
    abort_announce_kernel<<<1, 32, 0, ctx->stream>>>(c.peer_dev, c.rank, c.size, La.off_flags, kPathPartials, c.pepoch);
    ctx->launches++;
    char buf[200];
    snprintf(buf, sizeof(buf), "vr_comm_composite_partials: up to %zu partials this frame but max_partials = %zu "
             "(exchange %u aborted on all ranks)", ctx->n_partials_host, c.max_partials, c.pepoch);
    ctx->err = buf;
    ctx->n_partials_host = 0;
    return VR_ERR_NOMEM;
  }
  vr_status st = ensure_partial_scratch_pub(ctx, n_pixels, ctx->partial_cap ? ctx->partial_cap : 1);
  if (st != VR_OK) return st;
  const Layout L = make_layout(c.max_pixels, c.max_partials, c.rank == 0, c.size);
  c.pepoch += 1;
  const int par = c.pepoch & 1;
  PartialScratch sc{ ctx->px_count, ctx->px_end, ctx->sidx, ctx->rec, ctx->scan_blocks };
  cudaError_t e = cudaSuccess;
  // (1) local: order my list by (pixel, depth, list index) into the arena + per-pixel offsets
  vr_partial* sorted = reinterpret_cast<vr_partial*>(c.arena + L.off_psorted[par]);
  int* poff = reinterpret_cast<int*>(c.arena + L.off_poff[par]);
  if (ctx->partial_cap)
    ctx->launches += launch_partials_pixel_sort(ctx->partials, ctx->partial_count, ctx->partial_cap, n_pixels, sc,
                                                sorted, c.max_partials, poff, c.minmax_dev, ctx->stream, &e);
  else
  {
    // a rank without any block this frame: empty list
    cudaMemsetAsync(poff, 0, n_pixels * sizeof(int), ctx->stream);
    const int mm[2] = { 0x7fffffff, -1 };
    e = cudaMemcpyAsync(c.minmax_dev, mm, sizeof(mm), cudaMemcpyHostToDevice, ctx->stream);
  }
  if (e != cudaSuccess) return cfail(ctx, VR_ERR_CUDA, "partial pixel sort", e);
  publish_minmax_kernel<<<1, 32, 0, ctx->stream>>>(c.peer_dev, c.rank, c.size, L.off_flags, par, c.minmax_dev,
                                                   ctx->partial_count, c.max_partials);
  // (2) fused pull + merge + fold + gather
  MergeP2PParams p;
  std::memset(&p, 0, sizeof(p));
  p.peers = c.peer_dev;
  p.rank = c.rank;
  p.size = c.size;
  p.epoch = c.pepoch;
  p.n_pixels = n_pixels;
  p.off_flags = L.off_flags;
  p.off_poff = L.off_poff[par];
  p.off_psorted = L.off_psorted[par];
  p.off_pout = L.off_pout;
  p.sorted_cap = c.max_partials;
  p.out_cap = c.max_pixels;
  p.timeout_ns = c.timeout_ns;
  p.off_canvas_rgba = L.off_canvas_rgba;
  p.off_canvas_depth = L.off_canvas_depth;
  const int grid = ctx->sm_count * 4;
  if (cam)
  {
    fill_to_canvas_params_pub(cam, ctx->pW, ctx->pH, p.tp);
    if (c.rank == 0)
    {
      // Canvas::Clear of the frame, on rank 0's arena canvas, before this rank announces "ready"
      // (peers store their finished pixels only after that)
      vr_status st2 = vr_canvas_clear(ctx, ctx->pW, ctx->pH);
      if (st2 != VR_OK) return st2;
    }
    if (c.size <= 1) merge_fold_p2p_kernel<1, true><<<grid, 256, 0, ctx->stream>>>(p);
    else if (c.size <= 2) merge_fold_p2p_kernel<2, true><<<grid, 256, 0, ctx->stream>>>(p);
    else if (c.size <= 4) merge_fold_p2p_kernel<4, true><<<grid, 256, 0, ctx->stream>>>(p);
    else if (c.size <= 8) merge_fold_p2p_kernel<8, true><<<grid, 256, 0, ctx->stream>>>(p);
    else merge_fold_p2p_kernel<16, true><<<grid, 256, 0, ctx->stream>>>(p);
  }
  else if (c.size <= 1) merge_fold_p2p_kernel<1, false><<<grid, 256, 0, ctx->stream>>>(p);
  else if (c.size <= 2) merge_fold_p2p_kernel<2, false><<<grid, 256, 0, ctx->stream>>>(p);
  else if (c.size <= 4) merge_fold_p2p_kernel<4, false><<<grid, 256, 0, ctx->stream>>>(p);
  else if (c.size <= 8) merge_fold_p2p_kernel<8, false><<<grid, 256, 0, ctx->stream>>>(p);
  else merge_fold_p2p_kernel<16, false><<<grid, 256, 0, ctx->stream>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cfail(ctx, VR_ERR_CUDA, "merge_fold_p2p launch", e);
  ctx->launches += 2;
  Flags* f = reinterpret_cast<Flags*>(c.arena + L.off_flags);
  if (c.rank == 0)
  {
    wait_done_kernel<<<1, 32, 0, ctx->stream>>>(f, f->p_done, c.size, c.pepoch, kPathPartials, c.timeout_ns);
    ctx->launches++;
    // root-only result (PartialCompositor.cpp:580-595): the read side now sees the composited list
    // (fused canvas mode: the pixels went straight to the canvas, the list is empty)
    ctx->plist = reinterpret_cast<const vr_partial*>(c.arena + L.off_pout);
    ctx->plist_count = &f->p_out_count;
    ctx->plist_cap = cam ? 0 : c.max_pixels;
  }
  else
  {
    // non-root ranks end with an empty result, their own list stays intact (re-usable)
    ctx->plist = ctx->partials ? ctx->partials : reinterpret_cast<const vr_partial*>(c.arena);
    ctx->plist_count = ctx->partial_count;
    ctx->plist_cap = 0;
  }
  ctx->n_partials_host = 0;
  return VR_OK;
}

extern "C" vr_status vr_comm_composite_partials(vr_ctx* ctx)
{
  VR_ENTER(ctx);
  return comm_composite_partials_impl(ctx, nullptr);
}

extern "C" vr_status vr_comm_composite_partials_to_canvas(vr_ctx* ctx, const vr_camera* cam)
{
  VR_ENTER(ctx);
  if (!cam) return cfail(ctx, VR_ERR_INVALID, "vr_comm_composite_partials_to_canvas: camera is NULL", cudaSuccess);
  return comm_composite_partials_impl(ctx, cam);
}

extern "C" vr_status vr_comm_composite_images(vr_ctx* ctx, const int* vis_order)
{
  VR_ENTER_NOJOIN(ctx); // (the exchange stream orders itself after the trace and the previous exchange)
  return comm_composite_images_impl(ctx, vis_order, false);
}

extern "C" vr_status vr_comm_composite_images_to_canvas(vr_ctx* ctx, const int* vis_order)
{
  VR_ENTER_NOJOIN(ctx); // (the exchange stream orders itself after the trace and the previous exchange)
  return comm_composite_images_impl(ctx, vis_order, true);
}

// The renders of a batch (Scene::Render renders its m_renders in batches and loops over them, Scene.cpp:133-149;
// a Cinema database is 64+ cameras per cycle) are independent frames of the same blocks: one call issues all
// of them back to back -- trace (pushed into the exchange) + visibility-ordered exchange per frame -- so that
// the host's per-call overhead (one ABI crossing, one parameter set-up per frame instead of two) stays off
// the GPU's critical path and consecutive frames overlap on the side streams.  frames_rgba8_host (rank 0 only,
// may be NULL): frame k's final image, background-blended and quantised like Render::Save's input
// (vr_canvas_download_rgba8), is copied to frames_rgba8_host + k * W * H * 4 as soon as its exchange is done.
extern "C" vr_status vr_comm_render_frames(vr_ctx* ctx, int block_id, const vr_camera* cams, int n_frames, int width,
                                           int height, float sample_dist, float range_min, float range_max,
                                           const int* vis_orders, const float* bg_rgba, uint8_t* frames_rgba8_host)
{
  VR_ENTER_NOJOIN(ctx);
  Comm& c = ctx->comm;
  if (!c.on || !c.peer_dev) return cfail(ctx, VR_ERR_STATE, "vr_comm_render_frames: not connected", cudaSuccess);
  if (!cams || !vis_orders || n_frames < 0) return cfail(ctx, VR_ERR_INVALID, "vr_comm_render_frames: NULL argument", cudaSuccess);
  const int flags = (width % 4 == 0) ? (VR_FRAME_NO_CLEAR | VR_FRAME_PUSH) : 0;
  const size_t n = (size_t)width * height;
  if (c.rank == 0 && frames_rgba8_host && n > ctx->enc_cap)
  {
    // (before anything of this batch is queued: nothing here may block on a peer that has not been issued yet)
    VR_JOIN(ctx);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->enc_rgba);
    ctx->enc_rgba = nullptr;
    ctx->enc_cap = 0;
    if (cudaMalloc(&ctx->enc_rgba, n * sizeof(uchar4)) != cudaSuccess) return cfail(ctx, VR_ERR_NOMEM, "vr_comm_render_frames", cudaErrorMemoryAllocation);
    ctx->enc_cap = n;
  }
  for (int k = 0; k < n_frames; ++k)
  {
    vr_status st = vr_trace_to_image(ctx, block_id, cams + k, width, height, sample_dist, range_min, range_max, flags);
    if (st != VR_OK) return st;
    st = comm_composite_images_impl(ctx, vis_orders + (size_t)k * c.size, true);
    if (st != VR_OK) return st;
    if (c.rank == 0 && frames_rgba8_host)
    {
      // on the stream the exchange ran on, right behind it: encode + copy out while the next frames trace
      cudaStream_t xs = c.x_pending ? c.xstream : ctx->stream;
      cudaError_t e = launch_encode_rgba8(ctx->canvas_rgba, width, height, 1, bg_rgba, ctx->enc_rgba, xs);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(frames_rgba8_host + (size_t)k * n * 4, ctx->enc_rgba, n * sizeof(uchar4), cudaMemcpyDeviceToHost, xs);
      if (e != cudaSuccess) return cfail(ctx, VR_ERR_CUDA, "vr_comm_render_frames: encode/copy", e);
      ctx->launches++;
      if (xs == c.xstream) cudaEventRecord(c.ev_x[c.xserial & 7], xs); // joining the exchange now joins the copy too
    }
  }
  return VR_OK;
}

extern "C" vr_status vr_comm_composite_zbuffer(vr_ctx* ctx)
{
  VR_ENTER_NOJOIN(ctx); // (the exchange stream orders itself after the trace and the previous exchange)
  int order[kMaxRanks];
  for (int i = 0; i < kMaxRanks; ++i) order[i] = i; // rank order
  return comm_composite_images_impl(ctx, order, false, true);
}

extern "C" vr_status vr_comm_sync_depths(vr_ctx* ctx)
{
  VR_ENTER(ctx);
  Comm& c = ctx->comm;
  if (!c.on || !c.peer_dev) return cfail(ctx, VR_ERR_STATE, "vr_comm_sync_depths: not connected", cudaSuccess);
  if (ctx->W <= 0 || !ctx->canvas_depth) return cfail(ctx, VR_ERR_STATE, "vr_comm_sync_depths: no canvas", cudaSuccess);
  const size_t n = (size_t)ctx->W * ctx->H;
  if (n > c.max_pixels) return cfail(ctx, VR_ERR_INVALID, "vr_comm_sync_depths: canvas larger than max_pixels", cudaSuccess);
  cudaSetDevice(ctx->device);
  const Layout L = make_layout(c.max_pixels, c.max_partials, c.rank == 0, c.size);
  c.sepoch += 1;
  Flags* root_flags = reinterpret_cast<Flags*>(c.peer[0] + L.off_flags);
  float* staged = reinterpret_cast<float*>(c.peer[0] + L.off_sync_depth);
  if (c.rank == 0)
  {
    // every rank has finished pulling the previous broadcast before the staging copy is replaced
    if (c.sepoch > 1)
    {
      wait_done_kernel<<<1, 32, 0, ctx->stream>>>(root_flags, root_flags->s_done, c.size, c.sepoch - 1, kPathImage, c.timeout_ns);
      ctx->launches++;
    }
    cudaError_t e = cudaMemcpyAsync(staged, ctx->canvas_depth, n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream);
    if (e != cudaSuccess) return cfail(ctx, VR_ERR_CUDA, "vr_comm_sync_depths: stage", e);
    sync_post_kernel<<<1, 32, 0, ctx->stream>>>(root_flags, c.sepoch);
    ctx->launches++;
  }
  else
  {
    Flags* my_flags = reinterpret_cast<Flags*>(c.arena + L.off_flags);
    sync_pull_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(root_flags, my_flags, staged, ctx->canvas_depth, n,
                                                                 c.rank, c.sepoch, c.timeout_ns);
    ctx->launches++;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cfail(ctx, VR_ERR_CUDA, "vr_comm_sync_depths launch", e);
  return VR_OK;
}

extern "C" vr_status vr_comm_layers_composite_to_canvas(vr_ctx* ctx, const vr_camera* cam)
{
  VR_ENTER(ctx);
  Comm& c = ctx->comm;
  if (!c.on || !c.peer_dev) return cfail(ctx, VR_ERR_STATE, "vr_comm_layers_composite_to_canvas: not connected", cudaSuccess);
  if (!cam) return cfail(ctx, VR_ERR_INVALID, "vr_comm_layers_composite_to_canvas: camera is NULL", cudaSuccess);
  if (ctx->lW <= 0 || !ctx->layers_in_arena)
    return cfail(ctx, VR_ERR_STATE, "vr_comm_layers_composite_to_canvas: call vr_layers_begin (after vr_comm_connect) first", cudaSuccess);
  if ((size_t)ctx->lW * ctx->lH > c.max_pixels)
    return cfail(ctx, VR_ERR_INVALID, "vr_comm_layers_composite_to_canvas: frame larger than max_pixels", cudaSuccess);
  const bool too_many = ctx->ltab_host->n > kMaxSmemLayers / c.size;
  cudaSetDevice(ctx->device);
  const Layout L = make_layout(c.max_pixels, c.max_partials, c.rank == 0, c.size);
  c.lepoch += 1;
  const int par = (int)(c.lepoch % kLayerRing);
  vr_status st = upload_layer_table_pub(ctx); // into the arena table of this parity (bound at vr_layers_begin)
  if (st != VR_OK) return st;
  if (c.rank == 0)
  {
    // rank 0's arena canvas takes the frame; Canvas::Clear happens inside the fold kernel (outside the
    // layers' bounding box by rank 0 itself, inside it by the owner of each tile)
    st = ensure_frame_pub(ctx, ctx->lW, ctx->lH);
    if (st != VR_OK) return st;
  }
  LayerFoldParams p;
  std::memset(&p, 0, sizeof(p));
  p.rank = c.rank;
  p.size = c.size;
  p.epoch = c.lepoch;
  p.W = ctx->lW;
  p.H = ctx->lH;
  p.clear = 1;
  p.smem_layers = kMaxSmemLayers;
  p.light = c.fold_light == 1 ? 1 : 0;
  p.timeout_ns = c.timeout_ns;
  p.max_ctas = c.exchange_max_ctas;
  p.pushed = ctx->layers_pushed ? 1 : 0;
  if (p.pushed) p.light = 0; // (the samplers addressed the owners by 32x8 tiles)
  for (int r = 0; r < c.size; ++r)
  {
    p.table[r] = reinterpret_cast<const LayerTable*>(c.peer[r] + L.off_ltab[par]);
    if (p.pushed)
    {
      // every rank's entries of MY tiles sit in my own arena, pushed there by the samplers
      p.pool_rgba[r] = reinterpret_cast<const float4*>(c.arena + L.off_lrecv_rgba[par]) + (size_t)r * c.max_partials;
      p.pool_depth[r] = reinterpret_cast<const float*>(c.arena + L.off_lrecv_depth[par]) + (size_t)r * c.max_partials;
    }
    else
    {
      p.pool_rgba[r] = reinterpret_cast<const float4*>(c.peer[r] + L.off_lpool_rgba[par]);
      p.pool_depth[r] = reinterpret_cast<const float*>(c.peer[r] + L.off_lpool_depth[par]);
    }
    p.flags[r] = c.peer[r] + L.off_lflags;
  }
  p.canvas_rgba = reinterpret_cast<float4*>(c.peer[0] + L.off_canvas_rgba);
  p.canvas_depth = reinterpret_cast<float*>(c.peer[0] + L.off_canvas_depth);
  fill_to_canvas_params_pub(cam, ctx->lW, ctx->lH, p.tp);
  if (too_many || c.frame_poisoned)
  {
    // rank-local error (too many layers here, or a layer of this frame did not fit the arena pool): the
    // peers are inside this exchange already -- release them instead of leaving them to the time limit
    launch_layers_abort(p, ctx->stream);
    ctx->launches++;
    char buf[200];
    if (too_many)
      snprintf(buf, sizeof(buf), "vr_comm_layers_composite_to_canvas: %d layers on this rank, at most %d with %d ranks "
               "(exchange %u aborted on all ranks)", ctx->ltab_host->n, kMaxSmemLayers / c.size, c.size, c.lepoch);
    else
      snprintf(buf, sizeof(buf), "vr_comm_layers_composite_to_canvas: a layer of this frame did not fit max_partials "
               "(exchange %u aborted on all ranks)", c.lepoch);
    ctx->err = buf;
    c.frame_poisoned = false;
    ctx->lW = ctx->lH = 0;
    return VR_ERR_INVALID;
  }
  // on the exchange stream: after everything queued on the context so far (the traces that filled the layers,
  // the table upload), overlapping whatever the caller traces next
  cudaStream_t xs = c.xstream ? c.xstream : ctx->stream;
  if (xs != ctx->stream)
  {
    cudaEventRecord(c.ev_trace, ctx->stream);
    cudaStreamWaitEvent(xs, c.ev_trace, 0);
  }
  cudaError_t e = launch_layers_fold(p, true, ctx->sm_count, xs);
  if (e != cudaSuccess) return cfail(ctx, VR_ERR_CUDA, "layers_fold launch", e);
  ctx->launches++;
  if (c.rank == 0)
  {
    e = launch_layers_wait_done(c.arena + L.off_lflags, c.size, c.lepoch, c.timeout_ns, xs);
    if (e != cudaSuccess) return cfail(ctx, VR_ERR_CUDA, "layers wait launch", e);
    ctx->launches++;
  }
  if (xs != ctx->stream)
  {
    c.xserial += 1;
    cudaEventRecord(c.ev_x[c.xserial & 7], xs);
    c.x_pending = true;
  }
  ctx->lW = ctx->lH = 0; // the frame is consumed: vr_layers_begin starts the next one
  return VR_OK;
}
