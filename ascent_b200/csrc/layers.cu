// layers.cu -- path B of the volume plot without lists: dense per-block ray layers.
//
// The reference turns every structured block's rays into a compact std::vector<VolumePartial>
// (StructuredWrapper::render, src/libs/vtkh/rendering/VolumeRenderer.cpp:260-283, a serial
// push_back loop), concatenates the vectors, redistributes them by pixel range over MPI
// (vtkh_diy_partial_redistribute.hpp:58-152), std::sort's them by (pixel, depth)
// (PartialCompositor.cpp:343), folds each pixel's run front to back (:56-95, VolumePartial.hpp:86-95),
// gathers the result on rank 0 (vtkh_diy_partial_collect.hpp:57-141) and splats it onto the canvas
// (partials_to_canvas, VolumeRenderer.cpp:287-391).
//
// A structured block contributes AT MOST ONE partial per pixel, and only inside the screen
// rectangle its rays were generated for.  So the B200 formulation keeps each block's rays where
// they fall: the sampler stores {rgba, exit distance} of every ray of the block's rectangle into a
// dense "layer" (coalesced 16 + 4 byte stores, no atomics, no compaction; alpha = 0 marks a ray the
// reference would have dropped with its alpha < 0.001 test).  Compositing is then ONE kernel: a CTA
// owns a 32x8 pixel tile, finds the layers (of every rank) that overlap it, and each thread gathers
// its pixel's <= depth-complexity entries -- straight out of the peers' HBM over NVLink when they
// are remote -- orders them by (exit distance, rank, block) in registers/local memory, folds them
// with VolumePartial::blend and stores the finished canvas pixel (into rank 0's canvas).  No
// histogram, scan, scatter, sort or gather passes; traffic = the layer entries that are read once
// + the canvas pixels that are written once.  Results are bit-identical to the list pipeline
// (same entries, same order, same arithmetic), which is what tests/ check.
//
// Compiled with --fmad=false like the rest of the float fold.
#include <cstdio>
#include <cstring>

#include "vr_internal.h"

namespace vr
{
namespace
{

// A CTA owns a 32 x TH pixel tile.  TH = 8 (256 threads) when the fold has the GPU to itself; TH = 4 (128
// threads, <= 72 registers) is the footprint of ONE sampler CTA: the fold of frame k runs on the exchange
// stream while the sampler's persistent grids are busy with frame k+1, and such a CTA is placed as soon as any
// sampler CTA retires instead of waiting for two slots of one SM to fall free together.
constexpr int kTileW = 32;
constexpr int kMaxTileLayers = 192;              // layers overlapping one tile (smem list)
constexpr int kMaxSeg = 20;                      // entries per pixel ordered in shared memory (deeper: selection)
constexpr int kBatch = 8;                        // layer entries requested together per pixel
// per-thread sort keys live in shared memory, [slot][thread] so that a warp's accesses never conflict:
// exit distance and (tie-break order << 12 | index into the tile's layer list)
constexpr size_t key_bytes(int th) { return (size_t)kMaxSeg * kTileW * th * 8; }

__device__ __forceinline__ void blend(float4& a, const float4 o)
{
  // VolumePartial::blend, VolumePartial.hpp:86-95 (rgb in xyz, alpha in w)
  if (a.w >= 1.f || o.w == 0.f) return;
  const float opacity = (1.f - a.w);
  a.x += opacity * o.x;
  a.y += opacity * o.y;
  a.z += opacity * o.z;
  a.w += opacity * o.w;
  a.w = a.w > 1.f ? 1.f : a.w;
}

struct TileLayer
{
  int x0, y0, x1, y1;       // rectangle, exclusive upper corner
  int w;                    // row pitch of the layer
  int order;                // rank * kMaxLayers + index: the (rank, domain) tie-break
  const float4* rgba;       // entry (0,0) of the layer
  const float* depth;
};

// COMM: flags/peers are live, the frame starts from a cleared canvas: the owner of a tile writes all
// of its pixels into rank 0's canvas, rank 0 clears what lies outside the layers' bounding box;
// !COMM: one rank, every pixel of the frame is written (cleared or blended over).
template <bool COMM, int kTileH>
__global__ void __launch_bounds__(kTileW* kTileH, kTileH == 4 ? 7 : 4) layers_fold_kernel(const __grid_constant__ LayerFoldParams P)
{
  constexpr int kThreadsFold = kTileW * kTileH;
  constexpr size_t kKeyBytes = (size_t)kMaxSeg * kTileW * kTileH * 8;
  extern __shared__ unsigned char smem_raw[];
  float* s_kd = reinterpret_cast<float*>(smem_raw);                               // sort keys: exit distance
  int* s_ko = reinterpret_cast<int*>(smem_raw + kKeyBytes / 2);                   // sort keys: order | list index
  LayerDesc* s_desc = reinterpret_cast<LayerDesc*>(smem_raw + kKeyBytes);         // all ranks' tables
  __shared__ int s_first[kMaxCommRanks + 1];                          // table offsets per rank
  __shared__ TileLayer s_tile[kMaxTileLayers];
  __shared__ int s_ntile;
  __shared__ int s_total;

  bool go = true; // false: a peer aborted this exchange (or never showed up): fold nothing, still report "done"
  if (COMM)
  {
    LayerFlags* my_flags = reinterpret_cast<LayerFlags*>(P.flags[P.rank]);
    if (blockIdx.x == 0 && threadIdx.x < P.size)
    {
      LayerFlags* f = reinterpret_cast<LayerFlags*>(P.flags[threadIdx.x]);
      __threadfence_system();
      st_release_sys(&f->ready[P.rank], P.epoch);
    }
    go = wait_all_ready(reinterpret_cast<ExchangeError*>(my_flags->err), my_flags->ready, my_flags->aborted, 2u, P.size,
                        P.epoch, P.timeout_ns, 64);
  }

  // ---- every rank's layer table into shared memory (once per CTA)
  if (threadIdx.x == 0)
  {
    int total = 0;
    for (int r = 0; r < P.size; ++r)
    {
      s_first[r] = total;
      int n = go ? P.table[r]->n : 0;
      if (n > kMaxLayers) n = kMaxLayers;
      if (total + n > P.smem_layers) n = P.smem_layers - total; // flagged on the host side
      total += n;
    }
    s_first[P.size] = total;
    s_total = total;
  }
  __syncthreads();
  for (int r = 0; r < P.size; ++r)
  {
    const int n = s_first[r + 1] - s_first[r];
    for (int k = threadIdx.x; k < n; k += blockDim.x) s_desc[s_first[r] + k] = P.table[r]->d[k];
  }
  __syncthreads();

  const int lx = threadIdx.x % kTileW, ly = threadIdx.x / kTileW;
  float4* canvas = P.canvas_rgba;
  float* cdepth = P.canvas_depth;

  // ---- the tile-aligned bounding box of every layer of the frame: only tiles inside it can hold
  // a partial; everything outside is Canvas::Clear territory
  __shared__ int s_box[4];
  if (threadIdx.x < 32)
  {
    int x0 = 0x7fffffff, y0 = 0x7fffffff, x1 = 0, y1 = 0;
    for (int k = threadIdx.x; k < s_total; k += 32)
    {
      const LayerDesc d = s_desc[k];
      x0 = min(x0, d.x0); y0 = min(y0, d.y0); x1 = max(x1, d.x0 + d.w); y1 = max(y1, d.y0 + d.h);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o)); y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o));
      x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    }
    if (threadIdx.x == 0)
    {
      if (x1 <= x0 || y1 <= y0) x0 = y0 = x1 = y1 = 0;
      s_box[0] = x0 / kTileW; s_box[1] = y0 / kTileH;
      s_box[2] = (min(x1, P.W) + kTileW - 1) / kTileW; s_box[3] = (min(y1, P.H) + kTileH - 1) / kTileH;
    }
  }
  __syncthreads();
  const int btx0 = s_box[0], bty0 = s_box[1];
  const int btw = s_box[2] - s_box[0], bth = s_box[3] - s_box[1];
  const long long n_tiles = (long long)btw * bth;

  // ---- Canvas::Clear outside the box: streaming stores, no table look-ups.  One rank: only when the
  // frame starts from a cleared canvas; N ranks: rank 0 clears its own canvas (the others' pixels land
  // inside the box only), while the peers' entries are still crossing NVLink.
  if (COMM ? (P.rank == 0) : (P.clear != 0))
  {
    const int px0 = btx0 * kTileW, px1 = (btx0 + btw) * kTileW, py0 = bty0 * kTileH, py1 = (bty0 + bth) * kTileH;
    const size_t n_px = (size_t)P.W * P.H;
    for (size_t px = (size_t)blockIdx.x * blockDim.x + threadIdx.x; px < n_px; px += (size_t)gridDim.x * blockDim.x)
    {
      const unsigned uy = (unsigned)px / (unsigned)P.W; // 32-bit: W * H < 2^31
      const int y = (int)uy, x = (int)((unsigned)px - uy * (unsigned)P.W);
      if (y >= py0 && y < py1 && x >= px0 && x < px1) continue;
      canvas[px] = make_float4(0.f, 0.f, 0.f, 0.f);
      cdepth[px] = 1.001f;
    }
  }

  // tiles of the box are dealt round-robin to the ranks, then to the CTAs of a rank.  Pushed layers: by the
  // tile's ABSOLUTE index in the frame (what the samplers used to address the owners), walking the box's tile
  // rows and skipping the columns outside the box
  const int tpr = (P.W + kTileW - 1) / kTileW;
  const long long t_first = P.pushed ? (long long)bty0 * tpr : 0;
  const long long t_end = P.pushed ? (long long)(bty0 + bth) * tpr : n_tiles;
  long long t_begin = t_first + (long long)P.rank;
  if (P.pushed) t_begin = t_first + (((long long)P.rank - t_first) % P.size + P.size) % P.size;
  for (long long t = t_begin + (long long)blockIdx.x * P.size; t < t_end; t += (long long)gridDim.x * P.size)
  {
    int tile_x, tile_y;
    if (P.pushed)
    {
      tile_x = (int)(t % tpr); tile_y = (int)(t / tpr);
      if (tile_x < btx0 || tile_x >= btx0 + btw) continue; // (uniform for the CTA: no barrier is skipped by a part of it)
    }
    else { tile_x = btx0 + (int)(t % btw); tile_y = bty0 + (int)(t / btw); }
    const int tx0 = tile_x * kTileW, ty0 = tile_y * kTileH;
    if (threadIdx.x == 0) s_ntile = 0;
    __syncthreads();
    // ---- layers overlapping this tile
    for (int r = 0; r < P.size; ++r)
      for (int k = s_first[r] + threadIdx.x; k < s_first[r + 1]; k += blockDim.x)
      {
        const LayerDesc d = s_desc[k];
        if (d.x0 < tx0 + kTileW && d.x0 + d.w > tx0 && d.y0 < ty0 + kTileH && d.y0 + d.h > ty0)
        {
          const int slot = atomicAdd(&s_ntile, 1);
          if (slot < kMaxTileLayers)
          {
            TileLayer L;
            L.x0 = d.x0; L.y0 = d.y0; L.x1 = d.x0 + d.w; L.y1 = d.y0 + d.h; L.w = d.w;
            L.order = r * kMaxLayers + (k - s_first[r]);
            L.rgba = P.pool_rgba[r] + d.base;
            L.depth = P.pool_depth[r] + d.base;
            s_tile[slot] = L;
          }
        }
      }
    __syncthreads();
    // more overlapping layers than the tile list holds: walk the full table instead (slow, exact)
    const bool over = s_ntile > kMaxTileLayers;
    const int nt = over ? s_total : s_ntile;
    auto fetch = [&](int l) -> TileLayer {
      if (!over) return s_tile[l];
      int r = 0;
      while (l >= s_first[r + 1]) ++r;
      const LayerDesc d = s_desc[l];
      TileLayer L;
      L.x0 = d.x0; L.y0 = d.y0; L.x1 = d.x0 + d.w; L.y1 = d.y0 + d.h; L.w = d.w;
      L.order = r * kMaxLayers + (l - s_first[r]);
      L.rgba = P.pool_rgba[r] + d.base;
      L.depth = P.pool_depth[r] + d.base;
      return L;
    };

    const int x = tx0 + lx, y = ty0 + ly;
    if (x < P.W && y < P.H)
    {
      // ---- pass 1: this pixel's entries, insertion-sorted by (exit distance, rank, block).  Only the
      // sort key is kept -- exit distance and the entry's place in the tile's layer list, in SHARED memory
      // ([slot][thread]: conflict-free), not in per-thread local arrays -- so the kernel stays at ~50
      // registers and several CTAs per SM.  The key fields of up to kBatch overlapping layers (exit
      // distance, alpha) are requested together: independent loads, remote ones cross NVLink once per batch.
      int c = 0;
      bool deep = false;
      for (int l0 = 0; l0 < nt && !deep; l0 += kBatch)
      {
        float w[kBatch], d[kBatch];
        int key[kBatch];
        bool in[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u)
        {
          in[u] = false;
          if (l0 + u < nt)
          {
            const TileLayer L = fetch(l0 + u);
            if (!(x < L.x0 || x >= L.x1 || y < L.y0 || y >= L.y1))
            {
              const size_t e = (size_t)(y - L.y0) * L.w + (x - L.x0);
              w[u] = reinterpret_cast<const float*>(L.rgba + e)[3];
              d[u] = L.depth[e];
              key[u] = (L.order << 12) | (l0 + u); // order < 2^14 (16 ranks x 1024 layers), list index < 2^12
              in[u] = true;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u)
        {
          if (!in[u] || w[u] < 0.001f) continue; // the reference's `if(alpha < 0.001f) continue;` (:270-272)
          if (c == kMaxSeg) { deep = true; break; }
          int b = c - 1;
          while (b >= 0 && (s_kd[b * kThreadsFold + threadIdx.x] > d[u] ||
                            (s_kd[b * kThreadsFold + threadIdx.x] == d[u] && s_ko[b * kThreadsFold + threadIdx.x] > key[u])))
          {
            s_kd[(b + 1) * kThreadsFold + threadIdx.x] = s_kd[b * kThreadsFold + threadIdx.x];
            s_ko[(b + 1) * kThreadsFold + threadIdx.x] = s_ko[b * kThreadsFold + threadIdx.x];
            --b;
          }
          s_kd[(b + 1) * kThreadsFold + threadIdx.x] = d[u];
          s_ko[(b + 1) * kThreadsFold + threadIdx.x] = key[u];
          ++c;
        }
      }
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      float first_depth = 0.f;
      bool any = false;
      if (!deep)
      {
        // ---- pass 2: the colours, in order (the same 32-byte sectors pass 1 touched: cache hits)
        for (int a = 0; a < c; ++a)
        {
          const TileLayer L = fetch(s_ko[a * kThreadsFold + threadIdx.x] & 0xfff);
          const float4 q = L.rgba[(size_t)(y - L.y0) * L.w + (x - L.x0)];
          if (a == 0) { acc = q; first_depth = s_kd[threadIdx.x]; any = true; }
          else blend(acc, q);
        }
      }
      else
      {
        // more than kMaxSeg entries on this pixel: repeated selection of the next key, no storage
        float last_d = 0.f;
        int last_o = -1;
        for (;;)
        {
          int best = -1, best_o = 0;
          float best_d = 0.f;
          for (int l = 0; l < nt; ++l)
          {
            const TileLayer L = fetch(l);
            if (x < L.x0 || x >= L.x1 || y < L.y0 || y >= L.y1) continue;
            const size_t e = (size_t)(y - L.y0) * L.w + (x - L.x0);
            if (L.rgba[e].w < 0.001f) continue;
            const float d = L.depth[e];
            const bool after = !any || d > last_d || (d == last_d && L.order > last_o);
            if (!after) continue;
            if (best < 0 || d < best_d || (d == best_d && L.order < best_o)) { best = l; best_d = d; best_o = L.order; }
          }
          if (best < 0) break;
          const TileLayer L = fetch(best);
          const float4 q = L.rgba[(size_t)(y - L.y0) * L.w + (x - L.x0)];
          if (!any) { acc = q; first_depth = best_d; any = true; }
          else blend(acc, q);
          last_d = best_d; last_o = best_o;
        }
      }
      const size_t px = (size_t)y * P.W + x;
      if (any)
      {
        vr_partial part;
        part.pixel_id = (int)px;
        part.depth = first_depth;
        part.rgb[0] = acc.x; part.rgb[1] = acc.y; part.rgb[2] = acc.z;
        part.alpha = acc.w;
        const float4 in = (COMM || P.clear) ? make_float4(0.f, 0.f, 0.f, 0.f) : canvas[px];
        float4 o;
        float dimg;
        partial_to_canvas(part, P.tp, in, o, dimg);
        canvas[px] = o;
        cdepth[px] = dimg;
      }
      else if (COMM || P.clear)
      {
        // inside the box but no partial on this pixel: the cleared canvas value
        canvas[px] = make_float4(0.f, 0.f, 0.f, 0.f);
        cdepth[px] = 1.001f;
      }
    }
    __syncthreads();
  }

  if (COMM)
  {
    // ---- last CTA out tells rank 0 that my tiles have landed
    __syncthreads();
    if (threadIdx.x == 0)
    {
      LayerFlags* my_flags = reinterpret_cast<LayerFlags*>(P.flags[P.rank]);
      __threadfence_system();
      const unsigned prev = atomicAdd(&my_flags->cta_done, 1u);
      if (prev == gridDim.x - 1)
      {
        my_flags->cta_done = 0;
        LayerFlags* root = reinterpret_cast<LayerFlags*>(P.flags[0]);
        __threadfence_system();
        st_release_sys(&root->done[P.rank], P.epoch);
      }
    }
  }
}

__global__ void layers_wait_done_kernel(LayerFlags* f, int size, unsigned int epoch, unsigned long long timeout_ns)
{
  if (threadIdx.x < size && !wait_epoch(f->done + threadIdx.x, epoch, timeout_ns, 64))
    report_error(reinterpret_cast<ExchangeError*>(f->err), epoch, 2u, 1u, threadIdx.x);
}

// layer-path form of comm.cu's abort_announce_kernel
__global__ void layers_abort_kernel(LayerFoldParams P)
{
  const int t = threadIdx.x;
  if (t < P.size)
  {
    LayerFlags* f = reinterpret_cast<LayerFlags*>(P.flags[t]);
    ((volatile unsigned int*)f->aborted)[P.rank] = P.epoch;
    __threadfence_system();
    st_release_sys(&f->ready[P.rank], P.epoch);
  }
  if (t == 0)
  {
    LayerFlags* root = reinterpret_cast<LayerFlags*>(P.flags[0]);
    __threadfence_system();
    st_release_sys(&root->done[P.rank], P.epoch);
    report_error(reinterpret_cast<ExchangeError*>(reinterpret_cast<LayerFlags*>(P.flags[P.rank])->err), P.epoch, 2u, 2u,
                 P.rank);
  }
}

// layers -> the reference's compact list (StructuredWrapper::render :260-283), for callers that
// want vectors and for the parity tests: one warp-aggregated append per 32 entries
__global__ void layers_to_partials_kernel(const LayerTable* __restrict__ table,
                                          const float4* __restrict__ pool_rgba,
                                          const float* __restrict__ pool_depth, int W,
                                          vr_partial* __restrict__ out, unsigned long long* __restrict__ count,
                                          size_t cap)
{
  const int lane = threadIdx.x & 31;
  const int n = table->n;
  for (int l = blockIdx.y; l < n; l += gridDim.y)
  {
    const LayerDesc d = table->d[l];
    const size_t area = (size_t)d.w * d.h;
    for (size_t base = (size_t)blockIdx.x * blockDim.x; base < area; base += (size_t)gridDim.x * blockDim.x)
    {
      const size_t e = base + threadIdx.x;
      bool emit = false;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < area)
      {
        q = pool_rgba[d.base + e];
        emit = !(q.w < 0.001f);
      }
      const unsigned mask = __ballot_sync(0xffffffffu, emit);
      if (!mask) continue;
      unsigned long long b0 = 0;
      if (lane == 0) b0 = atomicAdd(count, (unsigned long long)__popc(mask));
      b0 = __shfl_sync(0xffffffffu, b0, 0);
      const unsigned long long slot = b0 + __popc(mask & ((1u << lane) - 1u));
      if (emit && slot < cap)
      {
        vr_partial p;
        const int y = d.y0 + (int)(e / d.w), x = d.x0 + (int)(e % d.w);
        p.pixel_id = y * W + x;
        p.depth = pool_depth[d.base + e];
        p.rgb[0] = q.x; p.rgb[1] = q.y; p.rgb[2] = q.z;
        p.alpha = q.w;
        out[slot] = p;
      }
    }
  }
}

} // namespace

template <bool COMM, int TH>
static cudaError_t launch_layers_fold_t(const LayerFoldParams& p, int sm_count, cudaStream_t s)
{
  const size_t smem = key_bytes(TH) + (size_t)p.smem_layers * sizeof(LayerDesc);
  // per device, not per process (a process may hold contexts on several GPUs): cheap enough to set per launch
  cudaFuncSetAttribute(layers_fold_kernel<COMM, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
  // (the kernel restricts itself to the layers' bounding box; the full frame bounds the grid)
  const long long tiles = (long long)((p.W + kTileW - 1) / kTileW) * ((p.H + TH - 1) / TH);
  // persistent grid: exactly the CTAs that are resident at once (registers and the table's smem decide)
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, layers_fold_kernel<COMM, TH>, kTileW * TH, smem);
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)sm_count * per_sm;
  const long long mine = (tiles + p.size - 1) / p.size;
  if (grid > mine) grid = mine > 0 ? mine : 1;
  if (p.max_ctas > 0 && grid > p.max_ctas) grid = p.max_ctas;
  layers_fold_kernel<COMM, TH><<<(int)grid, kTileW * TH, smem, s>>>(p);
  return cudaGetLastError();
}

void preload_layers_kernels()
{
  preload_kernel(layers_fold_kernel<true, 4>);
  preload_kernel(layers_fold_kernel<false, 4>);
  preload_kernel(layers_fold_kernel<true, 8>);
  preload_kernel(layers_fold_kernel<false, 8>);
  preload_kernel(layers_wait_done_kernel);
  preload_kernel(layers_abort_kernel);
  preload_kernel(layers_to_partials_kernel);
}

cudaError_t launch_layers_fold(const LayerFoldParams& p, bool comm, int sm_count, cudaStream_t s)
{
  if (p.light) return comm ? launch_layers_fold_t<true, 4>(p, sm_count, s) : launch_layers_fold_t<false, 4>(p, sm_count, s);
  return comm ? launch_layers_fold_t<true, 8>(p, sm_count, s) : launch_layers_fold_t<false, 8>(p, sm_count, s);
}

cudaError_t launch_layers_wait_done(unsigned char* flags, int size, unsigned int epoch, unsigned long long timeout_ns,
                                    cudaStream_t s)
{
  layers_wait_done_kernel<<<1, 32, 0, s>>>(reinterpret_cast<LayerFlags*>(flags), size, epoch, timeout_ns);
  return cudaGetLastError();
}

cudaError_t launch_layers_abort(const LayerFoldParams& p, cudaStream_t s)
{
  layers_abort_kernel<<<1, 32, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_layers_to_partials(const LayerTable* table, int n_layers, const float4* pool_rgba,
                                      const float* pool_depth, int W, vr_partial* out,
                                      unsigned long long* count, size_t cap, cudaStream_t s)
{
  if (n_layers <= 0) return cudaSuccess;
  dim3 grid(64, n_layers < 64 ? n_layers : 64);
  layers_to_partials_kernel<<<grid, 256, 0, s>>>(table, pool_rgba, pool_depth, W, out, count, cap);
  return cudaGetLastError();
}

} // namespace vr
