// composite.cu -- sort-last compositing kernels (single device) for sm_100a.
//
// Path A (uint8 images): Image::Init quantisation, the visibility-ordered truncating-uint8 fold
// of ImageCompositor::Blend/OrderedComposite, ZBufferComposite and ImageToCanvas
// (src/libs/vtkh/compositing/Image.hpp:80-113, ImageCompositor.hpp:15-86,
//  src/libs/vtkh/rendering/Renderer.cpp:265-283).
// Path B (float partials): PartialCompositor::composite_partials
// (src/libs/vtkh/compositing/PartialCompositor.cpp:329-488) re-thought for the GPU: the
// reference's serial std::sort over (pixel, depth) becomes a counting sort by pixel id
// (histogram -> exclusive scan -> scatter) followed by a per-pixel insertion sort on depth and
// the front-to-back VolumePartial::blend fold (VolumePartial.hpp:86-95), one thread per pixel.
//
// All of these are HBM-streaming integer/byte kernels: 16-byte vector loads/stores, grids sized
// in multiples of the SM count, no shared-memory staging needed (no reuse).
//
// Compiled with --fmad=false: the float fold must round like the reference's scalar code.
#include <cstring>

#include "vr_internal.h"

namespace vr
{
namespace
{

constexpr int kT = 256;

inline int grid_for(size_t work_items, int per_block, int max_blocks = 148 * 16)
{
  size_t b = (work_items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > (size_t)max_blocks) b = max_blocks;
  return (int)b;
}

// ---------------------------------------------------------------- canvas clear
__global__ void canvas_clear_kernel(float4* rgba, float* depth, size_t n)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    rgba[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    depth[i] = 1.001f; // VTKM_DEFAULT_CANVAS_DEPTH
  }
}

// ---------------------------------------------------------------- C1 Image::Init
__device__ __forceinline__ unsigned char f2uc(float c)
{
  // static_cast<unsigned char>(c * 255.f) on x86: cvttss2si then low byte
  return (unsigned char)(__float2int_rz(c * 255.f) & 0xff);
}
__global__ void quantize_kernel(const float4* __restrict__ rgba, const float* __restrict__ depth,
                                size_t n, uchar4* __restrict__ out, float* __restrict__ out_depth)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const float4 c = rgba[i];
    out[i] = make_uchar4(f2uc(c.x), f2uc(c.y), f2uc(c.z), f2uc(c.w));
    float d = depth[i];
    d = d < 0 ? fabsf(d) : d; // vtk-h rule, Image.hpp:110
    out_depth[i] = d;
  }
}

// ---------------------------------------------------------------- C2 ordered fold
struct FoldOrder
{
  int layer[64];
};

__device__ __forceinline__ unsigned blend_u8x4(unsigned front, unsigned back)
{
  const unsigned opacity = 255u - (front >> 24);
  unsigned r = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c)
  {
    const unsigned f = (front >> (8 * c)) & 0xffu;
    const unsigned b = (back >> (8 * c)) & 0xffu;
    const unsigned v = (f + ((opacity * b / 255u) & 0xffu)) & 0xffu; // wrapping uint8 +=
    r |= v << (8 * c);
  }
  return r;
}
__device__ __forceinline__ float std_min(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float blend_depth(float f, float b)
{
  const float d1 = std_min(f, 1.001f), d2 = std_min(b, 1.001f);
  return std_min(d1, d2);
}

// 4 pixels per thread: one 16-byte load per layer for colour, one for depth
__global__ void fold_images_kernel(const uint4* __restrict__ rgba, const float4* __restrict__ depth,
                                   size_t layer_stride4, const __grid_constant__ FoldOrder ord,
                                   int n_layers, size_t n4, uint4* __restrict__ out,
                                   float4* __restrict__ out_depth)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
  {
    uint4 f = __ldg(rgba + (size_t)ord.layer[0] * layer_stride4 + i);
    float4 fd = __ldg(depth + (size_t)ord.layer[0] * layer_stride4 + i);
    for (int l = 1; l < n_layers; ++l)
    {
      const uint4 b = __ldg(rgba + (size_t)ord.layer[l] * layer_stride4 + i);
      const float4 bd = __ldg(depth + (size_t)ord.layer[l] * layer_stride4 + i);
      f.x = blend_u8x4(f.x, b.x); f.y = blend_u8x4(f.y, b.y);
      f.z = blend_u8x4(f.z, b.z); f.w = blend_u8x4(f.w, b.w);
      fd.x = blend_depth(fd.x, bd.x); fd.y = blend_depth(fd.y, bd.y);
      fd.z = blend_depth(fd.z, bd.z); fd.w = blend_depth(fd.w, bd.w);
    }
    out[i] = f;
    out_depth[i] = fd;
  }
}
// scalar tail / unaligned fallback
__global__ void fold_images_scalar_kernel(const unsigned* __restrict__ rgba,
                                          const float* __restrict__ depth, size_t layer_stride,
                                          const __grid_constant__ FoldOrder ord, int n_layers,
                                          size_t begin, size_t n, unsigned* __restrict__ out,
                                          float* __restrict__ out_depth)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    unsigned f = rgba[(size_t)ord.layer[0] * layer_stride + i];
    float fd = depth[(size_t)ord.layer[0] * layer_stride + i];
    for (int l = 1; l < n_layers; ++l)
    {
      f = blend_u8x4(f, rgba[(size_t)ord.layer[l] * layer_stride + i]);
      fd = blend_depth(fd, depth[(size_t)ord.layer[l] * layer_stride + i]);
    }
    out[i] = f;
    out_depth[i] = fd;
  }
}

// ---------------------------------------------------------------- C4 z-buffer select
__global__ void zbuffer_kernel(unsigned* __restrict__ front, float* __restrict__ fdepth,
                               const unsigned* __restrict__ img, const float* __restrict__ depth,
                               size_t n)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const float d = depth[i];
    if (d > 1.f || fdepth[i] < d) continue; // ImageCompositor.hpp:63-66 (depth is >= 0 after Image::Init)
    fdepth[i] = d;
    front[i] = img[i];
  }
}

// ---------------------------------------------------------------- frame epilogue
// Render::RenderBackground -> Canvas::BlendBackground (Render.cpp:277-286), in place on the canvas
__global__ void blend_background_kernel(float4* __restrict__ canvas, size_t n, float4 bg)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    float4 c = canvas[i];
    if (c.w >= 1.f) continue;
    const float alpha = bg.w * (1.f - c.w);
    c.x = c.x + bg.x * alpha;
    c.y = c.y + bg.y * alpha;
    c.z = c.z + bg.z * alpha;
    c.w = alpha + c.w;
    canvas[i] = c;
  }
}
// PNGEncoder::Encode's conversion (ascent_png_encoder.cpp:257-281): (unsigned char)(c * 255.f), rows
// flipped; BLEND fuses the background blend in front of it without touching the canvas
template <bool BLEND>
__global__ void encode_rgba8_kernel(const float4* __restrict__ canvas, int W, int H, int flip, float4 bg,
                                    uchar4* __restrict__ out)
{
  const size_t n = (size_t)W * H;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    float4 c = canvas[i];
    if (BLEND && !(c.w >= 1.f))
    {
      const float alpha = bg.w * (1.f - c.w);
      c.x = c.x + bg.x * alpha;
      c.y = c.y + bg.y * alpha;
      c.z = c.z + bg.z * alpha;
      c.w = alpha + c.w;
    }
    const int y = (int)(i / (size_t)W), x = (int)(i % (size_t)W);
    const size_t o = (size_t)(flip ? H - y - 1 : y) * W + x;
    out[o] = make_uchar4((unsigned char)(__float2ll_rz(c.x * 255.f) & 0xff), (unsigned char)(__float2ll_rz(c.y * 255.f) & 0xff),
                         (unsigned char)(__float2ll_rz(c.z * 255.f) & 0xff), (unsigned char)(__float2ll_rz(c.w * 255.f) & 0xff));
  }
}

// ---------------------------------------------------------------- V10 ImageToCanvas
__global__ void image_to_canvas_kernel(const uchar4* __restrict__ rgba,
                                       const float* __restrict__ depth, size_t n,
                                       float4* __restrict__ canvas, float* __restrict__ cdepth)
{
  const float one_over_255 = 1.f / 255.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const uchar4 c = rgba[i];
    canvas[i] = make_float4((float)c.x * one_over_255, (float)c.y * one_over_255,
                            (float)c.z * one_over_255, (float)c.w * one_over_255);
    cdepth[i] = depth[i];
  }
}

// ---------------------------------------------------------------- synthetic braid (bench input)
template <typename T>
__global__ void braid_kernel(T* out, int nx, int ny, int nz, int i0, int j0, int k0, double dx,
                             double dy, double dz)
{
  const double PI = 3.14159265359;
  const size_t n = (size_t)nx * ny * nz;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride)
  {
    const int i = (int)(idx % nx), j = (int)((idx / nx) % ny), k = (int)(idx / ((size_t)nx * ny));
    const double cx = ((i0 + i) * dx) + (2.0 * PI);
    const double cy = ((j0 + j) * dy) - PI;
    const double cz = ((k0 + k) * dz) - (1.5 * PI);
    double cv = sin(cx) + sin(cy) + 2 * cos(sqrt((cx * cx) / 2.0 + cy * cy) / .75) +
      4 * cos(cx * cy / 4.0);
    cv += sin(cz) + 1.5 * cos(sqrt(cx * cx + cy * cy + cz * cz) / .75);
    out[idx] = (T)cv;
  }
}

// ---------------------------------------------------------------- P4 partial composite
// Every kernel of this pipeline reads the list length from device memory (*count_dev, clamped
// to cap): no host round trip between the tracer that appends partials and the compositor.
__device__ __forceinline__ size_t list_len(const unsigned long long* count_dev, size_t cap)
{
  const unsigned long long c = *count_dev;
  return c < cap ? (size_t)c : cap;
}

__global__ void partial_init_kernel(int* minmax, unsigned long long* out_count)
{
  if (minmax) { minmax[0] = 0x7fffffff; minmax[1] = -1; }
  if (out_count) *out_count = 0ull;
}

// histogram of partials per pixel + the list's min/max pixel id (PartialCompositor::merge,
// PartialCompositor.cpp:242-326 computes the same bounds for the redistribute step)
__global__ void px_count_kernel(const vr_partial* __restrict__ p,
                                const unsigned long long* __restrict__ count_dev, size_t cap,
                                int* __restrict__ cnt, int* __restrict__ minmax)
{
  const size_t n = list_len(count_dev, cap);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  int lo = 0x7fffffff, hi = -1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const int px = p[i].pixel_id;
    atomicAdd(cnt + px, 1);
    lo = min(lo, px);
    hi = max(hi, px);
  }
  if (minmax)
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && hi >= 0)
    {
      atomicMin(minmax, lo);
      atomicMax(minmax + 1, hi);
    }
  }
}

// exclusive scan in three launches, 4096 elements per block, 16-byte loads/stores
constexpr int kScanT = 256, kScanPer = 16, kScanTile = kScanT * kScanPer;
__device__ __forceinline__ int block_exclusive_scan(int v, int* total)
{
  __shared__ int warp_sums[kScanT / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0)
  {
    int s = lane < kScanT / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < kScanT / 32; o <<= 1)
    {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane < kScanT / 32) warp_sums[lane] = s;
  }
  __syncthreads();
  const int base = w ? warp_sums[w - 1] : 0;
  *total = warp_sums[kScanT / 32 - 1];
  __syncthreads();
  return base + inc - v;
}
// a thread's 16 consecutive elements as four int4 (n is padded: arrays are allocated to a multiple
// of kScanTile, the tail past n reads as whatever the memset left there -- zero)
__device__ __forceinline__ void load16(const int* __restrict__ in, size_t base, int v[kScanPer])
{
#pragma unroll
  for (int q = 0; q < kScanPer / 4; ++q)
  {
    const int4 t = *reinterpret_cast<const int4*>(in + base + 4 * q);
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}
__global__ void __launch_bounds__(kScanT) scan_reduce_kernel(const int* __restrict__ in,
                                                            int* __restrict__ block_sums)
{
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanPer;
  int v[kScanPer];
  load16(in, base, v);
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanPer; ++k) s += v[k];
  int total;
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
// single block, serial over chunks of kScanT
__global__ void __launch_bounds__(kScanT) scan_blocks_kernel(int* block_sums, int nb)
{
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b = 0; b < nb; b += kScanT)
  {
    const int i = b + threadIdx.x;
    const int v = i < nb ? block_sums[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, &total);
    if (i < nb) block_sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(kScanT) scan_apply_kernel(const int* __restrict__ in,
                                                           const int* __restrict__ block_sums,
                                                           int* __restrict__ out)
{
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanPer;
  int v[kScanPer];
  load16(in, base, v);
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanPer; ++k) s += v[k];
  int total;
  int ex = block_exclusive_scan(s, &total) + block_sums[blockIdx.x];
#pragma unroll
  for (int q = 0; q < kScanPer / 4; ++q)
  {
    int4 o;
    o.x = ex; ex += v[4 * q];
    o.y = ex; ex += v[4 * q + 1];
    o.z = ex; ex += v[4 * q + 2];
    o.w = ex; ex += v[4 * q + 3];
    *reinterpret_cast<int4*>(out + base + 4 * q) = o;
  }
}

// counting-sort scatter: the record itself moves to its pixel's segment (one coalesced read, one
// 24-byte write), so that everything downstream reads a pixel's partials contiguously.  `end`
// holds each pixel's start offset on entry and its END offset on exit (start = end - count).
__global__ void px_scatter_kernel(const vr_partial* __restrict__ p,
                                  const unsigned long long* __restrict__ count_dev, size_t cap,
                                  int* __restrict__ end, vr_partial* __restrict__ rec,
                                  int* __restrict__ sidx)
{
  const size_t n = list_len(count_dev, cap);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const vr_partial q = p[i];
    const int slot = atomicAdd(end + q.pixel_id, 1);
    rec[slot] = q;
    sidx[slot] = (int)i;
  }
}

__device__ __forceinline__ void partial_blend(vr_partial& a, const vr_partial& o)
{
  if (a.alpha >= 1.f || o.alpha == 0.f) return;
  const float opacity = (1.f - a.alpha);
  a.rgb[0] += opacity * o.rgb[0];
  a.rgb[1] += opacity * o.rgb[1];
  a.rgb[2] += opacity * o.rgb[2];
  a.alpha += opacity * o.alpha;
  a.alpha = a.alpha > 1.f ? 1.f : a.alpha;
}

// Order of one pixel's segment by (depth, list index): the (pixel, depth) key of
// VolumePartial::operator< with list order as the documented tie-break.  Segments are as long as
// the depth complexity of the scene, so the keys are insertion-sorted in a small local array;
// perm[a] = position in the segment of the a-th partial front to back.
constexpr int kMaxLocalSeg = 32;
__device__ __forceinline__ void order_segment(const vr_partial* __restrict__ rec,
                                              const int* __restrict__ sidx, int start, int c,
                                              unsigned char perm[kMaxLocalSeg])
{
  float kd[kMaxLocalSeg];
  int ks[kMaxLocalSeg];
  for (int a = 0; a < c; ++a)
  {
    const float da = rec[start + a].depth;
    const int ia = sidx[start + a];
    int b = a - 1;
    while (b >= 0 && (kd[b] > da || (kd[b] == da && ks[b] > ia)))
    {
      kd[b + 1] = kd[b]; ks[b + 1] = ks[b]; perm[b + 1] = perm[b];
      --b;
    }
    kd[b + 1] = da; ks[b + 1] = ia; perm[b + 1] = (unsigned char)a;
  }
}
// fallback for very deep pixels: sort the segment in place in global memory
__device__ __noinline__ void sort_segment_global(vr_partial* rec, int* sidx, int start, int c)
{
  for (int a = 1; a < c; ++a)
  {
    const vr_partial qa = rec[start + a];
    const int ia = sidx[start + a];
    int b = a - 1;
    while (b >= 0)
    {
      const float db = rec[start + b].depth;
      const int ib = sidx[start + b];
      if (db > qa.depth || (db == qa.depth && ib > ia))
      {
        rec[start + b + 1] = rec[start + b];
        sidx[start + b + 1] = ib;
        --b;
      }
      else break;
    }
    rec[start + b + 1] = qa;
    sidx[start + b + 1] = ia;
  }
}

struct FoldCanvas
{
  // partials_to_canvas (VolumeRenderer.cpp:287-391) fused into the fold; canvas == nullptr: off
  float4* canvas;
  float* cdepth;
  int clear; // 1: the canvas is to be treated as cleared (every pixel is written)
  ToCanvasParams tp;
};

// one thread per pixel: order the pixel's segment, fold it front to back, append the result to
// the composited list and (optionally) write the pixel of the final canvas
__global__ void px_fold_kernel(vr_partial* __restrict__ rec, int* __restrict__ sidx, size_t n_pixels,
                               const int* __restrict__ cnt, const int* __restrict__ end,
                               vr_partial* __restrict__ out, unsigned long long* __restrict__ out_count,
                               const __grid_constant__ FoldCanvas C)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  for (size_t base = (size_t)blockIdx.x * blockDim.x; base < n_pixels; base += stride)
  {
    const size_t px = base + threadIdx.x;
    const int c = px < n_pixels ? cnt[px] : 0;
    vr_partial result;
    if (c > 0)
    {
      const int start = end[px] - c;
      if (c == 1)
        result = rec[start];
      else if (c <= kMaxLocalSeg)
      {
        unsigned char perm[kMaxLocalSeg];
        order_segment(rec, sidx, start, c, perm);
        result = rec[start + perm[0]];
        for (int a = 1; a < c; ++a) partial_blend(result, rec[start + perm[a]]);
      }
      else
      {
        sort_segment_global(rec, sidx, start, c);
        result = rec[start];
        for (int a = 1; a < c; ++a) partial_blend(result, rec[start + a]);
      }
    }
    if (C.canvas && px < n_pixels)
    {
      if (c > 0)
      {
        const float4 in = C.clear ? make_float4(0.f, 0.f, 0.f, 0.f) : C.canvas[px];
        float4 o;
        float d;
        partial_to_canvas(result, C.tp, in, o, d);
        C.canvas[px] = o;
        C.cdepth[px] = d;
      }
      else if (C.clear)
      {
        C.canvas[px] = make_float4(0.f, 0.f, 0.f, 0.f);
        C.cdepth[px] = 1.001f;
      }
    }
    // warp-aggregated append
    const unsigned mask = __ballot_sync(0xffffffffu, c > 0);
    if (mask)
    {
      unsigned long long b0 = 0;
      if (lane == 0) b0 = atomicAdd(out_count, (unsigned long long)__popc(mask));
      b0 = __shfl_sync(0xffffffffu, b0, 0);
      if (c > 0) out[b0 + __popc(mask & ((1u << lane) - 1u))] = result;
    }
  }
}

// one thread per pixel: materialise the pixel's segment in front-to-back order, so that the whole
// list becomes ordered by (pixel, depth, list index) -- the form the multi-GPU merge pulls from
__global__ void px_sort_emit_kernel(vr_partial* __restrict__ rec, int* __restrict__ sidx,
                                    size_t n_pixels, const int* __restrict__ cnt,
                                    const int* __restrict__ end, vr_partial* __restrict__ out,
                                    size_t out_cap)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t px = (size_t)blockIdx.x * blockDim.x + threadIdx.x; px < n_pixels; px += stride)
  {
    const int c = cnt[px];
    if (c == 0) continue;
    const int start = end[px] - c;
    if ((size_t)(start + c) > out_cap) continue; // overflow is flagged by publish_minmax_kernel
    if (c == 1)
      out[start] = rec[start];
    else if (c <= kMaxLocalSeg)
    {
      unsigned char perm[kMaxLocalSeg];
      order_segment(rec, sidx, start, c, perm);
      for (int a = 0; a < c; ++a) out[start + a] = rec[start + perm[a]];
    }
    else
    {
      sort_segment_global(rec, sidx, start, c);
      for (int a = 0; a < c; ++a) out[start + a] = rec[start + a];
    }
  }
}

// ---------------------------------------------------------------- V9 partials_to_canvas
__global__ void partials_to_canvas_kernel(const vr_partial* __restrict__ p,
                                          const unsigned long long* __restrict__ count_dev,
                                          size_t max_n, const __grid_constant__ ToCanvasParams T,
                                          float4* __restrict__ canvas, float* __restrict__ cdepth)
{
  size_t n = (size_t)*count_dev;
  if (n > max_n) n = max_n;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
  {
    const vr_partial part = p[q];
    float4 o;
    float d;
    partial_to_canvas(part, T, canvas[part.pixel_id], o, d);
    canvas[part.pixel_id] = o;
    cdepth[part.pixel_id] = d;
  }
}

} // namespace

// ================================================================= launchers
// vr_partials_append: `n` partials of another producer go behind whatever the frame's list holds
__global__ void partial_append_kernel(const vr_partial* __restrict__ src, unsigned long long n, vr_partial* __restrict__ list,
                                      unsigned long long* __restrict__ count, unsigned long long capacity)
{
  __shared__ unsigned long long s_base;
  // one CTA (the lists of an external producer are host-sized): reserve, then copy
  if (threadIdx.x == 0) s_base = atomicAdd(count, n);
  __syncthreads();
  const unsigned long long base = s_base;
  for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x)
    if (base + i < capacity) list[base + i] = src[i];
}

cudaError_t launch_partial_append(const vr_partial* src_dev, size_t n, vr_partial* list, unsigned long long* count,
                                  size_t capacity, cudaStream_t s)
{
  if (n == 0) return cudaSuccess;
  partial_append_kernel<<<1, 1024, 0, s>>>(src_dev, (unsigned long long)n, list, count, (unsigned long long)capacity);
  return cudaGetLastError();
}

void preload_composite_kernels()
{
  preload_kernel(partial_append_kernel);
  preload_kernel(canvas_clear_kernel);
  preload_kernel(quantize_kernel);
  preload_kernel(fold_images_kernel);
  preload_kernel(fold_images_scalar_kernel);
  preload_kernel(zbuffer_kernel);
  preload_kernel(blend_background_kernel);
  preload_kernel(encode_rgba8_kernel<true>);
  preload_kernel(encode_rgba8_kernel<false>);
  preload_kernel(image_to_canvas_kernel);
  preload_kernel(partial_init_kernel);
  preload_kernel(px_count_kernel);
  preload_kernel(scan_reduce_kernel);
  preload_kernel(scan_blocks_kernel);
  preload_kernel(scan_apply_kernel);
  preload_kernel(px_scatter_kernel);
  preload_kernel(px_fold_kernel);
  preload_kernel(px_sort_emit_kernel);
  preload_kernel(partials_to_canvas_kernel);
}

cudaError_t launch_canvas_clear(float4* rgba, float* depth, size_t n, cudaStream_t s)
{
  canvas_clear_kernel<<<grid_for(n, kT), kT, 0, s>>>(rgba, depth, n);
  return cudaGetLastError();
}

cudaError_t launch_quantize(const float4* rgba, const float* depth, size_t n, uchar4* out,
                            float* out_depth, cudaStream_t s)
{
  quantize_kernel<<<grid_for(n, kT), kT, 0, s>>>(rgba, depth, n, out, out_depth);
  return cudaGetLastError();
}

cudaError_t launch_fold_images(const uchar4* rgba, const float* depth, size_t layer_stride,
                               const int* order, int n_layers, size_t n, uchar4* out,
                               float* out_depth, cudaStream_t s)
{
  if (n_layers < 1 || n_layers > 64) return cudaErrorInvalidValue;
  FoldOrder ord;
  for (int i = 0; i < n_layers; ++i) ord.layer[i] = order[i];
  const bool aligned = (layer_stride % 4 == 0) && (((uintptr_t)rgba | (uintptr_t)depth |
                                                    (uintptr_t)out | (uintptr_t)out_depth) % 16 == 0);
  size_t n4 = aligned ? n / 4 : 0;
  if (n4)
    fold_images_kernel<<<grid_for(n4, kT), kT, 0, s>>>(
      reinterpret_cast<const uint4*>(rgba), reinterpret_cast<const float4*>(depth), layer_stride / 4,
      ord, n_layers, n4, reinterpret_cast<uint4*>(out), reinterpret_cast<float4*>(out_depth));
  if (n4 * 4 < n)
    fold_images_scalar_kernel<<<grid_for(n - n4 * 4, kT), kT, 0, s>>>(
      reinterpret_cast<const unsigned*>(rgba), depth, layer_stride, ord, n_layers, n4 * 4, n,
      reinterpret_cast<unsigned*>(out), out_depth);
  return cudaGetLastError();
}

cudaError_t launch_zbuffer(uchar4* front, float* fdepth, const uchar4* img, const float* depth,
                           size_t n, cudaStream_t s)
{
  zbuffer_kernel<<<grid_for(n, kT), kT, 0, s>>>(reinterpret_cast<unsigned*>(front), fdepth,
                                                 reinterpret_cast<const unsigned*>(img), depth, n);
  return cudaGetLastError();
}

cudaError_t launch_blend_background(float4* canvas, size_t n, const float bg[4], cudaStream_t s)
{
  blend_background_kernel<<<grid_for(n, kT), kT, 0, s>>>(canvas, n, make_float4(bg[0], bg[1], bg[2], bg[3]));
  return cudaGetLastError();
}

cudaError_t launch_encode_rgba8(const float4* canvas, int W, int H, int flip, const float* bg, uchar4* out,
                                cudaStream_t s)
{
  const size_t n = (size_t)W * H;
  if (bg)
    encode_rgba8_kernel<true><<<grid_for(n, kT), kT, 0, s>>>(canvas, W, H, flip, make_float4(bg[0], bg[1], bg[2], bg[3]), out);
  else
    encode_rgba8_kernel<false><<<grid_for(n, kT), kT, 0, s>>>(canvas, W, H, flip, make_float4(0.f, 0.f, 0.f, 0.f), out);
  return cudaGetLastError();
}

cudaError_t launch_image_to_canvas(const uchar4* rgba, const float* depth, size_t n, float4* canvas,
                                   float* cdepth, cudaStream_t s)
{
  image_to_canvas_kernel<<<grid_for(n, kT), kT, 0, s>>>(rgba, depth, n, canvas, cdepth);
  return cudaGetLastError();
}

cudaError_t launch_synth_braid(void* field, int dtype, const int n[3], const int start[3],
                               const int global[3], cudaStream_t s)
{
  const double PI = 3.14159265359;
  const double dx = (double)(float)(4.0 * PI) / (double)(global[0] - 1);
  const double dy = (double)(float)(2.0 * PI) / (double)(global[1] - 1);
  const double dz = (double)(float)(3.0 * PI) / (double)(global[2] - 1);
  const size_t total = (size_t)n[0] * n[1] * n[2];
  if (dtype == VR_F32)
    braid_kernel<float><<<grid_for(total, kT), kT, 0, s>>>((float*)field, n[0], n[1], n[2], start[0],
                                                           start[1], start[2], dx, dy, dz);
  else
    braid_kernel<double><<<grid_for(total, kT), kT, 0, s>>>((double*)field, n[0], n[1], n[2],
                                                            start[0], start[1], start[2], dx, dy, dz);
  return cudaGetLastError();
}

// shared front half: histogram -> exclusive scan -> scatter of the records by pixel.
// On exit sc.px_end[px] is the END offset of pixel px's segment in sc.rec (start = end - count).
static int partials_bin_by_pixel(const vr_partial* in, const unsigned long long* count_dev, size_t cap,
                                 size_t n_pixels, const PartialScratch& sc, int* px_end, int* minmax,
                                 unsigned long long* out_count, cudaStream_t s)
{
  const int nb = (int)((n_pixels + kScanTile - 1) / kScanTile);
  partial_init_kernel<<<1, 1, 0, s>>>(minmax, out_count);
  cudaMemsetAsync(sc.px_count, 0, (size_t)nb * kScanTile * sizeof(int), s); // incl. the scan padding
  px_count_kernel<<<grid_for(cap, kT), kT, 0, s>>>(in, count_dev, cap, sc.px_count, minmax);
  scan_reduce_kernel<<<nb, kScanT, 0, s>>>(sc.px_count, sc.scan_blocks);
  scan_blocks_kernel<<<1, kScanT, 0, s>>>(sc.scan_blocks, nb);
  scan_apply_kernel<<<nb, kScanT, 0, s>>>(sc.px_count, sc.scan_blocks, px_end);
  px_scatter_kernel<<<grid_for(cap, kT), kT, 0, s>>>(in, count_dev, cap, px_end, sc.rec, sc.sidx);
  return 6;
}

size_t partial_scan_padded(size_t n_pixels)
{
  return (n_pixels + kScanTile - 1) / kScanTile * kScanTile;
}

int launch_partials_composite(const vr_partial* in, const unsigned long long* count_dev, size_t cap,
                              size_t n_pixels, const PartialScratch& sc, vr_partial* out,
                              unsigned long long* out_count, const ToCanvasParams* to_canvas,
                              float4* canvas, float* cdepth, int canvas_clear, cudaStream_t s,
                              cudaError_t* err)
{
  int launches = 0;
  FoldCanvas fc;
  memset(&fc, 0, sizeof(fc));
  if (to_canvas)
  {
    fc.canvas = canvas;
    fc.cdepth = cdepth;
    fc.clear = canvas_clear;
    fc.tp = *to_canvas;
  }
  if (cap == 0)
  {
    partial_init_kernel<<<1, 1, 0, s>>>(nullptr, out_count);
    launches = 1;
    if (to_canvas && canvas_clear)
    {
      canvas_clear_kernel<<<grid_for(n_pixels, kT), kT, 0, s>>>(canvas, cdepth, n_pixels);
      launches++;
    }
    *err = cudaGetLastError();
    return launches;
  }
  launches += partials_bin_by_pixel(in, count_dev, cap, n_pixels, sc, sc.px_end, nullptr, out_count, s);
  px_fold_kernel<<<grid_for(n_pixels, kT), kT, 0, s>>>(sc.rec, sc.sidx, n_pixels, sc.px_count, sc.px_end,
                                                       out, out_count, fc);
  launches += 1;
  *err = cudaGetLastError();
  return launches;
}

int launch_partials_pixel_sort(const vr_partial* in, const unsigned long long* count_dev, size_t cap,
                               size_t n_pixels, const PartialScratch& sc, vr_partial* sorted_out,
                               size_t sorted_cap, int* end_out, int* minmax, cudaStream_t s,
                               cudaError_t* err)
{
  int launches = partials_bin_by_pixel(in, count_dev, cap, n_pixels, sc, end_out, minmax, nullptr, s);
  px_sort_emit_kernel<<<grid_for(n_pixels, kT), kT, 0, s>>>(sc.rec, sc.sidx, n_pixels, sc.px_count, end_out,
                                                            sorted_out, sorted_cap);
  launches += 1;
  *err = cudaGetLastError();
  return launches;
}

cudaError_t launch_partials_to_canvas(const vr_partial* p, const unsigned long long* count_dev,
                                      size_t max_n, const ToCanvasParams& tp, float4* canvas,
                                      float* cdepth, cudaStream_t s)
{
  partials_to_canvas_kernel<<<grid_for(max_n ? max_n : 1, kT), kT, 0, s>>>(p, count_dev, max_n, tp,
                                                                            canvas, cdepth);
  return cudaGetLastError();
}

} // namespace vr
