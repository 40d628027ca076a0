// sampler.cu -- the ray-cast sampler of Ascent's `volume` plot, hand-written for sm_100a.
//
// One fused kernel per (block, camera) replaces the seven VTK-m worklets that
// vtkh::VolumeRenderer drives through MapperVolume::RenderCells
// (src/libs/vtkh/rendering/VolumeRenderer.cpp:516-528) and StructuredWrapper::render
// (:239-258): K1 ray generation over the block's screen subset, K2 canvas-depth clamp, K3
// bounds intersection, K4/K5/K6 march + trilinear/nearest sample + transfer function +
// front-to-back accumulation with early termination, then either K7 (blend over the canvas,
// projected entry depth), the alpha >= 0.001 partial emission of :260-283, or -- for a frame that
// starts from a cleared canvas, the normal case of RenderOneDomainPerRank (:482-536) -- K7 fused
// with Canvas::Clear, Image::Init (Image.hpp:80-113) and ImageToCanvas (Renderer.cpp:265-283),
// so that one launch produces the rank's final uint8 image (and canvas).
//
// Numerics: this translation unit is compiled with --fmad=false so every +,* below is a
// separately rounded IEEE f32 operation, exactly like the x86 build of VTK-m the reference's
// OpenMP path is (SURVEY section 7 "Float reproducibility").  That makes every inside/outside
// decision, cell index, colour index and termination test bit-identical to the oracle; where a
// fused multiply-add is wanted it is written explicitly (__fmaf_rn).
//
// Mapping to the machine (evidence: profiles/r1_*):
//  * a warp owns an 8x4 pixel tile: rays of a tile walk through neighbouring cells, so the eight
//    corner gathers of a warp fall into a handful of 32-byte sectors (L1 hit rate ~2/3);
//  * warps pull tiles from a global counter (persistent CTAs, grid = SMs x resident CTAs) so long
//    and short rays balance;
//  * the march is software-pipelined: the cell lookup and the eight gathers of sample k+1 are
//    issued before sample k is interpolated and blended.  The kernel is latency-bound (the sparse
//    default sampling touches ~1/4 of the block: DRAM runs at <10 % of peak), so doubling the loads
//    in flight per warp is what shortens a ray;
//  * the 16 KiB transfer function sits in shared memory (one LDS.128 per sample).
#include <cstring>

#include "vr_internal.h"

namespace vr
{

namespace
{

#ifndef VR_MIN_BLOCKS
#define VR_MIN_BLOCKS 9
#endif
constexpr int kThreads = 128;
constexpr int kTileW = 8, kTileH = 4;

template <typename FT>
__device__ __forceinline__ float load_scalar(const void* field, int i)
{
  return (float)__ldg(reinterpret_cast<const FT*>(field) + i);
}
template <typename FT>
__device__ __forceinline__ float load_scalar(const void* field, long long i)
{
  return (float)__ldg(reinterpret_cast<const FT*>(field) + i);
}

__device__ __forceinline__ float rcp_safe(float f)
{
  return 1.0f / ((fabsf(f) < 1e-8f) ? 1e-8f : f);
}

// RectilinearLocator::LocateCell for one axis: the unique cell with x[c] <= p < x[c+1], found by
// walking from the previous cell like VTK-m does (same result, same memory footprint).
__device__ __forceinline__ void locate_axis_rect(const float* __restrict__ ax, int dim, float maxp,
                                                 float p, int& cell, float& inv_sp)
{
  if (p == maxp)
  {
    cell = dim - 2;
    return; // inv_sp keeps its previous value, as in VTK-m
  }
  float minVal = __ldg(ax + cell);
  const int dir = (p - minVal >= 0.f) ? 1 : -1;
  float maxVal = __ldg(ax + cell + 1);
  while (!(p >= minVal && p < maxVal))
  {
    cell += dir;
    const int next_id = dir == 1 ? cell + 1 : cell;
    const float next = __ldg(ax + next_id);
    if (dir == 1) { minVal = maxVal; maxVal = next; }
    else          { maxVal = minVal; minVal = next; }
  }
  inv_sp = 1.f / (maxVal - minVal);
}

// k-th element of 0..n-1 visited from the middle outwards: m, m+1, m-1, m+2, ... with m = (n-1)/2
__device__ __forceinline__ int centre_out(int k, int n)
{
  const int m = (n - 1) >> 1;
  return (k & 1) ? m + ((k + 1) >> 1) : m - (k >> 1);
}

__device__ __forceinline__ unsigned char quant_u8(float c)
{
  // static_cast<unsigned char>(c * 255.f) on x86: cvttss2si then low byte
  return (unsigned char)(__float2int_rz(c * 255.f) & 0xff);
}

// 512 pixels per chunk: 4 x (32 lanes x 4 consecutive pixels).  A group of 4 pixels is either
// entirely inside the traced rectangle (tx0/tx1 are multiples of 4 when vec_ok) or outside it.
__device__ __forceinline__ void clear_chunk(const TraceParams& P, unsigned chunk, int lane)
{
  const long long n = (long long)P.W * P.H;
#pragma unroll
  for (int it = 0; it < 4; ++it)
  {
    const long long base = (long long)chunk * 512 + it * 128 + lane * 4;
    if (base >= n) continue;
    if (P.vec_ok)
    {
      const int j = (int)(base / P.W), i = (int)(base % P.W);
      if (j >= P.sy && j < P.sy + P.sh && i >= P.tx0 && i < P.tx1) continue;
      *reinterpret_cast<uint4*>(P.img_rgba + base) = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<float4*>(P.img_depth + base) = make_float4(1.001f, 1.001f, 1.001f, 1.001f);
      if (P.write_canvas)
      {
#pragma unroll
        for (int k = 0; k < 4; ++k) P.canvas_rgba[base + k] = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(P.canvas_depth + base) = make_float4(1.001f, 1.001f, 1.001f, 1.001f);
      }
    }
    else
    {
      for (int k = 0; k < 4; ++k)
      {
        const long long px = base + k;
        if (px >= n) break;
        const int j = (int)(px / P.W), i = (int)(px % P.W);
        if (j >= P.sy && j < P.sy + P.sh && i >= P.tx0 && i < P.tx1) continue;
        P.img_rgba[px] = make_uchar4(0, 0, 0, 0);
        P.img_depth[px] = 1.001f;
        if (P.write_canvas)
        {
          P.canvas_rgba[px] = make_float4(0.f, 0.f, 0.f, 0.f);
          P.canvas_depth[px] = 1.001f;
        }
      }
    }
  }
}

// MODE 5 (pushed frame): the quantised pixel does not stay on this GPU -- it is stored straight into
// the receive slot [owner of the pixel's 1024-pixel chunk][this rank] of the exchange (comm.cu), a
// posted NVLink store (a local one for the chunks this rank owns itself), so that the fold kernel
// later reads local memory only.
__device__ __forceinline__ void push_pixel(const TraceParams& P, long long pixel, uchar4 q, float depth)
{
  const unsigned px = (unsigned)pixel;
  const unsigned chunk = px >> 10;
  const unsigned turn = chunk / (unsigned)P.push_size;
  const unsigned owner = chunk - turn * (unsigned)P.push_size;
  const size_t at = (size_t)P.push_rank * P.push_share_px + (size_t)turn * 1024u + (px & 1023u);
  unsigned char* base = P.push_peers[owner];
  reinterpret_cast<uchar4*>(base + P.push_off_rgba)[at] = q;
  reinterpret_cast<float*>(base + P.push_off_depth)[at] = depth;
}

// the state of "the cell the ray is in": corner scalars in the pre-differenced form the
// reference keeps them in, the cell's lower-left point and inverse spacing
struct Cell
{
  float s0, s1m0, s2m3, s3, s4, s5m4, s6m7, s7; // point field (s0 doubles as the cell scalar)
  float blx, bly, blz;
  float isx, isy, isz;
  int cx, cy, cz;
};

// MARK: the demand-staging pre-pass (MODE 4): instead of gathering the corner scalars, flag the
// 128-byte lines of the field they live in (TraceParams::mark, one byte per line).
template <int KIND, typename FT, int ASSOC, typename IDX, bool MARK = false>
__device__ __forceinline__ void locate_and_load(const BlockDev& B, float px, float py, float pz,
                                                Cell& c, unsigned char* __restrict__ mark = nullptr)
{
  if (KIND == 0)
  {
    // UniformLocator::LocateCell
    float t0 = (px - B.min_point[0]) * B.inv_spacing[0];
    float t1 = (py - B.min_point[1]) * B.inv_spacing[1];
    float t2 = (pz - B.min_point[2]) * B.inv_spacing[2];
    if (t0 == (float)(B.dims[0] - 1)) t0 = (float)(B.dims[0] - 2);
    if (t1 == (float)(B.dims[1] - 1)) t1 = (float)(B.dims[1] - 2);
    if (t2 == (float)(B.dims[2] - 1)) t2 = (float)(B.dims[2] - 2);
    c.cx = (int)t0; c.cy = (int)t1; c.cz = (int)t2;
    c.isx = B.inv_spacing[0]; c.isy = B.inv_spacing[1]; c.isz = B.inv_spacing[2];
    // GetPoint(cellIndices[0]): origin + spacing*ijk evaluated in f64 and narrowed to f32 by
    // VTK-m; a single-rounding f32 fma gives the same bits whenever the exact sum fits in 53
    // bits (any grid whose |origin|/spacing ratio is below 2^18).
    c.blx = __fmaf_rn(B.spacing[0], (float)c.cx, B.origin[0]);
    c.bly = __fmaf_rn(B.spacing[1], (float)c.cy, B.origin[1]);
    c.blz = __fmaf_rn(B.spacing[2], (float)c.cz, B.origin[2]);
  }
  else
  {
    locate_axis_rect(B.axis[0], B.dims[0], B.max_point[0], px, c.cx, c.isx);
    locate_axis_rect(B.axis[1], B.dims[1], B.max_point[1], py, c.cy, c.isy);
    locate_axis_rect(B.axis[2], B.dims[2], B.max_point[2], pz, c.cz, c.isz);
    c.blx = __ldg(B.axis[0] + c.cx);
    c.bly = __ldg(B.axis[1] + c.cy);
    c.blz = __ldg(B.axis[2] + c.cz);
  }
  if (ASSOC == VR_POINT)
  {
    const IDX Nx = (IDX)B.dims[0], NxNy = (IDX)B.dims[0] * (IDX)B.dims[1];
    const IDX i0 = ((IDX)c.cz * (IDX)B.dims[1] + (IDX)c.cy) * Nx + (IDX)c.cx;
    const IDX i3 = i0 + Nx, i4 = i0 + NxNy, i7 = i4 + Nx;
    if (MARK)
    {
      constexpr int SH = sizeof(FT) == 4 ? 5 : 4; // elements per 128-byte line = 1 << SH
      const IDX rows[4] = { i0, i3, i4, i7 };
#pragma unroll
      for (int r = 0; r < 4; ++r)
      {
        const IDX l0 = rows[r] >> SH, l1 = (rows[r] + 1) >> SH;
        mark[l0] = 1;
        if (l1 != l0) mark[l1] = 1;
      }
      c.s0 = c.s3 = c.s4 = c.s7 = c.s6m7 = c.s5m4 = c.s1m0 = c.s2m3 = 0.f;
      return;
    }
    const float s0 = load_scalar<FT>(B.field, i0);
    const float s1 = load_scalar<FT>(B.field, i0 + 1);
    const float s3 = load_scalar<FT>(B.field, i3);
    const float s2 = load_scalar<FT>(B.field, i3 + 1);
    const float s4 = load_scalar<FT>(B.field, i4);
    const float s5 = load_scalar<FT>(B.field, i4 + 1);
    const float s7 = load_scalar<FT>(B.field, i7);
    const float s6 = load_scalar<FT>(B.field, i7 + 1);
    c.s0 = s0; c.s3 = s3; c.s4 = s4; c.s7 = s7;
    c.s6m7 = s6 - s7; c.s5m4 = s5 - s4; c.s1m0 = s1 - s0; c.s2m3 = s2 - s3;
  }
  else
  {
    const IDX ci = ((IDX)c.cz * (IDX)(B.dims[1] - 1) + (IDX)c.cy) * (IDX)(B.dims[0] - 1) + (IDX)c.cx;
    if (MARK)
    {
      constexpr int SH = sizeof(FT) == 4 ? 5 : 4;
      mark[ci >> SH] = 1;
      c.s0 = 0.f;
      return;
    }
    c.s0 = load_scalar<FT>(B.field, ci);
  }
}

// ---------------------------------------------------------------------------------------------
// Sparse march (uniform blocks, point fields, step >= 2.6 voxels -- the reference's default samples =
// 100 on every BASELINE config): the host has established (TraceParams::sparse) that every step of
// every ray leaves its cell, so each sample locates its cell from scratch -- exactly what the
// reference's "new cell" branch does -- and nothing is carried from one sample to the next except the
// position and the accumulated colour.  That makes a sample's eight gathers independent of all
// earlier samples: THREE samples are kept in flight per ray (slots A, B, C, rotated by unrolling, no
// register copies), against one in the general march.  float -> int of the non-negative cell
// coordinate is a round-toward-zero add of 2^23 (mantissa = integer part; exact below 2^22) on the
// FMA pipe instead of F2I/I2F on the quarter-rate conversion pipe.
__shared__ float4 s_lut[1024]; // the transfer function, one copy per CTA (every kernel of this file)

struct SparseSlot
{
  float s0, s1, s2, s3, s4, s5, s6, s7; // the eight corner scalars, raw
  float tx, ty, tz;                     // position inside the cell
};

// a gather that is skipped (result 0) when `on` is false: one predicated LDG, no branch
__device__ __forceinline__ float ldg_if(const float* p, bool on)
{
  float v;
  asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\t@q ld.global.nc.f32 %0, [%1];\n\t}"
      : "=f"(v) : "l"(p), "r"((int)on));
  return v;
}
__device__ __forceinline__ float ldg_if(const double* p, bool on)
{
  double v;
  asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\t@q ld.global.nc.f64 %0, [%1];\n\t}"
      : "=d"(v) : "l"(p), "r"((int)on));
  return (float)v;
}

// cell lookup + the eight gathers of one sample.  Branch-free: for a sample past the end of the ray
// (`on` false) the arithmetic runs on whatever the position is and only the loads are predicated off.
template <typename FT, typename IDX>
__device__ __forceinline__ void sparse_issue(const TraceParams& P, float px, float py, float pz, bool on, SparseSlot& s)
{
  const BlockDev& B = P.blk;
  // UniformLocator::LocateCell
  float t0 = (px - B.min_point[0]) * B.inv_spacing[0];
  float t1 = (py - B.min_point[1]) * B.inv_spacing[1];
  float t2 = (pz - B.min_point[2]) * B.inv_spacing[2];
  if (t0 == P.dims_m1[0]) t0 = P.dims_m2[0];
  if (t1 == P.dims_m1[1]) t1 = P.dims_m2[1];
  if (t2 == P.dims_m1[2]) t2 = P.dims_m2[2];
  const float kMagic = 8388608.f; // 2^23
  const float r0 = __fadd_rz(t0, kMagic), r1 = __fadd_rz(t1, kMagic), r2 = __fadd_rz(t2, kMagic);
  const int cx = __float_as_int(r0) - 0x4B000000, cy = __float_as_int(r1) - 0x4B000000,
            cz = __float_as_int(r2) - 0x4B000000;
  const float fcx = r0 - kMagic, fcy = r1 - kMagic, fcz = r2 - kMagic; // exact
  // GetPoint(cell): origin + spacing * ijk (see locate_and_load)
  const float blx = __fmaf_rn(B.spacing[0], fcx, B.origin[0]);
  const float bly = __fmaf_rn(B.spacing[1], fcy, B.origin[1]);
  const float blz = __fmaf_rn(B.spacing[2], fcz, B.origin[2]);
  s.tx = (px - blx) * B.inv_spacing[0];
  s.ty = (py - bly) * B.inv_spacing[1];
  s.tz = (pz - blz) * B.inv_spacing[2];
  // four row pointers, each one widening multiply-add away from the previous
  const IDX i0 = ((IDX)cz * (IDX)B.dims[1] + (IDX)cy) * (IDX)B.dims[0] + (IDX)cx;
  const FT* f0 = reinterpret_cast<const FT*>(B.field) + i0;
  const FT* f3 = f0 + B.dims[0];
  const FT* f4 = f0 + P.slice_elems;
  const FT* f7 = f4 + B.dims[0];
  s.s0 = ldg_if(f0, on);
  s.s1 = ldg_if(f0 + 1, on);
  s.s3 = ldg_if(f3, on);
  s.s2 = ldg_if(f3 + 1, on);
  s.s4 = ldg_if(f4, on);
  s.s5 = ldg_if(f4 + 1, on);
  s.s7 = ldg_if(f7, on);
  s.s6 = ldg_if(f7 + 1, on);
}

// interpolate + classify + blend one sample; returns true when the ray is saturated (alpha >= 1)
__device__ __forceinline__ bool sparse_shade(const TraceParams& P, const SparseSlot& s, float& c0, float& c1, float& c2,
                                             float& c3)
{
  const float cms_f = P.cms_f;
  const float s6m7 = s.s6 - s.s7, s5m4 = s.s5 - s.s4, s1m0 = s.s1 - s.s0, s2m3 = s.s2 - s.s3;
  const float l76 = s.s7 + s.tx * s6m7;
  const float l45 = s.s4 + s.tx * s5m4;
  const float ltop = l45 + s.ty * (l76 - l45);
  const float l01 = s.s0 + s.tx * s1m0;
  const float l32 = s.s3 + s.tx * s2m3;
  const float lbot = l01 + s.ty * (l32 - l01);
  float v = lbot + s.tz * (ltop - lbot);
  v = (v - P.range_min) * P.inv_delta_scalar;
  const float raw = v * cms_f;
  float fidx = fminf(fmaxf(raw, 0.f), cms_f); // (see the general march for the NaN / overflow cases)
  if (raw >= 9.2233720e18f) fidx = 0.f;
  const int idx = __float_as_int(__fadd_rz(fidx, 8388608.f)) - 0x4B000000;
  const float4 sc = s_lut[idx];
  const float alpha = sc.w * (1.f - c3);
  c0 = c0 + sc.x * alpha;
  c1 = c1 + sc.y * alpha;
  c2 = c2 + sc.z * alpha;
  c3 = alpha + c3;
  return c3 >= 1.f;
}

// ---------------------------------------------------------------------------------------------
// Brick march (uniform blocks, f32 point fields, step <= 2 voxels -- dense sampling, e.g. samples = 887 on
// the 512^3 block): consecutive samples of a ray, and the rays of a warp's 8x4 pixel packet, walk through
// neighbouring cells, so the warp stages the sub-volume its next kBrickSteps samples will touch in shared
// memory with ONE TMA tensor copy (cp.async.bulk.tensor.3d, a 16x10x10 box of the field addressed through a
// CUtensorMap the host encodes per block; completion on an mbarrier; out-of-range parts of the box are
// zero-filled by the hardware and never read) and fetches the eight corners of every sample with LDS at
// constant offsets from one address.  Two bricks per warp are in flight: while the lanes sample brick w,
// the copy engine fills brick w+1.  A packet ends when a warp ballot finds no lane with a live ray (early
// termination, exit, miss).  The brick is anchored at the low corner of the axis-aligned box around the
// cells of the window's first and last sample of every live lane, padded by a cell (rounding, "keep the
// cell while 0 <= t <= 1"); the window is halved until that box (plus the upper corners) fits.  A sample
// whose cell is not in the brick gathers from global memory like the general march.  The arithmetic per
// sample is the general march's, so the pixels are the same bits.
// brick = kBrickX x kBrickY x kBrickZ points.  TMA wants the box to start on a 16-byte boundary of the
// innermost (x) dimension -- measured: any other x coordinate raises "illegal instruction", y and z are
// free, negative coordinates included (profiles/experiments/tma_probe.cu) -- so x is anchored at a multiple
// of 4 points and the box is 4 points wider there than the 12 the window needs.
constexpr int kBrickX = 16, kBrickY = 10, kBrickZ = 10;
constexpr int kBrickFloats = kBrickX * kBrickY * kBrickZ; // 6400 bytes = 50 * 128
constexpr int kBrickSteps = 8;                           // samples per window (halved until the brick fits)

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes)
{
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity)
{
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_box(unsigned dst, const void* tmap, int x, int y, int z, unsigned bar)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
struct Brick
{
  int x0, y0, z0; // first point staged
  int on;         // a brick was requested for this window
};

// the brick of the window [a, a + k) steps ahead of the lanes' current positions (warp-uniform result)
__device__ __forceinline__ Brick brick_of_window(const TraceParams& P, bool live, float px, float py, float pz, float sx,
                                                 float sy, float sz, float a, int& k)
{
  const BlockDev& B = P.blk;
  Brick b;
  b.on = 0;
  b.x0 = b.y0 = b.z0 = 0;
  if (!__any_sync(0xffffffffu, live)) return b;
  for (;; k >>= 1)
  {
    const float e = a + (float)(k - 1);
    int lo[3], hi[3];
    const float q0[3] = { px + a * sx, py + a * sy, pz + a * sz }, q1[3] = { px + e * sx, py + e * sy, pz + e * sz };
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
      const int c0 = __float2int_rd((q0[d] - B.min_point[d]) * B.inv_spacing[d]);
      const int c1 = __float2int_rd((q1[d] - B.min_point[d]) * B.inv_spacing[d]);
      lo[d] = __reduce_min_sync(0xffffffffu, live ? min(c0, c1) - 1 : 0x7fffffff);
      hi[d] = __reduce_max_sync(0xffffffffu, live ? max(c0, c1) + 2 : -0x7fffffff); // + upper corner + a cell
    }
    const int x0 = lo[0] & ~3; // (two's complement: rounds towards -inf, also below zero)
    if (hi[0] - x0 < kBrickX && hi[1] - lo[1] < kBrickY && hi[2] - lo[2] < kBrickZ)
    {
      b.x0 = x0; b.y0 = lo[1]; b.z0 = lo[2];
      b.on = 1;
      return b;
    }
    if (k == 1) return b; // does not fit even for one step: these samples gather from global memory
  }
}

// `live`: this lane has a ray whose current position (px,py,pz at `distance`) is its first valid sample
template <typename IDX>
__device__ __forceinline__ void brick_march(const TraceParams& P, bool live, float px, float py, float pz, float sx,
                                            float sy, float sz, float distance, float max_distance, float& c0, float& c1,
                                            float& c2, float& c3, unsigned& my_samples, unsigned wbuf, unsigned wbar,
                                            unsigned& parity)
{
  const BlockDev& B = P.blk;
  const int lane = threadIdx.x & 31;
  if (!__any_sync(0xffffffffu, live)) return;
  const float minx = B.min_point[0], miny = B.min_point[1], minz = B.min_point[2];
  const float maxx = B.max_point[0], maxy = B.max_point[1], maxz = B.max_point[2];
  const float sd = P.sample_dist;
  const void* tmap = B.tmap;
  // the cell the ray is in (general march semantics: kept while 0 <= t <= 1)
  bool fresh = true;
  int cx = 0, cy = 0, cz = 0;
  float blx = 0.f, bly = 0.f, blz = 0.f;

  int k_cur = kBrickSteps;
  Brick cur = brick_of_window(P, live, px, py, pz, sx, sy, sz, 0.f, k_cur);
  unsigned slot = 0; // buffer / barrier of the current window: wbuf + slot * kBrickFloats * 4, wbar + slot * 8
  if (cur.on && lane == 0)
  {
    mbar_arrive_expect_tx(wbar, kBrickFloats * 4u);
    tma_load_box(wbuf, tmap, cur.x0, cur.y0, cur.z0, wbar);
  }
  for (;;)
  {
    // ---- the copy engine starts on the next window while this one is sampled
    int k_next = kBrickSteps;
    const bool live_next = live && (distance + (float)k_cur * sd < max_distance);
    const Brick nxt = brick_of_window(P, live_next, px, py, pz, sx, sy, sz, (float)k_cur, k_next);
    if (nxt.on && lane == 0)
    {
      const unsigned nb = wbar + (slot ^ 1u) * 8u;
      mbar_arrive_expect_tx(nb, kBrickFloats * 4u);
      tma_load_box(wbuf + (slot ^ 1u) * (kBrickFloats * 4u), tmap, nxt.x0, nxt.y0, nxt.z0, nb);
    }
    if (cur.on)
    {
      while (!mbar_try_wait(wbar + slot * 8u, (parity >> slot) & 1u)) {}
      parity ^= 1u << slot;
    }
    const unsigned bk = wbuf + slot * (kBrickFloats * 4u);
    for (int j = 0; j < k_cur; ++j)
    {
      if (live)
      {
        float tx = (px - blx) * B.inv_spacing[0], ty = (py - bly) * B.inv_spacing[1], tz = (pz - blz) * B.inv_spacing[2];
        if (fresh || fmaxf(tx, fmaxf(ty, tz)) > 1.f || fminf(tx, fminf(ty, tz)) < 0.f)
        {
          // UniformLocator::LocateCell (same arithmetic as locate_and_load)
          float t0 = (px - minx) * B.inv_spacing[0], t1 = (py - miny) * B.inv_spacing[1], t2 = (pz - minz) * B.inv_spacing[2];
          if (t0 == P.dims_m1[0]) t0 = P.dims_m2[0];
          if (t1 == P.dims_m1[1]) t1 = P.dims_m2[1];
          if (t2 == P.dims_m1[2]) t2 = P.dims_m2[2];
          cx = (int)t0; cy = (int)t1; cz = (int)t2;
          blx = __fmaf_rn(B.spacing[0], (float)cx, B.origin[0]);
          bly = __fmaf_rn(B.spacing[1], (float)cy, B.origin[1]);
          blz = __fmaf_rn(B.spacing[2], (float)cz, B.origin[2]);
          tx = (px - blx) * B.inv_spacing[0]; ty = (py - bly) * B.inv_spacing[1]; tz = (pz - blz) * B.inv_spacing[2];
          fresh = false;
        }
        float s0, s1, s2, s3, s4, s5, s6, s7;
        const int ux = cx - cur.x0, uy = cy - cur.y0, uz = cz - cur.z0;
        if (cur.on && (unsigned)ux < (unsigned)(kBrickX - 1) && (unsigned)uy < (unsigned)(kBrickY - 1) &&
            (unsigned)uz < (unsigned)(kBrickZ - 1))
        {
          // row pitch 16 floats = 64 bytes, slab pitch 160 floats = 640 bytes: all eight corners at constant
          // offsets from one address
          static_assert(kBrickX == 16 && kBrickY == 10, "the LDS offsets below assume a 16 x 10 x n brick");
          const unsigned q = bk + (unsigned)((uz * kBrickY + uy) * kBrickX + ux) * 4u;
          asm volatile("ld.shared.f32 %0, [%8];\n\tld.shared.f32 %1, [%8+4];\n\t"
                       "ld.shared.f32 %3, [%8+64];\n\tld.shared.f32 %2, [%8+68];\n\t"
                       "ld.shared.f32 %4, [%8+640];\n\tld.shared.f32 %5, [%8+644];\n\t"
                       "ld.shared.f32 %7, [%8+704];\n\tld.shared.f32 %6, [%8+708];"
                       : "=f"(s0), "=f"(s1), "=f"(s2), "=f"(s3), "=f"(s4), "=f"(s5), "=f"(s6), "=f"(s7)
                       : "r"(q) : "memory");
        }
        else
        {
          const IDX Nx = (IDX)B.dims[0], NxNy = (IDX)B.dims[0] * (IDX)B.dims[1];
          const float* f = reinterpret_cast<const float*>(B.field) + (((IDX)cz * (IDX)B.dims[1] + (IDX)cy) * Nx + (IDX)cx);
          s0 = __ldg(f); s1 = __ldg(f + 1); s3 = __ldg(f + Nx); s2 = __ldg(f + Nx + 1);
          s4 = __ldg(f + NxNy); s5 = __ldg(f + NxNy + 1); s7 = __ldg(f + NxNy + Nx); s6 = __ldg(f + NxNy + Nx + 1);
        }
        const float s6m7 = s6 - s7, s5m4 = s5 - s4, s1m0 = s1 - s0, s2m3 = s2 - s3;
        const float l76 = s7 + tx * s6m7;
        const float l45 = s4 + tx * s5m4;
        const float ltop = l45 + ty * (l76 - l45);
        const float l01 = s0 + tx * s1m0;
        const float l32 = s3 + tx * s2m3;
        const float lbot = l01 + ty * (l32 - l01);
        float v = lbot + tz * (ltop - lbot);
        v = (v - P.range_min) * P.inv_delta_scalar;
        const float raw = v * P.cms_f;
        float fidx = fminf(fmaxf(raw, 0.f), P.cms_f);
        if (raw >= 9.2233720e18f) fidx = 0.f;
        const float4 sc = s_lut[(int)fidx];
        const float alpha = sc.w * (1.f - c3);
        c0 = c0 + sc.x * alpha;
        c1 = c1 + sc.y * alpha;
        c2 = c2 + sc.z * alpha;
        c3 = alpha + c3;
        ++my_samples;
        // next sample of this ray, or the end of it (saturated / left the block / reached the depth limit)
        px = px + sx; py = py + sy; pz = pz + sz;
        distance = distance + sd;
        live = !(c3 >= 1.f) && !(px < minx || px > maxx) && !(py < miny || py > maxy) && !(pz < minz || pz > maxz) &&
               distance < max_distance;
      }
    }
    __syncwarp(); // every lane is done reading this window's brick before its buffer is filled again
    // ---- ray-packet termination: one ballot for the whole warp
    const bool more = __any_sync(0xffffffffu, live);
    if (!more)
    {
      if (nxt.on)
      {
        // a brick is still in flight: let it land so that barrier and buffer are free for the next packet
        while (!mbar_try_wait(wbar + (slot ^ 1u) * 8u, (parity >> (slot ^ 1u)) & 1u)) {}
        parity ^= 1u << (slot ^ 1u);
      }
      return;
    }
    cur = nxt;
    k_cur = k_next;
    slot ^= 1u;
  }
}

// One work item of a trace: tile `tile` of the launch P describes (an 8x4 pixel tile of the block's screen
// rectangle or, in MODE 2, a chunk of the frame's complement to clear).  P is the kernel's __grid_constant__
// for the one-block kernels and a per-warp copy in shared memory for trace_multi_kernel.
template <int KIND, typename FT, int ASSOC, int MODE, typename IDX, int MARCH>
__device__ __forceinline__ void trace_tile(const TraceParams& P, unsigned tile, unsigned& my_samples, unsigned& brick_parity,
                                           unsigned char* s_dyn)
{
  const BlockDev& B = P.blk;
  const int lane = threadIdx.x & 31;
  const int lx = lane & (kTileW - 1), ly = lane >> 3;
  const unsigned n_tiles = (unsigned)(P.tiles_x * P.tiles_y);
  const float cms_f = P.cms_f;
  const float minx = B.min_point[0], miny = B.min_point[1], minz = B.min_point[2];
  const float maxx = B.max_point[0], maxy = B.max_point[1], maxz = B.max_point[2];
  const float sd = P.sample_dist;

#define VR_INSIDE(x, y, z) \
  (!((x) < minx || (x) > maxx) && !((y) < miny || (y) > maxy) && !((z) < minz || (z) > maxz))

  {
    if (MODE == 2 && tile >= n_tiles)
    {
      // ---------------- Canvas::Clear for everything outside the traced rectangle, in the
      // frame's final formats (Image::Init of a cleared canvas: colour 0, depth 1.001)
      clear_chunk(P, tile - n_tiles, lane);
      return;
    }
    // tiles are handed out centre rows first, and centre-out within a row: rays through the middle of the
    // block's screen rectangle are the long ones, so the kernel's tail -- a warp resident at a time draws
    // only a handful of tiles -- is made of the short edge rays
    const int rj = (int)(tile / (unsigned)P.tiles_x), ri = (int)(tile % (unsigned)P.tiles_x);
    const int tj = P.tile_order ? rj : centre_out(rj, P.tiles_y), ti = P.tile_order ? ri : centre_out(ri, P.tiles_x);
    const int i = P.tx0 + ti * kTileW + lx;
    const int j = P.sy + tj * kTileH + ly;
    // tx0 <= sx: in MODE 2 the traced rectangle is widened to 4-pixel boundaries so that its
    // complement can be cleared (and, across GPUs, skipped) in 16-byte groups
    const bool in_rect = (i < P.tx1) && (j < P.sy + P.sh);
    const bool in_subset = in_rect && (i >= P.sx) && (i < P.sx + P.sw);
    const long long pixel = (long long)j * P.W + i;
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
    float max_distance = __int_as_float(0x7f800000);
    // (brick march: the march itself runs after the divergent ray set-up, with the whole warp present)
    bool b_live = false;
    float b_px = 0.f, b_py = 0.f, b_pz = 0.f, b_sx = 0.f, b_sy = 0.f, b_sz = 0.f, b_dist = 0.f;
    // K7 needs these after the march
    float dx = 0.f, dy = 0.f, dz = 0.f, distance0 = 0.f;
    if (in_subset)
    {
      // ---------------- K1: PerspectiveRayGen
      {
        const float fx = (2.f * (float)i - (float)P.W) / 2.0f;
        const float fy = (2.f * (float)j - (float)P.H) / 2.0f;
        dx = P.nlook[0] + P.delta_x[0] * fx + P.delta_y[0] * fy;
        dy = P.nlook[1] + P.delta_x[1] * fx + P.delta_y[1] * fy;
        dz = P.nlook[2] + P.delta_x[2] * fx + P.delta_y[2] * fy;
        if (dx == 0.f) dx += 0.0000001f;
        if (dy == 0.f) dy += 0.0000001f;
        if (dz == 0.f) dz += 0.0000001f;
        const float dot = dx * dx + dy * dy + dz * dz;
        const float sq = sqrtf(dot);
        dx = dx / sq; dy = dy / sq; dz = dz / sq;
      }
      const float ox = P.origin[0], oy = P.origin[1], oz = P.origin[2];
      float min_distance = 0.f;

      // ---------------- K2: RayMapCanvas
      if (P.use_depth)
      {
        float p0 = (float)(pixel % P.W), p1 = (float)(pixel / P.W);
        float p2 = P.canvas_depth[pixel];
        p0 = p0 * P.dbl_inv_w - 1.f;
        p1 = p1 * P.dbl_inv_h - 1.f;
        p2 = 2.f * p2 - 1.f;
        p2 -= 0.00001f;
        const float* m = P.inv_pv;
        const float q0 = m[0] * p0 + m[1] * p1 + m[2] * p2 + m[3] * 1.f;
        const float q1 = m[4] * p0 + m[5] * p1 + m[6] * p2 + m[7] * 1.f;
        const float q2 = m[8] * p0 + m[9] * p1 + m[10] * p2 + m[11] * 1.f;
        const float q3 = m[12] * p0 + m[13] * p1 + m[14] * p2 + m[15] * 1.f;
        const float rx = q0 / q3 - ox, ry = q1 / q3 - oy, rz = q2 / q3 - oz;
        max_distance = sqrtf(rx * rx + ry * ry + rz * rz);
      }

      // ---------------- K3: CalcRayStart
      {
        const float ix = rcp_safe(dx), iy = rcp_safe(dy), iz = rcp_safe(dz);
        const float odx = ox * ix, ody = oy * iy, odz = oz * iz;
        const float xmin = P.bmin[0] * ix - odx, ymin = P.bmin[1] * iy - ody, zmin = P.bmin[2] * iz - odz;
        const float xmax = P.bmax[0] * ix - odx, ymax = P.bmax[1] * iy - ody, zmax = P.bmax[2] * iz - odz;
        min_distance =
          fmaxf(fmaxf(fmaxf(fminf(ymin, ymax), fminf(xmin, xmax)), fminf(zmin, zmax)), min_distance);
        const float exit_distance = fminf(fminf(fmaxf(ymin, ymax), fmaxf(xmin, xmax)), fmaxf(zmin, zmax));
        max_distance = fminf(max_distance, exit_distance);
        if (max_distance < min_distance) min_distance = -1.f;
        else distance0 = min_distance;
      }

      // ---------------- K4/K5/K6: Sampler
      if (min_distance != -1.f)
      {
        float distance = min_distance + P.mesh_eps;
        float px = ox + distance * dx, py = oy + distance * dy, pz = oz + distance * dz;
        while (!VR_INSIDE(px, py, pz) && distance < max_distance)
        {
          distance += sd;
          px = ox + distance * dx; py = oy + distance * dy; pz = oz + distance * dz;
        }
        if (MARCH == 2)
        {
          // ---- brick march: only note where the ray starts; the warp marches together below
          b_live = VR_INSIDE(px, py, pz) && distance < max_distance;
          b_px = px; b_py = py; b_pz = pz; b_dist = distance;
          b_sx = sd * dx; b_sy = sd * dy; b_sz = sd * dz;
        }
        else if (MARCH == 1 && VR_INSIDE(px, py, pz) && distance < max_distance)
        {
          // ---- sparse march: three samples in flight (slots A, B, C)
          const float stepx = sd * dx, stepy = sd * dy, stepz = sd * dz;
          SparseSlot A, B_, C;
          bool vB, vC, vA;
          sparse_issue<FT, IDX>(P, px, py, pz, true, A);
#define VR_ADVANCE(valid_prev, valid_out, slot)                                              \
          px = px + stepx; py = py + stepy; pz = pz + stepz;                                 \
          distance = distance + sd;                                                          \
          valid_out = (valid_prev) && VR_INSIDE(px, py, pz) && distance < max_distance;      \
          sparse_issue<FT, IDX>(P, px, py, pz, valid_out, slot);
          VR_ADVANCE(true, vB, B_)
          VR_ADVANCE(vB, vC, C)
          for (;;)
          {
            ++my_samples;
            if (sparse_shade(P, A, c0, c1, c2, c3) || !vB) break;
            VR_ADVANCE(vC, vA, A)
            ++my_samples;
            if (sparse_shade(P, B_, c0, c1, c2, c3) || !vC) break;
            VR_ADVANCE(vA, vB, B_)
            ++my_samples;
            if (sparse_shade(P, C, c0, c1, c2, c3) || !vA) break;
            VR_ADVANCE(vB, vC, C)
          }
#undef VR_ADVANCE
        }
        else if (MARCH == 0 && VR_INSIDE(px, py, pz) && distance < max_distance)
        {
          const float stepx = sd * dx, stepy = sd * dy, stepz = sd * dz;
          // first sample: the reference enters its loop with newCell = true
          Cell cur;
          cur.cx = cur.cy = cur.cz = 0;
          cur.isx = cur.isy = cur.isz = 0.f;
          locate_and_load<KIND, FT, ASSOC, IDX, MODE == 4>(B, px, py, pz, cur, P.mark);
          float tx = (px - cur.blx) * cur.isx, ty = (py - cur.bly) * cur.isy, tz = (pz - cur.blz) * cur.isz;
          for (;;)
          {
            // ---- sample k+1: position, loop condition, cell change, gathers (issued early)
            const float npx = px + stepx, npy = py + stepy, npz = pz + stepz;
            const float ndist = distance + sd;
            const bool next_ok = VR_INSIDE(npx, npy, npz) && ndist < max_distance;
            Cell nxt = cur;
            float ntx = (npx - cur.blx) * cur.isx, nty = (npy - cur.bly) * cur.isy,
                  ntz = (npz - cur.blz) * cur.isz;
            if (next_ok)
            {
              const float mint = fminf(ntx, fminf(nty, ntz));
              const float maxt = fmaxf(ntx, fmaxf(nty, ntz));
              if (maxt > 1.f || mint < 0.f)
              {
                locate_and_load<KIND, FT, ASSOC, IDX, MODE == 4>(B, npx, npy, npz, nxt, P.mark);
                ntx = (npx - nxt.blx) * nxt.isx;
                nty = (npy - nxt.bly) * nxt.isy;
                ntz = (npz - nxt.blz) * nxt.isz;
              }
            }
            if (MODE == 4)
            {
              // pre-pass: every sample up to the exit is visited (no opacity, so no early exit): a
              // superset of what the real trace will touch
              if (!next_ok) break;
              cur = nxt;
              px = npx; py = npy; pz = npz;
              distance = ndist;
              tx = ntx; ty = nty; tz = ntz;
              continue;
            }
            // ---- sample k: interpolate, classify, blend
            float v;
            if (ASSOC == VR_POINT)
            {
              const float l76 = cur.s7 + tx * cur.s6m7;
              const float l45 = cur.s4 + tx * cur.s5m4;
              const float ltop = l45 + ty * (l76 - l45);
              const float l01 = cur.s0 + tx * cur.s1m0;
              const float l32 = cur.s3 + tx * cur.s2m3;
              const float lbot = l01 + ty * (l32 - l01);
              v = lbot + tz * (ltop - lbot);
            }
            else
              v = cur.s0;
            v = (v - P.range_min) * P.inv_delta_scalar;
            // static_cast<vtkm::Id>(v * size) then clamp to [0, size]: trunc and clamp commute;
            // x86's cvttss2si turns NaN and >= 2^63 into INT64_MIN, which the clamp maps to 0.
            const float raw = v * cms_f;
            float fidx = fminf(fmaxf(raw, 0.f), cms_f);
            if (raw >= 9.2233720e18f) fidx = 0.f;
            const float4 sc = s_lut[(int)fidx];
            const float alpha = sc.w * (1.f - c3);
            c0 = c0 + sc.x * alpha;
            c1 = c1 + sc.y * alpha;
            c2 = c2 + sc.z * alpha;
            c3 = alpha + c3;
            ++my_samples;
            if (c3 >= 1.f) break;
            if (!next_ok) break;
            cur = nxt;
            px = npx; py = npy; pz = npz;
            distance = ndist;
            tx = ntx; ty = nty; tz = ntz;
          }
        }
        if (MARCH != 2) { c0 = fminf(c0, 1.f); c1 = fminf(c1, 1.f); c2 = fminf(c2, 1.f); c3 = fminf(c3, 1.f); }
      }
    }
    if (MARCH == 2)
    {
      // this warp's two brick buffers (128-byte aligned: 6912 = 54 * 128) and its two mbarriers
      const unsigned wbuf = smem_addr(s_dyn) + (threadIdx.x >> 5) * (2u * kBrickFloats * 4u);
      const unsigned wbar = smem_addr(s_dyn) + (kThreads / 32) * (2u * kBrickFloats * 4u) + (threadIdx.x >> 5) * 16u;
      const bool had_ray = b_live; // (a ray that never reaches a valid sample keeps colour 0, as in the other marches)
      brick_march<IDX>(P, b_live, b_px, b_py, b_pz, b_sx, b_sy, b_sz, b_dist, max_distance, c0, c1, c2, c3, my_samples,
                       wbuf, wbar, brick_parity);
      if (had_ray) { c0 = fminf(c0, 1.f); c1 = fminf(c1, 1.f); c2 = fminf(c2, 1.f); c3 = fminf(c3, 1.f); }
    }
    if (in_subset)
    {
      const float ox = P.origin[0], oy = P.origin[1], oz = P.origin[2];
      if (MODE == 0)
      {
        // ---------------- K7: SurfaceConverter (blend over canvas, projected entry depth)
        const float ix_ = ox + distance0 * dx, iy_ = oy + distance0 * dy, iz_ = oz + distance0 * dz;
        const float* m = P.pv;
        const float n2 = m[8] * ix_ + m[9] * iy_ + m[10] * iz_ + m[11] * 1.f;
        const float n3 = m[12] * ix_ + m[13] * iy_ + m[14] * iz_ + m[15] * 1.f;
        const float depth = 0.5f * (n2 / n3) + 0.5f;
        const float4 in = P.canvas_rgba[pixel];
        const float a = 1.f - c3;
        float4 out;
        out.x = c0 + in.x * a;
        out.y = c1 + in.y * a;
        out.z = c2 + in.z * a;
        out.w = in.w * a + c3;
        out.x = fminf(1.f, fmaxf(out.x, 0.f));
        out.y = fminf(1.f, fmaxf(out.y, 0.f));
        out.z = fminf(1.f, fmaxf(out.z, 0.f));
        out.w = fminf(1.f, fmaxf(out.w, 0.f));
        P.canvas_depth[pixel] = depth;
        P.canvas_rgba[pixel] = out;
      }
      if (MODE == 2 || MODE == 5)
      {
        // ---------------- K7 over a cleared canvas (in = 0), fused with Image::Init
        // (Image.hpp:80-113: truncating uint8, depth < 0 -> |d|) and, for a single rank,
        // Renderer::ImageToCanvas (Renderer.cpp:265-283: k * (1/255))
        const float ix_ = ox + distance0 * dx, iy_ = oy + distance0 * dy, iz_ = oz + distance0 * dz;
        const float* m = P.pv;
        const float n2 = m[8] * ix_ + m[9] * iy_ + m[10] * iz_ + m[11] * 1.f;
        const float n3 = m[12] * ix_ + m[13] * iy_ + m[14] * iz_ + m[15] * 1.f;
        float depth = 0.5f * (n2 / n3) + 0.5f;
        const float a = 1.f - c3;
        float4 out; // same expressions as the unfused path so every bit agrees (0*a may be NaN for a = inf only)
        out.x = fminf(1.f, fmaxf(c0 + 0.f * a, 0.f));
        out.y = fminf(1.f, fmaxf(c1 + 0.f * a, 0.f));
        out.z = fminf(1.f, fmaxf(c2 + 0.f * a, 0.f));
        out.w = fminf(1.f, fmaxf(0.f * a + c3, 0.f));
        const uchar4 q = make_uchar4(quant_u8(out.x), quant_u8(out.y), quant_u8(out.z), quant_u8(out.w));
        depth = depth < 0.f ? fabsf(depth) : depth;
        if (MODE == 5)
          push_pixel(P, pixel, q, depth);
        else
        {
          P.img_rgba[pixel] = q;
          P.img_depth[pixel] = depth;
          if (P.write_canvas)
          {
            const float k = 1.f / 255.f;
            P.canvas_rgba[pixel] = make_float4((float)q.x * k, (float)q.y * k, (float)q.z * k, (float)q.w * k);
            P.canvas_depth[pixel] = depth;
          }
        }
      }
    } // in_subset
    else if (MODE == 5 && in_rect)
      push_pixel(P, pixel, make_uchar4(0, 0, 0, 0), 1.001f); // the widened rectangle's padding columns
    else if (MODE == 2 && in_rect)
    {
      P.img_rgba[pixel] = make_uchar4(0, 0, 0, 0);
      P.img_depth[pixel] = 1.001f;
      if (P.write_canvas)
      {
        P.canvas_rgba[pixel] = make_float4(0.f, 0.f, 0.f, 0.f);
        P.canvas_depth[pixel] = 1.001f;
      }
    }

    if (MODE == 3)
    {
      // ---------------- path B without lists: the ray's result goes to its place in the block's
      // dense layer; alpha = 0 marks what the reference drops (alpha < 0.001, VolumeRenderer.cpp:270)
      if (in_subset)
      {
        const size_t e = P.layer_base + (size_t)(j - P.sy) * P.sw + (size_t)(i - P.sx);
        const bool keep = !(c3 < 0.001f);
        const float4 entry = keep ? make_float4(c0, c1, c2, c3) : make_float4(0.f, 0.f, 0.f, 0.f);
        P.layer_rgba[e] = entry;
        P.layer_depth[e] = max_distance; // rays.MaxDistance: the exit distance (:262-263)
        if (P.lpush_peers)
        {
          // N ranks: the entry also goes straight to the rank that folds this pixel's tile (posted NVLink
          // stores, overlapped with the rest of the trace), same index, receive pool of source lpush_rank
          const int owner = (int)(((unsigned)(j >> 3) * (unsigned)P.lpush_tpr + (unsigned)(i >> 5)) % (unsigned)P.lpush_size);
          unsigned char* base = P.lpush_peers[owner];
          const size_t at = (size_t)P.lpush_rank * P.lpush_stride + e;
          reinterpret_cast<float4*>(base + P.lpush_off_rgba)[at] = entry;
          reinterpret_cast<float*>(base + P.lpush_off_depth)[at] = max_distance;
        }
      }
    }
    if (MODE == 1)
    {
      // ---------------- path B: keep rays with alpha >= 0.001 (VolumeRenderer.cpp:270-283);
      // depth = rays.MaxDistance (exit distance).  One atomic per warp, slots by ballot rank.
      const bool emit = in_subset && !(c3 < 0.001f);
      const unsigned mask = __ballot_sync(0xffffffffu, emit);
      if (mask)
      {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(P.partial_count, (unsigned long long)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, 0);
        const unsigned long long slot = base + __popc(mask & ((1u << lane) - 1u));
        if (emit && slot < P.partial_capacity)
        {
          vr_partial q;
          q.pixel_id = (int)pixel;
          q.depth = max_distance;
          q.rgb[0] = c0; q.rgb[1] = c1; q.rgb[2] = c2;
          q.alpha = c3;
          P.partials[slot] = q;
        }
      }
    }
  }
#undef VR_INSIDE
}

// MARCH: 0 = general march (any block, any step), 1 = sparse march, 2 = brick march (see above); each is
// its own kernel so that each gets the register budget it needs: 56 / 72 / 96 registers per thread
// (9 / 7 / 3 resident CTAs per SM; the brick march is bounded by its shared-memory bricks anyway).
template <int KIND, typename FT, int ASSOC, int MODE, typename IDX, int MARCH>
__global__ void __launch_bounds__(kThreads, MARCH == 0 ? VR_MIN_BLOCKS : (MARCH == 1 ? 7 : 3))
trace_kernel(const __grid_constant__ TraceParams P)
{
  extern __shared__ __align__(128) unsigned char s_dyn[]; // brick march: per-warp brick buffers + mbarriers
  unsigned brick_parity = 0;
  if (MARCH == 2)
  {
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(s_dyn + (size_t)(kThreads / 32) * 2 * kBrickFloats * 4);
    if (threadIdx.x < (kThreads / 32) * 2) mbar_init(smem_addr(bars + threadIdx.x), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
  }
  if (MODE != 4) // the staging pre-pass never classifies
  {
    for (int i = threadIdx.x; i < P.lut_size; i += kThreads) s_lut[i] = __ldg(P.lut + i);
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const unsigned n_work = (unsigned)(P.tiles_x * P.tiles_y) + (MODE == 2 ? (unsigned)P.n_clear_chunks : 0u);
  unsigned my_samples = 0;
  for (;;)
  {
    unsigned tile = 0;
    if (lane == 0) tile = atomicAdd(P.tile_counter, 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= n_work) break;
    trace_tile<KIND, FT, ASSOC, MODE, IDX, MARCH>(P, tile, my_samples, brick_parity, s_dyn);
  }
  if (P.end_stamp && threadIdx.x == 0) atomicMax(P.end_stamp, global_ns());
  if (P.sample_counter)
  {
    // one atomic per warp
    unsigned long long s = my_samples;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && s) atomicAdd(P.sample_counter, s);
  }
}

// Path B with MANY blocks per rank (c5: 512 blocks of 128^3 at 4096^2): one persistent launch over the
// (block, tile) work items of ALL the rank's blocks instead of one launch per block -- 512 launches of ~4000
// tiles each leave the GPU waiting for the host (5.6 us per launch) and every launch's CTAs reload the 16 KiB
// table.  table[b] is block b's TraceParams (device memory), tile_end[b] the running sum of the blocks' tile
// counts; a warp draws kMultiChunk consecutive work items per atomic, finds their block by bisection and keeps
// that block's parameters in its own slot of shared memory (re-read only when the block changes).  Same
// trace_tile, same arithmetic, same order of operations per ray as the one-block kernel: the same bits.
#ifndef VR_MULTI_SPARSE_CTAS
#define VR_MULTI_SPARSE_CTAS 6 // resident CTAs per SM of the sparse-march variant (A/B: profiles/experiments)
#endif
#ifndef VR_MULTI_CHUNK
#define VR_MULTI_CHUNK 8
#endif
constexpr unsigned kMultiChunk = VR_MULTI_CHUNK;
template <int KIND, typename FT, int ASSOC, typename IDX, int MARCH>
__global__ void __launch_bounds__(kThreads, MARCH == 0 ? VR_MIN_BLOCKS - 1 : VR_MULTI_SPARSE_CTAS)
trace_multi_kernel(const TraceParams* __restrict__ table, const unsigned* __restrict__ tile_end, int n_blocks,
                   unsigned* __restrict__ counter)
{
  __shared__ __align__(16) TraceParams s_P[kThreads / 32];
  static_assert(sizeof(TraceParams) % 16 == 0, "TraceParams is copied in 16-byte words");
  {
    const float4* lut = table[0].lut; // (one transfer function per context: the same table for every block)
    const int lut_size = table[0].lut_size;
    for (int i = threadIdx.x; i < lut_size; i += kThreads) s_lut[i] = __ldg(lut + i);
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned total = tile_end[n_blocks - 1];
  unsigned my_samples = 0, brick_parity = 0;
  int cur = -1;
  unsigned cur_begin = 0, cur_end = 0;
  for (;;)
  {
    unsigned w0 = 0;
    if (lane == 0) w0 = atomicAdd(counter, kMultiChunk);
    w0 = __shfl_sync(0xffffffffu, w0, 0);
    if (w0 >= total) break;
    const unsigned w1 = min(w0 + kMultiChunk, total);
    for (unsigned w = w0; w < w1; ++w)
    {
      if (w >= cur_end || w < cur_begin)
      {
        // first block whose running tile count exceeds w (blocks without tiles have equal neighbours)
        int lo = 0, hi = n_blocks - 1;
        while (lo < hi)
        {
          const int mid = (lo + hi) >> 1;
          if (__ldg(tile_end + mid) > w) hi = mid; else lo = mid + 1;
        }
        cur = lo;
        cur_begin = lo > 0 ? __ldg(tile_end + lo - 1) : 0u;
        cur_end = __ldg(tile_end + lo);
        __syncwarp();
        const uint4* src = reinterpret_cast<const uint4*>(table + cur);
        uint4* dst = reinterpret_cast<uint4*>(&s_P[warp]);
        for (int k = lane; k < (int)(sizeof(TraceParams) / 16); k += 32) dst[k] = __ldg(src + k);
        __syncwarp();
      }
      trace_tile<KIND, FT, ASSOC, 3, IDX, MARCH>(s_P[warp], w - cur_begin, my_samples, brick_parity, nullptr);
    }
  }
  const TraceParams& P0 = table[0];
  if (P0.end_stamp && threadIdx.x == 0) atomicMax(P0.end_stamp, global_ns());
  if (P0.sample_counter)
  {
    unsigned long long s = my_samples;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && s) atomicAdd(P0.sample_counter, s);
  }
}

constexpr size_t kBrickSmem = (size_t)(kThreads / 32) * 2 * kBrickFloats * 4 + (size_t)(kThreads / 32) * 2 * 8;

// grid < 0: load the kernel this dispatch would launch, launch nothing (preload_trace_kernels)
template <int KIND, typename FT, int ASSOC, typename IDX, int MARCH>
cudaError_t launch_march(const TraceParams& p, int mode, int grid, cudaStream_t s)
{
  const size_t smem = MARCH == 2 ? kBrickSmem : 0;
  auto go = [&](auto kernel) {
    if (grid < 0) { preload_kernel(kernel); return; }
    // static (16 KiB table) + dynamic (bricks) shared memory exceed the 48 KiB a kernel gets by default;
    // the opt-in is per device, so it is (cheaply) repeated per launch
    if (MARCH == 2) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBrickSmem);
    kernel<<<grid, kThreads, smem, s>>>(p);
  };
  if (mode == 0)      go(trace_kernel<KIND, FT, ASSOC, 0, IDX, MARCH>);
  else if (mode == 1) go(trace_kernel<KIND, FT, ASSOC, 1, IDX, MARCH>);
  else if (mode == 2) go(trace_kernel<KIND, FT, ASSOC, 2, IDX, MARCH>);
  else if (mode == 5) go(trace_kernel<KIND, FT, ASSOC, 5, IDX, MARCH>);
  else                go(trace_kernel<KIND, FT, ASSOC, 3, IDX, MARCH>);
  return cudaGetLastError();
}

// which march a launch runs (the host decided what the block and the step allow: TraceParams::march)
template <int KIND, typename FT, int ASSOC, typename IDX>
cudaError_t launch_mode(const TraceParams& p, int mode, int grid, cudaStream_t s)
{
  if (mode == 4) // the demand-staging pre-pass visits cells only: general march
  {
    if (grid < 0) preload_kernel(trace_kernel<KIND, FT, ASSOC, 4, IDX, 0>);
    else trace_kernel<KIND, FT, ASSOC, 4, IDX, 0><<<grid, kThreads, 0, s>>>(p);
    return cudaGetLastError();
  }
  if (KIND == 0 && ASSOC == VR_POINT)
  {
    if (p.march == 1) return launch_march<0, FT, VR_POINT, IDX, 1>(p, mode, grid, s);
    if (p.march == 2 && sizeof(FT) == 4) return launch_march<0, float, VR_POINT, IDX, 2>(p, mode, grid, s);
  }
  return launch_march<KIND, FT, ASSOC, IDX, 0>(p, mode, grid, s);
}
template <int KIND, typename FT, int ASSOC>
cudaError_t launch_idx(const TraceParams& p, int mode, int grid, cudaStream_t s)
{
  // 32-bit element indices whenever the block has fewer than 2^31 points (saves the 64-bit
  // multiply-add chains in the gather address arithmetic)
  const long long n = (long long)p.blk.dims[0] * p.blk.dims[1] * p.blk.dims[2];
  return n < (1ll << 31) ? launch_mode<KIND, FT, ASSOC, int>(p, mode, grid, s)
                         : launch_mode<KIND, FT, ASSOC, long long>(p, mode, grid, s);
}
template <int KIND, typename FT>
cudaError_t launch_assoc(const TraceParams& p, int mode, int grid, cudaStream_t s)
{
  return p.blk.assoc == VR_POINT ? launch_idx<KIND, FT, VR_POINT>(p, mode, grid, s)
                                 : launch_idx<KIND, FT, VR_CELL>(p, mode, grid, s);
}
template <int KIND>
cudaError_t launch_dtype(const TraceParams& p, int mode, int grid, cudaStream_t s)
{
  return p.blk.dtype == VR_F32 ? launch_assoc<KIND, float>(p, mode, grid, s)
                               : launch_assoc<KIND, double>(p, mode, grid, s);
}

} // namespace

namespace
{
template <int KIND, typename FT, int ASSOC, typename IDX>
cudaError_t launch_multi_march(const TraceParams& first, const TraceParams* table, const unsigned* tile_end, int n,
                               unsigned* counter, int grid, cudaStream_t s)
{
  auto go = [&](auto kernel) {
    if (grid < 0) { preload_kernel(kernel); return; }
    kernel<<<grid, kThreads, 0, s>>>(table, tile_end, n, counter);
  };
  if (KIND == 0 && ASSOC == VR_POINT && first.march == 1) go(trace_multi_kernel<0, FT, VR_POINT, IDX, 1>);
  else go(trace_multi_kernel<KIND, FT, ASSOC, IDX, 0>);
  return cudaGetLastError();
}
template <int KIND, typename FT>
cudaError_t launch_multi_assoc(const TraceParams& first, const TraceParams* table, const unsigned* tile_end, int n,
                               unsigned* counter, int grid, cudaStream_t s)
{
  // (blocks of 2^31 points or more go one launch each: vr_trace_blocks_to_layers batches 32-bit blocks only)
  return first.blk.assoc == VR_POINT ? launch_multi_march<KIND, FT, VR_POINT, int>(first, table, tile_end, n, counter, grid, s)
                                     : launch_multi_march<KIND, FT, VR_CELL, int>(first, table, tile_end, n, counter, grid, s);
}
cudaError_t launch_multi_dispatch(const TraceParams& first, const TraceParams* table, const unsigned* tile_end, int n,
                                  unsigned* counter, int grid, cudaStream_t s)
{
  if (first.blk.kind == 0)
    return first.blk.dtype == VR_F32 ? launch_multi_assoc<0, float>(first, table, tile_end, n, counter, grid, s)
                                     : launch_multi_assoc<0, double>(first, table, tile_end, n, counter, grid, s);
  return first.blk.dtype == VR_F32 ? launch_multi_assoc<1, float>(first, table, tile_end, n, counter, grid, s)
                                   : launch_multi_assoc<1, double>(first, table, tile_end, n, counter, grid, s);
}
} // namespace

// all n blocks share `first`'s kernel variant (grid kind, scalar type, association, 32-bit indices, march 0 or 1);
// `counter` has been zeroed on `s`
cudaError_t launch_trace_multi(const TraceParams& first, const TraceParams* table, const unsigned* tile_end, int n,
                               unsigned long long total_tiles, unsigned* counter, int sm_count, cudaStream_t s)
{
  if (n <= 0 || total_tiles == 0) return cudaSuccess;
  const int full = first.march == 1 ? VR_MULTI_SPARSE_CTAS : VR_MIN_BLOCKS - 1;
  const int ctas_per_sm = first.ctas_per_sm > 0 && first.ctas_per_sm < full ? first.ctas_per_sm : full;
  long long grid = (long long)sm_count * ctas_per_sm;
  const long long need = (long long)((total_tiles + 4ull * kMultiChunk - 1) / (4ull * kMultiChunk));
  if (grid > need) grid = need;
  return launch_multi_dispatch(first, table, tile_end, n, counter, (int)grid, s);
}

void preload_trace_kernels(const BlockDev& blk)
{
  TraceParams p;
  std::memset(&p, 0, sizeof(p));
  p.blk = blk;
  for (int march = 0; march < 3; ++march)
    for (int mode = 0; mode <= 5; ++mode)
    {
      p.march = march;
      if (blk.kind == 0) launch_dtype<0>(p, mode, -1, nullptr);
      else launch_dtype<1>(p, mode, -1, nullptr);
    }
  for (int march = 0; march < 2; ++march)
  {
    p.march = march;
    launch_multi_dispatch(p, nullptr, nullptr, 0, nullptr, -1, nullptr);
  }
  cudaGetLastError();
}

cudaError_t launch_trace(const TraceParams& p, int mode_partials, int sm_count, cudaStream_t s,
                         bool zero_counter)
{
  const long long n_tiles = (long long)p.tiles_x * p.tiles_y + (mode_partials == 2 ? p.n_clear_chunks : 0);
  if (n_tiles <= 0) return cudaSuccess;
  // persistent grid: SMs x resident CTAs (4 warps each), capped by the work available
  const bool std_march = mode_partials == 4 || p.march == 0;
  const int full = std_march ? VR_MIN_BLOCKS : (p.march == 1 ? 7 : 3); // what the registers allow
  const int ctas_per_sm = p.ctas_per_sm > 0 && p.ctas_per_sm < full ? p.ctas_per_sm : full; // (VR_CTAS_PER_SM: A/B runs)
  long long grid = (long long)sm_count * ctas_per_sm;
  const long long need = (n_tiles + 3) / 4;
  if (grid > need) grid = need;
  if (zero_counter)
  {
    cudaError_t e = cudaMemsetAsync(p.tile_counter, 0, sizeof(unsigned int), s);
    if (e != cudaSuccess) return e;
  }
  return p.blk.kind == 0 ? launch_dtype<0>(p, mode_partials, (int)grid, s)
                         : launch_dtype<1>(p, mode_partials, (int)grid, s);
}

} // namespace vr
