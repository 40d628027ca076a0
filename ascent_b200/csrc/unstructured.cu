// unstructured.cu -- the unstructured producer of path B.
//
// vtkh::VolumeRenderer renders a domain whose cell set is not structured through UnstructuredWrapper::render
// (src/libs/vtkh/rendering/VolumeRenderer.cpp:182-221: VTK-m's ConnectivityProxy::PartialTrace, then
// vtkm_to_partials :141-180) and forces path B for the whole scene (:874-903, m_has_unstructured).  VTK-m's
// ConnectivityTracer is not part of /root/reference, so what this kernel implements is the algorithm restated in
// oracle/raycast_oracle.c ("N4: unstructured cells"), whose conventions are the ones the reference's golden of this
// path decides (tout_multi_topo_single_ghost_vol_render100.png, a box of hexahedra with a notch: 99.9 % of its
// pixels come out uint8-equal) and which degenerates to the structured sampler on structured meshes:
//   K1-K3 exactly as sampler.cu (rays over the screen subset of the mesh's point bounds, canvas-depth clamp,
//   entry/exit of the bounds).  The ray is then cut into the stretches it spends INSIDE the mesh -- from an
//   entering crossing of the mesh boundary (the external faces, vr_umesh_geom.hpp) to the next leaving one; a ray
//   that leaves through a concavity and comes back starts a new stretch -- and every stretch is sampled from
//   entry + (entry mod sample distance), the phase reset at every entry, in steps of the sample distance while the
//   distance is <= the stretch's end.  A sample contributes when it lies inside a cell (hexahedron: inverse
//   trilinear map by four Newton steps from the cell centre; tetrahedron: barycentric coordinates; lowest cell id
//   on shared faces); the table index is v * 1024 clamped to 1023 (the structured sampler: v * 1023); blend and
//   early termination as in the structured sampler; one partial per ray with alpha >= 0.001, depth = the exit
//   distance of the bounds.
// Cells are found through uniform bins over the point bounds (ceil(cbrt(n_cells)) per axis), built on the device
// at publish time: bounds by atomic min/max, per-bin counts, a scan, a fill.  Compiled --fmad=false like the
// sampler: every comparison and every Newton step rounds like the oracle, so partials are bit-identical.
#include <cmath>
#include <cstring>

#include "vr_internal.h"
#include "vr_umesh_geom.hpp"

namespace vr
{
namespace
{
constexpr int kThreads = 128;
constexpr int kTileW = 8, kTileH = 4;
constexpr float kTol = 1e-4f;

__shared__ float4 s_ulut[1024];

template <typename FT>
__device__ __forceinline__ float ufld(const void* f, int i)
{
  return (float)__ldg(reinterpret_cast<const FT*>(f) + i);
}
__device__ __forceinline__ float urcp_safe(float f) { return 1.0f / ((fabsf(f) < 1e-8f) ? 1e-8f : f); }

__device__ __forceinline__ int bin_of(const UMeshDev& U, int a, float x)
{
  int b = (int)((x - U.bmin[a]) * U.ginv[a]);
  b = max(b, 0);
  b = min(b, U.g[a] - 1);
  return b;
}

// 3x3 solve by Cramer's rule: columns a, b, c; right-hand side r
__device__ __forceinline__ bool solve3(const float a[3], const float b[3], const float c[3], const float r[3], float out[3])
{
  const float c0 = b[1] * c[2] - b[2] * c[1], c1 = b[2] * c[0] - b[0] * c[2], c2 = b[0] * c[1] - b[1] * c[0];
  const float det = a[0] * c0 + a[1] * c1 + a[2] * c2;
  if (det == 0.f) return false;
  const float inv = 1.f / det;
  out[0] = (r[0] * c0 + r[1] * c1 + r[2] * c2) * inv;
  const float d0 = r[1] * c[2] - r[2] * c[1], d1 = r[2] * c[0] - r[0] * c[2], d2 = r[0] * c[1] - r[1] * c[0];
  out[1] = (a[0] * d0 + a[1] * d1 + a[2] * d2) * inv;
  const float e0 = b[1] * r[2] - b[2] * r[1], e1 = b[2] * r[0] - b[0] * r[2], e2 = b[0] * r[1] - b[1] * r[0];
  out[2] = (a[0] * e0 + a[1] * e1 + a[2] * e2) * inv;
  return true;
}

// parametric coordinates of p in cell c; true when inside (tolerance kTol).  Also the padded-bounds pre-test.
template <int SHAPE>
__device__ __forceinline__ bool pcoords(const UMeshDev& U, int c, const float p[3], float rst[3])
{
  const int* cn = U.conn + (size_t)c * SHAPE;
  float v[SHAPE][3];
  float lo[3] = { __int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000) };
  float hi[3] = { __int_as_float(0xff800000), __int_as_float(0xff800000), __int_as_float(0xff800000) };
#pragma unroll
  for (int k = 0; k < SHAPE; ++k)
  {
    const float* q = U.xyz + 3 * (size_t)__ldg(cn + k);
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
      v[k][a] = __ldg(q + a);
      lo[a] = fminf(lo[a], v[k][a]);
      hi[a] = fmaxf(hi[a], v[k][a]);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    const float pad = (hi[a] - lo[a]) * kTol;
    if (p[a] < lo[a] - pad || p[a] > hi[a] + pad) return false;
  }
  if (SHAPE == 4)
  {
    float e1[3], e2[3], e3[3], r[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
      e1[a] = v[1][a] - v[0][a];
      e2[a] = v[2][a] - v[0][a];
      e3[a] = v[3][a] - v[0][a];
      r[a] = p[a] - v[0][a];
    }
    if (!solve3(e1, e2, e3, r, rst)) return false;
    return rst[0] >= -kTol && rst[1] >= -kTol && rst[2] >= -kTol && rst[0] + rst[1] + rst[2] <= 1.f + kTol;
  }
  float r = 0.5f, s = 0.5f, t = 0.5f;
#pragma unroll 1
  for (int it = 0; it < 4; ++it)
  {
    float F[3], Jr[3], Js[3], Jt[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
      const float x01 = v[0][a] + r * (v[1 % SHAPE][a] - v[0][a]), x32 = v[3][a] + r * (v[2][a] - v[3][a]);
      const float x45 = v[4 % SHAPE][a] + r * (v[5 % SHAPE][a] - v[4 % SHAPE][a]);
      const float x76 = v[7 % SHAPE][a] + r * (v[6 % SHAPE][a] - v[7 % SHAPE][a]);
      const float xb = x01 + s * (x32 - x01), xt = x45 + s * (x76 - x45);
      F[a] = (xb + t * (xt - xb)) - p[a];
      const float d01 = v[1 % SHAPE][a] - v[0][a], d32 = v[2][a] - v[3][a];
      const float d45 = v[5 % SHAPE][a] - v[4 % SHAPE][a], d76 = v[6 % SHAPE][a] - v[7 % SHAPE][a];
      const float db = d01 + s * (d32 - d01), dt = d45 + s * (d76 - d45);
      Jr[a] = db + t * (dt - db);
      Js[a] = (x32 - x01) + t * ((x76 - x45) - (x32 - x01));
      Jt[a] = xt - xb;
    }
    float d[3];
    if (!solve3(Jr, Js, Jt, F, d)) return false;
    r = r - d[0]; s = s - d[1]; t = t - d[2];
  }
  rst[0] = r; rst[1] = s; rst[2] = t;
  return r >= -kTol && r <= 1.f + kTol && s >= -kTol && s <= 1.f + kTol && t >= -kTol && t <= 1.f + kTol;
}

// the lowest-numbered cell that contains p (bin lists are unordered: every candidate is tested), or -1
template <int SHAPE>
__device__ __forceinline__ int locate(const UMeshDev& U, const float p[3], float rst[3])
{
  const int bx = bin_of(U, 0, p[0]), by = bin_of(U, 1, p[1]), bz = bin_of(U, 2, p[2]);
  const size_t b = ((size_t)bz * U.g[1] + by) * U.g[0] + bx;
  const int k0 = __ldg(U.bin_start + b), k1 = __ldg(U.bin_start + b + 1);
  int found = -1;
  for (int k = k0; k < k1; ++k)
  {
    const int c = __ldg(U.bin_cells + k);
    if (found >= 0 && c > found) continue;
    float q[3];
    if (pcoords<SHAPE>(U, c, p, q)) { found = c; rst[0] = q[0]; rst[1] = q[1]; rst[2] = q[2]; }
  }
  return found;
}

template <typename FT, int SHAPE, int ASSOC>
__global__ void __launch_bounds__(kThreads) utrace_kernel(const __grid_constant__ TraceParams P, const __grid_constant__ UMeshDev U)
{
  for (int i = threadIdx.x; i < P.lut_size; i += kThreads) s_ulut[i] = __ldg(P.lut + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int lx = lane & (kTileW - 1), ly = lane >> 3;
  const unsigned n_tiles = (unsigned)(P.tiles_x * P.tiles_y);
  const float cms_f = P.cms_f;
  const float sd = P.sample_dist;
  const float minx = P.bmin[0], miny = P.bmin[1], minz = P.bmin[2];
  const float maxx = P.bmax[0], maxy = P.bmax[1], maxz = P.bmax[2];
  for (;;)
  {
    unsigned tile = 0;
    if (lane == 0) tile = atomicAdd(P.tile_counter, 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= n_tiles) break;
    const int tj = (int)(tile / (unsigned)P.tiles_x), ti = (int)(tile % (unsigned)P.tiles_x);
    const int i = P.sx + ti * kTileW + lx;
    const int j = P.sy + tj * kTileH + ly;
    const bool in_subset = (i < P.sx + P.sw) && (j < P.sy + P.sh);
    const long long pixel = (long long)j * P.W + i;
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
    float max_distance = __int_as_float(0x7f800000);
    if (in_subset)
    {
      // ---------------- K1 (sampler.cu)
      float dx, dy, dz;
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
      float min_distance = 0.f, exit_distance = 0.f;
      // ---------------- K2
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
      // ---------------- K3
      {
        const float ix = urcp_safe(dx), iy = urcp_safe(dy), iz = urcp_safe(dz);
        const float odx = ox * ix, ody = oy * iy, odz = oz * iz;
        const float xmin = minx * ix - odx, ymin = miny * iy - ody, zmin = minz * iz - odz;
        const float xmax = maxx * ix - odx, ymax = maxy * iy - ody, zmax = maxz * iz - odz;
        min_distance = fmaxf(fmaxf(fmaxf(fminf(ymin, ymax), fminf(xmin, xmax)), fminf(zmin, zmax)), min_distance);
        exit_distance = fminf(fminf(fmaxf(ymin, ymax), fmaxf(xmin, xmax)), fmaxf(zmin, zmax));
        max_distance = fminf(max_distance, exit_distance);
        if (max_distance < min_distance) min_distance = -1.f;
      }
      if (min_distance != -1.f)
      {
        // the stretches inside the mesh, each sampled from entry + (entry mod sample distance) -- see the file header
        const float o3[3] = { ox, oy, oz }, d3[3] = { dx, dy, dz };
        float hits[kMaxCrossings];
        const int nh = umesh_collect_crossings<SHAPE>(U, o3, d3, min_distance, exit_distance, hits);
        bool inside = false, opaque = false;
        float te = 0.f;
        for (int h = 0; h < nh && !opaque; ++h)
        {
          const float key = hits[h];
          if (key > 0.f)
          {
            if (!inside) { inside = true; te = key; }
            continue;
          }
          if (!inside) continue;
          inside = false;
          const float tx = -key;
          float distance = te + fmodf(te, sd);
          while (distance <= tx && distance < max_distance && !opaque)
          {
            const float p[3] = { ox + distance * dx, oy + distance * dy, oz + distance * dz };
            float rst[3];
            // (tried: testing the previous sample's cell first and taking it when the sample lies well inside -- for
            // warped hexahedra the inverse trilinear map is not unique, a lower-numbered neighbour can also claim the
            // point, so the result differed from the full search (2 of the parity tests failed); it also measured
            // 7x slower, 30 vs 4.4 ms: profiles/r2_v30_unstructured_time_*.json)
            const int c = locate<SHAPE>(U, p, rst);
            if (c >= 0)
            {
              float v;
              if (ASSOC == VR_CELL) v = ufld<FT>(U.field, c);
              else
              {
                const int* cn = U.conn + (size_t)c * SHAPE;
                if (SHAPE == 4)
                {
                  const float f0 = ufld<FT>(U.field, __ldg(cn));
                  v = f0 + rst[0] * (ufld<FT>(U.field, __ldg(cn + 1)) - f0) + rst[1] * (ufld<FT>(U.field, __ldg(cn + 2)) - f0) +
                      rst[2] * (ufld<FT>(U.field, __ldg(cn + 3)) - f0);
                }
                else
                {
                  const float s0 = ufld<FT>(U.field, __ldg(cn)), s1 = ufld<FT>(U.field, __ldg(cn + 1 % SHAPE));
                  const float s2 = ufld<FT>(U.field, __ldg(cn + 2)), s3 = ufld<FT>(U.field, __ldg(cn + 3));
                  const float s4 = ufld<FT>(U.field, __ldg(cn + 4 % SHAPE)), s5 = ufld<FT>(U.field, __ldg(cn + 5 % SHAPE));
                  const float s6 = ufld<FT>(U.field, __ldg(cn + 6 % SHAPE)), s7 = ufld<FT>(U.field, __ldg(cn + 7 % SHAPE));
                  const float l76 = s7 + rst[0] * (s6 - s7);
                  const float l45 = s4 + rst[0] * (s5 - s4);
                  const float ltop = l45 + rst[1] * (l76 - l45);
                  const float l01 = s0 + rst[0] * (s1 - s0);
                  const float l32 = s3 + rst[0] * (s2 - s3);
                  const float lbot = l01 + rst[1] * (l32 - l01);
                  v = lbot + rst[2] * (ltop - lbot);
                }
              }
              v = (v - P.range_min) * P.inv_delta_scalar;
              const float raw = v * (cms_f + 1.f);
              float fidx = fminf(fmaxf(raw, 0.f), cms_f);
              if (raw >= 9.2233720e18f) fidx = 0.f;
              const float4 sc = s_ulut[(int)fidx];
              const float alpha = sc.w * (1.f - c3);
              c0 = c0 + sc.x * alpha;
              c1 = c1 + sc.y * alpha;
              c2 = c2 + sc.z * alpha;
              c3 = alpha + c3;
              if (c3 >= 1.f) opaque = true;
            }
            distance += sd;
          }
        }
        c0 = fminf(c0, 1.f); c1 = fminf(c1, 1.f); c2 = fminf(c2, 1.f); c3 = fminf(c3, 1.f);
      }
    }
    // ---------------- vtkm_to_partials + the alpha >= 0.001 filter of path B (VolumeRenderer.cpp:141-180, :270-283):
    // one atomic per warp, slots by ballot rank
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

// ---- locator build
__device__ __forceinline__ int order_key(float f)
{
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__global__ void ubounds_kernel(const float* __restrict__ xyz, size_t n_points, int* __restrict__ keys /* min x3, max x3 */)
{
  int lo[3] = { 0x7fffffff, 0x7fffffff, 0x7fffffff }, hi[3] = { (int)0x80000000, (int)0x80000000, (int)0x80000000 };
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_points; i += (size_t)gridDim.x * blockDim.x)
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
      const int k = order_key(xyz[3 * i + a]);
      lo[a] = min(lo[a], k);
      hi[a] = max(hi[a], k);
    }
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    for (int o = 16; o > 0; o >>= 1)
    {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0)
    {
      atomicMin(keys + a, lo[a]);
      atomicMax(keys + 3 + a, hi[a]);
    }
  }
}

// pass 0: count the cells per bin; pass 1: list them (cursor = running copy of bin_start)
template <int SHAPE, bool FILL>
__global__ void ubins_kernel(const __grid_constant__ UMeshDev U, int* __restrict__ count_or_cursor, int* __restrict__ bin_cells)
{
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < U.n_cells; c += gridDim.x * blockDim.x)
  {
    float lo[3] = { __int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000) };
    float hi[3] = { __int_as_float(0xff800000), __int_as_float(0xff800000), __int_as_float(0xff800000) };
    for (int k = 0; k < SHAPE; ++k)
    {
      const float* q = U.xyz + 3 * (size_t)U.conn[(size_t)c * SHAPE + k];
      for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], q[a]); hi[a] = fmaxf(hi[a], q[a]); }
    }
    int b0[3], b1[3];
    for (int a = 0; a < 3; ++a) { b0[a] = bin_of(U, a, lo[a]); b1[a] = bin_of(U, a, hi[a]); }
    for (int z = b0[2]; z <= b1[2]; ++z)
      for (int y = b0[1]; y <= b1[1]; ++y)
        for (int x = b0[0]; x <= b1[0]; ++x)
        {
          const size_t b = ((size_t)z * U.g[1] + y) * U.g[0] + x;
          const int at = atomicAdd(count_or_cursor + b, 1);
          if (FILL) bin_cells[at] = c;
        }
  }
}

// which bins list a cell that owns an external face (the boundary crossings only look at those)
__global__ void ubinflag_kernel(const int* __restrict__ bin_start, const int* __restrict__ bin_cells,
                                const unsigned char* __restrict__ ext_mask, unsigned char* __restrict__ bin_ext, size_t n_bins)
{
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_bins; b += (size_t)gridDim.x * blockDim.x)
  {
    unsigned char any = 0;
    for (int k = bin_start[b]; k < bin_start[b + 1] && !any; ++k) any = ext_mask[bin_cells[k]] ? 1 : 0;
    bin_ext[b] = any;
  }
}

// exclusive scan of n counts into n + 1 starts, one CTA (publish-time work: n = bins ~ cells)
__global__ void __launch_bounds__(1024) uscan_kernel(const int* __restrict__ counts, int* __restrict__ starts, size_t n)
{
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (size_t base = 0; base < n; base += 1024)
  {
    const size_t i = base + threadIdx.x;
    const int v = i < n ? counts[i] : 0;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1)
    {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    int off = s_carry;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) off += s_warp[w];
    if (i < n) starts[i] = off + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = off + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) starts[n] = s_carry;
}

template <typename FT, int SHAPE>
cudaError_t launch_assoc(const TraceParams& p, const UMeshDev& u, int grid, cudaStream_t s)
{
  auto go = [&](auto kernel) {
    if (grid < 0) { preload_kernel(kernel); return; }
    kernel<<<grid, kThreads, 0, s>>>(p, u);
  };
  if (u.assoc == VR_POINT) go(utrace_kernel<FT, SHAPE, VR_POINT>);
  else go(utrace_kernel<FT, SHAPE, VR_CELL>);
  return cudaGetLastError();
}
template <typename FT>
cudaError_t launch_shape(const TraceParams& p, const UMeshDev& u, int grid, cudaStream_t s)
{
  return u.shape == 8 ? launch_assoc<FT, 8>(p, u, grid, s) : launch_assoc<FT, 4>(p, u, grid, s);
}
} // namespace

cudaError_t launch_utrace_partials(const TraceParams& p, const UMeshDev& u, int sm_count, cudaStream_t s)
{
  const long long n_tiles = (long long)p.tiles_x * p.tiles_y;
  if (n_tiles <= 0) return cudaSuccess;
  long long grid = (long long)sm_count * 8;
  const long long need = (n_tiles + 3) / 4;
  if (grid > need) grid = need;
  cudaError_t e = cudaMemsetAsync(p.tile_counter, 0, sizeof(unsigned int), s);
  if (e != cudaSuccess) return e;
  return u.dtype == VR_F32 ? launch_shape<float>(p, u, (int)grid, s) : launch_shape<double>(p, u, (int)grid, s);
}

void preload_unstructured_kernels()
{
  TraceParams p;
  std::memset(&p, 0, sizeof(p));
  UMeshDev u;
  std::memset(&u, 0, sizeof(u));
  for (int dtype = 0; dtype < 2; ++dtype)
    for (int shape = 4; shape <= 8; shape += 4)
      for (int assoc = 0; assoc < 2; ++assoc)
      {
        u.dtype = dtype; u.shape = shape; u.assoc = assoc;
        if (dtype == VR_F32) launch_shape<float>(p, u, -1, nullptr);
        else launch_shape<double>(p, u, -1, nullptr);
      }
  preload_kernel(ubounds_kernel);
  preload_kernel(ubins_kernel<8, false>);
  preload_kernel(ubins_kernel<8, true>);
  preload_kernel(ubins_kernel<4, false>);
  preload_kernel(ubins_kernel<4, true>);
  preload_kernel(uscan_kernel);
  preload_kernel(ubinflag_kernel);
  cudaGetLastError();
}

// point bounds of the mesh (device reduction; the six floats come back to the host, which needs them for the
// camera subset anyway).  keys: 6 ints of scratch on the device.
cudaError_t umesh_bounds(const float* xyz, size_t n_points, int* keys_dev, float bmin[3], float bmax[3], int sm_count,
                         cudaStream_t s)
{
  const int init[6] = { 0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000 };
  cudaError_t e = cudaMemcpyAsync(keys_dev, init, sizeof(init), cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return e;
  size_t grid = (n_points + 255) / 256;
  if (grid > (size_t)sm_count * 8) grid = (size_t)sm_count * 8;
  if (grid < 1) grid = 1;
  ubounds_kernel<<<(unsigned)grid, 256, 0, s>>>(xyz, n_points, keys_dev);
  int keys[6];
  e = cudaMemcpyAsync(keys, keys_dev, sizeof(keys), cudaMemcpyDeviceToHost, s);
  if (e != cudaSuccess) return e;
  e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return e;
  for (int a = 0; a < 3; ++a)
  {
    int k = keys[a];
    k = k >= 0 ? k : k ^ 0x7fffffff;
    std::memcpy(&bmin[a], &k, 4);
    k = keys[3 + a];
    k = k >= 0 ? k : k ^ 0x7fffffff;
    std::memcpy(&bmax[a], &k, 4);
  }
  return cudaGetLastError();
}

// bins: u.g / u.ginv / u.bmin / u.bmax / u.ext_mask set by the caller; allocates and fills u.bin_start / u.bin_cells /
// u.bin_ext
cudaError_t umesh_build_bins(UMeshDev& u, int** bin_start_out, int** bin_cells_out, unsigned char** bin_ext_out, int sm_count,
                             cudaStream_t s)
{
  const size_t nb = (size_t)u.g[0] * u.g[1] * u.g[2];
  int *counts = nullptr, *starts = nullptr, *cells = nullptr;
  cudaError_t e = cudaMalloc(&counts, (nb + 1) * sizeof(int));
  if (e != cudaSuccess) return e;
  e = cudaMalloc(&starts, (nb + 1) * sizeof(int));
  if (e != cudaSuccess) { cudaFree(counts); return e; }
  cudaMemsetAsync(counts, 0, (nb + 1) * sizeof(int), s);
  int grid = (u.n_cells + 255) / 256;
  if (grid > sm_count * 8) grid = sm_count * 8;
  if (grid < 1) grid = 1;
  if (u.shape == 8) ubins_kernel<8, false><<<grid, 256, 0, s>>>(u, counts, nullptr);
  else ubins_kernel<4, false><<<grid, 256, 0, s>>>(u, counts, nullptr);
  uscan_kernel<<<1, 1024, 0, s>>>(counts, starts, nb);
  int total = 0;
  e = cudaMemcpyAsync(&total, starts + nb, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaMalloc(&cells, (size_t)(total > 0 ? total : 1) * sizeof(int));
  if (e != cudaSuccess) { cudaFree(counts); cudaFree(starts); return e; }
  // the counts buffer becomes the fill cursor
  cudaMemcpyAsync(counts, starts, (nb + 1) * sizeof(int), cudaMemcpyDeviceToDevice, s);
  if (u.shape == 8) ubins_kernel<8, true><<<grid, 256, 0, s>>>(u, counts, cells);
  else ubins_kernel<4, true><<<grid, 256, 0, s>>>(u, counts, cells);
  unsigned char* flags = nullptr;
  e = cudaMalloc(&flags, nb);
  if (e == cudaSuccess)
  {
    size_t fgrid = (nb + 255) / 256;
    if (fgrid > (size_t)sm_count * 8) fgrid = (size_t)sm_count * 8;
    ubinflag_kernel<<<(unsigned)fgrid, 256, 0, s>>>(starts, cells, u.ext_mask, flags, nb);
    e = cudaStreamSynchronize(s);
  }
  cudaFree(counts);
  if (e != cudaSuccess) { cudaFree(starts); cudaFree(cells); if (flags) cudaFree(flags); return e; }
  u.bin_start = starts;
  u.bin_cells = cells;
  u.bin_ext = flags;
  *bin_start_out = starts;
  *bin_cells_out = cells;
  *bin_ext_out = flags;
  return cudaGetLastError();
}
} // namespace vr
