// vr_umesh_geom.hpp -- where a ray crosses the boundary of an explicit cell set (N4, csrc/unstructured.cu).
//
// VTK-m's ConnectivityTracer (behind UnstructuredWrapper::render, src/libs/vtkh/rendering/VolumeRenderer.cpp:
// 182-221) samples a ray only along the stretches it spends inside the mesh: it enters through an external face,
// walks from cell to cell, leaves through an external face, and looks for the next entry behind that.  This
// header finds those crossings for one ray without cell-to-cell connectivity: every external face near the ray is
// intersected (a quadrilateral as two triangles), a crossing ENTERS when the ray runs against the face's outward
// normal.  The arithmetic is the oracle's (oracle/raycast_oracle.c, um_boundary_crossings: that one tests every
// external face of the mesh by brute force), expression for expression, so that the crossings -- and with them
// every sample position -- come out bit-identical; what differs is how the candidate faces are found: here
// through the uniform bins of the cell locator, visited slab by slab along the ray's dominant axis.
//
// Plain functions over plain pointers, usable from device code and -- for the CPU test of the traversal
// (tests/umesh_geom_host.cpp, tests/test_umesh_crossings.py) -- from host code.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define VR_HD __host__ __device__ __forceinline__
#else
#define VR_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define VR_LD(p) __ldg(p)
#else
#define VR_LD(p) (*(p))
#endif

namespace vr
{
constexpr int kMaxCrossings = 32;

// faces in the cell's own point numbering (VTK hexahedron / tetrahedron; a triangle repeats its last point)
VR_HD int umesh_face_point(int shape, int face, int i)
{
  // (written as arithmetic-free switch tables so that they live in registers / immediates on the device)
  if (shape == 8)
  {
    switch (face * 4 + i)
    {
      case 0: return 0; case 1: return 3; case 2: return 2; case 3: return 1;
      case 4: return 4; case 5: return 5; case 6: return 6; case 7: return 7;
      case 8: return 0; case 9: return 1; case 10: return 5; case 11: return 4;
      case 12: return 1; case 13: return 2; case 14: return 6; case 15: return 5;
      case 16: return 2; case 17: return 3; case 18: return 7; case 19: return 6;
      case 20: return 3; case 21: return 0; case 22: return 4; default: return 7;
    }
  }
  switch (face * 4 + i)
  {
    case 0: return 0; case 1: return 2; case 2: return 1; case 3: return 1;
    case 4: return 0; case 5: return 1; case 6: return 3; case 7: return 3;
    case 8: return 1; case 9: return 2; case 10: return 3; case 11: return 3;
    case 12: return 2; case 13: return 0; case 14: return 3; default: return 3;
  }
}

// Moeller-Trumbore in f32, edges included (1e-6 in the barycentric coordinates): distance > 0, or +inf.
// n receives the unnormalised normal (b - a) x (c - a).
VR_HD float umesh_tri_hit(const float* o, const float* d, const float* a, const float* b, const float* c, float n[3])
{
  const float inf = INFINITY;
  const float e1[3] = { b[0] - a[0], b[1] - a[1], b[2] - a[2] }, e2[3] = { c[0] - a[0], c[1] - a[1], c[2] - a[2] };
  n[0] = e1[1] * e2[2] - e1[2] * e2[1];
  n[1] = e1[2] * e2[0] - e1[0] * e2[2];
  n[2] = e1[0] * e2[1] - e1[1] * e2[0];
  const float pv[3] = { d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0] };
  const float det = e1[0] * pv[0] + e1[1] * pv[1] + e1[2] * pv[2];
  if (fabsf(det) < 1e-12f) return inf;
  const float inv = 1.f / det;
  const float tv[3] = { o[0] - a[0], o[1] - a[1], o[2] - a[2] };
  const float u = (tv[0] * pv[0] + tv[1] * pv[1] + tv[2] * pv[2]) * inv;
  if (u < -1e-6f || u > 1.f + 1e-6f) return inf;
  const float qv[3] = { tv[1] * e1[2] - tv[2] * e1[1], tv[2] * e1[0] - tv[0] * e1[2], tv[0] * e1[1] - tv[1] * e1[0] };
  const float v = (d[0] * qv[0] + d[1] * qv[1] + d[2] * qv[2]) * inv;
  if (v < -1e-6f || u + v > 1.f + 1e-6f) return inf;
  const float t = (e2[0] * qv[0] + e2[1] * qv[1] + e2[2] * qv[2]) * inv;
  return t > 0.f ? t : inf;
}

// crossings are kept as signed distances (+t enters, -t leaves), sorted by distance with a leaving crossing
// before an entering one at the same distance; bit-identical repeats once; the nearest kMaxCrossings survive
VR_HD int umesh_insert_crossing(float* hits, int n, float key)
{
  const float t = fabsf(key);
  int at = 0;
  while (at < n)
  {
    const float h = hits[at], ht = fabsf(h);
    if (h == key) return n;
    if (ht > t || (ht == t && key < 0.f)) break;
    ++at;
  }
  if (at >= kMaxCrossings) return n;
  if (n < kMaxCrossings) ++n;
  for (int k = n - 1; k > at; --k) hits[k] = hits[k - 1];
  hits[at] = key;
  return n;
}

// the crossings of the external faces (bits of `mask`) of one cell, v = its points in cell order
template <int SHAPE>
VR_HD int umesh_cell_crossings(const float v[SHAPE][3], const int* cn, unsigned mask, const float* o, const float* d,
                               float* hits, int n)
{
  // centroid: points summed in cell order, times 1 / SHAPE (applied where it is used, like the oracle)
  float cen[3] = { 0.f, 0.f, 0.f };
  for (int k = 0; k < SHAPE; ++k)
    for (int q = 0; q < 3; ++q) cen[q] = cen[q] + v[k][q];
  const float w = SHAPE == 8 ? 0.125f : 0.25f;
  const int n_faces = SHAPE == 8 ? 6 : 4;
  for (int f = 0; f < n_faces; ++f)
  {
    if (!(mask & (1u << f))) continue;
    const int i0 = umesh_face_point(SHAPE, f, 0), i1 = umesh_face_point(SHAPE, f, 1);
    const int i2 = umesh_face_point(SHAPE, f, 2), i3 = umesh_face_point(SHAPE, f, 3);
    const float* a = v[i0];
    float nrm[3], n2[3];
    float t = umesh_tri_hit(o, d, a, v[i1], v[i2], nrm);
    if (SHAPE == 8 && VR_LD(cn + i3) != VR_LD(cn + i2))
    {
      const float t2 = umesh_tri_hit(o, d, a, v[i2], v[i3], n2);
      if (t2 < t) { t = t2; nrm[0] = n2[0]; nrm[1] = n2[1]; nrm[2] = n2[2]; }
    }
    if (t == INFINITY) continue;
    const float side = nrm[0] * (a[0] - cen[0] * w) + nrm[1] * (a[1] - cen[1] * w) + nrm[2] * (a[2] - cen[2] * w);
    float dn = d[0] * nrm[0] + d[1] * nrm[1] + d[2] * nrm[2];
    if (side < 0.f) dn = -dn;
    n = umesh_insert_crossing(hits, n, dn < 0.f ? t : -t);
  }
  return n;
}

// what the traversal needs of the mesh (UMeshDev of vr_internal.h has exactly these members)
template <class M>
VR_HD int umesh_bin_of(const M& U, int a, float x)
{
  int b = (int)((x - U.bmin[a]) * U.ginv[a]);
  b = b < 0 ? 0 : b;
  b = b > U.g[a] - 1 ? U.g[a] - 1 : b;
  return b;
}

// Every crossing of the ray o + t d, t in [t0, t1] (its span inside the mesh's point bounds), with the mesh
// boundary.  The bins are visited slab by slab along the ray's dominant axis A: within slab k the ray covers a
// small rectangle of bins in the other two axes (everything padded by 1/1000 of a bin, so that a face touching a
// bin the ray only grazes is still seen); every cell listed in those bins that owns an external face is tested.
// A cell listed in several visited bins produces the same keys again, which the insertion drops.
template <int SHAPE, class M>
VR_HD int umesh_collect_crossings(const M& U, const float o[3], const float d[3], float t0, float t1, float* hits)
{
  int n = 0;
  int A = 0;
  float ad = fabsf(d[0]);
  if (fabsf(d[1]) > ad) { A = 1; ad = fabsf(d[1]); }
  if (fabsf(d[2]) > ad) { A = 2; ad = fabsf(d[2]); }
  const int B = (A + 1) % 3, C = (A + 2) % 3;
  const float span = t1 - t0;
  const float tpad = span * 1e-3f + 1e-6f;
  const float ta_all = t0 - tpad, tb_all = t1 + tpad;
  const float hA = U.ginv[A] > 0.f ? 1.f / U.ginv[A] : 0.f;
  const float hB = U.ginv[B] > 0.f ? 1.f / U.ginv[B] : 0.f;
  const float hC = U.ginv[C] > 0.f ? 1.f / U.ginv[C] : 0.f;
  const float padA = hA * 1e-3f, padB = hB * 1e-3f, padC = hC * 1e-3f;
  const float pa0 = o[A] + ta_all * d[A], pa1 = o[A] + tb_all * d[A];
  const int k0 = umesh_bin_of(U, A, fminf(pa0, pa1) - padA), k1 = umesh_bin_of(U, A, fmaxf(pa0, pa1) + padA);
  const float inv_dA = 1.f / d[A]; // |d[A]| >= 1/sqrt(3): the dominant component of a unit vector
  for (int k = k0; k <= k1; ++k)
  {
    float ta = ta_all, tb = tb_all;
    if (hA > 0.f)
    {
      const float lo = U.bmin[A] + (float)k * hA - padA, hi = U.bmin[A] + (float)(k + 1) * hA + padA;
      const float s0 = (lo - o[A]) * inv_dA, s1 = (hi - o[A]) * inv_dA;
      ta = fmaxf(ta_all, fminf(s0, s1) - tpad);
      tb = fminf(tb_all, fmaxf(s0, s1) + tpad);
      // (the first and last slab also take what lies before / behind them: the bins are clamped at the bounds)
      if (k == k0) { if (d[A] > 0.f) ta = ta_all; else tb = tb_all; }
      if (k == k1) { if (d[A] > 0.f) tb = tb_all; else ta = ta_all; }
      if (ta > tb) continue;
    }
    const float b0 = o[B] + ta * d[B], b1 = o[B] + tb * d[B];
    const float c0 = o[C] + ta * d[C], c1 = o[C] + tb * d[C];
    const int jb0 = umesh_bin_of(U, B, fminf(b0, b1) - padB), jb1 = umesh_bin_of(U, B, fmaxf(b0, b1) + padB);
    const int jc0 = umesh_bin_of(U, C, fminf(c0, c1) - padC), jc1 = umesh_bin_of(U, C, fmaxf(c0, c1) + padC);
    for (int jc = jc0; jc <= jc1; ++jc)
      for (int jb = jb0; jb <= jb1; ++jb)
      {
        int ix[3];
        ix[A] = k; ix[B] = jb; ix[C] = jc;
        const size_t bin = ((size_t)ix[2] * U.g[1] + ix[1]) * U.g[0] + ix[0];
        if (!VR_LD(U.bin_ext + bin)) continue; // no cell of this bin touches the mesh boundary (most bins)
        const int q0 = VR_LD(U.bin_start + bin), q1 = VR_LD(U.bin_start + bin + 1);
        for (int q = q0; q < q1; ++q)
        {
          const int c = VR_LD(U.bin_cells + q);
          const unsigned mask = VR_LD(U.ext_mask + c);
          if (!mask) continue;
          const int* cn = U.conn + (size_t)c * SHAPE;
          float v[SHAPE][3];
          for (int p = 0; p < SHAPE; ++p)
          {
            const float* x = U.xyz + 3 * (size_t)VR_LD(cn + p);
            v[p][0] = VR_LD(x); v[p][1] = VR_LD(x + 1); v[p][2] = VR_LD(x + 2);
          }
          n = umesh_cell_crossings<SHAPE>(v, cn, mask, o, d, hits, n);
        }
      }
  }
  return n;
}
} // namespace vr
