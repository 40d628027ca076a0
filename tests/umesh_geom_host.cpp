// CPU harness for csrc/vr_umesh_geom.hpp + vr_umesh_faces.hpp (tests/test_umesh_crossings.py): the very functions
// the unstructured kernel runs per ray -- external-face mask, bin traversal, ray/face crossings -- compiled for
// the host, over bins built the way csrc/unstructured.cu builds them on the device, so that they can be checked
// against the oracle's brute force without a GPU.
#include <cmath>
#include <cstring>
#include <vector>

#include "../ascent_b200/csrc/vr_umesh_faces.hpp"

namespace
{
struct HostMesh
{
  const float* xyz;
  const int* conn;
  const unsigned char* ext_mask;
  float bmin[3], bmax[3], ginv[3];
  int g[3];
  const int* bin_start;
  const int* bin_cells;
  const unsigned char* bin_ext;
};

template <int SHAPE>
void build_bins(HostMesh& U, int n_cells, std::vector<int>& start, std::vector<int>& cells)
{
  const size_t nb = (size_t)U.g[0] * U.g[1] * U.g[2];
  start.assign(nb + 1, 0);
  for (int pass = 0; pass < 2; ++pass)
  {
    std::vector<int> cursor;
    if (pass == 1)
    {
      int run = 0;
      for (size_t b = 0; b <= nb; ++b) { const int c = start[b]; start[b] = run; run += c; }
      cells.assign((size_t)std::max(start[nb], 1), 0);
      cursor.assign(start.begin(), start.end() - 1);
    }
    for (int c = 0; c < n_cells; ++c)
    {
      float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
      for (int k = 0; k < SHAPE; ++k)
        for (int a = 0; a < 3; ++a)
        {
          const float x = U.xyz[3 * (size_t)U.conn[(size_t)c * SHAPE + k] + a];
          lo[a] = fminf(lo[a], x);
          hi[a] = fmaxf(hi[a], x);
        }
      int b0[3], b1[3];
      for (int a = 0; a < 3; ++a) { b0[a] = vr::umesh_bin_of(U, a, lo[a]); b1[a] = vr::umesh_bin_of(U, a, hi[a]); }
      for (int z = b0[2]; z <= b1[2]; ++z)
        for (int y = b0[1]; y <= b1[1]; ++y)
          for (int x = b0[0]; x <= b1[0]; ++x)
          {
            const size_t b = ((size_t)z * U.g[1] + y) * U.g[0] + x;
            if (pass == 0) start[b] += 1;
            else cells[(size_t)cursor[b]++] = c;
          }
    }
  }
}
} // namespace

// rays: n_rays x 8 floats (origin, direction, t0, t1); hits_out: n_rays x 32 floats; counts_out: n_rays ints.
// mask_out (n_cells bytes, may be null) receives the external-face mask.  bins_per_axis <= 0: ceil(cbrt(n_cells)).
extern "C" int umesh_crossings_host(const float* xyz, int n_points, const int* conn, int n_cells, int shape,
                                    int bins_per_axis, const float* rays, int n_rays, float* hits_out, int* counts_out,
                                    unsigned char* mask_out)
{
  if (shape != 8 && shape != 4) return -1;
  std::vector<unsigned char> mask = vr::umesh_external_mask(conn, (size_t)n_cells, shape);
  if (mask_out) std::memcpy(mask_out, mask.data(), mask.size());
  HostMesh U;
  U.xyz = xyz; U.conn = conn; U.ext_mask = mask.data();
  for (int a = 0; a < 3; ++a) { U.bmin[a] = INFINITY; U.bmax[a] = -INFINITY; }
  for (int i = 0; i < n_points; ++i)
    for (int a = 0; a < 3; ++a)
    {
      U.bmin[a] = fminf(U.bmin[a], xyz[3 * (size_t)i + a]);
      U.bmax[a] = fmaxf(U.bmax[a], xyz[3 * (size_t)i + a]);
    }
  int g = bins_per_axis > 0 ? bins_per_axis : (int)std::ceil(std::cbrt((double)n_cells));
  g = std::max(1, std::min(g, 256));
  for (int a = 0; a < 3; ++a)
  {
    U.g[a] = g;
    const float ext = U.bmax[a] - U.bmin[a];
    U.ginv[a] = ext > 0.f ? (float)g / ext : 0.f;
  }
  std::vector<int> start, cells;
  if (shape == 8) build_bins<8>(U, n_cells, start, cells);
  else build_bins<4>(U, n_cells, start, cells);
  U.bin_start = start.data();
  U.bin_cells = cells.data();
  // (csrc/unstructured.cu, ubinflag_kernel) which bins list a cell with an external face
  std::vector<unsigned char> flags(start.size() - 1, 0);
  for (size_t b = 0; b + 1 < start.size(); ++b)
    for (int k = start[b]; k < start[b + 1]; ++k)
      if (mask[(size_t)cells[(size_t)k]]) flags[b] = 1;
  U.bin_ext = flags.data();
  for (int r = 0; r < n_rays; ++r)
  {
    const float* q = rays + 8 * (size_t)r;
    float* h = hits_out + (size_t)vr::kMaxCrossings * r;
    counts_out[r] = shape == 8 ? vr::umesh_collect_crossings<8>(U, q, q + 3, q[6], q[7], h)
                               : vr::umesh_collect_crossings<4>(U, q, q + 3, q[6], q[7], h);
  }
  return 0;
}
