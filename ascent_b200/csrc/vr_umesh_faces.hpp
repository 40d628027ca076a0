// vr_umesh_faces.hpp -- which faces of an explicit cell set are EXTERNAL (belong to exactly one cell).
// Host side, publish time (vr_block_unstructured): one byte per cell, bit f = face f of vr_umesh_geom.hpp's
// numbering is external.  Faces are matched by their sorted point ids (one sort of n_cells * faces keys), which
// is also how VTK-m's ExternalFaces / MeshConnectivityBuilder find the mesh boundary the ConnectivityTracer
// enters through.
#pragma once
#include <algorithm>
#include <cstddef>
#include <vector>

#include "vr_umesh_geom.hpp"

namespace vr
{
struct UFaceKey
{
  int p[4];      // sorted point ids (-1 pads a triangle)
  int cell_face; // cell * 8 + face
};

inline std::vector<unsigned char> umesh_external_mask(const int* conn, size_t n_cells, int shape)
{
  const int n_faces = shape == 8 ? 6 : 4, nv = shape == 8 ? 4 : 3;
  std::vector<UFaceKey> keys(n_cells * (size_t)n_faces);
  for (size_t c = 0; c < n_cells; ++c)
    for (int f = 0; f < n_faces; ++f)
    {
      UFaceKey& k = keys[c * n_faces + f];
      for (int i = 0; i < 4; ++i) k.p[i] = i < nv ? conn[c * shape + umesh_face_point(shape, f, i)] : -1;
      std::sort(k.p, k.p + 4);
      k.cell_face = (int)(c * 8 + f);
    }
  std::sort(keys.begin(), keys.end(), [](const UFaceKey& a, const UFaceKey& b) {
    for (int i = 0; i < 4; ++i)
      if (a.p[i] != b.p[i]) return a.p[i] < b.p[i];
    return a.cell_face < b.cell_face;
  });
  std::vector<unsigned char> mask(n_cells, 0);
  for (size_t i = 0; i < keys.size();)
  {
    size_t j = i + 1;
    while (j < keys.size() && std::equal(keys[i].p, keys[i].p + 4, keys[j].p)) ++j;
    if (j - i == 1) mask[(size_t)keys[i].cell_face >> 3] |= (unsigned char)(1u << (keys[i].cell_face & 7));
    i = j;
  }
  return mask;
}
} // namespace vr
