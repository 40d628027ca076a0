// vr_umesh_faces.hpp -- which faces of an explicit cell set are EXTERNAL (belong to exactly one cell).
// Host side, publish time (vr_block_unstructured): one byte per cell, bit f = face f of vr_umesh_geom.hpp's
// numbering is external.  Faces are matched by their sorted point ids, which is also how VTK-m's ExternalFaces /
// MeshConnectivityBuilder find the mesh boundary the ConnectivityTracer enters through.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

#include "vr_umesh_geom.hpp"

namespace vr
{
// Faces are bucketed by their smallest point id (a counting sort: O(faces)); a point of a real mesh carries a
// handful of faces, so matching the faces of one bucket pairwise by their sorted ids is O(1) per face.
inline std::vector<unsigned char> umesh_external_mask(const int* conn, size_t n_cells, int shape)
{
  const int n_faces = shape == 8 ? 6 : 4, nv = shape == 8 ? 4 : 3;
  const size_t nf = n_cells * (size_t)n_faces;
  auto sorted_ids = [&](size_t face, int out[4]) {
    const size_t c = face / n_faces;
    const int f = (int)(face % n_faces);
    for (int i = 0; i < 4; ++i) out[i] = i < nv ? conn[c * shape + umesh_face_point(shape, f, i)] : -1;
    auto cs = [&](int i, int j) { if (out[j] < out[i]) std::swap(out[i], out[j]); };
    cs(0, 1); cs(2, 3); cs(0, 2); cs(1, 3); cs(1, 2); // 4-element sorting network
  };
  int max_id = 0;
  for (size_t i = 0; i < n_cells * (size_t)shape; ++i) max_id = std::max(max_id, conn[i]);
  // bucket = smallest REAL point id of the face (the -1 pad of a triangle sorts first and is skipped)
  std::vector<size_t> start((size_t)max_id + 2, 0);
  std::vector<int> low(nf);
  for (size_t c = 0, face = 0; c < n_cells; ++c)
    for (int f = 0; f < n_faces; ++f, ++face)
    {
      int lo = conn[c * shape + umesh_face_point(shape, f, 0)];
      for (int i = 1; i < nv; ++i) lo = std::min(lo, conn[c * shape + umesh_face_point(shape, f, i)]);
      low[face] = lo;
      start[(size_t)lo + 1] += 1;
    }
  for (size_t i = 1; i < start.size(); ++i) start[i] += start[i - 1];
  std::vector<size_t> order(nf), cursor(start.begin(), start.end() - 1);
  for (size_t face = 0; face < nf; ++face) order[cursor[(size_t)low[face]]++] = face;
  std::vector<unsigned char> external(nf, 0);
  {
    std::vector<int> ids;
    std::vector<unsigned char> matched;
    for (size_t b = 0; b + 1 < start.size(); ++b)
    {
      const size_t n = start[b + 1] - start[b];
      if (n == 0) continue;
      ids.resize(n * 4);
      matched.assign(n, 0);
      for (size_t i = 0; i < n; ++i) sorted_ids(order[start[b] + i], &ids[i * 4]);
      for (size_t i = 0; i < n; ++i)
        for (size_t j = i + 1; j < n; ++j)
          if (std::equal(&ids[i * 4], &ids[i * 4] + 4, &ids[j * 4])) matched[i] = matched[j] = 1;
      for (size_t i = 0; i < n; ++i)
        if (!matched[i]) external[order[start[b] + i]] = 1;
    }
  }
  std::vector<unsigned char> mask(n_cells, 0);
  for (size_t c = 0, face = 0; c < n_cells; ++c)
    for (int f = 0; f < n_faces; ++f, ++face)
      if (external[face]) mask[c] |= (unsigned char)(1u << f);
  return mask;
}

// FNV-1a over the connectivity: the boundary only depends on it, and a simulation republishes the same topology
// every cycle -- vr_block_unstructured keeps the last mask per block id and reuses it when the hash matches
inline uint64_t umesh_conn_hash(const int* conn, size_t n)
{
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i)
  {
    h ^= (uint64_t)(uint32_t)conn[i];
    h *= 1099511628211ull;
  }
  return h;
}
} // namespace vr
