// vr_radixk.hpp -- the radix-k schedule of the reference's surface compositor, in closed form.
//
// vtkh::RadixKCompositor::CompositeImpl (src/libs/vtkh/compositing/RadixKCompositor.cpp:138-180) decomposes the
// frame [1..W]x[1..H] into one block per rank with DIY's RegularDecomposer (2 dimensions), reduces with
// RegularSwapPartners(k = 8, distance halving) and the callback reduce_images (:35-124: split the current
// sub-image into k balanced pieces along the round's dimension, keep one, send the others, z-composite what
// arrives in link order), and finally pastes every block's piece into rank 0's frame in gid order (CollectImages,
// vtkh_diy_collect.hpp).  ImageCompositor::ZBufferComposite (ImageCompositor.hpp:49-76) lets an incoming fragment
// replace the held one when its depth is <=, so fragments of different ranks at EQUAL depth resolve by the order
// in which the tree visits them -- which differs from piece to piece.
//
// On NVSwitch the tree buys nothing (every rank reaches every peer at full bandwidth; one round of k = N is
// the optimal radix), so the kernel performs ONE round (comm.cu, fold_p2p_kernel<.., ZBUF>) -- but it visits the
// ranks' fragments in the order the reference's multi-round tree would, per piece, so the selected fragment is the
// reference's for every input (tests/test_oracle_radixk.py pins this closed form against the reference's own
// reduce_images + DIY run in one process).  Everything here is O(N^2) integer work on the host.
//
//   Seq_0(b) = [b];   Seq_{r+1}(g) = Seq_r(g) ++ Seq_r(q) for the partners q != g of g in round r, by position
//   pixel -> piece: per dimension, nested balanced splits of the INCLUSIVE interval (range_length = max - min, so
//   neighbouring pieces share their boundary pixel; the higher gid is pasted later and owns it).
#pragma once
#include <algorithm>
#include <vector>

namespace vr
{
namespace radixk
{
constexpr int kMagicK = 8;   // RadixKCompositor.cpp:144
constexpr int kMaxBlocks = 16; // = kMaxRanks

struct Schedule
{
  int n = 0;
  int divisions[2] = { 1, 1 };
  int n_rounds = 0;
  int round_dim[8], round_size[8], round_step[8];
  int lo[2][kMaxBlocks];               // first 0-based pixel of each block column / row (effective span)
  int seq[kMaxBlocks][kMaxBlocks] = {}; // seq[g][i]: i-th rank composited for the pixels gid g ends up owning
  unsigned char pos[kMaxBlocks][kMaxBlocks] = {}; // pos[g][rank]: inverse of seq[g]
};

// RegularDecomposer::factor (diy/decomposition.hpp:596-611): prime factors, ascending
inline void prime_factors(int n, std::vector<int>& f)
{
  while (n != 1)
    for (int i = 2; i <= n; ++i)
      if (n % i == 0)
      {
        f.push_back(i);
        n /= i;
        break;
      }
}

// RegularDecomposer<DiscreteBounds>::fill_divisions (diy/decomposition.hpp:514-594) for dim = 2, no user divisions
inline bool fill_divisions(int n, int W, int H, int divisions[2])
{
  struct Div { int dim, nb, b_size; };
  const int dmin[2] = { 1, 1 }, dmax[2] = { W, H };
  Div d[2] = { { 0, 1, dmax[0] - dmin[0] }, { 1, 1, dmax[1] - dmin[1] } };
  std::vector<int> f;
  prime_factors(n, f);
  for (int i = (int)f.size() - 1; i >= 0; --i)
  {
    // larger block first; ties: fewer blocks so far, then the lower dimension
    const bool swap = d[1].b_size != d[0].b_size ? d[1].b_size > d[0].b_size
                      : d[1].nb != d[0].nb       ? d[1].nb < d[0].nb
                                                 : d[1].dim < d[0].dim;
    if (swap) std::swap(d[0], d[1]);
    const int nn = d[0].nb * f[i];
    const int lo = dmin[d[0].dim];
    const int hi = nn == 1 ? dmax[d[0].dim] : lo + (dmax[d[0].dim] - lo + 1) / nn - 1;
    if (hi < lo) return false; // "Unable to decompose domain into n blocks"
    d[0].nb = nn;
    d[0].b_size = hi - lo;
  }
  divisions[d[0].dim] = d[0].nb;
  divisions[d[1].dim] = d[1].nb;
  return true;
}

// RegularPartners::factor(k, tot_b, kv) (diy/partners/common.hpp:170-201)
inline void factor_k(int k, int tot, std::vector<int>& kv)
{
  int rem = tot;
  while (rem > 1)
  {
    if (rem % k == 0) { kv.push_back(k); rem /= k; continue; }
    int j = k - 1;
    for (; j > 1; --j)
      if (rem % j == 0) { kv.push_back(j); rem /= j; break; }
    if (j == 1) { kv.push_back(rem); rem = 1; }
  }
}

inline bool make_schedule(int n, int W, int H, Schedule& s)
{
  if (n < 1 || n > kMaxBlocks || W < 1 || H < 1) return false;
  s.n = n;
  if (!fill_divisions(n, W, H, s.divisions)) return false;
  // rounds: per-dimension factorisations interleaved dimension by dimension (common.hpp:139-166); steps for
  // contiguous = false (:75-83)
  std::vector<int> per_dim[2];
  factor_k(kMagicK, s.divisions[0], per_dim[0]);
  factor_k(kMagicK, s.divisions[1], per_dim[1]);
  s.n_rounds = 0;
  size_t at[2] = { 0, 0 };
  int cur[2] = { s.divisions[0], s.divisions[1] };
  for (bool changed = true; changed;)
  {
    changed = false;
    for (int d = 0; d < 2; ++d)
      if (at[d] < per_dim[d].size())
      {
        const int size = per_dim[d][at[d]++];
        cur[d] /= size;
        s.round_dim[s.n_rounds] = d;
        s.round_size[s.n_rounds] = size;
        s.round_step[s.n_rounds] = cur[d];
        ++s.n_rounds;
        changed = true;
      }
  }
  // fold sequences
  std::vector<std::vector<int>> seq(n), nxt(n);
  for (int g = 0; g < n; ++g) seq[g].assign(1, g);
  for (int r = 0; r < s.n_rounds; ++r)
  {
    const int d = s.round_dim[r], size = s.round_size[r], step = s.round_step[r];
    for (int g = 0; g < n; ++g)
    {
      int c[2] = { g % s.divisions[0], g / s.divisions[0] };
      const int pos = c[d] / step % size; // RegularPartners::group_position
      const int first = c[d] - pos * step;
      nxt[g] = seq[g];
      for (int j = 0; j < size; ++j) // RegularPartners::fill: ascending position = link order
      {
        c[d] = first + j * step;
        const int q = c[0] + s.divisions[0] * c[1];
        if (q != g) nxt[g].insert(nxt[g].end(), seq[q].begin(), seq[q].end());
      }
    }
    seq.swap(nxt);
  }
  for (int g = 0; g < n; ++g)
  {
    if ((int)seq[g].size() != n) return false;
    for (int i = 0; i < n; ++i)
    {
      s.seq[g][i] = seq[g][i];
      s.pos[g][seq[g][i]] = (unsigned char)i;
    }
  }
  // pieces: reduce_images' balanced ranges (RadixKCompositor.cpp:66-92), nested round after round
  for (int d = 0; d < 2; ++d)
  {
    std::vector<std::pair<int, int>> spans(1, { 1, d == 0 ? W : H }), fine;
    for (int r = 0; r < s.n_rounds; ++r)
    {
      if (s.round_dim[r] != d) continue;
      fine.clear();
      for (const auto& sp : spans)
      {
        const int length = sp.second - sp.first, size = s.round_size[r];
        const int base = length / size, rem = length % size;
        int m = sp.first;
        for (int i = 0; i < size; ++i)
        {
          const int b = base + (i < rem ? 1 : 0);
          fine.push_back({ m, m + b });
          m += b;
        }
      }
      spans.swap(fine);
    }
    for (int i = 0; i < kMaxBlocks; ++i) s.lo[d][i] = 0x7fffffff;
    for (size_t i = 0; i < spans.size(); ++i) s.lo[d][i] = i == 0 ? 0 : spans[i].first - 1;
  }
  return true;
}
} // namespace radixk
} // namespace vr
