/*
 * radixk_ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Runs the REFERENCE's radix-k surface compositor -- its own `reduce_images<Image>` callback
 * (src/libs/apcomp/internal/RadixKCompositor.cpp:37-131, the same text as
 * src/libs/vtkh/compositing/RadixKCompositor.cpp:35-124 with vtkm::Bounds replaced by apcomp::Bounds),
 * its own CollectImages (internal/apcomp_diy_collect.hpp) and the DIY it vendors
 * (internal/diy/include/diy: RegularDecomposer, RegularSwapPartners, reduce, all_to_all) -- for N "ranks"
 * inside ONE process: DIY is compiled with DIY_NO_MPI (its own single-process mode) and the N image
 * blocks all live on the one process (ContiguousAssigner(1, N)), so every queue goes through DIY's
 * same-rank delivery while partners, rounds, link order and the per-round image split are exactly what
 * N MPI ranks would run.  The only lines that are ours mirror CompositeImpl (RadixKCompositor.cpp:138-180)
 * with "one block per rank" replaced by "N blocks on this rank".
 *
 * The reference source is compiled where it lies (#include of the .cpp below); nothing is copied.
 * Used by tests/test_oracle_radixk.py to pin oracle.radixk_* (and, through it, the product's
 * vr_radixk_schedule + the z-select kernel's tie-break).
 */
#define DIY_NO_MPI
/* MPICollect.hpp (raw MPI_Send/Recv gather) is included by the reference file but its only call site is
 * commented out (RadixKCompositor.cpp:169); it cannot compile without MPI, so its include guard is pre-set. */
#define APCOMP_MPI_COLLECT_HPP
#include <apcomp/apcomp_config.h>
#include <apcomp/internal/RadixKCompositor.cpp> // the reference's translation unit, in place
#include <diy/assigner.hpp>
#include <cstring>
#include <vector>

namespace
{
struct AddBlocks
{
  std::vector<apcomp::Image>& images;
  apcompdiy::Master& master;
  template <typename B, typename L>
  void operator()(int gid, const B&, const B&, const B&, const L& link) const
  {
    master.add(gid, new apcomp::ImageBlock<apcomp::Image>(images[gid]), new L(link));
  }
};
} // namespace

extern "C" {

/* rgba: [n][h][w][4] uint8, depth: [n][h][w] f32.  out = block 0's image after reduce + collect.
 * rounds_out (optional, >= 3*16 ints): per round {dim, k}, terminated by -1; divisions in [0],[1] first. */
__attribute__((visibility("default"))) int
ref_radixk_zbuffer(const unsigned char* rgba, const float* depth, int n, int width, int height,
                   unsigned char* out_rgba, float* out_depth, int* info_out)
{
  try
  {
    const size_t np = (size_t)width * height;
    std::vector<apcomp::Image> images(n);
    for (int i = 0; i < n; ++i)
      images[i].Init(rgba + i * np * 4, depth + i * np, width, height, true);

    apcompdiy::mpi::communicator comm; /* DIY_NO_MPI: size 1 */
    apcompdiy::DiscreteBounds global_bounds = apcomp::BoundsToDIY(images[0].m_orig_bounds);
    const int magic_k = 8; /* RadixKCompositor.cpp:144 */
    apcompdiy::Master master(comm, 1, -1, 0,
                             [](void* b) { delete reinterpret_cast<apcomp::ImageBlock<apcomp::Image>*>(b); });
    apcompdiy::ContiguousAssigner assigner(1, n);
    AddBlocks create{ images, master };
    apcompdiy::RegularDecomposer<apcompdiy::DiscreteBounds> decomposer(2, global_bounds, n);
    decomposer.decompose(comm.rank(), assigner, create);
    apcompdiy::RegularSwapPartners partners(decomposer, magic_k, false);
    apcompdiy::reduce(master, assigner, partners, apcomp::reduce_images<apcomp::Image>);
    apcompdiy::all_to_all(master, assigner, apcomp::CollectImages<apcomp::Image>(decomposer), magic_k);

    if (info_out)
    {
      int k = 0;
      info_out[k++] = decomposer.divisions[0];
      info_out[k++] = decomposer.divisions[1];
      for (size_t r = 0; r < partners.rounds() && r < 14; ++r)
      {
        info_out[k++] = partners.dim((int)r);
        info_out[k++] = partners.size((int)r);
      }
      info_out[k] = -1;
    }
    const apcomp::Image& res = images[0];
    if (res.m_pixels.size() != np * 4 || res.m_depths.size() != np) return 2;
    std::memcpy(out_rgba, res.m_pixels.data(), np * 4);
    std::memcpy(out_depth, res.m_depths.data(), np * sizeof(float));
    return 0;
  }
  catch (...)
  {
    return 1;
  }
}
}
