/*
 * apcomp_ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * extern "C" doorway onto the REFERENCE's own compositor, compiled in place from
 * /root/reference/src/libs/apcomp/{image,compositor,partial_compositor,apcomp}.cpp by
 * oracle/Makefile into oracle/_ref/libapcomp_ref.so (serial build: no APCOMP_PARALLEL, there
 * is no MPI here; the N-rank direct-send result equals the serial N-image fold).
 * Used to validate oracle/composite_oracle.c and as the "reference" CPU baseline in bench.py.
 * API used: src/libs/apcomp/compositor.hpp:19-87, src/libs/apcomp/partial_compositor.hpp:24-43.
 */
#include <apcomp/apcomp.hpp>
#include <apcomp/compositor.hpp>
#include <apcomp/partial_compositor.hpp>
#include <cstring>
#include <vector>

extern "C" {

struct ref_partial
{
  int pixel_id;
  float depth;
  float rgb[3];
  float alpha;
};

/* VIS_ORDER_BLEND: N float images -> composited uint8 image (t_apcomp_c_order.cpp:27-69) */
__attribute__((visibility("default"))) void
ref_composite_vis_order(const float* rgba, const float* depth, const int* vis_order, int n_images,
                        int width, int height, unsigned char* out_rgba, float* out_depth)
{
  apcomp::Compositor compositor;
  compositor.SetCompositeMode(apcomp::Compositor::VIS_ORDER_BLEND);
  const size_t np = (size_t)width * height;
  for (int i = 0; i < n_images; ++i)
    compositor.AddImage(rgba + i * np * 4, depth + i * np, width, height, vis_order[i]);
  apcomp::Image img = compositor.Composite();
  std::memcpy(out_rgba, img.m_pixels.data(), np * 4);
  std::memcpy(out_depth, img.m_depths.data(), np * sizeof(float));
}

/* same, uint8 inputs (AddImage(const unsigned char*...)) */
__attribute__((visibility("default"))) void
ref_composite_vis_order_u8(const unsigned char* rgba, const float* depth, const int* vis_order,
                           int n_images, int width, int height, unsigned char* out_rgba,
                           float* out_depth)
{
  apcomp::Compositor compositor;
  compositor.SetCompositeMode(apcomp::Compositor::VIS_ORDER_BLEND);
  const size_t np = (size_t)width * height;
  for (int i = 0; i < n_images; ++i)
    compositor.AddImage(rgba + i * np * 4, depth + i * np, width, height, vis_order[i]);
  apcomp::Image img = compositor.Composite();
  std::memcpy(out_rgba, img.m_pixels.data(), np * 4);
  std::memcpy(out_depth, img.m_depths.data(), np * sizeof(float));
}

/* Z_BUFFER_SURFACE_GL: t_apcomp_zbuffer.cpp:28-72 */
__attribute__((visibility("default"))) void
ref_composite_zbuffer(const float* rgba, const float* depth, int n_images, int width, int height,
                      unsigned char* out_rgba, float* out_depth)
{
  apcomp::Compositor compositor;
  compositor.SetCompositeMode(apcomp::Compositor::Z_BUFFER_SURFACE_GL);
  const size_t np = (size_t)width * height;
  for (int i = 0; i < n_images; ++i)
    compositor.AddImage(rgba + i * np * 4, depth + i * np, width, height);
  apcomp::Image img = compositor.Composite();
  std::memcpy(out_rgba, img.m_pixels.data(), np * 4);
  std::memcpy(out_depth, img.m_depths.data(), np * sizeof(float));
}

/* PartialCompositor<VolumePartial<float>>::composite (t_apcomp_volume_partials.cpp:29-70).
 * counts[i] partials for partial-image i, concatenated in `in`.  out must hold sum(counts). */
__attribute__((visibility("default"))) long long
ref_composite_partials(const ref_partial* in, const long long* counts, int n_lists,
                       ref_partial* out)
{
  typedef apcomp::VolumePartial<float> P;
  static_assert(sizeof(P) == sizeof(ref_partial), "VolumePartial<float> must be 24 B");
  std::vector<std::vector<P>> lists(n_lists);
  size_t off = 0;
  for (int i = 0; i < n_lists; ++i)
  {
    lists[i].resize(counts[i]);
    std::memcpy((void*)lists[i].data(), in + off, counts[i] * sizeof(P));
    off += counts[i];
  }
  std::vector<P> result;
  apcomp::PartialCompositor<P> compositor;
  compositor.composite(lists, result);
  std::memcpy(out, (void*)result.data(), result.size() * sizeof(P));
  return (long long)result.size();
}

__attribute__((visibility("default"))) int ref_openmp_enabled() { return apcomp::openmp_enabled() ? 1 : 0; }
}
