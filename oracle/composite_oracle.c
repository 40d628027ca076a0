/*
 * composite_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the sort-last compositing half of Ascent's `volume`
 * plot: the uint8 visibility-ordered image fold (path A) and the float depth-sorted
 * partial fold (path B).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file's shared object.
 *
 * PARITY PIN: this restatement is checked (tests/test_oracle_composite.py)
 *   (a) against the reference's own apcomp sources compiled in place into
 *       oracle/_ref/libapcomp_ref.so (oracle/Makefile) on random inputs, bit-exact, and
 *   (b) against the reference's golden PNGs src/tests/_baseline_images/apcomp/
 *       apcomp_{c_order,volume_partial,zbuffer}{,_mpi}.png for the scenes of
 *       src/tests/apcomp/t_apcomp_{c_order,volume_partials,zbuffer}.cpp.
 *
 * Ids (C1..C5, P1..P5) refer to SURVEY.md section 8(a).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))
#define STD_MIN(a, b) (((b) < (a)) ? (b) : (a))

typedef struct
{
  int32_t pixel_id;
  float depth;
  float rgb[3];
  float alpha;
} orc_partial;

/* ---- C1: Image::Init(const float*,...) src/libs/vtkh/compositing/Image.hpp:80-113
 * (apcomp twin src/libs/apcomp/image.cpp:56-93).  mode 0 = vtk-h rule (depth<0 -> |d|),
 * mode 1 = apcomp gl_depth rule (depth<0 -> 2.0), mode 2 = verbatim (apcomp vis-order). */
ORC_API void orc_image_init(const float* rgba, const float* depth, int n_pixels, int depth_mode,
                            uint8_t* out_rgba, float* out_depth)
{
#pragma omp parallel for
  for (int i = 0; i < n_pixels; ++i)
  {
    const int o = i * 4;
    out_rgba[o + 0] = (unsigned char)(rgba[o + 0] * 255.f);
    out_rgba[o + 1] = (unsigned char)(rgba[o + 1] * 255.f);
    out_rgba[o + 2] = (unsigned char)(rgba[o + 2] * 255.f);
    out_rgba[o + 3] = (unsigned char)(rgba[o + 3] * 255.f);
    float d = depth[i];
    if (depth_mode == 0) d = d < 0 ? fabsf(d) : d;
    else if (depth_mode == 1) d = d < 0 ? 2.f : d;
    out_depth[i] = d;
  }
}

/* ---- C2: ImageCompositor::Blend, src/libs/vtkh/compositing/ImageCompositor.hpp:15-47
 * (front over back, in place into front) */
ORC_API void orc_blend(uint8_t* front_rgba, float* front_depth, const uint8_t* back_rgba,
                       const float* back_depth, int n_pixels)
{
#pragma omp parallel for
  for (int i = 0; i < n_pixels; ++i)
  {
    const int o = i * 4;
    unsigned int alpha = front_rgba[o + 3];
    const unsigned int opacity = 255 - alpha;
    front_rgba[o + 0] += (unsigned char)(opacity * back_rgba[o + 0] / 255);
    front_rgba[o + 1] += (unsigned char)(opacity * back_rgba[o + 1] / 255);
    front_rgba[o + 2] += (unsigned char)(opacity * back_rgba[o + 2] / 255);
    front_rgba[o + 3] += (unsigned char)(opacity * back_rgba[o + 3] / 255);
    /* std::min(a,b) = (b < a) ? b : a  -- NOT fminf: a NaN first argument survives */
    float d1 = STD_MIN(front_depth[i], 1.001f);
    float d2 = STD_MIN(back_depth[i], 1.001f);
    front_depth[i] = STD_MIN(d1, d2);
  }
}

/* ---- C2/C3: OrderedComposite (ImageCompositor.hpp:78-86): sort by composite order,
 * left-fold Blend into the first.  DirectSendCompositor::CompositeVolume
 * (DirectSendCompositor.cpp:121-181) is per-pixel identical to this serial fold for any
 * tiling.  layers: n_images x (n_pixels*4) uint8, depths n_images x n_pixels.
 * vis_order[i] = composite order of image i.  Result in out_*. */
ORC_API void orc_ordered_composite(const uint8_t* layers_rgba, const float* layers_depth,
                                   const int* vis_order, int n_images, int n_pixels,
                                   uint8_t* out_rgba, float* out_depth)
{
  int* idx = (int*)malloc(sizeof(int) * (size_t)n_images);
  for (int i = 0; i < n_images; ++i) idx[i] = i;
  for (int i = 1; i < n_images; ++i)
  {
    int k = idx[i], j = i - 1;
    while (j >= 0 && vis_order[idx[j]] > vis_order[k]) { idx[j + 1] = idx[j]; --j; }
    idx[j + 1] = k;
  }
  memcpy(out_rgba, layers_rgba + (size_t)idx[0] * n_pixels * 4, (size_t)n_pixels * 4);
  memcpy(out_depth, layers_depth + (size_t)idx[0] * n_pixels, sizeof(float) * (size_t)n_pixels);
  for (int i = 1; i < n_images; ++i)
    orc_blend(out_rgba, out_depth, layers_rgba + (size_t)idx[i] * n_pixels * 4,
              layers_depth + (size_t)idx[i] * n_pixels, n_pixels);
  free(idx);
}

/* ---- C4: ImageCompositor::ZBufferComposite.  vtk-h (src/libs/vtkh/compositing/ImageCompositor.hpp:
 * 49-76) has one form: a fragment with depth > 1 never replaces -- call with gl_depth = 1 (its
 * abs(depth) is the identity: Image::Init has already made depths non-negative).  apcomp
 * (src/libs/apcomp/internal/ImageCompositor.hpp:67-122) switches on m_gl_depth: 1 = the same rule,
 * 0 = plain nearest-wins. */
ORC_API void orc_zbuffer_composite(uint8_t* front_rgba, float* front_depth,
                                   const uint8_t* img_rgba, const float* img_depth, int n_pixels,
                                   int gl_depth)
{
#pragma omp parallel for
  for (int i = 0; i < n_pixels; ++i)
  {
    const float depth = img_depth[i];
    if (gl_depth && depth > 1.f) continue;
    if (front_depth[i] < depth) continue;
    front_depth[i] = depth;
    memcpy(front_rgba + 4 * i, img_rgba + 4 * i, 4);
  }
}

/* ---- V10: Renderer::ImageToCanvas, src/libs/vtkh/rendering/Renderer.cpp:265-283 */
ORC_API void orc_image_to_canvas(const uint8_t* rgba, const float* depth, int n_pixels,
                                 float* canvas_rgba, float* canvas_depth)
{
  float one_over_255 = 1.f / 255.f;
  for (int i = 0; i < n_pixels * 4; ++i) canvas_rgba[i] = (float)rgba[i] * one_over_255;
  if (canvas_depth) memcpy(canvas_depth, depth, sizeof(float) * (size_t)n_pixels);
}

/* ---- Render::RenderBackground (src/libs/vtkh/rendering/Render.cpp:277-286) -> vtkm::rendering::
 * Canvas::BlendBackground [VTK-m 2.1.0, recalled: vtkm/rendering/Canvas.cxx BlendBackground worklet]:
 * per pixel, if alpha >= 1 keep; else a = bg.a * (1 - alpha); rgb += bg.rgb * a; alpha = a + alpha.
 * Pinned end to end by the reference's render goldens (tests/test_oracle_golden.py). */
ORC_API void orc_blend_background(float* canvas_rgba, int n_pixels, const float* bg)
{
  for (int i = 0; i < n_pixels; ++i)
  {
    float* c = canvas_rgba + 4 * (size_t)i;
    if (c[3] >= 1.f) continue;
    const float alpha = bg[3] * (1.f - c[3]);
    c[0] = c[0] + bg[0] * alpha;
    c[1] = c[1] + bg[1] * alpha;
    c[2] = c[2] + bg[2] * alpha;
    c[3] = alpha + c[3];
  }
}

/* ---- PNGEncoder::Encode(const float*,...), src/libs/png_utils/ascent_png_encoder.cpp:257-281 (called by
 * Render::Save, Render.cpp:299-312): (unsigned char)(c * 255.f) per channel, rows flipped vertically. */
ORC_API void orc_encode_rgba8(const float* canvas_rgba, int width, int height, int flip, uint8_t* out)
{
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x)
    {
      const size_t in = ((size_t)y * width + x) * 4;
      const size_t o = ((size_t)(flip ? height - y - 1 : y) * width + x) * 4;
      for (int k = 0; k < 4; ++k)
        out[o + k] = (uint8_t)(long long)(canvas_rgba[in + k] * 255.f); /* x86: cvttss2si, low byte */
    }
}

/* ---- P1: VolumePartial::blend, src/libs/vtkh/compositing/VolumePartial.hpp:86-95 */
static inline void partial_blend(orc_partial* a, const orc_partial* o)
{
  if (a->alpha >= 1.f || o->alpha == 0.f) return;
  const float opacity = (1.f - a->alpha);
  a->rgb[0] += opacity * o->rgb[0];
  a->rgb[1] += opacity * o->rgb[1];
  a->rgb[2] += opacity * o->rgb[2];
  a->alpha += opacity * o->alpha;
  a->alpha = a->alpha > 1.f ? 1.f : a->alpha;
}

/* sort key: (pixel_id, depth) lexicographic, VolumePartial.hpp:74-84.  The reference's
 * std::sort is unstable; the documented tie-break is input order, i.e.
 * (rank, domain, ray) order, realised here with a stable merge sort. */
static int partial_less(const orc_partial* a, const orc_partial* b)
{
  if (a->pixel_id != b->pixel_id) return a->pixel_id < b->pixel_id;
  return a->depth < b->depth;
}
static void merge_sort(orc_partial* a, orc_partial* tmp, int64_t n)
{
  if (n < 2) return;
  int64_t h = n / 2;
  merge_sort(a, tmp, h);
  merge_sort(a + h, tmp, n - h);
  int64_t i = 0, j = h, k = 0;
  while (i < h && j < n) tmp[k++] = partial_less(&a[j], &a[i]) ? a[j++] : a[i++];
  while (i < h) tmp[k++] = a[i++];
  while (j < n) tmp[k++] = a[j++];
  memcpy(a, tmp, sizeof(orc_partial) * (size_t)n);
}

/* ---- P2+P4: PartialCompositor::merge + composite_partials + BlendPartials,
 * src/libs/vtkh/compositing/PartialCompositor.cpp:242-326,329-488,56-95.
 * `partials` (n, concatenated in (rank,domain) order) is sorted in place.  Output order
 * follows the reference: singletons first (ascending pixel), then folded multi-partial
 * pixels (ascending pixel).  Returns the number of output partials. */
ORC_API int64_t orc_composite_partials(orc_partial* partials, int64_t n, orc_partial* out)
{
  if (n == 0) return 0;
  orc_partial* tmp = (orc_partial*)malloc(sizeof(orc_partial) * (size_t)n);
  merge_sort(partials, tmp, n);
  free(tmp);
  int64_t n_unique = 0, n_seg = 0;
  /* pass 1: singletons */
  for (int64_t i = 0; i < n; ++i)
  {
    int begin = (i == 0) || partials[i].pixel_id != partials[i - 1].pixel_id;
    int has_work = (i + 1 < n) && partials[i].pixel_id == partials[i + 1].pixel_id;
    if (begin && !has_work) out[n_unique++] = partials[i];
  }
  /* pass 2: folded segments */
  for (int64_t i = 0; i < n; ++i)
  {
    int begin = (i == 0) || partials[i].pixel_id != partials[i - 1].pixel_id;
    int has_work = (i + 1 < n) && partials[i].pixel_id == partials[i + 1].pixel_id;
    if (!(begin && has_work)) continue;
    orc_partial result = partials[i];
    int64_t j = i + 1;
    while (j < n && partials[j].pixel_id == result.pixel_id)
    {
      partial_blend(&result, &partials[j]);
      ++j;
    }
    out[n_unique + n_seg++] = result;
  }
  return n_unique + n_seg;
}

/* ---- P3: redistribute ownership, vtkh_diy_partial_redistribute.hpp:133-150 with
 * RegularDecomposer<DiscreteBounds>::point_to_gid (decomposition.hpp:648-666,
 * BoundsHelper::lower :37-46): 1-D regular split of [min_pixel,max_pixel] into n_ranks. */
ORC_API int orc_partial_owner(int pixel_id, int min_pixel, int max_pixel, int n_ranks)
{
  int width = (max_pixel - min_pixel + 1) / n_ranks;
  if (width < 1) width = 1; /* fewer pixels than ranks divides by zero in DIY; guard */
  int res = (pixel_id - min_pixel) / width;
  if (res >= n_ranks) res = n_ranks - 1;
  if (res < 0) res = 0;
  return res;
}
