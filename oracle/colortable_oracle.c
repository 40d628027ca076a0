/*
 * colortable_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, float32 like VTK-m's exec-side ColorTable, -ffp-contract=off) of
 * the 1024-entry transfer-function table the volume mapper
 * samples.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object.
 *
 * Where the algorithm lives: vtkm::cont::ColorTable (VTK-m v2.1.0, pinned at
 * scripts/build_ascent/build_ascent.sh:564; ABSENT from /root/reference).  Reference call sites:
 *   src/libs/vtkh/rendering/VolumeRenderer.cpp:64-91   convert_table: Sample(1024, Vec4ui_8) * 1/255
 *   src/libs/vtkh/rendering/VolumeRenderer.cpp:395-408 default table + the two AddPointAlpha calls
 *   src/libs/vtkh/rendering/VolumeRenderer.cpp:448-466 CorrectOpacity on the alpha nodes
 *   src/libs/ascent/runtimes/flow_filters/ascent_runtime_conduit_to_vtkm_parsing.cpp:199-305
 * Published algorithm restated here: K. Moreland, "Diverging Color Maps for Scientific
 * Visualization" (ISVC 2009), section 5 / appendix (RGB -> XYZ -> CIELAB -> Msh, hue adjustment,
 * white mid-point insertion) as VTK-m's exec/ColorTable.hxx implements it in Float32, and
 * ColorTable::Sample's position generation (float start + delta * i, last sample = range max).
 *
 * PARITY PIN: the colour path IS pinned against the reference's own output -- the colour bars VTK-m
 * draws into the golden images with Canvas::AddColorBar are ColorTable::Sample(bar_height) of the
 * same table: 179 samples of "Cool to Warm" in tout_render_mpi_3d_diy_volume100.png, 359 in
 * tout_render_3d_multi_default_runtime100.png, 359 of "Rainbow Desaturated" in the latter
 * (tests/golden/colorbars.npz, tests/test_oracle_colortable.py).  The alpha path (piecewise-linear
 * in f64 between alpha nodes) has no golden of its own; it is pinned through the render goldens.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

enum { CT_RGB = 0, CT_LAB = 1, CT_DIVERGING = 2 };

#define CT_MAX_NODES 64
typedef struct
{
  int space;
  int n_color;
  double color_x[CT_MAX_NODES];
  float color_rgb[CT_MAX_NODES][3];
  int n_alpha;
  double alpha_x[CT_MAX_NODES];
  float alpha_v[CT_MAX_NODES];
} orc_colortable;

/* ---- colour-space conversions, Float32 (vtkm/exec/ColorTable.hxx) */
static void rgb_to_lab(const float rgb[3], float lab[3])
{
  float r = rgb[0], g = rgb[1], b = rgb[2];
  /* sRGB gamma expansion */
  r = (r > 0.04045f) ? powf((r + 0.055f) / 1.055f, 2.4f) : r / 12.92f;
  g = (g > 0.04045f) ? powf((g + 0.055f) / 1.055f, 2.4f) : g / 12.92f;
  b = (b > 0.04045f) ? powf((b + 0.055f) / 1.055f, 2.4f) : b / 12.92f;
  /* Observer = 2 deg, Illuminant = D65 */
  float x = r * 0.4124f + g * 0.3576f + b * 0.1805f;
  float y = r * 0.2126f + g * 0.7152f + b * 0.0722f;
  float z = r * 0.0193f + g * 0.1192f + b * 0.9505f;
  const float ref_X = 0.9505f, ref_Y = 1.000f, ref_Z = 1.089f;
  float vx = x / ref_X, vy = y / ref_Y, vz = z / ref_Z;
  vx = (vx > 0.008856f) ? powf(vx, 1.0f / 3.0f) : (7.787f * vx) + (16.0f / 116.0f);
  vy = (vy > 0.008856f) ? powf(vy, 1.0f / 3.0f) : (7.787f * vy) + (16.0f / 116.0f);
  vz = (vz > 0.008856f) ? powf(vz, 1.0f / 3.0f) : (7.787f * vz) + (16.0f / 116.0f);
  lab[0] = (116.0f * vy) - 16.0f;
  lab[1] = 500.0f * (vx - vy);
  lab[2] = 200.0f * (vy - vz);
}

static void lab_to_rgb(const float lab[3], float rgb[3])
{
  float vy = (lab[0] + 16.0f) / 116.0f;
  float vx = lab[1] / 500.0f + vy;
  float vz = vy - lab[2] / 200.0f;
  vy = (powf(vy, 3.0f) > 0.008856f) ? powf(vy, 3.0f) : (vy - 16.0f / 116.0f) / 7.787f;
  vx = (powf(vx, 3.0f) > 0.008856f) ? powf(vx, 3.0f) : (vx - 16.0f / 116.0f) / 7.787f;
  vz = (powf(vz, 3.0f) > 0.008856f) ? powf(vz, 3.0f) : (vz - 16.0f / 116.0f) / 7.787f;
  const float ref_X = 0.9505f, ref_Y = 1.000f, ref_Z = 1.089f;
  const float x = ref_X * vx, y = ref_Y * vy, z = ref_Z * vz;
  float r = x * 3.2406f + y * -1.5372f + z * -0.4986f;
  float g = x * -0.9689f + y * 1.8758f + z * 0.0415f;
  float b = x * 0.0557f + y * -0.2040f + z * 1.0570f;
  /* sRGB gamma compression */
  r = (r > 0.0031308f) ? 1.055f * powf(r, 1.0f / 2.4f) - 0.055f : 12.92f * r;
  g = (g > 0.0031308f) ? 1.055f * powf(g, 1.0f / 2.4f) - 0.055f : 12.92f * g;
  b = (b > 0.0031308f) ? 1.055f * powf(b, 1.0f / 2.4f) - 0.055f : 12.92f * b;
  /* clip to the display gamut: scale down if any channel exceeds 1, clamp negatives */
  float m = r;
  if (m < g) m = g;
  if (m < b) m = b;
  if (m > 1.0f) { r /= m; g /= m; b /= m; }
  if (r < 0.f) r = 0.f;
  if (g < 0.f) g = 0.f;
  if (b < 0.f) b = 0.f;
  rgb[0] = r; rgb[1] = g; rgb[2] = b;
}

static void lab_to_msh(const float lab[3], float msh[3])
{
  const float L = lab[0], a = lab[1], b = lab[2];
  const float M = sqrtf(L * L + a * a + b * b);
  const float s = (M > 0.001f) ? acosf(L / M) : 0.0f;
  const float h = (s > 0.001f) ? atan2f(b, a) : 0.0f;
  msh[0] = M; msh[1] = s; msh[2] = h;
}

static void msh_to_lab(const float msh[3], float lab[3])
{
  const float M = msh[0], s = msh[1], h = msh[2];
  lab[0] = M * cosf(s);
  lab[1] = M * sinf(s) * cosf(h);
  lab[2] = M * sinf(s) * sinf(h);
}

static const float PI_F = 3.14159265358979323846f;

/* absolute difference of two angles, folded into [0, pi] */
static float angle_diff(float a1, float a2)
{
  float d = fabsf(a1 - a2);
  while (d >= 2.0f * PI_F) d -= 2.0f * PI_F;
  if (d > PI_F) d = (2.0f * PI_F) - d;
  return d;
}

/* hue for an unsaturated end point so that the interpolation towards it looks uniform (Moreland, eq. 6) */
static float adjust_hue(const float msh[3], float unsat_m)
{
  if (msh[0] >= unsat_m - 0.1f) return msh[2];
  const float spin = msh[1] * sqrtf(unsat_m * unsat_m - msh[0] * msh[0]) / (msh[0] * sinf(msh[1]));
  return (msh[2] > -0.3f * PI_F) ? msh[2] + spin : msh[2] - spin;
}

static void interp_diverging(const float rgb1[3], const float rgb2[3], float w, float out[3])
{
  float lab1[3], lab2[3], msh1[3], msh2[3];
  rgb_to_lab(rgb1, lab1);
  rgb_to_lab(rgb2, lab2);
  lab_to_msh(lab1, msh1);
  lab_to_msh(lab2, msh2);
  /* two distinct saturated colours: put white between them */
  if (msh1[1] > 0.05f && msh2[1] > 0.05f && angle_diff(msh1[2], msh2[2]) > 0.33f * PI_F)
  {
    float mmid = msh1[0] > msh2[0] ? msh1[0] : msh2[0];
    if (mmid < 88.0f) mmid = 88.0f;
    if (w < 0.5f)
    {
      msh2[0] = mmid; msh2[1] = 0.f; msh2[2] = 0.f;
      w = 2.0f * w;
    }
    else
    {
      msh1[0] = mmid; msh1[1] = 0.f; msh1[2] = 0.f;
      w = 2.0f * w - 1.0f;
    }
  }
  /* an unsaturated end has no meaningful hue: give it one */
  if (msh1[1] < 0.05f && msh2[1] > 0.05f) msh1[2] = adjust_hue(msh2, msh1[0]);
  else if (msh2[1] < 0.05f && msh1[1] > 0.05f) msh2[2] = adjust_hue(msh1, msh2[0]);
  float tmp[3], lab[3];
  for (int k = 0; k < 3; ++k) tmp[k] = (1.0f - w) * msh1[k] + w * msh2[k];
  msh_to_lab(tmp, lab);
  lab_to_rgb(lab, out);
}

static void interp_lab(const float rgb1[3], const float rgb2[3], float w, float out[3])
{
  float lab1[3], lab2[3], lab[3];
  rgb_to_lab(rgb1, lab1);
  rgb_to_lab(rgb2, lab2);
  for (int k = 0; k < 3; ++k) lab[k] = (1.0f - w) * lab1[k] + w * lab2[k];
  lab_to_rgb(lab, out);
}

/* exec ColorTable::MapThroughColorSpace(Float64 value), clamping on */
static void color_at(const orc_colortable* t, double x, float out[3])
{
  const int n = t->n_color;
  if (n == 0) { out[0] = out[1] = out[2] = 0.f; return; }
  if (x <= t->color_x[0] || n == 1) { memcpy(out, t->color_rgb[0], 3 * sizeof(float)); return; }
  if (x >= t->color_x[n - 1]) { memcpy(out, t->color_rgb[n - 1], 3 * sizeof(float)); return; }
  int first = 0, second = 1;
  for (; second < n - 1; ++first, ++second)
    if (x <= t->color_x[second]) break;
  if (x == t->color_x[second]) { memcpy(out, t->color_rgb[second], 3 * sizeof(float)); return; }
  const float w = (float)((x - t->color_x[first]) / (t->color_x[second] - t->color_x[first]));
  const float* a = t->color_rgb[first];
  const float* b = t->color_rgb[second];
  if (t->space == CT_DIVERGING) interp_diverging(a, b, w, out);
  else if (t->space == CT_LAB) interp_lab(a, b, w, out);
  else
    for (int k = 0; k < 3; ++k) out[k] = (1.0f - w) * a[k] + w * b[k];
}

/* exec ColorTable::MapThroughOpacitySpace(Float64 value): the weight is narrowed to Float32, bent
 * around the node's midpoint (0.5 for AddPointAlpha(x, a): the identity up to rounding) and, for
 * sharpness 0 -- the only kind AddPointAlpha(x, a) creates -- used for a Float32 linear blend */
static float alpha_at(const orc_colortable* t, double x)
{
  const int n = t->n_alpha;
  if (n == 0) return 1.0f;
  if (x <= t->alpha_x[0] || n == 1) return t->alpha_v[0];
  if (x >= t->alpha_x[n - 1]) return t->alpha_v[n - 1];
  int first = 0, second = 1;
  for (; second < n - 1; ++first, ++second)
    if (x <= t->alpha_x[second]) break;
  float w = (float)((x - t->alpha_x[first]) / (t->alpha_x[second] - t->alpha_x[first]));
  const float mid = 0.5f;
  if (w < mid) w = 0.5f * w / mid;
  else w = 0.5f + 0.5f * (w - mid) / (1.0f - mid);
  return (1.0f - w) * t->alpha_v[first] + w * t->alpha_v[second];
}

/* ColorTable::Sample(n, ArrayHandle<Vec4ui_8>): n positions over the table's range [0, 1] generated
 * in Float32 by ACCUMULATION (value = start; value += delta), the last one exactly the range max
 * -- the accumulated form, not start + delta * i, is what reproduces all 897 golden colour-bar
 * samples (i * delta differs at two rounding boundaries); each channel
 * static_cast<UInt8>(c * 255.0f + 0.5f).  out_u8: n x 4. */
ORC_API void orc_colortable_sample_u8(const orc_colortable* t, int n, uint8_t* out_u8)
{
  const float delta = 1.0f / (float)(n - 1);
  float value = 0.0f;
  for (int i = 0; i < n; ++i, value += delta)
  {
    const double x = (i == n - 1) ? 1.0 : (double)value;
    float c[4];
    color_at(t, x, c);
    c[3] = alpha_at(t, x);
    for (int k = 0; k < 4; ++k) out_u8[4 * i + k] = (uint8_t)(c[k] * 255.0f + 0.5f);
  }
}

/* convert_table, VolumeRenderer.cpp:64-91: uint8 samples -> float4 with 1/255.f */
ORC_API void orc_colortable_lut(const orc_colortable* t, int n, float* out_rgba)
{
  uint8_t u8[4 * 4096];
  if (n > 4096) n = 4096;
  orc_colortable_sample_u8(t, n, u8);
  const float k = 1.0f / 255.0f;
  for (int i = 0; i < 4 * n; ++i) out_rgba[i] = (float)u8[i] * k;
}

/* ---- table construction: presets + AddPoint / AddPointAlpha (same-position overwrite, B21) */
static void insert_color(orc_colortable* t, double x, float r, float g, float b)
{
  int i = 0;
  for (; i < t->n_color; ++i)
    if (t->color_x[i] == x) { t->color_rgb[i][0] = r; t->color_rgb[i][1] = g; t->color_rgb[i][2] = b; return; }
  if (t->n_color >= CT_MAX_NODES) return;
  i = t->n_color++;
  while (i > 0 && t->color_x[i - 1] > x)
  {
    t->color_x[i] = t->color_x[i - 1];
    memcpy(t->color_rgb[i], t->color_rgb[i - 1], 3 * sizeof(float));
    --i;
  }
  t->color_x[i] = x;
  t->color_rgb[i][0] = r; t->color_rgb[i][1] = g; t->color_rgb[i][2] = b;
}
static void insert_alpha(orc_colortable* t, double x, float a)
{
  int i = 0;
  for (; i < t->n_alpha; ++i)
    if (t->alpha_x[i] == x) { t->alpha_v[i] = a; return; }
  if (t->n_alpha >= CT_MAX_NODES) return;
  i = t->n_alpha++;
  while (i > 0 && t->alpha_x[i - 1] > x)
  {
    t->alpha_x[i] = t->alpha_x[i - 1];
    t->alpha_v[i] = t->alpha_v[i - 1];
    --i;
  }
  t->alpha_x[i] = x;
  t->alpha_v[i] = a;
}
static float clamp01(double v) { return (float)(v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v)); }

ORC_API void orc_colortable_add_point(orc_colortable* t, double x, double r, double g, double b)
{
  insert_color(t, x, clamp01(r), clamp01(g), clamp01(b));
}
ORC_API void orc_colortable_add_point_alpha(orc_colortable* t, double x, double a)
{
  insert_alpha(t, x, clamp01(a));
}
ORC_API void orc_colortable_clear_colors(orc_colortable* t) { t->n_color = 0; }
ORC_API void orc_colortable_clear_alpha(orc_colortable* t) { t->n_alpha = 0; }

/* presets [VTK-m ColorTablePresets, recalled; the first two are confirmed by the golden colour bars]:
 * 0 "Cool to Warm" (diverging), 1 "Rainbow Desaturated" (RGB), 2 "Black-Body Radiation" (RGB),
 * 3 grayscale (RGB).  Every preset starts with opaque alpha nodes (0,1),(1,1).  Returns 0 on success. */
ORC_API int orc_colortable_preset(orc_colortable* t, int which)
{
  memset(t, 0, sizeof(*t));
  insert_alpha(t, 0.0, 1.0f);
  insert_alpha(t, 1.0, 1.0f);
  switch (which)
  {
    case 0:
      t->space = CT_DIVERGING;
      insert_color(t, 0.0, (float)0.23137254902, (float)0.298039215686, (float)0.752941176471);
      insert_color(t, 0.5, (float)0.865, (float)0.865, (float)0.865);
      insert_color(t, 1.0, (float)0.705882352941, (float)0.0156862745098, (float)0.149019607843);
      return 0;
    case 1:
      t->space = CT_RGB;
      insert_color(t, 0.0, (float)0.278431372549, (float)0.278431372549, (float)0.858823529412);
      insert_color(t, 0.143, 0.0f, 0.0f, (float)0.360784313725);
      insert_color(t, 0.285, 0.0f, 1.0f, 1.0f);
      insert_color(t, 0.429, 0.0f, (float)0.501960784314, 0.0f);
      insert_color(t, 0.571, 1.0f, 1.0f, 0.0f);
      insert_color(t, 0.714, 1.0f, (float)0.380392156863, 0.0f);
      insert_color(t, 0.857, (float)0.419607843137, 0.0f, 0.0f);
      insert_color(t, 1.0, (float)0.878431372549, (float)0.301960784314, (float)0.301960784314);
      return 0;
    case 2:
      t->space = CT_RGB;
      insert_color(t, 0.0, 0.0f, 0.0f, 0.0f);
      insert_color(t, 0.4, (float)0.9, 0.0f, 0.0f);
      insert_color(t, 0.8, (float)0.9, (float)0.9, 0.0f);
      insert_color(t, 1.0, 1.0f, 1.0f, 1.0f);
      return 0;
    case 3:
      t->space = CT_RGB;
      insert_color(t, 0.0, 0.0f, 0.0f, 0.0f);
      insert_color(t, 1.0, 1.0f, 1.0f, 1.0f);
      return 0;
  }
  return 1;
}

/* VolumeRenderer::CorrectOpacity (VolumeRenderer.cpp:448-466) on every alpha node, in place:
 * alpha' = 1 - (1 - alpha)^(10 / samples), ratio formed in f32, pow in f64 */
ORC_API void orc_colortable_correct_opacity(orc_colortable* t, float samples)
{
  const float ratio = 10.f / samples; /* VTKH_OPACITY_CORRECTION, VolumeRenderer.cpp:25 */
  for (int i = 0; i < t->n_alpha; ++i)
    t->alpha_v[i] = clamp01(1. - pow(1. - (double)t->alpha_v[i], (double)ratio));
}
