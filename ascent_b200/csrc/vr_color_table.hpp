// vr_color_table.hpp -- host-side transfer-function sampling: what
// vtkm::cont::ColorTable::Sample(1024, Vec4ui_8) + vtkh::detail::convert_table
// (src/libs/vtkh/rendering/VolumeRenderer.cpp:64-91) hand the volume mapper.  Evaluated once per
// plot on the CPU; the device only ever sees the finished 1024 x float4 table (vr_set_tf).
//
// VTK-m's exec-side table works in Float32 (node colours, CIELAB/Msh conversions, blend weights)
// with Float64 node positions, and Sample() walks the range by accumulating a Float32 step.  Both
// details are visible in the reference's golden images (the colour bars are Sample(bar height) of
// the same tables) and are what tests/test_oracle_colortable.py holds this file to.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace vr
{
namespace ct
{

enum Space { kRGB = 0, kLab = 1, kDiverging = 2 };

struct Rgb
{
  float c[3];
};
struct Lab
{
  float L, a, b;
};
struct Msh
{
  float M, s, h;
};

constexpr float kPiF = 3.14159265358979323846f;

inline float srgb_expand(float v) { return v > 0.04045f ? std::pow((v + 0.055f) / 1.055f, 2.4f) : v / 12.92f; }
inline float srgb_compress(float v) { return v > 0.0031308f ? 1.055f * std::pow(v, 1.0f / 2.4f) - 0.055f : 12.92f * v; }
inline float lab_f(float t) { return t > 0.008856f ? std::pow(t, 1.0f / 3.0f) : (7.787f * t) + (16.0f / 116.0f); }
inline float lab_finv(float v)
{
  const float cube = std::pow(v, 3.0f);
  return cube > 0.008856f ? cube : (v - 16.0f / 116.0f) / 7.787f;
}

inline Lab to_lab(const Rgb& in)
{
  const float r = srgb_expand(in.c[0]), g = srgb_expand(in.c[1]), b = srgb_expand(in.c[2]);
  // D65, 2 degree observer
  const float X = r * 0.4124f + g * 0.3576f + b * 0.1805f;
  const float Y = r * 0.2126f + g * 0.7152f + b * 0.0722f;
  const float Z = r * 0.0193f + g * 0.1192f + b * 0.9505f;
  const float fx = lab_f(X / 0.9505f), fy = lab_f(Y / 1.000f), fz = lab_f(Z / 1.089f);
  return Lab{ (116.0f * fy) - 16.0f, 500.0f * (fx - fy), 200.0f * (fy - fz) };
}

inline Rgb to_rgb(const Lab& in)
{
  const float fy = (in.L + 16.0f) / 116.0f;
  const float fx = in.a / 500.0f + fy;
  const float fz = fy - in.b / 200.0f;
  const float Y = 1.000f * lab_finv(fy), X = 0.9505f * lab_finv(fx), Z = 1.089f * lab_finv(fz);
  float r = srgb_compress(X * 3.2406f + Y * -1.5372f + Z * -0.4986f);
  float g = srgb_compress(X * -0.9689f + Y * 1.8758f + Z * 0.0415f);
  float b = srgb_compress(X * 0.0557f + Y * -0.2040f + Z * 1.0570f);
  // out-of-gamut colours: scale into range, drop negatives
  const float top = std::max(r, std::max(g, b));
  if (top > 1.0f) { r /= top; g /= top; b /= top; }
  return Rgb{ { std::max(r, 0.f), std::max(g, 0.f), std::max(b, 0.f) } };
}

inline Msh to_msh(const Lab& in)
{
  Msh o;
  o.M = std::sqrt(in.L * in.L + in.a * in.a + in.b * in.b);
  o.s = o.M > 0.001f ? std::acos(in.L / o.M) : 0.0f;
  o.h = o.s > 0.001f ? std::atan2(in.b, in.a) : 0.0f;
  return o;
}
inline Lab to_lab(const Msh& in)
{
  return Lab{ in.M * std::cos(in.s), in.M * std::sin(in.s) * std::cos(in.h), in.M * std::sin(in.s) * std::sin(in.h) };
}

inline float hue_gap(float h1, float h2)
{
  float d = std::fabs(h1 - h2);
  while (d >= 2.0f * kPiF) d -= 2.0f * kPiF;
  return d > kPiF ? (2.0f * kPiF) - d : d;
}
// the hue an unsaturated colour should pretend to have when it is blended towards `sat`
inline float spun_hue(const Msh& sat, float unsat_M)
{
  if (sat.M >= unsat_M - 0.1f) return sat.h;
  const float spin = sat.s * std::sqrt(unsat_M * unsat_M - sat.M * sat.M) / (sat.M * std::sin(sat.s));
  return sat.h > -0.3f * kPiF ? sat.h + spin : sat.h - spin;
}

// Moreland's diverging interpolation: through Msh space, with white inserted between two saturated
// colours of clearly different hue
inline Rgb blend_diverging(const Rgb& c1, const Rgb& c2, float w)
{
  Msh m1 = to_msh(to_lab(c1)), m2 = to_msh(to_lab(c2));
  if (m1.s > 0.05f && m2.s > 0.05f && hue_gap(m1.h, m2.h) > 0.33f * kPiF)
  {
    const float Mmid = std::max(88.0f, std::max(m1.M, m2.M));
    if (w < 0.5f) { m2 = Msh{ Mmid, 0.f, 0.f }; w = 2.0f * w; }
    else { m1 = Msh{ Mmid, 0.f, 0.f }; w = 2.0f * w - 1.0f; }
  }
  if (m1.s < 0.05f && m2.s > 0.05f) m1.h = spun_hue(m2, m1.M);
  else if (m2.s < 0.05f && m1.s > 0.05f) m2.h = spun_hue(m1, m2.M);
  const Msh mix{ (1.0f - w) * m1.M + w * m2.M, (1.0f - w) * m1.s + w * m2.s, (1.0f - w) * m1.h + w * m2.h };
  return to_rgb(to_lab(mix));
}
inline Rgb blend_lab(const Rgb& c1, const Rgb& c2, float w)
{
  const Lab l1 = to_lab(c1), l2 = to_lab(c2);
  return to_rgb(Lab{ (1.0f - w) * l1.L + w * l2.L, (1.0f - w) * l1.a + w * l2.a, (1.0f - w) * l1.b + w * l2.b });
}

// nodes sorted by position; positions Float64, values Float32 (as vtkm::cont::ColorTable stores them)
struct Table
{
  int space = kRGB;
  std::vector<double> color_x;
  std::vector<Rgb> color;
  std::vector<double> alpha_x;
  std::vector<float> alpha;
};

// index of the node interval [k, k+1] that holds x (x strictly inside the node range)
inline int interval_of(const std::vector<double>& xs, double x)
{
  int k = 0;
  const int last = (int)xs.size() - 2;
  while (k < last && !(x <= xs[k + 1])) ++k;
  return k;
}

inline Rgb color_at(const Table& t, double x)
{
  if (t.color.empty()) return Rgb{ { 0.f, 0.f, 0.f } };
  if (t.color.size() == 1 || x <= t.color_x.front()) return t.color.front();
  if (x >= t.color_x.back()) return t.color.back();
  const int k = interval_of(t.color_x, x);
  if (x == t.color_x[k + 1]) return t.color[k + 1];
  const float w = (float)((x - t.color_x[k]) / (t.color_x[k + 1] - t.color_x[k]));
  const Rgb &lo = t.color[k], &hi = t.color[k + 1];
  if (t.space == kDiverging) return blend_diverging(lo, hi, w);
  if (t.space == kLab) return blend_lab(lo, hi, w);
  return Rgb{ { (1.0f - w) * lo.c[0] + w * hi.c[0], (1.0f - w) * lo.c[1] + w * hi.c[1], (1.0f - w) * lo.c[2] + w * hi.c[2] } };
}

// opacity nodes made by AddPointAlpha(x, a): midpoint 0.5, sharpness 0 -> linear, Float32 weight
inline float alpha_at(const Table& t, double x)
{
  if (t.alpha.empty()) return 1.0f;
  if (t.alpha.size() == 1 || x <= t.alpha_x.front()) return t.alpha.front();
  if (x >= t.alpha_x.back()) return t.alpha.back();
  const int k = interval_of(t.alpha_x, x);
  float w = (float)((x - t.alpha_x[k]) / (t.alpha_x[k + 1] - t.alpha_x[k]));
  const float midpoint = 0.5f;
  w = w < midpoint ? 0.5f * w / midpoint : 0.5f + 0.5f * (w - midpoint) / (1.0f - midpoint);
  return (1.0f - w) * t.alpha[k] + w * t.alpha[k + 1];
}

// ColorTable::Sample(n, Vec4ui_8): n >= 2 positions over [0, 1], stepped by Float32 accumulation,
// the last one pinned to the range end; channels rounded to uint8
inline void sample_u8(const Table& t, int n, uint8_t* rgba8)
{
  const float step = 1.0f / (float)(n - 1);
  float pos = 0.0f;
  for (int i = 0; i < n; ++i, pos += step)
  {
    const double x = (i + 1 == n) ? 1.0 : (double)pos;
    const Rgb c = color_at(t, x);
    const float ch[4] = { c.c[0], c.c[1], c.c[2], alpha_at(t, x) };
    for (int k = 0; k < 4; ++k) rgba8[4 * i + k] = (uint8_t)(ch[k] * 255.0f + 0.5f);
  }
}

} // namespace ct
} // namespace vr
