// vr_host_math.hpp -- host-side camera / matrix arithmetic of the tracer (f32, like VTK-m's
// vtkm::rendering::Camera and raytracing::Camera, which the reference calls at
// src/libs/vtkh/rendering/VolumeRenderer.cpp:239-247,298-339).  Evaluated once per (block,
// camera) on the CPU; the per-ray work is in sampler.cu.
#pragma once
#include <cmath>
#include <cstring>

#include "../../include/vr_b200.h"

namespace vr
{
namespace hm
{

struct Vec3
{
  float v[3];
  float& operator[](int i) { return v[i]; }
  const float& operator[](int i) const { return v[i]; }
};
inline Vec3 make3(const float* p) { return Vec3{ { p[0], p[1], p[2] } }; }
inline Vec3 sub(const Vec3& a, const Vec3& b) { return Vec3{ { a[0] - b[0], a[1] - b[1], a[2] - b[2] } }; }
inline float dot(const Vec3& a, const Vec3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline float magnitude(const Vec3& a) { return std::sqrt(dot(a, a)); }
inline Vec3 normal(const Vec3& a)
{
  const float r = 1.0f / std::sqrt(dot(a, a));
  return Vec3{ { r * a[0], r * a[1], r * a[2] } };
}
// a*b - c*d with one error-compensation step (vtkm::DifferenceOfProducts)
inline float dop(float a, float b, float c, float d)
{
  const float cd = c * d;
  const float err = std::fma(-c, d, cd);
  const float r = std::fma(a, b, -cd);
  return r + err;
}
inline Vec3 cross(const Vec3& x, const Vec3& y)
{
  return Vec3{ { dop(x[1], y[2], x[2], y[1]), dop(x[2], y[0], x[0], y[2]), dop(x[0], y[1], x[1], y[0]) } };
}

constexpr float kPi180 = (float)0.01745329251994329547437168059786927;

struct Mat4
{
  float m[16]; // row major
  float& operator()(int r, int c) { return m[r * 4 + c]; }
  float operator()(int r, int c) const { return m[r * 4 + c]; }
};
inline Mat4 identity()
{
  Mat4 a;
  std::memset(a.m, 0, sizeof(a.m));
  a(0, 0) = a(1, 1) = a(2, 2) = a(3, 3) = 1.f;
  return a;
}
inline Mat4 mul(const Mat4& a, const Mat4& b)
{
  Mat4 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
    {
      float s = a(i, 0) * b(0, j);
      for (int k = 1; k < 4; ++k) s = s + a(i, k) * b(k, j);
      r(i, j) = s;
    }
  return r;
}
inline void mulv(const Mat4& a, const float v[4], float out[4])
{
  float r[4];
  for (int i = 0; i < 4; ++i) r[i] = a(i, 0) * v[0] + a(i, 1) * v[1] + a(i, 2) * v[2] + a(i, 3) * v[3];
  std::memcpy(out, r, sizeof(r));
}

// LUP factorisation with a unit-diagonal upper factor, then one solve per identity column
inline Mat4 inverse(const Mat4& in)
{
  Mat4 A = in, out;
  int perm[4] = { 0, 1, 2, 3 };
  for (int top = 0; top < 4; ++top)
  {
    int best = top;
    float bestv = std::fabs(A(top, top));
    for (int r = top + 1; r < 4; ++r)
      if (bestv < std::fabs(A(r, top))) { bestv = std::fabs(A(r, top)); best = r; }
    if (best != top)
    {
      for (int c = 0; c < 4; ++c) std::swap(A(best, c), A(top, c));
      std::swap(perm[best], perm[top]);
    }
    for (int c = top + 1; c < 4; ++c) A(top, c) /= A(top, top);
    for (int r = top + 1; r < 4; ++r)
      for (int c = top + 1; c < 4; ++c) A(r, c) -= A(r, top) * A(top, c);
  }
  for (int col = 0; col < 4; ++col)
  {
    float y[4], x[4];
    for (int r = 0; r < 4; ++r)
    {
      float b = perm[r] == col ? 1.f : 0.f;
      for (int c = 0; c < r; ++c) b -= A(r, c) * y[c];
      y[r] = b / A(r, r);
    }
    for (int r = 3; r >= 0; --r)
    {
      float b = y[r];
      for (int c = r + 1; c < 4; ++c) b -= A(r, c) * x[c];
      x[r] = b;
    }
    for (int r = 0; r < 4; ++r) out(r, col) = x[r];
  }
  return out;
}

inline Mat4 view_matrix(const vr_camera& c)
{
  const Vec3 pos = make3(c.position);
  Vec3 vd = sub(pos, make3(c.look_at));
  Vec3 right = cross(make3(c.up), vd);
  Vec3 ru = cross(vd, right);
  vd = normal(vd); right = normal(right); ru = normal(ru);
  Mat4 m = identity();
  for (int k = 0; k < 3; ++k) { m(0, k) = right[k]; m(1, k) = ru[k]; m(2, k) = vd[k]; }
  m(0, 3) = -dot(right, pos);
  m(1, 3) = -dot(ru, pos);
  m(2, 3) = -dot(vd, pos);
  return m;
}

inline Mat4 projection_matrix(const vr_camera& c, int width, int height)
{
  const float n = c.near_plane, f = c.far_plane;
  Mat4 m = identity();
  const float aspect = (float)width / (float)height;
  const float t = std::tan((c.fov * kPi180) * 0.5f);
  const float size = n * t;
  const float left = -size * aspect, right = size * aspect, bottom = -size, top = size;
  m(0, 0) = 2.f * n / (right - left);
  m(1, 1) = 2.f * n / (top - bottom);
  m(0, 2) = (right + left) / (right - left);
  m(1, 2) = (top + bottom) / (top - bottom);
  m(2, 2) = -(f + n) / (f - n);
  m(3, 2) = -1.f;
  m(2, 3) = -(2.f * f * n) / (f - n);
  m(3, 3) = 0.f;
  Mat4 T = identity(), Z = identity();
  T(0, 3) = c.xpan; T(1, 3) = c.ypan;
  Z(0, 0) = c.zoom; Z(1, 1) = c.zoom;
  return mul(Z, mul(T, m));
}

inline Mat4 projview(const vr_camera& c, int w, int h)
{
  return mul(projection_matrix(c, w, h), view_matrix(c));
}

// ---------------------------------------------------------------------------------------------
// vtkm::rendering::Camera (3-D mode) as Ascent's render parsing drives it
// (ascent_runtime_conduit_to_vtkm_parsing.cpp:97-173, Render.cpp:314-347): ResetToBounds, Azimuth,
// Elevation, Zoom, and the cinema orbit of ascent_runtime_rendering_filters.cpp:906-960.
inline Mat4 translation(float x, float y, float z)
{
  Mat4 m = identity();
  m(0, 3) = x; m(1, 3) = y; m(2, 3) = z;
  return m;
}
// Transform3DRotate(angleDegrees, axis): Rodrigues matrix about the normalised axis
inline Mat4 rotation(float deg, const Vec3& axis)
{
  const float ang = kPi180 * deg;
  const Vec3 n = normal(axis);
  const float s = std::sin(ang), c = std::cos(ang), ic = 1 - c;
  Mat4 m = identity();
  m(0, 0) = n[0] * n[0] * ic + c;        m(0, 1) = n[0] * n[1] * ic - n[2] * s; m(0, 2) = n[0] * n[2] * ic + n[1] * s;
  m(1, 0) = n[1] * n[0] * ic + n[2] * s; m(1, 1) = n[1] * n[1] * ic + c;        m(1, 2) = n[1] * n[2] * ic - n[0] * s;
  m(2, 0) = n[2] * n[0] * ic - n[1] * s; m(2, 1) = n[2] * n[1] * ic + n[0] * s; m(2, 2) = n[2] * n[2] * ic + c;
  return m;
}
inline void camera_default(vr_camera& c)
{
  const vr_camera d = { { 0.f, 0.f, 1.f }, { 0.f, 0.f, 0.f }, { 0.f, 1.f, 0.f }, 60.f, 1.f, 0.f, 0.f, 0.01f, 1000.f };
  c = d;
}
inline void camera_reset_to_bounds(vr_camera& c, const double b[6])
{
  const Vec3 dir = normal(sub(make3(c.position), make3(c.look_at)));
  const Vec3 centre = { { (float)((b[0] + b[1]) / 2.0), (float)((b[2] + b[3]) / 2.0), (float)((b[4] + b[5]) / 2.0) } };
  const Vec3 extent = { { (float)(b[1] - b[0]), (float)(b[3] - b[2]), (float)(b[5] - b[4]) } };
  const float diag = magnitude(extent);
  for (int k = 0; k < 3; ++k)
  {
    c.look_at[k] = centre[k];
    c.position[k] = centre[k] + dir[k] * diag * 1.0f;
  }
  c.fov = 60.f;
  c.near_plane = 0.1f * diag;
  c.far_plane = diag * 10.0f;
  c.xpan = c.ypan = 0.f;
  c.zoom = 1.f;
}
// the camera position turned about the look-at point: T(look_at) * R * T(-look_at), applied as a point
inline void camera_turn(vr_camera& c, float deg, const Vec3& axis)
{
  const Mat4 m = mul(mul(translation(c.look_at[0], c.look_at[1], c.look_at[2]), rotation(deg, axis)),
                     translation(-c.look_at[0], -c.look_at[1], -c.look_at[2]));
  const float p[4] = { c.position[0], c.position[1], c.position[2], 1.f };
  float q[4];
  mulv(m, p, q);
  c.position[0] = q[0]; c.position[1] = q[1]; c.position[2] = q[2];
}
inline void camera_azimuth(vr_camera& c, float deg) { camera_turn(c, deg, make3(c.up)); }
inline void camera_elevation(vr_camera& c, float deg)
{
  camera_turn(c, deg, cross(sub(make3(c.position), make3(c.look_at)), make3(c.up)));
}
inline void camera_zoom(vr_camera& c, float z) { c.zoom *= std::pow(4.0f, z); }
inline void camera_cinema(vr_camera& c, const double b[6], float phi, float theta)
{
  camera_default(c);
  camera_reset_to_bounds(c, b);
  const Vec3 extent = { { (float)(b[1] - b[0]), (float)(b[3] - b[2]), (float)(b[5] - b[4]) } };
  const float radius = (float)((double)magnitude(extent) * 2.5 / 2.0);
  const Mat4 rot = mul(rotation(phi, Vec3{ { 0.f, 0.f, 1.f } }), rotation(theta, Vec3{ { 1.f, 0.f, 0.f } }));
  const float up_h[4] = { 0.f, 1.f, 0.f, 0.f }, pos_h[4] = { 0.f, 0.f, 1.f, 1.f };
  float u[4], q[4];
  mulv(rot, up_h, u);
  mulv(rot, pos_h, q);
  const Vec3 un = normal(Vec3{ { u[0], u[1], u[2] } });
  for (int k = 0; k < 3; ++k)
  {
    c.up[k] = un[k];
    c.look_at[k] = (float)((b[2 * k] + b[2 * k + 1]) / 2.0);
    c.position[k] = q[k] * radius + c.look_at[k];
  }
}

struct RayGen
{
  float nlook[3], delta_x[3], delta_y[3];
};

// delta_y_from_ru reproduces partials_to_canvas' quirk (VolumeRenderer.cpp:328)
inline RayGen raygen(const vr_camera& c, int width, int height, bool delta_y_from_ru)
{
  float fov_y = c.fov, fov_x = c.fov;
  if (width != height)
  {
    const float vd = std::tan(0.5f * (fov_y * kPi180));
    const float hd = ((float)width / (float)height) * vd;
    fov_x = (2.0f * std::atan(hd)) / kPi180;
  }
  const Vec3 look = normal(sub(make3(c.look_at), make3(c.position)));
  const float thx = std::tan((fov_x * kPi180) * .5f);
  const float thy = std::tan((fov_y * kPi180) * .5f);
  const Vec3 ru = normal(cross(look, make3(c.up)));
  const Vec3 rv = normal(cross(ru, look));
  RayGen g;
  const float sx = 2 * thx / (float)width, sy = 2 * thy / (float)height;
  for (int k = 0; k < 3; ++k)
  {
    g.delta_x[k] = ru[k] * sx;
    g.delta_y[k] = (delta_y_from_ru ? ru[k] : rv[k]) * sy;
  }
  if (c.zoom > 0)
    for (int k = 0; k < 3; ++k) { g.delta_x[k] = g.delta_x[k] / c.zoom; g.delta_y[k] = g.delta_y[k] / c.zoom; }
  const Vec3 nl = delta_y_from_ru ? look : normal(look);
  for (int k = 0; k < 3; ++k) g.nlook[k] = nl[k];
  return g;
}

// screen-space subset of a bounding box: {minx, miny, w, h}
inline void find_subset(const vr_camera& c, int W, int H, const double b[6], int out[4])
{
  const float x[2] = { (float)b[0], (float)b[1] }, y[2] = { (float)b[2], (float)b[3] },
              z[2] = { (float)b[4], (float)b[5] };
  const float* P = c.position;
  if (P[0] >= x[0] && P[0] <= x[1] && P[1] >= y[0] && P[1] <= y[1] && P[2] >= z[0] && P[2] <= z[1])
  {
    out[0] = 0; out[1] = 0; out[2] = W; out[3] = H;
    return;
  }
  const Mat4 pv = projview(c, W, H);
  float xmin = INFINITY, ymin = INFINITY, zmin = INFINITY;
  float xmax = -INFINITY, ymax = -INFINITY, zmax = -INFINITY;
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j)
      for (int k = 0; k < 2; ++k)
      {
        const float e[4] = { x[i], y[j], z[k], 1.f };
        float t[4];
        mulv(pv, e, t);
        for (int a = 0; a < 3; ++a) t[a] = t[a] / t[3];
        t[0] = (t[0] * 0.5f + 0.5f) * (float)W;
        t[1] = (t[1] * 0.5f + 0.5f) * (float)H;
        t[2] = (t[2] * 0.5f + 0.5f);
        zmin = std::fmin(zmin, t[2]);
        zmax = std::fmax(zmax, t[2]);
        if (t[2] < 0 || t[2] > 1) continue;
        xmin = std::fmin(xmin, t[0]); ymin = std::fmin(ymin, t[1]);
        xmax = std::fmax(xmax, t[0]); ymax = std::fmax(ymax, t[1]);
      }
  xmin -= .001f; xmax += .001f; ymin -= .001f; ymax += .001f;
  xmin = std::floor(std::fmin(std::fmax(0.f, xmin), (float)W));
  xmax = std::ceil(std::fmin(std::fmax(0.f, xmax), (float)W));
  ymin = std::floor(std::fmin(std::fmax(0.f, ymin), (float)H));
  ymax = std::ceil(std::fmin(std::fmax(0.f, ymax), (float)H));
  if (zmax < 0 || xmin >= xmax || ymin >= ymax)
  {
    out[0] = 0; out[1] = 0; out[2] = 1; out[3] = 1;
  }
  else
  {
    out[0] = (int)xmin; out[1] = (int)ymin;
    out[2] = (int)xmax - (int)xmin; out[3] = (int)ymax - (int)ymin;
  }
}

} // namespace hm
} // namespace vr
