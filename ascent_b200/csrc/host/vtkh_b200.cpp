// vtkh_b200.cpp -- see vtkh_b200.hpp.  Host-side driver logic of the volume plot (the parts of
// vtkh::VolumeRenderer / Renderer / Scene that stay on the CPU) over the C ABI of libvr_b200.so.
#include "vtkh_b200.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace vtkh_b200
{

// ------------------------------------------------------------------------------------------ context
class Context
{
public:
  static std::shared_ptr<Context> Get()
  {
    static std::weak_ptr<Context> g;
    std::shared_ptr<Context> c = g.lock();
    if (!c)
    {
      c.reset(new Context());
      g = c;
    }
    return c;
  }
  ~Context() { vr_destroy(h); }
  void Check(vr_status s) const
  {
    if (s != VR_OK) throw Error(std::string("vr_b200: ") + vr_last_error(h));
  }
  vr_ctx* h = nullptr;
private:
  Context()
  {
    // one rank <-> one GPU (ascent_main_runtime.cpp:190-192 picks rank % device_count)
    const char* e = std::getenv("VTKH_B200_DEVICE");
    if (!e) e = std::getenv("LOCAL_RANK");
    const int dev = e ? std::atoi(e) : 0;
    if (vr_create(dev, &h) != VR_OK) throw Error(std::string("vr_b200: ") + vr_last_error(nullptr));
  }
};

// ------------------------------------------------------------------------------------------ camera
namespace
{
const float kPi180 = (float)0.01745329251994329547437168059786927;

void normalize3(float v[3])
{
  const float r = 1.0f / std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  v[0] *= r; v[1] *= r; v[2] *= r;
}
// Rodrigues rotation matrix (3x3, row major) about a unit axis
void rotation(float deg, const float axis_in[3], float m[9])
{
  float n[3] = { axis_in[0], axis_in[1], axis_in[2] };
  normalize3(n);
  const float a = kPi180 * deg;
  const float s = std::sin(a), c = std::cos(a), t = 1.0f - c;
  m[0] = n[0] * n[0] * t + c;        m[1] = n[0] * n[1] * t - n[2] * s; m[2] = n[0] * n[2] * t + n[1] * s;
  m[3] = n[1] * n[0] * t + n[2] * s; m[4] = n[1] * n[1] * t + c;        m[5] = n[1] * n[2] * t - n[0] * s;
  m[6] = n[2] * n[0] * t - n[1] * s; m[7] = n[2] * n[1] * t + n[0] * s; m[8] = n[2] * n[2] * t + c;
}
} // namespace

Camera::Camera()
{
  // vtkm::rendering::Camera defaults: look_at 0, position (0,0,1), up +y, fov 60
  std::memset(&m_c, 0, sizeof(m_c));
  m_c.position[2] = 1.f;
  m_c.up[1] = 1.f;
  m_c.fov = 60.f;
  m_c.zoom = 1.f;
  m_c.near_plane = 0.01f;
  m_c.far_plane = 1000.f;
}
void Camera::SetLookAt(const float v[3]) { std::memcpy(m_c.look_at, v, 12); }
void Camera::SetPosition(const float v[3]) { std::memcpy(m_c.position, v, 12); }
void Camera::SetViewUp(const float v[3]) { std::memcpy(m_c.up, v, 12); }

void Camera::ResetToBounds(const Bounds& b)
{
  // look at the centre from one diagonal away along the current view direction
  float d[3] = { m_c.position[0] - m_c.look_at[0], m_c.position[1] - m_c.look_at[1], m_c.position[2] - m_c.look_at[2] };
  normalize3(d);
  const float center[3] = { (float)b.X.Center(), (float)b.Y.Center(), (float)b.Z.Center() };
  const float ext[3] = { (float)b.X.Length(), (float)b.Y.Length(), (float)b.Z.Length() };
  const float diag = std::sqrt(ext[0] * ext[0] + ext[1] * ext[1] + ext[2] * ext[2]);
  for (int k = 0; k < 3; ++k)
  {
    m_c.look_at[k] = center[k];
    m_c.position[k] = center[k] + d[k] * diag;
  }
  m_c.fov = 60.f;
  m_c.near_plane = 0.1f * diag;
  m_c.far_plane = diag * 10.f;
  m_c.xpan = m_c.ypan = 0.f;
  m_c.zoom = 1.f;
}
void Camera::RotateAboutLookAt(float deg, const float axis[3])
{
  float m[9];
  rotation(deg, axis, m);
  const float p[3] = { m_c.position[0] - m_c.look_at[0], m_c.position[1] - m_c.look_at[1], m_c.position[2] - m_c.look_at[2] };
  for (int r = 0; r < 3; ++r)
    m_c.position[r] = m[3 * r] * p[0] + m[3 * r + 1] * p[1] + m[3 * r + 2] * p[2] + m_c.look_at[r];
}
void Camera::Azimuth(float deg) { RotateAboutLookAt(deg, m_c.up); }
void Camera::Elevation(float deg)
{
  const double p[3] = { (double)m_c.position[0] - m_c.look_at[0], (double)m_c.position[1] - m_c.look_at[1],
                        (double)m_c.position[2] - m_c.look_at[2] };
  const double u[3] = { m_c.up[0], m_c.up[1], m_c.up[2] };
  const float axis[3] = { (float)(p[1] * u[2] - p[2] * u[1]), (float)(p[2] * u[0] - p[0] * u[2]),
                          (float)(p[0] * u[1] - p[1] * u[0]) };
  RotateAboutLookAt(deg, axis);
}
void Camera::Zoom(float z) { m_c.zoom = m_c.zoom * (float)std::pow(4.0, (double)z); }
void Camera::Pan(float dx, float dy) { m_c.xpan += dx; m_c.ypan += dy; }

// ------------------------------------------------------------------------------------------ colour table
namespace
{
const double kRefX = 0.9505, kRefY = 1.000, kRefZ = 1.089, kPi = 3.14159265358979323846;

void rgb_to_lab(const double rgb[3], double lab[3])
{
  double c[3];
  for (int k = 0; k < 3; ++k) c[k] = rgb[k] > 0.04045 ? std::pow((rgb[k] + 0.055) / 1.055, 2.4) : rgb[k] / 12.92;
  const double x = c[0] * 0.4124 + c[1] * 0.3576 + c[2] * 0.1805;
  const double y = c[0] * 0.2126 + c[1] * 0.7152 + c[2] * 0.0722;
  const double z = c[0] * 0.0193 + c[1] * 0.1192 + c[2] * 0.9505;
  auto f = [](double t) { return t > 0.008856 ? std::cbrt(t) : 7.787 * t + 16.0 / 116.0; };
  const double fx = f(x / kRefX), fy = f(y / kRefY), fz = f(z / kRefZ);
  lab[0] = 116.0 * fy - 16.0; lab[1] = 500.0 * (fx - fy); lab[2] = 200.0 * (fy - fz);
}
void lab_to_rgb(const double lab[3], double rgb[3])
{
  const double vy = (lab[0] + 16.0) / 116.0, vx = lab[1] / 500.0 + vy, vz = vy - lab[2] / 200.0;
  auto finv = [](double v) { const double v3 = v * v * v; return v3 > 0.008856 ? v3 : (v - 16.0 / 116.0) / 7.787; };
  const double x = kRefX * finv(vx), y = kRefY * finv(vy), z = kRefZ * finv(vz);
  double c[3] = { x * 3.2406 + y * -1.5372 + z * -0.4986, x * -0.9689 + y * 1.8758 + z * 0.0415,
                  x * 0.0557 + y * -0.2040 + z * 1.0570 };
  double m = 0.0;
  for (int k = 0; k < 3; ++k)
  {
    c[k] = c[k] > 0.0031308 ? 1.055 * std::pow(c[k], 1.0 / 2.4) - 0.055 : 12.92 * c[k];
    m = std::max(m, c[k]);
  }
  for (int k = 0; k < 3; ++k) rgb[k] = std::max((m > 1.0 ? c[k] / m : c[k]), 0.0);
}
void lab_to_msh(const double lab[3], double msh[3])
{
  msh[0] = std::sqrt(lab[0] * lab[0] + lab[1] * lab[1] + lab[2] * lab[2]);
  msh[1] = msh[0] > 0.001 ? std::acos(lab[0] / msh[0]) : 0.0;
  msh[2] = msh[1] > 0.001 ? std::atan2(lab[2], lab[1]) : 0.0;
}
void msh_to_lab(const double msh[3], double lab[3])
{
  lab[0] = msh[0] * std::cos(msh[1]);
  lab[1] = msh[0] * std::sin(msh[1]) * std::cos(msh[2]);
  lab[2] = msh[0] * std::sin(msh[1]) * std::sin(msh[2]);
}
double angle_diff(double a1, double a2)
{
  double d = std::fabs(a1 - a2);
  while (d >= 2.0 * kPi) d -= 2.0 * kPi;
  return d > kPi ? 2.0 * kPi - d : d;
}
double adjust_hue(const double msh[3], double unsat_m)
{
  if (msh[0] >= unsat_m - 0.1) return msh[2];
  const double spin = msh[1] * std::sqrt(unsat_m * unsat_m - msh[0] * msh[0]) / (msh[0] * std::sin(msh[1]));
  return msh[2] > -0.3 * kPi ? msh[2] + spin : msh[2] - spin;
}
void interp_diverging(const double rgb1[3], const double rgb2[3], double w, double out[3])
{
  double lab[3], msh1[3], msh2[3];
  rgb_to_lab(rgb1, lab); lab_to_msh(lab, msh1);
  rgb_to_lab(rgb2, lab); lab_to_msh(lab, msh2);
  if (msh1[1] > 0.05 && msh2[1] > 0.05 && angle_diff(msh1[2], msh2[2]) > 0.33 * kPi)
  {
    const double mmid = std::max(88.0, std::max(msh1[0], msh2[0]));
    if (w < 0.5) { msh2[0] = mmid; msh2[1] = 0.0; msh2[2] = 0.0; w = 2.0 * w; }
    else { msh1[0] = mmid; msh1[1] = 0.0; msh1[2] = 0.0; w = 2.0 * w - 1.0; }
  }
  if (msh1[1] < 0.05 && msh2[1] > 0.05) msh1[2] = adjust_hue(msh2, msh1[0]);
  else if (msh2[1] < 0.05 && msh1[1] > 0.05) msh2[2] = adjust_hue(msh1, msh2[0]);
  double tmp[3];
  for (int k = 0; k < 3; ++k) tmp[k] = (1.0 - w) * msh1[k] + w * msh2[k];
  msh_to_lab(tmp, lab);
  lab_to_rgb(lab, out);
}
std::string lower(std::string s)
{
  for (char& c : s) c = (char)std::tolower((unsigned char)c);
  return s;
}
} // namespace

ColorTable::ColorTable(const std::string& name)
{
  const std::string key = lower(name);
  m_alpha = { { 0.0, { 1.0, 0, 0 } }, { 1.0, { 1.0, 0, 0 } } };
  if (key == "cool to warm")
  {
    m_space = 2;
    m_rgb = { { 0.0, { 0.23137254902, 0.298039215686, 0.752941176471 } }, { 0.5, { 0.865, 0.865, 0.865 } },
              { 1.0, { 0.705882352941, 0.0156862745098, 0.149019607843 } } };
  }
  else if (key == "black-body radiation")
  {
    m_space = 0;
    m_rgb = { { 0.0, { 0, 0, 0 } }, { 0.4, { 0.9, 0, 0 } }, { 0.8, { 0.9, 0.9, 0 } }, { 1.0, { 1, 1, 1 } } };
  }
  else if (key == "grayscale")
  {
    m_space = 0;
    m_rgb = { { 0.0, { 0, 0, 0 } }, { 1.0, { 1, 1, 1 } } };
  }
  else
    throw Error("unknown color table preset '" + name + "'");
}
void ColorTable::Insert(std::vector<Node>& pts, const Node& n)
{
  for (Node& p : pts)
    if (p.x == n.x) { p = n; return; } // a point at an existing position overwrites it
  pts.push_back(n);
  std::stable_sort(pts.begin(), pts.end(), [](const Node& a, const Node& b) { return a.x < b.x; });
}
void ColorTable::AddPoint(double x, const float rgb[3])
{
  Node n{ x, { 0, 0, 0 } };
  for (int k = 0; k < 3; ++k) n.v[k] = std::min(1.0, std::max(0.0, (double)rgb[k]));
  Insert(m_rgb, n);
}
void ColorTable::AddPointAlpha(double x, float alpha)
{
  Insert(m_alpha, Node{ x, { std::min(1.0, std::max(0.0, (double)alpha)), 0, 0 } });
}
void ColorTable::ReverseColors()
{
  for (Node& n : m_rgb) n.x = 1.0 - n.x;
  std::stable_sort(m_rgb.begin(), m_rgb.end(), [](const Node& a, const Node& b) { return a.x < b.x; });
}
void ColorTable::ColorAt(double x, double out[3]) const
{
  out[0] = out[1] = out[2] = 0.0;
  if (m_rgb.empty()) return;
  const Node* hit = nullptr;
  if (x <= m_rgb.front().x) hit = &m_rgb.front();
  else if (x >= m_rgb.back().x) hit = &m_rgb.back();
  if (hit) { std::memcpy(out, hit->v, sizeof(hit->v)); return; }
  for (size_t i = 0; i + 1 < m_rgb.size(); ++i)
  {
    const Node &a = m_rgb[i], &b = m_rgb[i + 1];
    if (x < a.x || x > b.x) continue;
    if (x == a.x) { std::memcpy(out, a.v, sizeof(a.v)); return; }
    if (x == b.x) { std::memcpy(out, b.v, sizeof(b.v)); return; }
    const double w = (double)(float)((x - a.x) / (b.x - a.x));
    if (m_space == 2) interp_diverging(a.v, b.v, w, out);
    else if (m_space == 1)
    {
      double l1[3], l2[3], l[3];
      rgb_to_lab(a.v, l1); rgb_to_lab(b.v, l2);
      for (int k = 0; k < 3; ++k) l[k] = (1.0 - w) * l1[k] + w * l2[k];
      lab_to_rgb(l, out);
    }
    else
      for (int k = 0; k < 3; ++k) out[k] = (1.0 - w) * a.v[k] + w * b.v[k];
    return;
  }
}
double ColorTable::AlphaAt(double x) const
{
  if (m_alpha.empty()) return 1.0;
  if (x <= m_alpha.front().x) return m_alpha.front().v[0];
  if (x >= m_alpha.back().x) return m_alpha.back().v[0];
  for (size_t i = 0; i + 1 < m_alpha.size(); ++i)
  {
    const Node &a = m_alpha[i], &b = m_alpha[i + 1];
    if (x < a.x || x > b.x) continue;
    const double w = (x - a.x) / (b.x - a.x);
    return (1.0 - w) * a.v[0] + w * b.v[0];
  }
  return m_alpha.back().v[0];
}
void ColorTable::Sample(int n, std::vector<uint8_t>& rgba8) const
{
  rgba8.assign((size_t)n * 4, 0);
  const float delta = 1.0f / (float)(n - 1);
  for (int i = 0; i < n; ++i)
  {
    const double x = (i == n - 1) ? 1.0 : (double)(0.0f + delta * (float)i);
    double c[4];
    ColorAt(x, c);
    c[3] = AlphaAt(x);
    for (int k = 0; k < 4; ++k) rgba8[(size_t)i * 4 + k] = (uint8_t)(int)((float)c[k] * 255.0f + 0.5f);
  }
}

// ------------------------------------------------------------------------------------------ data set
void DataSet::AddDomainUniform(int domain_id, const int d[3], const float origin[3], const float spacing[3])
{
  Domain dom;
  dom.id = domain_id;
  dom.kind = 0;
  for (int k = 0; k < 3; ++k) { dom.dims[k] = d[k]; dom.origin[k] = origin[k]; dom.spacing[k] = spacing[k]; }
  m_domains.push_back(dom);
}
void DataSet::AddDomainRectilinear(int domain_id, const int d[3], const double* x, const double* y, const double* z)
{
  Domain dom;
  dom.id = domain_id;
  dom.kind = 1;
  const double* ax[3] = { x, y, z };
  for (int k = 0; k < 3; ++k)
  {
    dom.dims[k] = d[k];
    dom.origin[k] = 0.f; dom.spacing[k] = 0.f;
    dom.ax[k].assign(ax[k], ax[k] + d[k]);
  }
  m_domains.push_back(dom);
}
void DataSet::AddField(int i, const std::string& name, const void* data, int dtype, Assoc assoc, int where)
{
  if (i < 0 || i >= (int)m_domains.size()) throw Error("DataSet::AddField: no such domain");
  m_domains[i].fields.push_back(Field{ name, data, dtype, assoc, where });
}
const DataSet::Field* DataSet::Domain::Find(const std::string& n) const
{
  for (const Field& f : fields)
    if (f.name == n) return &f;
  return nullptr;
}
Bounds DataSet::GetDomainBounds(int i) const
{
  const Domain& d = m_domains.at(i);
  double b[6];
  for (int k = 0; k < 3; ++k)
  {
    if (d.kind == 0)
    {
      b[2 * k] = (double)d.origin[k];
      b[2 * k + 1] = (double)d.origin[k] + (double)d.spacing[k] * (double)(d.dims[k] - 1);
    }
    else
    {
      b[2 * k] = d.ax[k].front();
      b[2 * k + 1] = d.ax[k].back();
    }
  }
  return Bounds(b[0], b[1], b[2], b[3], b[4], b[5]);
}
Bounds DataSet::GetGlobalBounds() const
{
  Bounds b;
  for (int i = 0; i < (int)m_domains.size(); ++i) b.Include(GetDomainBounds(i));
  return b;
}
bool DataSet::GlobalFieldExists(const std::string& field) const
{
  for (const Domain& d : m_domains)
    if (d.Find(field)) return true;
  return false;
}
Range DataSet::GetGlobalRange(const std::string& field) const
{
  Range r;
  for (const Domain& d : m_domains)
  {
    const Field* f = d.Find(field);
    if (!f) continue;
    if (f->where != VR_HOST) throw Error("GetGlobalRange: field '" + field + "' lives on the device; call SetRange");
    const size_t n = f->assoc == Points ? (size_t)d.dims[0] * d.dims[1] * d.dims[2]
                                        : (size_t)(d.dims[0] - 1) * (d.dims[1] - 1) * (d.dims[2] - 1);
    if (f->dtype == VR_F64)
      for (size_t i = 0; i < n; ++i) r.Include(static_cast<const double*>(f->data)[i]);
    else
      for (size_t i = 0; i < n; ++i) r.Include((double)static_cast<const float*>(f->data)[i]);
  }
  return r;
}

// ------------------------------------------------------------------------------------------ render
void Render::ClearCanvas()
{
  std::fill(m_rgba->begin(), m_rgba->end(), 0.f);
  std::fill(m_depth->begin(), m_depth->end(), 1.001f);
  m_cleared = true;
}
Render MakeRender(int width, int height, const Camera& camera, const DataSet&, const std::string& image_name)
{
  if (width <= 0 || height <= 0) throw Error("MakeRender: bad image size");
  Render r;
  r.m_width = width;
  r.m_height = height;
  r.m_camera = camera;
  r.m_name = image_name;
  r.m_rgba = std::make_shared<std::vector<float>>((size_t)width * height * 4, 0.f);
  r.m_depth = std::make_shared<std::vector<float>>((size_t)width * height, 1.001f);
  return r;
}

// ------------------------------------------------------------------------------------------ compositor
Compositor::Compositor() : m_ctx(Context::Get()) {}
Compositor::~Compositor() {}
void Compositor::SetCompositeMode(CompositeMode m)
{
  if (!m_order.empty() || !m_depth.empty()) throw Error("Cannot change composite mode with images already added");
  m_mode = m;
}
void Compositor::ClearImages() { m_rgba.clear(); m_depth.clear(); m_order.clear(); m_w = m_h = 0; }
void Compositor::AddImage(const float* c, const float* d, int width, int height)
{
  if (m_mode == VIS_ORDER_BLEND) throw Error("AddImage: VIS_ORDER_BLEND needs a visibility order");
  AddImage(c, d, width, height, (int)m_order.size());
}
void Compositor::AddImage(const float* c, const float* d, int width, int height, int vis_order)
{
  if (!m_order.empty() && (width != m_w || height != m_h)) throw Error("AddImage: image sizes differ");
  m_w = width; m_h = height;
  const size_t n = (size_t)width * height;
  m_rgba.insert(m_rgba.end(), c, c + n * 4);
  m_depth.insert(m_depth.end(), d, d + n);
  m_order.push_back(vis_order);
}
Image Compositor::Composite()
{
  if (m_order.empty()) throw Error("Composite: no images");
  Image out;
  out.m_width = m_w; out.m_height = m_h;
  out.m_pixels.resize((size_t)m_w * m_h * 4);
  out.m_depths.resize((size_t)m_w * m_h);
  if (m_mode == VIS_ORDER_BLEND)
    m_ctx->Check(vr_composite_images(m_ctx->h, m_rgba.data(), m_depth.data(), m_order.data(), (int)m_order.size(),
                                     m_w, m_h, out.m_pixels.data(), out.m_depths.data()));
  else if (m_mode == Z_BUFFER_SURFACE)
    m_ctx->Check(vr_composite_zbuffer(m_ctx->h, m_rgba.data(), m_depth.data(), (int)m_order.size(), m_w, m_h,
                                      out.m_pixels.data(), out.m_depths.data()));
  else
    throw Error("CompositeZBufferBlend: not implemented"); // Compositor.cpp:215-218 asserts the same
  return out;
}

PartialCompositor<VolumePartial<float>>::PartialCompositor() : m_ctx(Context::Get()) {}
PartialCompositor<VolumePartial<float>>::~PartialCompositor() {}
void PartialCompositor<VolumePartial<float>>::composite(std::vector<std::vector<VolumePartial<float>>>& in,
                                                        std::vector<VolumePartial<float>>& out)
{
  std::vector<VolumePartial<float>> all;
  int max_px = -1;
  for (auto& v : in)
    for (auto& p : v)
    {
      all.push_back(p);
      max_px = std::max(max_px, p.m_pixel_id);
    }
  out.clear();
  if (all.empty()) return; // PartialCompositor.cpp:514-518
  int w = m_w, h = m_h;
  if (w <= 0 || h <= 0) { w = 4096; h = max_px / 4096 + 1; }
  out.resize(all.size());
  size_t n = 0;
  m_ctx->Check(vr_composite_partials(m_ctx->h, reinterpret_cast<const vr_partial*>(all.data()), all.size(), w, h,
                                     reinterpret_cast<vr_partial*>(out.data()), &n));
  out.resize(n);
}

// ------------------------------------------------------------------------------------------ volume renderer
VolumeRenderer::VolumeRenderer() : m_ctx(Context::Get())
{
  // VolumeRenderer.cpp:395-408: Cool to Warm with BOTH alpha points added at position 0
  m_color_table.AddPointAlpha(0.0, .02f);
  m_color_table.AddPointAlpha(.0, .5f);
  m_num_samples = 100;
}
VolumeRenderer::~VolumeRenderer()
{
  if (m_input && m_uploaded)
    for (int i = 0; i < m_input->GetNumberOfDomains(); ++i) vr_block_free(m_ctx->h, m_input->GetDomain(i).id);
}
uint64_t VolumeRenderer::KernelLaunches() const { return vr_kernel_launches(m_ctx->h); }

void VolumeRenderer::SetNumberOfSamples(const int num_samples)
{
  if (num_samples < 1) throw Error("Volume rendering samples must be at least 1: " + std::to_string(num_samples));
  m_num_samples = num_samples;
}
void VolumeRenderer::SetColorTable(const ColorTable& color_table) { m_color_table = color_table; }
void VolumeRenderer::SetInput(DataSet* input)
{
  m_input = input;
  m_uploaded = false;
}
void VolumeRenderer::CorrectOpacity()
{
  // alpha' = 1 - (1 - alpha)^(10/samples) on every alpha control point (VolumeRenderer.cpp:448-466)
  const float correction_scalar = 10.f;
  const float ratio = correction_scalar / (float)m_num_samples;
  m_corrected_color_table = m_color_table;
  for (int i = 0; i < m_corrected_color_table.GetNumberOfPointsAlpha(); ++i)
  {
    double x, a;
    m_corrected_color_table.GetPointAlpha(i, x, a);
    m_corrected_color_table.UpdatePointAlpha(i, x, 1.0 - std::pow(1.0 - a, (double)ratio));
  }
}
void VolumeRenderer::UploadInput()
{
  if (m_uploaded) return;
  for (int i = 0; i < m_input->GetNumberOfDomains(); ++i)
  {
    const DataSet::Domain& d = m_input->GetDomain(i);
    const DataSet::Field* f = d.Find(m_field_name);
    if (!f) continue;
    if (d.kind == 0)
      m_ctx->Check(vr_block_uniform(m_ctx->h, d.id, d.dims, d.origin, d.spacing, f->data, f->dtype, f->assoc, f->where));
    else
      m_ctx->Check(vr_block_rectilinear(m_ctx->h, d.id, d.dims, d.ax[0].data(), d.ax[1].data(), d.ax[2].data(), f->data,
                                        f->dtype, f->assoc, f->where));
  }
  m_uploaded = true;
}

void VolumeRenderer::PreExecute()
{
  if (!m_input) throw Error("VolumeRenderer: no input");
  if (m_field_name.empty()) throw Error("VolumeRenderer: no field set");
  // Renderer::PreExecute (Renderer.cpp:144-179): range = user's or the global field range
  Range local = m_range;
  if (!local.IsNonEmpty())
  {
    local = m_input->GetGlobalRange(m_field_name);
    if (m_comm.size > 1)
    {
      const double mine[2] = { local.Min, local.Max };
      std::vector<double> all(2 * (size_t)m_comm.size);
      m_comm.allgather(mine, all.data(), sizeof(mine));
      for (int r = 0; r < m_comm.size; ++r) { local.Include(all[2 * r]); local.Include(all[2 * r + 1]); }
    }
    if (!local.IsNonEmpty()) throw Error("VolumeRenderer: field '" + m_field_name + "' does not exist");
    m_range = local;
  }
  Bounds b = m_input->GetGlobalBounds();
  if (m_comm.size > 1)
  {
    double mine[6];
    b.ToArray(mine);
    std::vector<double> all(6 * (size_t)m_comm.size);
    m_comm.allgather(mine, all.data(), sizeof(mine));
    for (int r = 0; r < m_comm.size; ++r)
      b.Include(Bounds(all[6 * r], all[6 * r + 1], all[6 * r + 2], all[6 * r + 3], all[6 * r + 4], all[6 * r + 5]));
  }
  m_bounds = b;
  // VolumeRenderer::PreExecute (:599-612)
  CorrectOpacity();
  double gb[6];
  m_bounds.ToArray(gb);
  m_sample_dist = vr_sample_distance(gb, (float)m_num_samples);
  UploadInput();
  // Mapper::SetActiveColorTable / convert_table (:64-91): Sample(1024) -> uint8 -> * 1/255
  std::vector<uint8_t> u8;
  m_corrected_color_table.Sample(1024, u8);
  std::vector<float> lut(u8.size());
  const float inv = 1.f / 255.f;
  for (size_t i = 0; i < u8.size(); ++i) lut[i] = (float)u8[i] * inv;
  m_ctx->Check(vr_set_tf(m_ctx->h, lut.data(), 1024));
}

void VolumeRenderer::DoExecute()
{
  // VolumeRenderer.cpp:468-480: path A only when every rank holds exactly one (structured) domain
  int one = m_input->GetNumberOfDomains() == 1 ? 1 : 0;
  if (m_comm.size > 1)
  {
    std::vector<int> all((size_t)m_comm.size);
    m_comm.allgather(&one, all.data(), sizeof(int));
    for (int v : all) one = std::min(one, v);
  }
  m_used_path_a = one != 0;
  if (m_used_path_a) RenderOneDomainPerRank();
  else RenderMultipleDomainsPerRank();
}

void VolumeRenderer::RenderOneDomainPerRank()
{
  const DataSet::Domain& d = m_input->GetDomain(0);
  if (!d.Find(m_field_name)) return;
  if (m_comm.size > 1 && !m_comm_connected)
    throw Error("VolumeRenderer: multi-rank compositing needs ConnectComm (see tests/multi_rank_worker.py for the "
                "torch.distributed harness); this C++ mirror drives one rank");
  const float rmin = (float)m_range.Min, rmax = (float)m_range.Max;
  for (Render& r : m_renders)
  {
    const int W = r.GetWidth(), H = r.GetHeight();
    const vr_camera cam = r.GetCamera().ToVR();
    if (r.IsCleared() && m_do_composite)
    {
      // Canvas::Clear + RenderCells + Image::Init + ImageToCanvas: one launch
      m_ctx->Check(vr_trace_to_image(m_ctx->h, d.id, &cam, W, H, m_sample_dist, rmin, rmax, VR_FRAME_WRITE_CANVAS));
    }
    else
    {
      m_ctx->Check(vr_canvas_upload(m_ctx->h, W, H, r.GetColorBuffer().data(), r.GetDepthBuffer().data()));
      m_ctx->Check(vr_trace_to_canvas(m_ctx->h, d.id, &cam, m_sample_dist, rmin, rmax, 1));
      if (m_do_composite)
      {
        // Composite() with one image: Image::Init, no blend, ImageToCanvas (VolumeRenderer.cpp:652-688)
        void *rgba8 = nullptr, *depth = nullptr;
        m_ctx->Check(vr_image_from_canvas(m_ctx->h));
        m_ctx->Check(vr_image_ptrs(m_ctx->h, &rgba8, &depth));
        m_ctx->Check(vr_image_to_canvas_dev(m_ctx->h, static_cast<const uint8_t*>(rgba8), static_cast<const float*>(depth)));
      }
    }
    m_ctx->Check(vr_canvas_download(m_ctx->h, r.GetColorBuffer().data(), r.GetDepthBuffer().data()));
    r.Touch();
  }
}

void VolumeRenderer::RenderMultipleDomainsPerRank()
{
  if (m_comm.size > 1 && !m_comm_connected)
    throw Error("VolumeRenderer: multi-rank compositing needs ConnectComm; this C++ mirror drives one rank");
  const float rmin = (float)m_range.Min, rmax = (float)m_range.Max;
  for (Render& r : m_renders)
  {
    const int W = r.GetWidth(), H = r.GetHeight();
    const vr_camera cam = r.GetCamera().ToVR();
    if (!r.IsCleared())
      m_ctx->Check(vr_canvas_upload(m_ctx->h, W, H, r.GetColorBuffer().data(), r.GetDepthBuffer().data()));
    // wrapper->render(camera, canvas, partials) per domain (VolumeRenderer.cpp:561-577): the rays of
    // each structured block stay in a dense layer on the device
    m_ctx->Check(vr_layers_begin(m_ctx->h, W, H));
    std::vector<int> ids;
    for (int i = 0; i < m_input->GetNumberOfDomains(); ++i)
    {
      const DataSet::Domain& d = m_input->GetDomain(i);
      if (d.Find(m_field_name)) ids.push_back(d.id);
    }
    // the whole per-domain loop in one call: consecutive blocks overlap on the GPU
    m_ctx->Check(vr_trace_blocks_to_layers(m_ctx->h, (int)ids.size(), ids.data(), &cam, m_sample_dist, rmin, rmax,
                                           r.IsCleared() ? 0 : 1));
    // PartialCompositor::composite + partials_to_canvas (VolumeRenderer.cpp:580-595), one kernel
    m_ctx->Check(vr_layers_composite_to_canvas(m_ctx->h, &cam, r.IsCleared() ? 1 : 0));
    m_ctx->Check(vr_canvas_download(m_ctx->h, r.GetColorBuffer().data(), r.GetDepthBuffer().data()));
    r.Touch();
  }
}

void VolumeRenderer::Update()
{
  PreExecute();
  DoExecute();
}

// ------------------------------------------------------------------------------------------ scene
void Scene::AddRenderer(VolumeRenderer* renderer)
{
  if (m_volume) throw Error("Scenes only support a single volume plot"); // Scene.cpp:85-88
  m_volume = renderer;
}
void Scene::Render()
{
  if (!m_volume) return;
  // Scene.cpp:124-247: renders are processed in batches of 10; the volume goes last with
  // compositing on.  Results land in the renders' canvases.
  const int batch_size = 10;
  for (size_t begin = 0; begin < m_renders.size(); begin += batch_size)
  {
    const size_t end = std::min(m_renders.size(), begin + batch_size);
    std::vector<vtkh_b200::Render> batch(m_renders.begin() + begin, m_renders.begin() + end);
    m_volume->SetDoComposite(true);
    m_volume->SetRenders(batch);
    m_volume->Update();
    batch = m_volume->GetRenders();
    m_volume->ClearRenders();
    for (size_t i = begin; i < end; ++i) m_renders[i] = batch[i - begin];
  }
}

} // namespace vtkh_b200
