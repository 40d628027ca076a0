// vtkh_b200.cpp -- see vtkh_b200.hpp.  Host-side driver logic of the volume plot (the parts of
// vtkh::VolumeRenderer / Renderer / Scene that stay on the CPU) over the C ABI of libvr_b200.so.
#include "vtkh_b200.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
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
// State is the plain vr_camera; the arithmetic is the library's (vr_camera_*: the same Float32
// operation order as vtkm::rendering::Camera, shared with the Python harness).
Camera::Camera() { vr_camera_default(&m_c); }
void Camera::SetLookAt(const float v[3]) { std::memcpy(m_c.look_at, v, 12); }
void Camera::SetPosition(const float v[3]) { std::memcpy(m_c.position, v, 12); }
void Camera::SetViewUp(const float v[3]) { std::memcpy(m_c.up, v, 12); }
void Camera::ResetToBounds(const Bounds& b)
{
  double a[6];
  b.ToArray(a);
  vr_camera_reset_to_bounds(&m_c, a);
}
void Camera::Azimuth(float deg) { vr_camera_azimuth(&m_c, deg); }
void Camera::Elevation(float deg) { vr_camera_elevation(&m_c, deg); }
void Camera::Zoom(float z) { vr_camera_zoom(&m_c, z); }
void Camera::Pan(float dx, float dy) { m_c.xpan += dx; m_c.ypan += dy; }
Camera Camera::Cinema(const Bounds& b, float phi, float theta)
{
  Camera c;
  double a[6];
  b.ToArray(a);
  vr_camera_cinema(&c.m_c, a, phi, theta);
  return c;
}

// ------------------------------------------------------------------------------------------ colour table
namespace
{
std::string lower(std::string s)
{
  for (char& c : s) c = (char)std::tolower((unsigned char)c);
  return s;
}
} // namespace

ColorTable::ColorTable(const std::string& name)
{
  const std::string key = lower(name);
  m_alpha = { { 0.0, { 1.0f, 0, 0 } }, { 1.0, { 1.0f, 0, 0 } } };
  auto f = [](double v) { return (float)v; };
  if (key == "cool to warm")
  {
    m_space = 2;
    m_rgb = { { 0.0, { f(0.23137254902), f(0.298039215686), f(0.752941176471) } }, { 0.5, { f(0.865), f(0.865), f(0.865) } },
              { 1.0, { f(0.705882352941), f(0.0156862745098), f(0.149019607843) } } };
  }
  else if (key == "rainbow desaturated")
  {
    m_space = 0;
    m_rgb = { { 0.0, { f(0.278431372549), f(0.278431372549), f(0.858823529412) } }, { 0.143, { 0, 0, f(0.360784313725) } },
              { 0.285, { 0, 1, 1 } }, { 0.429, { 0, f(0.501960784314), 0 } }, { 0.571, { 1, 1, 0 } },
              { 0.714, { 1, f(0.380392156863), 0 } }, { 0.857, { f(0.419607843137), 0, 0 } },
              { 1.0, { f(0.878431372549), f(0.301960784314), f(0.301960784314) } } };
  }
  else if (key == "black-body radiation")
  {
    m_space = 0;
    m_rgb = { { 0.0, { 0, 0, 0 } }, { 0.4, { f(0.9), 0, 0 } }, { 0.8, { f(0.9), f(0.9), 0 } }, { 1.0, { 1, 1, 1 } } };
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
  for (int k = 0; k < 3; ++k) n.v[k] = std::min(1.0f, std::max(0.0f, rgb[k]));
  Insert(m_rgb, n);
}
void ColorTable::AddPointAlpha(double x, float alpha)
{
  Insert(m_alpha, Node{ x, { std::min(1.0f, std::max(0.0f, alpha)), 0, 0 } });
}
void ColorTable::ReverseColors()
{
  for (Node& n : m_rgb) n.x = 1.0 - n.x;
  std::stable_sort(m_rgb.begin(), m_rgb.end(), [](const Node& a, const Node& b) { return a.x < b.x; });
}
void ColorTable::Sample(int n, std::vector<uint8_t>& rgba8) const
{
  std::vector<double> cx, ax;
  std::vector<float> cc, av;
  for (const Node& p : m_rgb) { cx.push_back(p.x); cc.insert(cc.end(), p.v, p.v + 3); }
  for (const Node& p : m_alpha) { ax.push_back(p.x); av.push_back(p.v[0]); }
  rgba8.assign((size_t)n * 4, 0);
  if (vr_color_table_sample(m_space, (int)cx.size(), cx.data(), cc.data(), (int)ax.size(), ax.data(), av.data(), n,
                            rgba8.data(), nullptr) != VR_OK)
    throw Error("ColorTable::Sample: invalid table");
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
void DataSet::AddDomainUnstructured(int domain_id, size_t n_points, const void* xyz, int coord_dtype, size_t n_cells,
                                    int cell_shape, const void* connectivity, int index_bits)
{
  if (!xyz || !connectivity || n_points == 0 || n_cells == 0) throw Error("DataSet::AddDomainUnstructured: empty mesh");
  Domain dom;
  dom.id = domain_id;
  dom.kind = 2;
  for (int k = 0; k < 3; ++k) { dom.dims[k] = 0; dom.origin[k] = 0.f; dom.spacing[k] = 0.f; }
  dom.n_points = n_points; dom.n_cells = n_cells;
  dom.xyz = xyz; dom.conn = connectivity;
  dom.coord_dtype = coord_dtype; dom.cell_shape = cell_shape; dom.index_bits = index_bits;
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
  if (d.kind == 2)
  {
    // coords.GetBounds() of explicit points (f32, as the tracer sees them)
    float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (size_t i = 0; i < d.n_points * 3; ++i)
    {
      const float v = d.coord_dtype == VR_F32 ? static_cast<const float*>(d.xyz)[i] : (float)static_cast<const double*>(d.xyz)[i];
      lo[i % 3] = std::min(lo[i % 3], v);
      hi[i % 3] = std::max(hi[i % 3], v);
    }
    return Bounds(lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]);
  }
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
    const size_t n = d.kind == 2 ? (f->assoc == Points ? d.n_points : d.n_cells)
                     : f->assoc == Points ? (size_t)d.dims[0] * d.dims[1] * d.dims[2]
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
void Render::Save() const
{
  if (!m_png || m_png->empty()) throw Error("Render::Save: no PNG encoded (VolumeRenderer::SetEncodePNG)");
  const std::string path = m_name + ".png";
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw Error("Render::Save: cannot open " + path);
  const size_t n = std::fwrite(m_png->data(), 1, m_png->size(), f);
  std::fclose(f);
  if (n != m_png->size()) throw Error("Render::Save: short write to " + path);
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
  r.m_png = std::make_shared<std::vector<uint8_t>>();
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
  const float samples = (float)m_num_samples;
  m_corrected_color_table = m_color_table;
  for (int i = 0; i < m_corrected_color_table.GetNumberOfPointsAlpha(); ++i)
  {
    double x, a;
    m_corrected_color_table.GetPointAlpha(i, x, a);
    m_corrected_color_table.UpdatePointAlpha(i, x, vr_correct_opacity((float)a, samples));
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
    if (d.kind == 2)
    {
      m_has_unstructured = true;
      m_ctx->Check(vr_block_unstructured(m_ctx->h, d.id, d.n_points, d.xyz, d.coord_dtype, d.n_cells, d.cell_shape, d.conn,
                                         d.index_bits, f->data, f->dtype, f->assoc, VR_HOST));
    }
    else if (d.kind == 0)
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
  // ... and no rank holds an unstructured one (m_has_unstructured, :874-903 -> :470)
  int unstructured = m_has_unstructured ? 1 : 0;
  if (m_comm.size > 1)
  {
    std::vector<int> all((size_t)m_comm.size);
    m_comm.allgather(&unstructured, all.data(), sizeof(int));
    for (int v : all) unstructured = std::max(unstructured, v);
  }
  m_has_unstructured = unstructured != 0;
  m_used_path_a = one != 0 && !m_has_unstructured;
  if (m_used_path_a) RenderOneDomainPerRank();
  else RenderMultipleDomainsPerRank();
}

// A render whose host canvas is still the cleared canvas (Render::ClearCanvas) only needs the pixels inside the
// screen footprint of the GLOBAL bounds back: everything else the frame leaves as Canvas::Clear left it.
void VolumeRenderer::DownloadCanvas(Render& r, const vr_camera& cam, bool host_canvas_is_clear)
{
  const int W = r.GetWidth(), H = r.GetHeight();
  if (m_encode_png && m_comm.rank == 0)
  {
    // Render::RenderBackground + PNGEncoder::Encode on the device canvas: only the file crosses PCIe
    std::vector<uint8_t>& file = r.GetPNG();
    file.resize(vr_png_bound(W, H));
    size_t n = 0;
    m_ctx->Check(vr_canvas_encode_png(m_ctx->h, r.GetBackgroundColor(), file.data(), file.size(), &n));
    file.resize(n);
  }
  if (!host_canvas_is_clear)
  {
    m_ctx->Check(vr_canvas_download(m_ctx->h, r.GetColorBuffer().data(), r.GetDepthBuffer().data()));
    return;
  }
  double gb[6];
  m_bounds.ToArray(gb);
  int sub[4];
  vr_find_subset(&cam, W, H, gb, sub);
  const int x0 = sub[0] & ~3, x1 = std::min(W, (sub[0] + sub[2] + 3) & ~3);
  m_ctx->Check(vr_canvas_download_rect(m_ctx->h, x0, sub[1], x1, sub[1] + sub[3], r.GetColorBuffer().data(),
                                       r.GetDepthBuffer().data()));
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
    const bool was_clear = r.IsCleared();
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
    DownloadCanvas(r, cam, was_clear && m_do_composite);
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
    if (m_has_unstructured)
    {
      // a scene with an unstructured domain: every wrapper emits VolumePartials into ONE list
      // (UnstructuredWrapper / StructuredWrapper::render, :182-284), PartialCompositor folds it (:580-595)
      if (r.IsCleared()) m_ctx->Check(vr_canvas_clear(m_ctx->h, W, H));
      m_ctx->Check(vr_partials_begin(m_ctx->h, W, H));
      for (int i = 0; i < m_input->GetNumberOfDomains(); ++i)
      {
        const DataSet::Domain& d = m_input->GetDomain(i);
        if (d.Find(m_field_name)) m_ctx->Check(vr_trace_to_partials(m_ctx->h, d.id, &cam, m_sample_dist, rmin, rmax, 1));
      }
      m_ctx->Check(vr_partials_composite_to_canvas(m_ctx->h, &cam, r.IsCleared() ? 1 : 0));
      DownloadCanvas(r, cam, r.IsCleared());
      r.Touch();
      continue;
    }
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
    DownloadCanvas(r, cam, r.IsCleared());
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
