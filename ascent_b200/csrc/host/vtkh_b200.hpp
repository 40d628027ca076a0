// vtkh_b200.hpp -- C++ host side above the C ABI (include/vr_b200.h).
//
// Mirrors, for the `volume` plot path only, the interface a vtk-h user drives:
//   vtkh::DataSet            src/libs/vtkh/DataSet.hpp            (structured domains + fields)
//   vtkh::Render, MakeRender src/libs/vtkh/rendering/Render.hpp:22-114, Render.cpp:314-347
//   vtkh::VolumeRenderer     src/libs/vtkh/rendering/VolumeRenderer.hpp:15-55 / Renderer.hpp:19-81
//   vtkh::Scene              src/libs/vtkh/rendering/Scene.hpp
//   vtkh::Compositor         src/libs/vtkh/compositing/Compositor.hpp:14-87
//   vtkh::PartialCompositor  src/libs/vtkh/compositing/PartialCompositor.hpp
//   vtkh::VolumePartial      src/libs/vtkh/compositing/VolumePartial.hpp:48-119
// and the three VTK-m value types that cross that interface (vtkm::Bounds, vtkm::Range,
// vtkm::rendering::Camera, vtkm::cont::ColorTable -- only the members the reference's volume tests and
// Ascent's parse_camera / parse_color_table call).  Same method names, argument meaning and error
// behaviour (vtkh::Error thrown for what the reference throws for); every pixel is produced by the
// CUDA kernels behind libvr_b200.so -- there is no CPU rendering path in here.
//
// One process drives one GPU (one vr_ctx).  Multi-rank runs give the renderer a Comm (rank, size and
// an all-gather callback backed by MPI or torch.distributed); the pixels themselves never use it.
#pragma once
#include <cstdint>
#include <exception>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/vr_b200.h"

namespace vtkh_b200
{

class Error : public std::exception // vtkh::Error (src/libs/vtkh/Error.hpp)
{
public:
  explicit Error(const std::string& msg) : m_msg(msg) {}
  const char* what() const noexcept override { return m_msg.c_str(); }
private:
  std::string m_msg;
};

struct Range // vtkm::Range
{
  double Min = 1e300 * 1e10, Max = -1e300 * 1e10; // +inf / -inf: empty
  bool IsNonEmpty() const { return Min <= Max; }
  void Include(double v) { if (v < Min) Min = v; if (v > Max) Max = v; }
  double Length() const { return IsNonEmpty() ? Max - Min : 0.0; }
  double Center() const { return 0.5 * (Min + Max); }
};

struct Bounds // vtkm::Bounds
{
  Range X, Y, Z;
  Bounds() = default;
  Bounds(double x0, double x1, double y0, double y1, double z0, double z1)
  { X.Min = x0; X.Max = x1; Y.Min = y0; Y.Max = y1; Z.Min = z0; Z.Max = z1; }
  void Include(const Bounds& b)
  { X.Include(b.X.Min); X.Include(b.X.Max); Y.Include(b.Y.Min); Y.Include(b.Y.Max); Z.Include(b.Z.Min); Z.Include(b.Z.Max); }
  void ToArray(double out[6]) const { out[0] = X.Min; out[1] = X.Max; out[2] = Y.Min; out[3] = Y.Max; out[4] = Z.Min; out[5] = Z.Max; }
};

// vtkm::rendering::Camera, 3-D mode, as parse_camera drives it
// (ascent_runtime_conduit_to_vtkm_parsing.cpp:97-173).  f32 state like VTK-m's.
class Camera
{
public:
  Camera();
  void SetLookAt(const float v[3]);
  void SetPosition(const float v[3]);
  void SetViewUp(const float v[3]);
  void SetFieldOfView(float deg) { m_c.fov = deg; }
  void SetClippingRange(float nearp, float farp) { m_c.near_plane = nearp; m_c.far_plane = farp; }
  const float* GetLookAt() const { return m_c.look_at; }
  const float* GetPosition() const { return m_c.position; }
  const float* GetViewUp() const { return m_c.up; }
  float GetFieldOfView() const { return m_c.fov; }
  float GetZoom() const { return m_c.zoom; }
  void ResetToBounds(const Bounds& b);
  void Azimuth(float deg);
  void Elevation(float deg);
  void Zoom(float z);           // zoom *= 4^z
  void Pan(float dx, float dy);
  const vr_camera& ToVR() const { return m_c; }
  // one camera of the cinema orbit (CinemaManager::create_cinema_cameras, rendering_filters.cpp:906-960)
  static Camera Cinema(const Bounds& b, float phi_degrees, float theta_degrees);
private:
  vr_camera m_c;
};

// vtkm::cont::ColorTable: control points over [0,1]; presets by name; Sample(n) -> RGBA8
class ColorTable
{
public:
  explicit ColorTable(const std::string& name = "Cool to Warm");
  void AddPoint(double x, const float rgb[3]);
  void AddPointAlpha(double x, float alpha);
  void ClearColors() { m_rgb.clear(); }
  void ClearAlpha() { m_alpha.clear(); }
  void ReverseColors();
  int GetNumberOfPointsAlpha() const { return (int)m_alpha.size(); }
  void GetPointAlpha(int i, double& x, double& a) const { x = m_alpha[i].x; a = m_alpha[i].v[0]; }
  void UpdatePointAlpha(int i, double x, double a) { m_alpha[i].x = x; m_alpha[i].v[0] = (float)a; }
  void Sample(int n, std::vector<uint8_t>& rgba8) const; // ColorTable::Sample(n, Vec4ui_8)
private:
  struct Node { double x; float v[3]; }; // positions Float64, values Float32 (as VTK-m stores them)
  static void Insert(std::vector<Node>& pts, const Node& n);
  int m_space; // 0 rgb, 1 lab, 2 diverging
  std::vector<Node> m_rgb, m_alpha;
};

// vtkh::DataSet restricted to what SetInput accepts for the structured path (VolumeRenderer.cpp:868-907)
class DataSet
{
public:
  enum Assoc { Points = VR_POINT, Cells = VR_CELL };
  // AddDomain(vtkm::cont::DataSet, domain_id): the two coordinate kinds of the hot path
  void AddDomainUniform(int domain_id, const int point_dims[3], const float origin[3], const float spacing[3]);
  void AddDomainRectilinear(int domain_id, const int point_dims[3], const double* x, const double* y, const double* z);
  // an explicit cell set (vtkm::cont::CellSetSingleType of hexahedra or tetrahedra): caller-owned host arrays,
  // xyz interleaved; the volume renderer then takes the unstructured route for the whole scene (N4)
  void AddDomainUnstructured(int domain_id, size_t n_points, const void* xyz, int coord_dtype, size_t n_cells,
                             int cell_shape /* VR_HEXAHEDRON | VR_TETRA */, const void* connectivity, int index_bits);
  // data_set.AddField(...): caller-owned memory (zero copy on the host side; where = VR_HOST or VR_DEVICE)
  void AddField(int domain_index, const std::string& name, const void* data, int dtype, Assoc assoc, int where = VR_HOST);
  int GetNumberOfDomains() const { return (int)m_domains.size(); }
  bool IsEmpty() const { return m_domains.empty(); }
  Bounds GetDomainBounds(int domain_index) const;
  Bounds GetGlobalBounds() const;                        // local union (rank-global via Comm in the renderer)
  Range GetGlobalRange(const std::string& field) const;  // host fields only; throws for device fields
  bool GlobalFieldExists(const std::string& field) const;

  struct Field { std::string name; const void* data; int dtype; Assoc assoc; int where; };
  struct Domain
  {
    int id; int kind; int dims[3]; float origin[3], spacing[3]; // kind: 0 uniform, 1 rectilinear, 2 unstructured
    std::vector<double> ax[3];
    size_t n_points = 0, n_cells = 0;                           // kind 2
    const void* xyz = nullptr; const void* conn = nullptr;
    int coord_dtype = VR_F32, cell_shape = VR_HEXAHEDRON, index_bits = 32;
    std::vector<Field> fields;
    const Field* Find(const std::string& n) const;
  };
  const Domain& GetDomain(int i) const { return m_domains.at(i); }
private:
  std::vector<Domain> m_domains;
};

// vtkh::Render: camera + canvas (host float RGBA + depth, CanvasRayTracer's buffers)
class Render
{
public:
  int GetWidth() const { return m_width; }
  int GetHeight() const { return m_height; }
  const Camera& GetCamera() const { return m_camera; }
  void SetCamera(const Camera& c) { m_camera = c; }
  const std::string& GetImageName() const { return m_name; }
  std::vector<float>& GetColorBuffer() { return *m_rgba; }
  std::vector<float>& GetDepthBuffer() { return *m_depth; }
  const std::vector<float>& GetColorBuffer() const { return *m_rgba; }
  const std::vector<float>& GetDepthBuffer() const { return *m_depth; }
  void ClearCanvas(); // Canvas::Clear: colour 0, depth 1.001
  bool IsCleared() const { return m_cleared; }
  void Touch() { m_cleared = false; }
  // Render::RenderBackground + Render::Save (Render.cpp:277-312): the PNG the volume renderer encoded on the device
  // for this render (VolumeRenderer::SetEncodePNG), over the background colour; Save writes `<image name>.png`
  void SetBackgroundColor(const float bg[4]) { for (int k = 0; k < 4; ++k) m_bg[k] = bg[k]; }
  const float* GetBackgroundColor() const { return m_bg; }
  std::vector<uint8_t>& GetPNG() { return *m_png; }
  const std::vector<uint8_t>& GetPNG() const { return *m_png; }
  void Save() const; // throws if no PNG has been encoded
private:
  friend Render MakeRender(int, int, const Camera&, const DataSet&, const std::string&);
  int m_width = 0, m_height = 0;
  Camera m_camera;
  std::string m_name;
  std::shared_ptr<std::vector<float>> m_rgba, m_depth; // shared like vtkm ArrayHandles: copies alias
  std::shared_ptr<std::vector<uint8_t>> m_png;
  float m_bg[4] = { 0.f, 0.f, 0.f, 1.f };
  bool m_cleared = true;
};
Render MakeRender(int width, int height, const Camera& camera, const DataSet& data_set, const std::string& image_name);

// how ranks find each other (MPI_Comm in vtk-h, vtkh.cpp:58-72): O(ranks) scalars only
struct Comm
{
  int rank = 0, size = 1;
  // out = concatenation over ranks (rank order) of `bytes` bytes from `in`
  std::function<void(const void* in, void* out, size_t bytes)> allgather;
};

template <typename F> struct VolumePartial; // only float is used by the reference's volume path
template <> struct VolumePartial<float>
{
  int m_pixel_id; float m_depth; float m_pixel[3]; float m_alpha; // VolumePartial.hpp:48-56
};
static_assert(sizeof(VolumePartial<float>) == sizeof(vr_partial), "layout of VolumePartial<float>");

class Context; // owns the vr_ctx, shared by the objects below

struct Image // vtkh::Image as returned by Compositor::Composite
{
  int m_width = 0, m_height = 0;
  std::vector<unsigned char> m_pixels; // RGBA8
  std::vector<float> m_depths;
};

class Compositor
{
public:
  enum CompositeMode { Z_BUFFER_SURFACE, Z_BUFFER_BLEND, VIS_ORDER_BLEND }; // Compositor.hpp:14-18
  Compositor();
  ~Compositor();
  void SetCompositeMode(CompositeMode m);
  void ClearImages();
  void AddImage(const float* color_buffer, const float* depth_buffer, int width, int height);                // z-buffer modes
  void AddImage(const float* color_buffer, const float* depth_buffer, int width, int height, int vis_order); // VIS_ORDER_BLEND
  Image Composite();
private:
  std::shared_ptr<Context> m_ctx;
  CompositeMode m_mode = Z_BUFFER_SURFACE;
  int m_w = 0, m_h = 0;
  std::vector<float> m_rgba, m_depth;
  std::vector<int> m_order;
};

template <typename P> class PartialCompositor;
template <> class PartialCompositor<VolumePartial<float>>
{
public:
  PartialCompositor();
  ~PartialCompositor();
  // single rank: concatenate the per-domain vectors, order by (pixel, depth), fold front to back
  void composite(std::vector<std::vector<VolumePartial<float>>>& partial_images,
                 std::vector<VolumePartial<float>>& output_partials);
  void set_image_size(int width, int height) { m_w = width; m_h = height; } // optional: else from max pixel id
private:
  std::shared_ptr<Context> m_ctx;
  int m_w = 0, m_h = 0;
};

class VolumeRenderer
{
public:
  VolumeRenderer();
  virtual ~VolumeRenderer();
  std::string GetName() const { return "vtkh::VolumeRenderer"; }
  void SetNumberOfSamples(const int num_samples); // throws for <= 0 (VolumeRenderer.cpp:621-630)
  void SetColorTable(const ColorTable& color_table);
  void SetInput(DataSet* input);
  void SetField(const std::string& field_name) { m_field_name = field_name; }
  void SetRange(const Range& range) { m_range = range; }
  void SetDoComposite(bool do_composite) { m_do_composite = do_composite; }
  // Scene::Render's epilogue (RenderBackground + Save) on the device: each finished render also gets its PNG file
  void SetEncodePNG(bool on) { m_encode_png = on; }
  void AddRender(Render& render) { m_renders.push_back(render); }
  void SetRenders(const std::vector<Render>& renders) { m_renders = renders; }
  std::vector<Render> GetRenders() const { return m_renders; }
  int GetNumberOfRenders() const { return (int)m_renders.size(); }
  void ClearRenders() { m_renders.clear(); }
  Range GetRange() const { return m_range; }
  ColorTable GetColorTable() const { return m_color_table; }
  void SetComm(const Comm& comm) { m_comm = comm; }
  void Update();
  // evidence for tests/bench
  uint64_t KernelLaunches() const;
  bool UsedImagePath() const { return m_used_path_a; }
protected:
  void PreExecute();
  void DoExecute();
  void RenderOneDomainPerRank();
  void RenderMultipleDomainsPerRank();
  void CorrectOpacity();
  void UploadInput();
  void DownloadCanvas(Render& r, const vr_camera& cam, bool host_canvas_is_clear);
  bool m_has_unstructured = false; // SetInput's classification (VolumeRenderer.cpp:874-903)
  std::shared_ptr<Context> m_ctx;
  DataSet* m_input = nullptr;
  bool m_uploaded = false;
  std::string m_field_name;
  std::vector<Render> m_renders;
  Range m_range;
  Bounds m_bounds;
  ColorTable m_color_table, m_corrected_color_table;
  int m_num_samples = 100;
  float m_sample_dist = 0.f;
  bool m_do_composite = true;
  bool m_encode_png = false;
  bool m_used_path_a = false;
  Comm m_comm;
  bool m_comm_connected = false;
};

class Scene
{
public:
  void AddRender(Render& render) { m_renders.push_back(render); }
  void AddRenderer(VolumeRenderer* renderer);     // throws if a second volume is added (Scene.cpp:78-95)
  void Render();                                  // batches of <= 10 renders (Scene.cpp:124-247)
  std::vector<vtkh_b200::Render>& GetRenders() { return m_renders; }
private:
  std::vector<vtkh_b200::Render> m_renders;
  VolumeRenderer* m_volume = nullptr;
};

} // namespace vtkh_b200
