// t_vtkh_b200_volume_renderer.cpp -- C++ test driver over the host mirror (vtkh_b200.hpp).
//
// Reads like the reference's src/tests/vtkh/t_vtk-h_volume_renderer.cpp:79-123
// (vtkh_parallel_render: two uniform blocks of base size 32, "point_data_Float64", camera nudged and
// reset to the global bounds, "Cool to Warm" with alpha 0.01 -> 0.6, 512x512) but instead of only
// smoke-testing it dumps camera, range and canvas so that tests/test_host_cpp.py can compare the
// pixels with the CPU oracle.  No gtest in this image: modes are selected on the command line.
//   host   <out.bin>                      host-only classes (no GPU): camera, colour table, data set
//   render <num_blocks> <W> <H> <out.bin> the reference test body, any power-of-two block count
//   errors                                error behaviour (vtkh::Error equivalents)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <vector>

#include "vtkh_b200.hpp"

using namespace vtkh_b200;

struct Block
{
  int mins[3], maxs[3]; // cell index range, inclusive
};

// the same halving rule as the reference's test utility (t_vtkm_test_utils.hpp:14-105): cycle the
// dimensions, split every splittable division once per pass until there are enough blocks
static Block GetBlock(int block, int num_blocks, Block total)
{
  std::vector<Block> divs{ total };
  int dim = 0;
  while ((int)divs.size() < num_blocks)
  {
    const size_t cur = divs.size();
    for (size_t i = 0; i < cur && (int)divs.size() < num_blocks; ++i)
    {
      const int size = divs[i].maxs[dim] - divs[i].mins[dim] + 1;
      if (size <= 1) continue;
      Block right = divs[i];
      divs[i].maxs[dim] = divs[i].mins[dim] + size / 2 - 1;
      right.mins[dim] = divs[i].maxs[dim] + 1;
      divs.push_back(right);
    }
    dim = (dim + 1) % 3;
  }
  return divs.at(block);
}

struct TestDomain
{
  int dims[3];
  float origin[3], spacing[3];
  std::vector<double> point_data_Float64;
};

// CreateTestData(block, num_blocks, base_size) (t_vtkm_test_utils.hpp:196-252): unit spacing, origin
// at the block's first cell, point scalar = |point| + 1
static TestDomain CreateTestData(int block, int num_blocks, int base_size)
{
  Block total;
  for (int k = 0; k < 3; ++k) { total.mins[k] = 0; total.maxs[k] = num_blocks * base_size - 1; }
  const Block b = GetBlock(block, num_blocks, total);
  TestDomain d;
  for (int k = 0; k < 3; ++k)
  {
    d.origin[k] = (float)b.mins[k];
    d.spacing[k] = 1.f;
    d.dims[k] = b.maxs[k] - b.mins[k] + 2;
  }
  d.point_data_Float64.resize((size_t)d.dims[0] * d.dims[1] * d.dims[2]);
  size_t i = 0;
  for (int z = 0; z < d.dims[2]; ++z)
    for (int y = 0; y < d.dims[1]; ++y)
      for (int x = 0; x < d.dims[0]; ++x)
      {
        const double px = (double)(d.origin[0] + (float)x), py = (double)(d.origin[1] + (float)y),
                     pz = (double)(d.origin[2] + (float)z);
        d.point_data_Float64[i++] = std::sqrt(px * px + py * py + pz * pz) + 1.0;
      }
  return d;
}

static void put(FILE* f, const void* p, size_t n) { fwrite(p, 1, n, f); }

static int run_render(int num_blocks, int W, int H, const char* out)
{
  std::vector<TestDomain> storage(num_blocks);
  DataSet data_set;
  const int base_size = 32;
  for (int i = 0; i < num_blocks; ++i)
  {
    storage[i] = CreateTestData(i, num_blocks, base_size);
    data_set.AddDomainUniform(i, storage[i].dims, storage[i].origin, storage[i].spacing);
    data_set.AddField(i, "point_data_Float64", storage[i].point_data_Float64.data(), VR_F64, DataSet::Points);
  }
  Bounds bounds = data_set.GetGlobalBounds();

  Camera camera;
  float pos[3] = { camera.GetPosition()[0], camera.GetPosition()[1], camera.GetPosition()[2] };
  pos[0] += .1f;
  pos[1] += .1f;
  camera.SetPosition(pos);
  camera.ResetToBounds(bounds);
  Render render = MakeRender(W, H, camera, data_set, "volume");

  ColorTable color_map("Cool to Warm");
  color_map.AddPointAlpha(0.0, 0.01f);
  color_map.AddPointAlpha(1.0, 0.6f);

  VolumeRenderer tracer;
  tracer.SetColorTable(color_map);
  tracer.SetInput(&data_set);
  tracer.SetField("point_data_Float64");
  tracer.SetEncodePNG(true);
  const float bg[4] = { 0.2f, 0.3f, 0.4f, 1.f };
  render.SetBackgroundColor(bg);

  Scene scene;
  scene.AddRender(render);
  scene.AddRenderer(&tracer);
  scene.Render();

  const Render& done = scene.GetRenders()[0];
  {
    // Render::Save: the PNG encoded on the device, written next to the raw output
    FILE* pf = fopen((std::string(out) + ".png").c_str(), "wb");
    if (!pf) return 2;
    put(pf, done.GetPNG().data(), done.GetPNG().size());
    fclose(pf);
  }
  FILE* f = fopen(out, "wb");
  if (!f) return 2;
  const int hdr[4] = { W, H, num_blocks, tracer.UsedImagePath() ? 1 : 0 };
  put(f, hdr, sizeof(hdr));
  put(f, &camera.ToVR(), sizeof(vr_camera));
  const Range r = tracer.GetRange();
  const double rr[2] = { r.Min, r.Max };
  put(f, rr, sizeof(rr));
  const unsigned long long launches = tracer.KernelLaunches();
  put(f, &launches, sizeof(launches));
  put(f, done.GetColorBuffer().data(), done.GetColorBuffer().size() * 4);
  put(f, done.GetDepthBuffer().data(), done.GetDepthBuffer().size() * 4);
  fclose(f);
  std::cout << "render ok: " << num_blocks << " block(s), path " << (tracer.UsedImagePath() ? "A" : "B") << ", "
            << launches << " kernels launched\n";
  return 0;
}

// the same test body with the single domain handed over as an explicit cell set (hexahedra): the renderer must
// classify it as unstructured and take path B although there is one domain per rank (VolumeRenderer.cpp:874-903, :470)
static int run_render_unstructured(int W, int H, const char* out)
{
  TestDomain d = CreateTestData(0, 1, 16);
  const int nx = d.dims[0], ny = d.dims[1], nz = d.dims[2];
  std::vector<float> xyz((size_t)nx * ny * nz * 3);
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i)
      {
        const size_t p = ((size_t)k * ny + j) * nx + i;
        xyz[3 * p + 0] = d.origin[0] + d.spacing[0] * (float)i;
        xyz[3 * p + 1] = d.origin[1] + d.spacing[1] * (float)j;
        xyz[3 * p + 2] = d.origin[2] + d.spacing[2] * (float)k;
      }
  std::vector<long long> conn;
  for (int k = 0; k < nz - 1; ++k)
    for (int j = 0; j < ny - 1; ++j)
      for (int i = 0; i < nx - 1; ++i)
      {
        const long long i0 = ((long long)k * ny + j) * nx + i, up = (long long)nx * ny;
        const long long c[8] = { i0, i0 + 1, i0 + 1 + nx, i0 + nx, i0 + up, i0 + up + 1, i0 + up + 1 + nx, i0 + up + nx };
        conn.insert(conn.end(), c, c + 8);
      }
  DataSet data_set;
  data_set.AddDomainUnstructured(0, xyz.size() / 3, xyz.data(), VR_F32, conn.size() / 8, VR_HEXAHEDRON, conn.data(), 64);
  data_set.AddField(0, "point_data_Float64", d.point_data_Float64.data(), VR_F64, DataSet::Points);
  Bounds bounds = data_set.GetGlobalBounds();
  Camera camera;
  camera.ResetToBounds(bounds);
  camera.Azimuth(30.f);
  camera.Elevation(20.f);
  Render render = MakeRender(W, H, camera, data_set, "volume_unstructured");
  ColorTable color_map("Cool to Warm");
  color_map.AddPointAlpha(0.0, 0.01f);
  color_map.AddPointAlpha(1.0, 0.6f);
  VolumeRenderer tracer;
  tracer.SetColorTable(color_map);
  tracer.SetInput(&data_set);
  tracer.SetField("point_data_Float64");
  Scene scene;
  scene.AddRender(render);
  scene.AddRenderer(&tracer);
  scene.Render();
  const Render& done = scene.GetRenders()[0];
  FILE* f = fopen(out, "wb");
  if (!f) return 2;
  const int hdr[4] = { W, H, 1, tracer.UsedImagePath() ? 1 : 0 };
  put(f, hdr, sizeof(hdr));
  put(f, &camera.ToVR(), sizeof(vr_camera));
  const Range r = tracer.GetRange();
  const double rr[2] = { r.Min, r.Max };
  put(f, rr, sizeof(rr));
  const unsigned long long launches = tracer.KernelLaunches();
  put(f, &launches, sizeof(launches));
  put(f, done.GetColorBuffer().data(), done.GetColorBuffer().size() * 4);
  put(f, done.GetDepthBuffer().data(), done.GetDepthBuffer().size() * 4);
  fclose(f);
  std::cout << "render ok: unstructured, path " << (tracer.UsedImagePath() ? "A" : "B") << "\n";
  return 0;
}

static int run_host(const char* out)
{
  FILE* f = fopen(out, "wb");
  if (!f) return 2;
  // camera: defaults -> ResetToBounds -> Azimuth/Elevation/Zoom
  Camera c;
  Bounds b(-10, 10.5, -3, 7, 0, 31);
  c.ResetToBounds(b);
  c.Azimuth(45.f); c.Elevation(-10.f);
  c.Azimuth(10.f); c.Elevation(33.f);
  c.Zoom(0.5f);
  put(f, &c.ToVR(), sizeof(vr_camera));
  // colour tables: test table, default volume table (two alpha points at 0), rgb points
  std::vector<uint8_t> u8;
  ColorTable t1("Cool to Warm");
  t1.AddPointAlpha(0.0, 0.01f);
  t1.AddPointAlpha(1.0, 0.6f);
  t1.Sample(1024, u8);
  put(f, u8.data(), u8.size());
  ColorTable t2("Cool to Warm");
  t2.AddPointAlpha(0.0, .02f);
  t2.AddPointAlpha(.0, .5f);
  t2.Sample(1024, u8);
  put(f, u8.data(), u8.size());
  ColorTable t3("Cool to Warm");
  const float red[3] = { 1, 0, 0 }, green[3] = { 0, 1, 0 }, white[3] = { 1, 1, 1 };
  t3.AddPoint(0.0, red); t3.AddPoint(0.5, green); t3.AddPoint(1.0, white);
  t3.AddPointAlpha(0.0, 0.f); t3.AddPointAlpha(1.0, 1.f);
  t3.Sample(1024, u8);
  put(f, u8.data(), u8.size());
  // data set: bounds and range of the two-block test data
  std::vector<TestDomain> storage(2);
  DataSet ds;
  for (int i = 0; i < 2; ++i)
  {
    storage[i] = CreateTestData(i, 2, 32);
    ds.AddDomainUniform(i, storage[i].dims, storage[i].origin, storage[i].spacing);
    ds.AddField(i, "point_data_Float64", storage[i].point_data_Float64.data(), VR_F64, DataSet::Points);
  }
  double gb[6];
  ds.GetGlobalBounds().ToArray(gb);
  put(f, gb, sizeof(gb));
  const Range r = ds.GetGlobalRange("point_data_Float64");
  const double rr[2] = { r.Min, r.Max };
  put(f, rr, sizeof(rr));
  const int dims[6] = { storage[0].dims[0], storage[0].dims[1], storage[0].dims[2], storage[1].dims[0],
                        storage[1].dims[1], storage[1].dims[2] };
  put(f, dims, sizeof(dims));
  fclose(f);
  std::cout << "host ok\n";
  return 0;
}

template <typename F> static bool throws(F f)
{
  try { f(); }
  catch (const Error& e) { return true; }
  return false;
}

static int run_errors()
{
  int bad = 0;
  VolumeRenderer tracer;
  bad += !throws([&] { tracer.SetNumberOfSamples(0); });
  bad += !throws([&] { tracer.Update(); });                       // no input
  DataSet ds;
  TestDomain d = CreateTestData(0, 1, 8);
  ds.AddDomainUniform(0, d.dims, d.origin, d.spacing);
  ds.AddField(0, "point_data_Float64", d.point_data_Float64.data(), VR_F64, DataSet::Points);
  tracer.SetInput(&ds);
  bad += !throws([&] { tracer.Update(); });                       // no field set
  tracer.SetField("nope");
  bad += !throws([&] { tracer.Update(); });                       // unknown field
  Scene scene;
  VolumeRenderer second;
  scene.AddRenderer(&tracer);
  bad += !throws([&] { scene.AddRenderer(&second); });            // one volume per scene
  bad += !throws([&] { ColorTable t("no such table"); });
  Compositor comp;
  comp.SetCompositeMode(Compositor::VIS_ORDER_BLEND);
  bad += !throws([&] { comp.Composite(); });                      // no images
  std::vector<float> c(4 * 16, 0.f), z(16, 0.5f);
  comp.AddImage(c.data(), z.data(), 4, 4, 0);
  bad += !throws([&] { comp.SetCompositeMode(Compositor::Z_BUFFER_SURFACE); });
  bad += !throws([&] { comp.AddImage(c.data(), z.data(), 2, 8, 1); });
  std::cout << (bad ? "errors FAILED\n" : "errors ok\n");
  return bad;
}

int main(int argc, char** argv)
{
  try
  {
    if (argc >= 3 && !strcmp(argv[1], "host")) return run_host(argv[2]);
    if (argc >= 6 && !strcmp(argv[1], "render")) return run_render(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), argv[5]);
    if (argc >= 5 && !strcmp(argv[1], "render_unstructured")) return run_render_unstructured(atoi(argv[2]), atoi(argv[3]), argv[4]);
    if (argc >= 2 && !strcmp(argv[1], "errors")) return run_errors();
  }
  catch (const std::exception& e)
  {
    std::cerr << "exception: " << e.what() << "\n";
    return 3;
  }
  std::cerr << "usage: host <out> | render <blocks> <W> <H> <out> | render_unstructured <W> <H> <out> | errors\n";
  return 1;
}
