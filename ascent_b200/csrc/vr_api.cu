// vr_api.cu -- the C ABI of libvr_b200.so (include/vr_b200.h): context, block upload, frame
// buffers and the host-side driver logic around the kernels in sampler.cu / composite.cu /
// comm.cu.  No CPU fallback: every compute entry point launches CUDA kernels or fails.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cuda.h> // CUtensorMap types only: the encoder is looked up through the runtime, libcuda is not linked

#include "vr_color_table.hpp"
#include "vr_host_math.hpp"
#include "vr_internal.h"
#include "vr_radixk.hpp"
#include "vr_umesh_faces.hpp"

using namespace vr;

static std::string g_create_error;

static vr_status fail(vr_ctx* c, vr_status st, const char* fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  else g_create_error = buf;
  return st;
}

#define CK(call)                                                                                   \
  do                                                                                               \
  {                                                                                                \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? VR_ERR_NOMEM : VR_ERR_CUDA, "%s: %s",     \
                  #call, cudaGetErrorString(e_));                                                  \
  } while (0)

#define REQUIRE(cond, ...)                                                                         \
  do                                                                                               \
  {                                                                                                \
    if (!(cond)) return fail(ctx, VR_ERR_INVALID, __VA_ARGS__);                                    \
  } while (0)

static vr_status ensure_frame(vr_ctx* ctx, int W, int H);
namespace vr { vr_status comm_ahead_image(vr_ctx* ctx, uchar4** rgba, float** depth); } // comm.cu
namespace vr { vr_status comm_push_target(vr_ctx* ctx, bool ahead, int width, int height, vr::TraceParams& p); }
namespace vr { vr_status comm_check_errors(vr_ctx* ctx); }
namespace vr { void comm_join_previous_exchange(vr_ctx* ctx); }
namespace vr { void comm_join_for_image_trace(vr_ctx* ctx, bool ahead, cudaStream_t s); }
namespace vr { unsigned long long* comm_timeline_slot(vr_ctx* ctx, int k); }
static void fill_to_canvas_params(const vr_camera* cam, int W, int H, ToCanvasParams& tp);

// ================================================================= context
static void free_multi(vr_ctx* ctx);

extern "C" vr_status vr_create(int device, vr_ctx** out)
{
  vr_ctx* ctx = nullptr;
  if (!out) return fail(nullptr, VR_ERR_INVALID, "vr_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, VR_ERR_CUDA, "vr_create: no CUDA device (%s); this library has no CPU path",
                cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(nullptr, VR_ERR_INVALID, "vr_create: bad device %d", device);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return fail(nullptr, VR_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, VR_ERR_CUDA, "vr_create: device %d is sm_%d%d; built for sm_100a only", device,
                prop.major, prop.minor);
  if ((e = cudaSetDevice(device)) != cudaSuccess)
    return fail(nullptr, VR_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaGetLastError(); // a stale error of some earlier, unrelated runtime call must not fail this one
  ctx = new vr_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess)
  {
    delete ctx;
    return fail(nullptr, VR_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
  }
  ctx->stream = ctx->own_stream;
  preload_comm_kernels();
  preload_layers_kernels();
  preload_composite_kernels();
  preload_stage_kernels();
  preload_png_kernels();
  preload_unstructured_kernels();
  cudaGetLastError();
  if (const char* e = std::getenv("VR_CTAS_PER_SM")) ctx->ctas_per_sm = std::atoi(e); // tuning knob
  if (const char* e = std::getenv("VR_TILE_ORDER")) ctx->tile_order = std::atoi(e);
  if (const char* e = std::getenv("VR_COUNT_SAMPLES")) ctx->count_samples = std::atoi(e) != 0;
  if (const char* e = std::getenv("VR_NO_SPARSE")) ctx->no_sparse = std::atoi(e) != 0; // A/B knobs: general march only
  // The brick march (TMA-staged bricks) is bit-identical to the general march but measured SLOWER on B200 at
  // the dense operating point (0.75 vs 0.54 ms on c2 at samples = 887: the per-sample arithmetic, not the
  // gather latency, bounds that regime -- DESIGN.md section 4.1), so it is opt-in: VR_BRICK=1.
  ctx->no_brick = true;
  if (const char* e = std::getenv("VR_BRICK")) ctx->no_brick = std::atoi(e) == 0;
  // [0]: single launches, [1 + k]: launch k of a batched call, [1 + kMaxLayers + s]: image trace on side stream s
  cudaMalloc(&ctx->tile_counter, (size_t)(1 + vr::kMaxLayers + vr::Comm::kMaxTraceStreams) * sizeof(unsigned int));
  {
    // the exchange stream (fold of frame k overlaps the trace of frame k+1; its CTAs go first when SM slots
    // free up) and the two side streams of image-only traces -- see vr_internal.h
    vr::Comm& c = ctx->comm;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    cudaStreamCreateWithPriority(&c.xstream, cudaStreamNonBlocking, hi);
    cudaEventCreateWithFlags(&c.ev_trace, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c.ev_main, cudaEventDisableTiming);
    for (int k = 0; k < 8; ++k) cudaEventCreateWithFlags(&c.ev_x[k], cudaEventDisableTiming);
    for (int k = 0; k < vr::Comm::kMaxTraceStreams; ++k)
    {
      cudaStreamCreateWithFlags(&c.tstream[k], cudaStreamNonBlocking);
      cudaEventCreateWithFlags(&c.ev_t[k], cudaEventDisableTiming);
    }
    if (const char* e = std::getenv("VR_TRACE_STREAMS")) c.trace_streams = std::max(0, std::min(std::atoi(e), (int)vr::Comm::kMaxTraceStreams));
    if (const char* e = std::getenv("VR_FOLD_LIGHT")) c.fold_light = std::atoi(e);
    if (const char* e = std::getenv("VR_FOLD_GRID")) c.fold_grid = std::atoi(e);
    if (const char* e = std::getenv("VR_FOLD_NR8")) c.fold_nr8 = std::atoi(e) != 0;
    if (const char* e = std::getenv("VR_LAYER_PUSH")) c.layer_push = std::atoi(e) != 0;
    if (const char* e = std::getenv("VR_EXCHANGE_MAX_CTAS")) c.exchange_max_ctas = std::max(0, std::atoi(e));
  }
  for (int k = 0; k < vr::kAuxStreams; ++k)
  {
    cudaStreamCreateWithFlags(&ctx->aux[k], cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->ev_join[k], cudaEventDisableTiming);
  }
  cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  cudaMalloc(&ctx->sample_counter, sizeof(unsigned long long));
  cudaMalloc(&ctx->partial_count, sizeof(unsigned long long));
  cudaMalloc(&ctx->partial_count_tmp, sizeof(unsigned long long));
  cudaMalloc(&ctx->lut, 1024 * sizeof(float4));
  if ((e = cudaGetLastError()) != cudaSuccess || !ctx->lut)
  {
    delete ctx;
    return fail(nullptr, VR_ERR_NOMEM, "vr_create: small allocations failed: %s", cudaGetErrorString(e));
  }
  cudaMemset(ctx->partial_count, 0, sizeof(unsigned long long));
  cudaMemset(ctx->partial_count_tmp, 0, sizeof(unsigned long long));
  cudaMemset(ctx->sample_counter, 0, sizeof(unsigned long long));
  *out = ctx;
  return VR_OK;
}

static void free_block(Block& b)
{
  if (b.tmap_dev) cudaFree(b.tmap_dev);
  b.tmap_dev = nullptr;
  b.dev.tmap = nullptr;
  if (b.owned_field) cudaFree(b.owned_field);
  if (b.owned_axes) cudaFree(b.owned_axes);
  if (b.owned_xyz) cudaFree(b.owned_xyz);
  if (b.owned_conn) cudaFree(b.owned_conn);
  if (b.owned_bin_start) cudaFree(b.owned_bin_start);
  if (b.owned_bin_cells) cudaFree(b.owned_bin_cells);
  if (b.owned_ext_mask) cudaFree(b.owned_ext_mask);
  if (b.owned_bin_ext) cudaFree(b.owned_bin_ext);
  b.owned_ext_mask = b.owned_bin_ext = nullptr;
  b.owned_xyz = b.owned_conn = nullptr;
  b.owned_bin_start = b.owned_bin_cells = nullptr;
  if (b.line_want) cudaFree(b.line_want);
  if (b.line_have) cudaFree(b.line_have);
  if (b.n_have_dev) cudaFree(b.n_have_dev);
  if (b.n_have_host) cudaFreeHost(b.n_have_host);
  b.n_have_dev = b.n_have_host = nullptr;
  b.all_resident = false;
  b.owned_field = nullptr;
  b.owned_axes = nullptr;
  b.line_want = b.line_have = nullptr;
  b.staged_src = nullptr;
  b.n_lines = 0;
}

namespace vr { void comm_destroy(vr_ctx* ctx); }

extern "C" void vr_destroy(vr_ctx* ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->blocks) free_block(kv.second);
  comm_destroy(ctx);
  cudaFree(ctx->lut);
  if (!ctx->canvas_in_arena)
  {
    cudaFree(ctx->canvas_rgba);
    cudaFree(ctx->canvas_depth);
  }
  if (!ctx->img_in_arena)
  {
    cudaFree(ctx->img_rgba);
    cudaFree(ctx->img_depth);
    cudaFree(ctx->res_rgba);
    cudaFree(ctx->res_depth);
  }
  for (int k = 0; k < vr_ctx::kOwnLayerRing; ++k)
  {
    cudaFree(ctx->own_ltab[k]);
    cudaFree(ctx->own_lpool_rgba[k]);
    cudaFree(ctx->own_lpool_depth[k]);
  }
  delete ctx->ltab_host;
  cudaFree(ctx->partials);
  cudaFree(ctx->partials_tmp);
  cudaFree(ctx->partial_count);
  cudaFree(ctx->partial_count_tmp);
  cudaFree(ctx->px_count);
  cudaFree(ctx->px_end);
  cudaFree(ctx->sidx);
  cudaFree(ctx->rec);
  cudaFree(ctx->scan_blocks);
  cudaFree(ctx->enc_rgba);
  cudaFree(ctx->scratch_u64);
  cudaFree(ctx->tile_counter);
  free_multi(ctx);
  cudaFree(ctx->png_scratch);
  cudaFree(ctx->png_out);
  cudaFree(ctx->png_total);
  if (ctx->png_total_host) cudaFreeHost(ctx->png_total_host);
  cudaFree(ctx->sample_counter);
  for (int k = 0; k < vr::kAuxStreams; ++k)
  {
    if (ctx->aux[k]) cudaStreamDestroy(ctx->aux[k]);
    if (ctx->ev_join[k]) cudaEventDestroy(ctx->ev_join[k]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

extern "C" const char* vr_last_error(const vr_ctx* ctx)
{
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" vr_status vr_set_stream(vr_ctx* ctx, void* cuda_stream)
{
  VR_ENTER_RO(ctx);
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return VR_OK;
}

extern "C" vr_status vr_synchronize(vr_ctx* ctx)
{
  VR_ENTER_RO(ctx);
  CK(cudaStreamSynchronize(ctx->stream));
  return comm_check_errors(ctx); // an exchange that timed out / was aborted by a peer is reported here
}

extern "C" uint64_t vr_kernel_launches(const vr_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" vr_status vr_comm_join(vr_ctx* ctx)
{
  VR_ENTER_RO(ctx); // (the prologue is the join)
  return VR_OK;
}

extern "C" vr_status vr_comm_timeline(vr_ctx* ctx, uint64_t out_ns[16])
{
  VR_ENTER_RO(ctx);
  REQUIRE(out_ns, "vr_comm_timeline: NULL output");
  unsigned long long* src = comm_timeline_slot(ctx, 0);
  REQUIRE(src != nullptr, "vr_comm_timeline: no exchange arena, or VR_TIMELINE was not set at vr_comm_init");
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(out_ns, src, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return VR_OK;
}

// ================================================================= blocks
static bool encode_brick_tensor_map(const BlockDev& b, unsigned char out[128]);

static size_t field_bytes(const BlockDev& d)
{
  const size_t n = d.assoc == VR_POINT ? (size_t)d.dims[0] * d.dims[1] * d.dims[2]
                                       : (size_t)(d.dims[0] - 1) * (d.dims[1] - 1) * (d.dims[2] - 1);
  return n * (d.dtype == VR_F32 ? 4 : 8);
}

// `old`: the block this upload replaces (same id), or null.  A re-publish of a same-sized field
// (the in-situ steady state: one upload per simulation cycle) reuses the device allocation.
static vr_status upload_field(vr_ctx* ctx, Block& b, const void* field, int dtype, int assoc, int where,
                              Block* old)
{
  REQUIRE(field != nullptr, "block: field is NULL");
  REQUIRE(dtype == VR_F32 || dtype == VR_F64, "block: dtype must be VR_F32 or VR_F64");
  REQUIRE(assoc == VR_POINT || assoc == VR_CELL, "block: assoc must be VR_POINT or VR_CELL");
  REQUIRE(where == VR_HOST || where == VR_DEVICE || where == VR_HOST_MAPPED || where == VR_HOST_STAGED,
          "block: where must be VR_HOST, VR_DEVICE, VR_HOST_MAPPED or VR_HOST_STAGED");
  const int* d = b.dev.dims;
  REQUIRE(d[0] >= 2 && d[1] >= 2 && d[2] >= 2, "block: point dims must be >= 2 (got %d %d %d)", d[0],
          d[1], d[2]);
  const size_t n = assoc == VR_POINT ? (size_t)d[0] * d[1] * d[2]
                                     : (size_t)(d[0] - 1) * (d[1] - 1) * (d[2] - 1);
  const size_t bytes = n * (dtype == VR_F32 ? 4 : 8);
  b.dev.dtype = dtype;
  b.dev.assoc = assoc;
  if (where == VR_DEVICE)
    b.dev.field = field;
  else if (where == VR_HOST_STAGED)
  {
    // demand staging (stage.cu): nothing moves now; every trace first pulls the 128-byte lines its
    // rays will touch and that are not on the device yet
    void* dptr = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(&dptr, const_cast<void*>(field), 0);
    if (e != cudaSuccess)
    {
      cudaGetLastError();
      return fail(ctx, VR_ERR_INVALID, "block: VR_HOST_STAGED field is not page-locked mapped host memory "
                  "(cudaHostAlloc / cudaHostRegister): %s", cudaGetErrorString(e));
    }
    const size_t n_lines = (bytes + 127) / 128;
    if (old && old->owned_field && old->line_have && field_bytes(old->dev) == bytes && old->n_lines == n_lines)
    {
      // re-publish of a same-sized field (one per simulation cycle): keep the buffers, forget the lines
      b.owned_field = old->owned_field; old->owned_field = nullptr;
      b.line_want = old->line_want; old->line_want = nullptr;
      b.line_have = old->line_have; old->line_have = nullptr;
      b.n_have_dev = old->n_have_dev; old->n_have_dev = nullptr;
      b.n_have_host = old->n_have_host; old->n_have_host = nullptr;
      // the pinned mirror may still be the target of an async copy of the previous publish
      CK(cudaStreamSynchronize(ctx->stream));
    }
    else
    {
      CK(cudaMalloc(&b.owned_field, n_lines * 128));
      CK(cudaMalloc(&b.line_want, n_lines));
      CK(cudaMalloc(&b.line_have, n_lines));
      CK(cudaMalloc(&b.n_have_dev, sizeof(unsigned long long)));
      CK(cudaHostAlloc(&b.n_have_host, sizeof(unsigned long long), cudaHostAllocDefault));
      CK(cudaMemsetAsync(b.line_want, 0, n_lines, ctx->stream));
    }
    CK(cudaMemsetAsync(b.line_have, 0, n_lines, ctx->stream));
    CK(cudaMemsetAsync(b.n_have_dev, 0, sizeof(unsigned long long), ctx->stream));
    *b.n_have_host = 0;
    b.all_resident = false;
    b.n_lines = n_lines;
    b.staged_src = dptr;
    b.dev.field = b.owned_field;
  }
  else if (where == VR_HOST_MAPPED)
  {
    // page-locked host memory sampled in place over PCIe: the sparse default sampling touches about
    // a quarter of the field's 32-byte sectors, so pulling just those beats copying the whole block
    void* dptr = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(&dptr, const_cast<void*>(field), 0);
    if (e != cudaSuccess)
    {
      cudaGetLastError();
      return fail(ctx, VR_ERR_INVALID, "block: VR_HOST_MAPPED field is not page-locked mapped host memory "
                  "(cudaHostAlloc / cudaHostRegister): %s", cudaGetErrorString(e));
    }
    b.dev.field = dptr;
  }
  else
  {
    if (old && old->owned_field && field_bytes(old->dev) == bytes)
    {
      b.owned_field = old->owned_field; // stream order keeps earlier renders ahead of the copy
      old->owned_field = nullptr;
    }
    else
      CK(cudaMalloc(&b.owned_field, bytes));
    CK(cudaMemcpyAsync(b.owned_field, field, bytes, cudaMemcpyHostToDevice, ctx->stream));
    b.dev.field = b.owned_field;
  }
  return VR_OK;
}

// Strided Blueprint values (ascent_vtkh_data_adapter.cpp:1836-1887: a byte stride that is a multiple of the
// element size, handed to VTK-m as an ArrayHandleStride -- one component of an interleaved mcarray, a padded
// array): gathered ONCE per publish into a dense device array that the block entry points adopt with
// VR_DEVICE, so the sampler keeps its unit-stride, coalescible gathers.
extern "C" vr_status vr_field_gather_strided(vr_ctx* ctx, const void* src, int where, int dtype, size_t n_values,
                                             size_t element_stride, size_t element_offset, void** dense_dev_out)
{
  VR_ENTER_RO(ctx);
  REQUIRE(src && dense_dev_out, "vr_field_gather_strided: NULL argument");
  REQUIRE(dtype == VR_F32 || dtype == VR_F64, "vr_field_gather_strided: dtype must be VR_F32 or VR_F64");
  REQUIRE(where == VR_HOST || where == VR_DEVICE, "vr_field_gather_strided: where must be VR_HOST or VR_DEVICE");
  REQUIRE(element_stride >= 1 && n_values >= 1, "vr_field_gather_strided: stride and count must be positive");
  CK(cudaSetDevice(ctx->device));
  *dense_dev_out = nullptr;
  const size_t eb = dtype == VR_F32 ? 4 : 8;
  void* dense = nullptr;
  CK(cudaMalloc(&dense, n_values * eb));
  const unsigned char* from = static_cast<const unsigned char*>(src) + element_offset * eb;
  void* span_dev = nullptr;
  cudaError_t e = cudaSuccess;
  if (where == VR_HOST)
  {
    // the whole strided span crosses PCIe once (the DMA engine has no 4-byte gather worth using)
    const size_t span = ((n_values - 1) * element_stride + 1) * eb;
    e = cudaMalloc(&span_dev, span);
    if (e == cudaSuccess) e = cudaMemcpyAsync(span_dev, from, span, cudaMemcpyHostToDevice, ctx->stream);
    from = static_cast<const unsigned char*>(span_dev);
  }
  if (e == cudaSuccess) e = launch_gather_strided(from, (int)eb, element_stride, n_values, dense, ctx->sm_count, ctx->stream);
  ctx->launches++;
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); // (the caller may reuse its array; the span is freed)
  cudaFree(span_dev);
  if (e != cudaSuccess)
  {
    cudaFree(dense);
    return fail(ctx, VR_ERR_CUDA, "vr_field_gather_strided: %s", cudaGetErrorString(e));
  }
  *dense_dev_out = dense;
  return VR_OK;
}

extern "C" vr_status vr_field_free(vr_ctx* ctx, void* dense_dev)
{
  VR_ENTER_RO(ctx);
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(dense_dev);
  return VR_OK;
}

static void free_multi(vr_ctx* ctx)
{
  cudaFree(ctx->multi_table);
  cudaFree(ctx->multi_tile_end);
  if (ctx->multi_host) cudaFreeHost(ctx->multi_host);
  for (int k = 0; k < vr_ctx::kMultiSlots; ++k)
    if (ctx->multi_ev[k]) { cudaEventDestroy(ctx->multi_ev[k]); ctx->multi_ev[k] = nullptr; }
  ctx->multi_table = nullptr;
  ctx->multi_tile_end = nullptr;
  ctx->multi_host = nullptr;
  ctx->multi_cap = 0;
}

// first publish of a (grid kind, scalar type, association, index width) on this context: load the sampler
// kernels such a block can run (vr_internal.h, "kernel preloading")
static void preload_block_kernels(vr_ctx* ctx, const BlockDev& d)
{
  const long long n = (long long)d.dims[0] * d.dims[1] * d.dims[2];
  const unsigned bit = 1u << ((d.kind & 1) | (d.dtype & 1) << 1 | (d.assoc & 1) << 2 | (n >= (1ll << 31) ? 8 : 0));
  if (ctx->trace_preloaded & bit) return;
  preload_trace_kernels(d);
  ctx->trace_preloaded |= bit;
}

extern "C" vr_status vr_block_uniform(vr_ctx* ctx, int block_id, const int dims[3],
                                      const float origin[3], const float spacing[3],
                                      const void* field, int dtype, int assoc, int where)
{
  VR_ENTER_RO(ctx);
  REQUIRE(dims && origin && spacing, "vr_block_uniform: NULL argument");
  CK(cudaSetDevice(ctx->device));
  Block b;
  std::memset(&b.dev, 0, sizeof(b.dev));
  b.dev.kind = 0;
  for (int a = 0; a < 3; ++a)
  {
    REQUIRE(spacing[a] > 0.f, "vr_block_uniform: spacing must be positive");
    b.dev.dims[a] = dims[a];
    b.dev.origin[a] = origin[a];
    b.dev.spacing[a] = spacing[a];
    // UniformLocator: MaxPoint = Origin + spacing * (dims - 1) in f32
    b.dev.min_point[a] = origin[a];
    b.dev.max_point[a] = origin[a] + spacing[a] * (float)(dims[a] - 1);
    b.dev.inv_spacing[a] = 1.f / spacing[a];
    // coords.GetBounds(): f64 arithmetic on the f32-valued origin/spacing
    b.bounds[2 * a] = (double)origin[a];
    b.bounds[2 * a + 1] = (double)origin[a] + (double)spacing[a] * (double)(dims[a] - 1);
  }
  auto it = ctx->blocks.find(block_id);
  Block* old = it != ctx->blocks.end() ? &it->second : nullptr;
  vr_status st = upload_field(ctx, b, field, dtype, assoc, where, old);
  if (st != VR_OK) { free_block(b); return st; }
  // dense-sampling candidates get the brick march's tensor map now (f32 point field, x-rows on 16-byte
  // boundaries, resident in device memory); 128 bytes, copied synchronously
  if (dtype == VR_F32 && assoc == VR_POINT && dims[0] % 4 == 0 && where != VR_HOST_STAGED && where != VR_HOST_MAPPED &&
      (reinterpret_cast<uintptr_t>(b.dev.field) & 15) == 0 && !ctx->no_brick)
  {
    alignas(64) unsigned char m[128];
    if (encode_brick_tensor_map(b.dev, m) && cudaMalloc(&b.tmap_dev, 128) == cudaSuccess)
    {
      if (cudaMemcpy(b.tmap_dev, m, 128, cudaMemcpyHostToDevice) == cudaSuccess) b.dev.tmap = b.tmap_dev;
    }
    else
      cudaGetLastError();
  }
  if (old && (old->owned_field || old->owned_axes || old->line_want || old->line_have || old->tmap_dev))
  {
    cudaStreamSynchronize(ctx->stream);
    free_block(*old);
  }
  ctx->blocks[block_id] = b;
  preload_block_kernels(ctx, b.dev);
  return VR_OK;
}

extern "C" vr_status vr_block_rectilinear(vr_ctx* ctx, int block_id, const int dims[3],
                                          const double* x, const double* y, const double* z,
                                          const void* field, int dtype, int assoc, int where)
{
  VR_ENTER_RO(ctx);
  REQUIRE(dims && x && y && z, "vr_block_rectilinear: NULL argument");
  CK(cudaSetDevice(ctx->device));
  Block b;
  std::memset(&b.dev, 0, sizeof(b.dev));
  b.dev.kind = 1;
  const double* ax[3] = { x, y, z };
  const size_t total = (size_t)dims[0] + dims[1] + dims[2];
  std::vector<float> host(total);
  size_t off = 0;
  for (int a = 0; a < 3; ++a)
  {
    REQUIRE(dims[a] >= 2, "vr_block_rectilinear: point dims must be >= 2");
    b.dev.dims[a] = dims[a];
    for (int i = 0; i < dims[a]; ++i)
    {
      host[off + i] = (float)ax[a][i]; // RectilinearLocator reads f64 axes and narrows per access
      REQUIRE(i == 0 || ax[a][i] > ax[a][i - 1], "vr_block_rectilinear: axis %d not increasing", a);
    }
    b.dev.min_point[a] = (float)ax[a][0];
    b.dev.max_point[a] = (float)ax[a][dims[a] - 1];
    b.bounds[2 * a] = ax[a][0];
    b.bounds[2 * a + 1] = ax[a][dims[a] - 1];
    off += dims[a];
  }
  CK(cudaMalloc(&b.owned_axes, total * sizeof(float)));
  // staged through pageable memory: sync copy is fine for a few KiB
  CK(cudaMemcpy(b.owned_axes, host.data(), total * sizeof(float), cudaMemcpyHostToDevice));
  b.dev.axis[0] = b.owned_axes;
  b.dev.axis[1] = b.owned_axes + dims[0];
  b.dev.axis[2] = b.owned_axes + dims[0] + dims[1];
  auto it = ctx->blocks.find(block_id);
  Block* old = it != ctx->blocks.end() ? &it->second : nullptr;
  vr_status st = upload_field(ctx, b, field, dtype, assoc, where, old);
  if (st != VR_OK) { free_block(b); return st; }
  if (old && (old->owned_field || old->owned_axes || old->line_want || old->line_have))
  {
    cudaStreamSynchronize(ctx->stream);
    free_block(*old);
  }
  ctx->blocks[block_id] = b;
  preload_block_kernels(ctx, b.dev);
  return VR_OK;
}

static const std::vector<unsigned char>& cached_external_mask(vr_ctx* ctx, int block_id, const int* conn, size_t n_cells, int shape)
{
  auto& c = ctx->umask_cache[block_id];
  const unsigned long long h = umesh_conn_hash(conn, n_cells * (size_t)shape);
  if (c.mask.size() != n_cells || c.n_cells != n_cells || c.shape != shape || c.hash != h)
  {
    c.mask = umesh_external_mask(conn, n_cells, shape);
    c.hash = h;
    c.n_cells = n_cells;
    c.shape = shape;
  }
  return c.mask;
}

// N4: an explicit cell set.  Coordinates and connectivity are narrowed to f32 / int32 on the way in (VTK-m's
// unstructured tracer works in f32 as well); the cell locator is built on the device.
extern "C" vr_status vr_block_unstructured(vr_ctx* ctx, int block_id, size_t n_points, const void* xyz, int coord_dtype,
                                           size_t n_cells, int cell_shape, const void* connectivity, int index_bits,
                                           const void* field, int dtype, int assoc, int where)
{
  VR_ENTER_RO(ctx);
  REQUIRE(xyz && connectivity && field, "vr_block_unstructured: NULL argument");
  REQUIRE(cell_shape == VR_HEXAHEDRON || cell_shape == VR_TETRA,
          "vr_block_unstructured: cell shape must be VR_HEXAHEDRON (12) or VR_TETRA (10)");
  REQUIRE(coord_dtype == VR_F32 || coord_dtype == VR_F64, "vr_block_unstructured: coordinates must be VR_F32 or VR_F64");
  REQUIRE(index_bits == 32 || index_bits == 64, "vr_block_unstructured: connectivity must be 32- or 64-bit integers");
  REQUIRE(dtype == VR_F32 || dtype == VR_F64, "vr_block_unstructured: field must be VR_F32 or VR_F64");
  REQUIRE(assoc == VR_POINT || assoc == VR_CELL, "vr_block_unstructured: bad association");
  REQUIRE(where == VR_HOST || where == VR_DEVICE, "vr_block_unstructured: where must be VR_HOST or VR_DEVICE");
  REQUIRE(n_points >= 4 && n_cells >= 1 && n_points < (1ull << 31) && n_cells < (1ull << 31) / 8,
          "vr_block_unstructured: empty or too large mesh");
  REQUIRE(where == VR_HOST || (coord_dtype == VR_F32 && index_bits == 32),
          "vr_block_unstructured: device arrays are adopted in place and must be f32 coordinates / 32-bit connectivity");
  CK(cudaSetDevice(ctx->device));
  const int shape = cell_shape == VR_HEXAHEDRON ? 8 : 4;
  Block b;
  std::memset(&b.dev, 0, sizeof(b.dev));
  std::memset(&b.um, 0, sizeof(b.um));
  b.dev.kind = 2;
  b.dev.dtype = dtype;
  b.dev.assoc = assoc;
  const size_t n_field = assoc == VR_POINT ? n_points : n_cells;
  const size_t fb = n_field * (dtype == VR_F32 ? 4 : 8);
  cudaError_t e = cudaSuccess;
  std::vector<unsigned char> ext_mask; // which faces of which cells form the mesh boundary (host side, one sort)
  if (where == VR_DEVICE)
  {
    b.um.xyz = static_cast<const float*>(xyz);
    b.um.conn = static_cast<const int*>(connectivity);
    b.um.field = field;
    std::vector<int> hc(n_cells * (size_t)shape);
    e = cudaMemcpyAsync(hc.data(), connectivity, hc.size() * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return fail(ctx, VR_ERR_CUDA, "vr_block_unstructured: %s", cudaGetErrorString(e));
    for (size_t i = 0; i < hc.size(); ++i)
      if (hc[i] < 0 || (size_t)hc[i] >= n_points) return fail(ctx, VR_ERR_INVALID, "vr_block_unstructured: connectivity entry %zu out of range", i);
    ext_mask = cached_external_mask(ctx, block_id, hc.data(), n_cells, shape);
  }
  else
  {
    std::vector<float> hx(n_points * 3);
    if (coord_dtype == VR_F32) std::memcpy(hx.data(), xyz, hx.size() * 4);
    else
      for (size_t i = 0; i < hx.size(); ++i) hx[i] = (float)static_cast<const double*>(xyz)[i];
    std::vector<int> hc(n_cells * (size_t)shape);
    for (size_t i = 0; i < hc.size(); ++i)
    {
      const long long v = index_bits == 32 ? (long long)static_cast<const int*>(connectivity)[i]
                                           : static_cast<const long long*>(connectivity)[i];
      if (v < 0 || (size_t)v >= n_points) return fail(ctx, VR_ERR_INVALID, "vr_block_unstructured: connectivity entry %zu out of range", i);
      hc[i] = (int)v;
    }
    ext_mask = cached_external_mask(ctx, block_id, hc.data(), n_cells, shape);
    e = cudaMalloc(&b.owned_xyz, hx.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&b.owned_conn, hc.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&b.owned_field, fb);
    if (e == cudaSuccess) e = cudaMemcpyAsync(b.owned_xyz, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(b.owned_conn, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(b.owned_field, field, fb, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); // (the staging vectors die with this scope)
    if (e != cudaSuccess) { free_block(b); return fail(ctx, VR_ERR_CUDA, "vr_block_unstructured: %s", cudaGetErrorString(e)); }
    b.um.xyz = static_cast<const float*>(b.owned_xyz);
    b.um.conn = static_cast<const int*>(b.owned_conn);
    b.um.field = b.owned_field;
  }
  e = cudaMalloc(&b.owned_ext_mask, ext_mask.size());
  if (e == cudaSuccess) e = cudaMemcpyAsync(b.owned_ext_mask, ext_mask.data(), ext_mask.size(), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { free_block(b); return fail(ctx, VR_ERR_CUDA, "vr_block_unstructured (boundary): %s", cudaGetErrorString(e)); }
  b.um.ext_mask = b.owned_ext_mask;
  b.um.dtype = dtype;
  b.um.assoc = assoc;
  b.um.shape = shape;
  b.um.n_cells = (int)n_cells;
  b.um.n_points = (int)n_points;
  // point bounds (device reduction), bins per axis = ceil(cbrt(n_cells)) capped at 256
  int* keys = nullptr;
  e = cudaMalloc(&keys, 6 * sizeof(int));
  if (e == cudaSuccess) e = umesh_bounds(b.um.xyz, n_points, keys, b.um.bmin, b.um.bmax, ctx->sm_count, ctx->stream);
  cudaFree(keys);
  if (e != cudaSuccess) { free_block(b); return fail(ctx, VR_ERR_CUDA, "vr_block_unstructured (bounds): %s", cudaGetErrorString(e)); }
  int g = (int)std::ceil(std::cbrt((double)n_cells));
  g = std::max(1, std::min(g, 256));
  for (int a = 0; a < 3; ++a)
  {
    b.um.g[a] = g;
    const float ext = b.um.bmax[a] - b.um.bmin[a];
    b.um.ginv[a] = ext > 0.f ? (float)g / ext : 0.f;
    b.bounds[2 * a] = (double)b.um.bmin[a];
    b.bounds[2 * a + 1] = (double)b.um.bmax[a];
    b.dev.min_point[a] = b.um.bmin[a];
    b.dev.max_point[a] = b.um.bmax[a];
  }
  e = umesh_build_bins(b.um, &b.owned_bin_start, &b.owned_bin_cells, &b.owned_bin_ext, ctx->sm_count, ctx->stream);
  ctx->launches += 5;
  if (e != cudaSuccess) { free_block(b); return fail(ctx, VR_ERR_CUDA, "vr_block_unstructured (bins): %s", cudaGetErrorString(e)); }
  auto it = ctx->blocks.find(block_id);
  if (it != ctx->blocks.end())
  {
    cudaStreamSynchronize(ctx->stream);
    free_block(it->second);
  }
  ctx->blocks[block_id] = b;
  return VR_OK;
}

extern "C" vr_status vr_block_free(vr_ctx* ctx, int block_id)
{
  VR_ENTER_RO(ctx);
  auto it = ctx->blocks.find(block_id);
  REQUIRE(it != ctx->blocks.end(), "vr_block_free: unknown block %d", block_id);
  CK(cudaStreamSynchronize(ctx->stream));
  free_block(it->second);
  ctx->blocks.erase(it);
  ctx->umask_cache.erase(block_id);
  return VR_OK;
}

extern "C" vr_status vr_block_staged_bytes(vr_ctx* ctx, int block_id, size_t* bytes)
{
  VR_ENTER_RO(ctx);
  auto it = ctx->blocks.find(block_id);
  REQUIRE(it != ctx->blocks.end() && bytes, "vr_block_staged_bytes: unknown block %d", block_id);
  const Block& b = it->second;
  *bytes = 0;
  if (!b.staged_src) return VR_OK;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->scratch_u64) CK(cudaMalloc(&ctx->scratch_u64, sizeof(unsigned long long)));
  CK(launch_count_lines(b.line_have, b.n_lines, ctx->scratch_u64, ctx->stream));
  unsigned long long n = 0;
  CK(cudaMemcpyAsync(&n, ctx->scratch_u64, sizeof(n), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *bytes = (size_t)n * 128;
  return VR_OK;
}

extern "C" vr_status vr_block_bounds(vr_ctx* ctx, int block_id, double out[6])
{
  VR_ENTER_RO(ctx);
  auto it = ctx->blocks.find(block_id);
  REQUIRE(it != ctx->blocks.end() && out, "vr_block_bounds: unknown block %d", block_id);
  std::memcpy(out, it->second.bounds, sizeof(double) * 6);
  return VR_OK;
}

// ================================================================= transfer function
extern "C" vr_status vr_set_tf(vr_ctx* ctx, const float* rgba, int n_entries)
{
  VR_ENTER_RO(ctx);
  REQUIRE(rgba && n_entries >= 2 && n_entries <= 1024, "vr_set_tf: need 2..1024 entries (got %d)",
          n_entries);
  CK(cudaSetDevice(ctx->device));
  // small pageable copy: synchronous w.r.t. the host, ordered on the stream
  CK(cudaMemcpyAsync(ctx->lut, rgba, (size_t)n_entries * sizeof(float4), cudaMemcpyHostToDevice,
                     ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->lut_size = n_entries;
  return VR_OK;
}

// ================================================================= frame buffers
namespace vr { vr_status comm_bind_frame(vr_ctx* ctx, size_t n_pixels); }

static vr_status ensure_frame(vr_ctx* ctx, int W, int H)
{
  REQUIRE(W > 0 && H > 0 && (long long)W * H < (1ll << 31), "bad image size %d x %d", W, H);
  const size_t n = (size_t)W * H;
  if (n > ctx->cap_pixels)
  {
    CK(cudaStreamSynchronize(ctx->stream));
    REQUIRE(!ctx->canvas_in_arena, "image larger than the max_pixels given to vr_comm_init");
    cudaFree(ctx->canvas_rgba);
    cudaFree(ctx->canvas_depth);
    ctx->canvas_rgba = nullptr;
    ctx->canvas_depth = nullptr;
    CK(cudaMalloc(&ctx->canvas_rgba, n * sizeof(float4)));
    CK(cudaMalloc(&ctx->canvas_depth, n * sizeof(float)));
    if (!ctx->comm.on)
    {
      cudaFree(ctx->img_rgba); cudaFree(ctx->img_depth); cudaFree(ctx->res_rgba); cudaFree(ctx->res_depth);
      ctx->img_rgba = ctx->res_rgba = nullptr;
      ctx->img_depth = ctx->res_depth = nullptr;
      CK(cudaMalloc(&ctx->img_rgba, n * sizeof(uchar4)));
      CK(cudaMalloc(&ctx->img_depth, n * sizeof(float)));
      CK(cudaMalloc(&ctx->res_rgba, n * sizeof(uchar4)));
      CK(cudaMalloc(&ctx->res_depth, n * sizeof(float)));
    }
    ctx->cap_pixels = n;
    // fresh, uninitialised buffers: nothing about their contents may be assumed any more
    ctx->comm.clean_canvas.valid = false;
    ctx->comm.clean_res[0].valid = ctx->comm.clean_res[1].valid = false;
  }
  if (ctx->comm.on)
  {
    vr_status st = comm_bind_frame(ctx, n);
    if (st != VR_OK) return st;
  }
  if (W != ctx->W || H != ctx->H)
  {
    // a frame traced ahead belongs to the pending exchange's size: changing it now would hand that
    // exchange the wrong pixel count
    REQUIRE(!ctx->img_ahead, "frame size changed (%dx%d -> %dx%d) while a frame traced ahead is pending", ctx->W,
            ctx->H, W, H);
    ctx->comm.clean_canvas.valid = false;
    ctx->comm.clean_res[0].valid = ctx->comm.clean_res[1].valid = false;
  }
  ctx->W = W;
  ctx->H = H;
  return VR_OK;
}

extern "C" vr_status vr_canvas_clear(vr_ctx* ctx, int width, int height)
{
  VR_ENTER(ctx);
  CK(cudaSetDevice(ctx->device));
  vr_status st = ensure_frame(ctx, width, height);
  if (st != VR_OK) return st;
  CK(launch_canvas_clear(ctx->canvas_rgba, ctx->canvas_depth, (size_t)width * height, ctx->stream));
  ctx->launches++;
  return VR_OK;
}

extern "C" vr_status vr_canvas_upload(vr_ctx* ctx, int width, int height, const float* rgba,
                                      const float* depth)
{
  VR_ENTER(ctx);
  REQUIRE(rgba && depth, "vr_canvas_upload: NULL buffer");
  CK(cudaSetDevice(ctx->device));
  vr_status st = ensure_frame(ctx, width, height);
  if (st != VR_OK) return st;
  const size_t n = (size_t)width * height;
  CK(cudaMemcpyAsync(ctx->canvas_rgba, rgba, n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->canvas_depth, depth, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  return VR_OK;
}

extern "C" vr_status vr_canvas_download(vr_ctx* ctx, float* rgba, float* depth)
{
  VR_ENTER_RO(ctx);
  REQUIRE(ctx->W > 0, "vr_canvas_download: no canvas yet");
  const size_t n = (size_t)ctx->W * ctx->H;
  if (rgba) CK(cudaMemcpyAsync(rgba, ctx->canvas_rgba, n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  if (depth) CK(cudaMemcpyAsync(depth, ctx->canvas_depth, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return comm_check_errors(ctx);
}

// Only the rectangle [x0, x1) x [y0, y1) of the canvas, into the same pixels of FULL-FRAME host buffers: a frame
// that started from Canvas::Clear differs from the cleared canvas (colour 0, depth 1.001 -- which the caller's
// host canvas already holds after Render::ClearCanvas) only inside the screen footprint of the data, and the
// 20 B/pixel float canvas is what dominates the read-back (c2: 41 MB for a 10 MB footprint).
extern "C" vr_status vr_canvas_download_rect(vr_ctx* ctx, int x0, int y0, int x1, int y1, float* rgba, float* depth)
{
  VR_ENTER_RO(ctx);
  REQUIRE(ctx->W > 0, "vr_canvas_download_rect: no canvas yet");
  const int W = ctx->W, H = ctx->H;
  x0 = std::max(0, std::min(x0, W)); x1 = std::max(x0, std::min(x1, W));
  y0 = std::max(0, std::min(y0, H)); y1 = std::max(y0, std::min(y1, H));
  if (x1 > x0 && y1 > y0)
  {
    const size_t at = (size_t)y0 * W + x0;
    if (rgba)
      CK(cudaMemcpy2DAsync(rgba + 4 * at, (size_t)W * sizeof(float4), ctx->canvas_rgba + at, (size_t)W * sizeof(float4),
                           (size_t)(x1 - x0) * sizeof(float4), (size_t)(y1 - y0), cudaMemcpyDeviceToHost, ctx->stream));
    if (depth)
      CK(cudaMemcpy2DAsync(depth + at, (size_t)W * sizeof(float), ctx->canvas_depth + at, (size_t)W * sizeof(float),
                           (size_t)(x1 - x0) * sizeof(float), (size_t)(y1 - y0), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return comm_check_errors(ctx);
}

extern "C" vr_status vr_canvas_blend_background(vr_ctx* ctx, const float bg_rgba[4])
{
  VR_ENTER(ctx);
  REQUIRE(ctx->W > 0, "vr_canvas_blend_background: no canvas yet");
  REQUIRE(bg_rgba, "vr_canvas_blend_background: NULL colour");
  CK(cudaSetDevice(ctx->device));
  CK(launch_blend_background(ctx->canvas_rgba, (size_t)ctx->W * ctx->H, bg_rgba, ctx->stream));
  ctx->launches++;
  return VR_OK;
}

extern "C" vr_status vr_canvas_download_rgba8(vr_ctx* ctx, const float* bg_rgba, int flip_rows, uint8_t* out_rgba8)
{
  VR_ENTER_RO(ctx);
  REQUIRE(ctx->W > 0, "vr_canvas_download_rgba8: no canvas yet");
  REQUIRE(out_rgba8, "vr_canvas_download_rgba8: NULL output");
  CK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)ctx->W * ctx->H;
  if (n > ctx->enc_cap)
  {
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->enc_rgba);
    ctx->enc_rgba = nullptr;
    ctx->enc_cap = 0;
    CK(cudaMalloc(&ctx->enc_rgba, n * sizeof(uchar4)));
    ctx->enc_cap = n;
  }
  CK(launch_encode_rgba8(ctx->canvas_rgba, ctx->W, ctx->H, flip_rows ? 1 : 0, bg_rgba, ctx->enc_rgba, ctx->stream));
  ctx->launches++;
  CK(cudaMemcpyAsync(out_rgba8, ctx->enc_rgba, n * sizeof(uchar4), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VR_OK;
}

extern "C" vr_status vr_canvas_encode_png(vr_ctx* ctx, const float* bg_rgba, uint8_t* png_host, size_t capacity,
                                          size_t* png_bytes)
{
  VR_ENTER_RO(ctx);
  REQUIRE(ctx->W > 0, "vr_canvas_encode_png: no canvas yet");
  REQUIRE(png_bytes, "vr_canvas_encode_png: NULL size output");
  REQUIRE(ctx->W <= 16384, "vr_canvas_encode_png: width above 16384");
  CK(cudaSetDevice(ctx->device));
  const int W = ctx->W, H = ctx->H;
  const size_t n = (size_t)W * H;
  if (n > ctx->enc_cap)
  {
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->enc_rgba);
    ctx->enc_rgba = nullptr;
    ctx->enc_cap = 0;
    CK(cudaMalloc(&ctx->enc_rgba, n * sizeof(uchar4)));
    ctx->enc_cap = n;
  }
  const size_t need_scratch = png_scratch_bytes(W, H), need_out = png_capacity(W, H);
  if (need_scratch > ctx->png_scratch_cap || need_out > ctx->png_out_cap)
  {
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->png_scratch);
    cudaFree(ctx->png_out);
    ctx->png_scratch = ctx->png_out = nullptr;
    ctx->png_scratch_cap = ctx->png_out_cap = 0;
    CK(cudaMalloc(&ctx->png_scratch, need_scratch));
    CK(cudaMalloc(&ctx->png_out, need_out));
    ctx->png_scratch_cap = need_scratch;
    ctx->png_out_cap = need_out;
  }
  if (!ctx->png_total)
  {
    CK(cudaMalloc(&ctx->png_total, sizeof(unsigned long long)));
    CK(cudaMallocHost(&ctx->png_total_host, sizeof(unsigned long long)));
  }
  // PNGEncoder::Encode: (unsigned char)(c * 255.f), rows flipped (ascent_png_encoder.cpp:266-281)
  CK(launch_encode_rgba8(ctx->canvas_rgba, W, H, 1, bg_rgba, ctx->enc_rgba, ctx->stream));
  CK(launch_png_encode(ctx->enc_rgba, W, H, ctx->png_scratch, ctx->png_out, ctx->png_out_cap, ctx->png_total, ctx->sm_count,
                       ctx->stream));
  ctx->launches += 4;
  CK(cudaMemcpyAsync(ctx->png_total_host, ctx->png_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const size_t total = (size_t)*ctx->png_total_host;
  *png_bytes = total; // (also when it does not fit: the caller learns how much to provide)
  REQUIRE(total <= ctx->png_out_cap, "vr_canvas_encode_png: internal bound of the file size exceeded");
  if (!png_host) return VR_OK; // size query
  REQUIRE(total <= capacity, "vr_canvas_encode_png: the file needs %zu bytes, the buffer holds %zu", total, capacity);
  CK(cudaMemcpyAsync(png_host, ctx->png_out, total, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VR_OK;
}

extern "C" size_t vr_png_bound(int width, int height) { return png_capacity(width, height); }

extern "C" vr_status vr_canvas_ptrs(vr_ctx* ctx, void** rgba_dev, void** depth_dev)
{
  VR_ENTER(ctx);
  REQUIRE(ctx->W > 0, "vr_canvas_ptrs: no canvas yet");
  ctx->canvas_exposed = true; // the caller may write through these at any time: no more clean-rectangle shortcuts
  if (rgba_dev) *rgba_dev = ctx->canvas_rgba;
  if (depth_dev) *depth_dev = ctx->canvas_depth;
  return VR_OK;
}

// ================================================================= tracing
// VR_HOST_STAGED blocks: run the sampler's pre-pass for exactly this trace (same params, mode 4) and
// fetch the flagged lines that are still missing, on stream `s`, before the real launch.
static vr_status stage_for_trace(vr_ctx* ctx, int block_id, const TraceParams& p, cudaStream_t s,
                                 unsigned int* counter, bool zero_counter)
{
  auto it = ctx->blocks.find(block_id);
  if (it == ctx->blocks.end() || !it->second.staged_src) return VR_OK;
  Block& b = it->second;
  if (p.sw <= 0 || p.sh <= 0 || b.all_resident) return VR_OK;
  // most of the block is on the device already (many views per publish): pull the rest in one sweep
  // and stop paying for pre-passes.  The mirror may lag by one trace; it only steers this heuristic.
  if (*(volatile unsigned long long*)b.n_have_host * 4 >= b.n_lines * 3)
  {
    CK(launch_fetch_lines(b.line_want, b.line_have, b.staged_src, b.owned_field, b.n_lines, field_bytes(b.dev),
                          true, b.n_have_dev, ctx->sm_count, s));
    ctx->launches++;
    b.all_resident = true;
    return VR_OK;
  }
  TraceParams m = p;
  m.mark = b.line_want;
  m.tile_counter = counter;
  m.sample_counter = nullptr;
  m.n_clear_chunks = 0;
  m.tiles_x = (m.sw + 7) / 8; // the plain ray rectangle (mode 2 widens it for its clears)
  m.tx0 = m.sx;
  m.tx1 = m.sx + m.sw;
  CK(launch_trace(m, 4, ctx->sm_count, s, zero_counter));
  CK(launch_fetch_lines(b.line_want, b.line_have, b.staged_src, b.owned_field, b.n_lines, field_bytes(b.dev),
                        false, b.n_have_dev, ctx->sm_count, s));
  CK(cudaMemcpyAsync(b.n_have_host, b.n_have_dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  ctx->launches += 2;
  return VR_OK;
}

// The 3-D tensor map of a block's f32 point field with the brick march's box (sampler.cu: kBrickX/Y/Z):
// cuTensorMapEncodeTiled through cudaGetDriverEntryPoint.  Returns false when the driver cannot encode it.
static const int kBrickBox[3] = { 16, 10, 10 };
static bool encode_brick_tensor_map(const BlockDev& b, unsigned char out[128])
{
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn fn = nullptr;
  static bool looked_up = false;
  if (!looked_up)
  {
    looked_up = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_fn>(sym);
    else
      cudaGetLastError();
  }
  if (!fn) return false;
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
  alignas(64) CUtensorMap m;
  const cuuint64_t dims[3] = { (cuuint64_t)b.dims[0], (cuuint64_t)b.dims[1], (cuuint64_t)b.dims[2] };
  const cuuint64_t strides[2] = { (cuuint64_t)b.dims[0] * 4, (cuuint64_t)b.dims[0] * (cuuint64_t)b.dims[1] * 4 };
  const cuuint32_t box[3] = { (cuuint32_t)kBrickBox[0], (cuuint32_t)kBrickBox[1], (cuuint32_t)kBrickBox[2] };
  const cuuint32_t estr[3] = { 1, 1, 1 };
  const CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(b.field), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  std::memcpy(out, &m, 128);
  return true;
}

static vr_status fill_trace_params(vr_ctx* ctx, int block_id, const vr_camera* cam, float sample_dist,
                                   float range_min, float range_max, int use_depth, int W, int H,
                                   TraceParams& p, bool allow_unstructured = false)
{
  REQUIRE(cam != nullptr, "trace: camera is NULL");
  auto it = ctx->blocks.find(block_id);
  REQUIRE(it != ctx->blocks.end(), "trace: unknown block %d", block_id);
  REQUIRE(ctx->lut_size >= 2, "trace: no transfer function set (vr_set_tf)");
  const Block& b = it->second;
  // the reference renders unstructured domains as partials only (m_has_unstructured forces path B,
  // VolumeRenderer.cpp:874-903)
  REQUIRE(b.dev.kind != 2 || allow_unstructured,
          "trace: block %d is unstructured: render it with vr_trace_to_partials (path B)", block_id);
  std::memset(&p, 0, sizeof(p));
  p.blk = b.dev;
  int sub[4];
  hm::find_subset(*cam, W, H, b.bounds, sub);
  p.W = W; p.H = H;
  p.sx = sub[0]; p.sy = sub[1]; p.sw = sub[2]; p.sh = sub[3];
  p.tx0 = p.sx;
  p.tx1 = p.sx + p.sw;
  p.tiles_x = (p.sw + 7) / 8;
  p.tiles_y = (p.sh + 3) / 4;
  const hm::RayGen g = hm::raygen(*cam, W, H, false);
  for (int k = 0; k < 3; ++k)
  {
    p.origin[k] = cam->position[k];
    p.nlook[k] = g.nlook[k];
    p.delta_x[k] = g.delta_x[k];
    p.delta_y[k] = g.delta_y[k];
    p.bmin[k] = (float)b.bounds[2 * k];
    p.bmax[k] = (float)b.bounds[2 * k + 1];
  }
  const hm::Mat4 pv = hm::projview(*cam, W, H);
  std::memcpy(p.pv, pv.m, sizeof(p.pv));
  p.use_depth = use_depth ? 1 : 0;
  if (use_depth)
  {
    const hm::Mat4 inv = hm::inverse(pv);
    std::memcpy(p.inv_pv, inv.m, sizeof(p.inv_pv));
  }
  p.dbl_inv_w = 2.f / (float)W;
  p.dbl_inv_h = 2.f / (float)H;
  // first sample at entry + meshEpsilon = |block extent| * 1e-4 (VolumeRendererStructured::RenderOnDevice of the
  // pinned VTK-m; vr_set_first_sample_offset selects the older generation's 1e-4 absolute)
  const hm::Vec3 ext = { { (float)(b.bounds[1] - b.bounds[0]), (float)(b.bounds[3] - b.bounds[2]),
                           (float)(b.bounds[5] - b.bounds[4]) } };
  const float mag = hm::magnitude(ext);
  p.mesh_eps = ctx->first_sample_abs + ctx->first_sample_rel * mag;
  p.sample_dist = sample_dist > 0.f ? sample_dist : mag / 200.f;
  {
    // every sampling loop ends because the distance grows: a step that no longer changes a float of the size of
    // the distance to the far side of the block (zero extent with no sample distance given, or a sample count in
    // the millions) would spin on the device for ever -- refuse it here
    float far2 = 0.f;
    for (int k = 0; k < 3; ++k)
    {
      const float a = std::fabs((float)b.bounds[2 * k] - cam->position[k]);
      const float c = std::fabs((float)b.bounds[2 * k + 1] - cam->position[k]);
      far2 += std::max(a, c) * std::max(a, c);
    }
    const float far_side = std::sqrt(far2);
    REQUIRE(std::isfinite(p.sample_dist) && p.sample_dist > 0.f && !(p.sample_dist < far_side * 2.5e-7f), // (two ulps of the largest distance)
            "trace: sample distance %g is too small for a block %g away from the camera", (double)p.sample_dist,
            (double)far_side);
  }
  p.range_min = range_min;
  p.inv_delta_scalar = (range_max - range_min) != 0.f ? 1.f / (range_max - range_min) : range_min;
  p.march = 0;
  if (b.dev.kind == 0)
  {
    float min_inv = b.dev.inv_spacing[0], min_sp = b.dev.spacing[0], big = 0.f;
    for (int k = 0; k < 3; ++k)
    {
      p.dims_m1[k] = (float)(b.dev.dims[k] - 1);
      p.dims_m2[k] = (float)(b.dev.dims[k] - 2);
      min_inv = std::min(min_inv, b.dev.inv_spacing[k]);
      min_sp = std::min(min_sp, b.dev.spacing[k]);
      big = std::max(big, std::max(std::fabs(b.dev.min_point[k]), std::fabs(b.dev.max_point[k])));
    }
    // Along its dominant axis a unit direction has |d_k| >= 1/sqrt(3), so a step of >= 2.6 of the LARGEST
    // voxel edge moves every ray by >= 1.5 cells on some axis.  With the previous sample inside its cell
    // (0 <= t <= 1) the next one then lies outside [0,1] on that axis by half a cell, while the f32
    // rounding of p += step is below ulp(|p|) <= 1.2e-7 * |block coordinates| -- less than 1 % of a cell
    // when the smallest edge is >= 1e-4 of the coordinates' magnitude.  The reference therefore takes its
    // "new cell" branch on every sample, which is all the sparse march does.
    const bool well_conditioned = min_sp >= big * 1e-4f && b.dev.dims[0] < (1 << 22) && b.dev.dims[1] < (1 << 22) &&
                                  b.dev.dims[2] < (1 << 22);
    const long long slice = (long long)b.dev.dims[0] * b.dev.dims[1];
    p.slice_elems = slice < (1ll << 31) ? (int)slice : 0;
    if (b.dev.assoc == VR_POINT && well_conditioned && p.slice_elems > 0 && p.sample_dist * min_inv >= 2.6f &&
        !ctx->no_sparse)
      p.march = 1;
    // Dense steps: stage bricks of the field in shared memory with TMA bulk copies.  Needs f32 scalars whose
    // x-rows start on 16-byte boundaries (row length a multiple of 4, 16-byte aligned base) and, to pay off,
    // a step of at most two of the SMALLEST voxel edges (neighbouring samples then share cells / rows).
    float max_inv = b.dev.inv_spacing[0];
    for (int k = 1; k < 3; ++k) max_inv = std::max(max_inv, b.dev.inv_spacing[k]);
    if (p.march == 0 && b.dev.assoc == VR_POINT && b.dev.dtype == VR_F32 && b.dev.dims[0] % 4 == 0 &&
        (reinterpret_cast<uintptr_t>(b.dev.field) & 15) == 0 && p.sample_dist * max_inv <= 2.0f && !ctx->no_brick &&
        !b.staged_src && b.dev.tmap)
      p.march = 2;
  }
  p.lut = ctx->lut;
  p.lut_size = ctx->lut_size;
  p.cms_f = (float)(ctx->lut_size - 1);
  p.canvas_rgba = ctx->canvas_rgba;
  p.canvas_depth = ctx->canvas_depth;
  p.tile_counter = ctx->tile_counter;
  p.sample_counter = ctx->count_samples ? ctx->sample_counter : nullptr;
  p.ctas_per_sm = ctx->ctas_per_sm;
  p.tile_order = ctx->tile_order;
  return VR_OK;
}

extern "C" vr_status vr_trace_to_canvas(vr_ctx* ctx, int block_id, const vr_camera* cam,
                                        float sample_dist, float range_min, float range_max,
                                        int use_canvas_depth)
{
  VR_ENTER(ctx);
  REQUIRE(ctx->W > 0, "vr_trace_to_canvas: call vr_canvas_clear/upload first");
  CK(cudaSetDevice(ctx->device));
  TraceParams p;
  vr_status st = fill_trace_params(ctx, block_id, cam, sample_dist, range_min, range_max,
                                   use_canvas_depth, ctx->W, ctx->H, p);
  if (st != VR_OK) return st;
  st = stage_for_trace(ctx, block_id, p, ctx->stream, ctx->tile_counter, true);
  if (st != VR_OK) return st;
  CK(launch_trace(p, 0, ctx->sm_count, ctx->stream));
  ctx->launches++;
  return VR_OK;
}

extern "C" vr_status vr_render_image(vr_ctx* ctx, int block_id, const vr_camera* cam, int width,
                                     int height, float sample_dist, float range_min,
                                     float range_max, float* rgba_inout, float* depth_inout)
{
  VR_ENTER(ctx);
  REQUIRE(rgba_inout && depth_inout, "vr_render_image: NULL canvas");
  vr_status st = vr_canvas_upload(ctx, width, height, rgba_inout, depth_inout);
  if (st != VR_OK) return st;
  st = vr_trace_to_canvas(ctx, block_id, cam, sample_dist, range_min, range_max, 1);
  if (st != VR_OK) return st;
  return vr_canvas_download(ctx, rgba_inout, depth_inout);
}

extern "C" vr_status vr_trace_to_image(vr_ctx* ctx, int block_id, const vr_camera* cam, int width,
                                       int height, float sample_dist, float range_min,
                                       float range_max, int flags)
{
  VR_ENTER_NOJOIN(ctx);
  const bool ahead = (flags & VR_FRAME_AHEAD) != 0;
  CK(cudaSetDevice(ctx->device));
  // Where the launch goes.  A frame that also writes the canvas runs on the context's stream, after
  // everything the latest exchange does to the canvas.  An image-only frame of a connected context runs on
  // one of a few side streams (round-robin by frame): it neither waits for the previous frame's trace nor for the
  // exchanges still in flight, only for the exchange that frees its ring slot (comm_join_for_image_trace).
  cudaStream_t ts = ctx->stream;
  int side = -1;
  if (flags & VR_FRAME_WRITE_CANVAS)
  {
    ++ctx->api_serial;
    VR_JOIN(ctx);
  }
  else if (ctx->comm.on && ctx->comm.trace_streams > 0 && ctx->comm.tstream[0])
  {
    auto it = ctx->blocks.find(block_id);
    const bool staging = it != ctx->blocks.end() && it->second.staged_src && !it->second.all_resident;
    if (!staging) // (demand staging keeps per-block state on the device: one trace of the block at a time)
    {
      side = (int)((ctx->comm.epoch + (ahead ? 2u : 1u)) % (unsigned)ctx->comm.trace_streams);
      ts = ctx->comm.tstream[side];
      CK(cudaEventRecord(ctx->comm.ev_main, ctx->stream)); // after what the caller queued so far (publishes, ...)
      CK(cudaStreamWaitEvent(ts, ctx->comm.ev_main, 0));
    }
  }
  if (!(flags & VR_FRAME_WRITE_CANVAS)) comm_join_for_image_trace(ctx, ahead, ts);
  vr_status st = ensure_frame(ctx, width, height);
  if (st != VR_OK) return st;
  TraceParams p;
  st = fill_trace_params(ctx, block_id, cam, sample_dist, range_min, range_max, 0, width, height, p);
  if (st != VR_OK) return st;
  if (side >= 0) p.tile_counter = ctx->tile_counter + 1 + vr::kMaxLayers + side;
  p.vec_ok = (width % 4 == 0) ? 1 : 0;
  if (p.vec_ok && p.sw > 0)
  {
    p.tx0 = p.sx & ~3;
    p.tx1 = std::min(width, (p.sx + p.sw + 3) & ~3);
    p.tiles_x = (p.tx1 - p.tx0 + 7) / 8;
  }
  p.img_rgba = ctx->img_rgba;
  p.img_depth = ctx->img_depth;
  if (ahead)
  {
    // the frame AFTER the one whose exchange is still to be issued: next slot of the image ring
    REQUIRE(!(flags & VR_FRAME_WRITE_CANVAS), "vr_trace_to_image: VR_FRAME_AHEAD cannot write the canvas "
            "(it belongs to the frame being exchanged)");
    REQUIRE(!ctx->img_ahead, "vr_trace_to_image: one frame is already traced ahead; exchange a frame first");
    st = comm_ahead_image(ctx, &p.img_rgba, &p.img_depth);
    if (st != VR_OK) return st;
  }
  const bool push = (flags & VR_FRAME_PUSH) != 0;
  if (push)
  {
    // the image is scattered to the exchange's owners as it is produced (sampler mode 5)
    REQUIRE(!(flags & VR_FRAME_WRITE_CANVAS), "vr_trace_to_image: VR_FRAME_PUSH cannot write the canvas");
    REQUIRE(p.vec_ok, "vr_trace_to_image: VR_FRAME_PUSH needs a width that is a multiple of 4");
    st = comm_push_target(ctx, ahead, width, height, p);
    if (st != VR_OK) return st;
  }
  p.write_canvas = (flags & VR_FRAME_WRITE_CANVAS) ? 1 : 0;
  // (a width that is not a multiple of 4 has no announced rectangle -- the exchange reads the whole
  // image -- so such frames are always cleared)
  p.n_clear_chunks = (((flags & VR_FRAME_NO_CLEAR) && p.vec_ok) || push) ? 0 : (int)(((size_t)width * height + 511) / 512);
  p.end_stamp = comm_timeline_slot(ctx, 6);
  if (p.end_stamp) CK(cudaMemsetAsync(p.end_stamp, 0, sizeof(unsigned long long), ts));
  st = stage_for_trace(ctx, block_id, p, ts, p.tile_counter, true);
  if (st != VR_OK) return st;
  CK(launch_trace(p, push ? 5 : 2, ctx->sm_count, ts));
  ctx->launches++;
  if (side >= 0)
  {
    CK(cudaEventRecord(ctx->comm.ev_t[side], ts));
    ctx->comm.t_pending[side] = true;
  }
  // what the multi-GPU fold needs to know: outside this rectangle my image is empty
  int* rect = ahead ? ctx->img_rect_ahead : ctx->img_rect;
  rect[0] = p.tx0; rect[1] = p.sy;
  rect[2] = p.tx1; rect[3] = p.sy + p.sh;
  if (p.sw <= 0 || p.sh <= 0) rect[0] = rect[1] = rect[2] = rect[3] = 0;
  if (ahead) { ctx->img_ahead = true; ctx->img_pushed_ahead = push; }
  else ctx->img_pushed = push;
  return VR_OK;
}

// ----------------------------------------------------------------- partial list
static vr_status ensure_partials(vr_ctx* ctx, size_t need)
{
  if (need <= ctx->partial_cap) return VR_OK;
  size_t cap = std::max(need, ctx->partial_cap * 2);
  vr_partial* np = nullptr;
  CK(cudaMalloc(&np, cap * sizeof(vr_partial)));
  if (ctx->partials)
  {
    unsigned long long n = 0;
    CK(cudaMemcpyAsync(&n, ctx->partial_count, sizeof(n), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    n = std::min<unsigned long long>(n, ctx->partial_cap);
    if (n) CK(cudaMemcpyAsync(np, ctx->partials, n * sizeof(vr_partial), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->partials);
  }
  ctx->partials = np;
  ctx->partial_cap = cap;
  return VR_OK;
}

extern "C" vr_status vr_partials_begin(vr_ctx* ctx, int width, int height)
{
  VR_ENTER(ctx);
  REQUIRE(width > 0 && height > 0 && (long long)width * height < (1ll << 31), "bad image size");
  CK(cudaSetDevice(ctx->device));
  ctx->pW = width;
  ctx->pH = height;
  ctx->n_partials_host = 0;
  ctx->plist = nullptr; // read side sees the own list again
  CK(cudaMemsetAsync(ctx->partial_count, 0, sizeof(unsigned long long), ctx->stream));
  return VR_OK;
}

extern "C" vr_status vr_trace_to_partials(vr_ctx* ctx, int block_id, const vr_camera* cam,
                                          float sample_dist, float range_min, float range_max,
                                          int use_canvas_depth)
{
  VR_ENTER(ctx);
  REQUIRE(ctx->pW > 0, "vr_trace_to_partials: call vr_partials_begin first");
  REQUIRE(!use_canvas_depth || (ctx->W == ctx->pW && ctx->H == ctx->pH),
          "vr_trace_to_partials: canvas depth requested but canvas size differs");
  CK(cudaSetDevice(ctx->device));
  TraceParams p;
  vr_status st = fill_trace_params(ctx, block_id, cam, sample_dist, range_min, range_max,
                                   use_canvas_depth, ctx->pW, ctx->pH, p, true);
  if (st != VR_OK) return st;
  if (p.blk.kind == 2)
  {
    // UnstructuredWrapper::render + vtkm_to_partials (VolumeRenderer.cpp:141-221)
    ctx->n_partials_host += (size_t)p.sw * p.sh;
    st = ensure_partials(ctx, ctx->n_partials_host);
    if (st != VR_OK) return st;
    p.partials = ctx->partials;
    p.partial_count = ctx->partial_count;
    p.partial_capacity = ctx->partial_cap;
    CK(launch_utrace_partials(p, ctx->blocks[block_id].um, ctx->sm_count, ctx->stream));
    ctx->launches++;
    return VR_OK;
  }
  // worst case every ray of the subset emits: reserve before launching (no overflow possible)
  // the list can only be bounded from the host by the sum of subset sizes so far
  static_assert(sizeof(vr_partial) == 24, "vr_partial must be 24 bytes");
  ctx->n_partials_host += (size_t)p.sw * p.sh; // upper bound bookkeeping
  st = ensure_partials(ctx, ctx->n_partials_host);
  if (st != VR_OK) return st;
  p.partials = ctx->partials;
  p.partial_count = ctx->partial_count;
  p.partial_capacity = ctx->partial_cap;
  st = stage_for_trace(ctx, block_id, p, ctx->stream, ctx->tile_counter, true);
  if (st != VR_OK) return st;
  CK(launch_trace(p, 1, ctx->sm_count, ctx->stream));
  ctx->launches++;
  return VR_OK;
}

// Partials of ANOTHER producer (Devil Ray's volume integrator, dray/rendering/renderer.cpp:309-323, converts its
// own to apcomp::VolumePartial<float> -- the same 24-byte POD -- before PartialCompositor::composite) join the
// frame's list and are composited with everything vr_trace_to_partials emitted.
extern "C" vr_status vr_partials_append(vr_ctx* ctx, const vr_partial* partials, size_t n, int where)
{
  VR_ENTER(ctx);
  REQUIRE(ctx->pW > 0, "vr_partials_append: call vr_partials_begin first");
  REQUIRE(partials || n == 0, "vr_partials_append: NULL list");
  REQUIRE(where == VR_HOST || where == VR_DEVICE, "vr_partials_append: where must be VR_HOST or VR_DEVICE");
  if (n == 0) return VR_OK;
  CK(cudaSetDevice(ctx->device));
  const long long n_px = (long long)ctx->pW * ctx->pH;
  if (where == VR_HOST)
    for (size_t i = 0; i < n; ++i)
      REQUIRE(partials[i].pixel_id >= 0 && (long long)partials[i].pixel_id < n_px,
              "vr_partials_append: pixel id %d outside %dx%d", partials[i].pixel_id, ctx->pW, ctx->pH);
  ctx->n_partials_host += n;
  vr_status st = ensure_partials(ctx, ctx->n_partials_host);
  if (st != VR_OK) return st;
  const vr_partial* src = partials;
  vr_partial* tmp = nullptr;
  if (where == VR_HOST)
  {
    CK(cudaMalloc(&tmp, n * sizeof(vr_partial)));
    cudaError_t e = cudaMemcpyAsync(tmp, partials, n * sizeof(vr_partial), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { cudaFree(tmp); return fail(ctx, VR_ERR_CUDA, "vr_partials_append: %s", cudaGetErrorString(e)); }
    src = tmp;
  }
  cudaError_t e = launch_partial_append(src, n, ctx->partials, ctx->partial_count, ctx->partial_cap, ctx->stream);
  ctx->launches++;
  if (tmp)
  {
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); // (the caller's list may be reused; tmp is freed)
    cudaFree(tmp);
  }
  if (e != cudaSuccess) return fail(ctx, VR_ERR_CUDA, "vr_partials_append: %s", cudaGetErrorString(e));
  return VR_OK;
}

extern "C" vr_status vr_partials_count(vr_ctx* ctx, size_t* n)
{
  VR_ENTER_RO(ctx);
  REQUIRE(n, "vr_partials_count: NULL");
  unsigned long long c = 0;
  const unsigned long long* src = ctx->plist ? ctx->plist_count : ctx->partial_count;
  const size_t cap = ctx->plist ? ctx->plist_cap : ctx->partial_cap;
  CK(cudaMemcpyAsync(&c, src, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *n = (size_t)std::min<unsigned long long>(c, cap);
  return VR_OK;
}

extern "C" vr_status vr_partials_download(vr_ctx* ctx, vr_partial* out, size_t capacity, size_t* n)
{
  VR_ENTER_RO(ctx);
  size_t c = 0;
  vr_status st = vr_partials_count(ctx, &c);
  if (st != VR_OK) return st;
  if (n) *n = c;
  REQUIRE(c <= capacity, "vr_partials_download: capacity %zu < %zu partials", capacity, c);
  if (c && out)
  {
    const vr_partial* src = ctx->plist ? ctx->plist : ctx->partials;
    CK(cudaMemcpyAsync(out, src, c * sizeof(vr_partial), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return VR_OK;
}

extern "C" vr_status vr_render_partials(vr_ctx* ctx, int block_id, const vr_camera* cam, int width,
                                        int height, float sample_dist, float range_min,
                                        float range_max, const float* depth_in, vr_partial** out,
                                        size_t* n)
{
  VR_ENTER(ctx);
  REQUIRE(out && n, "vr_render_partials: NULL output");
  *out = nullptr;
  *n = 0;
  vr_status st;
  if (depth_in)
  {
    CK(cudaSetDevice(ctx->device));
    st = ensure_frame(ctx, width, height);
    if (st != VR_OK) return st;
    CK(cudaMemcpyAsync(ctx->canvas_depth, depth_in, (size_t)width * height * sizeof(float),
                       cudaMemcpyHostToDevice, ctx->stream));
  }
  st = vr_partials_begin(ctx, width, height);
  if (st != VR_OK) return st;
  ctx->n_partials_host = 0;
  st = vr_trace_to_partials(ctx, block_id, cam, sample_dist, range_min, range_max, depth_in != nullptr);
  if (st != VR_OK) return st;
  size_t c = 0;
  st = vr_partials_count(ctx, &c);
  if (st != VR_OK) return st;
  vr_partial* host = (vr_partial*)std::malloc(std::max<size_t>(c, 1) * sizeof(vr_partial));
  if (!host) return fail(ctx, VR_ERR_NOMEM, "vr_render_partials: host allocation failed");
  st = vr_partials_download(ctx, host, c, &c);
  if (st != VR_OK) { std::free(host); return st; }
  *out = host;
  *n = c;
  return VR_OK;
}

extern "C" void vr_free(void* p) { std::free(p); }

// ================================================================= ray layers (path B without lists)
namespace vr { vr_status comm_bind_layers(vr_ctx* ctx); }
namespace vr { void comm_layer_push_target(vr_ctx* ctx, vr::TraceParams& p); }

static vr_status ensure_layer_pool(vr_ctx* ctx, size_t need)
{
  if (need <= ctx->lpool_cap) return VR_OK;
  if (ctx->layers_in_arena)
  {
    // rank-local failure inside a collective frame: remember it, so that this rank's call of the layer
    // exchange releases its peers (abort) instead of leaving them waiting
    ctx->comm.frame_poisoned = true;
    return fail(ctx, VR_ERR_INVALID, "ray layers need %zu entries this frame but max_partials = %zu (vr_comm_init)", need,
                ctx->lpool_cap);
  }
  const size_t cap = std::max(need, ctx->lpool_cap * 2);
  float4* nr = nullptr;
  float* nd = nullptr;
  CK(cudaMalloc(&nr, cap * sizeof(float4)));
  CK(cudaMalloc(&nd, cap * sizeof(float)));
  if (ctx->lpool_used)
  {
    // layers traced earlier in this frame move along (rare: only while the pool is still growing)
    CK(cudaMemcpyAsync(nr, ctx->lpool_rgba, ctx->lpool_used * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(nd, ctx->lpool_depth, ctx->lpool_used * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(ctx->lpool_rgba);
  cudaFree(ctx->lpool_depth);
  ctx->lpool_rgba = ctx->own_lpool_rgba[ctx->own_lslot] = nr;
  ctx->lpool_depth = ctx->own_lpool_depth[ctx->own_lslot] = nd;
  ctx->lpool_cap = ctx->own_lpool_cap[ctx->own_lslot] = cap;
  return VR_OK;
}

extern "C" vr_status vr_layers_begin(vr_ctx* ctx, int width, int height)
{
  VR_ENTER_NOJOIN(ctx); // (touches only the next frame's layer buffer: may overlap the latest exchange)
  ++ctx->api_serial;
  vr::comm_join_previous_exchange(ctx);
  REQUIRE(width > 0 && height > 0 && (long long)width * height < (1ll << 31), "bad image size");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->ltab_host) ctx->ltab_host = new LayerTable(); // pageable on purpose: async copies stage it
  if (ctx->comm.on && ctx->comm.max_partials)
  {
    vr_status st = comm_bind_layers(ctx);
    if (st != VR_OK) return st;
  }
  else
  {
    // the next buffer of the context's own ring (its last reader, the fold of three frames ago, is done:
    // comm_join_previous_exchange above)
    ctx->layers_pushed = false;
    const int k = ctx->own_lslot = (ctx->own_lslot + 1) % vr_ctx::kOwnLayerRing;
    if (!ctx->own_ltab[k]) CK(cudaMalloc(&ctx->own_ltab[k], sizeof(LayerTable)));
    ctx->ltab = ctx->own_ltab[k];
    ctx->lpool_rgba = ctx->own_lpool_rgba[k];
    ctx->lpool_depth = ctx->own_lpool_depth[k];
    ctx->lpool_cap = ctx->own_lpool_cap[k];
  }
  ctx->ltab_host->n = 0;
  ctx->lpool_used = 0;
  ctx->lW = width;
  ctx->lH = height;
  return VR_OK;
}

extern "C" vr_status vr_trace_to_layer(vr_ctx* ctx, int block_id, const vr_camera* cam, float sample_dist,
                                       float range_min, float range_max, int use_canvas_depth)
{
  VR_ENTER_NOJOIN(ctx);
  ++ctx->api_serial;
  if (use_canvas_depth) VR_JOIN(ctx); // reads the canvas the latest exchange may still be writing
  else vr::comm_join_previous_exchange(ctx);
  REQUIRE(ctx->lW > 0, "vr_trace_to_layer: call vr_layers_begin first");
  REQUIRE(!use_canvas_depth || (ctx->W == ctx->lW && ctx->H == ctx->lH),
          "vr_trace_to_layer: canvas depth requested but canvas size differs");
  CK(cudaSetDevice(ctx->device));
  TraceParams p;
  vr_status st = fill_trace_params(ctx, block_id, cam, sample_dist, range_min, range_max, use_canvas_depth,
                                   ctx->lW, ctx->lH, p);
  if (st != VR_OK) return st;
  if (p.sw <= 0 || p.sh <= 0) return VR_OK; // block off screen: no rays, no layer
  LayerTable& T = *ctx->ltab_host;
  REQUIRE(T.n < kMaxLayers, "vr_trace_to_layer: more than %d layers in one frame", kMaxLayers);
  const size_t area = (size_t)p.sw * p.sh;
  // keep 16-byte alignment of every layer's first float4/float4-group
  const size_t base = (ctx->lpool_used + 3) & ~(size_t)3;
  st = ensure_layer_pool(ctx, base + area);
  if (st != VR_OK) return st;
  p.layer_rgba = ctx->lpool_rgba;
  p.layer_depth = ctx->lpool_depth;
  p.layer_base = base;
  comm_layer_push_target(ctx, p);
  st = stage_for_trace(ctx, block_id, p, ctx->stream, ctx->tile_counter, true);
  if (st != VR_OK) return st;
  CK(launch_trace(p, 3, ctx->sm_count, ctx->stream));
  ctx->launches++;
  LayerDesc& d = T.d[T.n++];
  d.x0 = p.sx; d.y0 = p.sy; d.w = p.sw; d.h = p.sh;
  d.base = base;
  ctx->lpool_used = base + area;
  return VR_OK;
}

// Every local block of the frame in one call (the loop of RenderMultipleDomainsPerRank,
// VolumeRenderer.cpp:557-578).  The launches are spread over a few side streams: each sampler launch
// is a persistent grid that fills the GPU, so the next block's CTAs move in as the previous block's
// long rays drain -- no idle tail between blocks -- and the host pays one counter memset per frame.
extern "C" vr_status vr_trace_blocks_to_layers(vr_ctx* ctx, int n_blocks, const int* block_ids,
                                               const vr_camera* cam, float sample_dist, float range_min,
                                               float range_max, int use_canvas_depth)
{
  VR_ENTER_NOJOIN(ctx);
  ++ctx->api_serial;
  if (use_canvas_depth) VR_JOIN(ctx);
  else vr::comm_join_previous_exchange(ctx);
  REQUIRE(ctx->lW > 0, "vr_trace_blocks_to_layers: call vr_layers_begin first");
  REQUIRE(n_blocks >= 0 && (n_blocks == 0 || block_ids), "vr_trace_blocks_to_layers: NULL block list");
  REQUIRE(!use_canvas_depth || (ctx->W == ctx->lW && ctx->H == ctx->lH),
          "vr_trace_blocks_to_layers: canvas depth requested but canvas size differs");
  CK(cudaSetDevice(ctx->device));
  LayerTable& T = *ctx->ltab_host;
  std::vector<TraceParams> ps;
  std::vector<int> block_of;
  ps.reserve(n_blocks);
  size_t used = ctx->lpool_used;
  for (int k = 0; k < n_blocks; ++k)
  {
    TraceParams p;
    vr_status st = fill_trace_params(ctx, block_ids[k], cam, sample_dist, range_min, range_max, use_canvas_depth,
                                     ctx->lW, ctx->lH, p);
    if (st != VR_OK) return st;
    if (p.sw <= 0 || p.sh <= 0) continue; // block off screen: no rays, no layer
    const size_t base = (used + 3) & ~(size_t)3;
    p.layer_base = base;
    used = base + (size_t)p.sw * p.sh;
    ps.push_back(p);
    block_of.push_back(block_ids[k]);
  }
  REQUIRE(T.n + (int)ps.size() <= kMaxLayers, "vr_trace_blocks_to_layers: more than %d layers in one frame",
          kMaxLayers);
  if (ps.empty()) return VR_OK;
  vr_status st = ensure_layer_pool(ctx, used);
  if (st != VR_OK) return st;
  const int n = (int)ps.size();
  // ---- many blocks: ONE persistent launch over the (block, tile) work items of all of them (sampler.cu,
  // trace_multi_kernel) when they all run the same kernel variant; otherwise one launch per block, below
  {
    // (VR_MULTI_MIN: smallest batch that takes this route, 0 = never; read per call so that tests can switch it)
    const char* mm = std::getenv("VR_MULTI_MIN");
    const int multi_min = mm ? std::atoi(mm) : 16;
    bool multi = multi_min > 0 && n >= multi_min;
    if (multi)
      for (int k = 0; k < n; ++k)
      {
        // short ray segments (a 128^3 block at the default sampling: ~14 samples per ray): the three-deep pipeline of
        // the sparse march never fills, and its 80 registers cost residency -- the general march (64 registers, 8 CTAs
        // per SM) is faster there (c5 on B200: render 2.09 vs 2.44 ms, profiles/r2_v20_c5_*.json).  Same bits either way
        // (tests/test_gpu_marches.py).
        TraceParams& q = ps[k];
        const float ex = q.bmax[0] - q.bmin[0], ey = q.bmax[1] - q.bmin[1], ez = q.bmax[2] - q.bmin[2];
        if (q.march == 1 && std::sqrt(ex * ex + ey * ey + ez * ez) < 32.f * q.sample_dist) q.march = 0;
      }
    for (int k = 0; k < n && multi; ++k)
    {
      const BlockDev& d = ps[k].blk;
      const BlockDev& d0 = ps[0].blk;
      const long long pts = (long long)d.dims[0] * d.dims[1] * d.dims[2];
      multi = d.kind == d0.kind && d.dtype == d0.dtype && d.assoc == d0.assoc && pts < (1ll << 31) &&
              ps[k].march == ps[0].march && ps[k].march <= 1 && !ctx->blocks[block_of[k]].staged_src;
    }
    if (multi)
    {
      std::vector<unsigned> tile_end(n);
      unsigned long long total = 0;
      for (int k = 0; k < n; ++k)
      {
        TraceParams& p = ps[k];
        p.layer_rgba = ctx->lpool_rgba;
        p.layer_depth = ctx->lpool_depth;
        comm_layer_push_target(ctx, p);
        p.tile_counter = nullptr;
        total += (unsigned long long)p.tiles_x * p.tiles_y;
        tile_end[k] = (unsigned)total;
      }
      if (total < (1ull << 32))
      {
        if (ctx->multi_cap < n)
        {
          CK(cudaStreamSynchronize(ctx->stream));
          free_multi(ctx);
          const int cap = std::max(n, 64);
          CK(cudaMalloc(&ctx->multi_table, (size_t)cap * sizeof(TraceParams)));
          CK(cudaMalloc(&ctx->multi_tile_end, (size_t)cap * sizeof(unsigned)));
          // pinned staging, a ring of kMultiSlots frames: the copies below never make the host wait for the GPU
          CK(cudaMallocHost(&ctx->multi_host, (size_t)vr_ctx::kMultiSlots * cap * (sizeof(TraceParams) + sizeof(unsigned))));
          for (int k = 0; k < vr_ctx::kMultiSlots; ++k) CK(cudaEventCreateWithFlags(&ctx->multi_ev[k], cudaEventDisableTiming));
          ctx->multi_cap = cap;
          ctx->multi_slot = 0;
        }
        const int slot = ctx->multi_slot++ % vr_ctx::kMultiSlots;
        CK(cudaEventSynchronize(ctx->multi_ev[slot])); // (the copy issued kMultiSlots batches ago: long done)
        unsigned char* stage = ctx->multi_host + (size_t)slot * ctx->multi_cap * (sizeof(TraceParams) + sizeof(unsigned));
        unsigned char* stage_ends = stage + (size_t)ctx->multi_cap * sizeof(TraceParams);
        std::memcpy(stage, ps.data(), (size_t)n * sizeof(TraceParams));
        std::memcpy(stage_ends, tile_end.data(), (size_t)n * sizeof(unsigned));
        CK(cudaMemcpyAsync(ctx->multi_table, stage, (size_t)n * sizeof(TraceParams), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->multi_tile_end, stage_ends, (size_t)n * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventRecord(ctx->multi_ev[slot], ctx->stream));
        CK(cudaMemsetAsync(ctx->tile_counter + 1, 0, sizeof(unsigned int), ctx->stream));
        CK(launch_trace_multi(ps[0], ctx->multi_table, ctx->multi_tile_end, n, total, ctx->tile_counter + 1, ctx->sm_count,
                              ctx->stream));
        ctx->launches++;
        for (int k = 0; k < n; ++k)
        {
          LayerDesc& d = T.d[T.n++];
          d.x0 = ps[k].sx; d.y0 = ps[k].sy; d.w = ps[k].sw; d.h = ps[k].sh;
          d.base = ps[k].layer_base;
        }
        ctx->lpool_used = used;
        return VR_OK;
      }
    }
  }
  const int n_streams = std::min(n, kAuxStreams);
  CK(cudaMemsetAsync(ctx->tile_counter + 1, 0, (size_t)n * sizeof(unsigned int), ctx->stream));
  CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
  for (int k = 0; k < n_streams; ++k) CK(cudaStreamWaitEvent(ctx->aux[k], ctx->ev_fork, 0));
  for (int k = 0; k < n; ++k)
  {
    TraceParams& p = ps[k];
    p.layer_rgba = ctx->lpool_rgba;
    p.layer_depth = ctx->lpool_depth;
    comm_layer_push_target(ctx, p);
    p.tile_counter = ctx->tile_counter + 1 + k;
    const bool staged = ctx->blocks[block_of[k]].staged_src != nullptr;
    if (staged)
    {
      st = stage_for_trace(ctx, block_of[k], p, ctx->aux[k % n_streams], p.tile_counter, true);
      if (st != VR_OK) return st;
    }
    CK(launch_trace(p, 3, ctx->sm_count, ctx->aux[k % n_streams], staged));
    ctx->launches++;
    LayerDesc& d = T.d[T.n++];
    d.x0 = p.sx; d.y0 = p.sy; d.w = p.sw; d.h = p.sh;
    d.base = p.layer_base;
  }
  for (int k = 0; k < n_streams; ++k)
  {
    CK(cudaEventRecord(ctx->ev_join[k], ctx->aux[k]));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[k], 0));
  }
  ctx->lpool_used = used;
  return VR_OK;
}

static vr_status upload_layer_table(vr_ctx* ctx)
{
  const LayerTable& T = *ctx->ltab_host;
  const size_t bytes = offsetof(LayerTable, d) + (size_t)T.n * sizeof(LayerDesc);
  CK(cudaMemcpyAsync(ctx->ltab, &T, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return VR_OK;
}
namespace vr { vr_status upload_layer_table_pub(vr_ctx* ctx) { return upload_layer_table(ctx); } }
namespace vr { vr_status ensure_frame_pub(vr_ctx* ctx, int W, int H) { return ensure_frame(ctx, W, H); } }

extern "C" vr_status vr_layers_composite_to_canvas(vr_ctx* ctx, const vr_camera* cam, int canvas_is_clear)
{
  // A fold over a cleared canvas only WRITES the canvas: it goes to the exchange stream (after this frame's
  // traces, after the previous fold) and overlaps whatever is traced next -- every entry point that touches
  // the canvas joins it first.  One that blends over the existing canvas stays on the context's stream.
  VR_ENTER_NOJOIN(ctx);
  ++ctx->api_serial;
  // (measured on B200, c3 on one GPU: 0.539 ms per frame overlapped against 0.495 ms one after the other --
  // with every block resident the sampler's launches already keep the SMs' issue slots busy, and the fold's
  // streaming traffic costs them L2 hits -- so the overlap is opt-in: VR_FOLD_OVERLAP=1)
  static const bool overlap = [] { const char* e = std::getenv("VR_FOLD_OVERLAP"); return e && std::atoi(e) != 0; }();
  const bool on_x = overlap && canvas_is_clear && ctx->comm.xstream;
  if (!on_x) VR_JOIN(ctx);
  REQUIRE(cam, "vr_layers_composite_to_canvas: camera is NULL");
  REQUIRE(ctx->lW > 0, "vr_layers_composite_to_canvas: call vr_layers_begin first");
  CK(cudaSetDevice(ctx->device));
  vr_status st;
  if (canvas_is_clear)
  {
    st = ensure_frame(ctx, ctx->lW, ctx->lH);
    if (st != VR_OK) return st;
  }
  REQUIRE(ctx->W == ctx->lW && ctx->H == ctx->lH, "vr_layers_composite_to_canvas: canvas (%dx%d) and layer frame "
          "(%dx%d) differ", ctx->W, ctx->H, ctx->lW, ctx->lH);
  st = upload_layer_table(ctx);
  if (st != VR_OK) return st;
  LayerFoldParams p;
  std::memset(&p, 0, sizeof(p));
  p.rank = 0;
  p.size = 1;
  p.W = ctx->lW;
  p.H = ctx->lH;
  p.clear = canvas_is_clear ? 1 : 0;
  p.smem_layers = std::max(ctx->ltab_host->n, 1);
  p.table[0] = ctx->ltab;
  p.pool_rgba[0] = ctx->lpool_rgba;
  p.pool_depth[0] = ctx->lpool_depth;
  p.canvas_rgba = ctx->canvas_rgba;
  p.canvas_depth = ctx->canvas_depth;
  fill_to_canvas_params(cam, ctx->lW, ctx->lH, p.tp);
  p.light = on_x ? 1 : 0; // alone on the GPU the 32x8 tiles are faster
  cudaStream_t xs = on_x ? ctx->comm.xstream : ctx->stream;
  if (on_x)
  {
    CK(cudaEventRecord(ctx->comm.ev_trace, ctx->stream));
    CK(cudaStreamWaitEvent(xs, ctx->comm.ev_trace, 0));
  }
  CK(launch_layers_fold(p, false, ctx->sm_count, xs));
  ctx->launches++;
  if (on_x)
  {
    ctx->comm.xserial += 1;
    CK(cudaEventRecord(ctx->comm.ev_x[ctx->comm.xserial & 7], xs));
    ctx->comm.x_pending = true;
  }
  return VR_OK;
}

extern "C" vr_status vr_layers_to_partials(vr_ctx* ctx)
{
  VR_ENTER(ctx);
  REQUIRE(ctx->lW > 0, "vr_layers_to_partials: call vr_layers_begin first");
  vr_status st = vr_partials_begin(ctx, ctx->lW, ctx->lH);
  if (st != VR_OK) return st;
  ctx->n_partials_host = ctx->lpool_used;
  st = ensure_partials(ctx, std::max<size_t>(ctx->lpool_used, 1));
  if (st != VR_OK) return st;
  st = upload_layer_table(ctx);
  if (st != VR_OK) return st;
  CK(launch_layers_to_partials(ctx->ltab, ctx->ltab_host->n, ctx->lpool_rgba, ctx->lpool_depth, ctx->lW,
                               ctx->partials, ctx->partial_count, ctx->partial_cap, ctx->stream));
  ctx->launches++;
  return VR_OK;
}

// ================================================================= image compositing
extern "C" vr_status vr_image_from_canvas(vr_ctx* ctx)
{
  VR_ENTER(ctx);
  REQUIRE(ctx->W > 0, "vr_image_from_canvas: no canvas yet");
  CK(cudaSetDevice(ctx->device));
  CK(launch_quantize(ctx->canvas_rgba, ctx->canvas_depth, (size_t)ctx->W * ctx->H, ctx->img_rgba,
                     ctx->img_depth, ctx->stream));
  ctx->launches++;
  ctx->img_rect[0] = ctx->img_rect[1] = 0; // a canvas of unknown history: the whole image counts
  ctx->img_rect[2] = ctx->img_rect[3] = 0x7fffffff;
  return VR_OK;
}

extern "C" vr_status vr_image_download(vr_ctx* ctx, uint8_t* rgba, float* depth)
{
  VR_ENTER_RO(ctx);
  REQUIRE(ctx->W > 0, "vr_image_download: no image yet");
  REQUIRE(!ctx->img_pushed, "vr_image_download: the pending image was pushed to the exchange (VR_FRAME_PUSH)");
  const size_t n = (size_t)ctx->W * ctx->H;
  if (rgba) CK(cudaMemcpyAsync(rgba, ctx->img_rgba, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (depth) CK(cudaMemcpyAsync(depth, ctx->img_depth, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VR_OK;
}

extern "C" vr_status vr_image_ptrs(vr_ctx* ctx, void** rgba8_dev, void** depth_dev)
{
  VR_ENTER(ctx);
  REQUIRE(ctx->W > 0, "vr_image_ptrs: no image yet");
  if (rgba8_dev) *rgba8_dev = ctx->img_rgba;
  if (depth_dev) *depth_dev = ctx->img_depth;
  return VR_OK;
}

// stable ascending order of the layers by vis_order (CompositeOrderSort, Image.hpp:325-331)
static void layer_order(const int* vis_order, int n, int* out)
{
  for (int i = 0; i < n; ++i) out[i] = i;
  std::stable_sort(out, out + n, [&](int a, int b) { return vis_order[a] < vis_order[b]; });
}

extern "C" vr_status vr_fold_images_dev(vr_ctx* ctx, const uint8_t* rgba, const float* depth,
                                        size_t layer_stride_px, const int* vis_order_host,
                                        int n_layers, size_t n_pixels, uint8_t* out_rgba,
                                        float* out_depth)
{
  VR_ENTER(ctx);
  REQUIRE(rgba && depth && vis_order_host && out_rgba && out_depth, "vr_fold_images_dev: NULL argument");
  REQUIRE(n_layers >= 1 && n_layers <= 64, "vr_fold_images_dev: 1..64 layers supported (got %d)", n_layers);
  CK(cudaSetDevice(ctx->device));
  int order[64];
  layer_order(vis_order_host, n_layers, order);
  CK(launch_fold_images((const uchar4*)rgba, depth, layer_stride_px, order, n_layers, n_pixels,
                        (uchar4*)out_rgba, out_depth, ctx->stream));
  ctx->launches++;
  return VR_OK;
}

extern "C" vr_status vr_composite_images(vr_ctx* ctx, const float* rgba, const float* depth,
                                         const int* vis_order, int n_images, int width, int height,
                                         uint8_t* out_rgba, float* out_depth)
{
  VR_ENTER(ctx);
  REQUIRE(rgba && depth && vis_order && out_rgba && out_depth, "vr_composite_images: NULL argument");
  REQUIRE(n_images >= 1 && n_images <= 64, "vr_composite_images: 1..64 images supported");
  CK(cudaSetDevice(ctx->device));
  vr_status st = ensure_frame(ctx, width, height);
  if (st != VR_OK) return st;
  const size_t n = (size_t)width * height;
  uchar4* layers = nullptr;
  float* ldepth = nullptr;
  CK(cudaMalloc(&layers, n * n_images * sizeof(uchar4)));
  cudaError_t e = cudaMalloc(&ldepth, n * n_images * sizeof(float));
  if (e != cudaSuccess) { cudaFree(layers); return fail(ctx, VR_ERR_NOMEM, "vr_composite_images: out of memory"); }
  vr_status rc = VR_OK;
  for (int i = 0; i < n_images && rc == VR_OK; ++i)
  {
    // AddImage(float*...) -> Image::Init: stage each float image through the device canvas
    e = cudaMemcpyAsync(ctx->canvas_rgba, rgba + i * n * 4, n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(ctx->canvas_depth, depth + i * n, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
      e = launch_quantize(ctx->canvas_rgba, ctx->canvas_depth, n, layers + i * n, ldepth + i * n, ctx->stream);
    ctx->launches++;
    if (e != cudaSuccess) rc = fail(ctx, VR_ERR_CUDA, "vr_composite_images: %s", cudaGetErrorString(e));
  }
  if (rc == VR_OK)
  {
    int order[64];
    layer_order(vis_order, n_images, order);
    e = launch_fold_images(layers, ldepth, n, order, n_images, n, ctx->res_rgba, ctx->res_depth, ctx->stream);
    ctx->launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_rgba, ctx->res_rgba, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_depth, ctx->res_depth, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, VR_ERR_CUDA, "vr_composite_images: %s", cudaGetErrorString(e));
  }
  cudaStreamSynchronize(ctx->stream);
  cudaFree(layers);
  cudaFree(ldepth);
  return rc;
}

extern "C" vr_status vr_composite_zbuffer(vr_ctx* ctx, const float* rgba, const float* depth, int n_images,
                                          int width, int height, uint8_t* out_rgba, float* out_depth)
{
  VR_ENTER(ctx);
  REQUIRE(rgba && depth && out_rgba && out_depth, "vr_composite_zbuffer: NULL argument");
  REQUIRE(n_images >= 1, "vr_composite_zbuffer: no images");
  CK(cudaSetDevice(ctx->device));
  vr_status st = ensure_frame(ctx, width, height);
  if (st != VR_OK) return st;
  const size_t n = (size_t)width * height;
  // serial Compositor in Z_BUFFER_SURFACE mode composites as images are added (Compositor.cpp:146-160):
  // the first image initialises the result, every further one is Image::Init'ed and z-selected in
  for (int i = 0; i < n_images; ++i)
  {
    CK(cudaMemcpyAsync(ctx->canvas_rgba, rgba + i * n * 4, n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->canvas_depth, depth + i * n, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (i == 0)
      CK(launch_quantize(ctx->canvas_rgba, ctx->canvas_depth, n, ctx->res_rgba, ctx->res_depth, ctx->stream));
    else
    {
      CK(launch_quantize(ctx->canvas_rgba, ctx->canvas_depth, n, ctx->img_rgba, ctx->img_depth, ctx->stream));
      CK(launch_zbuffer(ctx->res_rgba, ctx->res_depth, ctx->img_rgba, ctx->img_depth, n, ctx->stream));
      ctx->launches++;
    }
    ctx->launches++;
  }
  CK(cudaMemcpyAsync(out_rgba, ctx->res_rgba, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(out_depth, ctx->res_depth, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VR_OK;
}

extern "C" vr_status vr_zbuffer_composite_dev(vr_ctx* ctx, uint8_t* front_rgba, float* front_depth,
                                              const uint8_t* rgba, const float* depth,
                                              size_t n_pixels)
{
  VR_ENTER(ctx);
  REQUIRE(front_rgba && front_depth && rgba && depth, "vr_zbuffer_composite_dev: NULL argument");
  CK(cudaSetDevice(ctx->device));
  CK(launch_zbuffer((uchar4*)front_rgba, front_depth, (const uchar4*)rgba, depth, n_pixels, ctx->stream));
  ctx->launches++;
  return VR_OK;
}

extern "C" vr_status vr_image_to_canvas_dev(vr_ctx* ctx, const uint8_t* rgba, const float* depth)
{
  VR_ENTER(ctx);
  REQUIRE(ctx->W > 0 && rgba && depth, "vr_image_to_canvas_dev: no canvas or NULL argument");
  CK(cudaSetDevice(ctx->device));
  CK(launch_image_to_canvas((const uchar4*)rgba, depth, (size_t)ctx->W * ctx->H, ctx->canvas_rgba,
                            ctx->canvas_depth, ctx->stream));
  ctx->launches++;
  return VR_OK;
}

// ================================================================= partial compositing
static vr_status ensure_partial_scratch(vr_ctx* ctx, size_t n_pixels, size_t n_parts)
{
  if (n_pixels > ctx->scratch_px)
  {
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->px_count); cudaFree(ctx->px_end); cudaFree(ctx->scan_blocks);
    ctx->px_count = ctx->px_end = ctx->scan_blocks = nullptr;
    const size_t padded = partial_scan_padded(n_pixels);
    CK(cudaMalloc(&ctx->px_count, padded * sizeof(int)));
    CK(cudaMalloc(&ctx->px_end, padded * sizeof(int)));
    CK(cudaMalloc(&ctx->scan_blocks, (padded / 1024 + 2) * sizeof(int)));
    ctx->scratch_px = n_pixels;
  }
  if (n_parts > ctx->scratch_parts)
  {
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->sidx);
    cudaFree(ctx->rec);
    ctx->sidx = nullptr;
    ctx->rec = nullptr;
    CK(cudaMalloc(&ctx->sidx, n_parts * sizeof(int)));
    CK(cudaMalloc(&ctx->rec, n_parts * sizeof(vr_partial)));
    ctx->scratch_parts = n_parts;
  }
  if (n_parts > ctx->partial_tmp_cap)
  {
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->partials_tmp);
    ctx->partials_tmp = nullptr;
    CK(cudaMalloc(&ctx->partials_tmp, n_parts * sizeof(vr_partial)));
    ctx->partial_tmp_cap = n_parts;
  }
  return VR_OK;
}

namespace vr
{
vr_status ensure_partial_scratch_pub(vr_ctx* ctx, size_t n_pixels, size_t n_parts)
{
  return ensure_partial_scratch(ctx, n_pixels, n_parts);
}
void fill_to_canvas_params_pub(const vr_camera* cam, int W, int H, ToCanvasParams& tp)
{
  fill_to_canvas_params(cam, W, H, tp);
}
}

static void fill_to_canvas_params(const vr_camera* cam, int W, int H, ToCanvasParams& tp)
{
  const hm::RayGen g = hm::raygen(*cam, W, H, true);
  const hm::Mat4 pv = hm::projview(*cam, W, H);
  for (int k = 0; k < 3; ++k)
  {
    tp.origin[k] = cam->position[k];
    tp.look[k] = g.nlook[k];
    tp.delta_x[k] = g.delta_x[k];
    tp.delta_y[k] = g.delta_y[k];
  }
  std::memcpy(tp.pv, pv.m, sizeof(tp.pv));
  tp.W = W;
  tp.H = H;
}

static vr_status partials_composite_impl(vr_ctx* ctx, const vr_camera* cam, int canvas_is_clear)
{
  REQUIRE(ctx->pW > 0, "vr_partials_composite: call vr_partials_begin first");
  CK(cudaSetDevice(ctx->device));
  // no host sync: every kernel reads the list length from the device counter; the scratch is
  // sized by the list's capacity, which the host knows
  const size_t n_pixels = (size_t)ctx->pW * ctx->pH;
  vr_status st = ensure_partial_scratch(ctx, n_pixels, std::max<size_t>(ctx->partial_cap, 1));
  if (st != VR_OK) return st;
  ToCanvasParams tp;
  if (cam)
  {
    if (canvas_is_clear)
    {
      st = ensure_frame(ctx, ctx->pW, ctx->pH);
      if (st != VR_OK) return st;
    }
    REQUIRE(ctx->W == ctx->pW && ctx->H == ctx->pH, "vr_partials_composite_to_canvas: canvas (%dx%d) and partial "
            "frame (%dx%d) differ", ctx->W, ctx->H, ctx->pW, ctx->pH);
    fill_to_canvas_params(cam, ctx->W, ctx->H, tp);
  }
  PartialScratch sc{ ctx->px_count, ctx->px_end, ctx->sidx, ctx->rec, ctx->scan_blocks };
  cudaError_t e;
  // fold into the tmp list (<= 1 partial per pixel), then swap: the context's list becomes the
  // composited one.  The tmp counter is a second device scalar so the input count stays readable.
  ctx->launches += launch_partials_composite(ctx->partials, ctx->partial_count, ctx->partial_cap, n_pixels,
                                             sc, ctx->partials_tmp, ctx->partial_count_tmp, cam ? &tp : nullptr,
                                             ctx->canvas_rgba, ctx->canvas_depth, canvas_is_clear, ctx->stream, &e);
  CK(e);
  std::swap(ctx->partials, ctx->partials_tmp);
  std::swap(ctx->partial_cap, ctx->partial_tmp_cap);
  std::swap(ctx->partial_count, ctx->partial_count_tmp);
  ctx->n_partials_host = 0;
  ctx->plist = nullptr;
  return VR_OK;
}

extern "C" vr_status vr_partials_composite(vr_ctx* ctx)
{
  VR_ENTER(ctx);
  return partials_composite_impl(ctx, nullptr, 0);
}

extern "C" vr_status vr_partials_composite_to_canvas(vr_ctx* ctx, const vr_camera* cam, int canvas_is_clear)
{
  VR_ENTER(ctx);
  REQUIRE(cam, "vr_partials_composite_to_canvas: camera is NULL");
  return partials_composite_impl(ctx, cam, canvas_is_clear ? 1 : 0);
}

extern "C" vr_status vr_partials_to_canvas(vr_ctx* ctx, const vr_camera* cam)
{
  VR_ENTER(ctx);
  REQUIRE(cam, "vr_partials_to_canvas: camera is NULL");
  REQUIRE(ctx->W == ctx->pW && ctx->H == ctx->pH && ctx->W > 0,
          "vr_partials_to_canvas: canvas (%dx%d) and partial frame (%dx%d) differ", ctx->W, ctx->H,
          ctx->pW, ctx->pH);
  CK(cudaSetDevice(ctx->device));
  ToCanvasParams tp;
  fill_to_canvas_params(cam, ctx->W, ctx->H, tp);
  const vr_partial* list = ctx->plist ? ctx->plist : ctx->partials;
  const unsigned long long* list_count = ctx->plist ? ctx->plist_count : ctx->partial_count;
  const size_t list_cap = ctx->plist ? ctx->plist_cap : ctx->partial_cap;
  if (list_cap == 0) return VR_OK;
  CK(launch_partials_to_canvas(list, list_count, list_cap, tp,
                               ctx->canvas_rgba, ctx->canvas_depth, ctx->stream));
  ctx->launches++;
  return VR_OK;
}

extern "C" vr_status vr_composite_partials(vr_ctx* ctx, const vr_partial* in, size_t n_in, int width,
                                           int height, vr_partial* out, size_t* n_out)
{
  VR_ENTER(ctx);
  REQUIRE(out && n_out && (in || n_in == 0), "vr_composite_partials: NULL argument");
  vr_status st = vr_partials_begin(ctx, width, height);
  if (st != VR_OK) return st;
  for (size_t i = 0; i < n_in; ++i)
    REQUIRE(in[i].pixel_id >= 0 && (long long)in[i].pixel_id < (long long)width * height,
            "vr_composite_partials: pixel id %d outside %dx%d", in[i].pixel_id, width, height);
  ctx->n_partials_host = n_in;
  st = ensure_partials(ctx, std::max<size_t>(n_in, 1));
  if (st != VR_OK) return st;
  if (n_in)
    CK(cudaMemcpyAsync(ctx->partials, in, n_in * sizeof(vr_partial), cudaMemcpyHostToDevice, ctx->stream));
  unsigned long long c = n_in;
  CK(cudaMemcpyAsync(ctx->partial_count, &c, sizeof(c), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  st = vr_partials_composite(ctx);
  if (st != VR_OK) return st;
  return vr_partials_download(ctx, out, n_in, n_out);
}

// ================================================================= host-side helpers
extern "C" vr_status vr_set_first_sample_offset(vr_ctx* ctx, float abs_offset, float extent_rel)
{
  VR_ENTER_RO(ctx);
  REQUIRE(std::isfinite(abs_offset) && std::isfinite(extent_rel) && abs_offset >= 0.f && extent_rel >= 0.f,
          "vr_set_first_sample_offset: offsets must be finite and >= 0");
  ctx->first_sample_abs = abs_offset;
  ctx->first_sample_rel = extent_rel;
  return VR_OK;
}

extern "C" float vr_sample_distance(const double gb[6], float samples)
{
  // VolumeRenderer::PreExecute, VolumeRenderer.cpp:606-611
  const hm::Vec3 ext = { { (float)(gb[1] - gb[0]), (float)(gb[3] - gb[2]), (float)(gb[5] - gb[4]) } };
  return hm::magnitude(ext) / samples;
}

extern "C" void vr_visibility_order(const double* domain_bounds, int n, const vr_camera* cam,
                                    int* order_out)
{
  // FindMinDepth + DepthSort, VolumeRenderer.cpp:637-650,690-831: distance from the camera to each
  // domain's bounds centre, ascending; ties keep (rank, domain) order (stable).
  std::vector<float> depth(n);
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i)
  {
    const double* b = domain_bounds + 6 * i;
    double d2 = 0;
    for (int a = 0; a < 3; ++a)
    {
      const double c = (double)(float)((b[2 * a] + b[2 * a + 1]) / 2.0);
      const double d = c - (double)cam->position[a];
      d2 += d * d;
    }
    depth[i] = (float)std::sqrt(d2);
    idx[i] = i;
  }
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return depth[a] < depth[b]; });
  for (int i = 0; i < n; ++i) order_out[idx[i]] = i;
}

extern "C" int vr_radixk_schedule(int n_ranks, int width, int height, int divisions[2], int* lo_x, int* lo_y, int* seq)
{
  radixk::Schedule s;
  if (!radixk::make_schedule(n_ranks, width, height, s)) return -1;
  if (divisions) { divisions[0] = s.divisions[0]; divisions[1] = s.divisions[1]; }
  for (int i = 0; i < n_ranks; ++i)
  {
    if (lo_x) lo_x[i] = s.lo[0][i];
    if (lo_y) lo_y[i] = s.lo[1][i];
    if (seq)
      for (int k = 0; k < n_ranks; ++k) seq[i * n_ranks + k] = s.seq[i][k];
  }
  return 0;
}

extern "C" void vr_find_subset(const vr_camera* cam, int width, int height, const double bounds[6],
                               int out[4])
{
  hm::find_subset(*cam, width, height, bounds, out);
}

extern "C" void vr_camera_default(vr_camera* cam) { if (cam) hm::camera_default(*cam); }
extern "C" void vr_camera_reset_to_bounds(vr_camera* cam, const double bounds[6])
{
  if (cam && bounds) hm::camera_reset_to_bounds(*cam, bounds);
}
extern "C" void vr_camera_azimuth(vr_camera* cam, float degrees) { if (cam) hm::camera_azimuth(*cam, degrees); }
extern "C" void vr_camera_elevation(vr_camera* cam, float degrees) { if (cam) hm::camera_elevation(*cam, degrees); }
extern "C" void vr_camera_zoom(vr_camera* cam, float zoom) { if (cam) hm::camera_zoom(*cam, zoom); }
extern "C" void vr_camera_cinema(vr_camera* cam, const double bounds[6], float phi_degrees, float theta_degrees)
{
  if (cam && bounds) hm::camera_cinema(*cam, bounds, phi_degrees, theta_degrees);
}

extern "C" vr_status vr_color_table_sample(int color_space, int n_color, const double* color_x,
                                           const float* color_rgb, int n_alpha, const double* alpha_x,
                                           const float* alpha, int n_samples, uint8_t* rgba8_out, float* rgba_out)
{
  if (color_space < 0 || color_space > 2 || n_samples < 2 || n_color < 0 || n_alpha < 0) return VR_ERR_INVALID;
  if ((n_color && (!color_x || !color_rgb)) || (n_alpha && (!alpha_x || !alpha))) return VR_ERR_INVALID;
  ct::Table t;
  t.space = color_space;
  for (int i = 0; i < n_color; ++i)
  {
    if (i && !(color_x[i] > color_x[i - 1])) return VR_ERR_INVALID;
    t.color_x.push_back(color_x[i]);
    t.color.push_back(ct::Rgb{ { color_rgb[3 * i], color_rgb[3 * i + 1], color_rgb[3 * i + 2] } });
  }
  for (int i = 0; i < n_alpha; ++i)
  {
    if (i && !(alpha_x[i] > alpha_x[i - 1])) return VR_ERR_INVALID;
    t.alpha_x.push_back(alpha_x[i]);
    t.alpha.push_back(alpha[i]);
  }
  std::vector<uint8_t> u8((size_t)n_samples * 4);
  ct::sample_u8(t, n_samples, u8.data());
  if (rgba8_out) std::memcpy(rgba8_out, u8.data(), u8.size());
  if (rgba_out)
  {
    const float k = 1.0f / 255.0f; // convert_table's conversionToFloatSpace
    for (size_t i = 0; i < u8.size(); ++i) rgba_out[i] = (float)u8[i] * k;
  }
  return VR_OK;
}

extern "C" float vr_correct_opacity(float alpha, float samples)
{
  const float ratio = 10.f / samples; // VTKH_OPACITY_CORRECTION / samples (VolumeRenderer.cpp:25,453)
  return (float)(1. - std::pow(1. - (double)alpha, (double)ratio));
}

extern "C" vr_status vr_synth_braid_dev(vr_ctx* ctx, void* field_dev, int dtype, const int n[3],
                                        const int start[3], const int global[3])
{
  VR_ENTER_RO(ctx);
  REQUIRE(field_dev && n && start && global, "vr_synth_braid_dev: NULL argument");
  REQUIRE(dtype == VR_F32 || dtype == VR_F64, "vr_synth_braid_dev: bad dtype");
  CK(cudaSetDevice(ctx->device));
  CK(launch_synth_braid(field_dev, dtype, n, start, global, ctx->stream));
  ctx->launches++;
  return VR_OK;
}
