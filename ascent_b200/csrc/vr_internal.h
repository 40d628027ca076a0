// vr_internal.h -- shared declarations of libvr_b200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/vr_b200.h"

namespace vr
{
constexpr int kAuxStreams = 4;

// ---------------------------------------------------------------- device-side descriptors
struct BlockDev
{
  int kind;      // 0 uniform, 1 rectilinear
  int dtype;     // VR_F32 / VR_F64
  int assoc;     // VR_POINT / VR_CELL
  int dims[3];   // point dims
  float origin[3], spacing[3]; // uniform (f32, as Ascent hands them to VTK-m)
  float min_point[3], max_point[3], inv_spacing[3];
  const float* axis[3];        // rectilinear: axes narrowed to f32 once on upload
  const void* field;
  // brick march: the CUtensorMap (128 bytes in device memory, 64-byte aligned) of the field as a 3-D f32
  // tensor with a kBrick^3 box, encoded when the block is published; null when the block does not qualify
  const void* tmap;
};

// An explicit cell set (unstructured.cu): hexahedra (VTK vertex order) or tetrahedra, with the uniform bins that
// locate a sample's cell.  Passed by value as a __grid_constant__.
struct UMeshDev
{
  const float* xyz;   // n_points x 3
  const int* conn;    // n_cells x shape
  const void* field;
  int dtype, assoc;   // VR_F32 / VR_F64, VR_POINT / VR_CELL
  int shape;          // 8 or 4
  int n_cells, n_points;
  float bmin[3], bmax[3], ginv[3];
  int g[3];
  const int* bin_start; // g0 g1 g2 + 1
  const int* bin_cells;
  const unsigned char* ext_mask; // n_cells: bit f = face f is external (vr_umesh_faces.hpp)
  const unsigned char* bin_ext;  // g0 g1 g2: the bin lists a cell with an external face
};

// Everything one trace launch needs; passed by value as a __grid_constant__ (trace_multi_kernel reads a device
// table of them in 16-byte words, hence the alignment).
struct alignas(16) TraceParams
{
  BlockDev blk;
  // K1
  float origin[3], nlook[3], delta_x[3], delta_y[3];
  int W, H, sx, sy, sw, sh;
  int tiles_x, tiles_y;
  int tx0, tx1;        // traced x-range [tx0, tx1) >= [sx, sx+sw): widened to 4-pixel groups in mode 2
  // fused frame (mode 2): uint8 image out, optional float canvas, clear of the complement
  uchar4* img_rgba;
  float* img_depth;
  int write_canvas, vec_ok, n_clear_chunks;
  // K2
  int use_depth;
  float inv_pv[16];
  float dbl_inv_w, dbl_inv_h;
  // K3 (block bounds narrowed to f32)
  float bmin[3], bmax[3];
  // K4 (mesh_eps: the first sample sits at entry + mesh_eps; |extent| * 1e-4 unless vr_set_first_sample_offset says otherwise)
  float mesh_eps, sample_dist, range_min, inv_delta_scalar;
  // uniform blocks: (float)(dims - 1), (float)(dims - 2) for the locator's upper-face fix-up, and which
  // march the sampler runs.  Sparse: EVERY step of EVERY ray is guaranteed to leave its cell (step >= 2.6
  // voxels on a well-conditioned grid), so no sample ever tests for a cell change.  Brick: dense steps
  // (<= 2 voxels) through an f32 point field whose x-rows can be bulk-copied (16-byte aligned).
  float dims_m1[3], dims_m2[3];
  int march;        // 0 general, 1 sparse, 2 brick (TMA-staged): chosen by the host per launch
  int slice_elems;  // dims[0] * dims[1] when that fits 31 bits (sparse march: one z-slice of the field, in elements)
  float cms_f;      // (float)(lut_size - 1)
  int lut_size;
  const float4* lut;
  // K7
  float pv[16];
  float4* canvas_rgba;
  float* canvas_depth;
  // dense layer store (path B without lists, mode 3)
  float4* layer_rgba;
  float* layer_depth;
  unsigned long long layer_base;
  // pushed layers (N ranks): entry e of this rank's layer pool is ALSO stored into the receive pool
  // [source = lpush_rank] of the rank that will fold the entry's pixel -- the owner of its 32x8 screen tile,
  // (tile_y * lpush_tpr + tile_x) % lpush_size -- so that the fold reads local memory only (comm.cu, layers.cu)
  unsigned char* const* lpush_peers; // device table of arena base pointers (null: no push)
  unsigned long long lpush_off_rgba, lpush_off_depth; // byte offsets of this frame's receive pools inside an arena
  unsigned long long lpush_stride;   // entries per source in a receive pool (= max_partials)
  int lpush_rank, lpush_size, lpush_tpr;
  // partial emission (path B)
  vr_partial* partials;
  unsigned long long* partial_count;
  unsigned long long partial_capacity;
  // pushed frame (mode 5): the exchange's receive slots on every rank (comm.cu).  Pixel p of this
  // rank's image goes to rank (p / 1024) % push_size, slot entry
  // push_rank * push_share_px + (p / 1024 / push_size) * 1024 + p % 1024
  unsigned char* const* push_peers; // device table of arena base pointers
  unsigned long long push_off_rgba, push_off_depth; // byte offsets of the receive ring slot inside an arena
  unsigned int push_share_px;       // pixels per (owner, source) slot
  int push_rank, push_size;
  // demand staging pre-pass (mode 4): one byte per 128-byte line of the field
  unsigned char* mark;
  // dynamic tile scheduler + sample counter
  unsigned int* tile_counter;
  unsigned long long* sample_counter; // may be null
  unsigned long long* end_stamp;      // diagnostics: atomicMax of globaltimer at CTA exit (may be null)
  int ctas_per_sm;                    // 0 = default
  int tile_order;                     // 0 = centre-out (default), 1 = row-major (VR_TILE_ORDER, A/B runs)
};

// ---------------------------------------------------------------- host-side state
struct Block
{
  BlockDev dev;
  void* owned_field = nullptr; // device copy we own (null when adopted)
  void* tmap_dev = nullptr;    // the brick march's tensor map (owned)
  float* owned_axes = nullptr;
  double bounds[6];
  // VR_HOST_STAGED: the field stays in mapped host memory; owned_field is a device buffer of the
  // same size that holds only the 128-byte lines some ray has needed since the publish
  // kind 2 (vr_block_unstructured): the mesh; xyz / conn / bins owned unless adopted
  UMeshDev um;
  void* owned_xyz = nullptr;
  void* owned_conn = nullptr;
  int* owned_bin_start = nullptr;
  int* owned_bin_cells = nullptr;
  unsigned char* owned_ext_mask = nullptr;
  unsigned char* owned_bin_ext = nullptr;
  const void* staged_src = nullptr;   // device-visible alias of the host array
  unsigned char* line_want = nullptr; // lines the next trace will touch (pre-pass output)
  unsigned char* line_have = nullptr; // lines already fetched since the publish
  size_t n_lines = 0;
  unsigned long long* n_have_dev = nullptr;  // lines resident (device counter)
  unsigned long long* n_have_host = nullptr; // pinned mirror, refreshed asynchronously after every fetch
  bool all_resident = false;                 // every line is on the device: no more pre-passes
};

struct LayerTable; // layers.cu section below

struct Comm
{
  bool on = false;
  int rank = 0, size = 1;
  size_t max_pixels = 0, max_partials = 0;
  size_t arena_bytes = 0;
  unsigned char* arena = nullptr;               // own arena (cudaMalloc, IPC-exported)
  std::vector<unsigned char*> peer;             // mapped arenas, peer[rank] == arena
  unsigned char** peer_dev = nullptr;           // device copy of the pointer table
  bool local_peers = false;                     // the peers are contexts of this process (vr_comm_connect_local): no IPC mappings
  unsigned int epoch = 0;                       // image path frames composited so far
  unsigned int pepoch = 0;                      // partial path frames
  unsigned int lepoch = 0;                      // layer path frames
  unsigned int sepoch = 0;                      // depth broadcasts
  unsigned long long timeout_ns = 0;            // bound of every cross-rank wait inside the kernels
  // The image exchange runs on its own stream so that it overlaps the NEXT frame's trace (the folded
  // pixels drain into rank 0 over NVLink while the sampler is already busy): ev_trace orders it after the
  // trace that produced the image, ev_x[n & 7] marks the n-th exchange done.  Every entry point first makes the
  // context's stream wait for the latest exchange (x_pending), except an image-only vr_trace_to_image,
  // which only needs the one before (its ring slot is then free on every rank).
  cudaStream_t xstream = nullptr;
  cudaEvent_t ev_trace = nullptr, ev_x[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
  unsigned int xserial = 0;                     // exchanges queued on xstream so far (images and layers)
  bool x_pending = false;                       // exchange `xserial` may still be running on xstream
  unsigned int x_of_img_epoch[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }; // xserial of image exchange e, at [e & 7] (0: not on xstream)
  // Image-only traces (vr_trace_to_image without VR_FRAME_WRITE_CANVAS) of consecutive frames rotate
  // over a few side streams, so that the next frames' CTAs move in while this frame's long rays drain
  // (the sampler is a persistent grid: alone, its tail leaves the GPU half empty for ~a tile's duration).
  // ev_t[k] marks the latest trace on tstream[k]; the frame's exchange waits for it, and so does every
  // entry point that joins (t_pending).
  static constexpr int kMaxTraceStreams = 4;
  cudaStream_t tstream[kMaxTraceStreams] = { nullptr, nullptr, nullptr, nullptr };
  cudaEvent_t ev_t[kMaxTraceStreams] = { nullptr, nullptr, nullptr, nullptr }, ev_main = nullptr;
  bool t_pending[kMaxTraceStreams] = { false, false, false, false };
  int trace_streams = 3;                        // VR_TRACE_STREAMS=0..4: 0 keeps image-only traces on the context's stream
  // exchange kernel shape (A/B runs): 0 = all layers in registers, 256 threads (default); 1 / 2 = layers in
  // batches of four, <= 72 registers, 128 / 256 threads; fold_grid = CTAs per SM (0: 1)
  int fold_light = 0, fold_grid = 2;
  bool fold_nr8 = false;                        // VR_FOLD_NR8=1: instantiate the fold for 8 ranks even with fewer
  int exchange_max_ctas = 0;                    // VR_EXCHANGE_MAX_CTAS: cap of the exchange kernels' grids (0: none).  Their
                                                // CTAs spin on the peers' flags: several ranks sharing ONE device inside ONE
                                                // process (tests) must all fit on it at once
  bool layer_push = true;                       // VR_LAYER_PUSH=0: ray layers stay in their rank's arena, the fold pulls them
  bool timeline = false;                        // VR_TIMELINE=1: kernels leave globaltimer stamps in the flags
  bool frame_poisoned = false;                  // a rank-local error hit this frame: the next collective aborts
  // rank 0: "this buffer holds the cleared value outside the rectangle kept in the arena flags",
  // valid while api_serial has not moved and the frame size is the same
  struct Clean { bool valid = false; uint64_t serial = 0; int W = 0, H = 0; };
  Clean clean_res[2], clean_canvas;
  int* minmax_dev = nullptr;                    // {min,max} pixel id of the current list
};

} // namespace vr

// Entry-point prologues.  VR_ENTER counts the call in ctx->api_serial: rank 0's exchange kernels skip
// re-clearing the parts of the result image / canvas they left cleared last time, which is only sound
// if nothing else may have written those buffers since -- any entry point that is not explicitly
// marked read-only (VR_ENTER_RO: never writes the canvas or the composited image) invalidates that.
#define VR_JOIN(ctx)                                                                               \
  do                                                                                               \
  {                                                                                                \
    if ((ctx)->comm.x_pending)                                                                     \
    {                                                                                              \
      cudaStreamWaitEvent((ctx)->stream, (ctx)->comm.ev_x[(ctx)->comm.xserial & 7], 0);            \
      (ctx)->comm.x_pending = false;                                                               \
    }                                                                                              \
    for (int k_ = 0; k_ < vr::Comm::kMaxTraceStreams; ++k_)                                        \
      if ((ctx)->comm.t_pending[k_])                                                               \
      {                                                                                            \
        cudaStreamWaitEvent((ctx)->stream, (ctx)->comm.ev_t[k_], 0);                               \
        (ctx)->comm.t_pending[k_] = false;                                                         \
      }                                                                                            \
  } while (0)
#define VR_ENTER(ctx) do { if (!(ctx)) return VR_ERR_INVALID; ++(ctx)->api_serial; VR_JOIN(ctx); } while (0)
#define VR_ENTER_RO(ctx) do { if (!(ctx)) return VR_ERR_INVALID; VR_JOIN(ctx); } while (0)
// (vr_trace_to_image: joins selectively, see there)
#define VR_ENTER_NOJOIN(ctx) do { if (!(ctx)) return VR_ERR_INVALID; } while (0)

struct vr_ctx
{
  int device = 0;
  uint64_t api_serial = 0;
  bool canvas_exposed = false; // vr_canvas_ptrs handed the canvas out: its contents are never assumed
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  vr::TraceParams* multi_table = nullptr; // trace_multi_kernel: per-block parameters of the current batch
  unsigned* multi_tile_end = nullptr; //                     running sum of the blocks' tile counts
  int multi_cap = 0;
  // PNG encode on the device (png.cu): row slots + per-row checksums, the file, its size (device + pinned mirror)
  unsigned char* png_scratch = nullptr;
  unsigned char* png_out = nullptr;
  unsigned long long* png_total = nullptr;
  unsigned long long* png_total_host = nullptr;
  size_t png_scratch_cap = 0, png_out_cap = 0;
  static constexpr int kMultiSlots = 4;
  unsigned char* multi_host = nullptr; // pinned staging ring of the two arrays above
  cudaEvent_t multi_ev[kMultiSlots] = { nullptr, nullptr, nullptr, nullptr };
  int multi_slot = 0;
  unsigned trace_preloaded = 0; // bit (kind | dtype << 1 | assoc << 2 | 64-bit index << 3): sampler variants loaded
  int sm_count = 148;
  int ctas_per_sm = 0;        // trace kernel residency (0 = built-in default)
  int tile_order = 0;         // VR_TILE_ORDER=1: tiles in row-major order instead of centre-out
  bool count_samples = false; // accumulate the number of samples taken (debug/bench)
  bool no_sparse = false;     // never select the sampler's sparse march (A/B runs)
  bool no_brick = false;      // never select the brick march (A/B runs)

  std::map<int, vr::Block> blocks;
  float first_sample_abs = 0.f, first_sample_rel = 0.0001f; // vr_set_first_sample_offset
  // vr_block_unstructured: the external-face mask of the last publish per block id, reused while the
  // connectivity (its FNV-1a hash, size and cell shape) stays the same -- static topology republished every cycle
  struct UMaskCache { unsigned long long hash = 0; size_t n_cells = 0; int shape = 0; std::vector<unsigned char> mask; };
  std::map<int, UMaskCache> umask_cache;
  float4* lut = nullptr;
  int lut_size = 0;

  // frame buffers (sized W x H, grown on demand)
  int W = 0, H = 0;
  size_t cap_pixels = 0;
  float4* canvas_rgba = nullptr;
  float* canvas_depth = nullptr;
  uchar4* img_rgba = nullptr;  // quantised own image (lives in the arena when comm is on)
  float* img_depth = nullptr;
  uchar4* res_rgba = nullptr;  // composited result (rank 0)
  float* res_depth = nullptr;
  bool img_in_arena = false;
  bool canvas_in_arena = false;
  int img_rect[4] = { 0, 0, 0x7fffffff, 0x7fffffff }; // where the quantised image may be non-empty
  // the pending / ahead image was pushed into the owners' receive slots (VR_FRAME_PUSH, sampler mode 5)
  bool img_pushed = false, img_pushed_ahead = false;
  // an image traced one frame ahead of the pending exchange (VR_FRAME_AHEAD) and its rectangle
  bool img_ahead = false;
  int img_rect_ahead[4] = { 0, 0, 0x7fffffff, 0x7fffffff };

  // partial list
  vr_partial* partials = nullptr;
  size_t partial_cap = 0;
  unsigned long long* partial_count = nullptr; // device counter
  unsigned long long* partial_count_tmp = nullptr;
  // the list the read-side entry points (count/download/to_canvas) see: the own list above, or
  // the composited list in the exchange arena after vr_comm_composite_partials on rank 0
  const vr_partial* plist = nullptr;
  const unsigned long long* plist_count = nullptr;
  size_t plist_cap = 0;
  size_t n_partials_host = 0;                  // valid after a sync'ing call
  int pW = 0, pH = 0;
  // partial composite scratch
  int* px_count = nullptr;   // per pixel (padded)
  int* px_end = nullptr;
  int* sidx = nullptr;
  vr_partial* rec = nullptr;
  int* scan_blocks = nullptr;
  size_t scratch_px = 0, scratch_parts = 0, scratch_blocks = 0;
  vr_partial* partials_tmp = nullptr;
  size_t partial_tmp_cap = 0;

  // ray layers of the current frame
  vr::LayerTable* ltab_host = nullptr;   // pinned host copy (filled as layers are traced)
  vr::LayerTable* ltab = nullptr;        // device table (own allocation, or the arena's by parity)
  float4* lpool_rgba = nullptr;          // layer pool (own allocation or arena)
  float* lpool_depth = nullptr;
  size_t lpool_cap = 0, lpool_used = 0;
  bool layers_in_arena = false;
  bool layers_pushed = false;            // this frame's entries also go to the tile owners' receive pools
  // without an exchange arena the context owns a ring of three layer buffers (table + pools), like the
  // arena's: the fold of frame k runs on the exchange stream while frame k+1 is traced into the next one
  static constexpr int kOwnLayerRing = 3;
  vr::LayerTable* own_ltab[kOwnLayerRing] = { nullptr, nullptr, nullptr };
  float4* own_lpool_rgba[kOwnLayerRing] = { nullptr, nullptr, nullptr };
  float* own_lpool_depth[kOwnLayerRing] = { nullptr, nullptr, nullptr };
  size_t own_lpool_cap[kOwnLayerRing] = { 0, 0, 0 };
  int own_lslot = 0;
  int lW = 0, lH = 0;

  unsigned long long* scratch_u64 = nullptr;
  uchar4* enc_rgba = nullptr;  // encoded final image (vr_canvas_download_rgba8)
  size_t enc_cap = 0;

  unsigned int* tile_counter = nullptr;   // [0] single launches, [1 + k] launch k of a batched call
  unsigned long long* sample_counter = nullptr;
  // side streams of the batched block loop (vr_trace_blocks_to_layers)
  cudaStream_t aux[4] = { nullptr, nullptr, nullptr, nullptr };
  cudaEvent_t ev_fork = nullptr, ev_join[4] = { nullptr, nullptr, nullptr, nullptr };

  vr::Comm comm;
};

namespace vr
{
// sampler.cu
cudaError_t launch_trace(const TraceParams& p, int mode_partials, int sm_count, cudaStream_t s,
                         bool zero_counter = true);

// stage.cu
cudaError_t launch_trace_multi(const TraceParams& first, const TraceParams* table, const unsigned* tile_end, int n,
                               unsigned long long total_tiles, unsigned* counter, int sm_count, cudaStream_t s);
cudaError_t launch_fetch_lines(unsigned char* want, unsigned char* have, const void* src, void* dst,
                               size_t n_lines, size_t n_bytes, bool all, unsigned long long* n_have,
                               int sm_count, cudaStream_t s);

cudaError_t launch_gather_strided(const void* src, int elem_bytes, size_t stride, size_t n, void* dst, int sm_count,
                                  cudaStream_t s);
cudaError_t launch_count_lines(const unsigned char* have, size_t n_lines, unsigned long long* out, cudaStream_t s);

// composite.cu
cudaError_t launch_canvas_clear(float4* rgba, float* depth, size_t n, cudaStream_t s);
cudaError_t launch_quantize(const float4* rgba, const float* depth, size_t n, uchar4* out,
                            float* out_depth, cudaStream_t s);
cudaError_t launch_fold_images(const uchar4* rgba, const float* depth, size_t layer_stride,
                               const int* order /* layer index per fold step, by value */,
                               int n_layers, size_t n, uchar4* out, float* out_depth,
                               cudaStream_t s);
cudaError_t launch_zbuffer(uchar4* front, float* fdepth, const uchar4* img, const float* depth,
                           size_t n, cudaStream_t s);
cudaError_t launch_blend_background(float4* canvas, size_t n, const float bg[4], cudaStream_t s);
cudaError_t launch_encode_rgba8(const float4* canvas, int W, int H, int flip, const float* bg /* null: none */,
                                uchar4* out, cudaStream_t s);
cudaError_t launch_image_to_canvas(const uchar4* rgba, const float* depth, size_t n, float4* canvas,
                                   float* cdepth, cudaStream_t s);
cudaError_t launch_synth_braid(void* field, int dtype, const int n[3], const int start[3],
                               const int global[3], cudaStream_t s);

struct PartialScratch
{
  int* px_count;   // per pixel, padded to partial_scan_padded(n_pixels)
  int* px_end;     // per pixel segment end offset (same padding)
  int* sidx;       // original list index of each binned record
  vr_partial* rec; // records binned by pixel
  int* scan_blocks;
};
size_t partial_scan_padded(size_t n_pixels);
struct ToCanvasParams
{
  float origin[3], look[3], delta_x[3], delta_y[3];
  float pv[16];
  int W, H;
};
#ifdef __CUDACC__
// ---- cross-GPU flag protocol of the exchange kernels (comm.cu, layers.cu): system-scope release /
// acquire on epoch counters in the peers' arenas, every wait bounded in time
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v)
{
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p)
{
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *flag >= epoch; gives up after timeout_ns (0 = never) so that a peer that died or bailed
// out with a rank-local error cannot hang this GPU
__device__ __forceinline__ bool wait_epoch(const unsigned int* flag, unsigned int epoch, unsigned long long timeout_ns,
                                           unsigned sleep_ns = 32)
{
  if (ld_acquire_sys(flag) >= epoch) return true;
  const unsigned long long t0 = global_ns();
  for (;;)
  {
    __nanosleep(sleep_ns);
    if (ld_acquire_sys(flag) >= epoch) return true;
    if (timeout_ns && global_ns() - t0 > timeout_ns) return false;
  }
}
// what an exchange kernel leaves for the host when it could not complete: the first failing epoch,
// the reason (1 = a wait ran into the time limit, 2 = a peer aborted the exchange) and the peer
struct ExchangeError
{
  unsigned int epoch, path, reason, peer;
};
__device__ __forceinline__ void report_error(ExchangeError* e, unsigned epoch, unsigned path, unsigned reason, unsigned peer)
{
  if (atomicCAS(&e->epoch, 0u, epoch) == 0u)
  {
    e->path = path;
    e->reason = reason;
    e->peer = peer;
  }
}
// the wait of an exchange kernel's prologue: thread r < size waits for rank r's ready flag of `epoch`
// and checks that r did not abort it; returns (block-wide) whether the exchange can go ahead
__device__ __forceinline__ bool wait_all_ready(ExchangeError* err, const unsigned int* ready, const unsigned int* aborted,
                                               unsigned path, int size, unsigned epoch, unsigned long long timeout_ns,
                                               unsigned sleep_ns = 32)
{
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  if ((int)threadIdx.x < size)
  {
    if (!wait_epoch(ready + threadIdx.x, epoch, timeout_ns, sleep_ns))
    {
      s_bad = 1;
      report_error(err, epoch, path, 1u, threadIdx.x);
    }
    else if (((volatile const unsigned int*)aborted)[threadIdx.x] == epoch)
    {
      s_bad = 1;
      report_error(err, epoch, path, 2u, threadIdx.x);
    }
  }
  __syncthreads();
  return s_bad == 0;
}

// one pixel of partials_to_canvas (VolumeRenderer.cpp:287-391): recompute the ray like K1 (with the
// reference's delta_y-from-ru quirk baked into T by the host), project origin + depth*dir, blend
// the partial over the canvas value `in`
__device__ __forceinline__ void partial_to_canvas(const vr_partial& part, const ToCanvasParams& T,
                                                  float4 in, float4& o, float& image_depth)
{
  const int pixel_id = part.pixel_id;
  const int i = pixel_id % T.W, j = pixel_id / T.W;
  const float fx = (2.f * (float)i - (float)T.W) / 2.0f;
  const float fy = (2.f * (float)j - (float)T.H) / 2.0f;
  float dx = T.look[0] + T.delta_x[0] * fx + T.delta_y[0] * fy;
  float dy = T.look[1] + T.delta_x[1] * fx + T.delta_y[1] * fy;
  float dz = T.look[2] + T.delta_x[2] * fx + T.delta_y[2] * fy;
  const float r = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz); // vtkm::Normalize
  dx = r * dx; dy = r * dy; dz = r * dz;
  const float wd = part.depth;
  const float x = T.origin[0] + wd * dx, y = T.origin[1] + wd * dy, z = T.origin[2] + wd * dz;
  const float* m = T.pv;
  const float n2 = m[8] * x + m[9] * y + m[10] * z + m[11] * 1.f;
  const float n3 = m[12] * x + m[13] * y + m[14] * z + m[15] * 1.f;
  image_depth = 0.5f * (n2 / n3) + 0.49f;
  const float a = 1.f - part.alpha;
  o.x = part.rgb[0] + in.x * a;
  o.y = part.rgb[1] + in.y * a;
  o.z = part.rgb[2] + in.z * a;
  o.w = in.w * a + part.alpha;
}
#endif

// composite the list `in` (length *count_dev clamped to cap, pixel ids in [0, n_pixels)) into `out`
// (<= 1 per pixel); *out_count (device) receives the number written.  With to_canvas != nullptr the
// fold also writes the final canvas (partials_to_canvas fused).  Returns the number of launches.
int launch_partials_composite(const vr_partial* in, const unsigned long long* count_dev, size_t cap,
                              size_t n_pixels, const PartialScratch& sc, vr_partial* out,
                              unsigned long long* out_count, const ToCanvasParams* to_canvas,
                              float4* canvas, float* cdepth, int canvas_clear, cudaStream_t s,
                              cudaError_t* err);
// Reorder the list by (pixel, depth, list index) into sorted_out and write each pixel's segment END
// offset (padded array) and the list's {min,max} pixel id.
int launch_partials_pixel_sort(const vr_partial* in, const unsigned long long* count_dev, size_t cap,
                               size_t n_pixels, const PartialScratch& sc, vr_partial* sorted_out,
                               size_t sorted_cap, int* end_out, int* minmax, cudaStream_t s,
                               cudaError_t* err);

cudaError_t launch_partial_append(const vr_partial* src_dev, size_t n, vr_partial* list, unsigned long long* count,
                                  size_t capacity, cudaStream_t s);
cudaError_t launch_partials_to_canvas(const vr_partial* p, const unsigned long long* count_dev,
                                      size_t max_n, const ToCanvasParams& tp, float4* canvas,
                                      float* cdepth, cudaStream_t s);

// layers.cu -- dense per-block ray layers (path B without lists)
constexpr int kMaxCommRanks = 16;
constexpr int kMaxLayers = 1024;      // per rank and frame
constexpr int kMaxSmemLayers = 2048;  // all ranks' tables together (shared memory of the fold kernel)
struct LayerDesc
{
  int x0, y0, w, h;            // the block's ray rectangle on screen (K1 "subset")
  unsigned long long base;     // first entry in the rank's layer pool
};
struct LayerTable
{
  int n;
  int pad[3];
  LayerDesc d[kMaxLayers];
};
struct LayerFlags
{
  unsigned int ready[kMaxCommRanks];
  unsigned int done[kMaxCommRanks];
  unsigned int cta_done;
  unsigned int aborted[kMaxCommRanks]; // rank r aborted layer exchange <epoch> (rank-local error)
  unsigned int err[4];                 // ExchangeError of this rank's layer exchanges
};
struct LayerFoldParams
{
  int rank, size;
  unsigned int epoch;
  int W, H;
  int clear;        // single rank: treat the canvas as cleared (every pixel written)
  int smem_layers;  // capacity of the shared-memory table copy
  const LayerTable* table[kMaxCommRanks];
  const float4* pool_rgba[kMaxCommRanks];
  const float* pool_depth[kMaxCommRanks];
  unsigned char* flags[kMaxCommRanks];
  float4* canvas_rgba; // where finished pixels go (rank 0's canvas)
  float* canvas_depth;
  ToCanvasParams tp;
  unsigned long long timeout_ns; // bound of every cross-rank wait (0 = none)
  int light;        // 32x4 tiles / 128-thread CTAs (fit the slot of one sampler CTA) instead of 32x8 / 256
  int max_ctas;     // cap of the grid (0: none)
  int pushed;       // the pools are this rank's RECEIVE pools (the samplers pushed every entry to the owner of its
                    // tile): tiles are owned by absolute index, (tile_y * tiles_per_row + tile_x) % size
};
cudaError_t launch_layers_fold(const LayerFoldParams& p, bool comm, int sm_count, cudaStream_t s);
cudaError_t launch_layers_wait_done(unsigned char* flags, int size, unsigned int epoch, unsigned long long timeout_ns,
                                    cudaStream_t s);
cudaError_t launch_layers_abort(const LayerFoldParams& p, cudaStream_t s);
cudaError_t launch_layers_to_partials(const LayerTable* table, int n_layers, const float4* pool_rgba,
                                      const float* pool_depth, int W, vr_partial* out,
                                      unsigned long long* count, size_t cap, cudaStream_t s);

// comm.cu
struct FoldP2PParams
{
  unsigned char* const* peers; // device table of arena base pointers
  int rank, size;
  unsigned int epoch;
  size_t n_pixels;
  int W;
  int rect[4]; // {x0,y0,x1,y1}: my image is empty (colour 0, depth 1.001) outside of it
  size_t off_img_rgba, off_img_depth, off_res_rgba, off_res_depth, off_flags;
  int order[16]; // rank index per fold step (front to back)
  float4* canvas_rgba; // rank 0, fused ImageToCanvas (null: off)
  float* canvas_depth;
  int zbuffer;         // 1: select-nearest (opaque surfaces) instead of the ordered blend
  // rank 0: the result image of this parity / the canvas are known to be cleared outside the
  // rectangle the previous exchange left in the flags -> only that rectangle needs clearing again
  int track_res, track_canvas;
  // my image of this epoch was pushed into the owners' receive slots (sampler mode 5); a receive slot
  // holds share_groups 4-pixel groups per source rank, at off_recv_* of the owner's arena
  int pushed;
  size_t share_groups;
  size_t off_recv_rgba, off_recv_depth;
  unsigned long long timeout_ns; // bound of every cross-rank wait (0 = none)
  int timeline;                  // diagnostics: leave globaltimer stamps in the flags
  int light;                     // 0: fold_p2p_kernel; 1 / 2: fold_p2p_light_kernel with 128 / 256 threads
  int grid_per_sm;               // CTAs per SM of the fold (0: one)
  int max_ctas;                  // cap of the grid (0: none)
  int force_nr8;                 // diagnostics: the 8-rank instantiation whatever the size
  // zbuffer: the order in which the reference's radix-k tree visits the ranks' fragments, per piece of the
  // frame (vr_radixk.hpp): piece column / row starts, pieces per row, and pos[g][rank] = position of the
  // rank in the sequence of the piece gid g ends up owning (ties at equal depth go to the LAST position)
  struct ZSelect
  {
    int lo_x[16], lo_y[16];
    int div_x;
    unsigned char pos[16][16];
  } zs;
};
cudaError_t launch_fold_p2p(const FoldP2PParams& p, int sm_count, cudaStream_t s);

// ---- kernel preloading.  CUDA loads a kernel lazily at its first launch, and that load may have to
// synchronise the context -- which never returns while an exchange kernel of this context is resident and
// spinning on a peer whose own kernel the SAME host thread has not launched yet (one thread driving several
// contexts: vr_comm_connect_local; documented hazard of CUDA_MODULE_LOADING=LAZY).  vr_create therefore loads
// every kernel of the exchange, fold, composite and staging units up front, and a block's publish loads the
// sampler variants that block can run (cudaFuncGetAttributes forces the load; later calls are no-ops).
template <typename K>
inline void preload_kernel(K k)
{
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k));
}
void preload_comm_kernels();
void preload_layers_kernels();
void preload_composite_kernels();
void preload_stage_kernels();
void preload_trace_kernels(const BlockDev& blk);
void preload_png_kernels();
void preload_unstructured_kernels();

// unstructured.cu
cudaError_t launch_utrace_partials(const TraceParams& p, const UMeshDev& u, int sm_count, cudaStream_t s);
cudaError_t umesh_bounds(const float* xyz, size_t n_points, int* keys_dev, float bmin[3], float bmax[3], int sm_count,
                         cudaStream_t s);
cudaError_t umesh_build_bins(UMeshDev& u, int** bin_start_out, int** bin_cells_out, unsigned char** bin_ext_out, int sm_count,
                             cudaStream_t s);

// png.cu
unsigned png_slot_stride(int W);
size_t png_capacity(int W, int H);      // upper bound of the file size
size_t png_scratch_bytes(int W, int H);
cudaError_t launch_png_encode(const uchar4* rgba, int W, int H, unsigned char* scratch, unsigned char* out,
                              unsigned long long capacity, unsigned long long* total_dev, int sm_count, cudaStream_t s);
} // namespace vr
