/*
 * vr_b200.h -- C ABI of the B200-native volume-render backend (libvr_b200.so).
 *
 * Drop-in boundary for Ascent's `volume` plot hot path.  The
 * reference has no FFI for this path today; these entry points are what
 * vtkh::VolumeRenderer would bind at its four private call sites
 *   (1) m_mapper->RenderCells            src/libs/vtkh/rendering/VolumeRenderer.cpp:516-528
 *   (2) StructuredWrapper::render        src/libs/vtkh/rendering/VolumeRenderer.cpp:230-284
 *   (3) m_compositor->AddImage/Composite src/libs/vtkh/rendering/VolumeRenderer.cpp:652-688
 *   (4) PartialCompositor::composite +   src/libs/vtkh/rendering/VolumeRenderer.cpp:580-595
 *       partials_to_canvas               src/libs/vtkh/rendering/VolumeRenderer.cpp:287-391
 * (see INTEGRATION.md for the vtk-h side of the binding).
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Every call returns a vr_status;
 * vr_last_error(ctx) gives the message (the vtk-h wrapper rethrows it as vtkh::Error).  The
 * caller owns every host buffer; the library owns device copies.  One context drives one GPU
 * from one host thread (Ascent: one MPI rank per GPU, ascent_main_runtime.cpp:190-192).
 * All work is queued on the context's CUDA stream; entry points that fill HOST memory
 * synchronise that stream before returning, `_dev`/async ones do not.
 * There is no CPU fallback: without a usable sm_100 device vr_create fails.
 */
#ifndef VR_B200_H
#define VR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VR_API __attribute__((visibility("default")))

typedef struct vr_ctx vr_ctx;

typedef enum
{
  VR_OK = 0,
  VR_ERR_INVALID = 1, /* bad argument (vtk-h: Error thrown by the caller-side checks) */
  VR_ERR_CUDA = 2,    /* CUDA runtime failure */
  VR_ERR_NOMEM = 3,
  VR_ERR_STATE = 4    /* call made in the wrong order (e.g. composite before render) */
} vr_status;

enum { VR_F32 = 0, VR_F64 = 1 };     /* field scalar type (ascent_vtkh_data_adapter.cpp:1843-1887) */
enum { VR_POINT = 0, VR_CELL = 1 };  /* field association */
/* where a caller-supplied field lives: pageable/pinned host memory to copy, device memory to adopt, or
 * page-locked MAPPED host memory (cudaHostAlloc / cudaHostRegister) to sample in place over PCIe
 * (VR_HOST_MAPPED) or to stage on demand, only the lines the rays touch (VR_HOST_STAGED) */
enum { VR_HOST = 0, VR_DEVICE = 1, VR_HOST_MAPPED = 2, VR_HOST_STAGED = 3 };

/* vtkm::rendering::Camera as parse_camera fills it
 * (ascent_runtime_conduit_to_vtkm_parsing.cpp:97-173); f32 like VTK-m's. */
typedef struct
{
  float position[3];
  float look_at[3];
  float up[3];
  float fov;  /* vertical, degrees */
  float zoom; /* vtkm zoom factor (1 = none) */
  float xpan, ypan;
  float near_plane, far_plane;
} vr_camera;

/* vtkh::VolumePartial<float>, src/libs/vtkh/compositing/VolumePartial.hpp:48-56 (24-byte POD) */
typedef struct
{
  int32_t pixel_id;
  float depth;
  float rgb[3];
  float alpha;
} vr_partial;

/* ------------------------------------------------------------------ context */
VR_API vr_status vr_create(int device, vr_ctx** out);
VR_API void vr_destroy(vr_ctx* ctx);
VR_API const char* vr_last_error(const vr_ctx* ctx); /* ctx may be NULL: last vr_create error */
/* Use a caller-owned cudaStream_t (e.g. torch's current stream); NULL restores the own stream. */
VR_API vr_status vr_set_stream(vr_ctx* ctx, void* cuda_stream);
VR_API vr_status vr_synchronize(vr_ctx* ctx);
/* Number of this library's kernels launched on this context since creation (bench evidence). */
VR_API uint64_t vr_kernel_launches(const vr_ctx* ctx);

/* ------------------------------------------------------------------ blocks (domains)
 * Replaces the vtkm::cont::DataSet a StructuredWrapper holds (VolumeRenderer.cpp:223-229,
 * SetInput :868-907).  dims are POINT dims, x fastest: index (k*ny + j)*nx + i.  A cell field
 * has (nx-1)(ny-1)(nz-1) values.  The field is copied (or adopted, see below) once per publish
 * and reused by every render of the batch.
 *   where == VR_HOST   : `field` is host memory, copied to the device (cudaMemcpyAsync on the
 *                        context's stream: a page-locked array must stay valid and unchanged until the
 *                        next synchronising call -- vr_synchronize or any entry point that fills host
 *                        memory; pageable memory is staged by the runtime before the call returns).
 *   where == VR_DEVICE : `field` is device memory on this GPU and is used in place (zero copy;
 *                        must outlive the block).
 *   where == VR_HOST_MAPPED : `field` is page-locked, device-mapped host memory (the simulation's
 *                        own array after cudaHostRegister, cf. the zero-copy Blueprint path of
 *                        ascent_vtkh_data_adapter.cpp:1351-1355); no copy is made, the sampler
 *                        pulls only the sectors its rays touch across PCIe.  Worth it for a field
 *                        rendered once or a few times per publish; must outlive the block.
 *   where == VR_HOST_STAGED : same kind of memory, staged on demand: the publish moves nothing; each
 *                        trace of the block is preceded by a pre-pass of the sampler that flags the
 *                        128-byte lines its rays will read and by a gather of the flagged lines that
 *                        are not on the device yet (coalesced reads over PCIe into a device buffer
 *                        the sampler then uses).  Lines stay resident until the next publish, so a
 *                        batch of views only fetches what each view adds.  At the default sampling
 *                        (samples = 100) a frame touches about a quarter of a 512^3 block: the in-situ
 *                        publish + render drops from ~11 ms (dense copy) to ~4 ms.  Results are
 *                        bit-identical to the other modes.  The array must stay valid and unchanged
 *                        until the last render of this publish has completed.                   */
VR_API vr_status vr_block_uniform(vr_ctx* ctx, int block_id, const int dims[3],
                                  const float origin[3], const float spacing[3], const void* field,
                                  int dtype, int assoc, int where);
VR_API vr_status vr_block_rectilinear(vr_ctx* ctx, int block_id, const int dims[3], const double* x,
                                      const double* y, const double* z, const void* field,
                                      int dtype, int assoc, int where);
/* N4 -- a domain with an explicit cell set: hexahedra (VTK vertex order) or tetrahedra, point or cell field.
 * The reference renders such a domain through UnstructuredWrapper::render (VolumeRenderer.cpp:182-221: VTK-m's
 * ConnectivityProxy::PartialTrace + vtkm_to_partials :141-180) and then renders EVERY domain as partials
 * (m_has_unstructured, :874-903): an unstructured block is accepted by vr_trace_to_partials / vr_render_partials
 * only.  xyz: n_points x 3 interleaved; connectivity: n_cells x (8 | 4) point indices.  VR_HOST: copied (f64
 * coordinates and 64-bit indices are narrowed to f32 / int32); VR_DEVICE: adopted in place (f32 coordinates and
 * 32-bit connectivity only).  The cell locator (uniform bins over the point bounds) is built on the device by this
 * call, the mesh boundary (which faces belong to one cell only: a ray is sampled only along the stretches between
 * an entering and a leaving crossing of that boundary, concavities and cavities included; the nearest 32
 * crossings of a ray are kept) on the host, once per connectivity (reused while it does not change); the call
 * synchronises.  Algorithm and how it is pinned without VTK-m: DESIGN.md section 4.5.                       */
enum { VR_TETRA = 10, VR_HEXAHEDRON = 12 }; /* VTK / vtkm::CellShape ids */
VR_API vr_status vr_block_unstructured(vr_ctx* ctx, int block_id, size_t n_points, const void* xyz, int coord_dtype,
                                       size_t n_cells, int cell_shape, const void* connectivity, int index_bits,
                                       const void* field, int dtype, int assoc, int where);
VR_API vr_status vr_block_free(vr_ctx* ctx, int block_id);
/* Strided values (ascent_vtkh_data_adapter.cpp:1836-1887: Blueprint arrays whose byte stride is a multiple of the
 * element size -- one component of an interleaved mcarray, a padded array -- which the reference hands to VTK-m
 * as an ArrayHandleStride).  dense[i] = src[element_offset + i * element_stride] is gathered once into a new
 * dense device array (where = VR_HOST: the strided span is copied across PCIe first; VR_DEVICE: gathered in
 * place at HBM speed), to be published with vr_block_*(..., dense, dtype, assoc, VR_DEVICE) and released with
 * vr_field_free after the block has been freed or re-published.  The call synchronises: src may be reused.  */
VR_API vr_status vr_field_gather_strided(vr_ctx* ctx, const void* src, int where, int dtype, size_t n_values,
                                         size_t element_stride, size_t element_offset, void** dense_dev_out);
VR_API vr_status vr_field_free(vr_ctx* ctx, void* dense_dev);
/* Bytes of a VR_HOST_STAGED block fetched to the device since its last publish (0 for other kinds). Syncs. */
VR_API vr_status vr_block_staged_bytes(vr_ctx* ctx, int block_id, size_t* bytes);
/* coords.GetBounds(): xmin,xmax,ymin,ymax,zmin,zmax */
VR_API vr_status vr_block_bounds(vr_ctx* ctx, int block_id, double out[6]);

/* ------------------------------------------------------------------ transfer function
 * The 1024 x float4 table Mapper::SetActiveColorTable / convert_table build on the host
 * (VolumeRenderer.cpp:64-91, :518): already opacity-corrected and uint8-rounded.             */
VR_API vr_status vr_set_tf(vr_ctx* ctx, const float* rgba, int n_entries);
/* Where the structured sampler puts a ray's first sample: entry + abs_offset + extent_rel * |block extent|.
 * Default (0, 1e-4): VolumeRendererStructured's meshEpsilon of the VTK-m the reference pins (v2.1.0).  (1e-4, 0) is
 * what an older generation of VTK-m did -- the one that rendered the reference's three pure-volume golden images,
 * which this library then reproduces uint8 for uint8.  Applies to the traces issued after the
 * call; unstructured blocks are not affected.                                                            */
VR_API vr_status vr_set_first_sample_offset(vr_ctx* ctx, float abs_offset, float extent_rel);

/* ------------------------------------------------------------------ device canvas
 * The context keeps one float canvas (RGBA f32 + depth f32, W x H) in HBM: the stand-in for
 * vtkm::rendering::CanvasRayTracer's buffers (Render.hpp:80).                                 */
VR_API vr_status vr_canvas_clear(vr_ctx* ctx, int width, int height); /* colour 0, depth 1.001 */
VR_API vr_status vr_canvas_upload(vr_ctx* ctx, int width, int height, const float* rgba,
                                  const float* depth);
VR_API vr_status vr_canvas_download(vr_ctx* ctx, float* rgba, float* depth); /* syncs */
/* The same for the pixel rectangle [x0, x1) x [y0, y1) only, written into the same pixels of full-frame host
 * buffers.  A frame that started from Canvas::Clear equals the cleared canvas (colour 0, depth 1.001 -- what the
 * caller's host canvas holds after Render::ClearCanvas, Render.cpp) outside the screen footprint of the data:
 * vr_find_subset of the GLOBAL bounds (all ranks' domains, VolumeRenderer::PreExecute has them), united with the
 * previous frame's footprint when the camera moved.  Syncs.                                              */
VR_API vr_status vr_canvas_download_rect(vr_ctx* ctx, int x0, int y0, int x1, int y1, float* rgba, float* depth);
VR_API vr_status vr_canvas_ptrs(vr_ctx* ctx, void** rgba_dev, void** depth_dev);
/* Frame epilogue on the device (what Scene::Render does with each finished canvas on rank 0,
 * Scene.cpp:236-243, when annotations are off):
 * vr_canvas_blend_background = Render::RenderBackground -> vtkm Canvas::BlendBackground
 * (Render.cpp:277-286), in place on the device canvas.
 * vr_canvas_download_rgba8   = the float -> uint8 conversion PNGEncoder::Encode applies to the colour
 * buffer in Render::Save (Render.cpp:299-312, ascent_png_encoder.cpp:257-281): (unsigned char)(c*255.f),
 * rows flipped when flip_rows != 0; with bg_rgba != NULL the background blend is applied on the fly
 * (the canvas itself is left untouched).  Moves 4 B/pixel to the host instead of the 20 B/pixel float
 * canvas.  Syncs.                                                                               */
VR_API vr_status vr_canvas_blend_background(vr_ctx* ctx, const float bg_rgba[4]);
VR_API vr_status vr_canvas_download_rgba8(vr_ctx* ctx, const float* bg_rgba, int flip_rows,
                                          uint8_t* out_rgba8);
/* Render::Save's PNG encode on the device (Render.cpp:299-312 -> PNGEncoder::Encode + Save,
 * ascent_png_encoder.cpp:258-303: float -> uint8, rows flipped, lodepng with Huffman-only deflate): the canvas
 * (over bg_rgba when not NULL) becomes a complete PNG file -- RGBA, 8 bits, Sub-filtered scanlines, one
 * fixed-Huffman deflate block per scanline with run-length matches, Adler-32 and CRC-32 combined from per-row
 * partial sums -- and only the file crosses PCIe.  The decoded pixels equal vr_canvas_download_rgba8(ctx,
 * bg_rgba, 1, ..); the byte stream is a different, equally lossless deflate encoding than lodepng's.
 * *png_bytes receives the file size (always); png_host == NULL only queries it.  vr_png_bound(w, h) is an upper
 * bound of the size for any content.  No annotations or tEXt comments (annotations off).  Syncs. */
VR_API vr_status vr_canvas_encode_png(vr_ctx* ctx, const float* bg_rgba, uint8_t* png_host, size_t capacity,
                                      size_t* png_bytes);
VR_API size_t vr_png_bound(int width, int height);

/* ------------------------------------------------------------------ (1) path A render
 * MapperVolume::RenderCells for one block and one camera: ray generation over the block's
 * screen subset, canvas-depth clamp, bounds intersection, sampling, blend over the canvas and
 * projected entry depth (K1-K7 fused in one kernel), on the DEVICE canvas.
 * use_canvas_depth: 0 = canvas is known to be cleared (skip the per-ray un-projection, result
 * identical); 1 = clamp rays at the canvas depth (opaque plots rendered before the volume,
 * Scene.cpp:200-207).                                                                          */
VR_API vr_status vr_trace_to_canvas(vr_ctx* ctx, int block_id, const vr_camera* cam,
                                    float sample_dist, float range_min, float range_max,
                                    int use_canvas_depth);
/* Host-buffer form of the same call, exactly what VolumeRenderer.cpp:516-528 needs: canvas in,
 * canvas out (uploads, traces, downloads, syncs).                                              */
VR_API vr_status vr_render_image(vr_ctx* ctx, int block_id, const vr_camera* cam, int width,
                                 int height, float sample_dist, float range_min, float range_max,
                                 float* rgba_inout, float* depth_inout);

/* The whole per-rank body of RenderOneDomainPerRank (VolumeRenderer.cpp:482-536) for a frame
 * that starts from a CLEARED canvas -- Canvas::Clear, RenderCells (K1-K7), Image::Init
 * (Image.hpp:80-113) and, with VR_FRAME_WRITE_CANVAS, Renderer::ImageToCanvas
 * (Renderer.cpp:265-283) -- fused into ONE kernel launch.  Leaves the rank's quantised RGBA8 +
 * depth image in the context (vr_image_download / vr_comm_composite_images) and, if asked, the
 * k/255 float canvas.  Results are bit-identical to vr_canvas_clear + vr_trace_to_canvas(.., 0) +
 * vr_image_from_canvas [+ vr_image_to_canvas_dev].
 * VR_FRAME_NO_CLEAR: do not write the pixels outside the block's screen rectangle (they are left
 * undefined); only for images handed to vr_comm_composite_images, which never reads them.
 * VR_FRAME_AHEAD (multi-GPU, after vr_comm_init): this is the image of the frame AFTER the one whose
 * vr_comm_composite_images call is still to come -- it goes to the next slot of the rank's image ring
 * and becomes the pending image once that exchange has been issued.  Lets the renders of a batch
 * (Scene.cpp:133-149) be software-pipelined: trace(k+1), exchange(k), trace(k+2), exchange(k+1), ... so
 * that no rank idles in an exchange while it could be tracing.  At most one frame may be ahead.
 * VR_FRAME_PUSH (multi-GPU, after vr_comm_connect; implies NO_CLEAR; width % 4 == 0): every finished
 * pixel is stored straight into the exchange's receive slot on the GPU that owns the pixel (posted
 * NVLink stores from the sampler's epilogue, 8 bytes per pixel of the block's screen rectangle)
 * instead of into this rank's own image, so that vr_comm_composite_images folds from local memory
 * only -- no NVLink load round trips in the exchange.  The rank's own image is then not available to
 * vr_image_download.  Ranks may mix pushed and unpushed images within one exchange.            */
enum { VR_FRAME_WRITE_CANVAS = 1, VR_FRAME_NO_CLEAR = 2, VR_FRAME_AHEAD = 4, VR_FRAME_PUSH = 8 };
VR_API vr_status vr_trace_to_image(vr_ctx* ctx, int block_id, const vr_camera* cam, int width,
                                   int height, float sample_dist, float range_min, float range_max,
                                   int flags);

/* ------------------------------------------------------------------ (2) path B render
 * StructuredWrapper::render: same trace on a zeroed ray buffer, then keep rays with
 * alpha >= 0.001 as partials {pixel, exit distance, rgb, alpha}.  Partials are APPENDED to the
 * context's device partial list (one list per frame, all local domains).                       */
VR_API vr_status vr_partials_begin(vr_ctx* ctx, int width, int height);
VR_API vr_status vr_trace_to_partials(vr_ctx* ctx, int block_id, const vr_camera* cam,
                                      float sample_dist, float range_min, float range_max,
                                      int use_canvas_depth);
/* Partials of another producer join the frame's list (after vr_partials_begin, in any order with
 * vr_trace_to_partials): Devil Ray's volume integrator hands apcomp::VolumePartial<float> -- this POD -- to the
 * same PartialCompositor (dray/rendering/renderer.cpp:309-331), so its lists can be composited here, across ranks
 * too (vr_comm_composite_partials).  VR_HOST lists are copied (the call then synchronises), VR_DEVICE lists read
 * in stream order.  Pixel ids are validated for host lists.                                              */
VR_API vr_status vr_partials_append(vr_ctx* ctx, const vr_partial* partials, size_t n, int where);
VR_API vr_status vr_partials_count(vr_ctx* ctx, size_t* n); /* syncs */
VR_API vr_status vr_partials_download(vr_ctx* ctx, vr_partial* out, size_t capacity, size_t* n);
/* Host-buffer form: library-allocated array, release with vr_free.  Order within the array is
 * unspecified (the reference's is ray order; compositing sorts anyway).                        */
VR_API vr_status vr_render_partials(vr_ctx* ctx, int block_id, const vr_camera* cam, int width,
                                    int height, float sample_dist, float range_min,
                                    float range_max, const float* depth_in, vr_partial** out,
                                    size_t* n);
VR_API void vr_free(void* p);

/* ------------------------------------------------------------------ (2'+4') path B without lists
 * Same result as (2) + (4) for structured blocks, a different data structure: a block adds at most
 * one partial per pixel and only inside its screen rectangle, so every block's rays are kept as a
 * dense "layer" {rgba, exit distance} over that rectangle (alpha = 0 where the reference's
 * alpha < 0.001 test drops the ray) and composited by ONE kernel that gathers, per pixel, the
 * entries of the layers covering it, orders them by (exit distance, rank, block), folds them with
 * VolumePartial::blend and writes the canvas pixel (partials_to_canvas).  No compaction, sort,
 * scan or gather passes.                                                                        */
VR_API vr_status vr_layers_begin(vr_ctx* ctx, int width, int height);
VR_API vr_status vr_trace_to_layer(vr_ctx* ctx, int block_id, const vr_camera* cam, float sample_dist,
                                   float range_min, float range_max, int use_canvas_depth);
/* The whole per-domain loop of RenderMultipleDomainsPerRank (VolumeRenderer.cpp:557-578) in one call:
 * same layers as n_blocks calls of vr_trace_to_layer in block_ids order, launched so that consecutive
 * blocks overlap on the GPU.                                                                    */
VR_API vr_status vr_trace_blocks_to_layers(vr_ctx* ctx, int n_blocks, const int* block_ids,
                                           const vr_camera* cam, float sample_dist, float range_min,
                                           float range_max, int use_canvas_depth);
/* One rank: PartialCompositor::composite + partials_to_canvas (VolumeRenderer.cpp:580-595).
 * canvas_is_clear as for vr_partials_composite_to_canvas.                                       */
VR_API vr_status vr_layers_composite_to_canvas(vr_ctx* ctx, const vr_camera* cam, int canvas_is_clear);
/* The layers as the reference's compact list (what StructuredWrapper::render returns): replaces
 * the context's partial list; order unspecified.                                                */
VR_API vr_status vr_layers_to_partials(vr_ctx* ctx);

/* ------------------------------------------------------------------ (3) image compositing
 * Image::Init (Image.hpp:80-113): quantise the device canvas to RGBA8 (truncation) + depth
 * (negative -> |d|) into the context's exchange image.                                         */
VR_API vr_status vr_image_from_canvas(vr_ctx* ctx);
VR_API vr_status vr_image_download(vr_ctx* ctx, uint8_t* rgba, float* depth); /* syncs */
/* ImageCompositor::OrderedComposite over n_layers images resident on THIS device
 * (layer l at rgba + l*layer_stride_px*4, depth + l*layer_stride_px): sort by vis_order, fold
 * front-to-back with the truncating uint8 over operator, per pixel, in one pass.  All pointers
 * are device pointers; out may alias layer 0.                                                  */
VR_API vr_status vr_fold_images_dev(vr_ctx* ctx, const uint8_t* rgba, const float* depth,
                                    size_t layer_stride_px, const int* vis_order_host,
                                    int n_layers, size_t n_pixels, uint8_t* out_rgba,
                                    float* out_depth);
/* Host-buffer form of Compositor::{AddImage x n, Composite} in VIS_ORDER_BLEND mode
 * (Compositor.cpp:146-183,221-234): n float images -> composited RGBA8 + depth.               */
VR_API vr_status vr_composite_images(vr_ctx* ctx, const float* rgba, const float* depth,
                                     const int* vis_order, int n_images, int width, int height,
                                     uint8_t* out_rgba, float* out_depth);
/* ImageCompositor::ZBufferComposite (ImageCompositor.hpp:49-76), device pointers, in place.   */
VR_API vr_status vr_zbuffer_composite_dev(vr_ctx* ctx, uint8_t* front_rgba, float* front_depth,
                                          const uint8_t* rgba, const float* depth,
                                          size_t n_pixels);
/* Host-buffer form of Compositor::{AddImage x n, Composite} in Z_BUFFER_SURFACE mode for one rank
 * (Compositor.cpp:146-160: images are z-selected into the first one as they are added).         */
VR_API vr_status vr_composite_zbuffer(vr_ctx* ctx, const float* rgba, const float* depth,
                                      int n_images, int width, int height, uint8_t* out_rgba,
                                      float* out_depth);
/* Renderer::ImageToCanvas (Renderer.cpp:265-283): composited RGBA8 (device) -> device canvas.  */
VR_API vr_status vr_image_to_canvas_dev(vr_ctx* ctx, const uint8_t* rgba, const float* depth);

/* ------------------------------------------------------------------ (4) partial compositing
 * PartialCompositor::composite_partials (PartialCompositor.cpp:329-488) on the context's
 * device partial list: order by (pixel, depth) [ties: list order], fold front-to-back per
 * pixel with VolumePartial::blend; leaves <= 1 partial per covered pixel.                      */
VR_API vr_status vr_partials_composite(vr_ctx* ctx);
/* The two calls above and below fused for one rank (the body of the render loop at
 * VolumeRenderer.cpp:580-595): the fold writes the canvas pixel as it finishes it.
 * canvas_is_clear != 0: the frame starts from Canvas::Clear -- every pixel is written, no prior
 * vr_canvas_clear needed; 0: blend over the context's canvas as it is.  The composited list is
 * left in the context as after vr_partials_composite.                                          */
VR_API vr_status vr_partials_composite_to_canvas(vr_ctx* ctx, const vr_camera* cam, int canvas_is_clear);
/* partials_to_canvas (VolumeRenderer.cpp:287-391) from the composited list onto the device
 * canvas (which must have the frame's width/height).                                           */
VR_API vr_status vr_partials_to_canvas(vr_ctx* ctx, const vr_camera* cam);
/* Host-buffer form of PartialCompositor::composite for one rank: n_in partials in (any order),
 * composited partials out (capacity n_in).                                                     */
VR_API vr_status vr_composite_partials(vr_ctx* ctx, const vr_partial* in, size_t n_in, int width,
                                       int height, vr_partial* out, size_t* n_out);

/* ------------------------------------------------------------------ multi-GPU exchange
 * One process per GPU.  Every rank allocates one exchange arena, publishes its CUDA IPC handle
 * through whatever transport the host has (MPI_Allgather in Ascent, torch.distributed in the
 * harness) and maps its peers' arenas; after that images/partials move GPU-to-GPU over NVLink
 * inside the compositing kernels themselves (direct-send: DirectSendCompositor.cpp:121-181,
 * vtkh_diy_partial_redistribute.hpp:58-152) -- no host staging, no MPI in the data path.
 * max_pixels / max_partials size the arena (largest frame; longest per-rank partial list, resp.
 * the sum over a rank's blocks of (screen-rectangle area + 4) when ray layers are used; 0
 * partials = image path only) and MUST be identical on every rank: a rank addresses its peers'
 * arenas with its own layout (vr_comm_connect checks and fails otherwise).
 * Failure behaviour of the collectives: every cross-GPU wait inside the exchange kernels is bounded
 * (20 s; VR_COMM_TIMEOUT_MS overrides, 0 = unbounded), and a rank that hits a rank-local error in a
 * collective call (its partial list or layers do not fit the arena, ...) returns that error AND
 * releases its peers: their kernels skip the frame and their next synchronising call
 * (vr_synchronize, vr_canvas_download) returns VR_ERR_STATE naming the exchange and the rank.      */
#define VR_IPC_HANDLE_BYTES 64
VR_API vr_status vr_comm_init(vr_ctx* ctx, int rank, int n_ranks, size_t max_pixels,
                              size_t max_partials, void* handle_out /* VR_IPC_HANDLE_BYTES */);
VR_API vr_status vr_comm_connect(vr_ctx* ctx, const void* all_handles /* n_ranks * 64 B */);
/* The single-process form of vr_comm_connect (SURVEY 8(b) deployment shape (i): one process, one context per
 * GPU of the node): ctxs[r] must have been vr_comm_init'ed as rank r of n_ranks with identical sizes.  No IPC
 * handles travel; contexts on different devices reach each other's arenas through peer access.  Afterwards
 * every collective entry point must be issued on ALL contexts before any of them is synchronised.        */
VR_API vr_status vr_comm_connect_local(vr_ctx* const* ctxs, int n_ranks);
/* Path A, all ranks call collectively once per image: direct-send exchange of the quantised
 * images + visibility-ordered fold + gather to rank 0, fused in one kernel per rank (peer
 * loads of the owned tile from every rank, peer store of the folded tile into rank 0).
 * vis_order[r] = composite order of rank r's image (VolumeRenderer.cpp:833-867).
 * On rank 0 the composited image is left in the context (vr_image_result_*).
 * The exchange is queued on an internal stream, after everything queued on the context so far; the
 * next image-only vr_trace_to_image (no VR_FRAME_WRITE_CANVAS) may overlap it -- the folded pixels drain
 * into rank 0 while the sampler is already on the next frame -- and every other entry point first waits
 * for it on the context's stream, so callers see plain in-order semantics.                        */
VR_API vr_status vr_comm_composite_images(vr_ctx* ctx, const int* vis_order);
/* Make the context's stream wait for the exchange queued last (no host synchronisation): for callers
 * that time or order their own work on that stream against the exchange.                         */
VR_API vr_status vr_comm_join(vr_ctx* ctx);
/* The same plus Renderer::ImageToCanvas on rank 0 (vr_image_result_to_canvas) folded into the
 * exchange: rank 0 converts the pixels no rank covers while the others are still in flight.    */
VR_API vr_status vr_comm_composite_images_to_canvas(vr_ctx* ctx, const int* vis_order);
/* A batch of renders on path A, all ranks collectively (the loop over the renders of a batch in
 * vtkh::Scene::Render, Scene.cpp:133-149, around RenderOneDomainPerRank + Composite,
 * VolumeRenderer.cpp:482-536, 652-688): for k in [0, n_frames): vr_trace_to_image(block_id, cams[k]) pushed into
 * the exchange, then vr_comm_composite_images_to_canvas(vis_orders + k * n_ranks).  One ABI crossing for the
 * whole batch; consecutive frames overlap (trace of k+1 and k+2 while k is exchanged).  Rank 0's canvas holds
 * the last frame afterwards; frames_rgba8_host (rank 0, may be NULL; pinned memory makes the copies
 * asynchronous) receives EVERY frame as RGBA8, rows flipped, blended over bg_rgba (NULL: no background) -- the
 * input of Render::Save's PNG encoder (Render.cpp:299-312) -- frame k at offset k * width * height * 4.
 * Call vr_synchronize before reading the host frames.                                             */
VR_API vr_status vr_comm_render_frames(vr_ctx* ctx, int block_id, const vr_camera* cams, int n_frames, int width,
                                       int height, float sample_dist, float range_min, float range_max,
                                       const int* vis_orders, const float* bg_rgba, uint8_t* frames_rgba8_host);
/* Opaque surfaces, all ranks collectively (Compositor Z_BUFFER_SURFACE -> RadixKCompositor::
 * CompositeSurface, RadixKCompositor.cpp:35-180): the same fused exchange with
 * ImageCompositor::ZBufferComposite (ImageCompositor.hpp:49-76) as the per-pixel operator; the nearest
 * fragment's RGBA8 + depth land on rank 0 (vr_image_result_*).  One round of radix k = N over NVLink
 * instead of the reference's rounds of k <= 8, with equal-depth fragments of different ranks resolved
 * exactly as the reference's tree resolves them (the piece-by-piece visiting order of vr_radixk_schedule):
 * the same image for every input with depths >= 0.  Fails with VR_ERR_INVALID where the reference throws
 * "Unable to decompose domain" (frame too small for one block per rank).                          */
VR_API vr_status vr_comm_composite_zbuffer(vr_ctx* ctx);
/* The schedule behind it, host only (no context, no GPU): DIY's RegularDecomposer + RegularSwapPartners(k = 8,
 * distance halving) as RadixKCompositor::CompositeImpl (RadixKCompositor.cpp:138-180) configures them, and
 * reduce_images' balanced splits (:66-92), in closed form.  divisions[2]: blocks along x, y;  lo_x[n], lo_y[n]:
 * first 0-based pixel of each block column / row (INT_MAX beyond divisions[d]; the pixel a piece shares with its
 * lower neighbour counts for the higher gid, which CollectImages pastes later);  seq[n*n]: seq[g*n + i] = the
 * i-th rank whose fragment is z-composited for the pixels block g = cx + divisions[0] * cy ends up owning.
 * Returns 0, or -1 if the frame cannot be decomposed into n_ranks blocks (n_ranks <= 16).             */
VR_API int vr_radixk_schedule(int n_ranks, int width, int height, int divisions[2], int* lo_x, int* lo_y, int* seq);
/* Scene::SynchDepths (Scene.cpp:249-264), all ranks collectively: rank 0's canvas depth replaces every
 * other rank's canvas depth (NVLink pull), so the volume pass stops at the composited surfaces.   */
VR_API vr_status vr_comm_sync_depths(vr_ctx* ctx);
VR_API vr_status vr_image_result_download(vr_ctx* ctx, uint8_t* rgba, float* depth); /* syncs */
VR_API vr_status vr_image_result_to_canvas(vr_ctx* ctx);
/* Path B, collective: redistribute partials by pixel-range owner, sort+fold on the owner,
 * gather to rank 0 (P2-P5).  On rank 0 the result replaces the context's partial list.        */
VR_API vr_status vr_comm_composite_partials(vr_ctx* ctx);
/* Path B with partials_to_canvas fused in, for a frame that starts from a cleared canvas: the owner
 * of a pixel stores the finished canvas pixel straight into rank 0's canvas (20 B per covered pixel
 * over NVLink; no list, no gather).  On rank 0 the context's canvas holds the final image.       */
VR_API vr_status vr_comm_composite_partials_to_canvas(vr_ctx* ctx, const vr_camera* cam);
/* Path B on layers, collective, frame starts from a cleared canvas: every rank owns a round-robin
 * share of 32x8 pixel tiles, pulls the covering layers' entries out of every rank's arena over
 * NVLink, folds and stores finished pixels into rank 0's canvas.  vr_layers_begin must have been
 * called after vr_comm_connect (the layers then live in the exchange arena).                     */
VR_API vr_status vr_comm_layers_composite_to_canvas(vr_ctx* ctx, const vr_camera* cam);
/* Diagnostics (VR_TIMELINE=1 in the environment at vr_comm_init): %globaltimer stamps, in ns, that the
 * kernels of the last image exchange left on this GPU -- [0] fold kernel entered, [1] every rank's image
 * ready, [2] last fold CTA out, [3] rank 0's to-canvas kernel entered, [4] every rank's pixels landed,
 * [5] to-canvas done, [6] end of the last vr_trace_to_image kernel, [7] unused; CTA 0 of the fold
 * kernel: [8] prologue done, [9] own chunks folded, [10] rank 0's clears done, [11] fenced and counted;
 * [12..15] unused.  Syncs.                                                                       */
VR_API vr_status vr_comm_timeline(vr_ctx* ctx, uint64_t out_ns[16]);
/* Device pointers into this rank's arena for transports that move the bytes themselves
 * (NCCL send/recv baseline in the harness).                                                    */
VR_API vr_status vr_image_ptrs(vr_ctx* ctx, void** rgba8_dev, void** depth_dev);

/* ------------------------------------------------------------------ host-side helpers
 * The driver logic of VolumeRenderer that stays on the CPU and the camera
 * subset used by the tracer (K1), exported so hosts need not re-derive them.                   */
VR_API float vr_sample_distance(const double global_bounds[6], float samples);
VR_API void vr_visibility_order(const double* domain_bounds /* n x 6 */, int n_domains,
                                const vr_camera* cam, int* order_out);
VR_API void vr_find_subset(const vr_camera* cam, int width, int height, const double bounds[6],
                           int out_minx_miny_w_h[4]);
/* vtkm::rendering::Camera, 3-D mode, as Ascent drives it -- parse_camera
 * (ascent_runtime_conduit_to_vtkm_parsing.cpp:97-173), Render.cpp:314-347 (ResetToBounds on the scene
 * bounds), and one camera of the cinema orbit (ascent_runtime_rendering_filters.cpp:906-960).  K0.   */
VR_API void vr_camera_default(vr_camera* cam);
VR_API void vr_camera_reset_to_bounds(vr_camera* cam, const double bounds[6]);
VR_API void vr_camera_azimuth(vr_camera* cam, float degrees);
VR_API void vr_camera_elevation(vr_camera* cam, float degrees);
VR_API void vr_camera_zoom(vr_camera* cam, float zoom); /* zoom factor *= 4^zoom */
VR_API void vr_camera_cinema(vr_camera* cam, const double bounds[6], float phi_degrees, float theta_degrees);
/* vtkm::cont::ColorTable::Sample(n_samples, Vec4ui_8) for a table given by its nodes (positions
 * ascending in [0,1]; colour space 0 RGB, 1 CIELAB, 2 diverging/Msh), and convert_table's uint8 -> float4
 * (VolumeRenderer.cpp:64-91).  K8: what vr_set_tf expects.  Either output may be NULL.  Returns
 * VR_ERR_INVALID for unsorted nodes, n_samples < 2 or an unknown colour space.                      */
VR_API vr_status vr_color_table_sample(int color_space, int n_color, const double* color_x,
                                       const float* color_rgb /* n_color x 3 */, int n_alpha,
                                       const double* alpha_x, const float* alpha, int n_samples,
                                       uint8_t* rgba8_out, float* rgba_out);
/* VolumeRenderer::CorrectOpacity for one alpha node (VolumeRenderer.cpp:448-466).  V3.           */
VR_API float vr_correct_opacity(float alpha, float samples);
/* Synthetic braid field (Conduit blueprint::mesh::examples::braid) written straight into device
 * memory: window (i0,j0,k0)+(nx,ny,nz) of a (gx,gy,gz) global grid.  Bench input generator.    */
VR_API vr_status vr_synth_braid_dev(vr_ctx* ctx, void* field_dev, int dtype, const int n[3],
                                    const int start[3], const int global[3]);

#ifdef __cplusplus
}
#endif
#endif /* VR_B200_H */
