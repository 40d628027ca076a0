"""ctypes binding of libvr_b200.so (include/vr_b200.h).

The product path has no CPU fallback: if the shared object is missing or no B200 is visible,
loading / vr_create raises.  Nothing here imports the oracle.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("VR_LIB_NAME", "libvr_b200.so"))

VR_OK = 0
VR_F32, VR_F64 = 0, 1
VR_POINT, VR_CELL = 0, 1
VR_TETRA, VR_HEXAHEDRON = 10, 12
VR_HOST, VR_DEVICE, VR_HOST_MAPPED, VR_HOST_STAGED = 0, 1, 2, 3
IPC_HANDLE_BYTES = 64
FRAME_WRITE_CANVAS, FRAME_NO_CLEAR, FRAME_AHEAD, FRAME_PUSH = 1, 2, 4, 8

# every symbol include/vr_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "vr_create", "vr_destroy", "vr_last_error", "vr_set_stream", "vr_synchronize",
    "vr_kernel_launches", "vr_block_uniform", "vr_block_rectilinear", "vr_block_free",
    "vr_block_bounds", "vr_block_staged_bytes", "vr_set_tf", "vr_canvas_clear", "vr_canvas_upload", "vr_canvas_download",
    "vr_canvas_ptrs", "vr_canvas_blend_background", "vr_canvas_download_rgba8", "vr_trace_to_canvas", "vr_render_image", "vr_trace_to_image", "vr_partials_begin",
    "vr_trace_to_partials", "vr_partials_count", "vr_partials_download", "vr_render_partials",
    "vr_free", "vr_layers_begin", "vr_trace_to_layer", "vr_trace_blocks_to_layers", "vr_layers_composite_to_canvas",
    "vr_layers_to_partials", "vr_comm_layers_composite_to_canvas", "vr_image_from_canvas", "vr_image_download", "vr_fold_images_dev",
    "vr_composite_images", "vr_composite_zbuffer", "vr_zbuffer_composite_dev", "vr_image_to_canvas_dev",
    "vr_partials_composite", "vr_partials_composite_to_canvas", "vr_partials_to_canvas", "vr_composite_partials", "vr_comm_init",
    "vr_comm_connect", "vr_comm_composite_images", "vr_comm_composite_images_to_canvas",
    "vr_image_result_download", "vr_comm_composite_zbuffer", "vr_comm_sync_depths",
    "vr_image_result_to_canvas", "vr_comm_composite_partials", "vr_comm_composite_partials_to_canvas",
    "vr_image_ptrs",
    "vr_sample_distance", "vr_visibility_order", "vr_find_subset", "vr_synth_braid_dev",
    "vr_camera_default", "vr_camera_reset_to_bounds", "vr_camera_azimuth", "vr_camera_elevation",
    "vr_camera_zoom", "vr_camera_cinema", "vr_color_table_sample", "vr_correct_opacity", "vr_comm_timeline",
    "vr_comm_join", "vr_comm_render_frames", "vr_comm_connect_local", "vr_field_gather_strided", "vr_field_free", "vr_radixk_schedule", "vr_canvas_encode_png", "vr_png_bound", "vr_block_unstructured", "vr_canvas_download_rect", "vr_partials_append",
    "vr_set_first_sample_offset",
]


class VRError(RuntimeError):
    """Raised for any non-zero vr_status (the vtk-h wrapper would throw vtkh::Error)."""


class CameraStruct(C.Structure):
    """vr_camera"""
    _fields_ = [("position", C.c_float * 3), ("look_at", C.c_float * 3), ("up", C.c_float * 3),
                ("fov", C.c_float), ("zoom", C.c_float), ("xpan", C.c_float), ("ypan", C.c_float),
                ("near_plane", C.c_float), ("far_plane", C.c_float)]


PARTIAL_DTYPE = np.dtype([("pixel_id", "<i4"), ("depth", "<f4"), ("rgb", "<f4", (3,)),
                          ("alpha", "<f4")])

_lib = None


def load():
    """Load the shared object (building is __graft_entry__.build()'s / ascent_b200.build's job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VRError("libvr_b200.so is not built (python -m ascent_b200.build); "
                      "there is no CPU fallback for the volume-render path")
    lib = C.CDLL(LIB_PATH)
    vp, ip, fp, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double)
    cam = C.POINTER(CameraStruct)
    sz = C.c_size_t
    sig = {
        "vr_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "vr_destroy": (None, [vp]),
        "vr_last_error": (C.c_char_p, [vp]),
        "vr_set_stream": (C.c_int, [vp, vp]),
        "vr_synchronize": (C.c_int, [vp]),
        "vr_kernel_launches": (C.c_uint64, [vp]),
        "vr_block_uniform": (C.c_int, [vp, C.c_int, ip, fp, fp, vp, C.c_int, C.c_int, C.c_int]),
        "vr_block_rectilinear": (C.c_int, [vp, C.c_int, ip, dp, dp, dp, vp, C.c_int, C.c_int, C.c_int]),
        "vr_block_free": (C.c_int, [vp, C.c_int]),
        "vr_block_bounds": (C.c_int, [vp, C.c_int, dp]),
        "vr_set_tf": (C.c_int, [vp, fp, C.c_int]),
        "vr_canvas_clear": (C.c_int, [vp, C.c_int, C.c_int]),
        "vr_canvas_upload": (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
        "vr_canvas_download": (C.c_int, [vp, vp, vp]),
        "vr_canvas_ptrs": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp)]),
        "vr_trace_to_canvas": (C.c_int, [vp, C.c_int, cam, C.c_float, C.c_float, C.c_float, C.c_int]),
        "vr_render_image": (C.c_int, [vp, C.c_int, cam, C.c_int, C.c_int, C.c_float, C.c_float,
                                      C.c_float, vp, vp]),
        "vr_trace_to_image": (C.c_int, [vp, C.c_int, cam, C.c_int, C.c_int, C.c_float, C.c_float,
                                        C.c_float, C.c_int]),
        "vr_canvas_blend_background": (C.c_int, [vp, C.POINTER(C.c_float)]),
        "vr_canvas_download_rgba8": (C.c_int, [vp, C.POINTER(C.c_float), C.c_int, vp]),
        "vr_block_staged_bytes": (C.c_int, [vp, C.c_int, C.POINTER(sz)]),
        "vr_partials_begin": (C.c_int, [vp, C.c_int, C.c_int]),
        "vr_trace_to_partials": (C.c_int, [vp, C.c_int, cam, C.c_float, C.c_float, C.c_float, C.c_int]),
        "vr_partials_count": (C.c_int, [vp, C.POINTER(sz)]),
        "vr_partials_download": (C.c_int, [vp, vp, sz, C.POINTER(sz)]),
        "vr_render_partials": (C.c_int, [vp, C.c_int, cam, C.c_int, C.c_int, C.c_float, C.c_float,
                                         C.c_float, vp, C.POINTER(vp), C.POINTER(sz)]),
        "vr_free": (None, [vp]),
        "vr_layers_begin": (C.c_int, [vp, C.c_int, C.c_int]),
        "vr_trace_to_layer": (C.c_int, [vp, C.c_int, cam, C.c_float, C.c_float, C.c_float, C.c_int]),
        "vr_trace_blocks_to_layers": (C.c_int, [vp, C.c_int, ip, cam, C.c_float, C.c_float, C.c_float, C.c_int]),
        "vr_layers_composite_to_canvas": (C.c_int, [vp, cam, C.c_int]),
        "vr_layers_to_partials": (C.c_int, [vp]),
        "vr_comm_layers_composite_to_canvas": (C.c_int, [vp, cam]),
        "vr_image_from_canvas": (C.c_int, [vp]),
        "vr_image_download": (C.c_int, [vp, vp, vp]),
        "vr_fold_images_dev": (C.c_int, [vp, vp, vp, sz, ip, C.c_int, sz, vp, vp]),
        "vr_composite_images": (C.c_int, [vp, vp, vp, ip, C.c_int, C.c_int, C.c_int, vp, vp]),
        "vr_composite_zbuffer": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]),
        "vr_zbuffer_composite_dev": (C.c_int, [vp, vp, vp, vp, vp, sz]),
        "vr_image_to_canvas_dev": (C.c_int, [vp, vp, vp]),
        "vr_partials_composite": (C.c_int, [vp]),
        "vr_partials_composite_to_canvas": (C.c_int, [vp, cam, C.c_int]),
        "vr_partials_to_canvas": (C.c_int, [vp, cam]),
        "vr_composite_partials": (C.c_int, [vp, vp, sz, C.c_int, C.c_int, vp, C.POINTER(sz)]),
        "vr_comm_init": (C.c_int, [vp, C.c_int, C.c_int, sz, sz, vp]),
        "vr_comm_connect": (C.c_int, [vp, vp]),
        "vr_comm_composite_images": (C.c_int, [vp, ip]),
        "vr_comm_composite_images_to_canvas": (C.c_int, [vp, ip]),
        "vr_image_result_download": (C.c_int, [vp, vp, vp]),
        "vr_comm_composite_zbuffer": (C.c_int, [vp]),
        "vr_comm_sync_depths": (C.c_int, [vp]),
        "vr_image_result_to_canvas": (C.c_int, [vp]),
        "vr_comm_composite_partials": (C.c_int, [vp]),
        "vr_comm_composite_partials_to_canvas": (C.c_int, [vp, cam]),
        "vr_image_ptrs": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp)]),
        "vr_sample_distance": (C.c_float, [dp, C.c_float]),
        "vr_set_first_sample_offset": (C.c_int, [vp, C.c_float, C.c_float]),
        "vr_visibility_order": (None, [dp, C.c_int, cam, ip]),
        "vr_find_subset": (None, [cam, C.c_int, C.c_int, dp, ip]),
        "vr_synth_braid_dev": (C.c_int, [vp, vp, C.c_int, ip, ip, ip]),
        "vr_camera_default": (None, [cam]),
        "vr_camera_reset_to_bounds": (None, [cam, dp]),
        "vr_camera_azimuth": (None, [cam, C.c_float]),
        "vr_camera_elevation": (None, [cam, C.c_float]),
        "vr_camera_zoom": (None, [cam, C.c_float]),
        "vr_camera_cinema": (None, [cam, dp, C.c_float, C.c_float]),
        "vr_color_table_sample": (C.c_int, [C.c_int, C.c_int, dp, fp, C.c_int, dp, fp, C.c_int, vp, fp]),
        "vr_correct_opacity": (C.c_float, [C.c_float, C.c_float]),
        "vr_comm_timeline": (C.c_int, [vp, C.POINTER(C.c_uint64)]),
        "vr_comm_join": (C.c_int, [vp]),
        "vr_comm_connect_local": (C.c_int, [C.POINTER(vp), C.c_int]),
        "vr_field_gather_strided": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(vp)]),
        "vr_field_free": (C.c_int, [vp, vp]),
        "vr_canvas_encode_png": (C.c_int, [vp, fp, vp, C.c_size_t, C.POINTER(C.c_size_t)]),
        "vr_png_bound": (C.c_size_t, [C.c_int, C.c_int]),
        "vr_partials_append": (C.c_int, [vp, vp, C.c_size_t, C.c_int]),
        "vr_canvas_download_rect": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
        "vr_block_unstructured": (C.c_int, [vp, C.c_int, C.c_size_t, vp, C.c_int, C.c_size_t, C.c_int, vp, C.c_int, vp,
                                            C.c_int, C.c_int, C.c_int]),
        "vr_radixk_schedule": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "vr_comm_render_frames": (C.c_int, [vp, C.c_int, C.POINTER(CameraStruct), C.c_int, C.c_int, C.c_int, C.c_float,
                                            C.c_float, C.c_float, C.POINTER(C.c_int), fp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _i3(v):
    return (C.c_int * 3)(*[int(x) for x in v])


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def _d6(v):
    return (C.c_double * 6)(*[float(x) for x in v])


def comm_connect_local(ctxs):
    """vr_comm_connect_local: the contexts of ONE process (one per GPU), each already comm_init'ed as rank r"""
    arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    st = load().vr_comm_connect_local(arr, len(ctxs))
    if st != 0:
        raise VRError("vr_comm_connect_local failed (%d): %s" % (
            st, "; ".join(c.lib.vr_last_error(c.h).decode() for c in ctxs)))


def as_camera(cam):
    """Accept a CameraStruct, any ctypes struct with the same layout (the oracle's), or an object
    with a ``to_struct()`` method."""
    if isinstance(cam, CameraStruct):
        return cam
    if hasattr(cam, "to_struct"):
        return cam.to_struct()
    out = CameraStruct()
    C.memmove(C.byref(out), C.byref(cam), C.sizeof(CameraStruct))
    return out


def sample_distance(global_bounds, samples):
    return float(load().vr_sample_distance(_d6(global_bounds), C.c_float(samples)))


def visibility_order(domain_bounds, cam):
    db = np.ascontiguousarray(domain_bounds, np.float64).reshape(-1, 6)
    out = np.zeros(db.shape[0], np.int32)
    load().vr_visibility_order(db.ctypes.data_as(C.POINTER(C.c_double)), db.shape[0],
                               C.byref(as_camera(cam)), out.ctypes.data_as(C.POINTER(C.c_int)))
    return out


def radixk_schedule(n_ranks, W, H):
    """vr_radixk_schedule: {"divisions", "lo" [x starts, y starts], "seq" per block} or None where the
    reference cannot decompose the frame."""
    div = (C.c_int * 2)()
    lo_x = (C.c_int * n_ranks)()
    lo_y = (C.c_int * n_ranks)()
    seq = (C.c_int * (n_ranks * n_ranks))()
    if load().vr_radixk_schedule(n_ranks, W, H, div, lo_x, lo_y, seq) != 0:
        return None
    return {"divisions": list(div), "lo": [list(lo_x)[:div[0]], list(lo_y)[:div[1]]],
            "seq": [list(seq[g * n_ranks:(g + 1) * n_ranks]) for g in range(n_ranks)]}


def find_subset(cam, W, H, bounds):
    out = (C.c_int * 4)()
    load().vr_find_subset(C.byref(as_camera(cam)), W, H, _d6(bounds), out)
    return tuple(out)


class Context:
    """One vr_ctx == one GPU.  Thin, explicit wrapper: one method per C entry point."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        st = self.lib.vr_create(int(device), C.byref(h))
        if st != VR_OK:
            raise VRError(self.lib.vr_last_error(None).decode())
        self.h = h
        self.device = device
        self._keep = {}

    def close(self):
        if getattr(self, "h", None):
            self.lib.vr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != VR_OK:
            raise VRError(self.lib.vr_last_error(self.h).decode())

    # -- context
    def set_stream(self, cuda_stream):
        self._ck(self.lib.vr_set_stream(self.h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._ck(self.lib.vr_synchronize(self.h))

    def kernel_launches(self):
        return int(self.lib.vr_kernel_launches(self.h))

    # -- blocks
    def field_gather_strided(self, src, n_values, element_stride, element_offset=0, device_ptr=None, dtype=None):
        """vr_field_gather_strided: src = host numpy array (the whole strided buffer) or device_ptr; returns the
        dense device pointer (int) to publish with where=VR_DEVICE and to release with field_free"""
        out = C.c_void_p()
        if device_ptr is not None:
            self._ck(self.lib.vr_field_gather_strided(self.h, C.c_void_p(device_ptr), VR_DEVICE, dtype, n_values,
                                                      element_stride, element_offset, C.byref(out)))
        else:
            a = np.ascontiguousarray(src)
            dt = VR_F32 if a.dtype == np.float32 else VR_F64
            self._ck(self.lib.vr_field_gather_strided(self.h, a.ctypes.data, VR_HOST, dt, n_values, element_stride,
                                                      element_offset, C.byref(out)))
        return out.value

    def field_free(self, dense_ptr):
        self._ck(self.lib.vr_field_free(self.h, C.c_void_p(dense_ptr)))

    def block_unstructured(self, block_id, points, conn, field, assoc=VR_POINT):
        """vr_block_unstructured from host arrays: points [n,3] f32/f64, conn [n_cells, 8|4] int32/int64, field f32/f64"""
        pts = np.ascontiguousarray(points)
        assert pts.dtype in (np.float32, np.float64) and pts.ndim == 2 and pts.shape[1] == 3
        cn = np.ascontiguousarray(conn)
        assert cn.dtype in (np.int32, np.int64) and cn.ndim == 2 and cn.shape[1] in (4, 8)
        f = np.ascontiguousarray(field)
        assert f.dtype in (np.float32, np.float64)
        self._ck(self.lib.vr_block_unstructured(
            self.h, block_id, pts.shape[0], pts.ctypes.data, VR_F32 if pts.dtype == np.float32 else VR_F64, cn.shape[0],
            VR_HEXAHEDRON if cn.shape[1] == 8 else VR_TETRA, cn.ctypes.data, 32 if cn.dtype == np.int32 else 64,
            f.ctypes.data, VR_F32 if f.dtype == np.float32 else VR_F64, assoc, VR_HOST))

    def block_uniform(self, block_id, dims, origin, spacing, field, assoc=VR_POINT, device_ptr=None,
                      dtype=None, host_mapped=False, staged=False):
        """host_mapped / staged: `field` is page-locked mapped host memory (e.g. a pinned torch
        tensor's numpy view), sampled in place over PCIe (VR_HOST_MAPPED) or staged on demand, only
        the 128-byte lines the rays touch (VR_HOST_STAGED), instead of copied."""
        if device_ptr is not None:
            ptr, dt, where = C.c_void_p(device_ptr), dtype, VR_DEVICE
        else:
            field = np.ascontiguousarray(field)
            assert field.dtype in (np.float32, np.float64), "fields are f32 or f64"
            ptr, dt, where = C.c_void_p(field.ctypes.data), (VR_F64 if field.dtype == np.float64 else VR_F32), \
                (VR_HOST_STAGED if staged else VR_HOST_MAPPED if host_mapped else VR_HOST)
        self._ck(self.lib.vr_block_uniform(self.h, block_id, _i3(dims), _f3(origin), _f3(spacing), ptr,
                                           dt, assoc, where))

    def block_rectilinear(self, block_id, dims, axes, field, assoc=VR_POINT, device_ptr=None,
                          dtype=None, host_mapped=False, staged=False):
        ax = [np.ascontiguousarray(a, np.float64) for a in axes]
        if device_ptr is not None:
            ptr, dt, where = C.c_void_p(device_ptr), dtype, VR_DEVICE
        else:
            field = np.ascontiguousarray(field)
            assert field.dtype in (np.float32, np.float64), "fields are f32 or f64"
            ptr, dt, where = C.c_void_p(field.ctypes.data), (VR_F64 if field.dtype == np.float64 else VR_F32), \
                (VR_HOST_STAGED if staged else VR_HOST_MAPPED if host_mapped else VR_HOST)
        dpp = C.POINTER(C.c_double)
        self._ck(self.lib.vr_block_rectilinear(self.h, block_id, _i3(dims), ax[0].ctypes.data_as(dpp),
                                               ax[1].ctypes.data_as(dpp), ax[2].ctypes.data_as(dpp),
                                               ptr, dt, assoc, where))

    def block_staged_bytes(self, block_id):
        n = C.c_size_t(0)
        self._ck(self.lib.vr_block_staged_bytes(self.h, block_id, C.byref(n)))
        return int(n.value)

    def block_from_domain(self, block_id, dom):
        assoc = VR_CELL if dom.get("assoc") == "cell" else VR_POINT
        if dom["kind"] == "uniform":
            self.block_uniform(block_id, dom["dims"], dom["origin"], dom["spacing"], dom["field"], assoc)
        else:
            self.block_rectilinear(block_id, dom["dims"], dom["axes"], dom["field"], assoc)

    def block_free(self, block_id):
        self._ck(self.lib.vr_block_free(self.h, block_id))

    def block_bounds(self, block_id):
        out = (C.c_double * 6)()
        self._ck(self.lib.vr_block_bounds(self.h, block_id, out))
        return np.array(out[:], np.float64)

    # -- transfer function
    def set_tf(self, lut):
        lut = np.ascontiguousarray(lut, np.float32)
        assert lut.ndim == 2 and lut.shape[1] == 4
        self._ck(self.lib.vr_set_tf(self.h, lut.ctypes.data_as(C.POINTER(C.c_float)), lut.shape[0]))

    def set_first_sample_offset(self, abs_offset=0.0, extent_rel=1e-4):
        """first sample at entry + abs_offset + extent_rel * |block extent| (default: VTK-m's meshEpsilon)"""
        self._ck(self.lib.vr_set_first_sample_offset(self.h, C.c_float(abs_offset), C.c_float(extent_rel)))

    # -- canvas
    def canvas_clear(self, W, H):
        self._ck(self.lib.vr_canvas_clear(self.h, W, H))

    def canvas_encode_png(self, W, H, bg=None, out=None):
        """vr_canvas_encode_png: the canvas as PNG file bytes (encoded on the device).  out: a uint8 host buffer of
        at least vr_png_bound(W, H) bytes (e.g. pinned) -- then a view of it is returned instead of a copy."""
        cap = int(self.lib.vr_png_bound(W, H))
        buf = np.empty(cap, np.uint8) if out is None else out
        assert buf.dtype == np.uint8 and buf.size >= cap
        n = C.c_size_t(0)
        b = None if bg is None else np.ascontiguousarray(bg, np.float32).ctypes.data_as(C.POINTER(C.c_float))
        self._ck(self.lib.vr_canvas_encode_png(self.h, b, buf.ctypes.data, buf.size, C.byref(n)))
        return buf[:n.value].tobytes() if out is None else buf[:n.value]

    def canvas_upload(self, W, H, rgba, depth):
        rgba = np.ascontiguousarray(rgba, np.float32)
        depth = np.ascontiguousarray(depth, np.float32)
        self._ck(self.lib.vr_canvas_upload(self.h, W, H, rgba.ctypes.data, depth.ctypes.data))

    def canvas_download(self, W, H, rgba=None, depth=None):
        rgba = np.empty((H * W, 4), np.float32) if rgba is None else rgba
        depth = np.empty(H * W, np.float32) if depth is None else depth
        self._ck(self.lib.vr_canvas_download(self.h, rgba.ctypes.data, depth.ctypes.data))
        return rgba, depth

    def canvas_download_rect(self, rect, rgba, depth):
        """vr_canvas_download_rect: rect = (x0, y0, x1, y1); rgba [W*H, 4] / depth [W*H] full-frame host arrays"""
        self._ck(self.lib.vr_canvas_download_rect(self.h, int(rect[0]), int(rect[1]), int(rect[2]), int(rect[3]),
                                                  rgba.ctypes.data, depth.ctypes.data))

    def canvas_blend_background(self, bg):
        b = (C.c_float * 4)(*[float(x) for x in bg])
        self._ck(self.lib.vr_canvas_blend_background(self.h, b))

    def canvas_download_rgba8(self, W, H, bg=None, flip=True, out=None):
        out = np.empty((H, W, 4), np.uint8) if out is None else out
        b = (C.c_float * 4)(*[float(x) for x in bg]) if bg is not None else None
        self._ck(self.lib.vr_canvas_download_rgba8(self.h, b, int(flip), out.ctypes.data))
        return out

    def canvas_ptrs(self):
        a, b = C.c_void_p(), C.c_void_p()
        self._ck(self.lib.vr_canvas_ptrs(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- path A
    def trace_to_canvas(self, block_id, cam, sample_dist, rmin, rmax, use_canvas_depth=False):
        self._ck(self.lib.vr_trace_to_canvas(self.h, block_id, C.byref(as_camera(cam)), sample_dist,
                                             rmin, rmax, int(use_canvas_depth)))

    def render_image(self, block_id, cam, W, H, sample_dist, rmin, rmax, rgba, depth):
        assert rgba.dtype == np.float32 and depth.dtype == np.float32
        self._ck(self.lib.vr_render_image(self.h, block_id, C.byref(as_camera(cam)), W, H, sample_dist,
                                          rmin, rmax, rgba.ctypes.data, depth.ctypes.data))

    def trace_to_image(self, block_id, cam, W, H, sample_dist, rmin, rmax, write_canvas=False,
                       no_clear=False, ahead=False, push=False):
        flags = (FRAME_WRITE_CANVAS if write_canvas else 0) | (FRAME_NO_CLEAR if no_clear else 0) | \
            (FRAME_AHEAD if ahead else 0) | (FRAME_PUSH if push else 0)
        self._ck(self.lib.vr_trace_to_image(self.h, block_id, C.byref(as_camera(cam)), W, H, sample_dist,
                                            rmin, rmax, flags))

    # -- path B
    def partials_begin(self, W, H):
        self._ck(self.lib.vr_partials_begin(self.h, W, H))

    def trace_to_partials(self, block_id, cam, sample_dist, rmin, rmax, use_canvas_depth=False):
        self._ck(self.lib.vr_trace_to_partials(self.h, block_id, C.byref(as_camera(cam)), sample_dist,
                                               rmin, rmax, int(use_canvas_depth)))

    def partials_append(self, partials):
        """vr_partials_append: a host array of PARTIAL_DTYPE from another producer joins the frame's list"""
        p = np.ascontiguousarray(partials)
        assert p.dtype == PARTIAL_DTYPE
        self._ck(self.lib.vr_partials_append(self.h, p.ctypes.data, p.size, VR_HOST))

    def partials_count(self):
        n = C.c_size_t()
        self._ck(self.lib.vr_partials_count(self.h, C.byref(n)))
        return int(n.value)

    def partials_download(self):
        n = self.partials_count()
        out = np.zeros(max(n, 1), PARTIAL_DTYPE)
        got = C.c_size_t()
        self._ck(self.lib.vr_partials_download(self.h, out.ctypes.data, n, C.byref(got)))
        return out[:got.value]

    def render_partials(self, block_id, cam, W, H, sample_dist, rmin, rmax, depth_in=None):
        p, n = C.c_void_p(), C.c_size_t()
        dptr = None if depth_in is None else np.ascontiguousarray(depth_in, np.float32).ctypes.data
        self._ck(self.lib.vr_render_partials(self.h, block_id, C.byref(as_camera(cam)), W, H,
                                             sample_dist, rmin, rmax, dptr, C.byref(p), C.byref(n)))
        out = np.zeros(n.value, PARTIAL_DTYPE)
        if n.value:
            C.memmove(out.ctypes.data, p.value, n.value * 24)
        self.lib.vr_free(p)
        return out

    # -- path B on dense ray layers
    def layers_begin(self, W, H):
        self._ck(self.lib.vr_layers_begin(self.h, W, H))

    def trace_to_layer(self, block_id, cam, sample_dist, rmin, rmax, use_canvas_depth=False):
        self._ck(self.lib.vr_trace_to_layer(self.h, block_id, C.byref(as_camera(cam)), sample_dist, rmin,
                                            rmax, int(use_canvas_depth)))

    def trace_blocks_to_layers(self, block_ids, cam, sample_dist, rmin, rmax, use_canvas_depth=False):
        ids = np.ascontiguousarray(block_ids, np.int32)
        self._ck(self.lib.vr_trace_blocks_to_layers(self.h, len(ids), ids.ctypes.data_as(C.POINTER(C.c_int)),
                                                    C.byref(as_camera(cam)), sample_dist, rmin, rmax,
                                                    int(use_canvas_depth)))

    def layers_composite_to_canvas(self, cam, canvas_is_clear=True):
        self._ck(self.lib.vr_layers_composite_to_canvas(self.h, C.byref(as_camera(cam)), int(canvas_is_clear)))

    def layers_to_partials(self):
        self._ck(self.lib.vr_layers_to_partials(self.h))

    def comm_composite_zbuffer(self):
        self._ck(self.lib.vr_comm_composite_zbuffer(self.h))

    def comm_sync_depths(self):
        self._ck(self.lib.vr_comm_sync_depths(self.h))

    def comm_layers_composite_to_canvas(self, cam):
        self._ck(self.lib.vr_comm_layers_composite_to_canvas(self.h, C.byref(as_camera(cam))))

    # -- image compositing
    def image_from_canvas(self):
        self._ck(self.lib.vr_image_from_canvas(self.h))

    def image_download(self, W, H):
        rgba = np.empty((H * W, 4), np.uint8)
        depth = np.empty(H * W, np.float32)
        self._ck(self.lib.vr_image_download(self.h, rgba.ctypes.data, depth.ctypes.data))
        return rgba, depth

    def image_ptrs(self):
        a, b = C.c_void_p(), C.c_void_p()
        self._ck(self.lib.vr_image_ptrs(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def fold_images_dev(self, rgba_ptr, depth_ptr, layer_stride_px, vis_order, n_pixels, out_rgba_ptr,
                        out_depth_ptr):
        vo = np.ascontiguousarray(vis_order, np.int32)
        self._ck(self.lib.vr_fold_images_dev(self.h, rgba_ptr, depth_ptr, layer_stride_px,
                                             vo.ctypes.data_as(C.POINTER(C.c_int)), vo.size, n_pixels,
                                             out_rgba_ptr, out_depth_ptr))

    def composite_images(self, rgba, depth, vis_order, W, H):
        rgba = np.ascontiguousarray(rgba, np.float32)
        depth = np.ascontiguousarray(depth, np.float32)
        vo = np.ascontiguousarray(vis_order, np.int32)
        out = np.empty((H * W, 4), np.uint8)
        od = np.empty(H * W, np.float32)
        self._ck(self.lib.vr_composite_images(self.h, rgba.ctypes.data, depth.ctypes.data,
                                              vo.ctypes.data_as(C.POINTER(C.c_int)), vo.size, W, H,
                                              out.ctypes.data, od.ctypes.data))
        return out, od

    def composite_zbuffer(self, rgba, depth, W, H):
        rgba = np.ascontiguousarray(rgba, np.float32)
        depth = np.ascontiguousarray(depth, np.float32)
        out = np.empty((H * W, 4), np.uint8)
        od = np.empty(H * W, np.float32)
        self._ck(self.lib.vr_composite_zbuffer(self.h, rgba.ctypes.data, depth.ctypes.data,
                                               depth.reshape(-1, H * W).shape[0], W, H,
                                               out.ctypes.data, od.ctypes.data))
        return out, od

    def zbuffer_composite_dev(self, front_rgba, front_depth, rgba, depth, n_pixels):
        self._ck(self.lib.vr_zbuffer_composite_dev(self.h, front_rgba, front_depth, rgba, depth, n_pixels))

    def image_to_canvas_dev(self, rgba_ptr, depth_ptr):
        self._ck(self.lib.vr_image_to_canvas_dev(self.h, rgba_ptr, depth_ptr))

    # -- partial compositing
    def partials_composite(self):
        self._ck(self.lib.vr_partials_composite(self.h))

    def partials_composite_to_canvas(self, cam, canvas_is_clear=True):
        self._ck(self.lib.vr_partials_composite_to_canvas(self.h, C.byref(as_camera(cam)), int(canvas_is_clear)))

    def partials_to_canvas(self, cam):
        self._ck(self.lib.vr_partials_to_canvas(self.h, C.byref(as_camera(cam))))

    def composite_partials(self, partials, W, H):
        p = np.ascontiguousarray(partials)
        assert p.dtype == PARTIAL_DTYPE
        out = np.zeros(max(p.size, 1), PARTIAL_DTYPE)
        n = C.c_size_t()
        self._ck(self.lib.vr_composite_partials(self.h, p.ctypes.data, p.size, W, H, out.ctypes.data,
                                                C.byref(n)))
        return out[:n.value]

    # -- multi-GPU
    def comm_init(self, rank, size, max_pixels, max_partials=0):
        h = (C.c_ubyte * IPC_HANDLE_BYTES)()
        self._ck(self.lib.vr_comm_init(self.h, rank, size, max_pixels, max_partials, h))
        return bytes(h)

    def comm_connect(self, all_handles):
        buf = (C.c_ubyte * len(all_handles)).from_buffer_copy(all_handles)
        self._ck(self.lib.vr_comm_connect(self.h, buf))

    def comm_join(self):
        self._ck(self.lib.vr_comm_join(self.h))

    def comm_render_frames(self, block_id, cams, W, H, sample_dist, rmin, rmax, vis_orders, bg=None, out_rgba8=None):
        """vr_comm_render_frames: the frames of a batch in one ABI call.  cams: sequence of cameras (or a prepared
        ctypes array), vis_orders: (n_frames, n_ranks) int32, out_rgba8: host uint8 array (n_frames, H, W, 4) or None"""
        if isinstance(cams, C.Array):
            arr = cams
        else:
            arr = (CameraStruct * len(cams))(*[as_camera(c) for c in cams])
        vo = np.ascontiguousarray(vis_orders, np.int32)
        bgp = None
        if bg is not None:
            bga = np.ascontiguousarray(bg, np.float32)
            bgp = bga.ctypes.data_as(C.POINTER(C.c_float))
        self._ck(self.lib.vr_comm_render_frames(self.h, block_id, arr, len(arr), W, H, sample_dist, rmin, rmax,
                                                vo.ctypes.data_as(C.POINTER(C.c_int)), bgp,
                                                out_rgba8.ctypes.data if out_rgba8 is not None else None))

    def comm_timeline(self):
        out = (C.c_uint64 * 16)()
        self._ck(self.lib.vr_comm_timeline(self.h, out))
        return [int(x) for x in out]

    def comm_composite_images(self, vis_order):
        vo = np.ascontiguousarray(vis_order, np.int32)
        self._ck(self.lib.vr_comm_composite_images(self.h, vo.ctypes.data_as(C.POINTER(C.c_int))))

    def comm_composite_images_to_canvas(self, vis_order):
        vo = np.ascontiguousarray(vis_order, np.int32)
        self._ck(self.lib.vr_comm_composite_images_to_canvas(self.h, vo.ctypes.data_as(C.POINTER(C.c_int))))

    def image_result_download(self, W, H):
        rgba = np.empty((H * W, 4), np.uint8)
        depth = np.empty(H * W, np.float32)
        self._ck(self.lib.vr_image_result_download(self.h, rgba.ctypes.data, depth.ctypes.data))
        return rgba, depth

    def image_result_to_canvas(self):
        self._ck(self.lib.vr_image_result_to_canvas(self.h))

    def comm_composite_partials(self):
        self._ck(self.lib.vr_comm_composite_partials(self.h))

    def comm_composite_partials_to_canvas(self, cam):
        self._ck(self.lib.vr_comm_composite_partials_to_canvas(self.h, C.byref(as_camera(cam))))

    # -- bench input
    def synth_braid_dev(self, dev_ptr, dtype, n, start, glob):
        self._ck(self.lib.vr_synth_braid_dev(self.h, C.c_void_p(dev_ptr), dtype, _i3(n), _i3(start),
                                             _i3(glob)))
