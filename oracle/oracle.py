"""ctypes doorway onto the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs
may import this module.  The product package ``ascent_b200`` never does.

liboracle.so          = oracle/raycast_oracle.c + oracle/composite_oracle.c (our restatement)
_ref/libapcomp_ref.so = the reference's own apcomp sources (src/libs/apcomp/*.cpp) compiled in
                        place by oracle/Makefile; optional at run time (absent -> ``ref`` is None).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class Camera(C.Structure):
    """Field-for-field the same as ``vr_camera`` in include/vr_b200.h."""
    _fields_ = [("position", C.c_float * 3), ("look_at", C.c_float * 3), ("up", C.c_float * 3),
                ("fov", C.c_float), ("zoom", C.c_float), ("xpan", C.c_float), ("ypan", C.c_float),
                ("near_plane", C.c_float), ("far_plane", C.c_float)]


class Block(C.Structure):
    _fields_ = [("kind", C.c_int), ("dims", C.c_int * 3), ("origin", C.c_float * 3),
                ("spacing", C.c_float * 3), ("ax", C.c_void_p * 3), ("field", C.c_void_p),
                ("field_f64", C.c_int), ("cell_assoc", C.c_int)]


class Rays(C.Structure):
    _fields_ = [("n", C.c_int), ("subset", C.c_int * 4), ("dir", C.POINTER(C.c_float)),
                ("min_dist", C.POINTER(C.c_float)), ("max_dist", C.POINTER(C.c_float)),
                ("dist", C.POINTER(C.c_float)), ("rgba", C.POINTER(C.c_float)),
                ("pixel", C.POINTER(C.c_int64)), ("origin", C.c_float * 3),
                ("n_samples", C.c_int64)]


PARTIAL_DTYPE = np.dtype([("pixel_id", "<i4"), ("depth", "<f4"), ("rgb", "<f4", (3,)),
                          ("alpha", "<f4")])
assert PARTIAL_DTYPE.itemsize == 24


def build(force=False):
    """Compile liboracle.so (and _ref/ when /root/reference is present); make only rebuilds what is
    older than its sources."""
    have = os.path.exists(os.path.join(_HERE, "liboracle.so"))
    try:
        # (make's chatter goes to stderr: bench.py's stdout is one JSON line)
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), stdout=sys.stderr)
    except (OSError, subprocess.CalledProcessError):
        if not have:
            raise


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


build()
lib = C.CDLL(os.path.join(_HERE, "liboracle.so"))
_ref_path = os.path.join(_HERE, "_ref", "libapcomp_ref.so")
ref = C.CDLL(_ref_path) if os.path.exists(_ref_path) else None

lib.orc_correct_opacity.restype = C.c_double
lib.orc_correct_opacity.argtypes = [C.c_double, C.c_float]
lib.orc_sample_distance.restype = C.c_float
lib.orc_sample_distance.argtypes = [C.POINTER(C.c_double), C.c_float]
lib.orc_extract_partials.restype = C.c_int64
lib.orc_composite_partials.restype = C.c_int64
lib.orc_composite_partials.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
lib.orc_partial_owner.restype = C.c_int
lib.orc_num_threads.restype = C.c_int
lib.orc_set_num_threads.argtypes = [C.c_int]
lib.orc_set_first_sample_offset.argtypes = [C.c_float, C.c_float]
lib.orc_set_first_sample_offset.restype = None


def set_first_sample_offset(abs_offset=0.0, extent_rel=1e-4):
    """First sample of the structured sampler at entry + abs_offset + extent_rel * |block extent|.  Default (0, 1e-4):
    VTK-m's meshEpsilon form (the pinned v2.1.0; decided by the newest golden).  (1e-4, 0): the older generation of
    VTK-m that rendered the reference's three pure-volume goldens.  tests/test_oracle_golden.py pins both."""
    lib.orc_set_first_sample_offset(abs_offset, extent_rel)

if ref is not None:
    ref.ref_composite_partials.restype = C.c_longlong


def num_threads():
    return int(lib.orc_num_threads())


def use_all_cores():
    """OpenMP threads = the cores this process may run on, whatever OMP_NUM_THREADS says (torchrun
    exports OMP_NUM_THREADS=1 to its workers)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib.orc_set_num_threads(int(n))
    return num_threads()


# ----------------------------------------------------------------------------- camera (K0)
def make_camera(position, look_at, up, fov=60.0, zoom=1.0, near=0.01, far=1000.0, xpan=0., ypan=0.):
    c = Camera()
    c.position[:] = [float(v) for v in position]
    c.look_at[:] = [float(v) for v in look_at]
    c.up[:] = [float(v) for v in up]
    c.fov, c.zoom, c.xpan, c.ypan = fov, zoom, xpan, ypan
    c.near_plane, c.far_plane = near, far
    return c


def camera_default():
    c = Camera()
    lib.orc_camera_default(C.byref(c))
    return c


def camera_reset_to_bounds(bounds, cam=None):
    c = camera_default() if cam is None else cam
    b = (C.c_double * 6)(*[float(v) for v in bounds])
    lib.orc_camera_reset_to_bounds(C.byref(c), b)
    return c


def camera_azimuth(c, deg):
    lib.orc_camera_azimuth(C.byref(c), C.c_float(deg))


def camera_elevation(c, deg):
    lib.orc_camera_elevation(C.byref(c), C.c_float(deg))


def camera_zoom(c, z):
    lib.orc_camera_zoom(C.byref(c), C.c_float(z))


def camera_cinema(bounds, phi, theta):
    """One camera of the cinema orbit (CinemaManager::create_cinema_cameras)."""
    c = Camera()
    b = (C.c_double * 6)(*[float(v) for v in bounds])
    lib.orc_camera_cinema(C.byref(c), b, C.c_float(phi), C.c_float(theta))
    return c


def cinema_angles(phi, theta):
    """CinemaManager::create_cinema_angles (rendering_filters.cpp:882-893 with the phi/theta ranges
    of :568-574,613-619): phi in [-180, 180), theta in [0, 180), both as float."""
    ph = [float(np.float32(-180.0 + (360.0 / float(phi)) * a)) for a in range(phi)]
    th = [float(np.float32(0.0 + (180.0 / float(theta)) * a)) for a in range(theta)]
    return ph, th


def projview(c, w, h):
    m = np.zeros(16, np.float32)
    lib.orc_projview(C.byref(c), w, h, _ptr(m, C.c_float))
    return m.reshape(4, 4)


def find_subset(c, w, h, bounds):
    out = (C.c_int * 4)()
    b = (C.c_double * 6)(*[float(v) for v in bounds])
    lib.orc_find_subset(C.byref(c), w, h, b, out)
    return tuple(out)


# ----------------------------------------------------------------------------- blocks
class OracleBlock:
    """Keeps the numpy arrays alive next to the C struct."""

    def __init__(self, dims, field, origin=None, spacing=None, axes=None, cell_assoc=False):
        self.b = Block()
        self.b.dims[:] = [int(d) for d in dims]
        self.field = np.ascontiguousarray(field)
        assert self.field.dtype in (np.float32, np.float64)
        self.b.field = self.field.ctypes.data
        self.b.field_f64 = int(self.field.dtype == np.float64)
        self.b.cell_assoc = int(cell_assoc)
        if axes is None:
            self.b.kind = 0
            self.b.origin[:] = [float(np.float32(v)) for v in origin]
            self.b.spacing[:] = [float(np.float32(v)) for v in spacing]
        else:
            self.b.kind = 1
            self.axes = [np.ascontiguousarray(a, dtype=np.float64) for a in axes]
            for k in range(3):
                assert self.axes[k].size == dims[k]
                self.b.ax[k] = self.axes[k].ctypes.data
        n_expected = (np.prod([d - 1 for d in dims]) if cell_assoc else np.prod(dims))
        assert self.field.size == n_expected, (self.field.size, n_expected)

    def bounds(self):
        out = (C.c_double * 6)()
        lib.orc_block_bounds(C.byref(self.b), out)
        return np.array(out[:], np.float64)


def sample_distance(global_bounds, samples):
    b = (C.c_double * 6)(*[float(v) for v in global_bounds])
    return float(lib.orc_sample_distance(b, C.c_float(samples)))


def correct_opacity(alpha, samples):
    return float(lib.orc_correct_opacity(float(alpha), C.c_float(samples)))


class TraceResult:
    def __init__(self, rays):
        n = rays.n
        self.subset = tuple(rays.subset)
        self.n = n
        self.rgba = np.ctypeslib.as_array(rays.rgba, (n, 4)).copy()
        self.pixel = np.ctypeslib.as_array(rays.pixel, (n,)).copy()
        self.min_dist = np.ctypeslib.as_array(rays.min_dist, (n,)).copy()
        self.max_dist = np.ctypeslib.as_array(rays.max_dist, (n,)).copy()
        self.dist = np.ctypeslib.as_array(rays.dist, (n,)).copy()
        self.dir = np.ctypeslib.as_array(rays.dir, (n, 3)).copy()
        self.n_samples = int(rays.n_samples)


def trace_block(block, cam, W, H, lut, sample_dist, rmin, rmax, canvas_depth=None, keep=True):
    """K1+K2+K3+K4/5/6 for one block.  Returns (Rays struct, TraceResult|None); free with
    rays_free()."""
    lut = np.ascontiguousarray(lut, np.float32)
    rays = Rays()
    dptr = None if canvas_depth is None else _ptr(canvas_depth, C.c_float)
    lib.orc_trace_block(C.byref(block.b), C.byref(cam), W, H, _ptr(lut, C.c_float),
                        int(lut.shape[0]), C.c_float(sample_dist), C.c_float(rmin),
                        C.c_float(rmax), dptr, C.byref(rays))
    return rays, (TraceResult(rays) if keep else None)


def rays_free(rays):
    lib.orc_rays_free(C.byref(rays))


class UMesh(C.Structure):
    _fields_ = [("n_points", C.c_int), ("n_cells", C.c_int), ("shape", C.c_int), ("xyz", C.c_void_p),
                ("conn", C.c_void_p), ("field", C.c_void_p), ("field_f64", C.c_int), ("cell_assoc", C.c_int),
                ("bmin", C.c_float * 3), ("bmax", C.c_float * 3), ("ginv", C.c_float * 3), ("g", C.c_int * 3),
                ("bin_start", C.c_void_p), ("bin_cells", C.c_void_p),
                ("n_ext", C.c_int), ("ext_faces", C.c_void_p), ("ext_cell", C.c_void_p)]


class OracleUMesh:
    """N4: an explicit cell set (hexahedra or tetrahedra) with a point or cell field.  Keeps the arrays alive."""

    def __init__(self, points, conn, field, cell_assoc=False):
        self.points = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        self.conn = np.ascontiguousarray(conn, np.int32)
        assert self.conn.ndim == 2 and self.conn.shape[1] in (4, 8)
        self.field = np.ascontiguousarray(field)
        assert self.field.dtype in (np.float32, np.float64)
        assert self.field.size == (self.conn.shape[0] if cell_assoc else self.points.shape[0])
        m = self.m = UMesh()
        m.n_points, m.n_cells, m.shape = self.points.shape[0], self.conn.shape[0], self.conn.shape[1]
        m.xyz, m.conn, m.field = self.points.ctypes.data, self.conn.ctypes.data, self.field.ctypes.data
        m.field_f64 = int(self.field.dtype == np.float64)
        m.cell_assoc = int(cell_assoc)
        lib.orc_umesh_build(C.byref(m))

    def __del__(self):
        try:
            lib.orc_umesh_free(C.byref(self.m))
        except Exception:
            pass

    def bounds(self):
        out = (C.c_double * 6)()
        lib.orc_umesh_bounds(C.byref(self.m), out)
        return np.array(out[:], np.float64)


def trace_umesh(mesh, cam, W, H, lut, sample_dist, rmin, rmax, canvas_depth=None, keep=True,
                structured_conventions=False):
    lut = np.ascontiguousarray(lut, np.float32)
    rays = Rays()
    dptr = None if canvas_depth is None else _ptr(canvas_depth, C.c_float)
    lib.orc_trace_umesh(C.byref(mesh.m), C.byref(cam), W, H, _ptr(lut, C.c_float), int(lut.shape[0]),
                        C.c_float(sample_dist), C.c_float(rmin), C.c_float(rmax), dptr, int(structured_conventions),
                        C.byref(rays))
    return rays, (TraceResult(rays) if keep else None)


def render_umesh_partials(mesh, cam, W, H, lut, sample_dist, rmin, rmax, canvas_depth=None):
    """UnstructuredWrapper::render + vtkm_to_partials (VolumeRenderer.cpp:141-221): one partial per ray that
    gathered alpha >= 0.001, depth = the ray's exit distance."""
    rays, _ = trace_umesh(mesh, cam, W, H, lut, sample_dist, rmin, rmax, canvas_depth, keep=False)
    out = np.zeros(rays.n, PARTIAL_DTYPE)
    n = lib.orc_extract_partials(C.byref(rays), out.ctypes.data_as(C.c_void_p))
    rays_free(rays)
    return out[:n].copy()


def render_umesh_to_canvas(mesh, cam, W, H, lut, sample_dist, rmin, rmax, rgba, depth, use_depth=True):
    rays, _ = trace_umesh(mesh, cam, W, H, lut, sample_dist, rmin, rmax, depth if use_depth else None, keep=False)
    lib.orc_write_to_canvas(C.byref(rays), C.byref(cam), W, H, _ptr(rgba, C.c_float), _ptr(depth, C.c_float))
    rays_free(rays)


def new_canvas(W, H):
    """Canvas::Clear: colour 0, depth 1.001."""
    return np.zeros((H * W, 4), np.float32), np.full(H * W, 1.001, np.float32)


def render_to_canvas(block, cam, W, H, lut, sample_dist, rmin, rmax, rgba, depth, use_depth=True):
    """MapperVolume::RenderCells (path A per-domain body): trace + WriteToCanvas, in place."""
    rays, _ = trace_block(block, cam, W, H, lut, sample_dist, rmin, rmax,
                          depth if use_depth else None, keep=False)
    ns = int(rays.n_samples)
    lib.orc_write_to_canvas(C.byref(rays), C.byref(cam), W, H, _ptr(rgba, C.c_float),
                            _ptr(depth, C.c_float))
    rays_free(rays)
    return ns


def render_partials(block, cam, W, H, lut, sample_dist, rmin, rmax, canvas_depth):
    """StructuredWrapper::render (path B per-domain body): trace + partial extraction."""
    rays, _ = trace_block(block, cam, W, H, lut, sample_dist, rmin, rmax, canvas_depth, keep=False)
    out = np.zeros(rays.n, PARTIAL_DTYPE)
    n = lib.orc_extract_partials(C.byref(rays), out.ctypes.data_as(C.c_void_p))
    rays_free(rays)
    return out[:n].copy()


def partials_to_canvas(partials, cam, W, H, rgba, depth):
    p = np.ascontiguousarray(partials)
    lib.orc_partials_to_canvas(p.ctypes.data_as(C.c_void_p), C.c_int64(p.size), C.byref(cam), W, H,
                               _ptr(rgba, C.c_float), _ptr(depth, C.c_float))


def visibility_order(domain_bounds, cam):
    db = np.ascontiguousarray(domain_bounds, np.float64).reshape(-1, 6)
    n = db.shape[0]
    order = np.zeros(n, np.int32)
    depths = np.zeros(n, np.float32)
    lib.orc_visibility_order(_ptr(db, C.c_double), n, C.byref(cam), _ptr(order, C.c_int),
                             _ptr(depths, C.c_float))
    return order, depths


# ----------------------------------------------------------------------------- compositing
def image_init(rgba, depth, depth_mode=0):
    rgba = np.ascontiguousarray(rgba, np.float32).reshape(-1, 4)
    depth = np.ascontiguousarray(depth, np.float32).reshape(-1)
    n = depth.size
    out = np.zeros((n, 4), np.uint8)
    od = np.zeros(n, np.float32)
    lib.orc_image_init(_ptr(rgba, C.c_float), _ptr(depth, C.c_float), n, depth_mode,
                       _ptr(out, C.c_uint8), _ptr(od, C.c_float))
    return out, od


def ordered_composite(layers_rgba, layers_depth, vis_order):
    lr = np.ascontiguousarray(layers_rgba, np.uint8)
    ld = np.ascontiguousarray(layers_depth, np.float32)
    n_img = lr.shape[0]
    n = ld.reshape(n_img, -1).shape[1]
    vo = np.ascontiguousarray(vis_order, np.int32)
    out = np.zeros((n, 4), np.uint8)
    od = np.zeros(n, np.float32)
    lib.orc_ordered_composite(_ptr(lr, C.c_uint8), _ptr(ld, C.c_float), _ptr(vo, C.c_int), n_img, n,
                              _ptr(out, C.c_uint8), _ptr(od, C.c_float))
    return out, od


def zbuffer_composite(front_rgba, front_depth, img_rgba, img_depth, gl_depth=False):
    n = front_depth.size
    lib.orc_zbuffer_composite(_ptr(front_rgba, C.c_uint8), _ptr(front_depth, C.c_float),
                              _ptr(img_rgba, C.c_uint8), _ptr(img_depth, C.c_float), n,
                              int(gl_depth))


def image_to_canvas(rgba_u8, depth):
    n = depth.size
    out = np.zeros((n, 4), np.float32)
    od = np.zeros(n, np.float32)
    lib.orc_image_to_canvas(_ptr(np.ascontiguousarray(rgba_u8), C.c_uint8),
                            _ptr(np.ascontiguousarray(depth), C.c_float), n, _ptr(out, C.c_float),
                            _ptr(od, C.c_float))
    return out, od


def blend_background(canvas_rgba, bg):
    """Canvas::BlendBackground in place (Render::RenderBackground)."""
    assert canvas_rgba.dtype == np.float32 and canvas_rgba.flags.c_contiguous
    b = np.ascontiguousarray(bg, np.float32)
    lib.orc_blend_background(_ptr(canvas_rgba, C.c_float), canvas_rgba.size // 4, _ptr(b, C.c_float))


def encode_rgba8(canvas_rgba, W, H, flip=True):
    """PNGEncoder::Encode's float -> uint8 conversion (rows flipped like the PNG on disk)."""
    c = np.ascontiguousarray(canvas_rgba, np.float32)
    out = np.zeros((H, W, 4), np.uint8)
    lib.orc_encode_rgba8(_ptr(c, C.c_float), W, H, int(flip), _ptr(out, C.c_uint8))
    return out


def composite_partials(partial_lists):
    """PartialCompositor::composite for a list of per-domain partial arrays (single rank)."""
    allp = (np.concatenate(partial_lists) if len(partial_lists) else np.zeros(0, PARTIAL_DTYPE))
    allp = np.ascontiguousarray(allp)
    out = np.zeros(allp.size, PARTIAL_DTYPE)
    n = lib.orc_composite_partials(allp.ctypes.data_as(C.c_void_p), C.c_int64(allp.size),
                                   out.ctypes.data_as(C.c_void_p))
    return out[:n].copy()


def partial_owner(pixel_id, min_pixel, max_pixel, n_ranks):
    return int(lib.orc_partial_owner(int(pixel_id), int(min_pixel), int(max_pixel), int(n_ranks)))


# ----------------------------------------------------------------------------- _ref (real apcomp)
def ref_composite_vis_order(rgba_f32, depth, vis_order, W, H):
    assert ref is not None
    r = np.ascontiguousarray(rgba_f32, np.float32)
    d = np.ascontiguousarray(depth, np.float32)
    vo = np.ascontiguousarray(vis_order, np.int32)
    out = np.zeros((H * W, 4), np.uint8)
    od = np.zeros(H * W, np.float32)
    ref.ref_composite_vis_order(_ptr(r, C.c_float), _ptr(d, C.c_float), _ptr(vo, C.c_int),
                                int(vo.size), W, H, _ptr(out, C.c_uint8), _ptr(od, C.c_float))
    return out, od


def ref_composite_zbuffer(rgba_f32, depth, n_images, W, H):
    assert ref is not None
    r = np.ascontiguousarray(rgba_f32, np.float32)
    d = np.ascontiguousarray(depth, np.float32)
    out = np.zeros((H * W, 4), np.uint8)
    od = np.zeros(H * W, np.float32)
    ref.ref_composite_zbuffer(_ptr(r, C.c_float), _ptr(d, C.c_float), n_images, W, H,
                              _ptr(out, C.c_uint8), _ptr(od, C.c_float))
    return out, od


def ref_composite_partials(partial_lists):
    assert ref is not None
    counts = np.array([p.size for p in partial_lists], np.int64)
    allp = np.ascontiguousarray(np.concatenate(partial_lists))
    out = np.zeros(allp.size, PARTIAL_DTYPE)
    n = ref.ref_composite_partials(allp.ctypes.data_as(C.c_void_p), _ptr(counts, C.c_longlong),
                                   int(counts.size), out.ctypes.data_as(C.c_void_p))
    return out[:n].copy()


# ----------------------------------------------------------------------------- radix-k surface compositing (C4)
# Restatement of RadixKCompositor::CompositeImpl (src/libs/vtkh/compositing/RadixKCompositor.cpp:35-180; the
# apcomp copy is textually the same) and of the vendored DIY pieces it drives
# (src/libs/vtkh/compositing/internal/diy/include/diy: decomposition.hpp fill_divisions/factor,
# partners/common.hpp factor/fill/fill_steps with contiguous = false, reduce.hpp, and CollectImages in
# vtkh_diy_collect.hpp).  PINNED against the reference's own code run in one process
# (oracle/_ref/libradixk_ref.so, tests/test_oracle_radixk.py).
RADIXK_MAGIC_K = 8  # RadixKCompositor.cpp:144


def _diy_factor_ascending(n):
    """RegularDecomposer::factor (decomposition.hpp:596-611): prime factors, smallest first."""
    f = []
    while n != 1:
        for i in range(2, n + 1):
            if n % i == 0:
                f.append(i)
                n //= i
                break
    return f


def radixk_divisions(n, W, H):
    """RegularDecomposer<DiscreteBounds>(2, [1..W]x[1..H], n).fill_divisions (decomposition.hpp:514-594):
    the largest remaining prime factor always splits the dimension whose blocks are currently largest
    (ties: fewer blocks so far, then the lower dimension)."""
    dmin, dmax = [1, 1], [W, H]
    divs = [{"dim": d, "nb": 1, "b_size": dmax[d] - dmin[d]} for d in range(2)]
    for f in reversed(_diy_factor_ascending(n)):
        divs.sort(key=lambda v: (-v["b_size"], v["nb"], v["dim"]))
        v = divs[0]
        nn = v["nb"] * f
        lo = dmin[v["dim"]]
        hi = dmax[v["dim"]] if nn == 1 else lo + (dmax[v["dim"]] - lo + 1) // nn - 1
        if hi < lo:
            raise RuntimeError("Unable to decompose domain into %d blocks" % n)
        v["nb"], v["b_size"] = nn, hi - lo
    out = [0, 0]
    for v in divs:
        out[v["dim"]] = v["nb"]
    return out


def _diy_factor_k(k, tot):
    """RegularPartners::factor(k, tot_b, kv) (partners/common.hpp:170-201)."""
    kv, rem = [], tot
    while rem > 1:
        if rem % k == 0:
            kv.append(k)
            rem //= k
        else:
            for j in range(k - 1, 1, -1):
                if rem % j == 0:
                    kv.append(j)
                    rem //= j
                    break
            else:
                kv.append(rem)
                rem = 1
    return kv


def radixk_rounds(divisions, k=RADIXK_MAGIC_K):
    """[(dim, group size, step)] per round: per-dimension factorisations interleaved dimension by dimension
    (partners/common.hpp:139-166), steps for contiguous = false, i.e. distance halving (:75-83)."""
    per_dim = [_diy_factor_k(k, d) for d in divisions]
    kvs, at = [], [0] * len(divisions)
    while True:
        changed = False
        for d in range(len(divisions)):
            if at[d] < len(per_dim[d]):
                kvs.append((d, per_dim[d][at[d]]))
                at[d] += 1
                changed = True
        if not changed:
            break
    cur = list(divisions)
    rounds = []
    for d, size in kvs:
        cur[d] //= size
        rounds.append((d, size, cur[d]))
    return rounds


def _gid_to_coords(gid, divisions):
    c = []
    for d in divisions:
        c.append(gid % d)
        gid //= d
    return c


def _coords_to_gid(c, divisions):
    gid = 0
    for i in range(len(c) - 1, -1, -1):
        gid = gid * divisions[i] + c[i]
    return gid


def radixk_group(gid, rnd, divisions):
    """RegularPartners::fill (partners/common.hpp:86-113): the gids of `gid`'s group in a round, by
    ascending position; this is the out_link order of that round and the in_link order of the next."""
    d, size, step = rnd
    c = _gid_to_coords(gid, divisions)
    pos = c[d] // step % size
    first = c[d] - pos * step
    out = []
    for j in range(size):
        cc = list(c)
        cc[d] = first + j * step
        out.append(_coords_to_gid(cc, divisions))
    return out, pos


def _radixk_split(lo, hi, size):
    """reduce_images' balanced ranges (RadixKCompositor.cpp:66-92) of the INCLUSIVE pixel interval [lo, hi]:
    range_length = hi - lo (not + 1), so neighbouring pieces share their boundary pixel."""
    length = hi - lo
    base, rem = length // size, length % size
    out, m = [], lo
    for i in range(size):
        b = base + (1 if i < rem else 0)
        out.append((m, m + b))
        m += b
    return out


def radixk_simulate(rgba_u8, depth, W, H):
    """The literal algorithm: every block's image is split, swapped and z-composited round by round
    (reduce_images), then block 0 pastes every block's final piece into the full frame in gid order
    (CollectImages: its own piece first, then gids 1..n-1; later pastes overwrite the shared boundary pixels).
    rgba_u8 [n, H*W, 4] uint8, depth [n, H*W] f32 (row 0 = bounds y = 1).  Returns (rgba [H*W,4], depth [H*W])."""
    n = rgba_u8.shape[0]
    divisions = radixk_divisions(n, W, H)
    rounds = radixk_rounds(divisions)
    # a block's image: bounds (x0, y0, x1, y1) inclusive, 1-based + arrays over that rectangle
    imgs = [{"b": (1, 1, W, H), "c": rgba_u8[g].reshape(H, W, 4).copy(), "d": depth[g].reshape(H, W).copy()}
            for g in range(n)]

    def subset(im, b):
        x0, y0, x1, y1 = b
        ox, oy = im["b"][0], im["b"][1]
        return {"b": b, "c": im["c"][y0 - oy:y1 - oy + 1, x0 - ox:x1 - ox + 1].copy(),
                "d": im["d"][y0 - oy:y1 - oy + 1, x0 - ox:x1 - ox + 1].copy()}

    def zcomp(front, inc):  # ImageCompositor::ZBufferComposite (ImageCompositor.hpp:49-76)
        assert front["b"] == inc["b"]
        take = ~((inc["d"] > 1.0) | (front["d"] < inc["d"]))
        front["d"][take] = np.abs(inc["d"][take])
        front["c"][take] = inc["c"][take]

    inbox = [dict() for _ in range(n)]
    for r, rnd in enumerate(rounds + [None]):
        if r > 0:
            for g in range(n):
                group, _ = radixk_group(g, rounds[r - 1], divisions)
                for q in group:  # in_link order
                    if q != g:
                        zcomp(imgs[g], inbox[g][q])
            inbox = [dict() for _ in range(n)]
        if rnd is None:
            break
        d = rnd[0]
        nxt = [None] * n
        for g in range(n):
            group, _ = radixk_group(g, rnd, divisions)
            x0, y0, x1, y1 = imgs[g]["b"]
            pieces = _radixk_split(x0 if d == 0 else y0, x1 if d == 0 else y1, len(group))
            for i, q in enumerate(group):
                b = (pieces[i][0], y0, pieces[i][1], y1) if d == 0 else (x0, pieces[i][0], x1, pieces[i][1])
                piece = subset(imgs[g], b)
                if q == g:
                    nxt[g] = piece
                else:
                    inbox[q][g] = piece
        imgs = nxt
    out_c = np.zeros((H, W, 4), np.uint8)
    out_d = np.zeros((H, W), np.float32)
    for g in range(n):
        x0, y0, x1, y1 = imgs[g]["b"]
        out_c[y0 - 1:y1, x0 - 1:x1] = imgs[g]["c"]
        out_d[y0 - 1:y1, x0 - 1:x1] = imgs[g]["d"]
    return out_c.reshape(-1, 4), out_d.reshape(-1)


def radixk_schedule(n, W, H):
    """Closed form of the above, the shape the GPU kernel consumes.  Returns
      divisions [dx, dy];
      lo[d]  : for each coordinate c along dimension d, the first 0-based pixel of block column/row c's
               EFFECTIVE span (a boundary pixel shared by two pieces ends up with the higher gid's value,
               because CollectImages pastes in gid order);
      seq[g] : the order in which the ranks' fragments are z-composited for the pixels gid g ends up owning:
               Seq_0(b) = [b]; Seq_{r+1}(g) = Seq_r(g) ++ Seq_r(q) for q in group_r(g) by position, q != g.
    The composited pixel is the LAST fragment of seq with the minimum depth among those with depth <= 1
    (ImageCompositor.hpp:65: an incoming fragment replaces on <=), or seq[0]'s pixel if none is <= 1."""
    divisions = radixk_divisions(n, W, H)
    rounds = radixk_rounds(divisions)
    seq = [[g] for g in range(n)]
    for rnd in rounds:
        nxt = []
        for g in range(n):
            group, _ = radixk_group(g, rnd, divisions)
            s = list(seq[g])
            for q in group:
                if q != g:
                    s += seq[q]
            nxt.append(s)
        seq = nxt
    lo = []
    for d in range(2):
        spans = [(1, W if d == 0 else H)]
        for rd, size, _ in rounds:
            if rd == d:
                spans = [p for s in spans for p in _radixk_split(s[0], s[1], size)]
        # piece 0 starts at pixel 0; every other piece effectively starts AT the boundary pixel it shares with
        # its lower neighbour (1-based s[0] -> 0-based s[0] - 1): the higher gid is pasted later
        lo.append([0] + [s[0] - 1 for s in spans[1:]])
    return {"divisions": divisions, "rounds": rounds, "lo": lo, "seq": seq}


def radixk_zbuffer(rgba_u8, depth, W, H):
    """Z-composite n full-frame images through radixk_schedule (per pixel; the form the kernel implements)."""
    n = rgba_u8.shape[0]
    s = radixk_schedule(n, W, H)
    cx = np.searchsorted(np.asarray(s["lo"][0]), np.arange(W), side="right") - 1
    cy = np.searchsorted(np.asarray(s["lo"][1]), np.arange(H), side="right") - 1
    gid = (cx[None, :] + s["divisions"][0] * cy[:, None]).reshape(-1)
    out_c = np.zeros((H * W, 4), np.uint8)
    out_d = np.zeros(H * W, np.float32)
    for g in range(n):
        px = np.nonzero(gid == g)[0]
        if px.size == 0:
            continue
        order = s["seq"][g]
        fc = rgba_u8[order[0]][px].copy()
        fd = depth[order[0]][px].copy()
        for r in order[1:]:
            d = depth[r][px]
            take = ~((d > 1.0) | (fd < d))
            fd[take] = np.abs(d[take])
            fc[take] = rgba_u8[r][px][take]
        out_c[px], out_d[px] = fc, fd
    return out_c, out_d


_radixk_ref_path = os.path.join(_HERE, "_ref", "libradixk_ref.so")
radixk_ref = C.CDLL(_radixk_ref_path) if os.path.exists(_radixk_ref_path) else None


def ref_radixk_zbuffer(rgba_u8, depth, W, H):
    """The reference's own reduce_images + DIY, n blocks in one process (oracle/radixk_ref_harness.cpp)."""
    assert radixk_ref is not None
    c = np.ascontiguousarray(rgba_u8, np.uint8)
    d = np.ascontiguousarray(depth, np.float32)
    n = c.shape[0]
    out = np.zeros((H * W, 4), np.uint8)
    od = np.zeros(H * W, np.float32)
    info = (C.c_int * 64)()
    rc = radixk_ref.ref_radixk_zbuffer(c.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), n, W, H,
                                       out.ctypes.data_as(C.c_void_p), od.ctypes.data_as(C.c_void_p), info)
    if rc != 0:
        raise RuntimeError("reference radix-k failed (rc %d)" % rc)
    info = list(info)
    k = info.index(-1)
    return out, od, {"divisions": info[:2], "rounds": [(info[i], info[i + 1]) for i in range(2, k, 2)]}


# ----------------------------------------------------------------------------- colour table (K8)
class _CTable(C.Structure):
    _fields_ = [("space", C.c_int), ("n_color", C.c_int), ("color_x", C.c_double * 64),
                ("color_rgb", (C.c_float * 3) * 64), ("n_alpha", C.c_int), ("alpha_x", C.c_double * 64),
                ("alpha_v", C.c_float * 64)]


lib.orc_colortable_add_point.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]
lib.orc_colortable_add_point_alpha.argtypes = [C.c_void_p, C.c_double, C.c_double]
lib.orc_colortable_correct_opacity.argtypes = [C.c_void_p, C.c_float]

PRESETS = {"cool to warm": 0, "rainbow desaturated": 1, "black-body radiation": 2, "grayscale": 3}


class ColorTable:
    """vtkm::cont::ColorTable as the volume plot uses it (oracle/colortable_oracle.c)."""

    def __init__(self, name="cool to warm"):
        self.t = _CTable()
        if lib.orc_colortable_preset(C.byref(self.t), PRESETS[name.lower()]) != 0:
            raise ValueError(name)

    def add_point(self, x, rgb):
        lib.orc_colortable_add_point(C.byref(self.t), float(x), float(rgb[0]), float(rgb[1]), float(rgb[2]))
        return self

    def add_point_alpha(self, x, a):
        lib.orc_colortable_add_point_alpha(C.byref(self.t), float(x), float(a))
        return self

    def clear_colors(self):
        lib.orc_colortable_clear_colors(C.byref(self.t))
        return self

    def correct_opacity(self, samples):
        """VolumeRenderer::CorrectOpacity (in place)."""
        lib.orc_colortable_correct_opacity(C.byref(self.t), C.c_float(samples))
        return self

    def sample_u8(self, n=1024):
        out = np.zeros((n, 4), np.uint8)
        lib.orc_colortable_sample_u8(C.byref(self.t), int(n), _ptr(out, C.c_uint8))
        return out

    def lut(self, n=1024):
        out = np.zeros((n, 4), np.float32)
        lib.orc_colortable_lut(C.byref(self.t), int(n), _ptr(out, C.c_float))
        return out


def parse_color_table(node):
    """parse_color_table (ascent_runtime_conduit_to_vtkm_parsing.cpp:199-305) for the subset the volume
    tests use: a preset name, alpha / rgb control points appended in order."""
    name = str(node.get("name", "cool to warm")).lower()
    t = ColorTable(name if name in PRESETS else "cool to warm")
    cps = node.get("control_points", [])
    if any(p["type"] == "rgb" for p in cps) and "name" not in node:
        t.clear_colors()
    for p in cps:
        if p["type"] == "rgb":
            t.add_point(p["position"], p["color"])
        elif p["type"] == "alpha":
            t.add_point_alpha(p["position"], p["alpha"])
    return t


def default_volume_table():
    """VolumeRenderer ctor, VolumeRenderer.cpp:395-408 (both alpha points at x = 0)."""
    return ColorTable("cool to warm").add_point_alpha(0.0, 0.02).add_point_alpha(0.0, 0.5)


# ----------------------------------------------------------------------------- braid (input generator)
def braid_values(nx, ny, nz, i0=0, j0=0, k0=0, gx=None, gy=None, gz=None, dtype=np.float64):
    """Conduit blueprint::mesh::examples::braid vertex field [Conduit, recalled; corroborated by
    src/examples/tutorial/ascent_intro/cpp/blueprint_example3.cpp:61-86], evaluated point by point:
    window (i0,j0,k0)+(nx,ny,nz) of a (gx,gy,gz) grid, x fastest."""
    gx, gy, gz = gx or nx, gy or ny, gz or nz
    pi = 3.14159265359
    dx = float(np.float32(4.0 * pi)) / float(gx - 1)
    dy = float(np.float32(2.0 * pi)) / float(gy - 1)
    dz = float(np.float32(3.0 * pi)) / float(gz - 1)
    out = np.empty((nz, ny, nx), dtype)
    i = np.arange(i0, i0 + nx, dtype=np.float64)[None, :]
    j = np.arange(j0, j0 + ny, dtype=np.float64)[:, None]
    cx = i * dx + 2.0 * pi
    cy = j * dy - pi
    for k in range(nz):
        cz = (k0 + k) * dz - 1.5 * pi
        v = np.sin(cx) + np.sin(cy) + 2 * np.cos(np.sqrt((cx * cx) / 2.0 + cy * cy) / .75) + 4 * np.cos(cx * cy / 4.0)
        v = v + np.sin(cz) + 1.5 * np.cos(np.sqrt(cx * cx + cy * cy + cz * cz) / .75)
        out[k] = v
    return out.reshape(-1)
