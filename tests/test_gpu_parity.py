"""Parity of the CUDA path (through the C ABI, include/vr_b200.h) against the CPU oracle.

Tolerance (BASELINE.json north_star): >= 99.9 % of pixels within 1/255 per RGBA channel, every
pixel within 3/255, PSNR >= 50 dB; integer/index work (uint8 fold, visibility order, partial sort
order, pixel ids) bit-exact.  Because sampler.cu is compiled without FMA contraction the float
canvas is in fact expected to be bit-identical; that stronger property is asserted too."""
import numpy as np
import pytest

import scenes
from ascent_b200 import _lib, color_table, datasets
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = _lib.Context(0)
    yield c
    c.close()


def check_tolerance(mine, ref, exact=True):
    """mine/ref float RGBA in [0,1], shape (N,4)."""
    d = np.abs(mine.astype(np.float64) - ref.astype(np.float64)).max(axis=1)
    assert (d <= 1.0 / 255.0).mean() >= 0.999
    assert d.max() <= 3.0 / 255.0
    mse = ((mine.astype(np.float64) - ref.astype(np.float64)) ** 2).mean()
    psnr = 99.0 if mse == 0 else 10.0 * np.log10(1.0 / mse)
    assert psnr >= 50.0
    if exact:
        assert (mine.view(np.uint32) == ref.view(np.uint32)).all(axis=1).mean() >= 0.9999


def gpu_path_a_canvas(ctx, dom, sc, use_depth=False, canvas=None):
    W, H = sc["W"], sc["H"]
    ctx.block_from_domain(0, dom)
    ctx.set_tf(sc["lut"])
    if canvas is None:
        ctx.canvas_clear(W, H)
    else:
        ctx.canvas_upload(W, H, canvas[0], canvas[1])
    ctx.trace_to_canvas(0, sc["cam"], sc["sample_dist"], sc["rmin"], sc["rmax"], use_depth)
    return ctx.canvas_download(W, H)


def oracle_canvas(dom, sc, canvas=None, use_depth=True):
    W, H = sc["W"], sc["H"]
    if canvas is None:
        rgba, depth = O.new_canvas(W, H)
    else:
        rgba, depth = canvas[0].copy(), canvas[1].copy()
    ns = O.render_to_canvas(scenes.oracle_block(dom), sc["cam"], W, H, sc["lut"], sc["sample_dist"],
                            sc["rmin"], sc["rmax"], rgba, depth, use_depth=use_depth)
    return rgba, depth, ns


def depth_equal(a, b, where):
    a, b = a[where], b[where]
    return np.array_equal(a, b, equal_nan=True)


# ------------------------------------------------------------------ config c1 and friends
@pytest.mark.parametrize("res", [(256, 256), (1024, 1024), (640, 360)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_c1_braid_uniform(ctx, res, dtype):
    """BASELINE config 1: braid uniform 32^3, default camera, S=100 (1024^2 is the named size)."""
    dom = datasets.braid_uniform(32, dtype=dtype)
    b = datasets.domain_bounds(dom)
    sc = dict(W=res[0], H=res[1], cam=O.camera_reset_to_bounds(b),
              lut=color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut(),
              sample_dist=O.sample_distance(b, 100))
    sc["rmin"], sc["rmax"] = scenes.field_range([dom])
    rgba, depth = gpu_path_a_canvas(ctx, dom, sc)
    o_rgba, o_depth, _ = oracle_canvas(dom, sc)
    check_tolerance(rgba, o_rgba)
    assert depth_equal(depth, o_depth, o_rgba[:, 3] > 0)
    # pixels outside the block's screen subset are untouched
    assert np.array_equal(depth == np.float32(1.001), o_depth == np.float32(1.001))


def test_default_transfer_function(ctx):
    """the reference's default TF (VolumeRenderer.cpp:395-408) on c1."""
    dom = datasets.braid_uniform(32)
    b = datasets.domain_bounds(dom)
    sc = dict(W=300, H=300, cam=O.camera_reset_to_bounds(b),
              lut=color_table.default_volume_table().corrected_opacity(100).lut(),
              sample_dist=O.sample_distance(b, 100))
    sc["rmin"], sc["rmax"] = scenes.field_range([dom])
    rgba, _ = gpu_path_a_canvas(ctx, dom, sc)
    check_tolerance(rgba, oracle_canvas(dom, sc)[0])


@pytest.mark.parametrize("which", [0, 1])
def test_golden_scenes_multi_render(ctx, which):
    sc = scenes.multi_render_scene(which)
    rgba, depth = gpu_path_a_canvas(ctx, sc["doms"][0], sc)
    o_rgba, o_depth, _ = oracle_canvas(sc["doms"][0], sc)
    check_tolerance(rgba, o_rgba)


@pytest.mark.parametrize("samples", [1, 7, 100, 887])
def test_sample_counts(ctx, samples):
    """SetNumberOfSamples extremes: 1 sample per diagonal up to ~voxel-sized steps."""
    dom = datasets.braid_uniform(24, dtype=np.float32)
    b = datasets.domain_bounds(dom)
    cam = O.camera_reset_to_bounds(b)
    O.camera_azimuth(cam, 30.0)
    O.camera_elevation(cam, 20.0)
    sc = dict(W=200, H=160, cam=cam,
              lut=color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(samples).lut(),
              sample_dist=O.sample_distance(b, samples))
    sc["rmin"], sc["rmax"] = scenes.field_range([dom])
    rgba, _ = gpu_path_a_canvas(ctx, dom, sc)
    check_tolerance(rgba, oracle_canvas(dom, sc)[0])


def test_early_termination(ctx):
    """opaque TF: rays stop at alpha >= 1 after the same number of samples."""
    dom = datasets.braid_uniform(20, dtype=np.float32)
    b = datasets.domain_bounds(dom)
    tf = color_table.ColorTable("cool to warm")
    tf.add_point_alpha(0.0, 1.0)
    tf.add_point_alpha(1.0, 1.0)
    sc = dict(W=128, H=128, cam=O.camera_reset_to_bounds(b), lut=tf.lut(),
              sample_dist=O.sample_distance(b, 100))
    sc["rmin"], sc["rmax"] = scenes.field_range([dom])
    rgba, _ = gpu_path_a_canvas(ctx, dom, sc)
    o = oracle_canvas(dom, sc)[0]
    check_tolerance(rgba, o)
    assert o[:, 3].max() == 1.0


def test_camera_inside_volume_and_zoom(ctx):
    dom = datasets.braid_uniform(20, dtype=np.float32)
    b = datasets.domain_bounds(dom)
    for zoom, pos in [(1.0, [1., 2., 3.]), (2.5, None), (0.5, None)]:
        cam = O.camera_reset_to_bounds(b)
        cam.zoom = zoom
        if pos is not None:
            cam.position[:] = pos
            cam.look_at[:] = [0., 0., -1.]
        sc = dict(W=160, H=120, cam=cam,
                  lut=color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut(),
                  sample_dist=O.sample_distance(b, 100))
        sc["rmin"], sc["rmax"] = scenes.field_range([dom])
        rgba, _ = gpu_path_a_canvas(ctx, dom, sc)
        check_tolerance(rgba, oracle_canvas(dom, sc)[0])


def test_volume_behind_camera_is_a_no_op(ctx):
    dom = datasets.braid_uniform(8, dtype=np.float32)
    b = datasets.domain_bounds(dom)
    cam = O.make_camera([0, 0, 40], [0, 0, 80], [0, 1, 0], near=0.1, far=500.)
    sc = dict(W=64, H=64, cam=cam, lut=color_table.default_volume_table().lut(), sample_dist=0.3,
              rmin=-10., rmax=10.)
    rgba, depth = gpu_path_a_canvas(ctx, dom, sc)
    o_rgba, o_depth, _ = oracle_canvas(dom, sc)
    assert np.array_equal(rgba, o_rgba) and rgba.max() == 0.0


# ------------------------------------------------------------------ rectilinear / cell fields
@pytest.mark.parametrize("power", [1.0, 1.5, 0.6])
def test_rectilinear(ctx, power):
    """config c4 in small: rectilinear braid with warped axes, cinema cameras."""
    dom = datasets.braid_rectilinear(40, 36, 44, power=power, dtype=np.float32)
    b = datasets.domain_bounds(dom)
    lut = color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut()
    rmin, rmax = scenes.field_range([dom])
    for phi, theta in [(-180., 0.), (-45., 22.5), (90., 67.5), (135., 157.5)]:
        cam = scenes.cinema_camera(b, phi, theta)
        sc = dict(W=192, H=192, cam=cam, lut=lut, sample_dist=O.sample_distance(b, 100), rmin=rmin,
                  rmax=rmax)
        rgba, _ = gpu_path_a_canvas(ctx, dom, sc)
        check_tolerance(rgba, oracle_canvas(dom, sc)[0])


def test_mpi_golden_scene_rectilinear_f64(ctx):
    sc = scenes.mpi_volume_scene()
    for dom in sc["doms"]:
        rgba, _ = gpu_path_a_canvas(ctx, dom, sc)
        check_tolerance(rgba, oracle_canvas(dom, sc)[0])


@pytest.mark.parametrize("kind", ["uniform", "rectilinear"])
def test_cell_centred_field(ctx, kind):
    if kind == "uniform":
        dom = datasets.braid_uniform(21, dtype=np.float32)
    else:
        dom = datasets.braid_rectilinear(21, power=1.3, dtype=np.float32)
    n = 20
    pts = dom["field"].reshape(21, 21, 21)
    dom = dict(dom, field=np.ascontiguousarray(pts[:n, :n, :n]).reshape(-1), assoc="cell")
    b = datasets.domain_bounds(dom)
    cam = O.camera_reset_to_bounds(b)
    O.camera_azimuth(cam, -25.0)
    sc = dict(W=150, H=150, cam=cam,
              lut=color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut(),
              sample_dist=O.sample_distance(b, 100))
    sc["rmin"], sc["rmax"] = scenes.field_range([dom])
    rgba, _ = gpu_path_a_canvas(ctx, dom, sc)
    check_tolerance(rgba, oracle_canvas(dom, sc)[0])


# ------------------------------------------------------------------ K2 / K7 with a live canvas
def test_existing_canvas_depth_and_colour(ctx):
    """opaque geometry already on the canvas: rays stop at its depth, colour blends over it."""
    dom = datasets.braid_uniform(24, dtype=np.float32)
    b = datasets.domain_bounds(dom)
    W, H = 160, 128
    cam = O.camera_reset_to_bounds(b)
    rng = np.random.default_rng(3)
    rgba0, depth0 = O.new_canvas(W, H)
    yy, xx = np.mgrid[0:H, 0:W]
    disk = ((xx - 80) ** 2 + (yy - 64) ** 2) < 40 ** 2
    # image-space depths between near and far that land inside the volume
    pv = O.projview(cam, W, H)
    zs = []
    for z in (-5.0, 5.0):
        q = pv @ np.array([0, 0, z, 1], np.float32)
        zs.append(0.5 * q[2] / q[3] + 0.5)
    dvals = rng.uniform(min(zs), max(zs), disk.sum()).astype(np.float32)
    depth0.reshape(H, W)[disk] = dvals
    rgba0.reshape(H, W, 4)[disk] = [0.2, 0.4, 0.6, 1.0]
    sc = dict(W=W, H=H, cam=cam,
              lut=color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut(),
              sample_dist=O.sample_distance(b, 100))
    sc["rmin"], sc["rmax"] = scenes.field_range([dom])
    rgba, depth = gpu_path_a_canvas(ctx, dom, sc, use_depth=True, canvas=(rgba0, depth0))
    o_rgba, o_depth, _ = oracle_canvas(dom, sc, canvas=(rgba0, depth0))
    check_tolerance(rgba, o_rgba)
    # and through the host-buffer entry point vtk-h would call
    r2, d2 = rgba0.copy(), depth0.copy()
    ctx.render_image(0, cam, W, H, sc["sample_dist"], sc["rmin"], sc["rmax"], r2, d2)
    assert np.array_equal(r2, rgba)


# ------------------------------------------------------------------ path B (partials)
def partial_sets_equal(a, b):
    a = np.sort(a, order=["pixel_id", "depth"])
    b = np.sort(b, order=["pixel_id", "depth"])
    assert a.size == b.size
    assert np.array_equal(a["pixel_id"], b["pixel_id"])
    assert np.array_equal(a["depth"], b["depth"])
    return a, b


def test_render_partials_single_block(ctx):
    dom = datasets.braid_uniform(24, dtype=np.float32)
    b = datasets.domain_bounds(dom)
    W, H = 200, 150
    cam = O.camera_reset_to_bounds(b)
    O.camera_azimuth(cam, 15.0)
    lut = color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut()
    rmin, rmax = scenes.field_range([dom])
    sd = O.sample_distance(b, 100)
    ctx.block_from_domain(0, dom)
    ctx.set_tf(lut)
    mine = ctx.render_partials(0, cam, W, H, sd, rmin, rmax, None)
    _, depth = O.new_canvas(W, H)
    ref = O.render_partials(scenes.oracle_block(dom), cam, W, H, lut, sd, rmin, rmax, depth)
    a, r = partial_sets_equal(mine, ref)
    assert np.array_equal(a["rgb"], r["rgb"]) and np.array_equal(a["alpha"], r["alpha"])
    assert ref["alpha"].min() >= 0.001


@pytest.mark.parametrize("n_block,per_axis,az", [(12, 2, 0.0), (9, 3, 37.0)])
def test_multi_domain_path_b(ctx, n_block, per_axis, az):
    """configs c3(N=1)/c5 in small: several uniform blocks on one GPU -> partial compositing."""
    doms = datasets.braid_uniform_blocks(n_block, per_axis, dtype=np.float32)
    gb = datasets.union_bounds([datasets.domain_bounds(d) for d in doms])
    W, H = 240, 200
    cam = O.camera_reset_to_bounds(gb)
    O.camera_azimuth(cam, az)
    O.camera_elevation(cam, az / 3.0)
    sc = dict(doms=doms, cam=cam, W=W, H=H,
              lut=color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut(),
              sample_dist=O.sample_distance(gb, 100))
    sc["rmin"], sc["rmax"] = scenes.field_range(doms)
    ref_partials, ref_rgba, ref_depth = scenes.oracle_path_b(sc)

    ctx.set_tf(sc["lut"])
    for i, d in enumerate(doms):
        ctx.block_from_domain(i, d)
    ctx.canvas_clear(W, H)
    ctx.partials_begin(W, H)
    for i in range(len(doms)):
        ctx.trace_to_partials(i, cam, sc["sample_dist"], sc["rmin"], sc["rmax"], True)
    ctx.partials_composite()
    mine = ctx.partials_download()
    ctx.partials_to_canvas(cam)
    rgba, depth = ctx.canvas_download(W, H)
    for i in range(len(doms)):
        ctx.block_free(i)

    a = np.sort(mine, order="pixel_id")
    r = np.sort(ref_partials, order="pixel_id")
    assert np.array_equal(a["pixel_id"], r["pixel_id"])       # pixel ownership: exact
    assert np.array_equal(a["depth"], r["depth"])             # front partial's depth survives: exact
    assert np.abs(a["rgb"] - r["rgb"]).max() <= 1e-6 and np.abs(a["alpha"] - r["alpha"]).max() <= 1e-6
    check_tolerance(rgba, ref_rgba, exact=False)
    cov = ref_rgba[:, 3] > 0
    assert np.allclose(depth[cov], ref_depth[cov], rtol=0, atol=1e-6)


@pytest.mark.parametrize("clear", [True, False])
def test_fused_partial_composite_to_canvas(ctx, clear):
    """vr_partials_composite_to_canvas == vr_partials_composite + vr_partials_to_canvas, bit for bit,
    over a cleared canvas and over a canvas that already holds opaque geometry."""
    doms = datasets.braid_uniform_blocks(10, 2, dtype=np.float32)
    gb = datasets.union_bounds([datasets.domain_bounds(d) for d in doms])
    W, H = 260, 180
    cam = O.camera_reset_to_bounds(gb)
    O.camera_azimuth(cam, 21.0)
    lut = color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut()
    sd = O.sample_distance(gb, 100)
    rmin, rmax = scenes.field_range(doms)
    ctx.set_tf(lut)
    for i, d in enumerate(doms):
        ctx.block_from_domain(i, d)
    rng = np.random.default_rng(5)
    rgba0 = rng.random((H * W, 4), dtype=np.float32)
    depth0 = np.full(H * W, 1.001, np.float32)
    if clear:
        rgba0[:] = 0

    def run(fused):
        ctx.canvas_upload(W, H, rgba0, depth0)
        ctx.partials_begin(W, H)
        for i in range(len(doms)):
            ctx.trace_to_partials(i, cam, sd, rmin, rmax, False)
        if fused:
            if clear:  # poison: the fused call must overwrite every pixel
                ctx.canvas_upload(W, H, np.full((H * W, 4), 7., np.float32), np.full(H * W, 7., np.float32))
            ctx.partials_composite_to_canvas(cam, canvas_is_clear=clear)
        else:
            ctx.partials_composite()
            ctx.partials_to_canvas(cam)
        return ctx.canvas_download(W, H), np.sort(ctx.partials_download(), order="pixel_id")

    (c_a, d_a), p_a = run(False)
    (c_b, d_b), p_b = run(True)
    for i in range(len(doms)):
        ctx.block_free(i)
    assert np.array_equal(c_a, c_b) and np.array_equal(d_a, d_b)
    assert p_a.tobytes() == p_b.tobytes() and p_a.size > 1000


# ------------------------------------------------------------------ path B on dense ray layers
def layer_frame(ctx, doms, cam, W, H, sd, rmin, rmax):
    ctx.layers_begin(W, H)
    for i in range(len(doms)):
        ctx.trace_to_layer(i, cam, sd, rmin, rmax, False)


@pytest.mark.parametrize("n_block,per_axis,az,res", [(12, 2, 0.0, (240, 200)), (9, 3, 37.0, (333, 250)),
                                                     (6, 4, -120.0, (160, 90))])
def test_layers_equal_list_pipeline(ctx, n_block, per_axis, az, res):
    """vr_trace_to_layer + vr_layers_composite_to_canvas against the list pipeline and the oracle:
    same partials (as a set), same canvas, bit for bit."""
    doms = datasets.braid_uniform_blocks(n_block, per_axis, dtype=np.float32)
    gb = datasets.union_bounds([datasets.domain_bounds(d) for d in doms])
    W, H = res
    cam = O.camera_reset_to_bounds(gb)
    O.camera_azimuth(cam, az)
    O.camera_elevation(cam, az / 4.0)
    sc = dict(doms=doms, cam=cam, W=W, H=H,
              lut=color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut(),
              sample_dist=O.sample_distance(gb, 100))
    sc["rmin"], sc["rmax"] = scenes.field_range(doms)
    ctx.set_tf(sc["lut"])
    for i, d in enumerate(doms):
        ctx.block_from_domain(i, d)
    # the reference's lists
    ctx.partials_begin(W, H)
    for i in range(len(doms)):
        ctx.trace_to_partials(i, cam, sc["sample_dist"], sc["rmin"], sc["rmax"], False)
    lst = np.sort(ctx.partials_download(), order=["pixel_id", "depth"])
    # layers -> list
    layer_frame(ctx, doms, cam, W, H, sc["sample_dist"], sc["rmin"], sc["rmax"])
    ctx.layers_to_partials()
    lay = np.sort(ctx.partials_download(), order=["pixel_id", "depth"])
    assert lay.size == lst.size > 1000
    assert np.array_equal(lay["pixel_id"], lst["pixel_id"]) and np.array_equal(lay["depth"], lst["depth"])
    assert np.array_equal(np.sort(lay.view(np.uint8).reshape(-1, 24), axis=0),
                          np.sort(lst.view(np.uint8).reshape(-1, 24), axis=0))
    # layers -> canvas, cleared and over existing colours
    ref_partials, ref_rgba, ref_depth = scenes.oracle_path_b(sc)
    ctx.canvas_upload(W, H, np.full((H * W, 4), 9., np.float32), np.full(H * W, 9., np.float32))  # poison
    ctx.layers_composite_to_canvas(cam, canvas_is_clear=True)
    rgba, depth = ctx.canvas_download(W, H)
    assert np.array_equal(rgba, ref_rgba)
    cov = ref_rgba[:, 3] > 0
    assert np.array_equal(depth[cov], ref_depth[cov]) and (depth[~cov] == np.float32(1.001)).all()
    rng = np.random.default_rng(2)
    rgba0 = rng.random((H * W, 4), dtype=np.float32)
    depth0 = np.full(H * W, 1.001, np.float32)
    o_rgba, o_depth = rgba0.copy(), depth0.copy()
    O.partials_to_canvas(ref_partials, cam, W, H, o_rgba, o_depth)
    ctx.canvas_upload(W, H, rgba0, depth0)
    layer_frame(ctx, doms, cam, W, H, sc["sample_dist"], sc["rmin"], sc["rmax"])
    ctx.layers_composite_to_canvas(cam, canvas_is_clear=False)
    rgba, depth = ctx.canvas_download(W, H)
    assert np.array_equal(rgba, o_rgba) and np.array_equal(depth[cov], o_depth[cov])
    for i in range(len(doms)):
        ctx.block_free(i)


@pytest.mark.parametrize("n_block,per_axis,az,res,multi_min", [(12, 2, 0.0, (240, 200), None), (7, 4, 61.0, (320, 180), None),
                                                              (12, 2, 33.0, (240, 200), "2"), (7, 4, 61.0, (320, 180), "0")])
def test_batched_block_loop_equals_per_block_calls(ctx, n_block, per_axis, az, res, multi_min, monkeypatch):
    """vr_trace_blocks_to_layers (whole RenderMultipleDomainsPerRank loop: one launch per block overlapped on side
    streams, or -- from 16 blocks, VR_MULTI_MIN -- ONE persistent launch over the (block, tile) work items of all
    blocks) leaves the same layers and the same canvas as one vr_trace_to_layer per block."""
    if multi_min is not None:
        monkeypatch.setenv("VR_MULTI_MIN", multi_min)
    doms = datasets.braid_uniform_blocks(n_block, per_axis, dtype=np.float32)
    gb = datasets.union_bounds([datasets.domain_bounds(d) for d in doms])
    W, H = res
    cam = O.camera_reset_to_bounds(gb)
    O.camera_azimuth(cam, az)
    lut = color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut()
    sd = O.sample_distance(gb, 100)
    rmin, rmax = scenes.field_range(doms)
    ctx.set_tf(lut)
    for i, d in enumerate(doms):
        ctx.block_from_domain(i, d)
    layer_frame(ctx, doms, cam, W, H, sd, rmin, rmax)
    ctx.layers_composite_to_canvas(cam, canvas_is_clear=True)
    c_a, d_a = ctx.canvas_download(W, H)
    ctx.layers_to_partials()
    p_a = np.sort(ctx.partials_download(), order=["pixel_id", "depth"])
    for _ in range(2):  # twice: the per-launch counters are re-armed every call
        ctx.layers_begin(W, H)
        ctx.trace_blocks_to_layers(list(range(len(doms))), cam, sd, rmin, rmax, False)
        ctx.layers_composite_to_canvas(cam, canvas_is_clear=True)
        c_b, d_b = ctx.canvas_download(W, H)
        assert np.array_equal(c_a, c_b) and np.array_equal(d_a, d_b)
    ctx.layers_to_partials()
    p_b = np.sort(ctx.partials_download(), order=["pixel_id", "depth"])
    assert p_a.size == p_b.size > 1000
    assert np.array_equal(np.sort(p_a.view(np.uint8).reshape(-1, 24), axis=0),
                          np.sort(p_b.view(np.uint8).reshape(-1, 24), axis=0))
    # split over two calls in one frame == one call
    ctx.layers_begin(W, H)
    half = len(doms) // 2
    ctx.trace_blocks_to_layers(list(range(half)), cam, sd, rmin, rmax, False)
    ctx.trace_blocks_to_layers(list(range(half, len(doms))), cam, sd, rmin, rmax, False)
    ctx.layers_composite_to_canvas(cam, canvas_is_clear=True)
    c_c, d_c = ctx.canvas_download(W, H)
    assert np.array_equal(c_a, c_c) and np.array_equal(d_a, d_c)
    with pytest.raises(_lib.VRError):
        ctx.trace_blocks_to_layers([0, 9999], cam, sd, rmin, rmax, False)
    for i in range(len(doms)):
        ctx.block_free(i)


@pytest.mark.parametrize("kind", ["uniform", "rectilinear"])
def test_host_mapped_field_is_sampled_in_place(ctx, kind):
    """VR_HOST_MAPPED: a page-locked host field sampled over PCIe gives the same frame, bit for bit,
    as the copied field; the library sees later writes to the host array (no copy was made); plain
    pageable memory is refused."""
    import torch
    n = 40
    dom = datasets.braid_uniform(n, dtype=np.float32) if kind == "uniform" else \
        datasets.braid_rectilinear(n, power=1.5, dtype=np.float32)
    b = datasets.domain_bounds(dom)
    W, H = 320, 200
    cam = O.camera_reset_to_bounds(b)
    O.camera_azimuth(cam, 25.0)
    lut = color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut()
    sd = O.sample_distance(b, 100)
    rmin, rmax = scenes.field_range([dom])
    ctx.set_tf(lut)

    def publish(field, mapped):
        if kind == "uniform":
            ctx.block_uniform(0, dom["dims"], dom["origin"], dom["spacing"], field, host_mapped=mapped)
        else:
            ctx.block_rectilinear(0, dom["dims"], dom["axes"], field, host_mapped=mapped)

    def frame():
        ctx.trace_to_image(0, cam, W, H, sd, rmin, rmax, write_canvas=True)
        return ctx.canvas_download(W, H)

    publish(dom["field"], False)
    c_copy, d_copy = frame()
    pinned = torch.empty(dom["field"].size, dtype=torch.float32, pin_memory=True)
    host = pinned.numpy()
    host[:] = dom["field"].reshape(-1)
    publish(host, True)
    c_map, d_map = frame()
    assert np.array_equal(c_copy, c_map) and np.array_equal(d_copy, d_map)
    assert (c_map[:, 3] > 0).sum() > 1000
    # in place: a change of the host array shows up in the next frame without re-publishing
    ctx.synchronize()
    host[:] = np.float32(rmax)
    c_new, _ = frame()
    assert not np.array_equal(c_new, c_map)
    with pytest.raises(_lib.VRError):
        publish(dom["field"].copy(), True)
    ctx.block_free(0)


def _pinned(arr):
    import torch
    t = torch.empty(arr.size, dtype=torch.float32 if arr.dtype == np.float32 else torch.float64, pin_memory=True)
    h = t.numpy()
    h[:] = arr.reshape(-1)
    return t, h


@pytest.mark.parametrize("kind,dtype,assoc", [("uniform", np.float32, "point"), ("uniform", np.float64, "point"),
                                              ("rectilinear", np.float32, "point"), ("uniform", np.float32, "cell"),
                                              ("rectilinear", np.float64, "cell")])
def test_demand_staged_field_matches_copied_field(ctx, kind, dtype, assoc):
    """VR_HOST_STAGED: the publish copies nothing; every trace pulls the 128-byte lines its rays touch.
    Frames must be bit-identical to the copied field -- for several views per publish (lines accumulate),
    through every render entry point, and after a re-publish (stale lines must not survive: the first
    publish stages a poisoned field from other view points)."""
    n = 48
    dom = datasets.braid_uniform(n, dtype=dtype) if kind == "uniform" else \
        datasets.braid_rectilinear(n, power=1.5, dtype=dtype)
    if assoc == "cell":
        dom["assoc"] = "cell"
        dom["field"] = np.ascontiguousarray(dom["field"].reshape(n, n, n)[:n - 1, :n - 1, :n - 1]).reshape(-1)
    b = datasets.domain_bounds(dom)
    W, H = 333, 210
    lut = color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut()
    sd = O.sample_distance(b, 100)
    rmin, rmax = scenes.field_range([dom])
    ctx.set_tf(lut)
    A = _lib.VR_CELL if assoc == "cell" else _lib.VR_POINT

    def publish(field, staged):
        if kind == "uniform":
            ctx.block_uniform(0, dom["dims"], dom["origin"], dom["spacing"], field, assoc=A, staged=staged)
        else:
            ctx.block_rectilinear(0, dom["dims"], dom["axes"], field, assoc=A, staged=staged)

    def cams():
        out = []
        for az, el, zoom in [(0., 0., 0.), (35., 20., 0.), (-120., -40., 0.5), (35., 20., 0.)]:
            c = O.camera_reset_to_bounds(b)
            O.camera_azimuth(c, az)
            O.camera_elevation(c, el)
            O.camera_zoom(c, zoom)
            out.append(c)
        return out

    def frames():
        res = []
        for c in cams():
            ctx.trace_to_image(0, c, W, H, sd, rmin, rmax, write_canvas=True)
            res.append(ctx.canvas_download(W, H))
        # canvas-in/out with a depth clamp, the list path and the layer paths
        c = cams()[1]
        rgba0 = np.full((H * W, 4), 0.25, np.float32)
        depth0 = np.full(H * W, 0.97, np.float32)
        ctx.canvas_upload(W, H, rgba0, depth0)
        ctx.trace_to_canvas(0, c, sd, rmin, rmax, True)
        res.append(ctx.canvas_download(W, H))
        ctx.partials_begin(W, H)
        ctx.trace_to_partials(0, c, sd, rmin, rmax, False)
        res.append((np.sort(ctx.partials_download(), order=["pixel_id", "depth"]).view(np.uint8), np.zeros(1)))
        for batched in (False, True):
            ctx.layers_begin(W, H)
            if batched:
                ctx.trace_blocks_to_layers([0], c, sd, rmin, rmax, False)
            else:
                ctx.trace_to_layer(0, c, sd, rmin, rmax, False)
            ctx.layers_composite_to_canvas(c, canvas_is_clear=True)
            res.append(ctx.canvas_download(W, H))
        return res

    publish(dom["field"], False)
    ref = frames()
    assert (ref[0][0][:, 3] > 0).sum() > 1000
    keep, host = _pinned(dom["field"])
    # first publish: poison, looked at from a few view points, so that plenty of lines are resident
    host[:] = np.nan
    publish(host, True)
    ctx.trace_to_image(0, cams()[0], W, H, sd, rmin, rmax, write_canvas=True)
    ctx.trace_to_image(0, cams()[2], W, H, sd, rmin, rmax, write_canvas=True)
    ctx.synchronize()
    # second publish of the same array with the real values
    host[:] = dom["field"].reshape(-1)
    publish(host, True)
    got = frames()
    for k, ((ra, rd), (ga, gd)) in enumerate(zip(ref, got)):
        assert np.array_equal(ra, ga, equal_nan=True) and np.array_equal(rd, gd, equal_nan=True), "frame %d differs" % k
    with pytest.raises(_lib.VRError):
        publish(dom["field"].copy(), True)   # pageable memory cannot be staged
    ctx.block_free(0)
    del keep


def test_demand_staging_fetches_a_fraction_of_a_sparse_sampled_block(ctx):
    """the point of staging: at the default sampling a frame needs only part of the block.  Count the
    resident lines through the library's own statistic."""
    n = 160
    dom = datasets.braid_uniform(n, dtype=np.float32)
    b = datasets.domain_bounds(dom)
    W, H = 480, 270
    cam = O.camera_reset_to_bounds(b)
    ctx.set_tf(color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut())
    sd = O.sample_distance(b, 18)    # a step of ~15 voxels: the regime of a 512^3+ block at samples = 100
    rmin, rmax = scenes.field_range([dom])
    keep, host = _pinned(dom["field"])
    ctx.block_uniform(0, dom["dims"], dom["origin"], dom["spacing"], host, staged=True)
    assert ctx.block_staged_bytes(0) == 0
    ctx.trace_to_image(0, cam, W, H, sd, rmin, rmax, write_canvas=True)
    first = ctx.block_staged_bytes(0)
    assert 0 < first < 0.6 * dom["field"].nbytes
    ctx.trace_to_image(0, cam, W, H, sd, rmin, rmax, write_canvas=True)
    assert ctx.block_staged_bytes(0) == first          # same view: nothing new to fetch
    O.camera_azimuth(cam, 90.0)
    ctx.trace_to_image(0, cam, W, H, sd, rmin, rmax, write_canvas=True)
    assert first < ctx.block_staged_bytes(0) <= dom["field"].nbytes + 128
    # many views per publish: once most lines are resident the rest is pulled in one sweep
    ref = None
    for k in range(12):
        O.camera_azimuth(cam, 30.0)
        O.camera_elevation(cam, 7.0)
        ctx.trace_to_image(0, cam, W, H, sd, rmin, rmax, write_canvas=True)
        ctx.synchronize()
    assert ctx.block_staged_bytes(0) >= dom["field"].nbytes
    got = ctx.canvas_download(W, H)
    ctx.block_uniform(1, dom["dims"], dom["origin"], dom["spacing"], dom["field"])
    ctx.trace_to_image(1, cam, W, H, sd, rmin, rmax, write_canvas=True)
    ref = ctx.canvas_download(W, H)
    assert np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1])
    ctx.block_free(0)
    ctx.block_free(1)
    del keep


@pytest.mark.parametrize("n_slabs", [40, 200])
def test_layers_deep_pixels(ctx, n_slabs):
    """more entries per pixel than the in-register ordering holds (32) and more layers over one
    tile than the tile list holds (192): the exact fallbacks."""
    nx = 6
    doms = []
    z = np.linspace(-10.0, 10.0, n_slabs + 1).astype(np.float32)
    rng = np.random.default_rng(4)
    for k in range(n_slabs):
        sp = [20.0 / (nx - 1), 20.0 / (nx - 1), float(z[k + 1] - z[k])]
        doms.append(dict(kind="uniform", dims=(nx, nx, 2), origin=[-10., -10., float(z[k])], spacing=sp,
                         field=rng.random(nx * nx * 2, dtype=np.float32) * 0.3 + 0.2, assoc="point"))
    gb = datasets.union_bounds([datasets.domain_bounds(d) for d in doms])
    W, H = 64, 40
    cam = O.camera_reset_to_bounds(gb)
    O.camera_azimuth(cam, 3.0)
    sc = dict(doms=doms, cam=cam, W=W, H=H,
              lut=color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(400).lut(),
              sample_dist=O.sample_distance(gb, 400), rmin=0.0, rmax=1.0)
    ref_partials, ref_rgba, ref_depth = scenes.oracle_path_b(sc)
    ctx.set_tf(sc["lut"])
    for i, d in enumerate(doms):
        ctx.block_from_domain(1000 + i, d)
    ctx.layers_begin(W, H)
    for i in range(len(doms)):
        ctx.trace_to_layer(1000 + i, cam, sc["sample_dist"], 0.0, 1.0, False)
    ctx.layers_composite_to_canvas(cam, canvas_is_clear=True)
    rgba, depth = ctx.canvas_download(W, H)
    for i in range(len(doms)):
        ctx.block_free(1000 + i)
    per_px = np.bincount(np.concatenate([O.render_partials(scenes.oracle_block(d), cam, W, H, sc["lut"],
                                                           sc["sample_dist"], 0.0, 1.0, ref_depth * 0 + 1.001)["pixel_id"]
                                         for d in doms]), minlength=W * H)
    assert per_px.max() > 32
    assert np.array_equal(rgba, ref_rgba)


def test_deep_pixels_local_and_global_sort(ctx):
    """segments longer than the in-register sort limit (32) take the in-place global path"""
    rng = np.random.default_rng(9)
    W, H = 16, 8
    n = 6000
    p = np.zeros(n, O.PARTIAL_DTYPE)
    p["pixel_id"] = rng.integers(0, 40, n)          # ~150 partials per pixel
    p["depth"] = rng.random(n, dtype=np.float32)
    a = rng.random(n, dtype=np.float32) * 0.05
    p["alpha"] = a
    p["rgb"] = rng.random((n, 3), dtype=np.float32) * a[:, None]
    mine = np.sort(ctx.composite_partials(p, W, H), order="pixel_id")
    ref = np.sort(O.composite_partials([p]), order="pixel_id")
    assert mine.tobytes() == ref.tobytes()


# ------------------------------------------------------------------ compositing kernels
def test_quantise_and_fold_bit_exact(ctx):
    rng = np.random.default_rng(11)
    W, H, n_img = 333, 77, 5
    alpha = rng.random((n_img, H * W, 1), dtype=np.float32)
    rgba = np.concatenate([rng.random((n_img, H * W, 3), dtype=np.float32) * alpha, alpha], axis=2)
    rgba[rng.random((n_img, H * W)) < 0.25] = 0
    rgba[0, :10] = 1.0
    depth = (rng.random((n_img, H * W), dtype=np.float32) * 1.3 - 0.1).astype(np.float32)
    depth[1, :5] = np.nan
    depth[2, 5:9] = np.inf
    order = rng.permutation(n_img).astype(np.int32)
    out, od = ctx.composite_images(rgba, depth, order, W, H)
    q = [O.image_init(rgba[i], depth[i], 0) for i in range(n_img)]
    ref, rd = O.ordered_composite(np.stack([x[0] for x in q]), np.stack([x[1] for x in q]), order)
    assert np.array_equal(out, ref)
    assert np.array_equal(od, rd, equal_nan=True)


def test_apcomp_known_answer_scene_through_abi(ctx, golden_dir):
    """the reference's own c_order scene and golden (t_apcomp_c_order.cpp) on the GPU path."""
    import os
    g = np.load(os.path.join(golden_dir, "apcomp_goldens.npz"))["apcomp_c_order"]
    imgs = [scenes.apcomp_image(i) for i in range(4)]
    out, _ = ctx.composite_images(np.stack([i[0] for i in imgs]), np.stack([i[1] for i in imgs]),
                                  np.arange(4, dtype=np.int32), 1024, 1024)
    out = out.reshape(1024, 1024, 4)
    assert np.array_equal(out, g)
    assert tuple(out[600, 450]) == (255, 128, 65, 222)


def test_apcomp_partial_scene_through_abi(ctx, golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "apcomp_goldens.npz"))["apcomp_volume_partial"]
    lists = [scenes.apcomp_partials(i) for i in range(4)]
    res = ctx.composite_partials(np.concatenate(lists), 1024, 1024)
    img = scenes.partials_to_image(res, 1024, 1024)
    assert np.array_equal(img, g)
    assert tuple(img[600, 450]) == (255, 127, 63, 223)


@pytest.mark.parametrize("seed", [0, 1])
def test_partial_composite_random_with_ties(ctx, seed):
    """random partial lists incl. equal (pixel, depth) keys: order = list order (documented
    tie-break), empty/alpha-1/alpha-0 edge cases of VolumePartial::blend."""
    rng = np.random.default_rng(seed)
    W, H = 64, 48
    n = 20000
    p = np.zeros(n, O.PARTIAL_DTYPE)
    p["pixel_id"] = rng.integers(0, W * H // 2, n)
    p["depth"] = rng.integers(0, 6, n).astype(np.float32)  # many ties
    a = rng.random(n, dtype=np.float32)
    a[rng.random(n) < 0.1] = 1.0
    a[rng.random(n) < 0.1] = 0.0
    p["alpha"] = a
    p["rgb"] = rng.random((n, 3), dtype=np.float32) * a[:, None]
    mine = np.sort(ctx.composite_partials(p, W, H), order="pixel_id")
    ref = np.sort(O.composite_partials([p]), order="pixel_id")
    assert mine.tobytes() == ref.tobytes()
    assert ctx.composite_partials(p[:0], W, H).size == 0
    one = ctx.composite_partials(p[:1], W, H)
    assert one.tobytes() == p[:1].tobytes()


@pytest.mark.parametrize("bg", [(0., 0., 0., 1.), (1., 1., 1., 1.), (0.2, 0.4, 0.6, 0.5)])
def test_frame_epilogue_background_and_rgba8(ctx, bg):
    """Render::RenderBackground + the PNGEncoder conversion of Render::Save on the device: bit-exact
    vs the oracle (which the reference's render goldens pin), fused and unfused, flipped and not."""
    sc = scenes.multi_render_scene(0)
    W, H = sc["W"], sc["H"]
    rgba, _ = gpu_path_a_canvas(ctx, sc["doms"][0], sc)
    ref = rgba.copy()
    O.blend_background(ref, bg)
    fused = ctx.canvas_download_rgba8(W, H, bg=bg, flip=True)
    assert np.array_equal(fused, O.encode_rgba8(ref, W, H, flip=True))
    again, _ = ctx.canvas_download(W, H)
    assert np.array_equal(again, rgba)          # the fused form leaves the canvas alone
    ctx.canvas_blend_background(bg)
    blended, _ = ctx.canvas_download(W, H)
    assert np.array_equal(blended, ref)
    assert np.array_equal(ctx.canvas_download_rgba8(W, H, flip=False), O.encode_rgba8(ref, W, H, flip=False))
    # values outside [0,1] and NaN convert like the x86 cast (low byte of the truncated integer)
    rng = np.random.default_rng(5)
    odd = (rng.random((H * W, 4), dtype=np.float32) * 3 - 1).astype(np.float32)
    odd[::97, 0] = np.nan
    ctx.canvas_upload(W, H, odd, np.zeros(H * W, np.float32))
    assert np.array_equal(ctx.canvas_download_rgba8(W, H, flip=True), O.encode_rgba8(odd, W, H, flip=True))


def _zbuffer_oracle(rgba, depth):
    """vtk-h Compositor in Z_BUFFER_SURFACE mode on one rank (Compositor.cpp:146-160): Image::Init of
    every image, z-selected into the first one in the order they were added."""
    front, fd = O.image_init(rgba[0], depth[0], 0)
    for i in range(1, len(rgba)):
        q, d = O.image_init(rgba[i], depth[i], 0)
        O.zbuffer_composite(front, fd, q, d, gl_depth=True)
    return front, fd


def test_zbuffer_apcomp_scene_through_abi(ctx, golden_dir):
    """t_apcomp_zbuffer.cpp:26-68 (four overlapping opaque squares, depth 0.05*i over a 1.01
    background) through vr_composite_zbuffer: bit-exact vs the oracle and the reference's golden."""
    import os
    W = H = 1024
    rgba, depth = [], []
    for i in range(4):
        c = np.float32(0.1) + np.float32(i) * np.float32(0.1)
        px = np.zeros((H, W, 4), np.float32)
        dp = np.full((H, W), 1.01, np.float32)
        px[400:700, 200 + 100 * i:500 + 100 * i] = [c, c, c, 1.0]
        dp[400:700, 200 + 100 * i:500 + 100 * i] = np.float32(i) * np.float32(0.05)
        rgba.append(px.reshape(-1, 4))
        depth.append(dp.reshape(-1))
    out, od = ctx.composite_zbuffer(np.stack(rgba), np.stack(depth), W, H)
    ref, rd = _zbuffer_oracle(rgba, depth)
    assert np.array_equal(out, ref) and np.array_equal(od, rd)
    g = np.load(os.path.join(golden_dir, "apcomp_goldens.npz"))["apcomp_zbuffer"]
    assert np.array_equal(out.reshape(H, W, 4), g)
    assert tuple(out.reshape(H, W, 4)[600, 450]) == (25, 25, 25, 255)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_zbuffer_random_with_ties_and_far_fragments(ctx, seed):
    """ImageCompositor::ZBufferComposite (ImageCompositor.hpp:49-76) edge cases: equal depths (the
    later image wins), fragments beyond depth 1 (never replace, also not a farther one), negative
    depths (Image::Init makes them positive first)."""
    rng = np.random.default_rng(seed)
    W, H, n = 203, 77, 5
    rgba = rng.random((n, H * W, 4), dtype=np.float32)
    depth = rng.choice(np.array([0.0, 0.25, 0.25, 0.5, 0.75, 1.0, 1.001, 1.01, 1.5, -0.3], np.float32),
                       size=(n, H * W))
    depth[:, ::7] = rng.random((n, depth[:, ::7].shape[1]), dtype=np.float32)
    out, od = ctx.composite_zbuffer(rgba, depth, W, H)
    ref, rd = _zbuffer_oracle(list(rgba), list(depth))
    assert np.array_equal(out, ref) and np.array_equal(od, rd)


def test_path_a_single_rank_roundtrip(ctx):
    """Image::Init -> (single image, no blend) -> ImageToCanvas: canvas becomes k/255."""
    sc = scenes.multi_render_scene(0)
    dom = sc["doms"][0]
    gpu_path_a_canvas(ctx, dom, sc)
    ctx.image_from_canvas()
    u8, d = ctx.image_download(sc["W"], sc["H"])
    rp, dp = ctx.image_ptrs()
    ctx.image_to_canvas_dev(rp, dp)
    can, cd = ctx.canvas_download(sc["W"], sc["H"])
    o_u8, o_d, o_can = scenes.oracle_path_a(sc)
    assert np.array_equal(u8, o_u8)
    assert np.array_equal(can, o_can)


@pytest.mark.parametrize("res", [(256, 256), (1024, 1024), (642, 360), (1920, 1080), (37, 19)])
@pytest.mark.parametrize("az", [0.0, 33.0])
def test_fused_frame_equals_unfused_path_a(ctx, res, az):
    """vr_trace_to_image (clear + K1-K7 + Image::Init + ImageToCanvas in one launch) against the
    four separate calls and against the oracle: uint8 image, depth and canvas bit-identical."""
    dom = datasets.braid_uniform(32, dtype=np.float32)
    b = datasets.domain_bounds(dom)
    W, H = res
    cam = O.camera_reset_to_bounds(b)
    O.camera_azimuth(cam, az)
    O.camera_elevation(cam, az / 2)
    if az:
        cam.zoom = 1.7  # part of the volume leaves the screen
    lut = color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut()
    sd = O.sample_distance(b, 100)
    rmin, rmax = scenes.field_range([dom])
    ctx.block_from_domain(0, dom)
    ctx.set_tf(lut)
    # unfused
    ctx.canvas_clear(W, H)
    ctx.trace_to_canvas(0, cam, sd, rmin, rmax, False)
    ctx.image_from_canvas()
    u8_a, d_a = ctx.image_download(W, H)
    rp, dp = ctx.image_ptrs()
    ctx.image_to_canvas_dev(rp, dp)
    can_a, cd_a = ctx.canvas_download(W, H)
    # poison the buffers, then fused
    ctx.canvas_upload(W, H, np.full((H * W, 4), 0.37, np.float32), np.full(H * W, 0.5, np.float32))
    ctx.image_from_canvas()
    ctx.trace_to_image(0, cam, W, H, sd, rmin, rmax, write_canvas=True)
    u8_b, d_b = ctx.image_download(W, H)
    can_b, cd_b = ctx.canvas_download(W, H)
    assert np.array_equal(u8_a, u8_b)
    assert np.array_equal(d_a, d_b, equal_nan=True)
    assert np.array_equal(can_a, can_b)
    assert np.array_equal(cd_a, cd_b, equal_nan=True)
    # and the oracle's path A for one rank
    sc = dict(doms=[dom], W=W, H=H, cam=cam, lut=lut, sample_dist=sd, rmin=rmin, rmax=rmax,
              dom_bounds=[b])
    o_u8, o_d, o_can = scenes.oracle_path_a(sc)
    assert np.array_equal(u8_b, o_u8)
    assert np.array_equal(can_b, o_can)


def test_fused_frame_no_clear_leaves_outside_untouched(ctx):
    dom = datasets.braid_uniform(16, dtype=np.float32)
    b = datasets.domain_bounds(dom)
    W, H = 400, 300
    cam = O.camera_reset_to_bounds(b)
    lut = color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut()
    sd = O.sample_distance(b, 100)
    rmin, rmax = scenes.field_range([dom])
    ctx.block_from_domain(0, dom)
    ctx.set_tf(lut)
    ctx.canvas_upload(W, H, np.full((H * W, 4), 0.5, np.float32), np.full(H * W, 0.25, np.float32))
    ctx.image_from_canvas()                      # image = (127,127,127,127), depth .25 everywhere
    ctx.trace_to_image(0, cam, W, H, sd, rmin, rmax, no_clear=True)
    u8, d = ctx.image_download(W, H)
    ctx.trace_to_image(0, cam, W, H, sd, rmin, rmax)
    u8_full, d_full = ctx.image_download(W, H)
    sx, sy, sw, sh = _lib.find_subset(cam, W, H, b)
    x0, x1 = sx & ~3, min(W, (sx + sw + 3) & ~3)
    inside = np.zeros((H, W), bool)
    inside[sy:sy + sh, x0:x1] = True
    inside = inside.reshape(-1)
    assert np.array_equal(u8[inside], u8_full[inside])
    assert np.array_equal(d[inside], d_full[inside], equal_nan=True)
    assert (u8[~inside] == 127).all() and (d[~inside] == 0.25).all()
    assert (u8_full[~inside] == 0).all() and (d_full[~inside] == np.float32(1.001)).all()


def test_two_rank_path_a_on_one_gpu(ctx, golden_dir):
    """the reference's 2-rank MPI volume scene: each rank's image rendered and quantised on the
    GPU, folded in visibility order on the GPU; uint8 result bit-exact vs the oracle, and inside
    the reference golden's tolerance."""
    import os
    sc = scenes.mpi_volume_scene()
    W, H = sc["W"], sc["H"]
    layers, depths = [], []
    for dom in sc["doms"]:
        gpu_path_a_canvas(ctx, dom, sc)
        ctx.image_from_canvas()
        u8, d = ctx.image_download(W, H)
        layers.append(u8)
        depths.append(d)
    order = _lib.visibility_order(sc["dom_bounds"], sc["cam"])
    import torch
    lr = torch.from_numpy(np.stack(layers)).cuda()
    ld = torch.from_numpy(np.stack(depths)).cuda()
    out = torch.empty((H * W, 4), dtype=torch.uint8, device="cuda")
    od = torch.empty(H * W, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    ctx.fold_images_dev(lr.data_ptr(), ld.data_ptr(), H * W, order, H * W, out.data_ptr(), od.data_ptr())
    ctx.synchronize()
    o_u8, o_d, o_can = scenes.oracle_path_a(sc)
    assert np.array_equal(out.cpu().numpy(), o_u8)
    g = np.load(os.path.join(golden_dir, "tout_render_mpi_3d_diy_volume100.npz"))
    can, _ = O.image_to_canvas(out.cpu().numpy(), od.cpu().numpy())
    png = scenes.png_bytes(can, W, H)
    from test_oracle_golden import crop_stats
    assert crop_stats(png, g["rgb"], g["rects"]) <= 0.01


def test_errors_are_reported_not_thrown(ctx):
    with pytest.raises(_lib.VRError):
        ctx.block_free(12345)
    with pytest.raises(_lib.VRError):
        ctx.trace_to_canvas(999, O.camera_default(), 0.1, 0., 1.)
    with pytest.raises(_lib.VRError):
        ctx.set_tf(np.zeros((1, 4), np.float32))
    with pytest.raises(_lib.VRError):
        ctx.block_uniform(0, (1, 4, 4), [0, 0, 0], [1, 1, 1], np.zeros(16, np.float32))
    # a sample distance too small to advance a float distance would spin on the device: refused on the host
    dom = datasets.braid_uniform(8)
    b = datasets.domain_bounds(dom)
    ctx.set_tf(np.ones((16, 4), np.float32))
    ctx.block_from_domain(1, dom)
    ctx.canvas_clear(64, 64)
    with pytest.raises(_lib.VRError):
        ctx.trace_to_canvas(1, O.camera_reset_to_bounds(b), 1e-9, 0., 1., False)
    ctx.trace_to_canvas(1, O.camera_reset_to_bounds(b), 0.3, 0., 1., False)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_strided_blueprint_field_renders_like_the_dense_copy(ctx, dtype):
    """ascent_vtkh_data_adapter.cpp:1836-1887: Blueprint values with an element stride (component 1 of an
    interleaved 3-component mcarray here) reach the sampler through vr_field_gather_strided -- from host
    memory and from device memory -- and render bit-identically to the same values published densely."""
    import torch
    dom = datasets.braid_uniform(24, dtype=dtype)
    b = datasets.domain_bounds(dom)
    W, H = 256, 192
    cam = O.camera_reset_to_bounds(b)
    O.camera_azimuth(cam, 20.0)
    lut = color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut()
    sd = O.sample_distance(b, 100)
    rmin, rmax = scenes.field_range([dom])
    ctx.set_tf(lut)
    ctx.block_from_domain(0, dom)
    ctx.canvas_clear(W, H)
    ctx.trace_to_canvas(0, cam, sd, rmin, rmax, False)
    want_rgba, want_depth = ctx.canvas_download(W, H)
    n = dom["field"].size
    inter = np.empty((n, 3), dtype)
    inter[:, 0] = -7.0
    inter[:, 1] = dom["field"].reshape(-1)
    inter[:, 2] = 1e30
    dt = _lib.VR_F32 if dtype == np.float32 else _lib.VR_F64
    dev = torch.from_numpy(inter).cuda()
    for src in ("host", "device"):
        if src == "host":
            dense = ctx.field_gather_strided(inter, n, 3, 1)
        else:
            dense = ctx.field_gather_strided(None, n, 3, 1, device_ptr=dev.data_ptr(), dtype=dt)
        ctx.block_uniform(0, dom["dims"], dom["origin"], dom["spacing"], None, device_ptr=dense, dtype=dt)
        ctx.canvas_clear(W, H)
        ctx.trace_to_canvas(0, cam, sd, rmin, rmax, False)
        rgba, depth = ctx.canvas_download(W, H)
        ctx.block_free(0)
        ctx.field_free(dense)
        assert np.array_equal(rgba, want_rgba), src
        assert np.array_equal(depth, want_depth, equal_nan=True), src


@pytest.mark.parametrize("W,H,bg", [(256, 192, None), (203, 77, (0.1, 0.2, 0.3, 1.0)), (1920, 1080, (1.0, 1.0, 1.0, 1.0))])
def test_png_encoded_on_the_device(ctx, W, H, bg):
    """vr_canvas_encode_png (Render::Save, Render.cpp:299-312 / ascent_png_encoder.cpp:258-303, on the GPU): the file
    decodes (PIL) to exactly the pixels of vr_canvas_download_rgba8 -- the reference encoder's float -> uint8
    conversion with flipped rows, oracle-checked in test_frame_epilogue_background_and_rgba8 -- and is byte for
    byte the stream of the CPU model (tests/png_model.py): deflate blocks, Adler-32 and every chunk CRC."""
    import io
    from PIL import Image
    import png_model
    dom = datasets.braid_uniform(24, dtype=np.float32)
    b = datasets.domain_bounds(dom)
    cam = O.camera_reset_to_bounds(b)
    O.camera_azimuth(cam, 35.0)
    lut = color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut()
    sd = O.sample_distance(b, 100)
    rmin, rmax = scenes.field_range([dom])
    ctx.set_tf(lut)
    ctx.block_from_domain(0, dom)
    ctx.canvas_clear(W, H)
    ctx.trace_to_canvas(0, cam, sd, rmin, rmax, False)
    bgv = None if bg is None else np.array(bg, np.float32)
    want = np.asarray(ctx.canvas_download_rgba8(W, H, bgv, flip=True)).reshape(H, W, 4)
    png = ctx.canvas_encode_png(W, H, bgv)
    assert png[:8] == b"\x89PNG\r\n\x1a\n"
    img = Image.open(io.BytesIO(png))
    img.load()  # (verifies the chunk CRCs and the Adler-32)
    assert img.mode == "RGBA" and img.size == (W, H)
    assert np.array_equal(np.array(img), want)
    if W * H <= 100000:  # (the model is a pure-Python loop over every byte)
        model, _, _ = png_model.encode(want)
        assert len(png) == len(model) and png == model
    assert len(png) <= _lib.load().vr_png_bound(W, H)
    # the cleared background collapses (a few bytes per run); where the volume covers the frame the stream stays
    # near the raw size (Huffman-only, like the reference's lodepng settings)
    assert len(png) < (W * H * 4 // 2 if W == 1920 else W * H * 4)
    ctx.block_free(0)


@pytest.mark.parametrize("W,H,az", [(640, 360, 0.0), (203, 77, 40.0)])
def test_footprint_read_back_equals_the_full_canvas(ctx, W, H, az):
    """vr_canvas_download_rect: a frame that started from Canvas::Clear differs from the cleared canvas only inside
    the screen footprint of the data, so reading back that rectangle into a host canvas that holds the cleared canvas
    (Render::ClearCanvas) gives the whole frame -- also when the camera moves, if the previous footprint is included"""
    import bench
    dom = datasets.braid_uniform(20, dtype=np.float32)
    b = datasets.domain_bounds(dom)
    lut = color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut()
    sd = O.sample_distance(b, 100)
    rmin, rmax = scenes.field_range([dom])
    ctx.set_tf(lut)
    ctx.block_from_domain(0, dom)
    host_rgba, host_depth = O.new_canvas(W, H)
    prev = None
    for k in range(3):
        cam = O.camera_reset_to_bounds(b)
        O.camera_azimuth(cam, az + 25.0 * k)
        O.camera_zoom(cam, 0.1 * k)  # (vtkm zoom: factor 4^z)
        ctx.trace_to_image(0, cam, W, H, sd, rmin, rmax, write_canvas=True)
        rect = bench.footprint_rect(cam, W, H, [b])
        ctx.canvas_download_rect(bench.union_rect(prev, rect), host_rgba, host_depth)
        prev = rect
        full_rgba, full_depth = ctx.canvas_download(W, H)
        assert (rect[2] - rect[0]) * (rect[3] - rect[1]) < W * H
        assert np.array_equal(host_rgba, full_rgba) and np.array_equal(host_depth, full_depth), k
    ctx.block_free(0)
