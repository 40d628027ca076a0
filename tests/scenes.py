"""Scene builders shared by the oracle tests (CPU) and the parity tests (GPU).

Each builder returns plain data (domains, camera parameters, LUT, range, sample distance) so the
same inputs can be handed to the oracle and to the C-ABI."""
import numpy as np

from ascent_b200 import datasets
from oracle import oracle as O

MULTI_RENDER_TF = {"name": "blue", "control_points": [
    {"type": "rgb", "position": 0., "color": [1., 0., 0.]},
    {"type": "rgb", "position": 0.5, "color": [0., 1., 0.]},
    {"type": "rgb", "position": 1.0, "color": [1., 1., 1.]},
    {"type": "alpha", "position": 0., "alpha": 0.},
    {"type": "alpha", "position": 1.0, "alpha": 1.}]}

RAMP_TF = {"name": "cool to warm", "control_points": [
    {"type": "alpha", "position": 0., "alpha": 0.},
    {"type": "alpha", "position": 1.0, "alpha": 1.}]}


def oracle_block(dom):
    if dom["kind"] == "uniform":
        return O.OracleBlock(dom["dims"], dom["field"], origin=dom["origin"], spacing=dom["spacing"],
                             cell_assoc=dom.get("assoc") == "cell")
    return O.OracleBlock(dom["dims"], dom["field"], axes=dom["axes"],
                         cell_assoc=dom.get("assoc") == "cell")


def field_range(doms):
    return (min(float(d["field"].min()) for d in doms), max(float(d["field"].max()) for d in doms))


def png_bytes(canvas_rgba, W, H, bg=(0., 0., 0., 1.)):
    """Render::RenderBackground (BlendBackground: c + bg*(1-a)) then PNGEncoder
    ((uchar)(c*255.f), ascent_png_encoder.cpp:274-281); rows left un-flipped."""
    c = np.ascontiguousarray(canvas_rgba, np.float32).copy()
    O.blend_background(c, bg)
    return O.encode_rgba8(c, W, H, flip=False)


def multi_render_scene(which):
    """t_ascent_render_3d.cpp:1643-1780: braid uniform 20^3; r1 = 512^2 default camera,
    r2 = 400^2 fully specified camera."""
    dom = datasets.braid_uniform(20)
    bounds = datasets.domain_bounds(dom)
    cam = O.camera_reset_to_bounds(bounds)
    if which == 0:
        W = H = 512
    else:
        W = H = 400
        cam.look_at[:] = [1., 1., 1.]
        cam.position[:] = [0., 25., 15.]
        cam.up[:] = [0., -1., 0.]
        cam.fov = 60.
        O.camera_zoom(cam, 0.0)  # zoom 1.0 -> log4(1) = 0 (parsing.cpp:59-69)
        cam.near_plane, cam.far_plane = 0.1, 100.1
        O.camera_azimuth(cam, 10.0)
        O.camera_elevation(cam, -10.0)
    lut = O.parse_color_table(MULTI_RENDER_TF).correct_opacity(100).lut()
    rmin, rmax = field_range([dom])
    return dict(doms=[dom], cam=cam, W=W, H=H, lut=lut, rmin=rmin, rmax=rmax,
                sample_dist=O.sample_distance(bounds, 100), bounds=bounds)


def mpi_volume_scene():
    """t_ascent_mpi_render_3d.cpp:284-388: 2 ranks, rectilinear radial_vert, azimuth 45."""
    doms = [datasets.radial_example(32, r, 2) for r in range(2)]
    bl = [datasets.domain_bounds(d) for d in doms]
    gb = datasets.union_bounds(bl)
    cam = O.camera_reset_to_bounds(gb)
    O.camera_azimuth(cam, 45.0)
    tf = O.parse_color_table({"control_points": [
        {"type": "alpha", "position": 0., "alpha": 0.8},
        {"type": "alpha", "position": 1.0, "alpha": 0.0}]})
    rmin, rmax = field_range(doms)
    return dict(doms=doms, cam=cam, W=512, H=512, lut=tf.correct_opacity(100).lut(), rmin=rmin,
                rmax=rmax, sample_dist=O.sample_distance(gb, 100), bounds=gb, dom_bounds=bl)


def oracle_path_a(scene):
    """RenderOneDomainPerRank + Composite (VolumeRenderer.cpp:482-536,652-688) with one domain per
    rank: returns (uint8 image, depth, float canvas)."""
    W, H = scene["W"], scene["H"]
    layers, depths = [], []
    for dom in scene["doms"]:
        rgba, depth = O.new_canvas(W, H)
        O.render_to_canvas(oracle_block(dom), scene["cam"], W, H, scene["lut"], scene["sample_dist"],
                           scene["rmin"], scene["rmax"], rgba, depth)
        u8, d = O.image_init(rgba, depth, 0)
        layers.append(u8)
        depths.append(d)
    bl = [datasets.domain_bounds(d) for d in scene["doms"]]
    order, _ = O.visibility_order(np.array(bl), scene["cam"])
    out, od = O.ordered_composite(np.stack(layers), np.stack(depths), order)
    can, cd = O.image_to_canvas(out, od)
    return out, od, can


def oracle_path_b(scene):
    """RenderMultipleDomainsPerRank (VolumeRenderer.cpp:539-597), single rank: returns
    (composited partials, float canvas rgba, canvas depth)."""
    W, H = scene["W"], scene["H"]
    rgba, depth = O.new_canvas(W, H)
    plists = [O.render_partials(oracle_block(dom), scene["cam"], W, H, scene["lut"],
                                scene["sample_dist"], scene["rmin"], scene["rmax"], depth)
              for dom in scene["doms"]]
    res = O.composite_partials(plists)
    O.partials_to_canvas(res, scene["cam"], W, H, rgba, depth)
    return res, rgba, depth


# --- apcomp known-answer scenes (src/tests/apcomp/t_apcomp_test_utils.h:20-87) ----------------
APCOMP_COLORS = np.array([[1., 0., 0., .5], [0., 1., 0., .5], [0., 0., 1., .5], [0., 1., 1., .5]],
                         np.float32)


def apcomp_image(i, width=1024, height=1024, square=300, y=500, colors=APCOMP_COLORS):
    """gen_float32_image(..., depth=i*0.05, bottom_x=200+100*i, bottom_y=y-50*i, ...)"""
    px = np.zeros((height, width, 4), np.float32)
    dp = np.full((height, width), 1.01, np.float32)
    bx, by = 200 + 100 * i, y - 50 * i
    px[by:by + square, bx:bx + square] = colors[i]
    dp[by:by + square, bx:bx + square] = np.float32(i) * np.float32(0.05)
    return px.reshape(-1, 4), dp.reshape(-1)


def apcomp_partials(i, width=1024, square=300, y=500, colors=APCOMP_COLORS):
    """gen_float32_partials with the same geometry."""
    bx, by = 200 + 100 * i, y - 50 * i
    yy, xx = np.meshgrid(np.arange(by, by + square), np.arange(bx, bx + square), indexing="ij")
    p = np.zeros(square * square, O.PARTIAL_DTYPE)
    p["pixel_id"] = (yy * width + xx).reshape(-1)
    p["rgb"] = colors[i, :3]
    p["alpha"] = colors[i, 3]
    p["depth"] = np.float32(i) * np.float32(0.05)
    return p


def partials_to_image(partials, width, height):
    """partials_to_png (t_apcomp_test_utils.h:89-123) + PNGEncoder float path."""
    img = np.zeros((height * width, 4), np.float32)
    img[partials["pixel_id"], :3] = partials["rgb"]
    img[partials["pixel_id"], 3] = partials["alpha"]
    return (img * np.float32(255.0)).astype(np.uint8).reshape(height, width, 4)


def cinema_camera(bounds, phi, theta):
    """CinemaManager::create_cinema_cameras (ascent_runtime_rendering_filters.cpp:906-960), one angle
    pair: the oracle's restatement."""
    return O.camera_cinema(bounds, phi, theta)


GHOST_TF = {"name": "rainbow desaturated", "control_points": [
    {"type": "alpha", "position": 0., "alpha": 0.2},
    {"type": "alpha", "position": 1.0, "alpha": 0.5}]}


def ghost_volume_scene():
    """t_ascent_multi_topo.cpp:181-252 (single_ghost_vol_render): braid uniform 10^3 with a RAGGED ghost field (cells
    0 and 1 flagged, t_utils.hpp:517-553) -- vtk-h's ghost stripper cannot cut a structured sub-box out of that, so it
    thresholds the cells away and the volume renderer receives an explicit cell set: the unstructured path (N4).
    Volume plot of the vertex field `braid`, fixed range [-0.5, 0.5], azimuth 170, elevation 11, 1024^2, no
    annotations.  Returns the mesh as (points, hexahedra, point field)."""
    dom = datasets.braid_uniform(10, dtype=np.float64)
    bounds = datasets.domain_bounds(dom)
    cam = O.camera_reset_to_bounds(bounds)
    O.camera_azimuth(cam, 170.0)
    O.camera_elevation(cam, 11.0)
    pts, conn = datasets.structured_to_hexes(dom["dims"], dom["origin"], dom["spacing"], drop_cells=[0, 1])
    lut = O.parse_color_table(GHOST_TF).correct_opacity(100).lut()
    return dict(points=pts, conn=conn, field=dom["field"].reshape(-1), cam=cam, W=1024, H=1024, lut=lut, rmin=-0.5,
                rmax=0.5, sample_dist=O.sample_distance(bounds, 100), bounds=bounds, dom=dom)


def oracle_unstructured_path_b(scene, mesh=None):
    """RenderMultipleDomainsPerRank with an unstructured domain (VolumeRenderer.cpp:539-597 + UnstructuredWrapper
    :182-221): partials -> PartialCompositor -> partials_to_canvas.  Returns (partials, canvas rgba, canvas depth)."""
    W, H = scene["W"], scene["H"]
    um = mesh or O.OracleUMesh(scene["points"], scene["conn"], scene["field"], cell_assoc=scene.get("cell_assoc", False))
    rgba, depth = O.new_canvas(W, H)
    parts = O.render_umesh_partials(um, scene["cam"], W, H, scene["lut"], scene["sample_dist"], scene["rmin"],
                                    scene["rmax"], depth)
    res = O.composite_partials([parts])
    O.partials_to_canvas(res, scene["cam"], W, H, rgba, depth)
    return parts, rgba, depth
