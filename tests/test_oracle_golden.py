"""Pin the ray-cast oracle (oracle/raycast_oracle.c) against the reference's own golden images.

The goldens of the volume path were rendered by TWO generations of VTK-m that differ in one convention of the
structured sampler, where a ray's first sample sits behind its entry into the block:
  * entry + 1e-4 (absolute): render_0100.png, render_1100.png, tout_render_mpi_3d_diy_volume100.png.  With it the
    restatement reproduces these three uint8 FOR uint8 (every pixel of one crop, 99.8 % / 99.95 % of the others);
  * entry + 1e-4 * |block extent| (VTK-m's `meshEpsilon`, what SURVEY appendix B9 recalls of the pinned v2.1.0):
    tout_render_3d_multi_default_runtime100.png, a pseudocolor + volume scene whose pixels are the volume plot's
    alone wherever no ray meets the contour surface.  This is the DEFAULT of oracle and product.
Both are pinned below; `O.set_first_sample_offset` / `vr_set_first_sample_offset` switch between them.  The older
goldens stay inside the reference's own PNGCompare tolerance of the newer convention, which is why the reference
never regenerated them.

The goldens are whole Ascent renders with annotations burned in, so the comparison runs over
hand-picked annotation-free crops (tests/golden/make_golden.py) with the reference's own metric:
a pixel differs if any channel is off by more than 4/255 (ascent_png_compare.cpp:35,138-141),
and the test fails if the differing fraction exceeds the tolerance the reference test used
(check_test_image default 0.001; 0.01 where the test passes 0.01f)."""
import os

import numpy as np
import pytest

import scenes


def crop_diffs(mine, gold, rects):
    return np.concatenate([np.abs(mine[y0:y1, x0:x1, :3].astype(int) - gold[y0:y1, x0:x1].astype(int)).max(axis=2)
                           .reshape(-1) for (y0, y1, x0, x1) in rects])


def crop_stats(mine, gold, rects):
    bad = tot = 0
    for (y0, y1, x0, x1) in rects:
        d = np.abs(mine[y0:y1, x0:x1, :3].astype(int) - gold[y0:y1, x0:x1].astype(int)).max(axis=2)
        bad += int((d > 4).sum())
        tot += d.size
    return bad / tot


@pytest.mark.parametrize("which,name,tol", [(0, "render_0100", 0.001), (1, "render_1100", 0.01)])
def test_multi_render_golden(golden_dir, which, name, tol):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sc = scenes.multi_render_scene(which)
    _, _, canvas = scenes.oracle_path_a(sc)
    mine = scenes.png_bytes(canvas, sc["W"], sc["H"])
    assert crop_stats(mine, g["rgb"], g["rects"]) <= tol


def test_mpi_volume_golden(golden_dir):
    """2-rank path A: rectilinear sampler, visibility order, uint8 ordered fold."""
    g = np.load(os.path.join(golden_dir, "tout_render_mpi_3d_diy_volume100.npz"))
    sc = scenes.mpi_volume_scene()
    _, _, canvas = scenes.oracle_path_a(sc)
    mine = scenes.png_bytes(canvas, sc["W"], sc["H"])
    assert crop_stats(mine, g["rgb"], g["rects"]) <= 0.01


def test_default_camera_geometry():
    """SURVEY K0 corroboration: in render_0100.png the cube's front face spans
    10/(24.64*tan30) = 0.703 of the image -> subset rect of the 20^3 braid at 512^2."""
    from oracle import oracle as O
    sc = scenes.multi_render_scene(0)
    sx, sy, sw, sh = O.find_subset(sc["cam"], 512, 512, sc["bounds"])
    assert abs(sw / 512.0 - 0.703) < 0.01 and abs(sh / 512.0 - 0.703) < 0.01
    assert abs((sx + sw / 2) - 256) <= 1 and abs((sy + sh / 2) - 256) <= 1


import contextlib


@contextlib.contextmanager
def first_sample(abs_offset, extent_rel):
    from oracle import oracle as O
    O.set_first_sample_offset(abs_offset, extent_rel)
    try:
        yield
    finally:
        O.set_first_sample_offset()


OLDER_GENERATION = (1e-4, 0.0)   # entry + 1e-4
MESH_EPSILON = (0.0, 1e-4)       # entry + 1e-4 * |extent| (default)


def _golden_diff(golden_dir, which, name, offset=OLDER_GENERATION):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sc = scenes.multi_render_scene(which) if which < 2 else scenes.mpi_volume_scene()
    with first_sample(*offset):
        _, _, canvas = scenes.oracle_path_a(sc)
    return crop_diffs(scenes.png_bytes(canvas, sc["W"], sc["H"]), g["rgb"], g["rects"])


GOLDENS = [(0, "render_0100"), (1, "render_1100"), (2, "tout_render_mpi_3d_diy_volume100")]


@pytest.mark.parametrize("which,name,exact,within1,worst", [(0, "render_0100", 0.998, 0.9999, 1),
                                                            (1, "render_1100", 1.0, 1.0, 0),
                                                            (2, "tout_render_mpi_3d_diy_volume100", 0.9995, 0.9995, 2)])
def test_older_goldens_pin_the_oracle_uint8_for_uint8(golden_dir, which, name, exact, within1, worst):
    """The reference's PNGCompare only asks for <= 0.1 % (1 %) of pixels off by more than 4/255.  With the
    first-sample offset of the VTK-m generation that rendered them, the restated K0-K8 chain reproduces the three
    pure-volume goldens uint8 FOR uint8: every pixel of the render_1100 crop, 99.8 % of render_0100's and 99.95 % of
    the 2-rank rectilinear scene's, and -- the north star's own bar, here against the reference's images rather
    than against the oracle -- >= 99.9 % within 1/255 and none beyond 3/255.  What is left is not the volume:
    render_0100's off-by-one pixels sit exactly on the screen diagonals |x-256| == |y-256| (and two columns) where
    the golden shows the white bounding-box edges through rays whose opacity stops just short of 1 (green 255
    instead of 254), the mpi scene's 23 pixels (off by 2) are one 1-pixel-wide line of the same annotation
    crossing the top of the crops.  Everything else of the chain -- camera, ray generation, subset, bounds test,
    locator, trilinear form, table sampling and index, opacity correction, blend, termination, quantisation,
    visibility order, uint8 fold -- is thereby pinned exactly, whichever first-sample offset is in force."""
    d = _golden_diff(golden_dir, which, name)
    assert (d == 0).mean() >= exact, (d == 0).mean()
    assert (d <= 1).mean() >= within1 and (d <= 1).mean() >= 0.999, (d <= 1).mean()
    assert d.max() <= worst <= 3, d.max()


@pytest.mark.parametrize("which,name,within1,within2", [(0, "render_0100", 0.99, 0.997), (1, "render_1100", 0.98, 0.995),
                                                        (2, "tout_render_mpi_3d_diy_volume100", 0.995, 0.998)])
def test_older_goldens_under_the_default_offset(golden_dir, which, name, within1, within2):
    """with the default (meshEpsilon) offset the same three goldens are met to 98-99.7 % within 1/255 and >= 99.95 %
    within 4/255 -- far inside the reference's PNGCompare tolerance, which is why they could survive the change of
    convention unregenerated -- but only 64-92 % of their pixels are uint8-equal"""
    d = _golden_diff(golden_dir, which, name, MESH_EPSILON)
    assert (d <= 1).mean() >= within1, (d <= 1).mean()
    assert (d <= 2).mean() >= within2, (d <= 2).mean()
    assert (d <= 4).mean() >= 0.9995
    assert 0.6 < (d == 0).mean() < 0.93


def test_render_0100_residual_is_the_bounding_box_annotation(golden_dir):
    """nearly every pixel of the crop that is not uint8-equal lies on a screen diagonal through the image centre (the
    receding edges of the bounding box under the default camera) or on the columns x = 256 +- 99 (the vertical
    edges of its back face); all of them differ by one level, and only in the direction 'golden brighter' (white
    lines under a not quite opaque volume)"""
    g = np.load(os.path.join(golden_dir, "render_0100.npz"))
    sc = scenes.multi_render_scene(0)
    with first_sample(*OLDER_GENERATION):
        _, _, canvas = scenes.oracle_path_a(sc)
    mine = scenes.png_bytes(canvas, sc["W"], sc["H"])
    y0, y1, x0, x1 = [int(v) for v in g["rects"][0]]
    diff = g["rgb"][y0:y1, x0:x1].astype(int) - mine[y0:y1, x0:x1, :3].astype(int)
    ys, xs = np.nonzero(np.abs(diff).max(axis=2) > 0)
    assert 0 < ys.size < 300
    on_diagonal = np.abs(ys + y0 - 256) == np.abs(xs + x0 - 256)
    on_back_edge = np.abs(xs + x0 - 256) == 99
    # (the remaining ~30 form a 5-pixel staircase next to the top-left corner: a third line of the annotation)
    assert (on_diagonal | on_back_edge).mean() > 0.75 and on_diagonal.sum() > 20 and on_back_edge.sum() > 20
    assert diff.min() == 0 and diff.max() == 1


@pytest.mark.parametrize("which,name", GOLDENS)
def test_the_older_generations_offset_is_exactly_1e_4(golden_dir, which, name):
    """scanning the offset: the three pure-volume goldens peak at entry + 1e-4 (absolute) and degrade on either
    side of it; 1e-4 * |extent| (34.6x / 110.9x larger in these scenes) reproduces only 64-92 % of their pixels"""
    def exact(abs_offset, extent_rel):
        return float((_golden_diff(golden_dir, which, name, (abs_offset, extent_rel)) == 0).mean())

    best = exact(1e-4, 0.0)
    mesh_eps = exact(0.0, 1e-4)
    assert best >= 0.998 and mesh_eps < 0.93 and best - mesh_eps > 0.08, (best, mesh_eps)
    for other in (1e-5, 5e-4, 2e-3):
        assert exact(other, 0.0) <= best, other
    if which < 2:  # (the rectilinear scene is flat between 2e-5 and 2e-4; the two braid scenes are not)
        assert exact(2e-5, 0.0) < best and exact(2.5e-4, 0.0) < best


def _surface_free_pixels(blk, cam, W, H, bounds, margin=12):
    """pixels whose ray meets no iso-surface braid = 0: no sign change of the trilinear field along the ray (two
    passes of the sampler itself, 3000 samples, with tables that see only negative / only positive values), kept
    `margin` pixels away from any ray that has one (the reference's contour is a triangulation, not the trilinear
    zero set)"""
    from oracle import oracle as O
    fine = O.sample_distance(bounds, 3000)

    def seen(negative):
        lut = np.zeros((1024, 4), np.float32)
        lut[:, 0] = 1.0
        if negative:
            lut[:511, 3] = 1.0
        else:
            lut[512:, 3] = 1.0
        rgba, depth = O.new_canvas(W, H)
        O.render_to_canvas(blk, cam, W, H, lut, fine, -1.0, 1.0, rgba, depth)
        return rgba.reshape(H, W, 4)[..., 3] > 0

    surface = seen(True) & seen(False)
    acc = surface.copy()
    for dx in range(-margin, margin + 1):
        acc |= np.roll(surface, dx, 1)
    out = acc.copy()
    for dy in range(-margin, margin + 1):
        out |= np.roll(acc, dy, 0)
    return ~out


def test_newest_golden_decides_the_default_offset(golden_dir):
    """tout_render_3d_multi_default_runtime100.png (t_ascent_render_3d.cpp:1017-1122): an opaque contour under a
    volume plot (fixed range [-0.5, 0.5], "rainbow desaturated", alpha 0 -> 0.5, 1024^2, default camera).  Where no
    ray meets the contour the pixels are the volume plot's alone.  Those pixels -- found without the golden -- are
    reproduced to ~99 % uint8-equal under the default offset (the rest: annotation lines and surface fragments
    inside the margin), and on the pixels where the two generations' offsets give different bytes the golden sides
    with entry + 1e-4 * |extent| in all but a handful: this golden, the only one of the four that a change of VTK-m's
    contour filter forces to be regenerated, was rendered by the newer generation."""
    from ascent_b200 import datasets
    from oracle import oracle as O
    gold = np.load(os.path.join(golden_dir, "tout_render_3d_multi_default_runtime100.npz"))["rgb"].astype(int)
    H, W = gold.shape[:2]
    dom = datasets.braid_uniform(20)
    b = datasets.domain_bounds(dom)
    cam = O.camera_reset_to_bounds(b)
    lut = O.parse_color_table({"name": "rainbow desaturated", "control_points": [
        {"type": "alpha", "position": 0., "alpha": 0.}, {"type": "alpha", "position": 1., "alpha": .5}]}
    ).correct_opacity(100).lut()
    sd = O.sample_distance(b, 100)
    blk = scenes.oracle_block(dom)

    def frame(offset):
        with first_sample(*offset):
            rgba, depth = O.new_canvas(W, H)
            O.render_to_canvas(blk, cam, W, H, lut, sd, -0.5, 0.5, rgba, depth)
        u8, d8 = O.image_init(rgba, depth, 0)
        can, _ = O.image_to_canvas(u8, d8)
        return scenes.png_bytes(can, W, H)[..., :3].astype(int), u8.reshape(H, W, 4)[..., 3] > 0

    newer, covered = frame(MESH_EPSILON)
    older, _ = frame(OLDER_GENERATION)
    free = covered & _surface_free_pixels(blk, cam, W, H, b)
    assert free.sum() > 50000
    d_new = np.abs(newer - gold).max(axis=2)[free]
    d_old = np.abs(older - gold).max(axis=2)[free]
    assert (d_new == 0).mean() >= 0.985 and (d_new == 0).mean() > (d_old == 0).mean() + 0.008
    differ = free & (newer != older).any(axis=2)
    is_new = (newer == gold).all(axis=2) & differ
    is_old = (older == gold).all(axis=2) & differ
    assert differ.sum() > 500
    assert is_new.sum() >= 0.98 * differ.sum() and is_old.sum() <= 0.01 * differ.sum(), (differ.sum(), is_new.sum(), is_old.sum())


@pytest.mark.parametrize("which,name", GOLDENS)
def test_table_index_scale_is_pinned_by_the_goldens(golden_dir, which, name):
    """the structured sampler indexes the 1024-entry table with v * 1023; with v * 1024 -- the convention the
    reference's tracer of explicit cell sets turns out to use (tests/test_oracle_unstructured.py) -- the goldens of
    the structured path are missed on a large share of the pixels"""
    from oracle import oracle as O
    best = float((_golden_diff(golden_dir, which, name) == 0).mean())
    O.lib.orc_set_structured_index_extra(1)
    try:
        other = float((_golden_diff(golden_dir, which, name) == 0).mean())
    finally:
        O.lib.orc_set_structured_index_extra(0)
    assert best >= 0.998 and other < best - 0.05, (best, other)
