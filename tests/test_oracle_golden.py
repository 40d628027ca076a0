"""Pin the ray-cast oracle (oracle/raycast_oracle.c) against the reference's own golden images.

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


def _golden_diff(golden_dir, which, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sc = scenes.multi_render_scene(which) if which < 2 else scenes.mpi_volume_scene()
    _, _, canvas = scenes.oracle_path_a(sc)
    return crop_diffs(scenes.png_bytes(canvas, sc["W"], sc["H"]), g["rgb"], g["rects"])


GOLDENS = [(0, "render_0100"), (1, "render_1100"), (2, "tout_render_mpi_3d_diy_volume100")]


@pytest.mark.parametrize("which,name,exact,within1,worst", [(0, "render_0100", 0.998, 0.9999, 1),
                                                            (1, "render_1100", 1.0, 1.0, 0),
                                                            (2, "tout_render_mpi_3d_diy_volume100", 0.9995, 0.9995, 2)])
def test_goldens_pin_the_oracle_uint8_for_uint8(golden_dir, which, name, exact, within1, worst):
    """The reference's PNGCompare only asks for <= 0.1 % (1 %) of pixels off by more than 4/255.  The restated
    K0-K8 chain reproduces the reference's PNGs uint8 FOR uint8: every pixel of the render_1100 crop, 99.8 % of
    render_0100's and 99.95 % of the 2-rank rectilinear scene's, and -- the north star's own bar, here against the
    reference's images rather than against the oracle -- >= 99.9 % within 1/255 and none beyond 3/255.  What is
    left is not the volume: render_0100's off-by-one pixels sit exactly on the screen diagonals |x-256| == |y-256|
    (and two columns) where the golden shows the white bounding-box edges through rays whose opacity stops just short of 1 (green
    255 instead of 254), the mpi scene's 23 pixels (off by 2) are one 1-pixel-wide line of the same annotation
    crossing the top of the crops."""
    d = _golden_diff(golden_dir, which, name)
    assert (d == 0).mean() >= exact, (d == 0).mean()
    assert (d <= 1).mean() >= within1 and (d <= 1).mean() >= 0.999, (d <= 1).mean()
    assert d.max() <= worst <= 3, d.max()


def test_render_0100_residual_is_the_bounding_box_annotation(golden_dir):
    """nearly every pixel of the crop that is not uint8-equal lies on a screen diagonal through the image centre (the
    receding edges of the bounding box under the default camera) or on the columns x = 256 +- 99 (the vertical
    edges of its back face); all of them differ by one level, and only in the direction 'golden brighter' (white lines under
    a not quite opaque volume)"""
    g = np.load(os.path.join(golden_dir, "render_0100.npz"))
    sc = scenes.multi_render_scene(0)
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
def test_first_sample_offset_is_pinned_by_the_goldens(golden_dir, which, name):
    """The one convention of K4 the goldens decide: the first sample sits at entry + 1e-4, an ABSOLUTE offset.
    SURVEY appendix B9 recalled 1e-4 * |block extent| (34.6x / 110.9x larger in these scenes); scanning the offset
    through the oracle's test hook shows the recalled form reproduces only 64-92 % of the goldens' pixels exactly,
    against 99.8-100 % for the default, and that the match degrades on either side of 1e-4."""
    from oracle import oracle as O

    def exact(abs_offset, extent_rel):
        O.set_first_sample_offset(abs_offset, extent_rel)
        try:
            return float((_golden_diff(golden_dir, which, name) == 0).mean())
        finally:
            O.set_first_sample_offset()

    best = exact(1e-4, 0.0)
    recalled = exact(0.0, 1e-4)
    assert best >= 0.998 and recalled < 0.93 and best - recalled > 0.08, (best, recalled)
    for other in (1e-5, 5e-4, 2e-3):
        assert exact(other, 0.0) <= best, other
    if which < 2:  # (the rectilinear scene is flat between 2e-5 and 2e-4; the two braid scenes are not)
        assert exact(2e-5, 0.0) < best and exact(2.5e-4, 0.0) < best


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
