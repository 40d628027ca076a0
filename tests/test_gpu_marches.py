"""The sampler's three marches (general / sparse / brick, sampler.cu) must produce the same bits: every
scene is rendered by two contexts, one restricted to the general march, one free to pick any of the three
(VR_NO_SPARSE / VR_BRICK are read at vr_create), and the float canvases are compared bit for bit -- plus the
oracle."""
import os

import numpy as np
import pytest

import scenes
from ascent_b200 import _lib, color_table, datasets
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctxs():
    os.environ["VR_NO_SPARSE"] = "1"
    os.environ["VR_BRICK"] = "0"
    g = _lib.Context(0)
    os.environ["VR_NO_SPARSE"] = "0"
    os.environ["VR_BRICK"] = "1"   # the brick march is opt-in (slower than the general march on B200)
    p = _lib.Context(0)
    del os.environ["VR_BRICK"], os.environ["VR_NO_SPARSE"]
    yield g, p
    g.close()
    p.close()


def render(ctx, dom, cam, W, H, lut, sd, rmin, rmax, depth=None):
    ctx.set_tf(lut)
    ctx.block_from_domain(0, dom)
    if depth is None:
        ctx.canvas_clear(W, H)
    else:
        ctx.canvas_upload(W, H, np.zeros((H * W, 4), np.float32), depth)
    ctx.trace_to_canvas(0, cam, sd, rmin, rmax, depth is not None)
    out = ctx.canvas_download(W, H)
    ctx.partials_begin(W, H)
    ctx.trace_to_partials(0, cam, sd, rmin, rmax, False)
    parts = np.sort(ctx.partials_download(), order=["pixel_id", "depth"])
    ctx.block_free(0)
    return out, parts


@pytest.mark.parametrize("n,samples,dtype", [(48, 100, np.float32), (48, 400, np.float32), (64, 20, np.float32),
                                             (40, 887, np.float32), (33, 100, np.float32), (48, 100, np.float64),
                                             (64, 8, np.float64), (20, 1000, np.float32)])
@pytest.mark.parametrize("view", [(0., 0., 0.), (35., 20., 0.), (-120., -40., 0.5), (90., 0., 0.), (0., 89., 0.)])
def test_marches_agree_bit_for_bit(ctxs, n, samples, dtype, view):
    """steps from 0.05 to 8 voxels: dense ones take the brick march (f32, row length % 4 == 0), sparse ones
    (>= 2.6 voxels) the sparse march, the rest -- and f64 / odd row lengths -- the general march"""
    g, p = ctxs
    dom = datasets.braid_uniform(n, dtype=dtype)
    b = datasets.domain_bounds(dom)
    cam = O.camera_reset_to_bounds(b)
    O.camera_azimuth(cam, view[0])
    O.camera_elevation(cam, view[1])
    O.camera_zoom(cam, view[2])
    W, H = 320, 200
    lut = color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(samples).lut()
    sd = O.sample_distance(b, samples)
    rmin, rmax = scenes.field_range([dom])
    (ga, gd), gp = render(g, dom, cam, W, H, lut, sd, rmin, rmax)
    (pa, pd), pp = render(p, dom, cam, W, H, lut, sd, rmin, rmax)
    assert (ga[:, 3] > 0).sum() > 500
    assert np.array_equal(ga.view(np.uint32), pa.view(np.uint32)), "canvas colour differs between the marches"
    cov = ga[:, 3] > 0
    assert np.array_equal(gd[cov], pd[cov])
    assert gp.tobytes() == pp.tobytes(), "partials differ between the marches"
    o_rgba, o_depth = O.new_canvas(W, H)
    O.render_to_canvas(scenes.oracle_block(dom), cam, W, H, lut, sd, rmin, rmax, o_rgba, o_depth, use_depth=False)
    assert (pa.view(np.uint32) == o_rgba.view(np.uint32)).all(axis=1).mean() >= 0.9999
    # an opaque surface in the middle of the volume (canvas depth clamp): rays end early inside bricks
    depth = np.full(H * W, 0.9, np.float32)
    depth[::3] = 1.001
    (ga, gd), _ = render(g, dom, cam, W, H, lut, sd, rmin, rmax, depth)
    (pa, pd), _ = render(p, dom, cam, W, H, lut, sd, rmin, rmax, depth)
    assert np.array_equal(ga.view(np.uint32), pa.view(np.uint32))
