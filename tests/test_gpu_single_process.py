"""Deployment shape (i) of SURVEY 8(b): ONE process, one vr_ctx per rank, connected with
vr_comm_connect_local (no IPC) -- here all contexts sit on cuda:0, which exercises exactly the same kernels,
flags and rings as one context per GPU.  Covers path A per frame, the batch entry point
vr_comm_render_frames (device canvas and RGBA8 host frames) and the pushed/pulled layer exchange of path B,
each against the oracle (tests/scenes.py)."""
import os

import numpy as np
import pytest

from ascent_b200 import _lib, datasets
from oracle import oracle as O
import scenes

pytestmark = pytest.mark.gpu


def _contexts(n, max_pixels, max_partials=0):
    os.environ.setdefault("VR_COMM_TIMEOUT_MS", "30000")
    # all ranks share ONE device inside ONE process here: the exchange kernels spin on each other's flags, so
    # all of them must be resident at once -- small grids (one context per GPU needs no such cap)
    os.environ["VR_EXCHANGE_MAX_CTAS"] = "8"
    try:
        ctxs = [_lib.Context(0) for _ in range(n)]
    finally:
        del os.environ["VR_EXCHANGE_MAX_CTAS"]
    for r, c in enumerate(ctxs):
        c.comm_init(r, n, max_pixels, max_partials)
    _lib.comm_connect_local(ctxs)
    return ctxs


def _close(ctxs):
    for c in ctxs:
        c.close()


@pytest.fixture()
def scene4():
    """braid uniform split into 2 x 2 x 2 blocks; the four blocks of one z-layer, one per rank, camera off-axis"""
    doms = datasets.braid_uniform_blocks(12, 2, dtype=np.float32)[:4]
    bl = [datasets.domain_bounds(d) for d in doms]
    gb = datasets.union_bounds(bl)
    cam = O.camera_reset_to_bounds(gb)
    O.camera_azimuth(cam, 25.0)
    O.camera_elevation(cam, 15.0)
    lut = O.parse_color_table(scenes.RAMP_TF).correct_opacity(100).lut()
    rmin, rmax = scenes.field_range(doms)
    return dict(doms=doms, cam=cam, W=320, H=200, lut=lut, rmin=rmin, rmax=rmax,
                sample_dist=O.sample_distance(gb, 100), bounds=gb, dom_bounds=bl)


def test_path_a_one_process_four_contexts(scene4):
    sc = scene4
    W, H = sc["W"], sc["H"]
    n = len(sc["doms"])
    ctxs = _contexts(n, W * H)
    try:
        vis = np.ascontiguousarray(_lib.visibility_order(np.array(sc["dom_bounds"]), sc["cam"]), np.int32)
        for r, c in enumerate(ctxs):
            c.set_tf(sc["lut"])
            c.block_from_domain(0, sc["doms"][r])
        o_u8, o_d, o_can = scenes.oracle_path_a(sc)
        for push in (False, True):
            # every collective call is ISSUED on all contexts before any of them is synchronised
            for c in ctxs:
                c.trace_to_image(0, sc["cam"], W, H, sc["sample_dist"], sc["rmin"], sc["rmax"], no_clear=True,
                                 push=push)
            for c in ctxs:
                c.comm_composite_images_to_canvas(vis)
            can, cd = ctxs[0].canvas_download(W, H)
            for c in ctxs[1:]:
                c.synchronize()
            assert np.array_equal(can, o_can), "push=%s" % push
            u8, d = ctxs[0].image_result_download(W, H)
            assert np.array_equal(u8, o_u8)
            assert np.array_equal(d, o_d)
    finally:
        _close(ctxs)


def test_render_frames_batch_matches_per_frame(scene4):
    """vr_comm_render_frames: 5 cameras in one call per context; every frame's RGBA8 host image equals what the
    per-frame path + vr_canvas_download_rgba8 gives, and the canvas holds the last frame"""
    sc = scene4
    W, H = sc["W"], sc["H"]
    n = len(sc["doms"])
    ctxs = _contexts(n, W * H)
    try:
        cams = []
        for k in range(5):
            cam = O.camera_reset_to_bounds(sc["bounds"])
            O.camera_azimuth(cam, 10.0 + 35.0 * k)
            O.camera_elevation(cam, -20.0 + 12.0 * k)
            cams.append(cam)
        bounds = np.array(sc["dom_bounds"])
        vis = np.stack([np.ascontiguousarray(_lib.visibility_order(bounds, cam), np.int32) for cam in cams])
        bg = np.array([0.1, 0.2, 0.3, 1.0], np.float32)
        for r, c in enumerate(ctxs):
            c.set_tf(sc["lut"])
            c.block_from_domain(0, sc["doms"][r])
        # reference: one frame at a time
        want = []
        for k, cam in enumerate(cams):
            for c in ctxs:
                c.trace_to_image(0, cam, W, H, sc["sample_dist"], sc["rmin"], sc["rmax"], no_clear=True, push=True)
            for c in ctxs:
                c.comm_composite_images_to_canvas(vis[k])
            want.append(ctxs[0].canvas_download_rgba8(W, H, bg, flip=True))
            for c in ctxs[1:]:
                c.synchronize()
        last_can, last_cd = ctxs[0].canvas_download(W, H)
        # the batch: all frames of all contexts issued, then synchronised
        # (pinned: the per-frame copies must be asynchronous -- rank 0's call returns before its peers are issued)
        import torch
        out_t = torch.zeros((len(cams), H, W, 4), dtype=torch.uint8).pin_memory()
        out = out_t.numpy()
        for r, c in enumerate(ctxs):
            c.comm_render_frames(0, cams, W, H, sc["sample_dist"], sc["rmin"], sc["rmax"], vis, bg=bg,
                                 out_rgba8=out if r == 0 else None)
        for c in ctxs:
            c.synchronize()
        for k in range(len(cams)):
            assert np.array_equal(out[k].reshape(-1, 4), np.asarray(want[k]).reshape(-1, 4)), "frame %d" % k
        can, cd = ctxs[0].canvas_download(W, H)
        assert np.array_equal(can, last_can)
        assert np.array_equal(cd, last_cd, equal_nan=True)
    finally:
        _close(ctxs)


@pytest.mark.parametrize("push,multi", [("1", "16"), ("0", "16"), ("1", "2"), ("0", "2")])
def test_path_b_layers_two_contexts(push, multi):
    """four blocks per context (path B): ray layers pushed to / pulled by the tile owners, canvas bit-equal to
    the single-rank oracle of the same eight domains; multi = 2: the four blocks of a context go through ONE
    persistent launch over (block, tile) work items (trace_multi_kernel) instead of four launches"""
    os.environ["VR_LAYER_PUSH"] = push
    os.environ["VR_MULTI_MIN"] = multi
    try:
        doms = datasets.braid_uniform_blocks(10, 2, dtype=np.float32)
        bl = [datasets.domain_bounds(d) for d in doms]
        gb = datasets.union_bounds(bl)
        cam = O.camera_reset_to_bounds(gb)
        O.camera_azimuth(cam, -30.0)
        O.camera_elevation(cam, 20.0)
        lut = O.parse_color_table(scenes.RAMP_TF).correct_opacity(100).lut()
        rmin, rmax = scenes.field_range(doms)
        W, H = 300, 220
        sd = O.sample_distance(gb, 100)
        sc = dict(doms=doms, cam=cam, W=W, H=H, lut=lut, rmin=rmin, rmax=rmax, sample_dist=sd)
        _, o_rgba, o_depth = scenes.oracle_path_b(sc)
        ctxs = _contexts(2, W * H, W * H * 4 + 64)
        try:
            for r, c in enumerate(ctxs):
                c.set_tf(lut)
                for i in range(r, 8, 2):
                    c.block_from_domain(i, doms[i])
            for frame in range(3):  # (three frames: every buffer of the layer ring is used)
                for r, c in enumerate(ctxs):
                    c.layers_begin(W, H)
                    c.trace_blocks_to_layers(list(range(r, 8, 2)), cam, sd, rmin, rmax, False)
                for c in ctxs:
                    c.comm_layers_composite_to_canvas(cam)
                can, cd = ctxs[0].canvas_download(W, H)
                ctxs[1].synchronize()
                assert np.array_equal(can, o_rgba), "frame %d" % frame
                cov = o_rgba[:, 3] > 0
                assert np.array_equal(cd[cov], o_depth[cov]) and (cd[~cov] == np.float32(1.001)).all()
        finally:
            _close(ctxs)
    finally:
        del os.environ["VR_LAYER_PUSH"]
        del os.environ["VR_MULTI_MIN"]


@pytest.mark.parametrize("n,W,H", [(2, 200, 120), (3, 203, 77), (4, 256, 144), (6, 200, 120), (8, 320, 200)])
def test_zselect_resolves_ties_like_the_reference_radixk(n, W, H):
    """vr_comm_composite_zbuffer against oracle.radixk_zbuffer (pinned to the reference's own reduce_images + DIY,
    tests/test_oracle_radixk.py): depths drawn from a handful of values, so most pixels carry equal-depth fragments
    of several ranks -- the case where the reference's answer depends on the radix-k tree's visiting order -- plus
    fragments beyond the far plane (> 1), all-empty pixels, and a width that is not a multiple of 4."""
    ctxs = _contexts(n, W * H)
    try:
        for rep in range(2):
            g = np.random.default_rng(17 * n + rep)
            cols = g.random((n, H * W, 4), dtype=np.float32)
            deps = g.choice(np.array([0.2, 0.4, 0.4, 0.6, 0.8, 1.0, 1.001, 1.5], np.float32), (n, H * W))
            deps[:, :W] = np.float32(1.25)  # a row where NO rank has a fragment in (.., 1]: the piece owner's pixel stays
            for r, c in enumerate(ctxs):
                c.canvas_upload(W, H, cols[r], deps[r])
                c.image_from_canvas()
            for c in ctxs:
                c.comm_composite_zbuffer()
            u8, d = ctxs[0].image_result_download(W, H)
            for c in ctxs[1:]:
                c.synchronize()
            qs = [O.image_init(cols[r], deps[r], 0) for r in range(n)]
            want, wd = O.radixk_zbuffer(np.stack([q[0] for q in qs]), np.stack([q[1] for q in qs]), W, H)
            assert np.array_equal(d, wd), "depth (rep %d): %d pixels differ" % (rep, (d != wd).sum())
            assert np.array_equal(u8, want), "colour (rep %d): %d pixels differ" % (rep, (u8 != want).any(axis=1).sum())
            # and the rank-order fold the kernel used to do is a DIFFERENT image on this input
            if n > 1:
                front, fd = qs[0][0].copy(), qs[0][1].copy()
                for r in range(1, n):
                    O.zbuffer_composite(front, fd, qs[r][0], qs[r][1], gl_depth=True)
                assert not np.array_equal(front, want)
    finally:
        _close(ctxs)
