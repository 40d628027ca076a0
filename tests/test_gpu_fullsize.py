"""Parity at the sizes BASELINE.json names (c2, c3, c4, c5) and on a block of more than 2^31 points
(the 64-bit index instantiation of the sampler), through the C ABI against the CPU oracle.

Fields are generated on the GPU (vr_synth_braid_dev) and downloaded, so that both sides sample the
very same values; camera, transfer function and sample distance for the oracle come from oracle/.
Same tolerance as tests/test_gpu_parity.py: >= 99.9 % of pixels within 1/255, all within 3/255,
PSNR >= 50 dB -- and the stronger bit-identity that holds in practice."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (workload definitions shared with the bench: same scenes, same code path)
from ascent_b200 import _lib  # noqa: E402
from oracle import oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = _lib.Context(0)
    yield c
    c.close()


def synth_fields(ctx, wl):
    """the workload's blocks, generated on the GPU; returns (device tensors, host copies)"""
    import torch
    dev, host = [], []
    for b in bench.block_layout(wl):
        t = torch.empty(int(np.prod(b["dims"])), dtype=torch.float32, device="cuda")
        ctx.synth_braid_dev(t.data_ptr(), _lib.VR_F32, b["dims"], b["start"], b["glob"])
        ctx.synchronize()
        dev.append(t)
        host.append(t.cpu().numpy())
    return dev, host


def publish(ctx, wl, dev):
    axes = [bench.warp_axis(n) for n in bench.block_layout(wl)[0]["dims"]] if wl.get("rectilinear") else None
    for i, b in enumerate(bench.block_layout(wl)):
        if axes:
            ctx.block_rectilinear(i, b["dims"], axes, None, device_ptr=dev[i].data_ptr(), dtype=_lib.VR_F32)
        else:
            ctx.block_uniform(i, b["dims"], b["origin"], b["spacing"], None, device_ptr=dev[i].data_ptr(),
                              dtype=_lib.VR_F32)


def assert_parity(p, bit_share=0.9999):
    assert p["within1"] >= 0.999 and p["max"] <= 3.0 + 1e-6 and p["psnr"] >= 50.0, p
    assert p["bit_identical_pixels"] >= bit_share, p
    assert p["covered_pixels"] > 1000


def free_blocks(ctx, n):
    for i in range(n):
        ctx.block_free(i)


def test_c2_512cubed_1080p(ctx):
    """BASELINE config 2 as named: braid uniform 512^3 f32, 1920x1080, default camera, S = 100 --
    the frame bench.py times (fused clear + trace + Image::Init + ImageToCanvas)."""
    wl = bench.workload_c2()
    dev, host = synth_fields(ctx, wl)
    rng = (float(host[0].min()), float(host[0].max()))
    sp = bench.scene_params(wl, rng)
    sc = bench.OracleScene(wl, host, 1, rng=rng)
    assert bytes(sp["cam"]) == bytes(sc.cam) and sp["lut"].tobytes() == sc.lut.tobytes()
    assert sp["sample_dist"] == sc.sample_dist
    ctx.set_tf(sp["lut"])
    publish(ctx, wl, dev)
    W, H = wl["W"], wl["H"]
    ctx.trace_to_image(0, sp["cam"], W, H, sp["sample_dist"], *rng, write_canvas=True)
    g_rgba, g_depth = ctx.canvas_download(W, H)
    u8, _ = ctx.image_download(W, H)
    o_rgba, o_depth = sc.frame(0)
    p = bench.compare_canvas(g_rgba, g_depth, o_rgba, o_depth)
    assert_parity(p)
    assert p["bit_exact"], p  # path A returns k/255 of the uint8 image: every pixel identical
    assert np.array_equal(u8.astype(np.float32) * np.float32(1.0 / 255.0), o_rgba)
    # the dense operating point (S = 887, one sample per voxel) on the unfused float canvas
    bench.SAMPLES = 887
    try:
        sp = bench.scene_params(wl, rng)
        sc = bench.OracleScene(wl, host, 1, rng=rng)
    finally:
        bench.SAMPLES = 100
    ctx.set_tf(sp["lut"])
    ctx.canvas_clear(W, H)
    ctx.trace_to_canvas(0, sp["cam"], sp["sample_dist"], *rng, False)
    g_rgba, g_depth = ctx.canvas_download(W, H)
    o_rgba, o_depth = O.new_canvas(W, H)
    O.render_to_canvas(sc.obs[0], sc.cam, W, H, sc.lut, sc.sample_dist, *rng, o_rgba, o_depth)
    assert_parity(bench.compare_canvas(g_rgba, g_depth, o_rgba, o_depth))
    free_blocks(ctx, 1)


def test_c4_rectilinear_768cubed_cinema_views(ctx):
    """BASELINE config 4 as named: rectilinear 768^3 with warped axes, 1024^2, four of the 64 views of
    the phi = 8 x theta = 8 cinema orbit (front, oblique, from below, grazing)."""
    wl = bench.workload_c4()
    dev, host = synth_fields(ctx, wl)
    rng = (float(host[0].min()), float(host[0].max()))
    sp = bench.scene_params(wl, rng)
    sc = bench.OracleScene(wl, host, 1, rng=rng)
    ctx.set_tf(sp["lut"])
    publish(ctx, wl, dev)
    W, H = wl["W"], wl["H"]
    for v in (0, 9, 28, 63):
        assert bytes(sp["cams"][v]) == bytes(sc.cams[v])
        ctx.trace_to_image(0, sp["cams"][v], W, H, sp["sample_dist"], *rng, write_canvas=True)
        g_rgba, g_depth = ctx.canvas_download(W, H)
        o_rgba, o_depth = sc.frame(v)
        p = bench.compare_canvas(g_rgba, g_depth, o_rgba, o_depth)
        assert_parity(p)
        assert p["bit_exact"], (v, p)
    free_blocks(ctx, 1)


def test_c3_eight_blocks_4k_path_b_on_one_gpu(ctx):
    """BASELINE config 3 at N = 1: 8 blocks of 512^3, 3840x2160, path B (ray layers + fold)."""
    wl = bench.workload_c3()
    dev, host = synth_fields(ctx, wl)
    rng = (min(float(h.min()) for h in host), max(float(h.max()) for h in host))
    sp = bench.scene_params(wl, rng)
    sc = bench.OracleScene(wl, host, 1, rng=rng)
    ctx.set_tf(sp["lut"])
    publish(ctx, wl, dev)
    W, H = wl["W"], wl["H"]
    ctx.layers_begin(W, H)
    ctx.trace_blocks_to_layers(list(range(8)), sp["cam"], sp["sample_dist"], *rng, False)
    ctx.layers_composite_to_canvas(sp["cam"], canvas_is_clear=True)
    g_rgba, g_depth = ctx.canvas_download(W, H)
    o_rgba, o_depth = sc.frame(0)
    assert_parity(bench.compare_canvas(g_rgba, g_depth, o_rgba, o_depth))
    # and path A's building block at the same size: each block's fused uint8 image, folded in
    # visibility order on this one GPU (what 8 ranks exchange), against the oracle's 8-rank frame
    import torch
    n = W * H
    layers_c = torch.empty(8 * n, dtype=torch.int32, device="cuda")
    layers_d = torch.empty(8 * n, dtype=torch.float32, device="cuda")
    for i in range(8):
        ctx.trace_to_image(i, sp["cam"], W, H, sp["sample_dist"], *rng)
        c_ptr, d_ptr = ctx.image_ptrs()
        ctx.synchronize()
        layers_c[i * n:(i + 1) * n].copy_(torch.as_tensor(_DevBuf(c_ptr, n, "<i4"), device="cuda"))
        layers_d[i * n:(i + 1) * n].copy_(torch.as_tensor(_DevBuf(d_ptr, n, "<f4"), device="cuda"))
    torch.cuda.synchronize()
    order = _lib.visibility_order(np.array(sp["bounds"]), sp["cam"])
    out_c = torch.empty(n, dtype=torch.int32, device="cuda")
    out_d = torch.empty(n, dtype=torch.float32, device="cuda")
    ctx.fold_images_dev(layers_c.data_ptr(), layers_d.data_ptr(), n, order, n, out_c.data_ptr(), out_d.data_ptr())
    ctx.image_to_canvas_dev(out_c.data_ptr(), out_d.data_ptr())
    g_rgba, g_depth = ctx.canvas_download(W, H)
    sc8 = bench.OracleScene(wl, host, 8, rng=rng)
    o_rgba, o_depth = sc8.frame(0)
    p = bench.compare_canvas(g_rgba, g_depth, o_rgba, o_depth)
    assert p["bit_exact"], p
    free_blocks(ctx, 8)


class _DevBuf:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def test_c5_512_small_blocks_4096sq(ctx):
    """BASELINE config 5 at N = 1: 512 blocks of 128^3, 4096^2, path B on ray layers."""
    wl = bench.workload_c5()
    dev, host = synth_fields(ctx, wl)
    rng = (min(float(h.min()) for h in host), max(float(h.max()) for h in host))
    sp = bench.scene_params(wl, rng)
    sc = bench.OracleScene(wl, host, 1, rng=rng)
    ctx.set_tf(sp["lut"])
    publish(ctx, wl, dev)
    W, H = wl["W"], wl["H"]
    ctx.layers_begin(W, H)
    ctx.trace_blocks_to_layers(list(range(512)), sp["cam"], sp["sample_dist"], *rng, False)
    ctx.layers_composite_to_canvas(sp["cam"], canvas_is_clear=True)
    g_rgba, g_depth = ctx.canvas_download(W, H)
    o_rgba, o_depth = sc.frame(0)
    assert_parity(bench.compare_canvas(g_rgba, g_depth, o_rgba, o_depth))
    free_blocks(ctx, 512)


def test_block_with_more_than_2_pow_31_points(ctx):
    """A 1291 x 1290 x 1290 uniform block (2.148e9 points > 2^31, 8.6 GB): element indices no longer fit
    32 bits, so the sampler's `long long` instantiation runs.  Oblique camera so that rays cross the
    high-index end of the block; path A float canvas and the partial list against the oracle."""
    import torch
    dims = (1291, 1290, 1290)
    n = int(np.prod(dims))
    assert n > 2 ** 31
    t = torch.empty(n, dtype=torch.float32, device="cuda")
    ctx.synth_braid_dev(t.data_ptr(), _lib.VR_F32, dims, (0, 0, 0), dims)
    ctx.synchronize()
    host = t.cpu().numpy()
    # the far corner really holds field values (a 32-bit index would have wrapped long before)
    want = O.braid_values(4, 1, 1, dims[0] - 4, dims[1] - 1, dims[2] - 1, *dims, dtype=np.float32)
    assert np.allclose(host[-4:], want, atol=2e-5)
    spacing = [20.0 / (d - 1) for d in dims]
    ctx.block_uniform(0, dims, [-10.0] * 3, spacing, None, device_ptr=t.data_ptr(), dtype=_lib.VR_F32)
    ob = O.OracleBlock(dims, host, origin=[-10.0] * 3, spacing=spacing)
    bounds = ob.bounds()
    cam = O.camera_reset_to_bounds(bounds)
    O.camera_azimuth(cam, 155.0)
    O.camera_elevation(cam, -35.0)
    lut = O.parse_color_table(bench.RAMP_TF).correct_opacity(100).lut()
    sd = O.sample_distance(bounds, 100)
    rng = (float(host[::997].min()), float(host[::997].max()))
    W, H = 640, 480
    ctx.set_tf(lut)
    ctx.canvas_clear(W, H)
    ctx.trace_to_canvas(0, cam, sd, *rng, False)
    g_rgba, g_depth = ctx.canvas_download(W, H)
    o_rgba, o_depth = O.new_canvas(W, H)
    O.render_to_canvas(ob, cam, W, H, lut, sd, *rng, o_rgba, o_depth)
    assert_parity(bench.compare_canvas(g_rgba, g_depth, o_rgba, o_depth))
    gp = ctx.render_partials(0, cam, W, H, sd, *rng)
    op = O.render_partials(ob, cam, W, H, lut, sd, *rng, None)
    gp = np.sort(gp, order=["pixel_id", "depth"])
    op = np.sort(op, order=["pixel_id", "depth"])
    assert gp.size == op.size and np.array_equal(gp["pixel_id"], op["pixel_id"])
    assert np.array_equal(gp["depth"], op["depth"])
    assert (gp["alpha"].view(np.uint32) == op["alpha"].view(np.uint32)).mean() >= 0.9999
    ctx.block_free(0)
