"""N4 through the C ABI: vr_block_unstructured + vr_trace_to_partials against the unstructured oracle
(oracle/raycast_oracle.c "N4", pinned to the reference's ghost-field golden in tests/test_oracle_unstructured.py).
Partials must come out bit-identical: same rays, same sample positions, same Newton steps, same blend."""
import os

import numpy as np
import pytest

from ascent_b200 import _lib, color_table, datasets
from oracle import oracle as O
import scenes

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ctx():
    c = _lib.Context(0)
    yield c
    c.close()


def _sorted(p):
    return np.sort(p, order=["pixel_id", "depth"])


def _warped_mesh(n, seed, dtype, hollow=False):
    """braid on an n^3 grid, written as hexahedra whose interior points are pushed off the lattice (general, non
    axis-aligned hexahedra); the field is sampled at the lattice positions (any values do).  hollow: a block of cells
    is missing from the middle (a cavity: rays enter, leave, enter again) and three from a corner (a notch)"""
    dom = datasets.braid_uniform(n, dtype=dtype)
    drop = ()
    if hollow:
        drop = [((k * (n - 1) + j) * (n - 1) + i) for k in range(4, 6) for j in range(3, 6) for i in range(2, 6)] + [0, 1, n - 1]
    pts, conn = datasets.structured_to_hexes(dom["dims"], dom["origin"], dom["spacing"], drop_cells=drop)
    g = np.random.default_rng(seed)
    idx = np.arange(pts.shape[0])
    i, j, k = idx % n, (idx // n) % n, idx // (n * n)
    inner = (i > 0) & (i < n - 1) & (j > 0) & (j < n - 1) & (k > 0) & (k < n - 1)
    h = float(dom["spacing"][0])
    pts = pts.copy()
    pts[inner] += (g.random((int(inner.sum()), 3), dtype=np.float32) - 0.5) * np.float32(0.35 * h)
    return dom, pts, conn


def _same(got, want, what):
    """bit equality with a diagnosis instead of a bare False"""
    if got.size != want.size:
        common = np.intersect1d(got["pixel_id"], want["pixel_id"]).size
        raise AssertionError("%s: %d partials, oracle %d (%d pixels in common)" % (what, got.size, want.size, common))
    for f in ("pixel_id", "depth", "rgb", "alpha"):
        a, b = got[f], want[f]
        if not np.array_equal(a, b):
            bad = np.nonzero((a != b).reshape(a.shape[0], -1).any(axis=1))[0]
            raise AssertionError("%s: %s differs in %d of %d partials; first: pixel %d got %s want %s (alpha %s / %s)" % (
                what, f, bad.size, a.shape[0], int(want["pixel_id"][bad[0]]), a[bad[0]], b[bad[0]],
                got["alpha"][bad[0]], want["alpha"][bad[0]]))


@pytest.mark.parametrize("shape,dtype,assoc,az,hollow", [("hex", np.float32, "point", 25.0, False),
                                                          ("hex", np.float64, "cell", -50.0, False),
                                                          ("tet", np.float32, "point", 130.0, False),
                                                          ("tet", np.float64, "cell", 10.0, False),
                                                          ("hex", np.float32, "point", 200.0, True),
                                                          ("tet", np.float32, "point", -20.0, True),
                                                          ("hex", np.float64, "cell", 0.0, True)])
def test_partials_bit_exact_against_the_oracle(ctx, shape, dtype, assoc, az, hollow):
    dom, pts, conn = _warped_mesh(11, 3, dtype, hollow)
    if shape == "tet":
        conn = datasets.hexes_to_tets(conn)
    g = np.random.default_rng(5)
    field = dom["field"].reshape(-1) if assoc == "point" else g.random(conn.shape[0]).astype(dtype) * 4 - 2
    b = [pts[:, 0].min(), pts[:, 0].max(), pts[:, 1].min(), pts[:, 1].max(), pts[:, 2].min(), pts[:, 2].max()]
    W, H = 320, 240
    cam = O.camera_reset_to_bounds(b)
    O.camera_azimuth(cam, az)
    O.camera_elevation(cam, 20.0)
    lut = color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut()
    sd = O.sample_distance(b, 100)
    rmin, rmax = float(field.min()), float(field.max())
    um = O.OracleUMesh(pts, conn, field, cell_assoc=assoc == "cell")
    want = _sorted(O.render_umesh_partials(um, cam, W, H, lut, sd, rmin, rmax))
    ctx.set_tf(lut)
    ctx.block_unstructured(0, pts, conn.astype(np.int64) if shape == "tet" else conn, field,
                           assoc=_lib.VR_CELL if assoc == "cell" else _lib.VR_POINT)
    assert np.allclose(ctx.block_bounds(0), um.bounds())
    got = _sorted(ctx.render_partials(0, cam, W, H, sd, rmin, rmax, None))
    assert want.size > 5000
    _same(got, want, "%s %s hollow=%s" % (shape, assoc, hollow))
    # the other entry points refuse an unstructured block, as the reference never renders one to a canvas directly
    ctx.canvas_clear(W, H)
    with pytest.raises(_lib.VRError):
        ctx.trace_to_canvas(0, cam, sd, rmin, rmax, False)
    ctx.block_free(0)


def test_ghost_field_golden_through_the_abi(ctx, golden_dir):
    """t_ascent_multi_topo.cpp:181-252 end to end on the GPU: unstructured partials -> PartialCompositor ->
    partials_to_canvas -> background + uint8, against the oracle (bit-exact canvas) and against the reference's own
    PNG at the reference's tolerance and at what the restatement achieves (99.9 % of the pixels uint8-equal)"""
    sc = scenes.ghost_volume_scene()
    W, H = sc["W"], sc["H"]
    _, o_rgba, o_depth = scenes.oracle_unstructured_path_b(sc)
    ctx.set_tf(sc["lut"])
    ctx.block_unstructured(0, sc["points"], sc["conn"], sc["field"])
    ctx.canvas_clear(W, H)
    ctx.partials_begin(W, H)
    ctx.trace_to_partials(0, sc["cam"], sc["sample_dist"], sc["rmin"], sc["rmax"], True)
    ctx.partials_composite()
    ctx.partials_to_canvas(sc["cam"])
    rgba, depth = ctx.canvas_download(W, H)
    assert np.array_equal(rgba, o_rgba)
    cov = o_rgba[:, 3] > 0
    assert np.array_equal(depth[cov], o_depth[cov])
    g = np.load(os.path.join(golden_dir, "tout_multi_topo_single_ghost_vol_render100.npz"))["rgb"].astype(int)
    mine = np.asarray(ctx.canvas_download_rgba8(W, H, (0., 0., 0., 1.), flip=False)).reshape(H, W, 4)[..., :3].astype(int)
    d = np.abs(mine - g).max(axis=2)
    assert (d > 4).mean() <= 0.0005 and (d == 0).mean() >= 0.999
    ctx.block_free(0)


def test_unstructured_and_structured_domains_in_one_frame(ctx):
    """m_has_unstructured (VolumeRenderer.cpp:874-903): one unstructured domain makes EVERY domain render as
    partials; the list compositor folds them together by depth"""
    doms = datasets.braid_uniform_blocks(10, 2, dtype=np.float32)[:2]  # two neighbouring blocks
    bl = [datasets.domain_bounds(d) for d in doms]
    gb = datasets.union_bounds(bl)
    W, H = 300, 220
    cam = O.camera_reset_to_bounds(gb)
    O.camera_azimuth(cam, 40.0)
    O.camera_elevation(cam, -15.0)
    lut = color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut()
    sd = O.sample_distance(gb, 100)
    rmin, rmax = scenes.field_range(doms)
    pts, conn = datasets.structured_to_hexes(doms[1]["dims"], doms[1]["origin"], doms[1]["spacing"])
    field1 = doms[1]["field"].reshape(-1)
    # oracle
    rgba, depth = O.new_canvas(W, H)
    p0 = O.render_partials(scenes.oracle_block(doms[0]), cam, W, H, lut, sd, rmin, rmax, depth)
    p1 = O.render_umesh_partials(O.OracleUMesh(pts, conn, field1), cam, W, H, lut, sd, rmin, rmax, depth)
    res = O.composite_partials([p0, p1])
    O.partials_to_canvas(res, cam, W, H, rgba, depth)
    # GPU
    ctx.set_tf(lut)
    ctx.block_from_domain(0, doms[0])
    ctx.block_unstructured(1, pts, conn, field1)
    ctx.canvas_clear(W, H)
    ctx.partials_begin(W, H)
    ctx.trace_to_partials(0, cam, sd, rmin, rmax, True)
    ctx.trace_to_partials(1, cam, sd, rmin, rmax, True)
    ctx.partials_composite()
    mine = _sorted(ctx.partials_download())
    ctx.partials_to_canvas(cam)
    g_rgba, g_depth = ctx.canvas_download(W, H)
    want = _sorted(res)
    assert np.array_equal(mine["pixel_id"], want["pixel_id"]) and np.array_equal(mine["depth"], want["depth"])
    assert np.abs(mine["rgb"] - want["rgb"]).max() <= 1e-6 and np.abs(mine["alpha"] - want["alpha"]).max() <= 1e-6
    d = np.abs(g_rgba - rgba).max(axis=1)
    assert (d <= 1 / 255).mean() >= 0.999 and d.max() <= 3 / 255
    ctx.block_free(0)
    ctx.block_free(1)


def test_bad_meshes_are_rejected(ctx):
    pts = np.zeros((8, 3), np.float32)
    with pytest.raises(_lib.VRError):
        ctx.block_unstructured(0, pts, np.full((1, 8), 99, np.int32), np.zeros(8, np.float32))  # index out of range


def test_partials_of_another_producer_join_the_frame(ctx):
    """vr_partials_append (the Devil Ray consumer seam, dray/rendering/renderer.cpp:309-331): a list produced elsewhere
    -- here by the oracle, for the second of two neighbouring blocks -- is composited together with what
    vr_trace_to_partials emitted for the first one, exactly like two traced blocks"""
    doms = datasets.braid_uniform_blocks(10, 2, dtype=np.float32)[:2]
    gb = datasets.union_bounds([datasets.domain_bounds(d) for d in doms])
    W, H = 260, 200
    cam = O.camera_reset_to_bounds(gb)
    O.camera_azimuth(cam, 35.0)
    lut = color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(100).lut()
    sd = O.sample_distance(gb, 100)
    rmin, rmax = scenes.field_range(doms)
    _, depth = O.new_canvas(W, H)
    p0 = O.render_partials(scenes.oracle_block(doms[0]), cam, W, H, lut, sd, rmin, rmax, depth)
    p1 = O.render_partials(scenes.oracle_block(doms[1]), cam, W, H, lut, sd, rmin, rmax, depth)
    want = _sorted(O.composite_partials([p0, p1]))
    ctx.set_tf(lut)
    ctx.block_from_domain(0, doms[0])
    ctx.canvas_clear(W, H)
    ctx.partials_begin(W, H)
    ctx.partials_append(p1[:p1.size // 2])
    ctx.trace_to_partials(0, cam, sd, rmin, rmax, True)
    ctx.partials_append(p1[p1.size // 2:])
    ctx.partials_composite()
    got = _sorted(ctx.partials_download())
    assert got.size == want.size > 1000
    assert np.array_equal(got["pixel_id"], want["pixel_id"]) and np.array_equal(got["depth"], want["depth"])
    assert np.abs(got["rgb"] - want["rgb"]).max() <= 1e-6 and np.abs(got["alpha"] - want["alpha"]).max() <= 1e-6
    bad = p1[:4].copy()
    bad["pixel_id"][0] = W * H
    with pytest.raises(_lib.VRError):
        ctx.partials_append(bad)
    ctx.block_free(0)
