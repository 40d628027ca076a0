"""N4: the unstructured producer's oracle (oracle/raycast_oracle.c, "N4: unstructured cells").

VTK-m's ConnectivityTracer is absent from /root/reference, so the restatement is pinned from two sides:
  * against the reference's golden of the unstructured path, tout_multi_topo_single_ghost_vol_render100.png
    (annotation-free, whole 1024^2 frame: a box of hexahedra with a two-cell notch), which the restatement
    reproduces uint8 for uint8 on 99.9 % of the frame -- and which is what decides the three conventions that could
    not be known otherwise: the ray is sampled stretch by stretch inside the mesh (entering through an external
    face, leaving through one, re-entering behind a concavity), every stretch starts at entry + (entry mod sample
    distance), and the transfer-function table is indexed with v * 1024 (clamped), not v * 1023;
  * against the structured sampler (itself pinned to three goldens): on a structured mesh written out as
    hexahedra, with the structured sampler's conventions switched in, the two agree to rounding."""
import os

import numpy as np

from ascent_b200 import datasets
from oracle import oracle as O
import scenes


def _ghost_frame(sc, **kw):
    W, H = sc["W"], sc["H"]
    um = O.OracleUMesh(sc["points"], sc["conn"], sc["field"])
    rays, _ = O.trace_umesh(um, sc["cam"], W, H, sc["lut"], sc["sample_dist"], sc["rmin"], sc["rmax"], keep=False, **kw)
    rgba, depth = O.new_canvas(W, H)
    O.lib.orc_write_to_canvas(O.C.byref(rays), O.C.byref(sc["cam"]), W, H, O._ptr(rgba, O.C.c_float),
                              O._ptr(depth, O.C.c_float))
    O.rays_free(rays)
    return scenes.png_bytes(rgba, W, H)[..., :3].astype(int)


def test_ghost_volume_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "tout_multi_topo_single_ghost_vol_render100.npz"))["rgb"].astype(int)
    sc = scenes.ghost_volume_scene()
    _, rgba, _ = scenes.oracle_unstructured_path_b(sc)
    mine = scenes.png_bytes(rgba, sc["W"], sc["H"])[..., :3].astype(int)
    d = np.abs(mine - g).max(axis=2)
    assert (d > 4).mean() <= 0.02            # the reference test's own criterion
    assert (d > 4).mean() <= 0.0005          # what the restatement achieves: 0.036 %
    assert (d == 0).mean() >= 0.999          # 99.94 % of ALL pixels uint8-equal
    covered = g.sum(axis=2) > 0
    assert (d[covered] == 0).mean() >= 0.998 and (d[covered] <= 1).mean() >= 0.999
    # what is left: two strips 2-4 pixels wide where rays clip a corner of the notch and re-enter the mesh less than
    # one sample distance behind their exit (the reference's cell walk loses the re-entered cell's samples there,
    # DESIGN.md 4.5) and ~200 isolated pixels one level off
    ys, xs = np.nonzero(d > 1)
    assert ys.size < 600 and ((xs >= 700) & (ys <= 260)).mean() > 0.98


def test_conventions_the_golden_fixes(golden_dir):
    """with the structured sampler's conventions (first sample at bounds entry + 1e-4, table index v * 1023, no
    stretches) the same scene misses the golden on most pixels"""
    g = np.load(os.path.join(golden_dir, "tout_multi_topo_single_ghost_vol_render100.npz"))["rgb"].astype(int)
    sc = scenes.ghost_volume_scene()
    d = np.abs(_ghost_frame(sc, structured_conventions=True) - g).max(axis=2)
    covered = g.sum(axis=2) > 0
    assert (d > 4).mean() > 0.08 and (d[covered] == 0).mean() < 0.5


def test_hexahedra_of_a_structured_grid_degenerate_to_the_structured_sampler():
    dom = datasets.braid_uniform(12, dtype=np.float32)
    b = datasets.domain_bounds(dom)
    W, H = 256, 192
    cam = O.camera_reset_to_bounds(b)
    O.camera_azimuth(cam, 25.0)
    O.camera_elevation(cam, 10.0)
    lut = O.parse_color_table(scenes.RAMP_TF).correct_opacity(100).lut()
    sd = O.sample_distance(b, 100)
    rmin, rmax = scenes.field_range([dom])
    r0, t0 = O.trace_block(scenes.oracle_block(dom), cam, W, H, lut, sd, rmin, rmax)
    O.rays_free(r0)
    pts, conn = datasets.structured_to_hexes(dom["dims"], dom["origin"], dom["spacing"])
    um = O.OracleUMesh(pts, conn, dom["field"].reshape(-1))
    r1, t1 = O.trace_umesh(um, cam, W, H, lut, sd, rmin, rmax, structured_conventions=True)
    O.rays_free(r1)
    assert t0.subset == t1.subset and t0.n_samples == t1.n_samples
    d = np.abs(t0.rgba - t1.rgba).max(axis=1)
    assert d.max() < 1 / 255 and (d == 0).mean() > 0.99
    # (the structured block's upper bound is origin + spacing * (n - 1) in f64, the mesh's the f32 point itself)
    hit = t0.min_dist >= 0
    assert np.array_equal(hit, t1.min_dist >= 0)
    assert np.allclose(t0.min_dist[hit], t1.min_dist[hit], rtol=1e-6) and np.allclose(t0.max_dist[hit], t1.max_dist[hit], rtol=1e-6)


def test_tetrahedra_reproduce_a_linear_field_exactly_like_hexahedra():
    """a field that is linear in x, y, z is interpolated exactly by both cell types: the two meshes of the same grid
    (hexahedra; six tetrahedra per cell) give the same image up to rounding, cell-centred values likewise"""
    dims, origin, spacing = (7, 6, 5), (-3.0, -2.0, -1.0), (1.0, 0.8, 0.5)
    pts, conn = datasets.structured_to_hexes(dims, origin, spacing)
    tets = datasets.hexes_to_tets(conn)
    field = (0.3 * pts[:, 0] - 0.2 * pts[:, 1] + 0.5 * pts[:, 2]).astype(np.float32)
    b = [pts[:, 0].min(), pts[:, 0].max(), pts[:, 1].min(), pts[:, 1].max(), pts[:, 2].min(), pts[:, 2].max()]
    cam = O.camera_reset_to_bounds(b)
    O.camera_azimuth(cam, -40.0)
    O.camera_elevation(cam, 25.0)
    lut = O.parse_color_table(scenes.RAMP_TF).correct_opacity(100).lut()
    sd = O.sample_distance(b, 100)
    W, H = 200, 160
    out = []
    for c in (conn, tets):
        um = O.OracleUMesh(pts, c, field)
        r, t = O.trace_umesh(um, cam, W, H, lut, sd, float(field.min()), float(field.max()))
        O.rays_free(r)
        out.append(t)
    assert out[0].n_samples > 100000
    assert abs(out[0].n_samples - out[1].n_samples) <= out[0].n_samples * 1e-3
    d = np.abs(out[0].rgba - out[1].rgba).max(axis=1)
    assert (d <= 1 / 255).mean() >= 0.999
    # partial extraction: alpha >= 0.001, depth = exit distance
    um = O.OracleUMesh(pts, conn, field)
    p = O.render_umesh_partials(um, cam, W, H, lut, sd, float(field.min()), float(field.max()))
    assert p.size > 1000 and (p["alpha"] >= np.float32(0.001)).all()
